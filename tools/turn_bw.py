"""Turn kernel alone: GB/s (read + write) against the measured HBM copy peak, CUDA events on the launching stream."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
import torch
from pysangnom import cuda
lib = cuda.load()
out = {}
# rows padded to 256 bytes, like the planes of the device chain (sangnom_chain.cu): planes whose base or pitch is not a
# multiple of 16 bytes (a 1080-byte row) cannot be addressed by a tensor map and take the plain kernel
for sb, w, h, n in ((2, 1920, 2160, 48), (2, 2160, 3840, 24), (1, 1920, 1080, 96), (1, 1920, 2160, 96), (4, 3840, 2160, 12)):
    sp, dp = (w * sb + 255) // 256 * 256, (h * sb + 255) // 256 * 256
    a = torch.randint(0, 255, (n, h, sp), dtype=torch.uint8, device="cuda")
    b = torch.empty((n, w, dp), dtype=torch.uint8, device="cuda")
    planes = (cuda.SnTurnPlane * n)(*[cuda.SnTurnPlane(a[i].data_ptr(), sp, b[i].data_ptr(), dp, w, h, cuda.TURN_DST_PADDING_WRITABLE) for i in range(n)])
    stream = torch.cuda.Stream()
    for kind, name in ((0, "transpose"), (1, "turn_right"), (2, "turn_left")):
        for _ in range(3):
            lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
        e1.record(stream); torch.cuda.synchronize()
        out[f"{name}_{sb}B_{w}x{h}x{n}"] = round(2 * n * w * h * sb / (e0.elapsed_time(e1) / 10 / 1e3) / 1e9, 1)
print(json.dumps(out))
