"""Turn kernel alone: GB/s (read + write) against the measured HBM copy peak, CUDA events on the launching stream."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
import torch
from pysangnom import cuda
lib = cuda.load()
out = {}
for sb, w, h, n in ((2, 1920, 2160, 48), (2, 2160, 3840, 24), (1, 1920, 1080, 96), (4, 3840, 2160, 12)):
    a = torch.randint(0, 255, (n, h, w * sb), dtype=torch.uint8, device="cuda")
    b = torch.empty((n, w, h * sb), dtype=torch.uint8, device="cuda")
    planes = (cuda.SnTurnPlane * n)(*[cuda.SnTurnPlane(a[i].data_ptr(), w * sb, b[i].data_ptr(), h * sb, w, h) for i in range(n)])
    stream = torch.cuda.Stream()
    for kind, name in ((0, "transpose"), (1, "turn_right"), (2, "turn_left")):
        for _ in range(3):
            lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
        e1.record(stream); torch.cuda.synchronize()
        out[f"{name}_{sb}B_{w}x{h}x{n}"] = round(2 * a.numel() / (e0.elapsed_time(e1) / 10 / 1e3) / 1e9, 1)
print(json.dumps(out))
