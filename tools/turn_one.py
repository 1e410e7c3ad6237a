"""One turn-kernel launch on a few planes (for compute-sanitizer / debugging): python tools/turn_one.py <sample bytes> <w> <h> <kind> <planes>"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
import numpy as np, torch
from pysangnom import cuda
lib = cuda.load()
sb, w, h, kind, n = (int(x) for x in (sys.argv[1:] + ["2", "1920", "2160", "0", "2"])[:5])
rng = np.random.default_rng(1)
a = rng.integers(0, 255, size=(n, h, w * sb), dtype=np.uint8)
da = torch.from_numpy(a).cuda()
db = torch.zeros((n, w, h * sb), dtype=torch.uint8, device="cuda")
planes = (cuda.SnTurnPlane * n)(*[cuda.SnTurnPlane(da[i].data_ptr(), w * sb, db[i].data_ptr(), h * sb, w, h) for i in range(n)])
torch.cuda.synchronize()
rc = lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(0))
torch.cuda.synchronize()
dt = {1: np.uint8, 2: np.uint16, 4: np.float32}[sb]
ok = True
for i in range(n):
    src = a[i].view(dt)
    got = db[i].cpu().numpy().view(dt)
    exp = [src.T, np.rot90(src, -1), np.rot90(src, 1)][kind]
    ok &= bool(np.array_equal(got.view(np.uint8), np.ascontiguousarray(exp).view(np.uint8)))
print("rc", rc, "equal", ok)
