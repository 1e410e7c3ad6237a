"""Cost of cudaHostRegister / cudaHostUnregister on this box for frame-sized buffers (what pinning a recycled frame buffer costs once)."""
import time, ctypes, json, torch, numpy as np
rt = torch.cuda.cudart()
torch.zeros(1, device="cuda")
out = {}
for mb in (0.5, 3, 12, 50):
    n = int(mb * (1 << 20))
    bufs = [np.ones(n, np.uint8) for _ in range(8)]
    t0 = time.perf_counter()
    for b in bufs:
        rt.cudaHostRegister(b.ctypes.data, n, 1)
    t1 = time.perf_counter()
    for b in bufs:
        rt.cudaHostUnregister(b.ctypes.data)
    t2 = time.perf_counter()
    out[f"{mb}MB"] = {"register_ms": (t1 - t0) / 8 * 1e3, "unregister_ms": (t2 - t1) / 8 * 1e3}
print(json.dumps(out))
