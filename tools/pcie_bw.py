"""Pinned-copy bandwidth of this box (the end-to-end roofline denominator): H2D alone, D2H alone,
both at once, contiguous vs the 2-D strided copies the host path issues (kept field = every other row)."""
import sys, time, json
import torch

def bw(fn, nbytes, iters=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return nbytes * iters / (time.perf_counter() - t0) / 1e9

N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
res["h2d_GBps"] = bw(h2d, N)
res["d2h_GBps"] = bw(d2h, N)
res["both_total_GBps"] = bw(both, 2 * N)
# strided rows: 1920-byte rows, every other row of a 2 x pitch source (what the kept-field upload looks like)
rows, w = 270 * 1024, 1920
hs = h_in[: rows * 2 * w].view(rows, 2 * w)[:, :w]
ds = d_a[: rows * w].view(rows, w)
def h2d_2d():
    with torch.cuda.stream(s1): ds.copy_(hs, non_blocking=True)
res["h2d_strided_1920B_rows_GBps"] = bw(h2d_2d, rows * w)
print(json.dumps(res))
