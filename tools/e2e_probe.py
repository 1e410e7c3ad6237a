"""Probe the host path: run only the end-to-end leg for a few seconds, sample SM clocks meanwhile."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
import numpy as np, torch
import bench
from pysangnom import cuda
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p8"
Fe = int(sys.argv[2]) if len(sys.argv) > 2 else 592
inflight = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fmtname, w, h, kw, _, _ = bench.WORKLOADS[wl]
fmt = FORMATS[fmtname]; sb = fmt.sample_bytes
base = [make_frame(1, w, h, fmt, "noise", i) for i in range(4)]
thr = [cuda.threshold(a, fmt.bits, sb) for a in (kw.get("aa", 48), kw.get("aac", 0), kw.get("aac", 0))]
proc = [kw.get("luma", True)] + [kw.get("chroma", True)] * 2
keep, jobs = [], []
for n in range(Fe):
    for p in range(min(fmt.components, 3)):
        a = base[n % 4][p]
        s = cuda.pinned_empty(a.shape, a.dtype); s[...] = a
        d = cuda.pinned_empty(a.shape, a.dtype)
        keep += [s, d]
        jobs.append(cuda.make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], a.shape[1], a.shape[0],
                                  cuda.resolve_offset(kw.get("order", 1), n % 2 == 0), cuda.MODE_FIELD if proc[p] else cuda.MODE_COPY, thr[p], p, n))
arr = (cuda.SnPlaneJob * len(jobs))(*jobs)
ctx = cuda.Context(sb, w, h, max_frames_in_flight=inflight)
lib = cuda.load()
for _ in range(2): lib.sangnom_cuda_process_planes(ctx._h, arr, len(arr))
smp = bench.ClockSampler(0); smp.start()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 4.0:
    lib.sangnom_cuda_process_planes(ctx._h, arr, len(arr)); n += 1
dt = time.perf_counter() - t0
clk = smp.stop()
print(json.dumps({"workload": wl, "frames_per_call": Fe, "in_flight": inflight, "fps": Fe * n / dt, "clocks": clk,
                  "sm_samples": sorted(set(float(r[1]) for r in smp.rows if len(r) > 1))[:20]}))
