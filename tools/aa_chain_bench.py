#!/usr/bin/env python
"""BASELINE.json configs[2]: 1920x1080 YUV444P16, SangNom2(dh=true, aa=48) in a turn/dh/turn chain -> 3840x2160.
Times (a) the fused device chain through sangnom_cuda_chain_process with pinned host buffers (one upload, one
download), (b) the same four stages as separate host calls - pass 1 and pass 2 through sangnom_cuda_process_planes,
the turns NOT counted (a host filter would do them) - and (c) the turn kernel alone against the HBM copy peak.
Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200")]
import numpy as np
import torch
from pysangnom import cuda
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
fmt, w, h = FORMATS["YUV444P16"], 1920, 1080
sb = fmt.sample_bytes
base = [make_frame(1, w, h, fmt, "noise", i) for i in range(2)]
thr = [cuda.threshold(a, fmt.bits, sb) for a in (48, 0, 0)]
src_arena = cuda.PinnedArena(frames_n * 3 * w * h * sb + 4096)
dst_arena = cuda.PinnedArena(frames_n * 3 * 4 * w * h * sb + 4096)
mid_arena = cuda.PinnedArena(frames_n * 3 * 2 * w * h * sb + 4096)
jobs, j1, j2 = [], [], []
for k in range(frames_n):
    for p in range(3):
        s = src_arena.take((h, w), np.uint16); s[...] = base[k % 2][p]
        d = dst_arena.take((2 * h, 2 * w), np.uint16)
        m = mid_arena.take((2 * h, w), np.uint16)
        jobs.append(cuda.SnChainJob(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], w, h, 0, 0, thr[p], p, k))
        j1.append(cuda.make_job(s.ctypes.data, s.strides[0], m.ctypes.data, m.strides[0], w, 2 * h, 0, cuda.MODE_DH, thr[p], p, k))
        # pass 2 of the unfused path: a 2160 x 1920 plane (the turned pass-1 output; same bytes reinterpreted) -> 2160 x 3840
        mt = m.reshape(w, 2 * h)
        dt = d.reshape(2 * w, 2 * h)
        j2.append(cuda.make_job(mt.ctypes.data, mt.strides[0], dt.ctypes.data, dt.strides[0], 2 * h, 2 * w, 0, cuda.MODE_DH, thr[p], p, k))
arr = (cuda.SnChainJob * len(jobs))(*jobs)
out = {"workload": "1920x1080 YUV444P16 SangNom2(dh=true, aa=48) -> transpose -> SangNom2(dh=true) -> transpose [BASELINE.json configs[2]]",
       "frames_per_call": frames_n}
with cuda.Chain(sb, w, h, turn=cuda.TURN_TRANSPOSE) as ch:
    ch.process_jobs(arr)
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps):
        ch.process_jobs(arr)
    dt_ = time.perf_counter() - t0
    st = ch.stats()
    out["fused_chain_e2e_fps"] = frames_n * reps / dt_
    out["fused_h2d_bytes_per_frame"] = st["h2d_bytes"] // st["frames"]
    out["fused_d2h_bytes_per_frame"] = st["d2h_bytes"] // st["frames"]
with cuda.Context(sb, w, 2 * h) as c1, cuda.Context(sb, 2 * h, 2 * w) as c2:
    a1 = (cuda.SnPlaneJob * len(j1))(*j1); a2 = (cuda.SnPlaneJob * len(j2))(*j2)
    c1.process_jobs(a1); c2.process_jobs(a2)
    t0 = time.perf_counter()
    for _ in range(reps):
        c1.process_jobs(a1); c2.process_jobs(a2)
    dt_ = time.perf_counter() - t0
    out["separate_calls_e2e_fps_turns_not_counted"] = frames_n * reps / dt_
    out["separate_h2d_bytes_per_frame"] = (c1.stats()["h2d_bytes"] + c2.stats()["h2d_bytes"]) // c1.stats()["frames"]
    out["separate_d2h_bytes_per_frame"] = (c1.stats()["d2h_bytes"] + c2.stats()["d2h_bytes"]) // c1.stats()["frames"]
# turn kernel alone
lib = cuda.load()
n = 24
a = torch.randint(0, 255, (n, 2 * h, w * sb), dtype=torch.uint8, device="cuda")
b = torch.empty((n, w, 2 * h * sb), dtype=torch.uint8, device="cuda")
planes = (cuda.SnTurnPlane * n)(*[cuda.SnTurnPlane(a[i].data_ptr(), w * sb, b[i].data_ptr(), 2 * h * sb, w, 2 * h) for i in range(n)])
stream = torch.cuda.Stream()
for kind, name in ((0, "transpose"), (1, "turn_right")):
    lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        lib.sangnom_cuda_turn_planes_device(sb, kind, planes, n, C.c_void_p(stream.cuda_stream))
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"{name}_GBps"] = 2 * a.numel() / (ms / 1e3) / 1e9
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    out["hbm_copy_peak_GBps"] = json.load(open(peaks))["hbm_gbs"]
print(json.dumps(out))
