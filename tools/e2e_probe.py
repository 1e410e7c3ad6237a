"""Probe the host entry alone (no device-resident leg, no CPU arm): frames/s for a few seconds per variant.
   python tools/e2e_probe.py [workload] [frames per batch] [variant ...]
variants: pinned (src -> dst, both in pinned arenas), inplace (dst already holds the kept field: src == dst),
pageable (numpy memory), field (separated-field input, SN_MODE_DH). Environment knobs of the library apply
(SANGNOM_B200_COPY_THREADS, SANGNOM_TRACE)."""
import os, sys, time, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import bench
from pysangnom import cuda
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p8"
Fe = int(sys.argv[2]) if len(sys.argv) > 2 else 592
variants = sys.argv[3:] or ["pinned", "inplace", "pageable", "field"]
devices = [int(d) for d in os.environ.get("PROBE_DEVICES", "0").split(",")]
fmtname, w, h, kw, _, _ = bench.WORKLOADS[wl]
fmt = FORMATS[fmtname]; sb = fmt.sample_bytes
nplanes = min(fmt.components, 3)
base = [make_frame(1, w, h, fmt, "noise", i) for i in range(4)]
thr = [cuda.threshold(a, fmt.bits, sb) for a in (kw.get("aa", 48), kw.get("aac", 0), kw.get("aac", 0))]
proc = [kw.get("luma", True)] + [kw.get("chroma", True)] * 2
frame_bytes = sum(base[0][p].nbytes for p in range(nplanes))
lib = cuda.load()


def build(variant):
    sets = []
    for _ in range(2):
        if variant == "pageable":
            take_s = take_d = lambda shape, dt: np.empty(shape, dt)
            keep = None
        else:
            sa = cuda.PinnedArena(Fe * (frame_bytes + 64) + 4096)
            da = cuda.PinnedArena(Fe * (frame_bytes + 64) + 4096)
            take_s, take_d, keep = sa.take, da.take, (sa, da)
        jobs, bufs = [], []
        for n in range(Fe):
            off = cuda.resolve_offset(kw.get("order", 1), n % 2 == 0)
            for p in range(nplanes):
                a = base[n % 4][p]
                mode = cuda.MODE_FIELD if proc[p] else cuda.MODE_COPY
                if variant == "field" and proc[p]:
                    s = take_s((a.shape[0] // 2, a.shape[1]), a.dtype); s[...] = a[off::2]
                    mode = cuda.MODE_DH
                else:
                    s = take_s(a.shape, a.dtype); s[...] = a
                d = s if variant == "inplace" else take_d(a.shape, a.dtype)
                bufs += [s, d]
                jobs.append(cuda.make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], a.shape[1], a.shape[0], off, mode, thr[p], p, n))
        sets.append(((cuda.SnPlaneJob * len(jobs))(*jobs), bufs, keep))
    return sets


for variant in variants:
    sets = build(variant)
    ctx = cuda.Context(sb, w, h, device=devices if len(devices) > 1 else devices[0])
    for i in range(3):
        assert lib.sangnom_cuda_process_planes(ctx._h, sets[i % 2][0], len(sets[i % 2][0])) == 0, lib.sangnom_cuda_last_error(ctx._h)
    ctx.reset_stats()
    tick = C.c_uint64(); pending = []; n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 2.5:
        arr = sets[n % 2][0]
        assert lib.sangnom_cuda_submit(ctx._h, arr, len(arr), C.byref(tick)) == 0
        pending.append(int(tick.value)); n += 1
        if len(pending) == 2:
            assert lib.sangnom_cuda_wait(ctx._h, pending.pop(0)) == 0, lib.sangnom_cuda_last_error(ctx._h)
    while pending:
        assert lib.sangnom_cuda_wait(ctx._h, pending.pop(0)) == 0
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print(json.dumps({"workload": wl, "variant": variant, "devices": devices, "frames_per_batch": Fe, "fps": round(Fe * n / dt, 1),
                      "h2d_MB_per_frame": round(st["h2d_bytes"] / st["frames"] / 1e6, 3), "d2h_MB_per_frame": round(st["d2h_bytes"] / st["frames"] / 1e6, 3),
                      "host_copy_MB_per_frame": round(st["host_copy_bytes"] / st["frames"] / 1e6, 3),
                      "pcie_GBps_each_way": round(st["h2d_bytes"] / dt / 1e9, 1),
                      "copy_threads": os.environ.get("SANGNOM_B200_COPY_THREADS", "default")}), flush=True)
    ctx.close()
    del sets
