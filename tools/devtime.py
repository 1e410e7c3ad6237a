"""Device-resident frames/s of the row-sweep kernels for one or more bench workloads (no host legs, no CPU arm):
   python tools/devtime.py [workload[:frames]] ...     e.g.  python tools/devtime.py 1080p8 2160pf32:148 2160p10
Same job construction as bench.py's `value` leg (seeded noise frames, in place, pitch 256), CUDA events on the
launching stream, best of 3 runs of `iters` steps."""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import bench
from pysangnom import cuda
if os.environ.get("SANGNOM_DEV_LIB"):       # kernel experiments: a library built aside (never the product path)
    cuda.CUDA_LIB = os.path.abspath(os.environ["SANGNOM_DEV_LIB"])
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS


def run(wl, F, iters=5):
    fmtname, w, h, kw, dtype_tag, _ = bench.WORKLOADS[wl]
    fmt = FORMATS[fmtname]
    sb = fmt.sample_bytes
    nplanes = min(fmt.components, 3)
    proc = [kw.get("luma", True)] + [kw.get("chroma", True)] * 2
    thr = [cuda.threshold(a, fmt.bits, sb) for a in (kw.get("aa", 48), kw.get("aac", 0), kw.get("aac", 0))]
    base = [make_frame(1, w, h, fmt, "noise", i) for i in range(4)]
    dev = [[None] * nplanes for _ in range(4)]
    keep, jobs = [], []
    for n in range(F):
        for p in range(nplanes):
            if not proc[p]:
                continue
            a = base[n % 4][p]
            ph, pw = a.shape
            pitch = (pw * sb + 255) // 256 * 256
            if dev[n % 4][p] is None:
                t0 = torch.empty((ph, pitch), dtype=torch.uint8, device="cuda")
                t0[:, :pw * sb] = torch.from_numpy(a.view(np.uint8).reshape(ph, -1)).cuda()
                dev[n % 4][p] = t0
            t = dev[n % 4][p].clone()
            keep.append(t)
            jobs.append(cuda.make_job(0, 0, t.data_ptr(), pitch, pw, ph, cuda.resolve_offset(kw.get("order", 1), n % 2 == 0), cuda.MODE_INPLACE, thr[p], p, n))
    arr = (cuda.SnPlaneJob * len(jobs))(*jobs)
    ctx = cuda.Context(sb, w, h)
    ts = torch.cuda.Stream()
    for _ in range(3):
        ctx.process_jobs_device(arr, ts.cuda_stream)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(iters):
            ctx.process_jobs_device(arr, ts.cuda_stream)
        e1.record(ts)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    alg = bench.algorithmic_bytes_per_frame(wl)
    ctx.close()
    out = {"workload": wl, "frames": F, "ms_per_step": round(best, 4), "fps": round(F / best * 1000.0, 1), "alg_GBps": round(alg * F / best / 1e6, 1),
           "env": {k: v for k, v in os.environ.items() if k.startswith("SANGNOM_")}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:] or ["1080p8", "2160pf32", "2160p10"]:
        wl, _, f = spec.partition(":")
        run(wl, int(f) if f else bench.DEFAULT_FRAMES[wl])
