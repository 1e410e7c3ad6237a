#!/usr/bin/env python
"""Headline benchmark: SangNom2 frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1080p8|2160pf32|...]

A "step" is one pass of the hot path over one batch of synthetic frames of the workload.
  value      device-resident frames/s: planes already in HBM, kernels launched through the C ABI's device entry
             (sangnom_cuda_process_planes_device), timed with CUDA events on the launching stream; the K-step region
             is run several times back to back (>= 1 s in all) and the MEDIAN region counts; max over ranks.
  e2e        the same metric through the host entry (sangnom_cuda_submit / _wait) with pinned HOST buffers: full
             frames in, full frames out, uploads, kernels and downloads inside the timed region. At N > 1 the N GPUs
             sit behind ONE context (one pipeline per device, chunks of frames dealt round-robin) driven by rank 0.
             e2e.pcie_peak_gbs / e2e.frac: pinned-copy bandwidth per direction measured in the same run with both
             directions and all N GPUs busy, and the fraction of it the end-to-end leg reaches.
  secondary  the metric's second workload (BASELINE.json: "1080p8/2160p fp32 frames/s"): value, e2e and roofline of
             3840x2160 YUV420PS at the same N.
  roofline / cpu_baseline / clocks as the task contract asks; see DESIGN.md "Measurement".
N>1: launched by torchrun, one rank per GPU; every rank runs the same per-GPU batch on its own frame range
(weak scaling, no data-path collective); NCCL is used only for the barrier and the max-over-ranks reduction.
--impl reference times the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile)
through its own plugin API and stock code path (opt=-1 -> SSE2) on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "avisynth-sangnom2_b200")
for _p in (ROOT, PKG, os.path.join(ROOT, "tests")):     # tests/: the fake AviSynth host both plugin legs are driven by
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (format, width, height, script args, dtype tag, BASELINE.json config index)
    "1080p8": ("YUV420P8", 1920, 1080, dict(order=0, aa=48, aac=48), "u8", 1),
    "2160pf32": ("YUV420PS", 3840, 2160, dict(order=2, aa=48, aac=24), "f32", 3),
    "2160p10": ("YUV420P10", 3840, 2160, dict(order=1, aa=48, aac=48), "u16", 4),
    "480p8": ("YV12", 720, 480, dict(order=1, aa=48, chroma=False), "u8", 0),
}
DEFAULT_FRAMES = {"1080p8": 592, "2160pf32": 148, "2160p10": 148, "480p8": 1184}      # per step per GPU, device-resident leg
E2E_FRAMES = {"1080p8": 296, "2160pf32": 24, "2160p10": 48, "480p8": 1184}            # per step per GPU, the same at every N
SECONDARY = {"1080p8": "2160pf32"}                                                     # the metric's second workload


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_desc(name):
    fmtname, w, h, kw, dt, idx = WORKLOADS[name]
    args = ", ".join(f"{k}={v}" for k, v in kw.items())
    return f"{w}x{h} {fmtname} SangNom2({args}) [BASELINE.json configs[{idx}]]"


def config_of(name):
    """`config` of the JSON line - the same dict from both arms."""
    return {"workload": workload_desc(name), "frames_per_step_per_gpu": DEFAULT_FRAMES[name],
            "l2": f"inputs larger than L2: {algorithmic_bytes_per_frame(name) * DEFAULT_FRAMES[name] / 1e6:.0f} MB of planes per step per GPU",
            "sharding": "contiguous frame ranges per rank, no data-path collective"}


def algorithmic_bytes_per_frame(name):
    """W*H*s per processed plane: kept field read once + interpolated rows written once."""
    from pysangnom.formats import FORMATS
    fmtname, w, h, kw, _, _ = WORKLOADS[name]
    fmt = FORMATS[fmtname]
    total = 0
    for p in range(min(fmt.components, 3)):
        if p == 0 and not kw.get("luma", True):
            continue
        if p > 0 and not kw.get("chroma", True):
            continue
        ph, pw = fmt.plane_shape(w, h, p)
        total += pw * ph * fmt.sample_bytes
    return total


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def reference_fps(workload, threads, seconds_budget, opt=-1, warm=1):
    """Frames/s of the compiled reference plugin on host cores: one fresh filter instance per thread over
    disjoint frame ranges (what MT_MULTI_INSTANCE + Prefetch(threads) does in AviSynth+)."""
    from oracle import oracle as O
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    from fakehost import FakeHost

    plugin = O.reference_plugin_path()
    if plugin is None:
        return None
    fmtname, w, h, kw, _, _ = WORKLOADS[workload]
    fmt = FORMATS[fmtname]
    nsrc = 4
    frames = [make_frame(1, w, h, fmt, "noise", i) for i in range(nsrc)]
    hosts, filters = [], []
    for t in range(threads):
        host = FakeHost(poison_new_frames=False)
        host.load_plugin(plugin)
        src = host.source(w, h, fmt, nsrc, parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        filters.append(host.invoke("SangNom2", src, opt=opt, **kw))
        hosts.append(host)

    import fakehost as fh
    L = fh._load()
    err = C.create_string_buffer(256)

    def pull(t, n):
        f = L.fh_get_frame(hosts[t].env, filters[t].handle, n % nsrc, err, 256)
        L.fh_frame_release(f)

    for t in range(threads):
        for n in range(warm):
            pull(t, n)
    # calibrate on one thread, then run a fixed count per thread
    t0 = time.perf_counter()
    pull(0, 0)
    one = time.perf_counter() - t0
    per_thread = max(2, int(seconds_budget / max(one, 1e-6)))
    done = [0] * threads

    def worker(t):
        for n in range(per_thread):
            pull(t, n)
        done[t] = per_thread

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    for host in hosts:
        host.close()
    return {"fps": sum(done) / dt, "frames": sum(done), "seconds": dt, "threads": threads}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    wl = args.workload
    total = 0
    res = None
    t_all = time.perf_counter()
    per_step = max(1.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        reference_fps(wl, cores, per_step / 4)
    vals = []
    for _ in range(args.steps):
        res = reference_fps(wl, cores, per_step)
        if res is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference in the build container)"})
            return 0
        vals.append(res["fps"])
        total += res["frames"]
    fps = statistics.mean(vals)
    out = {
        "impl": "reference", "metric": "frames_per_second", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * res["seconds"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": WORKLOADS[wl][4], "data": "synthetic",
        "config": config_of(wl),
        "reference_path": "unmodified reference plugin, stock opt=-1 (SSE2), one instance per host thread",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": f"{res['frames']} frames per step over {cores} threads (~{per_step:.0f} s of CPU work per thread-step)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    emit(out)
    return 0



_ORIGINAL_AFFINITY = None


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank on the CPUs next to its GPU so that the pinned staging memory it allocates (first touch) and the
    copy threads sit on the GPU's NUMA node; matters for the end-to-end leg with several ranks per box."""
    if os.environ.get("SANGNOM_BENCH_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        global _ORIGINAL_AFFINITY
        _ORIGINAL_AFFINITY = os.sched_getaffinity(0)
        cpus &= _ORIGINAL_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception as e:       # no NVML / not permitted: stay where the launcher put us
        log(f"bench: no NUMA binding ({e})")
        return None


# ------------------------------------------------------------------------------------------------
class Run:
    """What every leg needs: ranks, the library, the workload's frames."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from pysangnom import cuda
        self.torch, self.dist, self.cuda, self.args = torch, dist, cuda, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torchrun (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libsangnom_cuda has no CPU path")
        torch.cuda.set_device(self.local)
        bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib = cuda.load()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


class Workload:
    def __init__(self, run, name):
        from pysangnom.clips import make_frame
        from pysangnom.formats import FORMATS
        from pysangnom.shard import frame_range
        cuda = run.cuda
        self.name = name
        fmtname, self.w, self.h, self.kw, self.dtype_tag, _ = WORKLOADS[name]
        self.fmt = FORMATS[fmtname]
        self.sb = self.fmt.sample_bytes
        self.nplanes = min(self.fmt.components, 3)
        self.proc = [self.kw.get("luma", True)] + [self.kw.get("chroma", True)] * 2
        self.thr = [cuda.threshold(a, self.fmt.bits, self.sb) for a in (self.kw.get("aa", 48), self.kw.get("aac", 0), self.kw.get("aac", 0))]
        self.base = [make_frame(1, self.w, self.h, self.fmt, "noise", i) for i in range(4)]      # a few distinct seeded frames, tiled over the batch
        self.frame_bytes = sum(self.base[0][p].nbytes for p in range(self.nplanes))
        self.alg = algorithmic_bytes_per_frame(name)
        self.F = (run.args.frames if name == run.args.workload and run.args.frames else DEFAULT_FRAMES[name])
        self.Fe = (run.args.e2e_frames if name == run.args.workload and run.args.e2e_frames else E2E_FRAMES[name])
        self.first, _ = frame_range(self.F * run.world, run.rank, run.world)      # this rank's contiguous frame range of the global clip

    def offset_of(self, n, cuda):
        return cuda.resolve_offset(self.kw.get("order", 1), n % 2 == 0)


def device_leg(run, wl, steps, warmup):
    """`value`: planes resident in HBM, in place, one launch per pass index per step."""
    torch, cuda = run.torch, run.cuda
    dev_planes, jobs = [], []
    for k in range(wl.F):
        n = wl.first + k
        for p in range(wl.nplanes):
            if not wl.proc[p]:
                continue
            a = wl.base[n % 4][p]
            ph, pw = a.shape
            pitch = (pw * wl.sb + 255) // 256 * 256
            t = torch.empty((ph, pitch), dtype=torch.uint8, device="cuda")
            t[:, :pw * wl.sb] = torch.from_numpy(a.view(np.uint8).reshape(ph, -1)).cuda()
            dev_planes.append(t)
            jobs.append(cuda.make_job(0, 0, t.data_ptr(), pitch, pw, ph, wl.offset_of(n, cuda), cuda.MODE_INPLACE, wl.thr[p], p, n))
    job_arr = (cuda.SnPlaneJob * len(jobs))(*jobs)
    ctx = cuda.Context(wl.sb, wl.w, wl.h, device=run.local)
    tstream = torch.cuda.Stream()                        # the launching stream; events are recorded on it
    stream = tstream.cuda_stream

    def step():
        ctx.process_jobs_device(job_arr, stream)

    for _ in range(max(warmup, 3)):
        step()
    run.barrier()
    # how many K-step regions make >= ~1 s of timed work (the clock sampler ticks every 200 ms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    step()
    e1.record(tstream)
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(101, max(3, 1200.0 / (one * steps)))) | 1
    reps = int(run.max_over_ranks([reps])[0])
    sampler = ClockSampler(run.local)
    if run.rank == 0:
        sampler.start()
    ctx.reset_stats()
    run.barrier()
    torch.cuda.cudart().cudaProfilerStart()           # `ncu --profile-from-start off` lists the timed regions only
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in events:
        a.record(tstream)
        for _ in range(steps):
            step()
        b.record(tstream)
    run.barrier()
    torch.cuda.cudart().cudaProfilerStop()
    regions = sorted(a.elapsed_time(b) for a, b in events)
    ms = regions[len(regions) // 2]                     # median K-step region of this rank
    launches = ctx.stats()["kernel_launches"] // reps
    ms_max = run.max_over_ranks([ms])[0]
    clocks = sampler.stop() if run.rank == 0 else None
    out = {"value": wl.F * run.world * steps / (ms_max / 1000.0), "ms_per_step": ms_max / steps, "ms_rank": ms,
           "launches": int(launches), "regions": reps, "region_ms_min_max": [regions[0], regions[-1]], "clocks": clocks,
           "dev_plane0": dev_planes[0]}
    ctx.close()
    return out


def roofline_of(wl, dev, steps):
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = wl.alg * wl.F * steps / (dev["ms_rank"] / 1000.0) / 1e9          # this rank's kernels over its own event time
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(wl.name)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": wl.alg * wl.F / max(1, dev["launches"] // steps),
            "note": "kernel is instruction-issue / latency bound, not HBM bound (DESIGN.md, Rooflines)"}


def host_sets(run, wl, frames, variant="frame"):
    """Two sets of pinned job lists over `frames` frames: one shared source arena, two destination arenas - the staging a
    batching host layer does; steps alternate between the sets so that step k+1 can be submitted while step k downloads.
    variant "field": separated-field sources (SN_MODE_DH); "inplace": the destination frames already hold the kept field
    (src == dst), as when a decoder writes fields straight into the output frames."""
    cuda = run.cuda
    field = variant == "field"
    if variant == "inplace":
        sets = []
        for _ in range(2):
            arena = cuda.PinnedArena(frames * (wl.frame_bytes + 64) + 4096)
            bufs, hjobs = [], []
            for k in range(frames):
                n = wl.first + k
                for p in range(wl.nplanes):
                    a = wl.base[n % 4][p]
                    d = arena.take(a.shape, a.dtype)
                    d[...] = a
                    bufs.append(d)
                    mode = cuda.MODE_FIELD if wl.proc[p] else cuda.MODE_COPY
                    hjobs.append(cuda.make_job(d.ctypes.data, d.strides[0], d.ctypes.data, d.strides[0], a.shape[1], a.shape[0],
                                               wl.offset_of(n, cuda), mode, wl.thr[p], p, n))
            sets.append(((cuda.SnPlaneJob * len(hjobs))(*hjobs), bufs, arena))
        return sets, None
    src_arena = cuda.PinnedArena(frames * ((wl.frame_bytes // 2 if field else wl.frame_bytes) + 64) + 4096)
    srcs = []
    for k in range(frames):
        n = wl.first + k
        off = wl.offset_of(n, cuda)
        row = []
        for p in range(wl.nplanes):
            a = wl.base[n % 4][p]
            if field and wl.proc[p]:
                s = src_arena.take((a.shape[0] // 2, a.shape[1]), a.dtype)
                s[...] = a[off::2]
            else:
                s = src_arena.take(a.shape, a.dtype)
                s[...] = a
            row.append(s)
        srcs.append(row)
    sets = []
    for _ in range(2):
        dst_arena = cuda.PinnedArena(frames * (wl.frame_bytes + 64) + 4096)
        dst_host, hjobs = [], []
        for k in range(frames):
            n = wl.first + k
            for p in range(wl.nplanes):
                a = wl.base[n % 4][p]
                s = srcs[k][p]
                d = dst_arena.take(a.shape, a.dtype)
                dst_host.append(d)
                mode = (cuda.MODE_DH if field else cuda.MODE_FIELD) if wl.proc[p] else cuda.MODE_COPY
                hjobs.append(cuda.make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], a.shape[1], a.shape[0],
                                           wl.offset_of(n, cuda), mode, wl.thr[p], p, n))
        sets.append(((cuda.SnPlaneJob * len(hjobs))(*hjobs), dst_host, dst_arena))
    return sets, src_arena


def stream_steps(run, ectx, job_arrays, esteps):
    """Step i+1 is submitted before step i is waited for - how a frame server keeps the GPUs fed. Wall clock."""
    lib = run.lib

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.sangnom_cuda_last_error(ectx._h).decode())
    ectx.reset_stats()
    tick = C.c_uint64()
    pending = []
    t0 = time.perf_counter()
    for i in range(esteps):
        arr = job_arrays[i % 2]
        check(lib.sangnom_cuda_submit(ectx._h, arr, len(arr), C.byref(tick)))
        pending.append(int(tick.value))
        if len(pending) == 2:
            check(lib.sangnom_cuda_wait(ectx._h, pending.pop(0)))
    while pending:
        check(lib.sangnom_cuda_wait(ectx._h, pending.pop(0)))
    return time.perf_counter() - t0, ectx.stats()


def e2e_leg(run, wl, dev, steps, full=True):
    """The metric through the host entry. N = 1: this rank's GPU. N > 1: rank 0 drives all N GPUs through ONE context
    (device list 0..N-1); the other ranks hold still at the barrier. `full`: also the synchronous-call, the
    separated-field and (N > 1) the one-context-per-rank variants."""
    torch, cuda, lib = run.torch, run.cuda, run.lib
    world = run.world
    esteps = max(4, steps)
    frames = wl.Fe * world                                 # weak scaling: Fe frames per GPU per step
    res = {}
    if run.rank == 0:
        sets, src_arena = host_sets(run, wl, frames)
        copy_threads = int(os.environ.get("SANGNOM_B200_COPY_THREADS", "0"))
        ectx = cuda.Context(wl.sb, wl.w, wl.h, device=list(range(world)) if world > 1 else run.local,
                            max_frames_in_flight=run.args.in_flight, copy_threads=copy_threads)

        def sync_step(i):
            arr = sets[i % 2][0]
            if lib.sangnom_cuda_process_planes(ectx._h, arr, len(arr)) != 0:
                raise RuntimeError(lib.sangnom_cuda_last_error(ectx._h).decode())

        for i in range(3):                                  # every pipeline slot has seen (and sized itself for) the largest chunk
            sync_step(i)
        e_s, est = stream_steps(run, ectx, [sets[0][0], sets[1][0]], esteps)
        # spot-check an e2e output plane against the device-resident result of the same frame (same bytes expected)
        chk = sets[(esteps - 1) % 2][1][0]
        ref_dev = dev["dev_plane0"][:, :chk.shape[1] * wl.sb].cpu().numpy().view(chk.dtype)
        if wl.proc[0] and not np.array_equal(ref_dev, chk):
            raise RuntimeError("e2e output differs from the device-resident output")
        res = {"value": frames * esteps / e_s, "unit": "frames/s", "h2d_bytes_per_step": est["h2d_bytes"] // esteps,
               "d2h_bytes_per_step": est["d2h_bytes"] // esteps, "host_copy_bytes_per_step": est["host_copy_bytes"] // esteps,
               "steps": esteps, "frames_per_step": frames, "devices_behind_the_context": ectx.device_count(),
               "api": "sangnom_cuda_submit/_wait, two steps in flight, pinned host arenas; kept rows up, interpolated rows down, "
                      "kept rows + border row copied src -> dst by the library's host threads"}
        if full:
            if world == 1:
                t0 = time.perf_counter()
                for i in range(esteps):
                    sync_step(i)                            # every step pays the pipeline's ramp (first upload, last download)
                res["sync_call_value"] = frames * esteps / (time.perf_counter() - t0)
            if world == 1 and all(wl.proc[:wl.nplanes]) and not wl.kw.get("dh", False):
                # the same frames from a double-rate producer that hands over SEPARATED FIELDS: SN_MODE_DH
                fsets, farena = host_sets(run, wl, frames, "field")
                sync = fsets[1][0]
                lib.sangnom_cuda_process_planes(ectx._h, sync, len(sync))
                f_s, fst = stream_steps(run, ectx, [fsets[0][0], fsets[1][0]], esteps)
                chk = fsets[(esteps - 1) % 2][1][0]
                if wl.proc[0] and not np.array_equal(ref_dev, chk):
                    raise RuntimeError("e2e output (field input) differs from the device-resident output")
                res["field_input"] = {"value": frames * esteps / f_s, "h2d_bytes_per_step": fst["h2d_bytes"] // esteps,
                                      "d2h_bytes_per_step": fst["d2h_bytes"] // esteps,
                                      "note": "separated-field input (SN_MODE_DH): same output frames, contiguous upload"}
                del fsets, farena
            # the destination frames already hold the kept field (src == dst): no host-side copy of kept rows at all -
            # what is left is the PCIe / host-memory traffic of the kept rows up and the interpolated rows down
            isets, _ = host_sets(run, wl, frames, "inplace")
            sync = isets[1][0]
            lib.sangnom_cuda_process_planes(ectx._h, sync, len(sync))
            i_s, ist = stream_steps(run, ectx, [isets[0][0], isets[1][0]], esteps)
            chk = isets[(esteps - 1) % 2][1][0]
            if wl.proc[0] and not np.array_equal(ref_dev, chk):
                raise RuntimeError("e2e output (in place) differs from the device-resident output")
            res["inplace"] = {"value": frames * esteps / i_s, "h2d_bytes_per_step": ist["h2d_bytes"] // esteps, "d2h_bytes_per_step": ist["d2h_bytes"] // esteps,
                              "host_copy_bytes_per_step": ist["host_copy_bytes"] // esteps,
                              "note": "dst already holds the kept field (src == dst): no host copy of kept rows"}
            del isets
        ectx.close()
        del sets, src_arena
    run.barrier()
    if full and world > 1:
        # for comparison: N independent contexts, one per rank (what round 1 measured)
        sets, src_arena = host_sets(run, wl, wl.Fe)
        ectx = cuda.Context(wl.sb, wl.w, wl.h, device=run.local, max_frames_in_flight=run.args.in_flight,
                            copy_threads=max(1, 8 // world))
        for i in range(3):
            lib.sangnom_cuda_process_planes(ectx._h, sets[i % 2][0], len(sets[i % 2][0]))
        run.barrier()
        e_s, _ = stream_steps(run, ectx, [sets[0][0], sets[1][0]], esteps)
        e_max = run.max_over_ranks([e_s])[0]
        if run.rank == 0:
            res["one_context_per_rank_value"] = wl.Fe * world * esteps / e_max
        ectx.close()
        del sets, src_arena
        run.barrier()
    return res


def pcie_roof(run):
    """Pinned-copy bandwidth of this box with all N GPUs busy at once: contiguous 256 MB copies, H2D and D2H running
    concurrently on every rank (what the end-to-end leg does). GB/s per direction, summed over the ranks."""
    torch = run.torch
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    both()
    run.barrier()
    iters = 6
    t0 = time.perf_counter()
    for _ in range(iters):
        both()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt_max = run.max_over_ranks([dt])[0]
    run.barrier()
    return {"per_direction_gbs_all_gpus": run.world * n * iters / dt_max / 1e9, "per_direction_gbs_per_gpu": n * iters / dt_max / 1e9,
            "how": f"{run.world} rank(s), each 256 MB pinned H2D + 256 MB D2H concurrently x {iters}, wall clock, slowest rank"}


def run_ours(args):
    run = Run(args)
    wl = Workload(run, args.workload)
    dev = device_leg(run, wl, args.steps, args.warmup)
    e2e = e2e_leg(run, wl, dev, args.steps, full=True)
    out = None
    if run.rank == 0:
        out = {
            "metric": "frames_per_second", "value": dev["value"], "unit": "frames/s", "n_gpus": run.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype_tag, "data": "synthetic", "config": config_of(wl.name),
            "timing": {"regions": dev["regions"], "region_ms_min_max": dev["region_ms_min_max"],
                       "note": f"the {args.steps}-step region was timed {dev['regions']} times back to back with CUDA events; value = median region, max over ranks"},
            "roofline": roofline_of(wl, dev, args.steps),
            "e2e": e2e, "gpu_launches": dev["launches"], "clocks": dev["clocks"],
        }
    del dev
    sec_name = SECONDARY.get(args.workload) if not args.no_secondary else None
    if sec_name:
        wl2 = Workload(run, sec_name)
        dev2 = device_leg(run, wl2, max(2, args.steps // 2), args.warmup)
        e2e2 = e2e_leg(run, wl2, dev2, max(2, args.steps // 2), full=False)
        if run.rank == 0:
            out["secondary"] = {"config": config_of(sec_name), "dtype": wl2.dtype_tag, "value": dev2["value"], "unit": "frames/s",
                                "ms_per_step": dev2["ms_per_step"], "steps": max(2, args.steps // 2), "roofline": roofline_of(wl2, dev2, max(2, args.steps // 2)),
                                "e2e": e2e2, "gpu_launches": dev2["launches"], "clocks": dev2["clocks"]}
        del dev2
    roof = pcie_roof(run)
    if run.rank == 0:
        for block, w_ in ((out, wl),) + (((out["secondary"], wl2),) if sec_name else ()):
            e = block["e2e"]
            gbs = max(e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]) * e["steps"] / (e["frames_per_step"] * e["steps"] / e["value"]) / 1e9
            e["pcie_gbs_busier_direction"] = gbs
            # every byte of the end-to-end leg crosses host DRAM: DMA reads of the kept rows, DMA writes of the interpolated
            # rows, and the host threads' read + write of the kept rows (src -> dst). On the measured boxes THIS saturates
            # (about 130 GB/s, whatever the number of GPUs), not PCIe
            e["host_memory_traffic_gbs"] = (e["h2d_bytes_per_step"] + e["d2h_bytes_per_step"] + 2 * e["host_copy_bytes_per_step"]) * e["value"] / e["frames_per_step"] / 1e9
            if "inplace" in e:
                i_ = e["inplace"]
                i_["host_memory_traffic_gbs"] = (i_["h2d_bytes_per_step"] + i_["d2h_bytes_per_step"] + 2 * i_["host_copy_bytes_per_step"]) * i_["value"] / e["frames_per_step"] / 1e9
            e["pcie_peak_gbs"] = roof["per_direction_gbs_all_gpus"]
            e["frac"] = gbs / roof["per_direction_gbs_all_gpus"]
        out["e2e"]["pcie_roof"] = roof
        if run.world == 1 and args.plugin_seconds > 0:
            out["e2e"]["plugin_value"] = plugin_fps(args.workload, args.plugin_seconds)
            if sec_name:
                out["secondary"]["e2e"]["plugin_value"] = plugin_fps(sec_name, args.plugin_seconds)
        if run.world == 1 and not args.no_cpu_baseline:
            if _ORIGINAL_AFFINITY:
                os.sched_setaffinity(0, _ORIGINAL_AFFINITY)      # the CPU baseline gets every host core
            cores = len(os.sched_getaffinity(0))
            ref = reference_fps(args.workload, cores, args.cpu_seconds)
            if ref is not None:
                ref0 = reference_fps(args.workload, cores, args.cpu_seconds / 2, opt=0)
                out["cpu_baseline"] = {"value": ref["fps"], "unit": "frames/s", "cores": cores, "kind": "reference",
                                       "sample": f"{ref['frames']} frames of the same workload in {ref['seconds']:.1f} s, one reference "
                                                 f"instance per host thread, stock opt=-1 (SSE2)",
                                       "opt0_cpp_path_fps": ref0["fps"]}
                if sec_name:
                    r2 = reference_fps(sec_name, cores, args.cpu_seconds / 2)
                    out["secondary"]["cpu_baseline"] = {"value": r2["fps"], "unit": "frames/s", "cores": cores, "kind": "reference",
                                                        "sample": f"{r2['frames']} frames in {r2['seconds']:.1f} s, stock opt=-1 (SSE2)"}
            else:
                out["cpu_baseline"] = cpu_port_baseline(args.workload, args.cpu_seconds)
        emit(out)
    if run.world > 1:
        run.dist.destroy_process_group()
    return 0


def plugin_fps(wl, seconds):
    """The drop-in path itself: our AviSynth plugin pulled frame by frame through the fake host (the host's own
    recycled frames, one filter instance, sequential GetFrame) - the same harness the reference arm is timed with.
    Warm-up runs until the plugin has pinned the host's recycled frame buffers (a one-time cost per buffer)."""
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    from fakehost import FakeHost
    import fakehost as fh
    fmtname, w, h, kw, _, _ = WORKLOADS[wl]
    fmt = FORMATS[fmtname]
    ours = os.path.join(PKG, "libsangnom2_b200.so")
    nsrc, total = 8, 1000000
    host = FakeHost(poison_new_frames=False)
    try:
        host.load_plugin(ours)
        src = host.looped_source(w, h, fmt, nsrc, total, parity_mode=2)
        for i in range(nsrc):
            src.set_frame(i, make_frame(1, w, h, fmt, "noise", i))
        flt = host.invoke("SangNom2", src, **kw)
        L = fh._load()
        err = C.create_string_buffer(256)

        def pull(n):
            f = L.fh_get_frame(host.env, flt.handle, n, err, 256)
            if not f:
                raise RuntimeError(err.value.decode())
            L.fh_frame_release(f)

        # warm-up: 16 of the plugin's batches (its default batch: about 1 GB of finished frames, 4..128), so that
        # every recycled frame buffer has been seen twice (and pinned, where the plugin pins); at most ~25 s
        frame_bytes = sum(int(np.prod(fmt.plane_shape(w, h, p))) for p in range(fmt.components)) * fmt.sample_bytes
        batch = int(os.environ.get("SANGNOM_B200_BATCH", max(4, min(128, (1024 << 20) // max(frame_bytes, 1)))))
        n, t_start = 0, time.perf_counter()
        while n < max(96, 16 * batch) and time.perf_counter() - t_start < 25:
            pull(n); n += 1
        warm = n
        t0, n0 = time.perf_counter(), n
        while time.perf_counter() - t0 < seconds and n < total:
            pull(n); n += 1
        dt = time.perf_counter() - t0
        return {"value": (n - n0) / dt, "unit": "frames/s", "frames": n - n0, "warmup_frames": warm,
                "note": "our AviSynth plugin through the fake host: the host's recycled frames, one instance, sequential GetFrame"}
    finally:
        host.close()


def cpu_port_baseline(wl, seconds):
    """Fallback when oracle/_ref is absent: our C restatement, single thread."""
    from oracle import oracle as O
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    fmtname, w, h, kw, _, _ = WORKLOADS[wl]
    fmt = FORMATS[fmtname]
    fr = make_frame(1, w, h, fmt, "noise", 0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds or n < 2:
        O.oracle_frame(fr, fmt.bits, order=kw.get("order", 1), aa=kw.get("aa", 48), aac=kw.get("aac", 0),
                       luma=kw.get("luma", True), chroma=kw.get("chroma", True), parity=n % 2 == 0)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"{n} frames in {dt:.1f} s, scalar C restatement"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p8", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (device-resident leg)")
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--in-flight", type=int, default=int(os.environ.get("SANGNOM_BENCH_INFLIGHT", "0")),
                    help="frames resident on one device in the host path (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the metric's second workload (2160p fp32)")
    ap.add_argument("--plugin-seconds", type=float, default=2.0, help="seconds of the plugin-path leg (0 = skip)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything a library prints there meanwhile (NCCL's version
    # banner, for one) is sent to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


_RESULT_FD = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


if __name__ == "__main__":
    sys.exit(main())
