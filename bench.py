#!/usr/bin/env python
"""Headline benchmark: SangNom2 frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1080p8|2160pf32|...]

A "step" is one pass of the hot path over one batch of synthetic frames of the workload.
  value  device-resident frames/s: planes already in HBM, kernels launched through the C ABI's device entry
         (sangnom_cuda_process_planes_device), timed with CUDA events on the launching stream; max over ranks.
  e2e    the same metric through the host entry (sangnom_cuda_process_planes) with pinned HOST buffers:
         kept-field upload, kernels and full-frame download all inside the timed region.
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
DEFAULT_FRAMES = {"1080p8": 592, "2160pf32": 148, "2160p10": 148, "480p8": 1184}
E2E_FRAMES = {"1080p8": 592, "2160pf32": 48, "2160p10": 96, "480p8": 1184}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_desc(name):
    fmtname, w, h, kw, dt, idx = WORKLOADS[name]
    args = ", ".join(f"{k}={v}" for k, v in kw.items())
    return f"{w}x{h} {fmtname} SangNom2({args}) [BASELINE.json configs[{idx}]]"


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
        "config": {"workload": workload_desc(wl), "reference_path": "unmodified reference plugin, stock opt=-1 (SSE2), one instance per host thread"},
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pysangnom import cuda
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    from pysangnom.shard import frame_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsangnom_cuda has no CPU path")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cuda.load()

    wl = args.workload
    fmtname, w, h, kw, dtype_tag, _ = WORKLOADS[wl]
    fmt = FORMATS[fmtname]
    sb = fmt.sample_bytes
    F = args.frames or DEFAULT_FRAMES[wl]                 # frames per step per GPU (weak scaling)
    Fe = args.e2e_frames or E2E_FRAMES[wl]
    if world > 1 and not args.e2e_frames:
        Fe = max(8, Fe // 2)                               # several ranks share the host: half the pinned staging per rank
    first, _ = frame_range(F * world, rank, world)        # this rank's contiguous frame range of the global clip
    nplanes = min(fmt.components, 3)
    proc = [kw.get("luma", True)] + [kw.get("chroma", True)] * 2
    thr = [cuda.threshold(a, fmt.bits, sb) for a in (kw.get("aa", 48), kw.get("aac", 0), kw.get("aac", 0))]

    def offset_of(n):
        return cuda.resolve_offset(kw.get("order", 1), n % 2 == 0)

    # ---- device-resident clip: a few distinct seeded frames tiled over the batch ----
    base = [make_frame(1, w, h, fmt, "noise", i) for i in range(4)]
    dev_planes, jobs = [], []
    for k in range(F):
        n = first + k
        for p in range(nplanes):
            if not proc[p]:
                continue
            a = base[n % 4][p]
            ph, pw = a.shape
            pitch = (pw * sb + 255) // 256 * 256
            t = torch.empty((ph, pitch), dtype=torch.uint8, device="cuda")
            t[:, :pw * sb] = torch.from_numpy(a.view(np.uint8).reshape(ph, -1)).cuda()
            dev_planes.append(t)
            jobs.append(cuda.make_job(0, 0, t.data_ptr(), pitch, pw, ph, offset_of(n), cuda.MODE_INPLACE, thr[p], p, n))
    job_arr = (cuda.SnPlaneJob * len(jobs))(*jobs)
    ctx = cuda.Context(sb, w, h, device=local)
    tstream = torch.cuda.Stream()                        # the launching stream; events are recorded on it
    stream = tstream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ctx.process_jobs_device(job_arr, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.cudart().cudaProfilerStart()           # `ncu --profile-from-start off` lists the timed region only
    e0.record(tstream)
    for _ in range(args.steps):
        step()
    e1.record(tstream)
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1)
    launches = ctx.stats()["kernel_launches"]
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = F * world * args.steps / (ms_max / 1000.0)

    # ---- end to end through the host entry, pinned host buffers ----
    # Two sets of pinned arenas (src + dst), planes of consecutive frames back to back - the staging a batching host
    # layer does ("batches prefetched frames into pinned host buffers"); steps alternate between the sets so that
    # step k+1 can be submitted while step k is still downloading (sangnom_cuda_submit / sangnom_cuda_wait).
    frame_bytes = sum(base[0][p].nbytes for p in range(nplanes))
    sets = []
    for _ in range(2):
        src_arena = cuda.PinnedArena(Fe * (frame_bytes + 64) + 4096)
        dst_arena = cuda.PinnedArena(Fe * (frame_bytes + 64) + 4096)
        dst_host, hjobs = [], []
        for k in range(Fe):
            n = first + k
            for p in range(nplanes):
                a = base[n % 4][p]
                s = src_arena.take(a.shape, a.dtype)
                s[...] = a
                d = dst_arena.take(a.shape, a.dtype)
                dst_host.append(d)
                mode = cuda.MODE_FIELD if proc[p] else cuda.MODE_COPY
                hjobs.append(cuda.make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], a.shape[1], a.shape[0],
                                           offset_of(n), mode, thr[p], p, n))
        sets.append(((cuda.SnPlaneJob * len(hjobs))(*hjobs), dst_host, src_arena, dst_arena))
    ectx = cuda.Context(sb, w, h, device=local, max_frames_in_flight=args.in_flight or int(os.environ.get("SANGNOM_BENCH_INFLIGHT", "0")))
    lib = cuda.load()

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.sangnom_cuda_last_error(ectx._h).decode())

    def estep_sync(i):
        arr = sets[i % 2][0]
        check(lib.sangnom_cuda_process_planes(ectx._h, arr, len(arr)))

    for i in range(4):                                  # every pipeline slot has seen (and sized itself for) the largest chunk
        estep_sync(i)
    barrier()
    esteps = max(2, args.steps)
    # (a) one synchronous call per step: every step pays the pipeline's ramp (first upload, last download)
    t0 = time.perf_counter()
    for i in range(esteps):
        estep_sync(i)
    torch.cuda.synchronize()
    sync_s = time.perf_counter() - t0
    barrier()
    # (b) streaming: step i+1 is submitted before step i is waited for - how a frame server keeps the GPU fed
    def stream(job_arrays):
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
        torch.cuda.synchronize()
        return time.perf_counter() - t0, ectx.stats()

    # spot-check an e2e output plane against the device-resident result of the same frame (same bytes expected)
    def spot_check(what):
        chk = sets[(esteps - 1) % 2][1][0]
        ref_dev = dev_planes[0][:, :chk.shape[1] * sb].cpu().numpy().view(chk.dtype)
        if proc[0] and not np.array_equal(ref_dev, chk):
            raise RuntimeError(f"e2e output ({what}) differs from the device-resident output")
        chk[...] = 0

    e_s, est = stream([sets[0][0], sets[1][0]])
    spot_check("frame input")
    # (c) the same frames from a double-rate producer that hands over SEPARATED FIELDS: only the kept
    # field exists on the host, so only it is uploaded (SN_MODE_DH, offset by field parity); output is identical.
    field_s, fst = None, None
    if all(proc[:nplanes]) and not kw.get("dh", False):
        barrier()
        fsets = []
        for i in range(2):
            farena = cuda.PinnedArena(Fe * (frame_bytes // 2 + 64) + 4096)
            fjobs = []
            for j, jb in enumerate(sets[i][0]):
                a = base[(first + j // nplanes) % 4][j % nplanes]
                f = farena.take((a.shape[0] // 2, a.shape[1]), a.dtype)
                f[...] = a[jb.offset::2]
                fjobs.append(cuda.make_job(f.ctypes.data, f.strides[0], jb.dst, jb.dst_pitch, jb.width, jb.dst_height, jb.offset,
                                           cuda.MODE_DH, jb.threshold, jb.plane, jb.frame))
            fsets.append(((cuda.SnPlaneJob * len(fjobs))(*fjobs), farena))
        check(lib.sangnom_cuda_process_planes(ectx._h, fsets[1][0], len(fsets[1][0])))
        field_s, fst = stream([fsets[0][0], fsets[1][0]])
        spot_check("field input")
    t_e = torch.tensor([e_s, sync_s, field_s or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = Fe * world * esteps / float(t_e[0].item())
    e2e_sync_value = Fe * world * esteps / float(t_e[1].item())
    e2e_field_value = Fe * world * esteps / float(t_e[2].item()) if field_s else None
    clocks = sampler.stop() if rank == 0 else None
    plugin_leg = plugin_fps(wl, args.plugin_seconds) if (rank == 0 and world == 1 and args.plugin_seconds > 0) else None

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_file):
            peak, peak_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg = algorithmic_bytes_per_frame(wl)
        achieved = alg * F * args.steps / (ms / 1000.0) / 1e9          # this rank's kernels over its own event time
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(wl)
        out = {
            "metric": "frames_per_second", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype_tag, "data": "synthetic",
            "config": {"workload": workload_desc(wl), "frames_per_step_per_gpu": F, "e2e_frames_per_step_per_gpu": Fe,
                       "l2": f"inputs larger than L2: {alg * F / 1e6:.0f} MB of planes per step per GPU",
                       "sharding": "contiguous frame ranges per rank, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg * F / max(1, launches // args.steps),
                         "note": "kernel is ALU-issue bound, not HBM bound (DESIGN.md, Roofline)"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": est["h2d_bytes"] // esteps,
                    "d2h_bytes_per_step": est["d2h_bytes"] // esteps, "steps": esteps,
                    "api": "sangnom_cuda_submit/_wait, two steps in flight, pinned host arenas",
                    "sync_call_value": e2e_sync_value,
                    "plugin_value": plugin_leg,
                    "field_input": None if e2e_field_value is None else {
                        "value": e2e_field_value, "h2d_bytes_per_step": fst["h2d_bytes"] // esteps, "d2h_bytes_per_step": fst["d2h_bytes"] // esteps,
                        "note": "separated-field input (SN_MODE_DH): same output frames, half the upload"}},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            if _ORIGINAL_AFFINITY:
                os.sched_setaffinity(0, _ORIGINAL_AFFINITY)      # the CPU baseline gets every host core
            cores = len(os.sched_getaffinity(0))
            ref = reference_fps(wl, cores, args.cpu_seconds)
            if ref is not None:
                ref0 = reference_fps(wl, cores, args.cpu_seconds / 2, opt=0)
                out["cpu_baseline"] = {"value": ref["fps"], "unit": "frames/s", "cores": cores, "kind": "reference",
                                       "sample": f"{ref['frames']} frames of the same workload in {ref['seconds']:.1f} s, one reference "
                                                 f"instance per host thread, stock opt=-1 (SSE2)",
                                       "opt0_cpp_path_fps": ref0["fps"]}
            else:
                out["cpu_baseline"] = cpu_port_baseline(wl, args.cpu_seconds)
        emit(out)
    ctx.close()
    ectx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def plugin_fps(wl, seconds):
    """The drop-in path itself: our AviSynth plugin pulled frame by frame through the fake host (PAGEABLE host frames,
    one filter instance, sequential GetFrame) - the same harness the reference arm is timed with."""
    from pysangnom.clips import make_frame
    from pysangnom.formats import FORMATS
    from fakehost import FakeHost
    import fakehost as fh
    fmtname, w, h, kw, _, _ = WORKLOADS[wl]
    fmt = FORMATS[fmtname]
    ours = os.path.join(PKG, "libsangnom2_b200.so")
    nsrc, total = 8, 100000
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

        n = 0
        while n < 96:
            pull(n); n += 1
        t0, n0 = time.perf_counter(), n
        while time.perf_counter() - t0 < seconds and n < total:
            pull(n); n += 1
        dt = time.perf_counter() - t0
        return {"value": (n - n0) / dt, "unit": "frames/s", "frames": n - n0,
                "note": "our AviSynth plugin through the fake host: pageable frames, one instance, sequential GetFrame"}
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
    ap.add_argument("--in-flight", type=int, default=0, help="frames resident on the device in the host path (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
