#!/usr/bin/env python
"""Frames/s through the AviSynth plugin surface (fake host, GetFrame pulls) for our plugin: the drop-in path with
PAGEABLE host frames, as a frame server would drive it. One filter instance (MT_NICE_FILTER), sequential pulls."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200"), os.path.join(ROOT, "tests")]
import bench
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS
from fakehost import FakeHost
import fakehost as fh

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p8"
batch = sys.argv[2] if len(sys.argv) > 2 else "64"
nframes = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
os.environ["SANGNOM_B200_BATCH"] = batch
fmtname, w, h, kw, _, _ = bench.WORKLOADS[wl]
fmt = FORMATS[fmtname]
OURS = os.path.join(ROOT, "avisynth-sangnom2_b200", "libsangnom2_b200.so")
host = FakeHost(poison_new_frames=False)
host.load_plugin(OURS)
nsrc = 8
src = host.source(w, h, fmt, nframes, parity_mode=2)
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
for n in range(int(batch) * 2):
    pull(n)
t0 = time.perf_counter()
for n in range(int(batch) * 2, nframes):
    pull(n)
dt = time.perf_counter() - t0
print(json.dumps({"workload": wl, "batch": int(batch), "plugin_fps": (nframes - int(batch) * 2) / dt}))
