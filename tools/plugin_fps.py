#!/usr/bin/env python
"""Frames/s through the AviSynth plugin surface (fake host, GetFrame pulls) for our plugin: the drop-in path with the
host's own (pageable, recycled) frames, as a frame server would drive it. One filter instance, sequential pulls.
Prints frames/s per window of frames, so that the one-time cost of pinning the recycled frame buffers is visible.
   python tools/plugin_fps.py [workload] [frames] [window]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200"), os.path.join(ROOT, "tests")]
import bench
from pysangnom.clips import make_frame
from pysangnom.formats import FORMATS
from fakehost import FakeHost
import fakehost as fh

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p8"
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
window = int(sys.argv[3]) if len(sys.argv) > 3 else 512
fmtname, w, h, kw, _, _ = bench.WORKLOADS[wl]
fmt = FORMATS[fmtname]
OURS = os.path.join(ROOT, "avisynth-sangnom2_b200", "libsangnom2_b200.so")
host = FakeHost(poison_new_frames=False)
host.load_plugin(OURS)
nsrc = 8
src = host.looped_source(w, h, fmt, nsrc, nframes, parity_mode=2)
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
rates = []
t0 = time.perf_counter()
for n in range(nframes):
    pull(n)
    if (n + 1) % window == 0:
        t1 = time.perf_counter(); rates.append(round(window / (t1 - t0), 1)); t0 = t1
print(json.dumps({"workload": wl, "frames": nframes, "window": window, "fps_per_window": rates,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("SANGNOM_")}}))
host.close()
