"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, opt=0).

Run in the container that has /root/reference:  python tests/golden/make_golden.py
For every case in tests/helpers.py:CASES the seeded input frames are regenerated, pushed through the
compiled reference plugin by the fake AviSynth host (fresh filter instance per frame = the parity
contract) and the SHA-256 of every input and output plane is recorded. Tiny cases also keep the full
output arrays. The GPU box has no /root/reference: there the tests compare against these fixtures.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "avisynth-sangnom2_b200"), os.path.join(ROOT, "tests")]

from helpers import CASES, case_frames, parity_of  # noqa: E402
from oracle import oracle as O  # noqa: E402
from fakehost import FakeHost  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_outputs(fmt, w, h, frames, kw, func="SangNom2"):
    outs = []
    with FakeHost(cpu_flags=0) as host:           # no SSE2 flag: the reference takes its C++ path
        host.load_plugin(O.reference_plugin_path())
        src = host.source(w, h, fmt, len(frames), parity_mode=2)
        for i, fr in enumerate(frames):
            src.set_frame(i, fr)
        for i in range(len(frames)):
            flt = host.invoke(func, src, **kw)    # fresh instance (zero-filled pool) per frame
            outs.append(flt.get_frame(i)[:3])
    return outs


def main():
    assert O.reference_plugin_path(), "build oracle/_ref first (make -f oracle/Makefile)"
    doc = {"generator": "tests/golden/make_golden.py", "reference": "Asd-g/AviSynth-SangNom2 v0.6.1 opt=0 (C++ path)", "cases": {}}
    for case in CASES:
        name, fmtname, w, h, kw, kind, nframes = case
        fmt, frames = case_frames(case)
        outs = reference_outputs(fmt, w, h, frames, kw)
        entry = {"format": fmtname, "width": w, "height": h, "args": kw, "kind": kind,
                 "input_sha256": [[sha(p) for p in fr] for fr in frames],
                 "output_sha256": [[sha(p) for p in fr] for fr in outs]}
        if w * h <= 4096:
            entry["output"] = [[p.tolist() for p in fr] for fr in outs]
        doc["cases"][name] = entry
        print(name, "ok")
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
