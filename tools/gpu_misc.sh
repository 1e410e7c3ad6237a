#!/bin/bash
# misc visit: full parity suite, turn bandwidth (TMA and plain kernel), host entry with merged 3-D transfers, plugin, AA chain
tag=${1:-m}; o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 $o/pytest_gpu_$tag.log
( python tools/turn_bw.py
  SANGNOM_TURN=plain python tools/turn_bw.py
  for t in 8 16; do SANGNOM_B200_COPY_THREADS=$t python tools/e2e_probe.py 1080p8 592 pinned; done
  python tools/e2e_probe.py 1080p8 592 inplace field pageable
  python tools/e2e_probe.py 2160pf32 48 pinned inplace
  python tools/plugin_fps.py 1080p8 4096 512
  python tools/plugin_fps.py 2160pf32 384 64
  python tools/aa_chain_bench.py 144
) 2>&1 | tee $o/misc_$tag.log
