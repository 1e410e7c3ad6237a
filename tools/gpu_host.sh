#!/bin/bash
# host-path visit: pin cost, host entry variants with different copy-thread counts, plugin path with and without pinning
tag=${1:-h}; o=gpurun_out
( python tools/pin_cost.py
  for t in 8 12 16; do SANGNOM_B200_COPY_THREADS=$t python tools/e2e_probe.py 1080p8 592 pinned; done
  python tools/e2e_probe.py 1080p8 592 inplace pageable field
  python tools/e2e_probe.py 2160pf32 48 pinned inplace
  python tools/plugin_fps.py 1080p8 4096 512
  SANGNOM_B200_PIN_MB=0 python tools/plugin_fps.py 1080p8 2048 512
  SANGNOM_B200_COPY_THREADS=16 python tools/plugin_fps.py 1080p8 4096 512
  python tools/plugin_fps.py 2160pf32 384 64
) 2>&1 | tee $o/host_$tag.log
