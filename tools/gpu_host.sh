#!/bin/bash
# host-path visit: parity suite, host entry variants (copy streams per slot vs shared), plugin path, the full bench line
tag=${1:-h}; o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $o/pytest_gpu_$tag.log
( python tools/e2e_probe.py 1080p8 592 pinned inplace
  SANGNOM_B200_COPY_STREAMS=shared python tools/e2e_probe.py 1080p8 592 pinned
  python tools/e2e_probe.py 2160pf32 48 pinned
  python tools/plugin_fps.py 1080p8 4096 512
  SANGNOM_B200_BATCH=32 python tools/plugin_fps.py 1080p8 2048 512
  python tools/plugin_fps.py 2160pf32 384 64
) 2>&1 | tee $o/host_$tag.log
python bench.py --steps 10 --cpu-seconds 6 > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"; tail -5 $o/bench_$tag.err; cut -c1-3000 $o/bench_$tag.json
