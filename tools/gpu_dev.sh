#!/bin/bash
# kernel iteration visit: GPU parity suite, then device-resident timing of the three bench workloads
tag=${1:-d}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 $o/pytest_gpu_$tag.log
python tools/devtime.py 1080p8 2160pf32 2160p10 2>&1 | tee $o/devtime_$tag.log
