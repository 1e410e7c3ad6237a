#!/bin/bash
# quick GPU check: parity tests, then kernel timing
tag=${1:-q}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$tag.log
python tools/quick_time2.py 2>&1 | tee gpurun_out/quick_$tag.log
