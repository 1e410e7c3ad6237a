#!/bin/bash
# One gpurun call: GPU tests, PCIe bandwidth, bench (both arms), launch list and one full ncu capture of the u8 kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_baseline.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python tools/pcie_bw.py > gpurun_out/pcie_bw_$tag.json 2> gpurun_out/pcie_bw_$tag.err; cat gpurun_out/pcie_bw_$tag.json
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cat gpurun_out/bench_ref_$tag.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 16 > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 16 > gpurun_out/ncu_launches_$tag.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 --frames 296 > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o gpurun_out/prof_u8_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 --frames 296 > gpurun_out/ncu_full_$tag.log 2>&1
echo done
