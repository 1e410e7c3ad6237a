#!/bin/bash
# one full ncu capture of the u8 kernel (3 launches = Y, U, V pass of 296 frames), after a clean plain run
tag=${1:-p}
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 --frames 296 > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o gpurun_out/prof_u8_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 --frames 296 > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log
