#!/bin/bash
# One GPU visit: parity tests, the headline bench, the other workloads, launch list, full ncu captures.
tag=${1:-r01c}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $o/pytest_gpu_$tag.log
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"
for wl in 2160pf32 2160p10 480p8; do
  python bench.py --workload $wl --steps 10 --cpu-seconds 6 > $o/bench_${wl}_$tag.json 2> $o/bench_${wl}_$tag.err; echo "bench $wl rc=$?"
done
# launch list of the headline command (short), only after the plain run above exited 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 16 > $o/ncu_launches_$tag.log 2>&1
# one full capture per kernel flavour: 3 launches (Y,U,V pass) after the warm-up launches
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_u8_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 > $o/ncu_full_u8_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_f32_$tag \
    python bench.py --workload 2160pf32 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4 > $o/ncu_full_f32_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_u16_$tag \
    python bench.py --workload 2160p10 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4 > $o/ncu_full_u16_$tag.log 2>&1
cat $o/bench_$tag.json
for wl in 2160pf32 2160p10 480p8; do cut -c1-400 $o/bench_${wl}_$tag.json; done
