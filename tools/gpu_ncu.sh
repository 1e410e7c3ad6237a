#!/bin/bash
# Full ncu captures (after tools/gpu_round.sh has shown the plain runs exit 0). gpurun merges at most 64 MiB back, so
# run it once per part:  tools/gpu_ncu.sh <tag> u8 | wide | turn
tag=${1:-r01}; part=${2:-u8}
o=gpurun_out
case $part in
u8)   # 3 launches = Y, U, V pass of one bench step, after the warm-up launches
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_u8_$tag \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 16 --plugin-seconds 0 > $o/ncu_full_u8_$tag.log 2>&1 ;;
wide) # the Y pass of the fp32 and of the 10-bit workload
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 1 -f -o $o/prof_f32_$tag \
      python bench.py --workload 2160pf32 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4 --plugin-seconds 0 > $o/ncu_full_f32_$tag.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 1 -f -o $o/prof_u16_$tag \
      python bench.py --workload 2160p10 --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 4 --plugin-seconds 0 > $o/ncu_full_u16_$tag.log 2>&1 ;;
turn)
  ncu --set full --clock-control none -k regex:turn -c 2 -f -o $o/prof_turn_$tag python tools/aa_chain_bench.py 12 > $o/ncu_turn_$tag.log 2>&1 ;;
esac
ls -la $o/*.ncu-rep
