#!/bin/bash
# Full ncu captures of the row-sweep kernels on tools/devtime.py (one workload per capture: passes of one step after
# the warm-up launches). gpurun merges at most 64 MiB back: one part per call.   tools/gpu_ncu2.sh <tag> wide|u8
# Each capture runs only after the same command exited 0 without ncu.
tag=${1:-r02}; part=${2:-wide}
o=gpurun_out
case $part in
wide)
  python tools/devtime.py 2160p10 2160pf32 > $o/devtime_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $o/devtime_plain_$tag.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 2 -f -o $o/prof_u16_$tag python tools/devtime.py 2160p10 > $o/ncu_u16_$tag.log 2>&1; echo "ncu u16 rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 1 -f -o $o/prof_f32_$tag python tools/devtime.py 2160pf32 > $o/ncu_f32_$tag.log 2>&1; echo "ncu f32 rc=$?" ;;
u8)
  python tools/devtime.py 1080p8 > $o/devtime_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $o/devtime_plain_$tag.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_u8_$tag python tools/devtime.py 1080p8 > $o/ncu_u8_$tag.log 2>&1; echo "ncu u8 rc=$?" ;;
esac
cat $o/devtime_plain_$tag.log; ls -la $o/*_$tag.ncu-rep
