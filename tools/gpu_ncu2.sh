#!/bin/bash
# Full ncu captures of the kernels on tools/devtime.py / tools/turn_bw.py: the Y, U and V pass of one step after the
# warm-up launches. gpurun merges at most 64 MiB back: one part per call.   tools/gpu_ncu2.sh <tag> u8|u16|f32|turn
# Each capture runs only after the same command exited 0 without ncu.
tag=${1:-r02}; part=${2:-u8}
o=gpurun_out
case $part in
u8)  wl=1080p8 ;;
u16) wl=2160p10 ;;
f32) wl=2160pf32 ;;
turn)
  python tools/turn_bw.py > $o/turn_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $o/turn_plain_$tag.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:turn -s 3 -c 2 -f -o $o/prof_turn_$tag python tools/turn_bw.py > $o/ncu_turn_$tag.log 2>&1; echo "ncu turn rc=$?"
  cat $o/turn_plain_$tag.log; ls -la $o/prof_turn_$tag.ncu-rep; exit 0 ;;
esac
python tools/devtime.py $wl > $o/devtime_plain_${part}_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $o/devtime_plain_${part}_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:sangnom -s 9 -c 3 -f -o $o/prof_${part}_$tag python tools/devtime.py $wl > $o/ncu_${part}_$tag.log 2>&1; echo "ncu $part rc=$?"
cat $o/devtime_plain_${part}_$tag.log; ls -la $o/prof_${part}_$tag.ncu-rep
