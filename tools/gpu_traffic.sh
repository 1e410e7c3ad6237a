#!/bin/bash
# DRAM traffic per launch (dram__bytes_read + write) of the Y, U and V pass of one step of each workload: small CSVs
tag=${1:-t}; o=gpurun_out
python tools/devtime.py 1080p8 2160p10 2160pf32 > $o/devtime_plain_traffic_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $o/devtime_plain_traffic_$tag.log; exit 1; }
cat $o/devtime_plain_traffic_$tag.log
for wl in 1080p8 2160p10 2160pf32; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sangnom -s 9 -c 3 --csv --log-file $o/traffic_${wl}_$tag.csv python tools/devtime.py $wl > $o/ncu_traffic_${wl}_$tag.log 2>&1; echo "ncu $wl rc=$?"
done
