#!/bin/bash
# tuning-knob sweep: column segment per block (cluster split) for the 8-bit and the 16-bit/fp32 kernel
for seg in 2048 1024 512; do
  echo "U8_SEG=$seg"; SANGNOM_U8_SEG=$seg python tools/quick_time2.py 2>&1 | sed -n 3p
done
for seg in 1024 512 256; do
  for wl in 2160pf32 2160p10; do
    echo -n "WIDE_SEG=$seg "; SANGNOM_WIDE_SEG=$seg python bench.py --workload $wl --steps 5 --no-cpu-baseline --e2e-frames 8 --plugin-seconds 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$wl', round(d['value']), 'fps')"
  done
done
