#!/bin/bash
# wide-kernel check: parity tests, then device-resident timing of the 16-bit and fp32 workloads
tag=${1:-w}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in 2160pf32 2160p10; do
  python bench.py --workload $wl --steps 5 --no-cpu-baseline --e2e-frames 8 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$wl', round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s')"
done 2>&1 | tee gpurun_out/wide_$tag.log
