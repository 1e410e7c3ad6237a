#!/bin/bash
# end-to-end experiments: DMA merging, chunk size, kept-only upload
o=gpurun_out; tag=${1:-e2e}
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'])"; }
( run A=0
  run A=0 SANGNOM_BENCH_INFLIGHT=222
  run A=0 SANGNOM_BENCH_INFLIGHT=111
  run SANGNOM_UPLOAD=field
  run SANGNOM_UPLOAD=field SANGNOM_BENCH_INFLIGHT=222
  SANGNOM_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "chunk of" | tail -8
  SANGNOM_TRACE=1 SANGNOM_UPLOAD=field python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "chunk of" | tail -8
) 2>&1 | tee $o/e2e_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 16 > $o/ncu_launches_$tag.log 2>&1
