#!/bin/bash
# One full single-GPU visit: parity tests, the headline bench (with its secondary workload), the other workloads, the
# reference arm, AA chain and turn-kernel numbers, the launch list of the timed region.
tag=${1:-r02}
o=gpurun_out
python -m pytest tests -m gpu -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $o/pytest_gpu_$tag.log
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"; tail -2 $o/bench_$tag.err
for wl in 2160p10 480p8; do
  python bench.py --workload $wl --steps 10 --cpu-seconds 6 > $o/bench_${wl}_$tag.json 2> $o/bench_${wl}_$tag.err; echo "bench $wl rc=$?"
done
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err; echo "ref rc=$?"
python tools/aa_chain_bench.py 144 > $o/aa_chain_$tag.json 2> $o/aa_chain_$tag.err; echo "chain rc=$?"
python tools/turn_bw.py > $o/turn_bw_$tag.json 2>&1
# launch list of the headline command (short), only after the plain run above exited 0
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-frames 16 --plugin-seconds 0 > $o/ncu_launches_$tag.log 2>&1
cut -c1-1200 $o/bench_$tag.json; echo; cat $o/turn_bw_$tag.json; cat $o/aa_chain_$tag.json
for wl in 2160p10 480p8; do cut -c1-300 $o/bench_${wl}_$tag.json; echo; done
