#!/bin/bash
# multi-GPU visit (gpurun --gpus N): multi-device parity tests, then the bench at N ranks
n=${1:-8}; tag=${2:-r02}; o=gpurun_out
python -m pytest tests -m gpu -q -k "multi_device or all_devices or sharding" > $o/pytest_multi_n${n}_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $o/pytest_multi_n${n}_$tag.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n > $o/bench_n${n}_$tag.json 2> $o/bench_n${n}_$tag.err; echo "bench rc=$?"; tail -2 $o/bench_n${n}_$tag.err
python - <<PY
import json
d=json.load(open("$o/bench_n${n}_$tag.json"))
e=d["e2e"]; s=d.get("secondary",{})
print("N=$n value",round(d["value"]),"e2e",round(e["value"]),"inplace",round(e.get("inplace",{}).get("value",0)),"per-rank ctx",round(e.get("one_context_per_rank_value",0)),
      "pcie roof",round(e["pcie_peak_gbs"],1),"frac",round(e["frac"],3),"host mem GB/s",round(e["host_memory_traffic_gbs"],1))
if s: print("  secondary value",round(s["value"]),"e2e",round(s["e2e"]["value"]),"frac",round(s["e2e"]["frac"],3))
PY
nvidia-smi topo -m > $o/topo_n${n}_$tag.txt 2>&1; lscpu | head -20 > $o/lscpu_n${n}_$tag.txt; free -g | head -2 >> $o/lscpu_n${n}_$tag.txt
