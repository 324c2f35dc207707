#!/bin/bash
# Concurrent pinned host<->device copies on 1, 2, 4, 8 GPUs (one process each, all started together): the aggregate
# bounds bench.py's end-to-end figure.  usage: profiles/pcie_multi.sh [GiB per direction] > out.jsonl
GIB=${1:-2}
NG=$(nvidia-smi -L | wc -l)
for G in 1 2 4 8; do
  [ $G -gt $NG ] && break
  rm -f /tmp/pcie_$G.*.json
  for ((i=0; i<G; i++)); do CUDA_VISIBLE_DEVICES=$i python profiles/pcie_probe.py $GIB > /tmp/pcie_$G.$i.json & done
  wait
  python - <<PY
import json, glob
rows = [json.load(open(f)) for f in sorted(glob.glob("/tmp/pcie_$G.*.json"))]
agg = {k: sum(r[k] for r in rows) for k in rows[0] if k.endswith("GBps") or k.endswith("GBps_each")}
print(json.dumps({"concurrent_gpus": $G, "sum_over_gpus": agg, "per_gpu_min": {k: min(r[k] for r in rows) for k in agg}}))
PY
done
