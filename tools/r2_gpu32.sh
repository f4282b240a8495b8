#!/bin/bash
# round 2, GPU call 32 (8 GPUs): short batches across 8 GPUs with the final tree: 2^20 and 2^21 keys in total
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_n8_short.jsonl
for PER in 131072 262144; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 --keys $PER --no-extra --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r2_n8_short.jsonl
done
python - <<PY
import json
for line in open("gpurun_out/r2_n8_short.jsonl"):
    j = json.loads(line)
    print(j["n_gpus"], j["config"]["keys_total"], "%.1f M/s" % (j["value"] / 1e6), "e2e %.1f M/s" % (j["e2e"]["value"] / 1e6), "parity", j["parity_spot_check"], j["parity_keys"])
PY
