#!/bin/bash
# BASELINE configs[4]: X25519 batch sweep 2^20 .. 2^28 keys in total, sharded over N GPUs by contiguous
# key range.  usage: tools/sweep.sh N   (run on the GPU box; writes gpurun_out/sweep_nN.jsonl)
N=${1:-1}
OUT=gpurun_out/sweep_n$N.jsonl
: > $OUT
for LG in ${SWEEP_LGS:-20 22 24 26 28}; do
  PER=$(( (1 << LG) / N ))
  PK=$(( 1048576 / N ))          # a fixed 2^20 sub-sample of every run is compared with the reference build
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --keys $PER --no-extra --no-cpu-baseline --parity-keys $PK >> $OUT
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 3 --warmup 3 --keys $PER --no-extra --no-cpu-baseline --parity-keys $PK 2>/dev/null | tail -1 >> $OUT
  fi
done
python - <<PY
import json
for line in open("$OUT"):
    j = json.loads(line)
    print(j["n_gpus"], j["config"]["keys_total"], "%.1f M/s" % (j["value"] / 1e6), "e2e %.1f M/s" % (j["e2e"]["value"] / 1e6),
          "frac %.3f" % j["roofline"]["frac"], "parity", j["parity_spot_check"], j["parity_keys"])
PY
