#!/bin/bash
# round 2, GPU call 13 (1 GPU): add-on libraries (NIST384 field, M383 user curve with ladder), compiled programs
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_extra_modulus.py tests/test_gpu_modprog.py -x -q 2>&1 | tail -12 > gpurun_out/r2_gpu13_pytest.txt; cat gpurun_out/r2_gpu13_pytest.txt
python - <<'PY'
import torch, time, numpy as np
from modarith_b200.rfc7748 import rfc7748
n = 1 << 19
g = torch.Generator(device="cuda"); g.manual_seed(1)
k = torch.randint(0, 256, (n, 48), dtype=torch.uint8, device="cuda", generator=g)
u = torch.randint(0, 256, (n, 48), dtype=torch.uint8, device="cuda", generator=g)
out = torch.empty_like(k)
rfc7748("M383", k, u, out); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): rfc7748("M383", k, u, out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3e3
print("M383 ladder (user curve, 12 limbs, add-on library): %.2f M scalar-mults/s at 2^19 keys" % (n / t / 1e6))
PY
