#!/usr/bin/env python3
"""Time ecnmul (P-256, Ed25519) for the default library and every built variant (tools/variants.py).

    python tools/bench_ecn.py            # on the GPU box; writes gpurun_out/ecn_variants.json

Each library runs in its own process (MODARITH_B200_LIB); besides the rate it prints a SHA-256 of
the outputs, which must be the same for every variant (the default build is the one `pytest -m gpu`
checks against the goldens)."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    from modarith_b200.ecn import ecnmul
    from modarith_b200.primes import PRIMES, X25519
    n = 1 << 18
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    e = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).to(dev)
    out = {}
    P = PRIMES["NIST256"]
    for curve, gx, gy in (("NIST256", P.wgx, P.wgy), ("ED25519", X25519.ed_gx, X25519.ed_gy)):
        x = torch.from_numpy(np.tile(np.frombuffer(gx.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        y = torch.from_numpy(np.tile(np.frombuffer(gy.to_bytes(32, "big"), dtype=np.uint8), (n, 1))).to(dev)
        xo, yo = ecnmul(curve, e, x, y)
        torch.cuda.synchronize()
        h = hashlib.sha256(xo.cpu().numpy().tobytes() + yo.cpu().numpy().tobytes()).hexdigest()[:16]
        best = 1e9
        for _ in range(3):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            ecnmul(curve, e, x, y)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1) * 1e-3)
        out[curve] = {"Mops": round(n / best / 1e6, 3), "sha": h}
        from modarith_b200.ecn import ecnmul2
        n2 = n // 2
        a2 = (e[:n2], x[:n2], y[:n2], e[n2:2 * n2], xo[:n2].contiguous(), yo[:n2].contiguous())
        x2, y2 = ecnmul2(curve, *a2)
        torch.cuda.synchronize()
        h2 = hashlib.sha256(x2.cpu().numpy().tobytes() + y2.cpu().numpy().tobytes()).hexdigest()[:16]
        best = 1e9
        for _ in range(3):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            ecnmul2(curve, *a2)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1) * 1e-3)
        out[curve + "_mul2"] = {"Mops": round(n2 / best / 1e6, 3), "sha": h2}
    print(json.dumps(out))


def main():
    vdir = os.path.join(ROOT, "modarith_b200", "build", "variants")
    libs = {"default": None}
    if os.path.isdir(vdir):
        for tag in sorted(os.listdir(vdir)):
            lib = os.path.join(vdir, tag, "libmodarith_b200.so")
            if os.path.exists(lib):
                libs[tag] = lib
    res = {}
    for tag, lib in libs.items():
        env = dict(os.environ)
        if lib:
            env["MODARITH_B200_LIB"] = lib
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        try:
            res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res[tag] = {"error": r.stderr[-500:]}
        print(tag, res[tag], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ecn_variants.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
