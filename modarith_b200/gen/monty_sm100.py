"""sm_100a backend for generalised-Mersenne / Montgomery-friendly moduli -- sits beside the
reference's monty.py and simd/monty_cuda.py.

    python -m modarith_b200.gen.monty_sm100 X448 [-o field.cuh]
    python -m modarith_b200.gen.monty_sm100 NIST256

Shape handling (monty.py:258-298 `process_prime`) is done on the saturated radix-2^32
representation: 2^448-2^224-1 reduces by half-length additions (gen/plan.py GenMersenne),
P-256 by a multiplication-free whole-quotient Montgomery reduction (gen/plan.py Montgomery).
"""
import sys

from .cli import main

if __name__ == "__main__":
    sys.exit(main("monty", sys.argv[1:]))
