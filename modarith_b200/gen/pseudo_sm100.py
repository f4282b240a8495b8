"""sm_100a backend for pseudo-Mersenne moduli 2^n - c -- sits beside the reference's
pseudo.py (C), pseudo_rust.py (Rust) and simd/pseudo_cuda.py (CUDA demo).

    python -m modarith_b200.gen.pseudo_sm100 X25519 [-o field.cuh]

Like pseudo.py (pseudo.py:1461-1473) it takes a prime name (or an expression such as
2**255-19), chooses a limb plan, self-tests the generated arithmetic against Python bignums
(the reference does this through test.so + ctypes, pseudo.py:1694-1855; here through the
PTX interpreter of gen/ptx.py because there is no GPU on the build machine) and writes the
field code.  Word length is fixed at 32 (the native multiplier width of the SM) as in
simd/pseudo_cuda.py:1328.
"""
import sys

from .cli import main

if __name__ == "__main__":
    sys.exit(main("pseudo", sys.argv[1:]))
