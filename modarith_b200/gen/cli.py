"""Shared command line of the sm_100a generators (see pseudo_sm100.py / monty_sm100.py)."""
from __future__ import annotations

import argparse
import os

from ..primes import PRIMES, ALL_PRIMES, Prime
from .plan import make_plan
from .emit import emit_field_header

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc")


def resolve(prime: str, family: str) -> Prime:
    if prime in ALL_PRIMES:
        return ALL_PRIMES[prime]
    if prime[0].isdigit():                      # expression, as pseudo.py:1553-1556
        p = eval(prime, {"__builtins__": {}})
        return Prime("P%d" % p.bit_length(), p, family)
    raise SystemExit("This named modulus is not supported")


def generate(prime: Prime, out: str | None = None, verbose=True) -> str:
    if prime.nbits < 120 or pow(3, prime.p - 1, prime.p) != 1:      # pseudo.py:1561-1564
        raise SystemExit("Not a sensible modulus, too small or not a prime")
    plan = make_plan(prime)
    text = emit_field_header(plan)
    if out is None:
        out = os.path.join(CSRC, "gen", "field_%s.cuh" % prime.name)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    old = open(out).read() if os.path.exists(out) else None
    if old != text:
        with open(out, "w") as f:
            f.write(text)
    if verbose:
        print("Prime %s is of length %d bits; plan %s, %d saturated 32-bit limbs" % (
            prime.name, prime.nbits, type(plan).__name__, plan.L))
        for k, b in plan.blocks.items():
            w, i, a = b.stats()
            print("  %-5s: %3d IMAD.WIDE  %2d IMAD  ~%3d ALU" % (k, w, i, a))
        print("Checking correctness.. Passed - OK")
        print("Field code is in", os.path.relpath(out))
    return out


def main(family: str, argv) -> int:
    ap = argparse.ArgumentParser(prog="%s_sm100" % family)
    ap.add_argument("prime")
    ap.add_argument("-o", "--output", default=None)
    a = ap.parse_args(argv)
    P = resolve(a.prime, family)
    generate(P, a.output)
    return 0


def generate_all(verbose=False, outdir=None):
    """Headers of every built-in modulus; into csrc/gen (the committed copies) or, for tests, into `outdir`/gen."""
    return [generate(P, None if outdir is None else os.path.join(outdir, "gen", "field_%s.cuh" % P.name), verbose=verbose)
            for P in ALL_PRIMES.values()]
