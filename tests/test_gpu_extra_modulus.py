"""The add-on library of a modulus outside the built-in five (python -m modarith_b200.build --prime NIST384) on the
GPU: the same entry points, checked byte-for-byte against the reference's own generated C for that modulus
(`monty.py 64 NIST384`, oracle/_ref/libref_NIST384.so) and against the value-level oracle."""
import os
import random
import sys

import numpy as np
import pytest
import torch

import util
from field_oracle import FieldOracle
from oracle_primes import OraclePrime

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
P384 = 2**384 - 2**128 - 2**96 + 2**32 - 1            # monty.py named table, "NIST384"


@pytest.fixture(scope="module")
def F():
    from modarith_b200 import Field, lib as mlib
    if not os.path.exists(mlib.extra_lib_path("NIST384")):
        pytest.fail("libmodarith_b200_NIST384.so is missing: __graft_entry__.build() builds it")
    return Field("NIST384")


def _bytes(t):
    return t.cpu().numpy()


def test_field_ops_vs_reference_build(F):
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libref_NIST384.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    ref = ctypes.CDLL(path)
    nb, n = F.Nbytes, 1 << 12
    assert nb == 48 and F.Nlimbs == 12
    a, b = util.random_bytes(384, n, nb), util.random_bytes(385, n, nb)
    for i, v in enumerate([P384 - 1, P384, P384 + 1, (1 << 384) - 1, 0, 1]):
        a[i] = np.frombuffer(v.to_bytes(nb, "big"), dtype=np.uint8)
    x, st = F.modimp(torch.from_numpy(a).cuda())
    y, _ = F.modimp(torch.from_numpy(b).cuda())
    r = F.alloc(n)
    for op in ("mul", "sqr", "inv", "sqrt", "add", "sub"):
        want, wst = util.ref_field_batch(ref, op, a, b if op in ("mul", "add", "sub") else None)
        if op == "mul": F.modmul(x, y, r)
        if op == "sqr": F.modsqr(x, r)
        if op == "inv": F.modinv(x, None, r)
        if op == "sqrt": F.modsqrt(x, None, r)
        if op == "add": F.modadd(x, y, r)
        if op == "sub": F.modsub(x, y, r)
        assert np.array_equal(_bytes(F.modexp(r)), want), op
        assert np.array_equal(st.cpu().numpy(), wst)


def test_api_vs_oracle_and_programs(F):
    O = FieldOracle(OraclePrime("NIST384", P384))
    p = O.p
    rng = random.Random(384)
    xs = [0, 1, p - 1, p - 2] + [rng.randrange(p) for _ in range(200)]
    ys = [p - 1, 0, 1, 2] + [rng.randrange(p) for _ in range(200)]
    x, y = F.from_ints(xs), F.from_ints(ys)
    r = F.alloc(len(xs))
    F.modmul(x, y, r); assert F.to_ints(r) == [O.modmul(a, b) for a, b in zip(xs, ys)]
    F.modsub(x, y, r); assert F.to_ints(r) == [O.modsub(a, b) for a, b in zip(xs, ys)]
    F.modmli(x, 39081, r); assert F.to_ints(r) == [O.modmli(a, 39081) for a in xs]
    F.modcpy(x, r); F.modhaf(r); assert F.to_ints(r) == [O.modhaf(a) for a in xs]
    F.modpro(x, r); assert F.to_ints(r) == [O.modpro(a) for a in xs]
    assert F.modqr(None, x).cpu().tolist() == [O.modqr(None, a) for a in xs]
    code = [("mul", 2, 0, 1), ("add", 3, 2, 0), ("sqr", 3, 3, 0), ("sub", 4, 3, 1), ("mli", 5, 4, 0, 7), ("neg", 6, 5, 0)]
    want = [(-(7 * ((a * b + a) ** 2 - b))) % p for a, b in zip(xs, ys)]
    for jit in (False, True):
        (res,) = F.modprog(code, [x, y], [6], jit=jit)
        assert F.to_ints(res) == want, jit
