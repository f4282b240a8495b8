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


# ---- a user-defined Montgomery curve: the add-on library's ladder ------------------------------------------------------
M383_P = 2**383 - 187


def _ref(name):
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libref_%s.so" % name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    return ctypes.CDLL(path)


def _m383_batch(n):
    k, u = util.random_bytes(38301, n, 48), util.random_bytes(38302, n, 48)
    rows = [0, 1, 12, M383_P - 1, M383_P, M383_P + 1, (1 << 383) - 1, (1 << 384) - 1]
    for i, v in enumerate(rows):
        u[i] = np.frombuffer(v.to_bytes(48, "little"), dtype=np.uint8)
    k[len(rows)] = 0
    k[len(rows) + 1] = 255
    return k, u


def test_user_curve_ladder_vs_reference_build():
    from modarith_b200 import lib as mlib
    from modarith_b200.rfc7748 import rfc7748
    if not os.path.exists(mlib.extra_lib_path("M383")):
        pytest.fail("libmodarith_b200_M383.so is missing: __graft_entry__.build() builds it")
    n = (1 << 14) + 77                      # more than one chunk round of the persistent grid's queues, ragged
    k, u = _m383_batch(n)
    want = util.ref_rfc7748_batch(_ref("M383"), k, u)
    dk, du = torch.from_numpy(k).cuda(), torch.from_numpy(u).cuda()
    got = rfc7748("M383", dk, du)
    assert np.array_equal(got.cpu().numpy(), want)
    # one key per thread, one inversion per key: same bytes
    lib = mlib.load_for("M383")
    out = torch.empty_like(dk)
    mlib.check(lib.mab_M383_rfc7748_perkey(dk.data_ptr(), du.data_ptr(), out.data_ptr(), n,
                                           torch.cuda.current_stream().cuda_stream), "perkey", lib)
    assert np.array_equal(out.cpu().numpy(), want)
    # host buffers: pageable (staged) and pinned (zero-copy)
    assert np.array_equal(rfc7748("M383", k, u), want)
    pk, pu = torch.from_numpy(k).pin_memory(), torch.from_numpy(u).pin_memory()
    assert np.array_equal(rfc7748("M383", pk, pu).numpy(), want)


def test_user_curve_point_validation_vs_reference_build():
    """The driver without TWIST_SECURE (rfc7748.c:228-251) -- what a user-defined curve that is not twist secure runs."""
    from modarith_b200.rfc7748 import rfc7748
    n = 1 << 12
    k, u = _m383_batch(n)
    want = util.ref_rfc7748_batch(_ref("M383_validate"), k, u)
    got = rfc7748("M383", torch.from_numpy(k).cuda(), torch.from_numpy(u).cuda(), validate=True)
    assert np.array_equal(got.cpu().numpy(), want)
    zero = (want == 0).all(axis=1).sum()
    assert n // 4 < zero < 3 * n // 4         # about half of all u are x-coordinates of points on the twist
