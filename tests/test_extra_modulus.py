"""`python -m modarith_b200.build --prime NAME[=EXPR]`: the add-on library for a modulus that is not one of the five
built in (the reference's "run the generator on another prime, compile what it printed", pseudo.py:1553-1564,
1694-1702).  CPU side: the library builds for sm_100a, exports the whole field ABI for its modulus, and every other
named modulus of the reference's tables resolves to a plan."""
import ctypes
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def nist384():
    from modarith_b200 import build
    return build.build_extra("NIST384", verbose=False)


def test_addon_library_exports_the_field_abi(nist384):
    from modarith_b200 import lib as mlib
    dll = ctypes.CDLL(nist384)
    names = ["mab_NIST384_%s" % n for n in mlib.FIELD_SIGNATURES] + ["mab_NIST384_modprog", "mab_NIST384_modprog_jit",
                                                                      "mab_NIST384_modprog_cubin"]
    for n in names + ["mab_error_string", "mab_jit_log", "mab_version"]:
        assert hasattr(dll, n), n
    par = mlib.params("NIST384")
    assert (par["nlimbs"], par["nbits"], par["nbytes"], par["montgomery"], par["has_curve"]) == (12, 384, 48, 1, 0)


def test_addon_programs_compile(nist384):
    from modarith_b200 import Field
    from modarith_b200.lib import MabError, ERR_NOJIT
    try:
        cub = Field.modprog_cubin("NIST384", [("mul", 2, 0, 1), ("sqr", 2, 2, 0), ("add", 3, 2, 0)], 2, [2, 3])
    except MabError as e:
        if "code %d" % ERR_NOJIT in str(e):
            pytest.skip("NVRTC is not installed on this machine")
        raise
    assert cub[:4] == b"\x7fELF" and b"k_prog_jit" in cub


def test_names_and_missing_libraries_fail_with_instructions():
    from modarith_b200 import build, lib as mlib
    with pytest.raises(ValueError):
        build.build_extra("no-such name")
    with pytest.raises(KeyError):
        build.build_extra("NOSUCHPRIME")
    with pytest.raises(mlib.MabError, match="--prime GM240"):
        mlib.load_for("GM240")


def test_every_named_modulus_of_the_reference_has_a_plan():
    """pseudo.py:1487-1550 / monty.py:1961-2108: every name either generator knows is a prime the sm_100a generator
    accepts (the fall-back plan makes it total, as monty.py is for the reference)."""
    from modarith_b200 import primes
    from modarith_b200.gen.plan import make_plan
    assert len(primes.REFERENCE_PRIMES) >= 30
    for name in primes.REFERENCE_PRIMES:
        P = primes.named(name)
        assert pow(3, P.p - 1, P.p) == 1, name                 # the reference's own sanity check, pseudo.py:1561-1564
        plan = make_plan(P)
        assert plan.L == (P.nbits + 31) // 32, name


# ---- an add-on library WITH a ladder: a user-defined Montgomery curve ------------------------------------------------
M383_P = 2**383 - 187
M383_CURVE = (516287, 3, 12)          # a24 = (A - 2) / 4 for A = 2065150, cofactor 2^3, base point u = 12


@pytest.fixture(scope="module")
def m383():
    from modarith_b200 import build
    return build.build_extra("M383", "2**383-187", verbose=False, curve=M383_CURVE)


def test_curve_addon_exports_the_ladder_abi(m383):
    from modarith_b200 import lib as mlib
    dll = ctypes.CDLL(m383)
    for n in ("rfc7748", "rfc7748_perkey", "rfc7748_validate", "rfc7748_host", "rfc7748_host_multi", "modmul", "modprog_jit"):
        assert hasattr(dll, "mab_M383_" + n), n
    par = mlib.params("M383")
    assert (par["nlimbs"], par["nbits"], par["nbytes"], par["has_curve"], par["pm1d2"]) == (12, 383, 48, 1, 2)
    from modarith_b200 import build
    with pytest.raises(ValueError):
        build.build_extra("BADCURVE", "2**383-187", curve=(516287, 5, 12))      # cofactor 2^2 or 2^3 only


def test_oracle_ladder_on_the_user_curve_matches_the_reference_build():
    """Pins the oracle for this curve: the reference's rfc7748.c with the curve's constants added to its "Describe
    Montgomery Curve parameters" section (oracle/build_ref.py), against the restatement, on random raw keys."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from field_oracle import rfc7748 as oracle_rfc7748
    from oracle_primes import OraclePrime
    sys.path.insert(0, os.path.dirname(__file__))
    import util
    path = os.path.join(ROOT, "oracle", "_ref", "libref_M383.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    ref = ctypes.CDLL(path)
    P = OraclePrime("M383", M383_P, a24=M383_CURVE[0], cof=M383_CURVE[1], generator=M383_CURVE[2])
    k, u = util.random_bytes(3831, 24, 48), util.random_bytes(3832, 24, 48)
    u[0] = 0
    u[1] = np.frombuffer((12).to_bytes(48, "little"), dtype=np.uint8)
    u[2] = np.frombuffer((M383_P - 1).to_bytes(48, "little"), dtype=np.uint8)
    k[3] = 255
    want = util.ref_rfc7748_batch(ref, k, u)
    for i in range(24):
        assert oracle_rfc7748(P, k[i].tobytes(), u[i].tobytes()) == want[i].tobytes(), i
    vref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_M383_validate.so"))
    wantv = util.ref_rfc7748_batch(vref, k, u)
    for i in range(24):
        assert oracle_rfc7748(P, k[i].tobytes(), u[i].tobytes(), twist_secure=False) == wantv[i].tobytes(), i
