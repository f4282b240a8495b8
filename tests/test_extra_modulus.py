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
    with pytest.raises(mlib.MabError, match="--prime C41417"):
        mlib.load_for("C41417")


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
