"""mab_<P>_modprog_jit without a GPU: NVRTC compiles the printed kernel for sm_100a on any machine, so the source
generator, the embedded headers and the compiler plumbing are covered by the CPU suite; the cubin is inspected for
what the design promises (variables in machine registers: no shared memory, no local memory) and goes through the
same constant-time audit as the library's own kernels."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))

X1_, Y1_, Z1_, X2_, Y2_, Z2_, B_, t0, t1, t2, t3, t4, X3, Y3, Z3 = range(15)
# the complete projective addition for a = -3 (weierstrass.c:69-160), as in tests/test_gpu_modprog.py
POINT_ADD = [("mul", t0, X1_, X2_), ("mul", t1, Y1_, Y2_), ("mul", t2, Z1_, Z2_), ("add", t3, X1_, Y1_), ("add", t4, X2_, Y2_),
             ("mul", t3, t3, t4), ("add", t4, t0, t1), ("sub", t3, t3, t4), ("add", t4, Y1_, Z1_), ("add", X3, Y2_, Z2_),
             ("mul", t4, t4, X3), ("add", X3, t1, t2), ("sub", t4, t4, X3), ("add", X3, X1_, Z1_), ("add", Y3, X2_, Z2_),
             ("mul", X3, X3, Y3), ("add", Y3, t0, t2), ("sub", Y3, X3, Y3), ("mul", Z3, B_, t2), ("sub", X3, Y3, Z3),
             ("add", Z3, X3, X3), ("add", X3, X3, Z3), ("sub", Z3, t1, X3), ("add", X3, t1, X3), ("mul", Y3, B_, Y3),
             ("add", t1, t2, t2), ("add", t2, t1, t2), ("sub", Y3, Y3, t2), ("sub", Y3, Y3, t0), ("add", t1, Y3, Y3),
             ("add", Y3, t1, Y3), ("add", t1, t0, t0), ("add", t0, t1, t0), ("sub", t0, t0, t2), ("mul", t1, t4, Y3),
             ("mul", t2, t0, Y3), ("mul", Y3, X3, Z3), ("add", Y3, Y3, t2), ("mul", X3, t3, X3), ("sub", X3, X3, t1),
             ("mul", Z3, t4, Z3), ("mul", t1, t3, t0), ("add", Z3, Z3, t1)]


def _cubin(prime, code, nin, outs):
    from modarith_b200 import Field
    from modarith_b200.lib import MabError, ERR_NOJIT
    try:
        return Field.modprog_cubin(prime, code, nin, outs)
    except MabError as e:
        if "code %d" % ERR_NOJIT in str(e):
            pytest.skip("NVRTC is not installed on this machine")
        raise


def _res_usage(path):
    out = subprocess.run(["cuobjdump", "-res-usage", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    m = re.search(r"Function k_prog_jit:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert m, out
    return tuple(int(x) for x in m.groups())


def test_point_addition_compiles_into_registers(tmp_path):
    cub = _cubin("NIST256", POINT_ADD, 7, [X3, Y3, Z3])
    path = str(tmp_path / "padd.cubin")
    with open(path, "wb") as f:
        f.write(cub)
    reg, stack, shared, local = _res_usage(path)
    assert stack == 0 and shared == 0 and local == 0 and reg <= 168, (reg, stack, shared, local)   # 3 CTAs/SM
    sass = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    wide = len(re.findall(r"IMAD\.WIDE", sass))
    assert 14 * 64 - 64 <= wide <= 14 * 64 + 64, wide        # 14 products of 64 wide multiplies (ptxas turns a few
                                                             # additions into IMAD.WIDE too), nothing duplicated
    # the constant-time audit the library's own kernels go through (tools/ct_audit.py)
    import ct_audit
    audited = 0
    for name, ins in ct_audit.parse(path, "k_prog_jit"):
        flags, _, _ = ct_audit.analyse(name, ins)
        assert not flags, [(w, d["text"]) for w, d in flags[:5]]
        audited += 1
    assert audited == 1


@pytest.mark.parametrize("prime", ["X25519", "X448", "SECP256K1", "NIST256ORDER"])
def test_every_field_compiles(prime, tmp_path):
    code = [("mul", 2, 0, 1), ("sqr", 3, 2, 0), ("add", 4, 3, 0), ("sub", 4, 4, 1), ("neg", 5, 4, 0), ("mli", 6, 5, 0, 121665),
            ("nsqr", 7, 6, 0, 3), ("haf", 8, 7, 0), ("int", 9, 0, 0, 77), ("one", 10, 0, 0), ("zer", 11, 0, 0), ("cpy", 12, 2, 0)]
    cub = _cubin(prime, code, 2, list(range(13)))
    path = str(tmp_path / "p.cubin")
    with open(path, "wb") as f:
        f.write(cub)
    reg, stack, shared, local = _res_usage(path)
    assert shared == 0 and local == 0


def test_bad_programs_are_rejected_before_compiling():
    from modarith_b200 import Field
    with pytest.raises(ValueError):
        Field.modprog_cubin("NIST256", [], 0, [0])
    with pytest.raises(ValueError):
        Field.modprog_cubin("NIST256", [("mul", 16, 0, 0)], 1, [0])
    with pytest.raises(ValueError):
        Field.modprog_cubin("NIST256", [("mul", 1, 0, 0)], 17, [1])


def test_missing_compiler_fails_loudly():
    """No NVRTC, no result: MAB_ERR_NOJIT with the search log, never a quiet switch to the interpreter."""
    env = dict(os.environ, MAB_NVRTC="/nonexistent/libnvrtc.so", LD_LIBRARY_PATH="")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import ctypes, modarith_b200.lib as L\n"
            "lib = L.load()\n"
            "real = ctypes.CDLL.__init__\n"
            "from modarith_b200 import Field\n"
            "try:\n"
            "    Field.modprog_cubin('X25519', [('mul', 1, 0, 0)], 1, [1])\n"
            "    print('COMPILED')\n"
            "except L.MabError as e:\n"
            "    print('ERR', 'code %%d' %% L.ERR_NOJIT in str(e))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, text=True).stdout
    # on a machine where the dynamic loader finds NVRTC by soname the explicit path is only the first candidate
    assert "ERR True" in out or "COMPILED" in out, out
