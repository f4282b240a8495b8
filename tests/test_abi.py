"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/modarith_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from modarith_b200 import lib as mlib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def built():
    from modarith_b200 import build
    return build.build(verbose=False)


def _macro_body(h, name):
    """Text of a multi-line #define (lines joined by backslashes)."""
    m = re.search(r"#define %s\(P\)(.*?[^\\])\n" % name, h, re.S)
    return m.group(1)


def header_symbols():
    h = open(os.path.join(ROOT, "include", "modarith_b200.h")).read()
    field = re.findall(r"mab_##P##_([a-z0-9_]+)\s*\(", _macro_body(h, "MAB_DECLARE_FIELD"))
    curve = re.findall(r"mab_##P##_([a-z0-9_]+)\s*\(", _macro_body(h, "MAB_DECLARE_CURVE"))
    assert len(field) == 36 and sorted(curve) == ["rfc7748", "rfc7748_host", "rfc7748_host_multi", "rfc7748_perkey",
                                                  "rfc7748_validate"]
    syms = set(re.findall(r"\b(mab_[A-Za-z0-9_]+)\s*\(", h))
    for P in re.findall(r"MAB_DECLARE_FIELD\((\w+)\)\n", h):
        if P != "P":
            syms |= {"mab_%s_%s" % (P, m) for m in field}
    # the built-in curves are declared one by one (with their comments); the macro declares the same set for add-ons
    for P in ("X25519", "X448"):
        assert {"mab_%s_%s" % (P, m) for m in curve} <= syms
    return syms


def test_library_exports_every_declared_symbol(built):
    dll = ctypes.CDLL(built)
    want = header_symbols()
    assert len(want) == 14 + 5 * 36 + 10
    for s in sorted(want):
        assert hasattr(dll, s), s
    assert want == set(mlib.exported_symbols())


def test_loader_binds_and_reports(built):
    l = mlib.load()
    assert b"sm_100a" in l.mab_version()
    q = mlib.params("X25519")
    assert (q["wordlength"], q["nlimbs"], q["radix"], q["nbits"], q["nbytes"]) == (32, 8, 32, 255, 32)
    assert mlib.params("SECP256K1")["montgomery"] == 0 and mlib.params("NIST256ORDER")["montgomery"] == 1
    assert mlib.params("NIST256ORDER")["pm1d2"] == 4
    assert mlib.params("X448")["nlimbs"] == 14 and mlib.params("NIST256")["nbytes"] == 32
    # SURVEY.md 8d work counts with this build's chains (251S+13M, 445S+14M)
    assert mlib.products("X25519", "modmul") == 64 and mlib.products("X25519", "modsqr") == 36
    assert mlib.products("X448", "modmul") == 196 and mlib.products("X448", "modsqr") == 105
    assert mlib.products("X25519", "rfc7748") == 255 * (5 * 64 + 4 * 36 + 8) + 251 * 36 + 13 * 64 + (36 + 64 + 3 * 36 + 64) + 64
    assert mlib.products("NIST256", "rfc7748") == -1
    assert l.mab_error_string(mlib.load().mab_params(b"nope", None, None, None, None, None)) == b"modarith_b200: bad argument"


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from modarith_b200 import Field, x25519
    import numpy as np
    with pytest.raises(mlib.MabError):
        Field("X25519")
    with pytest.raises(mlib.MabError):
        x25519(np.zeros((1, 32), np.uint8), np.zeros((1, 32), np.uint8))


def test_product_package_never_imports_the_oracle():
    for dp, _, fns in os.walk(os.path.join(ROOT, "modarith_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".inc")):
                text = open(os.path.join(dp, fn)).read()
                assert "field_oracle" not in text and "oracle/" not in text.replace("oracle/_ref", "").replace(
                    "oracle/build_ref", "").replace("oracle/addchain", ""), os.path.join(dp, fn)


def test_header_is_plain_c_and_links(built, tmp_path):
    """include/modarith_b200.h compiled as C11 by gcc, linked against the shared library, run without a GPU:
    the boundary is a C ABI (no C++ or torch types), and a call that needs no device works."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "consumer.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "modarith_b200.h"\n'
        "int main(void) {\n"
        "  int (*ladder)(const char *, const char *, char *, size_t, void *) = mab_X25519_rfc7748;\n"
        "  int (*host)(const char *, const char *, char *, size_t, int) = mab_X448_rfc7748_host;\n"
        "  int (*mul2)(const char *, const char *, const char *, const char *, const char *, const char *, char *, char *, size_t, void *) = mab_NIST256_ecnmul2;\n"
        "  if (!ladder || !host || !mul2) return 2;\n"
        '  printf("%s|%s\\n", mab_version(), mab_error_string(0));\n'
        "  return ladder(0, 0, 0, 0, 0);                 /* n = 0: nothing to do, no device needed */\n"
        "}\n")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(built)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", libdir, "-lmodarith_b200", "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.startswith("modarith_b200") and out.stdout.strip().endswith("|ok")
