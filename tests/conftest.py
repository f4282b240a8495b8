import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a box without a GPU: the gpu-marked tests are skipped, not failed (on a GPU box they run;
    the product itself never falls back -- see tests/test_abi.py)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_rfc():
    with open(os.path.join(ROOT, "tests", "golden", "rfc7748.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_ecn():
    with open(os.path.join(ROOT, "tests", "golden", "ecn.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_ecn2():
    with open(os.path.join(ROOT, "tests", "golden", "ecn2.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_testcurve():
    with open(os.path.join(ROOT, "tests", "golden", "testcurve.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_field():
    with open(os.path.join(ROOT, "tests", "golden", "field.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def hostsim(tmp_path_factory):
    """g++ build of the generated headers + device templates with the PTX blocks transcribed to C
    (tests/hostsim/hostsim.cpp).  Test scaffolding: lets the CPU suite run the device logic."""
    from modarith_b200.gen.cli import generate_all
    d = tmp_path_factory.mktemp("hostsim")
    generate_all(verbose=False, outdir=str(d))           # fresh headers in a scratch directory, not in the source tree
    out = str(d / "libhostsim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I", str(d),
                           "-I", os.path.join(ROOT, "modarith_b200", "csrc"),
                           "-o", out, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    return ctypes.CDLL(out)


def _ref(name):
    p = os.path.join(ROOT, "oracle", "_ref", "libref_%s.so" % name)
    if not os.path.exists(p):
        return None
    return ctypes.CDLL(p)


@pytest.fixture(scope="session")
def ref_libs():
    """The reference's own generated C (oracle/_ref), when it has been built."""
    libs = {n: _ref(n) for n in ("X25519", "X448", "NIST256", "X25519_generic", "X448_generic",
                                  "X25519_validate", "X448_validate", "SECP256K1", "NIST256ORDER", "NIST256_curve", "ED25519_curve")}
    return {k: v for k, v in libs.items() if v is not None}
