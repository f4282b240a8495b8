"""Argument checking of the host-side wrappers (no GPU needed): what the C ABI cannot take is refused with
ValueError / TypeError -- never with `assert`, which `python -O` strips -- and a host OUTPUT array that could
not be written in place is refused instead of being silently copied."""
import subprocess
import sys

import numpy as np
import pytest
import torch

from modarith_b200 import rfc7748 as R
from modarith_b200 import ecn, lib as mlib


def test_rfc7748_host_arguments():
    k = np.zeros((4, 32), dtype=np.uint8)
    with pytest.raises(ValueError):
        R.rfc7748("X25518", k, k)
    with pytest.raises(ValueError):
        R.rfc7748("X25519", k, np.zeros((5, 32), dtype=np.uint8))                 # shapes differ
    with pytest.raises(ValueError):
        R.rfc7748("X25519", np.zeros((4, 56), dtype=np.uint8), np.zeros((4, 56), dtype=np.uint8))
    with pytest.raises(TypeError):
        R.rfc7748("X25519", k.astype(np.int32), k)
    with pytest.raises(ValueError):                                                # strided output: would be a copy
        R.rfc7748("X25519", k, k, np.zeros((4, 64), dtype=np.uint8)[:, ::2])
    with pytest.raises(ValueError):
        R.rfc7748("X25519", k, k, np.zeros((4, 32), dtype=np.int32))
    with pytest.raises(ValueError):
        R.rfc7748("X25519", k, k, np.zeros((3, 32), dtype=np.uint8))
    with pytest.raises(ValueError):
        R.rfc7748("X25519", k, k, device="some")
    with pytest.raises(ValueError):
        R.rfc7748("X25519", k, k, validate=True)
    with pytest.raises(ValueError):
        R.rfc7748("X25519", torch.zeros((4, 64), dtype=torch.uint8)[:, ::2], torch.zeros((4, 32), dtype=torch.uint8))
    if not torch.cuda.is_available():
        with pytest.raises(mlib.MabError):                                         # valid arguments, no device: loud
            R.rfc7748("X25519", k, k)


def test_field_argument_checks_without_device():
    from modarith_b200.field import Field
    F = Field.__new__(Field)                      # the constructor needs a GPU; the checks do not
    F.Nlimbs, F.Nbytes, F.device = 8, 32, torch.device("cuda", 0)
    with pytest.raises(TypeError):
        F._chk(torch.zeros((8, 4), dtype=torch.int32))                             # CPU tensor
    with pytest.raises(TypeError):
        F._chk(np.zeros((8, 4), dtype=np.int32))
    with pytest.raises(TypeError):
        F._bytes(torch.zeros((4, 32), dtype=torch.uint8), 4, "b")
    with pytest.raises(ValueError):
        F._bits(torch.zeros(4, dtype=torch.int32), torch.zeros((8, 4), dtype=torch.int32))


def test_ecn_arguments():
    e = np.zeros((2, 32), dtype=np.uint8)
    with pytest.raises(ValueError):
        ecn.ecnmul("SECP256K1", e, e, e)
    if not torch.cuda.is_available():
        with pytest.raises(mlib.MabError):
            ecn.ecnmul("NIST256", e, e, e)


def test_checks_survive_python_O():
    """python -O removes assert statements; the wrappers must still refuse bad arguments."""
    code = ("import numpy as np\n"
            "from modarith_b200 import rfc7748 as R\n"
            "k = np.zeros((4, 32), dtype=np.uint8)\n"
            "try:\n"
            "    R.rfc7748('X25519', k, np.zeros((5, 32), dtype=np.uint8))\n"
            "except ValueError:\n"
            "    print('refused')\n")
    out = subprocess.run([sys.executable, "-O", "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         cwd=str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    assert out.stdout.strip() == "refused", out.stderr[-500:]
