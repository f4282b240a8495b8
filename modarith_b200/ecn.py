"""Batched group-law entry points: the reference's `ecnXXXset` + `ecnXXXmul` + `ecnXXXget`
(weierstrass.c:415-427,494-542,333-349; edwards.c likewise) for n independent points, and the double
multiplication `ecnXXXmul2` (weierstrass.c:545-572 / edwards.c:486-513).

    xo, yo = ecnmul("NIST256", e, x, y)      # [n, 32] uint8, big-endian like the reference's char* arguments
    xo, yo = ecnmul("ED25519", e, x, y)      # twisted Edwards (edwards.c), identity reported as (0, 1)
    xo, yo = ecnmul2("NIST256", e, x1, y1, f, x2, y2)   # e*(x1,y1) + f*(x2,y2): the verification block

CUDA tensors are used in place on the current stream and CUDA tensors come back; numpy arrays (or CPU
tensors) are copied to the current device and numpy arrays come back.  A point off the curve, a zero scalar
or a multiple of the group order give (0, 1) (ecnXXXget of O).  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as _lib

CURVES = ("NIST256", "ED25519")
NBYTES = 32


def _args(curve, arrays):
    if curve not in CURVES:
        raise ValueError("unsupported curve %r (have %s)" % (curve, ", ".join(CURVES)))
    on_device = all(isinstance(t, torch.Tensor) and t.is_cuda for t in arrays)
    if not on_device:
        if any(isinstance(t, torch.Tensor) and t.is_cuda for t in arrays):
            raise ValueError("mixed host and device arguments")
        if not torch.cuda.is_available():
            raise _lib.MabError("modarith_b200 needs a CUDA device: there is no CPU fallback")
        arrays = [torch.from_numpy(np.ascontiguousarray(np.asarray(t), dtype=np.uint8)).cuda() for t in arrays]
    e = arrays[0]
    for t in arrays:
        if t.dtype != torch.uint8 or t.dim() != 2 or t.shape[1] != NBYTES or t.shape != e.shape or not t.is_contiguous():
            raise ValueError("every argument must be a contiguous [n, %d] uint8 array of the same n" % NBYTES)
        if t.device != e.device:
            raise ValueError("arguments live on different devices")
    return arrays, on_device


def _call(name, arrays, on_device, xo, yo):
    lib = _lib.load()
    e = arrays[0]
    if not on_device and (xo is not None or yo is not None):
        raise ValueError("xo / yo can only be supplied with device tensors (host arguments get fresh numpy results)")
    for nm, t in (("xo", xo), ("yo", yo)):
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.uint8 or t.shape != e.shape \
                or t.device != e.device or not t.is_contiguous():
            raise ValueError("%s must be a contiguous uint8 CUDA tensor of shape %s on %s" % (nm, tuple(e.shape), e.device))
    xo = torch.empty_like(e) if xo is None else xo
    yo = torch.empty_like(e) if yo is None else yo
    stream = torch.cuda.current_stream(e.device).cuda_stream
    with torch.cuda.device(e.device):
        _lib.check(getattr(lib, name)(*(t.data_ptr() for t in arrays), xo.data_ptr(), yo.data_ptr(), e.shape[0], stream), name)
    if on_device:
        return xo, yo
    return xo.cpu().numpy(), yo.cpu().numpy()


def ecnmul(curve: str, e, x, y, xo=None, yo=None):
    """(xo, yo) = e * (x, y) for every row."""
    arrays, on_device = _args(curve, [e, x, y])
    return _call("mab_%s_ecnmul" % curve, arrays, on_device, xo, yo)


def ecnmul2(curve: str, e, x1, y1, f, x2, y2, xo=None, yo=None):
    """(xo, yo) = e * (x1, y1) + f * (x2, y2) for every row."""
    arrays, on_device = _args(curve, [e, x1, y1, f, x2, y2])
    return _call("mab_%s_ecnmul2" % curve, arrays, on_device, xo, yo)
