"""Batched short-Weierstrass scalar multiplication: the reference's
`ecnXXXset` + `ecnXXXmul` + `ecnXXXget` (weierstrass.c:415-427,494-542,333-349) for n independent points.

    xo, yo = ecnmul("NIST256", e, x, y)      # [n, 32] uint8 cuda tensors, big-endian like the reference's char*
    xo, yo = ecnmul("ED25519", e, x, y)      # twisted Edwards (edwards.c), identity reported as (0, 1)

    xo, yo = ecnmul2("NIST256", e, x1, y1, f, x2, y2)   # e*(x1,y1) + f*(x2,y2): ecnXXXmul2, the verification block

A point off the curve, a zero scalar or a multiple of the group order give (0, 1) (ecnXXXget of O).
"""
from __future__ import annotations

import torch

from . import lib as _lib


def ecnmul(curve: str, e, x, y, xo=None, yo=None):
    if curve not in ("NIST256", "ED25519"):
        raise ValueError("unsupported curve %r (have NIST256, ED25519)" % curve)
    lib = _lib.load()
    for t in (e, x, y):
        assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 2
        assert t.shape == e.shape and t.shape[1] == 32 and t.is_contiguous()
    n = e.shape[0]
    xo = torch.empty_like(x) if xo is None else xo
    yo = torch.empty_like(y) if yo is None else yo
    stream = torch.cuda.current_stream(e.device).cuda_stream
    with torch.cuda.device(e.device):
        fn = getattr(lib, "mab_%s_ecnmul" % curve)
        _lib.check(fn(e.data_ptr(), x.data_ptr(), y.data_ptr(), xo.data_ptr(), yo.data_ptr(), n, stream),
                   "mab_%s_ecnmul" % curve)
    return xo, yo


def ecnmul2(curve: str, e, x1, y1, f, x2, y2, xo=None, yo=None):
    """ecnXXXset x2 + ecnXXXmul2 + ecnXXXget (weierstrass.c:545-572 / edwards.c:486-513) for n independent pairs."""
    if curve not in ("NIST256", "ED25519"):
        raise ValueError("unsupported curve %r (have NIST256, ED25519)" % curve)
    lib = _lib.load()
    for t in (e, x1, y1, f, x2, y2):
        assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 2
        assert t.shape == e.shape and t.shape[1] == 32 and t.is_contiguous()
    n = e.shape[0]
    xo = torch.empty_like(x1) if xo is None else xo
    yo = torch.empty_like(y1) if yo is None else yo
    stream = torch.cuda.current_stream(e.device).cuda_stream
    with torch.cuda.device(e.device):
        fn = getattr(lib, "mab_%s_ecnmul2" % curve)
        _lib.check(fn(e.data_ptr(), x1.data_ptr(), y1.data_ptr(), f.data_ptr(), x2.data_ptr(), y2.data_ptr(),
                      xo.data_ptr(), yo.data_ptr(), n, stream), "mab_%s_ecnmul2" % curve)
    return xo, yo
