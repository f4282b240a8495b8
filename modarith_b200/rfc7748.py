"""Batched RFC 7748 entry points (rfc7748.c:156 `rfc7748(bk, bu, bv)`), one key per GPU thread.

    bv = x25519(bk, bu)      bk, bu: [n, 32] uint8 (cuda tensors, or numpy / pinned host arrays)
    bv = x448(bk, bu)        [n, 56]

Byte strings are little-endian exactly as in the reference; scalars are clamped and the
u-coordinate masked on the device (rfc7748.c:135-152,171-175), so raw random bytes are legal.
Device tensors go through `mab_<curve>_rfc7748` on the current stream; host arrays through
`mab_<curve>_rfc7748_host`, which pipelines H2D / ladder / D2H over three streams.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as _lib

_NBYTES = {"X25519": 32, "X448": 56}


def rfc7748(curve: str, bk, bu, bv=None, device=None, validate=False):
    """validate=True runs the driver as built without TWIST_SECURE (rfc7748.c:228-251): the result is
    all zero when bu is not on the curve (device tensors only)."""
    if curve not in _NBYTES:
        raise ValueError("unsupported curve %r" % curve)
    lib = _lib.load()
    nb = _NBYTES[curve]
    if isinstance(bk, torch.Tensor) and bk.is_cuda:
        assert bu.is_cuda and bk.dtype == torch.uint8 and bu.dtype == torch.uint8
        assert bk.dim() == 2 and bk.shape[1] == nb and bk.shape == bu.shape
        assert bk.is_contiguous() and bu.is_contiguous()
        n = bk.shape[0]
        if bv is None:
            bv = torch.empty_like(bk)
        assert bv.is_cuda and bv.is_contiguous() and bv.shape == bk.shape and bv.dtype == torch.uint8
        stream = torch.cuda.current_stream(bk.device).cuda_stream
        with torch.cuda.device(bk.device):
            name = "mab_%s_rfc7748%s" % (curve, "_validate" if validate else "")
            _lib.check(getattr(lib, name)(bk.data_ptr(), bu.data_ptr(), bv.data_ptr(), n, stream), name)
        return bv
    if validate:
        raise ValueError("validate=True is available for device tensors only")
    # host path
    if not torch.cuda.is_available():
        raise _lib.MabError("modarith_b200 needs a CUDA device: there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)

    def host(x):
        if isinstance(x, torch.Tensor):
            assert x.dtype == torch.uint8 and x.is_contiguous()
            return x, x.data_ptr(), tuple(x.shape)
        x = np.ascontiguousarray(x, dtype=np.uint8)
        return x, x.ctypes.data, x.shape

    k, kp, ks = host(bk)
    u, up, us = host(bu)
    assert ks == us and len(ks) == 2 and ks[1] == nb
    if bv is None:
        bv = torch.empty(ks, dtype=torch.uint8, pin_memory=True) if isinstance(bk, torch.Tensor) else np.empty(ks, dtype=np.uint8)
    v, vp, vs = host(bv)
    assert vs == ks
    _lib.check(getattr(lib, "mab_%s_rfc7748_host" % curve)(kp, up, vp, ks[0], dev), "mab_%s_rfc7748_host" % curve)
    return bv


def x25519(bk, bu, bv=None, device=None):
    return rfc7748("X25519", bk, bu, bv, device)


def x448(bk, bu, bv=None, device=None):
    return rfc7748("X448", bk, bu, bv, device)
