"""Batched RFC 7748 entry points (rfc7748.c:156 `rfc7748(bk, bu, bv)`), one key per GPU thread.

    bv = x25519(bk, bu)      bk, bu: [n, 32] uint8 (cuda tensors, or numpy / pinned host arrays)
    bv = x448(bk, bu)        [n, 56]

Byte strings are little-endian exactly as in the reference; scalars are clamped and the
u-coordinate masked on the device (rfc7748.c:135-152,171-175), so raw random bytes are legal.
Device tensors go through `mab_<curve>_rfc7748` on the current stream; host arrays through
`mab_<curve>_rfc7748_host` on one device, or `mab_<curve>_rfc7748_host_multi` on all of them
(`device="all"` or a device count): contiguous key ranges, one host thread per GPU, no exchange.

Like the reference no key or point is ever rejected; what IS rejected, with ValueError / TypeError,
is an argument the C ABI cannot take (wrong dtype, shape, device or a non-contiguous array) -- a
strided output would otherwise be written out of bounds or silently not at all.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as _lib

_NBYTES = {"X25519": 32, "X448": 56}


def _check_dev(name, t, nb, like=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor when bk is one" % name)
    if t.dtype != torch.uint8:
        raise TypeError("%s must be uint8, not %s" % (name, t.dtype))
    if t.dim() != 2 or t.shape[1] != nb:
        raise ValueError("%s must have shape [n, %d], not %s" % (name, nb, tuple(t.shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if like is not None and (t.shape != like.shape or t.device != like.device):
        raise ValueError("%s must match bk in shape and device" % name)


def _host_view(name, x, nb, writable=False):
    """(keep-alive object, address, shape) of a host array the C ABI can use IN PLACE.  Inputs may be copied
    into a contiguous uint8 array; an output must already be one, or the result would land in a copy."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            raise TypeError("%s is a CUDA tensor but bk is a host array" % name)
        if x.dtype != torch.uint8:
            raise TypeError("%s must be uint8, not %s" % (name, x.dtype))
        if not x.is_contiguous():
            raise ValueError("%s must be contiguous" % name)
        shape = tuple(x.shape)
        keep, addr = x, x.data_ptr()
    else:
        a = np.asarray(x)
        if writable:
            if a is not x and not isinstance(x, np.ndarray):
                raise TypeError("%s must be a numpy array or a CPU tensor" % name)
            if a.dtype != np.uint8 or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError("%s must be a writeable C-contiguous uint8 array (the result is written in place)" % name)
        else:
            if a.dtype != np.uint8:
                raise TypeError("%s must be uint8, not %s" % (name, a.dtype))
            a = np.ascontiguousarray(a)
        shape = tuple(a.shape)
        keep, addr = a, a.ctypes.data
    if len(shape) != 2 or shape[1] != nb:
        raise ValueError("%s must have shape [n, %d], not %s" % (name, nb, shape))
    return keep, addr, shape


def rfc7748(curve: str, bk, bu, bv=None, device=None, validate=False):
    """validate=True runs the driver as built without TWIST_SECURE (rfc7748.c:228-251): the result is
    all zero when bu is not on the curve (device tensors only).
    Host arrays: device = an index (default: the current device), or "all" / a count >= 2 given as
    ("all", count) to spread the batch over the GPUs of the box."""
    if curve in _NBYTES:
        lib, nb = _lib.load(), _NBYTES[curve]
    else:
        # a user-defined Montgomery curve: an add-on library built with its constants
        # (python -m modarith_b200.build --prime NAME[=<expression>] --a24 N --cof K [--generator G])
        import os
        if not os.path.exists(_lib.extra_lib_path(curve)):
            raise ValueError("unsupported curve %r (built in: %s; a curve of your own after `python -m modarith_b200.build "
                             "--prime %s=<modulus> --a24 <(A-2)/4> --cof <2|3>`)" % (curve, ", ".join(_NBYTES), curve))
        lib = _lib.load_for(curve)
        par = _lib.params(curve)
        if not par["has_curve"]:
            raise ValueError("the add-on library for %r was built without curve constants (--a24 / --cof)" % curve)
        nb = par["nbytes"]
    if isinstance(bk, torch.Tensor) and bk.is_cuda:
        _check_dev("bk", bk, nb)
        _check_dev("bu", bu, nb, bk)
        if bv is None:
            bv = torch.empty_like(bk)
        else:
            _check_dev("bv", bv, nb, bk)
        stream = torch.cuda.current_stream(bk.device).cuda_stream
        with torch.cuda.device(bk.device):
            name = "mab_%s_rfc7748%s" % (curve, "_validate" if validate else "")
            _lib.check(getattr(lib, name)(bk.data_ptr(), bu.data_ptr(), bv.data_ptr(), bk.shape[0], stream), name, lib)
        return bv
    if validate:
        raise ValueError("validate=True is available for device tensors only")
    # host path (arguments are validated before the device is touched)
    k, kp, ks = _host_view("bk", bk, nb)
    u, up, us = _host_view("bu", bu, nb)
    if ks != us:
        raise ValueError("bk and bu must have the same shape, got %s and %s" % (ks, us))
    if bv is None:
        bv = torch.empty(ks, dtype=torch.uint8, pin_memory=True) if isinstance(bk, torch.Tensor) else np.empty(ks, dtype=np.uint8)
    v, vp, vs = _host_view("bv", bv, nb, writable=True)
    if vs != ks:
        raise ValueError("bv must have the shape of bk, got %s" % (vs,))
    multi = None
    if isinstance(device, str):
        if device != "all":
            raise ValueError("device must be an index, 'all' or ('all', count)")
        multi = 0
    elif isinstance(device, tuple):
        if len(device) != 2 or device[0] != "all" or int(device[1]) < 1:
            raise ValueError("device must be an index, 'all' or ('all', count)")
        multi = int(device[1])
    if not torch.cuda.is_available():
        raise _lib.MabError("modarith_b200 needs a CUDA device: there is no CPU fallback")
    if multi is not None:
        name = "mab_%s_rfc7748_host_multi" % curve
        _lib.check(getattr(lib, name)(kp, up, vp, ks[0], multi), name, lib)
    else:
        dev = torch.cuda.current_device() if device is None else int(device)
        name = "mab_%s_rfc7748_host" % curve
        _lib.check(getattr(lib, name)(kp, up, vp, ks[0], dev), name, lib)
    del k, u, v
    return bv


def x25519(bk, bu, bv=None, device=None):
    return rfc7748("X25519", bk, bu, bv, device)


def x448(bk, bu, bv=None, device=None):
    return rfc7748("X448", bk, bu, bv, device)
