"""mab_<curve>_rfc7748_host_multi: the whole box from one C call (replaces a consumer's loop over keys,
rfc7748.c:301-304).  Contiguous key ranges, one host thread per device.  Needs >= 2 GPUs for the real thing; the
argument handling and the one-device route are tested on any GPU box."""
import ctypes

import numpy as np
import pytest
import torch

from modarith_b200 import lib as mlib
from modarith_b200.rfc7748 import rfc7748
from modarith_b200.shard import key_range
import util

pytestmark = pytest.mark.gpu


def _single(curve, k, u):
    return rfc7748(curve, torch.from_numpy(k).cuda(), torch.from_numpy(u).cuda()).cpu().numpy()


@pytest.mark.parametrize("curve,nb", [("X25519", 32), ("X448", 56)])
def test_multi_entry_on_whatever_is_there(curve, nb):
    l = mlib.load()
    have = l.mab_device_count()
    fn = getattr(l, "mab_%s_rfc7748_host_multi" % curve)
    n = 5000 + 7
    k, u = util.random_bytes(31, n, nb), util.random_bytes(32, n, nb)
    want = _single(curve, k, u)
    out = np.zeros_like(k)
    for ndev in sorted({0, 1, have}):
        out[:] = 0
        mlib.check(fn(k.ctypes.data, u.ctypes.data, out.ctypes.data, n, ndev))
        assert np.array_equal(out, want), ndev
    assert fn(k.ctypes.data, u.ctypes.data, out.ctypes.data, n, have + 1) == 100001      # MAB_ERR_BADARG
    assert fn(k.ctypes.data, u.ctypes.data, out.ctypes.data, 0, 0) == 0


@pytest.mark.parametrize("curve,nb", [("X25519", 32), ("X448", 56)])
def test_two_or_more_devices_against_one(curve, nb):
    l = mlib.load()
    have = l.mab_device_count()
    if have < 2:
        pytest.skip("needs at least two GPUs")
    fn = getattr(l, "mab_%s_rfc7748_host_multi" % curve)
    for n in (1, 3, 100003, (1 << 18) + 17):
        k, u = util.random_bytes(41 + n % 7, n, nb), util.random_bytes(42 + n % 7, n, nb)
        u[::101] = 0                                                  # low-order inputs scattered over the ranges
        want = _single(curve, k, u)
        for ndev in range(2, have + 1):
            # pageable buffers: staged per device
            out = np.zeros_like(k)
            mlib.check(fn(k.ctypes.data, u.ctypes.data, out.ctypes.data, n, ndev))
            assert np.array_equal(out, want), (n, ndev, "pageable")
            # page-locked buffers: every GPU reads its range of the same allocation in place
            hk, hu = torch.from_numpy(k).pin_memory(), torch.from_numpy(u).pin_memory()
            hv = torch.zeros((n, nb), dtype=torch.uint8).pin_memory()
            mlib.check(fn(hk.data_ptr(), hu.data_ptr(), hv.data_ptr(), n, ndev))
            assert np.array_equal(hv.numpy(), want), (n, ndev, "pinned")
            # the ranges are those of shard.key_range (what the torchrun form uses)
            lo, hi = key_range(ndev - 1, ndev, n)
            assert np.array_equal(hv.numpy()[lo:hi], want[lo:hi])
    # Python wrapper
    k, u = util.random_bytes(51, 70001, nb), util.random_bytes(52, 70001, nb)
    assert np.array_equal(rfc7748(curve, k, u, device="all"), _single(curve, k, u))
    assert np.array_equal(rfc7748(curve, k, u, device=("all", 2)), _single(curve, k, u))
