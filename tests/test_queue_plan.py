"""Host logic of the round-structured ladder's work queues (modarith_b200/csrc/mab_queue_plan.h), compiled with g++:
every plan the host can cut covers every group of a queue exactly once, in chunks the kernel's result slots can hold.
(The kernel walks the same functions; what it computes with them is checked by the -m gpu parity tests.)"""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

HARNESS = r'''
#include "mab_queue_plan.h"
extern "C" {
void t_plan(unsigned g, unsigned w, int kmax, int tail, unsigned* c4, unsigned* c2) {
  MabQueuePlan p = mab_queue_plan(g, w, kmax, tail);
  *c4 = p.c4; *c2 = p.c2;
}
unsigned t_nchunks(unsigned g, unsigned c4, unsigned c2) { return mab_queue_nchunks(g, c4, c2); }
int t_chunk(unsigned g, unsigned c4, unsigned c2, unsigned long long ci, unsigned* first) {
  unsigned f = 0xffffffffu;
  int k = mab_queue_chunk(g, c4, c2, ci, f);
  *first = f;
  return k;
}
}
'''


@pytest.fixture(scope="module")
def qp(tmp_path_factory):
    d = tmp_path_factory.mktemp("queue_plan")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    out = str(d / "libqp.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "modarith_b200", "csrc"), "-o", out, str(src)])
    lib = ctypes.CDLL(out)
    lib.t_chunk.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_uint)]
    return lib


def plan(lib, g, w, kmax, tail):
    c4, c2 = ctypes.c_uint(), ctypes.c_uint()
    lib.t_plan(g, w, kmax, tail, ctypes.byref(c4), ctypes.byref(c2))
    return c4.value, c2.value


def walk(lib, g, c4, c2):
    """every chunk the kernel would draw from the queue's counter, in order: [(K, first group)]"""
    out = []
    ci = 0
    while True:
        first = ctypes.c_uint()
        k = lib.t_chunk(g, c4, c2, ci, ctypes.byref(first))
        if k == 0:
            return out
        out.append((k, first.value))
        ci += 1
        assert ci <= g + 1


@pytest.mark.parametrize("kmax", [4, 2])
@pytest.mark.parametrize("tail", [0, 1, 2])
def test_every_plan_covers_the_queue_exactly_once(qp, kmax, tail):
    for w in (1, 2, 3, 4, 8, 16):
        for g in list(range(0, 260)) + [1000, 4095, 4096, 65537]:
            c4, c2 = plan(qp, g, w, kmax, tail)
            assert 4 * c4 + 2 * c2 <= g
            assert c4 % w == 0 and c2 % w == 0          # whole rounds of big chunks only
            if kmax < 4:
                assert c4 == 0
            chunks = walk(qp, g, c4, c2)
            assert len(chunks) == qp.t_nchunks(g, c4, c2) == c4 + c2 + (g - 4 * c4 - 2 * c2)
            nxt = 0
            for k, first in chunks:
                assert k <= kmax and first == nxt        # contiguous, disjoint, never more slots than the kernel has
                nxt += k
            assert nxt == g
            ks = [k for k, _ in chunks]
            assert ks == sorted(ks, reverse=True)        # big chunks first, single groups last
            # past the end the walk stays at the end (a warp may draw a counter value beyond the list)
            first = ctypes.c_uint()
            assert qp.t_chunk(g, c4, c2, len(chunks) + 5, ctypes.byref(first)) == 0


def test_reserved_round_of_single_groups(qp):
    """tail = 0: at least one round's worth (w) of single groups whenever the queue is longer than w;
    tail = 1 (default) gives that reserve up only where no K = 4 chunk is cut."""
    for w in (2, 3, 4):
        for g in range(w + 1, 200):
            c4, c2 = plan(qp, g, w, 4, 0)
            assert g - 4 * c4 - 2 * c2 >= w
            d4, d2 = plan(qp, g, w, 4, 1)
            if d4:
                assert (d4, d2) == (c4, c2)
            else:
                assert g - 2 * d2 < 2 * w


def test_documented_cases(qp):
    # 2^17 X25519 keys on 148 SMs x 4 queues of 3 warps: 6.9 groups per queue
    assert plan(qp, 7, 3, 4, 0) == (0, 0)
    assert plan(qp, 7, 3, 4, 1) == (0, 3)
    # 2^20 keys: 55.4 groups per queue: twelve K = 4 chunks (four per warp), seven single groups
    assert plan(qp, 55, 3, 4, 1) == (12, 0)
    assert plan(qp, 55, 3, 4, 2) == (12, 3)
