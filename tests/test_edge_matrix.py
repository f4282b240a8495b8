"""The reference's corner-case run, edge.py + edge.c, replayed for every built-in modulus.

edge.py:341-360 lists 17 operand pairs (a, b) in terms of p, n = bitlen(p), c = 2^n - p, a random r and
its inverse; edge.py:116-163 (`corner`) writes, for each pair, 12 expected values computed with Python's `%`:
a, a+b, a-b, b-a, a*b, a*a and the same six on the DOUBLED operands 2a, 2b.  edge.c:76-245 reads them back
and checks, through modcmp, the generated functions: modinv(modinv(a)), modadd, modsub both ways, modmul,
modsqr(modsqrt(modsqr(a))), then doubles x and y IN PLACE with modadd(x,x,x) / modadd(y,y,y) ("doubling will
not trigger reduction" -- the second stage feeds unreduced sums into inv/add/sub/mul/sqr) and repeats.

The expected values are regenerated here from the same formulas (nothing is copied from an edge.txt); the
flow is edge.c's, over all 17 pairs at once as one batch.  Three engines run it: the value-level oracle, the
host simulation of the generated device code (CPU suite), and the CUDA library through its C ABI (-m gpu).
"""
import random

import pytest

from field_oracle import FieldOracle
from modarith_b200.primes import ALL_PRIMES as PRIMES
import util

NAMES = list(PRIMES)


def edge_pairs(name):
    """edge.py:341-360 (the order is the reference's); r is random there, seeded here."""
    P = PRIMES[name]
    p, n = P.p, P.nbits
    c = (1 << n) - p
    r = random.Random(0xED6E + n).randrange(p)
    i = pow(r, -1, p)
    return [(p - 1, p - 1), (0, p - 1), (0, 0), (r, r), (r, i), (p, 1), (p - 1, 1), (p - 2, 2), (p - 1, r), (p - 2, r),
            (2 ** 64, 2 ** 64), (2 ** (n - 1), 2 ** (n - 1) - 1), (c, 1), (c, 2 ** n - 1), (2 ** n - 1, 0),
            (2 ** n - 1, 1), (2 ** n - 1, 2 ** n - 1)]


def corner(p, a, b):
    """The 12 values of edge.py:116-163 after the two echoed operands."""
    return [a % p, (a + b) % p, (a - b) % p, (b - a) % p, (a * b) % p, (a * a) % p,
            (2 * a) % p, (2 * a + 2 * b) % p, (2 * a - 2 * b) % p, (2 * b - 2 * a) % p, (2 * a * 2 * b) % p,
            (2 * a * 2 * a) % p]


def run_edge_c(E, name):
    """edge.c:76-245 on engine E (methods imp/add/sub/mul/sqr/inv/sqrt/cmp over a batch of 17 elements).
    Returns the list of (pair index, test index) that failed."""
    p = PRIMES[name].p
    pairs = edge_pairs(name)
    want = [corner(p, a, b) for a, b in pairs]
    x = E.imp([a for a, _ in pairs])
    y = E.imp([b for _, b in pairs])
    failed = []

    def check(t, z):
        w = E.imp([want[i][t] for i in range(len(pairs))])
        ok = E.cmp(w, z)
        failed.extend((i, t) for i, v in enumerate(ok) if v != 1)

    def stage(t0):
        check(t0 + 0, E.inv(E.inv(x)))
        check(t0 + 1, E.add(x, y))
        check(t0 + 2, E.sub(x, y))
        check(t0 + 3, E.sub(y, x))
        check(t0 + 4, E.mul(x, y))
        check(t0 + 5, E.sqr(E.sqrt(E.sqr(x))))

    stage(0)
    x = E.add(x, x)            # edge.c:161  modadd(x,x,x); modadd(y,y,y)
    y = E.add(y, y)
    stage(6)
    return failed


class OracleEngine:
    def __init__(self, name):
        self.O = FieldOracle(name)

    def imp(self, vals):
        return [self.O.modimp(v.to_bytes(self.O.nbytes, "big"))[0] for v in vals]

    def add(self, a, b): return [self.O.modadd(u, v) for u, v in zip(a, b)]
    def sub(self, a, b): return [self.O.modsub(u, v) for u, v in zip(a, b)]
    def mul(self, a, b): return [self.O.modmul(u, v) for u, v in zip(a, b)]
    def sqr(self, a): return [self.O.modsqr(u) for u in a]
    def inv(self, a): return [self.O.modinv(u) for u in a]
    def sqrt(self, a): return [self.O.modsqrt(u) for u in a]
    def cmp(self, a, b): return [self.O.modcmp(u, v) for u, v in zip(a, b)]


class SimEngine:
    """tests/hostsim: the generated field code with the PTX transcribed to C; stored forms as integers."""

    def __init__(self, lib, name):
        self.S = util.Sim(lib, name)

    def imp(self, vals): return [self.S.imp(v)[0] for v in vals]
    def add(self, a, b): return [self.S.raw("ADD", u, v)[0] for u, v in zip(a, b)]
    def sub(self, a, b): return [self.S.raw("SUB", u, v)[0] for u, v in zip(a, b)]
    def mul(self, a, b): return [self.S.raw("MUL", u, v)[0] for u, v in zip(a, b)]
    def sqr(self, a): return [self.S.raw("SQR", u)[0] for u in a]
    def inv(self, a): return [self.S.raw("INV", u)[0] for u in a]
    def sqrt(self, a): return [self.S.raw("SQRT", u)[0] for u in a]
    def cmp(self, a, b): return [self.S.raw("CMP", u, v)[1] for u, v in zip(a, b)]


class GpuEngine:
    """modarith_b200.Field = one mab_<PRIME>_<function> call of the C ABI per step, limb planes on the GPU."""

    def __init__(self, name):
        from modarith_b200 import Field
        self.F = Field(name)

    def _new(self, like): return self.F.alloc(like.shape[1])
    def imp(self, vals): return self.F.from_ints(vals)

    def add(self, a, b):
        r = self._new(a); self.F.modadd(a, b, r); return r

    def sub(self, a, b):
        r = self._new(a); self.F.modsub(a, b, r); return r

    def mul(self, a, b):
        r = self._new(a); self.F.modmul(a, b, r); return r

    def sqr(self, a):
        r = self._new(a); self.F.modsqr(a, r); return r

    def inv(self, a):
        r = self._new(a); self.F.modinv(a, None, r); return r

    def sqrt(self, a):
        r = self._new(a); self.F.modsqrt(a, None, r); return r

    def cmp(self, a, b): return self.F.modcmp(a, b).cpu().tolist()


def test_pairs_are_the_reference_list():
    for name in NAMES:
        pairs = edge_pairs(name)
        assert len(pairs) == 17
        n = PRIMES[name].nbits
        assert all(0 <= v < (1 << n) for ab in pairs for v in ab)      # edge.py:339: "positive and less than 2^n"


@pytest.mark.parametrize("name", NAMES)
def test_edge_matrix_oracle(name):
    assert run_edge_c(OracleEngine(name), name) == []


@pytest.mark.parametrize("name", NAMES)
def test_edge_matrix_hostsim(hostsim, name):
    assert run_edge_c(SimEngine(hostsim, name), name) == []


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_edge_matrix_gpu(name):
    """204 checks per modulus (17 pairs x 12 values) through the C ABI, doubled stage included."""
    assert run_edge_c(GpuEngine(name), name) == []


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_edge_matrix_gpu_in_place(name):
    """edge.c doubles its operands with the output aliasing both inputs (modadd(x,x,x), edge.c:161) and
    inverts in place (modinv(z,NULL,z), edge.c:95): the same run with those aliasings kept."""
    E = GpuEngine(name)
    F = E.F
    p = PRIMES[name].p
    pairs = edge_pairs(name)
    x, y = E.imp([a for a, _ in pairs]), E.imp([b for _, b in pairs])
    F.modadd(x, x, x)
    F.modadd(y, y, y)
    z = F.alloc(x.shape[1])
    F.modinv(x, None, z)
    F.modinv(z, None, z)
    assert F.to_ints(z) == [2 * a % p for a, _ in pairs]
    F.modsqr(x, z); F.modsqrt(z, None, z); F.modsqr(z, z)
    assert F.to_ints(z) == [4 * a * a % p for a, _ in pairs]
    F.modmul(x, y, x)
    assert F.to_ints(x) == [4 * a * b % p for a, b in pairs]
