"""oracle/field_oracle.py -- CPU restatement of the reference's batch hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline leg of bench.py, never by the product package.

Each function restates, on Python integers, the VALUE semantics of one generated
function of the reference (pseudo.py / monty.py emit the same API for every
modulus) or of the ladder driver rfc7748.c.  A field
element is held as the integer it represents modulo p; the reference's
unsaturated limbs, Montgomery factor R and lazily-reduced "< 2p" representatives
are invisible after redc/modexp, which is the level parity is pinned at
(SURVEY.md section 8c).  Where the reference's result depends on the
representative (modshr on an unreduced value) the docstring says so.

Pinned against: RFC 7748 vectors embedded in rfc7748.c:271,274 and
simd/rfc7748_simt.cu:245,249, the deterministic outputs of rfc7748.c:main, the
time.c checksums (pseudo.py:1862-1866), and oracle/_ref (the reference's own
generated C compiled in this container) on random inputs -- see
tests/test_oracle_pinned.py and tests/golden/.
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_primes import TABLE as PRIMES, OraclePrime as Prime, lookup as _lookup  # noqa: E402  (the oracle's own tables)


class FieldOracle:
    """Value-level model of the 32-function generated API for one modulus."""

    def __init__(self, prime: Prime | str):
        self.P = _lookup(prime)
        self.p = self.P.p
        self.nbytes = self.P.nbytes
        self.k = self.P.pm1d2
        self.pe = self.P.pe
        self.roi = self.P.roi

    # -- conversions ---------------------------------------------------------
    def nres(self, m):
        """pseudo.py:952-962 (copy) / monty.py:1386-1399 (multiply by R^2 mod p): value kept."""
        return m % self.p

    def redc(self, n):
        """pseudo.py:965-976 / monty.py:1402-1416: canonical residue in [0,p)."""
        return n % self.p

    def modfsb(self, n):
        """pseudo.py:272-283: n-=p, add back if negative; returns (value, 1 iff n was < p).
        Defined for 0 <= n < 2p."""
        return (n - self.p if n >= self.p else n), int(n < self.p)

    # -- ring operations -----------------------------------------------------
    def modadd(self, a, b):
        """pseudo.py:286-304."""
        return (a + b) % self.p

    def modsub(self, a, b):
        """pseudo.py:307-326."""
        return (a - b) % self.p

    def modneg(self, b):
        """pseudo.py:329-348."""
        return (-b) % self.p

    def modmul(self, a, b):
        """pseudo.py:616-659 / monty.py:663-872 (the R^-1 of Montgomery form cancels
        against the nres of the operands)."""
        return a * b % self.p

    def modsqr(self, a):
        """pseudo.py:663-702 / monty.py:982-1165."""
        return a * a % self.p

    def modmli(self, a, b: int):
        """pseudo.py:705-728 / monty.py:876-978: plain small integer b >= 0, no R factor."""
        assert b >= 0
        return a * b % self.p

    def modcpy(self, a):
        """pseudo.py:730-743."""
        return a

    def modnsqr(self, a, n: int):
        """pseudo.py:745-755: square n times."""
        for _ in range(n):
            a = a * a % self.p
        return a

    def modpro(self, w):
        """pseudo.py:758-785: progenitor w^PE, PE=(p-1-2^k)/2^(k+1) (pseudo.py:1574-1581).  The reference runs a
        straight-line sqr/mul program read from addchain's output; the chain only fixes the ORDER of the squarings
        and multiplications, never the value, so the oracle uses its own left-to-right square-and-multiply over
        modsqr / modmul (nothing of the product's chain finder is used here)."""
        r = 1
        for bit in bin(self.pe)[2:] if self.pe else "":
            r = self.modsqr(r)
            if bit == "1":
                r = self.modmul(r, w % self.p)
        return r % self.p

    def modinv(self, x, h=None):
        """pseudo.py:788-812 / monty.py:1225-1251.  0 -> 0."""
        t = self.modpro(x) if h is None else h
        s = x
        for _ in range(self.k - 1):          # only when PM1D2 > 1
            s = self.modsqr(s)
            s = self.modmul(s, x)
        t = self.modnsqr(t, self.k + 1)
        return self.modmul(s, t)

    def modis1(self, a):
        """pseudo.py:877-891."""
        return int(a % self.p == 1)

    def modis0(self, a):
        """pseudo.py:894-906."""
        return int(a % self.p == 0)

    def modzer(self):
        """pseudo.py:909-919."""
        return 0

    def modone(self):
        """pseudo.py:922-934."""
        return 1

    def modint(self, x: int):
        """pseudo.py:937-949."""
        return x % self.p

    def modqr(self, h, x):
        """pseudo.py:815-831: note the (h, x) argument order."""
        r = self.modsqr(self.modpro(x)) if h is None else self.modsqr(h)
        r = self.modmul(r, x)
        if self.k > 1:
            r = self.modnsqr(r, self.k - 1)
        return self.modis1(r) | self.modis0(x)

    def modsqrt(self, x, h=None):
        """pseudo.py:834-874 / monty.py:1272-1311: x*x^PE when k=1; constant-time
        Tonelli-Shanks with root of unity ROI (pseudo.py:1616-1630) when k>1.
        Deterministic for non-residues as well."""
        y = self.modpro(x) if h is None else h
        s = self.modmul(y, x)
        if self.k > 1:
            t = self.modmul(s, y)
            z = self.roi
            for kk in range(self.k, 1, -1):
                b = self.modnsqr(t, kk - 2)
                d = 1 - self.modis1(b)
                v = self.modmul(s, z)
                s = v if d else s            # modcmv(d, v, s)
                z = self.modsqr(z)
                v = self.modmul(t, z)
                t = v if d else t
        return s

    def modcsw(self, b, g, f):
        """pseudo.py:979-1014."""
        return (f, g) if b else (g, f)

    def modcmv(self, b, g, f):
        """pseudo.py:1017-1048: f <- g iff b."""
        return g if b else f

    def modshl(self, n, a):
        """pseudo.py:1052-1065: raw left shift of the limb vector.  As a field value this
        is a*2^n whenever the reference's limbs do not overflow."""
        return (a << n) % self.p

    def modshr(self, n, a):
        """pseudo.py:1068-1081: raw right shift, returns the shifted-out bits.  Only pinned
        for a canonical (redc'ed / modfsb'ed) input -- on an unreduced representative the
        reference's result depends on which representative its radix happened to hold."""
        a %= self.p
        return a >> n, a & ((1 << n) - 1)

    def modhaf(self, a):
        """pseudo.py:1084-1100: a/2 mod p."""
        a %= self.p
        return (a >> 1) if a % 2 == 0 else ((a + self.p) >> 1)

    def mod2r(self, r):
        """pseudo.py:1102-1112 / monty.py:1568-1577: 2^r, or 0 when r >= 8*Nbytes."""
        return 0 if r >= 8 * self.nbytes else (1 << r) % self.p

    def modexp(self, a) -> bytes:
        """pseudo.py:1115-1127: canonical value, big-endian, Nbytes."""
        return (a % self.p).to_bytes(self.nbytes, "big")

    def modimp(self, b: bytes):
        """pseudo.py:1130-1146: big-endian Nbytes, one modfsb (so the integer must be < 2p
        for the result to be a defined residue; any Nbytes string satisfies this for the
        three moduli here except the 38 values >= 2p of X25519, which still land on the right
        residue class); returns (value, 1 iff integer < p)."""
        v = int.from_bytes(b, "big")
        return v % self.p, int(v < self.p)

    def modsign(self, a):
        """pseudo.py:1149-1158."""
        return (a % self.p) & 1

    def modcmp(self, a, b):
        """pseudo.py:1161-1174."""
        return int((a - b) % self.p == 0)


def rfc7748(prime: Prime | str, bk: bytes, bu: bytes, twist_secure: bool = True) -> bytes:
    """rfc7748.c:156-256 restated: bv = clamp(bk) * bu on the Montgomery curve, all
    little-endian byte strings of Nbytes.  twist_secure=True is the TWIST_SECURE branch
    (rfc7748.c:225-227); False the cheap point validation of the #else branch (rfc7748.c:228-251),
    whose result is zero when bu is not on the curve."""
    F = FieldOracle(prime)
    P = F.P
    nb, nbits = P.nbytes, P.nbits
    ck = bytearray(bk)
    cu = bytearray(bu)[::-1]                      # reverse(): LE -> BE  (rfc7748.c:171)
    r = nbits % 8 or 8
    cu[0] &= (1 << r) - 1                         # mask()              (rfc7748.c:148-152,172)
    s = (8 - nbits % 8) % 8                       # clamp()             (rfc7748.c:135-141)
    ck[0] &= (-(1 << P.cof)) & 0xFF
    ck[nb - 1] &= 0xFF >> s
    ck[nb - 1] |= 0x80 >> s
    u, _ = F.modimp(bytes(cu))                    # rfc7748.c:178
    x1, x2, z2, x3, z3 = u, F.modone(), F.modzer(), u, F.modone()
    swap = 0
    for i in range(nbits - 1, -1, -1):            # rfc7748.c:186-221
        kt = (ck[i // 8] >> (i % 8)) & 1
        swap ^= kt
        x2, x3 = F.modcsw(swap, x2, x3)
        z2, z3 = F.modcsw(swap, z2, z3)
        swap = kt
        A = F.modadd(x2, z2)
        C = F.modadd(x3, z3)
        B = F.modsub(x2, z2)
        D = F.modsub(x3, z3)
        AA = F.modsqr(A)
        BB = F.modsqr(B)
        D = F.modmul(D, A)
        C = F.modmul(C, B)
        z3 = F.modsub(D, C)
        E = F.modsub(AA, BB)
        z2 = F.modmli(E, P.a24)
        x3 = F.modadd(D, C)
        z2 = F.modadd(z2, AA)
        z2 = F.modmul(z2, E)
        x3 = F.modsqr(x3)
        z3 = F.modsqr(z3)
        z3 = F.modmul(z3, x1)
        x2 = F.modmul(AA, BB)
    x2, x3 = F.modcsw(swap, x2, x3)
    z2, z3 = F.modcsw(swap, z2, z3)
    if twist_secure:
        A = F.modpro(z2)                          # rfc7748.c:226-227
        z2 = F.modinv(z2, A)
    else:                                         # rfc7748.c:228-251
        B = F.modmul(u, z2)
        A = F.modmul(B, z2)
        E = F.modpro(A)
        C = A
        D = F.modmul(E, z2)
        D = F.modsqr(D)
        D = F.modmul(D, u)
        for _ in range(P.cof - 2):
            C = F.modsqr(C)
            C = F.modmul(C, A)
        for _ in range(P.cof):
            E = F.modsqr(E)
        C = F.modmul(C, E)
        z2 = F.modmul(C, B)
        for _ in range(P.cof - 2):
            D = F.modsqr(D)
        A = F.modone()
        D = F.modadd(D, A)
        D, _ = F.modfsb(D)
        D, _ = F.modshr(1, D)                     # 1 for QR, else 0
        x2 = F.modmul(x2, D)
    x2 = F.modmul(x2, z2)                         # rfc7748.c:252
    return F.modexp(x2)[::-1]                     # rfc7748.c:254-255


def ecnmul(prime: Prime | str, e: bytes, x: bytes, y: bytes):
    """ecnXXXset(0,x,y,&P); ecnXXXmul(e,&P); ecnXXXget(&P,xo,yo) restated at value level
    (weierstrass.c:415-427 set with validation, :494-542 multiply, :297-349 affine/get; curve
    y^2 = x^3 - 3x + b, constants curve.py:157-166).  Big-endian Nbytes strings in and out.
    The point at infinity (point off the curve, e = 0 mod order) is reported as (0, 1)."""
    F = FieldOracle(prime)
    P, p, nb = F.P, F.p, F.nbytes
    xv, _ = F.modimp(x)
    yv, _ = F.modimp(y)
    on_curve = (yv * yv - (xv * xv * xv - 3 * xv + P.wb)) % p == 0      # setxy, weierstrass.c:364-396
    k = int.from_bytes(e, "big")

    def add(A, B):                      # affine group law with None = O
        if A is None:
            return B
        if B is None:
            return A
        (x1, y1), (x2, y2) = A, B
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1 - 3) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return x3, (lam * (x1 - x3) - y1) % p

    R, Q = None, ((xv, yv) if on_curve else None)
    while k:
        if k & 1:
            R = add(R, Q)
        Q = add(Q, Q)
        k >>= 1
    if R is None:
        return (0).to_bytes(nb, "big"), (1).to_bytes(nb, "big")        # ecnXXXaffine of O
    return R[0].to_bytes(nb, "big"), R[1].to_bytes(nb, "big")


def ecnmul_edwards(prime: Prime | str, e: bytes, x: bytes, y: bytes):
    """ecnXXXset(0,x,y,&P); ecnXXXmul(e,&P); ecnXXXget(&P,xo,yo) of edwards.c restated at value level
    (edwards.c:243-272 set with the on-curve check, :435-484 multiply, :184-241 affine/get) on the twisted
    Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 (curve.py:85-94).  The identity (and any point off the curve)
    is reported as (0, 1)."""
    F = FieldOracle(prime)
    P, p, nb = F.P, F.p, F.nbytes
    d = P.ed_d
    xv, _ = F.modimp(x)
    yv, _ = F.modimp(y)
    if (yv * yv - xv * xv - 1 - d * xv * xv * yv * yv) % p != 0:
        xv, yv = 0, 1
    k = int.from_bytes(e, "big")

    def add(A, B):
        (x1, y1), (x2, y2) = A, B
        t = d * x1 * x2 * y1 * y2 % p
        return ((x1 * y2 + y1 * x2) * pow(1 + t, -1, p) % p, (y1 * y2 + x1 * x2) * pow(1 - t, -1, p) % p)

    R, Q = (0, 1), (xv, yv)
    while k:
        if k & 1:
            R = add(R, Q)
        Q = add(Q, Q)
        k >>= 1
    return R[0].to_bytes(nb, "big"), R[1].to_bytes(nb, "big")


def ecnmul2(prime: Prime | str, e: bytes, x1: bytes, y1: bytes, f: bytes, x2: bytes, y2: bytes):
    """ecnXXXset(0,x1,y1,&P); ecnXXXset(0,x2,y2,&Q); ecnXXXmul2(e,&P,f,&Q,&R); ecnXXXget(&R,xo,yo)
    (weierstrass.c:545-572 / edwards.c:486-513: R = e*P + f*Q by a joint signed-digit scan) restated at
    value level as two single multiplications and one addition.  `prime` NIST256 -> the Weierstrass
    curve, X25519 -> Ed25519.  A point off the curve is replaced by the identity, as ecnXXXset does."""
    F = FieldOracle(prime)
    P, p, nb = F.P, F.p, F.nbytes
    a = ecnmul_edwards if P.wb is None else ecnmul
    ax, ay = a(prime, e, x1, y1)
    bx, by = a(prime, f, x2, y2)
    A = (int.from_bytes(ax, "big"), int.from_bytes(ay, "big"))
    B = (int.from_bytes(bx, "big"), int.from_bytes(by, "big"))
    if P.wb is None:                                  # Edwards: complete affine law, identity (0, 1)
        d = P.ed_d
        t = d * A[0] * B[0] * A[1] * B[1] % p
        R = ((A[0] * B[1] + A[1] * B[0]) * pow(1 + t, -1, p) % p, (A[1] * B[1] + A[0] * B[0]) * pow(1 - t, -1, p) % p)
        return R[0].to_bytes(nb, "big"), R[1].to_bytes(nb, "big")
    # Weierstrass: (0, 1) is how ecnXXXget reports O (it is not on the curve, so it cannot be a real point)
    inf = lambda T: T == (0, 1)
    if inf(A):
        R = B
    elif inf(B):
        R = A
    else:
        (u1, v1), (u2, v2) = A, B
        if u1 == u2 and (v1 + v2) % p == 0:
            R = (0, 1)
        else:
            lam = ((3 * u1 * u1 - 3) * pow(2 * v1, -1, p) if u1 == u2 else (v2 - v1) * pow(u2 - u1, -1, p)) % p
            u3 = (lam * lam - u1 - u2) % p
            R = (u3, (lam * (u1 - u3) - v1) % p)
    return R[0].to_bytes(nb, "big"), R[1].to_bytes(nb, "big")
