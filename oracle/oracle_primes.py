"""oracle/oracle_primes.py -- the oracle's OWN table of moduli and curve constants.

TEST INFRASTRUCTURE ONLY (see field_oracle.py).  Nothing here is imported from the product package: the
checker and the checked must not share code or data, so the constants are restated from the reference's
tables (file:line below) and tests/test_oracle_pinned.py compares the two tables with each other.
"""
from __future__ import annotations


class OraclePrime:
    """One modulus with the derived quantities the reference generators compute for it."""

    def __init__(self, name, p, **curve):
        self.name = name
        self.p = p
        self.nbits = p.bit_length()
        self.nbytes = (self.nbits + 7) // 8                     # pseudo.py:1611-1614
        k, t = 0, p - 1                                         # pseudo.py:1574-1579: PM1D2 = 2-adicity of p-1
        while t % 2 == 0:
            k += 1
            t //= 2
        self.pm1d2 = k
        self.pe = (p - 1 - (1 << k)) // (1 << (k + 1))          # pseudo.py:1580-1581: progenitor exponent
        if k == 1:                                              # pseudo.py:1616-1630: 2^k-th root of unity
            self.roi = p - 1
        elif k == 2:
            self.roi = pow(2, (p - 1) // 4, p)
        else:
            q = 2
            while pow(q, (p - 1) // 2, p) == 1:
                q += 1
            self.roi = pow(q, (p - 1) >> k, p)
        for key in ("a24", "cof", "generator", "ed_d", "ed_gx", "ed_gy", "ed_order", "wb", "wgx", "wgy", "worder"):
            setattr(self, key, curve.get(key))


TABLE = {q.name: q for q in (
    # rfc7748.c:120-124 (X25519: A24 121665, COF 3, GENERATOR 9); Ed25519 constants curve.py:85-94
    OraclePrime("X25519", (1 << 255) - 19, a24=121665, cof=3, generator=9,
                ed_d=0x52036CEE2B6FFE738CC740797779E89800700A4D4141D8AB75EB4DCA135978A3,
                ed_gx=0x216936D3CD6E53FEC0A4E231FDD6DC5C692CC7609525A7B2C9562D608F25D51A,
                ed_gy=0x6666666666666666666666666666666666666666666666666666666666666658,
                ed_order=0x1000000000000000000000000000000014DEF9DEA2F79CD65812631A5CF5D3ED),
    # rfc7748.c:127-131 (X448: A24 39081, COF 2, GENERATOR 5); modulus monty.py named table "X448"
    OraclePrime("X448", (1 << 448) - (1 << 224) - 1, a24=39081, cof=2, generator=5),
    # curve.py:157-166 (NIST256: p, q, B, X, Y as decimal / hex literals there)
    OraclePrime("NIST256", 115792089210356248762697446949407573530086143415290314195533631308867097853951,
                wb=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
                wgx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
                wgy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
                worder=115792089210356248762697446949407573529996955224135760342422259061068512044369),
    # monty.py:2066-2067 (SECP256K1 field prime)
    OraclePrime("SECP256K1", (1 << 256) - (1 << 32) - 977),
    # the order of the P-256 group, the reference's "00<decimal>" mode (monty.py:2110-2127; q of curve.py:159)
    OraclePrime("NIST256ORDER", 115792089210356248762697446949407573529996955224135760342422259061068512044369),
)}


def lookup(prime):
    """A table entry by name, or an ad-hoc modulus from any object with .p (and optional curve constants)."""
    if isinstance(prime, str):
        return TABLE[prime]
    if isinstance(prime, OraclePrime):
        return prime
    extra = {k: getattr(prime, k, None) for k in ("a24", "cof", "generator", "ed_d", "ed_gx", "ed_gy", "ed_order",
                                                   "wb", "wgx", "wgy", "worder")}
    return OraclePrime(getattr(prime, "name", "P%d" % prime.p.bit_length()), prime.p, **extra)
