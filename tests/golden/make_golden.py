#!/usr/bin/env python3
"""Generate tests/golden/*.json from the REFERENCE ITSELF (oracle/_ref, built by
oracle/build_ref.py from /root/reference in the build container).

The committed JSON files are what travels; /root/reference does not exist on the GPU box.
Run:  python oracle/build_ref.py && python tests/golden/make_golden.py

Contents
  rfc7748.json  * RFC 7748 6.1/6.2 keys embedded in rfc7748.c:271,274 and
                  simd/rfc7748_simt.cu:245,249 -> public keys and shared secrets
                * the deterministic demo of rfc7748.c:main (LCG rnd=5*rnd+1, rfc7748.c:263,
                  298-305,321-333): value after the 5000x2 loop and the DH secret
                * fixed edge rows (u in {0,1,p-1,p,p+1,2^n-1,generator}, k in {0,all-ones})
                * 64 random rows per curve (numpy PCG64 seed 7748, raw unclamped bytes)
                * "validate": 40 random + the edge rows through the driver compiled without
                  TWIST_SECURE (point-validation tail, rfc7748.c:228-251)
  field.json    per modulus, per operation: inputs (big-endian hex) -> modexp output of the
                reference's generated 64-bit C, on the edge.py-style operand set
                (edge.py:341-360 regenerated, not copied) plus 24 random operands
  checksums     the `time` binary checksums (pseudo.py:1862-1866) are recorded in
                tests/test_oracle_pinned.py from oracle/_ref/build_*.log
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
from modarith_b200.primes import PRIMES, ALL_PRIMES  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OPS = {"mul": 0, "sqr": 1, "inv": 2, "sqrt": 3, "add": 4, "sub": 5, "neg": 6, "pro": 7, "id": 8, "mli": 9,
       "haf": 10, "qr": 11}


def ref(name):
    lib = ctypes.CDLL(os.path.join(REF, "libref_%s.so" % name))
    return lib


def call_rfc(lib, nb, k: bytes, u: bytes) -> bytes:
    out = ctypes.create_string_buffer(nb)
    lib.ref_rfc7748(k, u, out)
    return out.raw[:nb]


def curve_vectors(name, sk1, sk2):
    P = PRIMES[name]
    nb = P.nbytes
    lib = ref(name)
    g = P.generator.to_bytes(nb, "little")
    out = {"nbytes": nb}
    k1, k2 = bytes.fromhex(sk1), bytes.fromhex(sk2)
    pk1, pk2 = call_rfc(lib, nb, k1, g), call_rfc(lib, nb, k2, g)
    out["rfc"] = {"sk1": sk1, "sk2": sk2, "pk1": pk1.hex(), "pk2": pk2.hex(),
                  "shared": call_rfc(lib, nb, k1, pk2).hex(), "shared_check": call_rfc(lib, nb, k2, pk1).hex()}
    # demo main(): rfc7748.c:296-339
    rnd = 1
    bk = bytearray(nb)
    for i in range(nb):
        rnd = (5 * rnd + 1) & 0xFFFF
        bk[i] = rnd % 256
    bu = g
    for _ in range(5000):
        bv = call_rfc(lib, nb, bytes(bk), bu)
        bu = call_rfc(lib, nb, bytes(bk), bv)
    alice, bob = bytearray(nb), bytearray(nb)
    for i in range(nb):
        rnd = (5 * rnd + 1) & 0xFFFF
        alice[i] = rnd % 256
        rnd = (5 * rnd + 1) & 0xFFFF
        bob[i] = rnd % 256
    apk, bpk = call_rfc(lib, nb, bytes(alice), g), call_rfc(lib, nb, bytes(bob), g)
    out["demo"] = {"key": bytes(bk).hex(), "loop5000": bu.hex(), "alice": bytes(alice).hex(), "bob": bytes(bob).hex(),
                   "ssa": call_rfc(lib, nb, bytes(alice), bpk).hex(), "ssb": call_rfc(lib, nb, bytes(bob), apk).hex()}
    # edge rows
    p = P.p
    us = [0, 1, p - 1, p, p + 1, (1 << P.nbits) - 1, (1 << (8 * nb)) - 1, P.generator, 2, p - 2, 1 << (P.nbits - 1)]
    ks = [0, (1 << (8 * nb)) - 1, 1, 8, int.from_bytes(k1, "little")]
    rows = []
    for u in us:
        for k in ks:
            kb, ub = k.to_bytes(nb, "little"), (u % (1 << (8 * nb))).to_bytes(nb, "little")
            rows.append({"k": kb.hex(), "u": ub.hex(), "out": call_rfc(lib, nb, kb, ub).hex()})
    out["edge"] = rows
    rng = np.random.Generator(np.random.PCG64(7748))
    k = rng.integers(0, 256, (64, nb), dtype=np.uint8)
    u = rng.integers(0, 256, (64, nb), dtype=np.uint8)
    out["random"] = [{"k": k[i].tobytes().hex(), "u": u[i].tobytes().hex(),
                      "out": call_rfc(lib, nb, k[i].tobytes(), u[i].tobytes()).hex()} for i in range(64)]
    # the driver built WITHOUT TWIST_SECURE (rfc7748.c:228-251): same inputs, plus the edge rows
    vlib = ref(name + "_validate")
    out["validate"] = [{"k": r["k"], "u": r["u"],
                        "out": call_rfc(vlib, nb, bytes.fromhex(r["k"]), bytes.fromhex(r["u"])).hex()}
                       for r in out["random"][:40] + out["edge"]]
    return out


def edge_operands(P):
    """The corner-case operand set of edge.py:341-360, regenerated for this modulus."""
    p, n = P.p, P.nbits
    c = (1 << n) - p
    r = 0x1234567890ABCDEFFEDCBA9876543210 % p
    vals = [p - 1, 0, 1, p - 2, 2, (1 << (n - 1)), c % p, (1 << 64) % p, r, pow(r, -1, p), (p + 1) // 2,
            (1 << 32) - 1, (1 << 96), p - c % p if c % p else 0, (1 << (8 * P.nbytes)) - 1, p, p + 1]
    lim = min(2 * p, 1 << (8 * P.nbytes))
    return [v for v in vals if 0 <= v < lim]


def field_vectors(name):
    P = ALL_PRIMES[name]
    nb = P.nbytes
    lib = ref(name + "_generic" if name in ("X25519", "X448") else name)
    rng = np.random.Generator(np.random.PCG64(256 + nb))
    edge = edge_operands(P)
    a_vals = list(edge)
    b_vals = edge[1:] + edge[:1]
    for _ in range(24):
        a_vals.append(int.from_bytes(rng.integers(0, 256, nb, dtype=np.uint8).tobytes(), "big") % min(2 * P.p, 1 << (8 * nb)))
        b_vals.append(int.from_bytes(rng.integers(0, 256, nb, dtype=np.uint8).tobytes(), "big") % min(2 * P.p, 1 << (8 * nb)))
    n = len(a_vals)
    A = b"".join(v.to_bytes(nb, "big") for v in a_vals)
    B = b"".join(v.to_bytes(nb, "big") for v in b_vals)
    out = {"nbytes": nb, "a": [v.to_bytes(nb, "big").hex() for v in a_vals],
           "b": [v.to_bytes(nb, "big").hex() for v in b_vals], "mli_int": 121665, "ops": {}}
    for op, code in OPS.items():
        buf = ctypes.create_string_buffer(n * nb)
        st = (ctypes.c_int * n)()
        lib.ref_field_batch(code, A, B if op in ("mul", "add", "sub") else None, 121665, buf, st,
                            ctypes.c_size_t(n), 1)
        res = [buf.raw[i * nb:(i + 1) * nb].hex() for i in range(n)]
        out["ops"][op] = {"out": res, "status": list(st)}
    return out


def ecn_vectors(curve="NIST256"):
    """weierstrass.c / edwards.c as patched by the reference's own curve.py (oracle/_ref/libref_<curve>_curve.so):
    ecnXXXset + ecnXXXmul + ecnXXXget on generator multiples, with off-curve points, zero scalars, the group
    order and order+-1, all-ones scalars and small scalars mixed in."""
    if curve == "NIST256":
        P = ALL_PRIMES["NIST256"]
        G = (P.wgx.to_bytes(32, "big"), P.wgy.to_bytes(32, "big"))
        order = P.worder
    else:
        P = ALL_PRIMES["X25519"]
        G = (P.ed_gx.to_bytes(32, "big"), P.ed_gy.to_bytes(32, "big"))
        order = P.ed_order
    lib = ref(curve + "_curve")
    rng = np.random.Generator(np.random.PCG64(1060))

    def mul(e, x, y):
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        lib.ref_ecnmul_batch(e, x, y, xo, yo, ctypes.c_size_t(1), 1)
        return xo.raw[:32], yo.raw[:32]

    rows = []
    pts = [G]
    for _ in range(10):
        pts.append(mul(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), *G))
    scalars = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, order - 1, order, order + 1, (8 * order) % (1 << 256), (1 << 256) - 1, 1 << 255,
               0x8888888888888888888888888888888888888888888888888888888888888888,
               0x7777777777777777777777777777777777777777777777777777777777777777]
    for k in scalars:
        x, y = pts[len(rows) % len(pts)]
        rows.append((k.to_bytes(32, "big"), x, y))
    for i in range(24):
        x, y = pts[i % len(pts)]
        rows.append((rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), x, y))
    bad = bytearray(pts[1][1]); bad[31] ^= 1
    rows.append((rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), pts[1][0], bytes(bad)))       # off the curve
    rows.append((rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), bytes(32), bytes(32)))          # (0,0)
    rows.append((rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), b"\xff" * 32, pts[2][1]))      # x >= p
    out = []
    for e, x, y in rows:
        xo, yo = mul(e, x, y)
        out.append({"e": e.hex(), "x": x.hex(), "y": y.hex(), "xo": xo.hex(), "yo": yo.hex()})
    return out


def ecn2_vectors(curve="NIST256"):
    """ecnXXXset x2 + ecnXXXmul2 + ecnXXXget of the reference (weierstrass.c:545-572 / edwards.c:486-513):
    R = e*P + f*Q on random pairs and on the cases its joint-digit scan has to get right: zero scalars, P = Q,
    Q = -P, e + f = order, order-1, all-ones scalars, a point that is not on the curve.  e = f = 0 is left out:
    the reference's leading-zero skip (`while (w[i]==0) i--`) then walks off the front of its digit array;
    the tests pin that case to the identity through the oracle instead."""
    if curve == "NIST256":
        P = ALL_PRIMES["NIST256"]
        G = (P.wgx.to_bytes(32, "big"), P.wgy.to_bytes(32, "big"))
        order, p = P.worder, P.p
        neg = lambda pt: (pt[0], ((p - int.from_bytes(pt[1], "big")) % p).to_bytes(32, "big"))
    else:
        P = ALL_PRIMES["X25519"]
        G = (P.ed_gx.to_bytes(32, "big"), P.ed_gy.to_bytes(32, "big"))
        order, p = P.ed_order, P.p
        neg = lambda pt: (((p - int.from_bytes(pt[0], "big")) % p).to_bytes(32, "big"), pt[1])
    lib = ref(curve + "_curve")
    rng = np.random.Generator(np.random.PCG64(2015))

    def mul(e, x, y):
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        lib.ref_ecnmul_batch(e, x, y, xo, yo, ctypes.c_size_t(1), 1)
        return xo.raw[:32], yo.raw[:32]

    def mul2(e, x1, y1, f, x2, y2):
        xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        lib.ref_ecnmul2_batch(e, x1, y1, f, x2, y2, xo, yo, ctypes.c_size_t(1), 1)
        return xo.raw[:32], yo.raw[:32]

    rb = lambda: rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    b = lambda k: (k % (1 << 256)).to_bytes(32, "big")
    A, B = mul(rb(), *G), mul(rb(), *G)
    k1 = int.from_bytes(rb(), "big") % order
    rows = [(rb(), A, rb(), B) for _ in range(16)]
    rows += [(b(0), A, rb(), B), (rb(), A, b(0), B), (b(1), A, b(1), B), (b(1), G, b(2), G),
             (rb(), A, rb(), A),                                   # P = Q
             (b(k1), A, b(order - k1), A),                         # e + f = order, same point -> O
             (b(k1), A, b(k1), neg(A)),                            # Q = -P, same scalar -> O
             (b(5), A, b(3), neg(A)),                              # 5P - 3P = 2P
             (b(order - 1), A, b(order - 1), B), (b((1 << 256) - 1), A, b((1 << 256) - 1), B),
             (b(order), A, b(7), B), (b(1 << 255), G, b((1 << 255) + 1), A),
             (b(0x5555555555555555555555555555555555555555555555555555555555555555), A,
              b(0xaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa), B)]
    bad = bytearray(A[1]); bad[31] ^= 1
    rows += [(rb(), (A[0], bytes(bad)), rb(), B), (rb(), A, rb(), (bytes(32), bytes(32)))]   # one operand off the curve
    out = []
    for e, Pt, f, Qt in rows:
        xo, yo = mul2(e, Pt[0], Pt[1], f, Qt[0], Qt[1])
        out.append({"e": e.hex(), "x1": Pt[0].hex(), "y1": Pt[1].hex(), "f": f.hex(), "x2": Qt[0].hex(), "y2": Qt[1].hex(),
                    "xo": xo.hex(), "yo": yo.hex()})
    return out


def testcurve_chain(curve, iters=150):
    """The reference's own group test (testcurve.c main): its per-curve constants order, r1, r2 (= order - r1),
    n1, n2 are read from /root/reference/testcurve.c at generation time; the generator comes from ecnXXXgen;
    the two timing loops (P = n1*P; then P = n1*P + n2*Q with Q the generator) are run `iters` times each
    instead of 10000 through the same reference functions (oracle/ref_curve_shim.c)."""
    import re
    src = open("/root/reference/testcurve.c").read()
    blk = src[src.index("#ifdef %s\n    const char order" % curve):]
    blk = blk[:blk.index("#endif")]
    consts = {k: re.search(r'const char %s\[\]=\s*"([0-9A-Fa-f]+)"' % k, blk).group(1) for k in ("order", "r1", "r2", "n1", "n2")}
    lib = ref(curve + "_curve")
    be = lambda h: int(h, 16).to_bytes(32, "big")
    bufs = [ctypes.create_string_buffer(32) for _ in range(6)]
    lib.ref_testcurve_chain(be(consts["n1"]), be(consts["n2"]), ctypes.c_int(iters), *bufs)
    gx, gy, x1, y1, x2, y2 = (b.raw[:32].hex() for b in bufs)
    out = dict(consts, iters=iters, gx=gx, gy=gy, x1=x1, y1=y1, x2=x2, y2=y2)
    # the two pass/fail checks of testcurve.c: order*G = O, r1*G + r2*G = O
    xo, yo = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    lib.ref_ecnmul_batch(be(consts["order"]), bytes.fromhex(gx), bytes.fromhex(gy), xo, yo, ctypes.c_size_t(1), 1)
    out["mul_test"] = [xo.raw[:32].hex(), yo.raw[:32].hex()]
    lib.ref_ecnmul2_batch(be(consts["r1"]), bytes.fromhex(gx), bytes.fromhex(gy), be(consts["r2"]), bytes.fromhex(gx), bytes.fromhex(gy),
                          xo, yo, ctypes.c_size_t(1), 1)
    out["mul2_test"] = [xo.raw[:32].hex(), yo.raw[:32].hex()]
    return out


def main():
    with open(os.path.join(HERE, "testcurve.json"), "w") as f:
        json.dump({"NIST256": testcurve_chain("NIST256"), "ED25519": testcurve_chain("ED25519")}, f, indent=1)
    with open(os.path.join(HERE, "ecn.json"), "w") as f:
        json.dump({"NIST256": ecn_vectors("NIST256"), "ED25519": ecn_vectors("ED25519")}, f, indent=1)
    with open(os.path.join(HERE, "ecn2.json"), "w") as f:
        json.dump({"NIST256": ecn2_vectors("NIST256"), "ED25519": ecn2_vectors("ED25519")}, f, indent=1)
    rfc = {
        "X25519": curve_vectors("X25519", "77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
                                "5dab087e624a8a4b79e17f8b83800ee66f3bb1292618b6fd1c2f8b27ff88e0eb"),
        "X448": curve_vectors("X448", "9a8f4925d1519f5775cf46b04b5800d4ee9ee8bae8bc5565d498c28dd9c9baf574a9419744897391006382a6f127ab1d9ac2d8c0a598726b",
                              "1c306a7ac2a0e2e0990b294470cba339e6453772b075811d8fad0d1d6927c120bb5ee8972b0d3e21374c9c921b09d1b0366f10b65173992d"),
    }
    with open(os.path.join(HERE, "rfc7748.json"), "w") as f:
        json.dump(rfc, f, indent=1)
    fld = {name: field_vectors(name) for name in ALL_PRIMES}
    with open(os.path.join(HERE, "field.json"), "w") as f:
        json.dump(fld, f, indent=1)
    print("wrote rfc7748.json, field.json")
    for c in rfc:
        print(c, rfc[c]["rfc"]["pk1"], rfc[c]["demo"]["loop5000"][:16], rfc[c]["demo"]["ssa"][:16])


if __name__ == "__main__":
    main()
