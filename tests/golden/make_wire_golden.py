#!/usr/bin/env python3
"""Generate the wire-format fixtures from the REFERENCE ITSELF.

  wycheproof_sig/ecdsa_sample_test.json, ed_sample_test.json
        small files in the Wycheproof signature schema (hand-made here: the real suites are not
        available offline); values are arbitrary hex, only the layout matters to parse.py
  wycheproof_sig/*.parsed.txt
        what the reference's own parse.py prints for them (run from /root/reference, unmodified)
  xdh_x25519_sample.json, xdh_x448_sample.json
        files in the Wycheproof XDH schema whose `shared` fields come from the reference's
        rfc7748() (oracle/_ref/libref_<curve>.so): the RFC 7748 keys of rfc7748.c:271-275, edge
        u-coordinates (low order -> all-zero output, marked "invalid"; non-canonical -> "acceptable"),
        random rows, and two rows whose `shared` is deliberately wrong (marked "invalid").

Run:  python oracle/build_ref.py && python tests/golden/make_wire_golden.py
"""
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from modarith_b200.primes import PRIMES  # noqa: E402

REFSRC = "/root/reference"
REF = os.path.join(ROOT, "oracle", "_ref")


def sig_doc(kind, rng):
    def hx(n):
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes().hex()
    groups = []
    tc = 1
    for g in range(3):
        key = {"curve": "edwards25519" if kind == "ed" else "secp256r1", "keySize": 255 if kind == "ed" else 256,
               "type": "EDDSAPublicKey" if kind == "ed" else "EcPublicKey"}
        if kind == "ed":
            key["pk"] = hx(32)
            key["sk"] = hx(32)
        else:
            key["uncompressed"] = "04" + hx(64)
            key["wx"] = hx(32)
            key["wy"] = hx(32)
        tests = []
        for t in range(4 if g != 1 else 1):
            tests.append({"tcId": tc, "comment": ["", "special case hash", "signature with special case values for r and s",
                                                   "Signature malleability"][(g + t) % 4] if kind == "ecdsa" else "case %d" % tc,
                          "msg": hx((3 * tc) % 11), "sig": hx(64 if kind == "ed" else 70),
                          "result": ["valid", "invalid", "acceptable"][tc % 3], "flags": []})
            tc += 1
        groups.append({"key" if kind == "ecdsa" else "key": key, "type": "EddsaVerify" if kind == "ed" else "EcdsaVerify", "tests": tests})
    return {"algorithm": "EDDSA" if kind == "ed" else "ECDSA", "generatorVersion": "sample", "numberOfTests": tc - 1,
            "header": ["synthetic file in the Wycheproof layout"], "notes": {}, "schema": "sample", "testGroups": groups}


def make_sig():
    out = os.path.join(HERE, "wycheproof_sig")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(2718)
    for kind in ("ecdsa", "ed"):
        name = "%s_sample_test.json" % kind
        path = os.path.join(out, name)
        with open(path, "w") as f:
            json.dump(sig_doc(kind, rng), f, indent=2)
            f.write("\n")
        with tempfile.TemporaryDirectory() as d:
            shutil.copy(path, os.path.join(d, name))
            # parse.py decides the key field from the first letters of argv[1]: run it inside the directory
            txt = subprocess.run([sys.executable, os.path.join(REFSRC, "parse.py"), name], cwd=d, check=True,
                                 stdout=subprocess.PIPE, text=True).stdout
        with open(os.path.join(out, name.replace(".json", ".parsed.txt")), "w") as f:
            f.write(txt)
        print(name, txt.count("\n") // 5, "records")


def make_xdh(curve, sk1, sk2):
    P = PRIMES[curve]
    nb, p = P.nbytes, P.p
    lib = ctypes.CDLL(os.path.join(REF, "libref_%s.so" % curve))

    def ladder(k, u):
        o = ctypes.create_string_buffer(nb)
        lib.ref_rfc7748(k, u, o)
        return o.raw[:nb]

    rng = np.random.default_rng(448 + nb)
    g = P.generator.to_bytes(nb, "little")
    k1, k2 = bytes.fromhex(sk1), bytes.fromhex(sk2)
    tests = []

    def add(comment, k, u, result=None, flags=(), shared=None):
        s = ladder(k, u) if shared is None else shared
        if result is None:
            result = "invalid" if s == bytes(nb) else "valid"
        tests.append({"tcId": len(tests) + 1, "comment": comment, "public": u.hex(), "private": k.hex(),
                      "shared": s.hex(), "result": result, "flags": list(flags)})

    pk1, pk2 = ladder(k1, g), ladder(k2, g)
    add("RFC 7748 section 6: Alice's secret with Bob's public key", k1, pk2)
    add("RFC 7748 section 6: Bob's secret with Alice's public key", k2, pk1)
    add("public key of Alice from the base point", k1, g)
    nbits = P.nbits
    top = (1 << (8 * nb)) - 1
    for name, uval in (("public key = 0", 0), ("public key = 1", 1), ("public key = p-1", p - 1), ("public key = p", p),
                       ("public key = p+1", p + 1), ("public key with all bits set", top), ("public key = 2", 2)):
        u = (uval & top).to_bytes(nb, "little")
        canonical = uval < p
        flags = [] if canonical else ["NonCanonicalPublic"]
        s = ladder(k1, u)
        res = "invalid" if s == bytes(nb) else ("valid" if canonical else "acceptable")
        add(name, k1, u, result=res, flags=(flags + (["ZeroSharedSecret", "LowOrderPublic"] if s == bytes(nb) else [])))
    for i in range(24):
        k = rng.integers(0, 256, nb, dtype=np.uint8).tobytes()
        u = rng.integers(0, 256, nb, dtype=np.uint8).tobytes()
        if nbits % 8 == 0 or i % 3:
            u = u[:-1] + bytes([u[-1] & (0xff >> ((8 - nbits % 8) % 8))]) if nbits % 8 else u
        add("random row %d" % i, k, u)
    # two wrong answers: a suite runner must report them as not matching
    k = rng.integers(0, 256, nb, dtype=np.uint8).tobytes()
    good = ladder(k, g)
    bad = bytes([good[0] ^ 1]) + good[1:]
    add("shared secret with one bit flipped", k, g, result="invalid", shared=bad)
    add("shared secret with the last byte changed", k, pk1, result="invalid", shared=ladder(k, pk1)[:-1] + b"\x5a")
    doc = {"algorithm": "XDH", "generatorVersion": "sample", "numberOfTests": len(tests),
           "header": ["synthetic file in the Wycheproof XDH layout; shared secrets from the reference's rfc7748()"],
           "schema": "xdh_comp_schema.json",
           "testGroups": [{"curve": "curve25519" if curve == "X25519" else "curve448", "type": "XdhComp", "tests": tests}]}
    path = os.path.join(HERE, "xdh_%s_sample.json" % curve.lower())
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print(path, len(tests), "tests")


if __name__ == "__main__":
    make_sig()
    make_xdh("X25519", "77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
             "5dab087e624a8a4b79e17f8b83800ee66f3bb1292618b6fd1c2f8b27ff88e0eb")
    make_xdh("X448", "9a8f4925d1519f5775cf46b04b5800d4ee9ee8bae8bc5565d498c28dd9c9baf574a9419744897391006382a6f127ab1d9ac2d8c0a598726b",
             "1c306a7ac2a0e2e0990b294470cba339e6453772b075811d8fad0d1d6927c120bb5ee8972b0d3e21374c9c921b09d1b0366f10b65173992d")
