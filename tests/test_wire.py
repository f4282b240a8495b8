"""Wire / file formats: hex helpers of rfc7748.c:44-107, the Wycheproof
converter parse.py, and external XDH suites through the batch ladder."""
import json
import os

import numpy as np
import pytest

from field_oracle import rfc7748 as oracle_rfc7748
from modarith_b200 import wire
from modarith_b200.primes import PRIMES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# -- a literal restatement of the C helpers, character by character (the checker) ------------------
def _char2int(c):
    if "0" <= c <= "9":
        return ord(c) - ord("0")
    if "A" <= c <= "F":
        return ord(c) - ord("A") + 10
    if "a" <= c <= "f":
        return ord(c) - ord("a") + 10
    return 0


def _from_hex_c(src, nbytes):                      # rfc7748.c:81-97
    lz = max(0, 2 * nbytes - len(src))
    pad = ["0"] * lz + [src[i - lz] for i in range(lz, 2 * nbytes)]
    return bytes((_char2int(pad[2 * i]) * 16 + _char2int(pad[2 * i + 1])) & 0xff for i in range(nbytes))


@pytest.mark.parametrize("nbytes", [32, 56])
def test_from_hex_follows_the_c_helper(nbytes):
    rng = np.random.default_rng(nbytes)
    alphabet = "0123456789abcdefABCDEFgz -"
    strings = ["", "0", "f", "F" * (2 * nbytes), "9" * (2 * nbytes + 5)]
    for _ in range(200):
        n = int(rng.integers(0, 2 * nbytes + 8))
        strings.append("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n)))
    got = wire.from_hex(strings, nbytes)
    assert got.shape == (len(strings), nbytes) and got.dtype == np.uint8
    for s, row in zip(strings, got):
        assert row.tobytes() == _from_hex_c(s, nbytes), s


def test_to_hex_and_reverse_round_trip():
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    hx = wire.to_hex(a)
    assert hx == [bytes(r).hex() for r in a]                       # lower case, 2*Nbytes digits (rfc7748.c:55-78)
    assert (wire.from_hex(hx, 32) == a).all()
    assert (wire.from_hex([h.upper() for h in hx], 32) == a).all()
    r = wire.reverse(a)
    assert (r == a[:, ::-1]).all() and (wire.reverse(r) == a).all()
    assert wire.to_hex(a[0]) == [bytes(a[0]).hex()]
    assert wire.from_hex([], 32).shape == (0, 32)


@pytest.mark.parametrize("kind", ["ecdsa", "ed"])
def test_signature_converter_matches_parse_py(kind):
    """The golden text is the output of the reference's own parse.py on the same file."""
    base = os.path.join(GOLD, "wycheproof_sig", "%s_sample_test" % kind)
    text = open(base + ".json").read()
    vecs = wire.parse_signature_vectors(text, kind)
    assert wire.signature_lines(vecs) == open(base + ".parsed.txt").read()
    doc = json.loads(text)
    want = [(g["key"]["pk" if kind == "ed" else "uncompressed"], t["comment"], t["msg"], t["sig"], t["result"])
            for g in doc["testGroups"] for t in g["tests"]]
    assert [(v.public_key, v.comment, v.msg, v.sig, v.result) for v in vecs] == want
    with pytest.raises(ValueError):
        wire.parse_signature_vectors(text, "rsa")
    assert wire.parse_signature_vectors("{}", kind) == []


@pytest.mark.parametrize("curve", ["X25519", "X448"])
def test_xdh_suite_loads_and_agrees_with_the_oracle(curve):
    path = os.path.join(GOLD, "xdh_%s_sample.json" % curve.lower())
    vecs = wire.load_xdh_vectors(path)
    assert len(vecs) == json.load(open(path))["numberOfTests"] == 36
    assert wire.load_xdh_vectors(open(path).read())[3].public == vecs[3].public
    nb = PRIMES[curve].nbytes
    wrong = 0
    for v in vecs:
        out = oracle_rfc7748(PRIMES[curve], bytes.fromhex(v.private), bytes.fromhex(v.public))
        if out.hex() != v.shared:
            wrong += 1
            assert v.result == "invalid" and "shared secret with" in v.comment
        if "ZeroSharedSecret" in v.flags:
            assert out == bytes(nb)
    assert wrong == 2


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ["X25519", "X448"])
def test_xdh_suite_through_the_batch_ladder(curve):
    vecs = wire.load_xdh_vectors(os.path.join(GOLD, "xdh_%s_sample.json" % curve.lower()))
    bv, passed = wire.run_xdh_vectors(curve, vecs)
    assert passed.all(), [v.comment for v, ok in zip(vecs, passed) if not ok]
    got = wire.to_hex(bv)
    for v, h in zip(vecs, got):
        if "shared secret with" in v.comment:
            assert h != v.shared                       # the two planted wrong answers are caught
        else:
            assert h == v.shared
    # a suite runner that sees a wrong "valid" answer reports it
    bad = [type(vecs[0])(**{**vecs[0].__dict__, "shared": "00" * PRIMES[curve].nbytes})]
    assert not wire.run_xdh_vectors(curve, bad)[1].any()
