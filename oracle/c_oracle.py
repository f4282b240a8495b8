"""ctypes wrapper of oracle/oracle.c (the plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
PRIME_ID = {"X25519": 0, "X448": 1, "NIST256": 2, "SECP256K1": 3, "NIST256ORDER": 4}
NBYTES = {"X25519": 32, "X448": 56, "NIST256": 32, "SECP256K1": 32, "NIST256ORDER": 32}
OPS = {"mul": 0, "sqr": 1, "inv": 2, "sqrt": 3, "add": 4, "sub": 5, "neg": 6, "pro": 7, "id": 8, "mli": 9,
       "haf": 10, "qr": 11}
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-fopenmp", "-fvisibility=hidden", "-o", SO, src])
    return SO


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def rfc7748_batch(curve, k, u):
    """k, u: [n, Nbytes] uint8 little-endian -> [n, Nbytes]."""
    k = np.ascontiguousarray(k, dtype=np.uint8)
    u = np.ascontiguousarray(u, dtype=np.uint8)
    out = np.zeros_like(k)
    load().oracle_rfc7748_batch(PRIME_ID[curve], k.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                                out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(k.shape[0]))
    return out


def field_batch(prime, op, a, b=None, ib=0):
    """a, b: [n, Nbytes] uint8 big-endian -> ([n, Nbytes] canonical big-endian, status[n])."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    n = a.shape[0]
    out = np.zeros_like(a)
    st = np.zeros(n, dtype=np.int32)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint8)
        bp = b.ctypes.data_as(ctypes.c_void_p)
    load().oracle_field_batch(PRIME_ID[prime], OPS[op], a.ctypes.data_as(ctypes.c_void_p), bp, ctypes.c_int(ib),
                              out.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    return out, st
