"""Host-side mirror of the reference's generated-code API, batched on one B200.

    F = Field("X25519")
    x = F.modimp(bytes_be)            # [n, Nbytes] uint8 cuda tensor -> limb planes
    F.modmul(x, y, z)                 # same names / argument order as the generated C
    out = F.modexp(z)                 # canonical big-endian bytes

Every method forwards to one `mab_<PRIME>_<function>` entry point of the C ABI
(include/modarith_b200.h) on the current torch CUDA stream.  torch is used only to own
device memory and streams.  A batch of n field elements is an int32 tensor of shape
[Nlimbs, n] on the GPU ("limb planes": plane j holds limb j of every element), opaque like
the reference's spint[Nlimbs]; outputs may alias inputs exactly as in the reference
(pseudo.py:1832-1845).
"""
from __future__ import annotations

import torch

from . import lib as _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class Field:
    def __init__(self, prime: str, device=None):
        import os
        if prime not in _lib.PRIMES and not os.path.exists(_lib.extra_lib_path(prime)):
            raise ValueError("unsupported modulus %r (built in: %s; any other one after `python -m modarith_b200.build "
                             "--prime %s[=<expression>]`)" % (prime, ", ".join(_lib.PRIMES), prime))
        self.lib = _lib.load_for(prime)
        if not torch.cuda.is_available():
            raise _lib.MabError("modarith_b200 needs a CUDA device: there is no CPU fallback")
        self.prime = prime
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        par = _lib.params(prime)
        self.Wordlength, self.Nlimbs, self.Radix = par["wordlength"], par["nlimbs"], par["radix"]
        self.Nbits, self.Nbytes = par["nbits"], par["nbytes"]

    # -- memory ----------------------------------------------------------------------------
    def alloc(self, n: int) -> torch.Tensor:
        return torch.empty((self.Nlimbs, n), dtype=torch.int32, device=self.device)

    def _ints(self, n):
        return torch.empty((n,), dtype=torch.int32, device=self.device)

    def _call(self, name, lead, planes):
        """planes: a tensor defining n/stride."""
        if planes.dtype != torch.int32 or planes.dim() != 2 or planes.shape[0] != self.Nlimbs:
            raise ValueError("expected int32 limb planes [%d, n]" % self.Nlimbs)
        if planes.shape[1] > 1 and planes.stride(1) != 1:
            raise ValueError("limb planes must be contiguous along the element axis")
        n, stride = planes.shape[1], planes.stride(0) if planes.shape[1] > 0 else 0
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = getattr(self.lib, "mab_%s_%s" % (self.prime, name))
        with torch.cuda.device(self.device):
            _lib.check(fn(*lead, n, max(stride, n), stream), "mab_%s_%s" % (self.prime, name), self.lib)

    def _chk(self, *ts):
        ref = None
        for t in ts:
            if t is None:
                continue
            if not isinstance(t, torch.Tensor) or not t.is_cuda:
                raise TypeError("field elements are CUDA tensors (int32 limb planes [%d, n])" % self.Nlimbs)
            if t.dtype != torch.int32 or t.dim() != 2 or t.shape[0] != self.Nlimbs:
                raise ValueError("expected int32 limb planes [%d, n] on the GPU, got %s %s" % (self.Nlimbs, t.dtype, tuple(t.shape)))
            if t.device != self.device:
                raise ValueError("operand lives on %s, this Field on %s" % (t.device, self.device))
            if ref is None:
                ref = t
            elif t.shape != ref.shape or t.stride() != ref.stride():
                raise ValueError("operands must share shape and pitch")
        return ref

    def _bits(self, b, g):
        if not isinstance(b, torch.Tensor) or not b.is_cuda or b.dtype != torch.int32 or b.shape != (g.shape[1],) \
                or not b.is_contiguous():
            raise ValueError("swap / move bits must be a contiguous int32 CUDA tensor of shape [n]")

    def _bytes(self, b, n, name):
        if not isinstance(b, torch.Tensor) or not b.is_cuda or b.dtype != torch.uint8:
            raise TypeError("%s must be a uint8 CUDA tensor" % name)
        if b.dim() != 2 or b.shape[1] != self.Nbytes or (n is not None and b.shape[0] != n) or not b.is_contiguous():
            raise ValueError("%s must be a contiguous [n, %d] byte array" % (name, self.Nbytes))
        if b.device != self.device:
            raise ValueError("%s lives on %s, this Field on %s" % (name, b.device, self.device))

    # -- the generated-code API (reference argument order) ------------------------------------
    def modfsb(self, n_):
        out = self._ints(n_.shape[1])
        self._call("modfsb", [_ptr(n_), _ptr(out)], self._chk(n_))
        return out

    def modadd(self, a, b, n_):
        self._call("modadd", [_ptr(a), _ptr(b), _ptr(n_)], self._chk(a, b, n_))

    def modsub(self, a, b, n_):
        self._call("modsub", [_ptr(a), _ptr(b), _ptr(n_)], self._chk(a, b, n_))

    def modneg(self, b, n_):
        self._call("modneg", [_ptr(b), _ptr(n_)], self._chk(b, n_))

    def modmul(self, a, b, c):
        self._call("modmul", [_ptr(a), _ptr(b), _ptr(c)], self._chk(a, b, c))

    def modsqr(self, a, c):
        self._call("modsqr", [_ptr(a), _ptr(c)], self._chk(a, c))

    def bench_modmul(self, a, b, c, iters: int):
        """measurement helper: c = a * b^iters with the running product in registers"""
        self._call("bench_modmul", [_ptr(a), _ptr(b), _ptr(c), int(iters)], self._chk(a, b, c))

    def modmli(self, a, b: int, c):
        self._call("modmli", [_ptr(a), int(b), _ptr(c)], self._chk(a, c))

    def modcpy(self, a, c):
        self._call("modcpy", [_ptr(a), _ptr(c)], self._chk(a, c))

    def modnsqr(self, a, n: int):
        self._call("modnsqr", [_ptr(a), int(n)], self._chk(a))

    def modpro(self, w, z):
        self._call("modpro", [_ptr(w), _ptr(z)], self._chk(w, z))

    def modinv(self, x, h, z):
        self._call("modinv", [_ptr(x), _ptr(h), _ptr(z)], self._chk(x, h, z))

    def modinv_perelement(self, x, z):
        """modinv with one progenitor chain per element (comparison; `modinv` shares chains)"""
        self._call("modinv_perelement", [_ptr(x), _ptr(z)], self._chk(x, z))

    def modqr(self, h, x):
        out = self._ints(x.shape[1])
        self._call("modqr", [_ptr(h), _ptr(x), _ptr(out)], self._chk(x, h))
        return out

    def modsqrt(self, x, h, r):
        self._call("modsqrt", [_ptr(x), _ptr(h), _ptr(r)], self._chk(x, h, r))

    def modis1(self, a):
        out = self._ints(a.shape[1])
        self._call("modis1", [_ptr(a), _ptr(out)], self._chk(a))
        return out

    def modis0(self, a):
        out = self._ints(a.shape[1])
        self._call("modis0", [_ptr(a), _ptr(out)], self._chk(a))
        return out

    def modzer(self, a):
        self._call("modzer", [_ptr(a)], self._chk(a))

    def modone(self, a):
        self._call("modone", [_ptr(a)], self._chk(a))

    def modint(self, x: int, a):
        self._call("modint", [int(x), _ptr(a)], self._chk(a))

    def nres(self, m, n_):
        self._call("nres", [_ptr(m), _ptr(n_)], self._chk(m, n_))

    def redc(self, n_, m):
        self._call("redc", [_ptr(n_), _ptr(m)], self._chk(n_, m))

    def modcsw(self, b, g, f):
        self._bits(b, g)
        self._call("modcsw", [_ptr(b), _ptr(g), _ptr(f)], self._chk(g, f))

    def modcmv(self, b, g, f):
        self._bits(b, g)
        self._call("modcmv", [_ptr(b), _ptr(g), _ptr(f)], self._chk(g, f))

    def modshl(self, n: int, a):
        self._call("modshl", [int(n), _ptr(a)], self._chk(a))

    def modshr(self, n: int, a):
        out = self._ints(a.shape[1])
        self._call("modshr", [int(n), _ptr(a), _ptr(out)], self._chk(a))
        return out

    def modhaf(self, a):
        self._call("modhaf", [_ptr(a)], self._chk(a))

    def mod2r(self, r: int, a):
        self._call("mod2r", [int(r), _ptr(a)], self._chk(a))

    def modexp(self, a, b=None):
        n = a.shape[1]
        if b is None:
            b = torch.empty((n, self.Nbytes), dtype=torch.uint8, device=self.device)
        self._bytes(b, n, "b")
        self._call("modexp", [_ptr(a), _ptr(b)], self._chk(a))
        return b

    def modimp(self, b, a=None):
        """b: [n, Nbytes] uint8 big-endian.  Returns (planes, status) with status[i]=1 iff < p."""
        self._bytes(b, None, "b")
        n = b.shape[0]
        if a is None:
            a = self.alloc(n)
        st = self._ints(n)
        self._call("modimp", [_ptr(b), _ptr(a), _ptr(st)], self._chk(a))
        return a, st

    def modsign(self, a):
        out = self._ints(a.shape[1])
        self._call("modsign", [_ptr(a), _ptr(out)], self._chk(a))
        return out

    def modcmp(self, a, b):
        out = self._ints(a.shape[1])
        self._call("modcmp", [_ptr(a), _ptr(b), _ptr(out)], self._chk(a, b))
        return out

    # -- a sequence of the calls above in ONE launch (mab_<P>_modprog) ---------------------------
    @staticmethod
    def _encode_program(code, out_regs):
        if not code or len(code) > _lib.PROG_MAX:
            raise ValueError("a program has 1 .. %d instructions" % _lib.PROG_MAX)
        if len(out_regs) > _lib.PROG_NREG or not out_regs:
            raise ValueError("1 .. %d outputs" % _lib.PROG_NREG)
        arr = (_lib.mab_insn * len(code))()
        for k, ins in enumerate(code):
            if len(ins) not in (4, 5) or ins[0] not in _lib.OPCODES:
                raise ValueError("instruction %d: expected (op, dst, a, b[, imm]) with a known op, got %r" % (k, ins))
            regs = ins[1:4]
            if any((not isinstance(r, int)) or r < 0 or r >= _lib.PROG_NREG for r in regs):
                raise ValueError("instruction %d: register numbers are 0 .. %d" % (k, _lib.PROG_NREG - 1))
            imm = int(ins[4]) if len(ins) == 5 else 0
            if imm < 0 or imm > 0x7FFFFFFF:
                raise ValueError("instruction %d: the small-integer operand must be in [0, 2^31)" % k)
            arr[k].op, arr[k].dst, arr[k].a, arr[k].b, arr[k].imm = _lib.OPCODES[ins[0]], regs[0], regs[1], regs[2], imm
        if any((not isinstance(r, int)) or r < 0 or r >= _lib.PROG_NREG for r in out_regs):
            raise ValueError("output registers are 0 .. %d" % (_lib.PROG_NREG - 1))
        return arr

    @classmethod
    def modprog_cubin(cls, prime, code, nin, out_regs):
        """The sm_100a cubin mab_<P>_modprog_jit compiles for this program (bytes).  Needs NVRTC, no device."""
        import ctypes
        lib = _lib.load_for(prime)
        arr = cls._encode_program(code, out_regs)
        if not (0 <= nin <= _lib.PROG_NREG):
            raise ValueError("at most %d inputs" % _lib.PROG_NREG)
        regs_p = (ctypes.c_ubyte * len(out_regs))(*out_regs)
        fn = getattr(lib, "mab_%s_modprog_cubin" % prime)
        cap = 1 << 22
        while True:
            buf = ctypes.create_string_buffer(cap)
            size = ctypes.c_size_t(cap)
            rc = fn(arr, len(code), nin, regs_p, len(out_regs), buf, ctypes.byref(size))
            if rc != 0 and size.value > cap:          # the cubin is larger than the buffer: its size came back
                cap = size.value
                continue
            _lib.check(rc, "mab_%s_modprog_cubin" % prime, lib)
            return buf.raw[:size.value]

    def modprog(self, code, inputs, out_regs, outputs=None, jit=False):
        """Run a straight-line program with its variables held on chip.

        jit=False: mab_<P>_modprog, the interpreter (variables in shared memory, no compiler needed);
        jit=True : mab_<P>_modprog_jit, the program compiled by NVRTC into a kernel of its own (variables in machine
                   registers; the first call of a new program compiles, later calls hit the cache).  Same results.

        code     : [(op, dst, a, b)] or [(op, dst, a, b, imm)] with op one of add sub neg mul sqr mli cpy nsqr pro
                   inv sqrt zer one int haf; dst / a / b are register numbers 0..15 (unused ones may be 0);
                   semantics of each op = the API function of the same name
        inputs   : limb-plane tensors, loaded into registers 0, 1, ... before the first instruction
        out_regs : registers to store after the last instruction; returns one plane tensor per entry
                   (written into `outputs` when given)
        E.g. (x + y)^2 - x*y:  F.modprog([("add", 2, 0, 1), ("sqr", 2, 2, 0), ("mul", 3, 0, 1), ("sub", 2, 2, 3)], [x, y], [2])"""
        import ctypes
        if len(inputs) > _lib.PROG_NREG:
            raise ValueError("at most %d inputs" % _lib.PROG_NREG)
        arr = self._encode_program(code, out_regs)
        ref = self._chk(*inputs) if inputs else None
        if outputs is None:
            if ref is None:
                raise ValueError("a program without inputs needs explicit output tensors (they define n)")
            outputs = [torch.empty_like(ref) for _ in out_regs]
        if len(outputs) != len(out_regs):
            raise ValueError("one output tensor per output register")
        ref = self._chk(*(list(inputs) + list(outputs)))
        n, stride = ref.shape[1], ref.stride(0) if ref.shape[1] > 0 else 0
        if ref.shape[1] > 1 and ref.stride(1) != 1:
            raise ValueError("limb planes must be contiguous along the element axis")
        ins_p = (ctypes.c_void_p * max(1, len(inputs)))(*[t.data_ptr() for t in inputs])
        out_p = (ctypes.c_void_p * len(outputs))(*[t.data_ptr() for t in outputs])
        regs_p = (ctypes.c_ubyte * len(out_regs))(*out_regs)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        name = "mab_%s_modprog%s" % (self.prime, "_jit" if jit else "")
        with torch.cuda.device(self.device):
            _lib.check(getattr(self.lib, name)(arr, len(code), ins_p, len(inputs), out_p, regs_p, len(outputs), n,
                                               max(stride, n), stream), name, self.lib)
        return outputs

    # -- conveniences for tests / small batches ------------------------------------------------
    def from_ints(self, values):
        """Python integers (< 2^(8*Nbytes)) -> planes via modimp."""
        import numpy as np
        raw = b"".join(int(v).to_bytes(self.Nbytes, "big") for v in values)
        b = torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).reshape(len(values), self.Nbytes).copy()).to(self.device)
        return self.modimp(b)[0]

    def to_ints(self, a):
        b = self.modexp(a).cpu().numpy()
        return [int.from_bytes(b[i].tobytes(), "big") for i in range(b.shape[0])]
