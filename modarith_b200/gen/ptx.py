"""Tiny PTX-level IR for the sm_100a field-code generators.

The reference generators print C text (pseudo.py:223-1174).  Ours print CUDA, and
the hot functions are single inline-PTX blocks built from 32-bit
`mad.lo.cc / madc.hi.cc / addc.cc / subc.cc` carry chains: ptxas fuses each
lo/hi pair on the same operands into one `IMAD.WIDE.U32[.X]` with the carry in a
predicate register, which is the densest way the INT32 multiplier pipe of
sm_100a can be driven.  Because there is no GPU where the generator runs, every
block is ALSO

  * interpreted here on Python integers (`Asm.run`) so the generator can check
    its own output against bignum arithmetic and trap any lost carry, and
  * printed as plain C (`emit_sim`) under `#ifdef MAB_HOSTSIM`, which the CPU
    test-suite compiles with g++ to exercise the hand-written device logic
    (ladder, import/export, fixed exponentiations) built on top.  That build is
    test scaffolding only; the shipped library never contains it.

Instruction set (all .u32): add sub madlo madhi mullo mulhi and or xor not shl
shr shfl shfr mov prmt, with optional carry-in / carry-out flags.
"""
from __future__ import annotations

import os

M32 = 0xFFFFFFFF
# How a carry capture (addc d, 0, 0) is emitted: "addc" leaves the choice to ptxas (SEL on the ALU pipe), "madc" writes
# madc.lo d, 0, 0, 0, which ptxas turns into IMAD.X on the multiplier pipe.  Multiplier-bound functions want the first,
# ALU-bound ones the second (profiles/r2_p256_capop.txt); a plan sets Asm.capture_op per function, MAB_CAPOP overrides
# the default of every function for experiments.
CAPTURE_OP = os.environ.get("MAB_CAPOP", "addc")


class LostCarry(Exception):
    pass


class Asm:
    def __init__(self, name=""):
        self.name = name
        self.ins = []          # (op, dst, srcs(tuple), cin, cout)
        self.ntmp = 0
        self.inputs = []       # external input names, in operand order
        self.outputs = []      # (external output name, internal reg)
        self.nocheck = set()   # instruction indices allowed to wrap (borrow-mask captures)
        self.zero_cin = set()  # instruction indices whose carry-in is an ordering link: must be 0
        self.capture_op = CAPTURE_OP

    # ---- registers ------------------------------------------------------
    def tmp(self, n=None):
        if n is None:
            r = "t%d" % self.ntmp
            self.ntmp += 1
            return r
        return [self.tmp() for _ in range(n)]

    def inp(self, *names):
        for n in names:
            if n not in self.inputs:
                self.inputs.append(n)
        return names[0] if len(names) == 1 else list(names)

    def out(self, ext, reg):
        self.outputs.append((ext, reg))

    # ---- instructions ---------------------------------------------------
    def _emit(self, op, dst, srcs, cin=False, cout=False):
        self.ins.append((op, dst, tuple(srcs), cin, cout))
        return dst

    def link_cc(self):
        """Ordering link for the NEXT carry-consuming instruction: the most recent carry-capable
        instruction (whose carry-out is provably zero and was not declared) is made to write CC, and
        the next instruction, emitted with cin=True, reads that zero.  Costs no instruction; it only
        gives ptxas a dependency, so that the consumer's chain cannot be started before the producer's
        chain has finished (ptxas otherwise runs many carry chains as a wavefront and runs out of
        predicate registers).  The interpreter checks that the carry really is zero."""
        for k in range(len(self.ins) - 1, -1, -1):
            op, d, srcs, cin, cout = self.ins[k]
            if op in ("add", "sub", "madlo", "madhi"):
                if not cout:
                    self.ins[k] = (op, d, srcs, cin, True)
                self.zero_cin.add(len(self.ins))
                return True
        return False

    def add(self, d, a, b, cin=False, cout=False):
        return self._emit("add", d, (a, b), cin, cout)

    def sub(self, d, a, b, cin=False, cout=False):
        return self._emit("sub", d, (a, b), cin, cout)

    def madlo(self, d, a, b, c, cin=False, cout=False):
        return self._emit("madlo", d, (a, b, c), cin, cout)

    def madhi(self, d, a, b, c, cin=False, cout=False):
        return self._emit("madhi", d, (a, b, c), cin, cout)

    def mullo(self, d, a, b):
        return self._emit("mullo", d, (a, b))

    def mulhi(self, d, a, b):
        return self._emit("mulhi", d, (a, b))

    def logic(self, op, d, a, b):
        assert op in ("and", "or", "xor")
        return self._emit(op, d, (a, b))

    def not_(self, d, a):
        return self._emit("not", d, (a,))

    def shl(self, d, a, n):
        return self._emit("shl", d, (a, n))

    def shr(self, d, a, n):
        return self._emit("shr", d, (a, n))

    def shfl(self, d, lo, hi, n):
        """d = high word of ((hi:lo) << n), 0 < n < 32  (shf.l.wrap.b32)."""
        return self._emit("shfl", d, (lo, hi, n))

    def shfr(self, d, lo, hi, n):
        """d = low word of ((hi:lo) >> n), 0 < n < 32  (shf.r.wrap.b32)."""
        return self._emit("shfr", d, (lo, hi, n))

    def mov(self, d, a):
        return self._emit("mov", d, (a,))

    def prmt(self, d, a, b, sel):
        return self._emit("prmt", d, (a, b, sel))

    # ---- carry-chain helpers -------------------------------------------
    def wide_chain(self, slots, last_carry_to=None, first=True, carry_in=False, keep_carry=False):
        """One carry chain of 32x32->64 multiply-accumulates.

        slots: list of (lo_dst, hi_dst, a, b, lo_addend, hi_addend); consecutive
        slots are adjacent 64-bit windows so the carry ripples upward.
        last_carry_to: (dst, addend) receiving the final carry, or None if the
        carry out of the last slot is provably zero (checked by the interpreter).
        carry_in: the first slot consumes the carry flag left by the preceding instruction.
        """
        n = len(slots)
        for k, (lo, hi, a, b, clo, chi) in enumerate(slots):
            self.madlo(lo, a, b, clo, cin=(k > 0 or carry_in), cout=True)
            last = (k == n - 1) and last_carry_to is None and not keep_carry      # keep_carry: the caller consumes CC
            self.madhi(hi, a, b, chi, cin=True, cout=not last)
        if last_carry_to is not None:
            d, c = last_carry_to
            self.add(d, c, 0, cin=True, cout=False)

    def add_chain(self, dsts, as_, bs, carry_to=None, carry_in=False, carry_out=False, wrap_ok=False):
        n = len(dsts)
        for k in range(n):
            last = k == n - 1
            if last and wrap_ok:
                self.nocheck.add(len(self.ins))
            self.add(dsts[k], as_[k], bs[k], cin=(k > 0 or carry_in),
                     cout=(not last) or carry_to is not None or carry_out)
        if carry_to is not None:
            d, c = carry_to
            self.add(d, c, 0, cin=True, cout=False)

    def sub_chain(self, dsts, as_, bs, borrow_to=None, wrap_ok=False):
        """borrow_to receives 0 or 0xFFFFFFFF (all-ones when the chain borrowed)."""
        n = len(dsts)
        for k in range(n):
            last = k == n - 1
            if last and wrap_ok:
                self.nocheck.add(len(self.ins))
            self.sub(dsts[k], as_[k], bs[k], cin=(k > 0), cout=(not last) or borrow_to is not None)
        if borrow_to is not None:
            self.nocheck.add(len(self.ins))
            self.sub(borrow_to, 0, 0, cin=True, cout=False)

    # ---- interpreter ----------------------------------------------------
    def run(self, env, strict=True):
        """Execute on Python ints.  env maps external input names -> values.
        With strict=True an instruction that produces a carry/borrow while not
        declaring carry-out raises LostCarry (chains are built so that every
        carry is either captured or provably zero)."""
        R = dict(env)
        cf = 0

        def val(x):
            if isinstance(x, int):
                return x & M32
            return R[x]

        for idx, (op, d, s, cin, cout) in enumerate(self.ins):
            c = cf if cin else 0
            if idx in self.zero_cin:
                assert cin, "ordering link without a consumer"
                if c and strict:
                    raise LostCarry("%s: ordering link at %d carries a non-zero carry" % (self.name, idx))
            if op == "add":
                t = val(s[0]) + val(s[1]) + c
                car = t >> 32
            elif op == "sub":
                t = val(s[0]) - val(s[1]) - c
                car = 1 if t < 0 else 0
            elif op == "madlo":
                t = ((val(s[0]) * val(s[1])) & M32) + val(s[2]) + c
                car = t >> 32
            elif op == "madhi":
                t = ((val(s[0]) * val(s[1])) >> 32) + val(s[2]) + c
                car = t >> 32
            else:
                car = None
                if op == "mullo":
                    t = val(s[0]) * val(s[1])
                elif op == "mulhi":
                    t = (val(s[0]) * val(s[1])) >> 32
                elif op == "and":
                    t = val(s[0]) & val(s[1])
                elif op == "or":
                    t = val(s[0]) | val(s[1])
                elif op == "xor":
                    t = val(s[0]) ^ val(s[1])
                elif op == "not":
                    t = ~val(s[0])
                elif op == "shl":
                    t = val(s[0]) << s[1]
                elif op == "shr":
                    t = val(s[0]) >> s[1]
                elif op == "shfl":
                    t = (((val(s[1]) << 32) | val(s[0])) << s[2]) >> 32
                elif op == "shfr":
                    t = ((val(s[1]) << 32) | val(s[0])) >> s[2]
                elif op == "mov":
                    t = val(s[0])
                elif op == "prmt":
                    bytes_ = [(val(s[0]) >> (8 * i)) & 0xFF for i in range(4)] + \
                             [(val(s[1]) >> (8 * i)) & 0xFF for i in range(4)]
                    t = 0
                    for i in range(4):
                        t |= bytes_[(s[2] >> (4 * i)) & 7] << (8 * i)
                else:
                    raise ValueError(op)
            if car is not None:
                if cout:
                    cf = car
                elif car and strict and idx not in self.nocheck:
                    raise LostCarry("%s: %s %s <- %s loses a carry" % (self.name, op, d, s))
            R[d] = t & M32
        return {ext: R[reg] if not isinstance(reg, int) else reg & M32 for ext, reg in self.outputs}

    # ---- statistics -----------------------------------------------------
    def stats(self):
        """Static instruction mix after ptxas fuses lo/hi pairs: (wide, imad32, alu)."""
        wide = imad = alu = 0
        i = 0
        ins = self.ins
        while i < len(ins):
            op, d, s, cin, cout = ins[i]
            if op in ("madlo", "mullo") and i + 1 < len(ins):
                op2, d2, s2 = ins[i + 1][0:3]
                if op2 == {"madlo": "madhi", "mullo": "mulhi"}[op] and s2[0:2] == s[0:2]:
                    wide += 1
                    i += 2
                    continue
            if op in ("madlo", "madhi", "mullo", "mulhi"):
                imad += 1
            elif op != "mov":
                alu += 1
            i += 1
        return wide, imad, alu

    # ---- emitters -------------------------------------------------------
    _PTX = {"add": "add", "sub": "sub", "madlo": "mad.lo", "madhi": "mad.hi"}

    def emit_ptx(self):
        """Return (asm_string_lines, output_exts, input_exts)."""
        outs = [ext for ext, _ in self.outputs]
        nout = len(outs)
        idx = {n: nout + i for i, n in enumerate(self.inputs)}

        def o(x):
            if isinstance(x, int):
                return "0x%x" % (x & M32)
            if x in idx:
                return "%%%d" % idx[x]
            return x

        lines = []
        if self.ntmp:
            lines.append(".reg .u32 t<%d>;" % self.ntmp)
        for (op, d, s, cin, cout) in self.ins:
            if op in ("add", "sub"):
                m = ("addc" if cin else "add") if op == "add" else ("subc" if cin else "sub")
                m += ".cc.u32" if cout else ".u32"
                if self.capture_op == "madc" and op == "add" and cin and not cout and s[0] == 0 and s[1] == 0:
                    # carry capture on the multiplier pipe (IMAD.X) instead of the ALU pipe (SEL): pays where a
                    # function is ALU-bound
                    lines.append("madc.lo.u32 %s, 0, 0, 0;" % o(d))
                    continue
                lines.append("%s %s, %s, %s;" % (m, o(d), o(s[0]), o(s[1])))
            elif op in ("madlo", "madhi"):
                m = ("madc" if cin else "mad") + (".lo" if op == "madlo" else ".hi")
                m += ".cc.u32" if cout else ".u32"
                lines.append("%s %s, %s, %s, %s;" % (m, o(d), o(s[0]), o(s[1]), o(s[2])))
            elif op == "mullo":
                lines.append("mul.lo.u32 %s, %s, %s;" % (o(d), o(s[0]), o(s[1])))
            elif op == "mulhi":
                lines.append("mul.hi.u32 %s, %s, %s;" % (o(d), o(s[0]), o(s[1])))
            elif op in ("and", "or", "xor"):
                lines.append("%s.b32 %s, %s, %s;" % (op, o(d), o(s[0]), o(s[1])))
            elif op == "not":
                lines.append("not.b32 %s, %s;" % (o(d), o(s[0])))
            elif op in ("shl", "shr"):
                lines.append("%s.b32 %s, %s, %d;" % (op, o(d), o(s[0]), s[1])
                             if op == "shl" else "shr.u32 %s, %s, %d;" % (o(d), o(s[0]), s[1]))
            elif op == "shfl":
                lines.append("shf.l.wrap.b32 %s, %s, %s, %d;" % (o(d), o(s[0]), o(s[1]), s[2]))
            elif op == "shfr":
                lines.append("shf.r.wrap.b32 %s, %s, %s, %d;" % (o(d), o(s[0]), o(s[1]), s[2]))
            elif op == "mov":
                lines.append("mov.u32 %s, %s;" % (o(d), o(s[0])))
            elif op == "prmt":
                lines.append("prmt.b32 %s, %s, %s, 0x%x;" % (o(d), o(s[0]), o(s[1]), s[2]))
            else:
                raise ValueError(op)
        for k, (ext, reg) in enumerate(self.outputs):
            lines.append("mov.u32 %%%d, %s;" % (k, o(reg)))
        return lines, outs, list(self.inputs)

    def emit_cuda(self, indent="    "):
        lines, outs, ins = self.emit_ptx()
        body = "".join('%s    "%s\\n\\t"\n' % (indent, ln) for ln in lines)
        s = '%sasm("{\\n\\t"\n%s%s    "}"\n' % (indent, body, indent)
        s += "%s    : %s\n" % (indent, ", ".join('"=r"(%s)' % e for e in outs))
        s += "%s    : %s);\n" % (indent, ", ".join('"r"(%s)' % e for e in ins))
        return s

    def emit_sim(self, indent="    "):
        """Plain C with the same semantics (host simulation build)."""
        def o(x):
            if isinstance(x, int):
                return "0x%xu" % (x & M32)
            return x.replace("[", "_").replace("]", "") + "_i" if x in self.inputs else x

        L = []
        for n in self.inputs:
            L.append("const uint32_t %s = %s;" % (o(n), n))
        if self.ntmp:
            L.append("uint32_t " + ", ".join("t%d" % i for i in range(self.ntmp)) + ";")
        L.append("uint64_t w_; uint32_t cf_ = 0; (void)cf_; (void)w_;")
        for (op, d, s, cin, cout) in self.ins:
            c = " + cf_" if cin else ""
            if op == "add":
                e = "(uint64_t)%s + %s%s" % (o(s[0]), o(s[1]), c)
                car = "(uint32_t)(w_ >> 32)"
            elif op == "sub":
                e = "(uint64_t)%s - %s%s" % (o(s[0]), o(s[1]), " - cf_" if cin else "")
                car = "(uint32_t)(w_ >> 63)"
            elif op == "madlo":
                e = "(uint64_t)(uint32_t)(%s * %s) + %s%s" % (o(s[0]), o(s[1]), o(s[2]), c)
                car = "(uint32_t)(w_ >> 32)"
            elif op == "madhi":
                e = "(((uint64_t)%s * %s) >> 32) + %s%s" % (o(s[0]), o(s[1]), o(s[2]), c)
                car = "(uint32_t)(w_ >> 32)"
            else:
                car = None
                if op == "mullo":
                    e = "(uint32_t)(%s * %s)" % (o(s[0]), o(s[1]))
                elif op == "mulhi":
                    e = "((uint64_t)%s * %s) >> 32" % (o(s[0]), o(s[1]))
                elif op in ("and", "or", "xor"):
                    e = "%s %s %s" % (o(s[0]), {"and": "&", "or": "|", "xor": "^"}[op], o(s[1]))
                elif op == "not":
                    e = "~%s" % o(s[0])
                elif op == "shl":
                    e = "%s << %d" % (o(s[0]), s[1])
                elif op == "shr":
                    e = "%s >> %d" % (o(s[0]), s[1])
                elif op == "shfl":
                    e = "((((uint64_t)%s << 32) | %s) << %d) >> 32" % (o(s[1]), o(s[0]), s[2])
                elif op == "shfr":
                    e = "(((uint64_t)%s << 32) | %s) >> %d" % (o(s[1]), o(s[0]), s[2])
                elif op == "mov":
                    e = o(s[0])
                elif op == "prmt":
                    e = "mab_sim_prmt(%s, %s, 0x%xu)" % (o(s[0]), o(s[1]), s[2])
                else:
                    raise ValueError(op)
            if car is None:
                L.append("%s = (uint32_t)(%s);" % (o(d), e))
            else:
                L.append("w_ = %s; %s = (uint32_t)w_;%s" % (e, o(d), " cf_ = %s;" % car if cout else ""))
        for ext, reg in self.outputs:
            L.append("%s = %s;" % (ext, o(reg)))
        return "".join(indent + ln + "\n" for ln in L)
