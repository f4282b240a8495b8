"""Limb plans: how one modulus maps onto 32-bit registers of an sm_100a thread.

The reference picks an UNSATURATED radix (getbase, pseudo.py:124-140: 29-bit limbs
x 9 for 255/256-bit moduli, 28 x 16 for 448 bits at WL=32) because portable C has
no carry flag.  sm_100a has one (`IMAD.WIDE.U32.X`, `IADD3.X` carry predicates),
and the roofline is counted in L^2 products with L = ceil(Nbits/32) (SURVEY.md
section 8d), so the plans here use L saturated 32-bit limbs -- 8 for 2^255-19 and
P-256, 14 for 2^448-2^224-1 -- which is 21 % / 23 % fewer wide multiplies than the
reference's WL=32 layout (64 vs 81, 196 vs 256 per modmul).  The price is that
there is no spare room for lazy "< 2p" results: every function returns a value
that fits the L words, and each family keeps its own invariant:

  PseudoMersenne  2^n - c, 32L - n spare bits (X25519: 1).  Weakly reduced:
                  any representative < 2^(32L).  2^(32L) == c*2^(32L-n) =: fold.
  GenMersenne     2^(2h) - 2^h - 1 with h a multiple of 32 (X448, h=224).  Weakly
                  reduced < 2^(2h); 2^(2h) == 2^h + 1, so reduction is additions of
                  half-length pieces: no multiplies (cf. monty.py:597-627 turning
                  shaped limbs into adds; we apply the shape directly instead of
                  through Montgomery form).
  Montgomery      p == -1 mod 2^32 (P-256): R = 2^(32L), ndash = 1 ("Montgomery
                  friendly", monty.py:740-751), values kept fully reduced in [0,p).

Each plan builds ptx.Asm blocks for mul/sqr/mli/add/sub/canon and checks them
against Python bignum arithmetic at generation time (`self_check`).
"""
from __future__ import annotations

import os
import random

from ..primes import Prime
from .ptx import Asm
from . import satmul

M32 = 0xFFFFFFFF


def words(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def value(ws):
    return sum(w << (32 * i) for i, w in enumerate(ws))


def _names(prefix, n):
    return ["%s[%d]" % (prefix, i) for i in range(n)]


class Plan:
    # reductions that consume the low half of the product first (Montgomery) order the merge after
    # the last product chain, see Asm.link_cc; MAB_LINK=0 switches it off (experiment)
    link_products = False
    family = "?"
    tight = None            # exclusive bound of mul/sqr outputs when it is below 2^(32L) (see PseudoMersenne)
    capture_ops = {}        # function -> "madc" where its carry captures belong on the multiplier pipe (see build)

    def __init__(self, prime: Prime):
        self.P = prime
        self.p = prime.p
        self.name = prime.name
        self.L = (prime.nbits + 31) // 32
        self.bound = 1 << (32 * self.L)      # exclusive bound on stored values (overridden)
        self.R = 1                           # Montgomery factor (1 = plain residues)
        # launch bound of the ladder kernel.  8 limbs: three resident CTAs per SM -- ptxas then takes the 146 registers it
        # wants for the round-structured kernel, and the same step loop (identical opcode mix) runs 1-7 % faster than the
        # 128-register allocation four CTAs force, at every batch size (profiles/r2_x25519_occupancy.txt); more limbs: two
        self.ladder_minblocks = int(os.environ.get("MAB_MINBLOCKS_" + prime.name, 3 if self.L <= 8 else 2))
        self.ladder_stash = os.environ.get("MAB_STASH_" + prime.name, "1" if self.L > 8 else "0") == "1"

    # -- representation -----------------------------------------------------
    def to_internal(self, v):                # field value -> stored integer
        return v * self.R % self.p

    def from_internal(self, w):              # stored integer -> field value
        return w * pow(self.R, -1, self.p) % self.p

    # -- blocks every plan provides ------------------------------------------
    def build(self):
        self.blocks = {
            "mul": self.build_mul(),
            "sqr": self.build_sqr(),
            "mli": self.build_mli(),
            "mla": self.build_mla(),
            "add": self.build_add(),
            "sub": self.build_sub(neg=False),
            "neg": self.build_sub(neg=True),
            "canon": self.build_canon(),
        }
        if self.tight:
            self.blocks["add_tt"] = self.build_add_tt()
            self.blocks["sub_tt"] = self.build_sub_tt()
        if getattr(self, "weak_bound", None):
            self.blocks["mul_w"] = self.build_mul_w()
            self.blocks["sqr_w"] = self.build_sqr_w()
        # which pipe a function's carry captures run on (ptx.py: "addc" = ptxas' choice, SEL on the ALU pipe; "madc" =
        # IMAD.X on the multiplier pipe): the plan's table, then MAB_CAPOP_<PRIME>_<FUNCTION>=madc|addc for experiments
        for fn, asm in self.blocks.items():
            v = os.environ.get("MAB_CAPOP_%s_%s" % (self.name.upper(), fn.upper()), self.capture_ops.get(fn))
            if v:
                asm.capture_op = v
        return self.blocks

    def _io(self, asm, ins, out="r"):
        regs = []
        for nm in ins:
            ws = _names(nm, self.L)
            asm.inp(*ws)
            regs.append(ws)
        return regs

    def _outs(self, asm, regs, out="r"):
        for k, r in enumerate(regs):
            asm.out("%s[%d]" % (out, k), r)

    def build_mul(self):
        asm = Asm(self.name + ".mul")
        a, b = self._io(asm, ["a", "b"])
        T = satmul.product(asm, a, b, link=self.link_products)
        self._outs(asm, self.reduce_wide(asm, T))
        return asm

    def build_sqr(self):
        asm = Asm(self.name + ".sqr")
        (a,) = self._io(asm, ["a"])
        T = satmul.square(asm, a, link=self.link_products)
        self._outs(asm, self.reduce_wide(asm, T))
        return asm

    def build_mli(self):
        asm = Asm(self.name + ".mli")
        (a,) = self._io(asm, ["a"])
        asm.inp("b")
        T = satmul.times_small(asm, a, "b")
        self._outs(asm, self.reduce_small(asm, T))
        return asm

    def build_mla(self):
        """r = a*b + c for a small integer b: modmli fused with the modadd that follows it in the
        ladder step (rfc7748.c:209,212: z2 = a24*E + AA)."""
        asm = Asm(self.name + ".mla")
        a, c = self._io(asm, ["a", "c"])
        asm.inp("b")
        T = satmul.times_small_add(asm, a, "b", c)
        self._outs(asm, self.reduce_small(asm, T))
        return asm

    # -- generation-time verification ---------------------------------------
    def sample_values(self, rng, n):
        """Stored-integer samples: edges of the representation plus random ones."""
        p, B = self.p, self.bound
        edge = [0, 1, 2, p - 1, p - 2, B - 1, B - 2, (B - 1) ^ 1, 19, 38, 1 << 32, (1 << 32) - 1]
        edge += [p, p + 1, 2 * p - 1, 2 * p, 2 * p + 1]
        edge += [value([M32] * k + [0] * (self.L - k)) for k in range(1, self.L)]
        edge += [value([0] * k + [M32] * (self.L - k)) for k in range(1, self.L)]
        edge = [e for e in edge if 0 <= e < B]
        out = list(edge)
        while len(out) < n:
            r = rng.random()
            if r < 0.6:
                out.append(rng.randrange(B))
            elif r < 0.8:
                out.append(value([rng.choice([0, M32, 1, 0x80000000, rng.getrandbits(32)])
                                  for _ in range(self.L)]) % B)
            else:
                out.append((B - 1 - rng.getrandbits(rng.randrange(1, 40))) % B)
        return out

    def run(self, block, **ins):
        env = {}
        for k, v in ins.items():
            if isinstance(v, list):
                for i, w in enumerate(v):
                    env["%s[%d]" % (k, i)] = w
            else:
                env[k] = v & M32
        out = self.blocks[block].run(env)
        r = value([out["r[%d]" % i] for i in range(self.L)])
        extra = {k: v for k, v in out.items() if not k.startswith("r[")}
        return r, extra

    def self_check(self, trials=400, seed=1):
        rng = random.Random(seed)
        p, L, B = self.p, self.L, self.bound
        Rinv = pow(self.R, -1, p)
        xs = self.sample_values(rng, trials)
        ys = self.sample_values(rng, trials)
        rng.shuffle(ys)
        for x, y in zip(xs, ys):
            ax, ay = words(x, L), words(y, L)
            r, _ = self.run("mul", a=ax, b=ay)
            assert r < B and r % p == x * y * Rinv % p, ("mul", hex(x), hex(y))
            r, _ = self.run("sqr", a=ax)
            assert r < B and r % p == x * x * Rinv % p, ("sqr", hex(x))
            for bb in (0, 1, 2, 121665, 39081, M32, y & M32):
                r, _ = self.run("mli", a=ax, b=bb)
                assert r < B and r % p == x * bb % p, ("mli", hex(x), bb)
            for bb in (0, 1, 121665, 39081, (1 << 31) - 1):
                r, _ = self.run("mla", a=ax, c=ay, b=bb)
                assert r < B and r % p == (x * bb + y) % p, ("mla", hex(x), hex(y), bb)
            r, _ = self.run("add", a=ax, b=ay)
            assert r < B and r % p == (x + y) % p, ("add", hex(x), hex(y))
            r, _ = self.run("sub", a=ax, b=ay)
            assert r < B and r % p == (x - y) % p, ("sub", hex(x), hex(y))
            r, _ = self.run("neg", b=ax)
            assert r < B and r % p == (-x) % p, ("neg", hex(x))
            r, e = self.run("canon", a=ax)
            assert r == x % p and e["lt"] == (1 if x < p else 0), ("canon", hex(x))
        # raw imports (modimp / mod2r hand over any value below 2^(32L), Field::import_raw): plain-residue plans
        # canonicalise it, Montgomery plans multiply the raw words by R^2 and rely on the product's own reduction
        top = 1 << (32 * L)
        raw = [top - 1, top - 2, p, p + 1, 2 * p - 1, 2 * p, 2 * p + 1, 3 * p, top >> 1, (top >> 1) + 1]
        raw = [v for v in raw if 0 <= v < top] + [rng.randrange(top) for _ in range(40)]
        for x in raw:
            ax = words(x, L)
            r, e = self.run("canon", a=ax)
            assert e["lt"] == (1 if x < p else 0), ("canon flag", hex(x))
            if self.R == 1:
                assert r == x % p, ("canon over the raw range", hex(x))
            else:
                r, _ = self.run("mul", a=ax, b=words(self.R * self.R % p, L))
                assert r == x * self.R % p, ("raw import through R^2", hex(x))
        # every pair of edge values through mul / add / sub (the shuffled pairing above does not guarantee that
        # the extremes meet: B-1 times B-1 is what stresses the last wrap of a fold)
        edge = [v for v in xs[:60] if v in (0, 1, 2, p - 1, p, p + 1, B - 1, B - 2, 2 * p - 1) or v >= B - (1 << 40)]
        edge = sorted(set(edge + [B - 1, p - 1, 0, 1]))[:14]
        for x in edge:
            for y in edge:
                if x >= B or y >= B:
                    continue
                ax, ay = words(x, L), words(y, L)
                r, _ = self.run("mul", a=ax, b=ay)
                assert r < B and r % p == x * y * Rinv % p, ("mul", hex(x), hex(y))
                r, _ = self.run("add", a=ax, b=ay)
                assert r < B and r % p == (x + y) % p, ("add", hex(x), hex(y))
                r, _ = self.run("sub", a=ax, b=ay)
                assert r < B and r % p == (x - y) % p, ("sub", hex(x), hex(y))
                r, _ = self.run("mla", a=ax, c=ay, b=M32 >> 1)
                assert r < B and r % p == (x * (M32 >> 1) + y) % p, ("mla", hex(x), hex(y))
        if "mul_w" in self.blocks:
            Wb = self.weak_bound
            wedge = [0, 1, p - 1, p, p + 1, Wb - 1, Wb - 2, Wb >> 1, (Wb >> 1) + 1, Wb - (1 << 32), Wb - (1 << 224)]
            ws = [v for v in wedge if 0 <= v < Wb]
            pairs = [(a, b) for a in ws for b in ws] + [(rng.randrange(Wb), rng.randrange(Wb)) for _ in range(trials)]
            for x, y in pairs:
                r, _ = self.run("mul_w", a=words(x, L), b=words(y, L))
                assert r < Wb and r % p == x * y * Rinv % p, ("mul_w", hex(x), hex(y))
                r, _ = self.run("sqr_w", a=words(x, L))
                assert r < Wb and r % p == x * x * Rinv % p, ("sqr_w", hex(x))
                r2, _ = self.run("canon", a=words(r, L))
                assert r2 == x * x * Rinv % p, ("canon after sqr_w", hex(x))
        if self.tight:
            # mul/sqr land below the tight bound for ANY stored operands; add_tt/sub_tt are exact on tight operands
            T = self.tight
            for x, y in zip(xs, ys):
                r, _ = self.run("mul", a=words(x, L), b=words(y, L))
                assert r < T, ("mul not tight", hex(x), hex(y))
                r, _ = self.run("sqr", a=words(x, L))
                assert r < T, ("sqr not tight", hex(x))
            edge = [0, 1, 2, 37, 38, 39, p - 1, p, p + 1, T - 1, T - 2, T - 38, T - 39, (1 << (32 * L - 1)) - 1,
                    1 << (32 * L - 1), (1 << (32 * L - 1)) + 1, M32, 1 << 32, (1 << 32) - 38]
            ts = [e for e in edge if 0 <= e < T]
            pairs = [(a, b) for a in ts for b in ts]
            while len(pairs) < len(ts) ** 2 + trials:
                pairs.append((rng.randrange(T), rng.randrange(T)))
            for x, y in pairs:
                r, _ = self.run("add_tt", a=words(x, L), b=words(y, L))
                assert r < B and r % p == (x + y) % p, ("add_tt", hex(x), hex(y))
                r, _ = self.run("sub_tt", a=words(x, L), b=words(y, L))
                assert r < B and r % p == (x - y) % p, ("sub_tt", hex(x), hex(y))
        return True

    # -- shared small pieces ------------------------------------------------
    def _cond_sub_p(self, asm, v, want_flag=None):
        """r = v - p if v >= p else v.  Returns regs; optionally writes (v < p) flag."""
        L = self.L
        pw = words(self.p, L)
        t = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(t, v, pw, borrow_to=m)          # m = all-ones iff v < p
        r = []
        for k in range(L):
            x, y, z = asm.tmp(), asm.tmp(), asm.tmp()
            asm.logic("xor", x, t[k], v[k])
            asm.logic("and", y, x, m)
            asm.logic("xor", z, y, t[k])              # m ? v : t   (one LOP3 after ptxas)
            r.append(z)
        if want_flag is not None:
            asm.logic("and", want_flag, m, 1)
        return r


class PseudoMersenne(Plan):
    """2^n - c on L saturated limbs; stored values are any representative < 2^(32L)."""
    family = "pseudo"

    def __init__(self, prime):
        super().__init__(prime)
        n = prime.nbits
        self.c = (1 << n) - prime.p
        self.xs = 32 * self.L - n
        self.fold = self.c << self.xs            # 2^(32L) == fold (mod p); cf. mm=m*2^xcess, pseudo.py:1594-1596
        assert self.fold * (self.fold + 2) < (1 << 32), "fold constant too large for this plan"
        assert self.bound <= 2 * self.p + self.fold + 1
        # A spare bit above Nbits (2^255-19 in 8 words) lets products be folded at bit Nbits instead of at
        # bit 32L for the same instruction count (_fold_top): they then stay below 2^Nbits + c*2^13, and
        # sums / differences of two such values need no second wrap (add_tt 10 instructions instead of
        # 20, sub_tt 18 instead of 21).  MAB_TIGHT=0 switches it off (experiment).
        if self.xs >= 1 and os.environ.get("MAB_TIGHT", "1") != "0" and os.environ.get("MAB_FOLD", "split") == "split":
            self.tight = (1 << n) + self.c * (1 << 13)
            assert 2 * self.tight < (1 << (32 * self.L)) + (1 << 32) - self.fold      # add_tt: wrapped sum + fold fits word 0

    def _fold_carry(self, asm, r, c, wide=False):
        """r (L words) += c * fold, where c*2^(32L) was dropped.  c is a register.
        wide=False needs c*fold < 2^32."""
        L = self.L
        o = asm.tmp(L)
        c2 = asm.tmp()
        if wide:
            asm.madlo(o[0], c, self.fold, r[0], cout=True)
            asm.madhi(o[1], c, self.fold, r[1], cin=True, cout=True)
            start = 2
        else:
            asm.madlo(o[0], c, self.fold, r[0], cout=True)
            start = 1
        for k in range(start, L):
            asm.add(o[k], r[k], 0, cin=True, cout=True)
        asm.add(c2, 0, 0, cin=True)
        # a second wrap leaves a tiny value (< c*fold), so this cannot carry past word 0
        # (word 1 when c is a full word); the interpreter checks
        f = asm.tmp()
        if wide:
            g = asm.tmp()
            asm.madlo(f, c2, self.fold, o[0], cout=True)
            asm.add(g, o[1], 0, cin=True)
            return [f, g] + o[2:]
        asm.madlo(f, c2, self.fold, o[0])
        return [f] + o[1:]

    def _fold_top(self, asm, r, top):
        """r (L words) + top * 2^(32L)  ->  L words below 2^Nbits + c * 2^13: everything at or above bit
        Nbits (`top` and the xs spare bits of the last word) is multiplied by c and added at the bottom.
        Same cost as _fold_carry, and no second wrap can occur."""
        L, xs = self.L, self.xs
        h = asm.tmp()
        asm.shfl(h, r[L - 1], top, xs)                    # (top : r[L-1]) << xs, high word = bits >= Nbits
        last = asm.tmp()
        asm.logic("and", last, r[L - 1], M32 >> xs)
        o = asm.tmp(L)
        asm.madlo(o[0], h, self.c, r[0], cout=True)
        for k in range(1, L - 1):
            asm.add(o[k], r[k], 0, cin=True, cout=True)
        asm.add(o[L - 1], last, 0, cin=True)
        return o

    def build_add_tt(self):
        """a + b for a, b below the tight bound: the sum wraps past 2^(32L) at most once, and when it does
        the wrapped value is so small that adding the fold constant cannot carry out of word 0."""
        asm = Asm(self.name + ".add_tt")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        s = asm.tmp(L)
        c = asm.tmp()
        asm.add_chain(s, a, b, carry_to=(c, 0))
        f = asm.tmp()
        asm.madlo(f, c, self.fold, s[0])
        self._outs(asm, [f] + s[1:])
        return asm

    def build_sub_tt(self):
        """a - b for a, b below the tight bound: after a borrow the difference is at least
        2^(32L) - tight > fold, so taking the fold constant off cannot borrow a second time."""
        asm = Asm(self.name + ".sub_tt")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)
        f = asm.tmp()
        asm.logic("and", f, m, self.fold)
        e = asm.tmp(L)
        asm.sub_chain(e, d, [f] + [0] * (L - 1))
        self._outs(asm, e)
        return asm

    def reduce_wide(self, asm, T):
        """2L words -> L words: lo + fold*hi (second_pass of pseudo.py:557-611 restated for a
        saturated radix).  MAB_FOLD=imad: two even/odd wide multiply chains on the multiplier
        pipe; MAB_FOLD=alu: shift-and-add on the ALU pipe (fold = 38 = 2^5+2^2+2^1), which
        frees 8 IMAD.WIDE slots per reduction at the price of ~54 ALU instructions."""
        L = self.L
        lo, hi = T[:L], T[L:]
        if os.environ.get("MAB_FOLD", "imad") == "alu":
            acc = list(lo) + [0]
            for sh in [k for k in range(31, -1, -1) if (self.fold >> k) & 1]:
                S = []
                for k in range(L + 1):
                    d = asm.tmp()
                    if k == 0:
                        asm.shl(d, hi[0], sh) if sh else asm.mov(d, hi[0])
                    elif k == L:
                        asm.shr(d, hi[L - 1], 32 - sh) if sh else asm.mov(d, 0)
                    else:
                        asm.shfl(d, hi[k - 1], hi[k], sh) if sh else asm.mov(d, hi[k])
                    S.append(d)
                new = asm.tmp(L + 1)
                asm.add_chain(new, acc, S)
                acc = new
            return self._fold_carry(asm, acc[:L], acc[L])
        r = asm.tmp(L + 1)
        cur = list(lo) + [0]
        # even windows (0,1),(2,3).. take hi[0],hi[2]..; odd windows (1,2).. take hi[1],hi[3]..
        ev = [(r[k], r[k + 1], hi[k], self.fold, cur[k], cur[k + 1]) for k in range(0, L, 2)]
        asm.wide_chain(ev, last_carry_to=(r[L], 0) if L % 2 == 0 else None)
        cur = list(r)
        r2 = asm.tmp(L + 1)
        od = [(r2[k], r2[k + 1], hi[k], self.fold, cur[k], cur[k + 1]) for k in range(1, L, 2)]
        asm.wide_chain(od, last_carry_to=None)
        res = [cur[0]] + r2[1:L + 1] if L % 2 == 0 else None
        assert res is not None, "odd limb counts not needed yet"
        return self._fold_carry(asm, res[:L], res[L])

    def reduce_small(self, asm, T):
        """L+1 words (top word < 2^32) -> L words."""
        return self._fold_carry(asm, T[:self.L], T[self.L], wide=True)

    # -- split fold: keeps every 64-bit window on the register pair it was accumulated in --------
    def _fold_split(self, asm, lo_even, hi, lo_odd=None):
        """lo_even[0..L-1] + 2^32*lo_odd + fold * hi[0..L-1]  ->  L words.

        Even words of `hi` fold onto the even-aligned windows (0,1),(2,3).. of lo_even with one
        wide carry chain; odd words fold onto the odd-aligned windows (1,2),(3,4).. of lo_odd
        (the product's ODD accumulator array, or fresh registers), so no window changes its
        register pairing (ptxas otherwise re-pairs with ~8 MOVs per reduction); the two results
        are merged by one add-with-carry chain."""
        L, f = self.L, self.fold
        assert L % 2 == 0
        r = asm.tmp(L)
        ce = asm.tmp()
        ev = [(r[k], r[k + 1], hi[k], f, lo_even[k], lo_even[k + 1]) for k in range(0, L, 2)]
        asm.wide_chain(ev, last_carry_to=(ce, 0))
        o = {k: asm.tmp() for k in range(1, L + 1)}
        if lo_odd is None:
            for k in range(1, L, 2):
                asm.mullo(o[k], hi[k], f)
                asm.mulhi(o[k + 1], hi[k], f)
        else:
            od = [(o[k], o[k + 1], hi[k], f, lo_odd[k], lo_odd[k + 1] if k + 1 < L else 0) for k in range(1, L, 2)]
            asm.wide_chain(od, last_carry_to=None)
        res = [r[0]] + asm.tmp(L - 1)
        top = asm.tmp()
        for k in range(1, L):
            asm.add(res[k], r[k], o[k], cin=(k > 1), cout=True)
        asm.add(top, ce, o[L], cin=True, cout=False)
        return self._fold_top(asm, res, top) if self.tight else self._fold_carry(asm, res, top)

    def build_mul(self):
        if os.environ.get("MAB_FOLD", "split") != "split":
            return super().build_mul()
        asm = Asm(self.name + ".mul")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        E, O = satmul.product_eo(asm, a, b)
        # U = high halves merged (the low halves stay apart until after the fold)
        U = asm.tmp(L)
        for k in range(L):
            asm.add(U[k], E.src(L + k), O.src(L + k), cin=(k > 0), cout=(k < L - 1))
        lo_even = [E.src(k) for k in range(L)]
        lo_odd = {k: O.src(k) for k in range(1, L)}
        self._outs(asm, self._fold_split(asm, lo_even, U, lo_odd))
        return asm

    def build_sqr(self):
        if os.environ.get("MAB_FOLD", "split") != "split":
            return super().build_sqr()
        asm = Asm(self.name + ".sqr")
        (a,) = self._io(asm, ["a"])
        L = self.L
        T = satmul.square(asm, a)
        self._outs(asm, self._fold_split(asm, T[:L], T[L:], None))
        return asm

    def build_add(self):
        asm = Asm(self.name + ".add")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        s = asm.tmp(L)
        c = asm.tmp()
        asm.add_chain(s, a, b, carry_to=(c, 0))
        self._outs(asm, self._fold_carry(asm, s, c))
        return asm

    def build_sub(self, neg):
        asm = Asm(self.name + (".neg" if neg else ".sub"))
        L = self.L
        if neg:
            (b,) = self._io(asm, ["b"])
            a = [0] * L
        else:
            a, b = self._io(asm, ["a", "b"])
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)           # value = d - 2^(32L)*[m]  == d - fold*[m]
        f = asm.tmp()
        asm.logic("and", f, m, self.fold)
        e = asm.tmp(L)
        m2 = asm.tmp()
        asm.sub_chain(e, d, [f] + [0] * (L - 1), borrow_to=m2)
        f2 = asm.tmp()
        asm.logic("and", f2, m2, self.fold)
        g = asm.tmp()
        asm.sub(g, e[0], f2)                          # cannot borrow: e[0] >= 2^32 - fold here
        self._outs(asm, [g] + e[1:])
        return asm

    def build_canon(self):
        """Weak representative (< 2^(32L) <= 2p+fold+1) -> canonical residue; also
        'lt' = 1 iff the input was already < p (modfsb's return, pseudo.py:272-283)."""
        asm = Asm(self.name + ".canon")
        (a,) = self._io(asm, ["a"])
        lt = asm.tmp()
        r = self._cond_sub_p(asm, a, want_flag=lt)
        r = self._cond_sub_p(asm, r)
        if self.bound > 3 * self.p:
            r = self._cond_sub_p(asm, r)
        self._outs(asm, r)
        asm.out("lt", lt)
        return asm


class PseudoMersenne33(Plan):
    """2^(32L) - c with c = 2^32 + cl, cl < 2^15 (secp256k1's field: c = 2^32 + 977), on L saturated limbs;
    stored values are any representative < 2^(32L) (< 2p), plain residues (R = 1).

    The reference builds this modulus through monty.py (Montgomery form, monty.py:2066-2067: "not exploitable"
    for its unsaturated radix); with saturated limbs the shape is: 2^(32L) == c, so a double-length product
    folds as lo + cl*hi + (hi << 32) -- L wide multiplies by the small constant on the even/odd accumulator
    windows (as in PseudoMersenne._fold_split) plus one word-shifted add chain -- then the word and a bit that
    stick out above 2^(32L) fold once more (one wide multiply), and a last wrap can only add c to a value that
    is tiny by then.  L^2 + L + 1 wide multiplies per modmul against 2 L^2 + L for the word-serial Montgomery
    fall-back this modulus used before."""
    family = "monty"

    def __init__(self, prime):
        super().__init__(prime)
        n, L = prime.nbits, self.L
        assert n == 32 * L and L % 2 == 0
        self.c = (1 << n) - prime.p
        self.cl = self.c - (1 << 32)
        assert 0 < self.cl < (1 << 15), "modulus is not 2^(32L) - 2^32 - small"
        assert self.bound <= 2 * self.p

    # -- folding what sticks out above 2^(32L) ----------------------------------------------------------
    def _fold_word(self, asm, r, t0, t1=None):
        """r (L words) + (t0 + 2^32*t1) * 2^(32L) -> L words, for a 32-bit word t0 and an optional bit t1:
        t*c = t*cl + (t << 32) is at most three words; one add chain, then a wrap can only happen onto a value
        below 2^67, where adding c once more stays inside three words."""
        L, cl = self.L, self.cl
        p0, p1 = asm.tmp(), asm.tmp()
        asm.mullo(p0, t0, cl)
        asm.mulhi(p1, t0, cl)                          # < cl
        if t1 is not None:
            q1 = asm.tmp()
            asm.madlo(q1, t1, cl, p1)                  # + t1*cl*2^32, cannot overflow (p1 < cl, t1 <= 1)
            p1 = q1
        v1, v2 = asm.tmp(), asm.tmp()
        asm.add(v1, p1, t0, cout=True)                 # + (t << 32): t0 at word 1, t1 at word 2
        asm.add(v2, t1 if t1 is not None else 0, 0, cin=True)
        o = asm.tmp(L)
        c2 = asm.tmp()
        asm.add_chain(o, r, [p0, v1, v2] + [0] * (L - 3), carry_to=(c2, 0))
        f0 = asm.tmp()
        asm.madlo(f0, c2, cl, o[0], cout=True)         # wrapped: o < 2^67, so o + c fits in three words
        f1, f2 = asm.tmp(), asm.tmp()
        asm.add(f1, o[1], c2, cin=True, cout=True)
        asm.add(f2, o[2], 0, cin=True)
        return [f0, f1, f2] + o[3:]

    def _fold_double(self, asm, lo_even, hi, lo_odd=None):
        """lo_even + 2^32*lo_odd + c*hi -> L words (hi: L words)."""
        L, cl = self.L, self.cl
        r = asm.tmp(L)
        ce = asm.tmp()
        ev = [(r[k], r[k + 1], hi[k], cl, lo_even[k], lo_even[k + 1]) for k in range(0, L, 2)]
        asm.wide_chain(ev, last_carry_to=(ce, 0))
        o = {k: asm.tmp() for k in range(1, L + 1)}
        if lo_odd is None:
            for k in range(1, L, 2):
                asm.mullo(o[k], hi[k], cl)
                asm.mulhi(o[k + 1], hi[k], cl)
        else:
            od = [(o[k], o[k + 1], hi[k], cl, lo_odd[k], lo_odd[k + 1] if k + 1 < L else 0) for k in range(1, L, 2)]
            asm.wide_chain(od, last_carry_to=None)
        res = [r[0]] + asm.tmp(L - 1)
        top = asm.tmp()
        for k in range(1, L):
            asm.add(res[k], r[k], o[k], cin=(k > 1), cout=True)
        asm.add(top, ce, o[L], cin=True, cout=False)               # lo + cl*hi = res + top*2^(32L), top <= cl
        # + (hi << 32): hi[k-1] onto word k, hi[L-1] onto the word above
        s = [res[0]] + asm.tmp(L - 1)
        for k in range(1, L):
            asm.add(s[k], res[k], hi[k - 1], cin=(k > 1), cout=True)
        t0, t1 = asm.tmp(), asm.tmp()
        asm.add(t0, top, hi[L - 1], cin=True, cout=True)
        asm.add(t1, 0, 0, cin=True)
        return self._fold_word(asm, s, t0, t1)

    def reduce_wide(self, asm, T):
        L = self.L
        return self._fold_double(asm, T[:L], T[L:], None)

    def reduce_small(self, asm, T):
        return self._fold_word(asm, T[:self.L], T[self.L])

    def build_mul(self):
        asm = Asm(self.name + ".mul")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        E, O = satmul.product_eo(asm, a, b)
        U = asm.tmp(L)
        for k in range(L):
            asm.add(U[k], E.src(L + k), O.src(L + k), cin=(k > 0), cout=(k < L - 1))
        lo_even = [E.src(k) for k in range(L)]
        lo_odd = {k: O.src(k) for k in range(1, L)}
        self._outs(asm, self._fold_double(asm, lo_even, U, lo_odd))
        return asm

    def build_add(self):
        asm = Asm(self.name + ".add")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        s = asm.tmp(L)
        c = asm.tmp()
        asm.add_chain(s, a, b, carry_to=(c, 0))
        self._outs(asm, self._fold_bit(asm, s, c))
        return asm

    def _fold_bit(self, asm, r, c):
        """r + c*2^(32L) for a bit c: add c*(2^32 + cl); a second wrap leaves a value below 2^34."""
        L, cl = self.L, self.cl
        f = asm.tmp()
        asm.madlo(f, c, cl, 0)
        o = asm.tmp(L)
        c2 = asm.tmp()
        asm.add_chain(o, r, [f, c] + [0] * (L - 2), carry_to=(c2, 0))
        g = asm.tmp()
        asm.madlo(g, c2, cl, 0)
        h0, h1 = asm.tmp(), asm.tmp()
        asm.add(h0, o[0], g, cout=True)
        asm.add(h1, o[1], c2, cin=True)
        return [h0, h1] + o[2:]

    def build_sub(self, neg):
        asm = Asm(self.name + (".neg" if neg else ".sub"))
        L, cl = self.L, self.cl
        if neg:
            (b,) = self._io(asm, ["b"])
            a = [0] * L
        else:
            a, b = self._io(asm, ["a", "b"])
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)           # value = d - 2^(32L)*[m] == d - c*[m]
        f, b1 = asm.tmp(), asm.tmp()
        asm.logic("and", f, m, cl)
        asm.logic("and", b1, m, 1)
        e = asm.tmp(L)
        m2 = asm.tmp()
        asm.sub_chain(e, d, [f, b1] + [0] * (L - 2), borrow_to=m2)
        f2, b2 = asm.tmp(), asm.tmp()
        asm.logic("and", f2, m2, cl)
        asm.logic("and", b2, m2, 1)
        g0, g1 = asm.tmp(), asm.tmp()
        asm.sub(g0, e[0], f2, cout=True)              # after a second borrow e >= 2^(32L) - c: no third one
        asm.sub(g1, e[1], b2, cin=True)
        self._outs(asm, [g0, g1] + e[2:])
        return asm

    def build_canon(self):
        asm = Asm(self.name + ".canon")
        (a,) = self._io(asm, ["a"])
        lt = asm.tmp()
        self._outs(asm, self._cond_sub_p(asm, a, want_flag=lt))
        asm.out("lt", lt)
        return asm


class PseudoMersenneBits(Plan):
    """2^n - c for ANY n (xs = 32L - n >= 2 spare bits in the top word, L even or odd) and c < 2^15 -- what
    pseudo.py's named table mostly holds (2^266 - 3, 2^206 - 5, 2^336 - 3, 2^414 - 17, 2^521 - 1; pseudo.py:1487-1550)
    and the word-aligned plan above cannot take.  Plain residues; stored values are any representative below
    2^n + 2^32.  Everything is folded at bit n: the part of a value at or above bit n is extracted with funnel
    shifts, multiplied by c and added at the bottom (second_pass, pseudo.py:557-611, restated for saturated limbs).

        mul / sqr   T = a*b < 2^(2n+1)          H = T >> n  (L funnel shifts),  R = (T mod 2^n) + c*H  (L wide
                    multiplies on even / odd windows, one merge),  then the few bits of R at or above bit n once more
        add / sub   one chain (sub: 2p added back under the borrow mask), then the top bits once

    L^2 + L wide multiplies per modmul where the fall-back plan needs 2 L^2 + L."""
    family = "pseudo"

    def __init__(self, prime):
        super().__init__(prime)
        n, L = prime.nbits, self.L
        self.c = (1 << n) - prime.p
        self.xs = 32 * L - n
        assert 2 <= self.xs <= 31 and L >= 3, "needs two spare bits in the top word"
        assert 0 < self.c < (1 << 15), "not of the form 2^n - c with a small c"
        self.bound = (1 << n) + (1 << 32)
        self.mask = M32 >> self.xs                      # bits of the top word below bit n

    # -- pieces -----------------------------------------------------------------------------------
    def _fold_top(self, asm, r, top=0, wide=False):
        """r (L words) + top * 2^(32L)  ->  L words below 2^n + c*h: everything at or above bit n (h, which must
        fit a word) times c, added at the bottom.  wide=True when c*h may exceed 32 bits (raw imports)."""
        L, xs = self.L, self.xs
        h = asm.tmp()
        if isinstance(top, int) and top == 0:
            asm.shr(h, r[L - 1], 32 - xs)
        else:
            asm.shfl(h, r[L - 1], top, xs)
        last = asm.tmp()
        asm.logic("and", last, r[L - 1], self.mask)
        o = asm.tmp(L)
        asm.madlo(o[0], h, self.c, r[0], cout=True)
        start = 1
        if wide:
            asm.madhi(o[1], h, self.c, r[1], cin=True, cout=True)
            start = 2
        for k in range(start, L - 1):
            asm.add(o[k], r[k], 0, cin=True, cout=True)
        asm.add(o[L - 1], last, 0, cin=True)
        return o, h

    def _fold_words(self, asm, v):
        """v: more than L words (a product: 2L, or L+1 after a small multiplication), value < 2^(2n+1)  ->  L words
        below 2^n + 2^32."""
        L, xs, c = self.L, self.xs, self.c
        nv = len(v)
        assert L < nv <= 2 * L
        # H = v >> n, word k = high word of (v[L+k] : v[L-1+k]) << xs
        m = nv - L + 1 if nv < 2 * L else L              # a full product is below 2^(2n+1): H has L words
        H = []
        for k in range(m):
            lo = v[L - 1 + k]
            hi = v[L + k] if L + k < nv else 0
            d = asm.tmp()
            if isinstance(hi, int) and hi == 0:
                asm.shr(d, lo, 32 - xs)
            else:
                asm.shfl(d, lo, hi, xs)
            H.append(d)
        last = asm.tmp()
        asm.logic("and", last, v[L - 1], self.mask)
        if m <= 2:
            # after a small multiplication: H is below 2^33, c*H below 2^48 -- one wide multiply, one add chain
            p0, p1 = asm.tmp(), asm.tmp()
            asm.mullo(p0, H[0], c)
            asm.mulhi(p1, H[0], c)
            if m == 2:
                q = asm.tmp()
                asm.madlo(q, H[1], c, p1)                  # H[1] is 0 or 1 and p1 < c: cannot carry
                p1 = q
            s = asm.tmp(L)
            asm.add_chain(s, list(v[:L - 1]) + [last], [p0, p1] + [0] * (L - 2))     # < 2^n + 2^48: no carry out
            r, _ = self._fold_top(asm, s)
            return r
        lo = list(v[:L - 1]) + [last] + [0, 0]
        # R = lo + c*H over L+1 words: even words of H chain onto the even-aligned windows of lo, odd words are
        # independent wide multiplies on the odd-aligned windows, one add-with-carry chain merges the two
        t = {}
        ev = []
        for k in range(0, m, 2):
            t[k], t[k + 1] = asm.tmp(), asm.tmp()
            ev.append((t[k], t[k + 1], H[k], c, lo[k], lo[k + 1]))
        ke = 2 * len(ev)                                   # first word above the even chain
        if ke <= L:                                        # L even: the chain ends at word L-1, its carry is word L
            assert ke == L
            t[ke] = asm.tmp()
            asm.wide_chain(ev, last_carry_to=(t[ke], 0))
        else:
            asm.wide_chain(ev, last_carry_to=None)         # L odd: the window (L-1, L) is the top, nothing leaves it
        o = {}
        for k in range(1, m, 2):
            o[k], o[k + 1] = asm.tmp(), asm.tmp()
            asm.mullo(o[k], H[k], c)
            asm.mulhi(o[k + 1], H[k], c)
        res = [t[0]]
        started = False
        for k in range(1, L + 1):
            a, b = t.get(k, lo[k] if k < len(lo) else 0), o.get(k, 0)
            if not started and isinstance(b, int) and b == 0:
                res.append(a)
                continue
            d = asm.tmp()
            asm.add(d, a, b, cin=started, cout=(k < L))
            started = True
            res.append(d)
        r, _ = self._fold_top(asm, res[:L], res[L])
        return r

    def reduce_wide(self, asm, T):
        return self._fold_words(asm, T)

    def reduce_small(self, asm, T):
        return self._fold_words(asm, T)

    def build_mla(self):
        asm = Asm(self.name + ".mla")
        a, c = self._io(asm, ["a", "c"])
        asm.inp("b")
        T = satmul.times_small(asm, a, "b")               # L+1 words
        L = self.L
        s = asm.tmp(L + 1)
        asm.add_chain(s, T, list(c) + [0])                 # a*b + c < 2^(n+33): still L+1 words
        self._outs(asm, self._fold_words(asm, s))
        return asm

    def build_add(self):
        asm = Asm(self.name + ".add")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        s = asm.tmp(L)
        asm.add_chain(s, a, b)                             # < 2^(n+1) + 2^33 < 2^(32L): no carry out
        r, _ = self._fold_top(asm, s)
        self._outs(asm, r)
        return asm

    def build_sub(self, neg):
        asm = Asm(self.name + (".neg" if neg else ".sub"))
        L = self.L
        if neg:
            (b,) = self._io(asm, ["b"])
            a = [0] * L
        else:
            a, b = self._io(asm, ["a", "b"])
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)                # value = d - 2^(32L)*[m]
        # a - b > -(2^n + 2^32) > -2p: add 2p back under the borrow mask (the carry out cancels the borrow)
        pw = words(2 * self.p, L)
        ops = []
        for k in range(L):
            if pw[k] == 0:
                ops.append(0)
            elif pw[k] == M32:
                ops.append(m)
            else:
                t = asm.tmp()
                asm.logic("and", t, m, pw[k])
                ops.append(t)
        e = asm.tmp(L)
        asm.add_chain(e, d, ops, wrap_ok=True)
        r, _ = self._fold_top(asm, e)                       # < 2^(n+1) + 2^32: the top bits once
        self._outs(asm, r)
        return asm

    def build_canon(self):
        """Any L-word value (a stored one, or a raw import up to 2^(32L)) -> canonical residue; lt = 1 iff the input
        was already < p (modfsb's return, pseudo.py:272-283)."""
        asm = Asm(self.name + ".canon")
        (a,) = self._io(asm, ["a"])
        r1, h1 = self._fold_top(asm, a, wide=True)         # < 2^n + c*2^xs
        r2, h2 = self._fold_top(asm, r1)                   # < 2^n + c = p + 2c
        lt = asm.tmp()
        r = self._cond_sub_p(asm, r2, want_flag=lt)
        # the flag describes the INPUT: it was below p iff nothing was folded and the subtraction borrowed
        nz, t, mz, nm, lt2 = asm.tmp(), asm.tmp(), asm.tmp(), asm.tmp(), asm.tmp()
        asm.logic("or", nz, h1, h2)
        asm.sub(t, 0, nz, cout=True)
        asm.nocheck.add(len(asm.ins))
        asm.sub(mz, 0, 0, cin=True)                        # all-ones iff something was folded
        asm.not_(nm, mz)
        asm.logic("and", lt2, lt, nm)
        self._outs(asm, r)
        asm.out("lt", lt2)
        return asm


class GenMersenne(Plan):
    """p = 2^(2h) - 2^h - 1, h = 32*H (X448: H=7, L=14).  Stored values are any
    representative < 2^(2h) (< 2p).  2^(2h) == 2^h + 1, so a double-length value
    lo + 2^(2h)*(h0 + 2^h*h1) reduces to lo + (h0+h1) + 2^h*(h0 + 2*h1): additions of
    half-length pieces only.  The reference reaches the same "no multiplies in the
    reduction" through Montgomery form with shaped limbs (monty.py:258-298,597-627,
    R = 2^476/2^504); applying the shape directly also removes nres/redc multiplies."""
    family = "monty"

    def __init__(self, prime):
        super().__init__(prime)
        n = prime.nbits
        assert n % 64 == 0 and prime.p == (1 << n) - (1 << (n // 2)) - 1
        self.H = self.L // 2

    def _fold_top(self, asm, r, c, tail=None):
        """r (L words) += c*(2^h + 1) for a dropped c*2^(2h); second wrap handled."""
        L, H = self.L, self.H
        o = asm.tmp(L)
        c2 = asm.tmp()
        for k in range(L):
            asm.add(o[k], r[k], c if k in (0, H) else 0, cin=(k > 0), cout=True)
        asm.add(c2, 0, 0, cin=True)
        n2 = L if tail is None else tail
        f = asm.tmp(n2)
        for k in range(n2):
            asm.add(f[k], o[k], c2 if k in (0, H) else 0, cin=(k > 0), cout=(k < n2 - 1))
        return f + o[n2:]

    def reduce_wide(self, asm, T):
        L, H = self.L, self.H
        lo, hi = T[:L], T[L:]
        h0, h1 = hi[:H], hi[H:]
        c = asm.tmp()
        s1 = asm.tmp(L)
        asm.add_chain(s1, lo, h0 + h0, carry_to=(c, 0))
        s2 = asm.tmp(L)
        c_b = asm.tmp()
        asm.add_chain(s2, s1, h1 + h1, carry_to=(c_b, c))
        s3 = asm.tmp(H)
        c_c = asm.tmp()
        asm.add_chain(s3, s2[H:], h1, carry_to=(c_c, c_b))
        # after a wrap the value is < 4*(2^h+1): the second pass stops at word H
        return self._fold_top(asm, s2[:H] + s3, c_c, tail=H + 1)

    def reduce_small(self, asm, T):
        return self._fold_top(asm, T[:self.L], T[self.L], tail=min(self.L, self.H + 2))

    def build_add(self):
        asm = Asm(self.name + ".add")
        a, b = self._io(asm, ["a", "b"])
        s = asm.tmp(self.L)
        c = asm.tmp()
        asm.add_chain(s, a, b, carry_to=(c, 0))
        self._outs(asm, self._fold_top(asm, s, c, tail=self.H + 1))
        return asm

    def build_sub(self, neg):
        asm = Asm(self.name + (".neg" if neg else ".sub"))
        L, H = self.L, self.H
        if neg:
            (b,) = self._io(asm, ["b"])
            a = [0] * L
        else:
            a, b = self._io(asm, ["a", "b"])
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)           # value = d - 2^(2h)*[m] == d - (2^h+1)*[m]
        b1 = asm.tmp()
        asm.logic("and", b1, m, 1)
        e = asm.tmp(L)
        m2 = asm.tmp()
        asm.sub_chain(e, d, [b1 if k in (0, H) else 0 for k in range(L)], borrow_to=m2)
        b2 = asm.tmp()
        asm.logic("and", b2, m2, 1)
        g = asm.tmp(L)
        asm.sub_chain(g, e, [b2 if k in (0, H) else 0 for k in range(L)])   # e >= p here: no borrow
        self._outs(asm, g)
        return asm

    def build_canon(self):
        asm = Asm(self.name + ".canon")
        (a,) = self._io(asm, ["a"])
        lt = asm.tmp()
        self._outs(asm, self._cond_sub_p(asm, a, want_flag=lt))
        asm.out("lt", lt)
        return asm


def _naf_terms(x, nbits):
    """Signed power-of-two terms of x (non-adjacent form), as [(bitpos, sign)]."""
    out = []
    k = 0
    while x:
        if x & 1:
            s = 2 - (x & 3)           # +1 or -1
            out.append((k, s))
            x -= s
        x >>= 1
        k += 1
    return out


class Montgomery(Plan):
    """Montgomery form with R = 2^(32L) for p == -1 (mod 2^32) whose complement
    d = 2^(32L) - p is a short signed sum of powers of two (P-256: d = 2^224 - 2^192 -
    2^96 + 1).  ndash = 1, the "Montgomery friendly" case of monty.py:740-751,
    2242-2244.  Instead of L word-serial rounds (monty.py:663-872 interleaves one per
    column) the whole quotient is produced at once:

        Q = T_lo * d^-1 mod R            (d^-1 is sparse too: shifts and adds)
        U = Q * d                        (shifts and adds; U_lo == T_lo by construction)
        a*b*R^-1 = T_hi + Q - U_hi       (< 2p, one conditional subtraction)

    so the reduction is ~60 add/sub-with-carry instructions on the ALU pipe and zero
    multiplies.  Stored values are fully reduced, in [0, p): 32L == Nbits leaves no
    spare bit for the reference's lazy "< 2p" results."""
    family = "monty"
    link_products = os.environ.get("MAB_LINK", "1") != "0"
    # Everything on this plan except the multiplication chain is ALU-bound (the reduction is add/sub-with-carry chains),
    # so the carry captures go to the multiplier pipe: modnsqr chain 138.6 -> 146.0 Gop/s, modinv per element 512 -> 530
    # Mop/s, ecnmul 21.66 -> 22.08 M/s (profiles/r2_p256_capop.txt, r2_ecn_capop.txt).  mul_w, which sits on the
    # boundary of the two regimes, loses 2 % with the same choice and keeps ptxas' SEL.
    capture_ops = {"mul": "madc", "sqr": "madc", "add": "madc", "sub": "madc", "sqr_w": "madc"}

    def __init__(self, prime):
        super().__init__(prime)
        L = self.L
        assert prime.nbits == 32 * L and prime.p % (1 << 32) == (1 << 32) - 1
        self.R = 1 << (32 * L)
        self.bound = prime.p
        self.d = self.R - prime.p
        self.dinv = pow(self.d, -1, self.R)
        self.d_terms = _naf_terms(self.d, 32 * L)
        self.dinv_terms = _naf_terms(self.dinv, 32 * L)
        assert len(self.d_terms) <= 8 and len(self.dinv_terms) <= 8, "modulus is not shaped"
        self.R2 = self.R * self.R % prime.p           # nres constant (cw, monty.py:2246-2248)
        # p + 1 = 2^g_shift * g, g a short signed sum of powers of two (see _redc)
        g, sh = prime.p + 1, 0
        while g % (1 << 32) == 0:
            g >>= 32
            sh += 32
        assert sh >= 32, "p + 1 has no whole zero word at the bottom"
        self.g_shift, self.g_terms = sh, _naf_terms(g, 32 * L)
        assert len(self.g_terms) <= 8 and all(k % 32 == 0 for k, _ in self.g_terms)
        self.weak_bound = self.R                      # mul_w / sqr_w: operands and results below R

    def _shifted(self, asm, src, k, N):
        """Words of (src << k) inside an N-word window: {index: reg}."""
        w0, b = divmod(k, 32)
        out = {}
        n = len(src)
        for j in range(n + (1 if b else 0)):
            idx = w0 + j
            if idx >= N:
                break
            if b == 0:
                out[idx] = src[j]
                continue
            lo = src[j - 1] if j > 0 else None
            hi = src[j] if j < n else None
            d = asm.tmp()
            if lo is None:
                asm.shl(d, hi, b)
            elif hi is None:
                asm.shr(d, lo, 32 - b)
            else:
                asm.shfl(d, lo, hi, b)
            out[idx] = d
        return out

    def _signed_sum(self, asm, src, terms, N, wrap_ok):
        """sum of +-(src << k) over an N-word window; positives first so the running
        value never goes negative when the exact result is non-negative."""
        acc = None
        for k, s in sorted(terms, key=lambda t: (-t[1], t[0])):
            sh = self._shifted(asm, src, k, N)
            if not sh:
                continue
            if acc is None:
                assert s > 0
                acc = [sh.get(i, 0) for i in range(N)]
                continue
            lo = min(sh)
            new = asm.tmp(N - lo)
            ops = [sh.get(i, 0) for i in range(lo, N)]
            if s > 0:
                asm.add_chain(new, acc[lo:], ops, wrap_ok=wrap_ok)
            else:
                asm.sub_chain(new, acc[lo:], ops, wrap_ok=wrap_ok)
            acc = acc[:lo] + new
        return acc

    def _redc(self, asm, T, weak=False):
        """T: 2L words -> L words: (T + Q*p) / R with Q = T_lo * (-p^-1) mod R.

        p + 1 = 2^s * g with g a short signed sum of powers of two (P-256: s = 96, g = 2^160 - 2^128 + 2^96 + 1),
        so (T + Q*p) / R = (T - Q + 2^s * Q*g) / R: H = Q*g is built with three add/sub chains that stop at the
        top of H (13 words), and ONE chain adds it to T from word s/32 upwards; the low L words of that sum equal Q
        by construction, so `- Q` needs no instruction at all.  51 add/sub-with-carry instructions for P-256 where
        the first formulation (U = Q*d over a 2L-word window, T_hi + Q - U_hi) took 60.
        weak=False: operands < p, result < 2p, conditional subtraction -> [0, p).
        weak=True : operands < R, result < R + p, p subtracted iff the carry word is set -> [0, R)."""
        L = self.L
        lo = T[:L]
        Q = self._signed_sum(asm, lo, self.dinv_terms, L, wrap_ok=True)
        sw = self.g_shift // 32
        # Q*g < 2^(32(2L-sw)) fits; the positive terms alone can exceed it by a bit for Q within 2^-64 of R, so the
        # chains work modulo 2^(32(2L-sw)) and the subtraction brings the value back
        H = self._signed_sum(asm, Q, self.g_terms, 2 * L - sw, wrap_ok=True)
        n = asm.tmp(2 * L - sw)
        top = asm.tmp()
        asm.add_chain(n, T[sw:], H, carry_to=(top, 0))
        hi = n[L - sw:] + [top]
        return self._weak_sub_p(asm, hi) if weak else self._cond_sub_p9(asm, hi)

    def _weak_sub_p(self, asm, v):
        """v: L+1 words, top word 0 or 1, value < R + p  ->  L words < R: subtract p iff the top word is set,
        i.e. add d = R - p under a mask and drop the carry."""
        L = self.L
        dw = words(self.d, L)
        m = asm.tmp()
        asm.sub(m, 0, v[L])                                  # all-ones iff top set
        asm.nocheck.add(len(asm.ins) - 1)
        ops = []
        for k in range(L):
            if dw[k] == 0:
                ops.append(0)
            elif dw[k] == M32:
                ops.append(m)
            elif dw[k] == 1:
                ops.append(v[L])
            else:
                t = asm.tmp()
                asm.logic("and", t, m, dw[k])
                ops.append(t)
        r = asm.tmp(L)
        asm.add_chain(r, v[:L], ops, wrap_ok=True)
        # the chain's carry-out cancels the top word; the interpreter cannot see that, the bignum self-check does
        return r

    def _cond_sub_p9(self, asm, v):
        """v: L+1 words, < 2p  ->  L words in [0,p)."""
        L = self.L
        pw = words(self.p, L) + [0]
        t = asm.tmp(L + 1)
        m = asm.tmp()
        asm.sub_chain(t, v, pw, borrow_to=m)
        r = []
        for k in range(L):
            x, y, z = asm.tmp(), asm.tmp(), asm.tmp()
            asm.logic("xor", x, t[k], v[k])
            asm.logic("and", y, x, m)
            asm.logic("xor", z, y, t[k])
            r.append(z)
        return r

    def reduce_wide(self, asm, T):
        return self._redc(asm, T)

    # Weakly reduced products for chains of multiplications (modpro, modnsqr): operands and results are any
    # representative below R = 2^(32L) (< 2p), and the final step looks at the carry word only -- 10 instructions
    # where the full conditional subtraction takes 18.  A chain ends with `canon`, which brings the value back
    # into [0, p), the invariant every other function of this plan keeps.
    def build_mul_w(self):
        asm = Asm(self.name + ".mul_w")
        a, b = self._io(asm, ["a", "b"])
        T = satmul.product(asm, a, b, link=self.link_products)
        self._outs(asm, self._redc(asm, T, weak=True))
        return asm

    def build_sqr_w(self):
        asm = Asm(self.name + ".sqr_w")
        (a,) = self._io(asm, ["a"])
        T = satmul.square(asm, a, link=self.link_products)
        self._outs(asm, self._redc(asm, T, weak=True))
        return asm

    def reduce_small(self, asm, T):
        """L+1 words (a*b, b < 2^32; no R factor involved, monty.py:876-978) -> [0,p):
        2^(32L) == d, so fold the top word through d's word-aligned terms until it fits."""
        L = self.L
        assert all(k % 32 == 0 for k, _ in self.d_terms)
        cur, top = T[:L], T[L]
        for _ in range(3):
            w = self._signed_sum_top(asm, cur, top)
            cur, top = w[:L], w[L]
        # third fold cannot carry (interpreter checks via the final compare width)
        return self._cond_sub_p9(asm, cur + [top])

    def _signed_sum_top(self, asm, cur, c):
        L = self.L
        acc = list(cur) + [0]
        for k, s in sorted(self.d_terms, key=lambda t: (-t[1], t[0])):
            lo = k // 32
            new = asm.tmp(L + 1 - lo)
            ops = [c] + [0] * (L - lo)
            if s > 0:
                asm.add_chain(new, acc[lo:], ops)
            else:
                asm.sub_chain(new, acc[lo:], ops)
            acc = acc[:lo] + new
        return acc

    def build_add(self):
        asm = Asm(self.name + ".add")
        a, b = self._io(asm, ["a", "b"])
        L = self.L
        s = asm.tmp(L + 1)
        asm.add_chain(s[:L], a, b, carry_to=(s[L], 0))
        self._outs(asm, self._cond_sub_p9(asm, s))
        return asm

    def build_sub(self, neg):
        asm = Asm(self.name + (".neg" if neg else ".sub"))
        L = self.L
        if neg:
            (b,) = self._io(asm, ["b"])
            a = [0] * L
        else:
            a, b = self._io(asm, ["a", "b"])
        d = asm.tmp(L)
        m = asm.tmp()
        asm.sub_chain(d, a, b, borrow_to=m)
        pw = words(self.p, L)
        pm = []
        for k in range(L):
            if pw[k] == 0:
                pm.append(0)
            elif pw[k] == M32:
                pm.append(m)
            else:
                t = asm.tmp()
                asm.logic("and", t, m, pw[k])
                pm.append(t)
        r = asm.tmp(L)
        asm.add_chain(r, d, pm, wrap_ok=True)
        self._outs(asm, r)
        return asm

    def build_canon(self):
        """Stored values are already canonical; still performs the conditional subtraction
        so that an out-of-range import (modimp of a value in [p, 2^(32L))) is handled here."""
        asm = Asm(self.name + ".canon")
        (a,) = self._io(asm, ["a"])
        lt = asm.tmp()
        self._outs(asm, self._cond_sub_p(asm, a, want_flag=lt))
        asm.out("lt", lt)
        return asm


class MontgomeryFull(Montgomery):
    """Montgomery form, R = 2^(32L), for ANY odd modulus below 2^(32L) -- the fallback that makes
    the generator total, as monty.py is for the reference (full Montgomery, ndash != 1:
    monty.py:740-751,2237-2244; e.g. group orders, monty.py:2110-2127).

    Default: word-serial Montgomery multiplication interleaved with the product on the even/odd
    accumulators (satmul.montgomery_interleaved): 2 L^2 wide multiplies + L plain ones per modmul,
    squaring by the same routine, small multiples via the lift b -> b*R (one short and one full
    multiplication).  MAB_MONTY=separate selects the first formulation, kept for comparison:

        Q = T_lo * (-p^-1 mod R)  mod R      L(L+1)/2 wide multiplies by a constant
        U = Q * p                            L^2 wide multiplies by a constant
        a*b*R^-1 = (T + U) / R               one 2L-word add chain, one conditional subtraction

    Stored values are fully reduced, in [0, p)."""

    capture_ops = {}        # multiplier-bound: ptxas' choice

    def __init__(self, prime):
        Plan.__init__(self, prime)
        L = self.L
        assert prime.p % 2 == 1
        self.R = 1 << (32 * L)
        self.bound = prime.p
        self.nprime = (-pow(prime.p, -1, self.R)) % self.R
        self.R2 = self.R * self.R % prime.p
        self.d_terms = self.dinv_terms = None

    def _redc(self, asm, T):
        L = self.L
        lo = T[:L]
        Q = satmul.product_low(asm, lo, words(self.nprime, L))
        U = satmul.product(asm, Q, words(self.p, L))
        s = asm.tmp(2 * L)
        top = asm.tmp()
        asm.add_chain(s, T, U, carry_to=(top, 0))          # low half becomes zero by construction
        return self._cond_sub_p9(asm, s[L:] + [top])

    def reduce_small(self, asm, T):
        raise NotImplementedError

    # MAB_MONTY=separate keeps the first formulation (product, then Q = T_lo*n', U = Q*p: 2.5 L^2 wide multiplies)
    interleaved = os.environ.get("MAB_MONTY", "interleaved") != "separate"

    def _mont(self, asm, a, b):
        """a*b*R^-1 mod p, fully reduced."""
        if not self.interleaved:
            return self._redc(asm, satmul.product(asm, a, b))
        L = self.L
        n0 = self.nprime & M32
        return self._cond_sub_p9(asm, satmul.montgomery_interleaved(asm, a, b, words(self.p, L), n0))

    def build_mul(self):
        asm = Asm(self.name + ".mul")
        a, b = self._io(asm, ["a", "b"])
        self._outs(asm, self._mont(asm, a, b))
        return asm

    def build_sqr(self):
        if not self.interleaved:
            return super().build_sqr()
        asm = Asm(self.name + ".sqr")
        (a,) = self._io(asm, ["a"])
        self._outs(asm, self._mont(asm, a, a))
        return asm

    def _small_times(self, asm, a, bname, c=None):
        """a*b (+c) for a small plain integer b: b is lifted into Montgomery form with one
        multiplication by R^2 (b*R), then multiplied in; no R factor remains (monty.py:876-978
        gets there with a Barrett-Dhem estimate instead)."""
        L = self.L
        r2 = words(self.R2, L)
        if self.interleaved:
            bm = self._mont(asm, r2, [bname] + [0] * (L - 1))        # (R^2 mod p) * b * R^-1 = b*R mod p; zero rows skipped
            r = self._mont(asm, a, bm)
        else:
            t = satmul.times_small(asm, r2, bname)                       # R^2 * b  (L+1 words)
            # Montgomery-reduce R^2*b (< 2^32 * p): pad to 2L words
            bm = self._redc(asm, t + [0] * (2 * L - (L + 1)))            # = b*R mod p
            T = satmul.product(asm, a, bm)
            r = self._redc(asm, T)
        if c is None:
            return r
        s = asm.tmp(L + 1)
        asm.add_chain(s[:L], r, c, carry_to=(s[L], 0))
        return self._cond_sub_p9(asm, s)

    def build_mli(self):
        asm = Asm(self.name + ".mli")
        (a,) = self._io(asm, ["a"])
        asm.inp("b")
        self._outs(asm, self._small_times(asm, a, "b"))
        return asm

    def build_mla(self):
        asm = Asm(self.name + ".mla")
        a, c = self._io(asm, ["a", "c"])
        asm.inp("b")
        self._outs(asm, self._small_times(asm, a, "b", c))
        return asm


class MontgomeryFriendly(MontgomeryFull):
    """Montgomery form, R = 2^(32L), for a modulus with p = -1 (mod 2^(32z)), z >= 1 whole words: p + 1 = 2^(32z) * q.
    The quotient digits are the low words themselves and only the L - z words of q are multiplied
    (satmul.montgomery_friendly): L^2 + L(L - z) wide multiplies per modmul instead of the fall-back plan's 2 L^2 + L.
    What monty.py's named table holds beyond the NIST shapes is mostly of this kind: 2^a 3^b - 1 (z = 6 .. 11 of
    14 .. 24 words), 3*67*2^246 - 1 and 5*2^248 - 1 (z = 7 of 8), 2^n - 2^m - 1 (z = m div 32), P-384 (z = 1)
    (monty.py:1961-2108)."""

    def __init__(self, prime):
        MontgomeryFull.__init__(self, prime)
        z = 0
        while (prime.p + 1) % (1 << (32 * (z + 1))) == 0:
            z += 1
        assert 1 <= z < self.L, "p + 1 has no whole zero word at the bottom"
        self.z = z
        self.q = (prime.p + 1) >> (32 * z)

    def _mont(self, asm, a, b):
        L = self.L
        return self._cond_sub_p9(asm, satmul.montgomery_friendly(asm, a, b, words(self.q, L - self.z), self.z))


def make_plan(prime: Prime) -> Plan:
    """Choose the cheapest plan whose preconditions hold for this modulus (cf. the radix /
    strategy decisions of pseudo.py:1569-1678 and monty.py:2140-2248); MontgomeryFull accepts
    any odd modulus, so every prime the reference generators take is covered."""
    p, n = prime.p, prime.nbits
    L = (n + 31) // 32
    c = (1 << n) - p
    cands = []
    if n % 64 == 0 and p == (1 << n) - (1 << (n // 2)) - 1 and (n // 2) % 32 == 0:
        cands.append(GenMersenne)
    if L % 2 == 0 and c < (1 << 12) and (c << (32 * L - n)) < (1 << 15):
        cands.append(PseudoMersenne)
    if L % 2 == 0 and n == 32 * L and 0 < c - (1 << 32) < (1 << 15):
        cands.append(PseudoMersenne33)
    if 0 < c < (1 << 15) and 32 * L - n >= 2 and os.environ.get("MAB_PMBITS", "1") != "0":
        cands.append(PseudoMersenneBits)
    cands += [Montgomery]
    if (p + 1) % (1 << 32) == 0 and os.environ.get("MAB_MFRIENDLY", "1") != "0":
        cands.append(MontgomeryFriendly)
    cands.append(MontgomeryFull)
    err = None
    for cls in cands:
        try:
            plan = cls(prime)
            plan.build()
            return plan
        except (AssertionError, NotImplementedError) as e:
            err = e
    raise ValueError("no limb plan for modulus %s: %r" % (prime.name, err))
