"""Full-width limb products for saturated radix-2^32 operands.

The reference's product rows (`getZM`/`getZS`, pseudo.py:351-554; `getZMU`,
monty.py:493-517) accumulate unsaturated limbs into a double-word `t` and
mask/shift a limb out per column.  On sm_100a the multiplier is the 32-bit IMAD
pipe and `IMAD.WIDE.U32.X` does a 32x32+64->64 multiply-accumulate with carry-in
and carry-out in predicate registers, so the densest schedule is L*L wide
products (L = ceil(Nbits/32) limbs, the algorithmic minimum of SURVEY.md
section 8d) arranged so that every product lands on a 64-bit aligned accumulator
window:

  * products a[i]*b[j] with i+j even accumulate into the EVEN array E, whose
    windows are words (0,1),(2,3),...;
  * products with i+j odd accumulate into the ODD array O, windows (1,2),(3,4),...;
  * within a row the windows are adjacent, so one carry chain ripples through the
    row and its final carry is captured in the word above;
  * E and O are merged once at the end with a single add-with-carry chain.

All carry-capture words are shown (by the interpreter in ptx.py, which traps any
lost carry) to hold only a few carry bits when they are written.
"""
from __future__ import annotations

import os

from .ptx import Asm

M32 = 0xFFFFFFFF
# generator tuning knobs (kernel experiments; defaults are what measured best on B200)
SMART_CAPTURE = os.environ.get("MAB_CAPTURE", "smart") == "smart"
ZERO_REG = os.environ.get("MAB_ZERO", "lit") == "reg"
# MAB_ROWS=fresh: within each accumulator array the row that opens a new top window is emitted BEFORE the row
# whose chain carries into that window, and the carry is added onto the live window (two add-with-carry
# instructions) instead of being captured in a word of its own that later needs a zero partner to form an
# aligned (carry, 0) addend pair (ptxas materialises that zero with HFMA2 on the multiplier pipe half the time)
ROW_ORDER = os.environ.get("MAB_ROWS", "natural")


class _Acc:
    """Accumulator words that start as the literal 0 until first written.  `ub` tracks an
    upper bound of every word so that carry captures which are provably zero are not emitted
    (the interpreter in ptx.py re-checks every such claim on concrete values)."""

    def __init__(self, asm: Asm, n: int):
        self.asm = asm
        self.reg = [asm.tmp() for _ in range(n)]
        self.live = [False] * n
        self.ub = [0] * n
        self._zero = None

    def src(self, k):
        if self.live[k]:
            return self.reg[k]
        if ZERO_REG:
            if self._zero is None:
                self._zero = self.asm.tmp()
                self.asm.mov(self._zero, 0)
            return self._zero
        return 0

    def dst(self, k):
        self.live[k] = True
        return self.reg[k]


def _row_chain(asm: Asm, acc: _Acc, prods, nwords, carry_in=False):
    """prods: list of (s, a, b) with s ascending in steps of 2; window (s, s+1).  carry_in: the first
    window also takes the carry flag left by the instruction emitted just before."""
    if not prods:
        assert not carry_in
        return
    fresh = (not carry_in) and all(not acc.live[s] and not acc.live[s + 1] for s, _, _ in prods)
    if fresh:
        # untouched windows: independent wide multiplies, no carries can arise
        for s, a, b in prods:
            asm.mullo(acc.dst(s), a, b)
            asm.mulhi(acc.dst(s + 1), a, b)
            acc.ub[s] = M32
            acc.ub[s + 1] = M32 - 1
        return
    slots = []
    cin = 1 if carry_in else 0
    for s, a, b in prods:
        clo, chi = acc.src(s), acc.src(s + 1)
        slots.append((acc.dst(s), acc.dst(s + 1), a, b, clo, chi))
        vmax = ((acc.ub[s + 1] << 32) | acc.ub[s]) + M32 * M32 + cin      # window after this product
        cin = 1 if vmax >> 64 else 0
        acc.ub[s] = M32
        acc.ub[s + 1] = M32 if cin else (vmax >> 32)
    top = prods[-1][0] + 2
    if top < nwords and (cin or not SMART_CAPTURE):
        if ROW_ORDER == "fresh" and top + 1 < nwords and acc.live[top] and acc.live[top + 1]:
            # the window above is live: add the carry onto it; the window holds few enough products that it cannot
            # overflow (bound tracked in ub, re-checked by the interpreter on concrete values)
            vmax = ((acc.ub[top + 1] << 32) | acc.ub[top]) + 1
            assert vmax >> 64 == 0, "carry into a live window could ripple further"
            asm.wide_chain(slots, last_carry_to=None, carry_in=carry_in, keep_carry=True)
            lo, hi = acc.reg[top], acc.reg[top + 1]
            nlo, nhi = asm.tmp(), asm.tmp()
            asm.add(nlo, lo, 0, cin=True, cout=True)
            asm.add(nhi, hi, 0, cin=True, cout=False)
            acc.reg[top], acc.reg[top + 1] = nlo, nhi
            acc.ub[top], acc.ub[top + 1] = vmax & M32 if (vmax >> 32) == 0 else M32, min(M32, vmax >> 32)
        elif acc.live[top] and acc.ub[top] == M32:
            # the word above the row is a full word of another row (rows shorter than the ones before them:
            # montgomery_friendly): the carry ripples on until a word that can absorb it
            asm.wide_chain(slots, last_carry_to=None, carry_in=carry_in, keep_carry=True)
            w = top
            while True:
                more = acc.live[w] and acc.ub[w] == M32 and w + 1 < nwords
                d = asm.tmp()
                asm.add(d, acc.src(w), 0, cin=True, cout=more)
                was = acc.ub[w] if acc.live[w] else 0
                acc.reg[w], acc.live[w] = d, True
                acc.ub[w] = min(M32, was + 1)
                if not more:
                    break
                w += 1
        else:
            c = acc.src(top)
            asm.wide_chain(slots, last_carry_to=(acc.dst(top), c), carry_in=carry_in)
            acc.ub[top] = min(M32, acc.ub[top] + 1)
    else:
        asm.wide_chain(slots, last_carry_to=None, carry_in=carry_in)


def _merge(asm: Asm, E: _Acc, O: _Acc, nwords, link=False):
    """T = E + O as one carry chain; returns list of nwords regs/literals.  link=True orders the
    chain after the last product chain (Asm.link_cc) -- for reductions that consume the low words
    first, where ptxas would otherwise pull every row's low windows forward."""
    T = []
    started = False
    linked = False
    for k in range(nwords):
        e, o = E.src(k), O.src(k)
        if not started and (isinstance(e, int) or isinstance(o, int)) and (e == 0 or o == 0):
            # nothing to add yet (word 0 has no odd part): pass through
            T.append(o if e == 0 else e)
            continue
        d = asm.tmp()
        if link and not started:
            linked = asm.link_cc()
        asm.add(d, e, o, cin=(started or linked), cout=(k < nwords - 1))
        started = True
        T.append(d)
    return T


def product_eo(asm: Asm, a, b):
    """a*b as the two unmerged accumulator arrays: value = E + O (2L words each; word k of
    either array is `X.src(k)`, the literal 0 where never written)."""
    L = len(a)
    assert len(b) == L
    n = 2 * L
    E, O = _Acc(asm, n), _Acc(asm, n)
    rows_e = rows_o = list(range(L))
    if ROW_ORDER == "fresh" and L % 2 == 0:
        # E: rows 1, 3, 5.. open the windows 8, 10, 12.. that rows 2, 4, 6.. carry into; O: rows 2, 4, 6.. open the
        # windows 9, 11, 13.. that rows 1, 3, 5.. carry into
        rows_e = [0, 1] + [r for k in range(1, L // 2) for r in (2 * k + 1, 2 * k)]
        rows_o = [0] + [r for k in range(1, L // 2) for r in (2 * k, 2 * k - 1)] + [L - 1]
    for k in range(L):
        i = rows_e[k]
        _row_chain(asm, E, [(i + j, a[j], b[i]) for j in range(L) if (i + j) % 2 == 0], n)
        i = rows_o[k]
        _row_chain(asm, O, [(i + j, a[j], b[i]) for j in range(L) if (i + j) % 2 == 1], n)
    return E, O


def product(asm: Asm, a, b, link=False):
    """Return the 2L words (registers) of a*b; a, b are lists of L register names."""
    E, O = product_eo(asm, a, b)
    return _merge(asm, E, O, 2 * len(a), link=link)


def square(asm: Asm, a, link=False):
    """Return the 2L words of a*a: 2 * sum_{i<j} a_i a_j 2^(32(i+j)) + sum a_i^2 2^(64 i).

    L(L-1)/2 off-diagonal wide products in even/odd chains, one funnel-shift
    doubling pass on the ALU pipe, then the L diagonal squares added with one
    wide carry chain: L(L+1)/2 wide products in total.
    """
    L = len(a)
    n = 2 * L
    E, O = _Acc(asm, n), _Acc(asm, n)
    for i in range(L):
        ev = [(i + j, a[i], a[j]) for j in range(i + 1, L) if (i + j) % 2 == 0]
        od = [(i + j, a[i], a[j]) for j in range(i + 1, L) if (i + j) % 2 == 1]
        _row_chain(asm, E, ev, n)
        _row_chain(asm, O, od, n)
    S = _merge(asm, E, O, n, link=link)            # off-diagonal sum, < 2^(64L-1)
    # double: D[k] = (S[k] << 1) | (S[k-1] >> 31)
    D = []
    for k in range(n):
        lo = S[k - 1] if k > 0 else 0
        hi = S[k]
        if isinstance(hi, int) and isinstance(lo, int):
            D.append(0)
            continue
        d = asm.tmp()
        if isinstance(lo, int):
            asm.shl(d, hi, 1)
        elif isinstance(hi, int):
            asm.shr(d, lo, 31)
        else:
            asm.shfl(d, lo, hi, 1)
        D.append(d)
    # add the diagonal: window (2i, 2i+1) += a_i^2, one chain
    T = [asm.tmp() for _ in range(n)]
    slots = [(T[2 * i], T[2 * i + 1], a[i], a[i], D[2 * i], D[2 * i + 1]) for i in range(L)]
    asm.wide_chain(slots, last_carry_to=None)
    return T


def times_small(asm: Asm, a, b):
    """Return L+1 words of a*b for one 32-bit multiplier b (register or literal)."""
    L = len(a)
    lo = [asm.tmp() for _ in range(L)]
    hi = [asm.tmp() for _ in range(L)]
    for j in range(L):
        asm.mullo(lo[j], a[j], b)
        asm.mulhi(hi[j], a[j], b)
    T = [lo[0]]
    for k in range(1, L):
        d = asm.tmp()
        asm.add(d, lo[k], hi[k - 1], cin=(k > 1), cout=True)
        T.append(d)
    d = asm.tmp()
    asm.add(d, hi[L - 1], 0, cin=True, cout=False)
    T.append(d)
    return T


def times_small_add(asm: Asm, a, b, c):
    """Return L+1 words of a*b + c for one 32-bit multiplier b: even limbs accumulate onto the
    even-aligned windows of c with one wide carry chain, odd limbs are independent wide multiplies,
    and one add-with-carry chain merges the two (L even)."""
    L = len(a)
    assert L % 2 == 0 and len(c) == L
    t = [asm.tmp() for _ in range(L)]
    ce = asm.tmp()
    asm.wide_chain([(t[k], t[k + 1], a[k], b, c[k], c[k + 1]) for k in range(0, L, 2)], last_carry_to=(ce, 0))
    o = {k: asm.tmp() for k in range(1, L + 1)}
    for k in range(1, L, 2):
        asm.mullo(o[k], a[k], b)
        asm.mulhi(o[k + 1], a[k], b)
    res = [t[0]] + [asm.tmp() for _ in range(L - 1)]
    top = asm.tmp()
    for k in range(1, L):
        asm.add(res[k], t[k], o[k], cin=(k > 1), cout=True)
    asm.add(top, ce, o[L], cin=True, cout=False)
    return res + [top]


def product_low(asm: Asm, a, b):
    """Low L words of a*b (mod 2^(32L)); b may hold integer literals (a constant operand).
    Only the products that reach below word L are formed: L(L+1)/2 wide multiplies, the ones
    landing on word L-1 as plain 32-bit multiplies."""
    L = len(a)
    lo_only = []                       # (a_j, b_i) whose low word lands on word L-1
    n = L + 1                          # one scratch word above the result absorbs what is dropped
    E, O = _Acc(asm, n), _Acc(asm, n)
    for i in range(L):
        if isinstance(b[i], int) and b[i] == 0:
            continue
        ev = [(i + j, a[j], b[i]) for j in range(L) if (i + j) % 2 == 0 and i + j <= L - 2]
        od = [(i + j, a[j], b[i]) for j in range(L) if (i + j) % 2 == 1 and i + j <= L - 2]
        for acc, prods in ((E, ev), (O, od)):
            if not prods:
                continue
            start = len(asm.ins)
            _row_chain(asm, acc, prods, n)
            # whatever leaves the low L words is dropped on purpose
            for k in range(start, len(asm.ins)):
                asm.nocheck.add(k)
        lo_only.append((a[L - 1 - i], b[i]))
    T = []
    started = False
    for k in range(L):
        e, o = E.src(k), O.src(k)
        if not started and (e == 0 or o == 0) and (isinstance(e, int) or isinstance(o, int)):
            T.append(o if (isinstance(e, int) and e == 0) else e)
            continue
        d = asm.tmp()
        if k == L - 1:
            asm.nocheck.add(len(asm.ins))
        asm.add(d, e, o, cin=started, cout=(k < L - 1))
        started = True
        T.append(d)
    top = T[L - 1]
    for (x, y) in lo_only:
        d = asm.tmp()
        asm.nocheck.add(len(asm.ins))
        asm.madlo(d, x, y, top)
        top = d
    T[L - 1] = top
    return T


def montgomery_interleaved(asm: Asm, a, b, pw, n0):
    """a*b*2^(-32L) mod p before the final conditional subtraction, as L+1 words (< 2p for a, b < p).

    Word-serial Montgomery multiplication on the same even/odd accumulators as `product_eo`
    (monty.py:663-872 interleaves one reduction step per column; the pairing of the two arrays follows
    the even/odd scheme known from GPU big-number libraries): round i adds the row a*b[i] at windows
    i, i+1, ..., then m = T[i]*n0 mod 2^32 and the row m*p at the same windows, which zeroes word i.
    The 64-bit window (i, i+1) belongs to array X (E for even i, O for odd i); the other array Y holds
    word i as the HIGH word of its window (i-1, i), final by now.  Y[i] is added into X[i] with one add
    whose carry-out has exactly the weight of Y's next window (i+1, i+2) -- it becomes the carry-in of
    the m*p chain on Y.  The carry out of word i when m*p[0] zeroes it stays inside X's window.
    2L^2 wide multiplies + L plain multiplies; b[i] may be the literal 0 (row skipped), a and pw may be
    literals."""
    L = len(a)
    assert len(b) == L and len(pw) == L
    n = 2 * L + 2
    E, O = _Acc(asm, n), _Acc(asm, n)
    for i in range(L):
        X, Y = (E, O) if i % 2 == 0 else (O, E)
        if not (isinstance(b[i], int) and b[i] == 0):
            _row_chain(asm, X, [(i + j, a[j], b[i]) for j in range(0, L, 2)], n)
            _row_chain(asm, Y, [(i + j, a[j], b[i]) for j in range(1, L, 2)], n)
        # fold the finished word i of Y into X's window; its carry belongs to word i+1 = Y's next window
        carry = False
        if Y.live[i]:
            d = asm.tmp()
            asm.add(d, X.src(i), Y.reg[i], cout=True)
            X.reg[i], X.live[i], X.ub[i] = d, True, M32
            carry = True
        if not X.live[i]:                      # nothing has reached word i yet (leading zero rows): m = 0
            assert not carry
            continue
        m = asm.tmp()
        asm.nocheck.add(len(asm.ins))
        asm.mullo(m, X.reg[i], n0)
        # pw[j] == 0 still occupies its window (the chain has to pass the carry on)
        _row_chain(asm, Y, [(i + j, pw[j], m) for j in range(1, L, 2)], n, carry_in=carry)
        _row_chain(asm, X, [(i + j, pw[j], m) for j in range(0, L, 2)], n)
    # words below L are zero (or dead); the result is E + O over words L .. 2L
    T = []
    started = False
    for k in range(L, 2 * L + 1):
        e, o = E.src(k), O.src(k)
        if not started and (isinstance(e, int) or isinstance(o, int)) and (e == 0 or o == 0):
            T.append(o if (isinstance(e, int) and e == 0) else e)
            continue
        d = asm.tmp()
        asm.add(d, e, o, cin=started, cout=(k < 2 * L))
        started = True
        T.append(d)
    return T


def montgomery_friendly(asm: Asm, a, b, q, z):
    """a*b*2^(-32L) mod p before the final conditional subtraction (L+1 words, < 2p) for a modulus with
    p = -1 (mod 2^(32z)), z >= 1: p + 1 = 2^(32z) * q.  Then -p^-1 = 1 (mod 2^(32z)), so the quotient digit of a
    z-word block is the block itself, and T + m*p = T - m + m*q*2^(32z): the low z words cancel WITHOUT a single
    instruction and only the L - z words of q are multiplied (monty.py:740-751 calls the one-word case "Montgomery
    friendly"; isogeny and MFP primes have z up to L - 1).  L^2 + L(L - z) wide multiplies and no plain ones, where
    `montgomery_interleaved` needs 2 L^2 + L.

    Rounds are taken z at a time on the even/odd accumulators of `product_eo`: the product rows a*b[i] of a block,
    then ONE add-with-carry chain E[k] + O[k] over the block's z words -- the sums are the digits m_k, and the chain's
    carry has exactly the weight of word i0 + z, where the first row m*q starts: it is that chain's carry-in -- then
    the rows m_k*q at word k + z.  Words below L are never touched again and never read."""
    L = len(a)
    nq = len(q)
    assert len(b) == L and 1 <= z < L and nq == L - z
    n = 2 * L + 2
    E, O = _Acc(asm, n), _Acc(asm, n)

    def arrays(s):
        return (E, O) if s % 2 == 0 else (O, E)

    for i0 in range(0, L, z):
        blk = list(range(i0, min(L, i0 + z)))
        for i in blk:
            if isinstance(b[i], int) and b[i] == 0:
                continue
            X, Y = arrays(i)
            _row_chain(asm, X, [(i + j, a[j], b[i]) for j in range(0, L, 2)], n)
            _row_chain(asm, Y, [(i + j, a[j], b[i]) for j in range(1, L, 2)], n)
        # the digits of the block: m_k = (E[k] + O[k] + carry) mod 2^32
        m = {}
        started = False
        for k in blk:
            e, o = E.src(k), O.src(k)
            if not started and (isinstance(e, int) or isinstance(o, int)) and (e == 0 or o == 0):
                m[k] = o if (isinstance(e, int) and e == 0) else e       # nothing to add: no carry can arise
                continue
            d = asm.tmp()
            asm.add(d, e, o, cin=started, cout=True)
            started = True
            m[k] = d
        # a short block (the last one when z does not divide L): the chain goes on over the words up to i0 + z - 1,
        # which merges them into E, so that its carry still has the weight of the word where the rows start
        for k in range(blk[-1] + 1, i0 + z):
            e, o = E.src(k), O.src(k)
            if not started and (isinstance(e, int) or isinstance(o, int)) and (e == 0 or o == 0):
                if isinstance(e, int) and e == 0 and not isinstance(o, int):
                    E.reg[k], E.live[k], E.ub[k] = O.reg[k], True, O.ub[k]
                    O.reg[k], O.live[k], O.ub[k] = asm.tmp(), False, 0
                continue
            d = asm.tmp()
            asm.add(d, e, o, cin=started, cout=True)
            started = True
            E.reg[k], E.live[k], E.ub[k] = d, True, M32
            O.reg[k], O.live[k], O.ub[k] = asm.tmp(), False, 0
        carry = started
        for k in blk:
            if isinstance(m[k], int) and m[k] == 0:                       # leading zero rows of b: digit 0, no carry yet
                assert not carry
                continue
            X, Y = arrays(k + z)
            ev = [(k + z + t, q[t], m[k]) for t in range(0, nq, 2)]         # windows of the parity of k + z
            od = [(k + z + t, q[t], m[k]) for t in range(1, nq, 2)]
            _row_chain(asm, X, ev, n, carry_in=carry)                     # the block's first row takes the digit chain's carry
            carry = False
            _row_chain(asm, Y, od, n)
    T = []
    started = False
    for k in range(L, 2 * L + 1):
        e, o = E.src(k), O.src(k)
        if not started and (isinstance(e, int) or isinstance(o, int)) and (e == 0 or o == 0):
            T.append(o if (isinstance(e, int) and e == 0) else e)
            continue
        d = asm.tmp()
        asm.add(d, e, o, cin=started, cout=(k < 2 * L))
        started = True
        T.append(d)
    return T
