"""Addition chains for the fixed exponentiations (modpro / modinv / modsqrt).

The reference shells out to the external Go tool `addchain`
(pseudo.py:1582-1587, monty.py:2166-2170) and turns its `tmp/double/add/shift`
program into straight-line modsqr/modmul/modnsqr calls (pseudo.py:758-785).
That tool is not vendored and cannot be installed offline, so this module is
our own finder.  The progenitor exponents PE=(p-1-2^k)/2^(k+1) of the moduli we
care about are long runs of ones, so a run-length decomposition is used:

  * split the exponent into maximal runs of 1 bits,
  * build x^(2^L-1) for every needed run length L with a doubling chain on the
    run lengths (x_{a+b} = x_a^(2^b) * x_b),
  * assemble left to right: acc = acc^(2^(gap+L)) * x_L.

For 2^252-3 this yields 251 squarings + 12 multiplies (the classic 25519
chain has 11); results never depend on the chain after `redc`.

The program is a list of ops over named variables, "x" = input, "z" = output:
    ("sqr", dst, src, n)      dst = src^(2^n)
    ("mul", dst, a, b)        dst = a*b
`to_reference_text` prints it in the format pseudo.py:759-783 parses, so the
same finder backs the `addchain` stand-in used to run the unmodified reference
generators when building oracle/_ref.
"""
from __future__ import annotations


def _runs(e: int):
    """Maximal runs of ones of e, most significant first: [(low_bit_pos, length)]."""
    out = []
    i = 0
    while e >> i:
        if (e >> i) & 1:
            j = i
            while (e >> j) & 1:
                j += 1
            out.append((i, j - i))
            i = j
        else:
            i += 1
    return out[::-1]


def _length_chain(lengths, have=None):
    """Addition chain on run lengths: [(L, a, b)] with L=a+b, a>=b; 1 is implicit.

    Halving ladder (L = 2*(L/2), or (L-1)+1 when odd) that reuses any pair of
    lengths already built.  Building x_L this way costs L-1 squarings in total.
    """
    have = {1} if have is None else have
    steps = []

    def need(L):
        if L in have:
            return
        pair = next(((a, L - a) for a in sorted(have, reverse=True)
                     if 2 * a >= L and (L - a) in have), None)
        if pair is None:
            if L % 2 == 0:
                need(L // 2)
                pair = (L // 2, L // 2)
            else:
                need(L - 1)
                pair = (L - 1, 1)
        have.add(L)
        steps.append((L, pair[0], pair[1]))

    for L in sorted(set(lengths)):
        need(L)
    return steps


def _pieces(L, have):
    """Greedy split of a run of L ones into already-built run lengths."""
    out = []
    while L:
        a = max(h for h in have if h <= L)
        out.append(a)
        L -= a
    return out


def find_chain(e: int):
    """Return a program computing z = x^e (e >= 1)."""
    assert e >= 1
    runs = _runs(e)
    have = {1}
    # The top run is built with the halving ladder: its squarings are the
    # exponent's own leading bits, so they are not overhead.
    steps = _length_chain([runs[0][1]], have)
    # A later run is appended piecewise from lengths already built (one multiply
    # per piece, no extra squarings); only if that would take many pieces is its
    # length built separately.
    for _, L in runs[1:]:
        if len(_pieces(L, have)) > 8:
            steps += _length_chain([L], have)
    prog = []
    name = {1: "x"}
    for (L, a, b) in steps:
        dst = f"r{L}"
        name[L] = dst
        prog.append(("sqr", dst, name[a], b))
        prog.append(("mul", dst, dst, name[b]))
    acc = name[runs[0][1]]
    pos = runs[0][0]                      # exponent of acc so far ends at bit `pos`
    for (rpos, L) in runs[1:]:
        top = rpos + L                    # run occupies bits [rpos, top)
        for a in _pieces(L, have):
            prog.append(("sqr", "z", acc, pos - (top - a)))
            prog.append(("mul", "z", "z", name[a]))
            acc = "z"
            top -= a
            pos = top
    if pos or acc != "z":
        prog.append(("sqr", "z", acc, pos))
    return _rename(prog)


def _rename(prog):
    """Allocate temporaries t0.. with liveness-based reuse; drop sqr-by-0 copies where possible."""
    # last use index of every variable
    last = {}
    for i, op in enumerate(prog):
        for v in op[2:4]:
            if isinstance(v, str):
                last[v] = i
    free, mapping, out = [], {"x": "x", "z": "z"}, []
    ntmp = 0
    for i, op in enumerate(prog):
        kind, dst = op[0], op[1]
        srcs = [v for v in op[2:4] if isinstance(v, str)]
        msrcs = [mapping[v] for v in srcs]
        if dst not in mapping:
            # a source dying here may donate its slot (ops are alias-safe: pseudo.py:1832-1845)
            donor = None
            for v in srcs:
                if last[v] == i and mapping[v].startswith("t") and v != dst:
                    donor = mapping[v]
                    break
            if donor is not None:
                mapping[dst] = donor
            elif free:
                mapping[dst] = free.pop()
            else:
                mapping[dst] = f"t{ntmp}"
                ntmp += 1
        mdst = mapping[dst]
        if kind == "sqr":
            out.append(("sqr", mdst, msrcs[0], op[3]))
        else:
            out.append(("mul", mdst, msrcs[0], msrcs[1]))
        for v in srcs:
            if last[v] == i and mapping[v].startswith("t") and mapping[v] != mdst:
                if mapping[v] not in free:
                    free.append(mapping[v])
    return out


def evaluate(prog, x: int, p: int) -> int:
    """Run a program on Python ints (used by tests and by the generators' self-check)."""
    env = {"x": x % p}
    for op in prog:
        if op[0] == "sqr":
            v = env[op[2]]
            for _ in range(op[3]):
                v = v * v % p
            env[op[1]] = v
        else:
            env[op[1]] = env[op[2]] * env[op[3]] % p
    return env["z"]


def cost(prog):
    """(squarings, multiplies) of a program."""
    s = sum(op[3] for op in prog if op[0] == "sqr")
    m = sum(1 for op in prog if op[0] == "mul")
    return s, m


def temporaries(prog):
    return sorted({v for op in prog for v in op[1:4] if isinstance(v, str) and v.startswith("t")},
                  key=lambda t: int(t[1:]))


def to_reference_text(prog) -> str:
    """`addchain gen` text as parsed by pseudo.py:759-783 (tmp/double/add/shift)."""
    lines = ["tmp " + " ".join(temporaries(prog))]
    for op in prog:
        if op[0] == "mul":
            lines.append(f"add {op[1]} {op[2]} {op[3]}")
        elif op[3] == 1:
            lines.append(f"double {op[1]} {op[2]}")
        else:
            # n == 0 is a copy: the reference emits modcpy then modnsqr(.,0)
            lines.append(f"shift {op[1]} {op[2]} {op[3]}")
    return "\n".join(lines) + "\n"
