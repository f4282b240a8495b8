#!/usr/bin/env python3
"""Stand-in for the external `addchain` tool the reference generators shell out to.

TEST INFRASTRUCTURE ONLY (part of oracle/): it lets the UNMODIFIED reference
scripts run offline when oracle/build_ref.py builds oracle/_ref.  The real tool
(github.com/mmcloughlin/addchain, Go, not vendored, no version pinned by the
reference: README.md:19-23, pseudo.py:1582-1586, monty.py:2166-2170) is absent
and cannot be installed.  The chain only decides the order of squarings and
multiplies inside modpro; no redc/modexp result depends on it.

  addchain search <N>   -> two lines that survive pseudo.py:51-121 `remove_unused`
  addchain gen <file>   -> tmp/double/add/shift program (pseudo.py:759-783)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from modarith_b200 import addchain as ac  # noqa: E402


def main(argv):
    if len(argv) >= 3 and argv[1] == "search":
        n = int(argv[2])
        sys.stdout.write("e = 2*1\nreturn e + %d\n" % (n - 2))
        return 0
    if len(argv) >= 3 and argv[1] == "gen":
        toks = open(argv[2]).read().split()
        n = int(toks[-1]) + 2
        sys.stdout.write(ac.to_reference_text(ac.find_chain(n)))
        return 0
    sys.stderr.write("addchain stand-in: unsupported invocation %r\n" % (argv[1:],))
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv))
