#!/usr/bin/env python3
"""Build oracle/_ref: the reference's OWN 64-bit C for the hot path, compiled here.

TEST INFRASTRUCTURE ONLY.  Runs the unmodified reference generators from
/root/reference (read-only) inside a scratch directory, with three edits applied
to SCRATCH COPIES of the scripts' module-level settings:

  cyclescounter=False   libcpucycles is not installed (pseudo.py:29)
  generic=False         what rfc7748.c:20 asks for (lazy add/sub, pseudo.py:1523-1528)
  PSCR=False            plain mask-XOR cswap; removes the `static R` data race when
                        the batch loop runs on many threads (pseudo.py:986-1005)

and with oracle/addchain_standin.py on PATH as `addchain` (the Go tool is absent).
The generated field.c is pasted into a scratch copy of rfc7748.c at its marker
(rfc7748.c:24-28), `#define COUNT_CLOCKS` (rfc7748.c:30) is commented out, our
ref_shim.c is appended, and the whole unit is compiled with gcc into

  oracle/_ref/libref_X25519.so          pseudo.py 64 X25519  + rfc7748.c   (generic=False)
  oracle/_ref/libref_X448.so            monty.py  64 X448    + rfc7748.c   (generic=False)
  oracle/_ref/libref_X25519_generic.so  pseudo.py 64 X25519  field only    (generic=True)
  oracle/_ref/libref_X448_generic.so    monty.py  64 X448    field only    (generic=True)
  oracle/_ref/libref_NIST256.so         monty.py  64 NIST256 field only    (generic=True)

Nothing from the reference is written into the repository outside oracle/_ref,
which is git-ignored (it still travels to the GPU box with the snapshot).
"""
import os
import re
import shutil
import stat
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MODARITH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

TARGETS = [
    # (name, script, prime argument, generic flag, has rfc7748)
    ("X25519", "pseudo.py", "X25519", False, True),          # ladder build: lazy add/sub
    # the same driver with `#define TWIST_SECURE` removed: the point-validation tail rfc7748.c:228-251
    # (generic=True: that tail calls modadd/modfsb/modshr on values the lazy add would leave unreduced)
    ("X25519_validate", "pseudo.py", "X25519", True, True),
    ("X448_validate", "monty.py", "X448", True, True),
    ("X448", "monty.py", "X448", False, True),
    # field builds (generic=True, the scripts' default): modadd/modsub reduce to < 2p, so a
    # single modexp after them is canonical -- these back the field-level golden vectors
    ("X25519_generic", "pseudo.py", "X25519", True, False),
    ("X448_generic", "monty.py", "X448", True, False),
    ("NIST256", "monty.py", "NIST256", True, False),
    # outside the hot path: an unshaped field prime and a group order ("00<decimal>" -> group.c,
    # monty.py:2110-2127); they back the vectors for the generator's full-Montgomery fall-back plan
    ("SECP256K1", "monty.py", "SECP256K1", True, False),
    ("NIST256ORDER", "monty.py", "00115792089210356248762697446949407573529996955224135760342422259061068512044369", True, False),
    # a modulus that is NOT compiled into the shipped library: backs the test of the add-on build
    # (python -m modarith_b200.build --prime NIST384; monty.py named table)
    ("NIST384", "monty.py", "NIST384", True, False),
    # 2^n - c with n not a multiple of 32 (the generator's bit-level pseudo-Mersenne plan): 2^414 - 17 on 13 limbs with
    # two spare bits, 2^521 - 1 on 17 limbs with 23 (pseudo.py named table)
    ("C41417", "pseudo.py", "C41417", True, False),
    ("NIST521", "pseudo.py", "NIST521", True, False),
    # p = -1 (mod 2^(32z)): the generator's Montgomery-friendly plan (5*2^248 - 1: z = 7 of 8 words)
    ("ED248", "monty.py", "ED248", True, False),
]
# a user-defined Montgomery curve (y^2 = x^3 + 2065150 x^2 + x over 2^383 - 187, a24 = (A - 2) / 4, cofactor 8, base
# point u = 12): backs the test of an add-on library WITH a ladder (python -m modarith_b200.build --prime M383=... --a24 ...)
M383 = (516287, 3, 12)
CURVE_TARGETS = [
    ("M383", "pseudo.py", "PM383", False, True, M383),
    ("M383_validate", "pseudo.py", "PM383", True, True, M383),
]


def _patch_settings(text, generic):
    def sub(name, value, t):
        new, n = re.subn(r"(?m)^%s=\w+" % name, "%s=%s" % (name, value), t, count=1)
        assert n == 1, name
        return new
    text = sub("cyclescounter", "False", text)
    text = sub("generic", "True" if generic else "False", text)
    text = sub("PSCR", "False", text)
    return text


def build_one(name, script, prime, generic, ladder, cflags, curve=None):
    """curve = (a24, cof, generator): a user-defined Montgomery curve -- its constants are added to the "Describe
    Montgomery Curve parameters" section of rfc7748.c exactly as its comment asks the user to (rfc7748.c:117-132)."""
    work = tempfile.mkdtemp(prefix="mab_ref_%s_" % name)
    try:
        bindir = os.path.join(work, "bin")
        os.mkdir(bindir)
        wrapper = os.path.join(bindir, "addchain")
        with open(wrapper, "w") as f:
            f.write("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(HERE, "addchain_standin.py")))
        os.chmod(wrapper, os.stat(wrapper).st_mode | stat.S_IEXEC)
        gen = os.path.join(work, script)
        with open(os.path.join(REF, script)) as f:
            src = f.read()
        with open(gen, "w") as f:
            f.write(_patch_settings(src, generic))
        env = dict(os.environ, PATH=bindir + os.pathsep + os.environ["PATH"])
        r = subprocess.run([sys.executable, gen, "64", prime], cwd=work, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log = r.stdout
        if "Passed - OK" not in log:
            raise RuntimeError("reference generator self-test did not pass for %s:\n%s" % (name, log))
        fname = "group.c" if prime.startswith("00") else "field.c"
        with open(os.path.join(work, fname)) as f:
            field = f.read()
        unit = field
        if ladder:
            with open(os.path.join(REF, "rfc7748.c")) as f:
                drv = f.read()
            drv = drv.replace("#define COUNT_CLOCKS", "//#define COUNT_CLOCKS", 1)
            if curve is not None:
                anchor = "// Describe Montgomery Curve parameters"
                assert drv.count(anchor) == 1
                drv = drv.replace(anchor, anchor + "\n#define A24 %d\n#define COF %d\n#define GENERATOR %d\n#define TWIST_SECURE\n"
                                  % tuple(curve))
            if name.endswith("_validate"):
                assert drv.count("#define TWIST_SECURE") == (3 if curve is not None else 2)
                drv = drv.replace("#define TWIST_SECURE", "//#define TWIST_SECURE")
            marker = "/*** Insert automatically generated code for modulus field.c here ***/"
            assert marker in drv
            unit = drv.replace(marker, marker + "\n" + field, 1)
        with open(os.path.join(HERE, "ref_shim.c")) as f:
            unit += "\n" + f.read()
        os.makedirs(OUT, exist_ok=True)
        csrc = os.path.join(OUT, "ref_%s.c" % name)
        with open(csrc, "w") as f:
            f.write(unit)
        so = os.path.join(OUT, "libref_%s.so" % name)
        cmd = ["gcc"] + cflags + ["-shared", "-fPIC", "-fopenmp", "-Dmain=ref_main", "-w"]
        if ladder:
            cmd.append("-DREF_HAS_RFC7748")
        cmd += ["-o", so, csrc]
        subprocess.check_call(cmd)
        # the timing binary's checksums are golden vectors (pseudo.py:1862-1866)
        timelog = ""
        tbin = os.path.join(work, "time")
        if os.path.exists(tbin) and os.environ.get("MODARITH_REF_TIME", "0") == "1":
            timelog = subprocess.run([tbin], stdout=subprocess.PIPE, text=True).stdout
        with open(os.path.join(OUT, "build_%s.log" % name), "w") as f:
            f.write(log + "\n" + timelog)
        return so
    finally:
        shutil.rmtree(work, ignore_errors=True)


def build_curve(cflags, curve="NIST256", template="weierstrass.c"):
    """oracle/_ref/libref_<curve>_curve.so: the reference's weierstrass.c / edwards.c as its own curve.py
    patches it (curve.py:85-94,157-166,335-351 -- it runs pseudo.py/monty.py for the field and the group order
    and pastes field.c / curve.c / point.h into the template and curve.h in the working directory)."""
    work = tempfile.mkdtemp(prefix="mab_ref_curve_")
    try:
        bindir = os.path.join(work, "bin")
        os.mkdir(bindir)
        wrapper = os.path.join(bindir, "addchain")
        with open(wrapper, "w") as f:
            f.write("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(HERE, "addchain_standin.py")))
        os.chmod(wrapper, os.stat(wrapper).st_mode | stat.S_IEXEC)
        py3 = os.path.join(bindir, "python3")
        with open(py3, "w") as f:
            f.write("#!/bin/sh\nexec %s \"$@\"\n" % sys.executable)
        os.chmod(py3, os.stat(py3).st_mode | stat.S_IEXEC)
        for fn in ("curve.py", template, "curve.h", "testcurve.c"):
            shutil.copy(os.path.join(REF, fn), os.path.join(work, fn))
        for gen in ("monty.py", "pseudo.py"):
            with open(os.path.join(REF, gen)) as f:
                src = f.read()
            with open(os.path.join(work, gen), "w") as f:
                f.write(_patch_settings(src, False))
        env = dict(os.environ, PATH=bindir + os.pathsep + os.environ["PATH"])
        r = subprocess.run([sys.executable, "curve.py", "64", curve], cwd=work, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if "Passed - OK" not in r.stdout:
            raise RuntimeError("curve.py did not complete:\n" + r.stdout[-3000:])
        with open(os.path.join(work, template)) as f:
            unit = f.read()
        with open(os.path.join(work, "curve.h")) as f:
            hdr = f.read()
        unit = unit.replace('#include "curve.h"', hdr)
        with open(os.path.join(HERE, "ref_curve_shim.c")) as f:
            unit += "\n" + f.read().replace("ecn_nist256_", "ecn_%s_" % curve.lower())
        os.makedirs(OUT, exist_ok=True)
        csrc = os.path.join(OUT, "ref_%s_curve.c" % curve)
        with open(csrc, "w") as f:
            f.write(unit)
        so = os.path.join(OUT, "libref_%s_curve.so" % curve)
        subprocess.check_call(["gcc"] + cflags + ["-shared", "-fPIC", "-fopenmp", "-w", "-o", so, csrc])
        return so
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    if not os.path.isdir(REF):
        print("reference tree %s not present: keeping any prebuilt oracle/_ref" % REF)
        return 0
    # x86-64-v3 (AVX2/BMI2/ADX-era) keeps the .so runnable on the GPU box's host
    cflags = os.environ.get("MODARITH_REF_CFLAGS", "-O3 -march=x86-64-v3").split()
    for t in TARGETS:
        so = build_one(*t, cflags)
        print("built", so)
    for t in CURVE_TARGETS:
        print("built", build_one(*t[:5], cflags, curve=t[5]))
    print("built", build_curve(cflags, "NIST256", "weierstrass.c"))
    print("built", build_curve(cflags, "ED25519", "edwards.c"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
