#!/usr/bin/env python3
"""Build tuning variants of libmodarith_b200.so side by side (kernel experiments).

    python tools/variants.py build  tag[:GENOPT=val,...][:-DMACRO=val,...] ...
    python tools/variants.py run    [bench args]      # on the GPU box: bench every built variant

Each variant is a full copy of csrc/ with regenerated field headers (generator options are
passed through environment variables read by gen/plan.py) compiled into
modarith_b200/build/variants/<tag>/libmodarith_b200.so.
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "modarith_b200", "build", "variants")


def build_variant(spec):
    parts = spec.split(":")
    tag = parts[0]
    genopts, defines = {}, []
    for p in parts[1:]:
        for item in p.split(","):
            if item.startswith("-D"):
                defines.append(item)
            elif item:
                k, v = item.split("=")
                genopts[k] = v
    d = os.path.join(VDIR, tag)
    shutil.rmtree(d, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "modarith_b200", "csrc"), os.path.join(d, "csrc"))
    os.makedirs(os.path.join(d, "include"), exist_ok=True)
    env = dict(os.environ, **genopts)
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from modarith_b200.gen.cli import generate\n"
            "from modarith_b200.primes import PRIMES\n"
            "import os\n"
            "for P in PRIMES.values(): generate(P, os.path.join(%r, 'csrc', 'gen', 'field_%%s.cuh' %% P.name), verbose=False)\n"
            % (ROOT, d))
    subprocess.check_call([sys.executable, "-c", code], env=env)
    from modarith_b200 import build as b
    objs = []
    procs = []
    for unit in b.UNITS:
        obj = os.path.join(d, unit.replace(".cu", ".o"))
        flags = [f for f in b.NVCC_FLAGS]
        flags[flags.index("-I") + 1] = os.path.join(d, "csrc")
        # csrc files include "../../include/modarith_b200.h" relative to themselves
        cmd = [b._nvcc()] + flags + defines + ["-Xptxas", "-v", "-c", os.path.join(d, "csrc", unit), "-o", obj]
        procs.append((unit, obj, subprocess.Popen(cmd, stdout=open(obj + ".log", "w"), stderr=subprocess.STDOUT)))
        objs.append(obj)
    for unit, obj, p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed for %s %s:\n%s" % (tag, unit, open(obj + ".log").read()[-3000:]))
    lib = os.path.join(d, "libmodarith_b200.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    regs = {}
    for line in open(os.path.join(d, "mab_capi_X25519.o.log")):
        if "Compiling entry function" in line:
            cur = line.split("'")[1]
        if "Used" in line and "registers" in line:
            regs[cur] = int(line.split("Used")[1].split()[0])
        if "bytes spill stores" in line and "k_rfc7748" in cur:
            regs[cur + ":spill"] = line.strip()
    lad = [v for k, v in regs.items() if "k_rfc7748" in k]
    print("built variant %-16s ladder regs/spill: %s" % (tag, lad))
    return lib


def run(argv):
    out = {}
    for tag in sorted(os.listdir(VDIR)):
        lib = os.path.join(VDIR, tag, "libmodarith_b200.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, MODARITH_B200_LIB=lib)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline"] + argv,
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            out[tag] = {"value": j["value"], "frac": j["roofline"]["frac"], "e2e": j["e2e"]["value"],
                        "parity": j["parity_spot_check"], "sm_mhz": (j.get("clocks") or {}).get("sm_mhz")}
        except Exception as e:
            out[tag] = {"error": (r.stderr or str(e))[-400:]}
        print(tag, out[tag], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for s in sys.argv[2:]:
            build_variant(s)
    else:
        run(sys.argv[2:])
