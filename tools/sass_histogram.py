#!/usr/bin/env python3
"""Opcode histograms of the built kernels (cuobjdump -sass of the objects the Makefile produced): the evidence that
the hot kernels are Blackwell-native (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) without having to disassemble the git-ignored objects.
usage: python tools/sass_histogram.py            # writes profiles/sass_<kernel>.txt for every entry function"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kaldi-hmm-gmm_b200", "csrc")
OUT = os.path.join(ROOT, "profiles")
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UTCATOMSWS", "MUFU", "FFMA2", "FADD2", "FMNMX3",
       "HMMA", "LDGSTS", "LDSM", "RED", "ATOM", "REDUX", "LDS", "STS", "LDG", "STG", "BAR")


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    objs = sorted(f for f in os.listdir(CSRC) if f.endswith(".o"))
    index = []
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
        fn, hist = None, None
        funcs = []
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                fn, hist = m.group(1), collections.Counter()
                funcs.append((fn, hist))
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
            if m and hist is not None:
                hist[m.group(1)] += 1
        for fn, hist in funcs:
            nice = demangle(fn)
            plain = nice.replace("khg::", "").replace("void ", "").replace("(bool)1", "1").replace("(bool)0", "0")
            short = re.sub(r"[^A-Za-z0-9_]+", "_", re.sub(r"\(.*", "", plain))[:80].strip("_")
            total = sum(hist.values())
            by_base = collections.Counter()
            for op, n in hist.items():
                by_base[op.split(".")[0]] += n
            path = os.path.join(OUT, f"sass_{short}.txt")
            with open(path, "w") as f:
                f.write(f"# {obj}: {nice}\n# cuobjdump -sass, sm_100a; {total} instructions\n")
                f.write("# Blackwell-native opcodes: " + ", ".join(f"{k}={by_base[k]}" for k in KEY if by_base[k]) + "\n")
                for op, n in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0])):
                    f.write(f"{n:7d}  {op}\n")
            index.append((short, total, {k: by_base[k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "MUFU") if by_base[k]}))
    with open(os.path.join(OUT, "sass_INDEX.txt"), "w") as f:
        f.write("# one line per kernel: instructions, Blackwell-native opcode counts (tools/sass_histogram.py)\n")
        for short, total, k in sorted(index, key=lambda x: x[0]):
            f.write(f"{short:82s} {total:7d}  {k}\n")
    print(f"wrote {len(index)} histograms to {OUT}")


if __name__ == "__main__":
    sys.exit(main())
