#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libruviii_b200.so (cuobjdump -sass): which kernels issue FP64 tensor instructions (DMMA),
cp.async (LDGSTS), TMA (UTMALDG), mbarrier (SYNCS), cluster / grid synchronisation, and -- for the record -- that no tcgen05
(UTC*MMA / LDTM) instruction exists: tcgen05 has no f64 kind, so the FP64 Gram runs on mma.sync DMMA.

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "ruviii_b200", "lib", "libruviii_b200.so")
WATCH = ["DMMA", "DFMA", "LDGSTS", "UTMALDG", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "ACQBULK", "USETMAXREG", "MUFU.RSQ64H",
         "RED.E.ADD", "ATOMG", "CCTL.IVALL", "MEMBAR", "SHFL", "LDG", "STG", "BAR.SYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("ruviii::", "")
            name = re.sub(r"^void ", "", name)
            name = re.sub(r"\((?!.*<).*$", "", name) if "<" not in name else re.sub(r">\(.*$", ">", name)
            name = re.sub(r"\(.*$", "", name) if name.count("(") and "<" not in name else name
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.endswith("64H") and op.startswith(w)):
                    cur[w] += 1
    print("SASS summary of %s (sm_100a), %d kernels" % (os.path.relpath(LIB, ROOT), len(kernels)))
    print("columns: instructions total | " + " ".join(WATCH))
    tot = collections.Counter()
    for name, c in kernels.items():
        tot.update(c)
        print("%-64s %7d | %s" % (name[:64], c["_total"], " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])))
    print("%-64s %7d | %s" % ("TOTAL", tot["_total"], " ".join("%s=%d" % (w, tot[w]) for w in WATCH)))
    print("tcgen05 (UTC*MMA / LDTM) instructions: %d -- expected 0: tcgen05.mma has no f64 kind" %
          (tot["UTCHMMA"] + tot["UTCQMMA"] + tot["UTCIMMA"] + tot["LDTM"]))


if __name__ == "__main__":
    sys.exit(main())
