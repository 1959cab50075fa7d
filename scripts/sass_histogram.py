#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (no GPU needed): `cuobjdump -sass libavcer_b200.so`, grouped by
kernel, counting the mnemonics that prove the Blackwell data path (tcgen05 MMA = UTCHMMA, TMA loads / stores = UTMALDG /
UTMASTG, bulk copies = UBLKCP, TMEM loads = LDTM, tcgen05 commit barriers = UTCBAR, mbarrier ops = SYNCS, classic tensor
ops = HMMA, FP64 = DADD/DMUL/DSETP) plus the total instruction count.

    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "avcer_b200", "libavcer_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "LDGSTS", "LDSM",
        "MUFU", "DADD", "DMUL", "DFMA", "DSETP", "FFMA2", "FADD2", "FMUL2", "BAR", "ACQBULK", "CCTL", "ERRBAR", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op, mods = m.group(1), m.group(2)
            if op in KEYS:
                cur[op] += 1
                if op in ("UTMALDG", "UTMASTG", "UTCBAR", "UTCHMMA") and mods:
                    cur[op + mods] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} ({len(kernels)} kernels), sm_100a")
    for (name, c), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"\(.*", "", pretty)
        items = ", ".join(f"{k}={v}" for k, v in sorted(c.items()) if k != "_total")
        print(f"{pretty}\n    instructions={c['_total']}" + (f"; {items}" if items else ""))
        tot.update(c)
    print("# totals: " + ", ".join(f"{k}={v}" for k, v in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
