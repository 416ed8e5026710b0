#!/usr/bin/env python
"""Per-kernel SASS evidence for libbvc.so: counts of the mnemonics that prove the Blackwell-native paths
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
HMMA = legacy mma.sync, which must not appear).  Runs anywhere cuobjdump is installed (no GPU):
    python tools/sass_summary.py > profiles/rNN_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "baby-vision-curriculum_b200", "libbvc.so")
PATS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MUFU.EX2", "SYNCS", "UTCBAR", "REDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            cnt[cur]["instr"] += 1
            for p in PATS:
                if m.group(1).startswith(p):
                    cnt[cur][p] += 1
    names = list(cnt)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    print(f"# SASS mnemonic counts per kernel of libbvc.so (`cuobjdump -sass`, sm_100a; {len(names)} kernels)\n")
    print("UTCHMMA = `tcgen05.mma`, LDTM / STTM = `tcgen05.ld` / `.st`, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / "
          "reduce, SYNCS = mbarrier ops, UTCBAR = `tcgen05.commit`; HMMA (legacy `mma.sync`) must be absent.\n")
    print("| kernel | instructions | " + " | ".join(PATS) + " |")
    print("|---|---|" + "---|" * len(PATS))
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = cnt[n]
        tot.update(c)
        short = re.sub(r"\(.*", "", d).replace("void ", "").replace("bvc::", "")
        print("| `" + short + "` | " + str(c["instr"]) + " | " + " | ".join(str(c[p]) if c[p] else "" for p in PATS) + " |")
    print("| **total** | " + str(tot["instr"]) + " | " + " | ".join(str(tot[p]) for p in PATS) + " |")
    if tot["HMMA"]:
        sys.exit("legacy HMMA found")


if __name__ == "__main__":
    main()
