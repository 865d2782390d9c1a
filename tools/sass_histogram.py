"""SASS opcode histogram of the in-tree library, per kernel: the Blackwell-native evidence named in
`/opt/skills/guides/B200_PROFILING.md` (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, LDGSTS = cp.async; HMMA would be the legacy mma.sync path).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "interactive-unet_b200", "libiunet_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDGSTS", "SYNCS", "HMMA",
         "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name.replace("void iu::", ""))
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            per[name][m.group(1).split(".")[0]] += 1
    print(f"SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)\n")
    print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{w:>7s}" for w in WATCH))
    tot = collections.Counter()
    for k, c in per.items():
        print(f"{k[:58]:58s} {sum(c.values()):7d} " + " ".join(f"{c.get(w, 0):7d}" for w in WATCH))
        tot.update(c)
    print(f"{'TOTAL':58s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(w, 0):7d}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
