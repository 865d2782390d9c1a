"""Print the per-launch durations of one network batch from an ncu `gpu__time_duration` CSV."""
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    return [(r["Kernel Name"], r["Grid Size"], float(r["Metric Value"].replace(",", ""))) for r in rows]


def one_batch(names, which=1):
    """Launches of network pass `which` (0-based): from its stem kernel up to the next pass's stem (or the end)."""
    idx = [i for i, (n, g, t) in enumerate(names) if "stem" in n]
    which = min(which, len(idx) - 1)
    end = idx[which + 1] if which + 1 < len(idx) else len(names)
    return names[idx[which]:end]


if __name__ == "__main__":
    cur = one_batch(load(sys.argv[1]))
    old = one_batch(load(sys.argv[2])) if len(sys.argv) > 2 else None
    tot = 0
    for k, (n, g, t) in enumerate(cur):
        short = re.sub(r"void iu::|\(.*", "", n)
        tot += t
        extra = f"   (before: {re.sub(r'void iu::|<.*', '', old[k][0])[:12]:12s} {old[k][2] / 1e3:8.1f})" if old else ""
        print(f"{k:3d} {short:36s} grid {g:14s} {t / 1e3:9.1f} us{extra}")
    print("batch total us", tot / 1e3, "" if not old else f"(before {sum(t for _, _, t in old) / 1e3:.1f})")
