"""Top SASS instructions by warp-stall samples from an `ncu --page source --csv` export (development aid).

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [N]
"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    samp = col["# Samples"]
    stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_")]
    total = sum(int(r[samp] or 0) for r in data)
    print(rows[0][1][:110] if rows[0] else "")
    print(f"total samples {total}")
    ranked = sorted(enumerate(data), key=lambda kv: -int(kv[1][samp] or 0))[:top]
    for idx, r in ranked:
        n = int(r[samp] or 0)
        reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        why = ", ".join(f"{name} {v}" for v, name in reasons if v)
        print(f"{100.0 * n / max(total, 1):5.1f}%  #{idx:5d}  {r[col['Source']].strip()[:70]:70s}  {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
