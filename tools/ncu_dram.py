"""DRAM bytes per launch of a kernel class from an `ncu --page raw --csv` export -> profiles/ncu_conv_dram.json,
the file `bench.py` reads `roofline.traffic` from (so the number follows the kernels instead of a constant).

    python tools/ncu_dram.py gpurun_out/r02_conv_full_raw.csv --edge 512 --slices 74 --source profiles/r02_ncu_full_conv_pass.txt
"""
import argparse
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--pattern", default=r"conv_(row|tc|halo|chain)")
    ap.add_argument("--edge", type=int, default=512)
    ap.add_argument("--slices", type=int, default=74)
    ap.add_argument("--source", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_conv_dram.json"))
    args = ap.parse_args()
    rows = list(csv.reader(l for l in open(args.csv) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    total, n = 0.0, 0
    for r in rows[2:]:
        if not re.search(args.pattern, r[kn]):
            continue
        total += float(r[rd].replace(",", "")) * SCALE[units[rd]] + float(r[wr].replace(",", "")) * SCALE[units[wr]]
        n += 1
    rec = {"edge": args.edge, "slices_per_pass": args.slices, "launches": n, "dram_bytes_per_pass": total,
           "dram_bytes_per_launch": total / max(n, 1),
           "source": args.source or os.path.relpath(args.csv, ROOT),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum, mean over the conv-class launches of one network "
                   "pass (ncu --set full --clock-control none)"}
    json.dump(rec, open(args.out, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
