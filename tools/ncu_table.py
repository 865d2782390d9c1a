"""Per-kernel table from an `ncu --page raw --csv` export (development aid).

    python tools/ncu_table.py gpurun_out/r01_conv_full_raw.csv
"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum"),
    ("rdMB", "dram__bytes_read.sum"),
    ("wrMB", "dram__bytes_write.sum"),
    # share of the kernel's active cycles in which the fp16/bf16 tensor sub-pipe (tcgen05 kind::f16) was busy
    ("tens%", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("lts%", "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsuwf%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("regs", "launch__registers_per_thread"),
]


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = [(n, hdr.index(c) if c in hdr else None) for n, c in COLS]
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print("  # kernel                             grid " + " ".join(f"{n:>7s}" for n, _ in idx))
    tot = 0.0
    for k, r in enumerate(rows[2:]):
        name = re.sub(r"void iu::|\(.*", "", r[kn])[:30]
        vals = [fnum(r[i]) if i is not None else float("nan") for _, i in idx]
        for j, (n, i) in enumerate(idx):          # normalise units to us / MB
            if i is None:
                continue
            u = units[i]
            if n == "us":
                vals[j] *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            if n in ("rdMB", "wrMB"):
                vals[j] *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        tot += vals[0]
        print(f"{k:3d} {name:30s} {r[gs]:>10s} " + " ".join(f"{v:7.1f}" for v in vals))
    print(f"total {tot:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
