"""Per-layer bound model of the conv stack against a measured launch list (development aid, runs on the CPU).

    python tools/layer_model.py [profiles/r02_launches_final.txt] [--ghz 1.65]

For each tensor-core conv launch of one pass (74 slices of 512 x 512, resnet34: 43 launches in round 1's lists, 41 since
decoder block 4 + head became one fused launch) it derives from the layer's shape and the kernel variant that runs it:

  t_mma   the MMA issue floor: number of tcgen05.mma (M = 128, K = 16) x the measured cycles per MMA of that N
          (`profiles/r01_pipe_probe.txt`: 42 cycles up to N = 32, N/2 from N = 128 on);
  t_smem  shared-memory traffic at 128 B/clk/SM: operand bytes the MMAs read (A 4 KB + B N*32 B each) plus the bytes
          TMA / cp.async write into the ring for them (finding 13 in `profiles/r01_findings.md`);
  t_hbm   algorithmic HBM bytes (inputs read once, halo rows of the row kernel included, output written once,
          residual read once) at the measured copy bandwidth;

and prints them beside the measured duration, with the largest of the three marked.  All per-SM figures assume the
work is spread evenly over 148 SMs.  It is a model, not a measurement: its use is to see which resource a layer is
closest to and how far the measured time is from that bound.
"""
import argparse
import re

SMS = 148
HBM = 6535.1e9
SMEM_BPC = 128.0
MMA_CYC = {16: 41.9, 32: 42.2, 48: 44.2, 64: 48.2, 96: 56.2, 128: 64.2, 192: 96.1, 256: 128.2}
BATCH = 74


def layers(r02=False):
    """(name, kind, out_hw, cout, [(cin, ksize, upsampled)], stride, residual)
    r02: the Cout-64 row layers add their shortcut in the epilogue (kind "rowepi": no identity K segment, the shortcut
    is read once from HBM) and decoder block 4 + head are one launch (their three models are summed by `main`)."""
    L = []
    for i in range(3):
        L.append((f"layer1.{i}.conv1", "row", 128, 64, [(64, 3, False)], 1, False))
        L.append((f"layer1.{i}.conv2", "rowepi" if r02 else "row", 128, 64, [(64, 3, False)], 1, True))
    for li, (c, hw, nb, kind) in enumerate([(128, 64, 4, "tap128x2"), (256, 32, 6, "tap256"), (512, 16, 3, "tap256")]):
        cin = c // 2
        L.append((f"layer{li + 2}.0.conv1 s2", kind, hw, c, [(cin, 3, False)], 2, False))
        L.append((f"layer{li + 2}.0.conv2+ds", kind, hw, c, [(c, 3, False), (cin, 1, False)], 1, False))
        for b in range(1, nb):
            L.append((f"layer{li + 2}.{b}.conv1", kind, hw, c, [(c, 3, False)], 1, False))
            L.append((f"layer{li + 2}.{b}.conv2", kind, hw, c, [(c, 3, False)], 1, True))
    L.append(("dec0.conv1", "halo", 32, 256, [(512, 3, True), (256, 3, False)], 1, False))
    L.append(("dec0.conv2", "tap256", 32, 256, [(256, 3, False)], 1, False))
    L.append(("dec1.conv1", "halo", 64, 128, [(256, 3, True), (128, 3, False)], 1, False))
    L.append(("dec1.conv2", "tap128x2", 64, 128, [(128, 3, False)], 1, False))
    L.append(("dec2.conv1", "row", 128, 64, [(128, 3, True), (64, 3, False)], 1, False))
    L.append(("dec2.conv2", "row", 128, 64, [(64, 3, False)], 1, False))
    L.append(("dec3.conv1", "row", 256, 32, [(64, 3, True), (64, 3, False)], 1, False))
    L.append(("dec3.conv2", "row", 256, 32, [(32, 3, False)], 1, False))
    L.append(("dec4.conv1", "row", 512, 16, [(32, 3, True)], 1, False))
    L.append(("dec4.conv2", "row", 512, 16, [(16, 3, False)], 1, False))
    L.append(("head", "row", 512, 16, [(16, 3, False)], 1, False))
    return L


def model(layer, ghz):
    name, kind, hw, cout, segs, stride, residual = layer
    px = BATCH * hw * hw
    mtiles = px / 128.0
    flops = sum(2.0 * px * k * k * cin * cout for cin, k, _ in segs)
    # ---- HBM: every input once (upsampled sources are stored at half resolution), output once, residual once
    out_bytes = px * (2 * 4 if name == "head" else cout * 2)
    in_bytes = 0.0
    for cin, k, up in segs:
        src_px = px / 4.0 if up else px * stride * stride
        halo = 1.0
        if kind in ("row", "rowepi"):
            rows = {64: 4, 32: 8, 16: 8}[cout]
            halo = (rows / 2 + 2) / (rows / 2) if up else (rows + 2) / rows      # input rows read per output row block
        in_bytes += src_px * cin * 2 * halo
    if residual:
        in_bytes += px * cout * 2
    w_bytes = sum(k * k * cin * cout * 2 for cin, k, _ in segs)
    hbm = in_bytes + out_bytes + w_bytes
    # ---- MMAs and shared-memory traffic
    n_mma = smem = cyc = 0.0
    if kind in ("row", "rowepi"):
        rows = {64: 4, 32: 8, 16: 8}[cout]
        for cin, k, up in segs:
            n = (4 if up else 3) * cout                      # vertical taps folded into N
            in_rows = (rows / 2 + 2) if up else (rows + 2)   # MMAs per row block and kx and 16 channels
            m = mtiles / rows * in_rows * 3 * (cin / 16.0)
            n_mma += m
            cyc += m * MMA_CYC[min(MMA_CYC, key=lambda v: abs(v - n))]
            smem += m * (4096 + n * 32)
        if residual and kind == "row":
            m = mtiles * (cout / 16.0)                       # identity K segment
            n_mma += m
            cyc += m * MMA_CYC[cout if cout in MMA_CYC else 64]
            smem += m * (4096 + cout * 32)
        smem += in_bytes                                      # gather writes (weights resident / streamed: small)
    else:
        n = 256 if kind == "tap256" else 128
        ksteps = sum(k * k * cin / 16.0 for cin, k, _ in segs)
        if kind == "halo":
            n = 128
        ntiles_n = cout / n
        m = mtiles * ksteps * ntiles_n
        n_mma = m
        cyc = m * MMA_CYC[n]
        smem = m * (4096 + n * 32)
        # ring fills per 64-channel chunk: A box 16 KB per M tile, B box n*128 B shared by the M tiles of the CTA tile
        chunks = sum(k * k * cin / 64.0 for cin, k, _ in segs)
        if kind == "halo":
            fills = mtiles / 2 * sum(cin / 64.0 * (18 * 18 * 128 + 9 * n * 128) for cin, _, _ in segs)
        else:
            share = 2.0 if kind == "tap128x2" else 1.0
            fills = mtiles * ntiles_n * chunks * (16384 + n * 128 / share)
        smem += fills
    f = ghz * 1e9
    return dict(name=name, kind=kind, gflop=flops / 1e9, t_mma=cyc / SMS / f * 1e6, t_smem=smem / SMS / SMEM_BPC / f * 1e6,
                t_hbm=hbm / HBM * 1e6, n_mma=n_mma)


def measured(path):
    out = []
    for line in open(path):
        m = re.search(r"(conv_\w+_kernel).*?([\d.]+) us\s*$", line)
        if m and "stem" not in m.group(1):
            out.append(float(m.group(2)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("launches", nargs="?", default="profiles/r02_launches_final.txt")
    ap.add_argument("--ghz", type=float, default=1.65, help="SM clock under the power cap")
    args = ap.parse_args()
    meas = measured(args.launches)
    r02 = len(meas) == 41
    rows = [model(l, args.ghz) for l in layers(r02)]
    if r02:
        # the fused tail: one launch; MMA and shared-memory work of its three layers add up, the two 16-channel
        # intermediates never reach HBM (input of conv1 + output of the head + weights remain)
        tail = rows[-3:]
        px = BATCH * 512 * 512
        fused = dict(name="dec4.conv1+conv2+head", kind="chain", gflop=sum(r["gflop"] for r in tail),
                     t_mma=sum(r["t_mma"] for r in tail) * 128 / 124, t_smem=sum(r["t_smem"] for r in tail) * 128 / 124,
                     t_hbm=(px / 4.0 * 32 * 2 * 10 / 8 + px * 2 * 4) / HBM * 1e6, n_mma=sum(r["n_mma"] for r in tail))
        rows = rows[:-3] + [fused]
    assert len(meas) == len(rows) and len(rows) in (41, 43), (len(meas), len(rows))
    print(f"one pass of {BATCH} slices of 512x512 at {args.ghz} GHz; times in us; '*' marks the model's largest bound")
    print(f"{'layer':22s} {'kernel':9s} {'GFLOP':>7s} {'t_mma':>7s} {'t_smem':>7s} {'t_hbm':>7s} {'meas':>7s} {'meas/bound':>10s}")
    tot = dict(t_mma=0.0, t_smem=0.0, t_hbm=0.0, bound=0.0, meas=0.0)
    for r, t in zip(rows, meas):
        bound = max(r["t_mma"], r["t_smem"], r["t_hbm"])
        mark = {k: ("*" if r[k] == bound else " ") for k in ("t_mma", "t_smem", "t_hbm")}
        print(f"{r['name']:22s} {r['kind']:9s} {r['gflop']:7.1f} {r['t_mma']:6.1f}{mark['t_mma']} {r['t_smem']:6.1f}{mark['t_smem']} "
              f"{r['t_hbm']:6.1f}{mark['t_hbm']} {t:7.1f} {t / bound:10.2f}")
        for k in ("t_mma", "t_smem", "t_hbm"):
            tot[k] += r[k]
        tot["bound"] += bound
        tot["meas"] += t
    print(f"{'total':22s} {'':9s} {sum(r['gflop'] for r in rows):7.1f} {tot['t_mma']:7.1f} {tot['t_smem']:7.1f} {tot['t_hbm']:7.1f} "
          f"{tot['meas']:7.1f} {tot['meas'] / tot['bound']:10.2f}")


if __name__ == "__main__":
    main()
