"""Multi-GPU check (torchrun, one rank per GPU): the z-slab sharded prediction over NCCL must equal the
single-GPU prediction of the same volume bit for bit, for 2 and 4 classes, with the input
given as this rank's slab (strips exchanged between the ranks) and as the replicated volume, and with the probability
exchange in one chunk and in several ragged chunks.  Rank 0 prints one line per case and the verdict.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactive_unet_b200 as iu  # noqa: E402
from interactive_unet_b200 import distributed as iud  # noqa: E402
from oracle import synth  # noqa: E402  (seeded synthetic weights / volume only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for n, c, max_batch in ((128, 2, 0), (256, 4, 0), (256, 2, 24), (256, 4, 7)):
        if n % world or n // world < 8:
            continue
        ref = synth.make_model(c)
        model = iu.UNet(num_classes=c)
        model.load_state_dict(ref.state_dict())
        model = model.to(dev).eval()
        eng = model.engine()
        vol = torch.from_numpy(synth.blob_volume(n, 21)[0]).to(dev)
        window = iu.gaussian_window_1d(n)
        t = n // world
        want_u8 = torch.empty((n, n, n, c), dtype=torch.uint8, device=dev)
        want_lab = torch.empty((n, n, n), dtype=torch.uint8, device=dev)
        eng.predict_volume(vol, axes=(0, 1, 2), window=window, out_u8=want_u8, out_labels=want_lab)
        for mode in ("slab", "volume"):
            src = dict(slab=vol[rank * t:(rank + 1) * t].clone()) if mode == "slab" else dict(volume=vol)
            with eng.limit_batch(max_batch):          # max_batch slices per network pass = per exchange chunk
                res = iud.predict_volume_sharded(eng, axes=(0, 1, 2), window=window, **src)
            z0 = res["z0"]
            same = torch.equal(res["u8"], want_u8[z0:z0 + t]) and torch.equal(res["labels"], want_lab[z0:z0 + t])
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            full = iud.gather_slabs(res["u8"], dst=0)
            if rank == 0:
                whole = torch.equal(full, want_u8)
                chunks = -(-t // eng.auto_batch(n, n, t)) if not max_batch else -(-t // min(max_batch, t))
                print(f"edge {n} classes {c} world {world} input {mode} exchange chunks {chunks}: slabs bit-identical on "
                      f"every rank = {bool(flag.item())}, gathered volume identical = {whole}", flush=True)
                ok = ok and bool(flag.item()) and whole
        # end-to-end form: pinned host slab in, pinned host results out, axis 0 reduced and copied out part by part
        h_u8 = torch.empty((t, n, n, c), dtype=torch.uint8).pin_memory()
        h_lab = torch.empty((t, n, n), dtype=torch.uint8).pin_memory()
        with eng.limit_batch(max_batch):
            iud.predict_slab_from_host(eng, vol[rank * t:(rank + 1) * t].cpu().pin_memory(), axes=(0, 1, 2), window=window,
                                       out_u8=h_u8, out_labels=h_lab)
        same = torch.equal(h_u8, want_u8[rank * t:(rank + 1) * t].cpu()) and \
            torch.equal(h_lab, want_lab[rank * t:(rank + 1) * t].cpu())
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"edge {n} classes {c} world {world} host slab -> host results (pipelined tail): bit-identical on "
                  f"every rank = {bool(flag.item())}", flush=True)
            ok = ok and bool(flag.item())
    if rank == 0:
        print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
