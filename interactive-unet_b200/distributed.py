"""z-slab sharded volume prediction across the GPUs of one node (one process per GPU).

The reference only sketches multi-GPU prediction in commented-out code (`predict.py:137-147,204-232`:
one whole block per GPU, results gathered on the host).  Here the partition follows SURVEY.md
section 8e: the unit of work is a 2-D slice, so every rank runs N/G slices per axis, and only the
cross-axis average couples ranks.

    rank r owns the z-slab  z in [r*T, (r+1)*T),  T = N / G   (its part of the OUTPUT)
    axis 0 (slices indexed by z): rank r runs its own slab's slices -> no exchange
    axis 1 (indexed by y, image (z,x)) and axis 2 (indexed by x, image (z,y)): rank r runs the slices
        y (resp. x) in [r*T, (r+1)*T); each image row is a z, so the engine's head epilogue writes the
        probabilities destination-major  [dest rank h][slice][z in slab h][col][C]  (row_block = T) and ONE
        all-to-all per off-slab axis delivers to rank h exactly its slab: [y or x][z local][col][C].
    then K4 (reduce / blend / quantise / argmax) runs locally on each slab.

The per-voxel arithmetic and its order (axis order of `axes`) are identical to the single-GPU path, so
the sharded result is bit-identical to it.  The exchange is `torch.distributed.all_to_all_single`
(NCCL over NVLink on GPUs; gloo in the CPU tests), T*T*N*C*4 bytes per ordered pair per axis.
"""
import torch
import torch.distributed as dist


def _all_to_all(recv, send, group):
    """all_to_all_single, with a send/recv fallback for backends that lack it."""
    try:
        dist.all_to_all_single(recv, send, group=group)
        return
    except (RuntimeError, NotImplementedError):
        pass
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sends = list(send.chunk(world))
    recvs = list(recv.chunk(world))
    recvs[rank].copy_(sends[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        ops.append(dist.P2POp(dist.isend, sends[peer], dist.get_global_rank(group, peer) if group else peer, group))
        ops.append(dist.P2POp(dist.irecv, recvs[peer], dist.get_global_rank(group, peer) if group else peer, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def volumes_for_rank(files, group=None):
    """Whole volumes are independent objects (`predict.py:163`: one loop iteration per store), so under `torchrun`
    `predict_volumes` gives rank r the files r, r + G, r + 2G, ... of the sorted list and needs no collective; a single
    process (or an uninitialised process group) keeps them all."""
    files = list(files)
    if not (dist.is_available() and dist.is_initialized()):
        return files
    return files[dist.get_rank(group)::dist.get_world_size(group)]


def predict_volume_sharded(engine, volume, axes=(0, 1, 2), window=None, want_u8=True, want_labels=True,
                           want_mean=False, group=None):
    """Predict this rank's z-slab of a replicated cubic volume.

    engine : an `Engine` (or any object with `predict_axis`, `reduce`, `num_classes`, `device`)
    volume : the WHOLE uint8 / fp32 volume `[N,N,N]`, identical on every rank (device tensor or numpy)
    Returns dict(z0, t, u8=[T,N,N,C] uint8, labels=[T,N,N] uint8, mean=[T,N,N,C] fp32) of device tensors.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = volume.shape[0]
    if n % world:
        raise ValueError(f"volume edge {n} is not divisible by the number of ranks {world}")
    t = n // world
    z0 = rank * t
    c = engine.num_classes
    dev = engine.device
    probs = {}
    for axis in axes:
        if axis == 0:
            probs[0] = engine.predict_axis(volume, 0, slice_begin=z0, slice_count=t)
        else:
            send = torch.empty((world, t, t, n, c), dtype=torch.float32, device=dev)
            engine.predict_axis(volume, axis, slice_begin=z0, slice_count=t, out=send, slice_total=t, row_block=t)
            recv = torch.empty_like(send)
            if world > 1:
                _all_to_all(recv.view(-1), send.view(-1), group)
            else:
                recv = send
            probs[axis] = recv           # [source rank g][slice in strip g][z local][col][C] == [y|x][z local][col][C]
    out = dict(z0=z0, t=t)
    out["u8"] = torch.empty((t, n, n, c), dtype=torch.uint8, device=dev) if want_u8 else None
    out["labels"] = torch.empty((t, n, n), dtype=torch.uint8, device=dev) if want_labels else None
    out["mean"] = torch.empty((t, n, n, c), dtype=torch.float32, device=dev) if want_mean else None
    engine.reduce(probs, list(axes), n, t=t, z0=z0, window=window, out_u8=out["u8"], out_labels=out["labels"],
                  out_mean=out["mean"])
    return out


def gather_slabs(slab, group=None, dst=0):
    """Concatenate every rank's slab tensor along z on rank `dst` (None elsewhere)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bufs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, bufs, dst=dst, group=group)
    return torch.cat(bufs, dim=0) if rank == dst else None
