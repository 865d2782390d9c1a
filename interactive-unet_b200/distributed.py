"""z-slab sharded volume prediction across the GPUs of one node (one process per GPU).

The reference only sketches multi-GPU prediction in commented-out code (`predict.py:137-147,204-232`:
one whole block per GPU, results gathered on the host).  Here the partition follows SURVEY.md
section 8e: the unit of work is a 2-D slice, so every rank runs N/G slices per axis, and only the
cross-axis average couples ranks.

    rank r owns the z-slab  z in [r*T, (r+1)*T),  T = N / G   (its part of the INPUT and of the OUTPUT)
    axis 0 (slices indexed by z): rank r runs its own slab's slices -> no exchange
    axis 1 (indexed by y, image (z,x)) and axis 2 (indexed by x, image (z,y)): rank r runs the slices
        y (resp. x) in [r*T, (r+1)*T).
        input : each rank holds (and, end to end, uploads) only its slab; the two uint8 strips
                vol[:, rT:(r+1)T, :] and vol[:, :, rT:(r+1)T] it needs are assembled by one uint8 all-to-all
                each (N^3/G bytes per rank) and read in place through the engine's strided slice source;
        output: each image row is a z, so the engine's head epilogue writes the probabilities
                destination-major  [dest rank h][slice][z in slab h][col][C]  (row_block = T); they travel
                in chunks of one network pass (`Engine.auto_batch` slices): the all-to-all of chunk k runs
                on NCCL's stream while the network computes chunk k+1, and lands directly in rank h's
                [y or x][z local][col][C] buffer.  No whole-axis send / receive staging exists, which is
                what lets 2048^3 fit on two GPUs.
    then K4 (reduce / blend / quantise / argmax) runs locally on each slab.

The per-voxel arithmetic and its order (axis order of `axes`) are identical to the single-GPU path, so
the sharded result is bit-identical to it.  Streams are ordered with events (engine stream <-> torch's
current stream <-> NCCL's stream); the host never waits inside the loop.
"""
import torch
import torch.distributed as dist


_tag = [0]


def _exchange(outs, ins, group, rank, world):
    """All-to-all over tensor lists: `ins[h]` goes to rank h, `outs[g]` is filled by rank g.  Returns the pending
    work handles (NCCL: one grouped send/recv; other backends: isend / irecv pairs, one tag per exchange so that
    several exchanges can be in flight between the same pair of ranks)."""
    _tag[0] = (_tag[0] + 1) % 30000
    if world == 1:
        outs[0].copy_(ins[0])
        return []
    if dist.get_backend(group) == "nccl":
        return [dist.all_to_all(outs, ins, group=group, async_op=True)]
    outs[rank].copy_(ins[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        gp = dist.get_global_rank(group, peer) if group is not None else peer
        ops.append(dist.P2POp(dist.isend, ins[peer], gp, group, _tag[0]))
        ops.append(dist.P2POp(dist.irecv, outs[peer], gp, group, _tag[0]))
    return list(dist.batch_isend_irecv(ops))


def _wait(works):
    """Wait for the handles and drop them: a gloo send / recv handle must not be waited on twice."""
    while works:
        works.pop().wait()


def _order(engine, name):
    """Stream-ordering hooks of the native engine (`wait_torch`, `torch_wait`); CPU test doubles have none."""
    fn = getattr(engine, name, None)
    if fn is not None:
        fn()


def volumes_for_rank(files, group=None):
    """Whole volumes are independent objects (`predict.py:163`: one loop iteration per store), so under `torchrun`
    `predict_volumes` gives rank r the files r, r + G, r + 2G, ... of the sorted list and needs no collective; a single
    process (or an uninitialised process group) keeps them all."""
    files = list(files)
    if not (dist.is_available() and dist.is_initialized()):
        return files
    return files[dist.get_rank(group)::dist.get_world_size(group)]


def exchange_strips(slab, axes, group=None):
    """From every rank's z-slab `[T,N,N]` (uint8 / fp32 device tensor) assemble the strips this rank predicts along the
    off-slab axes: {1: vol[:, rT:(r+1)T, :] as `[N,T,N]`, 2: vol[:, :, rT:(r+1)T] as `[N,N,T]`}."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t, n = int(slab.shape[0]), int(slab.shape[1])
    strips, works, keep = {}, [], []
    if 1 in axes:
        send = slab.view(t, world, t, n).permute(1, 0, 2, 3).contiguous()            # [dest][z local][y local][x]
        recv = torch.empty((world, t, t, n), dtype=slab.dtype, device=slab.device)     # [src][z local][y local][x]
        works += _exchange(list(recv.unbind(0)), list(send.unbind(0)), group, rank, world)
        strips[1] = recv.view(n, t, n)
        keep.append(send)
    if 2 in axes:
        send = slab.view(t, n, world, t).permute(2, 0, 1, 3).contiguous()            # [dest][z local][y][x local]
        recv = torch.empty((world, t, n, t), dtype=slab.dtype, device=slab.device)
        works += _exchange(list(recv.unbind(0)), list(send.unbind(0)), group, rank, world)
        strips[2] = recv.view(n, n, t)
        keep.append(send)
    _wait(works)
    return strips


def predict_volume_sharded(engine, volume=None, axes=(0, 1, 2), window=None, want_u8=True, want_labels=True,
                           want_mean=False, group=None, slab=None, host_u8=None, host_labels=None):
    """Predict this rank's z-slab of a cubic volume of edge N.

    engine : an `Engine` (or any object with `predict_slices`, `reduce`, `auto_batch`, `num_classes`, `device`)
    slab   : this rank's part `volume[rT:(r+1)T]` as a `[T,N,N]` device tensor -- the sharded input: strips for the
             off-slab axes are exchanged between the ranks (uint8 all-to-all);
    volume : alternatively the WHOLE volume `[N,N,N]`, identical on every rank (device tensor or numpy): every rank
             reads its slab and strips out of it in place, no input exchange.
    host_u8 / host_labels : optional (pinned) host tensors `[T,N,N,C]` / `[T,N,N]`: the slab's results are also
             copied there, part by part -- axis 0 (always computed last) runs one network pass at a time, each pass's z
             range is reduced as soon as it is predicted and copied out on a side stream under the next pass.
    Returns dict(z0, t, u8=[T,N,N,C] uint8, labels=[T,N,N] uint8, mean=[T,N,N,C] fp32) of device tensors.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = engine.device
    c = engine.num_classes
    axes = [int(a) for a in axes]
    if (volume is None) == (slab is None):
        raise ValueError("give either the replicated `volume` or this rank's `slab`")
    if slab is not None:
        slab = torch.as_tensor(slab).to(dev).contiguous()
        t, n = int(slab.shape[0]), int(slab.shape[1])
        if tuple(slab.shape) != (t, n, n) or t * world != n:
            raise ValueError(f"slab shape {tuple(slab.shape)} is not [N/{world}, N, N]")
        strips = exchange_strips(slab, axes, group)
        sources = {0: (slab, 0, (n * n, n, 1))}
        if 1 in strips:
            sources[1] = (strips[1], 0, (n, t * n, 1))           # [z][y local][x]: slice y, image (z, x)
        if 2 in strips:
            sources[2] = (strips[2], 0, (1, n * t, t))           # [z][y][x local]: slice x, image (z, y)
    else:
        volume = torch.as_tensor(volume).to(dev).contiguous()
        n = int(volume.shape[0])
        if tuple(volume.shape) != (n, n, n):
            raise ValueError("the sharded path predicts cubic volumes (predict.py:81)")
        if n % world:
            raise ValueError(f"volume edge {n} is not divisible by the number of ranks {world}")
        t = n // world
        z0 = rank * t
        sources = {0: (volume, z0 * n * n, (n * n, n, 1)), 1: (volume, z0 * n, (n, n * n, 1)),
                   2: (volume, z0, (1, n * n, n))}
    z0 = rank * t

    probs, in_flight, keep = {}, [], []
    pipelined = (host_u8 is not None or host_labels is not None) and 0 in axes and dev.type == "cuda"
    _order(engine, "wait_torch")                  # uploads / strip exchange queued on torch's stream come first
    # the off-slab axes run first so that their last exchanges overlap with axis 0's network passes; the order in
    # which axes are COMPUTED does not enter the arithmetic (K4 adds the buffers in the caller's order)
    for axis in [a for a in axes if a != 0] + [a for a in axes if a == 0]:
        src, off, strides = sources[axis]
        if axis == 0:
            probs[0] = torch.empty((t, n, n, c), dtype=torch.float32, device=dev)
            if not pipelined:
                engine.predict_slices(src, off, t, n, n, strides, probs[0], asynchronous=True, sync=False)
            continue
        dst = torch.empty((n, t, n, c), dtype=torch.float32, device=dev)     # [y | x][z local][col][C]
        probs[axis] = dst
        if world == 1:
            engine.predict_slices(src, off, t, n, n, strides, dst, row_block=t, asynchronous=True, sync=False)
            continue
        chunk = max(1, min(t, int(engine.auto_batch(n, n, t))))
        send = [torch.empty((world * chunk * t * n * c,), dtype=torch.float32, device=dev) for _ in range(2)]
        keep.append(send)
        pending = [None, None]
        for k, s0 in enumerate(range(0, t, chunk)):
            cnt = min(chunk, t - s0)
            if pending[k % 2] is not None:        # this buffer's previous exchange must have read it
                _wait(pending[k % 2])
                _order(engine, "wait_torch")
            buf = send[k % 2][:world * cnt * t * n * c].view(world, cnt, t, n, c)
            engine.predict_slices(src, off + s0 * strides[0], cnt, n, n, strides, buf, slice_offset=0, slice_total=cnt,
                                  row_block=t, asynchronous=True, sync=False)
            _order(engine, "torch_wait")          # the collective is queued behind this chunk's head kernel
            outs = [dst[g * t + s0:g * t + s0 + cnt] for g in range(world)]
            pending[k % 2] = _exchange(outs, list(buf.unbind(0)), group, rank, world)
            in_flight.append(pending[k % 2])
    out = dict(z0=z0, t=t)
    out["u8"] = torch.empty((t, n, n, c), dtype=torch.uint8, device=dev) if (want_u8 or host_u8 is not None) else None
    out["labels"] = torch.empty((t, n, n), dtype=torch.uint8, device=dev) if (want_labels or host_labels is not None) else None
    out["mean"] = torch.empty((t, n, n, c), dtype=torch.float32, device=dev) if want_mean else None
    if not pipelined:
        for w in in_flight:
            _wait(w)
        _order(engine, "wait_torch")
        engine.reduce(probs, list(axes), n, t=t, z0=z0, window=window, out_u8=out["u8"], out_labels=out["labels"],
                      out_mean=out["mean"])
        if host_u8 is not None:
            host_u8.copy_(out["u8"], non_blocking=True)
        if host_labels is not None:
            host_labels.copy_(out["labels"], non_blocking=True)
        if host_u8 is not None or host_labels is not None:
            torch.cuda.current_stream(dev).synchronize()
        return out
    # ---- axis 0 one network pass at a time; each pass's planes are reduced and copied out under the next pass
    src, off, strides = sources[0]
    chunk = max(1, min(t, int(engine.auto_batch(n, n, t))))
    copy_stream = torch.cuda.Stream(device=dev)
    ext = engine._ext_stream()
    for k, s0 in enumerate(range(0, t, chunk)):
        cnt = min(chunk, t - s0)
        engine.predict_slices(src, off + s0 * strides[0], cnt, n, n, strides, probs[0], slice_offset=s0, slice_total=t,
                              asynchronous=True, sync=False)
        if k == 0:                                # every off-slab exchange must have landed before the first reduce
            for w in in_flight:
                _wait(w)
            _order(engine, "wait_torch")
        engine.reduce(probs, list(axes), n, t=t, z0=z0, window=window, out_u8=out["u8"], out_labels=out["labels"],
                      out_mean=out["mean"], asynchronous=True, zoff=s0, zcount=cnt, sync=False)
        copy_stream.wait_event(ext.record_event())
        with torch.cuda.stream(copy_stream):
            if host_u8 is not None:
                host_u8[s0:s0 + cnt].copy_(out["u8"][s0:s0 + cnt], non_blocking=True)
            if host_labels is not None:
                host_labels[s0:s0 + cnt].copy_(out["labels"][s0:s0 + cnt], non_blocking=True)
    engine.synchronize()
    copy_stream.synchronize()
    return out


def predict_slab_from_host(engine, slab_host, axes=(0, 1, 2), window=None, out_u8=None, out_labels=None, group=None):
    """End-to-end form of `predict_volume_sharded` for host data: `slab_host` is this rank's `[T,N,N]` part of the
    volume in (pinned) host memory; the uint8 probabilities / labels of the slab are copied into the (pinned) host
    tensors `out_u8` `[T,N,N,C]` / `out_labels` `[T,N,N]`, overlapped with axis 0's network passes.  Per rank N^3/G
    bytes go up and (C+1) N^3/G come back."""
    dev = engine.device
    slab = torch.as_tensor(slab_host).to(dev, non_blocking=True)
    res = predict_volume_sharded(engine, slab=slab, axes=axes, window=window, want_u8=False, want_labels=False,
                                 group=group, host_u8=out_u8, host_labels=out_labels)
    return res["z0"], res["t"]


def gather_slabs(slab, group=None, dst=0):
    """Concatenate every rank's slab tensor along z on rank `dst` (None elsewhere)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bufs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, bufs, dst=dst, group=group)
    return torch.cat(bufs, dim=0) if rank == dst else None
