"""Read sharding across the GPUs of one box: reads are independent units, the index is
replicated per HBM, there is no collective on the data path; only the per-rank results are gathered
on rank 0 (sizes with one all_gather, payload with one padded all_gather). Works with the `nccl`
backend (device tensors) and with `gloo` (CPU tensors, used by the CPU tests)."""
import numpy as np


def shard_range(n, rank, world):
    """contiguous range [n*rank/world, n*(rank+1)/world) of read indexes"""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_bytes(payload, dst=0, device=None):
    """Gathers one bytes object per rank on `dst`; returns the list (rank order) there, else None."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if payload:
        buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != dst:
        return None
    return [bytes(o[:s].cpu().numpy().tobytes()) for o, s in zip(out, sizes)]


def align_sharded(aligner, batch, format_fn, device=None):
    """Aligns this rank's shard of `batch` ([(header, seq)]) and gathers the formatted lines on
    rank 0 in input order. Returns the list of lines on rank 0, None elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(len(batch), rank, world)
    mine = batch[lo:hi]
    res = aligner.align_batch(mine) if mine else []
    lines = [format_fn(h, r) for (h, _), r in zip(mine, res)]
    parts = gather_bytes("\n".join(lines).encode(), dst=0, device=device)
    if parts is None:
        return None
    out = []
    for p in parts:
        if p:
            out.extend(p.decode().split("\n"))
    return out
