"""Multi-GPU plumbing: one process per GPU, query batches partitioned by rank, index replicated.

The path shards by QUERY: every pattern is independent, so there is no collective
inside the query loop. The only exchange steps are
  * one broadcast of the finished device index blob from rank 0 (an ncclBroadcast over
    NVLink/NVSwitch when the process group is NCCL), after which every rank attaches the bytes it
    received as its own csfm_index (csfm_attach_blob), and
  * optionally one all_gather of the per-query results (8 B/query).
Everything here is backend-agnostic torch.distributed, so the protocol is covered on CPU with
gloo (tests/test_parallel_cpu.py); only `replicate_index` touches CUDA.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` items for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_patterns(data, offs, rank: int, world: int):
    """Slice a packed batch (bytes, offs[npat+1]) for `rank`: returns (bytes, offs rebased to 0, lo, hi)."""
    npat = len(offs) - 1
    lo, hi = shard_range(npat, rank, world)
    b0, b1 = int(offs[lo]), int(offs[hi])
    return data[b0:b1], offs[lo:hi + 1] - offs[lo], lo, hi


def broadcast_bytes(buf, src: int = 0, group=None):
    """Broadcast a 1-D uint8 tensor whose length only `src` knows. Non-source ranks pass a device
    (or None for CPU) instead of a tensor and receive a freshly allocated one."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if rank == src:
        dev = buf.device
        n = torch.tensor([buf.numel()], dtype=torch.int64, device=dev)
    else:
        dev = torch.device("cpu") if buf is None else torch.device(buf)
        n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(n, src, group=group)
    if rank != src:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src, group=group)
    return buf


class _DevMem:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def replicate_index(idx, device, src: int = 0, group=None):
    """Rank `src` passes its FMIndex, the others pass None; every rank returns an FMIndex over an
    identical device blob. Returns (index, broadcast_ms measured with CUDA events)."""
    import torch
    import torch.distributed as dist
    from .binding import FMIndex
    rank = dist.get_rank(group)
    dev = torch.device(device)
    if rank == src:
        ptr, nbytes = idx.blob()
        blob = torch.as_tensor(_DevMem(ptr, nbytes), device=dev)
    else:
        blob = dev
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    blob = broadcast_bytes(blob, src, group)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if rank != src:
        idx = FMIndex.attach_blob(blob.data_ptr(), blob.numel(), dev.index, keepalive=blob)
    return idx, ms


def gather_counts(local_counts, npat: int, group=None):
    """all_gather of per-rank result slices (sizes from shard_range) into one tensor of npat."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(npat, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    pad[: local_counts.numel()] = local_counts
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])
