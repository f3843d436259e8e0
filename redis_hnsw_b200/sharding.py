"""Multi-GPU plumbing for the search path: the index is replicated, the query batch is split into
contiguous slices, every rank writes its own slice of the results.  There is no collective on the query path; the
only collectives are the one-time broadcast of the index's device buffers (NCCL over NVLink) and the optional
all-gather of the result slices (80 bytes per query at k = 10).

Works with any torch.distributed backend: `nccl` on the GPUs, `gloo` in the CPU tests (tests/test_sharding_cpu.py)."""
import numpy as np


def query_slice(nq, rank, world):
    """[lo, hi) of the rank's contiguous query slice; the first nq % world ranks take one extra query."""
    base, rem = divmod(int(nq), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevPtr:
    """Zero-copy torch view of a raw device buffer."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def replicate_index(dev, rank, world, src=0):
    """Broadcast the device buffers of `dev` on rank `src` into the (same-parameter) index of every other rank.
    One ncclBroadcast per buffer: vector slab, adjacency rows, overflow links, levels, pool, meta.
    Returns (bytes broadcast, seconds of the broadcasts on this rank) — (0, 0.0) for a single rank."""
    if world == 1:
        return 0, 0.0
    import time

    import torch
    import torch.distributed as dist

    lay = (torch.from_numpy(dev.replica_layout().astype(np.int64)).cuda() if rank == src
           else torch.zeros(8, dtype=torch.int64, device="cuda"))
    dist.broadcast(lay, src)
    if rank != src:
        dev.prepare_replica(lay.cpu().numpy().astype(np.uint64))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    total = 0
    for ptr, nbytes in dev.device_buffers():
        if nbytes:
            dist.broadcast(torch.as_tensor(_DevPtr(ptr, nbytes), device="cuda"), src)
            total += nbytes
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    if rank != src:
        dev.adopt_replica()
    return total, secs


def result_checksum(ids, sims):
    """Order-sensitive 64-bit checksum of a result block (ids [n, k] u32, sims [n, k] f32 compared as bits): equal on two
    ranks iff they returned the same neighbours with the same sims in the same places (up to hash collisions)."""
    a = np.ascontiguousarray(ids, dtype=np.uint32).astype(np.uint64).ravel()
    b = np.ascontiguousarray(sims, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
    pos = np.arange(1, a.size + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = (a * np.uint64(0x9E3779B97F4A7C15) + b * np.uint64(0xC2B2AE3D27D4EB4F) + pos) * (pos | np.uint64(1))
        return int(np.bitwise_xor.reduce(h ^ (h >> np.uint64(29))))


def ranks_agree(checksum, world, device=None):
    """True iff every rank computed the same checksum (all-gather of one int64 per rank)."""
    if world == 1:
        return True, [checksum]
    import torch
    import torch.distributed as dist

    t = torch.tensor([checksum & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    vals = [int(p.item()) for p in parts]
    return all(v == vals[0] for v in vals), vals


def gather_results(ids, sims, counts, nq_total, rank, world, device=None):
    """All-gather the per-rank result slices (numpy arrays [n_r, k], [n_r, k], [n_r]) into full [nq_total, ...] arrays."""
    if world == 1:
        return ids, sims, counts
    import torch
    import torch.distributed as dist

    k = ids.shape[1]
    width = max(query_slice(nq_total, r, world)[1] - query_slice(nq_total, r, world)[0] for r in range(world))

    def pad(a, fill):
        out = np.full((width,) + a.shape[1:], fill, dtype=a.dtype)
        out[:a.shape[0]] = a
        return out

    packed = np.concatenate([pad(ids.astype(np.uint32), 0xFFFFFFFF).view(np.int32),
                             pad(sims.astype(np.float32), -np.inf).view(np.int32),
                             pad(counts.astype(np.uint32), 0).view(np.int32)[:, None]], axis=1)
    t = torch.from_numpy(packed)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out_ids = np.empty((nq_total, k), np.uint32)
    out_sims = np.empty((nq_total, k), np.float32)
    out_counts = np.empty(nq_total, np.uint32)
    for r, p in enumerate(parts):
        lo, hi = query_slice(nq_total, r, world)
        a = p.cpu().numpy()[:hi - lo]
        out_ids[lo:hi] = a[:, :k].view(np.uint32)
        out_sims[lo:hi] = a[:, k:2 * k].view(np.float32)
        out_counts[lo:hi] = a[:, 2 * k].view(np.uint32)
    return out_ids, out_sims, out_counts


def sharded_search(search_fn, Q, k, rank, world, gather=True, device=None):
    """Run `search_fn(Q_slice) -> (ids, sims, counts)` on the rank's slice of Q; optionally gather the full result."""
    lo, hi = query_slice(Q.shape[0], rank, world)
    ids, sims, counts = search_fn(Q[lo:hi])
    if not gather:
        return ids, sims, counts
    return gather_results(ids, sims, counts, Q.shape[0], rank, world, device)
