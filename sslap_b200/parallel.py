"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

The auction path shards in exactly one way today: INDEPENDENT PROBLEMS (a batch, or replicas of one instance) are dealt
out to the ranks, every rank runs the whole device-resident solve on its own GPU, and nothing crosses NVLink on the
data path — results are gathered once at the end ("replicas only").  `balanced_row_split` is the
nnz-balanced contiguous row partition the single-problem row-sharded path will use.
"""
import numpy as np


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) share of n_items for `rank` (earlier ranks take the remainder)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_row_split(indptr, parts: int):
    """Split rows 0..N-1 into `parts` contiguous ranges holding (nearly) equal numbers of CSR entries.
    Returns parts+1 row boundaries, non-decreasing, first 0 and last N."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    total = int(indptr[-1] - indptr[0])
    targets = indptr[0] + (np.arange(1, parts, dtype=np.int64) * total) // parts
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def solve_batch_sharded(problems, solve_fn, world: int = 1, rank: int = 0, gather_fn=None):
    """Deal `problems` out to the ranks (contiguous shards), solve the local shard with `solve_fn(list) -> list`, and —
    when `gather_fn` (e.g. a wrapper of torch.distributed.all_gather_object) is given — return the full result list in
    the original order on every rank; otherwise return only the local results."""
    lo, hi = shard_range(len(problems), world, rank)
    local = solve_fn(problems[lo:hi]) if hi > lo else []
    if gather_fn is None or world == 1:
        return local
    parts = gather_fn((lo, local))
    out = [None] * len(problems)
    for (start, res) in parts:
        out[start:start + len(res)] = res
    return out


def max_over_ranks(value: float, group=None) -> float:
    """max of a per-rank timing over all ranks (device tensor when NCCL is the backend, CPU tensor under gloo)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
