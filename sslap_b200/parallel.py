"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

Two ways of using several GPUs:
  * ONE problem, persons row-sharded: `init_row_sharding` builds the communicator of the C ABI (sslapb_comm_init /
    _connect) — every rank then calls `auction_solve` with the same full problem and the ranks split the bidding step of
    the large-frontier rounds, exchanging bids in-kernel over NVLink; results are bit-identical on every rank.
  * INDEPENDENT problems (a batch): dealt out to the ranks, every rank solves its shard with the one-warp-per-problem
    batch kernel, nothing crosses NVLink on the data path, results are gathered once (`solve_batch_sharded`).
`balanced_row_split` is the numpy statement of the nnz-balanced row partition the device computes
(sslapb_row_split_kernel) — tests compare the two.
"""
import numpy as np


def init_row_sharding(handle=None, capacity_rows: int = 1 << 20, group=None):
    """Build the row-sharding communicator over the ranks of a torch.distributed process group (any backend: only the
    128-byte export blobs travel through it).  Every rank must call this, then make identical `auction_solve` calls on
    `handle`.  Returns (rank, world)."""
    import torch.distributed as dist
    from . import _native as nat
    h = handle or nat.default_handle()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world > nat.COMM_MAX_RANKS:
        raise ValueError(f"row sharding supports up to {nat.COMM_MAX_RANKS} ranks")
    mine = h.comm_init(world, rank, capacity_rows)
    box = [None] * world
    dist.all_gather_object(box, mine, group=group)
    h.comm_connect(box)
    dist.barrier(group=group)                             # nobody solves before every rank has mapped its peers
    return rank, world


def connect_local(handles, capacity_rows: int):
    """Same-process communicator over `handles` (one per GPU, or several on ONE GPU with option "max_ctas" so that their
    persistent kernels are co-resident — the virtual-shard test).  The solves must then run concurrently, one thread per
    handle, because every sharded round waits for all ranks."""
    blobs = [h.comm_init(len(handles), r, capacity_rows) for r, h in enumerate(handles)]
    for h in handles:
        h.comm_connect(blobs)


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
