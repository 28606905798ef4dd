"""CPU tests (gloo, world_size 2) of the host-side multi-GPU logic: problem sharding, result gathering, max-over-ranks
timing.  The device solve is replaced by the oracle here (tests may use it as a stand-in; there is no GPU in CI)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle
    from sslap_b200 import parallel
    from sslap_b200.datagen import make_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    probs = [make_problem(40 + 3 * k, 0.2, "int" if k % 2 else "float", seed=k) for k in range(7)]

    def solve(chunk):
        return [oracle.auction_solve(loc=l, val=v)["sol"] for (l, v) in chunk]

    def gather(obj):
        box = [None] * world
        dist.all_gather_object(box, obj)
        return box

    res = parallel.solve_batch_sharded(probs, solve, world, rank, gather)
    t = parallel.max_over_ranks(1.0 + rank)
    lo, hi = parallel.shard_range(len(probs), world, rank)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), t=t, lo=lo, hi=hi, **{f"s{k}": r for k, r in enumerate(res)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_sharding_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from sslap_b200.datagen import make_problem
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [(int(o["lo"]), int(o["hi"])) for o in outs] == [(0, 4), (4, 7)]
    for o in outs:
        assert float(o["t"]) == 2.0                       # max over ranks of 1+rank
        for k in range(7):
            l, v = make_problem(40 + 3 * k, 0.2, "int" if k % 2 else "float", seed=k)
            assert np.array_equal(o[f"s{k}"], oracle.auction_solve(loc=l, val=v)["sol"])


def test_shard_range_and_balanced_row_split():
    from sslap_b200 import parallel
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    deg = rng.integers(1, 200, 5000)
    indptr = np.concatenate([[0], np.cumsum(deg)])
    for parts in (1, 2, 4, 8):
        b = parallel.balanced_row_split(indptr, parts)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0) and len(b) == parts + 1
        loads = np.diff(indptr[b])
        assert loads.max() - loads.min() <= 2 * deg.max()  # nnz-balanced up to one row
