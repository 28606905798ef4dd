"""Multi-GPU tests (-m gpu, need >= 2 devices; skipped on a one-GPU box): ONE problem row-sharded over 2 processes, one
GPU each, bids exchanged in-kernel over NVLink through CUDA-IPC mapped buffers — every rank must return
the single-GPU / oracle trajectory bit for bit; and a batch of independent problems dealt out to the ranks."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sslap_b200
    from sslap_b200 import _native as nat, parallel
    from sslap_b200.datagen import make_problem
    h = nat.Handle(rank)
    h.set_option("watchdog_ms", 30000)
    h.set_option("t_shard", 256)
    parallel.init_row_sharding(h, 1 << 17)
    res = {}
    for k, (n, d, mode, seed) in enumerate([(20000, 0.001, "float", 3), (1000, 0.01, "int", 0), (100000, 0.0002, "float", 5)]):
        loc, val = make_problem(n, d, mode, seed=seed)
        g = sslap_b200.auction_solve(loc=loc, val=val, size=(n, n), problem="min", cardinality_check=False, _handle=h,
                                     _raw_meta=True, return_prices=True)
        res[f"sol{k}"] = g["sol"]; res[f"prices{k}"] = g["prices"]
        res[f"its{k}"] = g["meta"]["its"]; res[f"sharded{k}"] = int(g["raw"].rounds_sharded)
        res[f"rows{k}"] = np.array([g["raw"].row_lo, g["raw"].row_hi])
    # batch shards: every rank solves its contiguous share of 12 problems, results gathered once
    probs = [make_problem(100 + 9 * k, 0.1, "float", seed=50 + k) for k in range(12)]

    def solve(chunk):
        return [r["sol"] for r in sslap_b200.auction_solve_batch([(l, v, (int(l[:, 0].max()) + 1,) * 2) for (l, v) in chunk],
                                                                 _handle=h)]

    def gather(obj):
        box = [None] * world
        dist.all_gather_object(box, obj)
        return box

    full = parallel.solve_batch_sharded(probs, solve, world, rank, gather)
    for k, s in enumerate(full):
        res[f"b{k}"] = s
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    h.close()
    dist.destroy_process_group()


def test_two_gpu_row_sharded_solve_and_batch_shards(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from sslap_b200.datagen import make_problem
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for k, (n, d, mode, seed) in enumerate([(20000, 0.001, "float", 3), (1000, 0.01, "int", 0), (100000, 0.0002, "float", 5)]):
        loc, val = make_problem(n, d, mode, seed=seed)
        want = oracle.auction_solve(loc=loc, val=val, problem="min", return_prices=True)
        for r, o in enumerate(outs):
            assert np.array_equal(o[f"sol{k}"], want["sol"]), (k, r)
            assert int(o[f"its{k}"]) == want["meta"]["its"]
            assert np.array_equal(o[f"prices{k}"], want["prices"])
            assert int(o[f"sharded{k}"]) > 0
        assert int(outs[0][f"rows{k}"][0]) == 0 and int(outs[1][f"rows{k}"][1]) == n
        assert int(outs[0][f"rows{k}"][1]) == int(outs[1][f"rows{k}"][0])
    for k in range(12):
        l, v = make_problem(100 + 9 * k, 0.1, "float", seed=50 + k)
        want = oracle.auction_solve(loc=l, val=v)["sol"]
        for o in outs:
            assert np.array_equal(o[f"b{k}"], want)
