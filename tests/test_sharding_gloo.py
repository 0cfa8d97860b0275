"""CPU tier: the N>1 path (contiguous shards, no data-path collective; gather / status / first-error
reductions) on world_size-2 gloo.  The per-shard compute is played by the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from py_vollib_vectorized_b200 import distributed as D


def test_shard_bounds_partition():
    for n in (0, 1, 7, 101, 1_000_003):
        for world in (1, 2, 3, 8):
            b = [D.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import get_oracle
        from oracle.chains import chain_cfg3
        O = get_oracle()
        c = chain_cfg3(n, seed=1, oracle=O)
        c["K"][[n // 3, n - 5]] = 0.0  # two zero-division rows in different shards
        cols = {k: c[k] for k in ("flag", "price", "S", "K", "t", "r", "q")}
        cols["r"] = np.array([0.03])  # a broadcast scalar must pass through the sharding untouched
        sh, lo, hi = D.shard_columns(cols, n, world, rank)
        assert sh["r"].size == 1 and sh["S"].size == hi - lo
        iv, st, zde = O.iv("black_scholes_merton", sh["price"], sh["S"], sh["K"], sh["t"], sh["r"], sh["flag"], sh["q"])
        full = D.gather_column(torch.from_numpy(iv), n, dst=0)
        below = D.global_status_indices(torch.from_numpy(st), lo, 1)
        above = D.global_status_indices(torch.from_numpy(st), lo, 2)
        first = D.reduce_first_zde(-1 if zde < 0 else zde + lo)
        if rank == 0:
            q.put((full.numpy(), below, above, first))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    n, world = 5001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, below, above, first = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from oracle import get_oracle
    from oracle.chains import chain_cfg3
    O = get_oracle()
    c = chain_cfg3(n, seed=1, oracle=O)
    c["K"][[n // 3, n - 5]] = 0.0
    iv, st, zde = O.iv("black_scholes_merton", c["price"], c["S"], c["K"], c["t"], 0.03, c["flag"], c["q"])
    assert np.array_equal(np.isnan(full), np.isnan(iv)) and np.array_equal(full[~np.isnan(iv)], iv[~np.isnan(iv)])
    assert below == np.flatnonzero(st & 1).tolist() and above == np.flatnonzero(st & 2).tolist()
    assert first == zde == n // 3
