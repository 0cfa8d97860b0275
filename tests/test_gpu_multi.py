"""GPU tier, needs >= 2 GPUs (skipped on a 1-GPU box): the N > 1 path on hardware.  One process per GPU over NCCL;
every rank runs the CUDA engine on its contiguous shard (no collective on the data path), then the optional steps of
py_vollib_vectorized_b200/distributed.py — gather of the output column to rank 0, global ascending status index
lists, min-reduction of the first ZeroDivisionError index — and rank 0's result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from py_vollib_vectorized_b200 import _engine as E, distributed as D
        from oracle import get_oracle
        from oracle.chains import chain_cfg3
        c = chain_cfg3(n, seed=1, oracle=get_oracle())  # inputs only: the compute below is the engine's
        c["K"][[n // 3, n - 5]] = 0.0  # two zero-division rows in different shards
        cols = {k: c[k] for k in ("flag", "price", "S", "K", "t", "r", "q")}
        sh, lo, hi = D.shard_columns(cols, n, world, rank)
        m = hi - lo
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sh.items()}
        iv = torch.empty(m, dtype=torch.float64, device=dev)
        status = torch.empty(m, dtype=torch.uint8, device=dev)
        first = torch.full((1,), np.iinfo(np.int64).max, dtype=torch.int64, device=dev)
        E.iv_dev("black_scholes_merton", m, d["flag"], d["price"], d["S"], d["K"], d["t"], d["r"], d["q"], iv, status, first_zde=first)
        torch.cuda.synchronize()
        zde = int(first.item())
        zde = -1 if zde == np.iinfo(np.int64).max else zde + lo
        full = D.gather_column(iv, n, dst=0)
        below = D.global_status_indices(status, lo, 1)
        above = D.global_status_indices(status, lo, 2)
        zero = D.global_status_indices(status, lo, 4)
        first_global = D.reduce_first_zde(zde, device=dev)
        if rank == 0:
            q.put((full.cpu().numpy(), below, above, zero, first_global))
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_iv_matches_single_process_oracle(oracle):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(torch.cuda.device_count(), 8)
    n = 2_000_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, below, above, zero, first = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    from oracle.chains import chain_cfg3
    c = chain_cfg3(n, seed=1, oracle=oracle)
    c["K"][[n // 3, n - 5]] = 0.0
    iv, st, zde = oracle.iv("black_scholes_merton", c["price"], c["S"], c["K"], c["t"], c["r"], c["flag"], c["q"], threads=oracle.max_threads)
    ok = (st & 4) == 0  # rows where numba would have raised are compared on the flag only
    assert np.array_equal(np.isnan(full[ok]), np.isnan(iv[ok])) and np.array_equal(full[ok][~np.isnan(iv[ok])], iv[ok][~np.isnan(iv[ok])])
    assert below == np.flatnonzero(st & 1).tolist() and above == np.flatnonzero(st & 2).tolist()
    assert zero == np.flatnonzero(st & 4).tolist() == [n // 3, n - 5]
    assert first == zde == n // 3
