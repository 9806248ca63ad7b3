"""world_size-2 gloo test of the chain sharding / final gather (host-side logic of the multi-GPU path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_chains, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from nhmc_b200 import dist as nd
    r, w, _ = nd.init_from_env("gloo")
    a, b = nd.chain_shard(n_chains, r, w)
    # every chain produces a tensor that depends only on its own seed, never on the sharding
    local = torch.stack([torch.randn(2, 3, generator=torch.Generator().manual_seed(nd.chain_seed(5678, c))) for c in range(a, b)]) \
        if b > a else torch.zeros(0, 2, 3)
    out = nd.gather_chains(local, n_chains)
    t = nd.max_over_ranks(10.0 + r, torch.device("cpu"))
    if r == 0:
        q.put((out.numpy(), t))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [5, 8])
def test_chain_sharding_and_gather_world2(n_chains):
    from nhmc_b200 import dist as nd
    assert [nd.chain_shard(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [nd.chain_shard(64, r, 8) for r in range(8)][-1] == (56, 64)
    assert sorted(sum(([*range(*nd.chain_shard(7, r, 4))] for r in range(4)), [])) == list(range(7))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chains, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([torch.randn(2, 3, generator=torch.Generator().manual_seed(5678 + c)).numpy() for c in range(n_chains)])
    assert np.array_equal(out, ref)                   # same chains, same order as a single process would give
    assert t == 11.0


def _worker_samples(rank, world, port, n_chains, counts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from nhmc_b200 import dist as nd
    r, w, _ = nd.init_from_env("gloo")
    a, b = nd.chain_shard(n_chains, r, w)
    S = counts[r]
    local = None
    if b > a:
        local = torch.stack([torch.arange(S * 4, dtype=torch.float32).reshape(S, 1, 2, 2) + 100 * c for c in range(a, b)])
    out = nd.gather_samples(local, n_chains, (1, 2, 2), torch.device("cpu"))
    if r == 0:
        q.put(None if out is None else out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chains,counts,expect_S", [(1, (3, 0), 3), (3, (4, 2), 2), (2, (0, 5), 0)])
def test_final_gather_with_idle_ranks_and_unequal_counts(n_chains, counts, expect_S):
    """a rank that owns no chain (chains < world) still takes part in the collective, ranks that collected
    different numbers of samples agree on the common count, nothing hangs."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_samples, args=(r, 2, port, n_chains, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if expect_S == 0:
        assert out is None
        return
    assert out.shape == (n_chains, expect_S, 1, 2, 2)
    for c in range(n_chains):
        assert np.array_equal(out[c].reshape(-1), np.arange(expect_S * 4, dtype=np.float32) + 100 * c)
