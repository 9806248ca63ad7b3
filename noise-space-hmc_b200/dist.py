"""Chain sharding across the GPUs of one box: one process per GPU, no collective in the sampler loop.

Chains (and images) are fully independent: chain c runs on rank c // ceil(C / world) with a full
model replica; the only communication is one gather of the collected samples / per-chain state at the end
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no distributed code at all.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world, local_rank); a single process needs no process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def chain_shard(n_chains, rank, world):
    """[start, stop) of the chains owned by `rank`: contiguous blocks, sizes differ by at most one."""
    base, extra = divmod(n_chains, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def chain_seed(base_seed, chain):
    """Seed of a chain's momentum / uniform streams: independent of how chains are sharded."""
    return int(base_seed) + int(chain)


def gather_chains(local, n_chains, dst=0):
    """Gather per-chain tensors [n_local, ...] from every rank onto `dst` in chain order; returns the
    [n_chains, ...] tensor on dst and None elsewhere.  Shards may have different sizes."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [chain_shard(n_chains, r, world) for r in range(world)]
    n_max = max(b - a for a, b in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)


def gather_samples(local, n_chains, sample_shape, device, dtype=torch.float32, dst=0):
    """Final gather of the collected samples (the only collective of a run).  local: [n_local, S_local, *sample_shape] of
    this rank's chains, or None / empty when the rank owns no chain (chains < world) or collected nothing.  EVERY rank
    makes the same collective calls: the ranks first agree on S = min over ranks that own chains (a run cut short by
    --max_iters can leave ranks with different counts), then all_gather [n_max, S, ...] blocks.  Returns
    [n_chains, S, *sample_shape] on dst (None elsewhere; None everywhere if S == 0)."""
    have = local is not None and local.numel() > 0
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local if have else None
    big = 1 << 30
    s_loc = torch.tensor([local.shape[1] if have else (0 if (local is not None and local.shape[0] > 0) else big)],
                         dtype=torch.int64, device=device)
    dist.all_reduce(s_loc, op=dist.ReduceOp.MIN)
    S = int(s_loc.item())
    if S >= big or S == 0:
        return None
    a, b = chain_shard(n_chains, dist.get_rank(), dist.get_world_size())
    if have:
        block = local[:, :S].to(device=device, dtype=dtype).contiguous()
    else:
        block = torch.zeros((b - a, S) + tuple(sample_shape), dtype=dtype, device=device)
    return gather_chains(block, n_chains, dst)


def max_over_ranks(value, device):
    """Device-timed durations are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
