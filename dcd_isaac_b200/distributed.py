"""Multi-GPU plumbing of the hot path: one process per GPU (torch.distributed, NCCL over NVLink on the B200
box, gloo in the CPU tests).

The path shards naturally: env i lives on rank i // (N/G); every rank steps, renders, runs GAE and
reduces the episode scores of its own slice with NO data-path collective.  The only exchange is the one the
reference's single-process semantics demand: the PLR buffer is a shared, order-dependent structure, so each
rank contributes its compact episode records (40 B each) to ONE all-gather per rollout and then every rank
applies the identical, canonically ordered (actor-major, time-minor) record list to its replica of the
sampler -- replicas stay bit-identical to the single-GPU result without a broadcast.  Random decisions
(replay decision, replay draws) come from the same seeded np.random stream on every rank.
"""
import numpy as np

from ._lib import EPISODE_DTYPE


def env_shard(num_envs_total, rank, world):
    """[lo, hi) of the contiguous env slice owned by `rank`."""
    assert num_envs_total % world == 0, 'num_envs must divide evenly over the ranks'
    per = num_envs_total // world
    return rank * per, (rank + 1) * per


def all_gather_episode_records(local, actor_offset, group=None, device=None):
    """All-gather variable-length episode-record arrays (numpy, EPISODE_DTYPE).  Actor indices are shifted by
    `actor_offset` (this rank's first global env); the result is ordered by rank, i.e. actor-major."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device('cpu')
    rec = np.array(local, dtype=np.dtype(EPISODE_DTYPE), copy=True).reshape(-1)
    rec['actor'] += actor_offset
    n_local = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(1, max(counts))
    buf = torch.zeros(mx, 10, dtype=torch.int32, device=dev)
    if len(rec):
        buf[:len(rec)] = torch.from_numpy(rec.view(np.int32).reshape(-1, 10)).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = [o[:c].cpu().numpy().view(np.dtype(EPISODE_DTYPE)).reshape(-1) for o, c in zip(out, counts)]
    return np.concatenate(parts) if parts else rec


def update_sampler_sharded(sampler, rollouts, rank, world, group=None):
    """LevelSampler.update_with_rollouts for env-sharded ranks: local score kernel, one all-gather, identical
    application everywhere.  `sampler.num_actors` must be the GLOBAL env count."""
    local = sampler.episode_records(rollouts)
    n_local = rollouts.rewards.shape[1]
    dev = rollouts.rewards.device if rollouts.rewards.is_cuda else None
    rec = all_gather_episode_records(local, rank * n_local, group=group, device=dev)
    sampler._apply_episode_records(rec)
    return rec
