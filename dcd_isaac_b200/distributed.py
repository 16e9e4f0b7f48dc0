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


# ---------------------------------------------------------------------------------------------------------------------
# Level buffer across ranks (SURVEY.md 8e collective 2, north_star "PLR score/level-buffer all-gather").
#
# With the envs sharded, a rank only BUILDS the levels of its own envs, but the level store and the sampler are
# replicated and must stay identical to the single-process objects: seeds are handed out in insertion order
# (level_store.py:39-58), duplicates collapse onto the first seed, and a replicated sampler may draw a seed whose level
# another rank built.  So whenever the single-process runner inserts the N current levels
# (_update_plr_with_current_unseen_levels, adversarial_runner.py:402-412) every rank contributes its slice to ONE
# all-gather of the level encodings (675 / 1875 bytes each, device tensors over NCCL) and inserts the full, env-ordered list
# into its replica: same seeds, same lineage, same staging sets everywhere, and `get_level(seed)` works on every rank.
# Replay draws come from the identically seeded np.random stream on every rank; the one place where the order of draws
# depends on other ranks' envs is the per-episode-end re-sampling inside a replay rollout (adversarial_runner.py:551-558,
# ascending env order within a step), which takes an all-gather of the step's done flags (N bytes).


def _gather_tensor(x, group=None):
    """all_gather_into_tensor of equally shaped per-rank tensors along dim 0."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class ShardedPLR(object):
    """One rank's replica of the runner's PLR objects.  `sampler.num_actors` is the GLOBAL env count; `venv` holds this
    rank's `num_envs_local` envs = global envs [rank * num_envs_local, (rank + 1) * num_envs_local)."""

    def __init__(self, sampler, level_store, rank, world, num_envs_local, group=None, device=None):
        self.sampler, self.level_store = sampler, level_store
        self.rank, self.world, self.n_local, self.group = int(rank), int(world), int(num_envs_local), group
        self.n_global = self.n_local * self.world
        self.lo, self.hi = self.rank * self.n_local, (self.rank + 1) * self.n_local
        self.device = device
        self.current_level_seeds = [-1] * self.n_global   # the runner's list (adversarial_runner.py:126), all envs

    def _dev(self):
        import torch
        return torch.device(self.device) if self.device is not None else torch.device('cpu')

    def _gather_ints(self, values, dtype=None):
        import torch
        t = torch.as_tensor(list(values), dtype=dtype or torch.int64).to(self._dev())
        return _gather_tensor(t, self.group).cpu().tolist() if self.world > 1 else t.cpu().tolist()

    def gather_levels(self, local_levels):
        """Per-rank list of levels (np.uint8 [W,W,3] arrays / a uint8 device tensor [n,W,W,3], or action strings) -> the
        env-ordered list of ALL levels in the store's key format (bytes / str)."""
        import torch
        import torch.distributed as dist
        if torch.is_tensor(local_levels) or not isinstance(local_levels[0], str):
            t = local_levels if torch.is_tensor(local_levels) else torch.from_numpy(np.stack([np.asarray(l, np.uint8) for l in local_levels]))
            t = t.to(self._dev()).contiguous()
            allv = _gather_tensor(t, self.group) if self.world > 1 else t
            flat = allv.reshape(allv.shape[0], -1).cpu().numpy()
            return [row.tobytes() for row in flat]
        if self.world == 1:
            return list(local_levels)
        parts = [None] * self.world
        dist.all_gather_object(parts, list(local_levels), group=self.group)
        return [l for p in parts for l in p]

    def insert_current_levels(self, local_levels, parent_seeds=None, solvable_local=None, reject_unsolvable=False):
        """_update_plr_with_current_unseen_levels (adversarial_runner.py:402-423) for sharded envs.  `parent_seeds`: the
        GLOBAL list (it comes from the replicated replay draw) or None.  Returns this rank's slice of the new seeds."""
        levels = self.gather_levels(local_levels)
        assert len(levels) == self.n_global
        seeds = self.level_store.insert(levels, parent_seeds=parent_seeds)
        self.current_level_seeds = list(seeds)
        solvable = None
        if solvable_local is not None:
            solvable = [bool(x) for x in self._gather_ints([int(bool(x)) for x in solvable_local])]
        obs_seeds = seeds
        if reject_unsolvable and solvable is not None:
            obs_seeds = [s for s, ok in zip(seeds, solvable) if ok]
            solvable = [True] * len(obs_seeds)
        self.sampler.observe_external_unseen_sample(obs_seeds, solvable)
        return seeds[self.lo:self.hi]

    def sample_replay_levels(self):
        """`[sample_replay_level() for _ in range(num_processes)]` + get_level (adversarial_runner.py:463-465): every rank
        makes all N draws (one launch), and returns (its seeds, its levels)."""
        seeds = self.sampler.sample_replay_levels(self.n_global)
        self.current_level_seeds = list(seeds)
        mine = seeds[self.lo:self.hi]
        return mine, [self.level_store.get_level(s) for s in mine]

    def resample_finished(self, done_local):
        """Per-episode-end replay re-sampling of one vector step (adversarial_runner.py:551-558): draws happen in ascending
        GLOBAL env order.  Returns {local env index: (seed, level)} for this rank's finished envs and updates
        current_level_seeds (apply it AFTER the step's level_seeds were stored, :576-588)."""
        done_all = self._gather_ints([int(bool(d)) for d in done_local], dtype=None)
        idx = [i for i, d in enumerate(done_all) if d]
        out = {}
        if idx:
            seeds = self.sampler.sample_replay_levels(len(idx))
            for i, s in zip(idx, seeds):
                self.current_level_seeds[i] = s
                if self.lo <= i < self.hi:
                    out[i - self.lo] = (s, self.level_store.get_level(s))
        return out

    def update_with_rollouts(self, storage):
        """level_sampler.update_with_rollouts(agent.storage) (adversarial_runner.py:616-617) on the rank's slice."""
        if self.world == 1:
            return self.sampler.update_with_rollouts(storage)
        return update_sampler_sharded(self.sampler, storage, self.rank, self.world, group=self.group)

    def reconcile(self):
        """_reconcile_level_store_and_samplers (adversarial_runner.py:425-429)."""
        self.level_store.reconcile_seeds(set(int(x) for x in self.sampler.seeds if x >= 0))
