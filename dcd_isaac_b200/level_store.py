"""LevelStore -- drop-in for level_replay/level_store.py:16-110 (seed <-> level map with content dedupe,
parent lineage, FIFO cap, reconciliation against the sampler's working buffer).

Host bookkeeping is a dict keyed by the level's bytes / action string exactly as in the reference (seeds are
handed out from 1, duplicates return the existing seed).  Byte-encoded levels are additionally mirrored in ONE
uint8 tensor in HBM ([capacity, *shape], slot = seed's row), so that replaying a batch of levels is a device
gather + mgplr_reset_to_encoding without touching the host (`get_levels_device`)."""
import numpy as np

INT32_MAX = 2147483647


class LevelStore(object):
    """Public surface as consumed by envs/runners/adversarial_runner.py (:101-128, 371-376, 402-437): `insert`,
    `get_level`, `reconcile_seeds`, `remove`, `len()`, and the maps `seed2level`, `level2seed`, `seed2parent`."""

    def __init__(self, max_size=None, data_info={}, device=None):
        self.max_size, self.data_info, self.device = max_size, data_info, device
        self.next_seed = 1                      # seeds are handed out from 1, never reused
        self.seed2level, self.level2seed, self.seed2parent = {}, {}, {}
        self._mirror, self._seed2slot, self._free_slots = None, {}, []   # HBM mirror (not pickled)

    @property
    def levels(self):
        return self.level2seed.keys()

    def __len__(self):
        return len(self.level2seed)

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if k not in ('_mirror', '_seed2slot', '_free_slots')}
        state.update(_mirror=None, _seed2slot={}, _free_slots=[])
        if state['device'] is not None:
            state['device'] = str(state['device'])
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    # -- insertion: content dedupe, lineage = parent's lineage + the parent level, FIFO eviction under max_size
    def _insert(self, level, parent_seed=None):
        if level is None:
            return None
        known = self.level2seed.get(level)
        if known is not None:
            return known
        while self.max_size is not None and len(self.level2seed) >= self.max_size:
            self._remove(next(iter(self.seed2level)))   # oldest surviving seed (dicts keep insertion order)
        seed, self.next_seed = self.next_seed, self.next_seed + 1
        lineage = [] if parent_seed is None else self.seed2parent[parent_seed] + [self.seed2level[parent_seed]]
        self.seed2level[seed], self.level2seed[level], self.seed2parent[seed] = level, seed, lineage
        return seed

    def insert(self, level, parent_seeds=None):
        if isinstance(level, (bytes, str)) or not hasattr(level, '__iter__'):
            return self._insert(level)
        parents = parent_seeds if parent_seeds is not None else [None] * len(level)
        return [self._insert(l, p) for l, p in zip(level, parents)]

    def _remove(self, level_seed):
        if level_seed is None or level_seed < 0:
            return
        self.level2seed.pop(self.seed2level.pop(level_seed))
        self.seed2parent.pop(level_seed)
        slot = self._seed2slot.pop(level_seed, None)
        if slot is not None:
            self._free_slots.append(slot)

    def remove(self, level_seed):
        for seed in (level_seed if hasattr(level_seed, '__iter__') else (level_seed,)):
            self._remove(seed)

    def reconcile_seeds(self, level_seeds):
        """Drop everything the sampler's working buffer no longer holds; an all-empty buffer ({-1}) changes nothing."""
        keep = set(level_seeds)
        if keep == {-1}:
            return
        for seed in [s for s in self.seed2level if s not in keep]:
            self._remove(seed)

    def get_level(self, level_seed):
        level = self.seed2level[level_seed]
        info = self.data_info
        if info and info.get('numpy', False):
            return np.frombuffer(level, dtype=info['dtype']).reshape(*info['shape'])
        return level

    # ------------------------------------------------------------------ device mirror
    def get_levels_device(self, seeds):
        """uint8 CUDA tensor [len(seeds), *shape] of byte-encoded levels, gathered on the device."""
        import torch
        if not (self.data_info and self.data_info.get('numpy', False)):
            raise ValueError('get_levels_device needs byte-encoded levels (data_info numpy=True)')
        shape = tuple(self.data_info['shape'])
        dev = torch.device(self.device if self.device is not None else 'cuda')
        missing = [s for s in dict.fromkeys(seeds) if s not in self._seed2slot]
        need = len(self._seed2slot) + len(missing)
        cap = 0 if self._mirror is None else self._mirror.shape[0]
        if need > cap:
            new_cap = max(64, 2 * need)
            mirror = torch.zeros((new_cap,) + shape, dtype=torch.uint8, device=dev)
            if self._mirror is not None:
                mirror[:cap].copy_(self._mirror)
            self._free_slots.extend(range(cap, new_cap))
            self._mirror = mirror
        if missing:
            slots = [self._free_slots.pop() for _ in missing]
            host = np.stack([np.frombuffer(self.seed2level[s], dtype=np.uint8).reshape(shape) for s in missing])
            self._mirror[torch.tensor(slots, device=dev)] = torch.from_numpy(host).to(dev)
            for s, sl in zip(missing, slots):
                self._seed2slot[s] = sl
        idx = torch.tensor([self._seed2slot[s] for s in seeds], dtype=torch.long, device=dev)
        return self._mirror.index_select(0, idx)
