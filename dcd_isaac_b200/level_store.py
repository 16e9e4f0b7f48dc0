"""LevelStore -- drop-in for level_replay/level_store.py:16-110 (seed <-> level map with content dedupe,
parent lineage, FIFO cap, reconciliation against the sampler's working buffer).

Host bookkeeping is a dict keyed by the level's bytes / action string exactly as in the reference (seeds are
handed out from 1, duplicates return the existing seed).  Byte-encoded levels are additionally mirrored in ONE
uint8 tensor in HBM ([capacity, *shape], slot = seed's row), so that replaying a batch of levels is a device
gather + mgplr_reset_to_encoding without touching the host (`get_levels_device`)."""
from collections import defaultdict

import numpy as np

INT32_MAX = 2147483647


class LevelStore(object):
    def __init__(self, max_size=None, data_info={}, device=None):
        self.max_size = max_size
        self.seed2level = defaultdict()
        self.level2seed = defaultdict()
        self.seed2parent = defaultdict()
        self.next_seed = 1
        self.levels = set()
        self.data_info = data_info
        self.device = device
        self._mirror = None       # device tensor [capacity, *shape] (not pickled)
        self._seed2slot = {}
        self._free_slots = []

    def __len__(self):
        return len(self.levels)

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_mirror'] = None
        st['_seed2slot'] = {}
        st['_free_slots'] = []
        if st.get('device') is not None:
            st['device'] = str(st['device'])
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)

    def _insert(self, level, parent_seed=None):
        if level is None:
            return None
        if level not in self.levels:
            if self.max_size is not None:  # FIFO if max size constraint
                while len(self.levels) >= self.max_size:
                    first_idx = list(self.seed2level)[0]
                    self._remove(first_idx)
            seed = self.next_seed
            self.seed2level[seed] = level
            if parent_seed is not None:
                self.seed2parent[seed] = self.seed2parent[parent_seed] + [self.seed2level[parent_seed]]
            else:
                self.seed2parent[seed] = []
            self.level2seed[level] = seed
            self.levels.add(level)
            self.next_seed += 1
            return seed
        return self.level2seed[level]

    def insert(self, level, parent_seeds=None):
        if hasattr(level, '__iter__') and not isinstance(level, (bytes, str)):
            idx = []
            for i, l in enumerate(level):
                ps = None
                if parent_seeds is not None:
                    ps = parent_seeds[i]
                idx.append(self._insert(l, ps))
            return idx
        return self._insert(level)

    def _remove(self, level_seed):
        if level_seed is None or level_seed < 0:
            return
        level = self.seed2level[level_seed]
        self.levels.remove(level)
        del self.seed2level[level_seed]
        del self.level2seed[level]
        del self.seed2parent[level_seed]
        slot = self._seed2slot.pop(level_seed, None)
        if slot is not None:
            self._free_slots.append(slot)

    def remove(self, level_seed):
        if hasattr(level_seed, '__iter__'):
            for i in level_seed:
                self._remove(i)
        else:
            self._remove(level_seed)

    def reconcile_seeds(self, level_seeds):
        old_seeds = set(self.seed2level)
        new_seeds = set(level_seeds)
        if len(new_seeds) == 1 and -1 in new_seeds:  # don't update if empty seeds
            return
        for seed in old_seeds - new_seeds:
            self._remove(seed)

    def get_level(self, level_seed):
        level = self.seed2level[level_seed]
        if self.data_info:
            if self.data_info.get('numpy', False):
                dtype = self.data_info['dtype']
                shape = self.data_info['shape']
                level = np.frombuffer(level, dtype=dtype).reshape(*shape)
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
