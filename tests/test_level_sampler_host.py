"""Host-only pieces of the LevelSampler drop-in (CPU; no kernel is involved): construction limits, staging bookkeeping of
observe_external_unseen_sample, the replay decision's use of the global np.random stream, pickling.  The record walk itself
runs on the device (tests/test_gpu_plr_parity.py compares it with oracle.plr_oracle.BufferOracle and the recorded reference
sessions)."""
import pickle

import numpy as np
import pytest

from dcd_isaac_b200.level_sampler import LevelSampler


def _mk(**kw):
    args = dict(num_actors=4, strategy='positive_value_loss', sample_full_distribution=True, seed_buffer_size=8)
    args.update(kw)
    return LevelSampler([], None, None, **args)


def test_pickle_round_trip():
    s = _mk()
    s.observe_external_unseen_sample([1, 2, 3])
    s2 = pickle.loads(pickle.dumps(s))
    assert s2.staging_seed_set == {1, 2, 3} and s2.running_sample_count == 3
    assert s2.seed2timestamp_buffer == {1: 1, 2: 2, 3: 3} and s2._dev is None


def test_buffer_limit_is_checked_at_construction():
    with pytest.raises(ValueError):
        _mk(seed_buffer_size=8193)
    with pytest.raises(ValueError):
        LevelSampler(list(range(9000)), None, None, strategy='value_l1')
    assert _mk(seed_buffer_size=8192).seed_buffer_size == 8192


def test_observe_resets_staleness_of_known_seeds():
    s = _mk(staleness_coef=0.3)
    s.observe_external_unseen_sample([5, 6], solvable=[True, False])
    assert s.track_solvable and s.staging_seed2solvable == {5: True, 6: False}
    # pretend 5 was admitted to slot 2
    s.staging_seed_set.discard(5); s.working_seed_set.add(5); s.seed2index[5] = 2; s.seeds[2] = 5
    s.seed_staleness[:] = 3
    s.observe_external_unseen_sample([5, 7])
    assert s.seed_staleness[2] == 0 and (np.delete(s.seed_staleness, 2) == 4).all()
    assert s.staging_seed_set == {6, 7} and s.seed2timestamp_buffer[7] == 4 and s.running_sample_count == 4


@pytest.mark.parametrize('schedule', ['fixed', 'proportionate'])
def test_replay_decision_draws_only_when_warm(schedule):
    """level_sampler.py:606-639: the uniform is drawn only when the fill test passes (Python `and`), so a recorded global
    np.random stream replays identically."""
    s = _mk(replay_schedule=schedule, rho=0.5, replay_prob=0.7)
    np.random.seed(0)
    pos = np.random.get_state()[2]
    assert s.sample_replay_decision() is False and np.random.get_state()[2] == pos   # cold: no draw
    s.working_seed_buffer_size = 6
    np.random.seed(0)
    ref = np.random.RandomState(0)
    u = ref.rand()
    want = u < (0.7 if schedule == 'fixed' else min(6 / 8, 0.7))
    assert s.sample_replay_decision() == want and np.random.get_state()[2] == ref.get_state()[2]   # exactly one uniform
    # fixed seed set (no full distribution): fixed schedule replays once everything is seen, whatever the draw
    f = LevelSampler([3, 4, 5], None, None, strategy='value_l1', replay_schedule='fixed', rho=0.5, replay_prob=0.0)
    assert f.sample_replay_decision() is False
    f.unseen_seed_weights[:] = 0
    assert f.sample_replay_decision() is True
