"""ctypes binding of libmgplr.so (the C ABI declared in include/mgplr.h).

There is NO CPU fallback: if the CUDA library is missing, cannot be loaded, or no CUDA device is
present, every entry point raises.  The library is built in-tree by `__graft_entry__.build()`
(`make -C dcd_isaac_b200/csrc`) for sm_100a only.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, 'libmgplr.so')
CSRC = os.path.join(HERE, 'csrc')

_lib = None


class MgplrError(RuntimeError):
    pass


class EnvConfig(C.Structure):
    """struct mgplr_env_config (include/mgplr.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        'width', 'agent_view_size', 'max_steps', 'max_episode_steps', 'see_through_walls', 'n_clutter',
        'resample_n_clutter', 'choose_goal_last', 'fixed_environment', 'n_editor_actions')]


class StepOut(C.Structure):
    """struct mgplr_step_out (include/mgplr.h): raw device pointers, 0 = not wanted."""
    FIELDS = ('image', 'direction', 'reward', 'flags', 'ep_return', 'ep_length', 'trunc_image',
              'trunc_direction', 'masks', 'bad_masks', 'cliffhanger_masks', 'image_u8', 'trunc_full_obs')
    _fields_ = [(n, C.c_void_p) for n in FIELDS]


class Episode(C.Structure):
    """struct mgplr_episode (include/mgplr.h)."""
    _fields_ = [('actor', C.c_int32), ('t_start', C.c_int32), ('t_end', C.c_int32), ('seed', C.c_int32),
                ('mean_score', C.c_float), ('max_score', C.c_float), ('reward_sum', C.c_float),
                ('value_sum', C.c_float), ('value_min', C.c_float), ('cliffhanger', C.c_int32)]


DONE_DTYPE = [('env', '<i4'), ('reward', '<f4'), ('ep_return', '<f4'), ('ep_length', '<i4')]
EPISODE_DTYPE = [('actor', '<i4'), ('t_start', '<i4'), ('t_end', '<i4'), ('seed', '<i4'), ('mean_score', '<f4'),
                 ('max_score', '<f4'), ('reward_sum', '<f4'), ('value_sum', '<f4'), ('value_min', '<f4'),
                 ('cliffhanger', '<i4')]

# every symbol include/mgplr.h declares; tests check that the .so exports all of them
SYMBOLS = (
    'mgplr_last_error', 'mgplr_abi_version', 'mgplr_venv_create', 'mgplr_venv_destroy', 'mgplr_venv_num_envs',
    'mgplr_venv_state_bytes', 'mgplr_seed', 'mgplr_reset', 'mgplr_step_adversary', 'mgplr_reset_agent',
    'mgplr_reset_random', 'mgplr_reset_to_encoding', 'mgplr_load_levels', 'mgplr_load_levels_at', 'mgplr_reset_to_actions', 'mgplr_mutate_edits',
    'mgplr_mutate_finalize', 'mgplr_step_env', 'mgplr_step_env_u8', 'mgplr_step_env_host', 'mgplr_step_env_host_u8', 'mgplr_rollout', 'mgplr_rollout_ex', 'mgplr_full_obs', 'mgplr_render_images', 'mgplr_get_encodings',
    'mgplr_get_metrics', 'mgplr_get_agent_state', 'mgplr_get_errors', 'mgplr_peek_rng', 'mgplr_gae',
    'mgplr_discounted_returns', 'mgplr_batched_value_loss',
    'mgplr_plr_episode_scores', 'mgplr_plr_episode_scores_ex', 'mgplr_plr_sample_weights', 'mgplr_plr_score_weights', 'mgplr_plr_sample_replay', 'mgplr_plr_apply_records',
    'mgplr_wide_create', 'mgplr_wide_destroy', 'mgplr_wide_load_levels', 'mgplr_wide_step', 'mgplr_wide_get_encodings',
)


def build(verbose=False):
    """Compile libmgplr.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(['make', '-C', CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise MgplrError('building libmgplr.so failed:\n' + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return SO_PATH


def load():
    """dlopen the library (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MgplrError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                         '(there is no CPU fallback)' % SO_PATH)
    L = C.CDLL(SO_PATH)
    L.mgplr_last_error.restype = C.c_char_p
    L.mgplr_venv_state_bytes.restype = C.c_int64
    L.mgplr_venv_destroy.restype = None
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    L.mgplr_venv_create.argtypes = [C.POINTER(EnvConfig), i32, i32, C.POINTER(vp)]
    L.mgplr_venv_destroy.argtypes = [vp]
    L.mgplr_venv_num_envs.argtypes = [vp]
    L.mgplr_venv_state_bytes.argtypes = [vp]
    L.mgplr_seed.argtypes = [vp, vp, vp, vp, i32, vp]
    L.mgplr_reset.argtypes = [vp, vp, vp, vp]
    L.mgplr_step_adversary.argtypes = [vp, vp, vp, vp, vp, vp]
    L.mgplr_reset_agent.argtypes = [vp, C.POINTER(StepOut), vp]
    L.mgplr_reset_random.argtypes = [vp, vp, C.POINTER(StepOut), vp]
    L.mgplr_reset_to_encoding.argtypes = [vp, vp, vp, i32, C.POINTER(StepOut), vp]
    L.mgplr_load_levels.argtypes = [vp, vp, i32, vp, i32, C.POINTER(StepOut), vp]
    L.mgplr_load_levels_at.argtypes = [vp, vp, vp, i32, i32, C.POINTER(StepOut), vp]
    L.mgplr_reset_to_actions.argtypes = [vp, vp, i32, vp, i32, C.POINTER(StepOut), vp]
    L.mgplr_mutate_edits.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    L.mgplr_mutate_finalize.argtypes = [vp, vp, C.POINTER(StepOut), vp]
    L.mgplr_step_env.argtypes = [vp, vp, i32, vp, i32, C.POINTER(StepOut), vp]
    L.mgplr_step_env_host.argtypes = [vp, vp, i32, i32, C.POINTER(StepOut), vp, vp, i32, vp, vp]
    L.mgplr_step_env_host_u8.argtypes = [vp, vp, i32, i32, C.POINTER(StepOut), vp, vp, i32, vp, vp]
    L.mgplr_step_env_u8.argtypes = [vp, vp, i32, vp, i32, C.POINTER(StepOut), vp]
    L.mgplr_rollout.argtypes = [vp, vp, i32, i32, C.POINTER(StepOut), vp]
    L.mgplr_rollout_ex.argtypes = [vp, vp, i32, i32, i32, C.POINTER(StepOut), vp]
    L.mgplr_full_obs.argtypes = [vp, vp, vp]
    L.mgplr_render_images.argtypes = [vp, vp, vp, i32, vp, vp]
    L.mgplr_get_encodings.argtypes = [vp, vp, vp]
    L.mgplr_get_metrics.argtypes = [vp, vp, vp]
    L.mgplr_get_agent_state.argtypes = [vp, vp, vp]
    L.mgplr_get_errors.argtypes = [vp, vp, i32, vp]
    L.mgplr_peek_rng.argtypes = [vp, i32, vp, i32]
    L.mgplr_gae.argtypes = [vp, vp, vp, vp, i32, i32, f64, f64, vp]
    L.mgplr_discounted_returns.argtypes = [vp, vp, vp, i32, i32, f64, vp]
    L.mgplr_batched_value_loss.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    L.mgplr_plr_episode_scores.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp, vp]
    L.mgplr_plr_episode_scores_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, f64, i32, i32, i32, vp, i32, vp, vp]
    L.mgplr_plr_sample_weights.argtypes = [vp, vp, vp, i32, i32, f64, f64, f64, i32, f64, vp, vp, vp]
    L.mgplr_plr_score_weights.argtypes = [vp, vp, i32, i32, f64, f64, vp, vp]
    L.mgplr_plr_sample_replay.argtypes = [vp, vp, vp, i32, i32, f64, f64, f64, i32, f64, vp, vp, i32, vp, vp]
    L.mgplr_plr_apply_records.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f64, f64, f64,
                                          i32, i32, i32, f64, f64, f64, i32, f64, vp, vp, vp]
    L.mgplr_wide_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    L.mgplr_wide_destroy.argtypes = [vp]
    L.mgplr_wide_destroy.restype = None
    L.mgplr_wide_load_levels.argtypes = [vp, vp, vp, i32, i32, C.POINTER(StepOut), vp]
    L.mgplr_wide_step.argtypes = [vp, vp, C.POINTER(StepOut), vp]
    L.mgplr_wide_get_encodings.argtypes = [vp, vp, vp]
    _lib = L
    return L


def check(rc, what=''):
    if rc != 0:
        msg = load().mgplr_last_error().decode('utf8', 'replace')
        raise MgplrError('%s failed (%d): %s' % (what or 'mgplr call', rc, msg))


def ptr(t):
    """Raw device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        assert t.is_contiguous(), 'tensor passed to libmgplr must be contiguous'
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
