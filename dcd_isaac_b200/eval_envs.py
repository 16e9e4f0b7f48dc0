"""Zero-shot evaluation environments: one constructor for what eval.py's Evaluator builds per
test env name (eval.py:176-204: gym_make + optional MultiGridFullyObsWrapper, then the vector-env wrappers)."""
from .mazes import MAZES
from .minigrid_envs import MINIGRID_ENVS, CudaMiniGridVecEnv
from .mst_maze import MST_MAZES, CudaMSTMazeVecEnv, CudaWideMSTMazeVecEnv
from .vec_env import CudaMazeVecEnv


def is_eval_env(env_name):
    """True for the env ids make_eval_venv builds (the drop-in routes only these away from the reference's subprocess envs)."""
    return env_name in MAZES or env_name in MST_MAZES or env_name in MINIGRID_ENVS


def make_eval_venv(env_name, num_processes, device='cuda:0', full_obs=False):
    """env_name: a fixed-bitmap maze of envs/multigrid/maze.py (MultiGrid-SixteenRooms-v0, -Labyrinth-v0, -Maze-v0, ...), a
    Kruskal perfect maze of envs/multigrid/mst_maze.py (Small / Medium / Large / XL), MiniGrid-SimpleCrossingS*N*-v0
    (envs/multigrid/crossing.py) or MiniGrid-FourRooms-v0 (envs/multigrid/fourrooms.py)."""
    if env_name in MAZES:
        return CudaMazeVecEnv(env_name, num_processes, device=device, full_obs=full_obs)
    if env_name in MST_MAZES:
        cls = CudaWideMSTMazeVecEnv if MST_MAZES[env_name] > 32 else CudaMSTMazeVecEnv
        return cls(env_name, num_processes, device=device, full_obs=full_obs)
    if env_name in MINIGRID_ENVS:
        return CudaMiniGridVecEnv(env_name, num_processes, device=device, full_obs=full_obs)
    raise KeyError('No registered env with id: %s' % env_name)
