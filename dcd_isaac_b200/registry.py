"""Registered MultiGrid adversarial environments: constructor arguments restated from the reference's
registrations (envs/multigrid/adversarial.py:584-786; defaults at :67-79).  `tl` is the TimeLimit
(`max_episode_steps`) the reference registry wraps around the env (envs/registration.py:118-120)."""

EDITOR_ACTION_SPACES = {  # adversarial.py:40-56
    'walls_none': ('-', '.'),
    'walls_none_goal': ('-', '.', 'g'),
    'walls_none_agent_goal': ('-', '.', 'a', 'g'),
}


def _spec(n_clutter=50, size=15, goal_last=True, see_through=True, max_steps=250, tl=250, resample=False,
          editor='walls_none_agent_goal', fixed=False):
    return dict(n_clutter=n_clutter, size=size, choose_goal_last=goal_last, see_through_walls=see_through,
                max_steps=max_steps, max_episode_steps=tl, resample_n_clutter=resample, editor_actions=editor,
                fixed_environment=fixed)


ENV_SPECS = {
    'MultiGrid-Adversarial-v0': _spec(goal_last=False),
    'MultiGrid-MiniAdversarial-v0': _spec(n_clutter=7, size=6, goal_last=False, max_steps=50, tl=50),
    'MultiGrid-MediumAdversarial-v0': _spec(n_clutter=30, size=10, goal_last=False, max_steps=200, tl=200),
    'MultiGrid-GoalLastAdversarial-v0': _spec(),
    'MultiGrid-GoalLastOpaqueWallsAdversarial-v0': _spec(see_through=False),
    'MultiGrid-GoalLastFewerBlocksAdversarial-v0': _spec(n_clutter=25),
    'MultiGrid-GoalLastFewerBlocksAdversarial-EditWN-v0': _spec(n_clutter=25, editor='walls_none'),
    'MultiGrid-GoalLastFewerBlocksAdversarial-EditWNG-v0': _spec(n_clutter=25, editor='walls_none_goal'),
    'MultiGrid-GoalLastVariableBlocksAdversarialEnv-v0': _spec(n_clutter=60, resample=True),
    'MultiGrid-GoalLastVariableBlocksAdversarialEnv-Edit-v0': _spec(n_clutter=60, resample=True, editor='walls_none_goal'),
    'MultiGrid-GoalLastEmptyAdversarialEnv-Edit-v0': _spec(n_clutter=0, editor='walls_none_goal'),
    'MultiGrid-GoalLastFewerBlocksOpaqueWallsAdversarial-v0': _spec(n_clutter=25, see_through=False),
    'MultiGrid-MiniGoalLastAdversarial-v0': _spec(n_clutter=7, size=6, max_steps=50, tl=50),
    'MultiGrid-FixedAdversarial-v0': _spec(goal_last=False, max_steps=50, tl=50, fixed=True),
    'MultiGrid-EmptyMiniFixedAdversarial-v0': _spec(n_clutter=0, size=6, goal_last=False, max_steps=50, tl=50, fixed=True),
    'MultiGrid-GoalLastAdversarialEnv30-v0': _spec(n_clutter=30, tl=50),
    'MultiGrid-GoalLastAdversarialEnv60-v0': _spec(n_clutter=60, tl=50),
}
# NoisyAdversarial (goal_noise=0.3, adversarial.py:588-590) draws from Python's `random`; not supported.


def env_spec(env_name, **overrides):
    if env_name not in ENV_SPECS:
        raise KeyError('No registered env with id: %s' % env_name)
    spec = dict(ENV_SPECS[env_name])
    spec.update({k: v for k, v in overrides.items() if v is not None})
    return spec
