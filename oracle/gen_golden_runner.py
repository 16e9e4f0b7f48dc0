"""Runner-level golden fixtures: EXECUTE the reference's envs/runners/adversarial_runner.py (AdversarialRunner.run ->
agent_rollout, :442-635,676-896) unmodified, over the reference's real spawn-subprocess vectorised env
(util.create_parallel_env), its LevelSampler / LevelStore and its RolloutStorage (TEST INFRASTRUCTURE ONLY).

The policy is oracle/scripted_agent.ScriptedAgent (seeded streams instead of a network), everything else is the
reference's code: the mask / bad-mask / cliffhanger bookkeeping of the last rollout step (:521-530,566-573), the
truncated-observation inserts (:546-549), level_seeds (:576-588), replay re-sampling on episode ends (:551-558), the
level store inserts (:402-412) and the sampler update (:616-622).  Each `run()` contributes one snapshot of the
student's rollout storage taken inside `agent.update()` (before after_update), plus the runner / sampler state after it.

  python oracle/gen_golden.py --only runner        (needs /root/reference)
"""
import gzip
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# name -> (command line for the reference's own arguments.parser, number of runner.run() calls, np.random seed)
COMMON = ['--num_env_steps', '1000000', '--test_env_names', '', '--log_interval', '1', '--screenshot_interval', '0']
CASES = {
    # DR, TimeLimit 50 inside a 120-step rollout: truncations (bad_masks), cliffhangers at the rollout end, reset_random
    'runner_dr_tl50_htl': (['--env_name', 'MultiGrid-GoalLastAdversarialEnv30-v0', '--ued_algo', 'domain_randomization',
                            '--num_processes', '6', '--num_steps', '120', '--handle_timelimits', 'true'], 2, 11),
    # the same bookkeeping without proper time limits (no cliffhanger marks, no truncated-value path); env max_steps ==
    # TimeLimit == 50 so 'truncated' is present with value False
    'runner_dr_mini_nohtl': (['--env_name', 'MultiGrid-MiniGoalLastAdversarial-v0', '--ued_algo', 'domain_randomization',
                              '--num_processes', '6', '--num_steps', '64', '--handle_timelimits', 'false'], 2, 12),
    # robust PLR on byte-encoded reset_random levels (use_reset_random_dr): non-replay and replay rollouts, replay
    # re-sampling on every episode end, staging -> working admission, level store reconcile
    'runner_plr_mini_bytes': (['--env_name', 'MultiGrid-MiniGoalLastAdversarial-v0', '--ued_algo', 'domain_randomization',
                               '--use_plr', 'true', '--use_reset_random_dr', 'true', '--num_processes', '4', '--num_steps', '64',
                               '--handle_timelimits', 'true', '--level_replay_strategy', 'value_l1',
                               '--level_replay_score_transform', 'rank', '--level_replay_temperature', '0.3',
                               '--level_replay_rho', '0.5', '--level_replay_prob', '0.6', '--level_replay_schedule', 'fixed',
                               '--level_replay_seed_buffer_size', '8', '--staleness_coef', '0.3',
                               '--no_exploratory_grad_updates', 'true'], 10, 13),
    # PLR on adversary-built levels stored as ACTION STRINGS (random adversary, the shipped mg_25b_robust_plr setup):
    # venv.reset / step_adversary through the runner, RolloutStorage.get_action_traj(as_string=True), reset_to_level(str)
    'runner_plr_mini_str': (['--env_name', 'MultiGrid-MiniGoalLastAdversarial-v0', '--ued_algo', 'domain_randomization',
                             '--use_plr', 'true', '--num_processes', '4', '--num_steps', '64', '--handle_timelimits', 'true',
                             '--level_replay_strategy', 'positive_value_loss', '--level_replay_score_transform', 'rank',
                             '--level_replay_temperature', '0.3', '--level_replay_rho', '0.5', '--level_replay_prob', '0.6',
                             '--level_replay_schedule', 'fixed', '--level_replay_seed_buffer_size', '8',
                             '--staleness_coef', '0.3', '--no_exploratory_grad_updates', 'true'], 10, 14),
    # the shipped 25-block robust-PLR configuration's env / scoring (MaxMC) at full rollout length: TimeLimit 250 in T=256
    'runner_plr_fb15_maxmc': (['--env_name', 'MultiGrid-GoalLastFewerBlocksAdversarial-v0', '--ued_algo', 'domain_randomization',
                               '--use_plr', 'true', '--use_reset_random_dr', 'true', '--num_processes', '4', '--num_steps', '256',
                               '--handle_timelimits', 'true', '--level_replay_strategy', 'grounded_signed_value_loss',
                               '--level_replay_score_transform', 'rank', '--level_replay_temperature', '0.1',
                               '--level_replay_rho', '0.5', '--level_replay_prob', '0.5', '--level_replay_schedule', 'fixed',
                               '--level_replay_seed_buffer_size', '8', '--staleness_coef', '0.3',
                               '--no_exploratory_grad_updates', 'true'], 6, 15),
}


def parse_args(argv):
    """The reference's own argument parser (arguments.py) on a fixture's command line."""
    from arguments import parser
    return parser.parse_args(COMMON + list(argv))


def build_runner(args, venv, ued_venv, storage_cls, runner_cls, plr_args, device='cpu'):
    """What train.py:72-107 builds, with ScriptedAgents in place of make_agent's networks.  `storage_cls`,
    `runner_cls` come from the reference here and from the drop-in / staged reference in the GPU tests."""
    from scripted_agent import ScriptedAgent
    N = args.num_processes
    htl = bool(args.handle_timelimits) and venv.get_max_episode_steps() is not None   # util/make_agent.py:205-208

    def storage(obs_space, action_space, steps, proper):
        st = storage_cls(model=None, num_steps=steps, num_processes=N, observation_space=obs_space,
                         action_space=action_space, recurrent_hidden_state_size=1, recurrent_arch='rnn',
                         use_proper_time_limits=proper)
        st.to(device)
        return st

    agent = ScriptedAgent(storage(venv.observation_space, venv.action_space, args.num_steps, htl), 7, N, seed=100 + args.seed,
                          forward_bias=0.45)
    adversary_env = None
    if args.ued_algo == 'domain_randomization' and args.use_plr and not args.use_reset_random_dr:   # train.py:85-87
        adv_steps = int(venv.adversary_observation_space['time_step'].high[0])
        adversary_env = ScriptedAgent(storage(venv.adversary_observation_space, venv.adversary_action_space, adv_steps, False),
                                      venv.adversary_action_space.n, N, seed=200 + args.seed)
    runner = runner_cls(args=args, venv=venv, agent=agent, ued_venv=ued_venv, adversary_agent=None,
                        adversary_env=adversary_env, flexible_protagonist=False, train=True, plr_args=plr_args, device=device)
    return runner, agent, adversary_env


def sampler_state(s):
    import numpy as np
    return dict(seeds=np.array(s.seeds).copy(), seed_scores=np.array(s.seed_scores).copy(),
                seed_staleness=np.array(s.seed_staleness).copy(), unseen_seed_weights=np.array(s.unseen_seed_weights).copy(),
                working_seed_set=sorted(int(x) for x in s.working_seed_set), staging_seed_set=sorted(int(x) for x in s.staging_seed_set),
                running_sample_count=int(s.running_sample_count), working_seed_buffer_size=int(s.working_seed_buffer_size))


def run_case(name, runner, agent, n_runs, seed):
    """Drive `n_runs` x AdversarialRunner.run() and collect everything a drop-in must reproduce."""
    import numpy as np
    np.random.seed(seed)
    runs = []
    for _ in range(n_runs):
        stats = runner.run()
        rec = dict(storage=agent.snapshots[-1], n_snapshots=len(agent.snapshots),
                   level_replay=bool(runner.sampled_level_info['level_replay']) if runner.sampled_level_info else False,
                   current_level_seeds=None if runner.current_level_seeds is None else [int(s) for s in runner.current_level_seeds],
                   total_episodes=int(runner.total_episodes_collected), total_seeds=int(runner.total_seeds_collected),
                   student_grad_updates=int(runner.student_grad_updates),
                   mean_agent_return=float(stats['mean_agent_return']),
                   stats={k: (None if stats.get(k) is None else float(stats[k])) for k in
                          ('num_blocks', 'passable_ratio', 'shortest_path_length', 'solved_path_length')},
                   encodings=np.stack(runner.venv.get_encodings()))
        if runner.level_samplers:
            rec['sampler'] = sampler_state(runner.level_samplers['agent'])
            rec['store_seeds'] = sorted(int(s) for s in runner.level_store.seed2level.keys())
            rec['store_levels'] = {int(s): (bytes(l) if isinstance(l, (bytes, bytearray)) else str(l))
                                   for s, l in runner.level_store.seed2level.items()}
        runs.append(rec)
    return runs


def gen_runner():
    rh.activate()
    import util
    from algos.storage import RolloutStorage
    from envs.runners.adversarial_runner import AdversarialRunner
    for name, (argv, n_runs, seed) in CASES.items():
        args = parse_args(argv)
        venv, ued_venv = util.create_parallel_env(args)
        plr_args = util.make_plr_args(args, venv.observation_space, venv.action_space) if args.use_plr else None
        runner, agent, _ = build_runner(args, venv, ued_venv, RolloutStorage, AdversarialRunner, plr_args)
        runs = run_case(name, runner, agent, n_runs, seed)
        venv.close()
        with gzip.open(os.path.join(GOLDEN, name + '.pkl.gz'), 'wb') as f:
            pickle.dump(dict(argv=list(argv), n_runs=n_runs, np_seed=seed, runs=runs), f, protocol=4)
        st = runs[-1]['storage']
        print('runner', name, 'runs', n_runs, 'replay', [int(r['level_replay']) for r in runs],
              'episodes', runs[-1]['total_episodes'], 'bad0', int((st['bad_masks'] == 0).sum()),
              'cliff0', int((st['cliffhanger_masks'] == 0).sum()), 'goals', int((st['rewards'] != 0).sum()))


if __name__ == '__main__':
    gen_runner()
