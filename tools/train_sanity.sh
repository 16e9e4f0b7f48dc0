#!/bin/bash
# Short DR vs robust-PLR training runs of the reference's UNMODIFIED train.py on the drop-in, shipped 25-block configs
# (train_scripts/grid_configs/minigrid/25_blocks/mg_25b_{dr,robust_plr}.json), 3 seeds each, all six concurrently for a fixed wall
# time.   gpurun --timeout 1200 -- 'bash tools/train_sanity.sh 540'
WALL=${1:-540}
O=$PWD/gpurun_out/r02/train; mkdir -p $O
REF=$PWD/baseline/_ref/reference
export PYTHONPATH=$PWD/oracle/shim:$REF:$PWD:$PYTHONPATH OMP_NUM_THREADS=2 TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1
COMMON="--env_name MultiGrid-GoalLastFewerBlocksAdversarial-v0 --ued_algo domain_randomization --num_processes 32 --num_env_steps 250000000 --num_steps 256 --ppo_epoch 5 --num_mini_batch 1 --handle_timelimits true --lr 1e-4 --gamma 0.995 --entropy_coef 0.01 --recurrent_arch lstm --recurrent_agent true --recurrent_adversary_env false --recurrent_hidden_size 256 --use_plr true --log_interval 10 --log_action_complexity true --log_plr_buffer_stats true --log_replay_complexity true --reject_unsolvable_seeds false --disable_checkpoint true --test_interval 100 --test_num_episodes 10 --screenshot_interval 0 --log_dir $O/logs"
DR="--level_replay_prob 0.0"
RPLR="--level_replay_prob 0.5 --level_replay_rho 0.5 --level_replay_temperature 0.1 --level_replay_seed_buffer_size 4000 --level_replay_score_transform rank --staleness_coef 0.3 --level_replay_strategy grounded_signed_value_loss --no_exploratory_grad_updates true"
cd $O
for seed in 1 2 3; do
  timeout -s INT $WALL python -m dcd_isaac_b200.dropin $REF/train.py $COMMON $DR --seed $seed --xpid dr_s$seed > $O/dr_s$seed.out 2>&1 &
  timeout -s INT $WALL python -m dcd_isaac_b200.dropin $REF/train.py $COMMON $RPLR --seed $seed --xpid rplr_s$seed > $O/rplr_s$seed.out 2>&1 &
done
wait
for x in dr_s1 dr_s2 dr_s3 rplr_s1 rplr_s2 rplr_s3; do
  cp $O/logs/$x/logs.csv $O/$x.csv 2>/dev/null; tail -2 $O/$x.out | cut -c1-300
  python - "$O/$x.csv" <<'PY'
import csv, sys
rows = [r for r in csv.DictReader(open(sys.argv[1])) if r.get('steps') not in (None, '', 'steps')]
if rows:
    k = [c for c in ('steps', 'sps', 'mean_agent_return', 'agent_plr_passable_mass', 'plr_passable_mass', 'solved_rate:MultiGrid-SixteenRooms-v0', 'solved_rate:MultiGrid-Labyrinth-v0') if c in rows[0]]
    for r in rows[:: max(1, len(rows) // 6)] + [rows[-1]]:
        print({c: r[c] for c in k})
PY
done
rm -rf $O/logs/*/screenshots
