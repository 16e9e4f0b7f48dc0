"""bench.py contract (CPU side): the reference arm runs without a GPU and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-envs', '2048', '--T', '64', '--envs', '8192', '--no-python-ref'], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and 'workload' in d['config']
    assert 'all of the workload' in d['cpu_baseline']['sample']   # one timed step covers envs x T env-steps


def test_reference_arm_python_leg():
    """cpu_baseline_python: the reference's own util.create_parallel_env + step_env loop, from
    /root/reference here and from the staged copy on the GPU box."""
    import pytest
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('no reference tree')
    env = dict(os.environ, MGPLR_BENCH_PYREF_PROCS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-envs', '1024', '--T', '24', '--envs', '1024'], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][0])
    p = d['cpu_baseline_python']
    assert p['kind'] == 'reference' and p['value'] > 0 and p['runs'][0]['num_processes'] == 4, p


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'], capture_output=True,
                         text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
