"""Stage an UNMODIFIED copy of the reference tree under the git-ignored baseline/_ref/reference (TEST / BENCH
INFRASTRUCTURE ONLY).

/root/reference does not exist on the GPU box; baseline/_ref/ is git-ignored (nothing of the reference enters the
history) but travels with the gpurun snapshot like the built .so files.  The staged copy is what
  * tests/test_gpu_reference_runner.py imports to run the reference's own AdversarialRunner / train.py / eval.py
    against the drop-in objects, and
  * bench.py's `cpu_baseline_python` leg times (util.create_parallel_env -> step_env loop on the box's host cores).
docs/ (images) is skipped.  `python oracle/stage_reference.py` or __graft_entry__.build() runs this when
/root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get('DCD_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref', 'reference')


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, 'envs', 'multigrid')):
        return None
    stamp = os.path.join(DST, '.staged')
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('docs', '.git', '__pycache__', '*.pyc'))
    with open(stamp, 'w') as f:
        f.write('unmodified copy of %s (docs/ skipped)\n' % SRC)
    return DST


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv))
