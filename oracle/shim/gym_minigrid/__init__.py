"""Stand-in for the subset of gym-minigrid==1.0.1 the reference hot path touches.

TEST INFRASTRUCTURE ONLY (oracle). gym-minigrid is an un-vendored third-party dependency
of the reference (requirements.txt:3), absent from this image. `minigrid.py` restates its
published algorithm for Grid.{get,set,wall_rect,encode,process_vis}, the WorldObj family,
the index tables and MiniGridEnv.{seed,_rand_int,_reward,put_obj,...}; call sites in the
reference: envs/multigrid/multigrid.py:39-40,54,58-61,76-101,110,156,334,341,603-606,694,
899,1001 and envs/multigrid/adversarial.py:31,172,222,343,366,409,494-518,532,564,577.
This is the one link of the parity chain that is restated rather than executed.
"""
from . import minigrid, rendering  # noqa: F401
