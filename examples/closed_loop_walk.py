#!/usr/bin/env python
"""Batched closed-loop MPC on the GPU (ideal plant): 256 Talos-shaped robots, each with its own RANDOM contact schedule, 50 ticks.
The knot appended every tick starts from the control of the nearest knot with the same contact phase (mpc_set_tail_warmstart(1); WALK_TAIL=copy
selects the reference scripts' us[1:] + [us[-1]]).  At the reference's mu_init = 1e-8 with one iteration per tick, 400 ticks of these random schedules:
244 of 256 robots survive with the phase-matched warm start, 12 with the scripts' (the REFERENCE gait is walked to the end by every robot:
examples/reference_gait_walk.py).  `tools/closed_loop_trace.py` prints per-tick statistics.
Usage: python examples/closed_loop_walk.py [batch] [ticks] [mu_init]      (mu_init defaults to 1e-4)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402
from mpc_benchmark_b200.closed_loop import ClosedLoop  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
MU = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
prob = problems.full_walk_batch(B, seed=1, stream_ticks=N, mu_init=MU)  # every robot's gait continues N knots past the horizon
s = BatchSolver(prob["robot"], prob["cfg"], B)
s.setup(prob["knots"], prob["terms"], prob["x0_nominal"])
t0 = time.time()
cold = s.run(prob["xs"], prob["us"], max_iters=30, gains=False)
print(f"cold solve: {time.time() - t0:.2f} s, converged {cold.conv.sum()}/{B}, median prim infeas {np.median(cold.prim_infeas):.2e}")
s.set_tail_warmstart(os.environ.get("WALK_TAIL", "phase") == "phase")  # the appended knot starts from a knot of the same contact phase
loop = ClosedLoop(s, prob["stream"])
t0 = time.time()
res = loop.run(N)
dt = time.time() - t0
print(f"{N} closed-loop ticks x {B} robots: {dt:.2f} s -> {B * N / dt:.0f} robot-ticks/s, median tick {np.median(loop.tick_ms):.1f} ms")
status = np.array([i.status for i in res.info])
print("base height range after the walk:", res.xs[:, 0, 2].min(), res.xs[:, 0, 2].max(), "| non-finite / failed instances:",
      int((status >= 2).sum()), "| median primal infeasibility", float(np.median(res.prim_infeas)))
