"""Per-tick statistics of a batched closed-loop walk (ideal plant): primal / dual infeasibility, accepted step lengths,
linesearch trials.  usage: python tools/closed_loop_trace.py [batch] [ticks] [iters_per_tick] [keep_multipliers] [mu_init]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402
from mpc_benchmark_b200.closed_loop import ClosedLoop  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
IT = int(sys.argv[3]) if len(sys.argv) > 3 else 1
KEEP = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
MU = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-8
prob = problems.full_walk_batch(B, seed=1, stream_ticks=N, mu_init=MU)
s = BatchSolver(prob["robot"], prob["cfg"], B)
s.setup(prob["knots"], prob["terms"], prob["x0_nominal"])
cold = s.run(prob["xs"], prob["us"], max_iters=40, gains=False)
print("cold: conv", int(cold.conv.sum()), "/", B, "prim", float(np.median(cold.prim_infeas)))
loop = ClosedLoop(s, prob["stream"], keep_multipliers=KEEP)
for t in range(N):
    loop.step(max_iters=IT)
    if t % 10 == 9 or t < 3:
        r = s.results(gains=False, multipliers=False)
        st = np.array([i.status for i in r.info])
        print(f"tick {t + 1:4d}: prim med {np.median(r.prim_infeas):9.2e} max {r.prim_infeas.max():9.2e} | dual med {np.median(r.dual_infeas):9.2e} | "
              f"alpha med {np.median(r.alpha):.3f} min {r.alpha.min():.2e} | ls mean {r.ls_evals.mean():.2f} | failed {int((st >= 2).sum())} | "
              f"base z {r.xs[:, 0, 2].min():.3f}..{r.xs[:, 0, 2].max():.3f}")
