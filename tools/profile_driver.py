"""Small driver for ncu captures (GPU box): `batch` full-dynamics walking instances, a short untimed preparation and ONE
warm MPC tick — the same call sequence as bench.py's timed step, at a size ncu's replay passes can afford.
usage: ncu ... python tools/profile_driver.py [batch] [prep_iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    prep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    prob = problems.full_walk_batch(batch, seed=5)
    s = BatchSolver(prob["robot"], prob["cfg"], batch, device=0)
    s.setup(prob["knots"], prob["terms"], prob["x0_nominal"])
    warm = s.run(prob["xs"], prob["us"], max_iters=prep, gains=False)
    s.set_x0(prob["x0"])
    s.reset_multipliers()
    r = s.run(problems.warm_tick_inputs(prob, warm.xs), warm.us, max_iters=1, gains=False)
    print("tick done: launches", s.last_launches, "kernel ms", s.kernel_ms(), "ls evals (mean)", float(r.ls_evals.mean()))
    s.close()


if __name__ == "__main__":
    main()
