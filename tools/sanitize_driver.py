"""Tiny solves of the three model kinds for compute-sanitizer (memcheck / racecheck) runs on the GPU box.
usage: compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402


def run(name, prob, iters):
    b = prob["x0"].shape[0]
    s = BatchSolver(prob["robot"], prob["cfg"], b, device=0)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    r = s.run(prob["xs"], prob["us"], max_iters=iters)
    print(name, "iters", r.num_iters, "prim", r.prim_infeas, "ls", r.ls_evals)
    s.close()


def main():
    run("full walk", problems.full_walk_batch(2, seed=3, T=12), 3)
    run("kino standing", problems.kino_standing_problem(batch=2, T=10), 2)
    run("cent standing", problems.cent_standing_problem(batch=2, T=20), 2)


if __name__ == "__main__":
    main()
