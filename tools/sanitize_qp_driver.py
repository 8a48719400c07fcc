"""Small batched QP solves for compute-sanitizer (memcheck / racecheck / synccheck) runs on the GPU box: random QPs with and without
box constraints, the whole-body fixture through IDSolver_ulim (device assembly) and through solve_from_state (k_rbd_terms).
usage: compute-sanitizer --tool racecheck python tools/sanitize_qp_driver.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_benchmark_b200 import pin, problems, proxqp, qp_utils  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402


def random_qps(n, ne, ni, batch, box, seed):
    rng = np.random.default_rng(seed)
    Lm = rng.normal(size=(batch, n, n))
    H = Lm @ Lm.transpose(0, 2, 1) / n + 0.1 * np.eye(n)
    g, A, x0 = rng.normal(size=(batch, n)), rng.normal(size=(batch, ne, n)), rng.normal(size=(batch, n))
    Cm = rng.normal(size=(batch, ni, n))
    s = np.einsum("bij,bj->bi", Cm, x0)
    qp = proxqp.dense.BatchQP(n, ne, ni, batch, box)
    qp.settings.eps_abs, qp.settings.max_iter, qp.settings.max_iter_in = 1e-6, 50, 30
    qp.init(H, g, A, np.einsum("bij,bj->bi", A, x0), Cm, s - 0.5, s + 0.5, x0 - 0.3 if box else None, x0 + 0.3 if box else None)
    r = qp.solve()
    print(f"random n {n} n_eq {ne} n_in {ni} box {box}: status {r.info.status.tolist()} newton {r.info.iter.tolist()}")
    qp.close()


def main():
    random_qps(20, 8, 12, 3, False, 1)
    random_qps(13, 5, 7, 3, True, 2)
    random_qps(62, 40, 18, 2, True, 3)
    d = np.load(os.path.join(ROOT, "tests", "golden", "qp_id_talos.npz"))
    B = 4
    s = qp_utils.IDSolver_ulim(pin.load_talos_like()[0], [1, 1], 2, 0.8, 0.1, 0.075, [0, 1], 6, False, batch=B)
    rbd = qp_utils.RBDTerms(nle=d["nle"][:B], Jc=d["Jc"][:B], dJv=d["dJv"][:B], vf=d["vf"][:B])
    s.solve(rbd, d["cs"][:B], None, d["a"][:B], d["forces"][:B], d["M"][:B])
    print("whole-body ID: status", s.qp.results.info.status.tolist(), "newton", s.qp.results.info.iter.tolist())
    pr = problems.full_standing_problem(batch=1, T=4)
    bs = BatchSolver(pr["robot"], pr["cfg"], 1)
    s.solve_from_state(bs, d["x"][:B], d["cs"][:B], d["a"][:B], d["forces"][:B])
    print("from state: status", s.qp.results.info.status.tolist(), "newton", s.qp.results.info.iter.tolist())
    bs.close()


if __name__ == "__main__":
    main()
