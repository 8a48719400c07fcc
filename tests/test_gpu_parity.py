"""GPU parity tests proper: the CUDA path, called through the C-ABI (BatchSolver -> libmpcb200.so), against the CPU
oracle on the same seeded inputs.  Tolerance: 1e-6 relative on trajectories (BASELINE.json north_star), iteration
count equal (north_star allows +-1)."""
import numpy as np
import pytest

from mpc_benchmark_b200 import _abi, problems
from mpc_benchmark_b200.batch import BatchSolver

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def rel(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())


def perturbed(prob, oracle, seed, sx=0.02, su=1.0):
    rng = np.random.default_rng(seed)
    xs, us = prob["xs"].copy(), prob["us"].copy()
    kind = prob["cfg"].kind
    for b in range(xs.shape[0]):
        for k in range(xs.shape[1]):
            if kind == _abi.KIND_CENT:
                xs[b, k] += rng.normal(size=9) * sx
            else:
                xs[b, k] = oracle.integrate(xs[b, k], rng.normal(size=56) * sx)
    us += rng.normal(size=us.shape) * su
    return xs, us


@pytest.mark.parametrize("maker", [problems.cent_standing_problem, problems.full_standing_problem])
def test_lq_blocks_match_oracle(oracle, maker):
    prob = maker(batch=2, T=8)
    cfg = prob["cfg"]
    nx, n, m, nc = _abi.DIMS[cfg.kind]
    xs, us = perturbed(prob, oracle, 3)
    s = BatchSolver(prob["robot"], cfg, 2)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    for inst in range(2):
        d = s.debug_lq(xs, us, inst)
        x = xs[inst].copy()
        x[0] = prob["x0"][inst]
        for k in range(cfg.T):
            o = oracle.eval_knot(prob["robot"], cfg, prob["knots"][inst * cfg.T + k], x[k], us[inst, k], x[k + 1])
            AB = np.hstack([o["A"], o["B"]])
            assert np.abs(AB - d["AB"][k]).max() < 1e-9
            H = o["H"] + 1e-9 * np.eye(n + m)
            assert np.abs(H - d["H"][k]).max() <= 1e-9 * max(1.0, np.abs(H).max())
            assert np.abs(o["gap"] - d["gap"][k]).max() < 1e-10
            assert np.abs(o["h"] - d["h"][k]).max() < 1e-8
            assert abs(o["cost"] - d["scal"][k, 0]) <= 1e-10 * max(1.0, abs(o["cost"]))
    s.close()


@pytest.mark.parametrize("maker,T,iters", [(problems.cent_standing_problem, 100, 100), (problems.full_standing_problem, 20, 100),
                                           (problems.full_standing_problem, 100, 100)])
def test_cold_solve_matches_oracle(oracle, maker, T, iters):
    prob = maker(batch=2, T=T)
    s = BatchSolver(prob["robot"], prob["cfg"], 2)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=iters)
    ref = oracle.solve(prob, max_iters=iters, inst_threads=2)
    assert list(res.num_iters) == [i.num_iters for i in ref["info"]]
    assert list(res.conv) == [bool(i.conv) for i in ref["info"]]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL
    assert rel(res.K, ref["K"]) < 1e-5
    for b in range(2):
        assert abs(res.prim_infeas[b] - ref["info"][b].prim_infeas) <= 1e-6 * max(1e-6, ref["info"][b].prim_infeas) + 1e-12
    s.close()


def test_mpc_tick_matches_oracle(oracle):
    """One-iteration warm-started tick (max_iters = 1, fulldynamic_talos.py:407,532-540) from a perturbed state."""
    prob = problems.full_standing_problem(batch=3, T=30)
    xs, us = perturbed(prob, oracle, 5, sx=0.005, su=0.5)
    prob["x0"] = xs[:, 0].copy()
    s = BatchSolver(prob["robot"], prob["cfg"], 3)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(xs, us, max_iters=1)
    ref = oracle.solve(prob, max_iters=1, xs=xs, us=us)
    assert list(res.num_iters) == [1, 1, 1]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL
    xdot, force = s.stage_data(0)
    assert rel(force, ref["stage0"][:, 56:]) < RTOL
    s.close()


def test_dmma_tile_gemm_matches_numpy():
    """FP64 tensor-core tile GEMM (mma.sync m8n8k4) used by the Riccati kernel, against numpy."""
    import ctypes as C

    from mpc_benchmark_b200 import _native

    rng = np.random.default_rng(0)
    for mt, nt, K, lda, ldb, ldc in [(7, 7, 56, 56, 56, 56), (7, 10, 56, 56, 88, 88), (10, 10, 56, 88, 88, 80), (1, 3, 4, 8, 24, 24)]:
        A = np.ascontiguousarray(rng.normal(size=(K, lda)))
        B = np.ascontiguousarray(rng.normal(size=(K, ldb)))
        Cm = np.zeros((8 * mt, ldc))
        _native.check(_native.lib().mpc_debug_gemm_tn(mt, nt, K, _native.ptr(A), lda, _native.ptr(B), ldb, _native.ptr(Cm), ldc), "gemm")
        ref = A[:, : 8 * mt].T @ B[:, : 8 * nt]
        assert np.abs(Cm[:, : 8 * nt] - ref).max() < 1e-12 * K
