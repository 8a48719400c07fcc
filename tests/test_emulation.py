"""Kernel-source host emulation vs the oracle (CPU, no GPU): the SAME .cuh sources that nvcc compiles for sm_100a are
compiled with g++ (PAR_FOR -> serial loop) and must reproduce the oracle's LQ blocks and ProxDDP iterates.  This checks
kernel LOGIC (composite-inertia derivative formulas, active-set compaction, block-eliminated KKT, explicit-inverse
Riccati); races and DMMA fragment layouts are covered by the -m gpu tests."""
import numpy as np
import pytest

import emu_lib
import golden_util
from mpc_benchmark_b200 import _abi, problems


def rel(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())


@pytest.mark.parametrize("maker", [problems.cent_standing_problem, problems.full_standing_problem, problems.kino_standing_problem])
def test_lq_blocks(oracle, maker):
    prob = maker(T=6)
    cfg = prob["cfg"]
    nx, n, m, nc = _abi.DIMS[cfg.kind]
    rng = np.random.default_rng(1)
    xs, us = prob["xs"].copy(), prob["us"].copy()
    for k in range(7):
        xs[0, k] = xs[0, k] + rng.normal(size=9) * 0.01 if cfg.kind == 0 else oracle.integrate(xs[0, k], rng.normal(size=56) * 0.02)
    us += rng.normal(size=us.shape)
    e = emu_lib.solve(prob, 1, dump=True, xs=xs, us=us)
    x = xs[0].copy()
    x[0] = prob["x0"][0]
    for k in range(6):
        o = oracle.eval_knot(prob["robot"], cfg, prob["knots"][k], x[k], us[0, k], x[k + 1])
        assert np.abs(np.hstack([o["A"], o["B"]]) - e["AB"][k]).max() < 1e-10
        H = o["H"] + 1e-9 * np.eye(n + m)
        assert np.abs(H - e["H"][k]).max() <= 1e-10 * max(1, np.abs(H).max())
        assert np.abs(o["gap"] - e["gap"][k]).max() < 1e-11 and np.abs(o["h"] - e["h"][k]).max() < 1e-9
        assert abs(o["cost"] - e["scal"][k, 0]) <= 1e-11 * max(1, abs(o["cost"]))


@pytest.mark.parametrize("maker,T", [(problems.cent_standing_problem, 100), (problems.full_standing_problem, 25), (problems.kino_standing_problem, 25),
                                     (problems.cent_standing_problem, 1), (problems.full_standing_problem, 1), (problems.kino_standing_problem, 2)])
def test_cold_solve(oracle, maker, T):
    prob = maker(T=T)
    r = oracle.solve(prob)
    e = emu_lib.solve(prob, 100)
    assert e["info"][0].num_iters == r["info"][0].num_iters and e["info"][0].conv == r["info"][0].conv == 1
    assert rel(e["xs"], r["xs"]) < 1e-9 and rel(e["us"], r["us"]) < 1e-9 and rel(e["K"], r["K"]) < 1e-7


def test_walking_instances_with_active_constraints(oracle):
    """Golden walking fixture: single-support knots, active wrench-cone / box rows, non-zero multipliers, backtracking."""
    prob, z = golden_util.load("walk_full.npz")
    e = emu_lib.solve(prob, 6)
    assert [i.num_iters for i in e["info"]] == list(z["sol_num_iters"])
    assert [i.ls_evals for i in e["info"]] == list(z["sol_ls_evals"])
    assert rel(e["xs"], z["sol_xs"]) < 1e-6 and rel(e["us"], z["sol_us"]) < 1e-6
    assert np.abs(z["sol_vs"]).max() > 1.0  # the fixture really has active constraints
    assert rel(e["vs"], z["sol_vs"]) < 1e-5
    t = emu_lib.solve(prob, 1, xs=z["sol_xs"], us=z["sol_us"])
    assert rel(t["xs"], z["tick_xs"]) < 1e-6 and rel(t["us"], z["tick_us"]) < 1e-6
    assert rel(t["stage0"], z["tick_stage0"]) < 1e-6


@pytest.mark.parametrize("kind,ticks,perturb", [(_abi.KIND_KINO, [99, 60], False), (_abi.KIND_CENT, [99, 150], True)])
def test_reference_gait_walking_kino_cent(oracle, kind, ticks, perturb):
    """BASELINE configs[0]-[1]: horizons of the reference walking loops (kinodynamic_talos.py:183-261, centroidal_talos.py:100-183):
    single-support knots, moving swing references / contact positions, ramping force references, terminal CoM equality."""
    prob = problems.walk_batch(kind, len(ticks), seed=2, ticks=ticks, mirror=[False, True], perturb=perturb)
    assert 0.1 < prob["ds_fraction"] < 0.9  # really mixes single and double support
    r = oracle.solve(prob, max_iters=5, inst_threads=2)
    e = emu_lib.solve(prob, 5)
    assert [i.num_iters for i in e["info"]] == [i.num_iters for i in r["info"]]
    assert [i.ls_evals for i in e["info"]] == [i.ls_evals for i in r["info"]]
    assert rel(e["xs"], r["xs"]) < 1e-7 and rel(e["us"], r["us"]) < 1e-7 and rel(e["vs"], r["vs"]) < 1e-5


def kino_many_active_rows(T=6):
    """A kinodynamic iterate with 59-65 ACTIVE constraint rows per knot (of 68): pulling contact forces put every wrench-cone row
    on the wrong side and every joint sits beyond a limit — more than the 56 rows of the Riccati kernel's fast shared-memory
    carving, so the knots take the carving that keeps [C D] in global memory."""
    prob = problems.kino_standing_problem(batch=1, T=T)
    rb = prob["robot"]
    hi, lo = np.array(rb.q_hi[:]), np.array(rb.q_lo[:])
    xs, us = prob["xs"].copy(), prob["us"].copy()
    rng = np.random.default_rng(0)
    for k in range(1, T + 1):
        xs[0, k, 7:29] = np.where(rng.random(22) < 0.5, hi + 0.05, lo - 0.05)
    us[0, :, 2] = us[0, :, 8] = -300.0
    us[0, :, :12] += rng.normal(size=(T, 12)) * 50
    return prob, xs, us


def test_kino_more_active_rows_than_fast_carving(oracle):
    prob, xs, us = kino_many_active_rows()
    r = oracle.solve(prob, max_iters=2, xs=xs, us=us)
    e = emu_lib.solve(prob, 2, xs=xs, us=us)
    assert e["info"][0].status != 3 and e["info"][0].num_iters == r["info"][0].num_iters == 2
    assert rel(e["xs"], r["xs"]) < 1e-7 and rel(e["us"], r["us"]) < 1e-7 and rel(e["vs"], r["vs"]) < 1e-6


@pytest.mark.parametrize("maker,T", [(problems.cent_standing_problem, 30), (problems.full_standing_problem, 12), (problems.kino_standing_problem, 12)])
def test_nonlinear_rollout_cold_solve(oracle, maker, T):
    """ROLLOUT_NONLINEAR (fused rollout + linesearch, solver_core.cuh rollout_trial) against the oracle's try_step_nonlinear."""
    prob = maker(batch=2, T=T)
    prob["cfg"].rollout = 1
    r = oracle.solve(prob)
    e = emu_lib.solve(prob, 100)
    assert e["info"][1].num_iters == r["info"][1].num_iters and e["info"][1].ls_evals == r["info"][1].ls_evals and e["info"][1].conv == 1
    assert rel(e["xs"], r["xs"]) < 1e-9 and rel(e["us"], r["us"]) < 1e-9 and rel(e["lams"], r["lams"]) < 1e-7


def test_nonlinear_rollout_walking_with_backtracking(oracle):
    """Walking fixture under ROLLOUT_NONLINEAR: active rows (their gains feed back through dx), several backtracking steps."""
    prob, z = golden_util.load("walk_full.npz")
    prob["cfg"].rollout = 1
    r = oracle.solve(prob, max_iters=4)
    e = emu_lib.solve(prob, 4)
    assert [i.ls_evals for i in e["info"]] == [i.ls_evals for i in r["info"]] and max(i.ls_evals for i in r["info"]) > 4
    assert rel(e["xs"], r["xs"]) < 1e-8 and rel(e["us"], r["us"]) < 1e-8 and rel(e["vs"], r["vs"]) < 1e-6
    lin = oracle.solve(dict(prob, cfg=golden_util.load("walk_full.npz")[0]["cfg"]), max_iters=4)
    assert rel(lin["xs"], r["xs"]) > 1e-9  # it really is a different rollout
