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


@pytest.mark.parametrize("maker", [problems.cent_standing_problem, problems.full_standing_problem, problems.kino_standing_problem])
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
                                           (problems.full_standing_problem, 100, 100), (problems.kino_standing_problem, 100, 100)])
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


@pytest.mark.parametrize("name", ["ref_flat_full.npz", "ref_flat_kino.npz", "ref_flat_cent.npz"])
def test_golden_reference_script_problems(name):
    """Descriptors flattened from the unmodified reference scripts (tests/golden/make_golden.py): cold solve on the GPU vs the
    committed oracle solution — same iteration count, trajectories within 1e-6 relative."""
    import golden_util

    prob, z = golden_util.load(name)
    s = BatchSolver(prob["robot"], prob["cfg"], 1)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"])
    assert list(res.num_iters) == list(z["sol_num_iters"]) and list(res.conv) == [bool(c) for c in z["sol_conv"]]
    assert rel(res.xs, z["sol_xs"]) < RTOL and rel(res.us, z["sol_us"]) < RTOL
    assert abs(res.prim_infeas[0] - z["sol_prim_infeas"][0]) < 1e-6 * max(1e-6, z["sol_prim_infeas"][0]) + 1e-10
    assert abs(res.dual_infeas[0] - z["sol_dual_infeas"][0]) < 1e-4 * max(1e-6, z["sol_dual_infeas"][0]) + 1e-10
    s.close()


def test_golden_walking_active_constraints():
    """Random-schedule walking instances: single support, active cone/box rows, non-zero multipliers, backtracking linesearch."""
    import golden_util

    prob, z = golden_util.load("walk_full.npz")
    B = prob["x0"].shape[0]
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=6)
    assert list(res.num_iters) == list(z["sol_num_iters"])
    assert [i.ls_evals for i in res.info] == list(z["sol_ls_evals"])
    assert rel(res.xs, z["sol_xs"]) < RTOL and rel(res.us, z["sol_us"]) < RTOL and rel(res.vs, z["sol_vs"]) < 1e-5
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    tick = s.run(z["sol_xs"], z["sol_us"], max_iters=1)
    assert rel(tick.xs, z["tick_xs"]) < RTOL and rel(tick.us, z["tick_us"]) < RTOL
    xdot, force = s.stage_data(0)
    assert rel(force, z["tick_stage0"][:, 56:]) < RTOL and rel(xdot, z["tick_stage0"][:, :56]) < RTOL
    s.close()


def test_large_batch_properties(oracle):
    """Size-independent checks at bench scale (batch 512 here): (1) instance results do not depend on the batch they are
    solved in or on their position in it (independent units): bitwise equality; (2) a sampled instance equals the oracle."""
    B = 512
    prob = problems.full_walk_batch(B, seed=21, T=100)
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=2, gains=False)
    s.close()
    pick = [0, 77, 300, 511]
    sub = dict(prob)
    T = prob["cfg"].T
    sub["knots"] = (_abi.Knot * (len(pick) * T))(*[prob["knots"][i * T + k] for i in pick for k in range(T)])
    sub["terms"] = (_abi.Term * len(pick))(*[prob["terms"][i] for i in pick])
    sub["x0"], sub["xs"], sub["us"] = prob["x0"][pick], prob["xs"][pick], prob["us"][pick]
    s2 = BatchSolver(sub["robot"], sub["cfg"], len(pick))
    s2.setup(sub["knots"], sub["terms"], sub["x0"])
    r2 = s2.run(sub["xs"], sub["us"], max_iters=2, gains=False)
    s2.close()
    assert np.array_equal(r2.xs, res.xs[pick]) and np.array_equal(r2.us, res.us[pick])
    ref = oracle.solve(sub, max_iters=2, inst_threads=4)
    assert rel(r2.xs, ref["xs"]) < RTOL and rel(r2.us, ref["us"]) < RTOL
    assert np.isfinite(res.xs).all() and (res.num_iters == 2).all()


def test_shim_end_to_end_centroidal(oracle):
    """The aligator-compatible surface on the GPU: build the centroidal problem through the public API, setup / run /
    results / workspace, horizon cycling — against the oracle on the flattened descriptor."""
    import mpc_benchmark_b200 as aligator
    from mpc_benchmark_b200 import flatten, pin
    from test_reference_scripts import _build_cent

    ns = _build_cent(aligator, pin, T=20)
    problem = ns["problem"]
    solver = aligator.SolverProxDDP(1e-5, 1e-8)
    solver.rollout_type = aligator.ROLLOUT_LINEAR
    solver.linear_solver_choice = aligator.LQ_SOLVER_PARALLEL
    solver.force_initial_condition = True
    solver.setNumThreads(2)
    solver.max_iters = 100
    solver.setup(problem)
    conv = solver.run(problem, [ns["x0"]] * 21, [ns["u0"] for _ in range(20)])
    flat = flatten.flatten_problem(problem, 1e-5, 1e-8, 100)
    ref = oracle.solve(dict(robot=flat.robot, cfg=flat.cfg, knots=flat.knots, terms=flat.terms, x0=flat.x0,
                            xs=np.tile(ns["x0"], (1, 21, 1)), us=np.tile(ns["u0"], (1, 20, 1))))
    r = solver.results
    assert conv and r.num_iters == ref["info"][0].num_iters
    assert rel(np.array(r.xs.tolist()), ref["xs"][0]) < RTOL and rel(np.array(r.us.tolist()), ref["us"][0]) < RTOL
    K0 = r.controlFeedbacks()[0]
    assert K0.shape == (12, 9) and rel(K0, ref["K"][0, 0]) < 1e-5
    xdot = solver.workspace.problem_data.stage_data[0].dynamics_data.continuous_data.xdot
    assert xdot.shape == (9,)
    # one MPC tick: rotate the horizon, shift the warm start, one iteration (centroidal_talos.py:454-462)
    xs, us = r.xs.tolist(), r.us.tolist()
    problem.replaceStageCircular(ns["stages"][0])
    solver.cycleProblem(problem, ns["stages"][0].createData())
    solver.max_iters = 1
    xs, us = xs[1:] + [xs[-1]], us[1:] + [us[-1]]
    problem.x0_init = xs[0]
    solver.setup(problem)
    solver.run(problem, xs, us)
    assert solver.results.num_iters <= 1 and np.isfinite(np.array(solver.results.xs.tolist())).all()


def test_closed_loop_ticks_match_host_driven_oracle(oracle):
    """Device-side closed-loop tick (mpc_tick: horizon rotation, warm-start shift, x0 from the model prediction) against the
    same loop driven on the host with the oracle (fulldynamic_talos.py:496-497,532-540)."""
    from mpc_benchmark_b200.closed_loop import ClosedLoop

    B, T = 3, 30
    prob = problems.full_walk_batch(B, seed=3, T=T, stream_ticks=3)  # every gait continues past the horizon
    assert bytes(prob["knots"]) == bytes(problems.full_walk_batch(B, seed=3, T=T)["knots"])

    def term_at(t):  # terminal references swapped every tick (fulldynamic_talos.py:499-510): the CoM target drifts forward
        com = np.array(prob["com0"], float)
        com[0] += 0.002 * (t + 1)
        return (_abi.Term * B)(*[problems.make_term(prob["lf"], prob["rf"], com)] * B)

    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    cold = s.run(prob["xs"], prob["us"], max_iters=8)
    ref = oracle.solve(prob, max_iters=8, inst_threads=3)
    assert rel(cold.xs, ref["xs"]) < RTOL
    loop = ClosedLoop(s, prob["stream"], term_stream=term_at)
    xs, us = ref["xs"], ref["us"]
    hp = dict(prob)
    knots = list(prob["knots"])
    for tick in range(3):
        loop.step(max_iters=1)
        # host-driven reference: rotate knots, swap the terminal, shift warm start, x0 <- previous xs[1]
        nxt = prob["stream"](tick)
        for b in range(B):
            knots[b * T:(b + 1) * T] = knots[b * T + 1:(b + 1) * T] + [nxt[b]]
        hp["knots"] = (_abi.Knot * (B * T))(*knots)
        hp["terms"] = term_at(tick)
        hp["x0"] = xs[:, 1].copy()
        xs_ws = np.concatenate([xs[:, 1:], xs[:, -1:]], axis=1)
        us_ws = np.concatenate([us[:, 1:], us[:, -1:]], axis=1)
        r = oracle.solve(hp, max_iters=1, inst_threads=3, xs=xs_ws, us=us_ws)
        xs, us = r["xs"], r["us"]
        got = s.results(gains=False, multipliers=False)
        assert rel(got.xs, xs) < RTOL and rel(got.us, us) < RTOL, tick
    s.close()


def test_bench_workload_tick_matches_oracle(oracle):
    """The bench's timed step on a few walking instances (bench.py): warm start = solve of the NOMINAL problem,
    multipliers reset (per-tick solver.setup), measured state forced at knot 0, one iteration.  Repeating the tick after
    mpc_reset_multipliers must reproduce it bit for bit (no state leaks between ticks)."""
    B, T = 4, 60
    prob = problems.full_walk_batch(B, seed=9, T=T)
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0_nominal"])
    warm = s.run(prob["xs"], prob["us"], max_iters=10, gains=False)
    nominal = dict(prob, x0=prob["x0_nominal"])
    wref = oracle.solve(nominal, max_iters=10, inst_threads=B)
    assert rel(warm.xs, wref["xs"]) < RTOL and rel(warm.us, wref["us"]) < RTOL
    xs_in = problems.warm_tick_inputs(prob, wref["xs"])
    assert np.array_equal(xs_in[:, 0], prob["x0"]) and np.array_equal(xs_in[:, 1:], wref["xs"][:, 1:])
    s.set_x0(prob["x0"])
    s.reset_multipliers()
    t1 = s.run(xs_in, wref["us"], max_iters=1)
    ref = oracle.solve(prob, max_iters=1, inst_threads=B, xs=xs_in, us=wref["us"])
    assert list(t1.num_iters) == [1] * B
    assert list(t1.ls_evals) == [i.ls_evals for i in ref["info"]]
    assert rel(t1.xs, ref["xs"]) < RTOL and rel(t1.us, ref["us"]) < RTOL and rel(t1.K, ref["K"]) < 1e-5
    s.reset_multipliers()
    t2 = s.run(xs_in, wref["us"], max_iters=1)
    assert np.array_equal(t1.xs, t2.xs) and np.array_equal(t1.us, t2.us) and np.array_equal(t1.vs, t2.vs)
    s.close()


def test_call_order_and_argument_errors():
    """Error behaviour of the C-ABI: non-zero status + message, raised as NativeError by the host layer."""
    from mpc_benchmark_b200 import _native

    prob = problems.cent_standing_problem(batch=2, T=10)
    s = BatchSolver(prob["robot"], prob["cfg"], 2)
    with pytest.raises(_native.NativeError):
        s.run(prob["xs"], prob["us"], max_iters=1)  # run before setup
    with pytest.raises(_native.NativeError):
        s.reset_multipliers()
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    with pytest.raises(_native.NativeError):
        s.update_knots((_abi.Knot * 10)(*[prob["knots"][0]] * 10), 8, 5)  # range beyond the horizon
    r0 = s.run(prob["xs"], prob["us"], max_iters=0)  # nothing to do: inputs come back unchanged
    assert list(r0.num_iters) == [0, 0] and np.array_equal(r0.us, prob["us"])
    s.close()


def test_non_finite_input_is_contained():
    """A NaN initial state flags that instance (status 2, mpc_info_t) and terminates it; its neighbours in the batch solve
    exactly as they do alone (independent units, no contamination through shared lists / reductions)."""
    B, T = 3, 20
    prob = problems.full_standing_problem(batch=B, T=T)
    bad = dict(prob)
    bad["x0"] = prob["x0"].copy()
    bad["x0"][1, 10] = np.nan
    s = BatchSolver(bad["robot"], bad["cfg"], B)
    s.setup(bad["knots"], bad["terms"], bad["x0"])
    r = s.run(bad["xs"], bad["us"], max_iters=5)
    status = [i.status for i in r.info]
    assert status[1] == 2 and status[0] != 2 and status[2] != 2
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    good = s.run(prob["xs"], prob["us"], max_iters=5)
    assert np.array_equal(r.xs[0], good.xs[0]) and np.array_equal(r.us[2], good.us[2])
    s.close()


@pytest.mark.parametrize("maker", [problems.cent_standing_problem, problems.full_standing_problem, problems.kino_standing_problem])
def test_shortest_horizons(oracle, maker):
    """Edge sizes: horizons of 1 and 2 knots, batch 1 (loop bounds, terminal-only coupling, prefetch one knot ahead)."""
    for T in (1, 2):
        prob = maker(batch=1, T=T)
        xs, us = perturbed(prob, oracle, 11 + T, sx=0.003, su=0.3)
        prob["x0"] = xs[:, 0].copy()
        s = BatchSolver(prob["robot"], prob["cfg"], 1)
        s.setup(prob["knots"], prob["terms"], prob["x0"])
        res = s.run(xs, us, max_iters=3)
        ref = oracle.solve(prob, max_iters=3, xs=xs, us=us)
        assert list(res.num_iters) == [i.num_iters for i in ref["info"]]
        assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL, T
        s.close()


def test_single_support_only_horizon(oracle):
    """A horizon that is in single support throughout (one rigid contact in every knot, 61 constraint rows)."""
    B, T = 2, 15
    prob = problems.full_standing_problem(batch=B, T=T)
    f_full = np.array([0, 0, prob["mass"] * problems.GRAVITY, 0, 0, 0.0])
    zero = np.zeros(6)
    lift = np.array(prob["rf"], float)
    lift[11] += 0.02
    for b in range(B):
        for k in range(T):
            prob["knots"][b * T + k] = problems.full_knot([True, False], prob["lf"], lift, f_full, zero)
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=4)
    ref = oracle.solve(prob, max_iters=4, inst_threads=B)
    assert list(res.num_iters) == [i.num_iters for i in ref["info"]]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL
    s.close()


# ---------------------------------------------------------------- reference-gait walking, kinodynamic and centroidal models
def _rotate(prob, knots, nxt):
    B, T = prob["x0"].shape[0], prob["cfg"].T
    out = list(knots)
    for b in range(B):
        out[b * T:(b + 1) * T] = out[b * T + 1:(b + 1) * T] + [nxt[b]]
    return out


@pytest.mark.parametrize("kind,ticks,perturb", [(_abi.KIND_KINO, [99, 60, 140], False), (_abi.KIND_CENT, [99, 150, 30], True),
                                                (_abi.KIND_FULL, [99, 70, 50], True)])
def test_reference_gait_cold_solve_and_ticks(oracle, kind, ticks, perturb):
    """BASELINE configs[0]-[2]: horizons of the reference walking loops (single-support knots, moving swing references / contact
    positions, ramping force references, terminal CoM equality; kinodynamic_talos.py:161-171,183-261, centroidal_talos.py:100-183,
    374-394): cold solve, then a warm tick with the multipliers RESET (solver.setup per tick: full:539, cent:461) and one with the
    multipliers SHIFTED (solver.cycleProblem without setup, kinodynamic_talos.py:488 -> mpc_shift_multipliers), each against the
    oracle run on the host-rotated problem."""
    B = len(ticks)
    prob = problems.walk_batch(kind, B, seed=2, ticks=ticks, mirror=[False, True, False], perturb=perturb)
    T = prob["cfg"].T
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=6)
    ref = oracle.solve(prob, max_iters=6, inst_threads=B)
    assert list(res.num_iters) == [i.num_iters for i in ref["info"]]
    assert list(res.ls_evals) == [i.ls_evals for i in ref["info"]]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL and rel(res.vs, ref["vs"]) < 1e-5
    assert rel(res.K, ref["K"]) < 1e-5
    # the stage entering each horizon: the next one of the same gait
    nxt_prob = problems.walk_batch(kind, B, seed=2, ticks=[t + 1 for t in ticks], mirror=[False, True, False], perturb=perturb)
    nxt = (_abi.Knot * B)(*[nxt_prob["knots"][b * T + T - 1] for b in range(B)])
    knots = list(prob["knots"])
    xs, us, vs, lams = ref["xs"], ref["us"], ref["vs"], ref["lams"]
    for keep in (False, True):
        s.tick(nxt, None, keep_multipliers=keep, max_iters=1)
        got = s.results(gains=False)
        knots = _rotate(prob, knots, nxt)
        hp = dict(prob, knots=(_abi.Knot * (B * T))(*knots), x0=xs[:, 1].copy())
        xs_ws = np.concatenate([xs[:, 1:], xs[:, -1:]], axis=1)
        us_ws = np.concatenate([us[:, 1:], us[:, -1:]], axis=1)
        if keep:
            # running knots / co-states of x_1..x_T one knot to the left, last entry repeated; terminal multiplier and lams[0] in place
            vs0 = np.concatenate([vs[:, 1:-1], vs[:, -2:-1], vs[:, -1:]], axis=1)
            lams0 = np.concatenate([lams[:, :1], lams[:, 2:], lams[:, -1:]], axis=1)
            assert np.abs(vs0).max() > 0 and np.abs(lams0).max() > 0
        else:
            vs0 = lams0 = None
        r = oracle.solve(hp, max_iters=1, inst_threads=B, xs=xs_ws, us=us_ws, vs=vs0, lams=lams0)
        assert list(got.ls_evals) == [i.ls_evals for i in r["info"]], keep
        assert rel(got.xs, r["xs"]) < RTOL and rel(got.us, r["us"]) < RTOL and rel(got.vs, r["vs"]) < 1e-5 and rel(got.lams, r["lams"]) < 1e-5, keep
        xs, us, vs, lams = r["xs"], r["us"], r["vs"], r["lams"]
    s.close()


def test_phase_matched_tail_warmstart_matches_oracle(oracle):
    """mpc_set_tail_warmstart(1): at the tick where the first double-support knot after a swing phase (robot 0: reference tick 110) / the first knot of
    the second swing (robot 1: tick 140) enters the horizon, the appended knot's control comes from the nearest knot with the same contact phase;
    the tick against the oracle started from exactly that warm start."""
    B, ticks = 2, [109, 139]
    prob = problems.walk_batch(_abi.KIND_FULL, B, seed=2, ticks=ticks, mirror=[False, False], perturb=True)
    T = prob["cfg"].T
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    s.run(prob["xs"], prob["us"], max_iters=4)
    ref = oracle.solve(prob, max_iters=4, inst_threads=B)
    nxt_prob = problems.walk_batch(_abi.KIND_FULL, B, seed=2, ticks=[t + 1 for t in ticks], mirror=[False, False], perturb=True)
    nxt = (_abi.Knot * B)(*[nxt_prob["knots"][b * T + T - 1] for b in range(B)])
    s.set_tail_warmstart(True)
    s.tick(nxt, None, keep_multipliers=False, max_iters=1)
    got = s.results(gains=False)
    knots = _rotate(prob, list(prob["knots"]), nxt)
    xs, us = ref["xs"], ref["us"]
    xs_ws = np.concatenate([xs[:, 1:], xs[:, -1:]], axis=1)
    us_ws = np.concatenate([us[:, 1:], us[:, -1:]], axis=1)
    replaced = []
    for b in range(B):
        ph = [(k.cs[0], k.cs[1]) for k in knots[b * T:(b + 1) * T]]
        assert ph[T - 1] != ph[T - 2]  # a contact switch has just entered the horizon
        for j in range(T - 2, -1, -1):
            if ph[j] == ph[T - 1]:
                replaced.append(j)
                us_ws[b, T - 1] = us_ws[b, j]
                break
    assert replaced and replaced[0] < T - 2  # robot 0: double-support knots at the front of the horizon
    hp = dict(prob, knots=(_abi.Knot * (B * T))(*knots), x0=xs[:, 1].copy())
    r = oracle.solve(hp, max_iters=1, inst_threads=B, xs=xs_ws, us=us_ws)
    assert list(got.ls_evals) == [i.ls_evals for i in r["info"]]
    assert rel(got.xs, r["xs"]) < RTOL and rel(got.us, r["us"]) < RTOL
    s.close()


def test_shift_multipliers_entry_point():
    """mpc_shift_multipliers (solver.cycleProblem, kinodynamic_talos.py:488): running-knot multipliers and the co-states of x_1..x_T move
    n knots to the left and repeat their last entry; the terminal multiplier vs[T] and the initial-condition co-state lams[0] stay."""
    prob = problems.walk_batch(_abi.KIND_KINO, 2, seed=4, ticks=[99, 80], mirror=[False, False], perturb=False)
    s = BatchSolver(prob["robot"], prob["cfg"], 2)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    a = s.run(prob["xs"], prob["us"], max_iters=3, gains=False)
    assert np.abs(a.vs).max() > 0 and np.abs(a.lams).max() > 0
    for n in (1, 3):
        s.shift_multipliers(n)
        b = s.results(gains=False)
        T = a.vs.shape[1] - 1
        assert np.array_equal(b.vs[:, :T - n], a.vs[:, n:T]) and np.array_equal(b.lams[:, 1:T + 1 - n], a.lams[:, 1 + n:])
        assert np.array_equal(b.vs[:, T - n:T], np.repeat(a.vs[:, T - 1:T], n, axis=1)) and np.array_equal(b.lams[:, T + 1 - n:], np.repeat(a.lams[:, T:], n, axis=1))
        assert np.array_equal(b.vs[:, T], a.vs[:, T]) and np.array_equal(b.lams[:, 0], a.lams[:, 0])
        a = b
    s.close()


def test_kino_more_active_rows_than_fast_carving(oracle):
    """59-65 active rows per knot (> the 56 of the kinodynamic Riccati's fast shared-memory carving, <= 68): solved, no status 3."""
    from test_emulation import kino_many_active_rows

    prob, xs, us = kino_many_active_rows()
    s = BatchSolver(prob["robot"], prob["cfg"], 1)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(xs, us, max_iters=2)
    ref = oracle.solve(prob, max_iters=2, xs=xs, us=us)
    assert res.info[0].status != 3 and list(res.num_iters) == [2]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL and rel(res.vs, ref["vs"]) < 1e-5
    s.close()


def test_batch_4096_sampled_instances(oracle):
    """BASELINE configs[4] at its full size (batch 4096, T = 100, 97 GB of workspace): eight sampled instances equal the oracle, every
    instance is finite, and the result of an instance does not depend on the batch it is solved in (independent units)."""
    B, T = 4096, 100
    prob = problems.walk_batch(_abi.KIND_FULL, B, seed=5, ticks=np.random.default_rng(5).integers(0, 100, size=B))
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=2, gains=False)
    s.close()
    assert np.isfinite(res.xs).all() and np.isfinite(res.us).all() and (res.num_iters == 2).all()
    pick = [0, 511, 1024, 2047, 2048, 3000, 3333, 4095]
    sub = dict(prob)
    sub["knots"] = (_abi.Knot * (len(pick) * T))(*[prob["knots"][i * T + k] for i in pick for k in range(T)])
    sub["terms"] = (_abi.Term * len(pick))(*[prob["terms"][i] for i in pick])
    sub["x0"], sub["xs"], sub["us"] = prob["x0"][pick], prob["xs"][pick], prob["us"][pick]
    ref = oracle.solve(sub, max_iters=2, inst_threads=8)
    assert rel(res.xs[pick], ref["xs"]) < RTOL and rel(res.us[pick], ref["us"]) < RTOL
    assert list(res.ls_evals[pick]) == [i.ls_evals for i in ref["info"]]
    s2 = BatchSolver(sub["robot"], sub["cfg"], len(pick))
    s2.setup(sub["knots"], sub["terms"], sub["x0"])
    r2 = s2.run(sub["xs"], sub["us"], max_iters=2, gains=False)
    s2.close()
    # launches that fill the GPU run the Riccati kernel with 128 threads per instance (two instances per SM), smaller ones with 256:
    # the merit's directional derivative is summed per thread, so the two agree to rounding, not bit for bit
    assert rel(r2.xs, res.xs[pick]) < 1e-11 and rel(r2.us, res.us[pick]) < 1e-11 and list(r2.ls_evals) == list(res.ls_evals[pick])
    # ... and inside one launch shape the result of an instance is bit-identical whatever batch it is solved in
    half = dict(sub)
    half["knots"] = (_abi.Knot * (4 * T))(*[sub["knots"][i * T + k] for i in range(4) for k in range(T)])
    half["terms"] = (_abi.Term * 4)(*[sub["terms"][i] for i in range(4)])
    half["x0"], half["xs"], half["us"] = sub["x0"][:4], sub["xs"][:4], sub["us"][:4]
    s3 = BatchSolver(half["robot"], half["cfg"], 4)
    s3.setup(half["knots"], half["terms"], half["x0"])
    r3 = s3.run(half["xs"], half["us"], max_iters=2, gains=False)
    s3.close()
    assert np.array_equal(r3.xs, r2.xs[:4]) and np.array_equal(r3.us, r2.us[:4])


def test_stairs_batch_sample(oracle):
    """BASELINE configs[3]: stair climbing (x_forward 0.3, z_height +0.10 per step, talos_utils.py:187-192), perturbed initial states,
    default_rng(4); batch 512 with four sampled instances against the oracle."""
    B, T = 512, 100
    prob = problems.full_stairs_batch(B, ticks=np.random.default_rng(4).integers(0, 100, size=B))
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=3, gains=False)
    s.close()
    assert np.isfinite(res.xs).all() and (res.num_iters == 3).all()
    pick = [3, 200, 333, 511]
    sub = problems.sub_problem(prob, 0, 1)
    sub["knots"] = (_abi.Knot * (len(pick) * T))(*[prob["knots"][i * T + k] for i in pick for k in range(T)])
    sub["terms"] = (_abi.Term * len(pick))(*[prob["terms"][i] for i in pick])
    sub["x0"], sub["xs"], sub["us"] = prob["x0"][pick], prob["xs"][pick], prob["us"][pick]
    ref = oracle.solve(sub, max_iters=3, inst_threads=4)
    assert rel(res.xs[pick], ref["xs"]) < RTOL and rel(res.us[pick], ref["us"]) < RTOL


@pytest.mark.parametrize("maker,T", [(problems.cent_standing_problem, 100), (problems.full_standing_problem, 100), (problems.kino_standing_problem, 40)])
def test_nonlinear_rollout_kernel_cold_solve(oracle, maker, T):
    """ROLLOUT_NONLINEAR: the fused rollout + linesearch kernel (k_rollout_ls, one CTA per instance: nonlinear dynamics knot after
    knot under the affine LQ policy, Armijo backtracking inside the kernel) against the oracle's try_step_nonlinear."""
    prob = maker(batch=2, T=T)
    prob["cfg"].rollout = 1
    s = BatchSolver(prob["robot"], prob["cfg"], 2)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=100)
    ref = oracle.solve(prob, max_iters=100, inst_threads=2)
    assert list(res.num_iters) == [i.num_iters for i in ref["info"]] and list(res.conv) == [bool(i.conv) for i in ref["info"]]
    assert [i.ls_evals for i in res.info] == [i.ls_evals for i in ref["info"]]
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL and rel(res.K, ref["K"]) < 1e-5
    s.close()


def test_nonlinear_rollout_kernel_walking_backtracking(oracle):
    """ROLLOUT_NONLINEAR on the walking fixture: active cone / box rows feed back through dx, the in-kernel linesearch backtracks."""
    import golden_util

    prob, z = golden_util.load("walk_full.npz")
    prob["cfg"].rollout = 1
    B = prob["x0"].shape[0]
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    res = s.run(prob["xs"], prob["us"], max_iters=4)
    ref = oracle.solve(prob, max_iters=4, inst_threads=4)
    assert [i.ls_evals for i in res.info] == [i.ls_evals for i in ref["info"]] and max(i.ls_evals for i in ref["info"]) > 4
    assert rel(res.xs, ref["xs"]) < RTOL and rel(res.us, ref["us"]) < RTOL and rel(res.vs, ref["vs"]) < 1e-5
    assert rel(res.xs, z["sol_xs"]) > 1e-9  # not the linear rollout's iterates
    s.close()


def test_nonlinear_rollout_through_the_shim(oracle):
    """solver.rollout_type = ROLLOUT_NONLINEAR (aligator's default) through the public surface."""
    import mpc_benchmark_b200 as aligator
    from mpc_benchmark_b200 import flatten, pin
    from test_reference_scripts import _build_cent

    ns = _build_cent(aligator, pin, T=20)
    solver = aligator.SolverProxDDP(1e-5, 1e-8)
    assert solver.rollout_type == aligator.ROLLOUT_NONLINEAR
    solver.max_iters = 100
    solver.setup(ns["problem"])
    conv = solver.run(ns["problem"], [ns["x0"]] * 21, [ns["u0"] for _ in range(20)])
    flat = flatten.flatten_problem(ns["problem"], 1e-5, 1e-8, 100, rollout=1)
    ref = oracle.solve(dict(robot=flat.robot, cfg=flat.cfg, knots=flat.knots, terms=flat.terms, x0=flat.x0,
                            xs=np.tile(ns["x0"], (1, 21, 1)), us=np.tile(ns["u0"], (1, 20, 1))))
    assert conv and solver.results.num_iters == ref["info"][0].num_iters
    assert rel(np.array(solver.results.xs.tolist()), ref["xs"][0]) < RTOL and rel(np.array(solver.results.us.tolist()), ref["us"][0]) < RTOL


@pytest.mark.parametrize("batch,parts", [(6, 2), (5, 3), (600, 2)])
def test_pipelined_host_call_equals_plain_run(batch, parts):
    """mpc_run_pipelined (sub-batches, uploads / downloads on a second stream beside the solve) returns what mpc_run + mpc_get_results
    return, bit for bit, for every instance — also when the parts are ragged and when they cross the Riccati kernel's launch shapes."""
    prob = problems.full_walk_batch(batch, seed=9, T=100 if batch < 100 else 30)
    s = BatchSolver(prob["robot"], prob["cfg"], batch)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    ref = s.run(prob["xs"], prob["us"], max_iters=2)
    s.reset_multipliers()
    T = prob["cfg"].T
    xs_o, us_o, k0_o = np.empty((batch, T + 1, 57)), np.empty((batch, T, 22)), np.empty((batch, 22, 56))
    info = s.run_pipelined(prob["xs"], prob["us"], xs_o, us_o, k0_o, max_iters=2, parts=parts, info=True)
    assert np.array_equal(xs_o, ref.xs) and np.array_equal(us_o, ref.us) and np.array_equal(k0_o, ref.K[:, 0])
    assert [i.num_iters for i in info] == list(ref.num_iters) and [i.ls_evals for i in info] == [i.ls_evals for i in ref.info]
    s.close()
