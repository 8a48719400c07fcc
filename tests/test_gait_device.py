"""Device-side gait / swing-foot reference generation (csrc/gait.cuh, mpc_gait_setup / mpc_gait_tick) against its
host mirror mpc_benchmark_b200/gait.py (GaitPlan: the restatement of talos_utils.update_timings / footTrajectory and of the scripts'
per-tick setReference / contact_poses / replaceStageCircular bookkeeping) — tick for tick, knot for knot, for all three models, both
mirror flags and measured sole placements that drift away from the references.  CPU: the kernel source under host emulation; GPU: the kernel."""
import ctypes as C

import numpy as np
import pytest

from mpc_benchmark_b200 import _abi, gait, problems

FIELDS = ("cs", "fcost", "w_lf", "w_rf", "lf_ref", "rf_ref", "f_ref", "u_ref", "cpos")


def _host_reference(kind, T, lf0, rf0, com0, mass, mirror, lf_meas, rf_meas, stairs=False):
    """What the host loop hands to the solver at every tick (the closed-loop tools do exactly this)."""
    ticks, B = lf_meas.shape[:2]
    kw = dict(x_forward=0.3, z_height=0.10, keep_forward=True) if stairs else {}
    plans = [gait.GaitPlan(kind, lf0, rf0, com0, nsteps=T, mirror=bool(mirror[b]), **kw) for b in range(B)]
    urefs = gait.force_ramp_refs(kind, mass, 34 if kind == _abi.KIND_KINO else 12, T) if kind != _abi.KIND_FULL else None
    fr = np.array([0, 0, mass * problems.GRAVITY / 2.0, 0, 0, 0.0])
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    out = []
    for t in range(ticks):
        ks, ts = [], []
        for b in range(B):
            p = plans[b]
            LF, RF, _, com_final = p.tick(lf_meas[t, b], rf_meas[t, b])
            for j in range(T):
                cs = p.h_phase[j]
                if kind == _abi.KIND_FULL:
                    ks.append(problems.full_knot(cs, p.h_lf[j], p.h_rf[j], fr, fr))
                else:
                    u = urefs[min(p.h_index[j], len(urefs) - 1)]
                    u = problems._swap_feet_u(u, kind) if mirror[b] else u
                    if kind == _abi.KIND_KINO:
                        ks.append(problems.kino_knot(cs, p.h_lf[j], p.h_rf[j], u))
                    else:
                        ks.append(problems.cent_knot(cs, p.h_lf[j] if cs[0] else lf0, p.h_rf[j] if cs[1] else rf0, u))
            if kind == _abi.KIND_FULL:
                ts.append(problems.make_term(LF[-1], RF[-1], com_final))
            elif kind == _abi.KIND_KINO:
                ts.append(problems.make_term(ident, ident, com_final))
            else:
                ts.append(problems.make_term(lf0, rf0))
        out.append((ks, ts))
    return out, urefs


def _measured_feet(rng, ticks, B, lf0, rf0, drift):
    """Sole placements that wander around the nominal ones (position noise + a slow yaw), as a plant would report them."""
    lf, rf = np.tile(np.asarray(lf0, float), (ticks, B, 1)), np.tile(np.asarray(rf0, float), (ticks, B, 1))
    for arr in (lf, rf):
        for b in range(B):
            yaw = np.cumsum(rng.normal(0, drift * 0.05, ticks))
            pos = np.cumsum(rng.normal(0, drift * 0.002, (ticks, 3)), axis=0)
            for t in range(ticks):
                R = gait.yaw_rotation(yaw[t]) @ arr[t, b, :9].reshape(3, 3)
                arr[t, b, :9] = R.reshape(9)
                arr[t, b, 9:] += pos[t]
    return lf, rf


def _compare(knots, terms, ref, T, B, tol=1e-12):
    for t, (ks, ts) in enumerate(ref):
        for i in range(B * T):
            a, b = knots[t][i], ks[i]
            for f in FIELDS:
                va, vb = np.array(getattr(a, f)), np.array(getattr(b, f))
                assert np.abs(va - vb).max() <= tol * max(1.0, np.abs(vb).max()), (t, i // T, i % T, f, va, vb)
        for i in range(B):
            for f in ("lf_ref", "rf_ref", "com_ref"):
                va, vb = np.array(getattr(terms[t][i], f)), np.array(getattr(ts[i], f))
                assert np.abs(va - vb).max() <= tol * max(1.0, np.abs(vb).max()), (t, i, f, va, vb)
            assert terms[t][i].has_com_cstr == ts[i].has_com_cstr


CASES = [(_abi.KIND_FULL, 400, False), (_abi.KIND_KINO, 330, False), (_abi.KIND_CENT, 330, False), (_abi.KIND_FULL, 260, True)]


@pytest.mark.parametrize("kind,ticks,stairs", CASES)
def test_emulated_gait_kernel_matches_host_plan(kind, ticks, stairs):
    import emu_lib

    T, B = 100, 2
    rb, q0, x0, lf0, rf0, com0, mass = problems.base_setup(None)
    rng = np.random.default_rng(3)
    lf, rf = _measured_feet(rng, ticks, B, lf0, rf0, drift=1.0)
    mirror = [False, True]
    ref, urefs = _host_reference(kind, T, lf0, rf0, com0, mass, mirror, lf, rf, stairs)
    kw = dict(x_forward=0.3, z_height=0.10, keep_forward=True) if stairs else {}
    g = gait.device_gait(kind, lf0, rf0, com0, mass, **kw)
    ks, ts = emu_lib.gait(kind, T, g, mirror, urefs, lf, rf)
    knots = [[ks[(t * B) * T + i] for i in range(B * T)] for t in range(ticks)]
    terms = [[ts[t * B + i] for i in range(B)] for t in range(ticks)]
    # every 7th tick plus the ticks around every contact switch at the front of the horizon
    sel = sorted(set(range(0, ticks, 7)) | {t for t in range(ticks) if any(abs(t - s) <= 1 for s in (20, 30, 100, 110, 120, 140, 200, 220, 230, 300, 320))})
    _compare([knots[t] for t in sel], [terms[t] for t in sel], [ref[t] for t in sel], T, B)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ticks", [(_abi.KIND_FULL, 150), (_abi.KIND_KINO, 130), (_abi.KIND_CENT, 130)])
def test_gpu_gait_kernel_matches_host_plan(kind, ticks):
    from mpc_benchmark_b200.batch import BatchSolver

    T, B = 100, 3
    rb, q0, x0, lf0, rf0, com0, mass = problems.base_setup(None)
    rng = np.random.default_rng(4)
    lf, rf = _measured_feet(rng, ticks, B, lf0, rf0, drift=1.0)
    mirror = [False, True, False]
    ref, urefs = _host_reference(kind, T, lf0, rf0, com0, mass, mirror, lf, rf)
    maker = {_abi.KIND_FULL: problems.full_standing_problem, _abi.KIND_KINO: problems.kino_standing_problem, _abi.KIND_CENT: problems.cent_standing_problem}[kind]
    prob = maker(batch=B, T=T)
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    s.gait_setup(gait.device_gait(kind, lf0, rf0, com0, mass), mirror, urefs)
    sel = set(range(0, ticks, 11)) | {19, 20, 21, 29, 30, 31, 99, 100, 101, 109, 110, 111, 119, 120, 121}
    for t in range(ticks):
        s.gait_tick(lf[t], rf[t])
        if t in sel:
            ks, ts = s.knots()
            _compare([list(ks)], [list(ts)], [ref[t]], T, B)
    s.close()


@pytest.mark.gpu
def test_gpu_closed_loop_with_device_gait_matches_host_driven_loop():
    """Five closed-loop ticks of the full-dynamics MPC with EVERYTHING per tick on the device (sole placements of the predicted state by
    forward kinematics, gait bookkeeping, warm-start shift, one ProxDDP iteration) against the same loop driven from the host
    (kinematics.foot_placements + GaitPlan + mpc_update_knots / mpc_update_terms)."""
    from mpc_benchmark_b200.batch import BatchSolver
    from mpc_benchmark_b200.kinematics import foot_placements

    T, B = 100, 2
    rb, q0, x0, lf0, rf0, com0, mass = problems.base_setup(None)
    prob = problems.full_standing_problem(batch=B, T=T)
    mirror = [False, True]
    fr = np.array([0, 0, mass * problems.GRAVITY / 2.0, 0, 0, 0.0])
    sols = []
    for device_side in (True, False):
        s = BatchSolver(prob["robot"], prob["cfg"], B)
        s.setup(prob["knots"], prob["terms"], prob["x0"])
        s.run(prob["xs"], prob["us"], max_iters=8)
        if device_side:
            s.gait_setup(gait.device_gait(_abi.KIND_FULL, lf0, rf0, com0, mass), mirror)
        else:
            plans = [gait.GaitPlan(_abi.KIND_FULL, lf0, rf0, com0, nsteps=T, mirror=m) for m in mirror]
        for t in range(5):
            if device_side:
                s.gait_tick()
            else:
                xs = s.results(gains=False, multipliers=False).xs
                knots, terms = (_abi.Knot * (B * T))(), (_abi.Term * B)()
                for b in range(B):
                    lf, rf = foot_placements(rb, xs[b, 1, :29])
                    LF, RF, _, com_final = plans[b].tick(lf, rf)
                    p = plans[b]
                    knots[b * T:(b + 1) * T] = [problems.full_knot(p.h_phase[j], p.h_lf[j], p.h_rf[j], fr, fr) for j in range(T)]
                    terms[b] = problems.make_term(LF[-1], RF[-1], com_final)
                s.update_knots(knots, 0, T)
                s.update_terms(terms)
            s.tick(None, None, keep_multipliers=False, max_iters=1)
        sols.append(s.results(gains=False, multipliers=False))
        s.close()
    assert np.abs(sols[0].xs - sols[1].xs).max() < 1e-9 and np.abs(sols[0].us - sols[1].us).max() < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ticks,phase_tail", [(_abi.KIND_FULL, 260, True), (_abi.KIND_KINO, 215, False), (_abi.KIND_CENT, 230, False)])
def test_gpu_reference_gait_closed_loop_walks(kind, ticks, phase_tail):
    """Closed loop along the reference gait with every tick on the device (mpc_gait_tick + mpc_tick), the reference's solver settings (mu_init = 1e-8,
    one ProxDDP iteration per tick, multipliers reset), 32 robots (half mirrored, perturbed initial states), ideal plant: through the first landing and
    into / through the second swing phase nobody fails, (nearly) full steps are taken and the base / CoM stays at its height.  Full dynamics: with the
    appended knot started from a knot of the same contact phase (mpc_set_tail_warmstart(1)); with the scripts' copy this loop has diverged by tick 260
    (DESIGN section 7).  Whole gaits: profiles/r2_gait_walk_{cent,kino,full}_gpu.txt."""
    from mpc_benchmark_b200.batch import BatchSolver

    B = 32
    maker = {_abi.KIND_FULL: problems.full_standing_problem, _abi.KIND_KINO: problems.kino_standing_problem, _abi.KIND_CENT: problems.cent_standing_problem}[kind]
    prob = maker(batch=B, mu_init=1e-8)
    rng = np.random.default_rng(1)
    if kind == _abi.KIND_CENT:
        prob["x0"] = prob["x0"] + rng.normal(size=prob["x0"].shape) * np.array([0.003] * 3 + [0.02 * prob["mass"]] * 3 + [0.02] * 3)
    else:
        x0 = problems.perturbed_x0(prob["robot"], prob["x0"][0], rng, B)
        prob["x0"] = prob["x0"] + 0.3 * (x0 - prob["x0"])
        prob["x0"][:, 3:7] /= np.linalg.norm(prob["x0"][:, 3:7], axis=1, keepdims=True)
    s = BatchSolver(prob["robot"], prob["cfg"], B)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    s.run(prob["xs"], prob["us"], max_iters=100, gains=False)
    urefs = gait.force_ramp_refs(kind, prob["mass"], 34 if kind == _abi.KIND_KINO else 12, prob["cfg"].T) if kind != _abi.KIND_FULL else None
    s.set_tail_warmstart(phase_tail)
    s.gait_setup(gait.device_gait(kind, prob["lf"], prob["rf"], prob["com0"], prob["mass"]), (np.arange(B) % 2).astype(bool), urefs)
    z0 = prob["x0"][:, 2].mean()
    small_steps = 0
    for t in range(ticks):
        s.gait_tick()
        s.tick(None, None, keep_multipliers=False, max_iters=1)
        if t % 10 == 9 or t == ticks - 1:
            r = s.results(gains=False, multipliers=False)
            st = np.array([i.status for i in r.info])
            assert (st < 2).all() and np.isfinite(r.xs).all(), t
            assert np.abs(r.xs[:, 0, 2] - z0).max() < 0.05, t  # base / CoM height
            small_steps += int(np.median(r.alpha) < 0.99)
    assert small_steps <= 6 and np.median(r.alpha) == 1.0
    ks, _ = s.knots()
    ph = gait.contact_phases(kind, prob["cfg"].T)  # robot 0 is not mirrored: knot 0 carries phase (ticks - 1) - (T - 1) of the schedule, past the first swing
    idx = ticks - prob["cfg"].T
    assert [bool(ks[0].cs[0]), bool(ks[0].cs[1])] == ph[idx] and any(p != [True, True] for p in ph[:idx])
    s.close()
