"""Closed-loop walking MPC on the CPU ORACLE with the reference's gait bookkeeping (gait.GaitPlan) and an ideal plant.

Reproduces the loop of fulldynamic_talos.py:438-550 without PyBullet: every tick the swing references are regenerated from the
measured foot placements, the horizon rotates by one stage, the previous solution shifted by one knot is the warm start, x0 is the
state the model predicted for the end of the previous tick, the multipliers are reset (`solver.setup`, full:539) and ONE ProxDDP
iteration runs (full:407).  The oracle's ablation switches (oracle/proxddp.hpp `SolverParams`) are set through the environment:
ORC_MU_DYN_SCALE, ORC_LS_MODE (0 Armijo, 1 non-monotone, 2 full steps), ORC_LS_ALPHA_MIN, ORC_DUAL_WEIGHT, ORC_REG_INIT.

usage: python tools/closed_loop_oracle.py [robots] [ticks] [iters_per_tick] [keep_multipliers] [mu_init] [plant: model|prediction] [full|kino]
This is TEST tooling around oracle/ (DESIGN "oracle-vs-Aligator ablation"); the product never runs it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from mpc_benchmark_b200 import _abi, gait, problems  # noqa: E402
from mpc_benchmark_b200.kinematics import foot_placements  # noqa: E402


def run(B=4, N=300, iters=1, keep=False, mu_init=1e-8, sigma=0.3, verbose=True, threads=8, plant="model", kind=_abi.KIND_FULL, want_state=False):
    """plant = "model": x_meas = f(x0, us[0]), the model's own integrator applied to the first control (an ideal plant that obeys the
    dynamics); "prediction": x_meas = xs[1] (what mpc_tick does without a measured state — it inherits the shooting gap of the plan)."""
    kino = kind == _abi.KIND_KINO  # kinodynamic_talos.py:362-500: same loop, contact wrenches in u, force references ramping through the double supports
    prob = (problems.kino_standing_problem if kino else problems.full_standing_problem)(batch=B, mu_init=mu_init)
    rb, cfg, T = prob["robot"], prob["cfg"], prob["cfg"].T
    lf0, rf0, com0, mass = prob["lf"], prob["rf"], prob["com0"], prob["mass"]
    rng = np.random.default_rng(1)
    x0 = problems.perturbed_x0(rb, prob["x0"][0], rng, B)
    x0 = prob["x0"] + sigma * (x0 - prob["x0"])  # small disturbance of the initial state (quaternion renormalised by the solver)
    x0[:, 3:7] /= np.linalg.norm(x0[:, 3:7], axis=1, keepdims=True)
    prob["x0"] = x0
    cold = oracle_lib.solve(prob, max_iters=100, inst_threads=threads)
    if verbose:
        print("cold: iters", [i.num_iters for i in cold["info"]], "conv", [i.conv for i in cold["info"]])
    plans = [gait.GaitPlan(kind, lf0, rf0, com0, nsteps=T) for _ in range(B)]
    urefs = gait.force_ramp_refs(kind, mass, 34, T) if kino else None
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    fr = np.array([0, 0, mass * problems.GRAVITY / 2.0, 0, 0, 0.0])
    xs, us, vs, lams = cold["xs"], cold["us"], cold["vs"], cold["lams"]
    hist = []
    cur_knots = list(prob["knots"])
    for t in range(N):
        if plant == "model":
            x_meas = np.stack([oracle_lib.eval_knot(rb, cfg, cur_knots[b * T], xs[b, 0], us[b, 0], xs[b, 1], derivs=False)["xnext"] for b in range(B)])
        else:
            x_meas = xs[:, 1].copy()
        knots = (_abi.Knot * (B * T))()
        terms = (_abi.Term * B)()
        for b in range(B):
            lf, rf = foot_placements(rb, x_meas[b, :29])
            LF, RF, _, com_final = plans[b].tick(lf, rf)
            p = plans[b]
            if kino:
                knots[b * T:(b + 1) * T] = [problems.kino_knot(p.h_phase[j], p.h_lf[j], p.h_rf[j], urefs[min(p.h_index[j], len(urefs) - 1)]) for j in range(T)]
                terms[b] = problems.make_term(ident, ident, com_final)
            else:
                knots[b * T:(b + 1) * T] = [problems.full_knot(p.h_phase[j], p.h_lf[j], p.h_rf[j], fr, fr) for j in range(T)]
                terms[b] = problems.make_term(LF[-1], RF[-1], com_final)
        hp = dict(prob, knots=knots, terms=terms, x0=x_meas)
        cur_knots = knots
        xs_ws = np.concatenate([xs[:, 1:], xs[:, -1:]], axis=1)
        us_ws = np.concatenate([us[:, 1:], us[:, -1:]], axis=1)
        if os.environ.get("ORC_TAIL_U") == "phase":  # experiment (NOT the reference's warm start): the appended knot starts from the control of the
            for b in range(B):                       # nearest knot of the horizon with the same contact phase instead of the previous knot's
                ph = plans[b].h_phase
                for j in range(T - 2, -1, -1):
                    if ph[j] == ph[T - 1]:
                        us_ws[b, T - 1] = us_ws[b, j]
                        break
        if keep:
            mode = int(os.environ.get("ORC_SHIFT_MODE", "1"))
            if mode == 0:  # round 1's mpc_shift_multipliers: all T + 1 slots one knot to the left, zeros appended (diverges within 40 ticks)
                vs0 = np.concatenate([vs[:, 1:], np.zeros_like(vs[:, :1])], axis=1)
                lams0 = np.concatenate([lams[:, 1:], np.zeros_like(lams[:, :1])], axis=1)
            elif mode == 1:  # mpc_shift_multipliers: terminal multiplier and initial-condition co-state stay in place, the last running knot is repeated
                vs0 = np.concatenate([vs[:, 1:-1], vs[:, -2:-1], vs[:, -1:]], axis=1)
                lams0 = np.concatenate([lams[:, :1], lams[:, 2:], lams[:, -1:]], axis=1)
            else:  # circular rotation of the running knots (the dropped knot's multipliers go to the appended one), head / tail in place
                vs0 = np.concatenate([vs[:, 1:-1], vs[:, :1], vs[:, -1:]], axis=1)
                lams0 = np.concatenate([lams[:, :1], lams[:, 2:], lams[:, 1:2]], axis=1)
        else:
            vs0 = lams0 = None
        r = oracle_lib.solve(hp, max_iters=iters, inst_threads=threads, xs=xs_ws, us=us_ws, vs=vs0, lams=lams0)
        xs, us, vs, lams = r["xs"], r["us"], r["vs"], r["lams"]
        prim = np.array([i.prim_infeas for i in r["info"]])
        dual = np.array([i.dual_infeas for i in r["info"]])
        alpha = np.array([i.alpha for i in r["info"]])
        ls = np.array([i.ls_evals for i in r["info"]])
        st = np.array([i.status for i in r["info"]])
        bad = int((st >= 2).sum() + (~np.isfinite(xs).all(axis=(1, 2))).sum())
        hist.append((t + 1, float(np.median(prim)), float(prim.max()), float(np.median(dual)), float(np.median(alpha)), float(alpha.min()),
                     float(ls.mean()), bad, float(xs[:, 0, 2].min()), float(xs[:, 0, 2].max())))
        if verbose and (t % 10 == 9 or t < 3):
            h = hist[-1]
            print(f"tick {h[0]:4d}: prim med {h[1]:9.2e} max {h[2]:9.2e} | dual med {h[3]:9.2e} | alpha med {h[4]:.3f} min {h[5]:.2e} | ls {h[6]:.2f} | "
                  f"failed {h[7]} | base z {h[8]:.3f}..{h[9]:.3f} | phase0 {plans[0].h_phase[0]}", flush=True)
        if not np.isfinite(xs).all() or xs[:, 0, 2].min() < 0.5:
            if verbose:
                print("DIVERGED at tick", t + 1)
            break
    if want_state:
        return hist, dict(prob=hp, xs=xs, us=us, vs=vs, lams=lams, plans=plans)
    return hist


if __name__ == "__main__":
    a = sys.argv[1:]
    run(B=int(a[0]) if len(a) > 0 else 4, N=int(a[1]) if len(a) > 1 else 300, iters=int(a[2]) if len(a) > 2 else 1,
        keep=bool(int(a[3])) if len(a) > 3 else False, mu_init=float(a[4]) if len(a) > 4 else 1e-8, plant=a[5] if len(a) > 5 else "model",
        kind=_abi.KIND_KINO if (len(a) > 6 and a[6] == "kino") else _abi.KIND_FULL)
