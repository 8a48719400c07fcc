"""Closed-loop CENTROIDAL walking MPC on the CPU ORACLE (BASELINE configs[0]: centroidal_talos.py flat-ground walk), ideal plant.

The loop of centroidal_talos.py:354-466 without PyBullet and without the whole-body tracking QP: every tick the foot references are
regenerated (talos_utils.footTrajectory restated in gait.py), the contact positions of the ACTIVE contacts follow them (cent:374-384),
the horizon rotates by one stage (cent:459-460), the previous solution shifted by one knot is the warm start, the multipliers are reset
(`solver.setup`, cent:461) and ONE ProxDDP iteration runs (cent:298).  Ideal plant: the centroidal state follows the model (x_meas =
f(x0, us[0])) and the feet track their references exactly (the pose handed to the trajectory generator is last tick's reference).
The model carries nothing of the synthetic Talos tree but its mass and nominal foot placements, which makes this the cleanest closed-loop
check of the ProxDDP restatement at the reference's settings.  TEST tooling around oracle/.

usage: python tools/closed_loop_oracle_cent.py [robots] [ticks] [iters_per_tick] [mu_init]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from mpc_benchmark_b200 import _abi, gait, problems  # noqa: E402


def run(B=2, N=450, iters=1, mu_init=1e-8, verbose=True, threads=8):
    prob = problems.cent_standing_problem(batch=B, mu_init=mu_init)
    rb, cfg, T = prob["robot"], prob["cfg"], prob["cfg"].T
    _, _, _, lf0, rf0, com0, mass = problems.base_setup(None)
    rng = np.random.default_rng(1)
    prob["x0"] = prob["x0"] + rng.normal(size=prob["x0"].shape) * np.array([0.003] * 3 + [0.02 * mass] * 3 + [0.02] * 3)
    cold = oracle_lib.solve(prob, max_iters=100, inst_threads=threads)
    if verbose:
        print("cold: iters", [i.num_iters for i in cold["info"]], "conv", [i.conv for i in cold["info"]])
    plans = [gait.GaitPlan(_abi.KIND_CENT, lf0, rf0, com0, nsteps=T, mirror=bool(b % 2)) for b in range(B)]
    urefs = gait.force_ramp_refs(_abi.KIND_CENT, mass, 12, T)
    feet = [(np.array(lf0, float), np.array(rf0, float)) for _ in range(B)]
    xs, us = cold["xs"], cold["us"]
    cur_knots = list(prob["knots"])
    hist = []
    for t in range(N):
        x_meas = np.stack([oracle_lib.eval_knot(rb, cfg, cur_knots[b * T], xs[b, 0], us[b, 0], xs[b, 1], derivs=False)["xnext"] for b in range(B)])
        knots = (_abi.Knot * (B * T))()
        terms = (_abi.Term * B)()
        for b in range(B):
            p = plans[b]
            LF, RF, _, _ = p.tick(feet[b][0], feet[b][1])
            feet[b] = (np.array(LF[1], float), np.array(RF[1], float))  # perfect tracking: next tick's measured pose = this tick's next reference
            ks = []
            for j in range(T):
                cs = p.h_phase[j]
                u = urefs[min(p.h_index[j], len(urefs) - 1)]
                if p.ft is not None and getattr(plans[b], "phases", None) is not None and b % 2:
                    u = problems._swap_feet_u(u, _abi.KIND_CENT)
                ks.append(problems.cent_knot(cs, p.h_lf[j] if cs[0] else lf0, p.h_rf[j] if cs[1] else rf0, u))
            knots[b * T:(b + 1) * T] = ks
            terms[b] = problems.make_term(lf0, rf0)
        hp = dict(prob, knots=knots, terms=terms, x0=x_meas)
        cur_knots = knots
        xs_ws = np.concatenate([xs[:, 1:], xs[:, -1:]], axis=1)
        us_ws = np.concatenate([us[:, 1:], us[:, -1:]], axis=1)
        r = oracle_lib.solve(hp, max_iters=iters, inst_threads=threads, xs=xs_ws, us=us_ws)
        xs, us = r["xs"], r["us"]
        prim = np.array([i.prim_infeas for i in r["info"]])
        dual = np.array([i.dual_infeas for i in r["info"]])
        alpha = np.array([i.alpha for i in r["info"]])
        ls = np.array([i.ls_evals for i in r["info"]])
        hist.append((t + 1, float(prim.max()), float(np.median(dual)), float(alpha.min()), float(ls.mean()), float(xs[:, 0, 2].min()), float(xs[:, 0, 2].max()),
                     float(np.abs(xs[:, 0, :2] - 0.5 * (feet[0][0][9:11] + feet[0][1][9:11])).max())))
        if verbose and (t % 20 == 19 or t < 3):
            h = hist[-1]
            print(f"tick {h[0]:4d}: prim max {h[1]:9.2e} | dual med {h[2]:9.2e} | alpha min {h[3]:.2e} | ls {h[4]:.2f} | com z {h[5]:.3f}..{h[6]:.3f} | "
                  f"phase0 {plans[0].h_phase[0]} | f_z L/R {us[0, 0, 2]:7.1f} {us[0, 0, 8]:7.1f}", flush=True)
        if not np.isfinite(xs).all() or xs[:, 0, 2].min() < 0.4 or xs[:, 0, 2].max() > 1.6:
            if verbose:
                print("DIVERGED at tick", t + 1)
            break
    return hist


if __name__ == "__main__":
    a = sys.argv[1:]
    run(B=int(a[0]) if len(a) > 0 else 2, N=int(a[1]) if len(a) > 1 else 450, iters=int(a[2]) if len(a) > 2 else 1, mu_init=float(a[3]) if len(a) > 3 else 1e-8)
