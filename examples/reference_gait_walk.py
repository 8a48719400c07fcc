#!/usr/bin/env python
"""Batched closed-loop walking MPC along the REFERENCE gait, everything per tick on the GPU: sole placements of the predicted state,
gait bookkeeping and all per-knot references (mpc_gait_tick), warm-start shift and one ProxDDP iteration (mpc_tick,
row f-2), ideal plant (x_meas = the model prediction), the reference's solver settings (mu_init = 1e-8, one iteration per tick).

    python examples/reference_gait_walk.py [kino|full|cent] [robots] [ticks] [y_gap]

kino (BASELINE configs[1], default): the whole 840-tick gait — three walking cycles — is walked; robots are perturbed copies, half of them
mirrored.  full (configs[2]): the whole 1000-tick gait with the appended knot's control taken from the nearest knot of the same contact phase
(mpc_set_tail_warmstart(1), the default here; WALK_TAIL=copy selects the scripts' us[1:] + [us[-1]], with which the loop degrades from tick 110 and
diverges around tick 250)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import _abi, gait, problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "kino"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 840
KEEP = os.environ.get("WALK_KEEP", "0") == "1"  # 1: multipliers kept and shifted (kinodynamic_talos.py:488 cycles the problem without solver.setup)
Y_GAP = float(sys.argv[4]) if len(sys.argv) > 4 else 0.18  # lateral foot spacing of the planned steps (full:355, kino:260)
kind = {"kino": _abi.KIND_KINO, "full": _abi.KIND_FULL, "cent": _abi.KIND_CENT}[model]
prob = {_abi.KIND_KINO: problems.kino_standing_problem, _abi.KIND_FULL: problems.full_standing_problem, _abi.KIND_CENT: problems.cent_standing_problem}[kind](batch=B, mu_init=1e-8)
# experiment switches (environment): Baumgarte gains of the rigid contacts and the foot-placement weight of the full-dynamics stages
if kind == _abi.KIND_FULL:
    for i in range(6):
        prob["cfg"].kd[i] *= float(os.environ.get("WALK_KD_SCALE", "1"))
        prob["cfg"].kp[i] *= float(os.environ.get("WALK_KP_SCALE", "1"))
rng = np.random.default_rng(1)
if kind == _abi.KIND_CENT:
    prob["x0"] = prob["x0"] + rng.normal(size=prob["x0"].shape) * np.array([0.003] * 3 + [0.02 * prob["mass"]] * 3 + [0.02] * 3)
else:
    x0 = problems.perturbed_x0(prob["robot"], prob["x0"][0], rng, B)
    prob["x0"] = prob["x0"] + 0.3 * (x0 - prob["x0"])
    prob["x0"][:, 3:7] /= np.linalg.norm(prob["x0"][:, 3:7], axis=1, keepdims=True)
mirror = (np.arange(B) % 2).astype(bool)
s = BatchSolver(prob["robot"], prob["cfg"], B)
s.setup(prob["knots"], prob["terms"], prob["x0"])
t0 = time.time()
cold = s.run(prob["xs"], prob["us"], max_iters=100, gains=False)
print(f"cold solve: {time.time() - t0:.2f} s, iterations {int(cold.num_iters.min())}..{int(cold.num_iters.max())}")
urefs = gait.force_ramp_refs(kind, prob["mass"], 34 if kind == _abi.KIND_KINO else 12, prob["cfg"].T) if kind != _abi.KIND_FULL else None
s.set_tail_warmstart(os.environ.get("WALK_TAIL", "phase") == "phase")  # WALK_TAIL=copy: the reference scripts' warm start of the appended knot
s.gait_setup(gait.device_gait(kind, prob["lf"], prob["rf"], prob["com0"], prob["mass"], x_forward=float(os.environ["WALK_X_FORWARD"]) if "WALK_X_FORWARD" in os.environ else None, y_gap=Y_GAP, w_lfrf=float(os.environ["WALK_W_FOOT"]) if "WALK_W_FOOT" in os.environ else None), mirror, urefs)
t0 = time.time()
for t in range(N):
    s.gait_tick()
    s.tick(None, None, keep_multipliers=KEEP, max_iters=1)
    if t % (60 if kind == _abi.KIND_KINO else 20) == 19 or t == N - 1:
        r = s.results(gains=False, multipliers=False)
        st = np.array([i.status for i in r.info])
        ks, _ = s.knots()
        print(f"tick {t + 1:4d}: phase at knot 0 [{int(ks[0].cs[0])} {int(ks[0].cs[1])}] | alpha median {np.median(r.alpha):.3f} min {r.alpha.min():.3f} | "
              f"prim infeas median {np.median(r.prim_infeas):.2e} max {r.prim_infeas.max():.2e} | base / CoM z {r.xs[:, 0, 2].min():.3f}..{r.xs[:, 0, 2].max():.3f} | "
              f"x {r.xs[:, 0, 0].min():+.3f}..{r.xs[:, 0, 0].max():+.3f} | failed {int((st >= 2).sum())}", flush=True)
dt = time.time() - t0
print(f"{N} ticks x {B} robots in {dt:.1f} s -> {B * N / dt:.0f} robot-ticks/s")
