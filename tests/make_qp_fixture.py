"""Generates tests/golden/qp_id_talos.npz: whole-body inverse-dynamics QP inputs of the reference's shape
(IDSolver_ulim, QP_utils.py:437-573; call site kinodynamic_talos.py:425-445) on the synthetic Talos-shaped model, plus the CPU
oracle's solution of each QP.  TEST INFRASTRUCTURE: uses the oracle (oracle/rbd.hpp for M, nle, contact dynamics; oracle/qp.hpp for
the solutions).  The fixture feeds the CPU and GPU parity tests and, as DATA, the QP leg of bench.py.

    python tests/make_qp_fixture.py

What plays the part of the pinocchio calls of kinodynamic_talos.py:425-431:
  M, nle              = crba / nonLinearEffects           -> oracle kinematics (rbd.hpp), AD/finite-difference verified there
  Jc (LOCAL, 6 x nv)  = getFrameJacobian(..., LOCAL)      -> central differences of the oracle's foot placements:
                                                             column j = log6(M(q)^-1 M(q (+) h e_j)) / h
  dJ v                = getFrameJacobianTimeVariation @ v -> directional derivative of Jc along v, central differences
  frame velocity      = getFrameVelocity                  -> Jc v
The desired acceleration / forces (a0, forces: the MPC's xdot and us[0] in the reference) are the oracle's constrained-dynamics
solution for a random torque, perturbed, so that the QP has something to correct."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from mpc_benchmark_b200 import problems  # noqa: E402

NV, NQ = 28, 29
MU, FOOT_L, FOOT_W = 0.8, 0.1, 0.075  # kinodynamic_talos.py:63-65 via SURVEY App. C2
WEIGHTS = (1.0, 1.0)  # weights_ID = [1, 1]: H = blockdiag(w0 I_nv, w1 I_12, 0)


def se3_inv_mul(A12, B12):
    Ra, pa = A12[:9].reshape(3, 3), A12[9:]
    Rb, pb = B12[:9].reshape(3, 3), B12[9:]
    R = Ra.T @ Rb
    p = Ra.T @ (pb - pa)
    return np.concatenate([R.reshape(-1), p])


def foot_jacobians(rb, x, h=1e-6):
    """LOCAL 6 x nv Jacobians of both soles by central differences on the manifold."""
    k0 = oracle_lib.kinematics(rb, x)
    J = np.zeros((2, 6, NV))
    for j in range(NV):
        dx = np.zeros(2 * NV)
        dx[j] = h
        kp = oracle_lib.kinematics(rb, oracle_lib.integrate(x, dx))
        km = oracle_lib.kinematics(rb, oracle_lib.integrate(x, -dx))
        for f, key in enumerate(("lf", "rf")):
            J[f, :, j] = (oracle_lib.log6(se3_inv_mul(k0[key], kp[key])) - oracle_lib.log6(se3_inv_mul(k0[key], km[key]))) / (2 * h)
    return J.reshape(12, NV)


def make(count=32, seed=11):
    rng = np.random.default_rng(seed)
    prob = problems.full_standing_problem(batch=1, T=4)
    rb, cfg, x_nom = prob["robot"], prob["cfg"], prob["x0"][0]
    out = {k: [] for k in ("x", "M", "nle", "Jc", "dJv", "vf", "a", "forces", "cs")}
    for i in range(count):
        dx = np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.03, 3 + 22), rng.normal(0, 0.08, NV)])
        x = oracle_lib.integrate(x_nom, dx)
        v = x[NQ:]
        cs = [(1, 1), (1, 1), (1, 0), (0, 1)][i % 4]
        k = oracle_lib.kinematics(rb, x)
        Jc = foot_jacobians(rb, x)
        hq = 1e-5
        step = np.concatenate([v, np.zeros(NV)])
        xp, xm = oracle_lib.integrate(x, hq * step), oracle_lib.integrate(x, -hq * step)
        dJv = (foot_jacobians(rb, xp) - foot_jacobians(rb, xm)) @ v / (2 * hq)
        tau = np.concatenate([np.zeros(6), rng.normal(0, 8.0, 22)])
        cd = oracle_lib.cdyn(rb, cfg, x, tau, cs, check=False)
        lam = cd["lam"].copy()
        for f in range(2):
            if not cs[f]:
                lam[6 * f:6 * f + 6] = 0
        out["x"].append(x); out["M"].append(k["M"]); out["nle"].append(k["b"]); out["Jc"].append(Jc); out["dJv"].append(dJv)
        out["vf"].append((Jc @ v).reshape(2, 6))
        out["a"].append(cd["a"] + rng.normal(0, 0.3, NV))
        out["forces"].append(lam + rng.normal(0, 5.0, 12) * np.repeat(cs, 6))
        out["cs"].append(cs)
    d = {k: np.array(v) for k, v in out.items()}
    d["cs"] = d["cs"].astype(np.int32)
    # gamma of IDSolver_ulim.computeMatrice (QP_utils.py:524-528): dJ v + Kd (v_lin + v_ang) on the linear rows, Kd = 1
    gamma = d["dJv"].copy().reshape(count, 2, 6)
    gamma[:, :, :3] += d["vf"][:, :, :3] + d["vf"][:, :, 3:]
    gamma *= d["cs"][:, :, None]
    d["gamma"] = gamma.reshape(count, 12)
    A, b, Cm, l = oracle_lib.qp_assemble_id(d["M"], d["nle"], d["Jc"], d["gamma"], d["a"], d["forces"], d["cs"], MU, FOOT_L, FOOT_W)
    n = 62
    H = np.zeros((n, n))
    H[:NV, :NV] = np.eye(NV) * WEIGHTS[0]
    H[NV:NV + 12, NV:NV + 12] = np.eye(12) * WEIGHTS[1]
    u = np.full(18, 1e5)
    st = oracle_lib.qp_default_settings(eps_abs=1e-3, eps_rel=0.0, max_iter=10, max_iter_in=10, check_duality_gap=1)  # QP_utils.py:502-508
    X, Y, Z, info = oracle_lib.qp_solve(H, np.zeros(n), A, b, Cm, l, u, settings=st)
    d.update(A=A, b=b, C=Cm, l=l, H=H, u=u, x_ref=X, y_ref=Y, z_ref=Z, status=np.array([i.status for i in info]), iters=np.array([i.iter for i in info]),
             iters_in=np.array([i.iter_in for i in info]), pri=np.array([i.pri_res for i in info]), dua=np.array([i.dua_res for i in info]))
    st2 = oracle_lib.qp_default_settings(eps_abs=1e-8, eps_rel=0.0, max_iter=100, max_iter_in=100, check_duality_gap=0)
    X2, _, _, info2 = oracle_lib.qp_solve(H, np.zeros(n), A, b, Cm, l, u, settings=st2)
    d.update(x_tight=X2, status_tight=np.array([i.status for i in info2]))
    return d


if __name__ == "__main__":
    d = make()
    # keep the file small: the assembled A / C are recomputed from the inputs by the tests
    keep = ("x", "M", "nle", "Jc", "dJv", "vf", "gamma", "a", "forces", "cs", "x_ref", "y_ref", "z_ref", "status", "iters", "iters_in", "pri", "dua", "x_tight", "status_tight")
    path = os.path.join(ROOT, "tests", "golden", "qp_id_talos.npz")
    np.savez_compressed(path, **{k: d[k] for k in keep})
    print(path, os.path.getsize(path), "bytes")
    print("status", d["status"], "iters", d["iters"], "inner", d["iters_in"])
    print("pri", d["pri"].max(), "dua", d["dua"].max(), "tight status", d["status_tight"], "|x_ref - x_tight|", np.abs(d["x_ref"] - d["x_tight"]).max())
