"""Batched dense QP (boundary proxsuite.proxqp.dense.QP as driven by QP_utils.py).

CPU (`-m "not gpu"`): the oracle (oracle/qp.hpp) against independent checks — KKT conditions, scipy's SLSQP, brute-force
active-set enumeration — the committed whole-body fixture, and the kernel SOURCE (csrc/qp.cuh) run serially under the host
emulation.  GPU (`-m gpu`): the CUDA path through the C-ABI / the proxqp shim against the oracle on the same inputs.
Tolerances: the two implementations run the same algorithm with the same iteration counts; they differ by rounding only, which the
penalty 1 / mu amplifies in the multipliers: 1e-9 relative on x, 1e-6 on y / z."""
import itertools
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "qp_id_talos.npz")
MU, FOOT_L, FOOT_W = 0.8, 0.1, 0.075


def rand_qp(rng, n, ne, ni, batch, box=False):
    Lm = rng.normal(size=(batch, n, n))
    H = Lm @ Lm.transpose(0, 2, 1) / n + 0.1 * np.eye(n)
    g = rng.normal(size=(batch, n))
    A = rng.normal(size=(batch, ne, n))
    x0 = rng.normal(size=(batch, n))
    b = np.einsum("bij,bj->bi", A, x0)
    Cm = rng.normal(size=(batch, ni, n))
    s = np.einsum("bij,bj->bi", Cm, x0)
    l, u = s - rng.uniform(0, 1, size=s.shape), s + rng.uniform(0, 1, size=s.shape)
    lb = x0 - rng.uniform(0, 0.5, size=x0.shape) if box else None
    ub = x0 + rng.uniform(0, 0.5, size=x0.shape) if box else None
    return H, g, A, b, Cm, l, u, lb, ub


def kkt_violation(H, g, A, b, Cm, l, u, x, y, z, lb=None, ub=None):
    n, ni = x.size, l.size
    zc, zb = z[:ni], z[ni:]
    dua = H @ x + g + A.T @ y + Cm.T @ zc + (zb if lb is not None else 0)
    s = Cm @ x
    pri = max(np.abs(A @ x - b).max(initial=0), np.maximum(s - u, 0).max(initial=0), np.maximum(l - s, 0).max(initial=0))
    comp = max(np.abs(np.maximum(zc, 0) * (u - s)).max(initial=0), np.abs(np.minimum(zc, 0) * (s - l)).max(initial=0))
    if lb is not None:
        pri = max(pri, np.maximum(x - ub, 0).max(), np.maximum(lb - x, 0).max())
        comp = max(comp, np.abs(np.maximum(zb, 0) * (ub - x)).max(), np.abs(np.minimum(zb, 0) * (x - lb)).max())
    return np.abs(dua).max(), pri, comp


def whole_body_qp(oracle):
    d = np.load(FIXTURE)
    A, b, Cm, l = oracle.qp_assemble_id(d["M"], d["nle"], d["Jc"], d["gamma"], d["a"], d["forces"], d["cs"], MU, FOOT_L, FOOT_W)
    n = 62
    H = np.zeros((n, n))
    H[:28, :28] = np.eye(28)
    H[28:40, 28:40] = np.eye(12)
    return d, H, np.zeros(n), A, b, Cm, l, np.full(18, 1e5)


# ------------------------------------------------------------------------------------------------ oracle pinned (CPU)
@pytest.mark.parametrize("shape", [(12, 4, 6, False), (20, 8, 12, False), (9, 3, 5, True), (16, 0, 8, False), (10, 4, 0, False)])
def test_oracle_satisfies_kkt(oracle, shape):
    n, ne, ni, box = shape
    rng = np.random.default_rng(1)
    H, g, A, b, Cm, l, u, lb, ub = rand_qp(rng, n, ne, ni, 6, box)
    st = oracle.qp_default_settings(eps_abs=1e-8, max_iter=200, max_iter_in=100, check_duality_gap=1)  # the gap test is what enforces complementarity
    X, Y, Z, info = oracle.qp_solve(H, g, A, b, Cm, l, u, lb, ub, settings=st)
    for i in range(6):
        assert info[i].status == 0
        dua, pri, comp = kkt_violation(H[i], g[i], A[i], b[i], Cm[i], l[i], u[i], X[i], Y[i], Z[i], lb[i] if box else None, ub[i] if box else None)
        assert dua < 1e-7 and pri < 1e-7 and comp < 1e-6, (dua, pri, comp)


def test_oracle_matches_scipy(oracle):
    from scipy.optimize import minimize

    rng = np.random.default_rng(2)
    H, g, A, b, Cm, l, u, _, _ = rand_qp(rng, 14, 5, 9, 4)
    st = oracle.qp_default_settings(eps_abs=1e-9, max_iter=200, max_iter_in=100)
    X, _, _, info = oracle.qp_solve(H, g, A, b, Cm, l, u, settings=st)
    for i in range(4):
        cons = [dict(type="eq", fun=lambda x, i=i: A[i] @ x - b[i], jac=lambda x, i=i: A[i]),
                dict(type="ineq", fun=lambda x, i=i: Cm[i] @ x - l[i], jac=lambda x, i=i: Cm[i]),
                dict(type="ineq", fun=lambda x, i=i: u[i] - Cm[i] @ x, jac=lambda x, i=i: -Cm[i])]
        r = minimize(lambda x: 0.5 * x @ H[i] @ x + g[i] @ x, np.zeros(14), jac=lambda x: H[i] @ x + g[i], constraints=cons, method="SLSQP",
                     options=dict(ftol=1e-13, maxiter=1000))
        assert info[i].status == 0  # (SLSQP may stop with 'positive directional derivative' AT the solution: compare the points)
        assert np.abs(r.x - X[i]).max() < 1e-6


def test_oracle_matches_active_set_enumeration(oracle):
    """Strictly convex QP with 4 two-sided inequalities: enumerate every active set (3^4), solve its equality-constrained KKT system,
    keep the feasible one with correctly signed multipliers — the unique solution."""
    rng = np.random.default_rng(3)
    n, ne, ni = 6, 2, 4
    H, g, A, b, Cm, l, u, _, _ = rand_qp(rng, n, ne, ni, 5)
    st = oracle.qp_default_settings(eps_abs=1e-10, max_iter=300, max_iter_in=100)
    X, _, _, info = oracle.qp_solve(H, g, A, b, Cm, l, u, settings=st)
    for q in range(5):
        best = None
        for act in itertools.product((0, 1, 2), repeat=ni):  # 0 free, 1 at lower, 2 at upper
            rows = [i for i in range(ni) if act[i]]
            E = np.vstack([A[q]] + [Cm[q][i:i + 1] for i in rows])
            rhs = np.concatenate([b[q]] + [[l[q][i]] if act[i] == 1 else [u[q][i]] for i in rows])
            K = np.block([[H[q], E.T], [E, np.zeros((E.shape[0], E.shape[0]))]])
            try:
                sol = np.linalg.solve(K, np.concatenate([-g[q], rhs]))
            except np.linalg.LinAlgError:
                continue
            x, mult = sol[:n], sol[n + ne:]
            s = Cm[q] @ x
            if (s < l[q] - 1e-9).any() or (s > u[q] + 1e-9).any():
                continue
            if any((act[i] == 1 and m > 1e-9) or (act[i] == 2 and m < -1e-9) for i, m in zip(rows, mult)):
                continue
            best = x
            break
        assert best is not None and info[q].status == 0
        assert np.abs(best - X[q]).max() < 1e-7


def test_whole_body_fixture_is_reproduced(oracle):
    """The committed fixture (tests/make_qp_fixture.py) is what the current oracle computes, at the reference's settings
    (eps_abs 1e-3, 10 x 10 iterations, duality-gap check: QP_utils.py:502-508), and every QP meets them."""
    d, H, g, A, b, Cm, l, u = whole_body_qp(oracle)
    st = oracle.qp_default_settings(eps_abs=1e-3, eps_rel=0.0, max_iter=10, max_iter_in=10, check_duality_gap=1)
    X, Y, Z, info = oracle.qp_solve(H, g, A, b, Cm, l, u, settings=st)
    assert all(i.status == 0 for i in info)
    assert np.array_equal(d["iters"], [i.iter for i in info])
    assert np.abs(X - d["x_ref"]).max() < 1e-9 * np.abs(d["x_ref"]).max()
    # physics of the answer: the corrected accelerations / forces / torques satisfy the rigid contact dynamics to eps_abs
    nv = 28
    for i in range(X.shape[0]):
        da, df, tau = X[i, :nv], X[i, nv:nv + 12], X[i, nv + 12:]
        anew, fnew = d["a"][i] + da, d["forces"][i] + df
        cs = np.repeat(d["cs"][i], 6)
        lhs = d["M"][i] @ anew + d["nle"][i] - d["Jc"][i].T @ (fnew * cs) - np.concatenate([np.zeros(6), tau])
        assert np.abs(lhs).max() < 2e-3
        for f in range(2):
            if d["cs"][i][f]:
                fx, fy, fz = fnew[6 * f:6 * f + 3]
                assert fz > -1e-3 and abs(fx) <= MU * fz + 2e-3


# --------------------------------------------------------------------- kernel source under host emulation (CPU)
@pytest.mark.parametrize("shape", [(20, 8, 12, False), (62, 40, 18, False), (13, 5, 7, True), (28, 0, 10, False), (17, 6, 0, False)])
def test_emulated_kernel_matches_oracle(oracle, shape):
    import emu_lib

    n, ne, ni, box = shape
    rng = np.random.default_rng(4)
    H, g, A, b, Cm, l, u, lb, ub = rand_qp(rng, n, ne, ni, 5, box)
    for eps, mi, mii, gap in [(1e-7, 100, 50, 0), (1e-3, 10, 10, 1)]:
        st = oracle.qp_default_settings(eps_abs=eps, max_iter=mi, max_iter_in=mii, check_duality_gap=gap)
        X, Y, Z, info = oracle.qp_solve(H, g, A, b, Cm, l, u, lb, ub, settings=st)
        X2, Y2, Z2, info2 = emu_lib.qp_solve(H, g, A, b, Cm, l, u, lb, ub, settings=st)
        assert [(i.status, i.iter, i.iter_in) for i in info] == [(i.status, i.iter, i.iter_in) for i in info2]
        assert np.abs(X - X2).max() < 1e-9 * max(1.0, np.abs(X).max())
        if Y.size:
            assert np.abs(Y - Y2).max() < 1e-6 * max(1.0, np.abs(Y).max())
        if Z.size:
            assert np.abs(Z - Z2).max() < 1e-6 * max(1.0, np.abs(Z).max())


def test_emulated_whole_body_assembly_and_solve(oracle):
    import emu_lib

    d, H, g, A, b, Cm, l, u = whole_body_qp(oracle)
    A2, b2, C2, l2 = emu_lib.qp_assemble_id(d["M"], d["nle"], d["Jc"], d["gamma"], d["a"], d["forces"], d["cs"], MU, FOOT_L, FOOT_W)
    assert np.array_equal(A, A2) and np.array_equal(Cm, C2)
    assert np.abs(l - l2).max() < 1e-12 and np.abs(b - b2).max() < 1e-10  # fused multiply-adds differ between the two builds
    st = oracle.qp_default_settings(eps_abs=1e-3, eps_rel=0.0, max_iter=10, max_iter_in=10, check_duality_gap=1)
    X2, _, _, info2 = emu_lib.qp_solve(H, g, A2, b2, C2, l2, u, settings=st)
    assert np.array_equal(d["iters"], [i.iter for i in info2])
    assert np.abs(X2 - d["x_ref"]).max() < 1e-9 * np.abs(d["x_ref"]).max()


def test_assembly_matches_reference_formulas(oracle):
    """IDSolver_ulim.computeMatrice written out with numpy exactly as QP_utils.py:514-552 does it."""
    d, H, g, A, b, Cm, l, u = whole_body_qp(oracle)
    nv, fs, nk, n = 28, 6, 2, 62
    S = np.zeros((nv, nv - 6))
    S[6:, :] = np.eye(nv - 6)
    Cmin = np.array([[-1, 0, MU, 0, 0, 0], [1, 0, MU, 0, 0, 0], [-1, 0, MU, 0, 0, 0], [1, 0, MU, 0, 0, 0], [0, 0, 1, 0, 0, 0],
                     [0, 0, FOOT_W, -1, 0, 0], [0, 0, FOOT_W, 1, 0, 0], [0, 0, FOOT_L, 0, -1, 0], [0, 0, FOOT_L, 0, 1, 0]])
    for i in range(d["M"].shape[0]):
        cs, f, a, M = d["cs"][i], d["forces"][i], d["a"][i], d["M"][i]
        Jc = d["Jc"][i] * np.repeat(cs, 6)[:, None]
        gamma = d["gamma"][i]
        Ar = np.zeros((nv + fs * nk, n))
        Ar[:nv, :nv] = M
        Ar[:nv, nv:nv + nk * fs] = -Jc.T
        Ar[:nv, nv + nk * fs:] = -S
        Ar[nv:, :nv] = Jc
        br = np.concatenate([-d["nle"][i] - M @ a + Jc.T @ f, -gamma - Jc @ a])
        lr, Cr = np.zeros(9 * nk), np.zeros((9 * nk, n))
        for k in range(nk):
            if cs[k]:
                fk = f[k * fs:(k + 1) * fs]
                lr[k * 9:(k + 1) * 9] = [fk[0] - fk[2] * MU, -fk[0] - fk[2] * MU, fk[1] - fk[2] * MU, -fk[1] - fk[2] * MU, -fk[2],
                                         fk[3] - fk[2] * FOOT_W, -fk[3] - fk[2] * FOOT_W, fk[4] - fk[2] * FOOT_L, -fk[4] - fk[2] * FOOT_L]
                Cr[k * 9:(k + 1) * 9, nv + k * fs:nv + (k + 1) * fs] = Cmin
        assert np.array_equal(Ar, A[i]) and np.array_equal(Cr, Cm[i])
        assert np.abs(lr - l[i]).max() < 1e-12 and np.abs(br - b[i]).max() < 1e-10


def test_proxqp_shim_surface():
    """The names QP_utils.py touches exist with the reference's signatures (no compute: no GPU here)."""
    import mpc_benchmark_b200.proxqp as proxsuite

    qp = proxsuite.proxqp.dense.QP(62, 40, 18, False, dense_backend=proxsuite.proxqp.dense.DenseBackend.PrimalDualLDLT)
    qp.settings.eps_abs = 1e-3
    qp.settings.eps_rel = 0
    qp.settings.primal_infeasibility_solving = True
    qp.settings.check_duality_gap = True
    qp.settings.verbose = False
    qp.settings.max_iter = 10
    qp.settings.max_iter_in = 10
    c = qp.settings.to_c(False)
    assert (c.eps_abs, c.max_iter, c.max_iter_in, c.check_duality_gap) == (1e-3, 10, 10, 1)
    with pytest.raises(RuntimeError):
        qp.solve()
    qb = proxsuite.proxqp.dense.QP(62, 40, 18, True)
    assert qb.box and qb.nz == 80


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
def _gpu_vs_oracle(oracle, H, g, A, b, Cm, l, u, lb, ub, st_kw, xtol=1e-9, mtol=1e-6):
    from mpc_benchmark_b200 import proxqp

    n, ne, ni, batch = g.shape[-1], b.shape[-1], l.shape[-1], g.shape[0]
    qp = proxqp.dense.BatchQP(n, ne, ni, batch, lb is not None)
    for k, v in st_kw.items():
        setattr(qp.settings, k, v)
    qp.init(H, g, A if ne else None, b if ne else None, Cm if ni else None, l if ni else None, u if ni else None, lb, ub)
    r = qp.solve()
    st = oracle.qp_default_settings(eps_abs=st_kw["eps_abs"], max_iter=st_kw["max_iter"], max_iter_in=st_kw["max_iter_in"],
                                    check_duality_gap=int(st_kw.get("check_duality_gap", False)))
    X, Y, Z, info = oracle.qp_solve(H, g, A, b, Cm, l, u, lb, ub, settings=st)
    # same algorithm, same decisions: iteration counts agree except where a rounding difference flips one Newton step on one instance
    # (north_star: "same iteration count +-1"); instances with identical counts must agree to rounding, the others to the tolerance asked
    it_o, in_o = np.array([i.iter for i in info]), np.array([i.iter_in for i in info])
    same = (np.asarray(r.info.iter_ext) == it_o) & (np.asarray(r.info.iter) == in_o) & (np.asarray(r.info.status) == [i.status for i in info])
    assert same.mean() >= 0.9 and np.abs(np.asarray(r.info.iter_ext) - it_o).max() <= 1
    scale = max(1.0, np.abs(X).max())
    assert np.abs(r.x - X)[same].max() < xtol * scale
    if (~same).any():
        assert np.abs(r.x - X)[~same].max() < 10 * st_kw["eps_abs"] * scale
    if Y.size:
        assert np.abs(r.y - Y)[same].max() < mtol * max(1.0, np.abs(Y).max())
    if Z.size:
        assert np.abs(r.z - Z)[same].max() < mtol * max(1.0, np.abs(Z).max())
    qp.close()
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(20, 8, 12, False), (62, 40, 18, False), (13, 5, 7, True), (28, 0, 10, False), (17, 6, 0, False), (64, 64, 32, False),
                                   (62, 40, 18, True)])
def test_gpu_random_qps_match_oracle(oracle, shape):
    n, ne, ni, box = shape
    rng = np.random.default_rng(6)
    H, g, A, b, Cm, l, u, lb, ub = rand_qp(rng, n, min(ne, n - 2), ni, 37, box)
    _gpu_vs_oracle(oracle, H, g, A, b, Cm, l, u, lb, ub, dict(eps_abs=1e-7, max_iter=100, max_iter_in=50))
    _gpu_vs_oracle(oracle, H, g, A, b, Cm, l, u, lb, ub, dict(eps_abs=1e-3, max_iter=10, max_iter_in=10, check_duality_gap=True))


@pytest.mark.gpu
def test_gpu_whole_body_id_solver_matches_fixture(oracle):
    """IDSolver_ulim mirror: device assembly + batched solve at the reference's settings against the committed oracle solutions."""
    from mpc_benchmark_b200 import pin, qp_utils

    d = np.load(FIXTURE)
    B = d["M"].shape[0]
    model = pin.load_talos_like()[0]
    solver = qp_utils.IDSolver_ulim(model, [1, 1], 2, MU, FOOT_L, FOOT_W, [0, 1], 6, False, batch=B)
    rbd = qp_utils.RBDTerms(nle=d["nle"], Jc=d["Jc"], dJv=d["dJv"], vf=d["vf"])
    v = d["x"][:, 29:]
    anew, fnew, tau = solver.solve(rbd, d["cs"], v, d["a"], d["forces"], d["M"])
    x = np.concatenate([anew - d["a"], fnew - d["forces"], tau], axis=1)
    assert np.abs(x - d["x_ref"]).max() < 1e-9 * np.abs(d["x_ref"]).max()
    assert np.array_equal(solver.qp.results.info.iter_ext, d["iters"])
    assert (solver.qp.results.info.status == 0).all()
    assert np.abs(solver.gamma(rbd, d["cs"]) - d["gamma"]).max() < 1e-12


@pytest.mark.gpu
def test_gpu_large_batch_properties():
    """Batch 4096 (tiled fixture with perturbed desired accelerations / forces): every QP meets the reference tolerances, its KKT
    residuals recomputed on the host agree with what the kernel reports, and identical QPs give identical answers."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mpc_benchmark_b200 import pin, qp_utils

    d = np.load(FIXTURE)
    B, reps = 4096, 4096 // d["M"].shape[0]
    rng = np.random.default_rng(7)
    tile = lambda a: np.tile(a, (reps,) + (1,) * (a.ndim - 1))  # noqa: E731
    a = tile(d["a"]) + rng.normal(0, 0.05, (B, 28))
    f = tile(d["forces"]) + rng.normal(0, 1.0, (B, 12)) * np.repeat(tile(d["cs"]), 6, axis=1)
    a[:32], f[:32] = d["a"], d["forces"]
    a[32:64], f[32:64] = d["a"], d["forces"]
    model = pin.load_talos_like()[0]
    solver = qp_utils.IDSolver_ulim(model, [1, 1], 2, MU, FOOT_L, FOOT_W, [0, 1], 6, False, batch=B)
    rbd = qp_utils.RBDTerms(nle=tile(d["nle"]), Jc=tile(d["Jc"]), dJv=tile(d["dJv"]), vf=tile(d["vf"]))
    anew, fnew, tau = solver.solve(rbd, tile(d["cs"]), None, a, f, tile(d["M"]))
    info = solver.qp.results.info
    assert (info.status == 0).all() and info.pri_res.max() <= 1e-3 and info.dua_res.max() <= 1e-3
    assert np.array_equal(anew[:32], anew[32:64]) and np.array_equal(tau[:32], tau[32:64])
    M, Jc, nle, cs = tile(d["M"]), tile(d["Jc"]), tile(d["nle"]), np.repeat(tile(d["cs"]), 6, axis=1)
    lhs = np.einsum("bij,bj->bi", M, anew) + nle - np.einsum("bji,bj->bi", Jc, fnew * cs)
    lhs[:, 6:] -= tau
    assert np.abs(lhs).max() < 2e-3


# ------------------------------------------------------------------------ IKIDSolver_f6 mirror (QP_utils.py:584-768)
def _ikid_inputs(B, seed=9):
    """Fixture states + synthetic task terms (base / torso Jacobians, momentum matrix, PD errors) of the right shapes."""
    d = np.load(FIXTURE)
    rng = np.random.default_rng(seed)
    reps = -(-B // d["M"].shape[0])
    tile = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:B])  # noqa: E731
    t = dict(M=tile(d["M"]), nle=tile(d["nle"]), Jc=tile(d["Jc"]), dJv=tile(d["dJv"]), cs=tile(d["cs"]), forces=tile(d["forces"]), v=tile(d["x"])[:, 29:])
    t.update(J_base=rng.normal(0, 0.5, (B, 3, 28)), dJv_base=rng.normal(0, 0.1, (B, 3)), J_torso=rng.normal(0, 0.5, (B, 3, 28)),
             dJv_torso=rng.normal(0, 0.1, (B, 3)), Ag=rng.normal(0, 2.0, (B, 6, 28)), dAgv=rng.normal(0, 0.2, (B, 6)), dH=rng.normal(0, 1.0, (B, 6)),
             q_diff=rng.normal(0, 0.02, (B, 28)), dq_diff=rng.normal(0, 0.05, (B, 28)), LF_diff=rng.normal(0, 0.01, (B, 6)), dLF_diff=rng.normal(0, 0.02, (B, 6)),
             RF_diff=rng.normal(0, 0.01, (B, 6)), dRF_diff=rng.normal(0, 0.02, (B, 6)), base_diff=rng.normal(0, 0.02, (B, 3)), dbase_diff=rng.normal(0, 0.05, (B, 3)),
             torso_diff=rng.normal(0, 0.02, (B, 3)), dtorso_diff=rng.normal(0, 0.05, (B, 3)))
    Kp = lambda k, m: np.eye(m) * k  # noqa: E731
    t["K_gains"] = [[Kp(100, 28), Kp(20, 28)], [Kp(400, 6), Kp(40, 6)], None, [Kp(100, 3), Kp(20, 3)]]
    t["weights"] = [1.0, 100.0, 0.1, 10.0, 1e-3]
    return t


def _ikid_reference_cost(t, i):
    """H[:nv,:nv], g[:nv] of instance i written as QP_utils.py:677-691 writes them (plain 2-D numpy, one instance)."""
    w, K, nv = t["weights"], t["K_gains"], 28
    Jc_left, Jc_right, Jc_base, Jc_torso, Ag = t["Jc"][i][:6], t["Jc"][i][6:], t["J_base"][i], t["J_torso"][i], t["Ag"][i]
    dJl_v, dJr_v = t["dJv"][i][:6], t["dJv"][i][6:]
    H = w[0] * np.eye(nv)
    H += w[1] * Jc_left.transpose() @ Jc_left
    H += w[1] * Jc_right.transpose() @ Jc_right
    H += w[2] * Ag.transpose() @ Ag
    H += w[3] * Jc_base.transpose() @ Jc_base
    H += w[3] * Jc_torso.transpose() @ Jc_torso
    g = w[0] * (-K[0][0] @ t["q_diff"][i] - K[0][1] @ t["dq_diff"][i])
    g += w[1] * (dJl_v - K[1][0] @ t["LF_diff"][i] - K[1][1] @ t["dLF_diff"][i]).transpose() @ Jc_left
    g += w[1] * (dJr_v - K[1][0] @ t["RF_diff"][i] - K[1][1] @ t["dRF_diff"][i]).transpose() @ Jc_right
    g -= w[2] * (t["dH"][i] - t["dAgv"][i]).transpose() @ Ag
    g += w[3] * (t["dJv_base"][i] - K[3][0] @ t["base_diff"][i] - K[3][1] @ t["dbase_diff"][i]).transpose() @ Jc_base
    g += w[3] * (t["dJv_torso"][i] - K[3][0] @ t["torso_diff"][i] - K[3][1] @ t["dtorso_diff"][i]).transpose() @ Jc_torso
    return H, g


class _NoDeviceQP:
    """Stands in for proxqp.BatchQP where no GPU exists: records what the mirror hands to the solver."""

    def __init__(self, *a, **k):
        import types

        self.settings, self.calls = types.SimpleNamespace(), {}

    def init(self, *a, **k):
        self.calls["init"] = a

    def assemble_id(self, *a):
        self.calls["assemble_id"] = a


def test_ikid_cost_assembly_matches_reference_formulas(monkeypatch):
    from mpc_benchmark_b200 import pin, proxqp, qp_utils

    monkeypatch.setattr(proxqp.dense, "BatchQP", _NoDeviceQP)
    B = 6
    t = _ikid_inputs(B)
    s = qp_utils.IKIDSolver_f6(pin.load_talos_like()[0], t["weights"], t["K_gains"], 2, MU, FOOT_L, FOOT_W, [0, 1], 2, 3, 6, False, batch=B)
    rbd = qp_utils.RBDTermsIKID(nle=t["nle"], Jc=t["Jc"], dJv=t["dJv"], J_base=t["J_base"], dJv_base=t["dJv_base"], J_torso=t["J_torso"],
                                dJv_torso=t["dJv_torso"], Ag=t["Ag"], dAgv=t["dAgv"])
    s.computeMatrice(rbd, t["cs"], t["v"], t["q_diff"], t["dq_diff"], t["LF_diff"], t["dLF_diff"], t["RF_diff"], t["dRF_diff"], t["base_diff"], t["dbase_diff"],
                     t["torso_diff"], t["dtorso_diff"], t["forces"], t["dH"], t["M"])
    for i in range(B):
        H, g = _ikid_reference_cost(t, i)
        assert np.abs(s.H[i, :28, :28] - H).max() < 1e-9 * np.abs(H).max()
        assert np.abs(s.g[i, :28] - g).max() < 1e-9 * max(1.0, np.abs(g).max())
        assert np.array_equal(np.diag(s.H[i])[28:40], np.full(12, t["weights"][4])) and not s.H[i, 40:, :].any() and not s.g[i, 28:].any()
    a = s.qp.calls["assemble_id"]
    assert not a[4].any()  # a = 0: the acceleration itself is the unknown (QP_utils.py:717,757)
    assert np.array_equal(a[3], t["dJv"] * np.repeat(t["cs"], 6, axis=1))  # gamma = dJ v on the active contacts, no Baumgarte term (QP_utils.py:722-727)
    assert s.l_box[40] == -100.0 and s.u_box[41] == 160.0 and s.l_box[0] == -100000  # torque limits from effortLimit[6:] (QP_utils.py:612-615)


@pytest.mark.gpu
def test_gpu_ikid_solver_matches_oracle(oracle):
    from mpc_benchmark_b200 import pin, qp_utils

    B = 48
    t = _ikid_inputs(B)
    s = qp_utils.IKIDSolver_f6(pin.load_talos_like()[0], t["weights"], t["K_gains"], 2, MU, FOOT_L, FOOT_W, [0, 1], 2, 3, 6, False, batch=B)
    rbd = qp_utils.RBDTermsIKID(nle=t["nle"], Jc=t["Jc"], dJv=t["dJv"], J_base=t["J_base"], dJv_base=t["dJv_base"], J_torso=t["J_torso"],
                                dJv_torso=t["dJv_torso"], Ag=t["Ag"], dAgv=t["dAgv"])
    anew, fnew, tau = s.solve(rbd, t["cs"], t["v"], t["q_diff"], t["dq_diff"], t["LF_diff"], t["dLF_diff"], t["RF_diff"], t["dRF_diff"], t["base_diff"],
                              t["dbase_diff"], t["torso_diff"], t["dtorso_diff"], t["forces"], t["dH"], t["M"])
    # the same QPs through the oracle: cost from the literal formulas, constraints from the oracle's assembly with a = 0, gamma = dJ v
    n = 62
    H, g = np.zeros((B, n, n)), np.zeros((B, n))
    for i in range(B):
        H[i, :28, :28], g[i, :28] = _ikid_reference_cost(t, i)
        H[i, np.arange(28, 40), np.arange(28, 40)] = t["weights"][4]
    gamma = t["dJv"] * np.repeat(t["cs"], 6, axis=1)
    A, b, Cm, l = oracle.qp_assemble_id(t["M"], t["nle"], t["Jc"], gamma, np.zeros((B, 28)), t["forces"], t["cs"], MU, FOOT_L, FOOT_W)
    st = oracle.qp_default_settings(eps_abs=1e-3, eps_rel=0.0, max_iter=100, max_iter_in=100, check_duality_gap=1)
    X, Y, Z, info = oracle.qp_solve(H, g, A, b, Cm, l, np.full(18, 1e5), s.l_box, s.u_box, settings=st)
    r = s.qp.results
    same = (r.info.iter_ext == [i.iter for i in info]) & (r.info.iter == [i.iter_in for i in info])
    assert same.mean() >= 0.9 and (r.info.status == [i.status for i in info]).all() and (r.info.status == 0).all()
    x = np.concatenate([anew, fnew - t["forces"], tau], axis=1)
    assert np.abs(x - X)[same].max() < 1e-8 * np.abs(X).max()
    assert np.abs(x - X).max() < 1e-2 * np.abs(X).max()
    eff = np.asarray(pin.load_talos_like()[0].effortLimit)[6:]
    assert (np.abs(tau) <= eff + 2e-3).all()  # the torque box holds


# ------------------------------------------------- rigid-body terms in front of the QP (csrc/rbd_terms.cuh; kinodynamic_talos.py:425-431)
def _check_rbd_terms(o, d):
    """Against the fixture: M, nle from the oracle's rigid-body code (exact to rounding); Jc, dJ v from central differences of the
    oracle's foot placements (their truncation error bounds the comparison)."""
    assert np.abs(o["M"] - d["M"]).max() < 1e-10 * np.abs(d["M"]).max()
    assert np.abs(o["nle"] - d["nle"]).max() < 1e-10 * np.abs(d["nle"]).max()
    assert np.abs(o["Jc"] - d["Jc"]).max() < 1e-8
    assert np.abs(o["vf"] - d["vf"]).max() < 1e-8
    assert np.abs(o["dJv"] - d["dJv"]).max() < 2e-5


def test_emulated_rbd_terms_match_oracle_fixture():
    import emu_lib
    from mpc_benchmark_b200 import problems

    d = np.load(FIXTURE)
    prob = problems.full_standing_problem(batch=1, T=4)
    _check_rbd_terms(emu_lib.rbd_terms(prob["robot"], prob["cfg"], d["x"]), d)


@pytest.mark.gpu
def test_gpu_rbd_terms_and_solve_from_state(oracle):
    """x -> (M, nle, Jc, dJ v, vf) on the device, and the whole kinodynamic_talos.py:425-445 step (rigid-body terms, assembly, QP) from
    the measured states alone, against the committed oracle solutions."""
    from mpc_benchmark_b200 import pin, problems, qp_utils
    from mpc_benchmark_b200.batch import BatchSolver

    d = np.load(FIXTURE)
    B = d["M"].shape[0]
    prob = problems.full_standing_problem(batch=1, T=4)
    s = BatchSolver(prob["robot"], prob["cfg"], 1)
    _check_rbd_terms(s.rbd_terms(d["x"]), d)
    solver = qp_utils.IDSolver_ulim(pin.load_talos_like()[0], [1, 1], 2, MU, FOOT_L, FOOT_W, [0, 1], 6, False, batch=B)
    anew, fnew, tau = solver.solve_from_state(s, d["x"], d["cs"], d["a"], d["forces"])
    x = np.concatenate([anew - d["a"], fnew - d["forces"], tau], axis=1)
    # the fixture's dJ v carries a finite-difference error of ~1e-5 that the exact device value does not: compare at that level
    assert np.abs(x - d["x_ref"]).max() < 1e-3 * np.abs(d["x_ref"]).max()
    assert (solver.qp.results.info.status == 0).all()
    # and exactly against the oracle fed with the device's own terms
    o = s.rbd_terms(d["x"])
    g = o["dJv"].reshape(B, 2, 6).copy()
    g[:, :, :3] += o["vf"][:, :, :3] + o["vf"][:, :, 3:]
    gamma = (g * d["cs"][:, :, None]).reshape(B, 12)
    A, b, Cm, l = oracle.qp_assemble_id(o["M"], o["nle"], o["Jc"], gamma, d["a"], d["forces"], d["cs"], MU, FOOT_L, FOOT_W)
    H = np.zeros((62, 62))
    H[:40, :40] = np.eye(40)
    st = oracle.qp_default_settings(eps_abs=1e-3, eps_rel=0.0, max_iter=10, max_iter_in=10, check_duality_gap=1)
    X, _, _, info = oracle.qp_solve(H, np.zeros(62), A, b, Cm, l, np.full(18, 1e5), settings=st)
    same = np.asarray(solver.qp.results.info.iter) == [i.iter_in for i in info]
    assert same.mean() >= 0.9 and np.abs(x - X)[same].max() < 1e-9 * np.abs(X).max()
    s.close()


def test_qp_fails_loudly_without_gpu():
    """No CPU fallback on the QP path either: without a CUDA device creating the native handle raises (and nothing is solved)."""
    import torch

    from mpc_benchmark_b200 import _native, proxqp

    if torch.cuda.is_available() or not os.path.exists(_native.LIB_PATH):
        pytest.skip("needs the built library and NO GPU")
    qp = proxqp.dense.QP(4, 2, 2)
    with pytest.raises(_native.NativeError):
        qp.init(np.eye(4), np.zeros(4), np.zeros((2, 4)), np.zeros(2), np.zeros((2, 4)), -np.ones(2), np.ones(2))


@pytest.mark.gpu
def test_gpu_qp_argument_errors():
    from mpc_benchmark_b200 import _native, proxqp

    with pytest.raises(_native.NativeError):
        proxqp.dense.BatchQP(65, 2, 2, 3)._handle()  # n above MPC_QP_MAXN
    qp = proxqp.dense.BatchQP(6, 2, 3, 4)
    with pytest.raises(RuntimeError):
        qp.solve()  # before init
    with pytest.raises(ValueError):
        qp.init(np.eye(6), np.zeros(6), None, None, np.zeros((3, 6)), np.zeros(3), np.zeros(3))  # A, b missing for n_eq > 0
    with pytest.raises(ValueError):
        qp.init(np.eye(6), np.zeros(5), np.zeros((2, 6)), np.zeros(2), np.zeros((3, 6)), np.zeros(3), np.zeros(3))  # wrong size
    qp.init(np.eye(6), np.ones(6), np.zeros((2, 6)), np.zeros(2), np.zeros((3, 6)), -np.ones(3), np.ones(3))
    r = qp.solve()
    assert np.allclose(r.x, -1.0, atol=1e-4) and (r.info.status == 0).all()  # unconstrained minimiser of 1/2 |x|^2 + 1'x
    # non-finite data is contained: status 2, no hang
    H = np.eye(6)
    H[0, 0] = np.nan
    qp.update(H=H)
    r = qp.solve()
    assert (r.info.status == 2).all()
    qp.close()
