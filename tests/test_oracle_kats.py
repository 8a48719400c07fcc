"""Known-answer / self-consistency tests that pin the CPU oracle.  The reference has no tests and its
dependencies cannot be installed, so these replace golden vectors: every analytic Jacobian is checked against
forward-mode AD of the templated value code, physics invariants are checked, and conventions the reference DOES pin
(cone rows of QP_utils.py:337-347, force sign of fulldynamic_talos.py:285) are asserted."""
import numpy as np
import pytest

from mpc_benchmark_b200 import _abi, problems


@pytest.fixture(scope="module")
def full(oracle):
    return problems.full_standing_problem(T=4)


def rand_state(oracle, x0, rng, s=0.2):
    return oracle.integrate(x0, rng.normal(size=56) * s)


def test_lie_group_jacobians_match_ad(oracle):
    rng = np.random.default_rng(0)
    for scale in [0.0, 1e-9, 1e-5, 1e-2, 0.5, 2.5]:
        xi = rng.normal(size=6) * scale
        assert oracle.check_jlog6(oracle.exp6(xi)) < 1e-11
        assert oracle.check_jexp6(xi) < 1e-12
        assert np.abs(oracle.log6(oracle.exp6(xi)) - xi).max() < 1e-12


def test_integrate_difference_roundtrip(oracle, full):
    rng = np.random.default_rng(1)
    x0 = full["x0"][0]
    for _ in range(5):
        x = rand_state(oracle, x0, rng)
        y = rand_state(oracle, x0, rng)
        z = oracle.integrate(x, oracle.difference(x, y))
        q_same = min(np.abs(z[3:7] - y[3:7]).max(), np.abs(z[3:7] + y[3:7]).max())
        assert np.abs(np.delete(z - y, [3, 4, 5, 6])).max() < 1e-12 and q_same < 1e-12


def test_rnea_crba_consistency(oracle, full):
    rng = np.random.default_rng(2)
    rb = full["robot"]
    for _ in range(3):
        x = rand_state(oracle, full["x0"][0], rng)
        k = oracle.kinematics(rb, x)
        M = k["M"]
        assert np.abs(M - M.T).max() < 1e-12 and np.linalg.eigvalsh(M).min() > 0
        a = rng.normal(size=28)
        assert np.abs(oracle.rnea(rb, x, a) - (M @ a + k["b"])).max() < 1e-10
        # A_g v = h: the linear momentum is mass * CoM velocity
        assert abs(k["mass"] - sum(rb.mass[b] for b in range(rb.nb))) < 1e-12


def test_centroidal_momentum_derivatives_match_ad(oracle, full):
    rng = np.random.default_rng(3)
    for _ in range(3):
        assert oracle.check_centroidal(full["robot"], rand_state(oracle, full["x0"][0], rng)) < 1e-10


@pytest.mark.parametrize("active", [(1, 1), (1, 0), (0, 1)])
def test_constrained_dynamics_derivatives_match_ad(oracle, full, active):
    rng = np.random.default_rng(4)
    x = rand_state(oracle, full["x0"][0], rng, 0.1)
    tau = np.concatenate([np.zeros(6), rng.normal(size=22) * 20])
    o = oracle.cdyn(full["robot"], full["cfg"], x, tau, active)
    assert o["errs"].max() < 1e-8, o["errs"]
    # the constraint rows hold: J a + gamma = a* up to the proximal regularisation mu * lambda
    for f in range(2):
        if not active[f]:
            assert np.abs(o["lam"][6 * f: 6 * f + 6]).max() == 0


def test_standing_fixed_point(oracle, full):
    """At half-sitting, v = 0, gravity-compensating torques: a = 0 and the vertical contact forces carry m g
    (f_half > 0 convention of fulldynamic_talos.py:285)."""
    rb, cfg, x0 = full["robot"], full["cfg"], full["x0"][0]
    k = oracle.kinematics(rb, x0)
    mg = k["mass"] * 9.81
    # the dynamics are affine in tau: a(tau) = a(0) + da_dtau tau.  Zero-acceleration joint torques by least squares.
    o0 = oracle.cdyn(rb, cfg, x0, np.zeros(28), (1, 1), check=False)
    u, *_ = np.linalg.lstsq(o0["da_dtau"][:, 6:], -o0["a"], rcond=None)
    tau = np.concatenate([np.zeros(6), u])
    o = oracle.cdyn(rb, cfg, x0, tau, (1, 1), check=False)
    assert np.abs(o["a"]).max() < 1e-6
    assert abs(o["lam"][2] + o["lam"][8] - mg) < 1e-2 * mg
    assert o["lam"][2] > 0.3 * mg and o["lam"][8] > 0.3 * mg


def test_cone_rows_match_reference_convention(oracle):
    """Rows 0-8 of the wrench cone agree with QP_utils.py:337-347 (written there as C w >= l, i.e. opposite sign)."""
    mu, L, W = 0.8, 0.1, 0.075
    A = oracle.cone_matrix(mu, L, W)
    qp = np.array([[0, 0, 1, 0, 0, 0], [-1, 0, mu, 0, 0, 0], [1, 0, mu, 0, 0, 0], [0, -1, mu, 0, 0, 0], [0, 1, mu, 0, 0, 0],
                   [0, 0, W, -1, 0, 0], [0, 0, W, 1, 0, 0], [0, 0, L, 0, -1, 0], [0, 0, L, 0, 1, 0]], float)
    assert A.shape == (17, 6)
    got = {tuple(np.round(-r, 12)) for r in A[:9]}
    assert got == {tuple(np.round(r, 12)) for r in qp}
    # a centred vertical force is strictly inside, a pure tangential one is outside
    assert (A @ np.array([0, 0, 100.0, 0, 0, 0]) < 0).all()
    assert (A @ np.array([100.0, 0, 1.0, 0, 0, 0]) > 0).any()


@pytest.mark.parametrize("cs", [(1, 1), (1, 0), (0, 1)])
def test_full_knot_jacobians_match_finite_differences(oracle, full, cs):
    rng = np.random.default_rng(5)
    rb, cfg = full["robot"], full["cfg"]
    lf, rf = full["lf"], full["rf"]
    fr = np.array([0, 0, 400.0, 0, 0, 0])
    kn = problems.full_knot(list(cs), lf, rf, fr, fr)
    x = rand_state(oracle, full["x0"][0], rng, 0.05)
    u = rng.normal(size=22) * 10
    xn = rand_state(oracle, x, rng, 0.02)
    o = oracle.eval_knot(rb, cfg, kn, x, u, xn)
    eps = 1e-6
    num_A, num_C, num_g = np.zeros((56, 78)), np.zeros((78, 78)), np.zeros(78)
    for j in range(78):
        def at(s):
            if j < 56:
                d = np.zeros(56); d[j] = s
                return oracle.eval_knot(rb, cfg, kn, oracle.integrate(x, d), u, xn, derivs=False)
            uu = u.copy(); uu[j - 56] += s
            return oracle.eval_knot(rb, cfg, kn, x, uu, xn, derivs=False)
        p, m = at(eps), at(-eps)
        num_A[:, j] = (p["gap"] - m["gap"]) / (2 * eps)
        num_C[:, j] = (p["h"] - m["h"]) / (2 * eps)
        num_g[j] = (p["cost"] - m["cost"]) / (2 * eps)
    AB = np.hstack([o["A"], o["B"]])
    CD = np.hstack([o["Cx"], o["Cu"]])
    g = np.concatenate([o["lx"], o["lu"]])
    assert np.abs(AB - num_A).max() < 2e-5 * max(1, np.abs(AB).max())
    assert np.abs(CD - num_C).max() < 2e-5 * max(1, np.abs(CD).max())
    assert np.abs(g - num_g).max() < 1e-5 * max(1, np.abs(g).max())
    # shooting-gap Jacobian wrt x_{k+1}
    E = -np.eye(56); E[:6, :6] = o["E6"]
    numE = np.zeros((56, 56))
    for j in range(56):
        d = np.zeros(56); d[j] = eps
        p = oracle.eval_knot(rb, cfg, kn, x, u, oracle.integrate(xn, d), derivs=False)["gap"]
        m = oracle.eval_knot(rb, cfg, kn, x, u, oracle.integrate(xn, -d), derivs=False)["gap"]
        numE[:, j] = (p - m) / (2 * eps)
    assert np.abs(E - numE).max() < 1e-6


def test_centroidal_knot_jacobians_match_finite_differences(oracle):
    p = problems.cent_standing_problem(T=3)
    rng = np.random.default_rng(6)
    rb, cfg = p["robot"], p["cfg"]
    for cs in [(1, 1), (1, 0)]:
        kn = problems.cent_knot(cs, p["lf"], p["rf"], p["us"][0, 0])
        x = p["x0"][0] + rng.normal(size=9) * 0.1
        u = p["us"][0, 0] + rng.normal(size=12) * 5
        o = oracle.eval_knot(rb, cfg, kn, x, u, x)
        eps = 1e-6
        z0 = np.concatenate([x, u])
        numA, numg = np.zeros((9, 21)), np.zeros(21)
        for j in range(21):
            zp, zm = z0.copy(), z0.copy()
            zp[j] += eps; zm[j] -= eps
            a = oracle.eval_knot(rb, cfg, kn, zp[:9], zp[9:], x, derivs=False)
            b = oracle.eval_knot(rb, cfg, kn, zm[:9], zm[9:], x, derivs=False)
            numA[:, j] = (a["gap"] - b["gap"]) / (2 * eps)
            numg[j] = (a["cost"] - b["cost"]) / (2 * eps)
        assert np.abs(np.hstack([o["A"], o["B"]]) - numA).max() < 1e-7
        assert np.abs(np.concatenate([o["lx"], o["lu"]]) - numg).max() < 1e-4 * max(1, np.abs(numg).max())
        assert (o["ctype"][:17] == 1).all() and ((o["ctype"][17:] == 1).all() == bool(cs[1]))


@pytest.mark.parametrize("cs", [(1, 1), (1, 0), (0, 1)])
def test_kinodynamic_knot_jacobians(oracle, cs):
    """Kinodynamic stage (kinodynamic_talos.py:107-173): dynamics Jacobians vs forward-mode AD of the templated value code,
    constraint / cost Jacobians vs central differences."""
    import ctypes as C

    p = problems.kino_standing_problem(T=3)
    rb, cfg = p["robot"], p["cfg"]
    rng = np.random.default_rng(8)
    kn = problems.kino_knot(list(cs), p["lf"], p["rf"], p["us"][0, 0])
    x = oracle.integrate(p["x0"][0], rng.normal(size=56) * 0.1)
    u = p["us"][0, 0] + rng.normal(size=34) * 5
    oracle.lib().orc_check_kino.restype = C.c_double
    err = oracle.lib().orc_check_kino(C.byref(rb), C.byref(cfg), C.byref(kn), oracle._p(np.ascontiguousarray(x)), oracle._p(np.ascontiguousarray(u)))
    assert err < 1e-8
    o = oracle.eval_knot(rb, cfg, kn, x, u, x)
    eps = 1e-6
    nC, ng = np.zeros((68, 90)), np.zeros(90)
    for j in range(90):
        def at(s):
            if j < 56:
                d = np.zeros(56); d[j] = s
                return oracle.eval_knot(rb, cfg, kn, oracle.integrate(x, d), u, x, derivs=False)
            uu = u.copy(); uu[j - 56] += s
            return oracle.eval_knot(rb, cfg, kn, x, uu, x, derivs=False)
        a, b = at(eps), at(-eps)
        nC[:, j] = (a["h"] - b["h"]) / (2 * eps)
        ng[j] = (a["cost"] - b["cost"]) / (2 * eps)
    assert np.abs(np.hstack([o["Cx"], o["Cu"]]) - nC).max() < 1e-6
    g = np.concatenate([o["lx"], o["lu"]])
    assert np.abs(g - ng).max() < 1e-6 * max(1, np.abs(g).max())
    n_active = sum(cs)
    assert (o["ctype"] == 1).sum() == 17 * n_active and (o["ctype"] == 0).sum() == 6 * n_active and (o["ctype"] == 2).sum() == 22


def test_cold_solves_converge(oracle):
    """centroidal cold solve converges to TOL = 1e-5 (centroidal_talos.py:265); so does the standing full model."""
    r = oracle.solve(problems.cent_standing_problem())
    i = r["info"][0]
    assert i.conv == 1 and i.num_iters <= 10 and max(i.prim_infeas, i.dual_infeas) <= 1e-5
    pf = problems.full_standing_problem()
    r = oracle.solve(pf, knot_threads=4)
    i = r["info"][0]
    assert i.conv == 1 and i.num_iters <= 15 and max(i.prim_infeas, i.dual_infeas) <= 1e-5
    mg = pf["mass"] * 9.81
    assert abs(r["stage0"][0, 56 + 2] + r["stage0"][0, 56 + 8] - mg) < 1e-3 * mg  # contact forces carry the weight
    pk = problems.kino_standing_problem()
    r = oracle.solve(pk, knot_threads=4)
    i = r["info"][0]
    assert i.conv == 1 and i.num_iters <= 15 and max(i.prim_infeas, i.dual_infeas) <= 1e-5
    assert abs(r["us"][0, 0, 2] + r["us"][0, 0, 8] - mg) < 1e-2 * mg  # kinodynamic wrenches carry the weight too


def test_golden_fixtures_reproduce(oracle):
    """The committed golden vectors are what the oracle computes today (guards against silent oracle drift)."""
    import golden_util

    for name in ["ref_flat_full.npz", "ref_flat_kino.npz", "ref_flat_cent.npz"]:
        prob, z = golden_util.load(name)
        r = oracle.solve(prob, knot_threads=4)
        assert [i.num_iters for i in r["info"]] == list(z["sol_num_iters"])
        assert np.abs(r["xs"] - z["sol_xs"]).max() <= 1e-9 * np.abs(z["sol_xs"]).max()
        assert np.abs(r["us"] - z["sol_us"]).max() <= 1e-8 * max(1.0, np.abs(z["sol_us"]).max())
