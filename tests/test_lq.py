"""The proximal Riccati recursion of the oracle against a dense solve of the whole dual-regularised KKT system
(SURVEY 7.3 'Riccati solution satisfies the full KKT system')."""
import numpy as np
import pytest


def random_lq(rng, n, m, nc, T, nact):
    nz = n + m
    H = np.zeros((T, nz, nz)); g = rng.normal(size=(T, nz))
    AB = np.zeros((T, n, nz)); f = rng.normal(size=(T, n)) * 0.1
    CD = np.zeros((T, nc, nz)); d = np.zeros((T, nc))
    E6 = np.zeros((T, 36))
    for k in range(T):
        R = rng.normal(size=(nz, nz)); H[k] = R @ R.T / nz + 0.1 * np.eye(nz)
        AB[k, :, :n] = np.eye(n) + 0.1 * rng.normal(size=(n, n)); AB[k, :, n:] = rng.normal(size=(n, m)) * 0.3
        rows = rng.choice(nc, size=min(nact, nc), replace=False)
        CD[k, rows] = rng.normal(size=(len(rows), nz)); d[k] = rng.normal(size=nc) * 0.01
        E = -np.eye(6) + 0.05 * rng.normal(size=(6, 6)); E6[k] = E.ravel()
    R = rng.normal(size=(n, n)); HT = R @ R.T / n + 0.1 * np.eye(n); gT = rng.normal(size=n)
    return H, g, AB, f, CD, d, E6, HT, gT


def dense_kkt_solution(n, m, nc, T, mu_d, mu, H, g, AB, f, CD, d, E6, HT, gT, CT, dT):
    """Unknowns: dx_1..dx_T, du_0..du_{T-1}, dv_0..dv_{T-1} (+ dv_T), dlam_1..dlam_T; dx_0 = 0."""
    nct = 0 if CT is None else CT.shape[0]
    nz = n + m
    ox = lambda k: (k - 1) * n
    ou = lambda k: T * n + k * m
    ov = lambda k: T * n + T * m + k * nc
    ovT = T * n + T * m + T * nc
    ol = lambda k: ovT + nct + (k - 1) * n
    N = ovT + nct + T * n
    K = np.zeros((N, N)); r = np.zeros(N)
    Efull = []
    for k in range(T):
        E = -np.eye(n)
        if n >= 6 and E6 is not None:
            E[:6, :6] = E6[k].reshape(6, 6)
        Efull.append(E)
    for k in range(T):
        A, B = AB[k][:, :n], AB[k][:, n:]
        C, D = CD[k][:, :n], CD[k][:, n:]
        Q, S, R = H[k][:n, :n], H[k][:n, n:], H[k][n:, n:]
        # stationarity wrt du_k
        K[ou(k):ou(k) + m, ou(k):ou(k) + m] += R
        if k > 0: K[ou(k):ou(k) + m, ox(k):ox(k) + n] += S.T
        K[ou(k):ou(k) + m, ol(k + 1):ol(k + 1) + n] += B.T
        K[ou(k):ou(k) + m, ov(k):ov(k) + nc] += D.T
        r[ou(k):ou(k) + m] = -g[k][n:]
        # stationarity wrt dx_k (k >= 1)
        if k > 0:
            K[ox(k):ox(k) + n, ox(k):ox(k) + n] += Q
            K[ox(k):ox(k) + n, ou(k):ou(k) + m] += S
            K[ox(k):ox(k) + n, ol(k + 1):ol(k + 1) + n] += A.T
            K[ox(k):ox(k) + n, ov(k):ov(k) + nc] += C.T
            K[ox(k):ox(k) + n, ol(k):ol(k) + n] += Efull[k - 1].T
            r[ox(k):ox(k) + n] = -g[k][:n]
        # dynamics row k: A dx + B du + E dx' - mu_d dlam' = -f
        if k > 0: K[ol(k + 1):ol(k + 1) + n, ox(k):ox(k) + n] += A
        K[ol(k + 1):ol(k + 1) + n, ou(k):ou(k) + m] += B
        K[ol(k + 1):ol(k + 1) + n, ox(k + 1):ox(k + 1) + n] += Efull[k]
        K[ol(k + 1):ol(k + 1) + n, ol(k + 1):ol(k + 1) + n] += -mu_d * np.eye(n)
        r[ol(k + 1):ol(k + 1) + n] = -f[k]
        # constraint rows
        if k > 0: K[ov(k):ov(k) + nc, ox(k):ox(k) + n] += C
        K[ov(k):ov(k) + nc, ou(k):ou(k) + m] += D
        K[ov(k):ov(k) + nc, ov(k):ov(k) + nc] += -mu * np.eye(nc)
        r[ov(k):ov(k) + nc] = -d[k]
    # terminal
    K[ox(T):ox(T) + n, ox(T):ox(T) + n] += HT
    K[ox(T):ox(T) + n, ol(T):ol(T) + n] += Efull[T - 1].T
    r[ox(T):ox(T) + n] = -gT
    if nct:
        K[ox(T):ox(T) + n, ovT:ovT + nct] += CT.T
        K[ovT:ovT + nct, ox(T):ox(T) + n] += CT
        K[ovT:ovT + nct, ovT:ovT + nct] += -mu * np.eye(nct)
        r[ovT:ovT + nct] = -dT
    sol = np.linalg.solve(K, r)
    dxs = np.vstack([np.zeros(n), sol[:T * n].reshape(T, n)])
    dus = sol[T * n:T * n + T * m].reshape(T, m)
    dvs = sol[ov(0):ov(0) + T * nc].reshape(T, nc)
    dls = sol[ol(1):ol(1) + T * n].reshape(T, n)
    return dxs, dus, dvs, dls, sol[ovT:ovT + nct]


@pytest.mark.parametrize("n,m,nc,T,nact,mu", [(9, 12, 34, 6, 5, 1e-2), (9, 12, 34, 5, 34, 1e-6), (14, 5, 8, 7, 3, 1e-8), (56, 22, 78, 4, 10, 1e-8)])
def test_riccati_matches_dense_kkt(oracle, n, m, nc, T, nact, mu):
    rng = np.random.default_rng(n + T)
    H, g, AB, f, CD, d, E6, HT, gT = random_lq(rng, n, m, nc, T, nact)
    CT, dT = (rng.normal(size=(3, n)), rng.normal(size=3) * 0.01) if n >= 14 else (None, None)
    o = oracle.riccati(n, m, nc, T, mu, mu, H, g, AB, f, CD, d, E6 if n >= 6 else None, HT, gT, CT, dT)
    dxs, dus, dvs, dls, dvT = dense_kkt_solution(n, m, nc, T, mu, mu, H, g, AB, f, CD, d, E6 if n >= 6 else None, HT, gT, CT, dT)
    sc = max(1.0, np.abs(dxs).max(), np.abs(dus).max())
    assert np.abs(o["dxs"] - dxs).max() < 1e-7 * sc
    assert np.abs(o["dus"] - dus).max() < 1e-7 * sc
    assert np.abs(o["dlams"][1:] - dls).max() < 1e-6 * max(1.0, np.abs(dls).max())
    assert np.abs(o["dvs"][:T] - dvs).max() < 1e-6 * max(1.0, np.abs(dvs).max())
    if CT is not None:
        assert np.abs(o["dvs"][T, :3] - dvT).max() < 1e-6 * max(1.0, np.abs(dvT).max())
