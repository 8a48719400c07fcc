"""Flat problem descriptors of the three reference OCPs (the constants of SURVEY App. C).

These builders produce exactly the `Config` / `Knot` / `Term` blocks that the aligator-compatible shim
(`mpc_benchmark_b200.api`) flattens out of the object graph the reference scripts build
(fulldynamic_talos.py:100-245, kinodynamic_talos.py:72-180, centroidal_talos.py:185-261); bench.py and
the tests use them directly to synthesise the BASELINE.json workloads.
"""
import numpy as np

from . import _abi
from .kinematics import center_of_mass, foot_placements
from .talos_like import half_sitting, talos_like_robot

GRAVITY = 9.81


def _set(arr, vals):
    vals = list(np.asarray(vals, float).reshape(-1))
    arr[: len(vals)] = vals


def base_setup(robot=None):
    rb = robot if robot is not None else talos_like_robot()
    q0 = half_sitting()
    x0 = np.concatenate([q0, np.zeros(_abi.NV)])
    lf, rf = foot_placements(rb, q0)
    com0, mass = center_of_mass(rb, q0)
    return rb, q0, x0, lf, rf, com0, mass


# ---------------------------------------------------------------- full dynamics
W_X_FULL = np.array(
    [0, 0, 0, 100, 100, 100] + [0.1] * 6 + [0.1] * 6 + [10, 10] + [1] * 4 + [1] * 4
    + [1] * 6 + [0.1, 0.1, 0.1, 0.1, 0.01, 0.01] * 2 + [10, 10] + [1] * 4 + [1] * 4,
    dtype=float,
)  # fulldynamic_talos.py:121-135


def full_config(rb, x0, lf, rf, T=100, dt=0.01, tol=1e-5, mu_init=1e-8, max_iters=100):
    c = _abi.Config()
    c.kind, c.T, c.dt = _abi.KIND_FULL, T, dt
    _set(c.x_ref, x0)
    _set(c.wx, W_X_FULL)
    _set(c.wu, [1e-4] * 22)  # full:136
    _set(c.w_cent, [0, 0, 10, 0, 0, 10])  # full:141-143
    _set(c.w_force, [1e-4] * 6)  # full:149-151
    _set(c.wx_term, W_X_FULL)  # full:235
    _set(c.w_cent_term, [0, 0, 10, 0, 0, 10])  # full:243
    _set(c.w_foot_term, [2000.0] * 6)  # full:244-245
    c.mu_fric, c.foot_L, c.foot_W = 0.8, 0.1, 0.075  # full:69-71
    _set(c.kp, [0, 0, 10, 0, 0, 0])  # full:93
    _set(c.kd, [50] * 6)  # full:94
    _set(c.contact_place[0], lf)  # full:83
    _set(c.contact_place[1], rf)
    c.mu_contact = 1e-10  # ProximalSettings(1e-9, 1e-10, 1), full:77
    c.tol, c.mu_init, c.max_iters, c.force_initial_condition = tol, mu_init, max_iters, 1
    return c


def full_knot(cs, lf_ref, rf_ref, f_ref_l, f_ref_r, w_lfrf=2000.0):
    """createStage(cs, ...) of fulldynamic_talos.py:153-232."""
    k = _abi.Knot()
    cl, cr = bool(cs[0]), bool(cs[1])
    k.cs[0], k.cs[1] = float(cl), float(cr)
    # force costs (full:187-201): both / left only / right only / none
    k.fcost[0] = 1.0 if cl else 0.0
    k.fcost[1] = 1.0 if cr else 0.0
    # pose-cost switches (full:179-182): w_RF on when the LEFT foot is in contact and vice versa
    _set(k.w_rf, [w_lfrf if cl else 0.0] * 6)
    _set(k.w_lf, [w_lfrf if cr else 0.0] * 6)
    _set(k.lf_ref, lf_ref)
    _set(k.rf_ref, rf_ref)
    _set(k.f_ref, np.concatenate([f_ref_l, f_ref_r]))
    return k


def full_contact_phases(T_ds=30, T_ss=80, total_steps=3, nsteps=100):
    """fulldynamic_talos.py:255-266."""
    ph = [[True, True]] * T_ds
    for _ in range(total_steps):
        ph += [[True, False]] * T_ss + [[True, True]] * T_ds + [[False, True]] * T_ss + [[True, True]] * T_ds
    ph += [[True, False]] * T_ss + [[True, True]] * T_ds
    ph += [[True, True]] * nsteps * 2
    return ph


def make_term(lf_ref, rf_ref, com_ref=None):
    t = _abi.Term()
    _set(t.lf_ref, lf_ref)
    _set(t.rf_ref, rf_ref)
    if com_ref is not None:
        _set(t.com_ref, com_ref)
        t.has_com_cstr = 1.0
    return t


def full_standing_problem(batch=1, T=100, robot=None, **kw):
    """Cold-solve problem of fulldynamic_talos.py:371-397: all-double-support standing horizon."""
    rb, q0, x0, lf, rf, com0, mass = base_setup(robot)
    cfg = full_config(rb, x0, lf, rf, T=T, **kw)
    f_half = mass * GRAVITY / 2.0
    fr = np.array([0, 0, f_half, 0, 0, 0.0])
    knot = full_knot([True, True], lf, rf, fr, fr)
    knots = (_abi.Knot * (batch * T))(*([knot] * (batch * T)))
    terms = (_abi.Term * batch)(*([make_term(lf, rf)] * batch))
    x0s = np.tile(x0, (batch, 1))
    xs = np.tile(x0, (batch, T + 1, 1))
    us = np.zeros((batch, T, 22))
    return dict(robot=rb, cfg=cfg, knots=knots, terms=terms, x0=x0s, xs=xs, us=us, lf=lf, rf=rf, com0=com0, mass=mass)


# ---------------------------------------------------------------- centroidal
def cent_config(rb, com0, mass, T=100, dt=0.01, tol=1e-5, mu_init=1e-8, max_iters=100):
    c = _abi.Config()
    c.kind, c.T, c.dt = _abi.KIND_CENT, T, dt
    _set(c.wu, [0.001] * 3 + [0.1] * 3 + [0.001] * 3 + [0.1] * 3)  # cent:193-200
    _set(c.w_com, [0, 0, 0])  # cent:187
    _set(c.w_linmom, [0.01, 0.01, 100])  # cent:188
    _set(c.w_linacc, [0.01] * 3)  # cent:189
    _set(c.w_angmom, [0.1, 0.1, 1000])  # cent:190
    _set(c.w_angacc, [0.01] * 3)  # cent:191
    _set(c.com_ref, com0)
    c.mass = mass
    c.mu_fric, c.foot_L, c.foot_W = 0.8, 0.1, 0.075  # cent:68-70
    c.tol, c.mu_init, c.max_iters, c.force_initial_condition = tol, mu_init, max_iters, 1
    return c


def cent_knot(cs, lf, rf, u_ref):
    """createStage(contact_state, LF_pose, RF_pose, ur) of centroidal_talos.py:208-247."""
    k = _abi.Knot()
    k.cs[0], k.cs[1] = float(bool(cs[0])), float(bool(cs[1]))
    _set(k.cpos, np.concatenate([np.asarray(lf)[9:12], np.asarray(rf)[9:12]]))
    _set(k.u_ref, u_ref)
    return k


def cent_standing_problem(batch=1, T=100, robot=None, **kw):
    """Cold-solve problem of centroidal_talos.py:252-288."""
    rb, q0, x0mb, lf, rf, com0, mass = base_setup(robot)
    cfg = cent_config(rb, com0, mass, T=T, **kw)
    u0 = np.zeros(12)
    u0[2] = u0[8] = mass * GRAVITY / 2.0  # cent:73-74
    # urefs[0] (cent:132-140 with i = 0, j = 0): un[2] = f_half, un[8] = f_half
    knot = cent_knot([True, True], lf, rf, u0)
    knots = (_abi.Knot * (batch * T))(*([knot] * (batch * T)))
    terms = (_abi.Term * batch)(*([make_term(lf, rf)] * batch))
    x0 = np.zeros(9)
    x0[:3] = com0
    return dict(robot=rb, cfg=cfg, knots=knots, terms=terms, x0=np.tile(x0, (batch, 1)), xs=np.tile(x0, (batch, T + 1, 1)),
                us=np.tile(u0, (batch, T, 1)), lf=lf, rf=rf, com0=com0, mass=mass)
