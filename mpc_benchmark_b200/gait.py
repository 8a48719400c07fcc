"""Gait / swing-foot reference generation of the reference MPC loops, restated on numpy.

What the three scripts do every tick before `solver.run` — `update_timings` (talos_utils.py:350-373), `footTrajectory.
updateTrajectory` (talos_utils.py:187-327: yaw-aligned footstep placement, degree-8 Bezier swing curve, geodesic rotation
interpolation — ndcurves is not installable offline, its two curves are restated here), the per-stage `setReference` /
`contact_poses` writes and `replaceStageCircular` (fulldynamic_talos.py:444-497; kinodynamic_talos.py:362-395;
centroidal_talos.py:354-384,459) — becomes `GaitPlan.tick`, which returns the flat per-knot parameter blocks of the horizon the
solver sees at that tick.  Used to synthesise the BASELINE.json walking / stairs workloads and by the closed-loop driver.
Placements are 12-vectors (rotation row-major, translation).  Host-side input generation only: no solver arithmetic here.
"""
import numpy as np

from . import _abi

GRAVITY = 9.81

# (T_ds, T_ss, full cycles, extra half cycle) of fulldynamic_talos.py:248-266, kinodynamic_talos.py:183-198, centroidal_talos.py:100-116
GAITS = {_abi.KIND_FULL: (30, 80, 3, True), _abi.KIND_KINO: (20, 80, 3, False), _abi.KIND_CENT: (20, 80, 1, False)}
X_FORWARD = {_abi.KIND_FULL: 0.0, _abi.KIND_KINO: 0.3, _abi.KIND_CENT: 0.2}  # full:352, kino:257, cent:175


def contact_phases(kind, nsteps=100, mirror=False):
    """The script's `contact_phases` list ([left, right] in contact).  `mirror` swaps the feet (first swing with the left foot)."""
    T_ds, T_ss, cycles, half = GAITS[kind]
    a, b = ([False, True], [True, False]) if mirror else ([True, False], [False, True])
    ph = [[True, True]] * T_ds
    for _ in range(cycles):
        ph += [a] * T_ss + [[True, True]] * T_ds + [b] * T_ss + [[True, True]] * T_ds
    if half:
        ph += [a] * T_ss + [[True, True]] * T_ds
    ph += [[True, True]] * nsteps * 2
    return [list(p) for p in ph]


def countdown_lists(phases, nsteps=100):
    """takeoff_RFs, takeoff_LFs, land_RFs, land_LFs seeded with `phase index + nsteps` (fulldynamic_talos.py:268-280)."""
    to_rf, to_lf, la_rf, la_lf = [], [], [], []
    for i in range(1, len(phases)):
        p, q = phases[i], phases[i - 1]
        if p == [True, False] and q == [True, True]:
            to_rf.append(i + nsteps)
        elif p == [False, True] and q == [True, True]:
            to_lf.append(i + nsteps)
        elif p == [True, True] and q == [True, False]:
            la_rf.append(i + nsteps)
        elif p == [True, True] and q == [False, True]:
            la_lf.append(i + nsteps)
    return to_rf, to_lf, la_rf, la_lf


def _scan(lst):
    for i in range(len(lst)):
        lst[i] -= 1
    if lst and lst[0] == -1:
        lst.pop(0)


def update_timings(la_lf, la_rf, to_lf, to_rf):
    """talos_utils.py:356-373: decrement every countdown, drop entries that reach -1, return the heads (or -1)."""
    for lst in (la_lf, la_rf, to_lf, to_rf):
        _scan(lst)
    head = lambda lst: lst[0] if lst else -1  # noqa: E731
    return head(to_rf), head(to_lf), head(la_rf), head(la_lf)


def force_ramp_refs(kind, mass, nu, nsteps=100):
    """`urefs` of kinodynamic_talos.py:200-237 / centroidal_talos.py:132-169: vertical-force references (entries 2 and 8 of u)
    ramping between the feet through the double-support phases.  The two scripts differ in the last ramp (which foot carries the
    full weight at the end of the last swing)."""
    T_ds, T_ss, cycles, _ = GAITS[kind]
    f_full, f_half = mass * GRAVITY, mass * GRAVITY / 2.0
    out = []

    def un(a, b):
        u = np.zeros(nu)
        u[2], u[8] = a, b
        return u

    for i in range(cycles):
        for j in range(T_ds):
            if i == 0:
                out.append(un(f_full * j / T_ds + f_half * (T_ds - j) / T_ds, f_half * (T_ds - j) / T_ds))
            else:
                out.append(un(f_full * (j + 1) / T_ds, f_full * (T_ds - j) / T_ds))
        out += [un(f_full, 0.0) for _ in range(T_ss)]
        for j in range(T_ds):
            out.append(un(f_full * (T_ds - j) / T_ds, f_full * (j + 1) / T_ds))
        out += [un(0.0, f_full) for _ in range(T_ss)]
    for j in range(T_ds):
        ramp_up, ramp_down = f_half * (j + 1) / float(T_ds), f_full * (T_ds - j) / float(T_ds) + f_half * j / float(T_ds)
        out.append(un(ramp_up, ramp_down) if kind == _abi.KIND_KINO else un(ramp_down, ramp_up))  # kino:226-230 vs cent:158-162
    out += [un(f_half, f_half) for _ in range(nsteps * 2)]
    return out


# ------------------------------------------------------------------ SE3 helpers on 12-vectors
def _R(p):
    return np.asarray(p, float)[:9].reshape(3, 3)


def _t(p):
    return np.asarray(p, float)[9:12]


def _pose(R, t):
    return np.concatenate([np.asarray(R, float).reshape(9), np.asarray(t, float)])


def yaw_rotation(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def extract_yaw(R):
    return np.arctan2(R[1, 0], R[0, 0])


def _log3(R):
    s = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    ct = 0.5 * (np.trace(R) - 1.0)
    sn = np.linalg.norm(s)
    if sn < 1e-12:
        return s
    return s * (np.arctan2(sn, ct) / sn)


def _exp3(w):
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W


def bezier_swing(start, final, apex, s):
    """Pose on the swing curve of talos_utils.py:281-296 at s in [0, 1]: degree-8 Bezier through 4 x start, the point
    0.75 start + 0.25 final lifted by `apex`, 4 x final (zero velocity / acceleration / jerk at both ends); rotation on the
    SO(3) geodesic from start to final (ndcurves.SE3Curve)."""
    from math import comb

    b = np.array([comb(8, i) * (1 - s) ** (8 - i) * s ** i for i in range(9)])
    mid = 0.75 * _t(start) + 0.25 * _t(final)
    mid[2] += apex
    t = b[:4].sum() * _t(start) + b[4] * mid + b[5:].sum() * _t(final)
    R0, R1 = _R(start), _R(final)
    R = R0 @ _exp3(s * _log3(R0.T @ R1))
    return _pose(R, t)


class FootTrajectory:
    """talos_utils.py:187-327 (`footTrajectory`), same constructor arguments and methods."""

    def __init__(self, start_pose_left, start_pose_right, T_ss, T_ds, nsteps, swing_apex, x_forward, y_forward, foot_angle, y_gap, z_height):
        self.translationRight = np.array([x_forward, -y_gap - y_forward, z_height])
        self.translationLeft = np.array([x_forward, y_gap, z_height])
        self.rotationDiff = yaw_rotation(foot_angle)
        self.start_pose_left, self.start_pose_right = np.array(start_pose_left, float), np.array(start_pose_right, float)
        self.final_pose_left, self.final_pose_right = self.start_pose_left.copy(), self.start_pose_right.copy()
        self.T_ds, self.T_ss, self.nsteps, self.swing_apex = T_ds, T_ss, nsteps, swing_apex

    def updateForward(self, x_f_left, x_f_right, y_gap, y_forward, z_height_left, z_height_right, swing_apex):
        self.translationRight = np.array([x_f_right, -y_gap - y_forward, z_height_right])
        self.translationLeft = np.array([x_f_left, y_gap, z_height_left])
        self.swing_apex = swing_apex

    def update_state(self, takeoff_RF, takeoff_LF, land_RF, land_LF, LF_pose, RF_pose):
        """The start / final pose bookkeeping of updateTrajectory (talos_utils.py:211-243) without sampling the curves."""
        LF_pose, RF_pose = np.asarray(LF_pose, float), np.asarray(RF_pose, float)
        if land_LF < 0:
            self.start_pose_left, self.final_pose_left = LF_pose.copy(), LF_pose.copy()
        if land_RF < 0:
            self.start_pose_right, self.final_pose_right = RF_pose.copy(), RF_pose.copy()
        if 0 <= takeoff_RF < self.T_ds:
            self.start_pose_right = RF_pose.copy()
            t = _t(LF_pose) + yaw_rotation(extract_yaw(_R(LF_pose))) @ self.translationRight
            self.final_pose_right = _pose(self.rotationDiff @ _R(LF_pose), t)
            self.start_pose_left = LF_pose.copy()
            yr = extract_yaw(_R(self.final_pose_right))
            self.final_pose_left = _pose(_R(self.final_pose_right), _t(self.final_pose_right) + yaw_rotation(yr) @ self.translationLeft)
        if 0 <= takeoff_LF < self.T_ds:
            self.start_pose_left = LF_pose.copy()
            t = _t(RF_pose) + yaw_rotation(extract_yaw(_R(RF_pose))) @ self.translationLeft
            self.final_pose_left = _pose(_R(RF_pose), t)
            self.start_pose_right = RF_pose.copy()
            yl = extract_yaw(_R(self.final_pose_left))
            t = _t(self.final_pose_left) + yaw_rotation(yl) @ self.translationRight
            self.final_pose_right = _pose(self.rotationDiff @ _R(self.final_pose_left), t)

    def foot_trajectory(self, T, time_to_land, initial_pose, final_pose, T_ss):
        """talos_utils.py:298-317: horizon index j has countdown t = time_to_land - j."""
        out = []
        for t in range(time_to_land, time_to_land - T, -1):
            if t <= 0:
                out.append(final_pose.copy())
            elif t > T_ss:
                out.append(initial_pose.copy())
            else:
                out.append(bezier_swing(initial_pose, final_pose, self.swing_apex, float(T_ss - t) / float(T_ss)))
        return out

    def updateTrajectory(self, takeoff_RF, takeoff_LF, land_RF, land_LF, LF_pose, RF_pose):
        self.update_state(takeoff_RF, takeoff_LF, land_RF, land_LF, LF_pose, RF_pose)
        n = self.nsteps
        LF = self.foot_trajectory(n, land_LF, self.start_pose_left, self.final_pose_left, self.T_ss) if land_LF > -1 \
            else [self.start_pose_left.copy() for _ in range(n)]
        RF = self.foot_trajectory(n, land_RF, self.start_pose_right, self.final_pose_right, self.T_ss) if land_RF > -1 \
            else [self.start_pose_right.copy() for _ in range(n)]
        return LF, RF


class GaitPlan:
    """The per-tick reference bookkeeping of one robot's MPC loop.

    `tick(lf_pose, rf_pose)` advances the countdowns by one tick and returns what the loop hands to the solver at that tick:
    `LF_refs`, `RF_refs` (nsteps placements), the contact phase entering the horizon (`contact_phases[t]`, the stage appended by
    `replaceStageCircular`) and `com_final` (fulldynamic_talos.py:499-500).  `horizon_refs()` applies the reference's write-then-
    rotate order: at solve time stage j carries ref j + 1 and the newly appended stage its construction default."""

    def __init__(self, kind, lf0, rf0, com0, nsteps=100, swing_apex=0.15, x_forward=None, y_forward=0.0, foot_yaw=0.0, y_gap=0.18,
                 z_height=0.0, mirror=False, keep_forward=False):
        self.kind, self.nsteps, self.com0 = kind, nsteps, np.asarray(com0, float)
        T_ds, T_ss, _, _ = GAITS[kind]
        self.phases = contact_phases(kind, nsteps, mirror)
        self.to_rf, self.to_lf, self.la_rf, self.la_lf = countdown_lists(self.phases, nsteps)
        xf = X_FORWARD[kind] if x_forward is None else x_forward
        self.ft = FootTrajectory(lf0, rf0, T_ss, T_ds, nsteps, swing_apex, xf, y_forward, foot_yaw, y_gap, z_height)
        self.lf0, self.rf0 = np.array(lf0, float), np.array(rf0, float)
        self.y_gap, self.y_forward, self.apex, self.keep_forward = y_gap, y_forward, swing_apex, keep_forward
        self.t = 0
        # the horizon: construction-default stages (phase 0, initial placements) that the stream of stages_full[t] replaces
        self.h_phase = [list(self.phases[0]) for _ in range(nsteps)]
        self.h_index = [0] * nsteps  # index into contact_phases / urefs of the stage in each slot
        self.h_lf = [self.lf0.copy() for _ in range(nsteps)]
        self.h_rf = [self.rf0.copy() for _ in range(nsteps)]

    def tick(self, lf_pose=None, rf_pose=None, sample=True):
        lf_pose = self.lf0 if lf_pose is None else lf_pose
        rf_pose = self.rf0 if rf_pose is None else rf_pose
        to_rf, to_lf, la_rf, la_lf = update_timings(self.la_lf, self.la_rf, self.to_lf, self.to_rf)
        if not self.keep_forward:
            # the scripts zero the forward step once no further landing is pending (full:448-449 also lowers the last left step by 1 cm)
            if self.kind == _abi.KIND_FULL and la_lf == -1:
                self.ft.updateForward(0, 0, self.y_gap, self.y_forward, -0.01, 0, self.apex)
            elif self.kind == _abi.KIND_KINO and la_rf == -1 and to_rf == -1:
                self.ft.updateForward(0, 0, self.y_gap, self.y_forward, 0, 0, self.apex)
            elif self.kind == _abi.KIND_CENT and la_rf == -1:
                self.ft.updateForward(0, 0, self.y_gap, self.y_forward, -0.01, 0, self.apex)
        t = min(self.t, len(self.phases) - 1)
        if sample:
            LF, RF = self.ft.updateTrajectory(to_rf, to_lf, la_rf, la_lf, lf_pose, rf_pose)
            # write-then-rotate (full:461-463 then :496): refs go to slots 0..n-1, then slot 0 is dropped and the new stage appended
            self.h_lf = LF[1:] + [self.lf0.copy()]
            self.h_rf = RF[1:] + [self.rf0.copy()]
            com_final = self.com0.copy()
            com_final[:2] = 0.5 * (_t(LF[-1])[:2] + _t(RF[-1])[:2])
        else:
            self.ft.update_state(to_rf, to_lf, la_rf, la_lf, lf_pose, rf_pose)
            LF = RF = com_final = None
        self.h_phase = self.h_phase[1:] + [list(self.phases[t])]
        self.h_index = self.h_index[1:] + [t]
        self.t += 1
        return LF, RF, self.phases[t], com_final


def device_gait(kind, lf0, rf0, com0, mass, swing_apex=0.15, x_forward=None, y_forward=0.0, foot_yaw=0.0, y_gap=0.18, z_height=0.0,
                keep_forward=False, w_lfrf=None):
    """Parameter block of the DEVICE gait generator (include/mpcb200.h `mpc_gait_t`, csrc/gait.cuh) for the reference gait of `kind`:
    the same arguments as GaitPlan, which it mirrors tick for tick (tests/test_gait_device.py)."""
    T_ds, T_ss, cycles, half = GAITS[kind]
    g = _abi.Gait()
    g.T_ds, g.T_ss, g.cycles, g.half_cycle, g.keep_forward, g.n_uref = T_ds, T_ss, cycles, int(half), int(keep_forward), 0
    g.x_forward = X_FORWARD[kind] if x_forward is None else x_forward
    g.y_forward, g.foot_yaw, g.y_gap, g.z_height, g.swing_apex = y_forward, foot_yaw, y_gap, z_height, swing_apex
    for i in range(12):
        g.lf0[i], g.rf0[i] = float(lf0[i]), float(rf0[i])
    for i in range(3):
        g.com0[i] = float(com0[i])
    g.f_half = mass * GRAVITY / 2.0
    g.w_lfrf = w_lfrf if w_lfrf is not None else (2000.0 if kind == _abi.KIND_FULL else 1e5)
    return g
