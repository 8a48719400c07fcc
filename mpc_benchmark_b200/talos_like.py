"""Talos-SHAPED synthetic robot tree.  NOT the real Talos parameters.

The reference loads `talos_reduced.urdf` from example-robot-data and locks 10 joints
(talos_utils.py:31-41); that package is not available offline, so only what the reference pins is
reproduced exactly: nq=29 / nv=28 / nu=22 (plot.py:488-490), joint order base, left leg(6),
right leg(6), torso(2), left arm(4), right arm(4) (fulldynamic_talos.py:121-134), base height 1.01927
(bullet_robot.py:22), sole frames `left_sole_link` / `right_sole_link`.  Link lengths, masses and
limits are plausible values; every reported number says "synthetic Talos-shaped model".
The tree is DATA: any 23-body free-flyer tree in the same schema can be passed to the solver.
"""
import numpy as np

from . import _abi

HALF_SITTING_LEGS = (0.0, 0.0, -0.411354, 0.859395, -0.448041, -0.001708)
HALF_SITTING = dict(
    base=(0.0, 0.0, 1.01927, 0.0, 0.0, 0.0, 1.0),
    leg_left=HALF_SITTING_LEGS,
    leg_right=HALF_SITTING_LEGS,
    torso=(0.0, 0.006761),
    arm_left=(0.25847, 0.173046, -0.0002, -0.525366),
    arm_right=(-0.25847, -0.173046, 0.0002, -0.525366),
)

_X, _Y, _Z = (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0)

# name, parent, joint translation in parent frame, axis, mass, com, box half-extents (for inertia)
_BODIES = [
    ("root_joint", -1, (0, 0, 0), _Z, 13.53, (-0.08, 0.0, -0.03), (0.12, 0.14, 0.10)),
]


def _leg(side, s):
    p0 = len(_BODIES)
    return [
        (f"leg_{side}_1_joint", 0, (-0.02, s * 0.085, -0.27105), _Z, 2.70, (0.02, s * 0.01, 0.02), (0.06, 0.06, 0.06)),
        (f"leg_{side}_2_joint", p0, (0, 0, 0), _X, 3.00, (-0.015, s * 0.02, 0.0), (0.07, 0.06, 0.06)),
        (f"leg_{side}_3_joint", p0 + 1, (0, 0, 0), _Y, 6.24, (0.01, s * 0.03, -0.17), (0.07, 0.07, 0.19)),
        (f"leg_{side}_4_joint", p0 + 2, (0, 0, -0.38), _Y, 3.76, (0.01, s * 0.01, -0.14), (0.06, 0.06, 0.16)),
        (f"leg_{side}_5_joint", p0 + 3, (0, 0, -0.325), _Y, 1.29, (-0.01, s * 0.005, 0.0), (0.05, 0.05, 0.04)),
        (f"leg_{side}_6_joint", p0 + 4, (0, 0, 0), _X, 1.60, (0.0, 0.0, -0.07), (0.10, 0.065, 0.03)),
    ]


def _arm(side, s, torso2):
    p0 = None  # filled by caller
    return [
        (f"arm_{side}_1_joint", torso2, (0.0, s * 0.1575, 0.232), _Z, 2.71, (-0.01, s * 0.07, -0.02), (0.05, 0.08, 0.05)),
        (f"arm_{side}_2_joint", -2, (0.00493, s * 0.1365, 0.04673), _X, 2.43, (0.02, s * 0.02, -0.03), (0.05, 0.05, 0.06)),
        (f"arm_{side}_3_joint", -2, (0, 0, 0), _Z, 2.53, (0.0, s * 0.005, -0.12), (0.05, 0.05, 0.12)),
        (f"arm_{side}_4_joint", -2, (0.02, 0, -0.273), _Y, 3.50, (-0.01, 0.0, -0.13), (0.05, 0.05, 0.17)),
    ]


def _build_table():
    bodies = list(_BODIES)
    for side, s in (("left", 1.0), ("right", -1.0)):
        p0 = len(bodies)
        leg = _leg(side, s)
        fixed = []
        for i, b in enumerate(leg):
            par = 0 if i == 0 else p0 + i - 1
            fixed.append((b[0], par) + b[2:])
        bodies += fixed
    t1 = len(bodies)
    bodies.append(("torso_1_joint", 0, (0, 0, 0.0722), _Z, 2.29, (0.0, 0.0, 0.03), (0.07, 0.09, 0.04)))
    bodies.append(("torso_2_joint", t1, (0, 0, 0), _Y, 17.55, (-0.04, 0.0, 0.20), (0.12, 0.17, 0.22)))
    t2 = t1 + 1
    for side, s in (("left", 1.0), ("right", -1.0)):
        p0 = len(bodies)
        arm = _arm(side, s, t2)
        for i, b in enumerate(arm):
            par = t2 if i == 0 else p0 + i - 1
            bodies.append((b[0], par) + b[2:])
    return bodies


BODY_TABLE = _build_table()
JOINT_NAMES = [b[0] for b in BODY_TABLE]
SOLE_OFFSET = (0.0, 0.0, -0.107)

# joint limits (rad) and effort limits (N m), order: legs(6+6), torso(2), arms(4+4)
_LEG_LO = (-0.35, -0.52, -2.10, 0.0, -1.31, -0.52)
_LEG_HI = (1.57, 0.52, 0.70, 2.62, 0.77, 0.52)
_LEG_TAU = (100.0, 160.0, 160.0, 300.0, 160.0, 100.0)
_ARM_LO_L = (-1.57, 0.01, -2.43, -2.23)
_ARM_HI_L = (0.79, 2.87, 2.43, -0.01)
_ARM_TAU = (44.0, 44.0, 22.0, 22.0)


def joint_limits():
    lo = list(_LEG_LO) + list(_LEG_LO) + [-1.26, -0.23] + list(_ARM_LO_L)
    hi = list(_LEG_HI) + list(_LEG_HI) + [1.26, 0.73] + list(_ARM_HI_L)
    # right arm mirrors the left (joints 1-3 flip sign, elbow keeps it)
    lo += [-_ARM_HI_L[0], -_ARM_HI_L[1], -_ARM_HI_L[2], _ARM_LO_L[3]]
    hi += [-_ARM_LO_L[0], -_ARM_LO_L[1], -_ARM_LO_L[2], _ARM_HI_L[3]]
    # right leg: yaw/roll mirrored
    lo[6], hi[6] = -_LEG_HI[0], -_LEG_LO[0]
    tau = list(_LEG_TAU) * 2 + [78.0, 78.0] + list(_ARM_TAU) * 2
    return np.array(lo), np.array(hi), np.array(tau)


def _place(R, p):
    return list(np.asarray(R, float).reshape(9)) + list(np.asarray(p, float).reshape(3))


def half_sitting():
    h = HALF_SITTING
    return np.array(h["base"] + h["leg_left"] + h["leg_right"] + h["torso"] + h["arm_left"] + h["arm_right"], dtype=float)


def talos_like_robot():
    """Return the `_abi.Robot` struct of the synthetic Talos-shaped tree."""
    rb = _abi.Robot()
    rb.nb = _abi.NB
    assert len(BODY_TABLE) == _abi.NB
    for b, (name, par, t, axis, mass, com, half) in enumerate(BODY_TABLE):
        rb.parent[b] = par
        rb.jplace[b][:] = _place(np.eye(3), t)
        rb.axis[b][:] = axis
        rb.mass[b] = mass
        rb.com[b][:] = com
        hx, hy, hz = half
        I = mass / 3.0 * np.diag([hy * hy + hz * hz, hx * hx + hz * hz, hx * hx + hy * hy])
        rb.inertia[b][:] = list(I.reshape(9))
    rb.foot_body[0] = JOINT_NAMES.index("leg_left_6_joint")
    rb.foot_body[1] = JOINT_NAMES.index("leg_right_6_joint")
    for f in range(2):
        rb.foot_place[f][:] = _place(np.eye(3), SOLE_OFFSET)
    lo, hi, tau = joint_limits()
    rb.q_lo[:] = list(lo)
    rb.q_hi[:] = list(hi)
    rb.tau_max[:] = list(tau)
    rb.gravity[:] = [0.0, 0.0, -9.81]
    return rb
