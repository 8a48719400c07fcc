"""URDF / SRDF -> `mpc_robot_t` (the model layer on the far side of the hot path).

The reference gets its robot tree from `example_robot_data.load("talos")` followed by
`buildReducedRobot(locked_joints, q_ref)` (talos_utils.py:31-41): a URDF (+ SRDF reference posture) parsed by
Pinocchio, with ten joints frozen at their half-sitting angles.  This module does that step without Pinocchio:

  parse_urdf(text)              links (mass, com, inertia), joints (type, parent/child, origin, axis, limits)
  parse_srdf_posture(text, n)   {joint name: value} of the group_state `n` ("half_sitting")
  build_robot(model, ...)       free-flyer tree -> `_abi.Robot` (= mpc_robot_t): fixed joints and LOCKED revolute joints are
                                merged into their parent body (lumped mass / centre of mass / rotational inertia, placements
                                composed through the frozen rotation), bodies ordered parents-first, siblings by joint name
                                (the order Pinocchio's URDF parser produces and the reference's joint ids rely on)

The CUDA kernels are compiled for the Talos-reduced shape (free-flyer + 22 revolute joints, two sole frames); other trees
are rejected with a message rather than silently truncated.  Conventions as in `include/mpcb200.h`: a body frame is its
joint frame (URDF child-link frame), `jplace` = joint placement in the parent BODY frame, inertia about the com in body axes.
"""
import xml.etree.ElementTree as ET

import numpy as np

from . import _abi


# ------------------------------------------------------------------ small SE(3) helpers (host side, numpy)
def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def axis_angle_to_R(axis, angle):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * K @ K


def _vec(s, n=3, default=0.0):
    if s is None:
        return np.full(n, default)
    v = np.array([float(x) for x in s.split()])
    assert v.size == n, s
    return v


class Link:
    def __init__(self, name, mass=0.0, com=None, inertia=None):
        self.name, self.mass = name, float(mass)
        self.com = np.zeros(3) if com is None else np.asarray(com, float)
        self.inertia = np.zeros((3, 3)) if inertia is None else np.asarray(inertia, float)  # about the com, link axes


class Joint:
    def __init__(self, name, jtype, parent, child, R, p, axis, lower, upper, effort):
        self.name, self.type, self.parent, self.child = name, jtype, parent, child
        self.R, self.p, self.axis = R, p, axis
        self.lower, self.upper, self.effort = lower, upper, effort


class UrdfModel:
    def __init__(self, name, links, joints):
        self.name, self.links, self.joints = name, links, joints
        children = {j.child for j in joints.values()}
        roots = [n for n in links if n not in children]
        if len(roots) != 1:
            raise ValueError(f"URDF must have exactly one root link, found {roots}")
        self.root = roots[0]


def parse_urdf(text):
    root = ET.fromstring(text)
    if root.tag != "robot":
        raise ValueError("not a URDF: root element is <%s>" % root.tag)
    links, joints = {}, {}
    for e in root.findall("link"):
        name = e.get("name")
        ine = e.find("inertial")
        if ine is None:
            links[name] = Link(name)
            continue
        o = ine.find("origin")
        xyz = _vec(o.get("xyz") if o is not None else None)
        Ri = rpy_to_R(_vec(o.get("rpy") if o is not None else None))
        m = float(ine.find("mass").get("value"))
        it = ine.find("inertia")
        g = lambda k: float(it.get(k, "0"))  # noqa: E731
        I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        links[name] = Link(name, m, xyz, Ri @ I @ Ri.T)
    for e in root.findall("joint"):
        name, jtype = e.get("name"), e.get("type")
        o = e.find("origin")
        R = rpy_to_R(_vec(o.get("rpy") if o is not None else None))
        p = _vec(o.get("xyz") if o is not None else None)
        ax = e.find("axis")
        axis = _vec(ax.get("xyz")) if ax is not None else np.array([1.0, 0, 0])
        lim = e.find("limit")
        lower = float(lim.get("lower", "-inf")) if lim is not None else -np.inf
        upper = float(lim.get("upper", "inf")) if lim is not None else np.inf
        effort = float(lim.get("effort", "0")) if lim is not None else 0.0
        if jtype == "continuous":
            lower, upper = -np.inf, np.inf
        if jtype not in ("revolute", "continuous", "fixed"):
            raise NotImplementedError(f"joint {name}: type '{jtype}' is not supported (revolute / continuous / fixed only)")
        joints[name] = Joint(name, jtype, e.find("parent").get("link"), e.find("child").get("link"), R, p, axis, lower, upper, effort)
    for j in joints.values():
        if j.parent not in links or j.child not in links:
            raise ValueError(f"joint {j.name} references an unknown link")
    return UrdfModel(root.get("name", "robot"), links, joints)


def parse_srdf_posture(text, name="half_sitting"):
    """Joint values of <group_state name=...> (SRDF), the source of `referenceConfigurations` in Pinocchio."""
    root = ET.fromstring(text)
    for gs in root.findall("group_state"):
        if gs.get("name") == name:
            return {j.get("name"): [float(x) for x in j.get("value").split()] for j in gs.findall("joint")}
    raise KeyError(f"SRDF has no group_state '{name}'")


def complete_joint_order(model):
    """Movable joints of the complete model in Pinocchio's order (depth first, siblings by joint name); index + 2 is the
    Pinocchio joint id (0 = universe, 1 = root_joint), which is what `locked_joints = [20, ...]` in talos_utils.py:35 means."""
    by_parent = {}
    for j in model.joints.values():
        by_parent.setdefault(j.parent, []).append(j)
    order = []

    def visit(link):
        for j in sorted(by_parent.get(link, []), key=lambda j: j.name):
            if j.type != "fixed":
                order.append(j.name)
            visit(j.child)

    visit(model.root)
    return order


class _Body:
    def __init__(self, name, parent, R, p, axis, lower, upper, effort):
        self.name, self.parent, self.R, self.p, self.axis = name, parent, R, p, axis
        self.lower, self.upper, self.effort = lower, upper, effort
        self.parts = []   # (mass, com, inertia about com) in the body frame
        self.frames = {}  # link name -> (R, p) in the body frame


def _lump(parts):
    m = sum(p[0] for p in parts)
    if m <= 0.0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    c = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for mi, ci, Ii in parts:
        d = ci - c
        I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
    return m, c, I


def build_robot(model, locked=None, q_locked=None, foot_frames=("left_sole_link", "right_sole_link"), gravity=(0.0, 0.0, -9.81)):
    """Reduced free-flyer tree as `_abi.Robot`.
    locked: joint NAMES or Pinocchio joint ids of the complete model (ints >= 2) to freeze; q_locked: {joint name: angle}
    (e.g. the SRDF half-sitting posture) — angles of the locked joints, 0 when absent.
    Returns (robot, info) with info = dict(joint_names, frames {link: (body index, placement 12)}, complete_order)."""
    order = complete_joint_order(model)
    locked_names = set()
    for l in (locked or []):
        if isinstance(l, (int, np.integer)):
            if not 2 <= l < len(order) + 2:
                raise ValueError(f"locked joint id {l} out of range (complete model has joint ids 2..{len(order) + 1})")
            locked_names.add(order[l - 2])
        else:
            if l not in model.joints:
                raise ValueError(f"locked joint '{l}' is not in the URDF")
            locked_names.add(l)
    q_locked = q_locked or {}
    by_parent = {}
    for j in model.joints.values():
        by_parent.setdefault(j.parent, []).append(j)

    bodies = [_Body("root_joint", -1, np.eye(3), np.zeros(3), np.array([0.0, 0, 1]), -np.inf, np.inf, 0.0)]

    def attach(link, b, R, p):
        """link rigidly attached to body b at placement (R, p) in b's frame"""
        L = model.links[link]
        bodies[b].frames[link] = (R, p)
        if L.mass > 0.0:
            bodies[b].parts.append((L.mass, R @ L.com + p, R @ L.inertia @ R.T))
        for j in sorted(by_parent.get(link, []), key=lambda j: j.name):
            Rj, pj = R @ j.R, R @ j.p + p  # joint frame in b's frame
            if j.type == "fixed" or j.name in locked_names:
                ang = 0.0
                if j.type != "fixed":
                    v = q_locked.get(j.name, 0.0)
                    ang = float(v[0] if isinstance(v, (list, tuple, np.ndarray)) else v)
                    Rj = Rj @ axis_angle_to_R(j.axis, ang)
                attach(j.child, b, Rj, pj)
            else:
                bodies.append(_Body(j.name, b, Rj, pj, j.axis / np.linalg.norm(j.axis), j.lower, j.upper, j.effort))
                attach(j.child, len(bodies) - 1, np.eye(3), np.zeros(3))

    attach(model.root, 0, np.eye(3), np.zeros(3))
    if len(bodies) != _abi.NB:
        raise NotImplementedError(f"reduced tree has {len(bodies) - 1} revolute joints; the kernels are built for {_abi.NB - 1} "
                                  "(Talos reduced: lock joints until 22 remain)")
    rb = _abi.Robot()
    rb.nb = _abi.NB
    frames = {}
    for i, b in enumerate(bodies):
        m, c, I = _lump(b.parts)
        if m <= 0.0:
            raise ValueError(f"body '{b.name}' has no mass after merging fixed links")
        rb.parent[i] = b.parent
        rb.jplace[i][:] = list(b.R.reshape(9)) + list(b.p)
        rb.axis[i][:] = list(b.axis)
        rb.mass[i] = m
        rb.com[i][:] = list(c)
        rb.inertia[i][:] = list(I.reshape(9))
        for ln, (R, p) in b.frames.items():
            frames[ln] = (i, list(R.reshape(9)) + list(p))
        if i > 0:
            if not (np.isfinite(b.lower) and np.isfinite(b.upper)):
                raise ValueError(f"joint '{b.name}' has no position limits (the reference builds BoxConstraints from them, full:209)")
            rb.q_lo[i - 1], rb.q_hi[i - 1], rb.tau_max[i - 1] = b.lower, b.upper, b.effort
    for f, ln in enumerate(foot_frames):
        if ln not in frames:
            raise ValueError(f"foot frame '{ln}' is not a link of the URDF")
        rb.foot_body[f] = frames[ln][0]
        rb.foot_place[f][:] = frames[ln][1]
    rb.gravity[:] = list(gravity)
    return rb, dict(joint_names=[b.name for b in bodies], frames=frames, complete_order=order)


def reduced_configuration(info, posture, base=(0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)):
    """q (nq = 29) of the reduced model from an SRDF posture: base pose + the remaining joints in body order."""
    q = list(base)
    for n in info["joint_names"][1:]:
        v = posture.get(n, 0.0)
        q.append(float(v[0] if isinstance(v, (list, tuple, np.ndarray)) else v))
    return np.array(q)


# ------------------------------------------------------------------ synthetic "complete" Talos-shaped URDF / SRDF (tests, examples)
def _f(x):
    return repr(float(x))


def _rpy_str(v):
    return " ".join(_f(x) for x in v)


def synthetic_complete_urdf(extra=True):
    """URDF + SRDF text of the synthetic Talos-shaped tree (talos_like.BODY_TABLE).  With `extra`, the ten joints the
    reference locks (arm 5-7 + gripper per side, head 1-2; talos_utils.py:35-36) are added below arm_*_4 / torso_2, carrying
    part of the mass of the body they merge into: every sub-link shares the lumped body's centre of mass and a proportional
    inertia, so locking them at angle 0 must reproduce `talos_like.talos_like_robot()` exactly."""
    from . import talos_like as tl

    lo, hi, tau = tl.joint_limits()
    out = ['<robot name="talos_like">']
    split = {}  # body index -> fraction kept by the body's own link
    if extra:
        for n in ("arm_left_4_joint", "arm_right_4_joint"):
            split[tl.JOINT_NAMES.index(n)] = 0.6
        split[tl.JOINT_NAMES.index("torso_2_joint")] = 0.9

    def link(name, mass, com, I):
        out.append(f'<link name="{name}"><inertial><origin xyz="{_rpy_str(com)}" rpy="0 0 0"/><mass value="{_f(mass)}"/>'
                   f'<inertia ixx="{_f(I[0, 0])}" ixy="{_f(I[0, 1])}" ixz="{_f(I[0, 2])}" iyy="{_f(I[1, 1])}" iyz="{_f(I[1, 2])}" izz="{_f(I[2, 2])}"/></inertial></link>')

    def joint(name, jtype, parent, child, xyz, axis=(0, 0, 1), lim=None):
        s = f'<joint name="{name}" type="{jtype}"><parent link="{parent}"/><child link="{child}"/><origin xyz="{_rpy_str(xyz)}" rpy="0 0 0"/>'
        if jtype != "fixed":
            s += f'<axis xyz="{_rpy_str(axis)}"/><limit lower="{_f(lim[0])}" upper="{_f(lim[1])}" effort="{_f(lim[2])}" velocity="10"/>'
        out.append(s + "</joint>")

    link_of = {}
    for b, (name, par, t, axis, mass, com, half) in enumerate(tl.BODY_TABLE):
        hx, hy, hz = half
        I = mass / 3.0 * np.diag([hy * hy + hz * hz, hx * hx + hz * hz, hx * hx + hy * hy])
        ln = "base_link" if b == 0 else name.replace("_joint", "_link")
        link_of[b] = ln
        f = split.get(b, 1.0)
        link(ln, f * mass, com, f * I)
        if b > 0:
            joint(name, "revolute", link_of[par], ln, t, axis, (lo[b - 1], hi[b - 1], tau[b - 1]))
        if b in split:
            rest = 1.0 - f
            if "arm" in name:
                side = "left" if "left" in name else "right"
                chain = [f"arm_{side}_5", f"arm_{side}_6", f"arm_{side}_7", f"gripper_{side}"]
                offs = [np.array([0.0, 0.0, -0.05 * (i + 1)]) for i in range(4)]
            else:
                chain, offs = ["head_1", "head_2"], [np.array([0.0, 0.0, 0.3]), np.array([0.0, 0.0, 0.05])]
            parent_link, origin_acc = ln, np.zeros(3)
            for i, cn in enumerate(chain):
                origin_acc = origin_acc + offs[i]
                link(cn + "_link", rest * mass / len(chain), np.asarray(com) - origin_acc, rest * I / len(chain))
                joint(cn + "_joint", "revolute", parent_link, cn + "_link", offs[i], (0, 1, 0), (-1.0, 1.0, 10.0))
                parent_link = cn + "_link"
    for f, side in enumerate(("left", "right")):
        out.append(f'<link name="{side}_sole_link"/>')
        joint(f"{side}_sole_fix_joint", "fixed", f"leg_{side}_6_link", f"{side}_sole_link", tl.SOLE_OFFSET)
    out.append("</robot>")
    q = tl.half_sitting()
    srdf = ['<robot name="talos_like"><group_state name="half_sitting" group="all">']
    for n, v in zip(tl.JOINT_NAMES[1:], q[7:]):
        srdf.append(f'<joint name="{n}" value="{_f(v)}"/>')
    srdf.append("</group_state></robot>")
    return "\n".join(out), "\n".join(srdf)
