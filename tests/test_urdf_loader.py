"""URDF / SRDF loader with joint locking (talos_utils.py:31-41) — host-side, no GPU.
Checks: (1) the synthetic tree survives URDF text -> `mpc_robot_t`; (2) locking the ten joints the reference locks, by the
reference's Pinocchio joint ids, reproduces the reduced tree; (3) with rotated joint origins and non-zero lock angles the
reduced tree has the same total mass, centre of mass and rotational inertia about the origin as the COMPLETE tree evaluated
by an independent forward kinematics over every URDF link; (4) error behaviour."""
import numpy as np
import pytest

from mpc_benchmark_b200 import _abi, pin, talos_like, urdf

REF_LOCKED = [20, 21, 22, 23, 28, 29, 30, 31, 32, 33]  # talos_utils.py:35-36


def fields(rb):
    return dict(parent=np.array(rb.parent[:]), jplace=np.array([list(r) for r in rb.jplace]), axis=np.array([list(r) for r in rb.axis]),
                mass=np.array(rb.mass[:]), com=np.array([list(r) for r in rb.com]), inertia=np.array([list(r) for r in rb.inertia]),
                foot_body=np.array(rb.foot_body[:]), foot_place=np.array([list(r) for r in rb.foot_place]), q_lo=np.array(rb.q_lo[:]),
                q_hi=np.array(rb.q_hi[:]), tau_max=np.array(rb.tau_max[:]), gravity=np.array(rb.gravity[:]))


def assert_same_robot(a, b, tol=1e-12):
    fa, fb = fields(a), fields(b)
    for k in fa:
        assert np.allclose(fa[k], fb[k], rtol=0, atol=tol), k


def test_round_trip_without_extra_joints():
    text, srdf = urdf.synthetic_complete_urdf(extra=False)
    rb, info = urdf.build_robot(urdf.parse_urdf(text))
    assert info["joint_names"] == list(talos_like.JOINT_NAMES)
    assert_same_robot(rb, talos_like.talos_like_robot())
    q = urdf.reduced_configuration(info, urdf.parse_srdf_posture(srdf), base=talos_like.half_sitting()[:7])
    assert np.allclose(q, talos_like.half_sitting())


def test_locking_the_reference_joint_ids_gives_the_reduced_tree():
    text, srdf = urdf.synthetic_complete_urdf(extra=True)
    m = urdf.parse_urdf(text)
    order = urdf.complete_joint_order(m)
    assert len(order) == 32 and [order[i - 2] for i in REF_LOCKED] == [
        "arm_left_5_joint", "arm_left_6_joint", "arm_left_7_joint", "gripper_left_joint", "arm_right_5_joint", "arm_right_6_joint",
        "arm_right_7_joint", "gripper_right_joint", "head_1_joint", "head_2_joint"]
    rb, info = urdf.build_robot(m, locked=REF_LOCKED, q_locked=urdf.parse_srdf_posture(srdf))
    assert info["joint_names"] == list(talos_like.JOINT_NAMES)
    assert_same_robot(rb, talos_like.talos_like_robot())
    model = pin.model_from_urdf(text, srdf, REF_LOCKED)
    assert model.getFrameId("left_sole_link") < len(model.frames) and model.names[2] == "leg_left_1_joint"
    assert model.effortLimit.shape == (28,) and np.allclose(model.referenceConfigurations["half_sitting"][7:], talos_like.half_sitting()[7:])


def rodrigues(axis, ang):
    return urdf.axis_angle_to_R(axis, ang)


def composite_of_complete(m, angles):
    """mass, first moment and inertia about the ROOT-link origin of the complete tree (every link, every joint at `angles`)."""
    by_parent = {}
    for j in m.joints.values():
        by_parent.setdefault(j.parent, []).append(j)
    tot = dict(m=0.0, mc=np.zeros(3), I=np.zeros((3, 3)))

    def visit(link, R, p):
        L = m.links[link]
        if L.mass > 0:
            c = R @ L.com + p
            tot["m"] += L.mass
            tot["mc"] += L.mass * c
            tot["I"] += R @ L.inertia @ R.T + L.mass * (c @ c * np.eye(3) - np.outer(c, c))
        for j in by_parent.get(link, []):
            Rj, pj = R @ j.R, R @ j.p + p
            if j.type != "fixed":
                Rj = Rj @ rodrigues(j.axis, angles.get(j.name, 0.0))
            visit(j.child, Rj, pj)

    visit(m.root, np.eye(3), np.zeros(3))
    return tot


def composite_of_reduced(rb, names, angles):
    R, p = [None] * rb.nb, [None] * rb.nb
    tot = dict(m=0.0, mc=np.zeros(3), I=np.zeros((3, 3)))
    for b in range(rb.nb):
        Rj, pj = np.array(rb.jplace[b][:9]).reshape(3, 3), np.array(rb.jplace[b][9:12])
        if b == 0:
            R[b], p[b] = Rj, pj
        else:
            par = rb.parent[b]
            R[b] = R[par] @ Rj @ rodrigues(np.array(rb.axis[b][:]), angles.get(names[b], 0.0))
            p[b] = R[par] @ pj + p[par]
        c = R[b] @ np.array(rb.com[b][:]) + p[b]
        I = np.array(rb.inertia[b][:]).reshape(3, 3)
        tot["m"] += rb.mass[b]
        tot["mc"] += rb.mass[b] * c
        tot["I"] += R[b] @ I @ R[b].T + rb.mass[b] * (c @ c * np.eye(3) - np.outer(c, c))
    return tot


def test_merged_inertias_match_independent_forward_kinematics():
    text, srdf = urdf.synthetic_complete_urdf(extra=True)
    # make the merge non-trivial: rotate the origins of locked joints and lock them at non-zero angles
    import re

    def set_origin(t, joint, xyz, rpy):
        t2, n = re.subn(r'(<joint name="%s" type="revolute">.*?<origin )xyz="[^"]*" rpy="[^"]*"' % joint, r'\1xyz="%s" rpy="%s"' % (xyz, rpy), t, count=1, flags=re.S)
        assert n == 1
        return t2

    text = set_origin(text, "arm_left_6_joint", "0.01 0.02 -0.05", "0.3 -0.2 0.5")
    text = set_origin(text, "head_2_joint", "0.0 0.03 0.05", "-0.4 0.1 0.2")
    assert 'rpy="0.3 -0.2 0.5"' in text and 'rpy="-0.4 0.1 0.2"' in text
    m = urdf.parse_urdf(text)
    rng = np.random.default_rng(0)
    order = urdf.complete_joint_order(m)
    angles = {n: float(rng.uniform(-0.5, 0.5)) for n in order}
    locked = [order[i - 2] for i in REF_LOCKED]
    rb, info = urdf.build_robot(m, locked=locked, q_locked={n: angles[n] for n in locked})
    a = composite_of_complete(m, angles)
    b = composite_of_reduced(rb, info["joint_names"], angles)
    assert abs(a["m"] - b["m"]) < 1e-10 and np.allclose(a["mc"], b["mc"], atol=1e-10) and np.allclose(a["I"], b["I"], atol=1e-9)
    # and the result differs from the un-rotated tree (the test is not vacuous)
    rb0, _ = urdf.build_robot(urdf.parse_urdf(urdf.synthetic_complete_urdf(extra=True)[0]), locked=REF_LOCKED)
    assert not np.allclose(np.array(rb.inertia[18][:]), np.array(rb0.inertia[18][:]), atol=1e-6)


def test_loader_errors():
    text, _ = urdf.synthetic_complete_urdf(extra=True)
    m = urdf.parse_urdf(text)
    with pytest.raises(NotImplementedError):
        urdf.build_robot(m)  # 32 joints: not the shape the kernels are built for
    with pytest.raises(ValueError):
        urdf.build_robot(m, locked=["no_such_joint"])
    with pytest.raises(ValueError):
        urdf.build_robot(m, locked=[99])
    with pytest.raises(NotImplementedError):
        urdf.parse_urdf(text.replace('name="head_1_joint" type="revolute"', 'name="head_1_joint" type="prismatic"'))
    with pytest.raises(ValueError):
        urdf.build_robot(urdf.parse_urdf(urdf.synthetic_complete_urdf(extra=False)[0]), foot_frames=("nope", "right_sole_link"))


def test_loaded_robot_has_the_abi_layout(oracle):
    """The loader's struct is consumed by the oracle exactly like the built-in tree (same bytes in, same dynamics out)."""
    text, srdf = urdf.synthetic_complete_urdf(extra=True)
    rb, _ = urdf.build_robot(urdf.parse_urdf(text), locked=REF_LOCKED)
    x = np.concatenate([talos_like.half_sitting(), np.zeros(28)])
    a = oracle.rnea(rb, x, np.zeros(28))
    b = oracle.rnea(talos_like.talos_like_robot(), x, np.zeros(28))
    assert np.allclose(a, b, atol=1e-9) and abs(a[2] - sum(rb.mass[:]) * 9.81) < 1e-6 * abs(a[2]) + 1e-6
