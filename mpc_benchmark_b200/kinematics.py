"""Host-side (numpy) forward kinematics of the robot tree.

Only what the reference scripts compute on the host BEFORE calling the solver: initial sole placements,
centre of mass and total mass (fulldynamic_talos.py:50-51,67-68; centroidal_talos.py:60-65).  This is
problem set-up, not the hot path; the hot-path kinematics live in csrc/ (CUDA).
"""
import numpy as np


def quat_to_R(q):
    x, y, z, w = q
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array(
        [
            [1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
            [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
            [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)],
        ]
    )


def rodrigues(axis, angle):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def body_placements(rb, q):
    """World placements (R, p) of every body for configuration q (nq=29)."""
    q = np.asarray(q, float)
    Rs, ps = [quat_to_R(q[3:7])], [q[0:3].copy()]
    for b in range(1, rb.nb):
        pl = np.array(rb.jplace[b][:])
        Rp, pp = pl[:9].reshape(3, 3), pl[9:]
        par = rb.parent[b]
        R = Rs[par] @ Rp @ rodrigues(rb.axis[b][:], q[6 + b])
        p = Rs[par] @ pp + ps[par]
        Rs.append(R)
        ps.append(p)
    return Rs, ps


def foot_placements(rb, q):
    """12-vectors (R row-major, p) of left/right sole frames."""
    Rs, ps = body_placements(rb, q)
    out = []
    for f in range(2):
        b = rb.foot_body[f]
        pl = np.array(rb.foot_place[f][:])
        R = Rs[b] @ pl[:9].reshape(3, 3)
        p = Rs[b] @ pl[9:] + ps[b]
        out.append(np.concatenate([R.reshape(9), p]))
    return out


def center_of_mass(rb, q):
    Rs, ps = body_placements(rb, q)
    m = 0.0
    mc = np.zeros(3)
    for b in range(rb.nb):
        c = Rs[b] @ np.array(rb.com[b][:]) + ps[b]
        m += rb.mass[b]
        mc += rb.mass[b] * c
    return mc / m, m
