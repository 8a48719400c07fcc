"""Minimal duck-typed stand-in for the `pinocchio` objects the reference scripts pass INTO Aligator
(fulldynamic_talos.py:27-97,146-147,154; kinodynamic_talos.py:35-69; centroidal_talos.py:34-85).

Pinocchio itself is not available offline; this is problem set-up glue (model/data carriers, SE3, forward
kinematics, CoM), not the hot path.  The robot tree is the `mpc_robot_t` data block of include/mpcb200.h.
"""
import numpy as np

from . import _abi, kinematics, talos_like

LOCAL = 0
WORLD = 1
LOCAL_WORLD_ALIGNED = 2


class ContactType:
    CONTACT_3D = 3
    CONTACT_6D = 6


class SE3:
    def __init__(self, rotation=None, translation=None):
        self.rotation = np.eye(3) if rotation is None else np.array(rotation, dtype=float).reshape(3, 3)
        self.translation = np.zeros(3) if translation is None else np.array(translation, dtype=float).reshape(3)

    @staticmethod
    def Identity():
        return SE3()

    @staticmethod
    def from12(v):
        v = np.asarray(v, float)
        return SE3(v[:9].reshape(3, 3), v[9:12])

    def to12(self):
        return np.concatenate([self.rotation.reshape(9), self.translation])

    def copy(self):
        return SE3(self.rotation.copy(), self.translation.copy())

    def __mul__(self, other):
        return SE3(self.rotation @ other.rotation, self.rotation @ other.translation + self.translation)

    def inverse(self):
        return SE3(self.rotation.T, -self.rotation.T @ self.translation)

    def __repr__(self):
        return f"SE3(R=\n{self.rotation},\n p={self.translation})"


class Motion:
    def __init__(self, v=None):
        self.np = np.zeros(6) if v is None else np.array(v, dtype=float).reshape(6)

    @property
    def linear(self):
        return self.np[:3]

    @property
    def angular(self):
        return self.np[3:]

    @staticmethod
    def Zero():
        return Motion()


class Force(Motion):
    pass


class ProximalSettings:
    def __init__(self, absolute_accuracy=1e-6, mu=0.0, max_iter=1):
        self.absolute_accuracy, self.mu, self.max_iter = absolute_accuracy, mu, max_iter


class _Corrector:
    def __init__(self):
        self.Kp = np.zeros(6)
        self.Kd = np.zeros(6)


class RigidConstraintData:
    def __init__(self):
        self.contact_force = Force()


class RigidConstraintModel:
    """pin.RigidConstraintModel(type, model, joint1_id, placement1, joint2_id, placement2, reference_frame)."""

    def __init__(self, ctype, model, joint1_id, joint1_placement, joint2_id=0, joint2_placement=None, reference_frame=LOCAL):
        self.type, self.joint1_id, self.joint1_placement = ctype, joint1_id, joint1_placement
        self.joint2_id, self.joint2_placement, self.reference_frame = joint2_id, joint2_placement, reference_frame
        self.corrector = _Corrector()
        self.name = ""

    def createData(self):
        return RigidConstraintData()


class Frame:
    def __init__(self, name, parent_joint, placement):
        self.name, self.parentJoint, self.parent, self.placement = name, parent_joint, parent_joint, placement


class Data:
    def __init__(self, model):
        self.oMi = [SE3() for _ in range(model.njoints)]
        self.oMf = [SE3() for _ in model.frames]
        self.com = [np.zeros(3)]
        self.mass = [0.0]


class Model:
    """Carrier of the robot tree with the pinocchio attribute names the scripts read."""

    def __init__(self, robot=None, joint_names=None, reference_configurations=None, link_frames=None):
        """robot: `_abi.Robot` (default: the synthetic Talos-shaped tree); joint_names / reference_configurations / link_frames
        come from `urdf.build_robot` when the tree was loaded from a URDF (`model_from_urdf`)."""
        self.robot = robot if robot is not None else talos_like.talos_like_robot()
        rb = self.robot
        names = list(joint_names) if joint_names is not None else list(talos_like.JOINT_NAMES)
        self.nq, self.nv = _abi.NQ, _abi.NV
        self.njoints = rb.nb + 1  # + universe
        self.names = np.array(["universe"] + names, dtype=object)
        big = 1e30
        self.lowerPositionLimit = np.concatenate([[-big] * 7, np.array(rb.q_lo[:])])
        self.upperPositionLimit = np.concatenate([[big] * 7, np.array(rb.q_hi[:])])
        self.effortLimit = np.concatenate([np.zeros(6), np.array(rb.tau_max[:])])
        self.referenceConfigurations = dict(reference_configurations) if reference_configurations is not None else {"half_sitting": talos_like.half_sitting()}
        self.gravity = Motion([rb.gravity[0], rb.gravity[1], rb.gravity[2], 0, 0, 0])
        # frames: universe, root_joint, the joints, then link frames (base_link, torso_2_link, soles, ...)
        self.frames = [Frame("universe", 0, SE3())]
        for j, n in enumerate(names):
            self.frames.append(Frame(n, j + 1, SE3()))
        if link_frames is None:
            self.frames.append(Frame("base_link", 1, SE3()))
            self.frames.append(Frame("torso_2_link", 1 + names.index("torso_2_joint"), SE3()))
            for f, n in enumerate(["left_sole_link", "right_sole_link"]):
                self.frames.append(Frame(n, rb.foot_body[f] + 1, SE3.from12(rb.foot_place[f][:])))
        else:
            for n, (body, place) in link_frames.items():
                self.frames.append(Frame(n, body + 1, SE3.from12(place)))

    def copy(self):
        return self  # immutable carrier

    def __deepcopy__(self, memo):
        return self  # shared by the value copies TrajOptProblem makes of its stages

    def createData(self):
        return Data(self)

    def getFrameId(self, name):
        for i, f in enumerate(self.frames):
            if f.name == name:
                return i
        return len(self.frames)

    def getJointId(self, name):
        names = list(self.names)
        return names.index(name) if name in names else len(names)

    def existFrame(self, name):
        return self.getFrameId(name) < len(self.frames)


def neutral(model):
    q = np.zeros(model.nq)
    q[6] = 1.0
    return q


def forwardKinematics(model, data, q, v=None, a=None):
    Rs, ps = kinematics.body_placements(model.robot, q)
    data.oMi[0] = SE3()
    for b in range(model.robot.nb):
        data.oMi[b + 1] = SE3(Rs[b], ps[b])


def updateFramePlacements(model, data):
    for i, f in enumerate(model.frames):
        data.oMf[i] = data.oMi[f.parentJoint] * f.placement


def framesForwardKinematics(model, data, q):
    forwardKinematics(model, data, q)
    updateFramePlacements(model, data)


def centerOfMass(model, data, q, v=None, *args):
    c, m = kinematics.center_of_mass(model.robot, np.asarray(q, float)[: model.nq])
    data.com[0] = c
    data.mass[0] = m
    return c


def computeTotalMass(model, data=None):
    return float(sum(model.robot.mass[b] for b in range(model.robot.nb)))


def load_talos_like():
    """Stand-in for talos_utils.loadTalos(): (rmodelComplete, rmodel, qComplete, q0) on the synthetic tree."""
    m = Model()
    q0 = m.referenceConfigurations["half_sitting"]
    return m, m, q0.copy(), q0.copy()


def model_from_urdf(urdf_text, srdf_text=None, locked_joints=(), posture="half_sitting", foot_frames=("left_sole_link", "right_sole_link")):
    """`example_robot_data.load(...)` + `buildReducedRobot(locked_joints, q_ref)` of talos_utils.py:31-41 without Pinocchio
    (SURVEY 8f row f-1): parse the URDF (+ SRDF posture), freeze `locked_joints` (names or complete-model joint ids) at the
    posture and return the reduced `Model`; `model.robot` is the `mpc_robot_t` the solver consumes."""
    from . import urdf

    m = urdf.parse_urdf(urdf_text)
    post = urdf.parse_srdf_posture(srdf_text, posture) if srdf_text else {}
    rb, info = urdf.build_robot(m, locked=list(locked_joints), q_locked=post, foot_frames=foot_frames)
    ref = {posture: urdf.reduced_configuration(info, post)} if srdf_text else None
    return Model(robot=rb, joint_names=info["joint_names"], reference_configurations=ref, link_frames=info["frames"])
