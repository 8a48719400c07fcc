"""Batched mirror of the reference's whole-body inverse-dynamics QP class `IDSolver_ulim` (QP_utils.py:437-573; call site
kinodynamic_talos.py:408-445) on the CUDA QP solver (proxqp.BatchQP -> libmpcb200.so).  SURVEY 8f row f-3.

Same constructor arguments, same `solve(...)` return values (new accelerations, new contact forces, joint torques), with a leading
batch axis on every array.  The reference reads M, nle, the LOCAL contact Jacobians, dJ v and the frame velocities out of a
`pinocchio.Data`; here they are passed as arrays (`RBDTerms`, produced by whatever plays pinocchio's part: the fixture generator
in tests/, a user's own pinocchio, ...), because the rigid-body library is not part of this row.
"""
from dataclasses import dataclass

import numpy as np

from . import proxqp


@dataclass
class RBDTerms:
    """What IDSolver_ulim.computeMatrice reads from `data` (QP_utils.py:515-528), batch-major."""
    nle: np.ndarray  # [B][nv]            data.nle
    Jc: np.ndarray   # [B][6 nk][nv]      getFrameJacobian(..., LOCAL) of every contact frame, stacked
    dJv: np.ndarray  # [B][6 nk]          getFrameJacobianTimeVariation(..., LOCAL) @ v
    vf: np.ndarray   # [B][nk][6]         getFrameVelocity(...): linear, angular


class IDSolver_ulim:
    def __init__(self, model, weights, nk, mu, L, W, contact_ids, force_size, verbose=False, batch=1, device=0):
        if force_size != 6 or nk != 2 or model.nv != 28:
            raise NotImplementedError("the device assembly covers the reference's configuration: nv = 28, two 6-D contacts (kinodynamic_talos.py:408-417)")
        kd = 1  # QP_utils.py:441
        self.baum_Kd = np.diag([kd, kd, kd])
        self.nk, self.contact_ids, self.mu, self.L, self.W, self.force_size = nk, contact_ids, mu, L, W, force_size
        self.model, self.batch = model, batch
        nv = model.nv
        n, neq, nin = 2 * nv - 6 + force_size * nk, nv + force_size * nk, 9 * nk  # QP_utils.py:451-453
        self.n, self.neq, self.nin = n, neq, nin
        u = np.ones(nin) * 100000  # QP_utils.py:486
        g = np.zeros(n)
        H = np.zeros((n, n))
        H[:nv, :nv] = np.eye(nv) * weights[0]
        H[nv:nv + force_size * nk, nv:nv + force_size * nk] = np.eye(force_size * nk) * weights[1]
        qp = proxqp.dense.BatchQP(n, neq, nin, batch, False, dense_backend=proxqp.dense.DenseBackend.PrimalDualLDLT, device=device)
        qp.settings.eps_abs = 1e-3  # QP_utils.py:502-508
        qp.settings.eps_rel = 0
        qp.settings.primal_infeasibility_solving = True
        qp.settings.check_duality_gap = True
        qp.settings.verbose = verbose
        qp.settings.max_iter = 10
        qp.settings.max_iter_in = 10
        qp.init(H, g, np.zeros((neq, n)), np.zeros(neq), np.zeros((nin, n)), np.zeros(nin), u)
        self.qp = qp

    def gamma(self, rbd, cs):
        """dJ v + Kd v_lin + Kd v_ang on the linear rows of the active contacts (QP_utils.py:524-528, as written there)."""
        B = self.batch
        g = np.array(rbd.dJv, float).reshape(B, self.nk, 6).copy()
        vf = np.asarray(rbd.vf, float).reshape(B, self.nk, 6)
        g[:, :, :3] += vf[:, :, :3] @ self.baum_Kd.T + vf[:, :, 3:] @ self.baum_Kd.T
        return (g * np.asarray(cs).reshape(B, self.nk, 1)).reshape(B, 6 * self.nk)

    def computeMatrice(self, rbd, cs, v, a, forces, M):
        """A, b, C, l of QP_utils.py:530-552, assembled by a CUDA kernel straight into the solver's device buffers."""
        self.qp.assemble_id(M, rbd.nle, rbd.Jc, self.gamma(rbd, cs), a, forces, cs, self.mu, self.L, self.W)

    def solve(self, rbd, cs, v, a, forces, M):
        self.computeMatrice(rbd, cs, v, a, forces, M)
        self.qp.solve()
        nv, nf = self.model.nv, self.force_size * self.nk
        x = self.qp.results.x
        anew = np.asarray(a, float).reshape(self.batch, nv) + x[:, :nv]
        new_forces = np.asarray(forces, float).reshape(self.batch, nf) + x[:, nv:nv + nf]
        torque = x[:, nv + nf:]
        return anew, new_forces, torque
