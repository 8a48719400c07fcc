"""mpc_benchmark_b200 — B200-native batched ProxDDP for the Talos walking MPC of edantec/MPC_benchmark.

Drop-in for the one hot path: `import mpc_benchmark_b200 as aligator` and
`from mpc_benchmark_b200 import manifolds, dynamics, constraints` replace the `aligator` imports of
centroidal_talos.py / kinodynamic_talos.py / fulldynamic_talos.py.  Compute runs in
libmpcb200.so (CUDA sm_100a, include/mpcb200.h); there is no CPU fallback.
"""
from . import constraints, dynamics, manifolds, pin  # noqa: F401
from .api import (  # noqa: F401
    LQ_SOLVER_PARALLEL, LQ_SOLVER_SERIAL, LQ_SOLVER_STAGEDENSE, ROLLOUT_LINEAR, ROLLOUT_NONLINEAR,
    AngularAccelerationResidual, AngularMomentumResidual, CenterOfMassTranslationResidual, CentroidalAccelerationResidual,
    CentroidalCoMResidual, CentroidalMomentumDerivativeResidual, CentroidalMomentumResidual, CentroidalWrenchConeResidual,
    ContactForceResidual, ContactMap, ControlErrorResidual, CostStack, DCMPositionResidual, FramePlacementResidual,
    FrameTranslationResidual, FrameVelocityResidual, LinearMomentumResidual, MultibodyWrenchConeResidual, QuadraticControlCost,
    QuadraticResidualCost, QuadraticStateCost, Results, SolverFDDP, SolverProxDDP, StageConstraint, StageModel, StateErrorResidual,
    TrajOptProblem, VerboseLevel, Workspace,
)

__all__ = [n for n in dir() if not n.startswith("_")]
