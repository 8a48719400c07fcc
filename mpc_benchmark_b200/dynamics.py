"""aligator.dynamics (fulldynamic_talos.py:100-111; kinodynamic_talos.py:107-112; centroidal_talos.py:202-205)."""
from .api import (  # noqa: F401
    CentroidalFwdDynamics, IntegratorEuler, IntegratorSemiImplEuler, KinodynamicsFwdDynamics, MultibodyConstraintFwdDynamics,
)
