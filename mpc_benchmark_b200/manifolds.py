"""aligator.manifolds (fulldynamic_talos.py:23,62; centroidal_talos.py:46)."""
from .api import MultibodyPhaseSpace, VectorSpace  # noqa: F401
