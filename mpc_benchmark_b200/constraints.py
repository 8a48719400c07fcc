"""aligator.constraints (fulldynamic_talos.py:207-225; kinodynamic_talos.py:162-171; centroidal_talos.py:245)."""
from .api import BoxConstraint, EqualityConstraintSet, NegativeOrthant  # noqa: F401
