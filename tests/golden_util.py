"""Load tests/golden/*.npz fixtures back into problem dicts (see tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np

from mpc_benchmark_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    z = np.load(os.path.join(HERE, "golden", name))
    raw = z["cfg"].tobytes()  # fixtures written before mpc_config_t grew its trailing `rollout` field: zero-extend (0 = ROLLOUT_LINEAR)
    cfg = _abi.Config.from_buffer_copy(raw + bytes(max(0, C.sizeof(_abi.Config) - len(raw))))
    rb = _abi.Robot.from_buffer_copy(z["robot"].tobytes())
    nk = z["knots"].size // C.sizeof(_abi.Knot)
    nt = z["terms"].size // C.sizeof(_abi.Term)
    knots = (_abi.Knot * nk).from_buffer_copy(z["knots"].tobytes())
    terms = (_abi.Term * nt).from_buffer_copy(z["terms"].tobytes())
    prob = dict(robot=rb, cfg=cfg, knots=knots, terms=terms, x0=np.ascontiguousarray(z["x0"]), xs=np.ascontiguousarray(z["xs"]),
                us=np.ascontiguousarray(z["us"]))
    return prob, z
