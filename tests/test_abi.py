"""The C-ABI boundary: struct layouts agree between include/mpcb200.h (as compiled into the oracle and into
libmpcb200.so) and the ctypes mirrors, and libmpcb200.so exports every symbol the header declares.  No GPU needed."""
import ctypes as C
import os
import re

import pytest

from mpc_benchmark_b200 import _abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = [(0, _abi.Robot), (1, _abi.Config), (2, _abi.Knot), (3, _abi.Term), (4, _abi.Info)]
QP_STRUCTS = [(0, _abi.QPSettings), (1, _abi.QPInfo)]


def test_struct_sizes_match_oracle(oracle):
    for which, s in STRUCTS:
        assert oracle.lib().orc_sizeof(which) == C.sizeof(s)
    for which, s in QP_STRUCTS:
        assert oracle.lib().orc_sizeof(5 + which) == C.sizeof(s)


def _header_functions():
    names = set()
    for hdr in ("mpcb200.h", "mpcqp_b200.h"):  # every header under include/
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mpc_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        pytest.skip("libmpcb200.so not built (run __graft_entry__.build())")
    lib = C.CDLL(_native.LIB_PATH)  # loading needs no GPU; compute entry points are not called here
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mpcb200.h but not exported"
    for n in _native.EXPORTS:
        assert n in names, f"{n} bound in _native.py but not declared in the header"
    lib.mpc_abi_sizeof.argtypes = [C.c_int32]
    for which, s in STRUCTS:
        assert lib.mpc_abi_sizeof(which) == C.sizeof(s)
    lib.mpc_qp_abi_sizeof.argtypes = [C.c_int32]
    for which, s in QP_STRUCTS:
        assert lib.mpc_qp_abi_sizeof(which) == C.sizeof(s)


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a machine without CUDA the product path must raise, not compute."""
    import torch

    if torch.cuda.is_available() or not os.path.exists(_native.LIB_PATH):
        pytest.skip("needs the built library and NO GPU")
    from mpc_benchmark_b200 import problems
    from mpc_benchmark_b200.batch import BatchSolver

    p = problems.cent_standing_problem(T=4)
    with pytest.raises(_native.NativeError):
        BatchSolver(p["robot"], p["cfg"], 1)
