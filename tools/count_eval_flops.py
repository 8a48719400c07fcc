"""Freeze F_eval: ALGORITHMIC FLOPs of one per-knot evaluation, counted by running the CPU oracle's
knot evaluation with an instrumented scalar type (oracle/count_flops.cpp: every `double` of the oracle headers becomes a counting
wrapper).  Writes profiles/eval_flops.json, which bench.py reads for `roofline_eval`.  usage: python tools/count_eval_flops.py"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_benchmark_b200 import _abi, problems  # noqa: E402


def main():
    so = os.path.join(ROOT, "oracle", "_build", "libcountflops.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-w", "-o", so, os.path.join(ROOT, "oracle", "count_flops.cpp")])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    out = {}
    rng = np.random.default_rng(0)
    for name, maker, nu in (("full", problems.full_standing_problem, 22), ("kino", problems.kino_standing_problem, 34), ("cent", problems.cent_standing_problem, 12)):
        prob = maker(batch=1, T=4)
        x = prob["xs"][0, 0].copy()
        xn = prob["xs"][0, 1].copy()
        if name != "cent":  # a generic (non-symmetric, moving) state so no term vanishes by accident
            x[7:29] += rng.normal(size=22) * 0.05
            x[29:] = rng.normal(size=28) * 0.1
        u = prob["us"][0, 0] + rng.normal(size=nu) * 1.0
        for variant, cs in (("ds", [True, True]), ("ss", [True, False])):
            kn = _abi.Knot.from_buffer_copy(prob["knots"][0])
            if name == "full":
                kn = problems.full_knot(cs, prob["lf"], prob["rf"], np.array(kn.f_ref[:6]), np.array(kn.f_ref[6:]))
            else:
                kn.cs[0], kn.cs[1] = float(cs[0]), float(cs[1])
                if name == "kino":
                    kn = problems.kino_knot(cs, prob["lf"], prob["rf"], np.array(kn.u_ref[:]))
            for deriv in (1, 0):
                cnt = (C.c_ulonglong * 5)()
                lib.orc_count_eval(C.byref(prob["robot"]), C.byref(prob["cfg"]), C.byref(kn), None, x.ctypes.data_as(dp), u.ctypes.data_as(dp),
                                   xn.ctypes.data_as(dp), deriv, cnt)
                add, mul, div, sq, trig = [int(c) for c in cnt]
                key = f"{name}_{variant}_{'deriv' if deriv else 'values'}"
                out[key] = add + mul + div + sq
                out[key + "_detail"] = {"add_sub": add, "mul": mul, "div": div, "sqrt": sq, "trig_calls_not_counted": trig}
        cnt = (C.c_ulonglong * 5)()
        lib.orc_count_eval(C.byref(prob["robot"]), C.byref(prob["cfg"]), None, C.byref(prob["terms"][0]), x.ctypes.data_as(dp), u.ctypes.data_as(dp),
                           xn.ctypes.data_as(dp), 1, cnt)
        out[f"{name}_term_deriv"] = int(sum(cnt[:4]))
    out["what"] = ("FLOPs (add/sub + mul + div + sqrt, one each; sin/cos/atan2 calls listed apart and not counted) of ONE knot evaluation of the CPU oracle "
                   "(oracle/knot.hpp eval_knot / eval_term) at a generic state, counted with the instrumented scalar of oracle/count_flops.cpp; "
                   "deriv = values + analytic Jacobians + Gauss-Newton Hessian (the k_eval<.,true> work), values = linesearch trial evaluation")
    path = os.path.join(ROOT, "profiles", "eval_flops.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        if not k.endswith("_detail") and k != "what":
            print(f"{k:24s} {v / 1e6:8.3f} MFLOP")


if __name__ == "__main__":
    main()
