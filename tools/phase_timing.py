"""In-kernel phase cycle counters of k_riccati and k_eval (instrumented build `make -C mpc_benchmark_b200/csrc phase`).
Usage (GPU box):  MPCB200_LIB=mpc_benchmark_b200/libmpcb200_phase.so python tools/phase_timing.py [batch]
Prints the cycles the LAST-written CTA spent in each phase; the numbers guide optimisation only, they are not bench values."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_benchmark_b200 import _native, problems  # noqa: E402
from mpc_benchmark_b200.batch import BatchSolver  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    prob = problems.full_walk_batch(batch, seed=5)
    s = BatchSolver(prob["robot"], prob["cfg"], batch, device=0)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    warm = s.run(prob["xs"], prob["us"], max_iters=3, gains=False)
    s.setup(prob["knots"], prob["terms"], prob["x0"])
    s.run(warm.xs, warm.us, max_iters=1, gains=False)
    out = (C.c_double * 64)()
    _native.check(_native.lib().mpc_debug_phases(s._h, out), "mpc_debug_phases")
    ph = np.array(out[:])
    print("riccati phases (cycles):", [int(v) for v in ph[:16]], "sum", int(ph[:16].sum()))
    print("eval<deriv>  phases (cycles):", [int(v) for v in ph[16:32]], "sum", int(ph[16:32].sum()))
    print("eval<values> phases (cycles):", [int(v) for v in ph[32:48]], "sum", int(ph[32:48].sum()))
    print("riccati sub-phases 16.. (cycles):", [int(v) for v in ph[48:]], "sum", int(ph[48:].sum()))
    print("kernel ms:", s.kernel_ms())


if __name__ == "__main__":
    main()
