"""Per-phase cycle counters of the batched QP kernel (instrumented build: make -C mpc_benchmark_b200/csrc qpphase, then
MPCB200_LIB=mpc_benchmark_b200/libmpcqp_phase.so python tools/qp_phase_timing.py).  Whole-body ID QPs of the bench's QP leg."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_benchmark_b200 import _native, pin, qp_utils  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = np.load(os.path.join(ROOT, "tests", "golden", "qp_id_talos.npz"))
reps = batch // 32
tile = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))  # noqa: E731
rng = np.random.default_rng(7)
cs = tile(d["cs"])
a = tile(d["a"]) + rng.normal(0, 0.05, (batch, 28))
f = tile(d["forces"]) + rng.normal(0, 1.0, (batch, 12)) * np.repeat(cs, 6, axis=1)
rbd = qp_utils.RBDTerms(nle=tile(d["nle"]), Jc=tile(d["Jc"]), dJv=tile(d["dJv"]), vf=tile(d["vf"]))
s = qp_utils.IDSolver_ulim(pin.load_talos_like()[0], [1, 1], 2, 0.8, 0.1, 0.075, [0, 1], 6, False, batch=batch)
for _ in range(3):
    s.solve(rbd, cs, None, a, f, tile(d["M"]))
L, h = _native.lib(), s.qp._handle()
out = np.zeros(8)
L.mpc_qp_debug_phases(h, _native.ptr(out))
s.solve(rbd, cs, None, a, f, tile(d["M"]))
ms = L.mpc_qp_last_device_ms(h)
L.mpc_qp_debug_phases(h, _native.ptr(out))
info = s.qp.results.info
names = ["load", "residuals", "AL gradient", "Newton matrix", "Cholesky", "substitutions", "linesearch", "multipliers+BCL"]
print(f"batch {batch}: kernel {ms:.3f} ms, mean outer {info.iter_ext.mean():.2f}, mean Newton {info.iter.mean():.2f}")
tot = out.sum()
for n_, c in zip(names, out):
    print(f"  {n_:18s} {c / batch:10.0f} cycles/QP  {100 * c / tot:5.1f} %")
print(f"  total              {tot / batch:10.0f} cycles/QP")
