"""Long-horizon closed-loop regression at the REFERENCE settings: mu_init = 1e-8, ONE ProxDDP iteration per tick,
multipliers reset every tick (fulldynamic_talos.py:375,407,539), reference gait bookkeeping, ideal plant that integrates the
model.  Runs on the CPU oracle (tools/closed_loop_oracle.py); the CUDA path mirrors the oracle bit-for-tolerance (test_gpu_parity).

Known limitation, documented in DESIGN.md ("oracle-vs-Aligator ablation"): the loop degrades once the first landing knot enters the
horizon (tick ~ 110) and diverges during the second step; with the tick solved to convergence instead of one iteration the same
loop walks.  The test is therefore an expected failure — it turns into a pass the day the one-iteration loop is fixed."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.xfail(reason="one-iteration-per-tick loop at mu_init = 1e-8 diverges during the second step (DESIGN.md ablation table)", strict=False)
def test_one_iteration_loop_survives_260_ticks():
    import closed_loop_oracle

    hist = closed_loop_oracle.run(B=1, N=260, iters=1, keep=False, mu_init=1e-8, verbose=False, threads=2)
    last = hist[-1]
    assert last[0] == 260 and last[7] == 0 and 0.95 < last[8] and last[9] < 1.08  # all ticks done, nothing failed, base height sane
    assert last[4] > 0.5  # the linesearch still accepts (nearly) full steps


def test_first_hundred_ticks_track():
    """The standing / weight-shift part of the loop (ticks 1-60, single-support knots entering the horizon from tick 30): full steps
    accepted, the base stays put."""
    import closed_loop_oracle

    hist = closed_loop_oracle.run(B=1, N=60, iters=1, keep=False, mu_init=1e-8, verbose=False, threads=2)
    assert len(hist) == 60 and all(h[7] == 0 for h in hist)
    assert min(h[5] for h in hist) >= 0.5 and 1.0 < hist[-1][8] < 1.04
