"""Long-horizon closed-loop regression at the REFERENCE settings: mu_init = 1e-8, ONE ProxDDP iteration per tick,
multipliers reset every tick (fulldynamic_talos.py:375,407,539), reference gait bookkeeping, ideal plant that integrates the
model.  Runs on the CPU oracle (tools/closed_loop_oracle.py); the CUDA path mirrors the oracle bit-for-tolerance (test_gpu_parity).

Documented in DESIGN.md section 7: WITH THE SCRIPTS' WARM START of the appended knot (us[1:] + [us[-1]]) the full-dynamics loop degrades once
the first landing knot enters the horizon (tick ~ 110) and diverges during the second step — that test stays here as an expected failure —
while with the appended knot started from the control of the nearest knot of the same contact phase (mpc_set_tail_warmstart(1)) the same loop
walks the whole gait, as do the centroidal and kinodynamic loops with the scripts' warm start (the passing tests below)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.xfail(reason="full dynamics with the scripts' warm start of the appended knot: diverges during the second step (DESIGN.md section 7)", strict=False)
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


def test_centroidal_loop_walks_the_whole_gait():
    """BASELINE configs[0] in closed loop at the reference's settings (mu_init = 1e-8, one iteration per tick, multipliers reset:
    centroidal_talos.py:270-277,298,461): 20 DS / 80 left / 20 DS / 80 right / ... for 450 ticks, ideal plant.  Full steps all the way,
    the CoM height holds, the stance foot carries the weight (profiles/r2_closed_loop_cent.txt)."""
    import closed_loop_oracle_cent

    hist = closed_loop_oracle_cent.run(B=2, N=450, iters=1, mu_init=1e-8, verbose=False, threads=2)
    assert len(hist) == 450
    assert all(0.89 < h[5] and h[6] < 0.92 for h in hist)  # CoM height
    # full steps except for a few ticks right after a contact switch reaches knot 0 (ticks 120-125, 202-208), from which the loop recovers
    assert sum(1 for h in hist if h[3] < 0.99) <= 25 and min(h[3] for h in hist) >= 0.01 and all(h[3] == 1.0 for h in hist[-150:])
    assert max(h[1] for h in hist) < 15.0  # primal infeasibility bounded (N / N m on the cone rows)


def test_kinodynamic_loop_walks_through_the_first_step():
    """BASELINE configs[1] in closed loop at the reference's settings (one iteration per tick, mu_init = 1e-8), multipliers reset:
    the first single-support phase arrives at the front of the horizon at tick 120 and leaves at tick 200 — exactly where the
    full-dynamics loop degrades; the kinodynamic loop keeps taking full steps (the whole 840-tick gait: profiles/r2_closed_loop_kino_reset.txt)."""
    import closed_loop_oracle
    from mpc_benchmark_b200 import _abi

    hist = closed_loop_oracle.run(B=1, N=215, iters=1, keep=False, mu_init=1e-8, verbose=False, threads=2, kind=_abi.KIND_KINO)
    assert len(hist) == 215 and all(h[7] == 0 for h in hist)
    # a few short backtracking bursts when a contact switch enters / leaves the horizon (ticks 19-26, 101-109, 120-127, 202), each
    # followed by full steps again; nothing fails, the infeasibility stays bounded and the base stays at its height
    assert sum(1 for h in hist if h[5] < 0.99) <= 40 and hist[-1][5] == 1.0 and sum(1 for h in hist[-50:] if h[5] < 0.99) <= 2
    assert max(h[2] for h in hist) < 10.0 and 1.0 < hist[-1][8] < 1.03


def test_kinodynamic_loop_with_kept_multipliers():
    """kinodynamic_talos.py:488 cycles the problem WITHOUT solver.setup: the multipliers are kept and shifted.  With the shift of
    mpc_shift_multipliers (running knots / co-states move, terminal multiplier and lams[0] stay, last entry repeated) the loop is healthy;
    round 1's shift of all T + 1 slots with zero fill diverges within 40 ticks (profiles/r2_closed_loop_kino_keep.txt)."""
    import closed_loop_oracle
    from mpc_benchmark_b200 import _abi

    hist = closed_loop_oracle.run(B=1, N=60, iters=1, keep=True, mu_init=1e-8, verbose=False, threads=2, kind=_abi.KIND_KINO)
    assert len(hist) == 60 and all(h[7] == 0 for h in hist)
    assert max(h[2] for h in hist) < 0.1 and sum(1 for h in hist if h[5] < 0.99) <= 12 and all(h[5] == 1.0 for h in hist[-20:])


def test_full_dynamics_loop_walks_with_phase_matched_tail(monkeypatch):
    """The full-dynamics loop at the reference's settings (mu_init = 1e-8, one iteration per tick, multipliers reset) WALKS once the knot
    appended every tick starts from the control of the nearest knot with the same contact phase (mpc_set_tail_warmstart(1)) instead of the
    previous knot's (the scripts' us[1:] + [us[-1]]): through the first landing (tick 110) and the first two swing phases, full steps at
    every tick, where the expected-failure test above has diverged by tick 247.  1000 ticks: profiles/r2_closed_loop_full_phase_warmstart.txt;
    256 robots on the GPU: profiles/r2_gait_walk_full_gpu.txt."""
    import closed_loop_oracle

    monkeypatch.setenv("ORC_TAIL_U", "phase")
    hist = closed_loop_oracle.run(B=1, N=260, iters=1, keep=False, mu_init=1e-8, verbose=False, threads=2)
    assert len(hist) == 260 and all(h[7] == 0 for h in hist)
    # (the primal infeasibility is reported BEFORE the step: a knot freshly appended at a contact switch shows up with 10 - 20 N on the cone rows once)
    assert sum(1 for h in hist if h[5] < 0.99) <= 5 and max(h[2] for h in hist) < 40.0 and hist[-1][2] < 5.0
    assert 1.0 < hist[-1][8] < 1.04
