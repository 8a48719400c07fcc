"""Batched closed-loop MPC driver on the device (the loop of fulldynamic_talos.py:438-550).

Every tick: append the next stage of each instance's gait schedule at the end of the horizon, shift the previous solution by
one knot as warm start, take the new initial state from the plant and run ONE ProxDDP iteration.  The plant is either
ideal (the model prediction xs[1]) or any host callback returning measured states.  PyBullet (bullet_robot.py) is out of
scope; the reference's 1 kHz Riccati feedback (fulldynamic_talos.py:522) is `BatchSolver.feedback(0)`.
"""
import time

import numpy as np

from . import _abi


class ClosedLoop:
    def __init__(self, solver, knot_stream, plant=None, keep_multipliers=False, term_stream=None, phase_matched_tail=None):
        """solver: a set-up BatchSolver holding the cold-solve result; knot_stream(t) -> ctypes Knot array [batch] with the stage
        entering the horizon at tick t; plant(t, xs, us, K0) -> measured states [batch, nx] or None for the ideal plant;
        term_stream(t) -> ctypes Term array [batch] or None: the terminal cost references / CoM equality swapped in at tick t
        (fulldynamic_talos.py:499-510); phase_matched_tail: True / False selects the control warm start of the knot appended every tick (the nearest
        knot of the same contact phase / the previous knot as in the scripts; BatchSolver.set_tail_warmstart), None leaves the solver's setting."""
        self.solver, self.knot_stream, self.plant, self.keep, self.term_stream = solver, knot_stream, plant, keep_multipliers, term_stream
        if phase_matched_tail is not None:
            solver.set_tail_warmstart(phase_matched_tail)
        self.t = 0
        self.tick_ms = []

    def step(self, max_iters=1):
        s = self.solver
        x_meas = None
        if self.plant is not None:
            r = s.results(gains=False, multipliers=False)
            x_meas = self.plant(self.t, r.xs, r.us, s.feedback(0))
        t0 = time.perf_counter()
        if self.term_stream is not None:
            terms = self.term_stream(self.t)
            if terms is not None:
                s.update_terms(terms)
        s.tick(self.knot_stream(self.t), x_meas, keep_multipliers=self.keep, max_iters=max_iters)
        self.tick_ms.append(1e3 * (time.perf_counter() - t0))
        self.t += 1

    def run(self, n_ticks, max_iters=1):
        for _ in range(n_ticks):
            self.step(max_iters)
        return self.solver.results(gains=False, multipliers=False)


def standing_stream(prob):
    """Knot stream that keeps appending the problem's last stage (double-support standing)."""
    T = prob["cfg"].T
    B = prob["x0"].shape[0]
    last = (_abi.Knot * B)(*[prob["knots"][b * T + T - 1] for b in range(B)])
    return lambda t: last
