// Host-side launch sequence of one solver.run() over a batch (shared by solver.cu and the test-only host
// emulation).  Mirrors the loop structure of aligator::SolverProxDDPTpl::run (fulldynamic_talos.py:393-397).
#pragma once
#include "solver_core.cuh"

namespace mpcdev {

// Backend: eval(bool deriv), decide_eval(), riccati(), apply_step(), decide_ls(), read_counters(int[2])
template <class Backend> int run_loop(Backend &be, int max_iters, const SolverConst &sc) {
  int launches = 0;
  const int guard = max_iters + sc.max_al_iters + 2;
  for (int pass = 0; pass < guard; pass++) {
    be.eval(true); be.decide_eval(); be.riccati();
    launches += 3;
    int c[2] = {0, 0};
    for (int ls = 0; ls <= sc.ls_max_steps; ls++) {
      be.apply_step(); be.eval(false); be.decide_ls();
      launches += 3;
      be.read_counters(c);
      if (c[0] == 0) break;
    }
    if (c[1] == 0) break;
  }
  return launches;
}

} // namespace mpcdev
