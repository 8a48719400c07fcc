// TEST-ONLY host emulation of the CUDA kernel sources (mpc_benchmark_b200/csrc/*.cuh compiled with
// -DMPC_HOST_EMU: PAR_FOR = serial loop, SYNC = no-op).  Lets the CPU test-suite check kernel LOGIC against the
// oracle without a GPU.  It is never loaded by the product package and is not a fallback.
#define MPC_HOST_EMU 1
#include "../../mpc_benchmark_b200/csrc/driver.hpp"
#include "../../mpc_benchmark_b200/csrc/ws_alloc.hpp"
#include "../../mpc_benchmark_b200/csrc/qp.cuh"
#include "../../mpc_benchmark_b200/csrc/rbd_terms.cuh"
#include "../../mpc_benchmark_b200/csrc/gait_host.hpp"
#include <cstdlib>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace mpcdev;

struct EmuBackend {
  Ws w;
  std::vector<double> smem;
  EmuBackend() : smem(40000, 0.0) {}
  void reset_counters() { for (int i = 0; i < 4; i++) w.counters[i] = 0; }
  void eval(bool d, const int32_t *list, int n) {
    for (int i = 0; i < n; i++)
      for (int k = 0; k <= w.T; k++) {
        switch (w.kind * 2 + (d ? 1 : 0)) {
        case MPC_KIND_FULL * 2 + 1: eval_dispatch<MPC_KIND_FULL, true>(w, list[i], k, smem.data()); break;
        case MPC_KIND_FULL * 2 + 0: eval_dispatch<MPC_KIND_FULL, false>(w, list[i], k, smem.data()); break;
        case MPC_KIND_KINO * 2 + 1: eval_dispatch<MPC_KIND_KINO, true>(w, list[i], k, smem.data()); break;
        case MPC_KIND_KINO * 2 + 0: eval_dispatch<MPC_KIND_KINO, false>(w, list[i], k, smem.data()); break;
        case MPC_KIND_CENT * 2 + 1: eval_dispatch<MPC_KIND_CENT, true>(w, list[i], k, smem.data()); break;
        default: eval_dispatch<MPC_KIND_CENT, false>(w, list[i], k, smem.data()); break;
        }
      }
  }
  void decide_eval(const int32_t *list, int n, int32_t *next_eval) { double red[8]; for (int i = 0; i < n; i++) mpcdev::decide_eval(w, list[i], red, next_eval); }
  void riccati(const int32_t *list, int n) {
    for (int i = 0; i < n; i++) {
      if (w.kind == MPC_KIND_FULL) riccati_dispatch<MPC_KIND_FULL>(w, list[i], smem.data());
      else if (w.kind == MPC_KIND_KINO) riccati_dispatch<MPC_KIND_KINO>(w, list[i], smem.data());
      else riccati_dispatch<MPC_KIND_CENT>(w, list[i], smem.data());
    }
  }
  void apply_step(const int32_t *list, int n) { w.counters[0] = 0; for (int i = 0; i < n; i++) mpcdev::apply_step(w, list[i]); }
  void decide_ls(const int32_t *list, int n, int32_t *ls_out, int32_t *next_eval) {
    double red[8];
    for (int i = 0; i < n; i++) mpcdev::decide_ls(w, list[i], red, ls_out, next_eval);
  }
  void rollout_ls(const int32_t *list, int n, int32_t *next_eval) {
    double red[8], xv[256];
    for (int i = 0; i < n; i++) {
      if (w.kind == MPC_KIND_FULL) rollout_linesearch<MPC_KIND_FULL>(w, list[i], smem.data(), xv, red, next_eval);
      else if (w.kind == MPC_KIND_KINO) rollout_linesearch<MPC_KIND_KINO>(w, list[i], smem.data(), xv, red, next_eval);
      else rollout_linesearch<MPC_KIND_CENT>(w, list[i], smem.data(), xv, red, next_eval);
    }
  }
  void read_counters(int *c) { for (int i = 0; i < 4; i++) c[i] = w.counters[i]; }
};

extern "C" int emu_solve(const mpc_robot_t *rb, const mpc_config_t *cfg, int batch, const mpc_knot_t *knots, const mpc_term_t *terms,
                         const double *x0, double *xs, double *us, double *K, double *vs, double *lams, mpc_info_t *info, double *stage0,
                         int max_iters, double *lq_dump /* optional: AB,H,g of instance 0 after the first derivative pass */) {
  static_assert(sizeof(FullWsT<true>) <= 40000 * 8 && sizeof(KinoWsT<true>) <= 40000 * 8, "smem");
  DevModel *model = new DevModel;
  const char *err = nullptr;
  if (build_dev_model(rb, cfg, model, &err)) return 1;
  EmuBackend be;
  Ws &w = be.w;
  std::memset(&w, 0, sizeof w);
  w.B = batch; w.T = cfg->T; w.kind = cfg->kind;
  dims_of_kind(cfg->kind, w.nx, w.n, w.m, w.nc);
  w.nz = w.n + w.m; w.model = model; w.sc = default_consts(cfg->tol, cfg->mu_init, cfg->rollout);
  std::vector<void *> allocs;
  alloc_ws(w, [&](size_t bytes) { void *p = calloc(bytes ? bytes : 8, 1); allocs.push_back(p); return p; });
  const size_t T1 = w.T + 1;
  std::memcpy(w.knots, knots, sizeof(mpc_knot_t) * batch * w.T);
  std::memcpy(w.terms, terms, sizeof(mpc_term_t) * batch);
  std::memcpy(w.x0, x0, 8 * batch * w.nx);
  if (vs) std::memcpy(w.vs, vs, 8 * batch * T1 * w.nc);
  if (lams) std::memcpy(w.lams, lams, 8 * batch * T1 * w.n);
  for (int b = 0; b < batch; b++) init_instance(w, b, xs, us, max_iters);
  if (lq_dump) { // single derivative pass, dump instance 0
    be.reset_counters(); be.eval(true, eval_list(w, 0), w.B); be.decide_eval(eval_list(w, 0), w.B, eval_list(w, 1));
    size_t o = 0;
    auto put = [&](const double *p, size_t n) { std::memcpy(lq_dump + o, p, 8 * n); o += n; };
    put(w.AB, (size_t)w.T * w.n * w.nz); put(w.H, T1 * w.nz * w.nz); put(w.g, T1 * w.nz); put(w.gap, (size_t)w.T * w.n); put(w.h, T1 * w.nc);
    put(w.scal, T1 * SC_COUNT);
  } else run_loop(be, w, max_iters, w.sc);
  if (getenv("EMU_NCA_HIST")) { // test tooling: active-row counts per knot of the last pass (sizes the kernel's shared-memory fast path)
    for (int b = 0; b < batch; b++) { fprintf(stderr, "nca[%d]:", b); for (size_t k = 0; k < T1; k++) fprintf(stderr, " %d", w.nca[b * T1 + k]); fprintf(stderr, "\n"); }
  }
  std::memcpy(xs, w.xs, 8 * batch * T1 * w.nx);
  std::memcpy(us, w.us, 8 * batch * w.T * w.m);
  if (K) std::memcpy(K, w.Kfb, 8 * (size_t)batch * w.T * w.m * w.n);
  if (vs) std::memcpy(vs, w.vs, 8 * batch * T1 * w.nc);
  if (lams) std::memcpy(lams, w.lams, 8 * batch * T1 * w.n);
  for (int b = 0; b < batch; b++) {
    const InstState &s = w.st[b];
    if (info) { mpc_info_t &o = info[b]; o.prim_infeas = s.prim_infeas; o.dual_infeas = s.dual_infeas; o.traj_cost = s.traj_cost; o.merit = s.merit; o.mu = s.mu; o.alpha = s.alpha; o.ls_evals = s.ls_evals; o.pad_ = 0;
      o.num_iters = s.num_iters; o.al_iters = s.al_iters; o.conv = s.conv; o.status = s.status; }
    if (stage0) { std::memcpy(stage0 + b * 68, w.xdot + b * T1 * 56, 8 * 56); std::memcpy(stage0 + b * 68 + 56, w.lamc + b * T1 * 12, 8 * 12); }
  }
  for (void *p : allocs) free(p);
  delete model;
  return 0;
}

// ---- batched dense QP (qp.cuh) run serially: same source as k_qp_solve / k_qp_assemble_id
extern "C" int emu_qp_solve(int n, int ne, int ni, int box, int batch, const mpc_qp_settings_t *st, const double *H, long sH, const double *g, long sg,
                            const double *A, long sA, const double *b, long sb, const double *C, long sC, const double *l, long sl, const double *u, long su,
                            const double *lb, long slb, const double *ub, long sub, double *x, double *y, double *z, mpc_qp_info_t *info) {
  QPArgs P;
  P.n = n; P.ne = ne; P.ni = ni; P.box = box; P.batch = batch; P.st = *st;
  P.H = H; P.sH = sH; P.g = g; P.sg = sg; P.A = A; P.sA = sA; P.b = b; P.sb = sb; P.C = C; P.sC = sC; P.l = l; P.sl = sl; P.u = u; P.su = su;
  P.lb = lb; P.slb = slb; P.ub = ub; P.sub = sub; P.x = x; P.y = y; P.z = z; P.info = info;
  std::vector<double> smem(qp_smem_doubles(n, ne, ni, box), 0.0);
  for (int i = 0; i < batch; i++) qp_solve_group(P, i, smem.data());
  return 0;
}
extern "C" void emu_qp_assemble_id(int batch, const double *M, const double *nle, const double *Jc, const double *gamma, const double *a, const double *forces,
                                   const int32_t *cs, double mu, double L, double W, double *A, double *b, double *C, double *l) {
  for (int i = 0; i < batch; i++)
    qp_assemble_id_group(M + i * 784, nle + i * 28, Jc + i * 336, gamma + i * 12, a + i * 28, forces + i * 12, cs + i * 2, mu, L, W, A + i * 2480, b + i * 40,
                         C + i * 1116, l + i * 18);
}

// ---- rigid-body terms in front of the whole-body QPs (rbd_terms.cuh), serial
extern "C" int emu_rbd_terms(const mpc_robot_t *rb, const mpc_config_t *cfg, int count, const double *x, double *M, double *nle, double *Jc, double *dJv, double *vf) {
  DevModel *model = new DevModel;
  const char *err = nullptr;
  if (build_dev_model(rb, cfg, model, &err)) { delete model; return 1; }
  std::vector<double> smem(sizeof(FullWsT<false>) / 8 + 8, 0.0);
  FullWsT<false> &w = *reinterpret_cast<FullWsT<false> *>(smem.data());
  for (int i = 0; i < count; i++)
    rbd_terms_group(*model, x + (size_t)i * 57, w, M + (size_t)i * 784, nle + (size_t)i * 28, Jc + (size_t)i * 336, dJv + (size_t)i * 12, vf + (size_t)i * 12);
  delete model;
  return 0;
}

// ---- device gait generator (gait.cuh) run serially: `ticks` ticks of `batch` robots; lf / rf [ticks][batch][12] measured sole placements;
// knots_out [ticks][batch][T], terms_out [ticks][batch]
extern "C" int emu_gait(int kind, int T, const mpc_gait_t *gait, int batch, const int32_t *mirror, const double *urefs, int ticks, const double *lf,
                        const double *rf, mpc_knot_t *knots_out, mpc_term_t *terms_out) {
  GaitCfg g;
  std::memset(&g, 0, sizeof g);
  gait_fill_cfg(*gait, kind, T, g);
  std::vector<int8_t> phases;
  gait_build_schedule(*gait, T, phases, g);
  g.phases = phases.data();
  g.urefs = urefs;
  std::vector<GaitRobot> rs(batch);
  for (int b = 0; b < batch; b++) {
    for (int i = 0; i < 12; i++) { rs[b].start_l[i] = rs[b].final_l[i] = rs[b].next_l[i] = gait->lf0[i]; rs[b].start_r[i] = rs[b].final_r[i] = rs[b].next_r[i] = gait->rf0[i]; }
    rs[b].mirror = mirror[b] != 0; rs[b].pad_ = 0;
  }
  double sm[16];
  for (int t = 0; t < ticks; t++)
    for (int b = 0; b < batch; b++)
      gait_tick_group(g, rs[b], t, lf + ((size_t)t * batch + b) * 12, rf + ((size_t)t * batch + b) * 12, knots_out + ((size_t)t * batch + b) * T,
                      terms_out + (size_t)t * batch + b, sm);
  return 0;
}
