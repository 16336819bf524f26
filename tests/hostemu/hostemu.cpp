// hostemu.cpp - TEST-ONLY single-lane host build of the device code.
//
// Compiles mjlab_b200/csrc/b2_{smooth,solve}.cuh with -DB2_HOSTEMU (one lane,
// no barriers) so the kernel logic can be checked against the oracle in a
// container without a GPU.  Not part of the product: nothing in mjlab_b200/
// loads this library, and the product raises if libb200sim.so is missing.
// B2_REAL selects float (bit-for-bit the arithmetic width of the device code)
// or double (to separate rounding effects from logic errors).
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../mjlab_b200/csrc/b2_smooth.cuh"
#include "../../mjlab_b200/csrc/b2_solve.cuh"
#include "../../mjlab_b200/csrc/b2_env.cuh"
#include "../../mjlab_b200/csrc/b2_trk.cuh"

extern "C" {

int b2h_sizeof_real(void) { return (int)sizeof(real); }
long long b2h_ls_pass_count(void) { return b2h_ls_passes; }
void b2h_flip_histogram(long long* out) { for (int i = 0; i < 66; ++i) out[i] = b2h_flip_hist[i]; }
int b2h_sizeof_model(void) { return (int)sizeof(KModel); }
int b2h_sizeof_data(void) { return (int)sizeof(KData); }

static int jpool_budget = 0;
void b2h_set_jpool_budget(int b) { jpool_budget = b; }

static int memo_consume = 0;
// 1: b2h_run lets the position / velocity stage reuse what a preceding pass left in the data arrays (smooth_world's memo)
void b2h_set_memo_consume(int v) { memo_consume = v; }
long long b2h_memo_hit_count(void) { return b2h_memo_hits; }

int b2h_run(const KModel* m, KData* d, int do_integrate, int world_lo, int world_hi) {
  SmoothSmem ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, 256);
  const SolveCaps caps = solve_caps(m->nlimited, m->nfriction, d->nconmax, d->njmax, m->maxchain);
  int arena = solve_arena_need(caps.rows, caps.runs, caps.dofs, caps.pool, d->nconmax > 0 ? d->nconmax : 1, m->cone);
  if (jpool_budget > 0) arena = jpool_budget;  // tests of the overflow path: an arena smaller than the worst case
  SolveSmem lv = solve_smem_layout(m->nv, caps.rows, m->maxchain, arena, 0, m->nfriction);
  PostSmem lp = post_smem_layout(m->nv);
  std::vector<real> a(ls.total), b(lv.fixed_total), ar(arena + 4), c(lp.total);
  int actl[2] = {0, 0};
  for (int w = world_lo; w < world_hi && w < d->nworld; ++w) {
    smooth_world(*m, *d, ls, 256, w, a.data(), memo_consume);
    solve_world(*m, *d, lv, w, b.data(), ar.data(), actl);
    post_world(*m, *d, lp, w, c.data(), do_integrate);
  }
  return 0;
}

int b2h_smooth(const KModel* m, KData* d) {
  SmoothSmem ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, 256);
  std::vector<real> a(ls.total);
  for (int w = 0; w < d->nworld; ++w) smooth_world(*m, *d, ls, 256, w, a.data());
  return 0;
}

// ---- env layer (b2_env.cuh), one call per stage over all envs ----
int b2h_sizeof_env(void) { return (int)sizeof(KEnv); }
int b2h_env_stage(const KModel* m, KData* d, const KEnv* e, int stage) {
  for (int w = 0; w < d->nworld; ++w) {
    switch (stage) {
      case 0: env_action_world(*m, *d, *e, w); break;
      case 1: env_airtime_world(*m, *d, *e, w); break;
      case 2: env_post_world(*m, *d, *e, w); break;
      case 3: env_obs_world(*m, *d, *e, w, 1); break;
      default: return 1;
    }
  }
  if (stage == 2 || stage == 3) {  // k_env_log
    real* log_env = reinterpret_cast<real*>(e->buf[B2E_B_LOG_ENV]);
    real* log = reinterpret_cast<real*>(e->buf[B2E_B_LOG]);
    for (int c = 0; c < B2E_NLOG; ++c) {
      real acc = 0;
      for (int w = 0; w < d->nworld; ++w) acc += log_env[(size_t)w * B2E_NLOG + c];
      if (stage == 3 || c == B2E_LOG_NRESET) log[c] = acc;
    }
    if (stage == 2) reinterpret_cast<int*>(e->buf[B2E_B_ANY_RESET])[0] = (int)log[B2E_LOG_NRESET];
    if (stage == 3) reinterpret_cast<long long*>(e->buf[B2E_B_STEP_COUNTER])[0] += 1;
  }
  return 0;
}
// masked forward: worlds with mask[w] == 0 are skipped
int b2h_forward_masked(const KModel* m, KData* d, const unsigned char* mask) {
  for (int w = 0; w < d->nworld; ++w)
    if (!mask || mask[w]) b2h_run(m, d, 0, w, w + 1);
  return 0;
}
float b2h_dc_clip(float effort, float vel, float limit, float sat, float vlim) { return b2_dc_clip(effort, vel, limit, sat, vlim); }

// ---- tracking env (b2_trk.cuh) ----
int b2h_sizeof_trk(void) { return (int)sizeof(KTrk); }
int b2h_trk_stage(const KModel* m, KData* d, const KTrk* e, int stage) {
  for (int w = 0; w < d->nworld; ++w) {
    switch (stage) {
      case 0: trk_action_world(*m, *d, *e, w); break;
      case 2: trk_post_world(*m, *d, *e, w); break;
      case 3: trk_obs_world(*m, *d, *e, w, 1); break;
      default: return 1;
    }
  }
  if (stage == 2 || stage == 3) {
    real* log_env = reinterpret_cast<real*>(e->buf[B2T_B_LOG_ENV]);
    real* log = reinterpret_cast<real*>(e->buf[B2T_B_LOG]);
    for (int c = 0; c < B2T_NLOG; ++c) {
      real acc = 0;
      for (int w = 0; w < d->nworld; ++w) acc += log_env[(size_t)w * B2T_NLOG + c];
      if (stage == 3 || c == B2T_LOG_NRESET) log[c] = acc;
    }
    if (stage == 2) reinterpret_cast<int*>(e->buf[B2T_B_ANY_RESET])[0] = (int)log[B2T_LOG_NRESET];
    if (stage == 3) {
      reinterpret_cast<long long*>(e->buf[B2T_B_STEP_COUNTER])[0] += 1;
      if (e->ip[B2T_IP_SAMPLING_MODE] == 2) {
        real* bfc = reinterpret_cast<real*>(e->buf[B2T_B_BIN_FAILED_COUNT]);
        real* cur = reinterpret_cast<real*>(e->buf[B2T_B_CURRENT_BIN_FAILED]);
        const real alpha = (real)e->fp[B2T_FP_ADAPTIVE_ALPHA];
        for (int b = 0; b < e->ip[B2T_IP_BIN_COUNT]; ++b) { bfc[b] = alpha * cur[b] + (1 - alpha) * bfc[b]; cur[b] = 0; }
      }
    }
  }
  return 0;
}
}
