// cport.cpp - multi-threaded CPU build of the physics step for bench.py's CPU baseline (TEST / MEASUREMENT INFRASTRUCTURE).
//
// The physics algorithm is written once, in mjlab_b200/csrc/b2_{smooth,solve}.cuh, against a small execution-model shim
// (b2_common.cuh): with -DB2_HOSTEMU the same source compiles as plain single-lane C++.  This file wraps that build in a
// thread pool over worlds so that the CPU baseline of bench.py measures a compiled, all-cores implementation of the SAME
// algorithm (float64), instead of the numpy oracle, whose interpreter overhead says nothing about CPUs.  It is not an
// independent checker - parity is checked against oracle/mjstep.py (numpy, written separately) - and nothing under
// mjlab_b200/ loads it.  Built by oracle/cport/Makefile into oracle/_build/libcport.so.
#include <atomic>
#include <thread>
#include <vector>

#define B2_HOSTEMU 1
#ifndef B2_REAL
#define B2_REAL double
#endif
#include "../../mjlab_b200/csrc/b2_smooth.cuh"
#include "../../mjlab_b200/csrc/b2_solve.cuh"

extern "C" {

int cport_sizeof_real(void) { return (int)sizeof(real); }

static int cport_run(const KModel* m, KData* d, int nstep, int nthread, int do_integrate);
// nstep physics steps for every world, worlds distributed over nthread threads (dynamic chunks of 1 world).
int cport_steps(const KModel* m, KData* d, int nstep, int nthread) { return cport_run(m, d, nstep, nthread, 1); }
// one forward pass (no integration) for every world: sim.forward() after resets (manager_based_rl_env.py:295-298)
int cport_forward(const KModel* m, KData* d, int nthread) { return cport_run(m, d, 1, nthread, 0); }
static int cport_run(const KModel* m, KData* d, int nstep, int nthread, int do_integrate) {
  SmoothSmem ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, 256);
  const SolveCaps caps = solve_caps(m->nlimited, m->nfriction, d->nconmax, d->njmax, m->maxchain);
  int arena = solve_arena_need(caps.rows, caps.runs, caps.dofs, caps.pool, d->nconmax > 0 ? d->nconmax : 1, m->cone);
  SolveSmem lv = solve_smem_layout(m->nv, caps.rows, m->maxchain, arena, 0, m->nfriction);
  PostSmem lp = post_smem_layout(m->nv);
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<real> a(ls.total), b(lv.fixed_total), ar(arena + 4), c(lp.total);
    int actl[2] = {0, 0};
    for (;;) {
      int w = next.fetch_add(1);
      if (w >= d->nworld) break;
      for (int s = 0; s < nstep; ++s) {
        smooth_world(*m, *d, ls, 256, w, a.data());
        solve_world(*m, *d, lv, w, b.data(), ar.data(), actl);
        post_world(*m, *d, lp, w, c.data(), do_integrate);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthread; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return 0;
}
}
