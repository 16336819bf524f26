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

extern "C" {

int b2h_sizeof_real(void) { return (int)sizeof(real); }
int b2h_sizeof_model(void) { return (int)sizeof(KModel); }
int b2h_sizeof_data(void) { return (int)sizeof(KData); }

static int jpool_budget = 0;
void b2h_set_jpool_budget(int b) { jpool_budget = b; }

int b2h_run(const KModel* m, KData* d, int do_integrate, int world_lo, int world_hi) {
  SmoothSmem ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, 256);
  int rows_cap = 4 * d->nconmax + 2 * m->nlimited;
  if (rows_cap > d->njmax) rows_cap = d->njmax;
  int wpad = m->maxchain | 1;
  int jpool_cap = rows_cap * wpad;
  if (jpool_budget > 0 && jpool_cap > jpool_budget) jpool_cap = jpool_budget;
  SolveSmem lv = solve_smem_layout(m->nv, rows_cap, 2 * m->nlimited + d->nconmax + 1, d->nconmax > 0 ? d->nconmax : 1,
                                   jpool_cap, 2 * m->nlimited + d->nconmax * m->maxchain);
  std::vector<real> a(ls.total), b(lv.total);
  for (int w = world_lo; w < world_hi && w < d->nworld; ++w) {
    smooth_world(*m, *d, ls, 256, w, a.data());
    solve_world(*m, *d, lv, w, b.data(), do_integrate);
  }
  return 0;
}

int b2h_smooth(const KModel* m, KData* d) {
  SmoothSmem ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, 256);
  std::vector<real> a(ls.total);
  for (int w = 0; w < d->nworld; ++w) smooth_world(*m, *d, ls, 256, w, a.data());
  return 0;
}
}
