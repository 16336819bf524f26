// b2_api.cu - C ABI (include/b200sim.h) + kernel launches for sm_100a.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// Launch shape: one warp per world, B2_WPB worlds per CTA; grid = ceil(nworld / B2_WPB).
// Dynamic shared memory = B2_WPB * per-world slab (layout computed on the host from the model sizes).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2_smooth.cuh"
#include "b2_solve.cuh"
#include "b2_env.cuh"
#include "b2_trk.cuh"

static_assert(sizeof(KModel) == sizeof(b2_model), "KModel must mirror b2_model");
static_assert(sizeof(KData) == sizeof(b2_data), "KData must mirror b2_data");
static_assert(sizeof(KEnv) == sizeof(b2_env), "KEnv must mirror b2_env");

#ifndef B2_SMOOTH_WPB
#define B2_SMOOTH_WPB 1   // worlds per CTA of k_smooth (finer scheduling granularity: -3 % vs 2)
#endif
#ifndef B2_WPB
#define B2_WPB 2
#endif
#define B2_NCAND_MAX 256
#ifndef B2_SOLVE_MINB
#define B2_SOLVE_MINB 8
#endif
#ifndef B2_ORDER_MODE_DEFAULT
#define B2_ORDER_MODE_DEFAULT 0
#endif

// Error text and the launch counter are per host thread (one host thread drives one GPU, sim.py one process per GPU; a
// process that drives several devices from several threads keeps their errors apart).
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#ifndef B2_SMOOTH_MINB
#define B2_SMOOTH_MINB 28   // worlds resident per SM the compiler must allow for (register cap 65536 / (MINB * 32))
#endif
__global__ void __launch_bounds__(B2_SMOOTH_WPB * 32, B2_SMOOTH_MINB) k_smooth(const KModel m, const KData d, const SmoothSmem L, int ncand_max, int memo_consume, const unsigned char* mask, const int* gate, int w0, int w1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* smem = reinterpret_cast<real*>(smem_raw);
  int warp = threadIdx.x >> 5;
  int w = w0 + blockIdx.x * B2_SMOOTH_WPB + warp;   // this launch covers the worlds [w0, w1)
  if (w >= w1) return;
  if (gate && *gate == 0) return;
  if (mask && !mask[w]) return;
  smooth_world(m, d, L, ncand_max, w, smem + (size_t)warp * L.total, memo_consume);
}

#ifdef B2_TIMING
// diagnostics build only (tools/solve_timeline.py): per-world start / end of the solver in globaltimer ns + the SM id
__device__ unsigned long long g_tim[3 * 16384];
extern "C" int b2_debug_lsstat(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_lsstat, sizeof(unsigned long long) * 8); }
extern "C" int b2_debug_phases(unsigned long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 8 * n); }
extern "C" int b2_debug_smooth_phases(unsigned long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, g_sphase, sizeof(unsigned long long) * 8 * n); }
__device__ __forceinline__ unsigned long long b2_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned b2_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
extern "C" int b2_debug_timing(unsigned long long* out, int n) { return (int)cudaMemcpyFromSymbol(out, g_tim, sizeof(unsigned long long) * 3 * n); }
#endif

// Number of solver CTAs for n worlds: all resident at once (the CTAs per SM the occupancy calculator grants, times the
// device's SM count - both queried in b2_check), each warp then pulls further worlds.  `cap` = that product.
#define B2_SOLVE_DEFAULT_CTAS (148 * B2_SOLVE_MINB)   // B200: 148 SMs; used until b2_check has queried the device
static __host__ __device__ inline int solve_grid(int n, int cap) {
  int g = (n + B2_WPB - 1) / B2_WPB;
  return g < cap ? g : cap;
}

// Rank p in the heaviest-first order -> solver slot (a permutation of [0, n) for every n; tests/test_abi.py checks it
// through b2_debug_solver_slot).  Light ranks start at L0 = max(G, n - G): with n < 2G (small or odd batches) every rank past
// the G heavy ones is "light".  The lightest world of all sits next to rank G-1, not next to rank 0: the first warps to come
// free then have a small CTA-mate, and the medium-heavy world they pull fits beside it.
static __host__ __device__ inline int solve_slot(int p, int n, int cap) {
  const int G = solve_grid(n, cap);
  const int L0 = G > n - G ? G : n - G;
  return p < G ? 2 * p : (p >= L0 ? 2 * (p - L0) + 1 : 2 * G + (p - G));
}

// Persistent warps: warp (CTA c, index i) starts on solver slot B2_WPB*c + i, then pulls the next unclaimed slot from the
// launch's work counter until the list is empty, so a warp-slot never idles while worlds are waiting (with one world per
// warp and two per CTA, the slot of a light world stayed blocked until its heavy CTA-mate was done).
//
// post_mode != 0 (1: forward, 2: step with integration; EXPERIMENT, B200SIM_FUSE_POST=1, measured slower - 0.67 ms against
// 0.42 + 0.10 ms for the two launches, profiles/r02_fused_post_experiment.txt: the sensor code evicts the Newton loop from
// the instruction caches of exactly the warps that set the launch's duration): the sensor / integration stage runs INSIDE
// this launch.  The span of
// the solver is the latency of its heaviest worlds and for its last ~40 % most warp slots are empty (per-world timeline,
// profiles/r02_timeline.txt), while the old separate k_post launch was one latency-bound wave of its own.  Every solved world
// is published in a ring (B2_DI_SOLVE_RING, release / acquire through __threadfence); a warp that finds the solver queue
// empty claims ring entries in order and runs post_world for them - on worlds that finished long ago while the heavy ones
// are still iterating - so the sensor stage fills the solver's tail instead of following it.  A claimed entry is waited
// for only when all earlier-finished worlds are already taken, i.e. at the very end of the launch.  All CTAs are resident
// (grid <= SMs x occupancy), so every claimed entry is eventually produced.
template <bool CG, bool FR, bool EL = false>
__global__ void __launch_bounds__(B2_WPB * 32, B2_SOLVE_MINB) k_solve_t(const KModel m, const KData d, const SolveSmem L, const PostSmem LP, int post_mode, const unsigned char* mask, const int* gate, int w0, int w1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* smem = reinterpret_cast<real*>(smem_raw);
  // [fixed part of world 0 | ... | fixed part of world B2_WPB-1 | row arena | arena control word]
  real* arena = smem + (size_t)B2_WPB * L.fixed_total;
  int* actl = reinterpret_cast<int*>(arena + L.arena_cap);
  if (threadIdx.x == 0) { actl[0] = 0; actl[1] = 0; }
  __syncthreads();
  if (gate && *gate == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = w1 - w0;
  int* next = d.di[B2_DI_SOLVE_NEXT] + w0;
  int* ring = d.di[B2_DI_SOLVE_RING] + w0;
  int* ctr = d.di[B2_DI_SOLVE_CTR] + 4 * (size_t)w0;
  int slot = blockIdx.x * B2_WPB + warp;
  while (slot < n) {
    int w = d.di[B2_DI_SOLVE_ORDER][w0 + slot];
    if (!(mask && !mask[w])) {
#ifdef B2_TIMING
      unsigned long long t0 = b2_gtime();
#endif
      solve_world_t<CG, FR, EL>(m, d, L, w, smem + (size_t)warp * L.fixed_total, arena, actl);
      __syncwarp();
#ifdef B2_TIMING
      if (lane == 0 && w < 16384) { g_tim[3 * w] = t0; g_tim[3 * w + 1] = b2_gtime(); g_tim[3 * w + 2] = b2_smid(); }
#endif
      if (lane == 0) {  // release this warp's arena block
        __threadfence_block();
        atomicAnd(reinterpret_cast<unsigned*>(actl), (warp & 1) ? 0x0000ffffu : 0xffff0000u);
      }
    }
    if (post_mode) {  // publish: every lane's results visible device-wide, then the ring entry
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        int pos = atomicAdd(&ctr[0], 1);
        *reinterpret_cast<volatile int*>(ring + pos) = w;
      }
    }
    if (lane == 0) slot = atomicAdd(next, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
  }
  if (!post_mode) return;
  real* psm = smem + (size_t)warp * L.fixed_total;   // the warp's solver scratch, free now (post layout <= fixed part, make_plan)
  for (;;) {
    int w = -1;
    if (lane == 0) {
      int idx = atomicAdd(&ctr[1], 1);
      if (idx < n) {
        // bounded like the arena wait: the bound only keeps a logic error from hanging the GPU
        for (int spins = 0; (w = *reinterpret_cast<volatile int*>(ring + idx)) < 0 && spins < (1 << 24); ++spins) __nanosleep(256);
        __threadfence();
      }
    }
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w < 0) break;
    if (mask && !mask[w]) continue;
    post_world(m, d, LP, w, psm, post_mode == 2);
    __syncwarp();
  }
}

// the instantiations under the names the profiles use: k_solve (Newton, no friction-loss rows: the shipped configuration),
// k_solve_cg, and the two that carry the friction-loss row code (models with dof_frictionloss > 0)
#define k_solve k_solve_t<false, false>
#define k_solve_cg k_solve_t<true, false>
#define k_solve_fr k_solve_t<false, true>
#define k_solve_cg_fr k_solve_t<true, true>
// elliptic friction cones (opt.cone = 1): Newton and CG, without / with friction-loss rows
#define k_solve_el k_solve_t<false, false, true>
#define k_solve_cg_el k_solve_t<true, false, true>
#define k_solve_fr_el k_solve_t<false, true, true>
#define k_solve_cg_fr_el k_solve_t<true, true, true>
typedef void (*solve_kernel_t)(const KModel, const KData, const SolveSmem, const PostSmem, int, const unsigned char*, const int*, int, int);
static solve_kernel_t solve_kernel(const KModel& km) {
  const bool cg = km.solver == 1, fr = km.nfriction > 0;
  if (km.cone == 1) return fr ? (cg ? (solve_kernel_t)k_solve_cg_fr_el : (solve_kernel_t)k_solve_fr_el) : (cg ? (solve_kernel_t)k_solve_cg_el : (solve_kernel_t)k_solve_el);
  return fr ? (cg ? (solve_kernel_t)k_solve_cg_fr : (solve_kernel_t)k_solve_fr) : (cg ? (solve_kernel_t)k_solve_cg : (solve_kernel_t)k_solve);
}
static void launch_solve(int grid, size_t smem, cudaStream_t st, const KModel& km, const KData& kd, const SolveSmem& lv, const PostSmem& lp,
                         int post_mode, const unsigned char* mask, const int* gate, int w0, int w1) {
  solve_kernel(km)<<<grid, B2_WPB * 32, smem, st>>>(km, kd, lv, lp, post_mode, mask, gate, w0, w1);
}

// Counting sort of the worlds by predicted solver cost, heaviest first (one CTA).  Which solver slot a world lands in has no
// effect on its results; the atomics only make the order inside a bucket arbitrary.
#define B2_ORDER_BINS 64
// Sort key.  Per-world timers (tools/solve_timeline.py): the latency of a world is ~42 us per Newton iteration almost
// regardless of its row count, so contacts alone (mode 0, round 1) leave 8-iteration worlds with 8 contacts in the middle
// of the order, where they become the launch's tail.  The iteration count of the world's previous solve predicts the next
// one well enough when it is ADDED to the contact count (a world with many contacts is never sorted as light, which is what
// went wrong when the product of the two was tried in round 1).  `mode`: 0 = contacts, k > 0 = contacts + k * previous iterations.
static __device__ __forceinline__ int order_key(int ncon, int niter_prev, int mode) {
  int k = ncon + mode * niter_prev;
  return k < B2_ORDER_BINS ? k : B2_ORDER_BINS - 1;
}
__global__ void __launch_bounds__(1024) k_order(const KData d, int w0, int w1, int cap, int mode) {
  __shared__ int hist[B2_ORDER_BINS], base[B2_ORDER_BINS];
  const int* ncon = d.di[B2_DI_NCON];
  const int* niter = d.di[B2_DI_SOLVER_NITER];   // of the world's previous solve
  int* order = d.di[B2_DI_SOLVE_ORDER];
  int* ring = d.di[B2_DI_SOLVE_RING];
  if (threadIdx.x < B2_ORDER_BINS) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    d.di[B2_DI_SOLVE_NEXT][w0] = solve_grid(w1 - w0, cap) * B2_WPB;   // first slot no warp starts on
    d.di[B2_DI_SOLVE_CTR][4 * (size_t)w0] = 0;
    d.di[B2_DI_SOLVE_CTR][4 * (size_t)w0 + 1] = 0;
  }
  __syncthreads();
  for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
    int k = order_key(ncon[w], niter[w], mode);
    atomicAdd(&hist[k], 1);
    ring[w] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = B2_ORDER_BINS - 1; k >= 0; --k) { base[k] = acc; acc += hist[k]; }
  }
  __syncthreads();
  for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
    int k = order_key(ncon[w], niter[w], mode);
    // Rank p in the heaviest-first order -> solver slot.  G CTAs of two warps are resident; warp (c, i) starts on slot
    // 2c + i and then pulls slots 2G, 2G+1, ... in order.  The two worlds of a CTA share one row arena and a world that
    // does not fit next to its CTA-mate waits for the mate to finish (per-world timers: 55 % of the latency of the
    // heaviest worlds was that wait while they were paired with each other), so the first wave pairs the c-th heaviest
    // world with the c-th lightest; the warps of the light worlds are free within microseconds and take the remaining
    // worlds longest-first, which keeps the tail of the launch short.
    int slot = solve_slot(atomicAdd(&base[k], 1), w1 - w0, cap);
    order[w0 + slot] = w;   // slots [w0, w1) hold a permutation of the worlds [w0, w1)
  }
}

#define B2_POST_WPB 4
// resident CTAs per SM the compiler must allow for (register cap 65536 / (B2_POST_MINB * 128)): the 1024 CTAs of a
// 4096-world launch have to fit in ONE round (148 * 7 = 1036); left alone the compiler used 96 registers -> 5 per SM, two rounds
#ifndef B2_POST_MINB
#define B2_POST_MINB 7
#endif
__global__ void __launch_bounds__(B2_POST_WPB * 32, B2_POST_MINB) k_post(const KModel m, const KData d, const PostSmem L, int do_integrate, const unsigned char* mask, const int* gate, int w0, int w1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* smem = reinterpret_cast<real*>(smem_raw);
  int warp = threadIdx.x >> 5;
  int w = w0 + blockIdx.x * B2_POST_WPB + warp;
  if (w >= w1) return;
  if (gate && *gate == 0) return;
  if (mask && !mask[w]) return;
  post_world(m, d, L, w, smem + (size_t)warp * L.total, do_integrate);
}

// masked reset, one warp per world
__global__ void k_reset(const KModel m, const KData d, const unsigned char* mask) {
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= d.nworld) return;
  if (mask && !mask[w]) return;
  reset_world(m, d, w);
}

__global__ void k_repeat_rows(float* dst, const float* src, int row_len, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dst[i] = src[i % row_len];
}

// static geoms: every world's geom_xpos / geom_xmat row gets the model constant (one thread per (world, geom, component))
__global__ void k_init_static(const KModel m, const KData d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)d.nworld * m.ngeom * 12;
  if (i >= total) return;
  int c = (int)(i % 12);
  long long wg = i / 12;
  int g = (int)(wg % m.ngeom);
  int w = (int)(wg / m.ngeom);
  // a geom is static iff it is not in dyn_geom; dyn_geom is ascending, so binary search
  const int* dyn = m.mi[B2_MI_DYN_GEOM];
  int lo = 0, hi = m.ndyn;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (dyn[mid] < g) lo = mid + 1; else hi = mid; }
  if (lo < m.ndyn && dyn[lo] == g) return;
  if (c < 3) d.df[B2_DF_GEOM_XPOS][((size_t)w * m.ngeom + g) * 3 + c] = (m.mf[B2_MF_SGEOM_XPOS] + (size_t)w * m.mf_stride[B2_MF_SGEOM_XPOS])[3 * g + c];
  else d.df[B2_DF_GEOM_XMAT][((size_t)w * m.ngeom + g) * 9 + (c - 3)] = (m.mf[B2_MF_SGEOM_XMAT] + (size_t)w * m.mf_stride[B2_MF_SGEOM_XMAT])[9 * g + c - 3];
}

struct Plan {
  SmoothSmem ls;
  SolveSmem lv;
  PostSmem lp;
  size_t smem_smooth, smem_solve, smem_post;
  int solve_cap;    // resident solver CTAs on this device: SMs x CTAs per SM granted by the occupancy calculator
  int order_mode;   // sort key of k_order (see order_key)
  int memo;         // the env step's first sub-step reuses the preceding forward pass's position / velocity stage (B200SIM_MEMO, default on)
  int fuse_post;    // run the sensor / integration stage inside k_solve's tail (experiment, default off: see k_solve)
  int chunks, chunk_min;
};

static_assert(B2_WPB <= 2, "the arena hand-over in k_solve assumes at most two worlds per CTA");

// Experiment knobs (tests flip them in-process) are read when a plan is BUILT, not on every enqueue: a plan is cached per
// (model sizes, data sizes, device) and rebuilt by b2_check, which Simulation calls at construction.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

static int make_plan(const b2_model* m, const b2_data* d, Plan* p) {
  p->ls = smooth_smem_layout(m->nq, m->nv, m->nu, m->nbody, m->njnt, B2_NCAND_MAX);
  // true upper bounds on rows / runs / pools of ONE world; the arena of a CTA holds one such world, and two worlds
  // whenever their actual demand fits (the common case: demand is an order of magnitude below the bound)
  const SolveCaps caps = solve_caps(m->nlimited, m->nfriction, d->nconmax, d->njmax, m->maxchain);
  const int rows_max = caps.rows, runs_max = caps.runs, pool_max = caps.pool, dofs_max = caps.dofs;
  int arena = solve_arena_need(rows_max, runs_max, dofs_max, pool_max, d->nconmax > 0 ? d->nconmax : 1, m->cone);
  // The strict bound assumes every contact couples the two longest kinematic chains (maxchain); real worlds stay an
  // order of magnitude below it (G1, random actions: median 30 rows, max 116 of 198).  Cap the arena so that
  // B2_SOLVE_MINB CTAs stay resident per SM, but never below half the strict bound.  A world beyond the cap loses rows
  // from the tail of its contact list exactly like njmax truncation does (counted in nefc_dropped).
  p->lp = post_smem_layout(m->nv);
  {
    SolveSmem probe = solve_smem_layout(m->nv, rows_max, m->maxchain, 0, p->lp.total, m->nfriction);
    long budget = (233472 / B2_SOLVE_MINB - 1024 - 16 - (long)probe.fixed_total * B2_WPB * (long)sizeof(real)) / (long)sizeof(real);
    if (budget < arena / 2) budget = arena / 2;
    if (arena > budget) arena = (int)budget & ~3;
  }
  arena = env_int("B200SIM_ARENA", arena);  // experiments
  {
    // whatever the cap, a world reduced to one row and no Jacobian pool must fit (solve_world shrinks towards that)
    int floor_ = solve_arena_need(1, runs_max, dofs_max, 0, d->nconmax > 0 ? d->nconmax : 1, m->cone) + 64;
    if (arena < floor_) arena = floor_;
  }
  arena = (arena + 3) & ~3;   // the two-ended allocator of k_solve works in units of 4 floats
  p->lv = solve_smem_layout(m->nv, rows_max, m->maxchain, arena, p->lp.total, m->nfriction);
  p->smem_smooth = (size_t)p->ls.total * sizeof(real) * B2_SMOOTH_WPB;
  p->smem_solve = ((size_t)p->lv.fixed_total * B2_WPB + p->lv.arena_cap) * sizeof(real) + 16;
  p->smem_post = (size_t)p->lp.total * sizeof(real) * B2_POST_WPB;
  p->smem_solve += (size_t)env_int("B200SIM_SMEM_PAD", 0);  // occupancy experiments only
  p->solve_cap = B2_SOLVE_DEFAULT_CTAS;
  p->order_mode = env_int("B200SIM_ORDER_MODE", B2_ORDER_MODE_DEFAULT);
  p->fuse_post = env_int("B200SIM_FUSE_POST", 0);
  p->memo = env_int("B200SIM_MEMO", 1);
  p->chunks = env_int("B200SIM_CHUNKS", 1);
  p->chunk_min = env_int("B200SIM_CHUNK_MIN", 512);
  return 0;
}

// Plan cache: one entry per (sizes the layouts depend on, device).  b2_check builds / refreshes the entry (and is the only
// place that touches the device attributes); the enqueue paths only look it up.
struct PlanKey { int v[12]; };
struct PlanEntry { PlanKey key; Plan plan; bool checked; };
static PlanEntry g_plans[32];
static int g_nplans = 0;

static PlanKey plan_key(const b2_model* m, const b2_data* d) {
  int dev = -1;
  cudaGetDevice(&dev);
  PlanKey k = {{m->nq, m->nv, m->nu, m->nbody, m->njnt, m->nlimited, m->maxchain, d->nconmax, d->njmax, dev, m->solver, m->nfriction + (m->cone << 20)}};
  return k;
}

static PlanEntry* plan_lookup(const b2_model* m, const b2_data* d, bool rebuild) {
  PlanKey k = plan_key(m, d);
  PlanEntry* e = nullptr;
  for (int i = 0; i < g_nplans; ++i)
    if (memcmp(&g_plans[i].key, &k, sizeof(k)) == 0) { e = &g_plans[i]; break; }
  if (e && !rebuild) return e;
  if (!e) {
    e = &g_plans[g_nplans < 32 ? g_nplans++ : 31];
    e->key = k;
    e->checked = false;
  }
  make_plan(m, d, &e->plan);
  return e;
}

extern "C" {

int b2_abi_version(void) { return 2; }
const char* b2_last_error(void) { return g_err; }
int b2_debug_solver_slot(int rank, int nworld) { return solve_slot(rank, nworld, B2_SOLVE_DEFAULT_CTAS); }
int b2_sizeof_model(void) { return (int)sizeof(b2_model); }
int b2_sizeof_data(void) { return (int)sizeof(b2_data); }
long long b2_launch_count(void) { return g_launches; }

static int check_sizes(const b2_model* m, const b2_data* d) {
  if (!m || !d) return fail("null model/data");
  if (m->nv > 64) return fail("nv=%d exceeds the 64-dof limit of the per-world dof masks", m->nv);
  if (m->nv < 1) return fail("nv must be >= 1");
  if (m->cone != 0 && m->cone != 1) return fail("unknown friction cone %d (0 pyramidal, 1 elliptic)", m->cone);
  if (m->solver != 2 && m->solver != 1) return fail("supported solvers: Newton, CG (PGS is not implemented)");
  if (m->integrator != B2_INT_EULER && m->integrator != B2_INT_IMPLICITFAST) return fail("unsupported integrator %d", m->integrator);
  if (d->nworld < 1 || d->nconmax < 0 || d->njmax < 1) return fail("bad nworld/nconmax/njmax");
  return 0;
}

int b2_check(const b2_model* m, const b2_data* d) {
  if (check_sizes(m, d)) return 1;
  int dev = 0, maxsmem = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
  PlanEntry* pe = plan_lookup(m, d, true);
  Plan& p = pe->plan;
  cudaDeviceGetAttribute(&maxsmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if ((int)p.smem_smooth > maxsmem || (int)p.smem_solve > maxsmem || (int)p.smem_post > maxsmem)
    return fail("shared memory per CTA (%zu / %zu B) exceeds the device limit %d B; lower njmax or model size",
                p.smem_smooth, p.smem_solve, maxsmem);
  cudaError_t e;
  e = cudaFuncSetAttribute(k_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_smooth);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(k_smooth): %s", cudaGetErrorString(e));
  const solve_kernel_t ksolve = solve_kernel(*reinterpret_cast<const KModel*>(m));
  e = cudaFuncSetAttribute(ksolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_solve);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(k_solve): %s", cudaGetErrorString(e));
  e = cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_post);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(k_post): %s", cudaGetErrorString(e));
  // The solver grid must be fully resident (its warps wait on each other's results in the fused sensor stage): size it
  // from what the occupancy calculator grants on THIS device, not from a compiled-in SM count.
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksolve, B2_WPB * 32, p.smem_solve);
  if (e != cudaSuccess || per_sm < 1 || nsm < 1) return fail("k_solve does not fit on an SM (%s)", cudaGetErrorString(e));
  p.solve_cap = per_sm * nsm;
  pe->checked = true;
  return 0;
}

static const Plan* get_plan(const b2_model* m, const b2_data* d) {
  PlanEntry* pe = plan_lookup(m, d, false);
  if (!pe->checked) {   // a caller that skipped b2_check: do it now (first enqueue only)
    if (b2_check(m, d)) return nullptr;
    pe = plan_lookup(m, d, false);
  }
  return &pe->plan;
}

// One physics step (do_integrate) or forward pass for the worlds [w0, w1) on stream st: four launches (kinematics ...
// qacc_smooth; the solver order; constraint rows + Newton solver; sensors + integration).
static void launch_range(const Plan& p, const KModel& km, const KData& kd, cudaStream_t st, int do_integrate, const unsigned char* mask,
                         const int* gate, int w0, int w1, int memo_consume = 0) {
  int n = w1 - w0;
  k_smooth<<<(n + B2_SMOOTH_WPB - 1) / B2_SMOOTH_WPB, B2_SMOOTH_WPB * 32, p.smem_smooth, st>>>(km, kd, p.ls, B2_NCAND_MAX, memo_consume, mask, gate, w0, w1);
  k_order<<<1, 1024, 0, st>>>(kd, w0, w1, p.solve_cap, p.order_mode);
  if (p.fuse_post) {
    launch_solve(solve_grid(n, p.solve_cap), p.smem_solve, st, km, kd, p.lv, p.lp, do_integrate ? 2 : 1, mask, gate, w0, w1);
    g_launches += 3;
  } else {
    launch_solve(solve_grid(n, p.solve_cap), p.smem_solve, st, km, kd, p.lv, p.lp, 0, mask, gate, w0, w1);
    k_post<<<(n + B2_POST_WPB - 1) / B2_POST_WPB, B2_POST_WPB * 32, p.smem_post, st>>>(km, kd, p.lp, do_integrate, mask, gate, w0, w1);
    g_launches += 4;
  }
}

static int run(const b2_model* m, b2_data* d, void* stream, int do_integrate, const unsigned char* mask = nullptr, const int* gate = nullptr) {
  const Plan* p = get_plan(m, d);
  if (!p) return 1;
  launch_range(*p, *reinterpret_cast<const KModel*>(m), *reinterpret_cast<const KData*>(d), (cudaStream_t)stream, do_integrate, mask, gate, 0, d->nworld);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- chunked, multi-stream execution (opt-in: B200SIM_CHUNKS=N) --------------------------------------------------------
// Worlds never interact, so the batch can be cut into chunks whose kernel CHAINS run on separate streams, forked from and
// joined to the caller's stream (graph edges under stream capture: the env step stays ONE graph).  The idea was to let one
// chunk's solver tail overlap the other chunks' kinematics / sensors / next solver.  Measured (G1, 4096 envs): no gain at
// 2 / 4 / 8 chunks - while a solver launch is resident it holds every register of the SMs it runs on, so the other chunks'
// kernels only start once it drains, and the chain of each chunk is still 4 x the latency of its heaviest world.  What did
// shorten the tail is inside k_solve (persistent warps + heavy/light pairing).  Kept as an experiment knob, default 1.
#define B2_MAX_CHUNKS 16
struct ChunkStreams {
  int dev = -1, n = 0;
  cudaStream_t side[B2_MAX_CHUNKS];
  cudaEvent_t fork, join[B2_MAX_CHUNKS];
};
static ChunkStreams g_cs[16];

static int chunk_count(const Plan& p, int nworld) {
  int c = p.chunks > 0 ? p.chunks : 1;   // default: one chunk (see the note above run_chunked)
  int min_worlds = p.chunk_min > 0 ? p.chunk_min : 512;   // below this many worlds per chunk the launches only add overhead
  if (c > B2_MAX_CHUNKS) c = B2_MAX_CHUNKS;
  while (c > 1 && nworld / c < min_worlds) c >>= 1;
  return c < 1 ? 1 : c;
}

static ChunkStreams* chunk_streams(int want) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  ChunkStreams& cs = g_cs[dev];
  if (cs.dev != dev) {
    if (cudaEventCreateWithFlags(&cs.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    cs.dev = dev;
    cs.n = 0;
  }
  while (cs.n < want) {
    if (cudaStreamCreateWithFlags(&cs.side[cs.n], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&cs.join[cs.n], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    cs.n += 1;
  }
  return &cs;
}

typedef void (*chunk_hook)(const void* ctx, cudaStream_t st, int w0, int w1);

// nsub x [physics step or forward; hook] for every world, chunk chains on side streams forked from / joined to `stream`.
// memo_first: the first of the nsub passes may reuse the stage outputs a preceding pass left in HBM (see smooth_world).
static int run_chunked(const b2_model* m, b2_data* d, void* stream, int nsub, int do_integrate, const unsigned char* mask, const int* gate,
                       chunk_hook hook, const void* hook_ctx, int memo_first = 0) {
  const Plan* pp = get_plan(m, d);
  if (!pp) return 1;
  const Plan& p = *pp;
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  cudaStream_t main_st = (cudaStream_t)stream;
  int nc = chunk_count(p, d->nworld);
  ChunkStreams* cs = nc > 1 ? chunk_streams(nc) : nullptr;
  if (!cs) {
    for (int s = 0; s < nsub; ++s) {
      launch_range(p, km, kd, main_st, do_integrate, mask, gate, 0, d->nworld, memo_first && p.memo && s == 0);
      if (hook) hook(hook_ctx, main_st, 0, d->nworld);
    }
  } else {
    // chunk boundaries on multiples of 4 worlds (the env kernels pack 4 worlds per CTA)
    int per = ((d->nworld + nc - 1) / nc + 3) & ~3;
    if (cudaEventRecord(cs->fork, main_st) != cudaSuccess) return fail("event record: %s", cudaGetErrorString(cudaGetLastError()));
    for (int c = 0; c < nc; ++c) {
      int w0 = c * per, w1 = w0 + per < d->nworld ? w0 + per : d->nworld;
      if (w0 >= w1) continue;
      cudaStream_t st = cs->side[c];
      if (cudaStreamWaitEvent(st, cs->fork, 0) != cudaSuccess) return fail("stream wait: %s", cudaGetErrorString(cudaGetLastError()));
      for (int s = 0; s < nsub; ++s) {
        launch_range(p, km, kd, st, do_integrate, mask, gate, w0, w1, memo_first && p.memo && s == 0);
        if (hook) hook(hook_ctx, st, w0, w1);
      }
      if (cudaEventRecord(cs->join[c], st) != cudaSuccess) return fail("event record: %s", cudaGetErrorString(cudaGetLastError()));
      if (cudaStreamWaitEvent(main_st, cs->join[c], 0) != cudaSuccess) return fail("stream wait: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// Stages on their own (bench.py brackets each with CUDA events, stage-wise tests): 0 = k_smooth + k_order, 1 = k_solve,
// 2 = k_post, 3 = k_solve with the fused sensor / integration stage (experiment), 4 = k_smooth alone, 5 = k_order alone,
// 6 = k_smooth as the first sub-step after a forward pass launches it (b2_substeps_after_forward).
int b2_stage(const b2_model* m, b2_data* d, int stage, int do_integrate, void* stream) {
  const Plan* pp = get_plan(m, d);
  if (!pp) return 1;
  const Plan& p = *pp;
  cudaStream_t st = (cudaStream_t)stream;
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  if (stage == 0) {
    k_smooth<<<(d->nworld + B2_SMOOTH_WPB - 1) / B2_SMOOTH_WPB, B2_SMOOTH_WPB * 32, p.smem_smooth, st>>>(km, kd, p.ls, B2_NCAND_MAX, 0, nullptr, nullptr, 0, d->nworld);
    k_order<<<1, 1024, 0, st>>>(kd, 0, d->nworld, p.solve_cap, p.order_mode);
    g_launches += 1;
  }
  else if (stage == 1) launch_solve(solve_grid(d->nworld, p.solve_cap), p.smem_solve, st, km, kd, p.lv, p.lp, 0, nullptr, nullptr, 0, d->nworld);
  else if (stage == 2) k_post<<<(d->nworld + B2_POST_WPB - 1) / B2_POST_WPB, B2_POST_WPB * 32, p.smem_post, st>>>(km, kd, p.lp, do_integrate, nullptr, nullptr, 0, d->nworld);
  else if (stage == 3) launch_solve(solve_grid(d->nworld, p.solve_cap), p.smem_solve, st, km, kd, p.lv, p.lp, do_integrate ? 2 : 1, nullptr, nullptr, 0, d->nworld);
  else if (stage == 4) k_smooth<<<(d->nworld + B2_SMOOTH_WPB - 1) / B2_SMOOTH_WPB, B2_SMOOTH_WPB * 32, p.smem_smooth, st>>>(km, kd, p.ls, B2_NCAND_MAX, 0, nullptr, nullptr, 0, d->nworld);
  else if (stage == 6) k_smooth<<<(d->nworld + B2_SMOOTH_WPB - 1) / B2_SMOOTH_WPB, B2_SMOOTH_WPB * 32, p.smem_smooth, st>>>(km, kd, p.ls, B2_NCAND_MAX, p.memo, nullptr, nullptr, 0, d->nworld);
  else if (stage == 5) k_order<<<1, 1024, 0, st>>>(kd, 0, d->nworld, p.solve_cap, p.order_mode);
  else return fail("unknown stage %d", stage);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int b2_step(const b2_model* m, b2_data* d, void* stream) { return run(m, d, stream, 1); }
int b2_forward(const b2_model* m, b2_data* d, void* stream) { return run(m, d, stream, 0); }

int b2_reset(const b2_model* m, b2_data* d, const unsigned char* mask, void* stream) {
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  int grid = (d->nworld + 3) / 4;
  k_reset<<<grid, 128, 0, (cudaStream_t)stream>>>(km, kd, mask);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int b2_init_static(const b2_model* m, b2_data* d, void* stream) {
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  long long total = (long long)d->nworld * m->ngeom * 12;
  if (total <= 0) return 0;
  k_init_static<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(km, kd);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int b2_repeat_rows(float* dst, const float* src, int row_len, int nworld, void* stream) {
  long long total = (long long)row_len * nworld;
  if (total <= 0) return 0;
  int grid = (int)((total + 255) / 256);
  k_repeat_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, src, row_len, total);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int b2_forward_masked(const b2_model* m, b2_data* d, const unsigned char* mask, const int* gate, void* stream) {
  return run_chunked(m, d, stream, 1, 0, mask, gate, nullptr, nullptr);
}

int b2_substeps(const b2_model* m, b2_data* d, int nsub, void* stream) {
  if (nsub < 0) return fail("nsub must be >= 0");
  return run_chunked(m, d, stream, nsub, 1, nullptr, nullptr, nullptr, nullptr);
}

int b2_substeps_after_forward(const b2_model* m, b2_data* d, int nsub, void* stream) {
  if (nsub < 0) return fail("nsub must be >= 0");
  return run_chunked(m, d, stream, nsub, 1, nullptr, nullptr, nullptr, nullptr, 1);
}

}  // extern "C"

#include "b2_env_api.inl"
#include "b2_trk_api.inl"
