// b2_common.cuh - execution-model shims, small math, shared-memory carving.
//
// Scheduling model: ONE WARP PER WORLD.  Every cooperative loop is written as
//   for (i = B2_LANE; i < n; i += B2_NLANE) {...}  B2_WSYNC();
// so the 32 lanes of a warp span the elements (bodies of a tree level, dofs,
// geom pairs, constraint rows) of one world, reductions are warp shuffles and
// the only barrier is __syncwarp().  World rows in HBM are world-major, so a
// warp's loads of its world's row are coalesced.
//
// The same source also compiles with -DB2_HOSTEMU as plain C++ with a single
// lane (B2_NLANE == 1): a DEBUG HARNESS used only by tests/ (logic can be
// checked in a container without a GPU).  It is never built into, nor loaded
// by, the product library.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef B2_REAL
#define B2_REAL float
#endif
typedef B2_REAL real;

#ifdef B2_HOSTEMU
#define B2_DEV static inline
#define B2_DEVN static
#define B2_UNROLL1
#define B2_UNROLL_HOT
#define B2_LANE 0
#define B2_NLANE 1
#define B2_WSYNC() ((void)0)
B2_DEV real b2_wsum(real x) { return x; }  // @region shuffle_reduce
B2_DEV real b2_wmax(real x) { return x; }
B2_DEV int b2_wsum_i(int x) { return x; }
B2_DEV void b2_wsum3(real&, real&, real&) {}
B2_DEV unsigned b2_ballot(bool p) { return p ? 1u : 0u; }
B2_DEV int b2_wor_i(int x) { return x; }
B2_DEV int b2_popc(unsigned x) { return __builtin_popcount(x); }
B2_DEV int b2_popcll(unsigned long long x) { return __builtin_popcountll(x); }
B2_DEV unsigned b2_lanemask_lt() { return 0u; }
B2_DEV real b2_rsqrt(real x) { return (real)1 / sqrt(x); }
#define B2_LDG(p) (*(p))
#else
#define B2_DEV __device__ __forceinline__
// Out-of-line device helpers + rolled loops on the Newton path: the solver kernel is instruction-fetch
// bound (ncu: stall_no_instruction), so hot code is kept small enough for the instruction caches.
#define B2_DEVN static __device__ __noinline__
#define B2_UNROLL1 _Pragma("unroll 1")
// hottest inner loops only: a small unroll factor buys instruction-level parallelism (loads of the next iteration issue
// under the FMAs of the current one); B2_HOT_UNROLL is a build-time knob for the tuning experiments in DESIGN.md
#ifndef B2_HOT_UNROLL
#define B2_HOT_UNROLL 2
#endif
#define B2_PRAGMA_(x) _Pragma(#x)
#define B2_PRAGMA(x) B2_PRAGMA_(x)
#define B2_UNROLL_HOT B2_PRAGMA(unroll B2_HOT_UNROLL)
#define B2_LANE ((int)(threadIdx.x & 31))
#define B2_NLANE 32
#define B2_WSYNC() __syncwarp()
B2_DEV real b2_wsum(real x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
B2_DEV real b2_wmax(real x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
B2_DEV int b2_wsum_i(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
B2_DEV int b2_wor_i(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
// Sums of three values over the warp with 9 shuffles instead of 15: after the xor-16 and xor-8 exchanges each lane
// carries ONE of the (padded to four) values, the remaining three butterfly steps finish that value, three broadcasts
// hand every lane all three sums.
B2_DEV void b2_wsum3(real& a, real& b, real& c) {
  const unsigned lane = threadIdx.x & 31;
  const bool hi = lane & 16, mid = lane & 8;
  real r1 = __shfl_xor_sync(0xffffffffu, hi ? a : c, 16);
  real r2 = __shfl_xor_sync(0xffffffffu, hi ? b : (real)0, 16);
  real A = (hi ? c : a) + r1, B = (hi ? (real)0 : b) + r2;      // !hi: (a, b) partial sums; hi: (c, 0)
  real C = (mid ? B : A) + __shfl_xor_sync(0xffffffffu, mid ? A : B, 8);
  C += __shfl_xor_sync(0xffffffffu, C, 4);
  C += __shfl_xor_sync(0xffffffffu, C, 2);
  C += __shfl_xor_sync(0xffffffffu, C, 1);
  a = __shfl_sync(0xffffffffu, C, 0);
  b = __shfl_sync(0xffffffffu, C, 8);
  c = __shfl_sync(0xffffffffu, C, 16);
}
B2_DEV unsigned b2_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
B2_DEV int b2_popc(unsigned x) { return __popc(x); }
B2_DEV int b2_popcll(unsigned long long x) { return __popcll(x); }
B2_DEV unsigned b2_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
B2_DEV real b2_rsqrt(real x) { return rsqrtf(x); }
#define B2_LDG(p) __ldg(p)
#endif

#define B2_MINVAL ((real)1e-15)
#ifndef B2_QUAT_UNIT_TOL
#define B2_QUAT_UNIT_TOL ((real)(sizeof(real) == 4 ? 3.6e-7 : 6.7e-16))   // 3 ulp of 1: after one rescale the COMPUTED norm is within 2 ulp of 1 (2e6 random float32 quaternions), so the second pass is a no-op
#endif
#define B2_MINIMP ((real)0.0001)
#define B2_MAXIMP ((real)0.9999)

// enums shared with the host compiler (mjlab_b200/mjcf.py)
enum { B2_GEOM_PLANE = 0, B2_GEOM_HFIELD = 1, B2_GEOM_SPHERE = 2, B2_GEOM_CAPSULE = 3, B2_GEOM_ELLIPSOID = 4,
       B2_GEOM_CYLINDER = 5, B2_GEOM_BOX = 6, B2_GEOM_MESH = 7 };
enum { B2_JNT_FREE = 0, B2_JNT_BALL = 1, B2_JNT_SLIDE = 2, B2_JNT_HINGE = 3 };
enum { B2_INT_EULER = 0, B2_INT_IMPLICITFAST = 3 };
enum { B2_OBJ_BODY = 1, B2_OBJ_XBODY = 2, B2_OBJ_JOINT = 3, B2_OBJ_GEOM = 5, B2_OBJ_SITE = 6 };
enum { B2_SENS_ACCELEROMETER = 1, B2_SENS_VELOCIMETER = 2, B2_SENS_GYRO = 3, B2_SENS_JOINTPOS = 8, B2_SENS_JOINTVEL = 9,
       B2_SENS_FRAMEPOS = 25, B2_SENS_FRAMEQUAT = 26, B2_SENS_FRAMELINVEL = 30, B2_SENS_FRAMEANGVEL = 31,
       B2_SENS_SUBTREECOM = 34, B2_SENS_SUBTREELINVEL = 35, B2_SENS_SUBTREEANGMOM = 36, B2_SENS_CONTACT = 42 };

// ---------------------------------------------------------------------------
// Kernel-side views of b2_model / b2_data (layout-identical when real==float).
// ---------------------------------------------------------------------------
#include "../../include/b200sim.h"

struct KModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nsensordata, npair, nmocap, nlevel;
  int integrator, cone, solver, iterations, ls_iterations, disableflags;
  int has_accelerometer, has_subtreevel;
  int nlimited, maxchain, nfriction;
  int nfield, nfq, ndyn, field_nx, field_ny;
  const real* mf[B2_MF_COUNT];
  int mf_stride[B2_MF_COUNT];
  const int* mi[B2_MI_COUNT];
};
struct KData {
  int nworld, nconmax, njmax;
  int flags;
  real* df[B2_DF_COUNT];
  int* di[B2_DI_COUNT];
};

#define MF(id) (m.mf[id] + (size_t)w * m.mf_stride[id])
#define MI(id) (m.mi[id])

// ---------------------------------------------------------------------------
// small vector / quaternion math (row-major 3x3)
// ---------------------------------------------------------------------------
// bit-for-bit equality (a copied NaN equals itself, -0 and +0 differ)
B2_DEV bool b2_same_bits(real a, real b) {
  if (sizeof(real) == 4) { unsigned x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return x == y; }
  unsigned long long x = 0, y = 0; memcpy(&x, &a, sizeof(real)); memcpy(&y, &b, sizeof(real)); return x == y;
}
B2_DEV real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }  // @region vecmath
B2_DEV void cross3(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
B2_DEV void quat_mul(real* r, const real* a, const real* b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
B2_DEV void quat_normalize(real* q) {
  real n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < B2_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  // MuJoCo rescales only when the norm is off by more than its rounding noise (mju_normalize4: |norm - 1| > mjMINVAL in
  // double).  The float analogue matters here: an unconditional rescale is NOT idempotent in float (about one unit
  // quaternion in six flips between two neighbours for ever), which would change qpos on every pass over an unchanged state.
  if (fabs(n - (real)1) <= B2_QUAT_UNIT_TOL) return;
  real s = (real)1 / n;
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
B2_DEV void quat2mat(real* R, const real* q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
B2_DEV void mulmv3(real* r, const real* R, const real* v) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
B2_DEV void mulmTv3(real* r, const real* R, const real* v) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  real y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  real z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
B2_DEV void axisangle2quat(real* q, const real* axis, real angle) {
  real s = sin(angle * (real)0.5);
  q[0] = cos(angle * (real)0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// spatial inertia (about a reference point) times motion vector [ang; lin]
// inertia layout: Ixx Iyy Izz Ixy Ixz Iyz hx hy hz m   (h = m * offset)
B2_DEV void inert_mul(real* r, const real* I, const real* v) {
  real a0 = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] + (I[7] * v[5] - I[8] * v[4]);
  real a1 = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + (I[8] * v[3] - I[6] * v[5]);
  real a2 = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] + (I[6] * v[4] - I[7] * v[3]);
  real l0 = I[9] * v[3] - (I[7] * v[2] - I[8] * v[1]);
  real l1 = I[9] * v[4] - (I[8] * v[0] - I[6] * v[2]);
  real l2 = I[9] * v[5] - (I[6] * v[1] - I[7] * v[0]);
  r[0] = a0; r[1] = a1; r[2] = a2; r[3] = l0; r[4] = l1; r[5] = l2;
}
B2_DEV void cross_motion(real* r, const real* v, const real* s) {
  real a[3], b[3], c[3];
  cross3(a, v, s); cross3(b, v, s + 3); cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
B2_DEV void cross_force(real* r, const real* v, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
B2_DEV real clampr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------
// Dense SPD factor + solve for one world (n <= 64), one warp.
//
// The matrix lives in shared memory, row-major with leading dimension ld = b2_ld(n):
// ld % 8 == 4, so (a) rows are 16-byte aligned for 128-bit loads and (b) the eight lanes
// of a quarter-warp reading consecutive rows hit eight distinct 4-bank groups, i.e. the
// LDS.128 of a lane's own row is conflict-free.
//
// Left-looking (Crout) order: column j of L needs, for every row i >= j,
//   s_i = A[i][j] - sum_{k<j} L[i][k] * L[j][k]
// Lane <-> row, the k loop only LOADS (own row: conflict-free, pivot row: broadcast), so
// it pipelines; one store per (row, column) and one __syncwarp per column.  The code is a
// few hundred bytes: an earlier fully unrolled register/shuffle variant was faster per
// instruction but its ~35 KB of straight-line code made the kernel instruction-fetch bound
// (ncu: 63 % of stall cycles "no_instruction", profiles/).
// ---------------------------------------------------------------------------
static inline
#ifndef B2_HOSTEMU
__host__ __device__
#endif
int b2_ld(int n) { int ld = (n + 3) & ~3; if ((ld & 7) != 4) ld += 4; return ld; }

#ifdef B2_HOSTEMU
B2_DEV real b2_shfl(real x, int) { return x; }
#else
B2_DEV real b2_shfl(real x, int src) { return __shfl_sync(0xffffffffu, x, src); }
#endif

// In-place lower Cholesky of A (lower triangle valid, upper ignored) with the forward substitution fused in:
// x (the right-hand side on entry) is treated as row n of the matrix, so on exit x = L^-1 b at no extra pass.
// invd[j] <- 1 / L[j][j].  Left-looking over PANELS OF TWO COLUMNS: one pass over a lane's row serves both columns
// (3 instead of 4 128-bit loads per 8 FMAs, half the per-column overhead); the pivot rows are read in place (their
// entries k < j are final).  j is even inside the panel loop, so the unaligned tail of the k range is 0 or 2 elements.
// Row i of a lower-triangular matrix stored PACKED: rows padded to a multiple of four entries (128-bit row loads stay
// aligned), row i = 4k + r starts at 4 (k + 1) (2k + r).  k_smooth keeps the mass matrix this way: 684 instead of 1260 words
// for the G1's 35 dofs, which is what lets an SM hold every world of a 4096-world batch at once (b2_smooth.cuh, layout).
static inline
#ifndef B2_HOSTEMU
__host__ __device__
#endif
int b2_poff(int i) { const int k = i >> 2, r = i & 3; return 4 * (k + 1) * (2 * k + r); }
#define B2_MROW(A, i) (PACKED ? (A) + b2_poff(i) : (A) + (i) * ld)

template <bool PACKED>
B2_DEVN void chol_factor_t(real* A, int n, int ld, real* invd, real* x) {  // @region chol_factor
  const int nrows = x ? n + 1 : n;
#ifdef B2_HOSTEMU
  for (int j = 0; j < n; ++j) {
    const real* rj = B2_MROW(A, j);
    real piv = rj[j];
    for (int k = 0; k < j; ++k) piv -= rj[k] * rj[k];
    real inv = b2_rsqrt(piv > B2_MINVAL ? piv : B2_MINVAL);
    invd[j] = inv;
    for (int i = j; i < nrows; ++i) {
      real* ri = i < n ? B2_MROW(A, i) : x;
      real s = ri[j];
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      ri[j] = s * inv;
    }
  }
#else
  int j = 0;
  B2_UNROLL1
  for (; j + 1 < n; j += 2) {
    B2_WSYNC();
    const real* r0 = B2_MROW(A, j);
    const real* r1 = B2_MROW(A, j + 1);
    const int j0 = j & ~3;
    real inv0 = 0, inv1 = 0, l10 = 0;
    B2_UNROLL1
    for (int base = j; base < nrows; base += B2_NLANE) {
      const int i = base + B2_LANE;
      real s0 = 0, s1 = 0;
      real* ri = i < n ? B2_MROW(A, i) : x;
      if (i < nrows) {
        s0 = ri[j]; s1 = ri[j + 1];
        real t0 = 0, t1 = 0;
        B2_UNROLL1
        for (int k = 0; k < j0; k += 4) {
          float4 a = *reinterpret_cast<const float4*>(ri + k);
          float4 b = *reinterpret_cast<const float4*>(r0 + k);
          float4 c = *reinterpret_cast<const float4*>(r1 + k);
          s0 = fmaf(-a.x, b.x, s0); t0 = fmaf(-a.y, b.y, t0); s0 = fmaf(-a.z, b.z, s0); t0 = fmaf(-a.w, b.w, t0);
          s1 = fmaf(-a.x, c.x, s1); t1 = fmaf(-a.y, c.y, t1); s1 = fmaf(-a.z, c.z, s1); t1 = fmaf(-a.w, c.w, t1);
        }
        if (j0 < j) {  // j - j0 == 2
          float2 a = *reinterpret_cast<const float2*>(ri + j0);
          float2 b = *reinterpret_cast<const float2*>(r0 + j0);
          float2 c = *reinterpret_cast<const float2*>(r1 + j0);
          s0 = fmaf(-a.x, b.x, s0); t0 = fmaf(-a.y, b.y, t0);
          s1 = fmaf(-a.x, c.x, s1); t1 = fmaf(-a.y, c.y, t1);
        }
        s0 += t0; s1 += t1;
      }
      if (base == j) {
        real piv0 = b2_shfl(s0, 0);
        inv0 = b2_rsqrt(piv0 > B2_MINVAL ? piv0 : B2_MINVAL);
        l10 = b2_shfl(s0, 1) * inv0;                 // L[j+1][j]
        real piv1 = b2_shfl(s1, 1) - l10 * l10;      // A[j+1][j+1] - sum_{k<=j} L[j+1][k]^2
        inv1 = b2_rsqrt(piv1 > B2_MINVAL ? piv1 : B2_MINVAL);
        if (B2_LANE == 0) { invd[j] = inv0; invd[j + 1] = inv1; }
      }
      const real l0 = s0 * inv0;
      const real l1 = fmaf(-l0, l10, s1) * inv1;
      if (i < nrows) {                               // columns j, j+1 are not read by anyone during this panel
        ri[j] = l0;
        if (i > j) ri[j + 1] = l1;
      }
    }
  }
  if (j < n) {  // n odd: last column on its own
    B2_WSYNC();
    const real* rj = B2_MROW(A, j);
    const int j0 = j & ~3;
    real inv = 0;
    B2_UNROLL1
    for (int base = j; base < nrows; base += B2_NLANE) {
      const int i = base + B2_LANE;
      real s = 0;
      real* ri = i < n ? B2_MROW(A, i) : x;
      if (i < nrows) {
        s = ri[j];
        real s2 = 0;
        B2_UNROLL1
        for (int k = 0; k < j0; k += 4) {
          float4 a = *reinterpret_cast<const float4*>(ri + k);
          float4 b = *reinterpret_cast<const float4*>(rj + k);
          s = fmaf(-a.x, b.x, s); s2 = fmaf(-a.y, b.y, s2); s = fmaf(-a.z, b.z, s); s2 = fmaf(-a.w, b.w, s2);
        }
        if (j0 < j) s = fmaf(-ri[j0], rj[j0], s);
        if (j0 + 1 < j) s2 = fmaf(-ri[j0 + 1], rj[j0 + 1], s2);
        if (j0 + 2 < j) s = fmaf(-ri[j0 + 2], rj[j0 + 2], s);
        s += s2;
      }
      if (base == j) {
        real piv = b2_shfl(s, 0);
        inv = b2_rsqrt(piv > B2_MINVAL ? piv : B2_MINVAL);
        if (B2_LANE == 0) invd[j] = inv;
      }
      if (i < nrows) ri[j] = s * inv;
    }
  }
#endif
  B2_WSYNC();
}

B2_DEV void chol_factor(real* A, int n, int ld, real* invd, real* x) { chol_factor_t<false>(A, n, ld, invd, x); }

// x <- L^-1 x (forward substitution) for a factor that is already in place.
B2_DEVN void chol_forwardsolve(const real* L, int n, int ld, const real* invd, real* x) {
#ifdef B2_HOSTEMU
  for (int k = 0; k < n; ++k) {
    real s = x[k];
    for (int i = 0; i < k; ++i) s -= L[k * ld + i] * x[i];
    x[k] = s * invd[k];
  }
#else
  const int lane = B2_LANE;
  const int row0 = lane, row1 = lane + 32;
  real x0 = row0 < n ? x[row0] : (real)0;
  real x1 = row1 < n ? x[row1] : (real)0;
  B2_UNROLL1
  for (int k = 0; k < n; ++k) {
    real l0 = row0 > k && row0 < n ? L[row0 * ld + k] : (real)0;
    real l1 = row1 > k && row1 < n ? L[row1 * ld + k] : (real)0;
    real xk = b2_shfl(k < 32 ? x0 : x1, k & 31) * invd[k];
    if (row0 == k) x0 = xk; else x0 = fmaf(-l0, xk, x0);
    if (row1 == k) x1 = xk; else x1 = fmaf(-l1, xk, x1);
  }
  if (row0 < n) x[row0] = x0;
  if (row1 < n) x[row1] = x1;
#endif
  B2_WSYNC();
}

// x <- L^-T x (backward substitution), x in shared memory.
template <bool PACKED>
B2_DEVN void chol_backsolve_t(const real* L, int n, int ld, const real* invd, real* x) {  // @region chol_solve
#ifdef B2_HOSTEMU
  for (int k = n - 1; k >= 0; --k) {
    real xk = x[k] * invd[k];
    x[k] = xk;
    for (int i = 0; i < k; ++i) x[i] -= B2_MROW(L, k)[i] * xk;
  }
#else
  // lane i keeps x[i] and x[i+32] in registers; x_k is broadcast with a dynamic-lane shuffle
  const int lane = B2_LANE;
  const int row0 = lane, row1 = lane + 32;
  real x0 = row0 < n ? x[row0] : (real)0;
  real x1 = row1 < n ? x[row1] : (real)0;
  B2_UNROLL1
  for (int k = n - 1; k >= 32; --k) {
    const real* Lk = B2_MROW(L, k);
    real l0 = Lk[row0];
    real l1 = row1 < k ? Lk[row1] : (real)0;
    real xk = b2_shfl(x1, k & 31) * invd[k];
    x0 = fmaf(-l0, xk, x0);
    if (row1 == k) x1 = xk; else x1 = fmaf(-l1, xk, x1);
  }
  B2_UNROLL1
  for (int k = (n < 32 ? n : 32) - 1; k >= 0; --k) {
    real l0 = row0 < k ? B2_MROW(L, k)[row0] : (real)0;
    real xk = b2_shfl(x0, k) * invd[k];
    if (row0 == k) x0 = xk; else x0 = fmaf(-l0, xk, x0);
  }
  if (row0 < n) x[row0] = x0;
  if (row1 < n) x[row1] = x1;
#endif
  B2_WSYNC();
}

B2_DEV void chol_backsolve(const real* L, int n, int ld, const real* invd, real* x) { chol_backsolve_t<false>(L, n, ld, invd, x); }

#ifndef B2_SYMV_BATCH
#define B2_SYMV_BATCH 4    // loads in flight per lane and batch (multiple of 4).  Measured (k_solve, G1 4096 worlds): 12 + straight-line tail 0.414 ms, one predicated 12-wide body 0.399, 8 -> 0.398, 4 -> 0.394: code size beats loads in flight (profiles/r02_k_solve_icache.txt)
#endif
// y = M x for a symmetric M [n][ld] in global memory: lane i walks column i (== row i), so every
// load of the warp is one coalesced segment.
B2_DEVN void symv_global(const real* M, int n, int ld, const real* x, real* y) {  // @region symv_global
#ifdef B2_HOSTEMU
  for (int i = 0; i < n; ++i) {
    real a = 0;
    for (int e = 0; e < n; ++e) a += M[e * ld + i] * x[e];
    y[i] = a;
  }
#else
  // Latency-bound on L2: keep many independent loads in flight (12 per batch), no dependent scalar tail.  ONE predicated
  // 12-wide batch body serves every batch, the partial last one included: the straight-line main + tail variants of round 1
  // cost 5.5 KB of the Newton loop's instruction footprint (38 KB against a 32 KB L1.5: profiles/r02_k_solve_icache.txt).
  const int i = B2_LANE;
  if (i < n) {
    const real* Mi = M + i;
    real a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    B2_UNROLL1
    for (int e = 0; e < n; e += B2_SYMV_BATCH) {
      real m[B2_SYMV_BATCH];
#pragma unroll
      for (int k = 0; k < B2_SYMV_BATCH; ++k) m[k] = e + k < n ? Mi[(e + k) * ld] : (real)0;
#pragma unroll
      for (int k = 0; k < B2_SYMV_BATCH; k += 4) {
        a0 = fmaf(m[k], e + k < n ? x[e + k] : (real)0, a0); a1 = fmaf(m[k + 1], e + k + 1 < n ? x[e + k + 1] : (real)0, a1);
        a2 = fmaf(m[k + 2], e + k + 2 < n ? x[e + k + 2] : (real)0, a2); a3 = fmaf(m[k + 3], e + k + 3 < n ? x[e + k + 3] : (real)0, a3);
      }
    }
    y[i] = (a0 + a1) + (a2 + a3);
  }
  // rows 32..n-1 (n > 32): lanes across the row (coalesced), one reduction per row
  B2_UNROLL1
  for (int r = 32; r < n; ++r) {
    real acc = 0;
    B2_UNROLL1
    for (int e = B2_LANE; e < n; e += B2_NLANE) acc = fmaf(M[r * ld + e], x[e], acc);
    acc = b2_wsum(acc);
    if (B2_LANE == 0) y[r] = acc;
  }
#endif
  B2_WSYNC();
}

// dst[0..count) = src[0..count) (global -> shared), count % 4 == 0, both 16-byte aligned:
// 128-bit loads, four in flight per lane so the L2 latency is paid once per batch.
B2_DEVN void copy_g2s(real* dst, const real* src, int count) {  // @region copy_g2s
#ifdef B2_HOSTEMU
  B2_UNROLL1
  for (int i = 0; i < count; ++i) dst[i] = src[i];
#else
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  const int n4 = count >> 2;
  int i = B2_LANE;
  B2_UNROLL1
  for (; i + 96 < n4; i += 128) {
    float4 v0 = s4[i], v1 = s4[i + 32], v2 = s4[i + 64], v3 = s4[i + 96];
    d4[i] = v0; d4[i + 32] = v1; d4[i + 64] = v2; d4[i + 96] = v3;
  }
  B2_UNROLL1
  for (; i < n4; i += 32) d4[i] = s4[i];
#endif
  B2_WSYNC();
}

// A (n x n, leading dimension ld, lower triangle valid) <- chol(A);  x <- A^-1 x.
B2_DEV void spd_factor_solve(real* A, int n, int ld, real* x, real* invd) {
  chol_factor(A, n, ld, invd, x);
  chol_backsolve(A, n, ld, invd, x);
}
// the same for a PACKED lower triangle (b2_poff)
B2_DEV void spd_factor_solve_packed(real* A, int n, real* x, real* invd) {
  chol_factor_t<true>(A, n, 0, invd, x);
  chol_backsolve_t<true>(A, n, 0, invd, x);
}
