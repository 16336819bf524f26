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

#ifndef B2_REAL
#define B2_REAL float
#endif
typedef B2_REAL real;

#ifdef B2_HOSTEMU
#define B2_DEV static inline
#define B2_LANE 0
#define B2_NLANE 1
#define B2_WSYNC() ((void)0)
B2_DEV real b2_wsum(real x) { return x; }
B2_DEV real b2_wmax(real x) { return x; }
B2_DEV int b2_wsum_i(int x) { return x; }
B2_DEV unsigned b2_ballot(bool p) { return p ? 1u : 0u; }
B2_DEV int b2_wor_i(int x) { return x; }
B2_DEV int b2_popc(unsigned x) { return __builtin_popcount(x); }
B2_DEV int b2_popcll(unsigned long long x) { return __builtin_popcountll(x); }
B2_DEV unsigned b2_lanemask_lt() { return 0u; }
B2_DEV real b2_rsqrt(real x) { return (real)1 / sqrt(x); }
#define B2_LDG(p) (*(p))
#else
#define B2_DEV __device__ __forceinline__
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
B2_DEV unsigned b2_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
B2_DEV int b2_popc(unsigned x) { return __popc(x); }
B2_DEV int b2_popcll(unsigned long long x) { return __popcll(x); }
B2_DEV unsigned b2_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
B2_DEV real b2_rsqrt(real x) { return (real)1 / sqrtf(x); }
#define B2_LDG(p) __ldg(p)
#endif

#define B2_MINVAL ((real)1e-15)
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
  const real* mf[B2_MF_COUNT];
  int mf_stride[B2_MF_COUNT];
  const int* mi[B2_MI_COUNT];
};
struct KData {
  int nworld, nconmax, njmax;
  real* df[B2_DF_COUNT];
  int* di[B2_DI_COUNT];
};

#define MF(id) (m.mf[id] + (size_t)w * m.mf_stride[id])
#define MI(id) (m.mi[id])

// ---------------------------------------------------------------------------
// small vector / quaternion math (row-major 3x3)
// ---------------------------------------------------------------------------
B2_DEV real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
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
// dense SPD helpers for one world (n <= 64), matrix row-major with stride n.
// ---------------------------------------------------------------------------
// In-place lower Cholesky A = L L^T (upper triangle left untouched).
B2_DEV void chol_factor(real* A, int n) {
  for (int k = 0; k < n; ++k) {
    B2_WSYNC();
    real dkk = A[k * n + k];
    real inv = b2_rsqrt(dkk > B2_MINVAL ? dkk : B2_MINVAL);
    B2_WSYNC();
    for (int i = k + B2_LANE; i < n; i += B2_NLANE) A[i * n + k] *= inv;  // diagonal becomes sqrt(dkk)
    B2_WSYNC();
    // trailing update of the lower triangle: one row per lane
    for (int i = k + 1 + B2_LANE; i < n; i += B2_NLANE) {
      real lik = A[i * n + k];
      for (int j = k + 1; j <= i; ++j) A[i * n + j] -= lik * A[j * n + k];
    }
  }
  B2_WSYNC();
}
// x <- (L L^T)^-1 x
B2_DEV void chol_solve(const real* L, real* x, int n) {
  for (int k = 0; k < n; ++k) {
    B2_WSYNC();
    real xk = x[k] / L[k * n + k];
    B2_WSYNC();
    if (B2_LANE == 0) x[k] = xk;
    for (int i = k + 1 + B2_LANE; i < n; i += B2_NLANE) x[i] -= L[i * n + k] * xk;
  }
  for (int k = n - 1; k >= 0; --k) {
    B2_WSYNC();
    real xk = x[k] / L[k * n + k];
    B2_WSYNC();
    if (B2_LANE == 0) x[k] = xk;
    for (int i = B2_LANE; i < k; i += B2_NLANE) x[i] -= L[k * n + i] * xk;
  }
  B2_WSYNC();
}
