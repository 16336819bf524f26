// b2_solve.cuh - stage 2 of the physics step, one warp per world:
//   constraint rows (friction loss, joint limits, pyramidal contacts) -> efc_J / efc_D / efc_aref
//   -> primal Newton solver with exact line search -> qacc, efc_force, qfrc_constraint
//   -> sensors (IMU trio, subtree angular momentum, contact sensors, ...)
//   -> integration (implicitfast | Euler), time, warm start
//
// Restates mj_makeConstraint / mj_fwdConstraint / mj_sensor* / mj_implicit as
// executed by mujoco_warp's `_src.constraint`, `_src.solver`, `_src.sensor`,
// `_src.derivative` modules.
#pragma once
#include "b2_common.cuh"

// Per-world shared-memory layout for stage 2 (offsets in reals).
//
// Constraint Jacobian rows are stored COMPACT: a row only keeps the entries of the dofs on the
// kinematic chains of the two bodies in contact (the other entries are structurally zero).
// Consecutive contacts between the same pair of bodies form a RUN and share one dof list, so
// J^T D J is accumulated run by run over w*(w+1)/2 dof pairs instead of nv*(nv+1)/2
// (G1 foot-ground contact: w = 12 of nv = 35).
struct SolveSmem {
  int H, cdof, qvel, qacc, Ma, grad, search, Mv, qfs, qas, qfc, tmp, invd, pairtab, actmask, gold, mgold, floss;  // fixed part, per world
  int fixed_total;  // reals per world
  int arena_cap;    // reals of the row arena shared by the worlds of one CTA (>= the demand of one worst-case world)
  int rows_max;     // bound on constraint rows per world
};

// Row storage is carved from an arena per CTA by actual demand (median world: ~30 rows; worst case: ~200), so the
// resident warps per SM are set by what worlds typically need, not by the worst case.  Arena demand in reals:
//   D, frc/aref, jar, jv/wgt, rowinfo [rows each] | runinfo [2 runs] | rundofs [dofs bytes] | cinfo [4 ncon] | (elliptic cone: cpar [12 ncon]) | jpool [pool]
static inline
#ifndef B2_HOSTEMU
__host__ __device__
#endif
int solve_arena_need(int rows, int runs, int dofs, int pool, int ncon, int cone = 0) {
  auto r4 = [](int n) { return (n + 3) & ~3; };
  int cd = r4(2 * r4(rows));
  return cd + 3 * r4(rows) + r4(2 * runs) + r4((dofs + 3) / 4) + r4(4 * ncon) + (cone ? r4(12 * ncon) : 0) + r4(pool);
}

// True upper bounds on the rows / runs / dof-list entries / Jacobian pool of ONE world (the arena of a CTA holds one such world).
struct SolveCaps { int rows, runs, dofs, pool; };
static inline SolveCaps solve_caps(int nlimited, int nfriction, int nconmax, int njmax, int maxchain) {
  SolveCaps c;
  const int single = 2 * nlimited + nfriction;   // rows of width 1: friction loss, joint limits
  c.rows = 4 * nconmax + single;
  if (c.rows > njmax) c.rows = njmax;
  if (c.rows < 1) c.rows = 1;
  c.runs = single + nconmax + 1;
  c.pool = single + 4 * nconmax * (maxchain | 1);
  if (c.pool > 65535) c.pool = 65535;
  c.dofs = single + nconmax * maxchain;
  return c;
}

static inline SolveSmem solve_smem_layout(int nv, int rows_max, int maxchain, int arena_cap, int min_fixed = 0, int nfriction = 0) {
  SolveSmem s;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  // qM stays in HBM/L2 (read coalesced through its symmetry), not staged
  s.H = take(nv * b2_ld(nv));
  s.cdof = s.H;  // cdof is only read while the Jacobian rows are assembled, H is first written in the Newton loop
  s.qvel = take(nv); s.qacc = take(nv); s.Ma = take(nv); s.grad = take(nv); s.search = take(nv); s.Mv = take(nv);
  s.qfs = take(nv); s.qas = take(nv); s.qfc = take(nv); s.tmp = take(nv); s.invd = take(64);  // 1 / diag(L)
  s.pairtab = take((maxchain * (maxchain + 1) / 2 + 1) / 2);  // ushort (a << 8 | b) per lower-triangle pair index
  s.actmask = take((rows_max + 31) / 32);  // bit r: row r was active at the last Hessian factorisation
  s.gold = take(nv); s.mgold = take(nv);   // CG solver: previous gradient and previous M^-1 gradient (Polak-Ribiere)
  s.floss = take(nfriction > 0 ? nv : 0);  // friction-loss rows (rows 0..nf-1, nf <= nv): the row's force bound
  s.fixed_total = o > min_fixed ? o : ((min_fixed + 3) & ~3);  // the fused sensor stage reuses the warp's fixed part as its scratch
  s.arena_cap = arena_cap;
  s.rows_max = rows_max;
  return s;
}

// impedance / reference parameters of one row (mj_makeImpedance restated)
B2_DEV void kbi(real timestep, const real* solref, const real* solimp, real pos, real& k, real& b, real& imp) {  // @region kbi
  real timeconst = solref[0], dampratio = solref[1];
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (timeconst > 0) timeconst = fmax(timeconst, 2 * timestep);
  dmin = clampr(dmin, B2_MINIMP, B2_MAXIMP);
  dmax = clampr(dmax, B2_MINIMP, B2_MAXIMP);
  width = fmax(B2_MINVAL, width);
  mid = clampr(mid, B2_MINIMP, B2_MAXIMP);
  power = fmax((real)1, power);
  if (solref[0] > 0) k = (real)1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  else k = -solref[0] / (dmax * dmax);
  if (solref[1] > 0) b = (real)2 / (dmax * timeconst);
  else b = -solref[1] / dmax;
  real x = fabs(pos) / width;
  if (x >= 1) {
    imp = dmax;
  } else {
    real y;
    if (power == 2) {  // MuJoCo's default solimp power: same value as pow(), without fetching the powf code
      if (x < mid) y = x * x / mid;
      else y = 1 - (1 - x) * (1 - x) / (1 - mid);
    } else if (x < mid) y = pow(x, power) / pow(mid, power - 1);
    else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
    imp = dmin + y * (dmax - dmin);
    real lo = fmin(dmin, dmax), hi = fmax(dmin, dmax);
    imp = clampr(imp, lo, hi);
  }
}

// y[r] = J[r,:] . x (- sub[r]) over compact rows; lanes over rows.
B2_DEVN void rows_dot(const real* jpool, const int* rowinfo, const int* runinfo, const unsigned char* rundofs, int nefc,  // @region rows_dot
                     const real* x, real* y, const real* sub) {
  B2_UNROLL1
  for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
    int info = rowinfo[r];
    int off = info & 0xffff, run = info >> 16;
    int ri = runinfo[2 * run];
    int dofoff = ri & 0xffff, wd = ri >> 16;
    const real* jr = jpool + off;
    const unsigned char* dl = rundofs + dofoff;
    real acc = 0;
    B2_UNROLL_HOT
    for (int a = 0; a < wd; ++a) acc += jr[a] * x[dl[a]];
    y[r] = sub ? acc - sub[r] : acc;
  }
  B2_WSYNC();
}

struct LsPnt { real alpha, cost, d1, d2; };
#if defined(B2_TIMING) && !defined(B2_HOSTEMU)
__device__ unsigned long long g_lsstat[8];   // [heavy?0:4] + {line searches, single evaluations, 3-point rounds, searches that reached refinement}
#define B2_LSSTAT(k) { if (B2_LANE == 0) atomicAdd(&g_lsstat[(nefc > 60 ? 0 : 4) + (k)], 1ull); }
#else
#define B2_LSSTAT(k)
#endif
#ifdef B2_HOSTEMU
static long long b2h_ls_passes = 0;   // test harness statistic: row passes of the line search (single or three-point)
static long long b2h_flip_hist[66] = {0};   // [k < 64]: Newton iterations (it > 0) whose active set differs in k rows from the last factorised one; [64]: >= 64; [65]: first iterations
#define B2_LS_COUNT() (++b2h_ls_passes)
#else
#define B2_LS_COUNT()
#endif
// float32 only: treat |d1| below B2_LS_FLOOR_MULT * eps * (|alpha| d2 + |q1|) - the rounding noise of the derivative itself - as
// converged.  MuJoCo's gradient tolerance (tolerance * ls_tolerance * |search| / scale ~ 1e-10) is below that noise in
// float32, so without the floor the search keeps refining digits that do not exist (measured on the G1 batch: 11.5 row
// passes per line search without it, 3.9 with it, same Newton iteration counts, same error against the float64 solution).  0 = off.
#ifndef B2_LS_FLOOR_MULT
#define B2_LS_FLOOR_MULT 16
#endif
#ifndef B2_LS_SKIP_P0
#define B2_LS_SKIP_P0 1     // float32 only: take the line search's alpha = 0 point from the Newton update instead of a row pass
#endif
#ifndef B2_GRAD_FLOOR_MULT
#define B2_GRAD_FLOOR_MULT 4   // float32 only: |gradient| < mult * eps * |(M a, qfrc_smooth, qfrc_constraint)| counts as converged; 0 = off
#endif
#ifndef B2_GRAD_KEEP_MULT
#define B2_GRAD_KEEP_MULT 100   // float32 only: do not stop on "no cost improvement" while |gradient| > mult * its rounding noise; 0 = off
#endif
#ifndef B2_FLOOR_MULT
#define B2_FLOOR_MULT 8   // rounding floor of the convergence test, in units of eps * |cost| (experiments: tools/floor_experiment.py)
#endif

// Friction-loss row (mj_constraintUpdate, mjCNSTR_FRICTION_DOF): Huber cost of the row's residual x with force bound f and
// R = 1 / D - quadratic inside |x| < R f, linear outside, where the force saturates at -/+ f.  Returns the cost; `force` is
// the row's efc_force, `wgt` its Hessian weight (D in the quadratic zone, else 0).
B2_DEV real fric_row(real x, real D, real f, real& force, real& wgt) {
  const real Rf = f / D;
  if (x <= -Rf) { force = f; wgt = 0; return f * (-(real)0.5 * Rf - x); }
  if (x >= Rf) { force = -f; wgt = 0; return f * (-(real)0.5 * Rf + x); }
  force = -D * x; wgt = D;
  return (real)0.5 * D * x * x;
}
// its contribution to the line search's quadratic q0 + alpha q1 + alpha^2 q2 at the trial point x = jar + alpha jv (mj PrimalEval)
B2_DEV void fric_quad(real x, real ja, real v, real D, real f, real& q0, real& q1, real& q2) {
  const real Rf = f / D;
  if (x <= -Rf) { q0 += f * (-(real)0.5 * Rf - ja); q1 -= f * v; }
  else if (x >= Rf) { q0 += f * (-(real)0.5 * Rf + ja); q1 += f * v; }
  else { const real dj = D * ja; q0 += (real)0.5 * dj * ja; q1 += dj * v; q2 += (real)0.5 * D * v * v; }
}

// ---- elliptic friction cones (opt.cone = mjCONE_ELLIPTIC; mj_constraintUpdate / PrimalEval / HessianCone restated) ----------
// A frictional contact (condim 3) owns three rows along its frame axes (normal, two tangents).  Its rows carry NEGATED D in the
// solver's D array: that marks them for the per-row passes (which skip them); the per-contact passes below recover D_j = -D[adr + j].
// With U = (mu x_n, f_1 x_t1, f_2 x_t2), N = U_0 and T = |U_1, U_2|: N >= mu T satisfied (no cost); mu N + T <= 0 fully violated
// (every row quadratic, weight D_j: the sticking contact); else the residual lies between the two cones and pays
// 0.5 Dm (N - mu T)^2, Dm = D_n / (mu^2 (1 + mu^2)): the sliding contact, force on the cone's surface.
// Per contact the solver keeps B2_CPAR reals (arena): [mu, Dm, f1, f2 | a0, a1, a2, b1, b2, Dm kd | -, -]; the second group is
// the cone Hessian of a sliding contact as two weighted rank-one terms, Hc = Dm a a^T + Dm kd b b^T with
// a = (mu, -mu f1 n1, -mu f2 n2), b = (0, -f1 n2, f2 n1), n = (U_1, U_2) / T, kd = mu^2 - mu N / T.
// All of this sits behind the template parameter EL: pyramidal models run instantiations that carry none of it.
#define B2_CPAR 12
struct ConeNone {};
struct ConeArgs { int* cinfo; real* cpar; int ncon; };   // cinfo: 4 ints per contact (row address | -1, dim | zone << 8, body1, body2)
template <bool EL> struct ConeSel { typedef ConeNone type; };
template <> struct ConeSel<true> { typedef ConeArgs type; };

B2_DEV int cone_zone(real N, real T, real mu) {
  if (N >= mu * T || (T <= 0 && N >= 0)) return 0;
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) return 1;
  return 2;
}
// One contact's residual ray jar + alpha jv in cone coordinates, set up once per line-search pass
struct ConeRay { real mu, Dm, U0, V0, UU, UV, VV, z0, z1, z2; };
B2_DEV bool cone_ray(const ConeArgs& ca, int c, const real* jar, const real* jv, const real* D, ConeRay& r) {
  const int adr = ca.cinfo[4 * c];
  if (adr < 0 || (ca.cinfo[4 * c + 1] & 0xff) < 2) return false;
  const real* cp = ca.cpar + B2_CPAR * c;
  r.mu = cp[0]; r.Dm = cp[1];
  const real x0 = jar[adr], x1 = jar[adr + 1], x2 = jar[adr + 2], v0 = jv[adr], v1 = jv[adr + 1], v2 = jv[adr + 2];
  const real u1 = x1 * cp[2], u2 = x2 * cp[3], w1 = v1 * cp[2], w2 = v2 * cp[3];
  r.U0 = x0 * r.mu; r.V0 = v0 * r.mu;
  r.UU = u1 * u1 + u2 * u2; r.UV = u1 * w1 + u2 * w2; r.VV = w1 * w1 + w2 * w2;
  const real D0 = -D[adr], D1 = -D[adr + 1], D2 = -D[adr + 2];   // the fully violated zone is an ordinary quadratic in alpha
  r.z0 = (real)0.5 * (D0 * x0 * x0 + D1 * x1 * x1 + D2 * x2 * x2);
  r.z1 = D0 * x0 * v0 + D1 * x1 * v1 + D2 * x2 * v2;
  r.z2 = (real)0.5 * (D0 * v0 * v0 + D1 * v1 * v1 + D2 * v2 * v2);
  return true;
}
// its contribution at alpha: a sticking contact adds to the pass's quadratic sums like any active row; a sliding one adds its
// own cost / slope / curvature (e0, e1, e2), which are summed apart - folding them into the quadratic's coefficients would
// cancel alpha e2 against itself in the slope, and e2 grows like 1 / T near the cone's axis (measured: 1e-3 force error in float32)
// (T and 1 / T are the IEEE square root and division: with rsqrt - 4.41 -> 4.29 ms per env step - a G1 world stopped an iteration early
// and missed the 1e-4 force parity (1.5e-3), also with rsqrt kept for the directions only: measured, reverted.  A non-inlined single copy of this function was measured: its
// by-reference arguments go through local memory, 4.80 against 4.41 ms per env step - reverted)
B2_DEV void cone_at(const ConeRay& r, real alpha, real& q0, real& q1, real& q2, real& e0, real& e1, real& e2) {
  const real N = r.U0 + alpha * r.V0;
  const real Tsq = r.UU + alpha * (2 * r.UV + alpha * r.VV);
  const real T = Tsq > 0 ? sqrt(Tsq) : (real)0;
  const int zone = cone_zone(N, T, r.mu);
  if (zone == 1) { q0 += r.z0; q1 += r.z1; q2 += r.z2; }
  else if (zone == 2) {
    const real rT = (real)1 / T;
    const real T1 = (r.UV + alpha * r.VV) * rT, T2 = (r.VV - T1 * T1) * rT;
    const real NmT = N - r.mu * T, s1 = r.V0 - r.mu * T1;
    e0 += (real)0.5 * r.Dm * NmT * NmT; e1 += r.Dm * NmT * s1; e2 += r.Dm * (s1 * s1 - NmT * r.mu * T2);
  }
}
// constraint update of the cone contacts at the current residual: returns this lane's share of the cost; with `write` also the
// rows' forces / Hessian weights (sticking), the zone (cinfo[4 c + 1] >> 8) and the sliding contacts' Hessian terms (cpar[4..9]).
// `changed` is set when a contact's zone differs from the last update's or the contact slides (its Hessian then moves with the
// residual): the factor of the previous Newton iteration cannot be reused.
B2_DEVN real cone_update(ConeArgs ca, const real* jar, const real* D, real* wgt, real* frc, bool write, bool& changed) {
  real cost = 0;
  B2_UNROLL1
  for (int c = B2_LANE; c < ca.ncon; c += B2_NLANE) {
    const int adr = ca.cinfo[4 * c], dz = ca.cinfo[4 * c + 1];
    if (adr < 0 || (dz & 0xff) < 2) continue;
    real* cp = ca.cpar + B2_CPAR * c;
    const real mu = cp[0], Dm = cp[1], f1 = cp[2], f2 = cp[3];
    const real x0 = jar[adr], x1 = jar[adr + 1], x2 = jar[adr + 2];
    const real N = x0 * mu, U1 = x1 * f1, U2 = x2 * f2;
    const real TT = U1 * U1 + U2 * U2;
    const real T = TT > 0 ? sqrt(TT) : (real)0;
    const int zone = cone_zone(N, T, mu);
    if (write) {
      ca.cinfo[4 * c + 1] = (dz & 0xff) | (zone << 8);
      if (zone != (dz >> 8) || zone == 2) changed = true;
    }
    if (zone == 1) {
      const real D0 = -D[adr], D1 = -D[adr + 1], D2 = -D[adr + 2];
      cost += (real)0.5 * (D0 * x0 * x0 + D1 * x1 * x1 + D2 * x2 * x2);
      if (write) {
        wgt[adr] = D0; wgt[adr + 1] = D1; wgt[adr + 2] = D2;
        frc[adr] = -D0 * x0; frc[adr + 1] = -D1 * x1; frc[adr + 2] = -D2 * x2;
      }
    } else if (zone == 2) {
      const real NmT = N - mu * T;
      cost += (real)0.5 * Dm * NmT * NmT;
      if (write) {
        const real rT = (real)1 / T;
        const real n1 = U1 * rT, n2 = U2 * rT;
        const real fn = -Dm * NmT * mu;
        frc[adr] = fn; frc[adr + 1] = -fn * n1 * f1; frc[adr + 2] = -fn * n2 * f2;
        cp[4] = mu; cp[5] = -mu * f1 * n1; cp[6] = -mu * f2 * n2;
        cp[7] = -f1 * n2; cp[8] = f2 * n1; cp[9] = Dm * (mu * mu - mu * N * rT);
      }
    }
  }
  return cost;
}

// FR: the world may hold friction-loss rows (rows 0..nf-1, bounds in floss).  A template parameter like CG below: models
// without friction loss run instantiations that carry none of this code.
template <bool FR, bool EL>
B2_DEVN LsPnt ls_eval(real alpha, real q0g, real q1g, real q2g, const real* jar, const real* jv, const real* D, int nefc, int nf, const real* floss, typename ConeSel<EL>::type ca) {  // @region linesearch
  B2_LS_COUNT();
  real q0 = 0, q1 = 0, q2 = 0;
  B2_UNROLL1   // (unrolling by 2 / 4 to overlap the passes' loads was measured: +2 % kernel time, code size again)
  for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
    real x = jar[r] + alpha * jv[r];
    if (EL && D[r] < 0) continue;   // cone rows: per contact below
    if (FR && r < nf) {
      fric_quad(x, jar[r], jv[r], D[r], floss[r], q0, q1, q2);
    } else if (x < 0) {
      real dj = D[r] * jar[r];
      q0 += (real)0.5 * dj * jar[r];
      q1 += dj * jv[r];
      q2 += (real)0.5 * D[r] * jv[r] * jv[r];
    }
  }
  real e0 = 0, e1 = 0, e2 = 0;
  if constexpr (EL) {
    B2_UNROLL1
    for (int c = B2_LANE; c < ca.ncon; c += B2_NLANE) {
      ConeRay cr;
      if (cone_ray(ca, c, jar, jv, D, cr)) cone_at(cr, alpha, q0, q1, q2, e0, e1, e2);
    }
    b2_wsum3(e0, e1, e2);
  }
  b2_wsum3(q0, q1, q2);
  q0 += q0g; q1 += q1g; q2 += q2g;
  LsPnt p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d1 = 2 * alpha * q2 + q1;
  p.d2 = 2 * q2;
  if constexpr (EL) { p.cost += e0; p.d1 += e1; p.d2 += e2; }
  if (p.d2 <= 0) p.d2 = B2_MINVAL;
  return p;
}

// Three points in one pass over the rows.  The refinement phase of the line search needs up to three evaluations whose
// abscissae are all known at the same time (the Newton steps from both bracket ends and the next midpoint); evaluating them
// together reads each row once and lets the three warp reductions overlap instead of running back to back - the line search
// is a chain of dependent reductions, and on the heaviest worlds (which set the kernel's duration) it is a third of the time.
// Each point's sums are accumulated in the same order as ls_eval does, so the values are bit-identical to three calls.
template <bool FR, bool EL>
B2_DEVN void ls_eval3(const real* alpha, LsPnt* out, real q0g, real q1g, real q2g, const real* jar, const real* jv, const real* D, int nefc, int nf, const real* floss, typename ConeSel<EL>::type ca) {
  B2_LS_COUNT();
  real a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0, c0 = 0, c1 = 0, c2 = 0;
  const real al0 = alpha[0], al1 = alpha[1], al2 = alpha[2];
  B2_UNROLL1
  for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
    const real ja = jar[r], v = jv[r], Dr = D[r];
    if (EL && Dr < 0) continue;
    if (FR && r < nf) {
      const real f = floss[r];
      fric_quad(ja + al0 * v, ja, v, Dr, f, a0, a1, a2);
      fric_quad(ja + al1 * v, ja, v, Dr, f, b0, b1, b2);
      fric_quad(ja + al2 * v, ja, v, Dr, f, c0, c1, c2);
      continue;
    }
    const real dj = Dr * ja;
    const real t0 = (real)0.5 * dj * ja, t1 = dj * v, t2 = (real)0.5 * Dr * v * v;
    if (ja + al0 * v < 0) { a0 += t0; a1 += t1; a2 += t2; }
    if (ja + al1 * v < 0) { b0 += t0; b1 += t1; b2 += t2; }
    if (ja + al2 * v < 0) { c0 += t0; c1 += t1; c2 += t2; }
  }
  real ea[3] = {0, 0, 0}, eb[3] = {0, 0, 0}, ec[3] = {0, 0, 0};   // sliding contacts: cost / slope / curvature at the three points
  if constexpr (EL) {
    B2_UNROLL1
    for (int c = B2_LANE; c < ca.ncon; c += B2_NLANE) {
      ConeRay cr;
      if (!cone_ray(ca, c, jar, jv, D, cr)) continue;
      cone_at(cr, al0, a0, a1, a2, ea[0], ea[1], ea[2]);
      cone_at(cr, al1, b0, b1, b2, eb[0], eb[1], eb[2]);
      cone_at(cr, al2, c0, c1, c2, ec[0], ec[1], ec[2]);
    }
    b2_wsum3(ea[0], ea[1], ea[2]);
    b2_wsum3(eb[0], eb[1], eb[2]);
    b2_wsum3(ec[0], ec[1], ec[2]);
  }
  b2_wsum3(a0, a1, a2);
  b2_wsum3(b0, b1, b2);
  b2_wsum3(c0, c1, c2);
  const real q0[3] = {a0 + q0g, b0 + q0g, c0 + q0g}, q1[3] = {a1 + q1g, b1 + q1g, c1 + q1g}, q2[3] = {a2 + q2g, b2 + q2g, c2 + q2g};
  for (int k = 0; k < 3; ++k) {
    const real al = k == 0 ? al0 : (k == 1 ? al1 : al2);
    LsPnt p;
    p.alpha = al;
    p.cost = al * al * q2[k] + al * q1[k] + q0[k];
    p.d1 = 2 * al * q2[k] + q1[k];
    p.d2 = 2 * q2[k];
    if constexpr (EL) {
      const real* e = k == 0 ? ea : (k == 1 ? eb : ec);
      p.cost += e[0]; p.d1 += e[1]; p.d2 += e[2];
    }
    if (p.d2 <= 0) p.d2 = B2_MINVAL;
    out[k] = p;
  }
}

// derivative tolerance at point p: MuJoCo's gtol, or the rounding noise of d1 in float32 when that is larger (see B2_LS_FLOOR_MULT)
B2_DEV real ls_gtol(const LsPnt& p, real gtol) {
  if (sizeof(real) == 4 && B2_LS_FLOOR_MULT > 0) {
    const real q1 = p.d1 - p.alpha * p.d2;
    const real noise = (real)B2_LS_FLOOR_MULT * (real)1.1920929e-7 * (fabs(p.alpha) * p.d2 + fabs(q1));
    return noise > gtol ? noise : gtol;
  }
  return gtol;
}

// Bracket update of MuJoCo's line search: move end point p to the best candidate on its side of the minimum.
B2_DEV bool ls_move_bracket(LsPnt& p, const LsPnt* cand) {
  bool flag = false;
  for (int i = 0; i < 3; ++i) {
    if (p.d1 < 0 && cand[i].d1 < 0 && p.d1 < cand[i].d1) { p = cand[i]; flag = true; }
    else if (p.d1 > 0 && cand[i].d1 > 0 && p.d1 > cand[i].d1) { p = cand[i]; flag = true; }
  }
  return flag;
}

// Exact 1-D minimisation along `search` (MuJoCo's iterative primal line search:
// Newton steps until the derivative changes sign, then bracket refinement).
// `cost0` / `slope0`: cost and directional derivative grad.search at alpha = 0, known from the Newton update.
template <bool FR, bool EL>
B2_DEVN real linesearch(real q0g, real q1g, real q2g, real snorm, real tolerance, real ls_tolerance, int ls_iterations,
                       real scale, const real* jar, const real* jv, const real* D, int nefc, real cost0, real slope0, int nf, const real* floss,
                       typename ConeSel<EL>::type ca) {
  if (snorm < B2_MINVAL) return 0;
  real gtol = tolerance * ls_tolerance * snorm / scale;
  int nev = 0;
  B2_LSSTAT(0) B2_LSSTAT(1)
  LsPnt p0;
  if (sizeof(real) == 4 && B2_LS_SKIP_P0 && slope0 < 0) {
    // float32 builds: the point alpha = 0 needs no pass over the rows.  Its cost is the cost the Newton update just
    // computed, its slope is grad.search, and because search solves H search = -grad its curvature search.H.search equals
    // -slope: the first trial point is the full Newton step alpha = 1.  (The float64 builds keep MuJoCo's evaluation so
    // that they reproduce the oracle digit for digit.)
    p0.alpha = 0; p0.cost = cost0; p0.d1 = slope0; p0.d2 = -slope0;
  } else {
    B2_LSSTAT(1)
    p0 = ls_eval<FR, EL>(0, q0g, q1g, q2g, jar, jv, D, nefc, nf, floss, ca);
  }
  LsPnt p1 = ls_eval<FR, EL>(p0.alpha - p0.d1 / p0.d2, q0g, q1g, q2g, jar, jv, D, nefc, nf, floss, ca); ++nev;
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.d1) < ls_gtol(p1, gtol)) return p1.alpha;
  real dir = p1.d1 < 0 ? (real)1 : (real)-1;
  LsPnt p2 = p1;
  bool p2set = false;
  while (p1.d1 * dir <= -ls_gtol(p1, gtol) && nev < ls_iterations) {
    p2 = p1; p2set = true;
    B2_LSSTAT(1)
    p1 = ls_eval<FR, EL>(p1.alpha - p1.d1 / p1.d2, q0g, q1g, q2g, jar, jv, D, nefc, nf, floss, ca); ++nev;
    if (fabs(p1.d1) < ls_gtol(p1, gtol)) return p1.alpha;
  }
  if (nev >= ls_iterations || !p2set) return p1.alpha;
  // Refinement.  The reference sequence is: evaluate p1next; loop { evaluate pmid; test; move the bracket ends, evaluating
  // the Newton step from every end that moved }.  Every abscissa of one round - the Newton steps of the ends that moved and
  // the midpoint of the NEXT iteration (needed iff the evaluation budget allows another iteration) - is known at the
  // same time, so each round is one ls_eval3 pass; `nev` advances exactly as in the one-at-a-time sequence.
  LsPnt p2next = p1, p1next, pmid;
  B2_LSSTAT(3)
  {
    B2_LSSTAT(2)
    ++nev;                                              // p1next
    const bool want_mid = nev < ls_iterations;
    real al[3] = {p1.alpha - p1.d1 / p1.d2, want_mid ? (real)0.5 * (p1.alpha + p2.alpha) : (real)0, 0};
    LsPnt out[3];
    ls_eval3<FR, EL>(al, out, q0g, q1g, q2g, jar, jv, D, nefc, nf, floss, ca);
    p1next = out[0]; pmid = out[1];
  }
  while (nev < ls_iterations) {
    ++nev;                                              // pmid (evaluated in the previous round)
    LsPnt cand[3] = {p1next, p2next, pmid};
    int best = -1;
    for (int i = 0; i < 3; ++i)
      if (fabs(cand[i].d1) < ls_gtol(cand[i], gtol) && (best < 0 || cand[i].cost < cand[best].cost)) best = i;
    if (best >= 0) return cand[best].alpha;
    const bool b1 = ls_move_bracket(p1, cand);
    const bool b2 = ls_move_bracket(p2, cand);
    if (!b1 && !b2) return pmid.cost < p0.cost ? pmid.alpha : (real)0;
    nev += (b1 ? 1 : 0) + (b2 ? 1 : 0);
    const bool want_mid = nev < ls_iterations;
    real al[3] = {b1 ? p1.alpha - p1.d1 / p1.d2 : (real)0, b2 ? p2.alpha - p2.d1 / p2.d2 : (real)0,
                  want_mid ? (real)0.5 * (p1.alpha + p2.alpha) : (real)0};
    LsPnt out[3];
    B2_LSSTAT(2)
    ls_eval3<FR, EL>(al, out, q0g, q1g, q2g, jar, jv, D, nefc, nf, floss, ca);
    if (b1) p1next = out[0];
    if (b2) p2next = out[1];
    pmid = out[2];
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}

// contact-frame force [normal, t1, t2] decoded from the pyramid edge forces
B2_DEV void contact_force(real* out, int adr, int dim, const real* fri, const real* efc_force, int cone) {  // @region sensor_helpers
  out[0] = out[1] = out[2] = 0;
  if (adr < 0) return;
  if (dim == 1) { out[0] = efc_force[adr]; return; }
  if (cone == 1) {   // elliptic: the rows are the frame axes
    for (int k = 0; k < dim && k < 3; ++k) out[k] = efc_force[adr + k];
    return;
  }
  for (int k = 1; k < dim; ++k) {
    int r = adr + 2 * (k - 1);
    real f1 = efc_force[r], f2 = efc_force[r + 1];
    out[0] += f1 + f2;
    if (k < 3) out[k] = (f1 - f2) * fri[k - 1];
  }
}

// is `root` the body itself or one of its ancestors?  ancmask != NULL (nbody <= 64): one table lookup; else walk up.
B2_DEV bool in_subtree(const int* body_parent, const int* ancmask, int body, int root) {
  if (ancmask) return ((unsigned)ancmask[2 * body + (root >> 5)] >> (root & 31)) & 1u;
  for (int b = body;; b = body_parent[b]) {
    if (b == root) return true;
    if (b == 0) return false;
  }
}
B2_DEV bool match_side(const int* body_parent, const int* ancmask, int objtype, int objid, int geom, int body) {
  if (objid < 0) return true;
  if (objtype == B2_OBJ_GEOM) return geom == objid;
  if (objtype == B2_OBJ_BODY) return body == objid;
  if (objtype == B2_OBJ_XBODY) return in_subtree(body_parent, ancmask, body, objid);
  return false;
}

// the same with the body's ancestor mask in hand (lo: bodies 0..31, hi: 32..63)
B2_DEV bool match_side_m(int objtype, int objid, int geom, int body, unsigned lo, unsigned hi) {
  if (objid < 0) return true;
  if (objtype == B2_OBJ_GEOM) return geom == objid;
  if (objtype == B2_OBJ_BODY) return body == objid;
  if (objtype == B2_OBJ_XBODY) return ((objid < 32 ? lo : hi) >> (objid & 31)) & 1u;
  return false;
}

B2_DEV void mat2quat(real* q, const real* R) {
  real tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    real s = sqrt(tr + 1) * 2;
    q[0] = (real)0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    real s = sqrt(1 + R[0] - R[4] - R[8]) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = (real)0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    real s = sqrt(1 + R[4] - R[0] - R[8]) * 2;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = (real)0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    real s = sqrt(1 + R[8] - R[0] - R[4]) * 2;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = (real)0.25 * s;
  }
  quat_normalize(q);
}

// ---------------------------------------------------------------------------
// Stage 2 for world w.
// ---------------------------------------------------------------------------
// `sm`: this world's fixed part; `arena` / `actl`: the CTA's row arena and its control words (actl[0] reals handed out,
// actl[1] worlds of the CTA that are done).
#if defined(B2_TIMING) && !defined(B2_HOSTEMU)
__device__ unsigned long long g_phase[8 * 16384];
#define B2_PH_BEGIN() long long ph_t = clock64(); unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define B2_PH(k) { long long ph_n = clock64(); ph_acc[k] += (unsigned long long)(ph_n - ph_t); ph_t = ph_n; }
#define B2_PH_END() if (B2_LANE == 0 && w < 16384) { for (int k = 0; k < 8; ++k) g_phase[8 * w + k] = ph_acc[k]; }
#else
#define B2_PH_BEGIN()
#define B2_PH(k)
#define B2_PH_END()
#endif

// CG: compile-time choice of the search direction (true = mjSOL_CG: M-preconditioned Polak-Ribiere; false = Newton).  A template
// and not a run-time branch: the Newton kernel's loop is instruction-fetch bound, and the CG code inside it cost 5 % (0.394 ->
// 0.414 ms per launch) even though it never ran.
template <bool CG, bool FR, bool EL = false>
B2_DEV void solve_world_t(const KModel& m, const KData& d, const SolveSmem& L, int w, real* sm, real* arena, int* actl) {  // @region setup_loads
  B2_PH_BEGIN()
  const int nv = m.nv, nb = m.nbody, nq = m.nq, nu = m.nu, njmax = d.njmax, nconmax = d.nconmax;
  real* H = sm + L.H; real* cdof = sm + L.cdof;
  real* qvel = sm + L.qvel; real* qacc = sm + L.qacc; real* Ma = sm + L.Ma; real* grad = sm + L.grad;
  real* search = sm + L.search; real* Mv = sm + L.Mv; real* qfs = sm + L.qfs; real* qas = sm + L.qas;
  real* qfc = sm + L.qfc; real* tmp = sm + L.tmp; real* invd = sm + L.invd;
  unsigned short* pairtab = (unsigned short*)(sm + L.pairtab);
  unsigned* actmask = (unsigned*)(sm + L.actmask);
  real* gold = sm + L.gold; real* mgold = sm + L.mgold;
  real* floss = sm + L.floss;   // FR only
  // rows of a frictional contact: the pyramid's 2 (dim - 1) edges, or (EL) the dim axes of its frame
  auto crows = [](int dim) { return dim == 1 ? 1 : (EL ? dim : 2 * (dim - 1)); };
  constexpr bool cg = CG;   // mjSOL_CG: M is the preconditioner, Polak-Ribiere directions (mj_solCG)
  const int ld = b2_ld(nv);
  const int* body_parent = MI(B2_MI_BODY_PARENTID);
  const int* body_root = MI(B2_MI_BODY_ROOTID);
  const int* dofmask = MI(B2_MI_BODY_DOFMASK);
  const int* geom_body = MI(B2_MI_GEOM_BODYID);
  const int* jnt_type = MI(B2_MI_JNT_TYPE);
  const int* jnt_qposadr = MI(B2_MI_JNT_QPOSADR);
  const int* jnt_dofadr = MI(B2_MI_JNT_DOFADR);
  const int* jnt_limited = MI(B2_MI_JNT_LIMITED);
  const real timestep = MF(B2_MF_OPT_TIMESTEP)[0];
  const bool export_J = (d.flags & 1) != 0;

  real* g_qpos = d.df[B2_DF_QPOS] + (size_t)w * nq;
  real* g_qvel = d.df[B2_DF_QVEL] + (size_t)w * nv;
  real* gJ = d.df[B2_DF_EFC_J] + (size_t)w * njmax * nv;
  real* g_pos = d.df[B2_DF_EFC_POS] + (size_t)w * njmax;
  real* g_mar = d.df[B2_DF_EFC_MARGIN] + (size_t)w * njmax;
  const real* g_scom = d.df[B2_DF_SUBTREE_COM] + (size_t)w * nb * 3;
  const real* g_cdof = d.df[B2_DF_CDOF] + (size_t)w * nv * 6;
  const real* M = d.df[B2_DF_QM] + (size_t)w * nv * b2_ld(nv);  // [nv][ld]; symmetric, so column i is read coalesced as M[e*ld + i]

  B2_UNROLL1
  for (int i = B2_LANE; i < 6 * nv; i += B2_NLANE) cdof[i] = g_cdof[i];
  B2_UNROLL1
  for (int i = B2_LANE; i < nv; i += B2_NLANE) {
    qvel[i] = g_qvel[i];
    qfs[i] = d.df[B2_DF_QFRC_SMOOTH][(size_t)w * nv + i];
    qas[i] = d.df[B2_DF_QACC_SMOOTH][(size_t)w * nv + i];
  }
  B2_WSYNC();

  // ---- arena demand of this world (counts only; the passes below fill the arrays) --------------------------  // @region arena_alloc
  const int ncon = d.di[B2_DI_NCON][w];
  const real* c_dist = d.df[B2_DF_CONTACT_DIST] + (size_t)w * nconmax;
  const real* c_inc = d.df[B2_DF_CONTACT_INCLUDEMARGIN] + (size_t)w * nconmax;
  const int* c_geom = d.di[B2_DI_CONTACT_GEOM] + (size_t)w * nconmax * 2;
  const int* c_dim = d.di[B2_DI_CONTACT_DIM] + (size_t)w * nconmax;
  const real* jnt_range = MF(B2_MF_JNT_RANGE);
  const real* jnt_margin = MF(B2_MF_JNT_MARGIN);
  int rows_cap, runs_cap, dofs_cap, jpool_cap;
  {
    int nlim = 0, crow = 0, ckept = 0, cdofs = 0, cpool = 0;
    if (FR) {   // friction-loss rows are rows of width 1 like the limit rows: counted with them
      const real* dfl = MF(B2_MF_DOF_FRICTIONLOSS);
      B2_UNROLL1
      for (int base = 0; base < nv; base += B2_NLANE) {
        int e = base + B2_LANE;
        nlim += b2_popc(b2_ballot(e < nv && dfl[e] > 0));
      }
    }
    B2_UNROLL1
    for (int base = 0; base < 2 * m.njnt; base += B2_NLANE) {
      int s = base + B2_LANE;
      bool on = false;
      int j = s >> 1, side = s & 1;
      if (s < 2 * m.njnt && jnt_limited[j] && (jnt_type[j] == B2_JNT_HINGE || jnt_type[j] == B2_JNT_SLIDE)) {
        real q = g_qpos[jnt_qposadr[j]];
        real dist = side == 0 ? q - jnt_range[2 * j] : jnt_range[2 * j + 1] - q;
        on = dist < jnt_margin[j];
      }
      nlim += b2_popc(b2_ballot(on));
    }
    B2_UNROLL1
    for (int c = B2_LANE; c < ncon; c += B2_NLANE) {
      if (c_dist[c] < c_inc[c]) {
        int nrow = crows(c_dim[c]);
        int b1 = geom_body[c_geom[2 * c]], b2 = geom_body[c_geom[2 * c + 1]];
        int wd = b2_popc((unsigned)(dofmask[2 * b1] | dofmask[2 * b2])) + b2_popc((unsigned)(dofmask[2 * b1 + 1] | dofmask[2 * b2 + 1]));
        crow += nrow; ckept += 1; cdofs += wd; cpool += nrow * (wd | 1);
      }
    }
    crow = b2_wsum_i(crow); ckept = b2_wsum_i(ckept); cdofs = b2_wsum_i(cdofs); cpool = b2_wsum_i(cpool);
    rows_cap = nlim + crow;
    if (rows_cap > njmax) rows_cap = njmax;
    if (rows_cap > L.rows_max) rows_cap = L.rows_max;
    if (rows_cap < 1) rows_cap = 1;
    runs_cap = nlim + ckept + 1;          // every kept contact could start a run (upper bound)
    dofs_cap = nlim + cdofs;
    jpool_cap = nlim + cpool;
    if (jpool_cap > 65535) jpool_cap = 65535;
  }
  int need = solve_arena_need(rows_cap, runs_cap, dofs_cap, jpool_cap, ncon > 0 ? ncon : 1, EL);
  // A world beyond the arena (the arena may be capped below the strict worst case, see make_plan) shrinks its Jacobian pool
  // first, then its row capacity: rows that do not fit are dropped from the tail and counted in nefc_dropped, as with njmax.
  // The loop ends with need <= arena_cap (make_plan guarantees that a one-row world fits), which the allocator below
  // relies on: a block larger than the arena could never be granted.
  B2_UNROLL1
  while (need > L.arena_cap) {
    const int over = ((need - L.arena_cap) + 3) & ~3;
    if (jpool_cap > 0) jpool_cap = jpool_cap > over ? jpool_cap - over : 0;
    else if (rows_cap > 1) rows_cap = rows_cap > 8 ? rows_cap - 8 : 1;
    else break;
    need = solve_arena_need(rows_cap, runs_cap, dofs_cap, jpool_cap, ncon > 0 ? ncon : 1, EL);
  }
  if (need > L.arena_cap) need = L.arena_cap;
  int aoff = 0;
#ifndef B2_HOSTEMU
  {
    // The CTA's two warps each hold one world at a time and share the arena: warp 0's block grows from the bottom, warp
    // 1's hangs from the top.  actl[0] packs both sizes (units of 4 floats, 16 bits each); a world whose block does not fit
    // next to the other warp's current one waits until that warp releases it (k_solve clears the field after each world).
    const int me = (threadIdx.x >> 5) & 1;
    const int cap4 = L.arena_cap >> 2;
    const int need4 = ((need + 3) >> 2) < cap4 ? ((need + 3) >> 2) : cap4;
    if (B2_LANE == 0) {
      unsigned* u = reinterpret_cast<unsigned*>(actl);
      // bounded wait (the CTA-mate releases its block within a millisecond; the bound only guards against a logic error
      // turning into a hung GPU - past it the block is taken regardless)
      for (int spins = 0;; ++spins) {
        unsigned old = *reinterpret_cast<volatile unsigned*>(u);
        int other = (int)(me ? (old & 0xffffu) : (old >> 16));
        if (need4 + other <= cap4 || spins > (1 << 22)) {
          unsigned nw = me ? ((old & 0xffffu) | ((unsigned)need4 << 16)) : ((old & 0xffff0000u) | (unsigned)need4);
          if (atomicCAS(u, old, nw) == old) break;
        } else {
          __nanosleep(200);
        }
      }
    }
    __syncwarp();
    aoff = me ? (cap4 - need4) * 4 : 0;
  }
#endif
  real* D; real* aref; real* jar; real* jv; int* rowinfo; int* runinfo; unsigned char* rundofs; int* cinfo; real* jpool; real* cpar = nullptr;
  {
    real* a = arena + aoff;
    const int r4 = (rows_cap + 3) & ~3;
    D = a; aref = a + r4; a += (2 * r4 + 3) & ~3;
    jar = a; a += r4; jv = a; a += r4; rowinfo = (int*)a; a += r4;
    runinfo = (int*)a; a += (2 * runs_cap + 3) & ~3;
    rundofs = (unsigned char*)a; a += ((dofs_cap + 3) / 4 + 3) & ~3;
    cinfo = (int*)a; a += (4 * (ncon > 0 ? ncon : 1) + 3) & ~3;
    if (EL) { cpar = a; a += (B2_CPAR * (ncon > 0 ? ncon : 1) + 3) & ~3; }
    jpool = a;
  }
  real* wgt = jv;          // D*active per row; consumed by the Hessian pass before the line search reuses jv
  int* rowsrc = (int*)jv;  // contact id, or -(jnt+1); dead before the line search writes jv
  real* frc = aref;        // -D*active*jar per row; aref itself is dead once jar has been initialised

  // ---- joint-limit rows (ordered: joint, lower then upper); each is its own run of width 1 ----  // @region limit_rows
  int nefc = 0;
  if (FR) {
    // ---- friction-loss rows first (MuJoCo's row order: friction, limits, contacts), one per dof with frictionloss > 0
    const real* dfl = MF(B2_MF_DOF_FRICTIONLOSS);
    B2_UNROLL1
    for (int base = 0; base < nv; base += B2_NLANE) {
      int e = base + B2_LANE;
      const real fl = e < nv ? dfl[e] : (real)0;
      const bool on = fl > 0;
      unsigned bal = b2_ballot(on);
      if (on) {
        int row = nefc + b2_popc(bal & b2_lanemask_lt());
        if (row < rows_cap) {
          jpool[row] = 1;
          rundofs[row] = (unsigned char)e;
          rowinfo[row] = row | (row << 16);
          runinfo[2 * row] = row | (1 << 16);
          runinfo[2 * row + 1] = row | (1 << 16);
          rowsrc[row] = -(0x10000 + e);
          floss[row] = fl;
          g_pos[row] = 0;
          g_mar[row] = 0;
          if (export_J) {
            real* Jr = gJ + (size_t)row * nv;
            B2_UNROLL1
            for (int k = 0; k < nv; ++k) Jr[k] = 0;
            Jr[e] = 1;
          }
        }
      }
      nefc += b2_popc(bal);
    }
    if (nefc > rows_cap) nefc = rows_cap;
  }
  const int nf = nefc;
  {
    B2_UNROLL1
    for (int base = 0; base < 2 * m.njnt; base += B2_NLANE) {
      int s = base + B2_LANE;
      bool on = false;
      real dist = 0;
      int j = s >> 1, side = s & 1;
      if (s < 2 * m.njnt && jnt_limited[j] && (jnt_type[j] == B2_JNT_HINGE || jnt_type[j] == B2_JNT_SLIDE)) {
        real q = g_qpos[jnt_qposadr[j]];
        dist = side == 0 ? q - jnt_range[2 * j] : jnt_range[2 * j + 1] - q;
        on = dist < jnt_margin[j];
      }
      unsigned bal = b2_ballot(on);
      if (on) {
        int row = nefc + b2_popc(bal & b2_lanemask_lt());
        if (row < rows_cap) {
          real val = side == 0 ? (real)1 : (real)-1;
          int dof = jnt_dofadr[j];
          jpool[row] = val;
          rundofs[row] = (unsigned char)dof;
          rowinfo[row] = row | (row << 16);
          runinfo[2 * row] = row | (1 << 16);
          runinfo[2 * row + 1] = row | (1 << 16);
          rowsrc[row] = -(j + 1);
          g_pos[row] = dist;
          g_mar[row] = jnt_margin[j];
          if (export_J) {
            real* Jr = gJ + (size_t)row * nv;
            B2_UNROLL1
            for (int e = 0; e < nv; ++e) Jr[e] = 0;
            Jr[dof] = val;
          }
        }
      }
      nefc += b2_popc(bal);
    }
    if (nefc > rows_cap) nefc = rows_cap;
  }
  const int nl = nefc;   // rows of width 1: nf friction-loss rows, then nl - nf joint-limit rows

  // ---- contact rows ---------------------------------------------------------------------------  // @region contact_rows
  const real* c_pos = d.df[B2_DF_CONTACT_POS] + (size_t)w * nconmax * 3;
  const real* c_frame = d.df[B2_DF_CONTACT_FRAME] + (size_t)w * nconmax * 9;
  const real* c_fri = d.df[B2_DF_CONTACT_FRICTION] + (size_t)w * nconmax * 5;
  const real* c_solref = d.df[B2_DF_CONTACT_SOLREF] + (size_t)w * nconmax * 2;
  const real* c_solimp = d.df[B2_DF_CONTACT_SOLIMP] + (size_t)w * nconmax * 5;
  int* c_adr = d.di[B2_DI_CONTACT_EFC_ADDRESS] + (size_t)w * nconmax;
  int ndrop = 0;
  {
    // row addresses by prefix sum over contacts; a contact is kept iff its rows fit below the cap
    int run = nefc;
    B2_UNROLL1
    for (int base = 0; base < ncon; base += B2_NLANE) {
      int c = base + B2_LANE;
      int nrow = 0;
      if (c < ncon && c_dist[c] < c_inc[c]) nrow = crows(c_dim[c]);
      int incl = nrow;
#ifndef B2_HOSTEMU
      B2_UNROLL1
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (B2_LANE >= o) incl += t;
      }
      int total = __shfl_sync(0xffffffffu, incl, 31);
#else
      int total = incl;
#endif
      if (c < ncon) {
        int adr = run + incl - nrow;
        bool keep = nrow > 0 && adr + nrow <= rows_cap;
        if (nrow > 0 && !keep) ndrop += nrow;
        int g1 = c_geom[2 * c], g2 = c_geom[2 * c + 1];
        cinfo[4 * c] = keep ? adr : -1;
        cinfo[4 * c + 1] = c_dim[c];
        cinfo[4 * c + 2] = geom_body[g1];
        cinfo[4 * c + 3] = geom_body[g2];
      }
      run += total;
    }
    ndrop = b2_wsum_i(ndrop);
    nefc = run - ndrop;
    B2_WSYNC();
  }
  // compact Jacobian rows, contact by contact (uniform loop), lanes over the run's dof list
  int nrun = nl;
  {
    int pooloff = nl, dofcount = nl;
    int pb1 = -1, pb2 = -1, cur_dofoff = 0, cur_w = 0, cur_run = -1, cur_row0 = 0, cur_nrow = 0;
    bool overflow = false;
    B2_UNROLL1
    for (int c = 0; c < ncon; ++c) {
      int adr = cinfo[4 * c];
      if (adr < 0) continue;
      int dim = cinfo[4 * c + 1], b1 = cinfo[4 * c + 2], b2 = cinfo[4 * c + 3];
      int nrow = crows(dim);
      unsigned lo1 = (unsigned)dofmask[2 * b1], hi1 = (unsigned)dofmask[2 * b1 + 1];
      unsigned lo2 = (unsigned)dofmask[2 * b2], hi2 = (unsigned)dofmask[2 * b2 + 1];
      if (!overflow && (b1 != pb1 || b2 != pb2)) {
        unsigned lo = lo1 | lo2, hi = hi1 | hi2;
        int cl = b2_popc(lo), wd = cl + b2_popc(hi);
        if (dofcount + wd > dofs_cap || nrun >= runs_cap) overflow = true;
        else {
          if (cur_run >= 0 && B2_LANE == 0) runinfo[2 * cur_run + 1] = cur_row0 | (cur_nrow << 16);
#ifndef B2_HOSTEMU
          if ((lo >> B2_LANE) & 1u) rundofs[dofcount + b2_popc(lo & b2_lanemask_lt())] = (unsigned char)B2_LANE;
          if ((hi >> B2_LANE) & 1u) rundofs[dofcount + cl + b2_popc(hi & b2_lanemask_lt())] = (unsigned char)(B2_LANE + 32);
#else
          { int k = 0; for (int e = 0; e < 64; ++e) { bool in = e < 32 ? ((lo >> e) & 1u) : ((hi >> (e - 32)) & 1u); if (in) rundofs[dofcount + k++] = (unsigned char)e; } }
#endif
          cur_run = nrun++; cur_dofoff = dofcount; cur_w = wd; cur_row0 = adr; cur_nrow = 0;
          dofcount += wd;
          if (B2_LANE == 0) runinfo[2 * cur_run] = cur_dofoff | (cur_w << 16);
          pb1 = b1; pb2 = b2;
          B2_WSYNC();
        }
      }
      const int wpad = cur_w | 1;  // odd row stride: rows of a run hit distinct banks
      if (!overflow && pooloff + nrow * wpad > jpool_cap) overflow = true;
      if (overflow) {
        // shared-memory pools exhausted: this and all later contacts lose their rows
        if (B2_LANE == 0) { cinfo[4 * c] = -1; }
        ndrop += nrow;
        continue;
      }
      int r1 = body_root[b1], r2 = body_root[b2];
      const real* p = c_pos + 3 * c;
      real o1[3] = {p[0] - g_scom[3 * r1], p[1] - g_scom[3 * r1 + 1], p[2] - g_scom[3 * r1 + 2]};
      real o2[3] = {p[0] - g_scom[3 * r2], p[1] - g_scom[3 * r2 + 1], p[2] - g_scom[3 * r2 + 2]};
      const real* F = c_frame + 9 * c;
      B2_UNROLL1
      for (int a = B2_LANE; a < cur_w; a += B2_NLANE) {
        int e = rundofs[cur_dofoff + a];
        bool in1 = e < 32 ? ((lo1 >> e) & 1u) : ((hi1 >> (e - 32)) & 1u);
        bool in2 = e < 32 ? ((lo2 >> e) & 1u) : ((hi2 >> (e - 32)) & 1u);
        real jd[3] = {0, 0, 0};
        const real* cd = cdof + 6 * e;
        if (in2) { real t[3]; cross3(t, cd, o2); jd[0] += cd[3] + t[0]; jd[1] += cd[4] + t[1]; jd[2] += cd[5] + t[2]; }
        if (in1) { real t[3]; cross3(t, cd, o1); jd[0] -= cd[3] + t[0]; jd[1] -= cd[4] + t[1]; jd[2] -= cd[5] + t[2]; }
        real jn = F[0] * jd[0] + F[1] * jd[1] + F[2] * jd[2];
        if (dim == 1) {
          jpool[pooloff + a] = jn;
        } else if (EL) {
          jpool[pooloff + a] = jn;
          B2_UNROLL1
          for (int k = 1; k < dim; ++k)
            jpool[pooloff + k * wpad + a] = k < 3 ? F[3 * k] * jd[0] + F[3 * k + 1] * jd[1] + F[3 * k + 2] * jd[2] : (real)0;
        } else {
          B2_UNROLL1
          for (int k = 1; k < dim; ++k) {
            real jt = k < 3 ? F[3 * k] * jd[0] + F[3 * k + 1] * jd[1] + F[3 * k + 2] * jd[2] : (real)0;
            real mu = c_fri[5 * c + k - 1];
            jpool[pooloff + (2 * (k - 1)) * wpad + a] = jn + mu * jt;
            jpool[pooloff + (2 * (k - 1) + 1) * wpad + a] = jn - mu * jt;
          }
        }
      }
      B2_UNROLL1
      for (int r = B2_LANE; r < nrow; r += B2_NLANE) {
        rowinfo[adr + r] = (pooloff + r * wpad) | (cur_run << 16);
        rowsrc[adr + r] = c;
        g_pos[adr + r] = c_dist[c];
        g_mar[adr + r] = c_inc[c];
      }
      pooloff += nrow * wpad;
      cur_nrow += nrow;
    }
    if (cur_run >= 0 && B2_LANE == 0) runinfo[2 * cur_run + 1] = cur_row0 | (cur_nrow << 16);
    B2_WSYNC();
    if (overflow) {
      // rows form a prefix: recompute nefc as the end of the last kept contact
      int last = nl;
      B2_UNROLL1
      for (int c = 0; c < ncon; ++c) {
        int adr = cinfo[4 * c];
        if (adr >= 0) last = adr + crows(cinfo[4 * c + 1]);
      }
      nefc = last;
    }
    B2_WSYNC();
    B2_UNROLL1
    for (int c = B2_LANE; c < ncon; c += B2_NLANE) c_adr[c] = cinfo[4 * c];
    if (B2_LANE == 0) {
      d.di[B2_DI_NEFC][w] = nefc;
      d.di[B2_DI_NL][w] = nl - nf;
      d.di[B2_DI_NF][w] = nf;
      d.di[B2_DI_NEFC_DROPPED][w] = ndrop;
    }
    // dense export for tests / API consumers that ask for efc_J
    if (export_J) {
      B2_UNROLL1
      for (int r = nl; r < nefc; ++r) {
        int info = rowinfo[r];
        int off = info & 0xffff, run = info >> 16;
        int ri = runinfo[2 * run];
        int dofoff = ri & 0xffff, wd = ri >> 16;
        B2_UNROLL1
        for (int e = B2_LANE; e < nv; e += B2_NLANE) gJ[(size_t)r * nv + e] = 0;
        B2_WSYNC();
        B2_UNROLL1
        for (int a = B2_LANE; a < wd; a += B2_NLANE) gJ[(size_t)r * nv + rundofs[dofoff + a]] = jpool[off + a];
      }
      B2_WSYNC();
    }
  }

  // pair-index -> (a, b) table of the lower triangle (a >= b), shared by every run: pr = a (a + 1) / 2 + b
  if (nefc > nl) {
    const int npairmax = m.maxchain * (m.maxchain + 1) / 2;
    B2_UNROLL1
    for (int pr = B2_LANE; pr < npairmax; pr += B2_NLANE) {
      int a = (int)((sqrt((real)(8 * pr + 1)) - 1) * (real)0.5);
      while (a * (a + 1) / 2 > pr) --a;
      while ((a + 1) * (a + 2) / 2 <= pr) ++a;
      pairtab[pr] = (unsigned short)((a << 8) | (pr - a * (a + 1) / 2));
    }
    B2_WSYNC();
  }
  // ---- per-row impedance, D, aref -----------------------------------------------------------------  // @region impedance
  real* g_efcD = d.df[B2_DF_EFC_D] + (size_t)w * njmax;
  real* g_aref = d.df[B2_DF_EFC_AREF] + (size_t)w * njmax;
  real* g_force = d.df[B2_DF_EFC_FORCE] + (size_t)w * njmax;
  if (nefc > 0) {
    rows_dot(jpool, rowinfo, runinfo, rundofs, nefc, qvel, jar, nullptr);  // jar <- J qvel (temporarily)
    const real* biw = MF(B2_MF_BODY_INVWEIGHT0);
    const real* diw = MF(B2_MF_DOF_INVWEIGHT0);
    const real* jnt_solref = MF(B2_MF_JNT_SOLREF);
    const real* jnt_solimp = MF(B2_MF_JNT_SOLIMP);
    const real impratio = MF(B2_MF_OPT_IMPRATIO)[0];
    B2_UNROLL1
    for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
      int src = rowsrc[r];
      real invw, k, b, imp;
      real pos = g_pos[r] - g_mar[r];
      if (FR && src <= -0x10000) {
        int e = -src - 0x10000;
        invw = diw[e];
        kbi(timestep, MF(B2_MF_DOF_SOLREF) + 2 * e, MF(B2_MF_DOF_SOLIMP) + 5 * e, pos, k, b, imp);
      } else if (src < 0) {
        int j = -src - 1;
        invw = diw[jnt_dofadr[j]];
        kbi(timestep, jnt_solref + 2 * j, jnt_solimp + 5 * j, pos, k, b, imp);
      } else {
        int c = src;
        invw = biw[2 * cinfo[4 * c + 2]] + biw[2 * cinfo[4 * c + 3]];
        if (!EL && cinfo[4 * c + 1] > 1) {
          real f0 = c_fri[5 * c];
          invw = (invw + f0 * f0 * invw) * 2 * f0 * f0 / impratio;
        }
        kbi(timestep, c_solref + 2 * c, c_solimp + 5 * c, pos, k, b, imp);
        if (EL && cinfo[4 * c + 1] > 1) {
          // elliptic cone (mj_makeImpedance): the friction rows take R_n / impratio (scaled by mu_1^2 / mu_j^2 for the later
          // axes) and have no position term; the negated D marks the contact's rows for the solver's per-contact passes
          const int j = r - cinfo[4 * c];
          real R = fmax(invw * (1 - imp) / imp, B2_MINVAL);
          if (j > 0) {
            R /= fmax(impratio, B2_MINVAL);
            if (j > 1) { const real f0 = c_fri[5 * c], fj = c_fri[5 * c + j - 1]; R *= f0 * f0 / (fj * fj); }
            k = 0;
          }
          const real Dr = (real)1 / R;
          const real ar = -b * jar[r] - k * imp * pos;
          D[r] = -Dr; aref[r] = ar;
          g_efcD[r] = Dr; g_aref[r] = ar;
          if (j == 0) {   // the contact's constants for the per-contact passes: mu = mu_1 sqrt(R_t1 / R_n), Dm, mu_1, mu_2
            const real f0 = c_fri[5 * c];
            const real mu = f0 * sqrt((R / fmax(impratio, B2_MINVAL)) / R);
            real* cp = cpar + B2_CPAR * c;
            cp[0] = mu; cp[1] = Dr / (mu * mu * (1 + mu * mu)); cp[2] = f0; cp[3] = c_fri[5 * c + 1];
          }
          continue;
        }
      }
      real R = fmax(invw * (1 - imp) / imp, B2_MINVAL);
      real Dr = (real)1 / R;
      real ar = -b * jar[r] - k * imp * pos;
      D[r] = Dr; aref[r] = ar;
      g_efcD[r] = Dr; g_aref[r] = ar;
    }
    B2_WSYNC();
  }

  typename ConeSel<EL>::type ca;
  if constexpr (EL) { ca.cinfo = cinfo; ca.cpar = cpar; ca.ncon = ncon; }
  bool cone_changed = false;
  // ---- solve ------------------------------------------------------------------------------------------  // @region solve_init
  B2_PH(0)
  int niter = 0;
  real* g_qacc = d.df[B2_DF_QACC] + (size_t)w * nv;
  real* g_qfc = d.df[B2_DF_QFRC_CONSTRAINT] + (size_t)w * nv;
  if (nefc == 0) {
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) { qacc[i] = qas[i]; qfc[i] = 0; }
    B2_WSYNC();
  } else {
    const real tolerance = MF(B2_MF_OPT_TOLERANCE)[0];
    const real ls_tolerance = MF(B2_MF_OPT_LS_TOLERANCE)[0];
    const real meaninertia = MF(B2_MF_STAT_MEANINERTIA)[0];
    const real scale = (real)1 / (meaninertia * (real)(nv > 1 ? nv : 1));
    const real* g_ws = d.df[B2_DF_QACC_WARMSTART] + (size_t)w * nv;
    // cost at qacc_smooth (Gauss term vanishes)
    rows_dot(jpool, rowinfo, runinfo, rundofs, nefc, qas, jar, aref);
    real cs = 0;
    B2_UNROLL1
    for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
      real x = jar[r];
      if (EL && D[r] < 0) continue;
      if (FR && r < nf) { real f_, w_; cs += fric_row(x, D[r], floss[r], f_, w_); }
      else if (x < 0) cs += (real)0.5 * D[r] * x * x;
    }
    if constexpr (EL) cs += cone_update(ca, jar, D, nullptr, nullptr, false, cone_changed);
    cs = b2_wsum(cs);
    // cost at the warm start
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) qacc[i] = g_ws[i];
    B2_WSYNC();
    rows_dot(jpool, rowinfo, runinfo, rundofs, nefc, qacc, jar, aref);
    real cw = 0;
    B2_UNROLL1
    for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
      real x = jar[r];
      if (EL && D[r] < 0) continue;
      if (FR && r < nf) { real f_, w_; cw += fric_row(x, D[r], floss[r], f_, w_); }
      else if (x < 0) cw += (real)0.5 * D[r] * x * x;
    }
    if constexpr (EL) cw += cone_update(ca, jar, D, nullptr, nullptr, false, cone_changed);
    real gw = 0;
    symv_global(M, nv, ld, qacc, Ma);
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) gw += (real)0.5 * (Ma[i] - qfs[i]) * (qacc[i] - qas[i]);
    cw = b2_wsum(cw + gw);
    B2_WSYNC();
    if (cw > cs) {
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) qacc[i] = qas[i];
      B2_WSYNC();
      rows_dot(jpool, rowinfo, runinfo, rundofs, nefc, qacc, jar, aref);
      symv_global(M, nv, ld, qacc, Ma);
    }
    real cost = 0, gauss = 0;
    // rounding floor of the cost difference in this arithmetic: below it "improvement" is noise
    const real eps_real = sizeof(real) == 4 ? (real)1.1920929e-7 : (real)2.220446e-16;
    B2_PH(1)
    B2_UNROLL1
    for (int it = 0;; ++it) {
      // -- constraint update, part 1: row weights, cost, qfrc_constraint, gradient  // @region newton_update
      // The Hessian M + J^T diag(D*active) J only depends on WHICH rows are active: when the active set equals the one
      // of the last factorisation, the factor still in H is reused (same result, no copy / J^T D J / Cholesky).
      real c1 = 0;
      bool same_set = it > 0;
#ifdef B2_HOSTEMU
      int nflip_stat = 0;
#endif
      B2_UNROLL1
      for (int base = 0; base < nefc; base += B2_NLANE) {
        int r = base + B2_LANE;
        bool act = false;
        if (r < nefc) {
          real x = jar[r];
          if (EL && D[r] < 0) {   // row of an elliptic contact: written by cone_update below
            wgt[r] = 0; frc[r] = 0;
          } else
          if (FR && r < nf) {   // friction loss: "active" = in the quadratic zone (the only state that enters the Hessian)
            real f_, w_;
            c1 += fric_row(x, D[r], floss[r], f_, w_);
            wgt[r] = w_; frc[r] = f_;
            act = w_ != 0;
          } else {
          act = x < 0;
          real wr = act ? D[r] : (real)0;
          wgt[r] = wr;
          frc[r] = -wr * x;
          c1 += (real)0.5 * wr * x * x;
          }
        }
        unsigned bal = b2_ballot(act);
#ifdef B2_HOSTEMU
        unsigned* slot = actmask + (base >> 5);
        unsigned prev = (*slot >> (base & 31)) & 1u;
        if (it > 0 && prev != bal) ++nflip_stat;
        if (it == 0 || prev != bal) same_set = false;
        *slot = (*slot & ~(1u << (base & 31))) | (bal << (base & 31));
#else
        if (it == 0 || actmask[base >> 5] != bal) same_set = false;
        B2_WSYNC();
        if (B2_LANE == 0) actmask[base >> 5] = bal;
#endif
      }
      if constexpr (EL) {
        // cone contacts: sticking ones are quadratic rows like any other (Hessian fixed by the zone); a sliding contact's Hessian
        // moves with the residual, so the factor is only reused when no contact slides and no zone changed
        B2_WSYNC();
        cone_changed = false;
        c1 += cone_update(ca, jar, D, wgt, frc, true, cone_changed);
        if (b2_ballot(cone_changed) != 0) same_set = false;
      }
      real g1 = 0;
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) { g1 += (real)0.5 * (Ma[i] - qfs[i]) * (qacc[i] - qas[i]); qfc[i] = 0; }
      gauss = b2_wsum(g1);
      real newcost = b2_wsum(c1) + gauss;
      B2_WSYNC();
      if (FR && nf > 0) {  // friction-loss rows: at most one per dof, so lanes over rows scatter without a race (J = 1)
        B2_UNROLL1
        for (int r = B2_LANE; r < nf; r += B2_NLANE) qfc[rundofs[r]] = frc[r];
        B2_WSYNC();
      }
      if (nl > nf) {  // limit rows, gathered per dof (deterministic even if two rows share a dof)
        B2_UNROLL1
        for (int e = B2_LANE; e < nv; e += B2_NLANE) {
          real ff = 0;
          B2_UNROLL1
          for (int r = nf; r < nl; ++r)
            if (rundofs[r] == e) ff += jpool[r] * frc[r];
          qfc[e] += ff;
        }
        B2_WSYNC();
      }
      B2_UNROLL1
      for (int run = nl; run < nrun; ++run) {  // contact runs: qfc += J^T f over the run's dofs
        int ri0 = runinfo[2 * run], ri1 = runinfo[2 * run + 1];
        int dofoff = ri0 & 0xffff, wd = ri0 >> 16, row0 = ri1 & 0xffff, nrow = ri1 >> 16;
        const int wpad = wd | 1;
        const int off0 = rowinfo[row0] & 0xffff;
        const unsigned char* dl = rundofs + dofoff;
        B2_UNROLL1
        for (int a = B2_LANE; a < wd; a += B2_NLANE) {
          real ff = 0;
          B2_UNROLL_HOT
          for (int q = 0; q < nrow; ++q) ff = fma(jpool[off0 + q * wpad + a], frc[row0 + q], ff);
          qfc[dl[a]] += ff;
        }
        B2_WSYNC();
      }
      real gn = 0, gref = 0;
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) {
        real g = Ma[i] - qfs[i] - qfc[i];
        grad[i] = g; tmp[i] = g; gn += g * g;
        if (sizeof(real) == 4 && B2_GRAD_FLOOR_MULT > 0) gref += Ma[i] * Ma[i] + qfs[i] * qfs[i] + qfc[i] * qfc[i];
      }
      B2_WSYNC();
      if (it > 0) {
        gn = sqrt(b2_wsum(gn));
        // float32: a gradient below the rounding noise of its own three terms is zero for all practical purposes
        bool grad_is_noise = false, grad_is_large = false;
        if (sizeof(real) == 4 && B2_GRAD_FLOOR_MULT > 0) {
          const real gnoise = eps_real * sqrt(b2_wsum(gref));
          grad_is_noise = gn < (real)B2_GRAD_FLOOR_MULT * gnoise;
          // float32 cannot resolve cost differences below ~eps |cost|, so "no improvement" is not evidence of convergence while
          // the gradient still stands far above ITS rounding noise: keep iterating then (the Newton direction and the line
          // search work from derivatives, not from cost differences).  Without this a world could stop one iteration early with
          // a residual that the implicit integrator turns into a 1e-2 rad/s velocity error on a light ankle dof
          // (profiles/r02_divergence.txt).  0 = off.
          grad_is_large = B2_GRAD_KEEP_MULT > 0 && gn > (real)B2_GRAD_KEEP_MULT * gnoise;
        }
        real improvement = scale * (cost - newcost);
        real floor_ = fmax(tolerance, B2_FLOOR_MULT * eps_real * scale * fabs(newcost));
#ifdef B2_TRACE
        printf("w%d it%d cost %.9g new %.9g impr %.3g floor %.3g gn*scale %.3g same %d nefc %d\n", w, it, (double)cost, (double)newcost, (double)improvement, (double)floor_, (double)(scale * gn), (int)same_set, nefc);
#endif
        cost = newcost;
        if ((improvement < floor_ && !grad_is_large) || scale * gn < tolerance || grad_is_noise) break;
      } else {
        cost = newcost;
      }
      if (it >= m.iterations) break;
#ifdef B2_HOSTEMU
      b2h_flip_hist[it == 0 ? 65 : (nflip_stat < 64 ? nflip_stat : 64)] += 1;
#endif
      B2_PH(2)
      // -- part 2: H = M + J^T diag(w) J, only when another Newton step will be taken  // @region hessian
      if (cg) {
        // CG: precondition with M alone - factorised once per solve - instead of the Hessian of the active set
        if (it == 0) {
          copy_g2s(H, M, nv * ld);
          spd_factor_solve(H, nv, ld, tmp, invd);
        } else {
          chol_forwardsolve(H, nv, ld, invd, tmp);
          chol_backsolve(H, nv, ld, invd, tmp);
        }
      } else if (same_set) {
        chol_forwardsolve(H, nv, ld, invd, tmp);
        chol_backsolve(H, nv, ld, invd, tmp);
      } else {
      copy_g2s(H, M, nv * ld);
      if (FR && nf > 0) {  // friction-loss rows in the quadratic zone: D on their dof's diagonal entry (one row per dof at most)
        B2_UNROLL1
        for (int r = B2_LANE; r < nf; r += B2_NLANE) { const int e = rundofs[r]; H[e * ld + e] += wgt[r]; }
        B2_WSYNC();
      }
      if (nl > nf) {
        B2_UNROLL1
        for (int e = B2_LANE; e < nv; e += B2_NLANE) {
          real hh = 0;
          B2_UNROLL1
          for (int r = nf; r < nl; ++r)
            if (rundofs[r] == e) hh += wgt[r];
          H[e * ld + e] += hh;
        }
        B2_WSYNC();
      }
      B2_UNROLL1
      for (int run = nl; run < nrun; ++run) {  // J^T diag(w) J over the run's dof pairs
        int ri0 = runinfo[2 * run], ri1 = runinfo[2 * run + 1];
        int dofoff = ri0 & 0xffff, wd = ri0 >> 16, row0 = ri1 & 0xffff, nrow = ri1 >> 16;
        const int wpad = wd | 1;
        const int off0 = rowinfo[row0] & 0xffff;
        const unsigned char* dl = rundofs + dofoff;
        const int npair = wd * (wd + 1) / 2;
        B2_UNROLL1
        for (int pr = B2_LANE; pr < npair; pr += B2_NLANE) {
          const int ab = pairtab[pr];
          const int a = ab >> 8, b = ab & 0xff;
          const real* ja = jpool + off0 + a;
          const real* jb = jpool + off0 + b;
          const real* wq = wgt + row0;
          real acc = 0, acc2 = 0;
          int q = 0;
          B2_UNROLL1
          for (; q + 2 <= nrow; q += 2) {
            acc = fma(wq[q] * ja[0], jb[0], acc);
            acc2 = fma(wq[q + 1] * ja[wpad], jb[wpad], acc2);
            ja += 2 * wpad; jb += 2 * wpad;
          }
          if (q < nrow) acc = fma(wq[q] * ja[0], jb[0], acc);
          int da = dl[a], db = dl[b];  // dof lists are ascending: da >= db
          H[da * ld + db] += acc + acc2;
        }
        B2_WSYNC();
      }
      if constexpr (EL) {
        // cone Hessians (mj HessianCone): a sliding contact adds J_c^T Hc J_c with Hc = Dm a a^T + Dm kd b b^T (cone_update left
        // a, b and the weights in cpar), i.e. two weighted combinations of the contact's three rows; same dof pairs as its run
        B2_UNROLL1
        for (int c = 0; c < ncon; ++c) {
          const int adr = cinfo[4 * c];
          if (adr < 0 || (cinfo[4 * c + 1] >> 8) != 2) continue;
          const real* cp = cpar + B2_CPAR * c;
          const real a0 = cp[4], a1 = cp[5], a2 = cp[6], b1 = cp[7], b2 = cp[8], wa = cp[1], wb = cp[9];
          const int info = rowinfo[adr];
          const int off0 = info & 0xffff, run = info >> 16;
          const int ri0 = runinfo[2 * run];
          const int dofoff = ri0 & 0xffff, wd = ri0 >> 16;
          const int wpad = wd | 1;
          const unsigned char* dl = rundofs + dofoff;
          const int npair = wd * (wd + 1) / 2;
          B2_UNROLL1
          for (int pr = B2_LANE; pr < npair; pr += B2_NLANE) {
            const int ab = pairtab[pr];
            const real* ja = jpool + off0 + (ab >> 8);
            const real* jb = jpool + off0 + (ab & 0xff);
            const real pa = a0 * ja[0] + a1 * ja[wpad] + a2 * ja[2 * wpad], pb = a0 * jb[0] + a1 * jb[wpad] + a2 * jb[2 * wpad];
            const real ta = b1 * ja[wpad] + b2 * ja[2 * wpad], tb = b1 * jb[wpad] + b2 * jb[2 * wpad];
            H[dl[ab >> 8] * ld + dl[ab & 0xff]] += wa * pa * pb + wb * ta * tb;
          }
          B2_WSYNC();
        }
      }
      B2_PH(3)
      spd_factor_solve(H, nv, ld, tmp, invd);  // @region factor_call
      }
      B2_PH(4)
      if (cg && it > 0) {
        // Polak-Ribiere: beta = grad . (Mgrad - Mgrad_old) / (grad_old . Mgrad_old), restarted (beta = 0) when negative
        real num = 0, den = 0;
        B2_UNROLL1
        for (int i = B2_LANE; i < nv; i += B2_NLANE) { num += grad[i] * (tmp[i] - mgold[i]); den += gold[i] * mgold[i]; }
        num = b2_wsum(num); den = b2_wsum(den);
        real beta = num / (den > B2_MINVAL ? den : B2_MINVAL);
        if (beta < 0) beta = 0;
        B2_UNROLL1
        for (int i = B2_LANE; i < nv; i += B2_NLANE) search[i] = -tmp[i] + beta * search[i];
      } else {
        B2_UNROLL1
        for (int i = B2_LANE; i < nv; i += B2_NLANE) search[i] = -tmp[i];
      }
      if (cg) {
        B2_UNROLL1
        for (int i = B2_LANE; i < nv; i += B2_NLANE) { gold[i] = grad[i]; mgold[i] = tmp[i]; }
      }
      B2_WSYNC();
      // -- line search  // @region ls_setup
      real sn = 0, q1 = 0, q2 = 0, gs = 0;
      symv_global(M, nv, ld, search, Mv);
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) {
        sn += search[i] * search[i];
        q1 += search[i] * (Ma[i] - qfs[i]);
        q2 += (real)0.5 * search[i] * Mv[i];
        gs += search[i] * grad[i];
      }
      sn = sqrt(b2_wsum(sn));
      b2_wsum3(q1, q2, gs);
      rows_dot(jpool, rowinfo, runinfo, rundofs, nefc, search, jv, nullptr);
      B2_PH(5)
      real alpha = linesearch<FR, EL>(gauss, q1, q2, sn, tolerance, ls_tolerance, m.ls_iterations, scale, jar, jv, D, nefc, cost, cg ? (real)0 : gs, nf, floss, ca);   // the alpha = 0 shortcut needs H search = -grad: Newton only
      B2_PH(6)
      if (alpha == 0) break;
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) { qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
      B2_UNROLL1
      for (int r = B2_LANE; r < nefc; r += B2_NLANE) jar[r] += alpha * jv[r];
      B2_WSYNC();
      niter = it + 1;
    }
    // qfc / efc_force at the final point: the loop exits right after a constraint update
    // (break at the convergence test or iteration cap) or with alpha == 0 (state unchanged).
    B2_UNROLL1
    for (int r = B2_LANE; r < nefc; r += B2_NLANE) {
      real x = jar[r];
      if (EL && D[r] < 0) { g_force[r] = frc[r]; continue; }   // cone rows: the last update's force (the loop ends right after an update)
      if (FR && r < nf) { real f_, w_; fric_row(x, D[r], floss[r], f_, w_); g_force[r] = f_; }
      else g_force[r] = x < 0 ? -D[r] * x : 0;
    }
  }
  B2_UNROLL1
  for (int i = B2_LANE; i < nv; i += B2_NLANE) { g_qacc[i] = qacc[i]; g_qfc[i] = qfc[i]; }
  if (B2_LANE == 0) d.di[B2_DI_SOLVER_NITER][w] = niter;
  B2_WSYNC();
  B2_PH(7)
  B2_PH_END()
}

// run-time dispatch for callers that are not templated themselves (the host builds: tests/hostemu, oracle/cport)
B2_DEV void solve_world(const KModel& m, const KData& d, const SolveSmem& L, int w, real* sm, real* arena, int* actl) {
  if (m.cone == 1) {   // elliptic cones
    if (m.nfriction > 0) {
      if (m.solver == 1) solve_world_t<true, true, true>(m, d, L, w, sm, arena, actl);
      else solve_world_t<false, true, true>(m, d, L, w, sm, arena, actl);
    } else if (m.solver == 1) solve_world_t<true, false, true>(m, d, L, w, sm, arena, actl);
    else solve_world_t<false, false, true>(m, d, L, w, sm, arena, actl);
  } else if (m.nfriction > 0) {
    if (m.solver == 1) solve_world_t<true, true>(m, d, L, w, sm, arena, actl);
    else solve_world_t<false, true>(m, d, L, w, sm, arena, actl);
  } else if (m.solver == 1) solve_world_t<true, false>(m, d, L, w, sm, arena, actl);
  else solve_world_t<false, false>(m, d, L, w, sm, arena, actl);
}

// Shared-memory layout of the post stage (sensors + integration): the integrator's matrix and a few vectors.
struct PostSmem {
  int H, qvel, qacc, qfs, qfc, tmp, invd, total;
};
static inline PostSmem post_smem_layout(int nv) {
  PostSmem s;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  s.H = take(nv * b2_ld(nv));
  s.qvel = take(nv); s.qacc = take(nv); s.qfs = take(nv); s.qfc = take(nv); s.tmp = take(nv); s.invd = take(64);
  s.total = o;
  return s;
}

// ---------------------------------------------------------------------------
// Stage 3 for world w: sensors, then integration.  A kernel of its own so that the Newton loop of stage 2 shares the
// instruction cache with nothing else (profiles/: k_solve was instruction-fetch bound while this code lived in it).
// Inputs come from the global outputs of stages 1-2 (qacc, qfrc_*, efc_force, contact efc addresses).
// ---------------------------------------------------------------------------
B2_DEV void post_world(const KModel& m, const KData& d, const PostSmem& L, int w, real* sm, int do_integrate) {
  const int nv = m.nv, nb = m.nbody, nq = m.nq, nu = m.nu, njmax = d.njmax, nconmax = d.nconmax;
  real* H = sm + L.H; real* qvel = sm + L.qvel; real* qacc = sm + L.qacc; real* qfs = sm + L.qfs; real* qfc = sm + L.qfc;
  real* tmp = sm + L.tmp; real* invd = sm + L.invd;
  const int ld = b2_ld(nv);
  const int* body_parent = MI(B2_MI_BODY_PARENTID);
  const int* body_root = MI(B2_MI_BODY_ROOTID);
  const int* geom_body = MI(B2_MI_GEOM_BODYID);
  const int* jnt_type = MI(B2_MI_JNT_TYPE);
  const int* jnt_qposadr = MI(B2_MI_JNT_QPOSADR);
  const int* jnt_dofadr = MI(B2_MI_JNT_DOFADR);
  const real timestep = MF(B2_MF_OPT_TIMESTEP)[0];
  real* g_qpos = d.df[B2_DF_QPOS] + (size_t)w * nq;
  real* g_qvel = d.df[B2_DF_QVEL] + (size_t)w * nv;
  const real* g_scom = d.df[B2_DF_SUBTREE_COM] + (size_t)w * nb * 3;
  const real* g_cdof = d.df[B2_DF_CDOF] + (size_t)w * nv * 6;
  const real* M = d.df[B2_DF_QM] + (size_t)w * nv * b2_ld(nv);
  const int ncon = d.di[B2_DI_NCON][w];
  const int nefc = d.di[B2_DI_NEFC][w];
  const real* c_dist = d.df[B2_DF_CONTACT_DIST] + (size_t)w * nconmax;
  const real* c_pos = d.df[B2_DF_CONTACT_POS] + (size_t)w * nconmax * 3;
  const real* c_frame = d.df[B2_DF_CONTACT_FRAME] + (size_t)w * nconmax * 9;
  const real* c_fri = d.df[B2_DF_CONTACT_FRICTION] + (size_t)w * nconmax * 5;
  const int* c_geom = d.di[B2_DI_CONTACT_GEOM] + (size_t)w * nconmax * 2;
  const int* c_dim = d.di[B2_DI_CONTACT_DIM] + (size_t)w * nconmax;
  const int* c_adr = d.di[B2_DI_CONTACT_EFC_ADDRESS] + (size_t)w * nconmax;
  const real* efc_force = d.df[B2_DF_EFC_FORCE] + (size_t)w * njmax;
  (void)nu; (void)c_dist;
  B2_UNROLL1
  for (int i = B2_LANE; i < nv; i += B2_NLANE) {
    size_t o = (size_t)w * nv + i;
    qvel[i] = g_qvel[i];
    qacc[i] = d.df[B2_DF_QACC][o];
    qfs[i] = d.df[B2_DF_QFRC_SMOOTH][o];
    qfc[i] = d.df[B2_DF_QFRC_CONSTRAINT][o];
  }
  B2_WSYNC();

  // ---- sensors -------------------------------------------------------------------------  // @region sensors
  if (m.nsensor > 0) {
    const int* ancmask = m.nbody <= 64 ? MI(B2_MI_BODY_ANCMASK) : nullptr;
    const int* s_type = MI(B2_MI_SENSOR_TYPE);
    const int* s_objtype = MI(B2_MI_SENSOR_OBJTYPE);
    const int* s_objid = MI(B2_MI_SENSOR_OBJID);
    const int* s_reftype = MI(B2_MI_SENSOR_REFTYPE);
    const int* s_refid = MI(B2_MI_SENSOR_REFID);
    const int* s_intprm = MI(B2_MI_SENSOR_INTPRM);
    const int* s_dim = MI(B2_MI_SENSOR_DIM);
    const int* s_adr = MI(B2_MI_SENSOR_ADR);
    const real* s_cutoff = MF(B2_MF_SENSOR_CUTOFF);
    const int* site_body = MI(B2_MI_SITE_BODYID);
    const int* level_adr = MI(B2_MI_LEVEL_ADR);
    const int* level_body = MI(B2_MI_LEVEL_BODY);
    const int* child_adr = MI(B2_MI_CHILD_ADR);
    const int* child_body = MI(B2_MI_CHILD_BODY);
    const int* body_dofnum = MI(B2_MI_BODY_DOFNUM);
    const int* body_dofadr = MI(B2_MI_BODY_DOFADR);
    const real* g_cvel = d.df[B2_DF_CVEL] + (size_t)w * nb * 6;
    const real* g_cdofdot = d.df[B2_DF_CDOF_DOT] + (size_t)w * nv * 6;
    const real* g_xipos = d.df[B2_DF_XIPOS] + (size_t)w * nb * 3;
    const real* g_ximat = d.df[B2_DF_XIMAT] + (size_t)w * nb * 9;
    const real* g_sxpos = d.df[B2_DF_SITE_XPOS] + (size_t)w * m.nsite * 3;
    const real* g_sxmat = d.df[B2_DF_SITE_XMAT] + (size_t)w * m.nsite * 9;
    real* g_cacc = d.df[B2_DF_CACC] + (size_t)w * nb * 6;
    real* g_slin = d.df[B2_DF_SUBTREE_LINVEL] + (size_t)w * nb * 3;
    real* g_sang = d.df[B2_DF_SUBTREE_ANGMOM] + (size_t)w * nb * 3;
    real* sd = d.df[B2_DF_SENSORDATA] + (size_t)w * m.nsensordata;
    const real* body_mass = MF(B2_MF_BODY_MASS);
    const real* body_inertia = MF(B2_MF_BODY_INERTIA);

    // The level sweeps below pass their results from level to level; through global memory every level costs a store -> load
    // round trip to L2, so they run on shared-memory copies (the integrator's matrix is not built yet: its space is free) and
    // the finished arrays go to HBM in one coalesced pass.
    real* s_cacc = H;                       // [6 nb]
    real* s_slin = H;                       // [3 nb]   (the cacc sweep is over by then)
    real* s_sang = H + 3 * nb;              // [3 nb]
    real* s_scom = H + 6 * nb;              // [3 nb]
    const bool sweeps_in_smem = 9 * nb <= nv * ld;
    if (!sweeps_in_smem) { s_cacc = g_cacc; s_slin = g_slin; s_sang = g_sang; s_scom = const_cast<real*>(g_scom); }
    if (m.has_accelerometer) {
      // cacc including qacc (rne_postconstraint, acceleration part), forward over levels
      const real* grav = MF(B2_MF_OPT_GRAVITY);
      B2_UNROLL1
      for (int lev = 0; lev < m.nlevel; ++lev) {
        B2_UNROLL1
        for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
          int b = level_body[k];
          real a[6];
          if (b == 0) {
            a[0] = a[1] = a[2] = 0; a[3] = -grav[0]; a[4] = -grav[1]; a[5] = -grav[2];
          } else {
            int p = body_parent[b];
            for (int c = 0; c < 6; ++c) a[c] = s_cacc[6 * p + c];
            B2_UNROLL1
            for (int e = 0; e < body_dofnum[b]; ++e) {
              int dd = body_dofadr[b] + e;
              for (int c = 0; c < 6; ++c) a[c] += g_cdofdot[6 * dd + c] * qvel[dd] + g_cdof[6 * dd + c] * qacc[dd];
            }
          }
          for (int c = 0; c < 6; ++c) s_cacc[6 * b + c] = a[c];
        }
        B2_WSYNC();
      }
      if (s_cacc != g_cacc) {
        B2_UNROLL1
        for (int i = B2_LANE; i < 6 * nb; i += B2_NLANE) g_cacc[i] = s_cacc[i];
      }
      B2_WSYNC();
    }
    if (m.has_subtreevel) {
      // subtree linear velocity: backward gather of body momenta
      // per-body com velocity and own angular momentum stored first
      B2_UNROLL1
      for (int b = B2_LANE; b < nb; b += B2_NLANE) {
        int r = body_root[b];
        const real* v = g_cvel + 6 * b;
        real off[3] = {g_xipos[3 * b] - g_scom[3 * r], g_xipos[3 * b + 1] - g_scom[3 * r + 1], g_xipos[3 * b + 2] - g_scom[3 * r + 2]};
        real t[3];
        cross3(t, v, off);
        real ms = body_mass[b];
        s_slin[3 * b] = ms * (v[3] + t[0]); s_slin[3 * b + 1] = ms * (v[4] + t[1]); s_slin[3 * b + 2] = ms * (v[5] + t[2]);
        if (sweeps_in_smem) { s_scom[3 * b] = g_scom[3 * b]; s_scom[3 * b + 1] = g_scom[3 * b + 1]; s_scom[3 * b + 2] = g_scom[3 * b + 2]; }
        const real* Ri = g_ximat + 9 * b;
        real wl[3], Iw[3];
        mulmTv3(wl, Ri, v);
        wl[0] *= body_inertia[3 * b]; wl[1] *= body_inertia[3 * b + 1]; wl[2] *= body_inertia[3 * b + 2];
        mulmv3(Iw, Ri, wl);
        s_sang[3 * b] = Iw[0]; s_sang[3 * b + 1] = Iw[1]; s_sang[3 * b + 2] = Iw[2];
      }
      B2_WSYNC();
      // angular momentum about each subtree's com: gather children bottom-up.  Keep momentum P, mass m and
      // L about own subtree com; combining child c into parent p:  L_p += L_c + (com_c - com_p) x P_c  is exact
      // when L is taken about com with absolute velocities: L_about_com = sum r_i x m_i v_i (since sum m_i r_i = 0).
      // Stage A: L_b (about xipos_b, own body) -> shift to subtree com later, so store body term about its own com.
      // Stage B: bottom-up: for parent p: P_p = P_own + sum P_c ; L_p(about com_p) = L_own + (xipos_p - com_p) x P_own
      //          + sum_c [ L_c + (com_c - com_p) x P_c ].
      B2_UNROLL1
      for (int lev = m.nlevel - 1; lev >= 0; --lev) {
        B2_UNROLL1
        for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
          int b = level_body[k];
          real P[3] = {s_slin[3 * b], s_slin[3 * b + 1], s_slin[3 * b + 2]};
          real Lb[3] = {s_sang[3 * b], s_sang[3 * b + 1], s_sang[3 * b + 2]};
          real dx[3] = {g_xipos[3 * b] - s_scom[3 * b], g_xipos[3 * b + 1] - s_scom[3 * b + 1], g_xipos[3 * b + 2] - s_scom[3 * b + 2]};
          real t[3];
          cross3(t, dx, P);
          Lb[0] += t[0]; Lb[1] += t[1]; Lb[2] += t[2];
          B2_UNROLL1
          for (int c = child_adr[b]; c < child_adr[b + 1]; ++c) {
            int ch = child_body[c];
            real Pc[3] = {s_slin[3 * ch], s_slin[3 * ch + 1], s_slin[3 * ch + 2]};
            real dc[3] = {s_scom[3 * ch] - s_scom[3 * b], s_scom[3 * ch + 1] - s_scom[3 * b + 1], s_scom[3 * ch + 2] - s_scom[3 * b + 2]};
            cross3(t, dc, Pc);
            Lb[0] += s_sang[3 * ch] + t[0]; Lb[1] += s_sang[3 * ch + 1] + t[1]; Lb[2] += s_sang[3 * ch + 2] + t[2];
            P[0] += Pc[0]; P[1] += Pc[1]; P[2] += Pc[2];
          }
          s_slin[3 * b] = P[0]; s_slin[3 * b + 1] = P[1]; s_slin[3 * b + 2] = P[2];
          s_sang[3 * b] = Lb[0]; s_sang[3 * b + 1] = Lb[1]; s_sang[3 * b + 2] = Lb[2];
        }
        B2_WSYNC();
      }
      // momentum -> velocity, and out to HBM
      const real* stm = MF(B2_MF_BODY_SUBTREEMASS);
      B2_UNROLL1
      for (int b = B2_LANE; b < nb; b += B2_NLANE) {
        real inv = stm[b] > B2_MINVAL ? (real)1 / stm[b] : (real)0;
        s_slin[3 * b] *= inv; s_slin[3 * b + 1] *= inv; s_slin[3 * b + 2] *= inv;
      }
      B2_WSYNC();
      if (s_slin != g_slin) {
        B2_UNROLL1
        for (int i = B2_LANE; i < 3 * nb; i += B2_NLANE) { g_slin[i] = s_slin[i]; g_sang[i] = s_sang[i]; }
      }
      B2_WSYNC();
    }
    B2_WSYNC();
    // Per-contact quantities every contact sensor needs - the two bodies, the contact-frame force, the world-frame force -
    // once per contact (lanes over contacts) instead of once per (sensor, contact) inside the sensor loops below: those run one
    // lane per sensor, so their length is the latency of this stage.  Scratch: the integrator's matrix, not built yet.
    const int cpre = ancmask ? (ncon < (nv * ld) / 12 ? ncon : (nv * ld) / 12) : 0;
    int* cbody = (int*)H;                  // [2 cpre] bodies of the two geoms
    unsigned* canc = (unsigned*)H + 2 * cpre;   // [4 cpre] their ancestor masks
    real* cfc = H + 6 * cpre;
    real* cfg = cfc + 3 * cpre;
    B2_UNROLL1
    for (int c = B2_LANE; c < cpre; c += B2_NLANE) {
      const int b1 = geom_body[c_geom[2 * c]], b2 = geom_body[c_geom[2 * c + 1]];
      cbody[2 * c] = b1; cbody[2 * c + 1] = b2;
      canc[4 * c] = (unsigned)ancmask[2 * b1]; canc[4 * c + 1] = (unsigned)ancmask[2 * b1 + 1];
      canc[4 * c + 2] = (unsigned)ancmask[2 * b2]; canc[4 * c + 3] = (unsigned)ancmask[2 * b2 + 1];
      real fc[3], fg[3];
      contact_force(fc, nefc > 0 ? c_adr[c] : -1, c_dim[c], c_fri + 5 * c, efc_force, m.cone);
      mulmTv3(fg, c_frame + 9 * c, fc);
      for (int e = 0; e < 3; ++e) { cfc[3 * c + e] = fc[e]; cfg[3 * c + e] = fg[e]; }
    }
    B2_WSYNC();
    B2_UNROLL1
    for (int s = B2_LANE; s < m.nsensor; s += B2_NLANE) {
      int t = s_type[s], oid = s_objid[s], adr = s_adr[s], dim = s_dim[s];
      real* out = sd + adr;
      B2_UNROLL1
      for (int e = 0; e < dim; ++e) out[e] = 0;
      if (t == B2_SENS_GYRO || t == B2_SENS_VELOCIMETER || t == B2_SENS_ACCELEROMETER) {
        int b = site_body[oid], r = body_root[b];
        const real* R = g_sxmat + 9 * oid;
        const real* v = g_cvel + 6 * b;
        real off[3] = {g_sxpos[3 * oid] - g_scom[3 * r], g_sxpos[3 * oid + 1] - g_scom[3 * r + 1], g_sxpos[3 * oid + 2] - g_scom[3 * r + 2]};
        real tt[3], lin[3], angl[3], linl[3];
        cross3(tt, v, off);
        lin[0] = v[3] + tt[0]; lin[1] = v[4] + tt[1]; lin[2] = v[5] + tt[2];
        mulmTv3(angl, R, v);
        mulmTv3(linl, R, lin);
        if (t == B2_SENS_GYRO) { out[0] = angl[0]; out[1] = angl[1]; out[2] = angl[2]; }
        else if (t == B2_SENS_VELOCIMETER) { out[0] = linl[0]; out[1] = linl[1]; out[2] = linl[2]; }
        else {
          const real* a = g_cacc + 6 * b;
          real al[3], acc[3], cor[3];
          cross3(tt, a, off);
          al[0] = a[3] + tt[0]; al[1] = a[4] + tt[1]; al[2] = a[5] + tt[2];
          mulmTv3(acc, R, al);
          cross3(cor, angl, linl);
          out[0] = acc[0] + cor[0]; out[1] = acc[1] + cor[1]; out[2] = acc[2] + cor[2];
        }
      } else if (t == B2_SENS_JOINTPOS) {
        out[0] = g_qpos[jnt_qposadr[oid]];
      } else if (t == B2_SENS_JOINTVEL) {
        out[0] = qvel[jnt_dofadr[oid]];
      } else if (t == B2_SENS_SUBTREECOM) {
        out[0] = g_scom[3 * oid]; out[1] = g_scom[3 * oid + 1]; out[2] = g_scom[3 * oid + 2];
      } else if (t == B2_SENS_SUBTREELINVEL) {
        out[0] = g_slin[3 * oid]; out[1] = g_slin[3 * oid + 1]; out[2] = g_slin[3 * oid + 2];
      } else if (t == B2_SENS_SUBTREEANGMOM) {
        out[0] = g_sang[3 * oid]; out[1] = g_sang[3 * oid + 1]; out[2] = g_sang[3 * oid + 2];
      } else if (t == B2_SENS_FRAMEPOS || t == B2_SENS_FRAMEQUAT || t == B2_SENS_FRAMELINVEL || t == B2_SENS_FRAMEANGVEL) {
        int ot = s_objtype[s];
        const real *pos, *R;
        int b;
        if (ot == B2_OBJ_SITE) { pos = g_sxpos + 3 * oid; R = g_sxmat + 9 * oid; b = site_body[oid]; }
        else if (ot == B2_OBJ_GEOM) { pos = d.df[B2_DF_GEOM_XPOS] + ((size_t)w * m.ngeom + oid) * 3; R = d.df[B2_DF_GEOM_XMAT] + ((size_t)w * m.ngeom + oid) * 9; b = geom_body[oid]; }
        else if (ot == B2_OBJ_XBODY) { pos = d.df[B2_DF_XPOS] + ((size_t)w * nb + oid) * 3; R = d.df[B2_DF_XMAT] + ((size_t)w * nb + oid) * 9; b = oid; }
        else { pos = g_xipos + 3 * oid; R = g_ximat + 9 * oid; b = oid; }
        if (t == B2_SENS_FRAMEPOS) { out[0] = pos[0]; out[1] = pos[1]; out[2] = pos[2]; }
        else if (t == B2_SENS_FRAMEQUAT) { mat2quat(out, R); }
        else {
          int r = body_root[b];
          const real* v = g_cvel + 6 * b;
          real off[3] = {pos[0] - g_scom[3 * r], pos[1] - g_scom[3 * r + 1], pos[2] - g_scom[3 * r + 2]};
          real tt[3];
          cross3(tt, v, off);
          if (t == B2_SENS_FRAMELINVEL) { out[0] = v[3] + tt[0]; out[1] = v[4] + tt[1]; out[2] = v[5] + tt[2]; }
          else { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
        }
      } else if (t == B2_SENS_CONTACT) {
        int bits = s_intprm[3 * s], reduce = s_intprm[3 * s + 1], num = s_intprm[3 * s + 2];
        int ot = s_objtype[s], rt = s_reftype[s], rid = s_refid[s];
        int per_slot = dim / (num > 0 ? num : 1);
        int nfound = 0;
        real net[3] = {0, 0, 0};
        // pass 1: count matches (and net force)
        B2_UNROLL1
        for (int c = 0; c < ncon; ++c) {
          int g1 = c_geom[2 * c], g2 = c_geom[2 * c + 1];
          real sgn = 0;
          if (c < cpre) {
            const int b1 = cbody[2 * c], b2 = cbody[2 * c + 1];
            const unsigned l1 = canc[4 * c], h1 = canc[4 * c + 1], l2 = canc[4 * c + 2], h2 = canc[4 * c + 3];
            if (match_side_m(ot, oid, g1, b1, l1, h1) && match_side_m(rt, rid, g2, b2, l2, h2)) sgn = 1;
            else if (match_side_m(ot, oid, g2, b2, l2, h2) && match_side_m(rt, rid, g1, b1, l1, h1)) sgn = -1;
          } else {
            const int b1 = geom_body[g1], b2 = geom_body[g2];
            if (match_side(body_parent, ancmask, ot, oid, g1, b1) && match_side(body_parent, ancmask, rt, rid, g2, b2)) sgn = 1;
            else if (match_side(body_parent, ancmask, ot, oid, g2, b2) && match_side(body_parent, ancmask, rt, rid, g1, b1)) sgn = -1;
          }
          if (sgn == 0) continue;
          ++nfound;
          if (reduce == 3 && (bits & 2)) {
            real fg[3];
            if (c < cpre) { fg[0] = cfg[3 * c]; fg[1] = cfg[3 * c + 1]; fg[2] = cfg[3 * c + 2]; }
            else {
              real fc[3];
              contact_force(fc, nefc > 0 ? c_adr[c] : -1, c_dim[c], c_fri + 5 * c, efc_force, m.cone);
              mulmTv3(fg, c_frame + 9 * c, fc);
            }
            net[0] += sgn * fg[0]; net[1] += sgn * fg[1]; net[2] += sgn * fg[2];
          }
        }
        if (nfound > 0) {
          if (reduce == 3) {
            int p = 0;
            if (bits & 1) out[p++] = (real)nfound;
            if (bits & 2) { out[p] = net[0]; out[p + 1] = net[1]; out[p + 2] = net[2]; p += 3; }
          } else {
            // slots: first `num` matches in contact order (reduce none) or ranked by dist / force
            int filled = 0;
            real lastkey = 0; int lastc = -1;
            B2_UNROLL1
            for (int slot = 0; slot < num && slot < nfound; ++slot) {
              int pick = -1; real pickkey = 0, picksgn = 0;
              B2_UNROLL1
              for (int c = 0; c < ncon; ++c) {
                int g1 = c_geom[2 * c], g2 = c_geom[2 * c + 1];
                real sgn = 0;
                if (c < cpre) {
                  const int b1 = cbody[2 * c], b2 = cbody[2 * c + 1];
                  const unsigned l1 = canc[4 * c], h1 = canc[4 * c + 1], l2 = canc[4 * c + 2], h2 = canc[4 * c + 3];
                  if (match_side_m(ot, oid, g1, b1, l1, h1) && match_side_m(rt, rid, g2, b2, l2, h2)) sgn = 1;
                  else if (match_side_m(ot, oid, g2, b2, l2, h2) && match_side_m(rt, rid, g1, b1, l1, h1)) sgn = -1;
                } else {
                  const int b1 = geom_body[g1], b2 = geom_body[g2];
                  if (match_side(body_parent, ancmask, ot, oid, g1, b1) && match_side(body_parent, ancmask, rt, rid, g2, b2)) sgn = 1;
                  else if (match_side(body_parent, ancmask, ot, oid, g2, b2) && match_side(body_parent, ancmask, rt, rid, g1, b1)) sgn = -1;
                }
                if (sgn == 0) continue;
                real key;
                if (reduce == 0) key = (real)c;
                else if (reduce == 1) key = c_dist[c];
                else {
                  real fc[3];
                  if (c < cpre) { fc[0] = cfc[3 * c]; fc[1] = cfc[3 * c + 1]; fc[2] = cfc[3 * c + 2]; }
                  else contact_force(fc, nefc > 0 ? c_adr[c] : -1, c_dim[c], c_fri + 5 * c, efc_force, m.cone);
                  key = -sqrt(dot3(fc, fc));
                }
                // strict ordering by (key, c) to enumerate matches in sorted order
                bool after_last = lastc < 0 || key > lastkey || (key == lastkey && c > lastc);
                if (!after_last) continue;
                if (pick < 0 || key < pickkey || (key == pickkey && c < pick)) { pick = c; pickkey = key; picksgn = sgn; }
              }
              if (pick < 0) break;
              lastkey = pickkey; lastc = pick;
              int p = slot * per_slot;
              int c = pick;
              if (bits & 1) out[p++] = (real)nfound;
              if (bits & 2) {
                real fc[3];
                if (c < cpre) { fc[0] = cfc[3 * c]; fc[1] = cfc[3 * c + 1]; fc[2] = cfc[3 * c + 2]; }
                else contact_force(fc, nefc > 0 ? c_adr[c] : -1, c_dim[c], c_fri + 5 * c, efc_force, m.cone);
                out[p] = picksgn * fc[0]; out[p + 1] = picksgn * fc[1]; out[p + 2] = picksgn * fc[2]; p += 3;
              }
              if (bits & 4) p += 3;
              if (bits & 8) out[p++] = c_dist[c];
              if (bits & 16) { out[p] = c_pos[3 * c]; out[p + 1] = c_pos[3 * c + 1]; out[p + 2] = c_pos[3 * c + 2]; p += 3; }
              if (bits & 32) { out[p] = picksgn * c_frame[9 * c]; out[p + 1] = picksgn * c_frame[9 * c + 1]; out[p + 2] = picksgn * c_frame[9 * c + 2]; p += 3; }
              if (bits & 64) { out[p] = picksgn * c_frame[9 * c + 3]; out[p + 1] = picksgn * c_frame[9 * c + 4]; out[p + 2] = picksgn * c_frame[9 * c + 5]; p += 3; }
              ++filled;
            }
            (void)filled;
          }
        }
      }
      real cut = s_cutoff[s];
      if (cut > 0 && t != B2_SENS_CONTACT)
        B2_UNROLL1
        for (int e = 0; e < dim; ++e) out[e] = clampr(out[e], -cut, cut);
    }
    B2_WSYNC();
  }

  // ---- integrate ---------------------------------------------------------------------------  // @region integrate
  if (do_integrate) {
    const real* damping = MF(B2_MF_DOF_DAMPING);
    real* g_ws = d.df[B2_DF_QACC_WARMSTART] + (size_t)w * nv;
    bool implicit = false;
    if (m.integrator == B2_INT_IMPLICITFAST) implicit = true;
    else {
      int anyd = 0;
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) anyd |= damping[i] > 0;
      implicit = b2_wor_i(anyd) != 0;
    }
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) tmp[i] = qacc[i];
    B2_WSYNC();
    if (implicit) {
      // (M - h*qDeriv) a = qfrc_smooth + qfrc_constraint, qDeriv diagonal: -damping (+ actuator velocity gain)
      copy_g2s(H, M, nv * ld);
      const int* trnid = MI(B2_MI_ACTUATOR_TRNID);
      const int* dof_actid = MI(B2_MI_DOF_ACTID);
      const int* biastype = MI(B2_MI_ACTUATOR_BIASTYPE);
      const int* flim = MI(B2_MI_ACTUATOR_FORCELIMITED);
      const real* biasprm = MF(B2_MF_ACTUATOR_BIASPRM);
      const real* gear = MF(B2_MF_ACTUATOR_GEAR);
      const real* frange = MF(B2_MF_ACTUATOR_FORCERANGE);
      const real* g_af = d.df[B2_DF_ACTUATOR_FORCE] + (size_t)w * nu;
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) {
        real qd = -damping[i];
        if (m.integrator == B2_INT_IMPLICITFAST) {
          int aid = dof_actid[i];
          int a_lo = aid >= 0 ? aid : 0, a_hi = aid >= 0 ? aid + 1 : (aid == -2 ? nu : 0);
          B2_UNROLL1
          for (int a = a_lo; a < a_hi; ++a) {
            if (biastype[a] != 1 || jnt_dofadr[trnid[2 * a]] != i) continue;
            real fa = g_af[a];
            if (flim[a] && (fa <= frange[2 * a] || fa >= frange[2 * a + 1])) continue;
            qd += gear[6 * a] * gear[6 * a] * biasprm[10 * a + 2];
          }
        }
        H[i * ld + i] -= timestep * qd;
        tmp[i] = qfs[i] + qfc[i];
      }
      B2_WSYNC();
      spd_factor_solve(H, nv, ld, tmp, invd);
    }
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) {
      real v = qvel[i] + timestep * tmp[i];
      qvel[i] = v;
      g_qvel[i] = v;
      g_ws[i] = qacc[i];
    }
    B2_WSYNC();
    B2_UNROLL1
    for (int j = B2_LANE; j < m.njnt; j += B2_NLANE) {
      int qa = jnt_qposadr[j], da = jnt_dofadr[j];
      if (jnt_type[j] == B2_JNT_FREE) {
        g_qpos[qa] += timestep * qvel[da]; g_qpos[qa + 1] += timestep * qvel[da + 1]; g_qpos[qa + 2] += timestep * qvel[da + 2];
        real q[4] = {g_qpos[qa + 3], g_qpos[qa + 4], g_qpos[qa + 5], g_qpos[qa + 6]};
        real wv[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
        real n = sqrt(dot3(wv, wv));
        if (n >= B2_MINVAL) {
          real ax[3] = {wv[0] / n, wv[1] / n, wv[2] / n}, dq[4], q2[4];
          axisangle2quat(dq, ax, timestep * n);
          quat_mul(q2, q, dq);
          q[0] = q2[0]; q[1] = q2[1]; q[2] = q2[2]; q[3] = q2[3];
        }
        quat_normalize(q);
        g_qpos[qa + 3] = q[0]; g_qpos[qa + 4] = q[1]; g_qpos[qa + 5] = q[2]; g_qpos[qa + 6] = q[3];
      } else {
        g_qpos[qa] += timestep * qvel[da];
      }
    }
    if (B2_LANE == 0) d.df[B2_DF_TIME][w] += timestep;
  }
}
