// b2_smooth.cuh - stage 1 of the physics step, one warp per world:
//   kinematics -> com_pos (subtree_com, cinert, cdof) -> CRB (dense qM)
//   -> collision (bounding-sphere broadphase + primitive narrowphase, ordered compaction)
//   -> com_vel (cvel, cdof_dot) -> passive -> actuation -> RNE bias -> qfrc_smooth -> qacc_smooth
//
// Restates the position/velocity stages of MuJoCo's mj_forward as executed by
// mujoco_warp's `_src.smooth`, `_src.collision_driver`, `_primitive_narrowphase`,
// `_src.passive`, `_src.forward` modules (call site
// /root/reference/src/mjlab/sim/sim.py:184-187).
//
// Roofline: HBM-bound in principle (per world: R = qpos,qvel,ctrl,xfrc rows;
// W = body/geom/site poses, cvel, cdof, qM, ...), in practice latency-bound on
// the tree recursions (depth 11 for G1), see DESIGN.md.
#pragma once
#include "b2_common.cuh"

// Per-world shared-memory layout for stage 1 (offsets in reals).
struct SmoothSmem {
  int qpos, qvel, xpos, xquat, xmat, xipos, xanchor, xaxis, scom, smass, cinert, crb, cdof, cdofdot, cvel, cacc, cfrc, M, f,
      aforce, cand, invd, total;
};

static inline SmoothSmem smooth_smem_layout(int nq, int nv, int nu, int nbody, int njnt, int ncand) {
  SmoothSmem s;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  // Shared memory per world sets how many worlds an SM holds, and this kernel is latency-bound: occupancy pays.  Only what
  // lives through the whole pass has its own slot; everything else time-shares ONE region whose largest tenant is the dense
  // mass matrix, which is why smooth_world assembles M last (after the bias forces), straight before its factorisation:
  //   position stage   : xpos | xquat | xanchor | xaxis | xmat      (last read: cdof)
  //   velocity / force : cand (collision) = cfrc (RNE) | cvel | cdofdot | cacc     (last read: qfrc_smooth)
  //   mass matrix      : M, packed lower triangle                     (crb is accumulated in place in cinert)
  s.qpos = take(nq); s.qvel = take(nv);
  s.xipos = take(3 * nbody);
  s.scom = take(3 * nbody); s.smass = take(nbody);
  s.cinert = take(10 * nbody); s.crb = s.cinert;
  s.cdof = take(6 * nv);
  s.f = take(nv); s.aforce = take(nu);
  s.invd = take(64);  // 1 / diag(L)
  const int r0 = o;
  s.xpos = take(3 * nbody); s.xquat = take(4 * nbody); s.xanchor = take(3 * njnt); s.xaxis = take(3 * njnt); s.xmat = take(9 * nbody);
  int end = o;
  o = r0;
  s.cand = take(ncand > 6 * nbody ? ncand : 6 * nbody);   // candidate pair ids (ints stored in real-sized slots)
  s.cfrc = s.cand;
  s.cvel = take(6 * nbody); s.cdofdot = take(6 * nv); s.cacc = take(6 * nbody);
  if (o > end) end = o;
  o = r0;
  s.M = take(b2_poff(nv));   // packed lower triangle
  if (o > end) end = o;
  s.total = end;
  return s;
}

#define B2_MAXCON_PER_PAIR 4

struct LaneContacts {
  int n;
  real dist[B2_MAXCON_PER_PAIR];
  real pos[B2_MAXCON_PER_PAIR][3];
  real normal[3];
  real tangent[3];  // zero => pick default
  real normal_b[3]; // contacts nb_from.. carry their own normal (capsule-box: one sphere-box test per contact)
  int nb_from;
};

B2_DEV void sphere_sphere(LaneContacts& lc, const real* p1, real r1, const real* p2, real r2, real margin) {  // @region narrowphase_fns
  real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  real c = sqrt(dot3(dif, dif));
  if (c > margin + r1 + r2) return;
  real n[3];
  if (c >= B2_MINVAL) { real ic = (real)1 / c; n[0] = dif[0] * ic; n[1] = dif[1] * ic; n[2] = dif[2] * ic; }
  else { n[0] = 1; n[1] = 0; n[2] = 0; }
  real dist = c - r1 - r2;
  int k = lc.n;
  if (k >= B2_MAXCON_PER_PAIR) return;
  lc.dist[k] = dist;
  real s = r1 + (real)0.5 * dist;
  lc.pos[k][0] = p1[0] + n[0] * s; lc.pos[k][1] = p1[1] + n[1] * s; lc.pos[k][2] = p1[2] + n[2] * s;
  // all contacts of a pair produced through this helper share one normal except
  // in the parallel-capsule branch, where normals coincide up to rounding.
  lc.normal[0] = n[0]; lc.normal[1] = n[1]; lc.normal[2] = n[2];
  lc.n = k + 1;
}

B2_DEV void plane_sphere(LaneContacts& lc, const real* pp, const real* pn, const real* sp, real r, real margin) {
  real d[3] = {sp[0] - pp[0], sp[1] - pp[1], sp[2] - pp[2]};
  real dist = dot3(d, pn) - r;
  if (dist > margin) return;
  int k = lc.n;
  real s = r + (real)0.5 * dist;
  lc.dist[k] = dist;
  lc.pos[k][0] = sp[0] - pn[0] * s; lc.pos[k][1] = sp[1] - pn[1] * s; lc.pos[k][2] = sp[2] - pn[2] * s;
  lc.n = k + 1;
}

// Sphere against box (MuJoCo's sphere-box routine as published: clamp the centre into the box; a centre inside the
// box leaves through the nearest face).  Appends at most one contact; normal from the sphere towards the box.
B2_DEV void sphere_box(LaneContacts& lc, const real* sp, real r, const real* bp, const real* bR, const real* bs, real margin) {
  real t[3] = {sp[0] - bp[0], sp[1] - bp[1], sp[2] - bp[2]}, c[3], dif[3];
  mulmTv3(c, bR, t);
  for (int j = 0; j < 3; ++j) dif[j] = clampr(c[j], -bs[j], bs[j]) - c[j];
  real dd = sqrt(dot3(dif, dif));
  if (dd - r > margin) return;
  real nb[3] = {0, 0, 0}, pb[3], dist;
  if (dd <= B2_MINVAL) {
    real closest = 2 * (bs[0] + bs[1] + bs[2]);
    int k = 0;
    for (int i = 0; i < 6; ++i) {
      real fd = fabs(((i & 1) ? bs[i >> 1] : -bs[i >> 1]) - c[i >> 1]);
      if (fd < closest) { closest = fd; k = i; }
    }
    real sg = (k & 1) ? (real)-1 : (real)1;
    if ((k >> 1) == 0) nb[0] = sg; else if ((k >> 1) == 1) nb[1] = sg; else nb[2] = sg;
    real s = (r - closest) * (real)0.5;
    pb[0] = c[0] + nb[0] * s; pb[1] = c[1] + nb[1] * s; pb[2] = c[2] + nb[2] * s;
    dist = -closest - r;
  } else {
    real id = (real)1 / dd;
    nb[0] = dif[0] * id; nb[1] = dif[1] * id; nb[2] = dif[2] * id;
    dist = dd - r;
    real s = r + (real)0.5 * dist;     // midway between the sphere's deepest point and the box surface
    pb[0] = c[0] + nb[0] * s; pb[1] = c[1] + nb[1] * s; pb[2] = c[2] + nb[2] * s;
  }
  int k = lc.n;
  if (k >= B2_MAXCON_PER_PAIR) return;
  real n[3], p[3];
  mulmv3(n, bR, nb);
  mulmv3(p, bR, pb);
  lc.dist[k] = dist;
  lc.pos[k][0] = bp[0] + p[0]; lc.pos[k][1] = bp[1] + p[1]; lc.pos[k][2] = bp[2] + p[2];
  if (k == 0) { lc.normal[0] = n[0]; lc.normal[1] = n[1]; lc.normal[2] = n[2]; }
  else { lc.normal_b[0] = n[0]; lc.normal_b[1] = n[1]; lc.normal_b[2] = n[2]; lc.nb_from = k; }
  lc.n = k + 1;
}

// g(t) = 1/2 d/dt dist^2(p + a t, box): nondecreasing and piecewise linear in t (the squared distance to a box is convex)
B2_DEV real capsule_box_slope(const real* p, const real* a, const real* bs, real t) {
  real s = 0;
  for (int j = 0; j < 3; ++j) {
    real q = p[j] + a[j] * t;
    real e = q > bs[j] ? q - bs[j] : (q < -bs[j] ? q + bs[j] : (real)0);
    s += a[j] * e;
  }
  return s;
}

#define B2_CAPBOX_BISECT 24

// Capsule against box: up to two contacts, each a sphere-box test at a point of the capsule's segment (as MuJoCo's
// capsule-box routine emits them).  The points: the two ends [ta, tb] of the stretch of the segment that is closest to the
// box when that stretch has a length (segment parallel to a face, or inside the box), else the closest point and the
// segment end farther from it.  [ta, tb] come from bisecting the monotone slope g(t).
B2_DEV void capsule_box(LaneContacts& lc, const real* cp, const real* cR, const real* cs, const real* bp, const real* bR,
                        const real* bs, real margin) {
  const real r = cs[0], h = cs[1];
  real ax[3] = {cR[2], cR[5], cR[8]};
  real t[3] = {cp[0] - bp[0], cp[1] - bp[1], cp[2] - bp[2]}, p[3], a[3];
  mulmTv3(p, bR, t);
  mulmTv3(a, bR, ax);
  real g_lo = capsule_box_slope(p, a, bs, -h), g_hi = capsule_box_slope(p, a, bs, h);
  real ta, tb;
  if (g_lo >= 0) ta = -h;
  else if (g_hi < 0) ta = h;
  else {
    real lo = -h, hi = h;
    B2_UNROLL1
    for (int it = 0; it < B2_CAPBOX_BISECT; ++it) {
      real mid = (real)0.5 * (lo + hi);
      if (capsule_box_slope(p, a, bs, mid) >= 0) hi = mid; else lo = mid;
    }
    ta = hi;
  }
  if (g_hi <= 0) tb = h;
  else if (g_lo > 0) tb = -h;
  else {
    real lo = -h, hi = h;
    B2_UNROLL1
    for (int it = 0; it < B2_CAPBOX_BISECT; ++it) {
      real mid = (real)0.5 * (lo + hi);
      if (capsule_box_slope(p, a, bs, mid) <= 0) lo = mid; else hi = mid;
    }
    tb = lo;
  }
  const real tol = (real)1e-4 * h;
  real t1, t2;
  if (tb - ta > tol) { t1 = ta; t2 = tb; }
  else { t1 = (real)0.5 * (ta + tb); t2 = t1 > 0 ? -h : h; }
  real q[3] = {cp[0] + ax[0] * t1, cp[1] + ax[1] * t1, cp[2] + ax[2] * t1};
  sphere_box(lc, q, r, bp, bR, bs, margin);
  if (fabs(t2 - t1) > tol) {
    q[0] = cp[0] + ax[0] * t2; q[1] = cp[1] + ax[1] * t2; q[2] = cp[2] + ax[2] * t2;
    sphere_box(lc, q, r, bp, bR, bs, margin);
  }
}

// Box against box (oracle/mjstep.py `_box_box`, statement by statement): separating-axis test over 3 + 3 face normals and 9
// edge cross products; a face axis clips the other box's incident face against the side planes of the reference face
// (Sutherland-Hodgman, <= 8 points), keeps the points at most `margin` above the reference face and reduces them to <= 4
// (deepest, farthest from it, farthest from that line on either side); an edge axis gives one contact between the closest
// points of the two supporting edges.  All contacts of the pair share one normal (box 1 -> box 2).
#define B2_BOXBOX_PARALLEL ((real)1e-10)
#define B2_BOXBOX_FACE_BIAS ((real)0.05)
B2_DEV real boxbox_radius(const real* R, const real* s, const real* L) {
  return s[0] * fabs(R[0] * L[0] + R[3] * L[1] + R[6] * L[2]) + s[1] * fabs(R[1] * L[0] + R[4] * L[1] + R[7] * L[2]) +
         s[2] * fabs(R[2] * L[0] + R[5] * L[1] + R[8] * L[2]);
}
B2_DEVN void box_box(LaneContacts& lc, const real* p1, const real* R1, const real* s1, const real* p2, const real* R2,
                     const real* s2, real margin) {
  const real d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  real best_s = -INFINITY, best_sign = 1;
  int best_code = 0;
  B2_UNROLL1
  for (int code = 0; code < 6; ++code) {
    const real* R = code < 3 ? R1 : R2;
    const int k = code < 3 ? code : code - 3;
    real L[3] = {R[k], R[3 + k], R[6 + k]};
    real dl = dot3(d, L);
    real sep = fabs(dl) - boxbox_radius(R1, s1, L) - boxbox_radius(R2, s2, L);
    if (sep > margin) return;
    if (sep > best_s) { best_s = sep; best_code = code; best_sign = dl >= 0 ? (real)1 : (real)-1; }
  }
  real edge_s = -INFINITY, en[3] = {0, 0, 0};
  int ei = 0, ej = 0;
  bool have_edge = false;
  B2_UNROLL1
  for (int i = 0; i < 3; ++i) {
    B2_UNROLL1
    for (int j = 0; j < 3; ++j) {
      real a[3] = {R1[i], R1[3 + i], R1[6 + i]}, b[3] = {R2[j], R2[3 + j], R2[6 + j]}, L[3];
      cross3(L, a, b);
      real n2 = dot3(L, L);
      if (n2 < B2_BOXBOX_PARALLEL) continue;
      real inv = (real)1 / sqrt(n2);
      L[0] *= inv; L[1] *= inv; L[2] *= inv;
      real dl = dot3(d, L);
      real sep = fabs(dl) - boxbox_radius(R1, s1, L) - boxbox_radius(R2, s2, L);
      if (sep > margin) return;
      if (sep > edge_s) {
        edge_s = sep; ei = i; ej = j; have_edge = true;
        real sg = dl >= 0 ? (real)1 : (real)-1;
        en[0] = L[0] * sg; en[1] = L[1] * sg; en[2] = L[2] * sg;
      }
    }
  }
  if (have_edge && edge_s > best_s + B2_BOXBOX_FACE_BIAS * fabs(best_s) + (real)1e-9) {
    real pa[3] = {p1[0], p1[1], p1[2]}, pb[3] = {p2[0], p2[1], p2[2]};
    for (int k = 0; k < 3; ++k) {
      if (k != ei) {
        real ax[3] = {R1[k], R1[3 + k], R1[6 + k]};
        real sg = (dot3(ax, en) >= 0 ? (real)1 : (real)-1) * s1[k];
        pa[0] += sg * ax[0]; pa[1] += sg * ax[1]; pa[2] += sg * ax[2];
      }
      if (k != ej) {
        real ax[3] = {R2[k], R2[3 + k], R2[6 + k]};
        real sg = (dot3(ax, en) >= 0 ? (real)1 : (real)-1) * s2[k];
        pb[0] -= sg * ax[0]; pb[1] -= sg * ax[1]; pb[2] -= sg * ax[2];
      }
    }
    real a[3] = {R1[ei], R1[3 + ei], R1[6 + ei]}, b[3] = {R2[ej], R2[3 + ej], R2[6 + ej]};
    real w[3] = {pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]};
    real bb = dot3(a, b), dd = dot3(a, w), e = dot3(b, w);
    real den = (real)1 - bb * bb;
    real t = clampr((bb * e - dd) / den, -s1[ei], s1[ei]);
    real u = clampr((e - bb * dd) / den, -s2[ej], s2[ej]);
    real qa[3] = {pa[0] + t * a[0], pa[1] + t * a[1], pa[2] + t * a[2]};
    real qb[3] = {pb[0] + u * b[0], pb[1] + u * b[1], pb[2] + u * b[2]};
    real df[3] = {qb[0] - qa[0], qb[1] - qa[1], qb[2] - qa[2]};
    real dist = dot3(df, en);
    if (dist > margin) return;
    lc.dist[0] = dist;
    lc.pos[0][0] = (real)0.5 * (qa[0] + qb[0]); lc.pos[0][1] = (real)0.5 * (qa[1] + qb[1]); lc.pos[0][2] = (real)0.5 * (qa[2] + qb[2]);
    lc.normal[0] = en[0]; lc.normal[1] = en[1]; lc.normal[2] = en[2];
    lc.n = 1;
    return;
  }
  // face contact: reference box r, incident box i
  const bool flip = best_code >= 3;
  const real *pr = flip ? p2 : p1, *Rr = flip ? R2 : R1, *sr = flip ? s2 : s1;
  const real *pi = flip ? p1 : p2, *Ri = flip ? R1 : R2, *si = flip ? s1 : s2;
  const int axis = flip ? best_code - 3 : best_code;
  const real nsg = flip ? -best_sign : best_sign;
  const real nref[3] = {Rr[axis] * nsg, Rr[3 + axis] * nsg, Rr[6 + axis] * nsg};
  int ks = 0;
  real kbest = -1;
  for (int k = 0; k < 3; ++k) {
    real v = fabs(Ri[k] * nref[0] + Ri[3 + k] * nref[1] + Ri[6 + k] * nref[2]);
    if (v > kbest) { ks = k; kbest = v; }
  }
  const real sgi = (Ri[ks] * nref[0] + Ri[3 + ks] * nref[1] + Ri[6 + ks] * nref[2]) > 0 ? (real)-1 : (real)1;
  const int iu = (ks + 1) % 3, iv = (ks + 2) % 3, ru = (axis + 1) % 3, rv = (axis + 2) % 3;
  const real ci[3] = {pi[0] + sgi * si[ks] * Ri[ks], pi[1] + sgi * si[ks] * Ri[3 + ks], pi[2] + sgi * si[ks] * Ri[6 + ks]};
  real poly[8][3], nxt[8][3];
  int np_ = 4;
  for (int v = 0; v < 4; ++v) {
    const real a = (v == 1 || v == 2) ? (real)1 : (real)-1, b = v >= 2 ? (real)1 : (real)-1;
    real q[3];
    for (int c = 0; c < 3; ++c) q[c] = ci[c] + a * si[iu] * Ri[3 * c + iu] + b * si[iv] * Ri[3 * c + iv] - pr[c];
    poly[v][0] = q[0] * Rr[ru] + q[1] * Rr[3 + ru] + q[2] * Rr[6 + ru];
    poly[v][1] = q[0] * Rr[rv] + q[1] * Rr[3 + rv] + q[2] * Rr[6 + rv];
    poly[v][2] = dot3(q, nref) - sr[axis];
  }
  B2_UNROLL1
  for (int pl = 0; pl < 4; ++pl) {
    const int c = pl >> 1;
    const real lim = c == 0 ? sr[ru] : sr[rv], sgn = (pl & 1) ? (real)-1 : (real)1;
    int nn = 0;
    B2_UNROLL1
    for (int k = 0; k < np_; ++k) {
      const real* cur = poly[k];
      const real* prv = poly[k == 0 ? np_ - 1 : k - 1];
      const real fc = sgn * cur[c] - lim, fp = sgn * prv[c] - lim;
      if ((fc <= 0) != (fp <= 0)) {
        const real tt = fp / (fp - fc);
        if (nn < 8) { for (int e = 0; e < 3; ++e) nxt[nn][e] = prv[e] + tt * (cur[e] - prv[e]); ++nn; }
      }
      if (fc <= 0 && nn < 8) { for (int e = 0; e < 3; ++e) nxt[nn][e] = cur[e]; ++nn; }
    }
    np_ = nn;
    for (int k = 0; k < nn; ++k) for (int e = 0; e < 3; ++e) poly[k][e] = nxt[k][e];
    if (np_ == 0) return;
  }
  int nc = 0;
  for (int k = 0; k < np_; ++k)
    if (poly[k][2] <= margin) { for (int e = 0; e < 3; ++e) nxt[nc][e] = poly[k][e]; ++nc; }
  int sel[4] = {0, 1, 2, 3}, nsel = nc < 4 ? nc : 4;
  if (nc > B2_MAXCON_PER_PAIR) {
    int i0 = 0;
    for (int k = 1; k < nc; ++k) if (nxt[k][2] < nxt[i0][2]) i0 = k;
    int i1 = 0; real dbest = -1;
    for (int k = 0; k < nc; ++k) {
      real dx = nxt[k][0] - nxt[i0][0], dy = nxt[k][1] - nxt[i0][1], d2 = dx * dx + dy * dy;
      if (d2 > dbest) { dbest = d2; i1 = k; }
    }
    const real ex = nxt[i1][0] - nxt[i0][0], ey = nxt[i1][1] - nxt[i0][1];
    int i2 = 0; real cbest = -1, cr2 = 0;
    for (int k = 0; k < nc; ++k) {
      real cr = ex * (nxt[k][1] - nxt[i0][1]) - ey * (nxt[k][0] - nxt[i0][0]);
      if (fabs(cr) > cbest) { cbest = fabs(cr); i2 = k; cr2 = cr; }
    }
    int i3 = 0; real obest = -INFINITY;
    for (int k = 0; k < nc; ++k) {
      real cr = ex * (nxt[k][1] - nxt[i0][1]) - ey * (nxt[k][0] - nxt[i0][0]);
      real o = cr2 > 0 ? -cr : cr;
      if (o > obest) { obest = o; i3 = k; }
    }
    sel[0] = i0; sel[1] = i1; sel[2] = i2; nsel = 3;
    if (obest > 0 && i3 != i0 && i3 != i1 && i3 != i2) { sel[3] = i3; nsel = 4; }
  }
  for (int q = 0; q < nsel; ++q) {
    const real* pt = nxt[sel[q]];
    const real cu = pt[0], cv = pt[1], cn = pt[2], hn = sr[axis] + (real)0.5 * cn;
    lc.dist[q] = cn;
    for (int c = 0; c < 3; ++c) lc.pos[q][c] = pr[c] + cu * Rr[3 * c + ru] + cv * Rr[3 * c + rv] + hn * nref[c];
  }
  const real fs = flip ? (real)-1 : (real)1;
  lc.normal[0] = fs * nref[0]; lc.normal[1] = fs * nref[1]; lc.normal[2] = fs * nref[2];
  lc.n = nsel;
}

// Narrowphase for one geom pair (types t1 <= t2). Normal points from geom1 to geom2.
B2_DEV void narrowphase(LaneContacts& lc, int t1, int t2, const real* p1, const real* R1, const real* s1, const real* p2,
                        const real* R2, const real* s2, real margin) {
  lc.n = 0;
  lc.nb_from = B2_MAXCON_PER_PAIR;
  lc.tangent[0] = lc.tangent[1] = lc.tangent[2] = 0;
  if (t1 == B2_GEOM_PLANE) {
    real n[3] = {R1[2], R1[5], R1[8]};
    lc.normal[0] = n[0]; lc.normal[1] = n[1]; lc.normal[2] = n[2];
    if (t2 == B2_GEOM_SPHERE) {
      plane_sphere(lc, p1, n, p2, s2[0], margin);
    } else if (t2 == B2_GEOM_CAPSULE) {
      real ax[3] = {R2[2], R2[5], R2[8]};
      real e[3];
      e[0] = p2[0] + ax[0] * s2[1]; e[1] = p2[1] + ax[1] * s2[1]; e[2] = p2[2] + ax[2] * s2[1];
      plane_sphere(lc, p1, n, e, s2[0], margin);
      e[0] = p2[0] - ax[0] * s2[1]; e[1] = p2[1] - ax[1] * s2[1]; e[2] = p2[2] - ax[2] * s2[1];
      plane_sphere(lc, p1, n, e, s2[0], margin);
      lc.tangent[0] = ax[0]; lc.tangent[1] = ax[1]; lc.tangent[2] = ax[2];
    } else if (t2 == B2_GEOM_BOX) {
      real d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      real dist0 = dot3(d, n);
      for (int i = 0; i < 8; ++i) {
        real v[3] = {(i & 1) ? s2[0] : -s2[0], (i & 2) ? s2[1] : -s2[1], (i & 4) ? s2[2] : -s2[2]};
        real c[3];
        mulmv3(c, R2, v);
        real ld = dot3(n, c);
        if (dist0 + ld > margin || ld > 0) continue;
        real dist = dist0 + ld;
        int k = lc.n;
        lc.dist[k] = dist;
        lc.pos[k][0] = c[0] + p2[0] - n[0] * dist * (real)0.5;
        lc.pos[k][1] = c[1] + p2[1] - n[1] * dist * (real)0.5;
        lc.pos[k][2] = c[2] + p2[2] - n[2] * dist * (real)0.5;
        lc.n = k + 1;
        if (lc.n >= 4) break;
      }
    }
  } else if (t1 == B2_GEOM_SPHERE && t2 == B2_GEOM_SPHERE) {
    sphere_sphere(lc, p1, s1[0], p2, s2[0], margin);
  } else if (t1 == B2_GEOM_SPHERE && t2 == B2_GEOM_CAPSULE) {
    real ax[3] = {R2[2], R2[5], R2[8]};
    real d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    real x = clampr(dot3(ax, d), -s2[1], s2[1]);
    real q[3] = {p2[0] + ax[0] * x, p2[1] + ax[1] * x, p2[2] + ax[2] * x};
    sphere_sphere(lc, p1, s1[0], q, s2[0], margin);
  } else if (t1 == B2_GEOM_CAPSULE && t2 == B2_GEOM_CAPSULE) {
    real a1[3] = {R1[2], R1[5], R1[8]}, a2[3] = {R2[2], R2[5], R2[8]};
    real dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    real ma = 1, mb = -dot3(a1, a2), mc = 1;
    real u = -dot3(a1, dif), v = dot3(a2, dif);
    real det = ma * mc - mb * mb;
    real l1 = s1[1], l2 = s2[1];
    if (fabs(det) >= B2_MINVAL) {
      real x1 = (mc * u - mb * v) / det;
      real x2 = (ma * v - mb * u) / det;
      if (x1 > l1) { x1 = l1; x2 = (v - mb * x1) / mc; }
      else if (x1 < -l1) { x1 = -l1; x2 = (v - mb * x1) / mc; }
      if (x2 > l2) { x2 = l2; x1 = (u - mb * x2) / ma; }
      else if (x2 < -l2) { x2 = -l2; x1 = (u - mb * x2) / ma; }
      x1 = clampr(x1, -l1, l1);
      real q1[3] = {p1[0] + a1[0] * x1, p1[1] + a1[1] * x1, p1[2] + a1[2] * x1};
      real q2[3] = {p2[0] + a2[0] * x2, p2[1] + a2[1] * x2, p2[2] + a2[2] * x2};
      sphere_sphere(lc, q1, s1[0], q2, s2[0], margin);
    } else {
      B2_UNROLL1
      for (int e = 0; e < 2 && lc.n < 2; ++e) {
        real sg = e == 0 ? (real)1 : (real)-1;
        real v1[3] = {p1[0] + a1[0] * l1 * sg, p1[1] + a1[1] * l1 * sg, p1[2] + a1[2] * l1 * sg};
        real dd[3] = {v1[0] - p2[0], v1[1] - p2[1], v1[2] - p2[2]};
        real x2 = clampr(dot3(dd, a2), -l2, l2);
        real q2[3] = {p2[0] + a2[0] * x2, p2[1] + a2[1] * x2, p2[2] + a2[2] * x2};
        sphere_sphere(lc, v1, s1[0], q2, s2[0], margin);
      }
      B2_UNROLL1
      for (int e = 0; e < 2 && lc.n < 2; ++e) {
        real sg = e == 0 ? (real)1 : (real)-1;
        real v2[3] = {p2[0] + a2[0] * l2 * sg, p2[1] + a2[1] * l2 * sg, p2[2] + a2[2] * l2 * sg};
        real dd[3] = {v2[0] - p1[0], v2[1] - p1[1], v2[2] - p1[2]};
        real x1 = clampr(dot3(dd, a1), -l1, l1);
        real q1[3] = {p1[0] + a1[0] * x1, p1[1] + a1[1] * x1, p1[2] + a1[2] * x1};
        sphere_sphere(lc, q1, s1[0], v2, s2[0], margin);
      }
    }
  } else if (t1 == B2_GEOM_SPHERE && t2 == B2_GEOM_BOX) {
    sphere_box(lc, p1, s1[0], p2, R2, s2, margin);
  } else if (t1 == B2_GEOM_CAPSULE && t2 == B2_GEOM_BOX) {
    capsule_box(lc, p1, R1, s1, p2, R2, s2, margin);
  } else if (t1 == B2_GEOM_BOX && t2 == B2_GEOM_BOX) {
    box_box(lc, p1, R1, s1, p2, R2, s2, margin);
  }
  // unsupported pairs (meshes, height fields, ...) are rejected on the host (engine.check_supported)
}

// Contact frame rows [normal; t1; t2] from a normal and an optional tangent hint.
B2_DEV void make_frame(real* F, const real* normal, const real* tangent) {
  real nn = b2_rsqrt(dot3(normal, normal));
  real n[3] = {normal[0] * nn, normal[1] * nn, normal[2] * nn};
  real t[3] = {tangent[0], tangent[1], tangent[2]};
  if (dot3(t, t) < (real)0.25) {
    t[0] = 0; t[1] = 0; t[2] = 0;
    if (n[1] < (real)0.5 && n[1] > (real)-0.5) t[1] = 1; else t[2] = 1;
  }
  real d = dot3(n, t);
  t[0] -= n[0] * d; t[1] -= n[1] * d; t[2] -= n[2] * d;
  real tn = b2_rsqrt(dot3(t, t));
  t[0] *= tn; t[1] *= tn; t[2] *= tn;
  F[0] = n[0]; F[1] = n[1]; F[2] = n[2];
  F[3] = t[0]; F[4] = t[1]; F[5] = t[2];
  cross3(F + 6, n, t);
}

#if defined(B2_TIMING) && !defined(B2_HOSTEMU)
// diagnostics build only (tools/solve_timeline.py): clock64 per phase of smooth_world
__device__ unsigned long long g_sphase[8 * 16384];
#define B2_SPH_BEGIN() long long sph_t = clock64(); unsigned long long sph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define B2_SPH(k) { long long sph_n = clock64(); sph_acc[k] += (unsigned long long)(sph_n - sph_t); sph_t = sph_n; }
#define B2_SPH_END() if (B2_LANE == 0 && w < 16384) { for (int k = 0; k < 8; ++k) g_sphase[8 * w + k] = sph_acc[k]; }
#else
#define B2_SPH_BEGIN()
#define B2_SPH(k)
#define B2_SPH_END()
#endif

// ---------------------------------------------------------------------------
// Stage 1 for world w.  sm: this warp's shared-memory slab.
// ---------------------------------------------------------------------------
#ifdef B2_HOSTEMU
static long long b2h_memo_hits = 0;
#endif
// memo_consume: reuse this world's position / velocity stage outputs in HBM when its state is bit for bit the one they were
// computed from (the env step's first sub-step after the forward pass that follows the resets): only the ctrl-dependent
// tail (actuator forces, qfrc_smooth, qacc_smooth) is redone, through the very same instructions as in the full pass.
B2_DEV void smooth_world(const KModel& m, const KData& d, const SmoothSmem& L, int ncand_max, int w, real* sm, int memo_consume = 0) {  // @region setup
  B2_SPH_BEGIN()
  const int nq = m.nq, nv = m.nv, nu = m.nu, nb = m.nbody, nj = m.njnt;
  real* qpos = sm + L.qpos; real* qvel = sm + L.qvel;
  real* xpos = sm + L.xpos; real* xquat = sm + L.xquat; real* xmat = sm + L.xmat; real* xipos = sm + L.xipos;
  real* xanchor = sm + L.xanchor; real* xaxis = sm + L.xaxis;
  real* scom = sm + L.scom; real* smass = sm + L.smass;
  real* cinert = sm + L.cinert; real* crb = sm + L.crb;
  real* cdof = sm + L.cdof; real* cdofdot = sm + L.cdofdot;
  real* cvel = sm + L.cvel; real* cacc = sm + L.cacc; real* cfrc = sm + L.cfrc;
  real* M = sm + L.M; real* f = sm + L.f; real* aforce = sm + L.aforce;
  int* cand = (int*)(sm + L.cand);

  const int* body_parent = MI(B2_MI_BODY_PARENTID);
  const int* body_root = MI(B2_MI_BODY_ROOTID);
  const int* body_mocap = MI(B2_MI_BODY_MOCAPID);
  const int* body_jntnum = MI(B2_MI_BODY_JNTNUM);
  const int* body_jntadr = MI(B2_MI_BODY_JNTADR);
  const int* body_dofnum = MI(B2_MI_BODY_DOFNUM);
  const int* body_dofadr = MI(B2_MI_BODY_DOFADR);
  const int* level_adr = MI(B2_MI_LEVEL_ADR);
  const int* level_body = MI(B2_MI_LEVEL_BODY);
  const int* child_adr = MI(B2_MI_CHILD_ADR);
  const int* child_body = MI(B2_MI_CHILD_BODY);
  const int* jnt_type = MI(B2_MI_JNT_TYPE);
  const int* jnt_qposadr = MI(B2_MI_JNT_QPOSADR);
  const int* jnt_dofadr = MI(B2_MI_JNT_DOFADR);
  const int* jnt_bodyid = MI(B2_MI_JNT_BODYID);
  const int* dof_bodyid = MI(B2_MI_DOF_BODYID);
  const int* dof_parentid = MI(B2_MI_DOF_PARENTID);

  real* g_qpos = d.df[B2_DF_QPOS] + (size_t)w * nq;
  const real* g_qvel = d.df[B2_DF_QVEL] + (size_t)w * nv;

  B2_UNROLL1
  for (int i = B2_LANE; i < nq; i += B2_NLANE) qpos[i] = g_qpos[i];
  B2_UNROLL1
  for (int i = B2_LANE; i < nv; i += B2_NLANE) qvel[i] = g_qvel[i];
  // xfrc_applied is almost always all zero; its 6*nbody loads join the state loads here so that their HBM latency is
  // paid once, instead of stalling the force stage later (it was 2.6 % of the kernel's stall samples there)
  int anyx_lane = 0;
  {
    const real* xf = d.df[B2_DF_XFRC_APPLIED] + (size_t)w * nb * 6;
    B2_UNROLL1
    for (int i = B2_LANE; i < 6 * nb; i += B2_NLANE) anyx_lane |= (xf[i] != 0);
  }
  B2_WSYNC();
  const int ld = b2_ld(nv);
  bool light = false;
  {
    real* memo_q = d.df[B2_DF_MEMO_QPOS] + (size_t)w * nq;
    real* memo_v = d.df[B2_DF_MEMO_QVEL] + (size_t)w * nv;
    const int* memo_ok = d.di[B2_DI_SMOOTH_MEMO] + w;
    if (memo_consume && m.nmocap == 0 && *memo_ok == 1) {
      int diff = anyx_lane;   // an external wrench needs the kinematic quantities: full pass
      B2_UNROLL1
      for (int i = B2_LANE; i < nq; i += B2_NLANE) diff |= !b2_same_bits(qpos[i], memo_q[i]);
      B2_UNROLL1
      for (int i = B2_LANE; i < nv; i += B2_NLANE) diff |= !b2_same_bits(qvel[i], memo_v[i]);
      // the plain pass would normalise the free joints' quaternions inside qpos again: only a fixed point may skip it
      B2_UNROLL1
      for (int j = B2_LANE; j < nj; j += B2_NLANE) {
        if (jnt_type[j] != B2_JNT_FREE) continue;
        const real* q0 = qpos + jnt_qposadr[j] + 3;
        real q[4] = {q0[0], q0[1], q0[2], q0[3]};
        quat_normalize(q);
        diff |= !(b2_same_bits(q[0], q0[0]) && b2_same_bits(q[1], q0[1]) && b2_same_bits(q[2], q0[2]) && b2_same_bits(q[3], q0[3]));
      }
      light = !b2_wor_i(diff);
#ifdef B2_HOSTEMU
      b2h_memo_hits += light;   // test harness statistic
#endif
    }
  }

  if (!light) {   // ======== position / velocity stage (skipped when the memo holds) ========
  // ---- kinematics, level by level ------------------------------------------  // @region kinematics
  const real* qpos0 = MF(B2_MF_QPOS0);
  const real* body_pos = MF(B2_MF_BODY_POS);
  const real* body_quat = MF(B2_MF_BODY_QUAT);
  const real* body_ipos = MF(B2_MF_BODY_IPOS);
  const real* body_iquat = MF(B2_MF_BODY_IQUAT);
  const real* jnt_pos = MF(B2_MF_JNT_POS);
  const real* jnt_axis = MF(B2_MF_JNT_AXIS);
  real* g_ximat = d.df[B2_DF_XIMAT] + (size_t)w * nb * 9;
  B2_UNROLL1
  for (int lev = 0; lev < m.nlevel; ++lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      real pos[3], quat[4], R[9];
      if (b == 0) {
        pos[0] = pos[1] = pos[2] = 0; quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
      } else {
        int p = body_parent[b];
        if (body_mocap[b] >= 0) {
          const real* mp = d.df[B2_DF_MOCAP_POS] + ((size_t)w * m.nmocap + body_mocap[b]) * 3;
          const real* mq = d.df[B2_DF_MOCAP_QUAT] + ((size_t)w * m.nmocap + body_mocap[b]) * 4;
          pos[0] = mp[0]; pos[1] = mp[1]; pos[2] = mp[2];
          quat[0] = mq[0]; quat[1] = mq[1]; quat[2] = mq[2]; quat[3] = mq[3];
          quat_normalize(quat);
        } else {
          real t[3];
          mulmv3(t, xmat + 9 * p, body_pos + 3 * b);
          pos[0] = xpos[3 * p] + t[0]; pos[1] = xpos[3 * p + 1] + t[1]; pos[2] = xpos[3 * p + 2] + t[2];
          quat_mul(quat, xquat + 4 * p, body_quat + 4 * b);
        }
        B2_UNROLL1
        for (int jj = 0; jj < body_jntnum[b]; ++jj) {
          int j = body_jntadr[b] + jj;
          int qa = jnt_qposadr[j];
          int t = jnt_type[j];
          if (t == B2_JNT_FREE) {
            pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
            quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
            quat_normalize(quat);
            // mj_kinematics normalises the quaternion inside qpos
            qpos[qa + 3] = quat[0]; qpos[qa + 4] = quat[1]; qpos[qa + 5] = quat[2]; qpos[qa + 6] = quat[3];
            g_qpos[qa + 3] = quat[0]; g_qpos[qa + 4] = quat[1]; g_qpos[qa + 5] = quat[2]; g_qpos[qa + 6] = quat[3];
            quat2mat(R, quat);
            xanchor[3 * j] = pos[0]; xanchor[3 * j + 1] = pos[1]; xanchor[3 * j + 2] = pos[2];
            xaxis[3 * j] = R[2]; xaxis[3 * j + 1] = R[5]; xaxis[3 * j + 2] = R[8];
          } else {
            quat2mat(R, quat);
            real a[3], ax[3];
            mulmv3(a, R, jnt_pos + 3 * j);
            a[0] += pos[0]; a[1] += pos[1]; a[2] += pos[2];
            mulmv3(ax, R, jnt_axis + 3 * j);
            xanchor[3 * j] = a[0]; xanchor[3 * j + 1] = a[1]; xanchor[3 * j + 2] = a[2];
            xaxis[3 * j] = ax[0]; xaxis[3 * j + 1] = ax[1]; xaxis[3 * j + 2] = ax[2];
            real dq = qpos[qa] - qpos0[qa];
            if (t == B2_JNT_HINGE) {
              real ql[4], q2[4];
              axisangle2quat(ql, jnt_axis + 3 * j, dq);
              quat_mul(q2, quat, ql);
              quat[0] = q2[0]; quat[1] = q2[1]; quat[2] = q2[2]; quat[3] = q2[3];
              quat2mat(R, quat);
              real t2[3];
              mulmv3(t2, R, jnt_pos + 3 * j);
              pos[0] = a[0] - t2[0]; pos[1] = a[1] - t2[1]; pos[2] = a[2] - t2[2];
            } else {  // slide
              pos[0] += ax[0] * dq; pos[1] += ax[1] * dq; pos[2] += ax[2] * dq;
            }
          }
        }
        quat_normalize(quat);
      }
      quat2mat(R, quat);
      for (int c = 0; c < 3; ++c) xpos[3 * b + c] = pos[c];
      for (int c = 0; c < 4; ++c) xquat[4 * b + c] = quat[c];
      for (int c = 0; c < 9; ++c) xmat[9 * b + c] = R[c];
      real t[3];
      mulmv3(t, R, body_ipos + 3 * b);
      xipos[3 * b] = pos[0] + t[0]; xipos[3 * b + 1] = pos[1] + t[1]; xipos[3 * b + 2] = pos[2] + t[2];
      real qi[4], Ri[9];
      quat_mul(qi, quat, body_iquat + 4 * b);
      quat2mat(Ri, qi);
      for (int c = 0; c < 9; ++c) g_ximat[9 * b + c] = Ri[c];
    }
    B2_WSYNC();
  }
  // body poses -> HBM (coalesced row writes)
  {
    real* g;
    g = d.df[B2_DF_XPOS] + (size_t)w * nb * 3;   for (int i = B2_LANE; i < 3 * nb; i += B2_NLANE) g[i] = xpos[i];
    g = d.df[B2_DF_XQUAT] + (size_t)w * nb * 4;  for (int i = B2_LANE; i < 4 * nb; i += B2_NLANE) g[i] = xquat[i];
    g = d.df[B2_DF_XMAT] + (size_t)w * nb * 9;   for (int i = B2_LANE; i < 9 * nb; i += B2_NLANE) g[i] = xmat[i];
    g = d.df[B2_DF_XIPOS] + (size_t)w * nb * 3;  for (int i = B2_LANE; i < 3 * nb; i += B2_NLANE) g[i] = xipos[i];
    g = d.df[B2_DF_XANCHOR] + (size_t)w * nj * 3; for (int i = B2_LANE; i < 3 * nj; i += B2_NLANE) g[i] = xanchor[i];
    g = d.df[B2_DF_XAXIS] + (size_t)w * nj * 3;  for (int i = B2_LANE; i < 3 * nj; i += B2_NLANE) g[i] = xaxis[i];
  }
  // the state this pass describes (after the in-place quaternion normalisation above): key of the memo
  {
    real* memo_q = d.df[B2_DF_MEMO_QPOS] + (size_t)w * nq;
    real* memo_v = d.df[B2_DF_MEMO_QVEL] + (size_t)w * nv;
    B2_UNROLL1
    for (int i = B2_LANE; i < nq; i += B2_NLANE) memo_q[i] = qpos[i];
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) memo_v[i] = qvel[i];
    if (B2_LANE == 0) d.di[B2_DI_SMOOTH_MEMO][w] = 1;
  }
  // geom and site poses straight to HBM
  {
    const int* geom_body = MI(B2_MI_GEOM_BODYID);
    const real* geom_pos = MF(B2_MF_GEOM_POS);
    const real* geom_quat = MF(B2_MF_GEOM_QUAT);
    real* gx = d.df[B2_DF_GEOM_XPOS] + (size_t)w * m.ngeom * 3;
    real* gm = d.df[B2_DF_GEOM_XMAT] + (size_t)w * m.ngeom * 9;
    // moving geoms only: static ones (terrain) hold the constant pose b2_init_static wrote
    const int* dyn_geom = MI(B2_MI_DYN_GEOM);
    B2_UNROLL1
    for (int gi = B2_LANE; gi < m.ndyn; gi += B2_NLANE) {
      int g = dyn_geom[gi];
      int b = geom_body[g];
      real t[3], q[4], R[9];
      mulmv3(t, xmat + 9 * b, geom_pos + 3 * g);
      gx[3 * g] = xpos[3 * b] + t[0]; gx[3 * g + 1] = xpos[3 * b + 1] + t[1]; gx[3 * g + 2] = xpos[3 * b + 2] + t[2];
      quat_mul(q, xquat + 4 * b, geom_quat + 4 * g);
      quat2mat(R, q);
      for (int c = 0; c < 9; ++c) gm[9 * g + c] = R[c];
    }
    const int* site_body = MI(B2_MI_SITE_BODYID);
    const real* site_pos = MF(B2_MF_SITE_POS);
    const real* site_quat = MF(B2_MF_SITE_QUAT);
    real* sx = d.df[B2_DF_SITE_XPOS] + (size_t)w * m.nsite * 3;
    real* smx = d.df[B2_DF_SITE_XMAT] + (size_t)w * m.nsite * 9;
    B2_UNROLL1
    for (int s = B2_LANE; s < m.nsite; s += B2_NLANE) {
      int b = site_body[s];
      real t[3], q[4], R[9];
      mulmv3(t, xmat + 9 * b, site_pos + 3 * s);
      sx[3 * s] = xpos[3 * b] + t[0]; sx[3 * s + 1] = xpos[3 * b + 1] + t[1]; sx[3 * s + 2] = xpos[3 * b + 2] + t[2];
      quat_mul(q, xquat + 4 * b, site_quat + 4 * s);
      quat2mat(R, q);
      for (int c = 0; c < 9; ++c) smx[9 * s + c] = R[c];
    }
  }

  B2_SPH(0)
  // ---- subtree com (backward gather over levels) ----------------------------  // @region subtree_com
  const real* body_mass = MF(B2_MF_BODY_MASS);
  const real* body_inertia = MF(B2_MF_BODY_INERTIA);
  B2_UNROLL1
  for (int lev = m.nlevel - 1; lev >= 0; --lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      real ms = body_mass[b];
      real a0 = ms * xipos[3 * b], a1 = ms * xipos[3 * b + 1], a2 = ms * xipos[3 * b + 2];
      B2_UNROLL1
      for (int c = child_adr[b]; c < child_adr[b + 1]; ++c) {
        int ch = child_body[c];
        a0 += scom[3 * ch]; a1 += scom[3 * ch + 1]; a2 += scom[3 * ch + 2];
        ms += smass[ch];
      }
      scom[3 * b] = a0; scom[3 * b + 1] = a1; scom[3 * b + 2] = a2; smass[b] = ms;
    }
    B2_WSYNC();
  }
  B2_UNROLL1
  for (int b = B2_LANE; b < nb; b += B2_NLANE) {
    real ms = smass[b];
    if (ms > B2_MINVAL) {
      real inv = (real)1 / ms;
      scom[3 * b] *= inv; scom[3 * b + 1] *= inv; scom[3 * b + 2] *= inv;
    } else if (b > 0) {
      scom[3 * b] = xipos[3 * b]; scom[3 * b + 1] = xipos[3 * b + 1]; scom[3 * b + 2] = xipos[3 * b + 2];
    } else {
      scom[0] = scom[1] = scom[2] = 0;
    }
  }
  B2_WSYNC();
  {
    real* g = d.df[B2_DF_SUBTREE_COM] + (size_t)w * nb * 3;
    B2_UNROLL1
    for (int i = B2_LANE; i < 3 * nb; i += B2_NLANE) g[i] = scom[i];
  }
  // ---- cinert, cdof -----------------------------------------------------------  // @region cinert_cdof
  B2_UNROLL1
  for (int b = B2_LANE; b < nb; b += B2_NLANE) {
    int r = body_root[b];
    real o[3] = {xipos[3 * b] - scom[3 * r], xipos[3 * b + 1] - scom[3 * r + 1], xipos[3 * b + 2] - scom[3 * r + 2]};
    real ms = body_mass[b];
    const real* Ri = g_ximat + 9 * b;  // written by the owner lane of b above; visible after the syncs
    real I0 = body_inertia[3 * b], I1 = body_inertia[3 * b + 1], I2 = body_inertia[3 * b + 2];
    real* ci = cinert + 10 * b;
    real oo = dot3(o, o);
    ci[0] = Ri[0] * Ri[0] * I0 + Ri[1] * Ri[1] * I1 + Ri[2] * Ri[2] * I2 + ms * (oo - o[0] * o[0]);
    ci[1] = Ri[3] * Ri[3] * I0 + Ri[4] * Ri[4] * I1 + Ri[5] * Ri[5] * I2 + ms * (oo - o[1] * o[1]);
    ci[2] = Ri[6] * Ri[6] * I0 + Ri[7] * Ri[7] * I1 + Ri[8] * Ri[8] * I2 + ms * (oo - o[2] * o[2]);
    ci[3] = Ri[0] * Ri[3] * I0 + Ri[1] * Ri[4] * I1 + Ri[2] * Ri[5] * I2 - ms * o[0] * o[1];
    ci[4] = Ri[0] * Ri[6] * I0 + Ri[1] * Ri[7] * I1 + Ri[2] * Ri[8] * I2 - ms * o[0] * o[2];
    ci[5] = Ri[3] * Ri[6] * I0 + Ri[4] * Ri[7] * I1 + Ri[5] * Ri[8] * I2 - ms * o[1] * o[2];
    ci[6] = ms * o[0]; ci[7] = ms * o[1]; ci[8] = ms * o[2]; ci[9] = ms;
  }
  B2_UNROLL1
  for (int j = B2_LANE; j < nj; j += B2_NLANE) {
    int b = jnt_bodyid[j], da = jnt_dofadr[j], r = body_root[b];
    real off[3] = {scom[3 * r] - xanchor[3 * j], scom[3 * r + 1] - xanchor[3 * j + 1], scom[3 * r + 2] - xanchor[3 * j + 2]};
    int t = jnt_type[j];
    if (t == B2_JNT_FREE) {
      for (int a = 0; a < 3; ++a) {
        real* cd = cdof + 6 * (da + a);
        cd[0] = cd[1] = cd[2] = cd[3] = cd[4] = cd[5] = 0; cd[3 + a] = 1;
        real ax[3] = {xmat[9 * b + a], xmat[9 * b + 3 + a], xmat[9 * b + 6 + a]};
        real* cr = cdof + 6 * (da + 3 + a);
        cr[0] = ax[0]; cr[1] = ax[1]; cr[2] = ax[2];
        cross3(cr + 3, ax, off);
      }
    } else if (t == B2_JNT_HINGE) {
      real* cd = cdof + 6 * da;
      cd[0] = xaxis[3 * j]; cd[1] = xaxis[3 * j + 1]; cd[2] = xaxis[3 * j + 2];
      cross3(cd + 3, xaxis + 3 * j, off);
    } else {  // slide
      real* cd = cdof + 6 * da;
      cd[0] = cd[1] = cd[2] = 0; cd[3] = xaxis[3 * j]; cd[4] = xaxis[3 * j + 1]; cd[5] = xaxis[3 * j + 2];
    }
  }
  B2_WSYNC();
  {
    real* g = d.df[B2_DF_CINERT] + (size_t)w * nb * 10;
    B2_UNROLL1
    for (int i = B2_LANE; i < 10 * nb; i += B2_NLANE) g[i] = cinert[i];
    g = d.df[B2_DF_CDOF] + (size_t)w * nv * 6;
    B2_UNROLL1
    for (int i = B2_LANE; i < 6 * nv; i += B2_NLANE) g[i] = cdof[i];
  }
  B2_SPH(1)
  // ---- collision ---------------------------------------------------------------  // @region collision
  {
    const int* pair_geom = MI(B2_MI_PAIR_GEOM);
    const int* geom_type = MI(B2_MI_GEOM_TYPE);
    const int* geom_condim = MI(B2_MI_GEOM_CONDIM);
    const int* geom_priority = MI(B2_MI_GEOM_PRIORITY);
    const real* geom_size = MF(B2_MF_GEOM_SIZE);
    const real* geom_rbound = MF(B2_MF_GEOM_RBOUND);
    const real* geom_margin = MF(B2_MF_GEOM_MARGIN);
    const real* geom_gap = MF(B2_MF_GEOM_GAP);
    const real* geom_friction = MF(B2_MF_GEOM_FRICTION);
    const real* geom_solref = MF(B2_MF_GEOM_SOLREF);
    const real* geom_solimp = MF(B2_MF_GEOM_SOLIMP);
    const real* geom_solmix = MF(B2_MF_GEOM_SOLMIX);
    const real* gx = d.df[B2_DF_GEOM_XPOS] + (size_t)w * m.ngeom * 3;
    const real* gm = d.df[B2_DF_GEOM_XMAT] + (size_t)w * m.ngeom * 9;
    B2_WSYNC();  // geom poses written above by other lanes
    // broadphase: ordered compaction of surviving pair ids
    int ncand = 0;
    B2_UNROLL1
    for (int base = 0; base < m.npair; base += B2_NLANE) {
      int k = base + B2_LANE;
      bool hit = false;
      if (k < m.npair) {
        int g1 = pair_geom[2 * k], g2 = pair_geom[2 * k + 1];
        real margin = fmax(geom_margin[g1], geom_margin[g2]);
        real dx[3] = {gx[3 * g2] - gx[3 * g1], gx[3 * g2 + 1] - gx[3 * g1 + 1], gx[3 * g2 + 2] - gx[3 * g1 + 2]};
        if (geom_type[g1] == B2_GEOM_PLANE) {
          real n[3] = {gm[9 * g1 + 2], gm[9 * g1 + 5], gm[9 * g1 + 8]};
          hit = !(dot3(dx, n) > margin + geom_rbound[g2]);
        } else {
          real bound = geom_rbound[g1] + geom_rbound[g2] + margin;
          hit = !(dot3(dx, dx) > bound * bound);
        }
      }
      unsigned bal = b2_ballot(hit);
      if (hit) {
        int slot = ncand + b2_popc(bal & b2_lanemask_lt());
        if (slot < ncand_max) cand[slot] = k;
      }
      ncand += b2_popc(bal);
    }
    // geom field (static terrain boxes): every query geom looks up the grid cells under its bounding sphere; a box met
    // in several cells is taken from the first cell the two cell ranges share.  Two passes (count, then write) keep the
    // candidate order deterministic: by query geom, cell, field index.
    const int* fq_geom = MI(B2_MI_FQ_GEOM);
    const int* f_geom = MI(B2_MI_FIELD_GEOM);
    const real* sg_xpos = MF(B2_MF_SGEOM_XPOS);
    const real* sg_xmat = MF(B2_MF_SGEOM_XMAT);
    if (m.nfield > 0) {
      const int* f_cell0 = MI(B2_MI_FIELD_CELL0);
      const int* f_adr = MI(B2_MI_FIELD_CELL_ADR);
      const int* f_item = MI(B2_MI_FIELD_CELL_ITEM);
      const real* f_aabb = MF(B2_MF_FIELD_AABB);
      const real* grid = MF(B2_MF_FIELD_GRID);
      B2_UNROLL1
      for (int base = 0; base < m.nfq; base += B2_NLANE) {
        int qi = base + B2_LANE;
        int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
        real c[3] = {0, 0, 0}, rb = 0;
        if (qi < m.nfq) {
          int g = fq_geom[qi];
          c[0] = gx[3 * g]; c[1] = gx[3 * g + 1]; c[2] = gx[3 * g + 2];
          rb = geom_rbound[g] + geom_margin[g];
          cx0 = (int)floor((c[0] - rb - grid[0]) * grid[3]); cx1 = (int)floor((c[0] + rb - grid[0]) * grid[3]);
          cy0 = (int)floor((c[1] - rb - grid[1]) * grid[3]); cy1 = (int)floor((c[1] + rb - grid[1]) * grid[3]);
          if (cx0 < 0) cx0 = 0;
          if (cy0 < 0) cy0 = 0;
          if (cx1 > m.field_nx - 1) cx1 = m.field_nx - 1;
          if (cy1 > m.field_ny - 1) cy1 = m.field_ny - 1;
        }
        int start = 0, total = 0;
        B2_UNROLL1
        for (int pass = 0; pass < 2; ++pass) {
          int k = 0;
          B2_UNROLL1
          for (int cx = cx0; cx <= cx1; ++cx) {
            B2_UNROLL1
            for (int cy = cy0; cy <= cy1; ++cy) {
              int cell = cx * m.field_ny + cy;
              B2_UNROLL1
              for (int it = f_adr[cell]; it < f_adr[cell + 1]; ++it) {
                int fi = f_item[it];
                int fx0 = f_cell0[2 * fi], fy0 = f_cell0[2 * fi + 1];
                if (cx != (cx0 > fx0 ? cx0 : fx0) || cy != (cy0 > fy0 ? cy0 : fy0)) continue;
                const real* a = f_aabb + 6 * fi;
                real ex = fabs(c[0] - a[0]) - a[3], ey = fabs(c[1] - a[1]) - a[4], ez = fabs(c[2] - a[2]) - a[5];
                ex = ex > 0 ? ex : 0; ey = ey > 0 ? ey : 0; ez = ez > 0 ? ez : 0;
                if (ex * ex + ey * ey + ez * ez > rb * rb) continue;
                if (pass == 1) {
                  int slot = start + k;
                  if (slot < ncand_max) cand[slot] = m.npair + qi * m.nfield + fi;
                }
                ++k;
              }
            }
          }
          if (pass == 0) {
            int incl = k;
#ifndef B2_HOSTEMU
            B2_UNROLL1
            for (int o = 1; o < 32; o <<= 1) {
              int t = __shfl_up_sync(0xffffffffu, incl, o);
              if (B2_LANE >= o) incl += t;
            }
            total = __shfl_sync(0xffffffffu, incl, 31);
#else
            total = incl;
#endif
            start = ncand + incl - k;
          }
        }
        ncand += total;
      }
    }
    if (ncand > ncand_max) ncand = ncand_max;
    B2_WSYNC();
    // narrowphase: ordered compaction of contacts into the world's contact rows
    const int nconmax = d.nconmax;
    real* c_dist = d.df[B2_DF_CONTACT_DIST] + (size_t)w * nconmax;
    real* c_pos = d.df[B2_DF_CONTACT_POS] + (size_t)w * nconmax * 3;
    real* c_frame = d.df[B2_DF_CONTACT_FRAME] + (size_t)w * nconmax * 9;
    real* c_inc = d.df[B2_DF_CONTACT_INCLUDEMARGIN] + (size_t)w * nconmax;
    real* c_fri = d.df[B2_DF_CONTACT_FRICTION] + (size_t)w * nconmax * 5;
    real* c_solref = d.df[B2_DF_CONTACT_SOLREF] + (size_t)w * nconmax * 2;
    real* c_solimp = d.df[B2_DF_CONTACT_SOLIMP] + (size_t)w * nconmax * 5;
    int* c_geom = d.di[B2_DI_CONTACT_GEOM] + (size_t)w * nconmax * 2;
    int* c_dim = d.di[B2_DI_CONTACT_DIM] + (size_t)w * nconmax;
    int ncon = 0, ndrop = 0;
    B2_UNROLL1
    for (int base = 0; base < ncand; base += B2_NLANE) {
      int ci = base + B2_LANE;
      LaneContacts lc;
      lc.n = 0;
      lc.nb_from = B2_MAXCON_PER_PAIR;
      int g1 = 0, g2 = 0;
      real margin = 0, gap = 0;
      if (ci < ncand) {
        int k = cand[ci];
        const real *x2, *R2;
        if (k < m.npair) {
          g1 = pair_geom[2 * k]; g2 = pair_geom[2 * k + 1];
          x2 = gx + 3 * g2; R2 = gm + 9 * g2;
        } else {   // (query geom, field box): the box pose is a model constant shared by all worlds
          k -= m.npair;
          g1 = fq_geom[k / m.nfield]; g2 = f_geom[k % m.nfield];
          x2 = sg_xpos + 3 * g2; R2 = sg_xmat + 9 * g2;
        }
        margin = fmax(geom_margin[g1], geom_margin[g2]);
        gap = fmax(geom_gap[g1], geom_gap[g2]);
        narrowphase(lc, geom_type[g1], geom_type[g2], gx + 3 * g1, gm + 9 * g1, geom_size + 3 * g1, x2, R2,
                    geom_size + 3 * g2, margin);
      }
      // exclusive prefix sum of lc.n over lanes
      int incl = lc.n;
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
      int start = ncon + incl - lc.n;
      if (lc.n > 0) {
        // contact parameter mixing (priority, solmix, max friction)
        int condim;
        real solref[2], solimp[5], fri[3];
        int p1 = geom_priority[g1], p2 = geom_priority[g2];
        if (p1 != p2) {
          int g = p1 > p2 ? g1 : g2;
          condim = geom_condim[g];
          solref[0] = geom_solref[2 * g]; solref[1] = geom_solref[2 * g + 1];
          for (int e = 0; e < 5; ++e) solimp[e] = geom_solimp[5 * g + e];
          for (int e = 0; e < 3; ++e) fri[e] = geom_friction[3 * g + e];
        } else {
          condim = geom_condim[g1] > geom_condim[g2] ? geom_condim[g1] : geom_condim[g2];
          real s1 = geom_solmix[g1], s2 = geom_solmix[g2], mix;
          if (s1 >= B2_MINVAL && s2 >= B2_MINVAL) mix = s1 / (s1 + s2);
          else if (s1 < B2_MINVAL && s2 < B2_MINVAL) mix = (real)0.5;
          else if (s1 < B2_MINVAL) mix = 0;
          else mix = 1;
          if (geom_solref[2 * g1] > 0 && geom_solref[2 * g2] > 0) {
            solref[0] = mix * geom_solref[2 * g1] + (1 - mix) * geom_solref[2 * g2];
            solref[1] = mix * geom_solref[2 * g1 + 1] + (1 - mix) * geom_solref[2 * g2 + 1];
          } else {
            solref[0] = fmin(geom_solref[2 * g1], geom_solref[2 * g2]);
            solref[1] = fmin(geom_solref[2 * g1 + 1], geom_solref[2 * g2 + 1]);
          }
          for (int e = 0; e < 5; ++e) solimp[e] = mix * geom_solimp[5 * g1 + e] + (1 - mix) * geom_solimp[5 * g2 + e];
          for (int e = 0; e < 3; ++e) fri[e] = fmax(geom_friction[3 * g1 + e], geom_friction[3 * g2 + e]);
        }
        real F[9];
        make_frame(F, lc.normal, lc.tangent);
        B2_UNROLL1
        for (int q = 0; q < lc.n; ++q) {
          int slot = start + q;
          if (slot >= nconmax) break;
          if (q == lc.nb_from) make_frame(F, lc.normal_b, lc.tangent);
          c_dist[slot] = lc.dist[q];
          c_pos[3 * slot] = lc.pos[q][0]; c_pos[3 * slot + 1] = lc.pos[q][1]; c_pos[3 * slot + 2] = lc.pos[q][2];
          for (int e = 0; e < 9; ++e) c_frame[9 * slot + e] = F[e];
          c_inc[slot] = margin - gap;
          c_fri[5 * slot] = fri[0]; c_fri[5 * slot + 1] = fri[0]; c_fri[5 * slot + 2] = fri[1];
          c_fri[5 * slot + 3] = fri[2]; c_fri[5 * slot + 4] = fri[2];
          c_solref[2 * slot] = solref[0]; c_solref[2 * slot + 1] = solref[1];
          for (int e = 0; e < 5; ++e) c_solimp[5 * slot + e] = solimp[e];
          c_geom[2 * slot] = g1; c_geom[2 * slot + 1] = g2;
          c_dim[slot] = condim;
        }
      }
      ncon += total;
    }
    if (ncon > nconmax) { ndrop = ncon - nconmax; ncon = nconmax; }
    if (B2_LANE == 0) {
      d.di[B2_DI_NCON][w] = ncon;
      d.di[B2_DI_NCON_DROPPED][w] = ndrop;
    }
  }

  B2_SPH(3)
  // ---- com_vel: cvel, cdof_dot (forward over levels) ----------------------------  // @region com_vel
  B2_UNROLL1
  for (int lev = 0; lev < m.nlevel; ++lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      real v[6] = {0, 0, 0, 0, 0, 0};
      if (b > 0) {
        int p = body_parent[b];
        for (int c = 0; c < 6; ++c) v[c] = cvel[6 * p + c];
        B2_UNROLL1
        for (int jj = 0; jj < body_jntnum[b]; ++jj) {
          int j = body_jntadr[b] + jj, da = jnt_dofadr[j];
          if (jnt_type[j] == B2_JNT_FREE) {
            for (int a = 0; a < 3; ++a) {
              for (int c = 0; c < 6; ++c) { v[c] += cdof[6 * (da + a) + c] * qvel[da + a]; cdofdot[6 * (da + a) + c] = 0; }
            }
            for (int a = 3; a < 6; ++a) cross_motion(cdofdot + 6 * (da + a), v, cdof + 6 * (da + a));
            for (int a = 3; a < 6; ++a)
              for (int c = 0; c < 6; ++c) v[c] += cdof[6 * (da + a) + c] * qvel[da + a];
          } else {
            cross_motion(cdofdot + 6 * da, v, cdof + 6 * da);
            for (int c = 0; c < 6; ++c) v[c] += cdof[6 * da + c] * qvel[da];
          }
        }
      }
      for (int c = 0; c < 6; ++c) cvel[6 * b + c] = v[c];
    }
    B2_WSYNC();
  }
  {
    real* g = d.df[B2_DF_CVEL] + (size_t)w * nb * 6;
    B2_UNROLL1
    for (int i = B2_LANE; i < 6 * nb; i += B2_NLANE) g[i] = cvel[i];
    g = d.df[B2_DF_CDOF_DOT] + (size_t)w * nv * 6;
    B2_UNROLL1
    for (int i = B2_LANE; i < 6 * nv; i += B2_NLANE) g[i] = cdofdot[i];
  }

  }   // !light
  // ---- actuation ------------------------------------------------------------------  // @region actuation
  {
    const int* trnid = MI(B2_MI_ACTUATOR_TRNID);
    const int* biastype = MI(B2_MI_ACTUATOR_BIASTYPE);
    const int* flim = MI(B2_MI_ACTUATOR_FORCELIMITED);
    const int* clim = MI(B2_MI_ACTUATOR_CTRLLIMITED);
    const real* gainprm = MF(B2_MF_ACTUATOR_GAINPRM);
    const real* biasprm = MF(B2_MF_ACTUATOR_BIASPRM);
    const real* gear = MF(B2_MF_ACTUATOR_GEAR);
    const real* frange = MF(B2_MF_ACTUATOR_FORCERANGE);
    const real* crange = MF(B2_MF_ACTUATOR_CTRLRANGE);
    const real* ctrl = d.df[B2_DF_CTRL] + (size_t)w * nu;
    real* g_af = d.df[B2_DF_ACTUATOR_FORCE] + (size_t)w * nu;
    real* g_al = d.df[B2_DF_ACTUATOR_LENGTH] + (size_t)w * nu;
    real* g_av = d.df[B2_DF_ACTUATOR_VELOCITY] + (size_t)w * nu;
    B2_UNROLL1
    for (int i = B2_LANE; i < nu; i += B2_NLANE) {
      int j = trnid[2 * i];
      real gr = gear[6 * i];
      real len = gr * qpos[jnt_qposadr[j]];
      real vel = gr * qvel[jnt_dofadr[j]];
      real c = ctrl[i];
      if (clim[i]) c = clampr(c, crange[2 * i], crange[2 * i + 1]);
      real force = gainprm[10 * i] * c;
      if (biastype[i] == 1) force += biasprm[10 * i] + biasprm[10 * i + 1] * len + biasprm[10 * i + 2] * vel;
      if (flim[i]) force = clampr(force, frange[2 * i], frange[2 * i + 1]);
      aforce[i] = force;
      g_af[i] = force; g_al[i] = len; g_av[i] = vel;
    }
    B2_WSYNC();
  }

  B2_SPH(4)
  // ---- RNE bias ------------------------------------------------------------------------  // @region rne
  if (!light) {
  const real* grav = MF(B2_MF_OPT_GRAVITY);
  B2_UNROLL1
  for (int lev = 0; lev < m.nlevel; ++lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      real a[6];
      if (b == 0) {
        a[0] = a[1] = a[2] = 0; a[3] = -grav[0]; a[4] = -grav[1]; a[5] = -grav[2];
        for (int c = 0; c < 6; ++c) { cacc[c] = a[c]; cfrc[c] = 0; }
      } else {
        int p = body_parent[b];
        for (int c = 0; c < 6; ++c) a[c] = cacc[6 * p + c];
        B2_UNROLL1
        for (int e = 0; e < body_dofnum[b]; ++e) {
          int dd = body_dofadr[b] + e;
          for (int c = 0; c < 6; ++c) a[c] += cdofdot[6 * dd + c] * qvel[dd];
        }
        for (int c = 0; c < 6; ++c) cacc[6 * b + c] = a[c];
        real Ia[6], Iv[6], cf[6];
        inert_mul(Ia, cinert + 10 * b, a);
        inert_mul(Iv, cinert + 10 * b, cvel + 6 * b);
        cross_force(cf, cvel + 6 * b, Iv);
        for (int c = 0; c < 6; ++c) cfrc[6 * b + c] = Ia[c] + cf[c];
      }
    }
    B2_WSYNC();
  }
  B2_UNROLL1
  for (int lev = m.nlevel - 1; lev >= 1; --lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      B2_UNROLL1
      for (int c = child_adr[b]; c < child_adr[b + 1]; ++c) {
        int ch = child_body[c];
        for (int e = 0; e < 6; ++e) cfrc[6 * b + e] += cfrc[6 * ch + e];
      }
    }
    B2_WSYNC();
  }
  }

  B2_SPH(5)
  // ---- passive + actuator + applied forces -> qfrc_smooth ------------------------------  // @region qfrc_smooth
  {
    const real* damping = MF(B2_MF_DOF_DAMPING);
    const real* stiffness = MF(B2_MF_JNT_STIFFNESS);
    const real* qspring = MF(B2_MF_QPOS_SPRING);
    const int* dof_jntid = MI(B2_MI_DOF_JNTID);
    const int* trnid = MI(B2_MI_ACTUATOR_TRNID);
    const int* dof_actid = MI(B2_MI_DOF_ACTID);
    const real* gear = MF(B2_MF_ACTUATOR_GEAR);
    const real* qfrc_applied = d.df[B2_DF_QFRC_APPLIED] + (size_t)w * nv;
    const real* xfrc = d.df[B2_DF_XFRC_APPLIED] + (size_t)w * nb * 6;
    const int* dofmask = MI(B2_MI_BODY_DOFMASK);
    // any external wrench at all? (common case: none; the loads were issued at the top of the kernel)
    const int anyx = b2_wor_i(anyx_lane);
    real* g_bias = d.df[B2_DF_QFRC_BIAS] + (size_t)w * nv;
    real* g_pas = d.df[B2_DF_QFRC_PASSIVE] + (size_t)w * nv;
    real* g_act = d.df[B2_DF_QFRC_ACTUATOR] + (size_t)w * nv;
    real* g_smooth = d.df[B2_DF_QFRC_SMOOTH] + (size_t)w * nv;
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) {
      const real* ci = cdof + 6 * i;
      const real* cf = cfrc + 6 * dof_bodyid[i];
      real bias;
      if (light) bias = g_bias[i];
      else bias = ci[0] * cf[0] + ci[1] * cf[1] + ci[2] * cf[2] + ci[3] * cf[3] + ci[4] * cf[4] + ci[5] * cf[5];
      real pas = -damping[i] * qvel[i];
      int j = dof_jntid[i];
      int jt = jnt_type[j];
      if ((jt == B2_JNT_HINGE || jt == B2_JNT_SLIDE) && stiffness[j] != 0) {
        int qa = jnt_qposadr[j];
        pas -= stiffness[j] * (qpos[qa] - qspring[qa]);
      }
      real act = 0;
      {
        int aid = dof_actid[i];
        int a_lo = aid >= 0 ? aid : 0, a_hi = aid >= 0 ? aid + 1 : (aid == -2 ? nu : 0);
        B2_UNROLL1
        for (int a = a_lo; a < a_hi; ++a)
          if (jnt_dofadr[trnid[2 * a]] == i) act += gear[6 * a] * aforce[a];
      }
      real fs = pas - bias + act + qfrc_applied[i];
      if (anyx) {
        B2_UNROLL1
        for (int b = 1; b < nb; ++b) {
          unsigned lo = (unsigned)dofmask[2 * b], hi = (unsigned)dofmask[2 * b + 1];
          bool in = i < 32 ? ((lo >> i) & 1u) : ((hi >> (i - 32)) & 1u);
          if (!in) continue;
          const real* x = xfrc + 6 * b;
          if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
          int r = body_root[b];
          real off[3] = {xipos[3 * b] - scom[3 * r], xipos[3 * b + 1] - scom[3 * r + 1], xipos[3 * b + 2] - scom[3 * r + 2]};
          real t[3];
          cross3(t, off, x);
          t[0] += x[3]; t[1] += x[4]; t[2] += x[5];
          fs += ci[3] * x[0] + ci[4] * x[1] + ci[5] * x[2] + ci[0] * t[0] + ci[1] * t[1] + ci[2] * t[2];
        }
      }
      g_bias[i] = bias; g_pas[i] = pas; g_act[i] = act; g_smooth[i] = fs;
      f[i] = fs;
    }
    B2_WSYNC();
  }
  if (!light) {
  // ---- composite rigid body inertias (accumulated in place: cinert is in HBM and dead here), dense qM ----  // @region crb_qM
  B2_UNROLL1
  for (int lev = m.nlevel - 1; lev >= 1; --lev) {
    B2_UNROLL1
    for (int k = level_adr[lev] + B2_LANE; k < level_adr[lev + 1]; k += B2_NLANE) {
      int b = level_body[k];
      B2_UNROLL1
      for (int c = child_adr[b]; c < child_adr[b + 1]; ++c) {
        int ch = child_body[c];
        for (int e = 0; e < 10; ++e) crb[10 * b + e] += crb[10 * ch + e];
      }
    }
    B2_WSYNC();
  }
  const int npack = b2_poff(nv);   // M is kept as a packed lower triangle (b2_common.cuh, b2_poff)
  B2_UNROLL1
  for (int i = B2_LANE; i < npack; i += B2_NLANE) M[i] = 0;
  B2_WSYNC();
  const real* armature = MF(B2_MF_DOF_ARMATURE);
  B2_UNROLL1
  for (int i = B2_LANE; i < nv; i += B2_NLANE) {
    real buf[6];
    inert_mul(buf, crb + 10 * dof_bodyid[i], cdof + 6 * i);
    const real* ci = cdof + 6 * i;
    real* Mi = M + b2_poff(i);
    Mi[i] = armature[i] + ci[0] * buf[0] + ci[1] * buf[1] + ci[2] * buf[2] + ci[3] * buf[3] + ci[4] * buf[4] + ci[5] * buf[5];
    B2_UNROLL1
    for (int j = dof_parentid[i]; j >= 0; j = dof_parentid[j]) {   // ancestors: j < i
      const real* cj = cdof + 6 * j;
      Mi[j] = cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
    }
  }
  B2_WSYNC();
  {  // dense symmetric copy for the solver and the integrator
    real* g = d.df[B2_DF_QM] + (size_t)w * nv * ld;
    int i = B2_LANE / ld, j = B2_LANE - i * ld;   // (row, column) of idx, advanced without a division per element
    B2_UNROLL1
    for (int idx = B2_LANE; idx < nv * ld; idx += B2_NLANE) {
      g[idx] = j >= nv ? (real)0 : (j <= i ? M[b2_poff(i) + j] : M[b2_poff(j) + i]);
      j += B2_NLANE;
      while (j >= ld) { j -= ld; ++i; }
    }
    B2_WSYNC();   // the factorisation below overwrites M in place
  }

  B2_SPH(2)
  } else {   // memo: the mass matrix of this state is in HBM
    const real* g = d.df[B2_DF_QM] + (size_t)w * nv * ld;
    const int npack = b2_poff(nv);
    B2_UNROLL1
    for (int i = B2_LANE; i < npack; i += B2_NLANE) M[i] = 0;
    B2_WSYNC();
    int i = B2_LANE / ld, j = B2_LANE - i * ld;
    B2_UNROLL1
    for (int idx = B2_LANE; idx < nv * ld; idx += B2_NLANE) {
      if (j <= i) M[b2_poff(i) + j] = g[idx];
      j += B2_NLANE;
      while (j >= ld) { j -= ld; ++i; }
    }
    B2_WSYNC();
  }

  B2_SPH(6)
  // ---- qacc_smooth = M^-1 qfrc_smooth (M is in HBM; factor in place) ---------------------  // @region factor_solve
  spd_factor_solve_packed(M, nv, f, sm + L.invd);
  {
    real* g = d.df[B2_DF_QACC_SMOOTH] + (size_t)w * nv;
    B2_UNROLL1
    for (int i = B2_LANE; i < nv; i += B2_NLANE) g[i] = f[i];
  }
  B2_SPH(7)
  B2_SPH_END()
}
