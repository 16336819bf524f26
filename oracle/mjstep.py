"""ORACLE (test infrastructure, never on the product path).

Float64 numpy restatement of MuJoCo's `mj_step` / `mj_forward` pipeline for the
feature subset the BASELINE configs feed to the engine.
One world at a time, plain Python loops: use it at small sizes only.

What it restates.  The reference delegates all physics arithmetic to
`mujoco_warp.{step,forward,reset_data}` at git f2f795796fc433adf8e235f01fae3747585ae5db
(call sites: /root/reference/src/mjlab/sim/sim.py:184,187,190,265,273,286;
pin: pyproject.toml:106, uv.lock:1378-1380), which itself restates MuJoCo C
3.4.1.dev843752024 (uv.lock:1334-1335).  Neither dependency is vendored under
/root/reference nor installable here, so this file follows MuJoCo's published
algorithm (the "Computation" chapter of the MuJoCo documentation; field names
from typings/mujoco/_structs.pyi:113-867,916).

Friction cones: pyramidal (the BASELINE configs) and elliptic (`MujocoCfg.cone`, src/mjlab/sim/sim.py:51-56); solvers:
Newton and CG; constraint rows: friction loss, joint limits, contacts.

PARITY UNPINNED: the reference's own tests hold no golden vectors for this path
 and the reference cannot be executed in this container or on the
GPU box, so this oracle is pinned only by analytic invariants
(tests/test_oracle_invariants.py) and by the reference tests' qualitative
assertions mirrored in tests/.

Pipeline order follows mj_forward + integrator:
  position: kinematics -> com_pos -> crb (qM) -> collision -> make_constraint -> transmission
  velocity: com_vel -> passive -> rne (qfrc_bias)
  actuation -> qfrc_smooth -> qacc_smooth -> constraint solve (Newton)
  sensors (pos/vel/acc) -> integrate (Euler | implicitfast)
"""

from __future__ import annotations

import math

import numpy as np

# enums (same numeric values as mjlab_b200.mjcf; restated so the oracle has no
# import dependency on the product package)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
INT_EULER, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST = range(4)
OBJ_BODY, OBJ_XBODY, OBJ_JOINT, OBJ_GEOM, OBJ_SITE = 1, 2, 3, 5, 6
SENS_ACCELEROMETER, SENS_VELOCIMETER, SENS_GYRO = 1, 2, 3
SENS_JOINTPOS, SENS_JOINTVEL = 8, 9
SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_FRAMELINVEL, SENS_FRAMEANGVEL = 25, 26, 30, 31
SENS_SUBTREECOM, SENS_SUBTREELINVEL, SENS_SUBTREEANGMOM = 34, 35, 36
SENS_CONTACT = 42
SOL_PGS, SOL_CG, SOL_NEWTON = range(3)
MINVAL = 1e-15
MINIMP, MAXIMP = 0.0001, 0.9999


# ---------------------------------------------------------------------------
# math helpers
# ---------------------------------------------------------------------------
def quat_mul(a, b):
  return np.array([
    a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
    a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
    a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
    a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
  ])


def quat_to_mat(q):
  w, x, y, z = q
  return np.array([
    [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
    [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
    [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
  ])


def mat_to_quat(R):
  tr = R[0, 0] + R[1, 1] + R[2, 2]
  if tr > 0:
    s = math.sqrt(tr + 1.0) * 2
    q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
  elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
    s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
    q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
  elif R[1, 1] > R[2, 2]:
    s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
    q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
  else:
    s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
    q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
  q = np.array(q)
  return q / np.linalg.norm(q)


def axis_angle_quat(axis, angle):
  s = math.sin(angle * 0.5)
  return np.array([math.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s])


def quat_integrate(q, w, h):
  """q <- normalize(q * exp(h*w/2)), w expressed in the local (body) frame."""
  n = np.linalg.norm(w)
  if n < MINVAL:
    return q / np.linalg.norm(q)
  q = quat_mul(q, axis_angle_quat(w / n, h * n))
  return q / np.linalg.norm(q)


def cross_motion(v, s):
  return np.concatenate([np.cross(v[:3], s[:3]), np.cross(v[:3], s[3:]) + np.cross(v[3:], s[:3])])


def cross_force(v, f):
  return np.concatenate([np.cross(v[:3], f[:3]) + np.cross(v[3:], f[3:]), np.cross(v[:3], f[3:])])


class Inert:
  """Spatial inertia about a reference point: (I 3x3 about ref, h = m*r, m)."""

  __slots__ = ("I", "h", "m")

  def __init__(self, I, h, m):
    self.I, self.h, self.m = I, h, m

  def mul(self, v):
    w, u = v[:3], v[3:]
    return np.concatenate([self.I @ w + np.cross(self.h, u), self.m * u - np.cross(self.h, w)])

  def add(self, o):
    return Inert(self.I + o.I, self.h + o.h, self.m + o.m)


def make_frame(normal, tangent=None):
  """Contact frame rows [normal, t1, t2] (normal-first, right-handed)."""
  n = normal / np.linalg.norm(normal)
  t = np.zeros(3) if tangent is None else np.array(tangent, float)
  if np.linalg.norm(t) < 0.5:
    t = np.zeros(3)
    if -0.5 < n[1] < 0.5:
      t[1] = 1.0
    else:
      t[2] = 1.0
  t = t - n * np.dot(n, t)
  t = t / np.linalg.norm(t)
  return np.stack([n, t, np.cross(n, t)])


# ---------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------
class Data:
  def __init__(self, m):
    self.qpos = np.array(m.qpos0, float)
    self.qvel = np.zeros(m.nv)
    self.ctrl = np.zeros(m.nu)
    self.qacc_warmstart = np.zeros(m.nv)
    self.qfrc_applied = np.zeros(m.nv)
    self.xfrc_applied = np.zeros((m.nbody, 6))
    self.mocap_pos = np.zeros((max(m.nmocap, 1), 3))
    self.mocap_quat = np.tile(np.array([1.0, 0, 0, 0]), (max(m.nmocap, 1), 1))
    self.time = 0.0
    self.qacc = np.zeros(m.nv)
    self.sensordata = np.zeros(m.nsensordata)
    self.contacts = []
    self.solver_niter = 0


# ---------------------------------------------------------------------------
# position stage
# ---------------------------------------------------------------------------
def kinematics(m, d):
  nb = m.nbody
  d.xpos = np.zeros((nb, 3))
  d.xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
  d.xmat = np.tile(np.eye(3), (nb, 1, 1))
  d.xanchor = np.zeros((m.njnt, 3))
  d.xaxis = np.zeros((m.njnt, 3))
  # normalise free-joint quaternions in qpos (mj_kinematics does this in place)
  for j in range(m.njnt):
    if m.jnt_type[j] == JNT_FREE:
      a = m.jnt_qposadr[j] + 3
      d.qpos[a:a + 4] /= np.linalg.norm(d.qpos[a:a + 4])
  for i in range(1, nb):
    p = m.body_parentid[i]
    if m.body_mocapid[i] >= 0:
      pos = d.mocap_pos[m.body_mocapid[i]].copy()
      quat = d.mocap_quat[m.body_mocapid[i]] / np.linalg.norm(d.mocap_quat[m.body_mocapid[i]])
    else:
      pos = d.xpos[p] + d.xmat[p] @ m.body_pos[i]
      quat = quat_mul(d.xquat[p], m.body_quat[i])
    for k in range(m.body_jntnum[i]):
      j = m.body_jntadr[i] + k
      qa = m.jnt_qposadr[j]
      t = m.jnt_type[j]
      if t == JNT_FREE:
        pos = d.qpos[qa:qa + 3].copy()
        quat = d.qpos[qa + 3:qa + 7].copy()
        d.xanchor[j] = pos
        d.xaxis[j] = quat_to_mat(quat)[:, 2]
      else:
        R = quat_to_mat(quat)
        d.xanchor[j] = pos + R @ m.jnt_pos[j]
        d.xaxis[j] = R @ m.jnt_axis[j]
        if t == JNT_HINGE:
          quat = quat_mul(quat, axis_angle_quat(m.jnt_axis[j], d.qpos[qa] - m.qpos0[qa]))
          pos = d.xanchor[j] - quat_to_mat(quat) @ m.jnt_pos[j]
        elif t == JNT_SLIDE:
          pos = pos + d.xaxis[j] * (d.qpos[qa] - m.qpos0[qa])
        else:
          raise NotImplementedError
    quat = quat / np.linalg.norm(quat)
    d.xpos[i], d.xquat[i], d.xmat[i] = pos, quat, quat_to_mat(quat)
  d.xipos = np.zeros((nb, 3))
  d.ximat = np.zeros((nb, 3, 3))
  for i in range(nb):
    d.xipos[i] = d.xpos[i] + d.xmat[i] @ m.body_ipos[i]
    d.ximat[i] = quat_to_mat(quat_mul(d.xquat[i], m.body_iquat[i]))
  # geoms of bodies welded to the world never move: evaluate them (through the same body FK) once per Data
  cache = getattr(d, "_static_geoms", None)
  if cache is None:
    weld0 = np.array([m.body_weldid[m.geom_bodyid[g]] == 0 and m.body_mocapid[m.geom_bodyid[g]] < 0 for g in range(m.ngeom)], bool)
    d.geom_xpos = np.zeros((m.ngeom, 3))
    d.geom_xmat = np.zeros((m.ngeom, 3, 3))
    todo = range(m.ngeom)
  else:
    weld0 = cache
    todo = np.flatnonzero(~weld0)
  for g in todo:
    b = m.geom_bodyid[g]
    d.geom_xpos[g] = d.xpos[b] + d.xmat[b] @ m.geom_pos[g]
    d.geom_xmat[g] = quat_to_mat(quat_mul(d.xquat[b], m.geom_quat[g]))
  d._static_geoms = weld0
  d.site_xpos = np.zeros((m.nsite, 3))
  d.site_xmat = np.zeros((m.nsite, 3, 3))
  for s in range(m.nsite):
    b = m.site_bodyid[s]
    d.site_xpos[s] = d.xpos[b] + d.xmat[b] @ m.site_pos[s]
    d.site_xmat[s] = quat_to_mat(quat_mul(d.xquat[b], m.site_quat[s]))


def com_pos(m, d):
  nb = m.nbody
  mass = getattr(d, "body_mass", m.body_mass)
  acc = mass[:, None] * d.xipos
  sub = mass.astype(float).copy()
  for i in range(nb - 1, 0, -1):
    acc[m.body_parentid[i]] += acc[i]
    sub[m.body_parentid[i]] += sub[i]
  d.subtree_mass = sub
  d.subtree_com = np.where(sub[:, None] > MINVAL, acc / np.maximum(sub[:, None], MINVAL), d.xipos)
  d.subtree_com[0] = acc[0] / sub[0] if sub[0] > MINVAL else np.zeros(3)
  d.cinert = []
  for i in range(nb):
    r = d.xipos[i] - d.subtree_com[m.body_rootid[i]]
    I = d.ximat[i] @ np.diag(m.body_inertia[i]) @ d.ximat[i].T + mass[i] * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
    d.cinert.append(Inert(I, mass[i] * r, mass[i]))
  d.cdof = np.zeros((m.nv, 6))
  for j in range(m.njnt):
    b = m.jnt_bodyid[j]
    da = m.jnt_dofadr[j]
    off = d.subtree_com[m.body_rootid[b]] - d.xanchor[j]
    t = m.jnt_type[j]
    if t == JNT_FREE:
      for a in range(3):
        d.cdof[da + a, 3 + a] = 1.0
        ax = d.xmat[b][:, a]
        d.cdof[da + 3 + a, :3] = ax
        d.cdof[da + 3 + a, 3:] = np.cross(ax, off)
    elif t == JNT_HINGE:
      d.cdof[da, :3] = d.xaxis[j]
      d.cdof[da, 3:] = np.cross(d.xaxis[j], off)
    elif t == JNT_SLIDE:
      d.cdof[da, 3:] = d.xaxis[j]


def crb(m, d):
  nb, nv = m.nbody, m.nv
  comp = [Inert(c.I.copy(), c.h.copy(), c.m) for c in d.cinert]
  for i in range(nb - 1, 0, -1):
    p = m.body_parentid[i]
    if p > 0:
      comp[p] = comp[p].add(comp[i])
  M = np.zeros((nv, nv))
  for i in range(nv):
    buf = comp[m.dof_bodyid[i]].mul(d.cdof[i])
    M[i, i] = m.dof_armature[i] + np.dot(d.cdof[i], buf)
    j = m.dof_parentid[i]
    while j >= 0:
      M[i, j] = M[j, i] = np.dot(d.cdof[j], buf)
      j = m.dof_parentid[j]
  d.qM = M


# ---------------------------------------------------------------------------
# collision
# ---------------------------------------------------------------------------
class Contact:
  __slots__ = ("dist", "pos", "frame", "geom", "includemargin", "friction", "solref", "solimp", "dim", "efc_address")


def _plane_sphere(ppos, pn, spos, r):
  dist = np.dot(spos - ppos, pn) - r
  return dist, spos - pn * (r + 0.5 * dist)


def _sphere_sphere(p1, r1, p2, r2, margin):
  dif = p2 - p1
  c = np.linalg.norm(dif)
  if c > margin + r1 + r2:
    return None
  n = dif / c if c >= MINVAL else np.array([1.0, 0, 0])
  dist = c - r1 - r2
  return dist, p1 + n * (r1 + 0.5 * dist), n


def _narrowphase(m, d, g1, g2, margin):
  """Returns list of (dist, pos, normal, tangent|None), normal from g1 to g2."""
  t1, t2 = m.geom_type[g1], m.geom_type[g2]
  p1, p2 = d.geom_xpos[g1], d.geom_xpos[g2]
  R1, R2 = d.geom_xmat[g1], d.geom_xmat[g2]
  s1, s2 = m.geom_size[g1], m.geom_size[g2]
  out = []
  if t1 == GEOM_PLANE:
    n = R1[:, 2]
    if t2 == GEOM_SPHERE:
      dist, pos = _plane_sphere(p1, n, p2, s2[0])
      if dist <= margin:
        out.append((dist, pos, n, None))
    elif t2 == GEOM_CAPSULE:
      axis = R2[:, 2]
      for sgn in (1.0, -1.0):
        dist, pos = _plane_sphere(p1, n, p2 + sgn * axis * s2[1], s2[0])
        if dist <= margin:
          out.append((dist, pos, n, axis))
    elif t2 == GEOM_BOX:
      dist0 = np.dot(p2 - p1, n)
      for i in range(8):
        vec = np.array([s2[0] if i & 1 else -s2[0], s2[1] if i & 2 else -s2[1], s2[2] if i & 4 else -s2[2]])
        corner = R2 @ vec
        ld = np.dot(n, corner)
        if dist0 + ld > margin or ld > 0:
          continue
        dist = dist0 + ld
        out.append((dist, corner + p2 - n * dist * 0.5, n, None))
        if len(out) >= 4:
          break
    else:
      raise NotImplementedError(f"plane vs geom type {t2}")
  elif t1 == GEOM_SPHERE and t2 == GEOM_SPHERE:
    r = _sphere_sphere(p1, s1[0], p2, s2[0], margin)
    if r:
      out.append((r[0], r[1], r[2], None))
  elif t1 == GEOM_SPHERE and t2 == GEOM_CAPSULE:
    axis = R2[:, 2]
    x = np.clip(np.dot(axis, p1 - p2), -s2[1], s2[1])
    r = _sphere_sphere(p1, s1[0], p2 + axis * x, s2[0], margin)
    if r:
      out.append((r[0], r[1], r[2], None))
  elif t1 == GEOM_CAPSULE and t2 == GEOM_CAPSULE:
    a1, a2 = R1[:, 2], R2[:, 2]
    dif = p1 - p2
    ma, mb, mc = 1.0, -np.dot(a1, a2), 1.0
    u, v = -np.dot(a1, dif), np.dot(a2, dif)
    det = ma * mc - mb * mb
    l1, l2 = s1[1], s2[1]
    if abs(det) >= MINVAL:
      x1 = (mc * u - mb * v) / det
      x2 = (ma * v - mb * u) / det
      if x1 > l1:
        x1 = l1
        x2 = (v - mb * x1) / mc
      elif x1 < -l1:
        x1 = -l1
        x2 = (v - mb * x1) / mc
      if x2 > l2:
        x2 = l2
        x1 = (u - mb * x2) / ma
      elif x2 < -l2:
        x2 = -l2
        x1 = (u - mb * x2) / ma
      x1 = min(max(x1, -l1), l1)
      r = _sphere_sphere(p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0], margin)
      if r:
        out.append((r[0], r[1], r[2], None))
    else:
      # parallel axes: test the four end-point projections, keep at most two
      for sgn in (1.0, -1.0):
        v1 = p1 + a1 * l1 * sgn
        x2 = np.clip(np.dot(v1 - p2, a2), -l2, l2)
        r = _sphere_sphere(v1, s1[0], p2 + a2 * x2, s2[0], margin)
        if r and len(out) < 2:
          out.append((r[0], r[1], r[2], None))
      for sgn in (1.0, -1.0):
        if len(out) >= 2:
          break
        v2 = p2 + a2 * l2 * sgn
        x1 = np.clip(np.dot(v2 - p1, a1), -l1, l1)
        r = _sphere_sphere(p1 + a1 * x1, s1[0], v2, s2[0], margin)
        if r:
          out.append((r[0], r[1], r[2], None))
  elif t1 == GEOM_SPHERE and t2 == GEOM_BOX:
    r = _sphere_box(p1, s1[0], p2, R2, s2, margin)
    if r:
      out.append((r[0], r[1], r[2], None))
  elif t1 == GEOM_CAPSULE and t2 == GEOM_BOX:
    for r in _capsule_box(p1, R1[:, 2], s1[0], s1[1], p2, R2, s2, margin):
      out.append((r[0], r[1], r[2], None))
  elif t1 == GEOM_BOX and t2 == GEOM_BOX:
    for r in _box_box(p1, R1, s1, p2, R2, s2, margin):
      out.append((r[0], r[1], r[2], None))
  else:
    raise NotImplementedError(f"geom pair types ({t1},{t2}) outside the supported subset")
  return out


def _sphere_box(sp, r, bp, bR, bs, margin):
  """Sphere against box, after MuJoCo's published sphere-box routine (engine_collision_box.c `mjraw_SphereBox`; the
  reference reaches it through mujoco_warp's collision_primitive.sphere_box): clamp the sphere centre into the box; when the
  centre is inside, leave through the nearest face.  Normal points from the sphere to the box."""
  c = bR.T @ (sp - bp)
  clamped = np.clip(c, -bs, bs)
  dif = clamped - c
  dd = np.linalg.norm(dif)
  if dd - r > margin:
    return None
  if dd <= MINVAL:
    closest, k = 2 * bs.sum(), 0
    for i in range(6):
      fd = abs((bs[i // 2] if i % 2 else -bs[i // 2]) - c[i // 2])
      if fd < closest:
        closest, k = fd, i
    nb = np.zeros(3)
    nb[k // 2] = -1.0 if k % 2 else 1.0
    pb = c + nb * (r - closest) * 0.5
    dist = -closest - r
  else:
    nb = dif / dd
    dist = dd - r
    pb = c + nb * (r + 0.5 * dist)
  return dist, bp + bR @ pb, bR @ nb


BOXBOX_MAXCON = 4
BOXBOX_PARALLEL = 1e-10   # |A_i x B_j|^2 below this: the edge pair is parallel, no edge axis
BOXBOX_FACE_BIAS = 0.05   # an edge axis must beat the best face axis by this fraction (+1e-9) to be chosen


def _box_box(p1, R1, s1, p2, R2, s2, margin):
  """Box against box: separating-axis test over the 15 candidate axes (3 + 3 face normals, 9 edge cross products), then
    * face axis: the incident face of the other box is clipped against the side planes of the reference face (Sutherland-
      Hodgman, <= 8 points), points deeper than `margin` below the reference face are kept and reduced to <= 4 (deepest, farthest
      from it, farthest from that line on either side) - all contacts share the face normal;
    * edge axis: one contact between the closest points of the two supporting edges.
  This is the textbook SAT + clipping manifold (the scheme MuJoCo's `mjc_BoxBox` and mujoco_warp's `box_box` follow as well; they
  keep up to 8 points and differ in tie-breaking), restated from the published algorithm - PARITY UNPINNED like the other pair
  types.  Returns [(dist, pos, normal)], normal from box 1 to box 2, pos midway between the two surfaces."""
  d = p2 - p1
  A = [R1[:, k] for k in range(3)]
  B = [R2[:, k] for k in range(3)]

  def radii(L):
    return (s1[0] * abs(np.dot(A[0], L)) + s1[1] * abs(np.dot(A[1], L)) + s1[2] * abs(np.dot(A[2], L)),
            s2[0] * abs(np.dot(B[0], L)) + s2[1] * abs(np.dot(B[1], L)) + s2[2] * abs(np.dot(B[2], L)))

  best_s, best_code, best_sign = -np.inf, 0, 1.0
  for code in range(6):
    L = A[code] if code < 3 else B[code - 3]
    dl = np.dot(d, L)
    ra, rb = radii(L)
    sep = abs(dl) - ra - rb
    if sep > margin:
      return []
    if sep > best_s:
      best_s, best_code, best_sign = sep, code, (1.0 if dl >= 0 else -1.0)
  edge_s, edge_i, edge_j, edge_n = -np.inf, 0, 0, None
  for i in range(3):
    for j in range(3):
      L = np.cross(A[i], B[j])
      n2 = np.dot(L, L)
      if n2 < BOXBOX_PARALLEL:
        continue
      L = L / np.sqrt(n2)
      dl = np.dot(d, L)
      ra, rb = radii(L)
      sep = abs(dl) - ra - rb
      if sep > margin:
        return []
      if sep > edge_s:
        edge_s, edge_i, edge_j, edge_n = sep, i, j, L * (1.0 if dl >= 0 else -1.0)
  if edge_n is not None and edge_s > best_s + BOXBOX_FACE_BIAS * abs(best_s) + 1e-9:
    n, i, j = edge_n, edge_i, edge_j
    pa, pb = p1.copy(), p2.copy()
    for k in range(3):
      if k != i:
        pa = pa + (1.0 if np.dot(A[k], n) >= 0 else -1.0) * s1[k] * A[k]
      if k != j:
        pb = pb - (1.0 if np.dot(B[k], n) >= 0 else -1.0) * s2[k] * B[k]
    w = pa - pb
    b = np.dot(A[i], B[j])
    dd, e = np.dot(A[i], w), np.dot(B[j], w)
    den = 1.0 - b * b
    t = min(max((b * e - dd) / den, -s1[i]), s1[i])
    u = min(max((e - b * dd) / den, -s2[j]), s2[j])
    qa, qb = pa + t * A[i], pb + u * B[j]
    dist = np.dot(qb - qa, n)
    if dist > margin:
      return []
    return [(dist, 0.5 * (qa + qb), n)]
  # face contact
  if best_code < 3:
    pr, Rr, sr, pi, Ri, si, axis, flip = p1, R1, s1, p2, R2, s2, best_code, False
    nref = A[axis] * best_sign
  else:
    pr, Rr, sr, pi, Ri, si, axis, flip = p2, R2, s2, p1, R1, s1, best_code - 3, True
    nref = -B[axis] * best_sign
  ks, kbest = 0, -1.0
  for k in range(3):
    v = abs(np.dot(Ri[:, k], nref))
    if v > kbest:
      ks, kbest = k, v
  sg = -1.0 if np.dot(Ri[:, ks], nref) > 0 else 1.0
  ci = pi + sg * si[ks] * Ri[:, ks]
  iu, iv = (ks + 1) % 3, (ks + 2) % 3
  ru, rv = (axis + 1) % 3, (axis + 2) % 3
  hu, hv = sr[ru], sr[rv]
  poly = []
  for a, b in ((-1.0, -1.0), (1.0, -1.0), (1.0, 1.0), (-1.0, 1.0)):
    q = ci + a * si[iu] * Ri[:, iu] + b * si[iv] * Ri[:, iv] - pr
    poly.append((np.dot(q, Rr[:, ru]), np.dot(q, Rr[:, rv]), np.dot(q, nref) - sr[axis]))
  for c, lim, sgn in ((0, hu, 1.0), (0, hu, -1.0), (1, hv, 1.0), (1, hv, -1.0)):   # keep sgn * coord <= lim
    nxt = []
    for k in range(len(poly)):
      cur, prv = poly[k], poly[k - 1]
      fc, fp = sgn * cur[c] - lim, sgn * prv[c] - lim
      if (fc <= 0) != (fp <= 0):
        tt = fp / (fp - fc)
        nxt.append(tuple(prv[e] + tt * (cur[e] - prv[e]) for e in range(3)))
      if fc <= 0:
        nxt.append(cur)
    poly = nxt
    if not poly:
      return []
  cand = [pt for pt in poly if pt[2] <= margin]
  if len(cand) > BOXBOX_MAXCON:
    i0 = min(range(len(cand)), key=lambda k: (cand[k][2], k))
    d2 = [(cand[k][0] - cand[i0][0]) ** 2 + (cand[k][1] - cand[i0][1]) ** 2 for k in range(len(cand))]
    i1 = max(range(len(cand)), key=lambda k: (d2[k], -k))
    ex, ey = cand[i1][0] - cand[i0][0], cand[i1][1] - cand[i0][1]
    cr = [ex * (cand[k][1] - cand[i0][1]) - ey * (cand[k][0] - cand[i0][0]) for k in range(len(cand))]
    i2 = max(range(len(cand)), key=lambda k: (abs(cr[k]), -k))
    sel = [i0, i1, i2]
    opp = [(-cr[k] if cr[i2] > 0 else cr[k]) for k in range(len(cand))]
    i3 = max(range(len(cand)), key=lambda k: (opp[k], -k))
    if opp[i3] > 0 and i3 not in sel:
      sel.append(i3)
    cand = [cand[k] for k in sel]
  normal = -nref if flip else nref
  out = []
  for cu, cv, cn in cand:
    x = pr + cu * Rr[:, ru] + cv * Rr[:, rv] + (sr[axis] + cn) * nref
    out.append((cn, x - nref * (0.5 * cn), normal))
  return out


CAPBOX_BISECT = 24


def _capsule_box(cp, ax, r, h, bp, bR, bs, margin):
  """Capsule against box: up to two contacts, each a sphere-box test at a point of the capsule's segment, which is how
  MuJoCo's capsule-box routine emits its contacts.  The choice of the two points is this engine's own, simpler rule (MuJoCo
  enumerates face / edge / corner configurations): the squared distance from the segment point p + a t to the box is convex
  in t, so its half-derivative g(t) is nondecreasing; [ta, tb] = the stretch where g == 0 within [-h, h] (one point unless
  the segment runs parallel to a face or inside the box), found by bisection.  A stretch with a length gives its two ends;
  a single closest point gives itself plus the segment end farther from it.  Parity for this pair type is unpinned."""
  p = bR.T @ (cp - bp)
  a = bR.T @ ax

  def g(t):
    q = p + a * t
    e = np.where(q > bs, q - bs, np.where(q < -bs, q + bs, 0.0))
    return float(np.dot(a, e))

  g_lo, g_hi = g(-h), g(h)
  if g_lo >= 0:
    ta = -h
  elif g_hi < 0:
    ta = h
  else:
    lo, hi = -h, h
    for _ in range(CAPBOX_BISECT):
      mid = 0.5 * (lo + hi)
      if g(mid) >= 0:
        hi = mid
      else:
        lo = mid
    ta = hi
  if g_hi <= 0:
    tb = h
  elif g_lo > 0:
    tb = -h
  else:
    lo, hi = -h, h
    for _ in range(CAPBOX_BISECT):
      mid = 0.5 * (lo + hi)
      if g(mid) <= 0:
        lo = mid
      else:
        hi = mid
    tb = lo
  tol = 1e-4 * h
  if tb - ta > tol:
    t1, t2 = ta, tb
  else:
    t1 = 0.5 * (ta + tb)
    t2 = -h if t1 > 0 else h
  out = []
  r1 = _sphere_box(cp + ax * t1, r, bp, bR, bs, margin)
  if r1:
    out.append(r1)
  if abs(t2 - t1) > tol:
    r2 = _sphere_box(cp + ax * t2, r, bp, bR, bs, margin)
    if r2:
      out.append(r2)
  return out


def _mix_params(m, d, g1, g2):
  fr = getattr(d, "geom_friction", m.geom_friction)
  p1, p2 = m.geom_priority[g1], m.geom_priority[g2]
  if p1 != p2:
    g = g1 if p1 > p2 else g2
    condim, solref, solimp, fri = m.geom_condim[g], m.geom_solref[g].copy(), m.geom_solimp[g].copy(), fr[g].copy()
  else:
    condim = max(m.geom_condim[g1], m.geom_condim[g2])
    s1, s2 = m.geom_solmix[g1], m.geom_solmix[g2]
    if s1 >= MINVAL and s2 >= MINVAL:
      mix = s1 / (s1 + s2)
    elif s1 < MINVAL and s2 < MINVAL:
      mix = 0.5
    elif s1 < MINVAL:
      mix = 0.0
    else:
      mix = 1.0
    r1, r2 = m.geom_solref[g1], m.geom_solref[g2]
    if r1[0] > 0 and r2[0] > 0:
      solref = mix * r1 + (1 - mix) * r2
    else:
      solref = np.minimum(r1, r2)
    solimp = mix * m.geom_solimp[g1] + (1 - mix) * m.geom_solimp[g2]
    fri = np.maximum(fr[g1], fr[g2])
  margin = max(m.geom_margin[g1], m.geom_margin[g2])
  gap = max(m.geom_gap[g1], m.geom_gap[g2])
  friction = np.array([fri[0], fri[0], fri[1], fri[2], fri[2]])
  return int(condim), solref, solimp, friction, margin, gap


def _field_candidates(m, d):
  """(query geom, field geom) pairs under the bounding sphere of each query geom, in the order the uniform grid yields
  them: by query geom, grid cell (x-major), field index; a box met in several cells comes from the first shared one."""
  x0, y0, _, inv = m.field_grid
  out = []
  for g in m.fq_geom:
    c = d.geom_xpos[g]
    rb = m.geom_rbound[g] + m.geom_margin[g]
    cx0 = max(int(np.floor((c[0] - rb - x0) * inv)), 0)
    cx1 = min(int(np.floor((c[0] + rb - x0) * inv)), m.field_nx - 1)
    cy0 = max(int(np.floor((c[1] - rb - y0) * inv)), 0)
    cy1 = min(int(np.floor((c[1] + rb - y0) * inv)), m.field_ny - 1)
    for cx in range(cx0, cx1 + 1):
      for cy in range(cy0, cy1 + 1):
        cell = cx * m.field_ny + cy
        for fi in m.field_cell_item[m.field_cell_adr[cell]:m.field_cell_adr[cell + 1]]:
          fx0, fy0 = m.field_cell0[fi]
          if cx != max(cx0, fx0) or cy != max(cy0, fy0):
            continue
          a = m.field_aabb[fi]
          e = np.maximum(np.abs(c - a[:3]) - a[3:], 0.0)
          if np.dot(e, e) > rb * rb:
            continue
          out.append((int(g), int(m.field_geom[fi])))
  return out


def collision(m, d, nconmax=None):
  d.contacts = []
  cands = [(int(a), int(b), False) for a, b in np.asarray(m.pair_geom).reshape(-1, 2)]
  if getattr(m, "nfield", 0) > 0:
    cands += [(a, b, True) for a, b in _field_candidates(m, d)]
  for g1, g2, from_field in cands:
    condim, solref, solimp, friction, margin, gap = _mix_params(m, d, g1, g2)
    # bounding-sphere broadphase (planes are unbounded; field pairs passed their sphere-vs-AABB test already)
    if from_field:
      pass
    elif m.geom_type[g1] == GEOM_PLANE:
      n = d.geom_xmat[g1][:, 2]
      if np.dot(d.geom_xpos[g2] - d.geom_xpos[g1], n) > margin + m.geom_rbound[g2]:
        continue
    else:
      bound = m.geom_rbound[g1] + m.geom_rbound[g2] + margin
      if np.sum((d.geom_xpos[g2] - d.geom_xpos[g1]) ** 2) > bound * bound:
        continue
    for dist, pos, normal, tangent in _narrowphase(m, d, g1, g2, margin):
      if nconmax is not None and len(d.contacts) >= nconmax:
        return
      c = Contact()
      c.dist, c.pos, c.frame = dist, pos, make_frame(normal, tangent)
      c.geom = (int(g1), int(g2))
      c.includemargin = margin - gap
      c.friction, c.solref, c.solimp, c.dim = friction, solref, solimp, condim
      c.efc_address = -1
      d.contacts.append(c)


# ---------------------------------------------------------------------------
# constraints
# ---------------------------------------------------------------------------
def _jac_point(m, d, body, point):
  """Translational Jacobian (3 x nv) of a world point attached to `body`."""
  jp = np.zeros((3, m.nv))
  off = point - d.subtree_com[m.body_rootid[body]]
  b = body
  while b > 0 and m.body_dofnum[b] == 0:
    b = m.body_parentid[b]
  if b == 0:
    return jp
  i = m.body_dofadr[b] + m.body_dofnum[b] - 1
  while i >= 0:
    jp[:, i] = d.cdof[i, 3:] + np.cross(d.cdof[i, :3], off)
    i = m.dof_parentid[i]
  return jp


def _kbi(m, solref, solimp, pos):
  timeconst, dampratio = solref
  dmin, dmax, width, mid, power = solimp
  if timeconst > 0:
    timeconst = max(timeconst, 2 * m.opt_timestep)
  dmin = min(max(dmin, MINIMP), MAXIMP)
  dmax = min(max(dmax, MINIMP), MAXIMP)
  width = max(MINVAL, width)
  mid = min(max(mid, MINIMP), MAXIMP)
  power = max(1.0, power)
  if solref[0] > 0:
    k = 1.0 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio)
  else:
    k = -solref[0] / (dmax * dmax)
  if solref[1] > 0:
    b = 2.0 / (dmax * timeconst)
  else:
    b = -solref[1] / dmax
  x = abs(pos) / width
  if x >= 1.0:
    imp = dmax
  else:
    if x < mid:
      y = (1.0 / mid ** (power - 1)) * x**power
    else:
      y = 1.0 - (1.0 / (1 - mid) ** (power - 1)) * (1 - x) ** power
    imp = min(max(dmin + y * (dmax - dmin), min(dmin, dmax)), max(dmin, dmax))
  return k, b, imp


def make_constraint(m, d, njmax=None):
  nv = m.nv
  rows_J, rows_pos, rows_margin, rows_invw, rows_solref, rows_solimp, rows_type, rows_id = ([] for _ in range(8))
  rows_floss = []
  # friction loss (mj_instantiateFriction: one row per dof with frictionloss > 0; rows come first: friction, limits, contacts)
  dfl = np.asarray(getattr(d, "dof_frictionloss", getattr(m, "dof_frictionloss", np.zeros(nv))), float)
  dof_solref = np.asarray(getattr(m, "dof_solref", np.tile([0.02, 1.0], (nv, 1))), float).reshape(nv, 2)
  dof_solimp = np.asarray(getattr(m, "dof_solimp", np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (nv, 1))), float).reshape(nv, 5)
  for i in range(nv):
    if dfl[i] > 0:
      J = np.zeros(nv)
      J[i] = 1.0
      rows_J.append(J)
      rows_pos.append(0.0)
      rows_margin.append(0.0)
      rows_invw.append(getattr(d, "dof_invweight0", m.dof_invweight0)[i])
      rows_solref.append(dof_solref[i])
      rows_solimp.append(dof_solimp[i])
      rows_type.append(2)
      rows_id.append(i)
      rows_floss.append(dfl[i])
  if njmax is not None and len(rows_J) > njmax:
    for lst in (rows_J, rows_pos, rows_margin, rows_invw, rows_solref, rows_solimp, rows_type, rows_id, rows_floss):
      del lst[njmax:]
  d.nf = len(rows_J)
  # joint limits
  for j in range(m.njnt):
    if not m.jnt_limited[j] or m.jnt_type[j] not in (JNT_HINGE, JNT_SLIDE):
      continue
    q = d.qpos[m.jnt_qposadr[j]]
    lo, hi = m.jnt_range[j]
    for side, dist in ((1.0, q - lo), (-1.0, hi - q)):
      if dist < m.jnt_margin[j]:
        J = np.zeros(nv)
        J[m.jnt_dofadr[j]] = side
        rows_J.append(J)
        rows_pos.append(dist)
        rows_margin.append(m.jnt_margin[j])
        rows_invw.append(getattr(d, "dof_invweight0", m.dof_invweight0)[m.jnt_dofadr[j]])
        rows_solref.append(m.jnt_solref[j])
        rows_solimp.append(m.jnt_solimp[j])
        rows_type.append(0)
        rows_id.append(j)
  if njmax is not None and len(rows_J) > njmax:
    for lst in (rows_J, rows_pos, rows_margin, rows_invw, rows_solref, rows_solimp, rows_type, rows_id):
      del lst[njmax:]
  virtual_adr = len(rows_J)
  cones = []
  # contacts (pyramidal cone, or elliptic with opt_cone = 1)
  biw = getattr(d, "body_invweight0", m.body_invweight0)
  for ci, c in enumerate(d.contacts):
    c.efc_address = -1
    if c.dist >= c.includemargin:
      continue
    b1, b2 = m.geom_bodyid[c.geom[0]], m.geom_bodyid[c.geom[1]]
    jd = _jac_point(m, d, b2, c.pos) - _jac_point(m, d, b1, c.pos)
    jc = c.frame @ jd  # rows: normal, t1, t2
    invw = biw[b1, 0] + biw[b2, 0]
    elliptic = int(getattr(m, "opt_cone", 0)) == 1 and c.dim > 1
    nrow = 1 if c.dim == 1 else (c.dim if elliptic else 2 * (c.dim - 1))
    # row addresses are a running sum over all active contacts; a contact is kept iff
    # its rows fit below njmax (so everything after the first overflow is dropped)
    adr = virtual_adr
    virtual_adr += nrow
    if njmax is not None and adr + nrow > njmax:
      continue
    c.efc_address = len(rows_J)
    if c.dim == 1:
      rows = [jc[0]]
    elif elliptic:
      # elliptic cone (mj_instantiateContact, mjCNSTR_CONTACT_ELLIPTIC): the contact-frame axes themselves are the rows -
      # normal first, then the tangents; the friction rows' R is derived from the normal row's below
      rows = [jc[k] for k in range(c.dim)]
      cones.append((ci, c.efc_address, c.dim))
    else:
      fri0 = c.friction[0]
      invw = (invw + fri0 * fri0 * invw) * 2 * fri0 * fri0 / m.opt_impratio
      rows = []
      for k in range(1, c.dim):
        rows.append(jc[0] + c.friction[k - 1] * jc[k])
        rows.append(jc[0] - c.friction[k - 1] * jc[k])
    for J in rows:
      rows_J.append(J)
      rows_pos.append(c.dist)
      rows_margin.append(c.includemargin)
      rows_invw.append(invw)
      rows_solref.append(c.solref)
      rows_solimp.append(c.solimp)
      rows_type.append(1)
      rows_id.append(ci)
  n = len(rows_J)
  d.nefc = n
  d.efc_J = np.array(rows_J).reshape(n, nv)
  d.efc_pos = np.array(rows_pos)
  d.efc_margin = np.array(rows_margin)
  d.efc_type = np.array(rows_type, int)
  d.efc_id = np.array(rows_id, int)
  d.efc_D = np.zeros(n)
  d.efc_R = np.zeros(n)
  d.efc_frictionloss = np.zeros(n)
  d.efc_frictionloss[: d.nf] = rows_floss
  d.efc_k, d.efc_b, d.efc_imp = np.zeros(n), np.zeros(n), np.zeros(n)
  for i in range(n):
    k, b, imp = _kbi(m, rows_solref[i], rows_solimp[i], rows_pos[i] - rows_margin[i])
    d.efc_k[i], d.efc_b[i], d.efc_imp[i] = k, b, imp
    d.efc_R[i] = max(rows_invw[i] * (1 - imp) / imp, MINVAL)
    d.efc_D[i] = 1.0 / d.efc_R[i]
  # elliptic cones (mj_makeImpedance): friction rows have no position term (K = 0), their regulariser follows from the
  # normal row's - R_t1 = R_n / impratio, R_tj = R_t1 mu_1^2 / mu_j^2 - and the cone's effective friction is
  # mu = mu_1 sqrt(R_t1 / R_n)
  d.efc_cones = []
  for ci, a, dim in cones:
    c = d.contacts[ci]
    d.efc_R[a + 1] = d.efc_R[a] / max(MINVAL, m.opt_impratio)
    for j in range(2, dim):
      d.efc_R[a + j] = d.efc_R[a + 1] * c.friction[0] ** 2 / c.friction[j - 1] ** 2
    for j in range(1, dim):
      d.efc_D[a + j] = 1.0 / d.efc_R[a + j]
      d.efc_k[a + j] = 0.0
    mu = c.friction[0] * math.sqrt(d.efc_R[a + 1] / d.efc_R[a])
    d.efc_cones.append((a, dim, mu, np.asarray(c.friction[: dim - 1], float).copy()))


def reference_constraint(m, d):
  vel = d.efc_J @ d.qvel if d.nefc else np.zeros(0)
  d.efc_vel = vel
  d.efc_aref = -d.efc_b * vel - d.efc_k * d.efc_imp * (d.efc_pos - d.efc_margin)


# ---------------------------------------------------------------------------
# velocity / acceleration stages
# ---------------------------------------------------------------------------
def com_vel(m, d):
  nb = m.nbody
  d.cvel = np.zeros((nb, 6))
  d.cdof_dot = np.zeros((m.nv, 6))
  for i in range(1, nb):
    v = d.cvel[m.body_parentid[i]].copy()
    for k in range(m.body_jntnum[i]):
      j = m.body_jntadr[i] + k
      da = m.jnt_dofadr[j]
      if m.jnt_type[j] == JNT_FREE:
        for a in range(3):
          v += d.cdof[da + a] * d.qvel[da + a]
        for a in range(3, 6):
          d.cdof_dot[da + a] = cross_motion(v, d.cdof[da + a])
        for a in range(3, 6):
          v += d.cdof[da + a] * d.qvel[da + a]
      else:
        d.cdof_dot[da] = cross_motion(v, d.cdof[da])
        v += d.cdof[da] * d.qvel[da]
    d.cvel[i] = v


def rne(m, d, with_acc=False):
  nb = m.nbody
  cacc = np.zeros((nb, 6))
  cacc[0, 3:] = -m.opt_gravity
  cfrc = np.zeros((nb, 6))
  for i in range(1, nb):
    a = cacc[m.body_parentid[i]].copy()
    for k in range(m.body_dofnum[i]):
      dd = m.body_dofadr[i] + k
      a += d.cdof_dot[dd] * d.qvel[dd]
      if with_acc:
        a += d.cdof[dd] * d.qacc[dd]
    cacc[i] = a
    cfrc[i] = d.cinert[i].mul(a) + cross_force(d.cvel[i], d.cinert[i].mul(d.cvel[i]))
  if with_acc:
    d.cacc = cacc
    return None
  for i in range(nb - 1, 0, -1):
    cfrc[m.body_parentid[i]] += cfrc[i]
  return np.array([np.dot(d.cdof[k], cfrc[m.dof_bodyid[k]]) for k in range(m.nv)])


def passive(m, d):
  damping = getattr(d, "dof_damping", m.dof_damping)
  f = -damping * d.qvel
  for j in range(m.njnt):
    if m.jnt_stiffness[j] != 0 and m.jnt_type[j] in (JNT_HINGE, JNT_SLIDE):
      qa = m.jnt_qposadr[j]
      f[m.jnt_dofadr[j]] -= m.jnt_stiffness[j] * (d.qpos[qa] - m.qpos_spring[qa])
  d.qfrc_passive = f


def actuation(m, d):
  gainprm = getattr(d, "actuator_gainprm", m.actuator_gainprm)
  biasprm = getattr(d, "actuator_biasprm", m.actuator_biasprm)
  frange = getattr(d, "actuator_forcerange", m.actuator_forcerange)
  d.actuator_force = np.zeros(m.nu)
  d.actuator_length = np.zeros(m.nu)
  d.actuator_velocity = np.zeros(m.nu)
  d.qfrc_actuator = np.zeros(m.nv)
  for i in range(m.nu):
    j = m.actuator_trnid[i, 0]
    gear = m.actuator_gear[i, 0]
    length = gear * d.qpos[m.jnt_qposadr[j]]
    vel = gear * d.qvel[m.jnt_dofadr[j]]
    ctrl = d.ctrl[i]
    if m.actuator_ctrllimited[i]:
      ctrl = min(max(ctrl, m.actuator_ctrlrange[i, 0]), m.actuator_ctrlrange[i, 1])
    force = gainprm[i, 0] * ctrl
    if m.actuator_biastype[i] == 1:
      force += biasprm[i, 0] + biasprm[i, 1] * length + biasprm[i, 2] * vel
    if m.actuator_forcelimited[i]:
      force = min(max(force, frange[i, 0]), frange[i, 1])
    d.actuator_force[i], d.actuator_length[i], d.actuator_velocity[i] = force, length, vel
    d.qfrc_actuator[m.jnt_dofadr[j]] += gear * force


def _apply_ft(m, d, force, torque, point, body):
  jp = _jac_point(m, d, body, point)
  jr = np.zeros((3, m.nv))
  b = body
  while b > 0 and m.body_dofnum[b] == 0:
    b = m.body_parentid[b]
  if b > 0:
    i = m.body_dofadr[b] + m.body_dofnum[b] - 1
    while i >= 0:
      jr[:, i] = d.cdof[i, :3]
      i = m.dof_parentid[i]
  return jp.T @ force + jr.T @ torque


def smooth_acc(m, d):
  d.qfrc_bias = rne(m, d)
  f = d.qfrc_passive - d.qfrc_bias + d.qfrc_actuator + d.qfrc_applied
  for b in range(1, m.nbody):
    if np.any(d.xfrc_applied[b] != 0):
      f = f + _apply_ft(m, d, d.xfrc_applied[b, :3], d.xfrc_applied[b, 3:], d.xipos[b], b)
  d.qfrc_smooth = f
  d.qacc_smooth = np.linalg.solve(d.qM, f) if m.nv else f


# ---------------------------------------------------------------------------
# Newton / CG solver (primal; friction-loss rows are Huber-type, limit and pyramidal rows are inequalities)
# ---------------------------------------------------------------------------
class _Pnt:
  __slots__ = ("alpha", "cost", "d1", "d2")


def _cone_zone(N, T, mu):
  """Zone of an elliptic-cone contact from its scaled residual: 0 = satisfied (inside the dual cone), 1 = fully violated
  (quadratic in every row), 2 = on the cone surface."""
  if N >= mu * T or (T <= 0 and N >= 0):
    return 0
  if mu * N + T <= 0 or (T <= 0 and N < 0):
    return 1
  return 2


def _row_update(jar, D, R, floss, nf, cones=(), hcones=None):
  """mj_constraintUpdate per row: (cost terms, force, rows that enter the Hessian).  Rows [0, nf) are friction-loss rows:
  quadratic for |jar| < R f, linear beyond with the force saturated at -/+ f; the others are active iff jar < 0.
  `cones`: elliptic contacts (address, dim, mu, friction); in the middle zone the cost is 0.5 Dm (N - mu T)^2 with
  N = mu jar_n, T = |friction_j jar_tj| and Dm = D_n / (mu^2 (1 + mu^2)); such a contact's rows leave the diagonal
  Hessian and its dense dim x dim block (second derivative of that cost) is appended to `hcones`."""
  act = jar < 0
  force = np.where(act, -D * jar, 0.0)
  cost = np.where(act, 0.5 * D * jar * jar, 0.0)
  for a, dim, mu, fri in cones:
    x = jar[a:a + dim]
    sc = np.concatenate(([mu], fri))
    U = x * sc
    N, T = U[0], float(np.linalg.norm(U[1:]))
    zone = _cone_zone(N, T, mu)
    if zone == 0:
      force[a:a + dim], cost[a:a + dim], act[a:a + dim] = 0.0, 0.0, False
    elif zone == 1:
      force[a:a + dim], cost[a:a + dim], act[a:a + dim] = -D[a:a + dim] * x, 0.5 * D[a:a + dim] * x * x, True
    else:
      Dm = D[a] / (mu * mu * (1 + mu * mu))
      NmT = N - mu * T
      cost[a:a + dim] = 0.0
      cost[a] = 0.5 * Dm * NmT * NmT
      force[a] = -Dm * NmT * mu
      force[a + 1:a + dim] = -force[a] / T * U[1:] * fri
      act[a:a + dim] = False
      if hcones is not None:
        Hc = np.zeros((dim, dim))
        Hc[0, 0] = 1.0
        Hc[0, 1:] = Hc[1:, 0] = -mu * U[1:] / T
        Hc[1:, 1:] = mu * N / T ** 3 * np.outer(U[1:], U[1:]) + (mu * mu - mu * N / T) * np.eye(dim - 1)
        hcones.append((a, dim, Dm * Hc * np.outer(sc, sc)))
  for i in range(nf):
    f, rf = floss[i], R[i] * floss[i]
    if jar[i] <= -rf:
      force[i], cost[i], act[i] = f, -0.5 * rf * f - f * jar[i], False
    elif jar[i] >= rf:
      force[i], cost[i], act[i] = -f, -0.5 * rf * f + f * jar[i], False
    else:
      force[i], cost[i], act[i] = -D[i] * jar[i], 0.5 * D[i] * jar[i] * jar[i], True
  return cost, force, act


def _ls_eval(alpha, quad_gauss, jar, jv, D, R=None, floss=None, nf=0, cones=()):
  x = jar + alpha * jv
  act = x < 0
  act[:nf] = False
  ec = np.zeros(3)   # elliptic contacts add cost / first / second derivative at alpha directly (mj PrimalEval)
  for a, dim, mu, fri in cones:
    act[a:a + dim] = False
    sc = np.concatenate(([mu], fri))
    U, V = jar[a:a + dim] * sc, jv[a:a + dim] * sc
    UU, UV, VV = np.dot(U[1:], U[1:]), np.dot(U[1:], V[1:]), np.dot(V[1:], V[1:])
    N = U[0] + alpha * V[0]
    Tsq = UU + alpha * (2 * UV + alpha * VV)
    T = math.sqrt(Tsq) if Tsq > 0 else 0.0
    zone = _cone_zone(N, T, mu)
    if zone == 1:
      Da, xa, va = D[a:a + dim], x[a:a + dim], jv[a:a + dim]
      ec += (0.5 * np.sum(Da * xa * xa), np.sum(Da * xa * va), np.sum(Da * va * va))
    elif zone == 2:
      Dm = D[a] / (mu * mu * (1 + mu * mu))
      N1, T1 = V[0], (UV + alpha * VV) / T
      T2 = VV / T - (UV + alpha * VV) ** 2 / T ** 3
      NmT = N - mu * T
      ec += (0.5 * Dm * NmT * NmT, Dm * NmT * (N1 - mu * T1), Dm * ((N1 - mu * T1) ** 2 - NmT * mu * T2))
  q0 = quad_gauss[0] + 0.5 * np.sum(D[act] * jar[act] ** 2)
  q1 = quad_gauss[1] + np.sum(D[act] * jar[act] * jv[act])
  q2 = quad_gauss[2] + 0.5 * np.sum(D[act] * jv[act] ** 2)
  for i in range(nf):   # mj PrimalEval, friction rows
    f, rf = floss[i], R[i] * floss[i]
    if -rf < x[i] < rf:
      q0 += 0.5 * D[i] * jar[i] * jar[i]
      q1 += D[i] * jar[i] * jv[i]
      q2 += 0.5 * D[i] * jv[i] * jv[i]
    elif x[i] <= -rf:
      q0 += f * (-0.5 * rf - jar[i])
      q1 += -f * jv[i]
    else:
      q0 += f * (-0.5 * rf + jar[i])
      q1 += f * jv[i]
  p = _Pnt()
  p.alpha = alpha
  p.cost = alpha * alpha * q2 + alpha * q1 + q0 + ec[0]
  p.d1 = 2 * alpha * q2 + q1 + ec[1]
  p.d2 = 2 * q2 + ec[2]
  if p.d2 <= 0:
    p.d2 = MINVAL
  return p


def _linesearch(m, d, st, search, tolerance, ls_tolerance, ls_iterations, scale):
  Mv = d.qM @ search
  jv = d.efc_J @ search
  st["Mv"], st["jv"] = Mv, jv
  snorm = np.linalg.norm(search)
  if snorm < MINVAL:
    return 0.0
  gtol = tolerance * ls_tolerance * snorm / scale
  quad_gauss = (st["gauss"], np.dot(search, st["Ma"]) - np.dot(search, d.qfrc_smooth), 0.5 * np.dot(search, Mv))
  nev = [0]

  def ev(a):
    nev[0] += 1
    return _ls_eval(a, quad_gauss, st["jar"], jv, d.efc_D, d.efc_R, d.efc_frictionloss, d.nf, d.efc_cones)

  p0 = _ls_eval(0.0, quad_gauss, st["jar"], jv, d.efc_D, d.efc_R, d.efc_frictionloss, d.nf, d.efc_cones)
  p1 = ev(p0.alpha - p0.d1 / p0.d2)
  if p0.cost < p1.cost:
    p1 = p0
  if abs(p1.d1) < gtol:
    return p1.alpha
  direction = 1.0 if p1.d1 < 0 else -1.0
  p2 = None
  while p1.d1 * direction <= -gtol and nev[0] < ls_iterations:
    p2 = p1
    p1 = ev(p1.alpha - p1.d1 / p1.d2)
    if abs(p1.d1) < gtol:
      return p1.alpha
  if nev[0] >= ls_iterations or p2 is None:
    return p1.alpha
  p2next = p1
  p1next = ev(p1.alpha - p1.d1 / p1.d2)

  def update_bracket(p, cands):
    flag = False
    for c in cands:
      if p.d1 < 0 and c.d1 < 0 and p.d1 < c.d1:
        p, flag = c, True
      elif p.d1 > 0 and c.d1 > 0 and p.d1 > c.d1:
        p, flag = c, True
    pn = ev(p.alpha - p.d1 / p.d2) if flag else None
    return p, pn, flag

  while nev[0] < ls_iterations:
    pmid = ev(0.5 * (p1.alpha + p2.alpha))
    cands = (p1next, p2next, pmid)
    best = None
    for c in cands:
      if abs(c.d1) < gtol and (best is None or c.cost < best.cost):
        best = c
    if best is not None:
      return best.alpha
    p1, n1, b1 = update_bracket(p1, cands)
    p2, n2, b2 = update_bracket(p2, cands)
    if b1:
      p1next = n1
    if b2:
      p2next = n2
    if not b1 and not b2:
      return pmid.alpha if pmid.cost < p0.cost else 0.0
  if p1.cost <= p2.cost and p1.cost < p0.cost:
    return p1.alpha
  if p2.cost <= p1.cost and p2.cost < p0.cost:
    return p2.alpha
  return 0.0


def solve(m, d):
  nv = m.nv
  if d.nefc == 0:
    d.qacc = d.qacc_smooth.copy()
    d.efc_force = np.zeros(0)
    d.qfrc_constraint = np.zeros(nv)
    d.solver_niter = 0
    return
  J, D, aref, M = d.efc_J, d.efc_D, d.efc_aref, d.qM
  scale = 1.0 / (m.stat_meaninertia * max(1, nv))

  R, floss, nf = d.efc_R, d.efc_frictionloss, d.nf
  cones = getattr(d, "efc_cones", ())

  def cost_of(qacc):
    jar = J @ qacc - aref
    Ma = M @ qacc
    return np.sum(_row_update(jar, D, R, floss, nf, cones)[0]) + 0.5 * np.dot(Ma - d.qfrc_smooth, qacc - d.qacc_smooth)

  # warm start: keep qacc_warmstart unless the unconstrained acceleration is cheaper
  qacc = d.qacc_warmstart.copy()
  if cost_of(qacc) > cost_of(d.qacc_smooth):
    qacc = d.qacc_smooth.copy()
  st = {"Ma": M @ qacc, "jar": J @ qacc - aref}

  def update(st, qacc):
    jar = st["jar"]
    hcones = []
    rowcost, force, act = _row_update(jar, D, R, floss, nf, cones, hcones)
    st["active"], st["force"] = act, force
    st["gauss"] = 0.5 * np.dot(st["Ma"] - d.qfrc_smooth, qacc - d.qacc_smooth)
    st["cost"] = np.sum(rowcost) + st["gauss"]
    st["qfrc_constraint"] = J.T @ force
    st["grad"] = st["Ma"] - d.qfrc_smooth - st["qfrc_constraint"]
    if cg:
      st["Mgrad"] = np.linalg.solve(M, st["grad"])      # CG: the mass matrix is the preconditioner (mj_solCG)
    else:
      H = M + (J[act].T * D[act]) @ J[act]
      for a, dim, Hc in hcones:   # cone Hessians of the contacts on their cone's surface (mj HessianCone)
        H = H + J[a:a + dim].T @ Hc @ J[a:a + dim]
      st["Mgrad"] = np.linalg.solve(H, st["grad"])

  cg = int(getattr(m, "opt_solver", SOL_NEWTON)) == SOL_CG
  update(st, qacc)
  niter = 0
  search = -st["Mgrad"]
  while niter < m.opt_iterations:
    alpha = _linesearch(m, d, st, search, m.opt_tolerance, m.opt_ls_tolerance, m.opt_ls_iterations, scale)
    if alpha == 0.0:
      break
    qacc = qacc + alpha * search
    st["Ma"] = st["Ma"] + alpha * st["Mv"]
    st["jar"] = st["jar"] + alpha * st["jv"]
    oldcost = st["cost"]
    gradold, Mgradold = st["grad"], st["Mgrad"]
    update(st, qacc)
    if cg:   # Polak-Ribiere with the M^-1 preconditioner, restarted when negative (MuJoCo's mj_solCG / mj_solPrimal)
      beta = np.dot(st["grad"], st["Mgrad"] - Mgradold) / max(MINVAL, np.dot(gradold, Mgradold))
      search = -st["Mgrad"] + max(0.0, beta) * search
    else:
      search = -st["Mgrad"]
    niter += 1
    improvement = scale * (oldcost - st["cost"])
    gradient = scale * np.linalg.norm(st["grad"])
    # the float64 rounding floor keeps the rule identical to the fp32 device code, where it matters
    floor = max(m.opt_tolerance, 8 * np.finfo(float).eps * scale * abs(st["cost"]))
    if improvement < floor or gradient < m.opt_tolerance:
      break
  d.qacc = qacc
  d.efc_force = st["force"]
  d.qfrc_constraint = st["qfrc_constraint"]
  d.solver_niter = niter


# ---------------------------------------------------------------------------
# sensors
# ---------------------------------------------------------------------------
def _object_velocity(m, d, body, pos, rot=None):
  """6D velocity [ang; lin] of a point `pos` on `body`, optionally in frame `rot`."""
  v = d.cvel[body]
  off = pos - d.subtree_com[m.body_rootid[body]]
  lin = v[3:] + np.cross(v[:3], off)
  ang = v[:3]
  if rot is not None:
    return np.concatenate([rot.T @ ang, rot.T @ lin])
  return np.concatenate([ang, lin])


def subtree_vel(m, d):
  nb = m.nbody
  mass = getattr(d, "body_mass", m.body_mass)
  lin = np.zeros((nb, 3))
  ang_body = np.zeros((nb, 3))
  for i in range(nb):
    v = _object_velocity(m, d, i, d.xipos[i])
    lin[i] = v[3:]
    ang_body[i] = d.ximat[i] @ (m.body_inertia[i] * (d.ximat[i].T @ v[:3]))
  mom = mass[:, None] * lin
  for i in range(nb - 1, 0, -1):
    mom[m.body_parentid[i]] += mom[i]
  d.subtree_linvel = mom / np.maximum(d.subtree_mass[:, None], MINVAL)
  # angular momentum of every subtree about its own com (direct sum over members)
  d.subtree_angmom = np.zeros((nb, 3))
  for j in range(nb):
    b = j
    while True:
      dx = d.xipos[j] - d.subtree_com[b]
      dv = lin[j] - d.subtree_linvel[b]
      d.subtree_angmom[b] += ang_body[j] + mass[j] * np.cross(dx, dv)
      if b == 0:
        break
      b = m.body_parentid[b]


def contact_force(m, d, c):
  """Contact-frame force [normal, t1, t2] from the pyramid edge forces."""
  out = np.zeros(3)
  if c.efc_address < 0:
    return out
  a = c.efc_address
  if c.dim == 1:
    out[0] = d.efc_force[a]
    return out
  if int(getattr(m, "opt_cone", 0)) == 1:   # elliptic: the rows are the contact-frame axes
    out[: min(c.dim, 3)] = d.efc_force[a:a + min(c.dim, 3)]
    return out
  for k in range(1, c.dim):
    f1, f2 = d.efc_force[a + 2 * (k - 1)], d.efc_force[a + 2 * (k - 1) + 1]
    out[0] += f1 + f2
    if k < 3:
      out[k] = (f1 - f2) * c.friction[k - 1]
  return out


def _in_subtree(m, body, root):
  b = body
  while True:
    if b == root:
      return True
    if b == 0:
      return False
    b = m.body_parentid[b]


def _match_side(m, objtype, objid, geom):
  if objid < 0:
    return True
  body = m.geom_bodyid[geom]
  if objtype == OBJ_GEOM:
    return geom == objid
  if objtype == OBJ_BODY:
    return body == objid
  if objtype == OBJ_XBODY:
    return _in_subtree(m, body, objid)
  return False


def sensors(m, d):
  d.sensordata = np.zeros(m.nsensordata)
  need_sub = any(t in (SENS_SUBTREELINVEL, SENS_SUBTREEANGMOM) for t in m.sensor_type)
  if need_sub:
    subtree_vel(m, d)
  need_acc = any(t == SENS_ACCELEROMETER for t in m.sensor_type)
  if need_acc:
    rne(m, d, with_acc=True)
  for s in range(m.nsensor):
    t, oid, adr, dim = m.sensor_type[s], m.sensor_objid[s], m.sensor_adr[s], m.sensor_dim[s]
    out = d.sensordata[adr:adr + dim]
    if t in (SENS_GYRO, SENS_VELOCIMETER, SENS_ACCELEROMETER):
      b = m.site_bodyid[oid]
      R = d.site_xmat[oid]
      v = _object_velocity(m, d, b, d.site_xpos[oid], R)
      if t == SENS_GYRO:
        out[:] = v[:3]
      elif t == SENS_VELOCIMETER:
        out[:] = v[3:]
      else:
        a = d.cacc[b]
        off = d.site_xpos[oid] - d.subtree_com[m.body_rootid[b]]
        lin = R.T @ (a[3:] + np.cross(a[:3], off))
        out[:] = lin + np.cross(v[:3], v[3:])
    elif t == SENS_JOINTPOS:
      out[0] = d.qpos[m.jnt_qposadr[oid]]
    elif t == SENS_JOINTVEL:
      out[0] = d.qvel[m.jnt_dofadr[oid]]
    elif t == SENS_SUBTREECOM:
      out[:] = d.subtree_com[oid]
    elif t == SENS_SUBTREELINVEL:
      out[:] = d.subtree_linvel[oid]
    elif t == SENS_SUBTREEANGMOM:
      out[:] = d.subtree_angmom[oid]
    elif t in (SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_FRAMELINVEL, SENS_FRAMEANGVEL):
      ot = m.sensor_objtype[s]
      if ot == OBJ_SITE:
        pos, R, b = d.site_xpos[oid], d.site_xmat[oid], m.site_bodyid[oid]
      elif ot == OBJ_GEOM:
        pos, R, b = d.geom_xpos[oid], d.geom_xmat[oid], m.geom_bodyid[oid]
      elif ot == OBJ_XBODY:
        pos, R, b = d.xpos[oid], d.xmat[oid], oid
      else:
        pos, R, b = d.xipos[oid], d.ximat[oid], oid
      if t == SENS_FRAMEPOS:
        out[:] = pos
      elif t == SENS_FRAMEQUAT:
        out[:] = mat_to_quat(R)
      else:
        v = _object_velocity(m, d, b, pos)
        out[:] = v[3:] if t == SENS_FRAMELINVEL else v[:3]
    elif t == SENS_CONTACT:
      bits, reduce, num = m.sensor_intprm[s]
      matches = []
      for c in d.contacts:
        g1, g2 = c.geom
        ot, rt, rid = m.sensor_objtype[s], m.sensor_reftype[s], m.sensor_refid[s]
        if _match_side(m, ot, oid, g1) and _match_side(m, rt, rid, g2):
          matches.append((c, 1.0))
        elif _match_side(m, ot, oid, g2) and _match_side(m, rt, rid, g1):
          matches.append((c, -1.0))
      if not matches:
        continue
      nfound = len(matches)
      if reduce == 3:  # netforce: one slot, global frame
        net = np.zeros(3)
        for c, sgn in matches:
          net += sgn * (c.frame.T @ contact_force(m, d, c))
        slots = [("net", net)]
      else:
        if reduce == 1:
          matches.sort(key=lambda cs: cs[0].dist)
        elif reduce == 2:
          matches.sort(key=lambda cs: -np.linalg.norm(contact_force(m, d, cs[0])))
        slots = [("c", cs) for cs in matches[:num]]
      p = 0
      per_slot = dim // num
      for k, (kind, val) in enumerate(slots):
        p = k * per_slot
        if bits & 1:
          out[p] = nfound
          p += 1
        if bits & 2:
          out[p:p + 3] = val if kind == "net" else val[1] * contact_force(m, d, val[0])
          p += 3
        if bits & 4:
          p += 3  # torque: zero for condim <= 3
        if kind == "c":
          c, sgn = val
          if bits & 8:
            out[p] = c.dist
            p += 1
          if bits & 16:
            out[p:p + 3] = c.pos
            p += 3
          if bits & 32:
            out[p:p + 3] = sgn * c.frame[0]
            p += 3
          if bits & 64:
            out[p:p + 3] = sgn * c.frame[1]
            p += 3
    cut = m.sensor_cutoff[s]
    if cut > 0 and t != SENS_CONTACT:
      np.clip(out, -cut, cut, out=out)


# ---------------------------------------------------------------------------
# top level
# ---------------------------------------------------------------------------
def forward(m, d, nconmax=None, njmax=None):
  kinematics(m, d)
  com_pos(m, d)
  crb(m, d)
  collision(m, d, nconmax)
  make_constraint(m, d, njmax)
  com_vel(m, d)
  passive(m, d)
  actuation(m, d)
  smooth_acc(m, d)
  reference_constraint(m, d)
  solve(m, d)
  sensors(m, d)


def integrate(m, d):
  h = m.opt_timestep
  nv = m.nv
  damping = getattr(d, "dof_damping", m.dof_damping)
  qacc = d.qacc
  if m.opt_integrator == INT_IMPLICITFAST:
    qd = -damping.astype(float).copy()
    biasprm = getattr(d, "actuator_biasprm", m.actuator_biasprm)
    frange = getattr(d, "actuator_forcerange", m.actuator_forcerange)
    for i in range(m.nu):
      if m.actuator_biastype[i] != 1:
        continue
      f = d.actuator_force[i]
      if m.actuator_forcelimited[i] and (f <= frange[i, 0] or f >= frange[i, 1]):
        continue
      g = m.actuator_gear[i, 0]
      qd[m.jnt_dofadr[m.actuator_trnid[i, 0]]] += g * g * biasprm[i, 2]
    qacc = np.linalg.solve(d.qM - h * np.diag(qd), d.qfrc_smooth + d.qfrc_constraint)
  elif m.opt_integrator == INT_EULER:
    if np.any(damping > 0):
      qacc = np.linalg.solve(d.qM + h * np.diag(damping), d.qfrc_smooth + d.qfrc_constraint)
  else:
    raise NotImplementedError("integrator")
  d.qvel = d.qvel + h * qacc
  for j in range(m.njnt):
    qa, da = m.jnt_qposadr[j], m.jnt_dofadr[j]
    if m.jnt_type[j] == JNT_FREE:
      d.qpos[qa:qa + 3] += h * d.qvel[da:da + 3]
      d.qpos[qa + 3:qa + 7] = quat_integrate(d.qpos[qa + 3:qa + 7], d.qvel[da + 3:da + 6], h)
    else:
      d.qpos[qa] += h * d.qvel[da]
  d.time += h
  d.qacc_warmstart = d.qacc.copy()


def step(m, d, nconmax=None, njmax=None):
  forward(m, d, nconmax, njmax)
  integrate(m, d)
