// mj_oracle.cpp — fp64 CPU restatement of `mj_step` / `mj_forward` / `mj_resetData`
// for the MJCF subset of the reference's levels.
//
// *** TEST INFRASTRUCTURE ONLY — PARITY UNPINNED ***
// This file is the CHECKER for the CUDA step path; nothing shipped may call it (only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do).
// The arithmetic it restates lives in the third-party dependency `mujoco==2.3.3`
// (reference requirements.txt:45), which is NOT vendored under the reference tree and is not
// installable here; the reference's own tests hold no golden numbers for it.
// So this restates MuJoCo's *published* pipeline (documentation "Computation" chapter), anchored
// on the reference's call sites:
//   mj.mj_step      MuJoCo_Gym/mujoco_parent.py:335,362   -> orc_step
//   mj.mj_forward   MuJoCo_Gym/mujoco_parent.py:350,355   -> orc_forward
//   mj.mj_resetData MuJoCo_Gym/mujoco_parent.py:349,354   -> orc_reset
//   data.qpos/qvel/ctrl/sensordata, data.body(n).xipos/xmat, data.geom(n).xpos/xmat,
//   data.contact[i].geom1/geom2, data.ncon   mujoco_parent.py:325,330,375,390-391,404-425,441-443,472-475
// It is pinned instead by analytic known-answer tests (tests/test_oracle_analytic.py).
//
// Formulation is deliberately the dense textbook one (world-frame Jacobians, dense M, dense
// Cholesky, spatial RNE about the world origin, exact piecewise-quadratic line search) so that it
// shares no structure with the warp-parallel tree recursions of the CUDA kernels.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// No product source is included: the oracle's own linear algebra, blob reader, kinematics, Jacobians and
// mass matrix live in orc_base.h; the only shared artefact is the PUBLIC blob format (include/mjb_blob.h),
// whose every field is pinned by an independent derivation from the XML (oracle/model_ref.py,
// tests/test_model_constants.py).
#include "orc_base.h"

using namespace orc;
using ModelView = orc::Model;

namespace {

const double kMinVal = 1e-15;

struct Contact {
  double dist = 0;
  V3 pos;
  V3 frame[3];  // normal, tangent1, tangent2
  int g1 = -1, g2 = -1, pair = -1;
  double includemargin = 0, friction[3] = {0, 0, 0}, solref[2] = {0, 0}, solimp[5] = {0, 0, 0, 0, 0};
  int condim = 3, efc_address = -1;
};

struct Sim {
  std::vector<uint8_t> blob;
  ModelView* m = nullptr;
  int nq, nv, nu;
  std::vector<double> qpos, qvel, ctrl, qacc, qacc_warmstart, qacc_smooth, qfrc_smooth, qfrc_bias, qfrc_passive,
      qfrc_actuator, qfrc_constraint, sensordata;
  double time = 0;
  Kin kin;
  std::vector<double> xpos, xmat, xipos, geom_xpos, geom_xmat, site_xpos, site_xmat;  // flat mirrors for the API
  std::vector<double> M, L;                                                             // dense mass matrix, Cholesky
  std::vector<Contact> contacts;
  // constraint rows
  int nefc = 0;
  std::vector<double> J, efc_pos, efc_margin, efc_D, efc_R, efc_aref, efc_force, efc_diagApprox;
  std::vector<int> efc_type, efc_id;  // type 0 = joint limit, 1 = pyramidal contact row, 2 = frictionless
  // body spatial velocity / acceleration about the world origin ([ang; lin])
  std::vector<double> bvel, bacc;
  // solver configuration
  int solver_mode = 0;  // 0 exact Newton to convergence, 1 Newton fixed iterations, 2 PGS fixed, 3 PGS converged
  int solver_iters = 100, ls_iters = 50;
  int last_solver_iters = 0;
  bool warmstart = true;
  int warmstart_mode = 1;
  int contact_geom_out[2 * 256];
};

V3 col(const M3& R, int c) { return V3(R(0, c), R(1, c), R(2, c)); }

// ---------------------------------------------------------------------------------------------
// kinematics + geom / site frames
void kinematics(Sim& s) {
  const ModelView& m = *s.m;
  // normalise free-joint quaternions in qpos (MuJoCo does this at the top of its kinematics pass)
  for (int j = 0; j < m.njnt; j++)
    if (m.jnt_type[j] == MJB_JNT_FREE) {
      double* q = &s.qpos[m.jnt_qposadr[j] + 3];
      double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      if (n < kMinVal) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
      else if (std::fabs(n - 1) > kMinVal) for (int i = 0; i < 4; i++) q[i] /= n;
    }
  forward_kinematics(m, s.qpos.data(), s.kin);
  for (int b = 0; b < m.nbody; b++) {
    for (int i = 0; i < 3; i++) { s.xpos[3 * b + i] = s.kin.xpos[b][i]; s.xipos[3 * b + i] = s.kin.xipos[b][i]; }
    for (int i = 0; i < 9; i++) s.xmat[9 * b + i] = s.kin.xmat[b].m[i];
  }
  for (int g = 0; g < m.ngeom; g++) {
    int b = m.geom_bodyid[g];
    V3 p = s.kin.xpos[b] + mulv(s.kin.xmat[b], V3(m.geom_pos[3 * g], m.geom_pos[3 * g + 1], m.geom_pos[3 * g + 2]));
    M3 R = mul(s.kin.xmat[b], rot_from_quat(m.geom_quat + 4 * g));
    for (int i = 0; i < 3; i++) s.geom_xpos[3 * g + i] = p[i];
    for (int i = 0; i < 9; i++) s.geom_xmat[9 * g + i] = R.m[i];
  }
  for (int t = 0; t < m.nsite; t++) {
    int b = m.site_bodyid[t];
    V3 p = s.kin.xpos[b] + mulv(s.kin.xmat[b], V3(m.site_pos[3 * t], m.site_pos[3 * t + 1], m.site_pos[3 * t + 2]));
    M3 R = mul(s.kin.xmat[b], rot_from_quat(m.site_quat + 4 * t));
    for (int i = 0; i < 3; i++) s.site_xpos[3 * t + i] = p[i];
    for (int i = 0; i < 9; i++) s.site_xmat[9 * t + i] = R.m[i];
  }
}

V3 gpos(const Sim& s, int g) { return V3(s.geom_xpos[3 * g], s.geom_xpos[3 * g + 1], s.geom_xpos[3 * g + 2]); }
M3 gmat(const Sim& s, int g) { M3 R; for (int i = 0; i < 9; i++) R.m[i] = s.geom_xmat[9 * g + i]; return R; }

// ---------------------------------------------------------------------------------------------
// spatial algebra about the world origin: vectors are [ang(3); lin(3)]
struct Sp { V3 a, l; };
Sp sp_add(Sp x, Sp y) { return {x.a + y.a, x.l + y.l}; }
Sp sp_scale(Sp x, double k) { return {x.a * k, x.l * k}; }
Sp sp_cross_motion(Sp v, Sp s) { return {cross(v.a, s.a), cross(v.a, s.l) + cross(v.l, s.a)}; }
Sp sp_cross_force(Sp v, Sp f) { return {cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)}; }
double sp_dot(Sp s, Sp f) { return dot(s.a, f.a) + dot(s.l, f.l); }

// motion subspace of a dof about the world origin
Sp dof_subspace(const Sim& s, int d) {
  V3 rot = s.kin.dof_axis_rot[d], lin = s.kin.dof_axis_lin[d];
  return {rot, lin + cross(s.kin.dof_anchor[d], rot)};
}
// is the dof axis fixed in the world (free-joint translations) rather than in the moving body?
bool dof_world_fixed(const ModelView& m, int d) {
  int j = m.dof_jntid[d];
  return m.jnt_type[j] == MJB_JNT_FREE && d - m.jnt_dofadr[j] < 3;
}
// spatial inertia (about the world origin) applied to a motion vector
Sp inertia_apply(const Sim& s, int b, Sp v) {
  const ModelView& m = *s.m;
  double mass = m.body_mass[b];
  V3 c = s.kin.xipos[b];
  M3 R = s.kin.ximat[b];
  V3 vc = v.l + cross(v.a, c);
  V3 p = vc * mass;
  V3 wl = mulTv(R, v.a);
  V3 Il(wl.x * m.body_inertia[3 * b], wl.y * m.body_inertia[3 * b + 1], wl.z * m.body_inertia[3 * b + 2]);
  V3 Lc = mulv(R, Il);
  return {Lc + cross(c, p), p};
}

// recursive Newton-Euler about the world origin; fills bvel / bacc and returns generalized force
// M*qacc_in + bias (qacc_in may be null -> pure bias).  Gravity enters as a base acceleration.
void rne(Sim& s, const double* qacc_in, std::vector<double>& out) {
  const ModelView& m = *s.m;
  int nb = m.nbody;
  std::vector<Sp> vel(nb), acc(nb), frc(nb);
  acc[0] = {V3(), V3(-m.gravity[0], -m.gravity[1], -m.gravity[2])};
  for (int b = 1; b < nb; b++) {
    int p = m.body_parentid[b];
    Sp v = vel[p], a = acc[p];
    for (int d = m.body_dofadr[b]; d >= 0 && d < m.body_dofadr[b] + m.body_dofnum[b]; d++) {
      Sp S = dof_subspace(s, d);
      v = sp_add(v, sp_scale(S, s.qvel[d]));
    }
    vel[b] = v;
    for (int d = m.body_dofadr[b]; d >= 0 && d < m.body_dofadr[b] + m.body_dofnum[b]; d++) {
      Sp S = dof_subspace(s, d);
      if (qacc_in) a = sp_add(a, sp_scale(S, qacc_in[d]));
      if (!dof_world_fixed(m, d)) {
        // frame carrying the axis: for a free joint the whole body; for hinge/slide joint j of a
        // multi-joint body the partial chain up to and including j.  With one joint per body
        // (every level of the reference) both are the full body velocity.
        Sp vf = vel[p];
        int j = m.dof_jntid[d];
        if (m.jnt_type[j] == MJB_JNT_FREE) vf = vel[b];
        else for (int d2 = m.body_dofadr[b]; d2 <= d; d2++) vf = sp_add(vf, sp_scale(dof_subspace(s, d2), s.qvel[d2]));
        a = sp_add(a, sp_scale(sp_cross_motion(vf, S), s.qvel[d]));
      }
    }
    acc[b] = a;
    Sp Iv = inertia_apply(s, b, vel[b]);
    frc[b] = sp_add(inertia_apply(s, b, acc[b]), sp_cross_force(vel[b], Iv));
  }
  for (int b = 0; b < nb; b++) {
    for (int i = 0; i < 3; i++) {
      s.bvel[6 * b + i] = vel[b].a[i]; s.bvel[6 * b + 3 + i] = vel[b].l[i];
      s.bacc[6 * b + i] = acc[b].a[i]; s.bacc[6 * b + 3 + i] = acc[b].l[i];
    }
  }
  out.assign(m.nv, 0.0);
  for (int b = nb - 1; b >= 1; b--) {
    for (int d = m.body_dofadr[b]; d >= 0 && d < m.body_dofadr[b] + m.body_dofnum[b]; d++)
      out[d] = sp_dot(dof_subspace(s, d), frc[b]);
    int p = m.body_parentid[b];
    frc[p] = sp_add(frc[p], frc[b]);
  }
}

// ---------------------------------------------------------------------------------------------
// narrow phase.  Normal points from geom1 to geom2; pos is the mid-point between the surfaces.
void make_frame(Contact& c) {
  V3 n = normalized(c.frame[0]);
  V3 t = c.frame[1];
  if (norm(t) < 0.5) {
    t = V3();
    if (n.y < 0.5 && n.y > -0.5) t.y = 1; else t.z = 1;
  }
  t = t - n * dot(n, t);
  t = normalized(t);
  c.frame[0] = n; c.frame[1] = t; c.frame[2] = cross(n, t);
}

int plane_sphere(V3 pp, V3 n, V3 c, double r, double margin, Contact* out) {
  double cdist = dot(c - pp, n), dist = cdist - r;
  if (dist > margin) return 0;
  out->dist = dist;
  out->pos = c - n * (r + dist * 0.5);
  out->frame[0] = n; out->frame[1] = V3();
  return 1;
}
int sphere_sphere(V3 c1, double r1, V3 c2, double r2, double margin, Contact* out) {
  V3 d = c2 - c1;
  double cd = norm(d), dist = cd - r1 - r2;
  if (dist > margin) return 0;
  V3 n = cd < kMinVal ? V3(1, 0, 0) : d * (1.0 / cd);
  out->dist = dist;
  out->pos = c1 + n * (r1 + dist * 0.5);
  out->frame[0] = n; out->frame[1] = V3();
  return 1;
}
int sphere_box(V3 c1, double r, V3 c2, const M3& R2, const double* size, double margin, Contact* out) {
  V3 cl = mulTv(R2, c1 - c2), cp;
  bool inside = true;
  for (int i = 0; i < 3; i++) {
    cp[i] = std::min(size[i], std::max(-size[i], cl[i]));
    if (cp[i] != cl[i]) inside = false;
  }
  V3 nl;
  double dist;
  if (inside) {
    int k = 0;
    double best = 1e300;
    for (int i = 0; i < 3; i++) {
      double pen = size[i] - std::fabs(cl[i]);
      if (pen < best) { best = pen; k = i; }
    }
    double sgn = cl[k] >= 0 ? 1.0 : -1.0;
    nl[k] = -sgn;
    dist = -best - r;
  } else {
    V3 d = cl - cp;
    double len = norm(d);
    dist = len - r;
    if (dist > margin) return 0;
    nl = d * (-1.0 / len);
  }
  V3 n = mulv(R2, nl);
  out->dist = dist;
  out->pos = c1 + n * (r + dist * 0.5);
  out->frame[0] = n; out->frame[1] = V3();
  return 1;
}
// signed distance from a point (box frame) to the box: Euclidean outside, -(depth to the nearest face) inside
double point_box_dist(V3 p, const double* size) {
  double s2 = 0, inside_pen = 1e300;
  for (int i = 0; i < 3; i++) {
    double e = std::fabs(p[i]) - size[i];
    if (e > 0) s2 += e * e;
    inside_pen = std::min(inside_pen, -e);
  }
  return s2 > 0 ? std::sqrt(s2) : -inside_pen;
}

// capsule - box.  MuJoCo's mjc_CapsuleBox (engine_collision_box.c of the un-vendored mujoco==2.3.3) places
// at most two spheres on the capsule axis and hands each to sphere - box: one at the point of the axis closest
// to the box, and a second "clamped to the farthest point of the capsule that is above the box" when the
// closest feature is a face.  Specification used here and by the CUDA kernel:
//   1. t* = argmin over the segment of the signed distance to the box;
//   2. if that point is inside the box: one contact there;
//   3. else let k be the dominant axis of (point - closest box point), i.e. the closest FACE, and I the part of
//      the segment whose projection lies inside that face's rectangle.  If t* lies in I (within 0.05 of the
//      half length): contacts at the two ends of I, deeper first (the distance is linear over I, so the deeper
//      end IS the closest point); ends closer than 1e-3 half lengths count once.  Otherwise (edge / vertex
//      feature away from the face): one contact at t*.
// Every sphere is margin-tested by sphere - box.  The minimiser is found EXACTLY here: the signed distance along
// the segment is convex and piecewise quadratic / linear, with breakpoints where a coordinate crosses a face
// plane or where two inside depths tie, so its minimum is at a breakpoint, an end, or the stationary point of
// one quadratic piece.  (The kernel uses a golden-section search on the same function.)
int capsule_box(V3 c1, const M3& R1, const double* size1, V3 c2, const M3& R2, const double* size2, double margin,
                Contact* out) {
  const double r = size1[0], h = size1[1];
  const V3 ax = col(R1, 2);
  const V3 c = mulTv(R2, c1 - c2), d = mulTv(R2, ax) * h;   // p(t) = c + t d in the box frame, t in [-1, 1]
  auto phi = [&](double t) { return point_box_dist(c + d * t, size2); };
  std::vector<double> cand = {-1.0, 1.0};
  auto add = [&](double t) { if (t > -1 && t < 1 && std::isfinite(t)) cand.push_back(t); };
  // face-plane crossings: c_k + t d_k = +- s_k
  for (int k = 0; k < 3; k++)
    if (std::fabs(d[k]) > 1e-300) { add((size2[k] - c[k]) / d[k]); add((-size2[k] - c[k]) / d[k]); }
  // ties of two inside depths: sa (c_a + t d_a) - s_a = sb (c_b + t d_b) - s_b
  for (int a = 0; a < 3; a++)
    for (int b = a + 1; b < 3; b++)
      for (int sa = -1; sa <= 1; sa += 2)
        for (int sb = -1; sb <= 1; sb += 2) {
          double den = sa * d[a] - sb * d[b];
          if (std::fabs(den) > 1e-300) add((sb * c[b] - size2[b] - sa * c[a] + size2[a]) / den);
        }
  // stationary points of the quadratic pieces: between consecutive breakpoints the set of violated coordinates and
  // their signs are fixed, f^2 = sum_k (sg_k (c_k + t d_k) - s_k)^2 over that set
  std::vector<double> brk = cand;
  std::sort(brk.begin(), brk.end());
  for (size_t i = 0; i + 1 < brk.size(); i++) {
    double a = brk[i], b = brk[i + 1];
    if (b - a < 1e-14) continue;
    V3 pm = c + d * (0.5 * (a + b));
    double num = 0, den = 0;
    for (int k = 0; k < 3; k++) {
      double sg = pm[k] > size2[k] ? 1.0 : (pm[k] < -size2[k] ? -1.0 : 0.0);
      if (sg == 0) continue;
      num += sg * d[k] * (sg * c[k] - size2[k]);
      den += d[k] * d[k];
    }
    if (den > 0) { double t = -num / den; if (t > a && t < b) cand.push_back(t); }
  }
  std::sort(cand.begin(), cand.end());
  double ts = cand[0], best = phi(cand[0]);
  for (double t : cand) { double f = phi(t); if (f < best) { best = f; ts = t; } }
  if (best > margin + r) return 0;
  auto sphere_at = [&](double t, Contact* o) { return sphere_box(c1 + ax * (t * h), r, c2, R2, size2, margin, o); };
  if (best <= 0) return sphere_at(ts, out);
  // closest face and the part of the segment over its rectangle
  V3 p = c + d * ts, sep;
  for (int k = 0; k < 3; k++) sep[k] = p[k] - std::min(size2[k], std::max(-size2[k], p[k]));
  int kf = 0;
  for (int k = 1; k < 3; k++) if (std::fabs(sep[k]) > std::fabs(sep[kf])) kf = k;
  double lo = -1, hi = 1;
  for (int j = 0; j < 3; j++) {
    if (j == kf) continue;
    if (std::fabs(d[j]) < 1e-12) { if (std::fabs(c[j]) > size2[j]) { lo = 1; hi = -1; } continue; }
    double t0 = (-size2[j] - c[j]) / d[j], t1 = (size2[j] - c[j]) / d[j];
    lo = std::max(lo, std::min(t0, t1)); hi = std::min(hi, std::max(t0, t1));
  }
  if (lo > hi || ts < lo - 0.05 || ts > hi + 0.05) return sphere_at(ts, out);
  double first = phi(lo) <= phi(hi) ? lo : hi, second = first == lo ? hi : lo;
  int n = sphere_at(first, out);
  if (hi - lo > 1e-3) n += sphere_at(second, out + n);
  return n;
}
int plane_box(V3 pp, V3 n, V3 c2, const M3& R2, const double* size, double margin, Contact* out) {
  double dist0 = dot(c2 - pp, n);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    V3 vec(size[0] * ((i & 1) ? 1 : -1), size[1] * ((i & 2) ? 1 : -1), size[2] * ((i & 4) ? 1 : -1));
    V3 corner = mulv(R2, vec);
    double ld = dot(n, corner);
    if (dist0 + ld > margin || ld > 0) continue;
    double cd = dist0 + ld;
    out[cnt].dist = cd;
    out[cnt].pos = corner + c2 - n * (cd * 0.5);
    out[cnt].frame[0] = n; out[cnt].frame[1] = V3();
    if (++cnt >= 4) break;
  }
  return cnt;
}
// box - box by the separating-axis test + face clipping / edge - edge closest points (the construction behind
// MuJoCo's mjc_BoxBox; un-vendored, restated from its documented behaviour: up to 8 contacts for face contact,
// 1 for edge - edge, normal from geom1 to geom2, position midway between the surfaces).  Specification shared
// with the CUDA kernel:
//   * sep(L) = |L . (c2 - c1)| - r1(L) - r2(L) for the 6 face normals and the 9 unit edge cross products
//     (pairs of edges with |cross| < 1e-6 are skipped); any sep > margin: no contact;
//   * the contact axis is the face axis of largest sep (lowest index on ties) unless an edge axis beats it by
//     more than 5 % of |sep| + 1e-6 (faces preferred, as in ODE-lineage implementations);
//   * face axis: the incident face (most anti-parallel face of the other box) is clipped against the four side
//     planes of the reference face; every vertex of the clipped polygon within `margin` of the reference plane
//     is a contact with dist = its signed distance to that plane;
//   * edge axis: one contact at the closest points of the two supporting edges, dist = sep.
// Here the clipping is Sutherland - Hodgman on the polygon (sequential); the kernel enumerates the candidate
// vertices of the same polygon in parallel.
int box_box(V3 c1, const M3& R1, const double* s1, V3 c2, const M3& R2, const double* s2, double margin, Contact* out) {
  const V3 dc = c2 - c1;
  V3 A[3], B[3];
  for (int i = 0; i < 3; i++) { A[i] = col(R1, i); B[i] = col(R2, i); }
  auto radius = [&](V3 L, const V3* ax, const double* s) { return s[0] * std::fabs(dot(L, ax[0])) + s[1] * std::fabs(dot(L, ax[1])) + s[2] * std::fabs(dot(L, ax[2])); };
  auto sep_of = [&](V3 L) { return std::fabs(dot(L, dc)) - radius(L, A, s1) - radius(L, B, s2); };
  double fsep = -1e300, esep = -1e300;
  int fax = -1, eax = -1;
  for (int i = 0; i < 6; i++) {
    double sp = sep_of(i < 3 ? A[i] : B[i - 3]);
    if (sp > margin) return 0;
    if (sp > fsep) { fsep = sp; fax = i; }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      V3 L = cross(A[i], B[j]);
      double len = norm(L);
      if (len < 1e-6) continue;
      double sp = sep_of(L * (1.0 / len));
      if (sp > margin) return 0;
      if (sp > esep) { esep = sp; eax = 3 * i + j; }
    }
  if (eax >= 0 && esep > fsep + 0.05 * std::fabs(fsep) + 1e-6) {
    const int i = eax / 3, j = eax % 3;
    V3 L = normalized(cross(A[i], B[j]));
    if (dot(L, dc) < 0) L = L * -1.0;
    V3 pa = c1, pb = c2;   // points on the supporting edges: farthest along +L on box 1, along -L on box 2
    for (int k = 0; k < 3; k++) {
      if (k != i) pa = pa + A[k] * (s1[k] * (dot(L, A[k]) >= 0 ? 1.0 : -1.0));
      if (k != j) pb = pb - B[k] * (s2[k] * (dot(L, B[k]) >= 0 ? 1.0 : -1.0));
    }
    // closest points of the lines pa + a A_i, pb + b B_j, parameters clamped to the edges
    const V3 w = pa - pb;
    const double ab = dot(A[i], B[j]), den = 1 - ab * ab;
    double a = (ab * dot(B[j], w) - dot(A[i], w)) / den, b = (dot(B[j], w) - ab * dot(A[i], w)) / den;
    a = std::min(s1[i], std::max(-s1[i], a)); b = std::min(s2[j], std::max(-s2[j], b));
    const V3 qa = pa + A[i] * a, qb = pb + B[j] * b;
    out[0].dist = esep; out[0].pos = (qa + qb) * 0.5; out[0].frame[0] = L; out[0].frame[1] = V3();
    return 1;
  }
  // face contact: reference box owns the axis
  const bool ref1 = fax < 3;
  const int ri = ref1 ? fax : fax - 3;
  const V3* RA = ref1 ? A : B; const V3* RB = ref1 ? B : A;
  const double* sa = ref1 ? s1 : s2; const double* sb = ref1 ? s2 : s1;
  const V3 ca = ref1 ? c1 : c2, cb = ref1 ? c2 : c1;
  V3 nA = RA[ri];
  if (dot(nA, cb - ca) < 0) nA = nA * -1.0;
  int bj = 0;
  for (int k = 1; k < 3; k++) if (std::fabs(dot(nA, RB[k])) > std::fabs(dot(nA, RB[bj]))) bj = k;
  const V3 mB = RB[bj] * (dot(nA, RB[bj]) > 0 ? -1.0 : 1.0);   // outward normal of the incident face
  const int bu = (bj + 1) % 3, bw = (bj + 2) % 3, au = (ri + 1) % 3, aw = (ri + 2) % 3;
  std::vector<V3> poly;
  const double sgn[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int q = 0; q < 4; q++) poly.push_back(cb + mB * sb[bj] + RB[bu] * (sgn[q][0] * sb[bu]) + RB[bw] * (sgn[q][1] * sb[bw]));
  for (int side = 0; side < 4; side++) {
    const V3 e = RA[side < 2 ? au : aw] * (side % 2 ? -1.0 : 1.0);
    const double lim = sa[side < 2 ? au : aw];
    std::vector<V3> nxt;
    for (size_t q = 0; q < poly.size(); q++) {
      const V3 p0 = poly[q], p1 = poly[(q + 1) % poly.size()];
      const double f0 = dot(p0 - ca, e) - lim, f1 = dot(p1 - ca, e) - lim;
      if (f0 <= 0) nxt.push_back(p0);
      if ((f0 < 0 && f1 > 0) || (f0 > 0 && f1 < 0)) nxt.push_back(p0 + (p1 - p0) * (f0 / (f0 - f1)));
    }
    poly.swap(nxt);
    if (poly.empty()) break;
  }
  int cnt = 0;
  for (size_t q = 0; q < poly.size() && cnt < 8; q++) {
    bool dup = false;
    for (size_t k = 0; k < q; k++) if (norm(poly[q] - poly[k]) < 1e-9) dup = true;
    if (dup) continue;
    const double dist = dot(poly[q] - ca, nA) - sa[ri];
    if (dist > margin) continue;
    out[cnt].dist = dist;
    out[cnt].pos = poly[q] - nA * (0.5 * dist);
    out[cnt].frame[0] = ref1 ? nA : nA * -1.0;
    out[cnt].frame[1] = V3();
    cnt++;
  }
  return cnt;
}

int narrowphase(const Sim& s, int g1, int g2, double margin, Contact* out) {
  const ModelView& m = *s.m;
  int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
  V3 p1 = gpos(s, g1), p2 = gpos(s, g2);
  M3 R1 = gmat(s, g1), R2 = gmat(s, g2);
  const double *s1 = &m.geom_size[3 * g1], *s2 = &m.geom_size[3 * g2];
  if (t1 == MJB_GEOM_PLANE) {
    V3 n = col(R1, 2);
    if (t2 == MJB_GEOM_SPHERE) return plane_sphere(p1, n, p2, s2[0], margin, out);
    if (t2 == MJB_GEOM_CAPSULE) {
      V3 ax = col(R2, 2);
      int k = plane_sphere(p1, n, p2 + ax * s2[1], s2[0], margin, out);
      k += plane_sphere(p1, n, p2 - ax * s2[1], s2[0], margin, out + k);
      for (int i = 0; i < k; i++) out[i].frame[1] = ax;
      return k;
    }
    if (t2 == MJB_GEOM_BOX) return plane_box(p1, n, p2, R2, s2, margin, out);
  } else if (t1 == MJB_GEOM_SPHERE) {
    if (t2 == MJB_GEOM_SPHERE) return sphere_sphere(p1, s1[0], p2, s2[0], margin, out);
    if (t2 == MJB_GEOM_CAPSULE) {
      V3 ax = col(R2, 2);
      double x = std::min(s2[1], std::max(-s2[1], dot(ax, p1 - p2)));
      return sphere_sphere(p1, s1[0], p2 + ax * x, s2[0], margin, out);
    }
    if (t2 == MJB_GEOM_BOX) return sphere_box(p1, s1[0], p2, R2, s2, margin, out);
  } else if (t1 == MJB_GEOM_CAPSULE) {
    if (t2 == MJB_GEOM_CAPSULE) {
      V3 a1 = col(R1, 2), a2 = col(R2, 2), dif = p1 - p2;
      double len1 = s1[1], len2 = s2[1];
      double ma = 1, mb = -dot(a1, a2), mc = 1, u = -dot(a1, dif), v = dot(a2, dif), det = ma * mc - mb * mb;
      if (std::fabs(det) >= kMinVal) {
        double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
        if (x1 > len1) { x1 = len1; x2 = (v - mb * len1) / mc; }
        else if (x1 < -len1) { x1 = -len1; x2 = (v + mb * len1) / mc; }
        if (x2 > len2) { x2 = len2; x1 = (u - mb * len2) / ma; }
        else if (x2 < -len2) { x2 = -len2; x1 = (u + mb * len2) / ma; }
        x1 = std::min(len1, std::max(-len1, x1));
        return sphere_sphere(p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0], margin, out);
      }
      // parallel axes: up to two contacts from the segment ends
      int k = 0;
      double xs[2] = {len1, -len1};
      for (int e = 0; e < 2 && k < 2; e++) {
        double x1 = xs[e], x2 = std::min(len2, std::max(-len2, (v - mb * x1) / mc));
        k += sphere_sphere(p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0], margin, out + k);
      }
      if (k < 2) {
        double ys[2] = {len2, -len2};
        for (int e = 0; e < 2 && k < 2; e++) {
          double x2 = ys[e], x1 = (u - mb * x2) / ma;
          if (x1 > len1 || x1 < -len1) continue;  // already covered by an end of segment 1
          k += sphere_sphere(p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0], margin, out + k);
        }
      }
      return k;
    }
    if (t2 == MJB_GEOM_BOX) return capsule_box(p1, R1, s1, p2, R2, s2, margin, out);
  } else if (t1 == MJB_GEOM_BOX && t2 == MJB_GEOM_BOX) {
    return box_box(p1, R1, s1, p2, R2, s2, margin, out);
  }
  return 0;
}

void collision(Sim& s) {
  const ModelView& m = *s.m;
  s.contacts.clear();
  Contact tmp[8];
  for (int k = 0; k < m.npair; k++) {
    int g1 = m.pair_geom1[k], g2 = m.pair_geom2[k];
    double margin = m.pair_margin[k];
    // bounding-sphere cull (conservative; the narrow phase decides)
    if (m.geom_type[g1] == MJB_GEOM_PLANE) {
      V3 n = col(gmat(s, g1), 2);
      if (dot(gpos(s, g2) - gpos(s, g1), n) > m.geom_rbound[g2] + margin) continue;
    } else if (m.geom_type[g2] == MJB_GEOM_BOX) {
      // conservative cull against the oriented box itself (bounding spheres of long boxes are useless)
      if (point_box_dist(mulTv(gmat(s, g2), gpos(s, g1) - gpos(s, g2)), &m.geom_size[3 * g2]) > m.geom_rbound[g1] + margin) continue;
    } else {
      if (norm(gpos(s, g2) - gpos(s, g1)) > m.geom_rbound[g1] + m.geom_rbound[g2] + margin) continue;
    }
    int n = narrowphase(s, g1, g2, margin, tmp);
    for (int i = 0; i < n; i++) {
      Contact c = tmp[i];
      make_frame(c);
      c.g1 = g1; c.g2 = g2; c.pair = k;
      c.includemargin = m.pair_includemargin[k];
      for (int a = 0; a < 3; a++) c.friction[a] = m.pair_friction[3 * k + a];
      for (int a = 0; a < 2; a++) c.solref[a] = m.pair_solref[2 * k + a];
      for (int a = 0; a < 5; a++) c.solimp[a] = m.pair_solimp[5 * k + a];
      c.condim = m.pair_condim[k];
      s.contacts.push_back(c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// constraints
double impedance(const double* solimp, double pos, double margin) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin == dmax || width <= kMinVal) return 0.5 * (dmin + dmax);
  double x = (pos - margin) / width;
  if (x < 0) x = -x;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
  else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}
void kb_from_solref(const ModelView& m, const double* solref, const double* solimp, double& K, double& B) {
  double dmax = solimp[1];
  if (solref[0] > 0) {
    double tc = std::max(solref[0], 2 * m.timestep), dr = solref[1];
    K = 1.0 / std::max(kMinVal, dmax * dmax * tc * tc * dr * dr);
    B = 2.0 / std::max(kMinVal, dmax * tc);
  } else {
    K = -solref[0] / std::max(kMinVal, dmax * dmax);
    B = -solref[1] / std::max(kMinVal, dmax);
  }
}

void make_constraints(Sim& s) {
  const ModelView& m = *s.m;
  int nv = m.nv;
  s.nefc = 0;
  s.J.clear(); s.efc_pos.clear(); s.efc_margin.clear(); s.efc_D.clear(); s.efc_R.clear(); s.efc_aref.clear();
  s.efc_type.clear(); s.efc_id.clear(); s.efc_diagApprox.clear();
  std::vector<double> K, Bv, imp;
  auto add_row = [&](const std::vector<double>& jrow, double pos, double margin, int type, int id, double dA,
                     const double* solref, const double* solimp) {
    s.J.insert(s.J.end(), jrow.begin(), jrow.end());
    s.efc_pos.push_back(pos); s.efc_margin.push_back(margin); s.efc_type.push_back(type); s.efc_id.push_back(id);
    s.efc_diagApprox.push_back(dA);
    double k, b;
    kb_from_solref(m, solref, solimp, k, b);
    double d = impedance(solimp, pos, margin);
    K.push_back(k); Bv.push_back(b); imp.push_back(d);
    s.efc_R.push_back(std::max(kMinVal, (1 - d) * dA / d));
    s.nefc++;
  };
  // joint limits (joint order; lower side first)
  std::vector<double> row(nv);
  for (int j = 0; j < m.njnt; j++) {
    if (!m.jnt_limited[j] || m.jnt_type[j] == MJB_JNT_FREE) continue;
    double value = s.qpos[m.jnt_qposadr[j]], margin = m.jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m.jnt_range[2 * j + (side + 1) / 2] - value);
      if (dist < margin) {
        std::fill(row.begin(), row.end(), 0.0);
        row[m.jnt_dofadr[j]] = -side;
        add_row(row, dist, margin, 0, j, m.dof_invweight0[m.jnt_dofadr[j]], &m.jnt_solref[2 * j], &m.jnt_solimp[5 * j]);
      }
    }
  }
  // contacts
  std::vector<double> jp1, jr1, jp2, jr2;
  for (size_t ci = 0; ci < s.contacts.size(); ci++) {
    Contact& c = s.contacts[ci];
    c.efc_address = s.nefc;
    int b1 = m.geom_bodyid[c.g1], b2 = m.geom_bodyid[c.g2];
    point_jacobian(m, s.kin, b1, c.pos, jp1, jr1);
    point_jacobian(m, s.kin, b2, c.pos, jp2, jr2);
    std::vector<double> jn(nv), jt1(nv), jt2(nv);
    for (int d = 0; d < nv; d++) {
      V3 dif(jp2[d] - jp1[d], jp2[nv + d] - jp1[nv + d], jp2[2 * nv + d] - jp1[2 * nv + d]);
      jn[d] = dot(c.frame[0], dif); jt1[d] = dot(c.frame[1], dif); jt2[d] = dot(c.frame[2], dif);
    }
    double tran = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
    if (c.condim == 1) {
      add_row(jn, c.dist, c.includemargin, 2, (int)ci, tran, c.solref, c.solimp);
    } else {
      int first = s.nefc;
      for (int k = 0; k < 2; k++) {
        double mu = c.friction[0];  // both tangential directions use the sliding coefficient
        const std::vector<double>& jt = k == 0 ? jt1 : jt2;
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int d = 0; d < nv; d++) row[d] = jn[d] + sgn * mu * jt[d];
          add_row(row, c.dist, c.includemargin, 1, (int)ci, tran + mu * mu * tran, c.solref, c.solimp);
        }
      }
      // pyramidal: one common regulariser for all edges, derived from the first row
      double mu = c.friction[0];
      double Rpy = 2 * mu * mu * s.efc_R[first];
      for (int k = 0; k < 4; k++) s.efc_R[first + k] = Rpy;
    }
  }
  // reference acceleration
  s.efc_D.resize(s.nefc); s.efc_aref.resize(s.nefc); s.efc_force.assign(s.nefc, 0.0);
  for (int i = 0; i < s.nefc; i++) {
    double vel = 0;
    for (int d = 0; d < nv; d++) vel += s.J[(size_t)i * nv + d] * s.qvel[d];
    s.efc_D[i] = 1.0 / s.efc_R[i];
    s.efc_aref[i] = -Bv[i] * vel - K[i] * imp[i] * (s.efc_pos[i] - s.efc_margin[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// solvers.  Primal problem: min_a 1/2 (a - a0)' M (a - a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
double primal_cost(const Sim& s, const std::vector<double>& a) {
  int nv = s.nv;
  double c = 0;
  for (int i = 0; i < nv; i++) {
    double ri = 0;
    for (int j = 0; j < nv; j++) ri += s.M[(size_t)i * nv + j] * (a[j] - s.qacc_smooth[j]);
    c += 0.5 * ri * (a[i] - s.qacc_smooth[i]);
  }
  for (int r = 0; r < s.nefc; r++) {
    double jar = -s.efc_aref[r];
    for (int d = 0; d < nv; d++) jar += s.J[(size_t)r * nv + d] * a[d];
    if (jar < 0) c += 0.5 * s.efc_D[r] * jar * jar;
  }
  return c;
}

// exact minimiser of the convex piecewise-quadratic phi(alpha) along `sv` (alpha >= 0)
double exact_linesearch(const Sim& s, const std::vector<double>& jar, const std::vector<double>& jv, double p1, double p2) {
  // phi'(alpha) = p1 + alpha p2 + sum_{jar_i + alpha jv_i < 0} D_i (jar_i + alpha jv_i) jv_i
  std::vector<double> bp;
  for (int i = 0; i < s.nefc; i++)
    if (jv[i] != 0) {
      double t = -jar[i] / jv[i];
      if (t > 0) bp.push_back(t);
    }
  std::sort(bp.begin(), bp.end());
  bp.push_back(1e300);
  double lo = 0;
  for (size_t k = 0; k < bp.size(); k++) {
    double hi = bp[k];
    double mid = (hi < 1e299) ? 0.5 * (lo + hi) : lo + 1.0;
    // active set on (lo, hi) evaluated at the interval's interior
    double c0 = p1, c1 = p2;
    for (int i = 0; i < s.nefc; i++)
      if (jar[i] + mid * jv[i] < 0) { c0 += s.efc_D[i] * jar[i] * jv[i]; c1 += s.efc_D[i] * jv[i] * jv[i]; }
    if (c1 > 0) {
      double a = -c0 / c1;
      if (a <= hi) return std::max(a, lo);
    }
    lo = hi;
  }
  return 0;
}

void solve_newton(Sim& s, bool fixed_iters) {
  int nv = s.nv, ne = s.nefc;
  std::vector<double>& a = s.qacc;
  // warm start: whichever of qacc_warmstart / qacc_smooth has the lower cost
  a = s.qacc_smooth;
  if (s.warmstart_mode == 2) a = s.qacc_warmstart;  // always (what the CUDA path does)
  else if (s.warmstart) {
    if (primal_cost(s, s.qacc_warmstart) < primal_cost(s, s.qacc_smooth)) a = s.qacc_warmstart;
  }
  std::vector<double> jar(ne), jv(ne), grad(nv), sv(nv), H((size_t)nv * nv), Mv(nv), Ma(nv);
  int maxit = fixed_iters ? s.solver_iters : 200;
  int it = 0;
  for (; it < maxit; it++) {
    for (int r = 0; r < ne; r++) {
      double x = -s.efc_aref[r];
      for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * a[d];
      jar[r] = x;
    }
    for (int i = 0; i < nv; i++) {
      double x = 0;
      for (int j = 0; j < nv; j++) x += s.M[(size_t)i * nv + j] * a[j];
      Ma[i] = x;
    }
    for (int i = 0; i < nv; i++) grad[i] = Ma[i] - s.qfrc_smooth[i];
    H = s.M;
    for (int r = 0; r < ne; r++) {
      if (jar[r] >= 0) { s.efc_force[r] = 0; continue; }
      double f = -s.efc_D[r] * jar[r];
      s.efc_force[r] = f;
      const double* Jr = &s.J[(size_t)r * nv];
      for (int i = 0; i < nv; i++) {
        grad[i] -= Jr[i] * f;
        if (Jr[i] != 0)
          for (int j = 0; j < nv; j++) H[(size_t)i * nv + j] += s.efc_D[r] * Jr[i] * Jr[j];
      }
    }
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
    if (!fixed_iters && std::sqrt(gn) < 1e-11) break;
    if (!cholesky(H, nv)) break;
    for (int i = 0; i < nv; i++) sv[i] = -grad[i];
    cholesky_solve(H, nv, sv.data());
    for (int r = 0; r < ne; r++) {
      double x = 0;
      for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * sv[d];
      jv[r] = x;
    }
    double p1 = 0, p2 = 0;
    for (int i = 0; i < nv; i++) {
      double x = 0;
      for (int j = 0; j < nv; j++) x += s.M[(size_t)i * nv + j] * sv[j];
      Mv[i] = x;
      p1 += sv[i] * (Ma[i] - s.qfrc_smooth[i]);
      p2 += sv[i] * x;
    }
    double alpha = exact_linesearch(s, jar, jv, p1, p2);
    if (alpha <= 0) break;
    for (int i = 0; i < nv; i++) a[i] += alpha * sv[i];
  }
  s.last_solver_iters = it;
  // final forces
  for (int r = 0; r < ne; r++) {
    double x = -s.efc_aref[r];
    for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * a[d];
    s.efc_force[r] = x < 0 ? -s.efc_D[r] * x : 0;
  }
}

void solve_pgs(Sim& s, bool fixed_iters) {
  int nv = s.nv, ne = s.nefc;
  // AR = J M^-1 J' + R ; b = J a0 - aref
  std::vector<double> MinvJt((size_t)ne * nv), AR((size_t)ne * ne), b(ne);
  for (int r = 0; r < ne; r++) {
    std::vector<double> x(s.J.begin() + (size_t)r * nv, s.J.begin() + (size_t)(r + 1) * nv);
    cholesky_solve(s.L, nv, x.data());
    for (int d = 0; d < nv; d++) MinvJt[(size_t)r * nv + d] = x[d];
  }
  for (int r = 0; r < ne; r++) {
    for (int c = 0; c < ne; c++) {
      double x = 0;
      for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * MinvJt[(size_t)c * nv + d];
      AR[(size_t)r * ne + c] = x;
    }
    AR[(size_t)r * ne + r] += s.efc_R[r];
    double x = -s.efc_aref[r];
    for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * s.qacc_smooth[d];
    b[r] = x;
  }
  std::vector<double>& f = s.efc_force;
  f.assign(ne, 0.0);
  if (s.warmstart) {
    for (int r = 0; r < ne; r++) {
      double x = -s.efc_aref[r];
      for (int d = 0; d < nv; d++) x += s.J[(size_t)r * nv + d] * s.qacc_warmstart[d];
      f[r] = x < 0 ? -s.efc_D[r] * x : 0;
    }
    // keep the warm start only if its dual cost beats zero force
    double cost = 0;
    for (int r = 0; r < ne; r++) {
      double x = 0;
      for (int c = 0; c < ne; c++) x += AR[(size_t)r * ne + c] * f[c];
      cost += f[r] * (0.5 * x + b[r]);
    }
    if (cost > 0) f.assign(ne, 0.0);
  }
  int maxit = fixed_iters ? s.solver_iters : 100000;
  int it = 0;
  for (; it < maxit; it++) {
    double change = 0;
    for (int r = 0; r < ne; r++) {
      double res = b[r];
      for (int c = 0; c < ne; c++) res += AR[(size_t)r * ne + c] * f[c];
      double nf = std::max(0.0, f[r] - res / AR[(size_t)r * ne + r]);
      change = std::max(change, std::fabs(nf - f[r]));
      f[r] = nf;
    }
    if (!fixed_iters && change < 1e-13) break;
  }
  s.last_solver_iters = it;
  s.qacc = s.qacc_smooth;
  for (int r = 0; r < ne; r++)
    for (int d = 0; d < nv; d++) s.qacc[d] += MinvJt[(size_t)r * nv + d] * f[r];
}

// ---------------------------------------------------------------------------------------------
// ray casting (rangefinder, touch)
double ray_plane(V3 pos, const M3& R, const double* size, V3 pnt, V3 vec) {
  V3 lp = mulTv(R, pnt - pos), lv = mulTv(R, vec);
  if (lv.z > -kMinVal) return -1;
  double x = -lp.z / lv.z;
  if (x < 0) return -1;
  double p0 = lp.x + x * lv.x, p1 = lp.y + x * lv.y;
  if ((size[0] <= 0 || std::fabs(p0) <= size[0]) && (size[1] <= 0 || std::fabs(p1) <= size[1])) return x;
  return -1;
}
// roots of a x^2 + 2 b x + c = 0; returns count, ascending
int quad_roots(double a, double b, double c, double x[2]) {
  if (a < kMinVal) return 0;
  double det = b * b - a * c;
  if (det < 0) return 0;
  double sq = std::sqrt(det);
  x[0] = (-b - sq) / a; x[1] = (-b + sq) / a;
  return 2;
}
double ray_sphere(V3 pos, double r, V3 pnt, V3 vec) {
  V3 d = pnt - pos;
  double x[2];
  if (!quad_roots(dot(vec, vec), dot(vec, d), dot(d, d) - r * r, x)) return -1;
  if (x[0] >= 0) return x[0];
  if (x[1] >= 0) return x[1];
  return -1;
}
double ray_capsule(V3 pos, const M3& R, const double* size, V3 pnt, V3 vec) {
  V3 lp = mulTv(R, pnt - pos), lv = mulTv(R, vec);
  double r = size[0], h = size[1], best = -1, x[2];
  auto consider = [&](double t) { if (t >= 0 && (best < 0 || t < best)) best = t; };
  // cylinder wall
  if (quad_roots(lv.x * lv.x + lv.y * lv.y, lp.x * lv.x + lp.y * lv.y, lp.x * lp.x + lp.y * lp.y - r * r, x))
    for (int i = 0; i < 2; i++) if (std::fabs(lp.z + x[i] * lv.z) <= h) consider(x[i]);
  // end caps
  for (int sgn = -1; sgn <= 1; sgn += 2) {
    V3 d = lp - V3(0, 0, sgn * h);
    if (quad_roots(dot(lv, lv), dot(lv, d), dot(d, d) - r * r, x))
      for (int i = 0; i < 2; i++) if (sgn * (lp.z + x[i] * lv.z) >= h) consider(x[i]);
  }
  return best;
}
double ray_box(V3 pos, const M3& R, const double* size, V3 pnt, V3 vec) {
  V3 lp = mulTv(R, pnt - pos), lv = mulTv(R, vec);
  double best = -1;
  for (int a = 0; a < 3; a++) {
    if (std::fabs(lv[a]) < kMinVal) continue;
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      double t = (sgn * size[a] - lp[a]) / lv[a];
      if (t < 0) continue;
      int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
      if (std::fabs(lp[a1] + t * lv[a1]) <= size[a1] && std::fabs(lp[a2] + t * lv[a2]) <= size[a2])
        if (best < 0 || t < best) best = t;
    }
  }
  return best;
}
double ray_geom(V3 pos, const M3& R, const double* size, V3 pnt, V3 vec, int type) {
  switch (type) {
    case MJB_GEOM_PLANE: return ray_plane(pos, R, size, pnt, vec);
    case MJB_GEOM_SPHERE: return ray_sphere(pos, size[0], pnt, vec);
    case MJB_GEOM_CAPSULE: return ray_capsule(pos, R, size, pnt, vec);
    case MJB_GEOM_BOX: return ray_box(pos, R, size, pnt, vec);
  }
  return -1;
}

V3 spos(const Sim& s, int t) { return V3(s.site_xpos[3 * t], s.site_xpos[3 * t + 1], s.site_xpos[3 * t + 2]); }
M3 smat(const Sim& s, int t) { M3 R; for (int i = 0; i < 9; i++) R.m[i] = s.site_xmat[9 * t + i]; return R; }

void apply_cutoff(Sim& s, int i) {
  const ModelView& m = *s.m;
  double cut = m.sensor_cutoff[i];
  if (cut <= 0) return;
  for (int k = 0; k < m.sensor_dim[i]; k++) {
    double& x = s.sensordata[m.sensor_adr[i] + k];
    if (m.sensor_datatype[i] == 0) x = std::min(cut, std::max(-cut, x));
    else if (m.sensor_datatype[i] == 1) x = std::min(cut, x);
  }
}

void sensors_pos(Sim& s) {
  const ModelView& m = *s.m;
  for (int i = 0; i < m.nsensor; i++) {
    int t = m.sensor_objid[i], adr = m.sensor_adr[i];
    int type = m.sensor_type[i];
    if (type == MJB_SENS_RANGEFINDER) {
      V3 pnt = spos(s, t), vec = col(smat(s, t), 2);
      int bex = m.site_bodyid[t];
      double best = -1;
      for (int g = 0; g < m.ngeom; g++) {
        if (m.geom_bodyid[g] == bex) continue;
        if (m.geom_rgba[4 * g + 3] == 0) continue;
        double x = ray_geom(gpos(s, g), gmat(s, g), &m.geom_size[3 * g], pnt, vec, m.geom_type[g]);
        if (x >= 0 && (best < 0 || x < best)) best = x;
      }
      s.sensordata[adr] = best;
      apply_cutoff(s, i);
    } else if (type == MJB_SENS_FRAMEXAXIS || type == MJB_SENS_FRAMEYAXIS || type == MJB_SENS_FRAMEZAXIS) {
      V3 ax = col(smat(s, t), type - MJB_SENS_FRAMEXAXIS);
      for (int k = 0; k < 3; k++) s.sensordata[adr + k] = ax[k];
      apply_cutoff(s, i);
    }
  }
}

void sensors_acc(Sim& s) {
  const ModelView& m = *s.m;
  bool need_acc = false;
  for (int i = 0; i < m.nsensor; i++) need_acc |= m.sensor_type[i] == MJB_SENS_ACCELEROMETER;
  if (need_acc) {
    std::vector<double> tmp;
    rne(s, s.qacc.data(), tmp);  // fills bvel / bacc with the post-constraint accelerations
  }
  for (int i = 0; i < m.nsensor; i++) {
    int t = m.sensor_objid[i], adr = m.sensor_adr[i], type = m.sensor_type[i];
    if (type == MJB_SENS_TOUCH) {
      int body = m.site_bodyid[t];
      double total = 0;
      for (const Contact& c : s.contacts) {
        int b1 = m.geom_bodyid[c.g1], b2 = m.geom_bodyid[c.g2];
        if (b1 != body && b2 != body) continue;
        if (c.efc_address < 0) continue;
        double fn = 0;
        int rows = c.condim == 1 ? 1 : 4;
        for (int k = 0; k < rows; k++) fn += s.efc_force[c.efc_address + k];
        if (fn <= kMinVal) continue;
        V3 ray = c.frame[0];
        if (b2 == body) ray = ray * -1.0;
        if (ray_geom(spos(s, t), smat(s, t), &m.site_size[3 * t], c.pos, ray, m.site_type[t]) >= 0) total += fn;
      }
      s.sensordata[adr] = total;
      apply_cutoff(s, i);
    } else if (type == MJB_SENS_ACCELEROMETER) {
      int b = m.site_bodyid[t];
      V3 w(s.bvel[6 * b], s.bvel[6 * b + 1], s.bvel[6 * b + 2]), vO(s.bvel[6 * b + 3], s.bvel[6 * b + 4], s.bvel[6 * b + 5]);
      V3 al(s.bacc[6 * b], s.bacc[6 * b + 1], s.bacc[6 * b + 2]), aO(s.bacc[6 * b + 3], s.bacc[6 * b + 4], s.bacc[6 * b + 5]);
      V3 r = spos(s, t);
      // classical acceleration of the body-fixed point at r from the spatial acceleration about the origin
      V3 acc = aO + cross(al, r) + cross(w, vO + cross(w, r));
      V3 loc = mulTv(smat(s, t), acc);
      for (int k = 0; k < 3; k++) s.sensordata[adr + k] = loc[k];
      apply_cutoff(s, i);
    }
  }
}

// ---------------------------------------------------------------------------------------------
void forward(Sim& s, bool with_sensors) {
  const ModelView& m = *s.m;
  int nv = m.nv;
  kinematics(s);
  mass_matrix(m, s.kin, s.M);
  s.L = s.M;
  cholesky(s.L, nv);
  collision(s);
  if (with_sensors) sensors_pos(s);
  // velocity-dependent terms
  rne(s, nullptr, s.qfrc_bias);
  for (int d = 0; d < nv; d++) s.qfrc_passive[d] = -m.dof_damping[d] * s.qvel[d];
  std::fill(s.qfrc_actuator.begin(), s.qfrc_actuator.end(), 0.0);
  for (int u = 0; u < m.nu; u++) {
    double c = s.ctrl[u];
    if (m.actuator_ctrllimited[u]) c = std::min(m.actuator_ctrlrange[2 * u + 1], std::max(m.actuator_ctrlrange[2 * u], c));
    s.qfrc_actuator[m.jnt_dofadr[m.actuator_trnid[u]]] += m.actuator_gear[u] * c;
  }
  for (int d = 0; d < nv; d++) s.qfrc_smooth[d] = s.qfrc_passive[d] - s.qfrc_bias[d] + s.qfrc_actuator[d];
  s.qacc_smooth = s.qfrc_smooth;
  cholesky_solve(s.L, nv, s.qacc_smooth.data());
  make_constraints(s);
  if (s.nefc == 0) {
    s.qacc = s.qacc_smooth;
  } else if (s.solver_mode == 0) solve_newton(s, false);
  else if (s.solver_mode == 1) solve_newton(s, true);
  else if (s.solver_mode == 2) solve_pgs(s, true);
  else solve_pgs(s, false);
  std::fill(s.qfrc_constraint.begin(), s.qfrc_constraint.end(), 0.0);
  for (int r = 0; r < s.nefc; r++)
    for (int d = 0; d < nv; d++) s.qfrc_constraint[d] += s.J[(size_t)r * nv + d] * s.efc_force[r];
  if (with_sensors) sensors_acc(s);
}

void integrate_pos(const ModelView& m, std::vector<double>& qpos, const double* qvel, double h) {
  for (int j = 0; j < m.njnt; j++) {
    int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == MJB_JNT_FREE) {
      for (int i = 0; i < 3; i++) qpos[qa + i] += h * qvel[da + i];
      V3 w(qvel[da + 3], qvel[da + 4], qvel[da + 5]);
      double ang = norm(w) * h;
      if (ang > 0) {
        // q <- q * exp(w h / 2): the angular velocity of a free joint is expressed in the body frame
        V3 u = normalized(w);
        double sn = std::sin(0.5 * ang), dq[4] = {std::cos(0.5 * ang), u.x * sn, u.y * sn, u.z * sn}, out[4];
        quat_mul(&qpos[qa + 3], dq, out);
        double nn = std::sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2] + out[3] * out[3]);
        for (int i = 0; i < 4; i++) qpos[qa + 3 + i] = out[i] / nn;
      }
    } else {
      qpos[qa] += h * qvel[da];
    }
  }
}

void euler(Sim& s) {
  const ModelView& m = *s.m;
  int nv = m.nv;
  double h = m.timestep;
  bool damped = false;
  for (int d = 0; d < nv; d++) damped |= m.dof_damping[d] > 0;
  std::vector<double> qacc = s.qacc;
  if (damped) {
    // implicit-in-velocity joint damping: (M + h D) a' = qfrc_smooth + qfrc_constraint
    std::vector<double> A = s.M;
    for (int d = 0; d < nv; d++) A[(size_t)d * nv + d] += h * m.dof_damping[d];
    cholesky(A, nv);
    for (int d = 0; d < nv; d++) qacc[d] = s.qfrc_smooth[d] + s.qfrc_constraint[d];
    cholesky_solve(A, nv, qacc.data());
  }
  for (int d = 0; d < nv; d++) s.qvel[d] += h * qacc[d];
  integrate_pos(m, s.qpos, s.qvel.data(), h);
  s.time += h;
  s.qacc_warmstart = s.qacc;
}

void rk4(Sim& s) {
  const ModelView& m = *s.m;
  int nv = m.nv;
  double h = m.timestep, t0 = s.time;
  const double A[3] = {0.5, 0.5, 1.0}, B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  std::vector<double> q0 = s.qpos, v0 = s.qvel;
  std::vector<std::vector<double>> Xv(4), F(4);
  Xv[0] = s.qvel; F[0] = s.qacc;
  for (int i = 1; i < 4; i++) {
    // the tableau is sub-diagonal: stage i starts from stage i-1's derivative scaled by A[i-1]
    s.qpos = q0;
    std::vector<double> dv(nv);
    for (int d = 0; d < nv; d++) dv[d] = A[i - 1] * Xv[i - 1][d];
    integrate_pos(m, s.qpos, dv.data(), h);
    for (int d = 0; d < nv; d++) s.qvel[d] = v0[d] + h * A[i - 1] * F[i - 1][d];
    s.time = t0 + h * A[i - 1];
    forward(s, false);
    Xv[i] = s.qvel; F[i] = s.qacc;
  }
  std::vector<double> dq(nv, 0.0), da(nv, 0.0);
  for (int i = 0; i < 4; i++)
    for (int d = 0; d < nv; d++) { dq[d] += B[i] * Xv[i][d]; da[d] += B[i] * F[i][d]; }
  s.qpos = q0;
  for (int d = 0; d < nv; d++) s.qvel[d] = v0[d] + h * da[d];
  integrate_pos(m, s.qpos, dq.data(), h);
  s.time = t0 + h;
  s.qacc_warmstart = s.qacc;
}

}  // namespace

// =================================================================================================
extern "C" {

void* orc_create(const void* blob, int64_t nbytes) {
  try {
    Sim* s = new Sim();
    s->blob.assign((const uint8_t*)blob, (const uint8_t*)blob + nbytes);
    s->m = new ModelView(s->blob.data());
    const ModelView& m = *s->m;
    s->nq = m.nq; s->nv = m.nv; s->nu = m.nu;
    s->qpos.assign(m.qpos0, m.qpos0 + m.nq);
    for (auto* v : {&s->qvel, &s->qacc, &s->qacc_warmstart, &s->qacc_smooth, &s->qfrc_smooth, &s->qfrc_bias,
                    &s->qfrc_passive, &s->qfrc_actuator, &s->qfrc_constraint})
      v->assign(m.nv, 0.0);
    s->ctrl.assign(m.nu, 0.0);
    s->sensordata.assign(std::max(1, m.nsensordata), 0.0);
    s->xpos.assign(3 * m.nbody, 0.0); s->xmat.assign(9 * m.nbody, 0.0); s->xipos.assign(3 * m.nbody, 0.0);
    s->geom_xpos.assign(3 * std::max(1, m.ngeom), 0.0); s->geom_xmat.assign(9 * std::max(1, m.ngeom), 0.0);
    s->site_xpos.assign(3 * std::max(1, m.nsite), 0.0); s->site_xmat.assign(9 * std::max(1, m.nsite), 0.0);
    s->bvel.assign(6 * m.nbody, 0.0); s->bacc.assign(6 * m.nbody, 0.0);
    for (int b = 0; b < m.nbody; b++) { s->xmat[9 * b] = s->xmat[9 * b + 4] = s->xmat[9 * b + 8] = 1; }
    return s;
  } catch (...) {
    return nullptr;
  }
}
void orc_destroy(void* p) {
  Sim* s = (Sim*)p;
  if (!s) return;
  delete s->m;
  delete s;
}
void orc_set_solver(void* p, int mode, int iters, int warmstart) {
  Sim* s = (Sim*)p;
  s->solver_mode = mode; s->solver_iters = iters; s->warmstart = warmstart != 0; s->warmstart_mode = warmstart;
}
void orc_reset(void* p) {
  Sim* s = (Sim*)p;
  const ModelView& m = *s->m;
  s->qpos.assign(m.qpos0, m.qpos0 + m.nq);
  for (auto* v : {&s->qvel, &s->qacc, &s->qacc_warmstart, &s->qacc_smooth, &s->qfrc_smooth, &s->qfrc_bias,
                  &s->qfrc_passive, &s->qfrc_actuator, &s->qfrc_constraint})
    std::fill(v->begin(), v->end(), 0.0);
  std::fill(s->ctrl.begin(), s->ctrl.end(), 0.0);
  std::fill(s->sensordata.begin(), s->sensordata.end(), 0.0);
  s->contacts.clear(); s->nefc = 0; s->time = 0;
}
void orc_forward(void* p) { forward(*(Sim*)p, true); }
void orc_step(void* p) {
  Sim* s = (Sim*)p;
  forward(*s, true);
  if (s->m->integrator == MJB_INT_RK4) rk4(*s); else euler(*s);
}
double* orc_array(void* p, const char* name, int* count) {
  Sim* s = (Sim*)p;
  std::vector<double>* v = nullptr;
  std::string n(name);
  if (n == "qpos") v = &s->qpos; else if (n == "qvel") v = &s->qvel; else if (n == "ctrl") v = &s->ctrl;
  else if (n == "qacc") v = &s->qacc; else if (n == "qacc_warmstart") v = &s->qacc_warmstart;
  else if (n == "qacc_smooth") v = &s->qacc_smooth; else if (n == "qfrc_bias") v = &s->qfrc_bias;
  else if (n == "qfrc_smooth") v = &s->qfrc_smooth; else if (n == "qfrc_constraint") v = &s->qfrc_constraint;
  else if (n == "qfrc_actuator") v = &s->qfrc_actuator; else if (n == "qfrc_passive") v = &s->qfrc_passive;
  else if (n == "sensordata") v = &s->sensordata; else if (n == "xpos") v = &s->xpos; else if (n == "xmat") v = &s->xmat;
  else if (n == "xipos") v = &s->xipos; else if (n == "geom_xpos") v = &s->geom_xpos;
  else if (n == "geom_xmat") v = &s->geom_xmat; else if (n == "site_xpos") v = &s->site_xpos;
  else if (n == "site_xmat") v = &s->site_xmat; else if (n == "M") v = &s->M;
  else if (n == "efc_J") v = &s->J; else if (n == "efc_force") v = &s->efc_force; else if (n == "efc_aref") v = &s->efc_aref;
  else if (n == "efc_D") v = &s->efc_D; else if (n == "efc_pos") v = &s->efc_pos; else if (n == "bvel") v = &s->bvel;
  else if (n == "bacc") v = &s->bacc;
  if (!v) { if (count) *count = -1; return nullptr; }
  if (count) *count = (int)v->size();
  return v->data();
}
double orc_time(void* p) { return ((Sim*)p)->time; }
int orc_ncon(void* p) { return (int)((Sim*)p)->contacts.size(); }
int orc_nefc(void* p) { return ((Sim*)p)->nefc; }
int orc_solver_iters(void* p) { return ((Sim*)p)->last_solver_iters; }
// contact i: geom ids + dist + pos[3] + frame[9]
int orc_contact(void* p, int i, int* geom, double* dist, double* pos, double* frame) {
  Sim* s = (Sim*)p;
  if (i < 0 || i >= (int)s->contacts.size()) return -1;
  const Contact& c = s->contacts[i];
  if (geom) { geom[0] = c.g1; geom[1] = c.g2; }
  if (dist) *dist = c.dist;
  if (pos) for (int k = 0; k < 3; k++) pos[k] = c.pos[k];
  if (frame) for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) frame[3 * r + k] = c.frame[r][k];
  return 0;
}
// Agent camera `cam` of the current kinematic state (call after orc_forward / orc_kinematics): the image formation
// documented in csrc/render_kernel.cuh, restated in fp64.  Reference: get_camera_data (mujoco_parent.py:518-556)
// renders with MuJoCo's OpenGL pipeline, which cannot run here -> image parity with the reference is UNPINNED;
// this function pins the CUDA raycaster.  out: u8 [height, width, 3], rows bottom-up.
int orc_render(void* p, int cam, int width, int height, uint8_t* out) {
  Sim* s = (Sim*)p;
  const ModelView& m = *s->m;
  if (cam < 0 || cam >= m.ncam || m.cam_mode[cam] != 0) return -1;
  kinematics(*s);
  const int b = m.cam_bodyid[cam];
  V3 o = s->kin.xpos[b] + mulv(s->kin.xmat[b], V3(m.cam_pos[3 * cam], m.cam_pos[3 * cam + 1], m.cam_pos[3 * cam + 2]));
  M3 R = mul(s->kin.xmat[b], rot_from_quat(m.cam_quat + 4 * cam));
  const double th = std::tan(0.5 * m.cam_fovy[cam] * 3.14159265358979323846 / 180.0), aspect = (double)width / height;
  const double ambient = 0.4, diffuse = 0.6;
  for (int iy = 0; iy < height; iy++)
    for (int ix = 0; ix < width; ix++) {
      V3 dl(((ix + 0.5) / width * 2 - 1) * th * aspect, ((iy + 0.5) / height * 2 - 1) * th, -1.0);
      V3 d = mulv(R, normalized(dl));
      double best = 1e300;
      int bi = -1;
      for (int g = 0; g < m.ngeom; g++) {
        if (m.geom_rgba[4 * g + 3] <= 0) continue;
        double t = ray_geom(gpos(*s, g), gmat(*s, g), &m.geom_size[3 * g], o, d, m.geom_type[g]);
        if (t >= 0 && t < best) { best = t; bi = g; }
      }
      uint8_t* px = out + ((size_t)iy * width + ix) * 3;
      if (bi < 0) { px[0] = px[1] = px[2] = 0; continue; }
      V3 hit = o + d * best, gp = gpos(*s, bi);
      M3 G = gmat(*s, bi);
      const double* sz = &m.geom_size[3 * bi];
      double nd;
      int type = m.geom_type[bi];
      if (type == MJB_GEOM_PLANE) nd = std::fabs(dot(V3(G(0, 2), G(1, 2), G(2, 2)), d));
      else if (type == MJB_GEOM_SPHERE) nd = std::fabs(dot(normalized(hit - gp), d));
      else {
        V3 lp = mulTv(G, hit - gp), ld = mulTv(G, d);
        if (type == MJB_GEOM_CAPSULE) {
          V3 n(lp.x, lp.y, lp.z - std::min(std::max(lp.z, -sz[1]), sz[1]));
          nd = std::fabs(dot(normalized(n), ld));
        } else {
          double ax = std::fabs(lp.x) / sz[0], ay = std::fabs(lp.y) / sz[1], az = std::fabs(lp.z) / sz[2];
          nd = (ax >= ay && ax >= az) ? std::fabs(ld.x) : (ay >= az ? std::fabs(ld.y) : std::fabs(ld.z));
        }
      }
      double I = ambient + diffuse * std::min(nd, 1.0);
      for (int c = 0; c < 3; c++) {
        double col = std::min(1.0, std::max(0.0, m.geom_rgba[4 * bi + c]));
        px[c] = (uint8_t)(col * I * 255.0 + 0.5);
      }
    }
  return 0;
}
// total mechanical energy (kinetic + gravitational potential), for the analytic tests
double orc_energy(void* p) {
  Sim* s = (Sim*)p;
  const ModelView& m = *s->m;
  kinematics(*s);
  mass_matrix(m, s->kin, s->M);
  double ke = 0, pe = 0;
  for (int i = 0; i < m.nv; i++)
    for (int j = 0; j < m.nv; j++) ke += 0.5 * s->qvel[i] * s->M[(size_t)i * m.nv + j] * s->qvel[j];
  for (int b = 1; b < m.nbody; b++)
    pe -= m.body_mass[b] * (m.gravity[0] * s->kin.xipos[b].x + m.gravity[1] * s->kin.xipos[b].y + m.gravity[2] * s->kin.xipos[b].z);
  return ke + pe;
}

}  // extern "C"
