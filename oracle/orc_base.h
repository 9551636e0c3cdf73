// orc_base.h — everything the fp64 oracle needs below the physics: small linear algebra, a reader for the
// packed model blob (the PUBLIC format of include/mjb_blob.h, the oracle's only contact with the product),
// kinematics, point Jacobians, the dense joint-space inertia and a dense Cholesky.
//
// TEST INFRASTRUCTURE ONLY.  Written for the oracle alone and deliberately NOT shared with the product: the
// shipped library has its own quaternion-based host kinematics (csrc/host_kin.h, used by the MJCF compiler)
// and tree-recursive device kinematics (csrc/step_kernel.cuh).  Here frames are composed as 3x3 rotation
// MATRICES (hinges by Rodrigues' formula), so a sign or ordering mistake in either quaternion path shows up
// as a parity failure instead of being shared.  Restates mj_kinematics / mj_comPos / mj_jac / mj_fullM of the
// un-vendored mujoco==2.3.3 (reference call sites: mj.mj_forward / mj.mj_step, MuJoCo_Gym/mujoco_parent.py:335,350).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/mjb_blob.h"

namespace orc {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(double s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { double n = norm(a); return n > 0 ? a * (1.0 / n) : a; }

// row-major 3x3, maps local -> world
struct M3 {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline V3 mulv(const M3& R, V3 v) { return V3(R(0, 0) * v.x + R(0, 1) * v.y + R(0, 2) * v.z, R(1, 0) * v.x + R(1, 1) * v.y + R(1, 2) * v.z, R(2, 0) * v.x + R(2, 1) * v.y + R(2, 2) * v.z); }
inline V3 mulTv(const M3& R, V3 v) { return V3(R(0, 0) * v.x + R(1, 0) * v.y + R(2, 0) * v.z, R(0, 1) * v.x + R(1, 1) * v.y + R(2, 1) * v.z, R(0, 2) * v.x + R(1, 2) * v.y + R(2, 2) * v.z); }
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C(r, c) = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
  return C;
}
// (w, x, y, z) -> rotation matrix; the quaternion is normalised first
inline M3 rot_from_quat(const double* q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = 1, x = 0, y = 0, z = 0;
  if (n > 1e-15) { w = q[0] / n; x = q[1] / n; y = q[2] / n; z = q[3] / n; }
  M3 R;
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z);     R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z);     R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y);     R(2, 1) = 2 * (y * z + w * x);     R(2, 2) = 1 - 2 * (x * x + y * y);
  return R;
}
// Rodrigues: rotation by `ang` about the unit vector `a`
inline M3 rot_axis_angle(V3 a, double ang) {
  double c = std::cos(ang), s = std::sin(ang), t = 1 - c;
  M3 R;
  R(0, 0) = c + a.x * a.x * t;       R(0, 1) = a.x * a.y * t - a.z * s; R(0, 2) = a.x * a.z * t + a.y * s;
  R(1, 0) = a.y * a.x * t + a.z * s; R(1, 1) = c + a.y * a.y * t;       R(1, 2) = a.y * a.z * t - a.x * s;
  R(2, 0) = a.z * a.x * t - a.y * s; R(2, 1) = a.z * a.y * t + a.x * s; R(2, 2) = c + a.z * a.z * t;
  return R;
}
// unit quaternion product a * b, (w, x, y, z) arrays (only the free-joint integrator needs quaternions)
inline void quat_mul(const double* a, const double* b, double* out) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  out[0] = w; out[1] = x; out[2] = y; out[3] = z;
}

// ---- model: typed pointers into the blob, by field name ----------------------------------------------------
struct Model {
  const void* blob = nullptr;
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nsensordata, npair, integrator, ncam = 0;
  double timestep;
  const double* gravity;
  const int32_t *body_parentid, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_invweight0;
  const int32_t *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp;
  const int32_t *dof_bodyid, *dof_jntid;
  const double *dof_armature, *dof_damping, *dof_invweight0;
  const int32_t *geom_type, *geom_bodyid;
  const double *geom_size, *geom_pos, *geom_quat, *geom_rbound, *geom_rgba;
  const int32_t *site_bodyid, *site_type;
  const double *site_pos, *site_quat, *site_size;
  const int32_t *sensor_type, *sensor_objid, *sensor_adr, *sensor_dim, *sensor_datatype;
  const double* sensor_cutoff;
  const int32_t *actuator_trnid, *actuator_ctrllimited;
  const double *actuator_gear, *actuator_ctrlrange;
  const int32_t *pair_geom1, *pair_geom2, *pair_condim;
  const double *pair_margin, *pair_includemargin, *pair_friction, *pair_solref, *pair_solimp;
  const double* qpos0;
  const int32_t *cam_bodyid = nullptr, *cam_mode = nullptr;
  const double *cam_pos = nullptr, *cam_quat = nullptr, *cam_fovy = nullptr;

  const int32_t* ints(const char* name, bool need = true) const {
    int n;
    const int32_t* p = mjb_blob_i32(blob, name, &n);
    if (n < 0 && need) throw std::runtime_error(std::string("oracle: blob lacks int field ") + name);
    return p;
  }
  const double* reals(const char* name, bool need = true) const {
    int n;
    const double* p = mjb_blob_f64(blob, name, &n);
    if (n < 0 && need) throw std::runtime_error(std::string("oracle: blob lacks f64 field ") + name);
    return p;
  }
  explicit Model(const void* b) : blob(b) {
    if (!b || std::memcmp(((const mjb_blob_header*)b)->magic, MJB_BLOB_MAGIC, 8) != 0) throw std::runtime_error("oracle: not a model blob");
    nq = ints("nq")[0]; nv = ints("nv")[0]; nu = ints("nu")[0]; nbody = ints("nbody")[0]; njnt = ints("njnt")[0];
    ngeom = ints("ngeom")[0]; nsite = ints("nsite")[0]; nsensor = ints("nsensor")[0]; nsensordata = ints("nsensordata")[0];
    npair = ints("npair")[0]; integrator = ints("opt_integrator")[0];
    timestep = reals("opt_timestep")[0]; gravity = reals("opt_gravity");
#define ORC_I(n) n = ints(#n)
#define ORC_F(n) n = reals(#n)
    ORC_I(body_parentid); ORC_I(body_jntnum); ORC_I(body_jntadr); ORC_I(body_dofnum); ORC_I(body_dofadr);
    ORC_F(body_pos); ORC_F(body_quat); ORC_F(body_ipos); ORC_F(body_iquat); ORC_F(body_mass); ORC_F(body_inertia); ORC_F(body_invweight0);
    ORC_I(jnt_type); ORC_I(jnt_bodyid); ORC_I(jnt_qposadr); ORC_I(jnt_dofadr); ORC_I(jnt_limited);
    ORC_F(jnt_pos); ORC_F(jnt_axis); ORC_F(jnt_range); ORC_F(jnt_margin); ORC_F(jnt_solref); ORC_F(jnt_solimp);
    ORC_I(dof_bodyid); ORC_I(dof_jntid); ORC_F(dof_armature); ORC_F(dof_damping); ORC_F(dof_invweight0);
    ORC_I(geom_type); ORC_I(geom_bodyid); ORC_F(geom_size); ORC_F(geom_pos); ORC_F(geom_quat); ORC_F(geom_rbound); ORC_F(geom_rgba);
    ORC_I(site_bodyid); ORC_I(site_type); ORC_F(site_pos); ORC_F(site_quat); ORC_F(site_size);
    ORC_I(sensor_type); ORC_I(sensor_objid); ORC_I(sensor_adr); ORC_I(sensor_dim); ORC_I(sensor_datatype); ORC_F(sensor_cutoff);
    ORC_I(actuator_trnid); ORC_I(actuator_ctrllimited); ORC_F(actuator_gear); ORC_F(actuator_ctrlrange);
    ORC_I(pair_geom1); ORC_I(pair_geom2); ORC_I(pair_condim);
    ORC_F(pair_margin); ORC_F(pair_includemargin); ORC_F(pair_friction); ORC_F(pair_solref); ORC_F(pair_solimp);
    ORC_F(qpos0);
    if (const int32_t* nc = ints("ncam", false)) {
      ncam = nc[0];
      ORC_I(cam_bodyid); ORC_I(cam_mode); ORC_F(cam_pos); ORC_F(cam_quat); ORC_F(cam_fovy);
    }
#undef ORC_I
#undef ORC_F
  }
};

inline V3 v3(const double* p) { return V3(p[0], p[1], p[2]); }

// ---- kinematics ----------------------------------------------------------------------------------------------
struct Kin {
  std::vector<V3> xpos, xipos;      // body frame origin / centre of mass, world
  std::vector<M3> xmat, ximat;      // body / inertial orientation
  // per dof: unit angular part `rot`, translational part `lin`, and for rotations a point on the axis
  std::vector<V3> dof_axis_rot, dof_axis_lin, dof_anchor;
};

inline void forward_kinematics(const Model& m, const double* qpos, Kin& k) {
  k.xpos.assign(m.nbody, V3()); k.xipos.assign(m.nbody, V3());
  k.xmat.assign(m.nbody, M3()); k.ximat.assign(m.nbody, M3());
  k.dof_axis_rot.assign(m.nv, V3()); k.dof_axis_lin.assign(m.nv, V3()); k.dof_anchor.assign(m.nv, V3());
  for (int b = 1; b < m.nbody; b++) {
    const int par = m.body_parentid[b];
    // pose of the body frame before its own joints act: parent * local offset
    V3 p = k.xpos[par] + mulv(k.xmat[par], v3(m.body_pos + 3 * b));
    M3 R = mul(k.xmat[par], rot_from_quat(m.body_quat + 4 * b));
    const int j0 = m.body_jntadr[b], jn = m.body_jntnum[b];
    for (int j = j0; j < j0 + jn; j++) {
      const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == MJB_JNT_FREE) {
        p = v3(qpos + qa);
        R = rot_from_quat(qpos + qa + 3);
        for (int i = 0; i < 3; i++) {
          V3 e; e[i] = 1;
          k.dof_axis_lin[da + i] = e;                                   // translations along the world axes
          k.dof_axis_rot[da + 3 + i] = V3(R(0, i), R(1, i), R(2, i));   // rotations about the body's own axes
          k.dof_anchor[da + 3 + i] = p;
        }
        continue;
      }
      const V3 axis_w = mulv(R, v3(m.jnt_axis + 3 * j));
      const V3 anchor_w = p + mulv(R, v3(m.jnt_pos + 3 * j));
      const double q = qpos[qa] - m.qpos0[qa];
      if (m.jnt_type[j] == MJB_JNT_HINGE) {
        // rotate the frame about the world-frame axis through the anchor: R <- Rot(axis_w, q) R, anchor stays put
        const M3 Rq = rot_axis_angle(axis_w, q);
        p = anchor_w + mulv(Rq, p - anchor_w);
        R = mul(Rq, R);
        k.dof_axis_rot[da] = axis_w; k.dof_anchor[da] = anchor_w;
      } else {  // slide
        p = p + axis_w * q;
        k.dof_axis_lin[da] = axis_w;
      }
    }
    k.xpos[b] = p; k.xmat[b] = R;
    k.xipos[b] = p + mulv(R, v3(m.body_ipos + 3 * b));
    k.ximat[b] = mul(R, rot_from_quat(m.body_iquat + 4 * b));
  }
}

// velocity of the world point `point` riding on `body`, and the body's angular velocity, per unit qvel:
// jacp / jacr are 3 x nv, row-major
inline void point_jacobian(const Model& m, const Kin& k, int body, V3 point, std::vector<double>& jacp, std::vector<double>& jacr) {
  jacp.assign(3 * m.nv, 0.0); jacr.assign(3 * m.nv, 0.0);
  while (body > 0) {
    const int d0 = m.body_dofadr[body];
    for (int d = d0; d0 >= 0 && d < d0 + m.body_dofnum[body]; d++) {
      const V3 w = k.dof_axis_rot[d];
      const V3 v = k.dof_axis_lin[d] + cross(w, point - k.dof_anchor[d]);
      for (int r = 0; r < 3; r++) { jacp[r * m.nv + d] = v[r]; jacr[r * m.nv + d] = w[r]; }
    }
    body = m.body_parentid[body];
  }
}

// M = sum over bodies of  m Jp' Jp + Jr' (R I R') Jr  at the centre of mass, + armature; dense nv x nv.  A body's
// Jacobian columns are non-zero only for the dofs between it and the world, so the products run over that list.
inline void mass_matrix(const Model& m, const Kin& k, std::vector<double>& M) {
  const int nv = m.nv;
  M.assign((size_t)nv * nv, 0.0);
  std::vector<double> jp, jr;
  std::vector<int> dofs;
  for (int b = 1; b < m.nbody; b++) {
    if (m.body_mass[b] <= 0) continue;
    dofs.clear();
    for (int a = b; a > 0; a = m.body_parentid[a])
      for (int d = m.body_dofadr[a]; d >= 0 && d < m.body_dofadr[a] + m.body_dofnum[a]; d++) dofs.push_back(d);
    if (dofs.empty()) continue;
    point_jacobian(m, k, b, k.xipos[b], jp, jr);
    const M3& R = k.ximat[b];
    double Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        Iw[3 * r + c] = R(r, 0) * m.body_inertia[3 * b] * R(c, 0) + R(r, 1) * m.body_inertia[3 * b + 1] * R(c, 1) + R(r, 2) * m.body_inertia[3 * b + 2] * R(c, 2);
    for (int i : dofs) {
      const double wi[3] = {jr[i], jr[nv + i], jr[2 * nv + i]};
      const double Iwi[3] = {Iw[0] * wi[0] + Iw[1] * wi[1] + Iw[2] * wi[2], Iw[3] * wi[0] + Iw[4] * wi[1] + Iw[5] * wi[2], Iw[6] * wi[0] + Iw[7] * wi[1] + Iw[8] * wi[2]};
      for (int j : dofs)
        M[(size_t)i * nv + j] += m.body_mass[b] * (jp[i] * jp[j] + jp[nv + i] * jp[nv + j] + jp[2 * nv + i] * jp[2 * nv + j]) +
                                 Iwi[0] * jr[j] + Iwi[1] * jr[nv + j] + Iwi[2] * jr[2 * nv + j];
    }
  }
  for (int d = 0; d < nv; d++) M[(size_t)d * nv + d] += m.dof_armature[d];
}

// dense Cholesky, column by column (A = L L', lower triangle overwritten); false when a pivot is not positive
inline bool cholesky(std::vector<double>& A, int n) {
  for (int c = 0; c < n; c++) {
    double piv = A[(size_t)c * n + c];
    for (int k = 0; k < c; k++) piv -= A[(size_t)c * n + k] * A[(size_t)c * n + k];
    if (!(piv > 0)) return false;
    piv = std::sqrt(piv);
    A[(size_t)c * n + c] = piv;
    for (int r = c + 1; r < n; r++) {
      double v = A[(size_t)r * n + c];
      for (int k = 0; k < c; k++) v -= A[(size_t)r * n + k] * A[(size_t)c * n + k];
      A[(size_t)r * n + c] = v / piv;
    }
  }
  return true;
}
inline void cholesky_solve(const std::vector<double>& L, int n, double* x) {
  for (int r = 0; r < n; r++) {
    double v = x[r];
    for (int k = 0; k < r; k++) v -= L[(size_t)r * n + k] * x[k];
    x[r] = v / L[(size_t)r * n + r];
  }
  for (int r = n - 1; r >= 0; r--) {
    double v = x[r];
    for (int k = r + 1; k < n; k++) v -= L[(size_t)k * n + r] * x[k];
    x[r] = v / L[(size_t)r * n + r];
  }
}

}  // namespace orc
