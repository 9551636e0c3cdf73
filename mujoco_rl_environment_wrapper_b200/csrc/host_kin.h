// host_kin.h — fp64 host kinematics + dense Jacobians + dense joint-space inertia
// (M = sum_b J_b^T I_b J_b + armature) over a ModelView.  Used by the MJCF
// compiler to evaluate the qpos0-dependent constants (body/dof invweight0, the
// quantities MuJoCo's compiler derives in its set-constants pass).  Deliberately
// the "textbook dense" formulation, NOT the tree recursion the CUDA kernels use.
#pragma once
#include <vector>

#include "hmath.h"
#include "model_view.h"

namespace mjb {

struct HostKin {
  int nbody = 0, nv = 0;
  std::vector<V3> xpos, xipos;       // body frame origin, inertial frame origin (world)
  std::vector<Quat> xquat;           // body orientation
  std::vector<M3> xmat, ximat;       // body / inertial orientation
  std::vector<V3> xanchor, xaxis;    // per joint
  std::vector<V3> dof_axis_lin;      // per dof: translation direction (zero for rotations)
  std::vector<V3> dof_axis_rot;      // per dof: rotation axis (zero for translations)
  std::vector<V3> dof_anchor;        // per dof: point the rotation axis passes through
};

inline void host_fk(const ModelView& m, const double* qpos, HostKin& k) {
  k.nbody = m.nbody; k.nv = m.nv;
  k.xpos.assign(m.nbody, V3()); k.xipos.assign(m.nbody, V3());
  k.xquat.assign(m.nbody, Quat()); k.xmat.assign(m.nbody, M3()); k.ximat.assign(m.nbody, M3());
  k.xanchor.assign(m.njnt, V3()); k.xaxis.assign(m.njnt, V3());
  k.dof_axis_lin.assign(m.nv, V3()); k.dof_axis_rot.assign(m.nv, V3()); k.dof_anchor.assign(m.nv, V3());
  for (int b = 1; b < m.nbody; b++) {
    int p = m.body_parentid[b];
    V3 bpos(m.body_pos[3 * b], m.body_pos[3 * b + 1], m.body_pos[3 * b + 2]);
    Quat bq{m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]};
    V3 pos; Quat quat;
    int jn = m.body_jntnum[b], ja = m.body_jntadr[b];
    if (jn == 1 && m.jnt_type[ja] == MJB_JNT_FREE) {
      int qa = m.jnt_qposadr[ja];
      pos = V3(qpos[qa], qpos[qa + 1], qpos[qa + 2]);
      quat = qnormalized(Quat{qpos[qa + 3], qpos[qa + 4], qpos[qa + 5], qpos[qa + 6]});
      k.xanchor[ja] = pos; k.xaxis[ja] = V3(0, 0, 1);
    } else {
      pos = k.xpos[p] + mulv(k.xmat[p], bpos);
      quat = qmul(k.xquat[p], bq);
      for (int j = ja; j < ja + jn; j++) {
        V3 jpos(m.jnt_pos[3 * j], m.jnt_pos[3 * j + 1], m.jnt_pos[3 * j + 2]);
        V3 jax(m.jnt_axis[3 * j], m.jnt_axis[3 * j + 1], m.jnt_axis[3 * j + 2]);
        V3 anchor = pos + rotq(quat, jpos);
        V3 axis = rotq(quat, jax);
        k.xanchor[j] = anchor; k.xaxis[j] = axis;
        int qa = m.jnt_qposadr[j];
        double q = qpos[qa] - m.qpos0[qa];
        if (m.jnt_type[j] == MJB_JNT_HINGE) {
          quat = qmul(quat, qaxisangle(jax, q));
          pos = anchor - rotq(quat, jpos);
        } else if (m.jnt_type[j] == MJB_JNT_SLIDE) {
          pos = pos + axis * q;
        }
      }
    }
    quat = qnormalized(quat);
    k.xpos[b] = pos; k.xquat[b] = quat; k.xmat[b] = q2m(quat);
    V3 ipos(m.body_ipos[3 * b], m.body_ipos[3 * b + 1], m.body_ipos[3 * b + 2]);
    Quat iq{m.body_iquat[4 * b], m.body_iquat[4 * b + 1], m.body_iquat[4 * b + 2], m.body_iquat[4 * b + 3]};
    k.xipos[b] = pos + mulv(k.xmat[b], ipos);
    k.ximat[b] = q2m(qmul(quat, iq));
    // dof axes (world frame)
    for (int j = ja; j < ja + jn; j++) {
      int da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == MJB_JNT_FREE) {
        for (int i = 0; i < 3; i++) {
          V3 e; e[i] = 1;
          k.dof_axis_lin[da + i] = e;
          k.dof_axis_rot[da + 3 + i] = V3(k.xmat[b](0, i), k.xmat[b](1, i), k.xmat[b](2, i));
          k.dof_anchor[da + 3 + i] = pos;
        }
      } else if (m.jnt_type[j] == MJB_JNT_HINGE) {
        k.dof_axis_rot[da] = k.xaxis[j]; k.dof_anchor[da] = k.xanchor[j];
      } else if (m.jnt_type[j] == MJB_JNT_SLIDE) {
        k.dof_axis_lin[da] = k.xaxis[j];
      }
    }
  }
}

// Jacobian of a world point attached to `body`: jacp, jacr are 3 x nv row-major.
inline void host_jac(const ModelView& m, const HostKin& k, int body, V3 point, std::vector<double>& jacp,
                     std::vector<double>& jacr) {
  jacp.assign(3 * m.nv, 0.0); jacr.assign(3 * m.nv, 0.0);
  // walk up the body chain; every dof of every ancestor (incl. self) contributes
  for (int b = body; b > 0; b = m.body_parentid[b]) {
    for (int d = m.body_dofadr[b]; d < m.body_dofadr[b] + m.body_dofnum[b]; d++) {
      V3 lin = k.dof_axis_lin[d], rot = k.dof_axis_rot[d];
      V3 v = lin + cross(rot, point - k.dof_anchor[d]);
      for (int r = 0; r < 3; r++) { jacp[r * m.nv + d] = v[r]; jacr[r * m.nv + d] = rot[r]; }
    }
  }
}

// dense joint-space inertia (nv x nv, row-major), armature included.  Only the dofs on a body's chain
// contribute to its Jacobian, so the double loop runs over that chain.
inline void host_mass_matrix(const ModelView& m, const HostKin& k, std::vector<double>& M) {
  int nv = m.nv;
  M.assign((size_t)nv * nv, 0.0);
  std::vector<double> jp, jr;
  std::vector<int> chain;
  for (int b = 1; b < m.nbody; b++) {
    double mass = m.body_mass[b];
    if (mass <= 0) continue;
    chain.clear();
    for (int bb = b; bb > 0; bb = m.body_parentid[bb])
      for (int d = m.body_dofadr[bb]; d >= 0 && d < m.body_dofadr[bb] + m.body_dofnum[bb]; d++) chain.push_back(d);
    if (chain.empty()) continue;
    host_jac(m, k, b, k.xipos[b], jp, jr);
    // world inertia = R diag(I) R^T
    M3 R = k.ximat[b], Iw;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int a = 0; a < 3; a++) s += R(r, a) * m.body_inertia[3 * b + a] * R(c, a);
        Iw(r, c) = s;
      }
    for (int i : chain) {
      double Ij[3];  // Iw * jr[:, i]
      for (int r = 0; r < 3; r++) Ij[r] = Iw(r, 0) * jr[i] + Iw(r, 1) * jr[nv + i] + Iw(r, 2) * jr[2 * nv + i];
      for (int j : chain) {
        double s = mass * (jp[i] * jp[j] + jp[nv + i] * jp[nv + j] + jp[2 * nv + i] * jp[2 * nv + j]) +
                   Ij[0] * jr[j] + Ij[1] * jr[nv + j] + Ij[2] * jr[2 * nv + j];
        M[(size_t)i * nv + j] += s;
      }
    }
  }
  for (int i = 0; i < nv; i++) M[(size_t)i * nv + i] += m.dof_armature[i];
}

// in-place dense Cholesky (lower), returns false if not positive definite
inline bool host_cholesky(std::vector<double>& A, int n) {
  for (int j = 0; j < n; j++) {
    double s = A[(size_t)j * n + j];
    for (int k2 = 0; k2 < j; k2++) s -= A[(size_t)j * n + k2] * A[(size_t)j * n + k2];
    if (s <= 0) return false;
    double d = std::sqrt(s);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double t = A[(size_t)i * n + j];
      for (int k2 = 0; k2 < j; k2++) t -= A[(size_t)i * n + k2] * A[(size_t)j * n + k2];
      A[(size_t)i * n + j] = t / d;
    }
  }
  return true;
}
inline void host_chol_solve(const std::vector<double>& L, int n, double* x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k2 = 0; k2 < i; k2++) s -= L[(size_t)i * n + k2] * x[k2];
    x[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k2 = i + 1; k2 < n; k2++) s -= L[(size_t)k2 * n + i] * x[k2];
    x[i] = s / L[(size_t)i * n + i];
  }
}

}  // namespace mjb
