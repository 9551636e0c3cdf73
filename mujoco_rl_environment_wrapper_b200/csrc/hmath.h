// hmath.h — small fp64 vector / quaternion / 3x3 helpers for the host-side
// MJCF compiler.  Conventions follow MuJoCo's documented ones: quaternions are
// (w, x, y, z), matrices are row-major, frames map local -> world.
#pragma once
#include <cmath>

namespace mjb {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return a * s; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) {
  double n = norm(a);
  return n > 0 ? a * (1.0 / n) : a;
}

struct Quat {
  double w = 1, x = 0, y = 0, z = 0;
};
inline Quat qmul(Quat a, Quat b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
inline Quat qnormalized(Quat q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < 1e-15) return Quat();
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline Quat qaxisangle(V3 axis, double angle) {
  double s = std::sin(angle * 0.5);
  return {std::cos(angle * 0.5), axis.x * s, axis.y * s, axis.z * s};
}

struct M3 {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline M3 q2m(Quat q) {
  M3 R;
  double ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  R.m[0] = ww + xx - yy - zz;
  R.m[4] = ww - xx + yy - zz;
  R.m[8] = ww - xx - yy + zz;
  R.m[1] = 2 * (q.x * q.y - q.w * q.z);
  R.m[2] = 2 * (q.x * q.z + q.w * q.y);
  R.m[3] = 2 * (q.x * q.y + q.w * q.z);
  R.m[5] = 2 * (q.y * q.z - q.w * q.x);
  R.m[6] = 2 * (q.x * q.z - q.w * q.y);
  R.m[7] = 2 * (q.y * q.z + q.w * q.x);
  return R;
}
inline V3 mulv(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
          R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
inline V3 mulTv(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
          R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
inline M3 mulm(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C.m[3 * r + c] = A.m[3 * r] * B.m[c] + A.m[3 * r + 1] * B.m[3 + c] + A.m[3 * r + 2] * B.m[6 + c];
  return C;
}
inline M3 transpose(const M3& A) {
  M3 C;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C.m[3 * r + c] = A.m[3 * c + r];
  return C;
}
inline V3 rotq(Quat q, V3 v) { return mulv(q2m(q), v); }

// rotation matrix -> quaternion (Shepperd's method)
inline Quat m2q(const M3& R) {
  Quat q;
  double tr = R.m[0] + R.m[4] + R.m[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2;
    q.w = 0.25 * s;
    q.x = (R.m[7] - R.m[5]) / s;
    q.y = (R.m[2] - R.m[6]) / s;
    q.z = (R.m[3] - R.m[1]) / s;
  } else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) {
    double s = std::sqrt(1.0 + R.m[0] - R.m[4] - R.m[8]) * 2;
    q.w = (R.m[7] - R.m[5]) / s;
    q.x = 0.25 * s;
    q.y = (R.m[1] + R.m[3]) / s;
    q.z = (R.m[2] + R.m[6]) / s;
  } else if (R.m[4] > R.m[8]) {
    double s = std::sqrt(1.0 + R.m[4] - R.m[0] - R.m[8]) * 2;
    q.w = (R.m[2] - R.m[6]) / s;
    q.x = (R.m[1] + R.m[3]) / s;
    q.y = 0.25 * s;
    q.z = (R.m[5] + R.m[7]) / s;
  } else {
    double s = std::sqrt(1.0 + R.m[8] - R.m[0] - R.m[4]) * 2;
    q.w = (R.m[3] - R.m[1]) / s;
    q.x = (R.m[2] + R.m[6]) / s;
    q.y = (R.m[5] + R.m[7]) / s;
    q.z = 0.25 * s;
  }
  return qnormalized(q);
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi sweeps.
// A (row-major, symmetric) -> eigenvalues d[3] (descending) and a proper
// rotation V whose columns are the eigenvectors.
inline void eig3(const M3& Ain, double d[3], M3& V) {
  M3 A = Ain;
  V = M3();
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = std::fabs(A(0, 1)) + std::fabs(A(0, 2)) + std::fabs(A(1, 2));
    if (off < 1e-300) break;
    double scale = std::fabs(A(0, 0)) + std::fabs(A(1, 1)) + std::fabs(A(2, 2));
    if (off < 1e-18 * scale) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (std::fabs(A(p, q)) < 1e-300) continue;
        double theta = (A(q, q) - A(p, p)) / (2 * A(p, q));
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        M3 Jr;
        Jr(p, p) = c; Jr(q, q) = c; Jr(p, q) = s; Jr(q, p) = -s;
        A = mulm(transpose(Jr), mulm(A, Jr));
        V = mulm(V, Jr);
      }
  }
  d[0] = A(0, 0); d[1] = A(1, 1); d[2] = A(2, 2);
  // sort descending, permuting columns of V
  for (int i = 0; i < 2; i++)
    for (int j = i + 1; j < 3; j++)
      if (d[j] > d[i]) {
        double td = d[i]; d[i] = d[j]; d[j] = td;
        for (int r = 0; r < 3; r++) { double tv = V(r, i); V(r, i) = V(r, j); V(r, j) = tv; }
      }
  // make it a proper rotation
  V3 c0(V(0, 0), V(1, 0), V(2, 0)), c1(V(0, 1), V(1, 1), V(2, 1));
  V3 c2 = cross(c0, c1);
  V(0, 2) = c2.x; V(1, 2) = c2.y; V(2, 2) = c2.z;
}

}  // namespace mjb
