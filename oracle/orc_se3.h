// TEST INFRASTRUCTURE ONLY (see orc_common.h).  fp64 SE3 / quaternion helpers and the Huber kernel
// shared by the g2o restatements (orc_lba.cpp, orc_pose.cpp).  Restates (paths relative to
// /root/reference): Thirdparty/g2o/g2o/types/se3quat.h:98-120, 217-285 (map, *, exp,
// normalizeRotation), Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91 (RobustKernelHuber,
// float dsqr), and the Eigen quaternion <-> matrix conversions those use.
#pragma once
#include <cmath>

namespace {


struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

inline void quat_normalize(Quat& q) {  // se3quat.h:280-285
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
inline Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline void quat_rot(const Quat& q, const double v[3], double out[3]) {  // Eigen _transformVector
  double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
inline void quat_to_R(const Quat& q, double R[9]) {  // Eigen toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
inline Quat R_to_quat(const double R[9]) {  // Eigen quaternion from rotation matrix
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
inline void se3_map(const SE3& T, const double p[3], double out[3]) {
  quat_rot(T.r, p, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}
// SE3Quat::exp(update) * T   (se3quat.h:223-257, :98-104; oplusImpl)
inline SE3 se3_exp_mul(const double u[6], const SE3& T) {
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double Om[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Om[i * 3 + k] * Om[k * 3 + j];
      Om2[i * 3 + j] = s;
    }
  double R[9], V[9];
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0 ? 1.0 : 0.0);
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  SE3 E;
  E.r = R_to_quat(R);
  for (int i = 0; i < 3; i++) E.t[i] = V[i * 3] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
  quat_normalize(E.r);  // SE3Quat(q,t) ctor
  SE3 out;
  double rt[3];
  quat_rot(E.r, T.t, rt);
  for (int i = 0; i < 3; i++) out.t[i] = E.t[i] + rt[i];
  out.r = quat_mul(E.r, T.r);
  quat_normalize(out.r);
  return out;
}

struct Huber {
  double delta; float dsqr;
  explicit Huber(float th) : delta(th), dsqr((float)((double)th * (double)th)) {}
  // robust_kernel_impl.cpp:78-91: rho[0], rho[1]
  inline void robustify(double e, double& rho0, double& rho1) const {
    if (e <= dsqr) { rho0 = e; rho1 = 1.; }
    else { const double s = std::sqrt(e); rho0 = 2 * s * delta - dsqr; rho1 = delta / s; }
  }
};

}  // namespace
