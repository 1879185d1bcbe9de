// fp64 SE3 / quaternion device helpers and the Huber kernel shared by the g2o-style solvers
// (lba.cu, pose_opt.cu).  Follow Thirdparty/g2o/g2o/types/se3quat.h:98-120, 217-285 (map, *, exp,
// normalizeRotation), Eigen's quaternion <-> matrix conversions, and
// Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91 of the reference.
#pragma once
#include <cuda_runtime.h>

namespace orbb200 {

struct DQuat { double x, y, z, w; };

__device__ __forceinline__ void q_normalize(DQuat& q) {  // se3quat.h:280-285
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
__device__ __forceinline__ DQuat q_mul(const DQuat& a, const DQuat& b) {
  DQuat r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}
__device__ __forceinline__ void q_rot(const DQuat& q, const double* v, double* o) {
  double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  o[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  o[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
__device__ __forceinline__ void q_to_R(const DQuat& q, double* R) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ DQuat R_to_q(const double* R) {
  DQuat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
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

struct HuberD { double delta, dsqr; };
__device__ __forceinline__ void robustify(const HuberD& h, double e, double& rho0, double& rho1) {
  if (e <= h.dsqr) { rho0 = e; rho1 = 1.; }
  else { const double s = sqrt(e); rho0 = 2 * s * h.delta - h.dsqr; rho1 = h.delta / s; }
}


// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:73-76): P_out = SE3Quat::exp(u) * P_in,
// u = (omega, upsilon), P = quaternion xyzw + translation (se3quat.h:223-257, :98-104).
__device__ __forceinline__ void se3_exp_mul(const double* u, const double* P, double* out) {
  const double w0 = u[0], w1 = u[1], w2 = u[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[i * 3 + j] = Om[i * 3] * Om[j] + Om[i * 3 + 1] * Om[3 + j] + Om[i * 3 + 2] * Om[6 + j];
  double R[9], V[9];
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
    const double c = (theta - sin(theta)) / pow(theta, 3.0);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  DQuat qe = R_to_q(R);
  q_normalize(qe);
  double te[3];
  for (int i = 0; i < 3; i++) te[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
  const DQuat q0 = {P[0], P[1], P[2], P[3]};
  const double t0[3] = {P[4], P[5], P[6]};
  double rt[3];
  q_rot(qe, t0, rt);
  DQuat qn = q_mul(qe, q0);
  q_normalize(qn);
  out[0] = qn.x; out[1] = qn.y; out[2] = qn.z; out[3] = qn.w;
  out[4] = te[0] + rt[0]; out[5] = te[1] + rt[1]; out[6] = te[2] + rt[2];
}

}  // namespace orbb200
