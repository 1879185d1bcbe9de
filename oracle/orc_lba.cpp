// TEST INFRASTRUCTURE ONLY (see orc_common.h).  fp64 CPU restatement of the g2o
// Levenberg-Marquardt loop that Optimizer::LocalBundleAdjustment runs
// (src/Optimizer.cc:1410-1411 `optimizer.optimize(10)`), on the flat graph of
// include/orb_b200.h (interface types only).
//
// Restates (paths relative to /root/reference):
//   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419   optimize()
//   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-194  solve(), lambda init, scale
//   Thirdparty/g2o/g2o/core/block_solver.hpp:354-486 (Schur solve), :502-560 (buildSystem), :564-604
//   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-120     constructQuadraticForm
//   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91    RobustKernelHuber (float dsqr)
//   Thirdparty/g2o/g2o/types/se3quat.h:98-120, 217-285       SE3Quat map / * / exp / normalizeRotation
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:190-197, 228-274  stereo edge
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:73-76    VertexSE3Expmap::oplusImpl
//   src/OptimizableTypes.cpp:139-160, include/OptimizableTypes.h:99-104   mono edge
//   src/CameraModels/Pinhole.cpp:42-48, 71-81                project / projectJac (float parameters)
//   src/CameraModels/KannalaBrandt8.cpp:46-65, 145-175       project (atan2f / sqrtf in float) / projectJac
//   src/OptimizableTypes.cpp:192-213, include/OptimizableTypes.h:150-185   EdgeSE3ProjectXYZToBody (second camera)
//   src/Optimizer.cc:1366-1400                               body edges: obs, Huber (mono delta), mTrl, pCamera = mpCamera2
//   src/Optimizer.cc:1275-1276, 1305-1364                    Huber deltas, information, edge set-up
// The reduced system is solved with a skyline LDL^T on the dense S instead of
// Eigen's SimplicialLDLT (un-vendored); same solution up to rounding.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/orb_b200.h"

#include "orc_se3.h"
#include "orc_lm_ops.h"

namespace {

// GeometricCamera::project / projectJac of the two camera models on float parameters
// p = fx, fy, cx, cy [, k0..k3] (std::vector<float> mvParameters promoted to double as the expressions do).
inline void cam_project(int model, const float* p, const float* k, const double X[3], double uv[2]) {
  if (model == ORB_CAM_KB8) {  // KannalaBrandt8.cpp:46-65
    const double x2_plus_y2 = X[0] * X[0] + X[1] * X[1];
    const double theta = atan2f(sqrtf((float)x2_plus_y2), (float)X[2]);
    const double psi = atan2f((float)X[1], (float)X[0]);
    const double theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2,
                 theta9 = theta7 * theta2;
    const double r = theta + k[0] * theta3 + k[1] * theta5 + k[2] * theta7 + k[3] * theta9;
    uv[0] = p[0] * r * cos(psi) + p[2];
    uv[1] = p[1] * r * sin(psi) + p[3];
  } else {  // Pinhole.cpp:42-48
    uv[0] = p[0] * X[0] / X[2] + p[2];
    uv[1] = p[1] * X[1] / X[2] + p[3];
  }
}
inline void cam_project_jac(int model, const float* p, const float* k, const double X[3], double J[6]) {
  if (model == ORB_CAM_KB8) {  // KannalaBrandt8.cpp:145-175
    const double x2 = X[0] * X[0], y2 = X[1] * X[1], z2 = X[2] * X[2];
    const double r2 = x2 + y2, r = sqrt(r2), r3 = r2 * r;
    const double theta = atan2(r, X[2]);
    const double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta4 * theta,
                 theta6 = theta2 * theta4, theta7 = theta6 * theta, theta8 = theta4 * theta4, theta9 = theta8 * theta;
    const double f = theta + theta3 * k[0] + theta5 * k[1] + theta7 * k[2] + theta9 * k[3];
    const double fd = 1 + 3 * k[0] * theta2 + 5 * k[1] * theta4 + 7 * k[2] * theta6 + 9 * k[3] * theta8;
    J[0] = p[0] * (fd * X[2] * x2 / (r2 * (r2 + z2)) + f * y2 / r3);
    J[3] = p[1] * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
    J[1] = p[0] * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
    J[4] = p[1] * (fd * X[2] * y2 / (r2 * (r2 + z2)) + f * x2 / r3);
    J[2] = -p[0] * fd * X[0] / (r2 + z2);
    J[5] = -p[1] * fd * X[1] / (r2 + z2);
  } else {  // Pinhole.cpp:71-81
    J[0] = p[0] / X[2]; J[1] = 0.; J[2] = -p[0] * X[0] / (X[2] * X[2]);
    J[3] = 0.; J[4] = p[1] / X[2]; J[5] = -p[1] * X[1] / (X[2] * X[2]);
  }
}

struct Problem {
  const lba_graph_view* g;
  std::vector<SE3> pose;
  std::vector<double> pt;
  std::vector<int> free_idx;  // kf -> free pose index or -1
  int nf = 0;
  std::vector<double> err;    // per edge, 3 comps
  Huber hm, hs;
  // system
  std::vector<double> Hpp, Hll, W, b;  // Hpp nf*36, Hll nmp*9, W nedges*18 (6x3), b (6nf+3nmp)
  std::vector<double> x;
  std::vector<std::vector<int>> lm_edges;

  explicit Problem(const lba_graph_view* gv)
      : g(gv), hm((float)std::sqrt(5.991)), hs((float)std::sqrt(7.815)) {  // Optimizer.cc:1275-1276
    pose.resize(g->n_kf);
    free_idx.assign(g->n_kf, -1);
    for (int k = 0; k < g->n_kf; k++) {
      const double* p = g->kf_pose + 7 * k;
      pose[k].r = Quat{p[0], p[1], p[2], p[3]};
      quat_normalize(pose[k].r);
      pose[k].t[0] = p[4]; pose[k].t[1] = p[5]; pose[k].t[2] = p[6];
      if (!g->kf_fixed[k]) free_idx[k] = nf++;
    }
    pt.assign(g->mp_pos, g->mp_pos + 3 * (size_t)g->n_mp);
    err.assign(3 * (size_t)g->n_edges, 0.0);
    lm_edges.resize(g->n_mp);
    for (int e = 0; e < g->n_edges; e++) lm_edges[g->e_mp[e]].push_back(e);
    Hpp.resize((size_t)nf * 36); Hll.resize((size_t)g->n_mp * 9); W.resize((size_t)g->n_edges * 18);
    b.resize((size_t)6 * nf + 3 * (size_t)g->n_mp);
    x.assign(b.size(), 0.0);
  }

  inline int dim(int e) const { return g->e_stereo[e] == LBA_EDGE_STEREO ? 3 : 2; }
  inline bool is_stereo(int e) const { return g->e_stereo[e] == LBA_EDGE_STEREO; }
  inline bool is_body(int e) const { return g->e_stereo[e] == LBA_EDGE_BODY; }
  // e->pCamera of a 2-D edge: mpCamera (mono, Optimizer.cc:1326) or mpCamera2 (body, :1387)
  struct Cam { int model; const float* p; const float* k; };
  inline Cam edge_cam(int e) const {
    static const float zero4[4] = {0, 0, 0, 0};
    const int k = g->e_kf[e];
    if (is_body(e))
      return Cam{g->kf_cam2_model ? g->kf_cam2_model[k] : ORB_CAM_PINHOLE, g->kf_cam2 + 8 * (size_t)k, g->kf_cam2 + 8 * (size_t)k + 4};
    return Cam{g->kf_cam_model ? g->kf_cam_model[k] : ORB_CAM_PINHOLE, g->kf_cam + 5 * (size_t)k,
               g->kf_cam_dist ? g->kf_cam_dist + 4 * (size_t)k : zero4};
  }
  inline SE3 trw(int k) const {  // mTrl * Tcw: SE3Quat::operator* (se3quat.h:104-110)
    const SE3 Trl = trl(k);
    SE3 T;
    quat_rot(Trl.r, pose[k].t, T.t);
    for (int i = 0; i < 3; i++) T.t[i] += Trl.t[i];
    T.r = quat_mul(Trl.r, pose[k].r);
    quat_normalize(T.r);
    return T;
  }
  inline SE3 trl(int k) const {  // e->mTrl (:1384-1385)
    const double* p = g->kf_trl + 7 * (size_t)k;
    SE3 T; T.r = Quat{p[0], p[1], p[2], p[3]}; T.t[0] = p[4]; T.t[1] = p[5]; T.t[2] = p[6];
    return T;
  }

  // computeError of both edge types
  void compute_errors() {
    for (int e = 0; e < g->n_edges; e++) {
      const int k = g->e_kf[e];
      double Xc[3];
      se3_map(pose[k], &pt[3 * (size_t)g->e_mp[e]], Xc);
      const float* cam = g->kf_cam + 5 * k;
      const double* obs = g->e_obs + 3 * (size_t)e;
      double* r = &err[3 * (size_t)e];
      if (!is_stereo(e)) {
        // EdgeSE3ProjectXYZ::computeError (OptimizableTypes.h:99-104): obs - pCamera->project(Tcw.map(Xw));
        // EdgeSE3ProjectXYZToBody::computeError (:156-161): obs - pCamera->project((mTrl * Tcw).map(Xw))
        double Xe[3] = {Xc[0], Xc[1], Xc[2]};
        if (is_body(e)) se3_map(trw(k), &pt[3 * (size_t)g->e_mp[e]], Xe);
        const Cam c = edge_cam(e);
        double uv[2];
        cam_project(c.model, c.p, c.k, Xe, uv);
        r[0] = obs[0] - uv[0]; r[1] = obs[1] - uv[1]; r[2] = 0;
        continue;
      }
      if (g->e_stereo[e]) {
        // types_six_dof_expmap.cpp:190-197: `const float invz = 1.0f/trans_xyz[2]` -- the quotient is formed in double
        // (trans_xyz[2] is a double) and THEN rounded to float; bf arrives as `const float&`, so bf*invz is a float product;
        // fx.. are doubles set from floats
        const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
        const float bf = (float)(double)cam[4];
        const float invz = (float)(1.0 / Xc[2]);
        const double u = Xc[0] * invz * fx + cx;
        const double v = Xc[1] * invz * fy + cy;
        r[0] = obs[0] - u; r[1] = obs[1] - v; r[2] = obs[2] - (u - bf * invz);
      } else {
        // Pinhole::project(Vector3d): float parameters promoted
        r[0] = obs[0] - (cam[0] * Xc[0] / Xc[2] + cam[2]);
        r[1] = obs[1] - (cam[1] * Xc[1] / Xc[2] + cam[3]);
        r[2] = 0;
      }
    }
  }
  inline double chi2(int e) const {
    const double s = g->e_inv_sigma2[e];
    const double* r = &err[3 * (size_t)e];
    double c = r[0] * (s * r[0]) + r[1] * (s * r[1]);
    if (is_stereo(e)) c += r[2] * (s * r[2]);
    return c;
  }
  double robust_chi2() const {  // sparse_optimizer.cpp:100-114
    double chi = 0, r0, r1;
    for (int e = 0; e < g->n_edges; e++) {
      (is_stereo(e) ? hs : hm).robustify(chi2(e), r0, r1);  // body edges: thHuberMono (:1380-1382)
      chi += r0;
    }
    return chi;
  }

  // linearizeOplus of one edge: A = d err / d point (d x 3), B = d err / d pose (d x 6), row-major
  void edge_jac(int e, double* A, double* B) const {
      const int k = g->e_kf[e], l = g->e_mp[e];
      double Xc[3], R[9];
      se3_map(pose[k], &pt[3 * (size_t)l], Xc);
      quat_to_R(pose[k].r, R);
      const float* cam = g->kf_cam + 5 * k;
      const double x = Xc[0], y = Xc[1], z = Xc[2];
      if (is_stereo(e)) {
        const double fx = cam[0], fy = cam[1], bf = cam[4];
        const double z_2 = z * z;
        for (int c = 0; c < 3; c++) {
          A[0 * 3 + c] = -fx * R[0 * 3 + c] / z + fx * x * R[2 * 3 + c] / z_2;
          A[1 * 3 + c] = -fy * R[1 * 3 + c] / z + fy * y * R[2 * 3 + c] / z_2;
          A[2 * 3 + c] = A[0 * 3 + c] - bf * R[2 * 3 + c] / z_2;
        }
        B[0] = x * y / z_2 * fx; B[1] = -(1 + (x * x / z_2)) * fx; B[2] = y / z * fx;
        B[3] = -1. / z * fx; B[4] = 0; B[5] = x / z_2 * fx;
        B[6] = (1 + y * y / z_2) * fy; B[7] = -x * y / z_2 * fy; B[8] = -x / z * fy;
        B[9] = 0; B[10] = -1. / z * fy; B[11] = y / z_2 * fy;
        B[12] = B[0] - bf * y / z_2; B[13] = B[1] + bf * x / z_2; B[14] = B[2];
        B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf / z_2;
      } else if (is_body(e)) {
        // EdgeSE3ProjectXYZToBody::linearizeOplus (OptimizableTypes.cpp:192-213):
        //   Xi = -projectJac(X_r) * (Trl * Tlw).rotation();  Xj = -projectJac(X_r) * Rrl * SE3deriv(X_l)
        const SE3 Trl = trl(k);
        double Xr[3], Rrl[9], Rrw[9], J[6];
        se3_map(Trl, Xc, Xr);
        quat_to_R(Trl.r, Rrl);
        quat_to_R(trw(k).r, Rrw);
        const Cam c = edge_cam(e);
        cam_project_jac(c.model, c.p, c.k, Xr, J);
        for (int i = 0; i < 6; i++) J[i] = -J[i];
        for (int r = 0; r < 2; r++)
          for (int cc = 0; cc < 3; cc++)
            A[r * 3 + cc] = J[r * 3] * Rrw[cc] + J[r * 3 + 1] * Rrw[3 + cc] + J[r * 3 + 2] * Rrw[6 + cc];
        double JR[6];
        for (int r = 0; r < 2; r++)
          for (int cc = 0; cc < 3; cc++)
            JR[r * 3 + cc] = J[r * 3] * Rrl[cc] + J[r * 3 + 1] * Rrl[3 + cc] + J[r * 3 + 2] * Rrl[6 + cc];
        const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
        for (int r = 0; r < 2; r++)
          for (int cc = 0; cc < 6; cc++)
            B[r * 6 + cc] = JR[r * 3] * D[cc] + JR[r * 3 + 1] * D[6 + cc] + JR[r * 3 + 2] * D[12 + cc];
      } else {
        // -projectJac (Pinhole.cpp:71-81 / KannalaBrandt8.cpp:145-175), times R / SE3deriv (OptimizableTypes.cpp:139-160)
        const Cam c = edge_cam(e);
        double J[6];
        cam_project_jac(c.model, c.p, c.k, Xc, J);
        for (int i = 0; i < 6; i++) J[i] = -J[i];
        for (int r = 0; r < 2; r++)
          for (int cc = 0; cc < 3; cc++)
            A[r * 3 + cc] = J[r * 3] * R[cc] + J[r * 3 + 1] * R[3 + cc] + J[r * 3 + 2] * R[6 + cc];
        const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
        for (int r = 0; r < 2; r++)
          for (int cc = 0; cc < 6; cc++)
            B[r * 6 + cc] = J[r * 3] * D[cc] + J[r * 3 + 1] * D[6 + cc] + J[r * 3 + 2] * D[12 + cc];
      }
  }

  // linearizeOplus + constructQuadraticForm for every edge, copyB
  void build_system() {
    std::fill(Hpp.begin(), Hpp.end(), 0.0);
    std::fill(Hll.begin(), Hll.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    for (int e = 0; e < g->n_edges; e++) {
      const int k = g->e_kf[e], l = g->e_mp[e], d = dim(e);
      double A[9], B[18];  // d x 3, d x 6 row-major
      edge_jac(e, A, B);
      double rho0, rho1;
      (is_stereo(e) ? hs : hm).robustify(chi2(e), rho0, rho1);
      const double s = g->e_inv_sigma2[e];
      const double ws = rho1 * s;  // weightedOmega = rho[1] * information
      const double* r = &err[3 * (size_t)e];
      double omega_r[3];
      for (int i = 0; i < d; i++) omega_r[i] = -(s * r[i]) * rho1;
      // landmark (from) part
      double* Hl = &Hll[9 * (size_t)l];
      double* bl = &b[(size_t)6 * nf + 3 * (size_t)l];
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
          double acc = 0;
          for (int q = 0; q < d; q++) acc += A[q * 3 + i] * ws * A[q * 3 + j];
          Hl[i * 3 + j] += acc;
        }
        double acc = 0;
        for (int q = 0; q < d; q++) acc += A[q * 3 + i] * omega_r[q];
        bl[i] += acc;
      }
      const int f = free_idx[k];
      double* We = &W[18 * (size_t)e];
      if (f >= 0) {
        double* Hp = &Hpp[36 * (size_t)f];
        double* bp = &b[6 * (size_t)f];
        for (int i = 0; i < 6; i++) {
          for (int j = 0; j < 6; j++) {
            double acc = 0;
            for (int q = 0; q < d; q++) acc += B[q * 6 + i] * ws * B[q * 6 + j];
            Hp[i * 6 + j] += acc;
          }
          double acc = 0;
          for (int q = 0; q < d; q++) acc += B[q * 6 + i] * omega_r[q];
          bp[i] += acc;
          for (int j = 0; j < 3; j++) {  // Hpl block (pose rows, landmark cols) = B^T wOmega A
            double a2 = 0;
            for (int q = 0; q < d; q++) a2 += B[q * 6 + i] * ws * A[q * 3 + j];
            We[i * 3 + j] = a2;
          }
        }
      } else {
        for (int i = 0; i < 18; i++) We[i] = 0;
      }
    }
  }
};

// Skyline LDL^T of a dense symmetric matrix (lower triangle, row-major n x n),
// in place; returns false on a zero pivot (Eigen SimplicialLDLT reports failure
// only then).  first[i] = first structurally non-zero column of row i.
bool skyline_ldlt(std::vector<double>& A, int n, std::vector<int>& first, std::vector<double>& D) {
  first.resize(n);
  D.resize(n);
  for (int i = 0; i < n; i++) {
    int f = 0;
    while (f < i && A[(size_t)i * n + f] == 0.0) f++;
    first[i] = f;
  }
  for (int i = 0; i < n; i++) {
    double* Li = &A[(size_t)i * n];
    for (int j = first[i]; j < i; j++) {
      const double* Lj = &A[(size_t)j * n];
      double s = Li[j];
      for (int k = std::max(first[i], first[j]); k < j; k++) s -= Li[k] * Lj[k];  // Li[k] still holds L*D
      Li[j] = s;  // = L_ij * D_j
    }
    double d = Li[i];
    for (int j = first[i]; j < i; j++) {
      const double ld = Li[j];
      const double l = ld / D[j];
      d -= ld * l;
      Li[j] = l;
    }
    if (d == 0.0) return false;
    D[i] = d;
  }
  return true;
}

void skyline_solve(const std::vector<double>& L, const std::vector<int>& first, const std::vector<double>& D,
                   int n, double* x /* in: rhs, out: solution */) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = first[i]; k < i; k++) s -= L[(size_t)i * n + k] * x[k];
    x[i] = s;
  }
  for (int i = 0; i < n; i++) x[i] /= D[i];
  for (int i = n - 1; i >= 0; i--) {
    const double xi = x[i];
    for (int k = first[i]; k < i; k++) x[k] -= L[(size_t)i * n + k] * xi;
  }
}

inline bool inv3(const double* m, double* o) {  // Eigen 3x3 inverse (cofactors)
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

// Schur complement for damping lambda: S (n x n dense, n = 6 nf), bs, Dinv
void schur(const Problem& P, double lambda, std::vector<double>& S, std::vector<double>& bs,
           std::vector<double>& Dinv, const uint8_t* lm_mask) {
  const int nf = P.nf, n = 6 * nf;
  S.assign((size_t)n * n, 0.0);
  bs.assign(P.b.begin(), P.b.begin() + n);
  for (int f = 0; f < nf; f++)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++)
        S[(size_t)(6 * f + i) * n + 6 * f + j] = P.Hpp[36 * (size_t)f + i * 6 + j] + (i == j ? lambda : 0.0);
  Dinv.assign((size_t)P.g->n_mp * 9, 0.0);
  for (int l = 0; l < P.g->n_mp; l++) {
    if (lm_mask && !lm_mask[l]) continue;
    double Dm[9];
    for (int i = 0; i < 9; i++) Dm[i] = P.Hll[9 * (size_t)l + i] + ((i % 4 == 0) ? lambda : 0.0);
    double* Di = &Dinv[9 * (size_t)l];
    inv3(Dm, Di);
    const double* bl = &P.b[(size_t)n + 3 * (size_t)l];
    double db[3];
    for (int i = 0; i < 3; i++) db[i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
    const std::vector<int>& es = P.lm_edges[l];
    for (int e1 : es) {
      const int f1 = P.free_idx[P.g->e_kf[e1]];
      if (f1 < 0) continue;
      const double* W1 = &P.W[18 * (size_t)e1];
      double Y[18];
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 3; j++)
          Y[i * 3 + j] = W1[i * 3] * Di[j] + W1[i * 3 + 1] * Di[3 + j] + W1[i * 3 + 2] * Di[6 + j];
      for (int i = 0; i < 6; i++) bs[6 * f1 + i] -= W1[i * 3] * db[0] + W1[i * 3 + 1] * db[1] + W1[i * 3 + 2] * db[2];
      for (int e2 : es) {
        const int f2 = P.free_idx[P.g->e_kf[e2]];
        if (f2 < 0) continue;
        const double* W2 = &P.W[18 * (size_t)e2];
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++)
            S[(size_t)(6 * f1 + i) * n + 6 * f2 + j] -= Y[i * 3] * W2[j * 3] + Y[i * 3 + 1] * W2[j * 3 + 1] + Y[i * 3 + 2] * W2[j * 3 + 2];
      }
    }
  }
}


// The operations g2o's LM driver calls on its Solver / SparseOptimizer (optimization_algorithm_levenberg.cpp:61-169), over a
// Problem: buildSystem, setLambda + solve (Schur + LDL^T + landmark back-substitution, block_solver.hpp:354-486), update,
// push / pop / discardTop, the diagonal computeLambdaInit scans and the sum computeScale forms.  orc_lba_solve below runs
// its restated control law on them; the orc_lba_stepper_* entry points expose the SAME operations so that the reference's
// own control law -- optimization_algorithm_levenberg.cpp compiled unmodified into oracle/_ref/libref_lm.so -- can drive
// them too (tests/test_ref_lm.py holds the two drivers equal).
struct Stepper {
  Problem P;
  const int nf, n;
  const size_t nvec;
  std::vector<double> S, bs, Dinv, D;
  std::vector<int> first;
  std::vector<std::vector<SE3>> pose_stack;
  std::vector<std::vector<double>> pt_stack;
  const volatile uint8_t* stop = nullptr;   // pbStopFlag -> optimizer.setForceStopFlag (Optimizer.cc:1153-1154)
  explicit Stepper(const lba_graph_view* g) : P(g), nf(P.nf), n(6 * P.nf), nvec(P.b.size()) {}
  double max_diagonal() const {  // computeLambdaInit: max |H_jj| over all free vertices
    double mx = 0;
    for (int f = 0; f < nf; f++)
      for (int j = 0; j < 6; j++) mx = std::max(std::fabs(P.Hpp[36 * (size_t)f + j * 7]), mx);
    for (int l = 0; l < P.g->n_mp; l++)
      for (int j = 0; j < 3; j++) mx = std::max(std::fabs(P.Hll[9 * (size_t)l + j * 4]), mx);
    return mx;
  }
  void push() { pose_stack.push_back(P.pose); pt_stack.push_back(P.pt); }
  void pop() { P.pose = pose_stack.back(); P.pt = pt_stack.back(); discard_top(); }
  void discard_top() { pose_stack.pop_back(); pt_stack.pop_back(); }
  bool solve(double lambda) {  // setLambda(lambda) + solve(): x is left stale when the factorisation fails
    const lba_graph_view* g = P.g;
    schur(P, lambda, S, bs, Dinv, nullptr);
    const bool ok = skyline_ldlt(S, n, first, D);
    if (!ok) return false;
    for (int i = 0; i < n; i++) P.x[i] = bs[i];
    skyline_solve(S, first, D, n, P.x.data());
    for (int l = 0; l < g->n_mp; l++) {  // xl = Dinv (bl - W^T xp)
      double c[3] = {P.b[(size_t)n + 3 * (size_t)l], P.b[(size_t)n + 3 * (size_t)l + 1], P.b[(size_t)n + 3 * (size_t)l + 2]};
      for (int e : P.lm_edges[l]) {
        const int f = P.free_idx[g->e_kf[e]];
        if (f < 0) continue;
        const double* We = &P.W[18 * (size_t)e];
        for (int j = 0; j < 3; j++)
          for (int i = 0; i < 6; i++) c[j] -= We[i * 3 + j] * P.x[6 * f + i];
      }
      const double* Di = &Dinv[9 * (size_t)l];
      for (int i = 0; i < 3; i++) P.x[(size_t)n + 3 * (size_t)l + i] = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
    }
    return true;
  }
  void update(const double* x) {  // oplus on every free vertex (sparse_optimizer.cpp:421-432)
    const lba_graph_view* g = P.g;
    for (int k = 0; k < g->n_kf; k++)
      if (P.free_idx[k] >= 0) P.pose[k] = se3_exp_mul(&x[6 * (size_t)P.free_idx[k]], P.pose[k]);
    for (size_t i = 0; i < 3 * (size_t)g->n_mp; i++) P.pt[i] += x[(size_t)n + i];
  }
  double scale(double lambda) const {  // computeScale
    double sc = 0;
    for (size_t j = 0; j < nvec; j++) sc += P.x[j] * (lambda * P.x[j] + P.b[j]);
    return sc;
  }
  void results(double* kf_pose_out, double* mp_pos_out) const {
    for (int k = 0; k < P.g->n_kf; k++) {
      double* o = kf_pose_out + 7 * k;
      o[0] = P.pose[k].r.x; o[1] = P.pose[k].r.y; o[2] = P.pose[k].r.z; o[3] = P.pose[k].r.w;
      o[4] = P.pose[k].t[0]; o[5] = P.pose[k].t[1]; o[6] = P.pose[k].t[2];
    }
    memcpy(mp_pos_out, P.pt.data(), sizeof(double) * 3 * (size_t)P.g->n_mp);
  }
};

const orc_lm_ops* lba_ops() {
  static const orc_lm_ops ops = {
      [](void* h) { static_cast<Stepper*>(h)->P.compute_errors(); },
      [](void* h) { return static_cast<Stepper*>(h)->P.robust_chi2(); },
      [](void* h) { static_cast<Stepper*>(h)->P.build_system(); },
      // free vertices: poses first (6-dimensional), then landmarks (3)
      [](void* h) { Stepper* s = static_cast<Stepper*>(h); return s->nf + s->P.g->n_mp; },
      [](void* h, int v) { return v < static_cast<Stepper*>(h)->nf ? 6 : 3; },
      [](void* h, int v, int i, int j) {
        Stepper* s = static_cast<Stepper*>(h);
        return v < s->nf ? s->P.Hpp[36 * (size_t)v + i * 6 + j] : s->P.Hll[9 * (size_t)(v - s->nf) + i * 3 + j];
      },
      [](void* h, double lambda) { return static_cast<Stepper*>(h)->solve(lambda) ? 1 : 0; },
      [](void* h) { return static_cast<Stepper*>(h)->P.x.data(); },
      [](void* h) { return static_cast<Stepper*>(h)->P.b.data(); },
      [](void* h) { return static_cast<Stepper*>(h)->nvec; },
      [](void* h, const double* x) { static_cast<Stepper*>(h)->update(x); },
      [](void* h) { static_cast<Stepper*>(h)->push(); },
      [](void* h) { static_cast<Stepper*>(h)->pop(); },
      [](void* h) { static_cast<Stepper*>(h)->discard_top(); },
      [](void* h) { Stepper* s = static_cast<Stepper*>(h); return (s->stop && *s->stop) ? 1 : 0; },
  };
  return &ops;
}

}  // namespace

extern "C" const orc_lm_ops* orc_lba_stepper_ops() { return lba_ops(); }

extern "C" {

// Full optimize(max_iters).  Outputs as lba_solve of include/orb_b200.h.  lm = the driver of the LM control law (NULL: the
// restated one, orc_lm_restated; tests/test_ref_lm.py passes the reference's own object code, oracle/_ref/libref_lm.so).
int orc_lba_solve_lm(const lba_graph_view* g, const volatile uint8_t* stop, int max_iters, double lambda_init,
                     double* kf_pose_out, double* mp_pos_out, double* chi2_out, uint8_t* depth_pos_out,
                     lba_stats* stats, double* trace /* per trial: lambda, tempChi, rho, accepted; cap 4*128 */, orc_lm_driver lm) {
  auto t_begin = std::chrono::steady_clock::now();
  if (!lm) lm = orc_lm_restated;
  Stepper st(g);
  st.stop = stop;
  Problem& P = st.P;
  orc_lm_report rep = {};
  rep.trace = trace;
  const int iters = lm(orc_lba_stepper_ops(), &st, max_iters, lambda_init, &rep);
  const int stopped = (stop && *stop) ? 1 : 0;
  st.results(kf_pose_out, mp_pos_out);
  for (int e = 0; e < g->n_edges; e++) {
    if (chi2_out) chi2_out[e] = P.chi2(e);  // errors of the last evaluated trial
    if (depth_pos_out) {
      double Xc[3];
      se3_map(P.pose[g->e_kf[e]], &P.pt[3 * (size_t)g->e_mp[e]], Xc);
      // EdgeSE3ProjectXYZToBody::isDepthPositive (OptimizableTypes.h:135-139): depth in the second camera
      if (P.is_body(e)) se3_map(P.trw(g->e_kf[e]), &P.pt[3 * (size_t)g->e_mp[e]], Xc);
      depth_pos_out[e] = Xc[2] > 0.0;
    }
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->iterations = iters; stats->trials = rep.trials; stats->stopped = stopped;
    stats->chi2_initial = rep.chi_first; stats->chi2_final = rep.chi_final; stats->lambda_final = rep.lambda_final;
    stats->n_free_kf = st.nf;
    stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  }
  return iters;
}

// The Huber kernel of orc_se3.h for delta = (float)th (Optimizer.cc:1275-1276 hands g2o a float): rho[0], rho[1] --
// tests/test_ref_edges.py holds it against RobustKernelHuber's object code (oracle/_ref/libref_g2o.so).
void orc_huber(float th, double e, double* rho2) { Huber(th).robustify(e, rho2[0], rho2[1]); }

// se3_exp_mul of orc_se3.h -- VertexSE3Expmap::oplusImpl: SE3Quat::exp(update) * estimate -- on its own:
// tests/test_ref_edges.py holds it against the reference's se3quat.h as object code (oracle/_ref/libref_g2o.so).
void orc_se3_oplus(const double* pose7, const double* update6, double* out7) {
  SE3 T; T.r = Quat{pose7[0], pose7[1], pose7[2], pose7[3]}; quat_normalize(T.r);
  T.t[0] = pose7[4]; T.t[1] = pose7[5]; T.t[2] = pose7[6];
  const SE3 o = se3_exp_mul(update6, T);
  out7[0] = o.r.x; out7[1] = o.r.y; out7[2] = o.r.z; out7[3] = o.r.w; out7[4] = o.t[0]; out7[5] = o.t[1]; out7[6] = o.t[2];
}

int orc_lba_solve(const lba_graph_view* g, const volatile uint8_t* stop, int max_iters, double lambda_init,
                  double* kf_pose_out, double* mp_pos_out, double* chi2_out, uint8_t* depth_pos_out,
                  lba_stats* stats, double* trace) {
  return orc_lba_solve_lm(g, stop, max_iters, lambda_init, kf_pose_out, mp_pos_out, chi2_out, depth_pos_out, stats, trace, nullptr);
}

// The normal equations at the input estimates as build_system() leaves them: Hpp n_kf x 36 (zero rows for fixed
// keyframes), Hll n_mp x 9, W n_edges x 18 (6 x 3 per edge: B^T wOmega A), bp n_kf x 6, bl n_mp x 3 --
// tests/test_ref_edges.py holds them against g2o's constructQuadraticForm() object code (oracle/_ref/libref_g2o.so).
int orc_lba_system(const lba_graph_view* g, double* Hpp, double* Hll, double* W, double* bp, double* bl) {
  Problem P(g);
  P.compute_errors();
  P.build_system();
  memset(Hpp, 0, sizeof(double) * 36 * (size_t)g->n_kf);
  memset(bp, 0, sizeof(double) * 6 * (size_t)g->n_kf);
  for (int k = 0; k < g->n_kf; k++) {
    const int f = P.free_idx[k];
    if (f < 0) continue;
    memcpy(Hpp + 36 * (size_t)k, &P.Hpp[36 * (size_t)f], sizeof(double) * 36);
    memcpy(bp + 6 * (size_t)k, &P.b[6 * (size_t)f], sizeof(double) * 6);
  }
  memcpy(Hll, P.Hll.data(), sizeof(double) * 9 * (size_t)g->n_mp);
  memcpy(W, P.W.data(), sizeof(double) * 18 * (size_t)g->n_edges);
  memcpy(bl, &P.b[(size_t)6 * P.nf], sizeof(double) * 3 * (size_t)g->n_mp);
  return P.nf;
}

// One edge at the input estimates: err (3, third entry 0 for 2-D edges), A = d err / d point (d x 3), B = d err / d pose
// (d x 6), isDepthPositive -- what tests/test_ref_edges.py holds against the reference's own computeError() /
// linearizeOplus() (oracle/_ref/libref_edges.so).
int orc_lba_edge(const lba_graph_view* g, int e, double* err3, double* A9, double* B18, uint8_t* depth_pos) {
  if (!g || e < 0 || e >= g->n_edges) return -1;
  Problem P(g);
  P.compute_errors();
  for (int i = 0; i < 3; i++) err3[i] = P.err[3 * (size_t)e + i];
  for (int i = 0; i < 9; i++) A9[i] = 0;
  for (int i = 0; i < 18; i++) B18[i] = 0;
  P.edge_jac(e, A9, B18);
  double Xc[3];
  se3_map(P.pose[g->e_kf[e]], &P.pt[3 * (size_t)g->e_mp[e]], Xc);
  if (P.is_body(e)) se3_map(P.trw(g->e_kf[e]), &P.pt[3 * (size_t)g->e_mp[e]], Xc);
  *depth_pos = Xc[2] > 0.0;
  return 0;
}

// One linearisation at the input estimates: robust chi2, and for damping
// `lambda` the reduced system restricted to landmarks with lm_mask != 0 (NULL =
// all).  H_pp/b_p of edges whose landmark is masked out are excluded too, so
// shard results add up to the full system.  S is 6nf x 6nf row-major.
int orc_lba_reduced_system(const lba_graph_view* g, double lambda, const uint8_t* lm_mask, double* S_out,
                           double* bs_out, double* chi2_robust) {
  Problem P(g);
  P.compute_errors();
  if (chi2_robust) {
    double chi = 0, r0, r1;
    for (int e = 0; e < g->n_edges; e++) {
      if (lm_mask && !lm_mask[g->e_mp[e]]) continue;
      (P.is_stereo(e) ? P.hs : P.hm).robustify(P.chi2(e), r0, r1);
      chi += r0;
    }
    *chi2_robust = chi;
  }
  if (lm_mask) {
    // rebuild with masked-out edges removed: emulate by zero weight
    std::vector<float> w(g->e_inv_sigma2, g->e_inv_sigma2 + g->n_edges);
    lba_graph_view g2 = *g;
    for (int e = 0; e < g->n_edges; e++)
      if (!lm_mask[g->e_mp[e]]) w[e] = 0.f;
    g2.e_inv_sigma2 = w.data();
    Problem Q(&g2);
    Q.compute_errors();
    Q.build_system();
    std::vector<double> S, bs, Dinv;
    schur(Q, lambda, S, bs, Dinv, lm_mask);
    const int n = 6 * Q.nf;
    // lambda on the pose diagonal is added once per call: remove it so shards sum, caller adds it back
    for (int i = 0; i < n; i++) S[(size_t)i * n + i] -= lambda;
    memcpy(S_out, S.data(), sizeof(double) * (size_t)n * n);
    memcpy(bs_out, bs.data(), sizeof(double) * n);
    return n;
  }
  P.build_system();
  std::vector<double> S, bs, Dinv;
  schur(P, lambda, S, bs, Dinv, nullptr);
  const int n = 6 * P.nf;
  for (int i = 0; i < n; i++) S[(size_t)i * n + i] -= lambda;
  memcpy(S_out, S.data(), sizeof(double) * (size_t)n * n);
  memcpy(bs_out, bs.data(), sizeof(double) * n);
  return n;
}

}  // extern "C"
