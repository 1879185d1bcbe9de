// TEST INFRASTRUCTURE ONLY (see orc_common.h).  fp64 CPU restatement of the optimisation core of
// Optimizer::LocalInertialBA (reference src/Optimizer.cc:2383-2958): g2o
// Levenberg-Marquardt (user lambda init) with BlockSolverX over VertexPose (ImuCamPose) /
// VertexVelocity / VertexGyroBias / VertexAccBias and marginalised VertexSBAPointXYZ, edges
// EdgeMono / EdgeStereo / EdgeInertial / EdgeGyroRW / EdgeAccRW.
// PARITY UNPINNED BY THE REFERENCE; pinned by a numerical-gradient check of the whole system and by
// recovery of a synthetic visual-inertial trajectory (tests/test_lia_oracle.py).  No CUDA twin yet.
//
// Restates (paths relative to /root/reference):
//   src/Optimizer.cc:2503-2520 (solver, lambda), :2580-2611 (inertial edges, Huber on the last one),
//   :2636-2735 (visual edges), :2748-2751 (optimize)
//   src/G2oTypes.cc:25-70 (ImuCamPose), :172-219 (Project, ProjectStereo, isDepthPositive, Update --
//     whose NormalizeRotation(Rwb) discards its result, i.e. does nothing), :349-373, :397-427
//     (EdgeMono / EdgeStereo Jacobians), :492-594 (EdgeInertial), :774-856 (SO3 helpers)
//   include/G2oTypes.h:68-71 (NormalizeRotation = U V^T), :342-365, :636-700 (random-walk edges)
//   src/ImuTypes.cc:298-332 (GetDeltaRotation / Velocity / Position: float, first-order bias update)
//   Thirdparty/g2o/g2o/core/base_multi_edge.hpp (constructQuadraticForm), block_solver.hpp, levenberg.cpp
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "orc_lia.h"
#include "orc_lm_ops.h"

namespace {

// ---------------------------------------------------------------- small dense helpers (row-major 3x3)
inline void mm(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  memcpy(C, T, sizeof(T));
}
inline void tr(const double* A, double* T) {
  double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
  memcpy(T, t, sizeof(t));
}
inline void mv(const double* A, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
  memcpy(o, t, sizeof(t));
}
inline void mtv(const double* A, const double* v, double* o) {  // A^T v
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
  memcpy(o, t, sizeof(t));
}
inline void skew(const double* w, double* W) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1]; W[3] = w[2]; W[4] = 0; W[5] = -w[0]; W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
inline bool inv3(const double* m, double* o) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (det == 0) return false;
  const double id = 1.0 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}
// NormalizeRotation: svd.matrixU() * svd.matrixV().transpose() = the orthogonal polar factor; computed
// with Higham's Newton iteration (quadratic, R is always close to a rotation here)
void normalize_rotation(double* R) {
  for (int it = 0; it < 20; it++) {
    double Ri[9], RiT[9], N[9];
    if (!inv3(R, Ri)) return;
    tr(Ri, RiT);
    double diff = 0;
    for (int i = 0; i < 9; i++) { N[i] = 0.5 * (R[i] + RiT[i]); diff = std::max(diff, fabs(N[i] - R[i])); }
    memcpy(R, N, sizeof(N));
    if (diff < 1e-15) break;
  }
}
void exp_so3(const double* w, double* R) {  // G2oTypes.cc:780-798
  const double d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], d = sqrt(d2);
  double W[9], W2[9];
  skew(w, W);
  mm(W, W, W2);
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = d < 1e-5 ? I + W[i] + 0.5 * W2[i] : I + W[i] * sin(d) / d + W2[i] * (1.0 - cos(d)) / d2;
  }
  normalize_rotation(R);
}
void log_so3(const double* R, double* w) {  // :800-814
  const double t = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (t - 1.0) * 0.5f;
  if (costheta > 1 || costheta < -1) return;
  const double theta = acos(costheta), s = sin(theta);
  if (fabs(s) < 1e-5) return;
  for (int i = 0; i < 3; i++) w[i] = theta * w[i] / s;
}
void right_jac(const double* v, double* J) {  // :835-849
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mm(W, W, W2);
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    J[i] = d < 1e-5 ? I : I - W[i] * (1.0 - cos(d)) / d2 + W2[i] * (d - sin(d)) / (d2 * d);
  }
}
void inv_right_jac(const double* v, double* J) {  // :821-833
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mm(W, W, W2);
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    J[i] = d < 1e-5 ? I : I + W[i] / 2 + W2[i] * (1.0 / d2 - (1.0 + cos(d)) / (2.0 * d * sin(d)));
  }
}

// n x n symmetric eigen-decomposition by cyclic Jacobi (Eigen::SelfAdjointEigenSolver stands behind
// the reference's information clamp); V columns = eigenvectors
void jacobi_eig(std::vector<double>& A, int n, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}
// dense inverse by Gauss-Jordan with partial pivoting
bool invert(std::vector<double> A, int n, std::vector<double>& Ai) {
  Ai.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Ai[i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > fabs(A[p * n + c])) p = r;
    if (A[p * n + c] == 0) return false;
    if (p != c) for (int k = 0; k < n; k++) { std::swap(A[c * n + k], A[p * n + k]); std::swap(Ai[c * n + k], Ai[p * n + k]); }
    const double d = 1.0 / A[c * n + c];
    for (int k = 0; k < n; k++) { A[c * n + k] *= d; Ai[c * n + k] *= d; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      const double f = A[r * n + c];
      if (f == 0) continue;
      for (int k = 0; k < n; k++) { A[r * n + k] -= f * A[c * n + k]; Ai[r * n + k] -= f * Ai[c * n + k]; }
    }
  }
  return true;
}
// in-place LDL^T solve of a dense SPD system; false on a non-positive pivot
bool ldlt_solve(std::vector<double> A, int n, const double* b, double* x) {
  std::vector<double> D(n);
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k] * D[k];
    if (!(d > 0)) return false;
    D[j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * n + j];
      for (int k = 0; k < j; k++) v -= A[i * n + k] * A[j * n + k] * D[k];
      A[i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= A[i * n + k] * x[k]; x[i] = v; }
  for (int i = 0; i < n; i++) x[i] /= D[i];
  for (int i = n - 1; i >= 0; i--) { double v = x[i]; for (int k = i + 1; k < n; k++) v -= A[k * n + i] * x[k]; x[i] = v; }
  return true;
}

struct Huber {
  double delta; float dsqr; bool on;
  void robustify(double e, double& r0, double& r1) const {
    if (!on || e <= dsqr) { r0 = e; r1 = 1.; }
    else { const double s = sqrt(e); r0 = 2 * s * delta - dsqr; r1 = delta / s; }
  }
};
Huber make_huber(double delta) { Huber h; h.delta = delta; h.dsqr = (float)(delta * delta); h.on = true; return h; }

struct Pose { double Rwb[9], twb[3], Rcw[9], tcw[3]; };

struct Problem {
  const lia_graph_view* g;
  std::vector<Pose> pose;
  std::vector<double> vel, bg, ba, pt;
  std::vector<int> ip, iv, ig, ia;  // offsets into the pose-side vector, -1 = fixed / absent
  int np = 0;                        // pose-side dimension
  std::vector<double> info;          // n_inertial x 81
  std::vector<double> infoG, infoA;  // n_inertial x 9
  Huber hm, hs, hi;
  double grav[3];
  // system
  std::vector<double> H, b, Hll, bl, W;  // H np x np, Hll n_mp x 9, W n_edges x 18 (6x3, pose rows)
  std::vector<double> verr, ierr;        // visual 3 / edge, inertial 15 / edge (9 + 3 + 3)
  std::vector<std::vector<int>> lm_edges;

  explicit Problem(const lia_graph_view* gv) : g(gv) {
    hm = make_huber((float)sqrt(5.991)); hs = make_huber((float)sqrt(7.815)); hi = make_huber(sqrt(16.92));  // :2646-2649, :2595
    grav[0] = 0; grav[1] = 0; grav[2] = -(double)9.81f;  // IMU::GRAVITY_VALUE is a float
    const int K = g->n_kf;
    pose.resize(K); vel.assign(g->kf_vel, g->kf_vel + 3 * (size_t)K);
    bg.assign(g->kf_bg, g->kf_bg + 3 * (size_t)K); ba.assign(g->kf_ba, g->kf_ba + 3 * (size_t)K);
    ip.assign(K, -1); iv.assign(K, -1); ig.assign(K, -1); ia.assign(K, -1);
    for (int k = 0; k < K; k++) {
      memcpy(pose[k].Rwb, g->kf_Rwb + 9 * k, 72); memcpy(pose[k].twb, g->kf_twb + 3 * k, 24);
      memcpy(pose[k].Rcw, g->kf_Rcw + 9 * k, 72); memcpy(pose[k].tcw, g->kf_tcw + 3 * k, 24);
      if (g->kf_fixed[k]) continue;
      ip[k] = np; np += 6;
      if (g->kf_has_imu[k]) { iv[k] = np; np += 3; ig[k] = np; np += 3; ia[k] = np; np += 3; }
    }
    pt.assign(g->mp_pos, g->mp_pos + 3 * (size_t)g->n_mp);
    lm_edges.resize(g->n_mp);
    for (int e = 0; e < g->n_edges; e++) lm_edges[g->e_mp[e]].push_back(e);
    verr.assign(3 * (size_t)g->n_edges, 0.0); ierr.assign(15 * (size_t)g->n_inertial, 0.0);
    info.resize(81 * (size_t)g->n_inertial); infoG.resize(9 * (size_t)g->n_inertial); infoA.resize(9 * (size_t)g->n_inertial);
    for (int i = 0; i < g->n_inertial; i++) {  // EdgeInertial ctor (:500-508), InfoG / InfoA (:2601-2609)
      const float* C = g->i_C + 225 * (size_t)i;
      std::vector<double> C9(81), I9;
      for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) C9[r * 9 + c] = C[r * 15 + c];
      invert(C9, 9, I9);
      for (int r = 0; r < 9; r++) for (int c = r + 1; c < 9; c++) { const double m = (I9[r * 9 + c] + I9[c * 9 + r]) / 2; I9[r * 9 + c] = I9[c * 9 + r] = m; }
      std::vector<double> A = I9, V;
      jacobi_eig(A, 9, V);
      double* O = &info[81 * (size_t)i];
      for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) {
        double s = 0;
        for (int k = 0; k < 9; k++) { const double ev = A[k * 9 + k] < 1e-12 ? 0.0 : A[k * 9 + k]; s += V[r * 9 + k] * ev * V[c * 9 + k]; }
        O[r * 9 + c] = g->i_last[i] ? s * 1e-2 : s;
      }
      double G3[9], A3[9];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { G3[r * 3 + c] = C[(9 + r) * 15 + 9 + c]; A3[r * 3 + c] = C[(12 + r) * 15 + 12 + c]; }
      inv3(G3, &infoG[9 * (size_t)i]); inv3(A3, &infoA[9 * (size_t)i]);
    }
    H.resize((size_t)np * np); b.resize(np); Hll.resize(9 * (size_t)g->n_mp); bl.resize(3 * (size_t)g->n_mp);
    W.resize(18 * (size_t)g->n_edges);
  }

  // IMU::Preintegrated::GetDeltaRotation / Velocity / Position (ImuTypes.cc:298-332), float arithmetic
  void preint(int i, int k1, double* dR, double* dV, double* dP, double* dbg_out) const {
    const float* b0 = g->i_bias + 6 * (size_t)i;
    float dbg[3], dba[3];
    for (int c = 0; c < 3; c++) { dbg[c] = (float)bg[3 * k1 + c] - b0[3 + c]; dba[c] = (float)ba[3 * k1 + c] - b0[c]; }
    const float *JRg = g->i_JRg + 9 * (size_t)i, *JVg = g->i_JVg + 9 * (size_t)i, *JVa = g->i_JVa + 9 * (size_t)i;
    const float *JPg = g->i_JPg + 9 * (size_t)i, *JPa = g->i_JPa + 9 * (size_t)i;
    float w[3];
    for (int r = 0; r < 3; r++) w[r] = JRg[r * 3] * dbg[0] + JRg[r * 3 + 1] * dbg[1] + JRg[r * 3 + 2] * dbg[2];
    double wd[3] = {w[0], w[1], w[2]}, E[9], R0[9];
    exp_so3(wd, E);
    for (int c = 0; c < 9; c++) R0[c] = g->i_dR[9 * (size_t)i + c];
    mm(R0, E, dR);
    normalize_rotation(dR);
    for (int c = 0; c < 9; c++) dR[c] = (double)(float)dR[c];  // the reference hands a Matrix3f over
    for (int r = 0; r < 3; r++) {
      const float v = g->i_dV[3 * (size_t)i + r] + (JVg[r * 3] * dbg[0] + JVg[r * 3 + 1] * dbg[1] + JVg[r * 3 + 2] * dbg[2]) +
                      (JVa[r * 3] * dba[0] + JVa[r * 3 + 1] * dba[1] + JVa[r * 3 + 2] * dba[2]);
      const float p = g->i_dP[3 * (size_t)i + r] + (JPg[r * 3] * dbg[0] + JPg[r * 3 + 1] * dbg[1] + JPg[r * 3 + 2] * dbg[2]) +
                      (JPa[r * 3] * dba[0] + JPa[r * 3 + 1] * dba[1] + JPa[r * 3 + 2] * dba[2]);
      dV[r] = v; dP[r] = p;
    }
    if (dbg_out) for (int c = 0; c < 3; c++) dbg_out[c] = dbg[c];
  }

  void compute_errors() {
    for (int e = 0; e < g->n_edges; e++) {  // EdgeMono / EdgeStereo::computeError
      const Pose& P = pose[g->e_kf[e]];
      double Xc[3];
      mv(P.Rcw, &pt[3 * (size_t)g->e_mp[e]], Xc);
      for (int c = 0; c < 3; c++) Xc[c] += P.tcw[c];
      const double u = g->fx * Xc[0] / Xc[2] + g->cx, v = g->fy * Xc[1] / Xc[2] + g->cy;
      double* r = &verr[3 * (size_t)e];
      const double* o = g->e_obs + 3 * (size_t)e;
      r[0] = o[0] - u; r[1] = o[1] - v;
      r[2] = g->e_stereo[e] ? o[2] - (u - (double)g->bf * (1 / Xc[2])) : 0.0;
    }
    for (int i = 0; i < g->n_inertial; i++) {  // EdgeInertial / EdgeGyroRW / EdgeAccRW::computeError
      const int k1 = g->i_kf1[i], k2 = g->i_kf2[i];
      double dR[9], dV[3], dP[3];
      preint(i, k1, dR, dV, dP, nullptr);
      const double dt = g->i_dT[i];
      double Rbw1[9], T[9], eR[9], dRt[9];
      tr(pose[k1].Rwb, Rbw1); tr(dR, dRt);
      mm(dRt, Rbw1, T); mm(T, pose[k2].Rwb, eR);
      double* r = &ierr[15 * (size_t)i];
      log_so3(eR, r);
      double a[3], c[3];
      for (int q = 0; q < 3; q++) {
        a[q] = vel[3 * k2 + q] - vel[3 * k1 + q] - grav[q] * dt;
        c[q] = pose[k2].twb[q] - pose[k1].twb[q] - vel[3 * k1 + q] * dt - grav[q] * dt * dt / 2;
      }
      mv(Rbw1, a, a); mv(Rbw1, c, c);
      for (int q = 0; q < 3; q++) { r[3 + q] = a[q] - dV[q]; r[6 + q] = c[q] - dP[q]; }
      for (int q = 0; q < 3; q++) { r[9 + q] = bg[3 * k2 + q] - bg[3 * k1 + q]; r[12 + q] = ba[3 * k2 + q] - ba[3 * k1 + q]; }
    }
  }
  double vchi2(int e) const {
    const double s = g->e_inv_sigma2[e];
    const double* r = &verr[3 * (size_t)e];
    return r[0] * (s * r[0]) + r[1] * (s * r[1]) + (g->e_stereo[e] ? r[2] * (s * r[2]) : 0.0);
  }
  static double quad(const double* O, const double* r, int n) {
    double c = 0;
    for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += O[i * n + j] * r[j]; c += r[i] * s; }
    return c;
  }
  double robust_chi2() const {
    double chi = 0, r0, r1;
    for (int e = 0; e < g->n_edges; e++) { (g->e_stereo[e] ? hs : hm).robustify(vchi2(e), r0, r1); chi += r0; }
    for (int i = 0; i < g->n_inertial; i++) {
      const double* r = &ierr[15 * (size_t)i];
      const double ci = quad(&info[81 * (size_t)i], r, 9);
      if (g->i_last[i]) { hi.robustify(ci, r0, r1); chi += r0; } else chi += ci;
      chi += quad(&infoG[9 * (size_t)i], r + 9, 3) + quad(&infoA[9 * (size_t)i], r + 12, 3);
    }
    return chi;
  }

  // H += Ja^T (w O) Jb for two variable blocks (offsets oa / ob, dims da / db), J row-major d x dim
  void add_block(int oa, int da, const double* Ja, int ob, int db_, const double* Jb, const double* O, int d, double w) {
    for (int i = 0; i < da; i++)
      for (int j = 0; j < db_; j++) {
        double s = 0;
        for (int p = 0; p < d; p++) {
          double t = 0;
          for (int q = 0; q < d; q++) t += O[p * d + q] * Jb[q * db_ + j];
          s += Ja[p * da + i] * t;
        }
        H[(size_t)(oa + i) * np + ob + j] += w * s;
        if (oa != ob) H[(size_t)(ob + j) * np + oa + i] += w * s;
      }
  }
  void add_b(int oa, int da, const double* Ja, const double* O, const double* r, int d, double w) {
    for (int i = 0; i < da; i++) {
      double s = 0;
      for (int p = 0; p < d; p++) { double t = 0; for (int q = 0; q < d; q++) t += O[p * d + q] * r[q]; s += Ja[p * da + i] * t; }
      b[oa + i] -= w * s;
    }
  }

  // EdgeMono / EdgeStereo::linearizeOplus (G2oTypes.cc:290-345): A = d err / d point (d x 3), B = d err / d pose (d x 6),
  // row-major, third row zero for a mono edge
  void vis_jac(int e, double A[9], double B[18]) const {
    const int k = g->e_kf[e], l = g->e_mp[e], d = g->e_stereo[e] ? 3 : 2;
    double Rbc[9];
    tr(g->Rcb, Rbc);
    const Pose& P = pose[k];
    double Xc[3], Xb[3];
    mv(P.Rcw, &pt[3 * (size_t)l], Xc);
    for (int c = 0; c < 3; c++) Xc[c] += P.tcw[c];
    mv(Rbc, Xc, Xb);
    for (int c = 0; c < 3; c++) Xb[c] += g->tbc[c];
    const double x = Xc[0], y = Xc[1], z = Xc[2];
    double pj[9] = {g->fx / z, 0, -g->fx * x / (z * z), 0, g->fy / z, -g->fy * y / (z * z), 0, 0, 0};
    if (d == 3) { pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + (double)g->bf * (1.0 / (z * z)); }
    double PR[9];
    mm(pj, P.Rcw, A);
    for (int c = 0; c < 9; c++) A[c] = -A[c];                       // _jacobianOplusXi = -proj_jac * Rcw
    mm(pj, g->Rcb, PR);
    const double D[18] = {0, Xb[2], -Xb[1], 1, 0, 0, -Xb[2], 0, Xb[0], 0, 1, 0, Xb[1], -Xb[0], 0, 0, 0, 1};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 6; c++) B[r * 6 + c] = PR[r * 3] * D[c] + PR[r * 3 + 1] * D[6 + c] + PR[r * 3 + 2] * D[12 + c];
  }

  void build_system() {
    std::fill(H.begin(), H.end(), 0.0); std::fill(b.begin(), b.end(), 0.0);
    std::fill(Hll.begin(), Hll.end(), 0.0); std::fill(bl.begin(), bl.end(), 0.0);
    for (int e = 0; e < g->n_edges; e++) {  // EdgeMono / EdgeStereo::linearizeOplus + constructQuadraticForm
      const int k = g->e_kf[e], l = g->e_mp[e], d = g->e_stereo[e] ? 3 : 2;
      double A[9], B[18];
      vis_jac(e, A, B);
      double r0, r1;
      (d == 3 ? hs : hm).robustify(vchi2(e), r0, r1);
      const double s = g->e_inv_sigma2[e], ws = r1 * s;
      const double* r = &verr[3 * (size_t)e];
      double* Hl = &Hll[9 * (size_t)l];
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) { double a = 0; for (int q = 0; q < d; q++) a += A[q * 3 + i] * ws * A[q * 3 + j]; Hl[i * 3 + j] += a; }
        double a = 0;
        for (int q = 0; q < d; q++) a += A[q * 3 + i] * (s * r[q]);
        bl[3 * (size_t)l + i] -= r1 * a;
      }
      double* We = &W[18 * (size_t)e];
      const int o = ip[k];
      if (o >= 0) {
        for (int i = 0; i < 6; i++) {
          for (int j = 0; j < 6; j++) { double a = 0; for (int q = 0; q < d; q++) a += B[q * 6 + i] * ws * B[q * 6 + j]; H[(size_t)(o + i) * np + o + j] += a; }
          double a = 0;
          for (int q = 0; q < d; q++) a += B[q * 6 + i] * (s * r[q]);
          b[o + i] -= r1 * a;
          for (int j = 0; j < 3; j++) { double a2 = 0; for (int q = 0; q < d; q++) a2 += B[q * 6 + i] * ws * A[q * 3 + j]; We[i * 3 + j] = a2; }
        }
      } else {
        for (int i = 0; i < 18; i++) We[i] = 0;
      }
    }
    for (int i = 0; i < g->n_inertial; i++) {  // EdgeInertial::linearizeOplus (:534-594) + BaseMultiEdge quadratic form
      const int k1 = g->i_kf1[i], k2 = g->i_kf2[i];
      double dR[9], dV[3], dP[3], dbg[3];
      preint(i, k1, dR, dV, dP, dbg);
      const double dt = g->i_dT[i];
      const double* Rwb1 = pose[k1].Rwb; const double* Rwb2 = pose[k2].Rwb;
      double Rbw1[9], dRt[9], T[9], eR[9], er[3], invJr[9];
      tr(Rwb1, Rbw1); tr(dR, dRt);
      mm(dRt, Rbw1, T); mm(T, Rwb2, eR);
      log_so3(eR, er);
      inv_right_jac(er, invJr);
      double JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
      for (int c = 0; c < 9; c++) {
        JRg[c] = g->i_JRg[9 * (size_t)i + c]; JVg[c] = g->i_JVg[9 * (size_t)i + c]; JVa[c] = g->i_JVa[9 * (size_t)i + c];
        JPg[c] = g->i_JPg[9 * (size_t)i + c]; JPa[c] = g->i_JPa[9 * (size_t)i + c];
      }
      // six Jacobian blocks, 9 x dim, row-major
      double J0[54] = {0}, J1[27] = {0}, J2[27] = {0}, J3[27] = {0}, J4[54] = {0}, J5[27] = {0};
      double Rwb2t[9], M[9], a[3], c[3], S[9];
      tr(Rwb2, Rwb2t);
      mm(Rwb2t, Rwb1, M); mm(invJr, M, M);
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[r * 6 + q] = -M[r * 3 + q];
      for (int q = 0; q < 3; q++) {
        a[q] = vel[3 * k2 + q] - vel[3 * k1 + q] - grav[q] * dt;
        c[q] = pose[k2].twb[q] - pose[k1].twb[q] - vel[3 * k1 + q] * dt - 0.5 * grav[q] * dt * dt;
      }
      mv(Rbw1, a, a); mv(Rbw1, c, c);
      skew(a, S);
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[(3 + r) * 6 + q] = S[r * 3 + q];
      skew(c, S);
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[(6 + r) * 6 + q] = S[r * 3 + q];
      for (int r = 0; r < 3; r++) J0[(6 + r) * 6 + 3 + r] = -1.0;
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J1[(3 + r) * 3 + q] = -Rbw1[r * 3 + q]; J1[(6 + r) * 3 + q] = -Rbw1[r * 3 + q] * dt; }
      {
        double v[3], Jr[9], eRt[9], X[9];
        mv(JRg, dbg, v);
        right_jac(v, Jr);
        tr(eR, eRt);
        mm(invJr, eRt, X); mm(X, Jr, X); mm(X, JRg, X);
        for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J2[r * 3 + q] = -X[r * 3 + q]; J2[(3 + r) * 3 + q] = -JVg[r * 3 + q]; J2[(6 + r) * 3 + q] = -JPg[r * 3 + q]; }
      }
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J3[(3 + r) * 3 + q] = -JVa[r * 3 + q]; J3[(6 + r) * 3 + q] = -JPa[r * 3 + q]; }
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J4[r * 6 + q] = invJr[r * 3 + q];
      mm(Rbw1, Rwb2, M);
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J4[(6 + r) * 6 + 3 + q] = M[r * 3 + q];
      for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J5[(3 + r) * 3 + q] = Rbw1[r * 3 + q];
      const double* r9 = &ierr[15 * (size_t)i];
      const double* O = &info[81 * (size_t)i];
      double w = 1.0, r0;
      if (g->i_last[i]) hi.robustify(quad(O, r9, 9), r0, w);
      const int off[6] = {ip[k1], iv[k1], ig[k1], ia[k1], ip[k2], iv[k2]};
      const int dim[6] = {6, 3, 3, 3, 6, 3};
      const double* J[6] = {J0, J1, J2, J3, J4, J5};
      for (int p = 0; p < 6; p++) {
        if (off[p] < 0) continue;
        add_b(off[p], dim[p], J[p], O, r9, 9, w);
        for (int q = p; q < 6; q++) {
          if (off[q] < 0) continue;
          add_block(off[p], dim[p], J[p], off[q], dim[q], J[q], O, 9, w);
        }
      }
      // EdgeGyroRW / EdgeAccRW: e = b2 - b1, J = (-I, I)
      const double nI[9] = {-1, 0, 0, 0, -1, 0, 0, 0, -1}, pI[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      const int og[2] = {ig[k1], ig[k2]}, oa[2] = {ia[k1], ia[k2]};
      const double* JJ[2] = {nI, pI};
      for (int p = 0; p < 2; p++) {
        if (og[p] >= 0) {
          add_b(og[p], 3, JJ[p], &infoG[9 * (size_t)i], r9 + 9, 3, 1.0);
          for (int q = p; q < 2; q++) if (og[q] >= 0) add_block(og[p], 3, JJ[p], og[q], 3, JJ[q], &infoG[9 * (size_t)i], 3, 1.0);
        }
        if (oa[p] >= 0) {
          add_b(oa[p], 3, JJ[p], &infoA[9 * (size_t)i], r9 + 12, 3, 1.0);
          for (int q = p; q < 2; q++) if (oa[q] >= 0) add_block(oa[p], 3, JJ[p], oa[q], 3, JJ[q], &infoA[9 * (size_t)i], 3, 1.0);
        }
      }
    }
  }

  // x: pose-side update (np) then points (3 n_mp); oplus of every vertex type
  void update(const std::vector<double>& x) {
    for (int k = 0; k < g->n_kf; k++) {
      if (ip[k] >= 0) {  // ImuCamPose::Update (:190-218); its NormalizeRotation(Rwb) has no effect
        Pose& P = pose[k];
        const double* u = &x[ip[k]];
        double d[3], E[9];
        mv(P.Rwb, u + 3, d);
        for (int c = 0; c < 3; c++) P.twb[c] += d[c];
        exp_so3(u, E);
        mm(P.Rwb, E, P.Rwb);
        double Rbw[9], tbw[3];
        tr(P.Rwb, Rbw);
        mv(Rbw, P.twb, tbw);
        for (int c = 0; c < 3; c++) tbw[c] = -tbw[c];
        mm(g->Rcb, Rbw, P.Rcw);
        mv(g->Rcb, tbw, P.tcw);
        for (int c = 0; c < 3; c++) P.tcw[c] += g->tcb[c];
      }
      if (iv[k] >= 0) for (int c = 0; c < 3; c++) { vel[3 * k + c] += x[iv[k] + c]; bg[3 * k + c] += x[ig[k] + c]; ba[3 * k + c] += x[ia[k] + c]; }
    }
    for (size_t i = 0; i < 3 * (size_t)g->n_mp; i++) pt[i] += x[(size_t)np + i];
  }
};

// What g2o's LM driver calls (orc_lm_ops.h) over a Problem: the block solver's Schur complement on the pose side
// (block_solver.hpp:373-486), the dense LDL^T, the landmark back-substitution, oplus, push / pop.
struct LiaStepper {
  Problem P;
  const int np;
  const size_t nvec;
  std::vector<double> x, ball, S, bs, Dinv;
  struct State { std::vector<Pose> pose; std::vector<double> vel, bg, ba, pt; };
  std::vector<State> stack;
  explicit LiaStepper(const lia_graph_view* g)
      : P(g), np(P.np), nvec((size_t)P.np + 3 * (size_t)g->n_mp), x(nvec, 0.0), ball(nvec, 0.0), bs(P.np), Dinv(9 * (size_t)g->n_mp) {}
  void build_system() {
    P.build_system();
    memcpy(ball.data(), P.b.data(), sizeof(double) * np);
    memcpy(ball.data() + np, P.bl.data(), sizeof(double) * 3 * (size_t)P.g->n_mp);
  }
  bool solve(double lambda) {
    const lia_graph_view* g = P.g;
    S = P.H;
    for (int i = 0; i < np; i++) { S[(size_t)i * np + i] += lambda; bs[i] = P.b[i]; }
    bool ok2 = true;
    for (int l = 0; l < g->n_mp; l++) {
      double Hl[9];
      memcpy(Hl, &P.Hll[9 * (size_t)l], 72);
      Hl[0] += lambda; Hl[4] += lambda; Hl[8] += lambda;
      double* Di = &Dinv[9 * (size_t)l];
      if (!inv3(Hl, Di)) { ok2 = false; break; }
      const double* bll = &P.bl[3 * (size_t)l];
      for (int ea : P.lm_edges[l]) {
        const int oa = P.ip[g->e_kf[ea]];
        if (oa < 0) continue;
        const double* Wa = &P.W[18 * (size_t)ea];
        double Y[18];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) Y[i * 3 + j] = Wa[i * 3] * Di[j] + Wa[i * 3 + 1] * Di[3 + j] + Wa[i * 3 + 2] * Di[6 + j];
        for (int i = 0; i < 6; i++) bs[oa + i] -= Y[i * 3] * bll[0] + Y[i * 3 + 1] * bll[1] + Y[i * 3 + 2] * bll[2];
        for (int eb : P.lm_edges[l]) {
          const int ob = P.ip[g->e_kf[eb]];
          if (ob < 0) continue;
          const double* Wb = &P.W[18 * (size_t)eb];
          for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++)
            S[(size_t)(oa + i) * np + ob + j] -= Y[i * 3] * Wb[j * 3] + Y[i * 3 + 1] * Wb[j * 3 + 1] + Y[i * 3 + 2] * Wb[j * 3 + 2];
        }
      }
    }
    if (ok2) ok2 = ldlt_solve(S, np, bs.data(), x.data());
    if (ok2) {
      for (int l = 0; l < g->n_mp; l++) {
        double c[3] = {P.bl[3 * (size_t)l], P.bl[3 * (size_t)l + 1], P.bl[3 * (size_t)l + 2]};
        for (int e : P.lm_edges[l]) {
          const int o = P.ip[g->e_kf[e]];
          if (o < 0) continue;
          const double* We = &P.W[18 * (size_t)e];
          for (int j = 0; j < 3; j++) for (int i = 0; i < 6; i++) c[j] -= We[i * 3 + j] * x[o + i];
        }
        const double* Di = &Dinv[9 * (size_t)l];
        for (int i = 0; i < 3; i++) x[(size_t)np + 3 * (size_t)l + i] = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
      }
    }
    return ok2;
  }
  void update(const double* xv) { P.update(std::vector<double>(xv, xv + nvec)); }
  void push() { stack.push_back(State{P.pose, P.vel, P.bg, P.ba, P.pt}); }
  void pop() { State& s = stack.back(); P.pose = s.pose; P.vel = s.vel; P.bg = s.bg; P.ba = s.ba; P.pt = s.pt; stack.pop_back(); }
};

const orc_lm_ops* lia_ops() {
  static const orc_lm_ops ops = {
      [](void* h) { static_cast<LiaStepper*>(h)->P.compute_errors(); },
      [](void* h) { return static_cast<LiaStepper*>(h)->P.robust_chi2(); },
      [](void* h) { static_cast<LiaStepper*>(h)->build_system(); },
      // (computeLambdaInit only scans these when no user lambda is set; LocalInertialBA always sets one, :2517-2520)
      [](void* h) { return 1 + static_cast<LiaStepper*>(h)->P.g->n_mp; },
      [](void* h, int v) { return v == 0 ? static_cast<LiaStepper*>(h)->np : 3; },
      [](void* h, int v, int i, int j) {
        LiaStepper* s = static_cast<LiaStepper*>(h);
        return v == 0 ? s->P.H[(size_t)i * s->np + j] : s->P.Hll[9 * (size_t)(v - 1) + i * 3 + j];
      },
      [](void* h, double lambda) { return static_cast<LiaStepper*>(h)->solve(lambda) ? 1 : 0; },
      [](void* h) { return static_cast<LiaStepper*>(h)->x.data(); },
      [](void* h) { return static_cast<LiaStepper*>(h)->ball.data(); },
      [](void* h) { return static_cast<LiaStepper*>(h)->nvec; },
      [](void* h, const double* x) { static_cast<LiaStepper*>(h)->update(x); },
      [](void* h) { static_cast<LiaStepper*>(h)->push(); },
      [](void* h) { static_cast<LiaStepper*>(h)->pop(); },
      [](void* h) { static_cast<LiaStepper*>(h)->stack.pop_back(); },
      nullptr,
  };
  return &ops;
}

}  // namespace

extern "C" {

// optimizer.optimize(opt_it).  Outputs: kf_out n_kf x 21 (Rcw 9, tcw 3, vel 3, bg 3, ba 3), mp_out n_mp x 3,
// chi2_out / depth_pos_out per visual edge (e->chi2(), isDepthPositive()), stats[6] = iterations, trials,
// err (activeRobustChi2 before), err_end (after), lambda_final, pose-side dimension.  lm = the driver of the LM control
// law (NULL: orc_lm_restated; tests/test_ref_lm.py passes the reference's own object code); trace optional (128 x 4).
int orc_lia_solve_lm(const lia_graph_view* g, double* kf_out, double* mp_out, double* chi2_out, uint8_t* depth_pos_out,
                     double* stats, orc_lm_driver lm, double* trace) {
  if (!lm) lm = orc_lm_restated;
  LiaStepper st(g);
  Problem& P = st.P;
  const int np = st.np;
  orc_lm_report rep = {};
  rep.trace = trace;
  const int iters = lm(lia_ops(), &st, g->iterations, g->lambda_init, &rep);  // setUserLambdaInit(lambda_init)
  const int trials = rep.trials;
  const double chi_first = rep.chi_first, currentChi = rep.chi_final, lambda = rep.lambda_final;
  for (int k = 0; k < g->n_kf; k++) {
    double* o = kf_out + 21 * (size_t)k;
    memcpy(o, P.pose[k].Rcw, 72); memcpy(o + 9, P.pose[k].tcw, 24);
    memcpy(o + 12, &P.vel[3 * k], 24); memcpy(o + 15, &P.bg[3 * k], 24); memcpy(o + 18, &P.ba[3 * k], 24);
  }
  memcpy(mp_out, P.pt.data(), sizeof(double) * 3 * (size_t)g->n_mp);
  for (int e = 0; e < g->n_edges; e++) {
    if (chi2_out) chi2_out[e] = P.vchi2(e);  // errors of the last evaluated trial
    if (depth_pos_out) {
      const Pose& Q = P.pose[g->e_kf[e]];
      const double* X = &P.pt[3 * (size_t)g->e_mp[e]];
      depth_pos_out[e] = (Q.Rcw[6] * X[0] + Q.Rcw[7] * X[1] + Q.Rcw[8] * X[2] + Q.tcw[2]) > 0.0;
    }
  }
  if (stats) { stats[0] = iters; stats[1] = trials; stats[2] = chi_first; stats[3] = currentChi; stats[4] = lambda; stats[5] = np; }
  return iters;
}

// For the tests: robust chi2 and the right-hand side b (pose side, then points) at the input state, and the
// robust chi2 after applying `delta` (same layout) with the vertices' own oplus.
int orc_lia_linearize(const lia_graph_view* g, const double* delta, double* chi2_out, double* b_out, double* chi2_delta_out) {
  Problem P(g);
  P.compute_errors();
  if (chi2_out) *chi2_out = P.robust_chi2();
  if (b_out) {
    P.build_system();
    memcpy(b_out, P.b.data(), sizeof(double) * P.np);
    memcpy(b_out + P.np, P.bl.data(), sizeof(double) * 3 * (size_t)g->n_mp);
  }
  if (delta && chi2_delta_out) {
    std::vector<double> x(delta, delta + P.np + 3 * (size_t)g->n_mp);
    P.update(x);
    P.compute_errors();
    *chi2_delta_out = P.robust_chi2();
  }
  return P.np;
}

int orc_lia_solve(const lia_graph_view* g, double* kf_out, double* mp_out, double* chi2_out, uint8_t* depth_pos_out,
                  double* stats) {
  return orc_lia_solve_lm(g, kf_out, mp_out, chi2_out, depth_pos_out, stats, nullptr, nullptr);
}

// One visual edge at the input state: err (3), A = d err / d point, B = d err / d pose (3 x 3 / 3 x 6 row-major, third
// row zero for EdgeMono), isDepthPositive -- tests/test_ref_edges.py holds it against the reference's EdgeMono / EdgeStereo
// (src/G2oTypes.cc, include/G2oTypes.h) as object code.
int orc_lia_edge(const lia_graph_view* g, int e, double* err3, double* A9, double* B18, uint8_t* depth_pos) {
  if (!g || e < 0 || e >= g->n_edges) return -1;
  Problem P(g);
  P.compute_errors();
  for (int i = 0; i < 3; i++) err3[i] = P.verr[3 * (size_t)e + i];
  P.vis_jac(e, A9, B18);
  if (!g->e_stereo[e]) { for (int i = 6; i < 9; i++) A9[i] = 0; for (int i = 12; i < 18; i++) B18[i] = 0; }
  const Pose& Q = P.pose[g->e_kf[e]];
  const double* X = &P.pt[3 * (size_t)g->e_mp[e]];
  if (depth_pos) *depth_pos = (Q.Rcw[6] * X[0] + Q.Rcw[7] * X[1] + Q.Rcw[8] * X[2] + Q.tcw[2]) > 0.0;
  return 0;
}

}  // extern "C"
