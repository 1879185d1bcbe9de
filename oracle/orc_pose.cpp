// TEST INFRASTRUCTURE ONLY (see orc_common.h).  fp64 CPU restatement of
// Optimizer::PoseOptimization (reference src/Optimizer.cc:814-1115; SURVEY.md
// 8(f-2)) for the Pinhole single-camera layout (`!pFrame->mpCamera2`): one
// VertexSE3Expmap, unary EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose
// edges, BlockSolver_6_3 + LinearSolverDense + g2o's Levenberg-Marquardt.
// PARITY UNPINNED BY THE REFERENCE (no tests ship with it); pinned by an
// independent numpy/scipy check of the normal equations and a finite-difference
// check of the Jacobians in tests/test_pose_oracle.py.
//
// Restates (paths relative to /root/reference):
//   src/Optimizer.cc:814-1115                                   edge set-up, 4 rounds, chi2 classification
//   include/OptimizableTypes.h:41-45, src/OptimizableTypes.cpp:49-63   mono edge error / Jacobian
//   src/CameraModels/Pinhole.cpp:42-48, 71-81                    project / projectJac (float parameters)
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:339-346, 375-404  stereo edge (float invz in the error)
//   Thirdparty/g2o/g2o/core/base_unary_edge.hpp:44-70            constructQuadraticForm
//   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-194  LM control
//   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419         optimize()
//   Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:64-112      dense LDLT, isPositive()
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/orb_b200.h"
#include "orc_se3.h"
#include "orc_lm_ops.h"

namespace {

// Eigen::LDLT (diagonal pivoting) of a 6x6 SPD matrix; false when a pivot is negative
// (isPositive() == false -> LinearSolverDense::solve fails) or vanishes.
bool ldlt6_solve(const double Hin[36], const double b[6], double x[6]) {
  double A[36];
  memcpy(A, Hin, sizeof(A));
  int perm[6] = {0, 1, 2, 3, 4, 5};
  double D[6];
  for (int k = 0; k < 6; k++) {
    int piv = k;
    for (int i = k + 1; i < 6; i++)
      if (std::fabs(A[i * 6 + i]) > std::fabs(A[piv * 6 + piv])) piv = i;
    if (piv != k) {  // symmetric row/column swap
      for (int j = 0; j < 6; j++) std::swap(A[k * 6 + j], A[piv * 6 + j]);
      for (int j = 0; j < 6; j++) std::swap(A[j * 6 + k], A[j * 6 + piv]);
      std::swap(perm[k], perm[piv]);
    }
    const double d = A[k * 6 + k];
    if (!(d > 0)) return false;
    D[k] = d;
    for (int i = k + 1; i < 6; i++) A[i * 6 + k] /= d;
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j < 6; j++) A[i * 6 + j] -= A[i * 6 + k] * d * A[j * 6 + k];
  }
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = b[perm[i]];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * 6 + j] * y[j];
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) y[i] -= A[j * 6 + i] * y[j];
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
  return true;
}

struct PoseProblem {
  const pose_opt_view* v;
  SE3 T;
  std::vector<uint8_t> level;   // 0 active, 1 excluded (setLevel)
  std::vector<double> err;      // 3 per edge; what e->chi2() reads
  bool robust = true;
  Huber hm, hs;
  double H[36], b[6];

  explicit PoseProblem(const pose_opt_view* pv)
      : v(pv), hm((float)std::sqrt(5.991)), hs((float)std::sqrt(7.815)) {  // Optimizer.cc:851-852
    level.assign(v->n, 0);
    err.assign(3 * (size_t)v->n, 0.0);
  }
  bool stereo(int e) const { return v->obs[3 * e + 2] >= 0; }  // mvuRight[i] < 0 -> monocular (:867)

  void compute_error(int e) {
    const double Xw[3] = {v->xw[3 * e], v->xw[3 * e + 1], v->xw[3 * e + 2]};  // GetWorldPos().cast<double>()
    double Xc[3];
    se3_map(T, Xw, Xc);
    double* r = &err[3 * (size_t)e];
    const double ou = v->obs[3 * e], ov = v->obs[3 * e + 1];
    if (stereo(e)) {
      const double fx = v->fx, fy = v->fy, cx = v->cx, cy = v->cy, bf = v->bf;  // e->fx = pFrame->fx ... (:906-910)
      const float invz = (float)(1.0 / Xc[2]);   // :340 `1.0f/trans_xyz[2]`: double quotient, rounded to float
      const double pu = Xc[0] * invz * fx + cx;
      const double pv = Xc[1] * invz * fy + cy;
      r[0] = ou - pu; r[1] = ov - pv; r[2] = (double)v->obs[3 * e + 2] - (pu - bf * invz);
    } else {
      r[0] = ou - (v->fx * Xc[0] / Xc[2] + v->cx);
      r[1] = ov - (v->fy * Xc[1] / Xc[2] + v->cy);
      r[2] = 0;
    }
  }
  void compute_active_errors() {
    for (int e = 0; e < v->n; e++) if (!level[e]) compute_error(e);
  }
  double chi2(int e) const {
    const double s = v->inv_sigma2[e];
    const double* r = &err[3 * (size_t)e];
    double c = r[0] * (s * r[0]) + r[1] * (s * r[1]);
    if (stereo(e)) c += r[2] * (s * r[2]);
    return c;
  }
  double active_robust_chi2() const {
    double chi = 0, r0, r1;
    for (int e = 0; e < v->n; e++) {
      if (level[e]) continue;
      if (robust) { (stereo(e) ? hs : hm).robustify(chi2(e), r0, r1); chi += r0; }
      else chi += chi2(e);
    }
    return chi;
  }
  void jacobian(int e, double B[18]) const {  // d x 6 row-major
    const double Xw[3] = {v->xw[3 * e], v->xw[3 * e + 1], v->xw[3 * e + 2]};
    double Xc[3];
    se3_map(T, Xw, Xc);
    const double x = Xc[0], y = Xc[1];
    if (stereo(e)) {
      const double fx = v->fx, fy = v->fy, bf = v->bf;
      const double invz = 1.0 / Xc[2], invz_2 = invz * invz;
      B[0] = x * y * invz_2 * fx; B[1] = -(1 + (x * x * invz_2)) * fx; B[2] = y * invz * fx;
      B[3] = -invz * fx; B[4] = 0; B[5] = x * invz_2 * fx;
      B[6] = (1 + y * y * invz_2) * fy; B[7] = -x * y * invz_2 * fy; B[8] = -x * invz * fy;
      B[9] = 0; B[10] = -invz * fy; B[11] = y * invz_2 * fy;
      B[12] = B[0] - bf * y * invz_2; B[13] = B[1] + bf * x * invz_2; B[14] = B[2];
      B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * invz_2;
    } else {
      const double z = Xc[2];
      const double J[6] = {-(v->fx / z), -0., -(-v->fx * x / (z * z)), -0., -(v->fy / z), -(-v->fy * y / (z * z))};
      const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
      for (int r = 0; r < 2; r++)
        for (int c = 0; c < 6; c++) B[r * 6 + c] = J[r * 3] * D[c] + J[r * 3 + 1] * D[6 + c] + J[r * 3 + 2] * D[12 + c];
    }
  }
  void build_system() {
    std::fill(H, H + 36, 0.0);
    std::fill(b, b + 6, 0.0);
    for (int e = 0; e < v->n; e++) {
      if (level[e]) continue;
      const int d = stereo(e) ? 3 : 2;
      double B[18];
      jacobian(e, B);
      double rho0 = 0, rho1 = 1;
      if (robust) (stereo(e) ? hs : hm).robustify(chi2(e), rho0, rho1);
      const double s = v->inv_sigma2[e];
      const double ws = rho1 * s;
      const double* r = &err[3 * (size_t)e];
      for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) {
          double acc = 0;
          for (int q = 0; q < d; q++) acc += B[q * 6 + i] * ws * B[q * 6 + j];
          H[i * 6 + j] += acc;
        }
        double acc = 0;
        for (int q = 0; q < d; q++) acc += B[q * 6 + i] * (s * r[q]);
        b[i] -= rho1 * acc;
      }
    }
  }
  // ---- what g2o's LM driver calls (orc_lm_ops.h)
  double x[6] = {0, 0, 0, 0, 0, 0};  // the solver's _x survives a failed solve
  std::vector<SE3> stack;
  bool solve(double lambda) {        // BlockSolverX + LinearSolverDense: (H + lambda I) x = b by Eigen::LDLT
    double Hl[36];
    memcpy(Hl, H, sizeof(Hl));
    for (int j = 0; j < 6; j++) Hl[j * 7] += lambda;
    return ldlt6_solve(Hl, b, x);    // a failed solve leaves x as it was
  }
};

const orc_lm_ops* pose_ops() {
  static const orc_lm_ops ops = {
      [](void* h) { static_cast<PoseProblem*>(h)->compute_active_errors(); },
      [](void* h) { return static_cast<PoseProblem*>(h)->active_robust_chi2(); },
      [](void* h) { static_cast<PoseProblem*>(h)->build_system(); },
      [](void*) { return 1; },
      [](void*, int) { return 6; },
      [](void* h, int, int i, int j) { return static_cast<PoseProblem*>(h)->H[i * 6 + j]; },
      [](void* h, double lambda) { return static_cast<PoseProblem*>(h)->solve(lambda) ? 1 : 0; },
      [](void* h) { return static_cast<PoseProblem*>(h)->x; },
      [](void* h) { return static_cast<PoseProblem*>(h)->b; },
      [](void*) { return (size_t)6; },
      [](void* h, const double* x) { PoseProblem* P = static_cast<PoseProblem*>(h); P->T = se3_exp_mul(x, P->T); },
      [](void* h) { PoseProblem* P = static_cast<PoseProblem*>(h); P->stack.push_back(P->T); },
      [](void* h) { PoseProblem* P = static_cast<PoseProblem*>(h); P->T = P->stack.back(); P->stack.pop_back(); },
      [](void* h) { static_cast<PoseProblem*>(h)->stack.pop_back(); },
      nullptr,
  };
  return &ops;
}

}  // namespace

// The restated control law of optimization_algorithm_levenberg.cpp:61-169 (+ sparse_optimizer.cpp:354-419's loop around it)
// over an orc_lm_ops table: the driver every oracle optimiser uses unless it is handed another one.
extern "C" int orc_lm_restated(const orc_lm_ops* ops, void* h, int max_iters, double lambda_init, orc_lm_report* rep) {
  double lambda = -1, ni = 2, currentChi = 0, chi_first = 0;
  int nBad = 0, iters = 0;
  const size_t nvec = ops->vector_size(h);
  auto terminate = [&]() { return ops->terminate && ops->terminate(h); };
  for (int it = 0; it < max_iters && !terminate(); it++) {
    ops->compute_errors(h);
    currentChi = ops->robust_chi2(h);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) chi_first = currentChi;
    ops->build_system(h);
    if (it == 0) {
      if (lambda_init > 0) lambda = lambda_init;
      else {  // computeLambdaInit: tau * max |H_jj| over all vertices
        double mx = 0;
        for (int v = 0, nv = ops->n_vertices(h); v < nv; v++)
          for (int j = 0, d = ops->vertex_dim(h, v); j < d; j++) mx = std::max(std::fabs(ops->hessian(h, v, j, j)), mx);
        lambda = 1e-5 * mx;
      }
      ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      ops->push(h);
      const bool ok2 = ops->solve(h, lambda) != 0;
      ops->update(h, ops->x(h));  // g2o updates even when the solve failed; x then holds stale values
      ops->compute_errors(h);
      tempChi = ops->robust_chi2(h);
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;  // computeScale
      const double *x = ops->x(h), *b = ops->b(h);
      for (size_t j = 0; j < nvec; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      const bool accept = rho > 0 && std::isfinite(tempChi);
      if (rep) {
        if (rep->trace && rep->trace_rows < 128) {
          double* r = rep->trace + 4 * (size_t)rep->trace_rows;
          r[0] = lambda; r[1] = tempChi; r[2] = rho; r[3] = accept ? 1 : 0;
        }
        rep->trace_rows++;
        rep->trials++;
      }
      if (accept) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        lambda *= std::max(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
        ops->discard_top(h);
      } else {
        lambda *= ni;
        ni *= 2;
        ops->pop(h);
      }
      qmax++;
    } while (rho < 0 && qmax < 10 && !terminate());
    iters++;
    if (qmax == 10 || rho == 0) break;  // Terminate
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
    else nBad = 0;
    if (nBad >= 3) break;
  }
  if (rep) { rep->iterations = iters; rep->lambda_final = lambda; rep->chi_first = chi_first; rep->chi_final = currentChi; }
  return iters;
}

extern "C" {

// One edge of the frame's graph at its input pose: err (3), Jacobian d err / d pose (d x 6 row-major) -- held against the
// reference's EdgeSE3ProjectXYZOnlyPose::computeError / linearizeOplus object code (tests/test_ref_edges.py).
int orc_pose_edge(const pose_opt_view* v, int e, double* err3, double* B18) {
  if (!v || e < 0 || e >= v->n) return -1;
  PoseProblem P(v);
  P.T.r = Quat{v->pose[0], v->pose[1], v->pose[2], v->pose[3]};
  quat_normalize(P.T.r);
  P.T.t[0] = v->pose[4]; P.T.t[1] = v->pose[5]; P.T.t[2] = v->pose[6];
  P.compute_error(e);
  for (int i = 0; i < 3; i++) err3[i] = P.err[3 * (size_t)e + i];
  for (int i = 0; i < 18; i++) B18[i] = 0;
  P.jacobian(e, B18);
  return 0;
}

// The 6 x 6 system build_system() forms at the input pose with every edge active and robust (the first linearisation of
// the first round) -- tests/test_ref_edges.py holds it against BaseUnaryEdge::constructQuadraticForm's object code.
int orc_pose_system(const pose_opt_view* v, double* H36, double* b6) {
  PoseProblem P(v);
  P.T.r = Quat{v->pose[0], v->pose[1], v->pose[2], v->pose[3]};
  quat_normalize(P.T.r);
  P.T.t[0] = v->pose[4]; P.T.t[1] = v->pose[5]; P.T.t[2] = v->pose[6];
  P.compute_active_errors();
  P.build_system();
  memcpy(H36, P.H, sizeof(P.H));
  memcpy(b6, P.b, sizeof(P.b));
  return 0;
}

// Returns nInitialCorrespondences - nBad (Optimizer.cc:1114).  pose_out: quaternion xyzw + translation
// (SE3quat_recov before the cast to float); outlier_out[n] = mvbOutlier of the edges; chi2_out[n] (optional)
// = the chi2 each edge was last classified with; stats_out (optional) = {rounds run, LM iterations, LM trials}.
// lm = the driver that runs optimizer.optimize(10) on the frame's graph (NULL: orc_lm_restated); trace (optional): the
// trials of all rounds, 4 doubles each (lambda, chi2, -, accepted), cap 128 rows.
int orc_pose_optimize_lm(const pose_opt_view* v, double* pose_out, uint8_t* outlier_out, double* chi2_out,
                         int* stats_out, orc_lm_driver lm, double* trace) {
  if (!lm) lm = orc_lm_restated;
  int trace_rows = 0;
  const int n = v->n;
  if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = 0;
  SE3 T0;
  T0.r = Quat{v->pose[0], v->pose[1], v->pose[2], v->pose[3]};
  quat_normalize(T0.r);  // SE3Quat(q, t) constructor
  T0.t[0] = v->pose[4]; T0.t[1] = v->pose[5]; T0.t[2] = v->pose[6];
  auto write_pose = [&](const SE3& T) {
    pose_out[0] = T.r.x; pose_out[1] = T.r.y; pose_out[2] = T.r.z; pose_out[3] = T.r.w;
    pose_out[4] = T.t[0]; pose_out[5] = T.t[1]; pose_out[6] = T.t[2];
  };
  for (int e = 0; e < n; e++) outlier_out[e] = 0;  // mvbOutlier[i] = false (:869, :903)
  if (n < 3) { write_pose(T0); return 0; }         // :1000-1001 (the frame pose is left untouched)
  PoseProblem P(v);
  const float chi2Mono = 5.991, chi2Stereo = 7.815;  // :1005-1006
  int nBad = 0;
  for (int round = 0; round < 4; round++) {
    P.T = T0;  // every round restarts from the frame pose (:1012-1013)
    // ---- optimizer.optimize(10)
    {
      orc_lm_report rep = {};
      rep.trace = trace; rep.trace_rows = trace_rows;
      const int iters = lm(pose_ops(), &P, 10, 0.0, &rep);
      trace_rows = rep.trace_rows;
      if (stats_out) { stats_out[1] += iters; stats_out[2] += rep.trials; }
    }
    // ---- classification (:1018-1096)
    nBad = 0;
    for (int e = 0; e < n; e++) {
      if (outlier_out[e]) P.compute_error(e);  // excluded edges carry a stale error
      const float chi2 = (float)P.chi2(e);
      if (chi2_out) chi2_out[e] = P.chi2(e);
      if (chi2 > (P.stereo(e) ? chi2Stereo : chi2Mono)) { outlier_out[e] = 1; P.level[e] = 1; nBad++; }
      else { outlier_out[e] = 0; P.level[e] = 0; }
    }
    if (round == 2) P.robust = false;  // e->setRobustKernel(0) (:1041-1042)
    if (stats_out) stats_out[0]++;
    if (n < 10) break;                 // optimizer.edges().size() < 10 (:1098-1099)
  }
  write_pose(P.T);
  return n - nBad;
}

int orc_pose_optimize(const pose_opt_view* v, double* pose_out, uint8_t* outlier_out, double* chi2_out,
                      int* stats_out) {
  return orc_pose_optimize_lm(v, pose_out, outlier_out, chi2_out, stats_out, nullptr, nullptr);
}

}  // extern "C"
