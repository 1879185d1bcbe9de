// Optimizer::PoseOptimization (reference src/Optimizer.cc:814-1115) on the GPU:
// motion-only bundle adjustment of one frame pose, fp64, Pinhole single-camera layout.
//
// One CTA runs the whole function for one frame -- 4 rounds x optimize(10) x up to 10 LM trials --
// inside ONE kernel launch (a batch of frames = a grid of CTAs): the threads own the edges
// (error, Jacobian, 6x6 normal-equation partials), fixed-order block reductions produce chi2 and
// H/b, thread 0 plays g2o's control law (lambda, rho, accept/reject, stop rules), the 6x6 LDL^T
// solve and the SE3 exponential.  Nothing returns to the host between trials.
//
// Follows, expression by expression (paths relative to the reference):
//   src/Optimizer.cc:851-852, 867-925, 1000-1114      edges, deltas, rounds, classification
//   include/OptimizableTypes.h:41-45, src/OptimizableTypes.cpp:49-63, src/CameraModels/Pinhole.cpp:42-48,71-81
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:339-346, 375-404 (float invz in the error)
//   Thirdparty/g2o/g2o/core/base_unary_edge.hpp:44-70, optimization_algorithm_levenberg.cpp:61-194,
//   sparse_optimizer.cpp:354-419, Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:64-112
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "se3_dev.cuh"
#include "orb_engine.h"

namespace orbb200 {

#define CUDA_TRYP(expr)                                                                \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return ORB_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

constexpr int PO_THREADS = 256, PO_WARPS = PO_THREADS / 32;

struct PoseJob {
  int n;
  size_t off;               // element offset of this frame's edges in the packed arrays
  float fx, fy, cx, cy, bf;
  double pose[7];
};

struct PoseDev {
  const PoseJob* jobs;
  const float* xw;          // packed edges x 3
  const float* obs;         // packed edges x 3
  const float* is2;         // packed edges
  double* err;              // packed edges x 3 (scratch: what e->chi2() reads)
  uint8_t* level;           // packed edges (scratch)
  uint8_t* outlier;         // packed edges (out)
  double* pose_out;         // jobs x 7
  int* inliers;             // jobs
  int* stats;               // jobs x 3
  HuberD hm, hs;
};

struct EdgeIn {
  double Xw[3];
  float u, v, ur, is2;
};

__device__ __forceinline__ EdgeIn load_edge(const PoseDev& D, size_t e) {
  EdgeIn E;
  E.Xw[0] = D.xw[3 * e]; E.Xw[1] = D.xw[3 * e + 1]; E.Xw[2] = D.xw[3 * e + 2];  // GetWorldPos().cast<double>()
  E.u = D.obs[3 * e]; E.v = D.obs[3 * e + 1]; E.ur = D.obs[3 * e + 2];
  E.is2 = D.is2[e];
  return E;
}

__device__ __forceinline__ void map_point(const double* T, const double* Xw, double* Xc) {
  const DQuat q = {T[0], T[1], T[2], T[3]};
  q_rot(q, Xw, Xc);
  Xc[0] += T[4]; Xc[1] += T[5]; Xc[2] += T[6];
}

// computeError of the two edge types; returns chi2 = r^T (invSigma2 I) r
__device__ __forceinline__ double edge_error(const PoseJob& J, const EdgeIn& E, const double* T, double* r) {
  double Xc[3];
  map_point(T, E.Xw, Xc);
  const double s = E.is2;
  if (E.ur >= 0) {
    const double fx = J.fx, fy = J.fy, cx = J.cx, cy = J.cy, bf = J.bf;
    const float invz = __double2float_rn(__ddiv_rn(1.0, Xc[2]));  // types_six_dof_expmap.cpp:340: double quotient rounded to float
    const double pu = Xc[0] * invz * fx + cx;
    const double pv = Xc[1] * invz * fy + cy;
    r[0] = (double)E.u - pu; r[1] = (double)E.v - pv; r[2] = (double)E.ur - (pu - bf * invz);
    return r[0] * (s * r[0]) + r[1] * (s * r[1]) + r[2] * (s * r[2]);
  }
  r[0] = (double)E.u - (J.fx * Xc[0] / Xc[2] + J.cx);
  r[1] = (double)E.v - (J.fy * Xc[1] / Xc[2] + J.cy);
  r[2] = 0;
  return r[0] * (s * r[0]) + r[1] * (s * r[1]);
}

__device__ __forceinline__ double chi2_of(const EdgeIn& E, const double* r) {
  const double s = E.is2;
  double c = r[0] * (s * r[0]) + r[1] * (s * r[1]);
  if (E.ur >= 0) c += r[2] * (s * r[2]);
  return c;
}

// fixed-order block reduction of NV values per thread; result valid in every thread
template <int NV>
__device__ __forceinline__ void block_reduce(double* v, double (*sm)[28]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < NV; k++) sm[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double t = 0;
    for (int w = 0; w < PO_WARPS; w++) t += sm[w][k];
    v[k] = t;
  }
}

// Eigen::LDLT (diagonal pivoting) of the damped 6x6 system; false when a pivot is not positive
// (LinearSolverDense::solve returns false -> the trial is rejected)
__device__ bool ldlt6_solve(const double* Hin, const double* b, double* x) {
  double A[36];
  for (int i = 0; i < 36; i++) A[i] = Hin[i];
  int perm[6] = {0, 1, 2, 3, 4, 5};
  double Dg[6];
  for (int k = 0; k < 6; k++) {
    int piv = k;
    for (int i = k + 1; i < 6; i++)
      if (fabs(A[i * 6 + i]) > fabs(A[piv * 6 + piv])) piv = i;
    if (piv != k) {
      for (int j = 0; j < 6; j++) { const double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t; }
      for (int j = 0; j < 6; j++) { const double t = A[j * 6 + k]; A[j * 6 + k] = A[j * 6 + piv]; A[j * 6 + piv] = t; }
      const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double d = A[k * 6 + k];
    if (!(d > 0)) return false;
    Dg[k] = d;
    for (int i = k + 1; i < 6; i++) A[i * 6 + k] /= d;
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j < 6; j++) A[i * 6 + j] -= A[i * 6 + k] * d * A[j * 6 + k];
  }
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = b[perm[i]];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * 6 + j] * y[j];
  for (int i = 0; i < 6; i++) y[i] /= Dg[i];
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) y[i] -= A[j * 6 + i] * y[j];
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
  return true;
}

__global__ void __launch_bounds__(PO_THREADS) pose_opt_kernel(const PoseDev D) {
  __shared__ double sm[PO_WARPS][28];
  __shared__ double s_T[7], s_Tn[7], s_T0[7], s_H[36], s_b[6], s_x[6];
  __shared__ double s_lambda, s_ni, s_cur, s_ini, s_rho;
  __shared__ int s_cont, s_stop, s_qmax, s_nbadlm, s_nbad, s_ok;
  const PoseJob J = D.jobs[blockIdx.x];
  const int n = J.n, tid = threadIdx.x;
  const size_t off = J.off;
  double* pose_out = D.pose_out + 7 * (size_t)blockIdx.x;
  int* stats = D.stats + 3 * (size_t)blockIdx.x;
  int rounds = 0, lm_iters = 0, lm_trials = 0;  // meaningful in thread 0

  if (tid == 0) {
    DQuat q = {J.pose[0], J.pose[1], J.pose[2], J.pose[3]};
    q_normalize(q);  // SE3Quat(q, t) constructor
    s_T0[0] = q.x; s_T0[1] = q.y; s_T0[2] = q.z; s_T0[3] = q.w;
    s_T0[4] = J.pose[4]; s_T0[5] = J.pose[5]; s_T0[6] = J.pose[6];
  }
  for (int e = tid; e < n; e += PO_THREADS) { D.outlier[off + e] = 0; D.level[off + e] = 0; }  // :869, :903
  __syncthreads();
  if (n < 3) {  // :1000-1001
    if (tid < 7) pose_out[tid] = s_T0[tid];
    if (tid == 0) { D.inliers[blockIdx.x] = 0; stats[0] = stats[1] = stats[2] = 0; }
    return;
  }
  bool robust = true;
  for (int round = 0; round < 4; round++) {
    if (tid < 7) s_T[tid] = s_T0[tid];  // every round restarts from the frame pose (:1012-1013)
    if (tid < 6) s_x[tid] = 0;
    __syncthreads();
    // ------------------------------------------------------------ optimizer.optimize(10)
    for (int it = 0; it < 10; it++) {
      // computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate
      double acc[28];
#pragma unroll
      for (int k = 0; k < 28; k++) acc[k] = 0;
      {
        double T[7];
        for (int k = 0; k < 7; k++) T[k] = s_T[k];
        for (int e = tid; e < n; e += PO_THREADS) {
          if (D.level[off + e]) continue;
          const EdgeIn E = load_edge(D, off + e);
          double r[3];
          const double c2 = edge_error(J, E, T, r);
          D.err[3 * (off + e)] = r[0]; D.err[3 * (off + e) + 1] = r[1]; D.err[3 * (off + e) + 2] = r[2];
          double rho0 = c2, rho1 = 1.0;
          if (robust) robustify(E.ur >= 0 ? D.hs : D.hm, c2, rho0, rho1);
          acc[27] += rho0;
          // linearizeOplus
          double Xc[3], B[18];
          map_point(T, E.Xw, Xc);
          const double x = Xc[0], y = Xc[1];
          int d;
          if (E.ur >= 0) {
            d = 3;
            const double fx = J.fx, fy = J.fy, bf = J.bf;
            const double invz = 1.0 / Xc[2], invz_2 = invz * invz;
            B[0] = x * y * invz_2 * fx; B[1] = -(1 + (x * x * invz_2)) * fx; B[2] = y * invz * fx;
            B[3] = -invz * fx; B[4] = 0; B[5] = x * invz_2 * fx;
            B[6] = (1 + y * y * invz_2) * fy; B[7] = -x * y * invz_2 * fy; B[8] = -x * invz * fy;
            B[9] = 0; B[10] = -invz * fy; B[11] = y * invz_2 * fy;
            B[12] = B[0] - bf * y * invz_2; B[13] = B[1] + bf * x * invz_2; B[14] = B[2];
            B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * invz_2;
          } else {
            d = 2;
            const double z = Xc[2];
            const double Jp[6] = {-(J.fx / z), -0., -(-J.fx * x / (z * z)), -0., -(J.fy / z), -(-J.fy * y / (z * z))};
            const double Dv[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
            for (int rr = 0; rr < 2; rr++)
              for (int c = 0; c < 6; c++)
                B[rr * 6 + c] = Jp[rr * 3] * Dv[c] + Jp[rr * 3 + 1] * Dv[6 + c] + Jp[rr * 3 + 2] * Dv[12 + c];
            for (int c = 12; c < 18; c++) B[c] = 0;
          }
          // constructQuadraticForm: A += B^T (rho1 Omega) B (upper triangle), b -= rho1 B^T Omega r
          const double s = E.is2, ws = rho1 * s;
          int k = 0;
          for (int i = 0; i < 6; i++) {
            for (int j = i; j < 6; j++, k++) {
              double a = 0;
              for (int q = 0; q < d; q++) a += B[q * 6 + i] * ws * B[q * 6 + j];
              acc[k] += a;
            }
            double a = 0;
            for (int q = 0; q < d; q++) a += B[q * 6 + i] * (s * r[q]);
            acc[21 + i] -= rho1 * a;
          }
        }
      }
      block_reduce<28>(acc, sm);
      if (tid == 0) {
        int k = 0;
        for (int i = 0; i < 6; i++)
          for (int j = i; j < 6; j++, k++) { s_H[i * 6 + j] = acc[k]; s_H[j * 6 + i] = acc[k]; }
        for (int i = 0; i < 6; i++) s_b[i] = acc[21 + i];
        s_cur = acc[27]; s_ini = acc[27];
        if (it == 0) {  // computeLambdaInit: tau * max |diagonal|
          double mx = 0;
          for (int j = 0; j < 6; j++) mx = fmax(fabs(s_H[j * 7]), mx);
          s_lambda = 1e-5 * mx;
          s_ni = 2; s_nbadlm = 0;
        }
        s_qmax = 0;
      }
      __syncthreads();
      // ---- trials
      while (true) {
        if (tid == 0) {
          double Hl[36];
          for (int i = 0; i < 36; i++) Hl[i] = s_H[i];
          for (int j = 0; j < 6; j++) Hl[j * 7] += s_lambda;
          double x[6];
          for (int i = 0; i < 6; i++) x[i] = s_x[i];
          s_ok = ldlt6_solve(Hl, s_b, x) ? 1 : 0;  // a failed solve leaves x as it was
          for (int i = 0; i < 6; i++) s_x[i] = x[i];
          se3_exp_mul(x, s_T, s_Tn);
        }
        __syncthreads();
        double chi[1] = {0};
        {
          double T[7];
          for (int k = 0; k < 7; k++) T[k] = s_Tn[k];
          for (int e = tid; e < n; e += PO_THREADS) {
            if (D.level[off + e]) continue;
            const EdgeIn E = load_edge(D, off + e);
            double r[3];
            const double c2 = edge_error(J, E, T, r);
            D.err[3 * (off + e)] = r[0]; D.err[3 * (off + e) + 1] = r[1]; D.err[3 * (off + e) + 2] = r[2];
            double rho0 = c2, rho1 = 1.0;
            if (robust) robustify(E.ur >= 0 ? D.hs : D.hm, c2, rho0, rho1);
            chi[0] += rho0;
          }
        }
        block_reduce<1>(chi, sm);
        if (tid == 0) {
          double tempChi = chi[0];
          if (!s_ok) tempChi = DBL_MAX;
          double rho = s_cur - tempChi;
          double scale = 0;  // computeScale (:181-188), then `scale += 1e-3` (:124-125): in that order
          for (int j = 0; j < 6; j++) scale += s_x[j] * (s_lambda * s_x[j] + s_b[j]);
          scale += 1e-3;
          rho /= scale;
          if (rho > 0 && isfinite(tempChi)) {
            double alpha = 1. - pow((2 * rho - 1), 3.0);
            alpha = fmin(alpha, 2. / 3.);
            s_lambda *= fmax(1. / 3., alpha);
            s_ni = 2;
            s_cur = tempChi;
            for (int k = 0; k < 7; k++) s_T[k] = s_Tn[k];
          } else {
            s_lambda *= s_ni;
            s_ni *= 2;
          }
          s_qmax++;
          lm_trials++;
          s_rho = rho;
          s_cont = (rho < 0 && s_qmax < 10) ? 1 : 0;
        }
        __syncthreads();
        if (!s_cont) break;
      }
      if (tid == 0) {
        lm_iters++;
        int stop = 0;
        if (s_qmax == 10 || s_rho == 0) stop = 1;
        else {
          if ((s_ini - s_cur) * 1e3 < s_ini) s_nbadlm++;
          else s_nbadlm = 0;
          if (s_nbadlm >= 3) stop = 1;
        }
        s_stop = stop;
      }
      __syncthreads();
      if (s_stop) break;  // rewritten only after the next iteration's barriers
    }
    // ------------------------------------------------------------ classification (:1018-1096)
    if (tid == 0) s_nbad = 0;
    __syncthreads();
    {
      double T[7];
      for (int k = 0; k < 7; k++) T[k] = s_T[k];
      int bad = 0;
      for (int e = tid; e < n; e += PO_THREADS) {
        const EdgeIn E = load_edge(D, off + e);
        double r[3];
        if (D.outlier[off + e]) {  // excluded edges carry a stale error: e->computeError()
          edge_error(J, E, T, r);
          D.err[3 * (off + e)] = r[0]; D.err[3 * (off + e) + 1] = r[1]; D.err[3 * (off + e) + 2] = r[2];
        } else {
          r[0] = D.err[3 * (off + e)]; r[1] = D.err[3 * (off + e) + 1]; r[2] = D.err[3 * (off + e) + 2];
        }
        const float chi2 = (float)chi2_of(E, r);
        const float th = E.ur >= 0 ? 7.815f : 5.991f;
        const int out = chi2 > th ? 1 : 0;
        D.outlier[off + e] = (uint8_t)out;
        D.level[off + e] = (uint8_t)out;
        bad += out;
      }
      if (bad) atomicAdd(&s_nbad, bad);
    }
    if (round == 2) robust = false;  // e->setRobustKernel(0) (:1041-1042)
    rounds++;
    __syncthreads();
    if (n < 10) break;  // optimizer.edges().size() < 10 (:1098-1099)
  }
  if (tid < 7) pose_out[tid] = s_T[tid];
  if (tid == 0) {
    D.inliers[blockIdx.x] = n - s_nbad;
    stats[0] = rounds; stats[1] = lm_iters; stats[2] = lm_trials;
  }
}

struct PoseOpt {
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  size_t cap_edges = 0;
  int cap_jobs = 0;
  PoseJob *d_jobs = nullptr, *h_jobs = nullptr;
  float *d_xw = nullptr, *d_obs = nullptr, *d_is2 = nullptr, *h_xw = nullptr, *h_obs = nullptr, *h_is2 = nullptr;
  double *d_err = nullptr, *d_pose = nullptr, *h_pose = nullptr;
  uint8_t *d_level = nullptr, *d_outlier = nullptr, *h_outlier = nullptr;
  int *d_inl = nullptr, *d_stats = nullptr, *h_inl = nullptr, *h_stats = nullptr;
  long long launches = 0;

  explicit PoseOpt(int dev) : device(dev) {}
  ~PoseOpt() {
    release();
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
  void release() {
    cudaFree(d_jobs); cudaFree(d_xw); cudaFree(d_obs); cudaFree(d_is2); cudaFree(d_err); cudaFree(d_pose);
    cudaFree(d_level); cudaFree(d_outlier); cudaFree(d_inl); cudaFree(d_stats);
    cudaFreeHost(h_jobs); cudaFreeHost(h_xw); cudaFreeHost(h_obs); cudaFreeHost(h_is2); cudaFreeHost(h_pose);
    cudaFreeHost(h_outlier); cudaFreeHost(h_inl); cudaFreeHost(h_stats);
    d_jobs = h_jobs = nullptr; d_xw = d_obs = d_is2 = h_xw = h_obs = h_is2 = nullptr;
    d_err = d_pose = h_pose = nullptr; d_level = d_outlier = h_outlier = nullptr;
    d_inl = d_stats = h_inl = h_stats = nullptr;
    cap_edges = 0; cap_jobs = 0;
  }
  int ensure(int jobs, size_t edges) {
    if (!stream) {
      CUDA_TRYP(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      CUDA_TRYP(cudaEventCreate(&ev0));
      CUDA_TRYP(cudaEventCreate(&ev1));
    }
    if (jobs <= cap_jobs && edges <= cap_edges) return 0;
    jobs = std::max(jobs, cap_jobs);
    edges = std::max<size_t>(std::max(edges, cap_edges), 1);
    release();
    CUDA_TRYP(cudaMalloc(&d_jobs, sizeof(PoseJob) * jobs));
    CUDA_TRYP(cudaMallocHost(&h_jobs, sizeof(PoseJob) * jobs));
    CUDA_TRYP(cudaMalloc(&d_xw, sizeof(float) * 3 * edges));
    CUDA_TRYP(cudaMalloc(&d_obs, sizeof(float) * 3 * edges));
    CUDA_TRYP(cudaMalloc(&d_is2, sizeof(float) * edges));
    CUDA_TRYP(cudaMallocHost(&h_xw, sizeof(float) * 3 * edges));
    CUDA_TRYP(cudaMallocHost(&h_obs, sizeof(float) * 3 * edges));
    CUDA_TRYP(cudaMallocHost(&h_is2, sizeof(float) * edges));
    CUDA_TRYP(cudaMalloc(&d_err, sizeof(double) * 3 * edges));
    CUDA_TRYP(cudaMalloc(&d_level, edges));
    CUDA_TRYP(cudaMalloc(&d_outlier, edges));
    CUDA_TRYP(cudaMallocHost(&h_outlier, edges));
    CUDA_TRYP(cudaMalloc(&d_pose, sizeof(double) * 7 * jobs));
    CUDA_TRYP(cudaMallocHost(&h_pose, sizeof(double) * 7 * jobs));
    CUDA_TRYP(cudaMalloc(&d_inl, sizeof(int) * jobs));
    CUDA_TRYP(cudaMalloc(&d_stats, sizeof(int) * 3 * jobs));
    CUDA_TRYP(cudaMallocHost(&h_inl, sizeof(int) * jobs));
    CUDA_TRYP(cudaMallocHost(&h_stats, sizeof(int) * 3 * jobs));
    cap_jobs = jobs; cap_edges = edges;
    return 0;
  }

  int run(int batch, const pose_opt_view* views, double* pose_out, uint8_t* const* outlier_out, int* inliers_out,
          int* stats_out) {
    if (batch <= 0 || !views || !pose_out || !outlier_out || !inliers_out) {
      set_last_error("pose_optimize: bad argument");
      return ORB_E_ARG;
    }
    size_t total = 0;
    for (int k = 0; k < batch; k++) {
      const pose_opt_view& v = views[k];
      if (v.n < 0 || (v.n > 0 && (!v.xw || !v.obs || !v.inv_sigma2 || !outlier_out[k]))) {
        set_last_error("pose_optimize: bad view " + std::to_string(k));
        return ORB_E_ARG;
      }
      total += (size_t)v.n;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
      return ORB_E_NODEVICE;
    }
    CUDA_TRYP(cudaSetDevice(device));
    int rc = ensure(batch, total);
    if (rc) return rc;
    size_t off = 0;
    for (int k = 0; k < batch; k++) {
      const pose_opt_view& v = views[k];
      PoseJob& J = h_jobs[k];
      J.n = v.n; J.off = off;
      J.fx = v.fx; J.fy = v.fy; J.cx = v.cx; J.cy = v.cy; J.bf = v.bf;
      for (int i = 0; i < 7; i++) J.pose[i] = v.pose[i];
      if (v.n) {
        memcpy(h_xw + 3 * off, v.xw, sizeof(float) * 3 * (size_t)v.n);
        memcpy(h_obs + 3 * off, v.obs, sizeof(float) * 3 * (size_t)v.n);
        memcpy(h_is2 + off, v.inv_sigma2, sizeof(float) * (size_t)v.n);
      }
      off += (size_t)v.n;
    }
    cudaStream_t s = stream;
    CUDA_TRYP(cudaMemcpyAsync(d_jobs, h_jobs, sizeof(PoseJob) * batch, cudaMemcpyHostToDevice, s));
    if (total) {
      CUDA_TRYP(cudaMemcpyAsync(d_xw, h_xw, sizeof(float) * 3 * total, cudaMemcpyHostToDevice, s));
      CUDA_TRYP(cudaMemcpyAsync(d_obs, h_obs, sizeof(float) * 3 * total, cudaMemcpyHostToDevice, s));
      CUDA_TRYP(cudaMemcpyAsync(d_is2, h_is2, sizeof(float) * total, cudaMemcpyHostToDevice, s));
    }
    PoseDev D;
    D.jobs = d_jobs; D.xw = d_xw; D.obs = d_obs; D.is2 = d_is2; D.err = d_err; D.level = d_level;
    D.outlier = d_outlier; D.pose_out = d_pose; D.inliers = d_inl; D.stats = d_stats;
    // const float deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815) (:851-852); RobustKernelHuber keeps
    // delta as double and delta^2 as float
    const float dm = (float)sqrt(5.991), ds = (float)sqrt(7.815);
    D.hm.delta = dm; D.hm.dsqr = (float)((double)dm * (double)dm);
    D.hs.delta = ds; D.hs.dsqr = (float)((double)ds * (double)ds);
    CUDA_TRYP(cudaEventRecord(ev0, s));
    pose_opt_kernel<<<batch, PO_THREADS, 0, s>>>(D);
    CUDA_TRYP(cudaEventRecord(ev1, s));
    launches += 1;
    CUDA_TRYP(cudaGetLastError());
    CUDA_TRYP(cudaMemcpyAsync(h_pose, d_pose, sizeof(double) * 7 * batch, cudaMemcpyDeviceToHost, s));
    CUDA_TRYP(cudaMemcpyAsync(h_inl, d_inl, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
    CUDA_TRYP(cudaMemcpyAsync(h_stats, d_stats, sizeof(int) * 3 * batch, cudaMemcpyDeviceToHost, s));
    if (total) CUDA_TRYP(cudaMemcpyAsync(h_outlier, d_outlier, total, cudaMemcpyDeviceToHost, s));
    CUDA_TRYP(cudaStreamSynchronize(s));
    memcpy(pose_out, h_pose, sizeof(double) * 7 * batch);
    off = 0;
    for (int k = 0; k < batch; k++) {
      inliers_out[k] = h_inl[k];
      if (stats_out) for (int i = 0; i < 3; i++) stats_out[3 * k + i] = h_stats[3 * k + i];
      if (views[k].n) memcpy(outlier_out[k], h_outlier + off, (size_t)views[k].n);
      off += (size_t)views[k].n;
    }
    return batch;
  }
};

}  // namespace orbb200

struct orb_poseopt { orbb200::PoseOpt p; explicit orb_poseopt(int dev) : p(dev) {} };

extern "C" {

int poseopt_create(int device, orb_poseopt** out) {
  if (!out || device < 0) { orbb200::set_last_error("poseopt_create: bad argument"); return ORB_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    orbb200::set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
    return ORB_E_NODEVICE;
  }
  *out = new orb_poseopt(device);
  return ORB_OK;
}

void poseopt_destroy(orb_poseopt* h) { delete h; }

int pose_optimize(orb_poseopt* h, const pose_opt_view* v, double* pose_out, uint8_t* outlier_out) {
  if (!h || !v) return ORB_E_ARG;
  uint8_t dummy = 0;
  uint8_t* outs[1] = {outlier_out ? outlier_out : &dummy};
  if (!outlier_out && v->n > 0) { orbb200::set_last_error("pose_optimize: outlier_out is NULL"); return ORB_E_ARG; }
  int inl = 0;
  const int rc = h->p.run(1, v, pose_out, outs, &inl, nullptr);
  return rc < 0 ? rc : inl;
}

int pose_optimize_batch(orb_poseopt* h, int batch, const pose_opt_view* views, double* pose_out,
                        uint8_t* const* outlier_out, int* inliers_out, int* stats_out) {
  if (!h) return ORB_E_ARG;
  return h->p.run(batch, views, pose_out, outlier_out, inliers_out, stats_out);
}

long long poseopt_kernel_launches(const orb_poseopt* h) { return h ? h->p.launches : 0; }

float poseopt_last_ms(orb_poseopt* h) {
  if (!h || !h->p.ev1 || cudaEventSynchronize(h->p.ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->p.ev0, h->p.ev1) != cudaSuccess) return 0.f;
  return ms;
}

}  // extern "C"
