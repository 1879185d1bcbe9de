// Optimizer::LocalInertialBA's optimizer.optimize(opt_it) (reference src/Optimizer.cc:2383-2958 with the
// vertex / edge types of src/G2oTypes.cc and include/G2oTypes.h) as ONE cooperating
// thread group: the whole Levenberg-Marquardt loop -- errors, robust chi2, linearisation, Schur complement
// of the marginalised points onto the (pose, velocity, gyro bias, acc bias) side, dense LDL^T, back
// substitution, oplus, accept / reject -- runs inside one kernel launch (lia.cu), like pose_opt.cu does
// for the motion-only problem.  The window is small by construction (<= 25 keyframes x 15 unknowns), so
// one CTA owns it; a batch of windows would be a grid of CTAs.
//
// Every sum is formed in a fixed order (per-edge terms are stored and gathered by the owner of the destination;
// inertial edges are processed colour by colour), so a solve is bitwise reproducible -- no floating-point atomics.
//
// Written against a Backend (thread index, barrier, block sum, a counter add for the failure flag) so that the same
// source runs single-threaded on the host (lia_debug_host) where the CPU tests hold it against the oracle.
//
// Follows (paths relative to the reference): G2oTypes.cc:172-219 (Project / ProjectStereo / isDepthPositive /
// ImuCamPose::Update, whose NormalizeRotation(Rwb) discards its result), :349-373 and :397-427 (visual
// Jacobians), :492-594 (EdgeInertial), :774-856 (SO3 helpers); G2oTypes.h:636-700 (random-walk edges);
// ImuTypes.cc:298-332 (float preintegration getters); g2o base_multi_edge.hpp, block_solver.hpp, levenberg.cpp.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "introsort_emul.h"  // ORB_HD

namespace orbb200 {

struct LiaDev {
  // graph (device or host pointers)
  int n_kf, n_mp, n_edges, n_inertial, np, iterations;
  double lambda_init;
  const uint8_t* kf_fixed; const uint8_t* kf_has_imu;
  const int *ip, *iv, *ig, *ia;            // pose-side offsets per keyframe, -1 = fixed / absent
  double Rcb[9], tcb[3], tbc[3];
  double fx, fy, cx, cy, bf;
  const int *e_kf, *e_mp; const uint8_t* e_stereo; const double* e_obs; const float* e_is2;
  const int *lm_ptr, *lm_edges;            // CSR: edges of every map point
  // fixed-order accumulation (host-built, lia_host.h): edges of every keyframe, the (ea, eb) edge pairs of every
  // ordered pair of free poses in map-point order, the free keyframes, and a colouring of the inertial edges such
  // that edges of one colour share no keyframe
  const int *kf_ptr, *kf_edges, *pair_ptr, *pair_ea, *pair_eb, *free_kf, *i_color;
  int n_free, n_colors;
  const int *i_kf1, *i_kf2;
  const float *i_dR, *i_dV, *i_dP, *i_JRg, *i_JVg, *i_JVa, *i_JPg, *i_JPa, *i_bias, *i_dT;
  const uint8_t* i_last;
  const double *info, *infoG, *infoA;      // n_inertial x 81 / 9 / 9 (built on the host like the edge constructors)
  // state
  double *pose, *pose_bak;                 // n_kf x 24: Rwb 9, twb 3, Rcw 9, tcw 3
  double *vel, *bg, *ba, *pt, *vel_bak, *bg_bak, *ba_bak, *pt_bak;
  // system
  double *H, *b, *Hll, *bl, *W, *Dinv, *S, *bs, *x, *verr, *ierr, *Dg;
  double *Hpe, *Ye;                        // per visual edge: its 6x6 pose block + 6 gradient entries; W Dinv (6x3)
  // results
  double* chi2_out; uint8_t* depth_pos_out; double* stats;  // stats[6]: iterations, trials, err, err_end, lambda, np
  double huber_mono_delta, huber_mono_dsqr, huber_stereo_delta, huber_stereo_dsqr, huber_in_delta, huber_in_dsqr;
};

// ---------------------------------------------------------------- 3x3 helpers (row-major)
ORB_HD void l_mm(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  for (int i = 0; i < 9; i++) C[i] = T[i];
}
ORB_HD void l_tr(const double* A, double* T) {
  const double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
  for (int i = 0; i < 9; i++) T[i] = t[i];
}
ORB_HD void l_mv(const double* A, const double* v, double* o) {
  const double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2], t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2],
               t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
ORB_HD void l_skew(const double* w, double* W) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1]; W[3] = w[2]; W[4] = 0; W[5] = -w[0]; W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
ORB_HD bool l_inv3(const double* m, double* o) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (det == 0) return false;
  const double id = 1.0 / det;
  double t[9];
  t[0] = (m[4] * m[8] - m[5] * m[7]) * id; t[1] = (m[2] * m[7] - m[1] * m[8]) * id; t[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * id; t[4] = (m[0] * m[8] - m[2] * m[6]) * id; t[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * id; t[7] = (m[1] * m[6] - m[0] * m[7]) * id; t[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  for (int i = 0; i < 9; i++) o[i] = t[i];
  return true;
}
// NormalizeRotation (G2oTypes.h:68-71): U V^T of the SVD = orthogonal polar factor, by Newton iteration
ORB_HD void l_normalize(double* R) {
  for (int it = 0; it < 20; it++) {
    double Ri[9], N[9];
    if (!l_inv3(R, Ri)) return;
    double diff = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { N[i * 3 + j] = 0.5 * (R[i * 3 + j] + Ri[j * 3 + i]); diff = fmax(diff, fabs(N[i * 3 + j] - R[i * 3 + j])); }
    for (int i = 0; i < 9; i++) R[i] = N[i];
    if (diff < 1e-15) break;
  }
}
ORB_HD void l_exp(const double* w, double* R) {  // ExpSO3, G2oTypes.cc:780-798
  const double d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], d = sqrt(d2);
  double W[9], W2[9];
  l_skew(w, W);
  l_mm(W, W, W2);
  const double a = d < 1e-5 ? 1.0 : sin(d) / d, c = d < 1e-5 ? 0.5 : (1.0 - cos(d)) / d2;
  for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i] * a + W2[i] * c;
  l_normalize(R);
}
ORB_HD void l_log(const double* R, double* w) {  // LogSO3, :800-814
  const double t = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (t - 1.0) * 0.5f;
  if (costheta > 1 || costheta < -1) return;
  const double theta = acos(costheta), s = sin(theta);
  if (fabs(s) < 1e-5) return;
  for (int i = 0; i < 3; i++) w[i] = theta * w[i] / s;
}
ORB_HD void l_jr(const double* v, double* J, bool inverse) {  // RightJacobianSO3 / InverseRightJacobianSO3, :816-849
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  l_skew(v, W);
  l_mm(W, W, W2);
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    if (d < 1e-5) J[i] = I;
    else if (inverse) J[i] = I + W[i] / 2 + W2[i] * (1.0 / d2 - (1.0 + cos(d)) / (2.0 * d * sin(d)));
    else J[i] = I - W[i] * (1.0 - cos(d)) / d2 + W2[i] * (d - sin(d)) / (d2 * d);
  }
}
ORB_HD void l_huber(double delta, double dsqr, bool on, double e, double& r0, double& r1) {
  if (!on || e <= dsqr) { r0 = e; r1 = 1.; }
  else { const double s = sqrt(e); r0 = 2 * s * delta - dsqr; r1 = delta / s; }
}
ORB_HD double l_quad(const double* O, const double* r, int n) {
  double c = 0;
  for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += O[i * n + j] * r[j]; c += r[i] * s; }
  return c;
}

// IMU::Preintegrated::GetDeltaRotation / Velocity / Position for the bias estimate of keyframe k1
// (ImuTypes.cc:298-332): float arithmetic, results handed over as float
ORB_HD void l_preint(const LiaDev& D, int i, int k1, double* dR, double* dV, double* dP, double* dbg_out) {
  const float* b0 = D.i_bias + 6 * (size_t)i;
  float dbg[3], dba[3];
  for (int c = 0; c < 3; c++) { dbg[c] = (float)D.bg[3 * k1 + c] - b0[3 + c]; dba[c] = (float)D.ba[3 * k1 + c] - b0[c]; }
  const float *JRg = D.i_JRg + 9 * (size_t)i, *JVg = D.i_JVg + 9 * (size_t)i, *JVa = D.i_JVa + 9 * (size_t)i;
  const float *JPg = D.i_JPg + 9 * (size_t)i, *JPa = D.i_JPa + 9 * (size_t)i;
  double wd[3], E[9], R0[9];
  for (int r = 0; r < 3; r++) wd[r] = (double)(JRg[r * 3] * dbg[0] + JRg[r * 3 + 1] * dbg[1] + JRg[r * 3 + 2] * dbg[2]);
  l_exp(wd, E);
  for (int c = 0; c < 9; c++) R0[c] = D.i_dR[9 * (size_t)i + c];
  l_mm(R0, E, dR);
  l_normalize(dR);
  for (int c = 0; c < 9; c++) dR[c] = (double)(float)dR[c];
  for (int r = 0; r < 3; r++) {
    const float v = D.i_dV[3 * (size_t)i + r] + (JVg[r * 3] * dbg[0] + JVg[r * 3 + 1] * dbg[1] + JVg[r * 3 + 2] * dbg[2]) +
                    (JVa[r * 3] * dba[0] + JVa[r * 3 + 1] * dba[1] + JVa[r * 3 + 2] * dba[2]);
    const float p = D.i_dP[3 * (size_t)i + r] + (JPg[r * 3] * dbg[0] + JPg[r * 3 + 1] * dbg[1] + JPg[r * 3 + 2] * dbg[2]) +
                    (JPa[r * 3] * dba[0] + JPa[r * 3 + 1] * dba[1] + JPa[r * 3 + 2] * dba[2]);
    dV[r] = v; dP[r] = p;
  }
  if (dbg_out) for (int c = 0; c < 3; c++) dbg_out[c] = dbg[c];
}

ORB_HD double l_vchi2(const LiaDev& D, int e) {
  const double s = D.e_is2[e];
  const double* r = D.verr + 3 * (size_t)e;
  return r[0] * (s * r[0]) + r[1] * (s * r[1]) + (D.e_stereo[e] ? r[2] * (s * r[2]) : 0.0);
}

// computeError of one visual edge (EdgeMono / EdgeStereo)
ORB_HD void l_visual_error(const LiaDev& D, int e) {
  const double* P = D.pose + 24 * (size_t)D.e_kf[e];
  const double* X = D.pt + 3 * (size_t)D.e_mp[e];
  double Xc[3];
  l_mv(P + 12, X, Xc);
  for (int c = 0; c < 3; c++) Xc[c] += P[21 + c];
  const double u = D.fx * Xc[0] / Xc[2] + D.cx, v = D.fy * Xc[1] / Xc[2] + D.cy;
  double* r = D.verr + 3 * (size_t)e;
  const double* o = D.e_obs + 3 * (size_t)e;
  r[0] = o[0] - u; r[1] = o[1] - v;
  r[2] = D.e_stereo[e] ? o[2] - (u - D.bf * (1 / Xc[2])) : 0.0;
}

// computeError of EdgeInertial + EdgeGyroRW + EdgeAccRW i
ORB_HD void l_inertial_error(const LiaDev& D, int i) {
  const int k1 = D.i_kf1[i], k2 = D.i_kf2[i];
  const double *P1 = D.pose + 24 * (size_t)k1, *P2 = D.pose + 24 * (size_t)k2;
  double dR[9], dV[3], dP[3];
  l_preint(D, i, k1, dR, dV, dP, nullptr);
  const double dt = D.i_dT[i];
  const double grav[3] = {0, 0, -(double)9.81f};
  double Rbw1[9], dRt[9], T[9], eR[9];
  l_tr(P1, Rbw1); l_tr(dR, dRt);
  l_mm(dRt, Rbw1, T); l_mm(T, P2, eR);
  double* r = D.ierr + 15 * (size_t)i;
  l_log(eR, r);
  double a[3], c[3];
  for (int q = 0; q < 3; q++) {
    a[q] = D.vel[3 * k2 + q] - D.vel[3 * k1 + q] - grav[q] * dt;
    c[q] = P2[9 + q] - P1[9 + q] - D.vel[3 * k1 + q] * dt - grav[q] * dt * dt / 2;
  }
  l_mv(Rbw1, a, a); l_mv(Rbw1, c, c);
  for (int q = 0; q < 3; q++) { r[3 + q] = a[q] - dV[q]; r[6 + q] = c[q] - dP[q]; }
  for (int q = 0; q < 3; q++) { r[9 + q] = D.bg[3 * k2 + q] - D.bg[3 * k1 + q]; r[12 + q] = D.ba[3 * k2 + q] - D.ba[3 * k1 + q]; }
}

// H += w Ja^T O Jb (and its transpose block), b -= w Ja^T O r.  Plain read-modify-write: the caller guarantees that
// no other thread touches the blocks of these vertices (inertial edges are processed one colour at a time)
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_add_block(BE& be, const LiaDev& D, int oa, int da, const double* Ja, int ob, int db_, const double* Jb,
                        const double* O, int d, double w) {
  for (int i = 0; i < da; i++)
    for (int j = 0; j < db_; j++) {
      double s = 0;
      for (int p = 0; p < d; p++) {
        double t = 0;
        for (int q = 0; q < d; q++) t += O[p * d + q] * Jb[q * db_ + j];
        s += Ja[p * da + i] * t;
      }
      D.H[(size_t)(oa + i) * D.np + ob + j] += w * s;
      if (oa != ob) D.H[(size_t)(ob + j) * D.np + oa + i] += w * s;
    }
}
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_add_b(BE& be, const LiaDev& D, int oa, int da, const double* Ja, const double* O, const double* r, int d, double w) {
  for (int i = 0; i < da; i++) {
    double s = 0;
    for (int p = 0; p < d; p++) { double t = 0; for (int q = 0; q < d; q++) t += O[p * d + q] * r[q]; s += Ja[p * da + i] * t; }
    D.b[oa + i] += -w * s;
  }
}

// linearizeOplus + constructQuadraticForm of the edges of map point l (one thread owns the point)
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_build_point(BE& be, const LiaDev& D, int l) {
  double Hl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, blv[3] = {0, 0, 0}, Rbc[9];
  l_tr(D.Rcb, Rbc);
  for (int q = D.lm_ptr[l]; q < D.lm_ptr[l + 1]; q++) {
    const int e = D.lm_edges[q], k = D.e_kf[e], d = D.e_stereo[e] ? 3 : 2;
    const double* P = D.pose + 24 * (size_t)k;
    double Xc[3], Xb[3];
    l_mv(P + 12, D.pt + 3 * (size_t)l, Xc);
    for (int c = 0; c < 3; c++) Xc[c] += P[21 + c];
    l_mv(Rbc, Xc, Xb);
    for (int c = 0; c < 3; c++) Xb[c] += D.tbc[c];
    const double x = Xc[0], y = Xc[1], z = Xc[2];
    double pj[9] = {D.fx / z, 0, -D.fx * x / (z * z), 0, D.fy / z, -D.fy * y / (z * z), 0, 0, 0};
    if (d == 3) { pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + D.bf * (1.0 / (z * z)); }
    double A[9], B[18], PR[9];
    l_mm(pj, P + 12, A);
    for (int c = 0; c < 9; c++) A[c] = -A[c];
    l_mm(pj, D.Rcb, PR);
    const double Dv[18] = {0, Xb[2], -Xb[1], 1, 0, 0, -Xb[2], 0, Xb[0], 0, 1, 0, Xb[1], -Xb[0], 0, 0, 0, 1};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 6; c++) B[r * 6 + c] = PR[r * 3] * Dv[c] + PR[r * 3 + 1] * Dv[6 + c] + PR[r * 3 + 2] * Dv[12 + c];
    double r0, r1;
    if (d == 3) l_huber(D.huber_stereo_delta, D.huber_stereo_dsqr, true, l_vchi2(D, e), r0, r1);
    else l_huber(D.huber_mono_delta, D.huber_mono_dsqr, true, l_vchi2(D, e), r0, r1);
    const double s = D.e_is2[e], ws = r1 * s;
    const double* r = D.verr + 3 * (size_t)e;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) { double a = 0; for (int p = 0; p < d; p++) a += A[p * 3 + i] * ws * A[p * 3 + j]; Hl[i * 3 + j] += a; }
      double a = 0;
      for (int p = 0; p < d; p++) a += A[p * 3 + i] * (s * r[p]);
      blv[i] -= r1 * a;
    }
    double* We = D.W + 18 * (size_t)e;
    const int o = D.ip[k];
    if (o >= 0) {
      double* He = D.Hpe + 42 * (size_t)e;  // gathered per keyframe by l_gather_poses
      for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) { double a = 0; for (int p = 0; p < d; p++) a += B[p * 6 + i] * ws * B[p * 6 + j]; He[i * 6 + j] = a; }
        double a = 0;
        for (int p = 0; p < d; p++) a += B[p * 6 + i] * (s * r[p]);
        He[36 + i] = -r1 * a;
        for (int j = 0; j < 3; j++) { double a2 = 0; for (int p = 0; p < d; p++) a2 += B[p * 6 + i] * ws * A[p * 3 + j]; We[i * 3 + j] = a2; }
      }
    } else {
      for (int i = 0; i < 18; i++) We[i] = 0;
    }
  }
  for (int i = 0; i < 9; i++) D.Hll[9 * (size_t)l + i] = Hl[i];
  for (int i = 0; i < 3; i++) D.bl[3 * (size_t)l + i] = blv[i];
}

// EdgeInertial::linearizeOplus (:534-594) + BaseMultiEdge quadratic form, and the two random-walk edges
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_build_inertial(BE& be, const LiaDev& D, int i) {
  const int k1 = D.i_kf1[i], k2 = D.i_kf2[i];
  const double *P1 = D.pose + 24 * (size_t)k1, *P2 = D.pose + 24 * (size_t)k2;
  double dR[9], dV[3], dP[3], dbg[3];
  l_preint(D, i, k1, dR, dV, dP, dbg);
  const double dt = D.i_dT[i];
  const double grav[3] = {0, 0, -(double)9.81f};
  double Rbw1[9], dRt[9], T[9], eR[9], er[3], invJr[9];
  l_tr(P1, Rbw1); l_tr(dR, dRt);
  l_mm(dRt, Rbw1, T); l_mm(T, P2, eR);
  l_log(eR, er);
  l_jr(er, invJr, true);
  double JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
  for (int c = 0; c < 9; c++) {
    JRg[c] = D.i_JRg[9 * (size_t)i + c]; JVg[c] = D.i_JVg[9 * (size_t)i + c]; JVa[c] = D.i_JVa[9 * (size_t)i + c];
    JPg[c] = D.i_JPg[9 * (size_t)i + c]; JPa[c] = D.i_JPa[9 * (size_t)i + c];
  }
  double J0[54], J1[27], J2[27], J3[27], J4[54], J5[27];
  for (int c = 0; c < 54; c++) { J0[c] = 0; J4[c] = 0; }
  for (int c = 0; c < 27; c++) { J1[c] = 0; J2[c] = 0; J3[c] = 0; J5[c] = 0; }
  double Rwb2t[9], M[9], a[3], c3[3], Sk[9];
  l_tr(P2, Rwb2t);
  l_mm(Rwb2t, P1, M); l_mm(invJr, M, M);
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[r * 6 + q] = -M[r * 3 + q];
  for (int q = 0; q < 3; q++) {
    a[q] = D.vel[3 * k2 + q] - D.vel[3 * k1 + q] - grav[q] * dt;
    c3[q] = P2[9 + q] - P1[9 + q] - D.vel[3 * k1 + q] * dt - 0.5 * grav[q] * dt * dt;
  }
  l_mv(Rbw1, a, a); l_mv(Rbw1, c3, c3);
  l_skew(a, Sk);
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[(3 + r) * 6 + q] = Sk[r * 3 + q];
  l_skew(c3, Sk);
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J0[(6 + r) * 6 + q] = Sk[r * 3 + q];
  for (int r = 0; r < 3; r++) J0[(6 + r) * 6 + 3 + r] = -1.0;
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J1[(3 + r) * 3 + q] = -Rbw1[r * 3 + q]; J1[(6 + r) * 3 + q] = -Rbw1[r * 3 + q] * dt; }
  {
    double v[3], Jr[9], eRt[9], X[9];
    l_mv(JRg, dbg, v);
    l_jr(v, Jr, false);
    l_tr(eR, eRt);
    l_mm(invJr, eRt, X); l_mm(X, Jr, X); l_mm(X, JRg, X);
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J2[r * 3 + q] = -X[r * 3 + q]; J2[(3 + r) * 3 + q] = -JVg[r * 3 + q]; J2[(6 + r) * 3 + q] = -JPg[r * 3 + q]; }
  }
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) { J3[(3 + r) * 3 + q] = -JVa[r * 3 + q]; J3[(6 + r) * 3 + q] = -JPa[r * 3 + q]; }
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J4[r * 6 + q] = invJr[r * 3 + q];
  l_mm(Rbw1, P2, M);
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J4[(6 + r) * 6 + 3 + q] = M[r * 3 + q];
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) J5[(3 + r) * 3 + q] = Rbw1[r * 3 + q];
  const double* r9 = D.ierr + 15 * (size_t)i;
  const double* O = D.info + 81 * (size_t)i;
  double w = 1.0, r0;
  if (D.i_last[i]) l_huber(D.huber_in_delta, D.huber_in_dsqr, true, l_quad(O, r9, 9), r0, w);
  const int off[6] = {D.ip[k1], D.iv[k1], D.ig[k1], D.ia[k1], D.ip[k2], D.iv[k2]};
  const int dim[6] = {6, 3, 3, 3, 6, 3};
  const double* J[6] = {J0, J1, J2, J3, J4, J5};
  for (int p = 0; p < 6; p++) {
    if (off[p] < 0) continue;
    l_add_b(be, D, off[p], dim[p], J[p], O, r9, 9, w);
    for (int q = p; q < 6; q++)
      if (off[q] >= 0) l_add_block(be, D, off[p], dim[p], J[p], off[q], dim[q], J[q], O, 9, w);
  }
  const double nI[9] = {-1, 0, 0, 0, -1, 0, 0, 0, -1}, pI[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const int og[2] = {D.ig[k1], D.ig[k2]}, oa[2] = {D.ia[k1], D.ia[k2]};
  const double* JJ[2] = {nI, pI};
  for (int p = 0; p < 2; p++) {
    if (og[p] >= 0) {
      l_add_b(be, D, og[p], 3, JJ[p], D.infoG + 9 * (size_t)i, r9 + 9, 3, 1.0);
      for (int q = p; q < 2; q++) if (og[q] >= 0) l_add_block(be, D, og[p], 3, JJ[p], og[q], 3, JJ[q], D.infoG + 9 * (size_t)i, 3, 1.0);
    }
    if (oa[p] >= 0) {
      l_add_b(be, D, oa[p], 3, JJ[p], D.infoA + 9 * (size_t)i, r9 + 12, 3, 1.0);
      for (int q = p; q < 2; q++) if (oa[q] >= 0) l_add_block(be, D, oa[p], 3, JJ[p], oa[q], 3, JJ[q], D.infoA + 9 * (size_t)i, 3, 1.0);
    }
  }
}

// Pose blocks of the visual edges: entry `ent` (36 of H, 6 of b) of free keyframe f = the sum over the keyframe's
// edges in input order.  One owner per destination, so H / b start from these sums (they were zeroed).
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_gather_poses(BE& be, const LiaDev& D) {
  const int tid = be.tid(), nt = be.nthreads();
  for (int item = tid; item < D.n_free * 42; item += nt) {
    const int f = item / 42, ent = item - f * 42, k = D.free_kf[f], o = D.ip[k];
    double sum = 0;
    for (int q = D.kf_ptr[k]; q < D.kf_ptr[k + 1]; q++) sum += D.Hpe[42 * (size_t)D.kf_edges[q] + ent];
    if (ent < 36) D.H[(size_t)(o + ent / 6) * D.np + o + ent % 6] = sum;
    else D.b[o + ent - 36] = sum;
  }
}

// Schur complement of the map points onto the pose side, fixed order:
//   (1) per map point: Dinv = (Hll + lambda I)^-1 and Y_e = W_e Dinv for its edges;
//   (2) per (free keyframe, row): bs -= sum over the keyframe's edges of Y_e b_l;
//       per (ordered pair of free keyframes, i, j): S -= sum over the pair's edge pairs (map-point order) of Y_ea W_eb^T.
// S and bs hold Hpp + lambda I and b on entry.  Barriers inside.
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_schur(BE& be, const LiaDev& D, double lambda) {
  const int tid = be.tid(), nt = be.nthreads(), np = D.np;
  for (int l = tid; l < D.n_mp; l += nt) {
    double Hl[9], Di[9];
    for (int c = 0; c < 9; c++) Hl[c] = D.Hll[9 * (size_t)l + c];
    Hl[0] += lambda; Hl[4] += lambda; Hl[8] += lambda;
    const bool ok = l_inv3(Hl, Di);
    if (!ok) {
      be.count(&D.stats[6]);
      for (int c = 0; c < 9; c++) Di[c] = 0;
    }
    for (int c = 0; c < 9; c++) D.Dinv[9 * (size_t)l + c] = Di[c];
    for (int qa = D.lm_ptr[l]; qa < D.lm_ptr[l + 1]; qa++) {
      const int ea = D.lm_edges[qa];
      if (D.ip[D.e_kf[ea]] < 0) continue;
      const double* Wa = D.W + 18 * (size_t)ea;
      double* Y = D.Ye + 18 * (size_t)ea;
      for (int i = 0; i < 6; i++) for (int j = 0; j < 3; j++) Y[i * 3 + j] = Wa[i * 3] * Di[j] + Wa[i * 3 + 1] * Di[3 + j] + Wa[i * 3 + 2] * Di[6 + j];
    }
  }
  be.sync();
  for (int item = tid; item < D.n_free * 6; item += nt) {
    const int f = item / 6, i = item - f * 6, k = D.free_kf[f];
    double sum = 0;
    for (int q = D.kf_ptr[k]; q < D.kf_ptr[k + 1]; q++) {
      const int e = D.kf_edges[q];
      const double* Y = D.Ye + 18 * (size_t)e + 3 * i;
      const double* bll = D.bl + 3 * (size_t)D.e_mp[e];
      sum += Y[0] * bll[0] + Y[1] * bll[1] + Y[2] * bll[2];
    }
    D.bs[D.ip[k] + i] -= sum;
  }
  for (int item = tid; item < D.n_free * D.n_free * 36; item += nt) {
    const int key = item / 36, ij = item - key * 36, i = ij / 6, j = ij - i * 6;
    const int q0 = D.pair_ptr[key], q1 = D.pair_ptr[key + 1];
    if (q0 == q1) continue;
    double sum = 0;
    for (int q = q0; q < q1; q++) {
      const double* Y = D.Ye + 18 * (size_t)D.pair_ea[q] + 3 * i;
      const double* Wb = D.W + 18 * (size_t)D.pair_eb[q] + 3 * j;
      sum += Y[0] * Wb[0] + Y[1] * Wb[1] + Y[2] * Wb[2];
    }
    const int oa = D.ip[D.free_kf[key / D.n_free]], ob = D.ip[D.free_kf[key % D.n_free]];
    D.S[(size_t)(oa + i) * np + ob + j] -= sum;
  }
  be.sync();
}

// errors of every edge, then the robust chi2 (block sum)
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD double l_errors_and_chi2(BE& be, const LiaDev& D) {
  const int tid = be.tid(), nt = be.nthreads();
  for (int e = tid; e < D.n_edges; e += nt) l_visual_error(D, e);
  for (int i = tid; i < D.n_inertial; i += nt) l_inertial_error(D, i);
  be.sync();
  double part = 0, r0, r1;
  for (int e = tid; e < D.n_edges; e += nt) {
    if (D.e_stereo[e]) l_huber(D.huber_stereo_delta, D.huber_stereo_dsqr, true, l_vchi2(D, e), r0, r1);
    else l_huber(D.huber_mono_delta, D.huber_mono_dsqr, true, l_vchi2(D, e), r0, r1);
    part += r0;
  }
  for (int i = tid; i < D.n_inertial; i += nt) {
    const double* r = D.ierr + 15 * (size_t)i;
    const double ci = l_quad(D.info + 81 * (size_t)i, r, 9);
    l_huber(D.huber_in_delta, D.huber_in_dsqr, D.i_last[i] != 0, ci, r0, r1);
    part += r0 + l_quad(D.infoG + 9 * (size_t)i, r + 9, 3) + l_quad(D.infoA + 9 * (size_t)i, r + 12, 3);
  }
  return be.sum(part);
}

// oplus of every vertex with the update x (pose side, then points)
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void l_update(BE& be, const LiaDev& D) {
  const int tid = be.tid(), nt = be.nthreads();
  for (int k = tid; k < D.n_kf; k += nt) {
    if (D.ip[k] >= 0) {  // ImuCamPose::Update (:190-218)
      double* P = D.pose + 24 * (size_t)k;
      const double* u = D.x + D.ip[k];
      double d[3], E[9], Rbw[9], tbw[3];
      l_mv(P, u + 3, d);
      for (int c = 0; c < 3; c++) P[9 + c] += d[c];
      l_exp(u, E);
      l_mm(P, E, P);
      l_tr(P, Rbw);
      l_mv(Rbw, P + 9, tbw);
      for (int c = 0; c < 3; c++) tbw[c] = -tbw[c];
      l_mm(D.Rcb, Rbw, P + 12);
      l_mv(D.Rcb, tbw, P + 21);
      for (int c = 0; c < 3; c++) P[21 + c] += D.tcb[c];
    }
    if (D.iv[k] >= 0)
      for (int c = 0; c < 3; c++) { D.vel[3 * k + c] += D.x[D.iv[k] + c]; D.bg[3 * k + c] += D.x[D.ig[k] + c]; D.ba[3 * k + c] += D.x[D.ia[k] + c]; }
  }
  for (int i = tid; i < 3 * D.n_mp; i += nt) D.pt[i] += D.x[(size_t)D.np + i];
  be.sync();
}

// The whole optimize(opt_it).
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
ORB_HD void lia_solve_core(BE& be, const LiaDev& D) {
  const int tid = be.tid(), nt = be.nthreads(), np = D.np;
  double lambda = D.lambda_init, ni = 2, chi_first = 0, currentChi = 0;
  int nBad = 0, trials = 0, iters = 0;
  for (int it = 0; it < D.iterations; it++) {
    currentChi = l_errors_and_chi2(be, D);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) { chi_first = currentChi; lambda = D.lambda_init; ni = 2; nBad = 0; }
    // buildSystem
    for (size_t i = tid; i < (size_t)np * np; i += nt) D.H[i] = 0;
    for (int i = tid; i < np; i += nt) D.b[i] = 0;
    be.sync();
    for (int l = tid; l < D.n_mp; l += nt) l_build_point(be, D, l);
    be.sync();
    l_gather_poses(be, D);
    be.sync();
    for (int c = 0; c < D.n_colors; c++) {  // edges of one colour touch disjoint vertex blocks
      for (int i = tid; i < D.n_inertial; i += nt)
        if (D.i_color[i] == c) l_build_inertial(be, D, i);
      be.sync();
    }
    double rho = 0;
    int qmax = 0;
    do {
      // push
      for (int i = tid; i < 24 * D.n_kf; i += nt) D.pose_bak[i] = D.pose[i];
      for (int i = tid; i < 3 * D.n_kf; i += nt) { D.vel_bak[i] = D.vel[i]; D.bg_bak[i] = D.bg[i]; D.ba_bak[i] = D.ba[i]; }
      for (int i = tid; i < 3 * D.n_mp; i += nt) D.pt_bak[i] = D.pt[i];
      // S = Hpp + lambda I, bs = b; then the Schur terms of every map point
      for (size_t i = tid; i < (size_t)np * np; i += nt) D.S[i] = D.H[i] + ((i / np == i % np) ? lambda : 0.0);
      for (int i = tid; i < np; i += nt) D.bs[i] = D.b[i];
      if (tid == 0) D.stats[6] = 0;  // failure flag
      be.sync();
      l_schur(be, D, lambda);
      // dense LDL^T of S (lower triangle in place, D in Dg), right-looking, columns in order
      bool ok2 = D.stats[6] == 0;
      be.sync();
      for (int j = 0; j < np && ok2; j++) {
        const double d = D.S[(size_t)j * np + j];
        if (!(d > 0)) { ok2 = false; break; }
        be.sync();
        if (tid == 0) D.Dg[j] = d;
        for (int i = j + 1 + tid; i < np; i += nt) D.S[(size_t)i * np + j] /= d;
        be.sync();
        // trailing update: A[i][k] -= L[i][j] d L[k][j], k in (j, i]
        const int m = np - j - 1;
        for (long long t = tid; t < (long long)m * m; t += nt) {
          const int i = j + 1 + (int)(t / m), k = j + 1 + (int)(t % m);
          if (k <= i) D.S[(size_t)i * np + k] -= D.S[(size_t)i * np + j] * d * D.S[(size_t)k * np + j];
        }
        be.sync();
      }
      if (ok2) {
        if (tid == 0) {  // triangular solves: the system is at most a few hundred unknowns
          for (int i = 0; i < np; i++) { double v = D.bs[i]; for (int k = 0; k < i; k++) v -= D.S[(size_t)i * np + k] * D.x[k]; D.x[i] = v; }
          for (int i = 0; i < np; i++) D.x[i] /= D.Dg[i];
          for (int i = np - 1; i >= 0; i--) { double v = D.x[i]; for (int k = i + 1; k < np; k++) v -= D.S[(size_t)k * np + i] * D.x[k]; D.x[i] = v; }
        }
        be.sync();
        for (int l = tid; l < D.n_mp; l += nt) {  // x_l = Dinv (b_l - W^T x_p)
          double c[3] = {D.bl[3 * (size_t)l], D.bl[3 * (size_t)l + 1], D.bl[3 * (size_t)l + 2]};
          for (int q = D.lm_ptr[l]; q < D.lm_ptr[l + 1]; q++) {
            const int e = D.lm_edges[q], o = D.ip[D.e_kf[e]];
            if (o < 0) continue;
            const double* We = D.W + 18 * (size_t)e;
            for (int j = 0; j < 3; j++) for (int i = 0; i < 6; i++) c[j] -= We[i * 3 + j] * D.x[o + i];
          }
          const double* Di = D.Dinv + 9 * (size_t)l;
          for (int i = 0; i < 3; i++) D.x[(size_t)np + 3 * (size_t)l + i] = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
        }
      }
      be.sync();
      l_update(be, D);  // g2o updates even after a failed solve (x keeps its previous content)
      tempChi = l_errors_and_chi2(be, D);
      if (!ok2) tempChi = DBL_MAX;
      double part = 0;
      for (int j = tid; j < np; j += nt) part += D.x[j] * (lambda * D.x[j] + D.b[j]);
      for (int j = tid; j < 3 * D.n_mp; j += nt) part += D.x[(size_t)np + j] * (lambda * D.x[(size_t)np + j] + D.bl[j]);
      const double scale = be.sum(part) + 1e-3;
      rho = (currentChi - tempChi) / scale;
      if (rho > 0 && tempChi < DBL_MAX && tempChi == tempChi) {
        double alpha = 1. - pow((2 * rho - 1), 3.0);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2;
        for (int i = tid; i < 24 * D.n_kf; i += nt) D.pose[i] = D.pose_bak[i];  // pop
        for (int i = tid; i < 3 * D.n_kf; i += nt) { D.vel[i] = D.vel_bak[i]; D.bg[i] = D.bg_bak[i]; D.ba[i] = D.ba_bak[i]; }
        for (int i = tid; i < 3 * D.n_mp; i += nt) D.pt[i] = D.pt_bak[i];
        be.sync();
      }
      qmax++; trials++;
    } while (rho < 0 && qmax < 10);
    iters++;
    if (qmax == 10 || rho == 0) break;
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) break;
  }
  // e->chi2() of the last evaluated trial, isDepthPositive at the final estimate
  for (int e = tid; e < D.n_edges; e += nt) {
    D.chi2_out[e] = l_vchi2(D, e);
    const double* P = D.pose + 24 * (size_t)D.e_kf[e];
    const double* X = D.pt + 3 * (size_t)D.e_mp[e];
    D.depth_pos_out[e] = (P[18] * X[0] + P[19] * X[1] + P[20] * X[2] + P[23]) > 0.0;
  }
  if (tid == 0) { D.stats[0] = iters; D.stats[1] = trials; D.stats[2] = chi_first; D.stats[3] = currentChi; D.stats[4] = lambda; D.stats[5] = np; }
  be.sync();
}

struct LiaHostBackend {
  int tid() const { return 0; }
  int nthreads() const { return 1; }
  void sync() {}
  void count(double* p) { *p += 1.0; }
  double sum(double v) { return v; }
};

}  // namespace orbb200
