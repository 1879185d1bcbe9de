// Host-side preparation of a LocalInertialBA graph for csrc/lia_core.h: what the reference's vertex and edge
// constructors do once per call (vertex indexing, the map-point CSR, the information matrices of the inertial
// edges -- EdgeInertial constructor G2oTypes.cc:500-508, InfoG / InfoA Optimizer.cc:2601-2609).  Plain C++ so
// that lia.cu (nvcc) and the ThreadSanitizer harness of the CPU tests (g++, tests/native/lia_threads.cpp) share it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "lia_core.h"

namespace orbb200 {

inline bool invert_dense(std::vector<double> A, int n, std::vector<double>& Ai) {
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

// symmetric eigen-decomposition by cyclic Jacobi rotations (A is overwritten by the eigenvalues on its diagonal)
inline void jacobi_eigen(std::vector<double>& A, int n, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) { const double x = A[k * n + p], y = A[k * n + q]; A[k * n + p] = c * x - s * y; A[k * n + q] = s * x + c * y; }
        for (int k = 0; k < n; k++) { const double x = A[p * n + k], y = A[q * n + k]; A[p * n + k] = c * x - s * y; A[q * n + k] = s * x + c * y; }
        for (int k = 0; k < n; k++) { const double x = V[k * n + p], y = V[k * n + q]; V[k * n + p] = c * x - s * y; V[k * n + q] = s * x + c * y; }
      }
  }
}

struct LiaHost {
  std::vector<int> ip, iv, ig, ia, lm_ptr, lm_edges;
  std::vector<int> kf_ptr, kf_edges, pair_ptr, pair_ea, pair_eb, free_kf, i_color;  // fixed-order accumulation (lia_core.h)
  int n_colors = 0;
  std::vector<double> info, infoG, infoA, pose, e_obs;
  int np = 0;
};

inline int lia_check(const lia_graph_view* g, std::string& err) {
  if (!g || g->n_kf <= 0 || g->n_mp < 0 || g->n_edges < 0 || g->n_inertial < 0 || g->iterations < 0 || !(g->lambda_init > 0) ||
      !g->kf_Rwb || !g->kf_twb || !g->kf_Rcw || !g->kf_tcw || !g->kf_fixed || !g->kf_has_imu || !g->kf_vel || !g->kf_bg ||
      !g->kf_ba || (g->n_mp && !g->mp_pos) ||
      (g->n_edges && (!g->e_kf || !g->e_mp || !g->e_stereo || !g->e_obs || !g->e_inv_sigma2)) ||
      (g->n_inertial && (!g->i_kf1 || !g->i_kf2 || !g->i_dR || !g->i_dV || !g->i_dP || !g->i_JRg || !g->i_JVg || !g->i_JVa ||
                         !g->i_JPg || !g->i_JPa || !g->i_bias || !g->i_dT || !g->i_C || !g->i_last))) {
    err = "lia_solve: bad graph view";
    return ORB_E_ARG;
  }
  for (int e = 0; e < g->n_edges; e++)
    if (g->e_kf[e] < 0 || g->e_kf[e] >= g->n_kf || g->e_mp[e] < 0 || g->e_mp[e] >= g->n_mp) { err = "lia_solve: edge index out of range"; return ORB_E_ARG; }
  for (int i = 0; i < g->n_inertial; i++)
    if (g->i_kf1[i] < 0 || g->i_kf1[i] >= g->n_kf || g->i_kf2[i] < 0 || g->i_kf2[i] >= g->n_kf || !g->kf_has_imu[g->i_kf1[i]] ||
        !g->kf_has_imu[g->i_kf2[i]]) { err = "lia_solve: inertial edge between keyframes without IMU vertices"; return ORB_E_ARG; }
  return 0;
}

// Vertex indexing, the map-point CSR and the information matrices of the inertial edges
// (EdgeInertial constructor G2oTypes.cc:500-508; InfoG / InfoA Optimizer.cc:2601-2609)
inline int lia_prepare(const lia_graph_view* g, LiaHost& Hs, std::string& err) {
  const int K = g->n_kf;
  Hs.ip.assign(K, -1); Hs.iv.assign(K, -1); Hs.ig.assign(K, -1); Hs.ia.assign(K, -1);
  Hs.np = 0;
  Hs.pose.resize(24 * (size_t)K);
  for (int k = 0; k < K; k++) {
    double* P = &Hs.pose[24 * (size_t)k];
    memcpy(P, g->kf_Rwb + 9 * k, 72); memcpy(P + 9, g->kf_twb + 3 * k, 24);
    memcpy(P + 12, g->kf_Rcw + 9 * k, 72); memcpy(P + 21, g->kf_tcw + 3 * k, 24);
    if (g->kf_fixed[k]) continue;
    Hs.ip[k] = Hs.np; Hs.np += 6;
    if (g->kf_has_imu[k]) { Hs.iv[k] = Hs.np; Hs.np += 3; Hs.ig[k] = Hs.np; Hs.np += 3; Hs.ia[k] = Hs.np; Hs.np += 3; }
  }
  if (Hs.np == 0) { err = "lia_solve: no free vertex"; return ORB_E_ARG; }
  Hs.lm_ptr.assign(g->n_mp + 1, 0);
  for (int e = 0; e < g->n_edges; e++) Hs.lm_ptr[g->e_mp[e] + 1]++;
  for (int l = 0; l < g->n_mp; l++) Hs.lm_ptr[l + 1] += Hs.lm_ptr[l];
  Hs.lm_edges.resize(std::max(g->n_edges, 1));
  {
    std::vector<int> cur(Hs.lm_ptr.begin(), Hs.lm_ptr.end() - 1);
    for (int e = 0; e < g->n_edges; e++) Hs.lm_edges[cur[g->e_mp[e]]++] = e;
  }
  // ---- structures of the fixed-order sums: edges per keyframe (input order), free keyframes, edge pairs per ordered
  //      pair of free keyframes (map points ascending, the pairs of one point in list order), inertial edge colours
  Hs.kf_ptr.assign(K + 1, 0);
  for (int e = 0; e < g->n_edges; e++) Hs.kf_ptr[g->e_kf[e] + 1]++;
  for (int k = 0; k < K; k++) Hs.kf_ptr[k + 1] += Hs.kf_ptr[k];
  Hs.kf_edges.resize(std::max(g->n_edges, 1));
  {
    std::vector<int> cur(Hs.kf_ptr.begin(), Hs.kf_ptr.end() - 1);
    for (int e = 0; e < g->n_edges; e++) Hs.kf_edges[cur[g->e_kf[e]]++] = e;
  }
  std::vector<int> frank(K, -1);
  Hs.free_kf.clear();
  for (int k = 0; k < K; k++)
    if (Hs.ip[k] >= 0) { frank[k] = (int)Hs.free_kf.size(); Hs.free_kf.push_back(k); }
  const int F = (int)Hs.free_kf.size();
  Hs.pair_ptr.assign((size_t)F * F + 1, 0);
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int> cur;
    if (pass == 1) {
      for (size_t i = 0; i < (size_t)F * F; i++) Hs.pair_ptr[i + 1] += Hs.pair_ptr[i];
      Hs.pair_ea.resize(std::max(Hs.pair_ptr[(size_t)F * F], 1)); Hs.pair_eb.resize(Hs.pair_ea.size());
      cur.assign(Hs.pair_ptr.begin(), Hs.pair_ptr.end() - 1);
    }
    for (int l = 0; l < g->n_mp; l++)
      for (int qa = Hs.lm_ptr[l]; qa < Hs.lm_ptr[l + 1]; qa++) {
        const int ea = Hs.lm_edges[qa], fa = frank[g->e_kf[ea]];
        if (fa < 0) continue;
        for (int qb = Hs.lm_ptr[l]; qb < Hs.lm_ptr[l + 1]; qb++) {
          const int eb = Hs.lm_edges[qb], fb = frank[g->e_kf[eb]];
          if (fb < 0) continue;
          const size_t key = (size_t)fa * F + fb;
          if (pass == 0) Hs.pair_ptr[key + 1]++;
          else { Hs.pair_ea[cur[key]] = ea; Hs.pair_eb[cur[key]] = eb; cur[key]++; }
        }
      }
  }
  Hs.i_color.assign(std::max(g->n_inertial, 1), 0);
  Hs.n_colors = 0;
  {
    std::vector<std::vector<uint8_t>> used;  // used[c][k]: colour c already has an edge at keyframe k
    for (int i = 0; i < g->n_inertial; i++) {
      const int k1 = g->i_kf1[i], k2 = g->i_kf2[i];
      int c = 0;
      while (c < (int)used.size() && (used[c][k1] || used[c][k2])) c++;
      if (c == (int)used.size()) used.emplace_back(K, 0);
      used[c][k1] = used[c][k2] = 1;
      Hs.i_color[i] = c;
    }
    Hs.n_colors = (int)used.size();
  }
  const int nI = g->n_inertial;
  Hs.info.assign(81 * (size_t)std::max(nI, 1), 0.0); Hs.infoG.assign(9 * (size_t)std::max(nI, 1), 0.0);
  Hs.infoA.assign(9 * (size_t)std::max(nI, 1), 0.0);
  for (int i = 0; i < nI; i++) {
    const float* C = g->i_C + 225 * (size_t)i;
    std::vector<double> C9(81), I9, V;
    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) C9[r * 9 + c] = C[r * 15 + c];
    if (!invert_dense(C9, 9, I9)) { err = "lia_solve: singular preintegration covariance"; return ORB_E_ARG; }
    for (int r = 0; r < 9; r++) for (int c = r + 1; c < 9; c++) { const double m = (I9[r * 9 + c] + I9[c * 9 + r]) / 2; I9[r * 9 + c] = I9[c * 9 + r] = m; }
    std::vector<double> A = I9;
    jacobi_eigen(A, 9, V);
    double* O = &Hs.info[81 * (size_t)i];
    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += V[r * 9 + k] * (A[k * 9 + k] < 1e-12 ? 0.0 : A[k * 9 + k]) * V[c * 9 + k];
      O[r * 9 + c] = g->i_last[i] ? s * 1e-2 : s;  // Optimizer.cc:2592-2593
    }
    double G3[9], A3[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { G3[r * 3 + c] = C[(9 + r) * 15 + 9 + c]; A3[r * 3 + c] = C[(12 + r) * 15 + 12 + c]; }
    if (!l_inv3(G3, &Hs.infoG[9 * (size_t)i]) || !l_inv3(A3, &Hs.infoA[9 * (size_t)i])) { err = "lia_solve: singular random-walk covariance"; return ORB_E_ARG; }
  }
  return 0;
}

inline void lia_fill_scalars(const lia_graph_view* g, const LiaHost& Hs, LiaDev& D) {
  D.n_kf = g->n_kf; D.n_mp = g->n_mp; D.n_edges = g->n_edges; D.n_inertial = g->n_inertial; D.np = Hs.np;
  D.iterations = g->iterations; D.lambda_init = g->lambda_init;
  D.n_free = (int)Hs.free_kf.size(); D.n_colors = Hs.n_colors;
  for (int i = 0; i < 9; i++) D.Rcb[i] = g->Rcb[i];
  for (int i = 0; i < 3; i++) { D.tcb[i] = g->tcb[i]; D.tbc[i] = g->tbc[i]; }
  D.fx = g->fx; D.fy = g->fy; D.cx = g->cx; D.cy = g->cy; D.bf = g->bf;
  // thHuberMono / thHuberStereo are floats (Optimizer.cc:2646-2649), the inertial delta a double (:2595)
  const float hm = (float)sqrt(5.991), hs = (float)sqrt(7.815);
  const double hi = sqrt(16.92);
  D.huber_mono_delta = hm; D.huber_mono_dsqr = (float)((double)hm * (double)hm);
  D.huber_stereo_delta = hs; D.huber_stereo_dsqr = (float)((double)hs * (double)hs);
  D.huber_in_delta = hi; D.huber_in_dsqr = (float)(hi * hi);
}

inline void lia_write_out(const lia_graph_view* g, const double* pose, const double* vel, const double* bg, const double* ba,
                          const double* pt, double* kf_out, double* mp_out) {
  for (int k = 0; k < g->n_kf; k++) {
    double* o = kf_out + 21 * (size_t)k;
    memcpy(o, pose + 24 * (size_t)k + 12, 72); memcpy(o + 9, pose + 24 * (size_t)k + 21, 24);
    memcpy(o + 12, vel + 3 * k, 24); memcpy(o + 15, bg + 3 * k, 24); memcpy(o + 18, ba + 3 * k, 24);
  }
  if (g->n_mp) memcpy(mp_out, pt, sizeof(double) * 3 * (size_t)g->n_mp);
}

// All state / system / result buffers of one solve in host memory, wired into a LiaDev
struct LiaHostBuffers {
  std::vector<double> pose, pose_bak, vel, bg, ba, pt, vel_bak, bg_bak, ba_bak, pt_bak, H, b, Hll, bl, W, Dinv, Sm, bs, x, verr,
      ierr, Dg, chi, st, Hpe, Ye;
  std::vector<uint8_t> dp;
  LiaDev D;
  LiaHostBuffers(const lia_graph_view* g, const LiaHost& Hs) {
    const size_t K = g->n_kf, L = g->n_mp, E = g->n_edges, NI = g->n_inertial, np = Hs.np;
    const size_t L1 = std::max<size_t>(L, 1), E1 = std::max<size_t>(E, 1), N1 = std::max<size_t>(NI, 1);
    pose = Hs.pose; pose_bak.resize(24 * K);
    vel.assign(g->kf_vel, g->kf_vel + 3 * K); bg.assign(g->kf_bg, g->kf_bg + 3 * K); ba.assign(g->kf_ba, g->kf_ba + 3 * K);
    pt.assign(3 * L1, 0.0);
    if (L) memcpy(pt.data(), g->mp_pos, 24 * L);
    vel_bak.resize(3 * K); bg_bak.resize(3 * K); ba_bak.resize(3 * K); pt_bak.resize(3 * L1);
    H.resize(np * np); b.resize(np); Hll.resize(9 * L1); bl.resize(3 * L1); W.resize(18 * E1); Dinv.resize(9 * L1);
    Sm.resize(np * np); bs.resize(np); x.assign(np + 3 * L, 0.0); verr.resize(3 * E1); ierr.resize(15 * N1); Dg.resize(np);
    chi.resize(E1); st.assign(8, 0.0); dp.resize(E1); Hpe.assign(42 * E1, 0.0); Ye.assign(18 * E1, 0.0);
    lia_fill_scalars(g, Hs, D);
    D.kf_fixed = g->kf_fixed; D.kf_has_imu = g->kf_has_imu; D.ip = Hs.ip.data(); D.iv = Hs.iv.data(); D.ig = Hs.ig.data(); D.ia = Hs.ia.data();
    D.e_kf = g->e_kf; D.e_mp = g->e_mp; D.e_stereo = g->e_stereo; D.e_obs = g->e_obs; D.e_is2 = g->e_inv_sigma2;
    D.lm_ptr = Hs.lm_ptr.data(); D.lm_edges = Hs.lm_edges.data();
    D.kf_ptr = Hs.kf_ptr.data(); D.kf_edges = Hs.kf_edges.data(); D.pair_ptr = Hs.pair_ptr.data(); D.pair_ea = Hs.pair_ea.data();
    D.pair_eb = Hs.pair_eb.data(); D.free_kf = Hs.free_kf.data(); D.i_color = Hs.i_color.data();
    D.i_kf1 = g->i_kf1; D.i_kf2 = g->i_kf2; D.i_dR = g->i_dR; D.i_dV = g->i_dV; D.i_dP = g->i_dP; D.i_JRg = g->i_JRg; D.i_JVg = g->i_JVg;
    D.i_JVa = g->i_JVa; D.i_JPg = g->i_JPg; D.i_JPa = g->i_JPa; D.i_bias = g->i_bias; D.i_dT = g->i_dT; D.i_last = g->i_last;
    D.info = Hs.info.data(); D.infoG = Hs.infoG.data(); D.infoA = Hs.infoA.data();
    D.pose = pose.data(); D.pose_bak = pose_bak.data(); D.vel = vel.data(); D.bg = bg.data(); D.ba = ba.data(); D.pt = pt.data();
    D.vel_bak = vel_bak.data(); D.bg_bak = bg_bak.data(); D.ba_bak = ba_bak.data(); D.pt_bak = pt_bak.data();
    D.H = H.data(); D.b = b.data(); D.Hll = Hll.data(); D.bl = bl.data(); D.W = W.data(); D.Dinv = Dinv.data(); D.S = Sm.data();
    D.bs = bs.data(); D.x = x.data(); D.verr = verr.data(); D.ierr = ierr.data(); D.Dg = Dg.data();
    D.chi2_out = chi.data(); D.depth_pos_out = dp.data(); D.stats = st.data(); D.Hpe = Hpe.data(); D.Ye = Ye.data();
  }
};

}  // namespace orbb200
