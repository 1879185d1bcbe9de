// Per-point body of Frame::isInFrustum (reference src/Frame.cc:512-570, Nleft == -1) with
// MapPoint::PredictScale (src/MapPoint.cc:531-546) and the distance-invariance getters
// (:505-515).  Written once for both sides: the frustum kernel (frustum.cu) runs it per thread,
// frustum_debug_host runs it in a loop so the CPU tests can hold exactly this source against the
// oracle.  Strict IEEE single (the library is built with -fmad=false; host code has no FMA target);
// 3-term sums in Eigen's fixed-size order a0 + (a1 + a2) (coefficient-based product / redux unroller).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "introsort_emul.h"  // ORB_HD

namespace orbb200 {

struct FrustumFrame {
  float Rcw[9], tcw[3], Ow[3];
  float fx, fy, cx, cy, bf;
  float min_x, max_x, min_y, max_y;
  float log_scale_factor;
  int n_levels;
};

struct FrustumPoint {
  uint8_t in_view;
  uint8_t full;  // 1: proj_xr / depth / level / view_cos are valid (the reference writes them only then)
  float proj_x, proj_y, proj_xr, view_cos, depth;
  int level;
};

ORB_HD float frustum_sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

// std::log(float) of the reference.  glibc's logf on the host; on the device the correctly rounded
// value through the double-precision log (CUDA's logf has a 1-ulp error bound, which could flip a
// ceil() on an exact power of the scale factor).
ORB_HD float frustum_logf(float x) {
#ifdef __CUDA_ARCH__
  return (float)log((double)x);
#else
  return logf(x);
#endif
}

ORB_HD FrustumPoint frustum_point(const FrustumFrame& F, const float* P, const float* Pn, float min_dist,
                                  float max_dist, float cos_limit) {
  FrustumPoint o;
  o.in_view = 0; o.full = 0; o.proj_x = -1.0f; o.proj_y = -1.0f;                 // :514-516
  o.proj_xr = 0; o.view_cos = 0; o.depth = 0; o.level = 0;
  float Pc[3];
  for (int r = 0; r < 3; r++)
    Pc[r] = frustum_sum3(F.Rcw[3 * r] * P[0], F.Rcw[3 * r + 1] * P[1], F.Rcw[3 * r + 2] * P[2]) + F.tcw[r];  // :522
  const float Pc_dist = sqrtf(frustum_sum3(Pc[0] * Pc[0], Pc[1] * Pc[1], Pc[2] * Pc[2]));
  const float PcZ = Pc[2];
  const float invz = 1.0f / PcZ;
  if (PcZ < 0.0f) return o;                                                     // :528-529
  const float u = F.fx * Pc[0] / Pc[2] + F.cx;                                  // Pinhole.cpp:43-49
  const float v = F.fy * Pc[1] / Pc[2] + F.cy;
  if (u < F.min_x || u > F.max_x) return o;                                     // :533-536
  if (v < F.min_y || v > F.max_y) return o;
  o.proj_x = u; o.proj_y = v;                                                   // :538-539
  const float maxDistance = 1.2f * max_dist, minDistance = 0.8f * min_dist;     // MapPoint.cc:505-515
  const float PO[3] = {P[0] - F.Ow[0], P[1] - F.Ow[1], P[2] - F.Ow[2]};
  const float dist = sqrtf(frustum_sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
  if (dist < minDistance || dist > maxDistance) return o;                       // :547-548
  const float viewCos = frustum_sum3(PO[0] * Pn[0], PO[1] * Pn[1], PO[2] * Pn[2]) / dist;  // :553
  if (viewCos < cos_limit) return o;
  const float ratio = max_dist / dist;                                          // MapPoint.cc:531-546
  int nScale = (int)ceilf(frustum_logf(ratio) / F.log_scale_factor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= F.n_levels) nScale = F.n_levels - 1;
  o.in_view = 1; o.full = 1;                                                    // :560-568
  o.proj_xr = u - F.bf * invz;
  o.depth = Pc_dist;
  o.level = nScale;
  o.view_cos = viewCos;
  return o;
}

inline FrustumFrame frustum_frame_of(const orb_frustum_view& v) {
  FrustumFrame F;
  for (int i = 0; i < 9; i++) F.Rcw[i] = v.Rcw[i];
  for (int i = 0; i < 3; i++) { F.tcw[i] = v.tcw[i]; F.Ow[i] = v.Ow[i]; }
  F.fx = v.fx; F.fy = v.fy; F.cx = v.cx; F.cy = v.cy; F.bf = v.bf;
  F.min_x = v.min_x; F.max_x = v.max_x; F.min_y = v.min_y; F.max_y = v.max_y;
  F.log_scale_factor = v.log_scale_factor; F.n_levels = v.n_levels;
  return F;
}

}  // namespace orbb200
