// TEST INFRASTRUCTURE ONLY (see orc_common.h) -- CPU restatement of Frame::isInFrustum
// (reference src/Frame.cc:512-570, the Nleft == -1 branch) with MapPoint::PredictScale
// (src/MapPoint.cc:531-546) and Get{Min,Max}DistanceInvariance (:505-515), SURVEY.md 8(f-3),
// on flat arrays.  PARITY UNPINNED BY THE REFERENCE.  The Eigen expressions (Eigen is
// un-vendored: find_package(Eigen3 3.1.0), CMakeLists.txt:41) are restated with Eigen's
// published evaluation order for fixed-size 3-vectors: coefficient-based product and
// redux_novec_unroller, i.e. a 3-term sum is a0 + (a1 + a2); strict IEEE single, no FMA.
// An independent numpy float32 reading in tests/test_frustum_oracle.py pins it.
#include <math.h>
#include <stdint.h>

#include "../include/orb_b200.h"

namespace {
inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
}

extern "C" {

// Outputs follow the MapPoint members the reference writes: track_in_view, proj_x, proj_y are always
// written (mTrackProjX/Y = -1 or the projection, exactly as :514-516, :538-539); proj_xr, depth,
// scale_level, view_cos only where the point is in view (:556-565) -- other entries are left untouched,
// like the stale members of the reference.  Returns the number of points in view.
int orc_is_in_frustum(const orb_frustum_view* v, float viewingCosLimit, uint8_t* track_in_view, float* proj_x,
                      float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos, float* depth) {
  int n_in = 0;
  for (int i = 0; i < v->n; i++) {
    track_in_view[i] = 0; proj_x[i] = -1; proj_y[i] = -1;                       // :514-516
    const float* P = v->world_pos + 3 * i;
    const float* R = v->Rcw;
    float Pc[3];
    for (int r = 0; r < 3; r++) Pc[r] = sum3(R[3 * r] * P[0], R[3 * r + 1] * P[1], R[3 * r + 2] * P[2]) + v->tcw[r];  // :522
    const float Pc_dist = sqrtf(sum3(Pc[0] * Pc[0], Pc[1] * Pc[1], Pc[2] * Pc[2]));  // :523
    const float PcZ = Pc[2];
    const float invz = 1.0f / PcZ;
    if (PcZ < 0.0f) continue;                                                    // :528-529
    const float u = v->fx * Pc[0] / Pc[2] + v->cx;                               // Pinhole::project(Vector3f), Pinhole.cpp:50-56
    const float w = v->fy * Pc[1] / Pc[2] + v->cy;
    if (u < v->min_x || u > v->max_x) continue;                                  // :533-536
    if (w < v->min_y || w > v->max_y) continue;
    proj_x[i] = u; proj_y[i] = w;                                                // :538-539
    const float maxDistance = 1.2f * v->max_dist[i];                             // MapPoint.cc:505-515
    const float minDistance = 0.8f * v->min_dist[i];
    const float PO[3] = {P[0] - v->Ow[0], P[1] - v->Ow[1], P[2] - v->Ow[2]};
    const float dist = sqrtf(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
    if (dist < minDistance || dist > maxDistance) continue;                      // :547-548
    const float* Pn = v->normal + 3 * i;
    const float viewCos = sum3(PO[0] * Pn[0], PO[1] * Pn[1], PO[2] * Pn[2]) / dist;  // :553
    if (viewCos < viewingCosLimit) continue;
    // MapPoint::PredictScale(dist, Frame*): float log (std::log(float) through the global using-directive
    // of DBoW2/TemplatedVocabulary.h:36), float division, ceil, clamp
    const float ratio = v->max_dist[i] / dist;
    int nScale = (int)ceilf(logf(ratio) / v->log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= v->n_levels) nScale = v->n_levels - 1;
    track_in_view[i] = 1;                                                        // :560-568
    proj_xr[i] = u - v->bf * invz;
    depth[i] = Pc_dist;
    scale_level[i] = nScale;
    view_cos[i] = viewCos;
    n_in++;
  }
  return n_in;
}

}  // extern "C"
