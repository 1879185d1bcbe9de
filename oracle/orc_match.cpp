// TEST INFRASTRUCTURE ONLY (see orc_common.h).  CPU restatement of the
// reference's Hamming / projection matchers on the flat views of
// include/orb_b200.h (interface types only; no product code is used).
//
// Restates (paths relative to /root/reference):
//   src/ORBmatcher.cc:43-141     SearchByProjection(Frame&, vector<MapPoint*>&, ...)   (Nleft == -1)
//   src/ORBmatcher.cc:215-221    RadiusByViewingCos
//   src/ORBmatcher.cc:907-1146   SearchForTriangulation                              (no mpCamera2)
//   src/ORBmatcher.cc:1676-1887  SearchByProjection(Frame& Cur, const Frame& Last,..) (Nleft == -1)
//   src/ORBmatcher.cc:2012-2053  ComputeThreeMaxima
//   src/ORBmatcher.cc:2058-2074  DescriptorDistance
//   src/Frame.cc:385-416         AssignFeaturesToGrid, :725-735 PosInGrid
//   src/Frame.cc:657-723         GetFeaturesInArea
//   src/CameraModels/Pinhole.cpp:50-56 project(Vector3f), :107-129 epipolarConstrain (line test only;
//       F12 is an input because its Eigen 3x3 inverses are not restated)
//   Sophus SE3f * Vector3f = Eigen Quaternion::_transformVector + translation
// Float semantics: strict IEEE single, no FMA (-ffp-contract=off).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/orb_b200.h"

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;
const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:44-45

// ORBmatcher.cc:2058-2074
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  const int32_t* pa = (const int32_t*)a;
  const int32_t* pb = (const int32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

struct Grid {
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  const orb_frame_view* F;
  // Frame.cc:385-416 / :725-735
  explicit Grid(const orb_frame_view* f) : F(f) {
    for (int i = 0; i < F->n; i++) {
      const orb_keypoint& kp = F->keys[i];
      int px = (int)std::round((kp.x - F->min_x) * F->grid_w_inv);
      int py = (int)std::round((kp.y - F->min_y) * F->grid_h_inv);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
  // Frame.cc:657-723
  void in_area(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const float factorX = r, factorY = r;
    const int nMinCellX = std::max(0, (int)std::floor((x - F->min_x - factorX) * F->grid_w_inv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - F->min_x + factorX) * F->grid_w_inv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - F->min_y - factorY) * F->grid_h_inv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - F->min_y + factorY) * F->grid_h_inv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int idx : cell[ix][iy]) {
          const orb_keypoint& kp = F->keys[idx];
          if (bCheckLevels) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
          }
          const float distx = kp.x - x, disty = kp.y - y;
          if (std::fabs(distx) < factorX && std::fabs(disty) < factorY) out.push_back(idx);
        }
  }
};

// ORBmatcher.cc:2012-2053 on bin sizes
void three_maxima(const int* size, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = size[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

}  // namespace

extern "C" {

int orc_ham_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
// ComputeThreeMaxima alone (pinned against the reference's object code by tests/test_ref_parity.py)
void orc_three_maxima(const int* sizes, int L, int* ind /*3, in: initial values*/) {
  three_maxima(sizes, L, ind[0], ind[1], ind[2]);
}

// ORBmatcher.cc:43-141
int orc_match_project_local(const orb_frame_view* F, const orb_mappoint_view* mps, float th, float nn_ratio,
                            int far_points, float th_far, int32_t* assign) {
  Grid grid(F);
  std::vector<uint8_t> taken(F->n, 0);
  if (F->kp_taken) memcpy(taken.data(), F->kp_taken, F->n);
  for (int i = 0; i < F->n; i++) assign[i] = -1;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  std::vector<int> vIndices;
  for (int iMP = 0; iMP < mps->n; iMP++) {
    if (!mps->track_in_view[iMP]) continue;
    if (far_points && mps->depth[iMP] > th_far) continue;
    if (mps->is_bad[iMP]) continue;
    const int nPredictedLevel = mps->scale_level[iMP];
    float r = (mps->view_cos[iMP] > 0.998) ? 2.5f : 4.0f;
    if (bFactor) r *= th;
    const float rs = r * F->scale_factors[nPredictedLevel];
    grid.in_area(mps->proj_x[iMP], mps->proj_y[iMP], rs, nPredictedLevel - 1, nPredictedLevel, vIndices);
    if (vIndices.empty()) continue;
    const uint8_t* d = mps->desc + (size_t)iMP * 32;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : vIndices) {
      if (taken[idx]) continue;
      if (F->u_right && F->u_right[idx] > 0) {
        const float er = std::fabs(mps->proj_xr[iMP] - F->u_right[idx]);
        if (er > rs) continue;
      }
      const int dist = descriptor_distance(d, F->desc + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
        bestLevel = F->keys[idx].octave; bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = F->keys[idx].octave; bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
      if (bestLevel != bestLevel2 || bestDist <= nn_ratio * bestDist2) {
        assign[bestIdx] = iMP;
        // the slot now holds this MapPoint: later candidates skip it iff Observations()>0
        taken[bestIdx] = mps->has_obs[iMP] ? 1 : 0;
        nmatches++;
      }
    }
  }
  return nmatches;
}

// ORBmatcher.cc:1676-1887
int orc_match_project_last(const orb_frame_view* C, const orb_lastframe_view* L, const float* Tcw, int forward,
                           int backward, float th, int check_ori, int32_t* assign) {
  Grid grid(C);
  std::vector<uint8_t> taken(C->n, 0);
  if (C->kp_taken) memcpy(taken.data(), C->kp_taken, C->n);
  for (int i = 0; i < C->n; i++) assign[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  const float qx = Tcw[0], qy = Tcw[1], qz = Tcw[2], qw = Tcw[3];
  std::vector<int> vIndices2;
  for (int i = 0; i < L->n; i++) {
    if (!L->has_mp[i]) continue;
    // Sophus SE3f * Vector3f: Eigen Quaternion::_transformVector, then + t
    const float vx = L->world_pos[3 * i], vy = L->world_pos[3 * i + 1], vz = L->world_pos[3 * i + 2];
    float ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
    ux += ux; uy += uy; uz += uz;
    const float cx_ = qy * uz - qz * uy, cy_ = qz * ux - qx * uz, cz_ = qx * uy - qy * ux;
    const float xc = (vx + qw * ux + cx_) + Tcw[4];
    const float yc = (vy + qw * uy + cy_) + Tcw[5];
    const float zc = (vz + qw * uz + cz_) + Tcw[6];
    const float invzc = (float)(1.0 / zc);
    if (invzc < 0) continue;
    const float u = C->fx * xc / zc + C->cx;
    const float v = C->fy * yc / zc + C->cy;
    if (u < C->min_x || u > C->max_x) continue;
    if (v < C->min_y || v > C->max_y) continue;
    const int nLastOctave = L->octave[i];
    const float radius = th * C->scale_factors[nLastOctave];
    if (forward) grid.in_area(u, v, radius, nLastOctave, -1, vIndices2);
    else if (backward) grid.in_area(u, v, radius, 0, nLastOctave, vIndices2);
    else grid.in_area(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
    if (vIndices2.empty()) continue;
    const uint8_t* dMP = L->desc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      if (taken[i2]) continue;
      if (C->u_right && C->u_right[i2] > 0) {
        const float ur = u - C->bf * invzc;
        const float er = std::fabs(ur - C->u_right[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(dMP, C->desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      assign[bestIdx2] = i;
      taken[bestIdx2] = L->has_obs[i] ? 1 : 0;
      nmatches++;
      if (check_ori) rotHist[rot_bin(L->angle[i], C->keys[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
    three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { assign[idx] = -2; nmatches--; }
    }
  }
  return nmatches;
}

// ORBmatcher.cc:907-1146 (pinhole, no second camera; F12 and epipole are inputs)
int orc_match_triangulate(const orb_frame_view* K1, const orb_frame_view* K2, const orb_featvec_view* fv1,
                          const orb_featvec_view* fv2, const float* F12, const float* ep, int only_stereo,
                          int coarse, int check_ori, int32_t* pairs, int cap) {
  std::vector<int> vMatches12(K1->n, -1);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  int a = 0, b = 0;
  while (a < fv1->n_nodes && b < fv2->n_nodes) {
    if (fv1->node_ids[a] == fv2->node_ids[b]) {
      for (int p1 = fv1->ptr[a]; p1 < fv1->ptr[a + 1]; p1++) {
        const int idx1 = fv1->idx[p1];
        if (K1->kp_taken && K1->kp_taken[idx1]) continue;
        const bool bStereo1 = K1->u_right && K1->u_right[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const orb_keypoint& kp1 = K1->keys[idx1];
        const uint8_t* d1 = K1->desc + (size_t)idx1 * 32;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int p2 = fv2->ptr[b]; p2 < fv2->ptr[b + 1]; p2++) {
          const int idx2 = fv2->idx[p2];
          if (K2->kp_taken && K2->kp_taken[idx2]) continue;  // vbMatched2 is never set (:949, :1004)
          const bool bStereo2 = K2->u_right && K2->u_right[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = descriptor_distance(d1, K2->desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;
          const orb_keypoint& kp2 = K2->keys[idx2];
          if (!bStereo1 && !bStereo2) {
            const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
            if (distex * distex + distey * distey < 100 * K2->scale_factors[kp2.octave]) continue;
          }
          bool ok = coarse != 0;
          if (!ok) {  // Pinhole::epipolarConstrain, Pinhole.cpp:114-128
            const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = la * kp2.x + lb * kp2.y + lc;
            const float den = la * la + lb * lb;
            if (den != 0) {
              const float dsqr = num * num / den;
              ok = dsqr < 3.84 * K2->level_sigma2[kp2.octave];
            }
          }
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          vMatches12[idx1] = bestIdx2;
          nmatches++;
          if (check_ori) rotHist[rot_bin(kp1.angle, K2->keys[bestIdx2].angle)].push_back(idx1);
        }
      }
      a++; b++;
    } else if (fv1->node_ids[a] < fv2->node_ids[b]) {
      while (a < fv1->n_nodes && fv1->node_ids[a] < fv2->node_ids[b]) a++;  // lower_bound
    } else {
      while (b < fv2->n_nodes && fv2->node_ids[b] < fv1->node_ids[a]) b++;
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
    three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { vMatches12[idx] = -1; nmatches--; }
    }
  }
  int m = 0;
  for (int i = 0; i < K1->n; i++) {
    if (vMatches12[i] < 0) continue;
    if (m < cap) { pairs[2 * m] = i; pairs[2 * m + 1] = vMatches12[i]; }
    m++;
  }
  return m;
}

}  // extern "C"
