// Array formulation of ORBextractor::DistributeOctTree (reference
// src/ORBextractor.cc:555-779, DivideNode :480-536) that one CTA executes per
// (frame, pyramid level).  The reference walks a std::list and copies KeyPoints
// into child vectors; here the list is an array indexed by list position and the
// points only carry the position of the node that owns them:
//
//   * one "split step" = a set of nodes processed in a known order (BFS pass:
//     every node with >1 point, in list order; overshoot phase: the nodes sorted
//     by (count, UL.x), largest first, cut where the list reaches N);
//   * children are created in (processing order, n1..n4) order and the
//     reference push_front()s them, so child number c of T lands at list
//     position T-1-c and every untouched node keeps its relative order behind
//     them -- both are prefix sums;
//   * "first max response wins" (:757-776) over a vKeys vector that preserves
//     candidate order == max response, ties to the smallest candidate order key
//     (cell row, cell col, y, x), so no point ever has to be moved.
//
// The body is written against a Backend (thread index, barrier, atomics, scan)
// so the same code runs as one CUDA CTA (octree kernel in orb_extract.cu) and
// single-threaded on the host (orb_debug_octree_host, used by the CPU tests to
// pin this formulation against the list-based oracle).
#pragma once
#include <stdint.h>

#include "introsort_emul.h"

namespace orbb200 {

struct OctreeLevelParams {
  int bandW, bandH;          // maxBorderX-minBorderX, maxBorderY-minBorderY
  int N;                     // mnFeaturesPerLevel[level]
  int nIni;                  // round(bandW / bandH)
  float hX;                  // (float)bandW / nIni
  int wCell, hCell, nCols;   // FAST cell grid (candidate order key)
  int node_cap;              // capacity of every per-node array
};

// Scratch carved by the host; every per-node array has node_cap entries (x4
// where noted).
struct OctreeScratch {
  int* pt_node;       // [n]   list position of the node owning the point
  uint8_t* pt_q;      // [n]   quadrant chosen in the current step
  int* nd[2][5];      // [buffer][ulx,uly,brx,bry,cnt][node_cap]
  int* childcnt;      // [4*node_cap]
  int* cidx;          // [4*node_cap] creation index of child (processing k, q)
  int* eidx;          // [4*node_cap] index among children with >1 point
  int* remap;         // [4*node_cap] new list position per (old node, q)
  int* rank;          // [node_cap]   processing rank of a node or -1
  int* proc;          // [node_cap]   node processed k-th
  int* surv;          // [node_cap]
  int* tmp;           // [node_cap]
  int* expand_pos;    // [node_cap]   nodes with >1 point, creation order
  SortNode* sortbuf;  // [node_cap]
  int* sortwork;      // [6*node_cap] range lists of the level-synchronous sort
  unsigned long long* best;  // [node_cap]
};

// Candidate: xy = x' | y'<<16 (band-relative, i.e. the reference's coordinates
// before minBorder is added back, ORBextractor.cc:863-868), score = FAST score.
struct Cand {
  uint32_t xy;
  uint32_t score;
};

ORB_HD uint32_t cand_order_key(int x, int y, const OctreeLevelParams& p) {
  // cv::FAST tests cell-local columns [3, w-3): band-relative x in
  // [j*wCell+3, (j+1)*wCell+3)  =>  j = (x-3)/wCell.  The reference visits cells
  // row-major and cv::FAST emits row-major inside a cell.
  int cj = (x - 3) / p.wCell, ci = (y - 3) / p.hCell;
  int lx = (x - 3) - cj * p.wCell, ly = (y - 3) - ci * p.hCell;
  return ((uint32_t)(ci * p.nCols + cj) << 14) | ((uint32_t)ly << 7) | (uint32_t)lx;
}

ORB_HD int quadrant_of(int x, int y, int ulx, int uly, int brx, int bry) {
  // DivideNode :480-536: halfX = ceil((UR.x-UL.x)/2) etc.
  const int mx = ulx + ((brx - ulx + 1) >> 1);
  const int my = uly + ((bry - uly + 1) >> 1);
  return (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
}

// Returns the number of selected keypoints; out[3*pos..] = {x', y', score} in
// final list order (== order of the reference's vResultKeys).
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
#ifdef __CUDACC__
__host__ __device__
#endif
int octree_select(BE& be, const Cand* cand, int n, const OctreeLevelParams& p,
                         const OctreeScratch& s, int* out) {
  const int tid = be.tid(), nt = be.nthreads();
  const int N = p.N;
  int cur = 0;

  // ---- initial nodes and point assignment (:557-583)
  {
    int** A = (int**)s.nd[0];
    for (int i = tid; i < p.nIni; i += nt) {
      A[0][i] = (int)(p.hX * (float)i);
      A[2][i] = (int)(p.hX * (float)(i + 1));
      A[1][i] = 0;
      A[3][i] = p.bandH;
      A[4][i] = 0;
    }
    be.sync();
    for (int i = tid; i < n; i += nt) {
      const int x = (int)(cand[i].xy & 0xffffu);
      const int node = (int)((float)x / p.hX);
      s.pt_node[i] = node;
      be.atomic_add(&A[4][node], 1);
    }
    be.sync();
    // erase empty nodes, keep order (:587-598)
    for (int i = tid; i < p.nIni; i += nt) s.surv[i] = A[4][i] > 0 ? 1 : 0;
    be.sync();
    const int M0 = be.exclusive_scan(s.surv, p.nIni);
    int** B = (int**)s.nd[1];
    for (int i = tid; i < p.nIni; i += nt) {
      if (A[4][i] > 0) {
        const int pos = s.surv[i];
        for (int f = 0; f < 5; f++) B[f][pos] = A[f][i];
      }
    }
    be.sync();
    for (int i = tid; i < n; i += nt) s.pt_node[i] = s.surv[s.pt_node[i]];
    be.sync();
    cur = 1;
    // M0 is carried below
    int M = M0;
    int mode = 0;  // 0: BFS passes (:607-676), 1: overshoot phase (:687-750)
    int P = 0;     // entries of expand_pos
    bool finish = (M == 0);

    while (!finish) {
      int** Acur = (int**)s.nd[cur];
      int** Bnew = (int**)s.nd[cur ^ 1];
      int* ulx = Acur[0]; int* uly = Acur[1]; int* brx = Acur[2]; int* bry = Acur[3];
      int* cnt = Acur[4];

      // 1. children point counts of every node that could be split
      for (int j = tid; j < 4 * M; j += nt) s.childcnt[j] = 0;
      be.sync();
      // 4 points per trip: the four independent (pt_node, xy) loads are issued together, which is
      // what bounds this latency-bound kernel (one CTA walks ~10^4 points a dozen times)
      for (int i0 = tid; i0 < n; i0 += 4 * nt) {
        int node[4];
        uint32_t xy[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * nt;
          node[u] = i < n ? s.pt_node[i] : -1;
          xy[u] = i < n ? cand[i].xy : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (node[u] < 0 || cnt[node[u]] <= 1) continue;
          const int q = quadrant_of((int)(xy[u] & 0xffffu), (int)(xy[u] >> 16), ulx[node[u]], uly[node[u]],
                                    brx[node[u]], bry[node[u]]);
          s.pt_q[i0 + u * nt] = (uint8_t)q;
          be.atomic_add(&s.childcnt[4 * node[u] + q], 1);
        }
      }
      be.sync();

      // 2. processing order: rank[node] in [0,K) or -1; proc[k] = node
      int K;
      if (mode == 0) {
        for (int j = tid; j < M; j += nt) s.tmp[j] = cnt[j] > 1 ? 1 : 0;
        be.sync();
        K = be.exclusive_scan(s.tmp, M);
        for (int j = tid; j < M; j += nt) {
          if (cnt[j] > 1) { s.rank[j] = s.tmp[j]; s.proc[s.tmp[j]] = j; }
          else s.rank[j] = -1;
        }
        be.sync();
      } else {
        for (int e = tid; e < P; e += nt) {
          const int node = s.expand_pos[e];
          s.sortbuf[e] = make_sort_node(cnt[node], ulx[node], node);
        }
        be.sync();
        introsort_levels(be, s.sortbuf, P, s.sortwork, p.node_cap);  // ends with a barrier
        // split from the back (:701) until the list holds >= N nodes (:746)
        for (int k = tid; k < P; k += nt) {
          const int node = s.sortbuf[P - 1 - k].id;
          int ne = 0;
          for (int q = 0; q < 4; q++) ne += s.childcnt[4 * node + q] > 0;
          s.tmp[k] = ne - 1;
        }
        be.sync();
        be.exclusive_scan(s.tmp, P);
        int* counter = be.shared_int(0);
        if (tid == 0) *counter = 0;
        be.sync();
        int local = 0;
        for (int k = tid; k < P; k += nt) {
          const int node = s.sortbuf[P - 1 - k].id;
          int ne = 0;
          for (int q = 0; q < 4; q++) ne += s.childcnt[4 * node + q] > 0;
          if (M + s.tmp[k] + (ne - 1) < N) local++;
        }
        if (local) be.atomic_add(counter, local);
        be.sync();
        K = *counter + 1;
        if (K > P) K = P;
        be.sync();
        for (int j = tid; j < M; j += nt) s.rank[j] = -1;
        be.sync();
        for (int k = tid; k < K; k += nt) {
          const int node = s.sortbuf[P - 1 - k].id;
          s.rank[node] = k;
          s.proc[k] = node;
        }
        be.sync();
      }

      // 3. creation index of each non-empty child, and index among expandable ones
      for (int j = tid; j < 4 * K; j += nt) {
        const int c = s.childcnt[4 * s.proc[j >> 2] + (j & 3)];
        s.cidx[j] = c > 0 ? 1 : 0;
        s.eidx[j] = c > 1 ? 1 : 0;
      }
      be.sync();
      const int T = be.exclusive_scan(s.cidx, 4 * K);
      const int E = be.exclusive_scan(s.eidx, 4 * K);

      // 4. untouched nodes keep their order behind the new children
      for (int j = tid; j < M; j += nt) s.surv[j] = s.rank[j] < 0 ? 1 : 0;
      be.sync();
      const int U = be.exclusive_scan(s.surv, M);

      // 5. write the new list
      for (int j = tid; j < 4 * K; j += nt) {
        const int node = s.proc[j >> 2], q = j & 3;
        const int c = s.childcnt[4 * node + q];
        if (c > 0) {
          const int pos = T - 1 - s.cidx[j];
          const int x0 = ulx[node], y0 = uly[node], x1 = brx[node], y1 = bry[node];
          const int mx = x0 + ((x1 - x0 + 1) >> 1), my = y0 + ((y1 - y0 + 1) >> 1);
          Bnew[0][pos] = (q & 1) ? mx : x0;
          Bnew[2][pos] = (q & 1) ? x1 : mx;
          Bnew[1][pos] = (q & 2) ? my : y0;
          Bnew[3][pos] = (q & 2) ? y1 : my;
          Bnew[4][pos] = c;
          s.remap[4 * node + q] = pos;
          if (c > 1) s.expand_pos[s.eidx[j]] = pos;
        }
      }
      for (int j = tid; j < M; j += nt) {
        if (s.rank[j] < 0) {
          const int pos = T + s.surv[j];
          for (int f = 0; f < 5; f++) Bnew[f][pos] = Acur[f][j];
          s.remap[4 * j] = pos;
        }
      }
      be.sync();
      for (int i0 = tid; i0 < n; i0 += 4 * nt) {
        int node[4], q[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * nt;
          node[u] = i < n ? s.pt_node[i] : -1;
          q[u] = i < n ? (int)s.pt_q[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (node[u] >= 0) s.pt_node[i0 + u * nt] = s.remap[4 * node[u] + (s.rank[node[u]] >= 0 ? q[u] : 0)];
      }
      be.sync();

      const int Mnew = T + U;
      if (Mnew >= N || Mnew == M) finish = true;             // :680-683 / :749-750
      else if (mode == 0 && Mnew + 3 * E > N) mode = 1;       // :684
      P = E;
      M = Mnew;
      cur ^= 1;
    }

    // ---- best point per node (:755-776): max response, first in vKeys order
    int** F = (int**)s.nd[cur];
    (void)F;
    for (int j = tid; j < M; j += nt) s.best[j] = 0ull;
    be.sync();
    for (int i = tid; i < n; i += nt) {
      const uint32_t xy = cand[i].xy;
      const uint32_t key = cand_order_key((int)(xy & 0xffffu), (int)(xy >> 16), p);
      const unsigned long long v =
          ((unsigned long long)(cand[i].score + 1u) << 32) | (unsigned long long)(0xffffffffu - key);
      be.atomic_max64(&s.best[s.pt_node[i]], v);
    }
    be.sync();
    for (int i = tid; i < n; i += nt) {
      const uint32_t xy = cand[i].xy;
      const uint32_t key = cand_order_key((int)(xy & 0xffffu), (int)(xy >> 16), p);
      const unsigned long long v =
          ((unsigned long long)(cand[i].score + 1u) << 32) | (unsigned long long)(0xffffffffu - key);
      const int node = s.pt_node[i];
      if (s.best[node] == v) {
        out[3 * node + 0] = (int)(xy & 0xffffu);
        out[3 * node + 1] = (int)(xy >> 16);
        out[3 * node + 2] = (int)cand[i].score;
      }
    }
    be.sync();
    return M;
  }
}

// ------------------------------------------------------------ host backend
struct HostBackend {
  int shared[4];
  int tid() const { return 0; }
  int nthreads() const { return 1; }
  void sync() {}
  int atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
  void atomic_max64(unsigned long long* p, unsigned long long v) { if (v > *p) *p = v; }
  int* shared_int(int i) { return &shared[i]; }
  int exclusive_scan(int* d, int n) {
    int acc = 0;
    for (int i = 0; i < n; i++) { int v = d[i]; d[i] = acc; acc += v; }
    return acc;
  }
};

}  // namespace orbb200
