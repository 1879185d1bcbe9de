// CPU stand-in for "one CTA runs octree_kernel": the product's DistributeOctTree source (csrc/octree_core.h) executed by
// T real threads under ThreadSanitizer -- a pthread barrier as __syncthreads(), std::atomic_ref as the shared / global
// atomics, one thread doing the block scan between barriers.  The selection must be identical to the single-threaded
// run (the algorithm is order-independent by construction) and free of unordered conflicting accesses.
// Input: a flat binary file written by tests/test_octree_threads.py (level parameters + candidate triples).
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../orb_slam3_b200/csrc/octree_core.h"

using namespace orbb200;

struct Shared {
  pthread_barrier_t bar;
  int ints[4];
  int total;
};

struct ThreadsBackend {
  int t, T;
  Shared* sh;
  int tid() const { return t; }
  int nthreads() const { return T; }
  void sync() { pthread_barrier_wait(&sh->bar); }
  int atomic_add(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
  void atomic_max64(unsigned long long* p, unsigned long long v) {
    std::atomic_ref<unsigned long long> a(*p);
    unsigned long long old = a.load(std::memory_order_relaxed);
    while (old < v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
  }
  int* shared_int(int i) { return &sh->ints[i]; }
  int exclusive_scan(int* d, int n) {  // in place; every thread gets the total (the CTA version has the same barriers)
    sync();
    if (t == 0) {
      int acc = 0;
      for (int i = 0; i < n; i++) { const int v = d[i]; d[i] = acc; acc += v; }
      sh->total = acc;
    }
    sync();
    const int total = sh->total;
    sync();
    return total;
  }
};

struct Scratch {
  std::vector<int> pt_node, ints;
  std::vector<uint8_t> pt_q;
  std::vector<SortNode> sortbuf;
  std::vector<int> sortwork;
  std::vector<unsigned long long> best;
  std::vector<int> out;
  OctreeScratch s;
  Scratch(int n, const OctreeLevelParams& p) {
    const size_t nc = p.node_cap;
    pt_node.resize(n + 1); ints.resize(nc * (10 + 16 + 5)); pt_q.resize(n + 1); sortbuf.resize(nc); sortwork.resize(6 * nc); best.resize(nc); out.assign(3 * nc, 0);
    s.pt_node = pt_node.data(); s.pt_q = pt_q.data();
    int* q = ints.data();
    for (int b = 0; b < 2; b++) for (int f = 0; f < 5; f++) { s.nd[b][f] = q; q += nc; }
    s.childcnt = q; q += 4 * nc; s.cidx = q; q += 4 * nc; s.eidx = q; q += 4 * nc; s.remap = q; q += 4 * nc;
    s.rank = q; q += nc; s.proc = q; q += nc; s.surv = q; q += nc; s.tmp = q; q += nc; s.expand_pos = q; q += nc;
    s.sortbuf = sortbuf.data(); s.sortwork = sortwork.data(); s.best = best.data();
  }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const int T = atoi(argv[2]);
  int hdr[7];
  if (fread(hdr, 4, 7, f) != 7) return 2;  // n, band_w, band_h, N, w_cell, h_cell, n_cols
  const int n = hdr[0];
  std::vector<int> xys(3 * (size_t)std::max(n, 1));
  if (n && fread(xys.data(), 4, 3 * (size_t)n, f) != 3 * (size_t)n) return 2;
  fclose(f);
  OctreeLevelParams p;
  p.bandW = hdr[1]; p.bandH = hdr[2]; p.N = hdr[3];
  p.nIni = (int)roundf((float)p.bandW / (float)p.bandH);
  p.hX = (float)p.bandW / p.nIni;
  p.wCell = hdr[4]; p.hCell = hdr[5]; p.nCols = hdr[6];
  p.node_cap = p.N + 4 * p.nIni + 16;
  std::vector<Cand> cand(std::max(n, 1));
  for (int i = 0; i < n; i++) { cand[i].xy = (uint32_t)xys[3 * i] | ((uint32_t)xys[3 * i + 1] << 16); cand[i].score = (uint32_t)xys[3 * i + 2]; }
  Scratch A(n, p), B(n, p);
  HostBackend hb;
  const int m_ref = octree_select(hb, cand.data(), n, p, A.s, A.out.data());
  Shared sh;
  memset(sh.ints, 0, sizeof(sh.ints));
  sh.total = 0;
  pthread_barrier_init(&sh.bar, nullptr, T);
  std::vector<int> m_thr(T, -1);
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t]() { ThreadsBackend be{t, T, &sh}; m_thr[t] = octree_select(be, cand.data(), n, p, B.s, B.out.data()); });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&sh.bar);
  int same = 1;
  for (int t = 0; t < T; t++) same &= (m_thr[t] == m_ref);
  for (int i = 0; i < 3 * m_ref && same; i++) same &= (A.out[i] == B.out[i]);
  printf("%d %d %d\n", m_ref, m_thr[0], same);
  return 0;
}
