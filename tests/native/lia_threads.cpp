// CPU stand-in for "many threads run lia_kernel": the product's LocalInertialBA source (csrc/lia_core.h) executed by T
// real threads with a pthread barrier as __syncthreads(), built with -fsanitize=thread.  ThreadSanitizer then checks
// what a single-threaded host run cannot: that every pair of conflicting accesses in the kernel body is separated by
// a barrier -- the accumulation into H / b / S is plain read-modify-write by one owner per destination (per-edge
// terms gathered in a fixed order, inertial edges colour by colour), the only atomic is a failure counter.  The graph comes
// from a flat binary file written by tests/test_lia_threads.py; results are compared with the
// single-threaded run of the same source.
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -I<repo>/include tests/native/lia_threads.cpp
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../orb_slam3_b200/csrc/lia_host.h"

using namespace orbb200;

struct Shared {
  pthread_barrier_t bar;
  std::vector<double> partial;
};

struct ThreadsBackend {
  int t, T;
  Shared* sh;
  int tid() const { return t; }
  int nthreads() const { return T; }
  void sync() { pthread_barrier_wait(&sh->bar); }
  void count(double* p) {  // atomicAdd(p, 1.0): the only atomic left (integer-valued failure counter)
    std::atomic_ref<double> a(*p);
    double old = a.load(std::memory_order_relaxed);
    while (!a.compare_exchange_weak(old, old + 1.0, std::memory_order_relaxed)) {}
  }
  double sum(double v) {
    sh->partial[t] = v;
    sync();
    double s = 0;
    for (int i = 0; i < T; i++) s += sh->partial[i];
    sync();
    return s;
  }
};

static std::vector<std::vector<uint8_t>> g_blobs;
static const void* next_blob(FILE* f, size_t* n_out = nullptr) {
  uint64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(2); }
  g_blobs.emplace_back(n ? n : 1);
  if (n && fread(g_blobs.back().data(), 1, n, f) != n) { fprintf(stderr, "short file\n"); exit(2); }
  if (n_out) *n_out = n;
  return g_blobs.back().data();
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: lia_threads <graph.bin> <threads>\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const int T = atoi(argv[2]);
  lia_graph_view g;
  memset(&g, 0, sizeof(g));
  int32_t hdr[5];
  if (fread(hdr, 4, 5, f) != 5) return 2;
  g.n_kf = hdr[0]; g.n_mp = hdr[1]; g.n_edges = hdr[2]; g.n_inertial = hdr[3]; g.iterations = hdr[4];
  double dh[16];
  if (fread(dh, 8, 16, f) != 16) return 2;  // Rcb 9, tcb 3, tbc 3, lambda
  for (int i = 0; i < 9; i++) g.Rcb[i] = dh[i];
  for (int i = 0; i < 3; i++) { g.tcb[i] = dh[9 + i]; g.tbc[i] = dh[12 + i]; }
  g.lambda_init = dh[15];
  float cam[5];
  if (fread(cam, 4, 5, f) != 5) return 2;
  g.fx = cam[0]; g.fy = cam[1]; g.cx = cam[2]; g.cy = cam[3]; g.bf = cam[4];
  g.kf_Rwb = (const double*)next_blob(f); g.kf_twb = (const double*)next_blob(f); g.kf_Rcw = (const double*)next_blob(f);
  g.kf_tcw = (const double*)next_blob(f); g.kf_fixed = (const uint8_t*)next_blob(f); g.kf_has_imu = (const uint8_t*)next_blob(f);
  g.kf_vel = (const double*)next_blob(f); g.kf_bg = (const double*)next_blob(f); g.kf_ba = (const double*)next_blob(f);
  g.mp_pos = (const double*)next_blob(f); g.e_kf = (const int32_t*)next_blob(f); g.e_mp = (const int32_t*)next_blob(f);
  g.e_stereo = (const uint8_t*)next_blob(f); g.e_obs = (const double*)next_blob(f); g.e_inv_sigma2 = (const float*)next_blob(f);
  g.i_kf1 = (const int32_t*)next_blob(f); g.i_kf2 = (const int32_t*)next_blob(f); g.i_dR = (const float*)next_blob(f);
  g.i_dV = (const float*)next_blob(f); g.i_dP = (const float*)next_blob(f); g.i_JRg = (const float*)next_blob(f);
  g.i_JVg = (const float*)next_blob(f); g.i_JVa = (const float*)next_blob(f); g.i_JPg = (const float*)next_blob(f);
  g.i_JPa = (const float*)next_blob(f); g.i_bias = (const float*)next_blob(f); g.i_dT = (const float*)next_blob(f);
  g.i_C = (const float*)next_blob(f); g.i_last = (const uint8_t*)next_blob(f);
  fclose(f);
  std::string err;
  LiaHost Hs;
  if (lia_check(&g, err) || lia_prepare(&g, Hs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
  // reference: one thread
  LiaHostBuffers R(&g, Hs);
  { LiaHostBackend be; lia_solve_core(be, R.D); }
  // T threads on a second set of buffers
  LiaHostBuffers M(&g, Hs);
  Shared sh;
  sh.partial.assign(T, 0.0);
  pthread_barrier_init(&sh.bar, nullptr, T);
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t]() { ThreadsBackend be{t, T, &sh}; lia_solve_core(be, M.D); });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&sh.bar);
  double dpose = 0, dpt = 0, dchi = 0;
  for (size_t i = 0; i < R.pose.size(); i++) dpose = std::max(dpose, fabs(R.pose[i] - M.pose[i]));
  for (size_t i = 0; i < R.vel.size(); i++) dpose = std::max(dpose, std::max(fabs(R.vel[i] - M.vel[i]), std::max(fabs(R.bg[i] - M.bg[i]), fabs(R.ba[i] - M.ba[i]))));
  for (size_t i = 0; i < R.pt.size(); i++) dpt = std::max(dpt, fabs(R.pt[i] - M.pt[i]));
  for (int e = 0; e < g.n_edges; e++) dchi = std::max(dchi, fabs(R.chi[e] - M.chi[e]));
  printf("%d %d %d %d %.17g %.17g %.3e %.3e %.3e\n", (int)R.st[0], (int)M.st[0], (int)R.st[1], (int)M.st[1], R.st[3], M.st[3], dpose, dpt, dchi);
  return 0;
}
