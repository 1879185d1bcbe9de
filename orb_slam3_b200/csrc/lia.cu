// Optimizer::LocalInertialBA's optimize() (reference src/Optimizer.cc:2383-2958): host
// preparation of the flat graph (what the reference's vertex / edge constructors do once per call), one
// kernel launch that runs the whole LM loop of csrc/lia_core.h in a single CTA, and the host debug hook
// that runs the same source single-threaded.
// The device path is held against the oracle on the B200 by tests/test_lia_gpu.py; the same core runs on the
// host in tests/test_lia_core_host.py.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "lia_core.h"
#include "lia_host.h"
#include "orb_engine.h"

namespace orbb200 {

#define CUDA_TRYI(expr)                                                                \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return ORB_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

constexpr int LIA_THREADS = 512;  // 128 registers per thread: the per-edge Jacobian code spills at 64

struct LiaCtaBackend {
  double* sm;  // LIA_THREADS / 32 partials
  __device__ int tid() const { return threadIdx.x; }
  __device__ int nthreads() const { return blockDim.x; }
  __device__ void sync() { __syncthreads(); }
  __device__ void count(double* p) { atomicAdd(p, 1.0); }  // failure counter: integer-valued, order-independent
  __device__ double sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < nw; w++) t += sm[w];
    __syncthreads();
    return t;
  }
};

__global__ void __launch_bounds__(LIA_THREADS) lia_kernel(const LiaDev D) {
  __shared__ double sm[LIA_THREADS / 32];
  LiaCtaBackend be{sm};
  lia_solve_core(be, D);
}

// One arena of doubles / ints / bytes on the device, re-used across calls
struct Lia {
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void* d_arena = nullptr;
  size_t arena_bytes = 0;
  long long launches = 0;
  explicit Lia(int dev) : device(dev) {}
  ~Lia() {
    cudaFree(d_arena);
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

}  // namespace orbb200

using namespace orbb200;

struct orb_lia { Lia s; explicit orb_lia(int dev) : s(dev) {} };

namespace {

// host-side carving of one byte arena: the same layout is used for the host image and the device copy
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t o = (off + 15) & ~(size_t)15; off = o + bytes; return o; }
};

}  // namespace

extern "C" {

int lia_create(int device, orb_lia** out) {
  if (!out || device < 0) { set_last_error("lia_create: bad argument"); return ORB_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
    return ORB_E_NODEVICE;
  }
  *out = new orb_lia(device);
  return ORB_OK;
}

void lia_destroy(orb_lia* h) { delete h; }

int lia_solve(orb_lia* h, const lia_graph_view* g, double* kf_out, double* mp_out, double* chi2_out,
              uint8_t* depth_pos_out, double* stats) {
  if (!h || !kf_out || (g && g->n_mp && !mp_out)) { set_last_error("lia_solve: bad argument"); return ORB_E_ARG; }
  std::string err;
  int rc = lia_check(g, err);
  if (rc) { set_last_error(err); return rc; }
  LiaHost Hs;
  if ((rc = lia_prepare(g, Hs, err))) { set_last_error(err); return rc; }
  Lia& S = h->s;
  CUDA_TRYI(cudaSetDevice(S.device));
  if (!S.stream) {
    CUDA_TRYI(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
    CUDA_TRYI(cudaEventCreate(&S.ev0));
    CUDA_TRYI(cudaEventCreate(&S.ev1));
  }
  const size_t K = g->n_kf, L = g->n_mp, E = g->n_edges, NI = g->n_inertial, np = Hs.np;
  // ---- layout: inputs first (one H2D of the prefix), then state / system / results
  Carver c;
  const size_t o_fixed = c.take(K), o_imu = c.take(K), o_ip = c.take(4 * K), o_iv = c.take(4 * K), o_ig = c.take(4 * K),
               o_ia = c.take(4 * K), o_ekf = c.take(4 * E), o_emp = c.take(4 * E), o_est = c.take(E), o_eobs = c.take(24 * E),
               o_eis2 = c.take(4 * E), o_lmptr = c.take(4 * (L + 1)), o_lmed = c.take(4 * std::max<size_t>(E, 1)),
               o_k1 = c.take(4 * NI), o_k2 = c.take(4 * NI), o_dR = c.take(36 * NI), o_dV = c.take(12 * NI), o_dP = c.take(12 * NI),
               o_JRg = c.take(36 * NI), o_JVg = c.take(36 * NI), o_JVa = c.take(36 * NI), o_JPg = c.take(36 * NI),
               o_JPa = c.take(36 * NI), o_bias = c.take(24 * NI), o_dT = c.take(4 * NI), o_last = c.take(NI),
               o_info = c.take(648 * std::max<size_t>(NI, 1)), o_infoG = c.take(72 * std::max<size_t>(NI, 1)),
               o_infoA = c.take(72 * std::max<size_t>(NI, 1)), o_pose = c.take(192 * K), o_vel = c.take(24 * K),
               o_bg = c.take(24 * K), o_ba = c.take(24 * K), o_pt = c.take(24 * std::max<size_t>(L, 1)),
               o_kfptr = c.take(4 * (K + 1)), o_kfed = c.take(4 * Hs.kf_edges.size()), o_pptr = c.take(4 * Hs.pair_ptr.size()),
               o_pea = c.take(4 * Hs.pair_ea.size()), o_peb = c.take(4 * Hs.pair_eb.size()), o_free = c.take(4 * std::max<size_t>(Hs.free_kf.size(), 1)),
               o_icol = c.take(4 * Hs.i_color.size());
  const size_t in_bytes = c.off;
  const size_t o_poseb = c.take(192 * K), o_velb = c.take(24 * K), o_bgb = c.take(24 * K), o_bab = c.take(24 * K),
               o_ptb = c.take(24 * std::max<size_t>(L, 1)), o_H = c.take(8 * np * np), o_b = c.take(8 * np),
               o_Hll = c.take(72 * std::max<size_t>(L, 1)), o_bl = c.take(24 * std::max<size_t>(L, 1)),
               o_W = c.take(144 * std::max<size_t>(E, 1)), o_Dinv = c.take(72 * std::max<size_t>(L, 1)), o_S = c.take(8 * np * np),
               o_bs = c.take(8 * np), o_x = c.take(8 * (np + 3 * L)), o_verr = c.take(24 * std::max<size_t>(E, 1)),
               o_ierr = c.take(120 * std::max<size_t>(NI, 1)), o_Dg = c.take(8 * np), o_chi2 = c.take(8 * std::max<size_t>(E, 1)),
               o_dpos = c.take(std::max<size_t>(E, 1)), o_stats = c.take(64),
               o_Hpe = c.take(336 * std::max<size_t>(E, 1)), o_Ye = c.take(144 * std::max<size_t>(E, 1));
  const size_t total = c.off + 16;
  if (total > S.arena_bytes) {
    cudaFree(S.d_arena);
    S.d_arena = nullptr; S.arena_bytes = 0;
    CUDA_TRYI(cudaMalloc(&S.d_arena, total));
    S.arena_bytes = total;
  }
  std::vector<uint8_t> img(in_bytes, 0);
  auto put = [&](size_t off, const void* src, size_t bytes) { if (bytes) memcpy(img.data() + off, src, bytes); };
  put(o_fixed, g->kf_fixed, K); put(o_imu, g->kf_has_imu, K);
  put(o_ip, Hs.ip.data(), 4 * K); put(o_iv, Hs.iv.data(), 4 * K); put(o_ig, Hs.ig.data(), 4 * K); put(o_ia, Hs.ia.data(), 4 * K);
  put(o_ekf, g->e_kf, 4 * E); put(o_emp, g->e_mp, 4 * E); put(o_est, g->e_stereo, E); put(o_eobs, g->e_obs, 24 * E);
  put(o_eis2, g->e_inv_sigma2, 4 * E); put(o_lmptr, Hs.lm_ptr.data(), 4 * (L + 1)); put(o_lmed, Hs.lm_edges.data(), 4 * E);
  put(o_k1, g->i_kf1, 4 * NI); put(o_k2, g->i_kf2, 4 * NI); put(o_dR, g->i_dR, 36 * NI); put(o_dV, g->i_dV, 12 * NI);
  put(o_dP, g->i_dP, 12 * NI); put(o_JRg, g->i_JRg, 36 * NI); put(o_JVg, g->i_JVg, 36 * NI); put(o_JVa, g->i_JVa, 36 * NI);
  put(o_JPg, g->i_JPg, 36 * NI); put(o_JPa, g->i_JPa, 36 * NI); put(o_bias, g->i_bias, 24 * NI); put(o_dT, g->i_dT, 4 * NI);
  put(o_last, g->i_last, NI); put(o_info, Hs.info.data(), 648 * NI); put(o_infoG, Hs.infoG.data(), 72 * NI);
  put(o_infoA, Hs.infoA.data(), 72 * NI); put(o_pose, Hs.pose.data(), 192 * K); put(o_vel, g->kf_vel, 24 * K);
  put(o_bg, g->kf_bg, 24 * K); put(o_ba, g->kf_ba, 24 * K); put(o_pt, g->mp_pos, 24 * L);
  put(o_kfptr, Hs.kf_ptr.data(), 4 * (K + 1)); put(o_kfed, Hs.kf_edges.data(), 4 * Hs.kf_edges.size());
  put(o_pptr, Hs.pair_ptr.data(), 4 * Hs.pair_ptr.size()); put(o_pea, Hs.pair_ea.data(), 4 * Hs.pair_ea.size());
  put(o_peb, Hs.pair_eb.data(), 4 * Hs.pair_eb.size()); put(o_free, Hs.free_kf.data(), 4 * Hs.free_kf.size());
  put(o_icol, Hs.i_color.data(), 4 * Hs.i_color.size());
  uint8_t* base = (uint8_t*)S.d_arena;
  cudaStream_t s = S.stream;
  CUDA_TRYI(cudaMemcpyAsync(base, img.data(), in_bytes, cudaMemcpyHostToDevice, s));
  CUDA_TRYI(cudaMemsetAsync(base + o_stats, 0, 64, s));
  CUDA_TRYI(cudaMemsetAsync(base + o_x, 0, 8 * (np + 3 * L), s));
  LiaDev D;
  lia_fill_scalars(g, Hs, D);
#define AT(T, off) reinterpret_cast<T*>(base + (off))
  D.kf_fixed = AT(uint8_t, o_fixed); D.kf_has_imu = AT(uint8_t, o_imu);
  D.ip = AT(int, o_ip); D.iv = AT(int, o_iv); D.ig = AT(int, o_ig); D.ia = AT(int, o_ia);
  D.e_kf = AT(int, o_ekf); D.e_mp = AT(int, o_emp); D.e_stereo = AT(uint8_t, o_est); D.e_obs = AT(double, o_eobs);
  D.e_is2 = AT(float, o_eis2); D.lm_ptr = AT(int, o_lmptr); D.lm_edges = AT(int, o_lmed);
  D.i_kf1 = AT(int, o_k1); D.i_kf2 = AT(int, o_k2); D.i_dR = AT(float, o_dR); D.i_dV = AT(float, o_dV); D.i_dP = AT(float, o_dP);
  D.i_JRg = AT(float, o_JRg); D.i_JVg = AT(float, o_JVg); D.i_JVa = AT(float, o_JVa); D.i_JPg = AT(float, o_JPg);
  D.i_JPa = AT(float, o_JPa); D.i_bias = AT(float, o_bias); D.i_dT = AT(float, o_dT); D.i_last = AT(uint8_t, o_last);
  D.info = AT(double, o_info); D.infoG = AT(double, o_infoG); D.infoA = AT(double, o_infoA);
  D.pose = AT(double, o_pose); D.pose_bak = AT(double, o_poseb); D.vel = AT(double, o_vel); D.bg = AT(double, o_bg);
  D.ba = AT(double, o_ba); D.pt = AT(double, o_pt); D.vel_bak = AT(double, o_velb); D.bg_bak = AT(double, o_bgb);
  D.ba_bak = AT(double, o_bab); D.pt_bak = AT(double, o_ptb);
  D.H = AT(double, o_H); D.b = AT(double, o_b); D.Hll = AT(double, o_Hll); D.bl = AT(double, o_bl); D.W = AT(double, o_W);
  D.Dinv = AT(double, o_Dinv); D.S = AT(double, o_S); D.bs = AT(double, o_bs); D.x = AT(double, o_x);
  D.verr = AT(double, o_verr); D.ierr = AT(double, o_ierr); D.Dg = AT(double, o_Dg);
  D.chi2_out = AT(double, o_chi2); D.depth_pos_out = AT(uint8_t, o_dpos); D.stats = AT(double, o_stats);
  D.kf_ptr = AT(int, o_kfptr); D.kf_edges = AT(int, o_kfed); D.pair_ptr = AT(int, o_pptr); D.pair_ea = AT(int, o_pea);
  D.pair_eb = AT(int, o_peb); D.free_kf = AT(int, o_free); D.i_color = AT(int, o_icol);
  D.Hpe = AT(double, o_Hpe); D.Ye = AT(double, o_Ye);
#undef AT
  CUDA_TRYI(cudaEventRecord(S.ev0, s));
  lia_kernel<<<1, LIA_THREADS, 0, s>>>(D);
  CUDA_TRYI(cudaEventRecord(S.ev1, s));
  S.launches += 1;
  CUDA_TRYI(cudaGetLastError());
  std::vector<double> pose(24 * K), vel(3 * K), bg(3 * K), ba(3 * K), pt(3 * std::max<size_t>(L, 1)), chi(std::max<size_t>(E, 1)), st(8);
  std::vector<uint8_t> dp(std::max<size_t>(E, 1));
  CUDA_TRYI(cudaMemcpyAsync(pose.data(), base + o_pose, 192 * K, cudaMemcpyDeviceToHost, s));
  CUDA_TRYI(cudaMemcpyAsync(vel.data(), base + o_vel, 24 * K, cudaMemcpyDeviceToHost, s));
  CUDA_TRYI(cudaMemcpyAsync(bg.data(), base + o_bg, 24 * K, cudaMemcpyDeviceToHost, s));
  CUDA_TRYI(cudaMemcpyAsync(ba.data(), base + o_ba, 24 * K, cudaMemcpyDeviceToHost, s));
  if (L) CUDA_TRYI(cudaMemcpyAsync(pt.data(), base + o_pt, 24 * L, cudaMemcpyDeviceToHost, s));
  if (E) {
    CUDA_TRYI(cudaMemcpyAsync(chi.data(), base + o_chi2, 8 * E, cudaMemcpyDeviceToHost, s));
    CUDA_TRYI(cudaMemcpyAsync(dp.data(), base + o_dpos, E, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRYI(cudaMemcpyAsync(st.data(), base + o_stats, 64, cudaMemcpyDeviceToHost, s));
  CUDA_TRYI(cudaStreamSynchronize(s));
  lia_write_out(g, pose.data(), vel.data(), bg.data(), ba.data(), pt.data(), kf_out, mp_out);
  if (chi2_out && E) memcpy(chi2_out, chi.data(), 8 * E);
  if (depth_pos_out && E) memcpy(depth_pos_out, dp.data(), E);
  if (stats) { for (int i = 0; i < 6; i++) stats[i] = st[i]; stats[6] = stats[7] = 0; }
  return (int)st[0];
}

long long lia_kernel_launches(const orb_lia* h) { return h ? h->s.launches : 0; }

float lia_last_ms(orb_lia* h) {
  if (!h || !h->s.ev1 || cudaEventSynchronize(h->s.ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->s.ev0, h->s.ev1) != cudaSuccess) return 0.f;
  return ms;
}

int lia_debug_host(const lia_graph_view* g, double* kf_out, double* mp_out, double* chi2_out, uint8_t* depth_pos_out,
                   double* stats) {
  if (!kf_out || (g && g->n_mp && !mp_out)) return ORB_E_ARG;
  std::string err;
  int rc = lia_check(g, err);
  if (rc) { set_last_error(err); return rc; }
  LiaHost Hs;
  if ((rc = lia_prepare(g, Hs, err))) { set_last_error(err); return rc; }
  const size_t E = g->n_edges;
  LiaHostBuffers B(g, Hs);
  LiaDev& D = B.D;
  LiaHostBackend be;
  lia_solve_core(be, D);
  lia_write_out(g, B.pose.data(), B.vel.data(), B.bg.data(), B.ba.data(), B.pt.data(), kf_out, mp_out);
  if (chi2_out && E) memcpy(chi2_out, B.chi.data(), 8 * E);
  if (depth_pos_out && E) memcpy(depth_pos_out, B.dp.data(), E);
  if (stats) { for (int i = 0; i < 6; i++) stats[i] = B.st[i]; stats[6] = stats[7] = 0; }
  return (int)B.st[0];
}

}  // extern "C"
