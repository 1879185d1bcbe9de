// Frame::ComputeStereoMatches (reference src/Frame.cc:811-981) on the
// device-resident results of two extractor handles: the keypoints / descriptors
// and the un-blurred pyramids of the left and right image never leave HBM (the
// reference reads mvImagePyramid on the host, which is what forces the pyramid
// D2H mirror of orb_pyramid()).
//
//   stereo_rows_kernel    one CTA per frame: row table of the right keypoints
//                         (Frame.cc:820-838) as CSR built with shared-memory atomics.
//                         The order inside a row is irrelevant here: "first strictly
//                         smaller distance wins" over a list in ascending iR (:873-895)
//                         is the minimum of (distance, iR), which the match kernel
//                         reduces directly.
//   stereo_match_kernel   one warp per left keypoint: Hamming search over the row's
//                         candidates, 11x11 L1 window slid over +-5 px at the keypoint's
//                         pyramid level, parabola sub-pixel fit (:898-964).
//   stereo_reject_kernel  one CTA per frame: median of the window distances by a
//                         two-level histogram select, 1.5*1.4*median gate (:968-982).
//
// Float arithmetic follows the reference expression by expression (IEEE single,
// no FMA: the library is built with -fmad=false).
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/orb_b200.h"
#include "orb_engine.h"

namespace orbb200 {

#define CUDA_TRYS(expr)                                                                \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return ORB_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

constexpr int ST_TH_HIGH = 100, ST_TH_LOW = 50;       // ORBmatcher.cc:35-36
constexpr int ST_W = 5, ST_L = 5;                      // Frame.cc:907, :913
constexpr int ST_MAX_ROWS_PER_KP = 20;                 // ceil(y+r)-floor(y-r)+1 with r = 2*scale <= 2*1.2^7... (checked on the host)
constexpr int ST_WARPS = 8;

struct StereoParams {
  const uint8_t *pyr_l, *pyr_r;
  size_t pyr_stride;
  const orb_keypoint *kl, *kr;
  const uint8_t *dl, *dr;
  const int *nl, *nr;
  int cap;                 // keypoint slots per frame in the extractor outputs
  int rows, nlevels;
  int lw[16], lpitch[16];
  unsigned long long loff[16];
  float scale[16], inv_scale[16];
  float bf, maxD;
  int* row_off;            // [batch][rows + 1]
  int* row_items;          // [batch][items_cap]
  int items_cap;
  float *u_right, *depth;  // [batch][cap]
  int* sad;                // [batch][cap], -1 = not in vDistIdx
  int* kept;               // [batch]
};

__device__ __forceinline__ int st_popc256(const uint4 a0, const uint4 a1, const uint8_t* __restrict__ b) {
  const uint4* pb = reinterpret_cast<const uint4*>(b);
  const uint4 b0 = pb[0], b1 = pb[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ void st_row_span(const orb_keypoint& kp, const float* scale, int rows, int& minr, int& maxr) {
  const float r = __fmul_rn(2.0f, scale[kp.octave]);          // :831
  maxr = min((int)ceilf(__fadd_rn(kp.y, r)), rows - 1);       // :832
  minr = max((int)floorf(__fsub_rn(kp.y, r)), 0);             // :833
}

__global__ void __launch_bounds__(256) stereo_rows_kernel(const __grid_constant__ StereoParams P) {
  extern __shared__ int s_cnt[];  // [rows + 1]
  const int f = blockIdx.x, rows = P.rows;
  const int nr = P.nr[f];
  const orb_keypoint* kr = P.kr + (size_t)f * P.cap;
  int* row_off = P.row_off + (size_t)f * (rows + 1);
  int* items = P.row_items + (size_t)f * P.items_cap;
  for (int y = threadIdx.x; y <= rows; y += 256) s_cnt[y] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nr; i += 256) {
    int minr, maxr;
    st_row_span(kr[i], P.scale, rows, minr, maxr);
    for (int y = minr; y <= maxr; y++) atomicAdd(&s_cnt[y], 1);
  }
  __syncthreads();
  // exclusive scan of the row counts by warp 0, 32 rows per trip
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 0; base <= rows; base += 32) {
      const int y = base + lane;
      const int v = y <= rows ? s_cnt[y] : 0;
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (y <= rows) { s_cnt[y] = carry + incl - v; row_off[y] = carry + incl - v; }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nr; i += 256) {
    int minr, maxr;
    st_row_span(kr[i], P.scale, rows, minr, maxr);
    for (int y = minr; y <= maxr; y++) {
      const int pos = atomicAdd(&s_cnt[y], 1);
      if (pos < P.items_cap) items[pos] = i;
    }
  }
}

__global__ void __launch_bounds__(32 * ST_WARPS) stereo_match_kernel(const __grid_constant__ StereoParams P) {
  __shared__ int s_sad[ST_WARPS][2 * ST_L + 1];
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iL = blockIdx.x * ST_WARPS + warp;
  const int nl = P.nl[f];
  if (iL >= nl) return;
  const size_t fo = (size_t)f * P.cap;
  float* out_u = P.u_right + fo;
  float* out_d = P.depth + fo;
  int* out_s = P.sad + fo;
  if (lane == 0) { out_u[iL] = -1.0f; out_d[iL] = -1.0f; out_s[iL] = -1; }  // :813-814
  const orb_keypoint kpL = P.kl[fo + iL];
  const int levelL = kpL.octave;
  const float uL = kpL.x, vL = kpL.y;
  const int row = min(max((int)vL, 0), P.rows - 1);                         // vRowIndices[vL] :856
  const int* row_off = P.row_off + (size_t)f * (P.rows + 1);
  const int c0 = row_off[row], c1 = min(row_off[row + 1], P.items_cap);
  if (c0 >= c1) return;
  const float minU = __fsub_rn(uL, P.maxD), maxU = uL;                     // minD = 0 (:842)
  if (maxU < 0) return;
  const orb_keypoint* kr = P.kr + fo;
  const uint8_t* dr = P.dr + fo * 32;
  const int* items = P.row_items + (size_t)f * P.items_cap;
  const uint4* pl = reinterpret_cast<const uint4*>(P.dl + (fo + iL) * 32);
  const uint4 a0 = pl[0], a1 = pl[1];
  // best = min over (distance, iR): the reference keeps the first strictly smaller distance of a
  // list in ascending iR (:873-895)
  unsigned best = ((unsigned)ST_TH_HIGH << 16) | 0xffffu;
  for (int c = c0 + lane; c < c1; c += 32) {
    const int iR = items[c];
    const orb_keypoint k = kr[iR];
    if (k.octave < levelL - 1 || k.octave > levelL + 1) continue;
    if (!(k.x >= minU && k.x <= maxU)) continue;
    const int dist = st_popc256(a0, a1, dr + (size_t)iR * 32);
    if (dist < ST_TH_HIGH) best = min(best, ((unsigned)dist << 16) | (unsigned)iR);
  }
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  const int bestDist = (int)(best >> 16);
  if (!(bestDist < (ST_TH_HIGH + ST_TH_LOW) / 2)) return;                   // thOrbDist (:816, :898)
  const int bestIdxR = (int)(best & 0xffffu);

  // coordinates at the keypoint's pyramid level (:901-905)
  const float uR0 = kr[bestIdxR].x;
  const float sf = P.inv_scale[levelL];
  const float scaleduL = roundf(__fmul_rn(uL, sf));
  const float scaledvL = roundf(__fmul_rn(vL, sf));
  const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
  const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)ST_L), (float)ST_W);
  const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)ST_L), (float)ST_W), 1.0f);
  if (iniu < 0 || endu >= (float)P.lw[levelL]) return;                      // :918-919
  const int pitch = P.lpitch[levelL];
  const uint8_t* IL = P.pyr_l + (size_t)f * P.pyr_stride + P.loff[levelL];
  const uint8_t* IR = P.pyr_r + (size_t)f * P.pyr_stride + P.loff[levelL];
  const int y0 = (int)scaledvL - ST_W, xl0 = (int)scaleduL - ST_W, xr0 = (int)scaleduR0 - ST_L - ST_W;
  if (lane < 2 * ST_L + 1) s_sad[warp][lane] = 0;
  __syncwarp();
  // 11 window rows x 11 shifts = 121 row sums of 11 absolute differences, spread over the lanes
  for (int idx = lane; idx < (2 * ST_W + 1) * (2 * ST_L + 1); idx += 32) {
    const int r = idx / (2 * ST_L + 1), k = idx - r * (2 * ST_L + 1);
    const uint8_t* a = IL + (size_t)(y0 + r) * pitch + xl0;
    const uint8_t* b = IR + (size_t)(y0 + r) * pitch + xr0 + k;
    int s = 0;
#pragma unroll
    for (int c = 0; c < 2 * ST_W + 1; c++) s += abs((int)a[c] - (int)b[c]);
    atomicAdd(&s_sad[warp][k], s);
  }
  __syncwarp();
  if (lane != 0) return;
  int bestSad = INT_MAX, bestinc = 0;
  for (int k = 0; k < 2 * ST_L + 1; k++) {                                  // :921-933
    const float dist = (float)s_sad[warp][k];
    if (dist < (float)bestSad) { bestSad = (int)dist; bestinc = k - ST_L; }
  }
  if (bestinc == -ST_L || bestinc == ST_L) return;                          // :935-936
  const float dist1 = (float)s_sad[warp][ST_L + bestinc - 1];
  const float dist2 = (float)s_sad[warp][ST_L + bestinc];
  const float dist3 = (float)s_sad[warp][ST_L + bestinc + 1];
  const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3),
                                 __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
  if (deltaR < -1 || deltaR > 1) return;
  float bestuR = __fmul_rn(P.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));  // :950
  float disparity = __fsub_rn(uL, bestuR);
  if (disparity >= 0 && disparity < P.maxD) {
    if (disparity <= 0) { disparity = 0.01; bestuR = (float)((double)uL - 0.01); }  // :956-960
    out_d[iL] = __fdiv_rn(P.bf, disparity);
    out_u[iL] = bestuR;
    out_s[iL] = bestSad;
  }
}

__global__ void __launch_bounds__(256) stereo_reject_kernel(const __grid_constant__ StereoParams P) {
  __shared__ int s_hi[256], s_lo[128];
  __shared__ int s_m, s_bin, s_rem, s_median, s_kept;
  const int f = blockIdx.x;
  const int nl = P.nl[f];
  const size_t fo = (size_t)f * P.cap;
  const int* sad = P.sad + fo;
  s_hi[threadIdx.x] = 0;
  if (threadIdx.x < 128) s_lo[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_m = 0; s_kept = 0; }
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int s = sad[i];
    if (s >= 0) { atomicAdd(&s_hi[min(s >> 7, 255)], 1); mine++; }
  }
  if (mine) atomicAdd(&s_m, mine);
  __syncthreads();
  const int m = s_m;
  if (m == 0) { if (threadIdx.x == 0) P.kept[f] = 0; return; }  // the reference reads vDistIdx[0] of an empty list here
  if (threadIdx.x == 0) {
    int k = m / 2, b = 0;                                       // vDistIdx[size/2] of the sorted list (:969)
    while (k >= s_hi[b]) { k -= s_hi[b]; b++; }
    s_bin = b; s_rem = k;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int s = sad[i];
    if (s >= 0 && min(s >> 7, 255) == s_bin) atomicAdd(&s_lo[s & 127], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = s_rem, b = 0;
    while (k >= s_lo[b]) { k -= s_lo[b]; b++; }
    s_median = (s_bin << 7) | b;
  }
  __syncthreads();
  const float median = (float)s_median;
  const float thDist = 1.5f * 1.4f * median;                    // :970
  int kept = 0;
  for (int i = threadIdx.x; i < nl; i += 256) {
    const int s = sad[i];
    if (s < 0) continue;
    if ((float)s < thDist) kept++;
    else { P.u_right[fo + i] = -1.0f; P.depth[fo + i] = -1.0f; }  // :972-981
  }
  if (kept) atomicAdd(&s_kept, kept);
  __syncthreads();
  if (threadIdx.x == 0) P.kept[f] = s_kept;
}

struct Stereo {
  int device;
  int cap_batch = 0, cap_kp = 0, cap_rows = 0;
  int *d_row_off = nullptr, *d_row_items = nullptr, *d_sad = nullptr, *d_kept = nullptr;
  float *d_u = nullptr, *d_depth = nullptr;
  int* h_kept = nullptr;
  cudaEvent_t ev = nullptr, ev0 = nullptr, ev1 = nullptr;
  long long launches = 0;

  explicit Stereo(int dev) : device(dev) {}
  ~Stereo() {
    release();
    if (ev) cudaEventDestroy(ev);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
  void release() {
    cudaFree(d_row_off); cudaFree(d_row_items); cudaFree(d_sad); cudaFree(d_kept); cudaFree(d_u); cudaFree(d_depth);
    if (h_kept) cudaFreeHost(h_kept);
    d_row_off = d_row_items = d_sad = d_kept = nullptr; d_u = d_depth = nullptr; h_kept = nullptr;
    cap_batch = cap_kp = cap_rows = 0;
  }
  int ensure(int batch, int kp, int rows) {
    if (!ev) {
      CUDA_TRYS(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      CUDA_TRYS(cudaEventCreate(&ev0));
      CUDA_TRYS(cudaEventCreate(&ev1));
    }
    // the per-frame stride of the outputs is the extractor's keypoint slot count
    if (batch <= cap_batch && kp == cap_kp && rows <= cap_rows) return 0;
    batch = std::max(batch, cap_batch); rows = std::max(rows, cap_rows);
    release();
    CUDA_TRYS(cudaMalloc(&d_row_off, sizeof(int) * (size_t)batch * (rows + 1)));
    CUDA_TRYS(cudaMalloc(&d_row_items, sizeof(int) * (size_t)batch * kp * ST_MAX_ROWS_PER_KP));
    CUDA_TRYS(cudaMalloc(&d_sad, sizeof(int) * (size_t)batch * kp));
    CUDA_TRYS(cudaMalloc(&d_kept, sizeof(int) * batch));
    CUDA_TRYS(cudaMalloc(&d_u, sizeof(float) * (size_t)batch * kp));
    CUDA_TRYS(cudaMalloc(&d_depth, sizeof(float) * (size_t)batch * kp));
    CUDA_TRYS(cudaMallocHost(&h_kept, sizeof(int) * batch));
    cap_batch = batch; cap_kp = kp; cap_rows = rows;
    return 0;
  }

  int run(Engine& L, Engine& R, int batch, float bf, float b, float* u_right, float* depth, int cap, int* kept,
          int on_device, cudaStream_t user) {
    if (!L.initialized || !R.initialized || batch <= 0 || batch > L.last_batch || batch > R.last_batch) {
      set_last_error("stereo_match: both extractors must hold the results of an extract of >= batch frames");
      return ORB_E_ARG;
    }
    if (L.device != device || R.device != device || L.cap_rows != R.cap_rows || L.cap_cols != R.cap_cols ||
        L.nlevels != R.nlevels || L.out_cap != R.out_cap || L.scale_factor != R.scale_factor) {
      set_last_error("stereo_match: left and right extractor differ in device, image size or parameters");
      return ORB_E_ARG;
    }
    if (!(b > 0) || !(bf > 0)) { set_last_error("stereo_match: mb and mbf must be positive"); return ORB_E_ARG; }
    if (!on_device && (!u_right || !depth || cap <= 0)) { set_last_error("stereo_match: bad output buffers"); return ORB_E_ARG; }
    if (L.out_cap > 0xffff) { set_last_error("stereo_match: more than 65535 keypoint slots per frame"); return ORB_E_CAPACITY; }
    // rows a right keypoint can be filed under (:831-836): ceil(y+r) - floor(y-r) + 1 <= 2r + 3
    if (2.0f * 2.0f * L.scale[L.nlevels - 1] + 3.0f > (float)ST_MAX_ROWS_PER_KP) {
      set_last_error("stereo_match: scale pyramid too deep for the row table");
      return ORB_E_CAPACITY;
    }
    CUDA_TRYS(cudaSetDevice(device));
    const int rows = L.levels[0].h;
    int rc = ensure(std::max(batch, L.cap_batch), L.out_cap, rows);
    if (rc) return rc;
    cudaStream_t s = user ? user : (L.last_stream ? L.last_stream : L.stream);
    // both extractions must have finished on their own streams
    for (Engine* e : {&L, &R}) {
      cudaStream_t es = e->last_stream ? e->last_stream : e->stream;
      if (es != s) {
        CUDA_TRYS(cudaEventRecord(ev, es));
        CUDA_TRYS(cudaStreamWaitEvent(s, ev, 0));
      }
    }
    StereoParams P;
    P.pyr_l = L.d_pyr; P.pyr_r = R.d_pyr; P.pyr_stride = L.pyr_frame_bytes;
    P.kl = L.d_kps; P.kr = R.d_kps; P.dl = L.d_desc; P.dr = R.d_desc; P.nl = L.d_n; P.nr = R.d_n;
    P.cap = L.out_cap; P.rows = rows; P.nlevels = L.nlevels;
    for (int l = 0; l < L.nlevels; l++) {
      P.lw[l] = L.levels[l].w; P.lpitch[l] = L.levels[l].pitch; P.loff[l] = L.levels[l].img_off;
      P.scale[l] = L.scale[l]; P.inv_scale[l] = L.inv_scale[l];
    }
    P.bf = bf; P.maxD = bf / b;                                  // minZ = mb, maxD = mbf/minZ (:841-843)
    P.row_off = d_row_off; P.row_items = d_row_items; P.items_cap = cap_kp * ST_MAX_ROWS_PER_KP;
    P.u_right = d_u; P.depth = d_depth; P.sad = d_sad; P.kept = d_kept;
    CUDA_TRYS(cudaEventRecord(ev0, s));
    stereo_rows_kernel<<<batch, 256, sizeof(int) * (rows + 1), s>>>(P);
    stereo_match_kernel<<<dim3((L.out_cap + ST_WARPS - 1) / ST_WARPS, batch), 32 * ST_WARPS, 0, s>>>(P);
    stereo_reject_kernel<<<batch, 256, 0, s>>>(P);
    CUDA_TRYS(cudaEventRecord(ev1, s));
    launches += 3;
    CUDA_TRYS(cudaGetLastError());
    if (on_device) return batch;
    const size_t w = sizeof(float) * (size_t)std::min(cap, L.out_cap);
    CUDA_TRYS(cudaMemcpy2DAsync(u_right, sizeof(float) * (size_t)cap, d_u, sizeof(float) * (size_t)L.out_cap, w, batch,
                                cudaMemcpyDeviceToHost, s));
    CUDA_TRYS(cudaMemcpy2DAsync(depth, sizeof(float) * (size_t)cap, d_depth, sizeof(float) * (size_t)L.out_cap, w, batch,
                                cudaMemcpyDeviceToHost, s));
    CUDA_TRYS(cudaMemcpyAsync(h_kept, d_kept, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
    CUDA_TRYS(cudaStreamSynchronize(s));
    if (kept) for (int i = 0; i < batch; i++) kept[i] = h_kept[i];
    return batch;
  }
};

}  // namespace orbb200

using orbb200::Stereo;

struct orb_stereo { Stereo s; explicit orb_stereo(int dev) : s(dev) {} };

extern "C" {

int stereo_create(int device, orb_stereo** out) {
  if (!out || device < 0) { orbb200::set_last_error("stereo_create: bad argument"); return ORB_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    orbb200::set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
    return ORB_E_NODEVICE;
  }
  *out = new orb_stereo(device);
  return ORB_OK;
}

void stereo_destroy(orb_stereo* h) { delete h; }

int stereo_match(orb_stereo* h, orb_extractor* left, orb_extractor* right, float bf, float b, float* u_right,
                 float* depth, int cap) {
  if (!h || !left || !right) return ORB_E_ARG;
  int kept = 0;
  const int rc = h->s.run(left->e, right->e, 1, bf, b, u_right, depth, cap, &kept, 0, nullptr);
  return rc < 0 ? rc : kept;
}

int stereo_match_batch(orb_stereo* h, orb_extractor* left, orb_extractor* right, int batch, float bf, float b,
                       float* u_right, float* depth, int cap, int* kept, int on_device, void* cuda_stream) {
  if (!h || !left || !right) return ORB_E_ARG;
  return h->s.run(left->e, right->e, batch, bf, b, u_right, depth, cap, kept, on_device, (cudaStream_t)cuda_stream);
}

int stereo_device_results(orb_stereo* h, const float** d_u_right, const float** d_depth, const int** d_kept,
                          int* stride) {
  if (!h || !h->s.d_u) return ORB_E_ARG;
  if (d_u_right) *d_u_right = h->s.d_u;
  if (d_depth) *d_depth = h->s.d_depth;
  if (d_kept) *d_kept = h->s.d_kept;
  if (stride) *stride = h->s.cap_kp;
  return ORB_OK;
}

long long stereo_kernel_launches(const orb_stereo* h) { return h ? h->s.launches : 0; }
float stereo_last_ms(orb_stereo* h) {
  if (!h || !h->s.ev1 || cudaEventSynchronize(h->s.ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->s.ev0, h->s.ev1) != cudaSuccess) return 0.f;
  return ms;
}

}  // extern "C"
