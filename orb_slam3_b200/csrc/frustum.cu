// Frame::isInFrustum for all local map points of a frame (reference src/Frame.cc:512-570 called
// from Tracking::SearchLocalPoints, src/Tracking.cc:3367-3390): one thread per
// map point, SoA outputs in the layout SearchByProjection(Frame&, vector<MapPoint*>&) reads
// (orb_mappoint_view), so the projection matcher can consume them without leaving the device.
// HBM-bound streaming kernel: 32 B in, 25 B out per point.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "../../include/orb_b200.h"
#include "frustum_core.h"
#include "orb_engine.h"

namespace orbb200 {

#define CUDA_TRYF(expr)                                                                \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return ORB_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

struct FrustumDev {
  const float *world_pos, *normal, *min_dist, *max_dist;
  uint8_t *in_view, *full;
  float *proj_x, *proj_y, *proj_xr, *view_cos, *depth;
  int* level;
  int* count;
  int n;
};

__global__ void __launch_bounds__(256) frustum_kernel(const __grid_constant__ FrustumFrame F, const FrustumDev D,
                                                      float cos_limit) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  int in = 0;
  if (i < D.n) {
    const float P[3] = {D.world_pos[3 * i], D.world_pos[3 * i + 1], D.world_pos[3 * i + 2]};
    const float Pn[3] = {D.normal[3 * i], D.normal[3 * i + 1], D.normal[3 * i + 2]};
    const FrustumPoint o = frustum_point(F, P, Pn, D.min_dist[i], D.max_dist[i], cos_limit);
    D.in_view[i] = o.in_view; D.full[i] = o.full;
    D.proj_x[i] = o.proj_x; D.proj_y[i] = o.proj_y;
    if (o.full) { D.proj_xr[i] = o.proj_xr; D.view_cos[i] = o.view_cos; D.depth[i] = o.depth; D.level[i] = o.level; }
    in = o.in_view;
  }
  if (D.count) {  // device-resident callers read the count on the device; the host path counts while it scatters
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(D.count, __popc(m));
  }
}

struct Frustum {
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_h2d = nullptr;
  bool h2d_pending = false;
  int cap = 0;
  // One input blob and one result blob per handle, each with a pinned mirror, laid out with the stride of the
  // current call (n rounded up to 4): a host-buffer call is ONE H2D, one launch and ONE D2H -- at the 5000
  // points of a local map the fixed cost per cudaMemcpyAsync is what the call consists of.
  //   input : world_pos[3s] normal[3s] min_dist[s] max_dist[s]                      (floats)
  //   result: count, 3 pad | level[s] (int) | proj_x proj_y proj_xr view_cos depth [5s] (float) | in_view[s] full[s] (bytes)
  size_t stride = 0;
  float *d_in = nullptr, *h_in = nullptr;
  uint8_t *d_res = nullptr, *h_res = nullptr;
  long long launches = 0;

  explicit Frustum(int dev) : device(dev) {}
  ~Frustum() {
    release();
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_h2d) cudaEventDestroy(ev_h2d);
  }
  void release() {
    cudaFree(d_in); cudaFree(d_res);
    cudaFreeHost(h_in); cudaFreeHost(h_res);
    d_in = h_in = nullptr; d_res = h_res = nullptr;
    cap = 0;
  }
  static size_t res_bytes(size_t s) { return 16 + 4 * s + 20 * s + 2 * s; }
  int* count_of(uint8_t* base) const { return reinterpret_cast<int*>(base); }
  int* level_of(uint8_t* base) const { return reinterpret_cast<int*>(base + 16); }
  float* out_of(uint8_t* base) const { return reinterpret_cast<float*>(base + 16 + 4 * stride); }
  uint8_t* flags_of(uint8_t* base) const { return base + 16 + 24 * stride; }
  int ensure(int n) {
    if (!stream) {
      CUDA_TRYF(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      CUDA_TRYF(cudaEventCreate(&ev0));
      CUDA_TRYF(cudaEventCreate(&ev1));
      CUDA_TRYF(cudaEventCreateWithFlags(&ev_h2d, cudaEventDisableTiming));
    }
    if (n <= cap) return 0;
    if (h2d_pending) { CUDA_TRYF(cudaEventSynchronize(ev_h2d)); h2d_pending = false; }
    release();
    const size_t m = ((size_t)std::max(n, 1024) + 3) & ~(size_t)3;
    CUDA_TRYF(cudaMalloc(&d_in, sizeof(float) * 8 * m));
    CUDA_TRYF(cudaMallocHost(&h_in, sizeof(float) * 8 * m));
    CUDA_TRYF(cudaMalloc(&d_res, res_bytes(m)));
    CUDA_TRYF(cudaMallocHost(&h_res, res_bytes(m)));
    // the "written only when in view" outputs start from a defined state
    CUDA_TRYF(cudaMemset(d_res, 0, res_bytes(m)));
    cap = (int)m;
    return 0;
  }

  static int check_view(const orb_frustum_view* v) {
    if (!v || v->n < 0 || (v->n > 0 && (!v->world_pos || !v->normal || !v->min_dist || !v->max_dist)) ||
        v->n_levels <= 0 || !(v->log_scale_factor > 0)) {
      set_last_error("frame_is_in_frustum: bad view");
      return ORB_E_ARG;
    }
    return 0;
  }

  // device check, stream and buffers for n points
  int prepare(int n) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
      return ORB_E_NODEVICE;
    }
    CUDA_TRYF(cudaSetDevice(device));
    return ensure(n);
  }

  // uploads, launches, leaves the results on the device (after prepare())
  int enqueue(const orb_frustum_view* v, float cos_limit, cudaStream_t s, bool device_count = true) {
    const size_t n = (size_t)v->n;
    stride = std::max<size_t>((n + 3) & ~(size_t)3, 4);
    const size_t c = stride;
    // the pinned staging buffer of an earlier enqueue-only call may still be in flight
    if (h2d_pending) { CUDA_TRYF(cudaEventSynchronize(ev_h2d)); h2d_pending = false; }
    if (n) {
      memcpy(h_in, v->world_pos, sizeof(float) * 3 * n);
      memcpy(h_in + 3 * c, v->normal, sizeof(float) * 3 * n);
      memcpy(h_in + 6 * c, v->min_dist, sizeof(float) * n);
      memcpy(h_in + 7 * c, v->max_dist, sizeof(float) * n);
      CUDA_TRYF(cudaMemcpyAsync(d_in, h_in, sizeof(float) * 8 * c, cudaMemcpyHostToDevice, s));
      CUDA_TRYF(cudaEventRecord(ev_h2d, s));
      h2d_pending = true;
    }
    if (device_count) CUDA_TRYF(cudaMemsetAsync(count_of(d_res), 0, sizeof(int), s));
    FrustumDev D;
    D.world_pos = d_in; D.normal = d_in + 3 * c; D.min_dist = d_in + 6 * c; D.max_dist = d_in + 7 * c;
    uint8_t* fl = flags_of(d_res);
    float* o = out_of(d_res);
    D.in_view = fl; D.full = fl + c;
    D.proj_x = o; D.proj_y = o + c; D.proj_xr = o + 2 * c; D.view_cos = o + 3 * c; D.depth = o + 4 * c;
    D.level = level_of(d_res); D.count = device_count ? count_of(d_res) : nullptr; D.n = v->n;
    const FrustumFrame F = frustum_frame_of(*v);
    CUDA_TRYF(cudaEventRecord(ev0, s));
    if (n) {
      frustum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(F, D, cos_limit);
      launches += 1;
    }
    CUDA_TRYF(cudaEventRecord(ev1, s));
    CUDA_TRYF(cudaGetLastError());
    return 0;
  }
};

}  // namespace orbb200

using orbb200::Frustum;

struct orb_frustum { Frustum f; explicit orb_frustum(int dev) : f(dev) {} };

extern "C" {

int frustum_create(int device, orb_frustum** out) {
  if (!out || device < 0) { orbb200::set_last_error("frustum_create: bad argument"); return ORB_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    orbb200::set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
    return ORB_E_NODEVICE;
  }
  *out = new orb_frustum(device);
  return ORB_OK;
}

void frustum_destroy(orb_frustum* h) { delete h; }

int frame_is_in_frustum(orb_frustum* h, const orb_frustum_view* v, float cos_limit, uint8_t* track_in_view,
                        float* proj_x, float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos,
                        float* depth) {
  if (!h) return ORB_E_ARG;
  int rc = Frustum::check_view(v);
  if (rc) return rc;
  if (v->n > 0 && (!track_in_view || !proj_x || !proj_y || !proj_xr || !scale_level || !view_cos || !depth)) {
    orbb200::set_last_error("frame_is_in_frustum: NULL output");
    return ORB_E_ARG;
  }
  Frustum& f = h->f;
  if ((rc = f.prepare(v->n))) return rc;
  cudaStream_t s = f.stream;
  if ((rc = f.enqueue(v, cos_limit, s, false))) return rc;
  const size_t n = (size_t)v->n, c = f.stride;
  if (cudaMemcpyAsync(f.h_res, f.d_res, Frustum::res_bytes(c), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    orbb200::set_last_error(std::string("frame_is_in_frustum: ") + cudaGetErrorString(cudaGetLastError()));
    return ORB_E_CUDA;
  }
  const uint8_t* h_flags = f.flags_of(f.h_res);
  const float* h_out = f.out_of(f.h_res);
  const int* h_level = f.level_of(f.h_res);
  int n_in_view = 0;
  for (size_t i = 0; i < n; i++) {
    track_in_view[i] = h_flags[i];
    n_in_view += h_flags[i];
    proj_x[i] = h_out[i]; proj_y[i] = h_out[c + i];
    if (h_flags[c + i]) {  // members the reference only writes for points in view
      proj_xr[i] = h_out[2 * c + i]; view_cos[i] = h_out[3 * c + i]; depth[i] = h_out[4 * c + i];
      scale_level[i] = h_level[i];
    }
  }
  return n_in_view;
}

int frame_is_in_frustum_device(orb_frustum* h, const orb_frustum_view* v, float cos_limit, void* cuda_stream) {
  if (!h) return ORB_E_ARG;
  int rc = Frustum::check_view(v);
  if (rc) return rc;
  if ((rc = h->f.prepare(v->n))) return rc;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->f.stream;
  rc = h->f.enqueue(v, cos_limit, s);
  return rc ? rc : v->n;
}

int frustum_device_results(orb_frustum* h, const uint8_t** d_track_in_view, const float** d_proj_x,
                           const float** d_proj_y, const float** d_proj_xr, const int32_t** d_scale_level,
                           const float** d_view_cos, const float** d_depth, const int32_t** d_count) {
  if (!h || !h->f.d_res || !h->f.stride) return ORB_E_ARG;
  Frustum& f = h->f;
  const size_t c = f.stride;  // layout of the latest call
  const float* o = f.out_of(f.d_res);
  if (d_track_in_view) *d_track_in_view = f.flags_of(f.d_res);
  if (d_proj_x) *d_proj_x = o;
  if (d_proj_y) *d_proj_y = o + c;
  if (d_proj_xr) *d_proj_xr = o + 2 * c;
  if (d_view_cos) *d_view_cos = o + 3 * c;
  if (d_depth) *d_depth = o + 4 * c;
  if (d_scale_level) *d_scale_level = f.level_of(f.d_res);
  if (d_count) *d_count = f.count_of(f.d_res);
  return ORB_OK;
}

long long frustum_kernel_launches(const orb_frustum* h) { return h ? h->f.launches : 0; }

float frustum_last_ms(orb_frustum* h) {
  if (!h || !h->f.ev1 || cudaEventSynchronize(h->f.ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->f.ev0, h->f.ev1) != cudaSuccess) return 0.f;
  return ms;
}

int frustum_debug_host(const orb_frustum_view* v, float cos_limit, uint8_t* track_in_view, float* proj_x,
                       float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos, float* depth) {
  if (Frustum::check_view(v)) return ORB_E_ARG;
  const orbb200::FrustumFrame F = orbb200::frustum_frame_of(*v);
  int n_in = 0;
  for (int i = 0; i < v->n; i++) {
    const orbb200::FrustumPoint o = orbb200::frustum_point(F, v->world_pos + 3 * i, v->normal + 3 * i, v->min_dist[i],
                                                           v->max_dist[i], cos_limit);
    track_in_view[i] = o.in_view; proj_x[i] = o.proj_x; proj_y[i] = o.proj_y;
    if (o.full) { proj_xr[i] = o.proj_xr; view_cos[i] = o.view_cos; depth[i] = o.depth; scale_level[i] = o.level; }
    n_in += o.in_view;
  }
  return n_in;
}

}  // extern "C"
