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
  const unsigned m = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(D.count, __popc(m));
}

struct Frustum {
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_h2d = nullptr;
  bool h2d_pending = false;
  int cap = 0;
  float *d_in = nullptr, *h_in = nullptr;     // world_pos[3n] normal[3n] min[n] max[n]
  uint8_t *d_flags = nullptr, *h_flags = nullptr;  // in_view[n] full[n]
  float *d_out = nullptr, *h_out = nullptr;   // proj_x proj_y proj_xr view_cos depth, n each
  int *d_level = nullptr, *h_level = nullptr, *d_count = nullptr, *h_count = nullptr;
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
    cudaFree(d_in); cudaFree(d_flags); cudaFree(d_out); cudaFree(d_level); cudaFree(d_count);
    cudaFreeHost(h_in); cudaFreeHost(h_flags); cudaFreeHost(h_out); cudaFreeHost(h_level); cudaFreeHost(h_count);
    d_in = h_in = d_out = h_out = nullptr; d_flags = h_flags = nullptr;
    d_level = h_level = d_count = h_count = nullptr;
    cap = 0;
  }
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
    const size_t m = (size_t)std::max(n, 1024);
    CUDA_TRYF(cudaMalloc(&d_in, sizeof(float) * 8 * m));
    CUDA_TRYF(cudaMallocHost(&h_in, sizeof(float) * 8 * m));
    CUDA_TRYF(cudaMalloc(&d_flags, 2 * m));
    CUDA_TRYF(cudaMallocHost(&h_flags, 2 * m));
    CUDA_TRYF(cudaMalloc(&d_out, sizeof(float) * 5 * m));
    CUDA_TRYF(cudaMallocHost(&h_out, sizeof(float) * 5 * m));
    CUDA_TRYF(cudaMalloc(&d_level, sizeof(int) * m));
    CUDA_TRYF(cudaMallocHost(&h_level, sizeof(int) * m));
    CUDA_TRYF(cudaMalloc(&d_count, sizeof(int)));
    CUDA_TRYF(cudaMallocHost(&h_count, sizeof(int)));
    // the "written only when in view" outputs start from a defined state
    CUDA_TRYF(cudaMemset(d_out, 0, sizeof(float) * 5 * m));
    CUDA_TRYF(cudaMemset(d_level, 0, sizeof(int) * m));
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
  int enqueue(const orb_frustum_view* v, float cos_limit, cudaStream_t s) {
    const size_t n = (size_t)v->n, c = (size_t)cap;
    // the pinned staging buffer of an earlier enqueue-only call may still be in flight
    if (h2d_pending) { CUDA_TRYF(cudaEventSynchronize(ev_h2d)); h2d_pending = false; }
    if (n) {
      memcpy(h_in, v->world_pos, sizeof(float) * 3 * n);
      memcpy(h_in + 3 * c, v->normal, sizeof(float) * 3 * n);
      memcpy(h_in + 6 * c, v->min_dist, sizeof(float) * n);
      memcpy(h_in + 7 * c, v->max_dist, sizeof(float) * n);
      CUDA_TRYF(cudaMemcpyAsync(d_in, h_in, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
      CUDA_TRYF(cudaMemcpyAsync(d_in + 3 * c, h_in + 3 * c, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
      CUDA_TRYF(cudaMemcpyAsync(d_in + 6 * c, h_in + 6 * c, sizeof(float) * n, cudaMemcpyHostToDevice, s));
      CUDA_TRYF(cudaMemcpyAsync(d_in + 7 * c, h_in + 7 * c, sizeof(float) * n, cudaMemcpyHostToDevice, s));
      CUDA_TRYF(cudaEventRecord(ev_h2d, s));
      h2d_pending = true;
    }
    CUDA_TRYF(cudaMemsetAsync(d_count, 0, sizeof(int), s));
    FrustumDev D;
    D.world_pos = d_in; D.normal = d_in + 3 * c; D.min_dist = d_in + 6 * c; D.max_dist = d_in + 7 * c;
    D.in_view = d_flags; D.full = d_flags + c;
    D.proj_x = d_out; D.proj_y = d_out + c; D.proj_xr = d_out + 2 * c; D.view_cos = d_out + 3 * c; D.depth = d_out + 4 * c;
    D.level = d_level; D.count = d_count; D.n = v->n;
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
  if ((rc = f.enqueue(v, cos_limit, s))) return rc;
  const size_t n = (size_t)v->n, c = (size_t)f.cap;
  if (n) {
    if (cudaMemcpyAsync(f.h_flags, f.d_flags, n, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(f.h_flags + c, f.d_flags + c, n, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(f.h_level, f.d_level, sizeof(int) * n, cudaMemcpyDeviceToHost, s) != cudaSuccess) {
      orbb200::set_last_error("frame_is_in_frustum: D2H failed");
      return ORB_E_CUDA;
    }
    for (int k = 0; k < 5; k++)
      if (cudaMemcpyAsync(f.h_out + k * c, f.d_out + k * c, sizeof(float) * n, cudaMemcpyDeviceToHost, s) != cudaSuccess) {
        orbb200::set_last_error("frame_is_in_frustum: D2H failed");
        return ORB_E_CUDA;
      }
  }
  if (cudaMemcpyAsync(f.h_count, f.d_count, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    orbb200::set_last_error(std::string("frame_is_in_frustum: ") + cudaGetErrorString(cudaGetLastError()));
    return ORB_E_CUDA;
  }
  for (size_t i = 0; i < n; i++) {
    track_in_view[i] = f.h_flags[i];
    proj_x[i] = f.h_out[i]; proj_y[i] = f.h_out[c + i];
    if (f.h_flags[c + i]) {  // members the reference only writes for points in view
      proj_xr[i] = f.h_out[2 * c + i]; view_cos[i] = f.h_out[3 * c + i]; depth[i] = f.h_out[4 * c + i];
      scale_level[i] = f.h_level[i];
    }
  }
  return *f.h_count;
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
  if (!h || !h->f.d_out) return ORB_E_ARG;
  const Frustum& f = h->f;
  const size_t c = (size_t)f.cap;
  if (d_track_in_view) *d_track_in_view = f.d_flags;
  if (d_proj_x) *d_proj_x = f.d_out;
  if (d_proj_y) *d_proj_y = f.d_out + c;
  if (d_proj_xr) *d_proj_xr = f.d_out + 2 * c;
  if (d_view_cos) *d_view_cos = f.d_out + 3 * c;
  if (d_depth) *d_depth = f.d_out + 4 * c;
  if (d_scale_level) *d_scale_level = f.d_level;
  if (d_count) *d_count = f.d_count;
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
