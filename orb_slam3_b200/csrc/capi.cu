// extern "C" surface of liborbb200.so (include/orb_b200.h) -- extractor part.
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

#include "../../include/orb_b200.h"
#include "octree_core.h"
#include "orb_engine.h"

using orbb200::Engine;

extern "C" {

const char* orb_version(void) { return "orb_slam3_b200 0.1 (sm_100a)"; }
const char* orb_last_error(void) { return orbb200::last_error(); }

int orb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device,
               orb_extractor** out) {
  if (!out || nfeatures <= 0 || nlevels <= 0 || nlevels > 16 || !(scale_factor > 1.0f) || device < 0) {
    orbb200::set_last_error("orb_create: bad argument");
    return ORB_E_ARG;
  }
  *out = new orb_extractor(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, device);
  return ORB_OK;
}

void orb_destroy(orb_extractor* h) { delete h; }

int orb_get_levels(const orb_extractor* h) { return h->e.nlevels; }
float orb_get_scale_factor(const orb_extractor* h) { return (float)h->e.scale_factor; }
static int copy_tab(const std::vector<float>& v, float* out) {
  if (!out) return ORB_E_ARG;
  memcpy(out, v.data(), sizeof(float) * v.size());
  return (int)v.size();
}
int orb_get_scale_factors(const orb_extractor* h, float* out) { return copy_tab(h->e.scale, out); }
int orb_get_inverse_scale_factors(const orb_extractor* h, float* out) { return copy_tab(h->e.inv_scale, out); }
int orb_get_scale_sigma_squares(const orb_extractor* h, float* out) { return copy_tab(h->e.sigma2, out); }
int orb_get_inverse_scale_sigma_squares(const orb_extractor* h, float* out) { return copy_tab(h->e.inv_sigma2, out); }
int orb_get_features_per_level(const orb_extractor* h, int* out) {
  if (!out) return ORB_E_ARG;
  memcpy(out, h->e.quota.data(), sizeof(int) * h->e.quota.size());
  return (int)h->e.quota.size();
}

int orb_extract(orb_extractor* h, const uint8_t* img, int rows, int cols, size_t step, int lap0, int lap1,
                orb_keypoint* kps, uint8_t* desc, int cap, int* n) {
  if (!h || !n) return ORB_E_ARG;
  if (!img || rows <= 0 || cols <= 0) return ORB_E_EMPTY;
  int lap[2] = {lap0, lap1};
  int mono = 0;
  const uint8_t* imgs[1] = {img};
  int rc = h->e.extract_batch_host(1, imgs, rows, cols, step, lap, kps, desc, cap, n, &mono);
  return rc < 0 ? rc : mono;
}

int orb_extract_batch(orb_extractor* h, int batch, const uint8_t* const* imgs, int rows, int cols, size_t step,
                      const int* lap, orb_keypoint* kps, uint8_t* desc, int cap, int* n, int* mono_index) {
  if (!h) return ORB_E_ARG;
  return h->e.extract_batch_host(batch, imgs, rows, cols, step, lap, kps, desc, cap, n, mono_index);
}

int orb_extract_batch_device(orb_extractor* h, int batch, const uint8_t* d_imgs, size_t frame_stride, int rows,
                             int cols, size_t step, const int* lap, void* cuda_stream) {
  if (!h) return ORB_E_ARG;
  return h->e.extract_batch_device(batch, d_imgs, frame_stride, rows, cols, step, lap, (cudaStream_t)cuda_stream);
}

int orb_device_results(orb_extractor* h, const orb_keypoint** d_kps, const uint8_t** d_desc, const int** d_n,
                       const int** d_mono_index, int* cap_per_frame) {
  if (!h || !h->e.initialized) return ORB_E_ARG;
  if (d_kps) *d_kps = h->e.d_kps;
  if (d_desc) *d_desc = h->e.d_desc;
  if (d_n) *d_n = h->e.d_n;
  if (d_mono_index) *d_mono_index = h->e.d_mono;
  if (cap_per_frame) *cap_per_frame = h->e.out_cap;
  return ORB_OK;
}

int orb_download_results(orb_extractor* h, int frame, orb_keypoint* kps, uint8_t* desc, int cap, int* n) {
  if (!h || !h->e.initialized || frame < 0 || frame >= h->e.last_batch || !kps || !desc || !n) return ORB_E_ARG;
  Engine& e = h->e;
  cudaSetDevice(e.device);
  cudaStream_t s = e.last_stream ? e.last_stream : e.stream;
  int cnt[2] = {0, 0};
  if (cudaMemcpyAsync(&cnt[0], e.d_n + frame, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaMemcpyAsync(&cnt[1], e.d_mono + frame, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    orbb200::set_last_error(cudaGetErrorString(cudaGetLastError()));
    return ORB_E_CUDA;
  }
  *n = cnt[0];
  if (cnt[0] > cap) { orbb200::set_last_error("keypoint buffer too small"); return ORB_E_CAPACITY; }
  if (cudaMemcpyAsync(kps, e.d_kps + (size_t)frame * e.out_cap, sizeof(orb_keypoint) * cnt[0], cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaMemcpyAsync(desc, e.d_desc + (size_t)frame * e.out_cap * 32, (size_t)32 * cnt[0], cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    orbb200::set_last_error(cudaGetErrorString(cudaGetLastError()));
    return ORB_E_CUDA;
  }
  return cnt[1];
}

int orb_synchronize(orb_extractor* h) {
  if (!h || !h->e.initialized) return ORB_E_ARG;
  cudaSetDevice(h->e.device);
  cudaError_t e = cudaStreamSynchronize(h->e.last_stream ? h->e.last_stream : h->e.stream);
  if (e != cudaSuccess) {
    orbb200::set_last_error(cudaGetErrorString(e));
    return ORB_E_CUDA;
  }
  return ORB_OK;
}

int orb_pyramid(orb_extractor* h, int frame, int level, const uint8_t** ptr, int* rows, int* cols, size_t* step) {
  if (!h || !h->e.initialized || frame < 0 || frame >= h->e.last_batch || level < 0 || level >= h->e.nlevels)
    return ORB_E_ARG;
  int rc = h->e.fetch_pyramid();
  if (rc) return rc;
  const orbb200::LevelDev& L = h->e.levels[level];
  if (ptr) *ptr = h->e.h_pyr + (size_t)frame * h->e.pyr_frame_bytes + L.img_off;
  if (rows) *rows = L.h;
  if (cols) *cols = L.w;
  if (step) *step = (size_t)L.pitch;
  return ORB_OK;
}

int orb_set_profiling(orb_extractor* h, int enabled) {
  if (!h) return ORB_E_ARG;
  h->e.profiling = enabled != 0;
  return ORB_OK;
}
int orb_stage_times(orb_extractor* h, double* ms, long long* launches, int reset) {
  if (!h) return ORB_E_ARG;
  return h->e.collect_times(ms, launches, reset != 0);
}
const char* orb_stage_name(int stage) {
  static const char* names[ORB_NUM_STAGES] = {"h2d", "pyramid", "fast", "octree", "blur", "layout", "describe", "d2h"};
  return (stage >= 0 && stage < ORB_NUM_STAGES) ? names[stage] : "?";
}
long long orb_kernel_launches(const orb_extractor* h) { return h ? h->e.total_launches : 0; }

int orb_debug_candidates(orb_extractor* h, int frame, int level, int* xys, int cap) {
  if (!h) return ORB_E_ARG;
  return h->e.debug_candidates(frame, level, xys, cap);
}

int orb_debug_octree_host(const int* xys, int n, int band_w, int band_h, int n_features, int w_cell, int h_cell,
                          int n_cols, int* out_xys, int out_cap) {
  using namespace orbb200;
  OctreeLevelParams p;
  p.bandW = band_w; p.bandH = band_h; p.N = n_features;
  p.nIni = (int)roundf((float)band_w / (float)band_h);
  if (p.nIni < 1) return ORB_E_ARG;
  p.hX = (float)band_w / p.nIni;
  p.wCell = w_cell; p.hCell = h_cell; p.nCols = n_cols;
  p.node_cap = p.N + 4 * p.nIni + 16;
  std::vector<Cand> cand(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    cand[i].xy = (uint32_t)xys[3 * i] | ((uint32_t)xys[3 * i + 1] << 16);
    cand[i].score = (uint32_t)xys[3 * i + 2];
  }
  const size_t nc = p.node_cap;
  std::vector<int> pt_node(n + 1), ints(nc * (10 + 16 + 5));
  std::vector<uint8_t> pt_q(n + 1);
  std::vector<SortNode> sortbuf(nc);
  std::vector<int> sortwork(6 * nc);
  std::vector<unsigned long long> best(nc);
  OctreeScratch s;
  s.pt_node = pt_node.data(); s.pt_q = pt_q.data();
  int* q = ints.data();
  for (int b = 0; b < 2; b++) for (int f = 0; f < 5; f++) { s.nd[b][f] = q; q += nc; }
  s.childcnt = q; q += 4 * nc; s.cidx = q; q += 4 * nc; s.eidx = q; q += 4 * nc; s.remap = q; q += 4 * nc;
  s.rank = q; q += nc; s.proc = q; q += nc; s.surv = q; q += nc; s.tmp = q; q += nc; s.expand_pos = q; q += nc;
  s.sortbuf = sortbuf.data(); s.sortwork = sortwork.data(); s.best = best.data();
  std::vector<int> out(3 * nc);
  HostBackend be;
  int m = octree_select(be, cand.data(), n, p, s, out.data());
  for (int i = 0; i < m && i < out_cap; i++) {
    out_xys[3 * i] = out[3 * i]; out_xys[3 * i + 1] = out[3 * i + 1]; out_xys[3 * i + 2] = out[3 * i + 2];
  }
  return m;
}

// the level-synchronous form (what the octree CTA runs), single-threaded on the host
int orb_debug_introsort_levels(const int* count, const int* ulx, int n, int* perm_out) {
  std::vector<orbb200::SortNode> v(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) v[i] = orbb200::make_sort_node(count[i], ulx[i], i);
  std::vector<int> work(6 * (size_t)(n / 8 + 8));
  orbb200::HostBackend be;
  orbb200::introsort_levels(be, v.data(), n, work.data(), n / 8 + 8);
  for (int i = 0; i < n; i++) perm_out[i] = v[i].id;
  return n;
}

int orb_debug_introsort(const int* count, const int* ulx, int n, int* perm_out) {
  std::vector<orbb200::SortNode> v(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) v[i] = orbb200::make_sort_node(count[i], ulx[i], i);
  orbb200::introsort_emul(v.data(), n);
  for (int i = 0; i < n; i++) perm_out[i] = v[i].id;
  return n;
}

int orb_debug_sincos_device(int device, const float* x, size_t n, float* cos_out, float* sin_out) {
  if (!x || !cos_out || !sin_out || n == 0) return ORB_E_ARG;
  return orbb200::debug_sincos_device(device, x, n, cos_out, sin_out);
}

int orb_debug_sincos_host(const float* x, size_t n, float* cos_out, float* sin_out, int fused) {
  if (!x || !cos_out || !sin_out) return ORB_E_ARG;
  orbb200::debug_sincos_host(x, n, cos_out, sin_out, fused);
  return 0;
}

}  // extern "C"
