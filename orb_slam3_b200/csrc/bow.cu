// Frame::ComputeBoW (reference src/Frame.cc:738-745) = DBoW2 TemplatedVocabulary<FORB>::transform with
// levelsup = 4; the per-feature and per-frame source lives in bow_core.h.
//
//   bow_descend_kernel   thread per feature: L levels x k children Hamming distances against the
//                        resident vocabulary (1.1 M nodes x 32 B for the ORB vocabulary = 35 MB in HBM);
//                        latency-bound gathers, k independent 32-byte loads in flight per thread.
//   bow_assemble_kernel  one CTA per frame: bitonic sort of (word | node, feature) keys in shared
//                        memory, sequential per-word / per-vector double sums in the reference's order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "octree_core.h"  // HostBackend
#include "bow_core.h"
#include "cta_backend.cuh"
#include "orb_engine.h"

namespace orbb200 {

#define CUDA_TRYB(expr)                                                                \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return ORB_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

constexpr int BOW_ASM_THREADS = 512;

__global__ void __launch_bounds__(128) bow_descend_kernel(const BowVocab V, const uint8_t* __restrict__ desc, int n,
                                                          int levelsup, int* __restrict__ word,
                                                          double* __restrict__ weight, int* __restrict__ nid) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  int wd, nd;
  double w;
  bow_descend_one(V, desc + (size_t)i * 32, levelsup, wd, w, nd);
  word[i] = wd; weight[i] = w; nid[i] = nd;
}

__global__ void __launch_bounds__(BOW_ASM_THREADS) bow_assemble_kernel(int n, int P, const int* word, const double* weight,
                                                                       const int* nid, BowFrameOut o) {
  extern __shared__ unsigned long long bow_smem[];
  __shared__ int s_ints[48];
  unsigned long long* kw = bow_smem;
  unsigned long long* kn = bow_smem + P;
  int* flag = reinterpret_cast<int*>(bow_smem + 2 * (size_t)P);
  CtaBackend be{s_ints};
  bow_assemble(be, n, P, word, weight, nid, kw, kn, flag, o);
}

// Frames with more keys than fit in shared memory (monocular initialisation extracts 5 x nFeatures, Tracking.cc:2536):
// the same source on global-memory scratch (L2-resident; slower per stage, no limit on n)
__global__ void __launch_bounds__(BOW_ASM_THREADS) bow_assemble_global_kernel(int n, int P, const int* word, const double* weight,
                                                                              const int* nid, BowFrameOut o,
                                                                              unsigned long long* keys) {
  __shared__ int s_ints[48];
  CtaBackend be{s_ints};
  bow_assemble(be, n, P, word, weight, nid, keys, keys + P, reinterpret_cast<int*>(keys + 2 * (size_t)P), o);
}

static int next_pow2(int n) { int p = 2; while (p < n) p <<= 1; return p; }

static int check_vocab(const orb_vocab_view* v) {
  if (!v || v->n_nodes < 2 || v->L < 1 || !v->child_ptr || !v->child_ids || !v->desc || !v->weight || !v->word_id ||
      v->child_ptr[1] <= v->child_ptr[0]) {
    set_last_error("vocabulary view: need >= 2 nodes, a root with children and all arrays");
    return ORB_E_ARG;
  }
  return 0;
}

struct Vocab {
  int device;
  BowVocab V{};            // device pointers
  int n_children = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int cap = 0;
  uint8_t* d_desc = nullptr;
  int *d_word = nullptr, *d_nid = nullptr, *d_bow_ids = nullptr, *d_fv_nodes = nullptr, *d_fv_ptr = nullptr,
      *d_fv_idx = nullptr, *d_counts = nullptr;
  double *d_weight = nullptr, *d_bow_vals = nullptr, *d_norm = nullptr;
  unsigned long long* d_keys = nullptr;  // global-memory sort scratch of oversized frames (P * 20 bytes)
  size_t keys_bytes = 0;
  int *h_ints = nullptr;   // pinned: counts[3] | bow_ids | fv_nodes | fv_ptr | fv_idx
  double* h_vals = nullptr;
  long long launches = 0;

  explicit Vocab(int dev) : device(dev) {}
  ~Vocab() {
    release_scratch();
    cudaFree((void*)V.child_ptr); cudaFree((void*)V.child_ids); cudaFree((void*)V.desc);
    cudaFree((void*)V.weight); cudaFree((void*)V.word_id);
    if (stream) cudaStreamDestroy(stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
  void release_scratch() {
    cudaFree(d_desc); cudaFree(d_word); cudaFree(d_nid); cudaFree(d_bow_ids); cudaFree(d_fv_nodes);
    cudaFree(d_fv_ptr); cudaFree(d_fv_idx); cudaFree(d_counts); cudaFree(d_weight); cudaFree(d_bow_vals);
    cudaFree(d_norm); cudaFree(d_keys); cudaFreeHost(h_ints); cudaFreeHost(h_vals);
    d_keys = nullptr; keys_bytes = 0;
    d_desc = nullptr; d_word = d_nid = d_bow_ids = d_fv_nodes = d_fv_ptr = d_fv_idx = d_counts = nullptr;
    d_weight = d_bow_vals = d_norm = nullptr; h_ints = nullptr; h_vals = nullptr;
    cap = 0;
  }
  int upload(const orb_vocab_view* v) {
    CUDA_TRYB(cudaSetDevice(device));
    CUDA_TRYB(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_TRYB(cudaEventCreate(&ev0));
    CUDA_TRYB(cudaEventCreate(&ev1));
    const size_t nn = (size_t)v->n_nodes;
    n_children = v->child_ptr[nn];
    for (size_t i = 0; i < nn; i++)
      if (v->child_ptr[i + 1] < v->child_ptr[i]) { set_last_error("vocabulary view: child_ptr not monotone"); return ORB_E_ARG; }
    for (int c = 0; c < n_children; c++)
      if (v->child_ids[c] <= 0 || v->child_ids[c] >= v->n_nodes) { set_last_error("vocabulary view: child id out of range"); return ORB_E_ARG; }
    // every allocation is handed to V at once, so that ~Vocab() frees it if a later step fails
    int *cp = nullptr, *ci = nullptr, *wi = nullptr;
    uint8_t* ds = nullptr;
    double* wt = nullptr;
    V.n_nodes = v->n_nodes; V.L = v->L;
    CUDA_TRYB(cudaMalloc(&cp, sizeof(int) * (nn + 1)));
    V.child_ptr = cp;
    CUDA_TRYB(cudaMalloc(&ci, sizeof(int) * std::max(n_children, 1)));
    V.child_ids = ci;
    CUDA_TRYB(cudaMalloc(&ds, 32 * nn));
    V.desc = ds;
    CUDA_TRYB(cudaMalloc(&wt, sizeof(double) * nn));
    V.weight = wt;
    CUDA_TRYB(cudaMalloc(&wi, sizeof(int) * nn));
    V.word_id = wi;
    CUDA_TRYB(cudaMemcpy(cp, v->child_ptr, sizeof(int) * (nn + 1), cudaMemcpyHostToDevice));
    CUDA_TRYB(cudaMemcpy(ci, v->child_ids, sizeof(int) * n_children, cudaMemcpyHostToDevice));
    CUDA_TRYB(cudaMemcpy(ds, v->desc, 32 * nn, cudaMemcpyHostToDevice));
    CUDA_TRYB(cudaMemcpy(wt, v->weight, sizeof(double) * nn, cudaMemcpyHostToDevice));
    CUDA_TRYB(cudaMemcpy(wi, v->word_id, sizeof(int) * nn, cudaMemcpyHostToDevice));
    return 0;
  }
  int ensure(int n) {
    if (n <= cap) return 0;
    release_scratch();
    const size_t m = (size_t)std::max(n, 2048);
    CUDA_TRYB(cudaMalloc(&d_desc, 32 * m));
    CUDA_TRYB(cudaMalloc(&d_word, sizeof(int) * m));
    CUDA_TRYB(cudaMalloc(&d_nid, sizeof(int) * m));
    CUDA_TRYB(cudaMalloc(&d_weight, sizeof(double) * m));
    CUDA_TRYB(cudaMalloc(&d_bow_ids, sizeof(int) * (m + 1)));
    CUDA_TRYB(cudaMalloc(&d_bow_vals, sizeof(double) * (m + 1)));
    CUDA_TRYB(cudaMalloc(&d_fv_nodes, sizeof(int) * (m + 1)));
    CUDA_TRYB(cudaMalloc(&d_fv_ptr, sizeof(int) * (m + 2)));
    CUDA_TRYB(cudaMalloc(&d_fv_idx, sizeof(int) * m));
    CUDA_TRYB(cudaMalloc(&d_counts, sizeof(int) * 4));
    CUDA_TRYB(cudaMalloc(&d_norm, sizeof(double)));
    CUDA_TRYB(cudaMallocHost(&h_ints, sizeof(int) * (4 * m + 8)));
    CUDA_TRYB(cudaMallocHost(&h_vals, sizeof(double) * (m + 1)));
    cap = (int)m;
    return 0;
  }

  // d_in: descriptors already on the device (or nullptr: upload `desc`)
  int run(const uint8_t* desc, const uint8_t* d_in, cudaStream_t wait_on, int n, int levelsup, int32_t* bow_ids,
          double* bow_vals, int32_t* n_words, int32_t* fv_node_ids, int32_t* fv_ptr, int32_t* fv_idx,
          int32_t* n_fv_nodes, int cap_words) {
    if (n < 0 || levelsup < 0 || !bow_ids || !bow_vals || !n_words || !fv_node_ids || !fv_ptr || !fv_idx || !n_fv_nodes ||
        (n > 0 && !desc && !d_in)) {
      set_last_error("bow_transform: bad argument");
      return ORB_E_ARG;
    }
    CUDA_TRYB(cudaSetDevice(device));
    int rc = ensure(n);
    if (rc) return rc;
    const int P = next_pow2(n);
    const size_t smem = (size_t)P * 20;
    const bool in_smem = smem <= 200 * 1024;  // up to 8192 features per frame
    if (!in_smem && smem > keys_bytes) {
      cudaFree(d_keys);
      d_keys = nullptr; keys_bytes = 0;
      CUDA_TRYB(cudaMalloc(&d_keys, smem));
      keys_bytes = smem;
    }
    cudaStream_t s = stream;
    if (wait_on && wait_on != s) {  // descriptors produced on another stream
      CUDA_TRYB(cudaEventRecord(ev0, wait_on));
      CUDA_TRYB(cudaStreamWaitEvent(s, ev0, 0));
    }
    const uint8_t* dd = d_in;
    if (!d_in && n) {
      CUDA_TRYB(cudaMemcpyAsync(d_desc, desc, 32 * (size_t)n, cudaMemcpyHostToDevice, s));
      dd = d_desc;
    }
    BowFrameOut o;
    o.bow_ids = d_bow_ids; o.bow_vals = d_bow_vals; o.n_words = d_counts; o.fv_node_ids = d_fv_nodes;
    o.fv_ptr = d_fv_ptr; o.fv_idx = d_fv_idx; o.n_fv_nodes = d_counts + 1; o.used = d_counts + 2; o.norm = d_norm;
    CUDA_TRYB(cudaEventRecord(ev0, s));
    if (n) {
      bow_descend_kernel<<<(n + 127) / 128, 128, 0, s>>>(V, dd, n, levelsup, d_word, d_weight, d_nid);
      launches++;
    }
    if (in_smem) {
      CUDA_TRYB(raise_dynamic_smem((const void*)bow_assemble_kernel, smem, device));
      bow_assemble_kernel<<<1, BOW_ASM_THREADS, smem, s>>>(n, P, d_word, d_weight, d_nid, o);
    } else {
      bow_assemble_global_kernel<<<1, BOW_ASM_THREADS, 0, s>>>(n, P, d_word, d_weight, d_nid, o, d_keys);
    }
    launches++;
    CUDA_TRYB(cudaEventRecord(ev1, s));
    CUDA_TRYB(cudaGetLastError());
    const size_t m = (size_t)n;
    CUDA_TRYB(cudaMemcpyAsync(h_ints, d_counts, sizeof(int) * 3, cudaMemcpyDeviceToHost, s));
    if (n) {
      CUDA_TRYB(cudaMemcpyAsync(h_ints + 4, d_bow_ids, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
      CUDA_TRYB(cudaMemcpyAsync(h_ints + 4 + cap, d_fv_nodes, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
      CUDA_TRYB(cudaMemcpyAsync(h_ints + 4 + 2 * (size_t)cap, d_fv_ptr, sizeof(int) * (m + 1), cudaMemcpyDeviceToHost, s));
      CUDA_TRYB(cudaMemcpyAsync(h_ints + 6 + 3 * (size_t)cap, d_fv_idx, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
      CUDA_TRYB(cudaMemcpyAsync(h_vals, d_bow_vals, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRYB(cudaStreamSynchronize(s));
    const int nw = h_ints[0], nn = h_ints[1], used = h_ints[2];
    if (nw > cap_words || nn > cap_words) { set_last_error("bow_transform: cap_words too small"); return ORB_E_CAPACITY; }
    *n_words = nw; *n_fv_nodes = nn;
    memcpy(bow_ids, h_ints + 4, sizeof(int) * nw);
    memcpy(bow_vals, h_vals, sizeof(double) * nw);
    memcpy(fv_node_ids, h_ints + 4 + cap, sizeof(int) * nn);
    if (n) memcpy(fv_ptr, h_ints + 4 + 2 * (size_t)cap, sizeof(int) * (nn + 1));
    else fv_ptr[0] = 0;
    memcpy(fv_idx, h_ints + 6 + 3 * (size_t)cap, sizeof(int) * used);
    return used;
  }
};

}  // namespace orbb200

using orbb200::Vocab;

struct orb_vocab { Vocab v; explicit orb_vocab(int dev) : v(dev) {} };

extern "C" {

int vocab_create(int device, const orb_vocab_view* view, orb_vocab** out) {
  if (!out || device < 0) { orbb200::set_last_error("vocab_create: bad argument"); return ORB_E_ARG; }
  int rc = orbb200::check_vocab(view);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    orbb200::set_last_error("no CUDA device: orb_slam3_b200 has no CPU path");
    return ORB_E_NODEVICE;
  }
  orb_vocab* h = new orb_vocab(device);
  rc = h->v.upload(view);
  if (rc) { delete h; return rc; }
  *out = h;
  return ORB_OK;
}

void vocab_destroy(orb_vocab* h) { delete h; }

int bow_transform(orb_vocab* h, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_vals,
                  int32_t* n_words, int32_t* fv_node_ids, int32_t* fv_ptr, int32_t* fv_idx, int32_t* n_fv_nodes,
                  int cap_words) {
  if (!h) return ORB_E_ARG;
  return h->v.run(desc, nullptr, nullptr, n, levelsup, bow_ids, bow_vals, n_words, fv_node_ids, fv_ptr, fv_idx,
                  n_fv_nodes, cap_words);
}

int bow_transform_extracted(orb_vocab* h, orb_extractor* ex, int frame, int levelsup, int32_t* bow_ids,
                            double* bow_vals, int32_t* n_words, int32_t* fv_node_ids, int32_t* fv_ptr,
                            int32_t* fv_idx, int32_t* n_fv_nodes, int cap_words) {
  if (!h || !ex || !ex->e.initialized || frame < 0 || frame >= ex->e.last_batch || ex->e.device != h->v.device) {
    orbb200::set_last_error("bow_transform_extracted: the extractor holds no such frame on this device");
    return ORB_E_ARG;
  }
  orbb200::Engine& e = ex->e;
  cudaStream_t es = e.last_stream ? e.last_stream : e.stream;
  int n = 0;
  if (cudaSetDevice(e.device) != cudaSuccess ||
      cudaMemcpyAsync(&n, e.d_n + frame, sizeof(int), cudaMemcpyDeviceToHost, es) != cudaSuccess ||
      cudaStreamSynchronize(es) != cudaSuccess) {
    orbb200::set_last_error("bow_transform_extracted: reading the keypoint count failed");
    return ORB_E_CUDA;
  }
  return h->v.run(nullptr, e.d_desc + (size_t)frame * e.out_cap * 32, es, n, levelsup, bow_ids, bow_vals, n_words,
                  fv_node_ids, fv_ptr, fv_idx, n_fv_nodes, cap_words);
}

long long bow_kernel_launches(const orb_vocab* h) { return h ? h->v.launches : 0; }

float bow_last_ms(orb_vocab* h) {
  if (!h || !h->v.ev1 || cudaEventSynchronize(h->v.ev1) != cudaSuccess) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->v.ev0, h->v.ev1) != cudaSuccess) return 0.f;
  return ms;
}

int bow_debug_host(const orb_vocab_view* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids,
                   double* bow_vals, int32_t* n_words, int32_t* fv_node_ids, int32_t* fv_ptr, int32_t* fv_idx,
                   int32_t* n_fv_nodes, int cap_words) {
  if (orbb200::check_vocab(v) || n < 0 || (n > 0 && !desc)) return ORB_E_ARG;
  orbb200::BowVocab V;
  V.n_nodes = v->n_nodes; V.L = v->L; V.child_ptr = v->child_ptr; V.child_ids = v->child_ids; V.desc = v->desc;
  V.weight = v->weight; V.word_id = v->word_id;
  std::vector<int> word(n + 1), nid(n + 1);
  std::vector<double> weight(n + 1);
  std::vector<uint32_t> aligned(8);
  for (int i = 0; i < n; i++) {
    memcpy(aligned.data(), desc + (size_t)i * 32, 32);
    orbb200::bow_descend_one(V, reinterpret_cast<const uint8_t*>(aligned.data()), levelsup, word[i], weight[i], nid[i]);
  }
  const int P = orbb200::next_pow2(n);
  std::vector<unsigned long long> kw(P), kn(P);
  std::vector<int> flag(P), ids(n + 1), fnodes(n + 1), fptr(n + 2), fidx(n + 1);
  std::vector<double> vals(n + 1);
  int counts[3] = {0, 0, 0};
  double norm = 0;
  orbb200::BowFrameOut o;
  o.bow_ids = ids.data(); o.bow_vals = vals.data(); o.n_words = &counts[0]; o.fv_node_ids = fnodes.data();
  o.fv_ptr = fptr.data(); o.fv_idx = fidx.data(); o.n_fv_nodes = &counts[1]; o.used = &counts[2]; o.norm = &norm;
  orbb200::HostBackend be;
  orbb200::bow_assemble(be, n, P, word.data(), weight.data(), nid.data(), kw.data(), kn.data(), flag.data(), o);
  if (counts[0] > cap_words || counts[1] > cap_words) return ORB_E_CAPACITY;
  *n_words = counts[0]; *n_fv_nodes = counts[1];
  memcpy(bow_ids, ids.data(), sizeof(int) * counts[0]);
  memcpy(bow_vals, vals.data(), sizeof(double) * counts[0]);
  memcpy(fv_node_ids, fnodes.data(), sizeof(int) * counts[1]);
  memcpy(fv_ptr, fptr.data(), sizeof(int) * (counts[1] + 1));
  memcpy(fv_idx, fidx.data(), sizeof(int) * counts[2]);
  return counts[2];
}

}  // extern "C"
