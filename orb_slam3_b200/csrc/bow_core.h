// DBoW2's TemplatedVocabulary<FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195, :1218-1258; FORB.cpp:81-101;
// BowVector.cpp:34-84; FeatureVector.cpp:31-45) as two pieces of host/device source:
//
//   bow_descend_one   one feature down the tree (the per-thread body of bow_descend_kernel);
//   bow_assemble<BE>  the two std::maps of one frame, built by sorting (key, feature index) pairs:
//                     BowVector = per word the weights added in feature order (addWeight), then the
//                     L1 norm accumulated in ascending word order (normalize) -- both sums are
//                     sequential in exactly the reference's order, so the doubles are bit-identical;
//                     FeatureVector = per node the feature indices in ascending order (push_back).
//
// The body is written against a Backend (thread index, barrier, block scan) so the same source
// runs as one CUDA CTA (bow.cu) and single-threaded on the host (bow_debug_host, CPU tests).
#pragma once
#include <math.h>
#include <stdint.h>

#include "introsort_emul.h"  // ORB_HD

namespace orbb200 {

struct BowVocab {
  int n_nodes, L;
  const int* child_ptr;
  const int* child_ids;
  const uint8_t* desc;
  const double* weight;
  const int* word_id;
};

ORB_HD int bow_popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// FORB::distance: Hamming distance of two 32-byte descriptors (4-byte aligned)
ORB_HD int bow_distance(const uint32_t* a, const uint8_t* b) {
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(b);
  int d = 0;
  for (int i = 0; i < 8; i++) d += bow_popc32(a[i] ^ pb[i]);
  return d;
}

// transform(feature, word_id, weight, &nid, levelsup) :1218-1258
ORB_HD void bow_descend_one(const BowVocab& V, const uint8_t* feature, int levelsup, int& word, double& w, int& nid) {
  uint32_t f[8];
  const uint32_t* pf = reinterpret_cast<const uint32_t*>(feature);
  for (int i = 0; i < 8; i++) f[i] = pf[i];
  const int nid_level = V.L - levelsup;
  nid = 0;  // nid_level <= 0: root
  int final_id = 0, current_level = 0;
  int c0 = V.child_ptr[0], c1 = V.child_ptr[1];
  do {
    ++current_level;
    final_id = V.child_ids[c0];
    int best_d = bow_distance(f, V.desc + (size_t)final_id * 32);
    for (int c = c0 + 1; c < c1; c++) {
      const int id = V.child_ids[c];
      const int d = bow_distance(f, V.desc + (size_t)id * 32);
      if (d < best_d) { best_d = d; final_id = id; }  // first minimum in children order
    }
    if (current_level == nid_level) nid = final_id;
    c0 = V.child_ptr[final_id]; c1 = V.child_ptr[final_id + 1];
  } while (c0 != c1);  // !isLeaf()
  word = V.word_id[final_id];
  w = V.weight[final_id];
}

struct BowFrameOut {
  int* bow_ids; double* bow_vals; int* n_words;
  int* fv_node_ids; int* fv_ptr; int* fv_idx; int* n_fv_nodes;
  int* used;
  double* norm;  // one scratch double (L1 norm broadcast)
};

// Scratch: kw / kn = P 64-bit keys each (P = power of two >= n), flag = P ints.
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
#ifdef __CUDACC__
__host__ __device__
#endif
void bow_assemble(BE& be, int n, int P, const int* word, const double* weight, const int* nid,
                  unsigned long long* kw, unsigned long long* kn, int* flag, const BowFrameOut& o) {
  const int tid = be.tid(), nt = be.nthreads();
  const unsigned long long SENT = ~0ull;
  // (word, i) and (node, i) keys of the features that are not stopped (w > 0, :1157)
  for (int i = tid; i < P; i += nt) {
    const bool ok = i < n && weight[i] > 0;
    kw[i] = ok ? (((unsigned long long)(unsigned)word[i] << 32) | (unsigned)i) : SENT;
    kn[i] = ok ? (((unsigned long long)(unsigned)nid[i] << 32) | (unsigned)i) : SENT;
  }
  be.sync();
  // bitonic sort of both key arrays, ascending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nt) {
        const int x = i ^ j;
        if (x > i) {
          const bool asc = (i & k) == 0;
          unsigned long long a = kw[i], b = kw[x];
          if ((a > b) == asc) { kw[i] = b; kw[x] = a; }
          a = kn[i]; b = kn[x];
          if ((a > b) == asc) { kn[i] = b; kn[x] = a; }
        }
      }
      be.sync();
    }
  }
  // ---- BowVector: segment heads -> rank; head thread adds its word's weights in feature order
  for (int i = tid; i < P; i += nt)
    flag[i] = (kw[i] != SENT && (i == 0 || (kw[i] >> 32) != (kw[i - 1] >> 32))) ? 1 : 0;
  be.sync();
  const int nw = be.exclusive_scan(flag, P);
  for (int i = tid; i < P; i += nt) {
    if (kw[i] == SENT) continue;
    const unsigned wd = (unsigned)(kw[i] >> 32);
    if (i != 0 && (unsigned)(kw[i - 1] >> 32) == wd) continue;
    double acc = 0.0;
    bool first = true;
    for (int q = i; q < P && kw[q] != SENT && (unsigned)(kw[q] >> 32) == wd; q++) {
      const double wv = weight[(int)(kw[q] & 0xffffffffu)];
      if (first) { acc = wv; first = false; }  // insert(id, v)
      else acc += wv;                          // vit->second += v
    }
    o.bow_ids[flag[i]] = (int)wd;
    o.bow_vals[flag[i]] = acc;
  }
  be.sync();
  if (tid == 0) {  // BowVector::normalize(L1): ascending word order, one accumulator
    double norm = 0.0;
    for (int q = 0; q < nw; q++) norm += fabs(o.bow_vals[q]);
    *o.norm = norm;
    *o.n_words = nw;
  }
  be.sync();
  {
    const double norm = *o.norm;
    if (norm > 0.0)
      for (int q = tid; q < nw; q += nt) o.bow_vals[q] /= norm;
  }
  be.sync();
  // ---- FeatureVector: node heads -> rank; fv_idx is the sorted order itself
  for (int i = tid; i < P; i += nt)
    flag[i] = (kn[i] != SENT && (i == 0 || (kn[i] >> 32) != (kn[i - 1] >> 32))) ? 1 : 0;
  be.sync();
  const int nn = be.exclusive_scan(flag, P);
  int used_local = 0;
  for (int i = tid; i < P; i += nt) {
    if (kn[i] == SENT) continue;
    used_local++;
    o.fv_idx[i] = (int)(kn[i] & 0xffffffffu);  // valid keys sort to the front: position == rank
    if (i == 0 || (kn[i - 1] >> 32) != (kn[i] >> 32)) {
      o.fv_node_ids[flag[i]] = (int)(kn[i] >> 32);
      o.fv_ptr[flag[i]] = i;
    }
  }
  be.sync();
  // number of used features = first sentinel position
  for (int i = tid; i < P; i += nt)
    if (kn[i] != SENT && (i + 1 == P || kn[i + 1] == SENT)) { o.fv_ptr[nn] = i + 1; *o.used = i + 1; }
  if (tid == 0) {
    *o.n_fv_nodes = nn;
    if (nn == 0) { o.fv_ptr[0] = 0; *o.used = 0; }
  }
  be.sync();
  (void)used_local;
}

}  // namespace orbb200
