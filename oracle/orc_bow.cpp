// TEST INFRASTRUCTURE ONLY (see orc_common.h) -- CPU restatement of DBoW2's
// TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// as Frame::ComputeBoW calls it (reference src/Frame.cc:738-745; vendored
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195 and :1218-1258, FORB.cpp:81-101,
// BowVector.cpp:34-84, FeatureVector.cpp:31-45), SURVEY.md 8(f-4), with the same containers
// (std::map) on a flat vocabulary.  TF_IDF / TF weighting, L1 scoring (the ORB vocabulary).
// Pinned by the reference's own DBoW2 object code (oracle/_ref/libref_bow.so, tests/test_ref_bow.py: bit for bit) and by
// an independent Python reading (tests/test_bow_oracle.py).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <map>
#include <vector>

#include "../include/orb_b200.h"

namespace {
int forb_distance(const uint8_t* a, const uint8_t* b) {  // FORB.cpp:81-101
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    unsigned int v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}
}  // namespace

extern "C" {

int orc_bow_transform(const orb_vocab_view* voc, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids,
                      double* bow_vals, int32_t* n_words, int32_t* fv_node_ids, int32_t* fv_ptr, int32_t* fv_idx,
                      int32_t* n_fv_nodes, int cap_words) {
  std::map<unsigned, double> v;                          // BowVector
  std::map<unsigned, std::vector<unsigned>> fv;          // FeatureVector
  int used = 0;
  for (int i = 0; i < n; i++) {
    const uint8_t* f = desc + (size_t)i * 32;
    // transform(feature, word_id, weight, &nid, levelsup) :1218-1258
    const int nid_level = voc->L - levelsup;
    unsigned nid = 0;                                     // nid_level <= 0 -> root
    int final_id = 0, current_level = 0;
    do {
      ++current_level;
      const int c0 = voc->child_ptr[final_id], c1 = voc->child_ptr[final_id + 1];
      final_id = voc->child_ids[c0];
      double best_d = forb_distance(f, voc->desc + (size_t)final_id * 32);
      for (int c = c0 + 1; c < c1; c++) {
        const int id = voc->child_ids[c];
        const double d = forb_distance(f, voc->desc + (size_t)id * 32);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid = (unsigned)final_id;
    } while (voc->child_ptr[final_id] != voc->child_ptr[final_id + 1]);  // !isLeaf()
    const unsigned id = (unsigned)voc->word_id[final_id];
    const double w = voc->weight[final_id];
    if (w > 0) {                                          // not stopped (:1157-1161)
      auto vit = v.lower_bound(id);                       // BowVector::addWeight
      if (vit != v.end() && !(id < vit->first)) vit->second += w;
      else v.insert(vit, std::make_pair(id, w));
      fv[nid].push_back((unsigned)i);                     // FeatureVector::addFeature
      used++;
    }
  }
  // mustNormalize(L1) (:1194): BowVector::normalize
  double norm = 0.0;
  for (auto& kv : v) norm += fabs(kv.second);
  if (norm > 0.0)
    for (auto& kv : v) kv.second /= norm;
  if ((int)v.size() > cap_words || (int)fv.size() > cap_words) return ORB_E_CAPACITY;
  int k = 0;
  for (auto& kv : v) { bow_ids[k] = (int32_t)kv.first; bow_vals[k] = kv.second; k++; }
  *n_words = k;
  k = 0;
  int pos = 0;
  for (auto& kv : fv) {
    fv_node_ids[k] = (int32_t)kv.first;
    fv_ptr[k] = pos;
    for (unsigned f : kv.second) fv_idx[pos++] = (int32_t)f;
    k++;
  }
  fv_ptr[k] = pos;
  *n_fv_nodes = k;
  return used;
}

}  // extern "C"
