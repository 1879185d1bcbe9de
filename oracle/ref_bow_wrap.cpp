// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE's vendored DBoW2 as object code
// (oracle/_ref/libref_bow.so): Thirdparty/DBoW2/DBoW2/{FORB,BowVector,FeatureVector,ScoringObject}.cpp and
// DUtils/{Random,Timestamp}.cpp compiled UNMODIFIED, TemplatedVocabulary.h instantiated for FORB, against the functional
// cv::Mat of oracle/cvcompat/.  The vocabulary enters through the reference's own loadFromTextFile (the ORBvoc.txt
// format); transform() is what Frame::ComputeBoW calls (src/Frame.cc:738-745).  tests/test_ref_bow.py holds the oracle
// (orc_bow.cpp) -- and through it the CUDA path -- against it.  Nothing in the product links this.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;  // = ORB_SLAM3::ORBVocabulary (ORBVocabulary.h)

extern "C" {

void* ref_voc_load_text(const char* path) {
  RefVocabulary* v = new RefVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void ref_voc_free(void* h) { delete (RefVocabulary*)h; }
unsigned ref_voc_size(void* h) { return ((RefVocabulary*)h)->size(); }

// desc: n x 32 bytes.  Outputs in map order (ascending ids): BowVector (word id, value) pairs; FeatureVector as CSR
// (node id, feature indices).  Returns 0, or -1 when a capacity is too small.
int ref_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, int cap, uint32_t* bow_ids, double* bow_vals,
                      int* n_words, uint32_t* fv_nodes, int* fv_ptr, int* fv_idx, int* n_nodes) {
  RefVocabulary* voc = (RefVocabulary*)h;
  std::vector<cv::Mat> features;  // Converter::toDescriptorVector: one 1 x 32 row per feature
  features.reserve(n);
  for (int i = 0; i < n; i++) {
    cv::Mat row(1, 32, CV_8U);
    memcpy(row.ptr(), desc + 32 * (size_t)i, 32);
    features.push_back(row);
  }
  DBoW2::BowVector bow;
  DBoW2::FeatureVector fv;
  voc->transform(features, bow, fv, levelsup);
  if ((int)bow.size() > cap || (int)fv.size() > cap) return -1;
  int k = 0;
  for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it, ++k) { bow_ids[k] = it->first; bow_vals[k] = it->second; }
  *n_words = k;
  k = 0;
  int p = 0;
  fv_ptr[0] = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++k) {
    fv_nodes[k] = it->first;
    for (size_t j = 0; j < it->second.size(); j++) { if (p >= cap) return -1; fv_idx[p++] = (int)it->second[j]; }
    fv_ptr[k + 1] = p;
  }
  *n_nodes = k;
  return 0;
}

}  // extern "C"
