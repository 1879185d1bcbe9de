// Replacement for Frame::ComputeBoW (reference src/Frame.cc:738-745) and KeyFrame::ComputeBoW
// (src/KeyFrame.cc:98-107), SURVEY.md 8(f-4): the call into the vendored DBoW2
// (ORBVocabulary::transform(features, mBowVec, mFeatVec, 4), TemplatedVocabulary.h:1127-1195) becomes
// bow_transform() of liborbb200.so against a vocabulary uploaded once per process; the two std::maps
// are rebuilt from the flat outputs (already in key order, so every insert is an O(1) hinted insert).
// The vocabulary object stays the reference's (loading, scoring, KeyFrameDatabase are untouched).
// Syntax-checked against the reference's headers over stand-ins for its third-party libraries (tests/test_shim_syntax.py); not linked here -- see INTEGRATION.md.
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Converter.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "orb_b200.h"

namespace ORB_SLAM3 {

namespace {

// Flattens the TemplatedVocabulary once (m_nodes is protected: the two-line friend declaration /
// accessor added to ORBVocabulary.h is described in INTEGRATION.md) and keeps the device handle.
orb_vocab* device_vocabulary(ORBVocabulary* voc) {
  static std::mutex mu;
  static ORBVocabulary* cached = nullptr;
  static orb_vocab* handle = nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (handle && cached == voc) return handle;
  std::vector<int32_t> child_ptr, child_ids, word_id;
  std::vector<uint8_t> desc;
  std::vector<double> weight;
  int L = 0;
  voc->flatten(L, child_ptr, child_ids, desc, weight, word_id);  // walks m_nodes in id order
  orb_vocab_view v;
  v.n_nodes = (int32_t)weight.size(); v.L = L;
  v.child_ptr = child_ptr.data(); v.child_ids = child_ids.data(); v.desc = desc.data();
  v.weight = weight.data(); v.word_id = word_id.data();
  if (handle) vocab_destroy(handle);
  if (vocab_create(/*device=*/0, &v, &handle) != ORB_OK)
    throw std::runtime_error(std::string("vocab_create: ") + orb_last_error());
  cached = voc;
  return handle;
}

void transform_on_device(ORBVocabulary* voc, const cv::Mat& descriptors, DBoW2::BowVector& bow, DBoW2::FeatureVector& fv) {
  const int n = descriptors.rows;
  bow.clear(); fv.clear();
  if (n == 0) return;
  cv::Mat d = descriptors.isContinuous() ? descriptors : descriptors.clone();
  std::vector<int32_t> ids(n), nodes(n), ptr(n + 1), idx(n);
  std::vector<double> vals(n);
  int32_t n_words = 0, n_nodes = 0;
  int used;
  {
    // Tracking (Frame::ComputeBoW) and LocalMapping (KeyFrame::ComputeBoW) call this concurrently and the handle
    // owns per-call scratch (stream, device buffers, pinned results): one transform at a time per vocabulary
    static std::mutex call_mu;
    std::lock_guard<std::mutex> lk(call_mu);
    used = bow_transform(device_vocabulary(voc), d.data, n, /*levelsup=*/4, ids.data(), vals.data(), &n_words,
                         nodes.data(), ptr.data(), idx.data(), &n_nodes, n);
  }
  if (used == ORB_E_CAPACITY) {
    // caller-side capacity (cannot happen with cap_words = n; oversized frames such as the 5 x nFeatures of
    // monocular initialisation, Tracking.cc:2536, are handled on the device): use the reference's own transform.
    std::vector<cv::Mat> vCurrentDesc = Converter::toDescriptorVector(descriptors);
    voc->transform(vCurrentDesc, bow, fv, 4);
    return;
  }
  if (used < 0) throw std::runtime_error(std::string("bow_transform: ") + orb_last_error());
  for (int k = 0; k < n_words; k++) bow.insert(bow.end(), DBoW2::BowVector::value_type((DBoW2::WordId)ids[k], vals[k]));
  for (int k = 0; k < n_nodes; k++) {
    auto it = fv.insert(fv.end(), DBoW2::FeatureVector::value_type((DBoW2::NodeId)nodes[k], std::vector<unsigned int>()));
    it->second.assign(idx.begin() + ptr[k], idx.begin() + ptr[k + 1]);
  }
}

}  // namespace

void Frame::ComputeBoW() {
  if (mBowVec.empty()) transform_on_device(mpORBvocabulary, mDescriptors, mBowVec, mFeatVec);  // :740-744
}

void KeyFrame::ComputeBoW() {
  if (mBowVec.empty() || mFeatVec.empty()) transform_on_device(mpORBvocabulary, mDescriptors, mBowVec, mFeatVec);  // :100-106
}

}  // namespace ORB_SLAM3
