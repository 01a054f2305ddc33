// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the
// product path (caesar-mrcnn_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// CPU restatement of tf.image.non_max_suppression as TensorFlow 1.13 implements it
// (NonMaxSuppressionV3, core/kernels/non_max_suppression_op.cc — third-party, NOT under
// /root/reference; pinned by requirements.txt:19 `tensorflow==1.13.2`).  The reference calls it at
//   mrcnn/model.py:392-395  (ProposalLayer, thr = RPN_NMS_THRESHOLD, max = proposal_count)
//   mrcnn/model.py:826-830  (refine_detections_graph per-class NMS, thr = DETECTION_NMS_THRESHOLD)
//
// Semantics restated:
//   * candidates with score > -inf are pushed, in index order, into a
//     std::priority_queue<Candidate, std::deque<Candidate>, cmp(score <)>  — NO index tie-break
//     in 1.13, so the pop order among equal scores is whatever libstdc++'s heap produces; we use
//     the very same container here so the order is reproduced by construction.
//   * pop the best candidate; it is suppressed iff IoU(candidate, s) > thr for some already
//     selected s (scanned newest -> oldest); boxes with area <= 0 have IoU 0 with everything.
//   * stop when max_out boxes are selected.
//   * all arithmetic in float32, no FMA contraction (compile with -ffp-contract=off).
//
// parity: UNPINNED against TensorFlow itself (TF cannot run in this image); pinned only against
// torchvision.ops.nms where the two rules coincide (tests/test_oracle.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <queue>
#include <vector>

namespace {

struct Candidate {
  int box_index;
  float score;
};

inline float iou_f32(const float* boxes, int i, int j) {
  const float* a = boxes + 4 * i;
  const float* b = boxes + 4 * j;
  const float ymin_i = std::min<float>(a[0], a[2]);
  const float xmin_i = std::min<float>(a[1], a[3]);
  const float ymax_i = std::max<float>(a[0], a[2]);
  const float xmax_i = std::max<float>(a[1], a[3]);
  const float ymin_j = std::min<float>(b[0], b[2]);
  const float xmin_j = std::min<float>(b[1], b[3]);
  const float ymax_j = std::max<float>(b[0], b[2]);
  const float xmax_j = std::max<float>(b[1], b[3]);
  const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
  const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
  if (area_i <= 0 || area_j <= 0) return 0.0f;
  const float iymin = std::max<float>(ymin_i, ymin_j);
  const float ixmin = std::max<float>(xmin_i, xmin_j);
  const float iymax = std::min<float>(ymax_i, ymax_j);
  const float ixmax = std::min<float>(xmax_i, xmax_j);
  const float inter = std::max<float>(iymax - iymin, 0.0f) * std::max<float>(ixmax - ixmin, 0.0f);
  return inter / (area_i + area_j - inter);
}

}  // namespace

extern "C" {

// boxes [n,4] (y1,x1,y2,x2) float32, scores [n] float32.  Writes up to max_out selected indices
// (selection order) into `selected`, returns the count.  If pop_order != nullptr it receives the
// full sequence of popped candidate indices (length returned through n_popped) — the order the
// device kernels must reproduce.
int oracle_nms_tf113(const float* boxes, const float* scores, int n, int max_out, float iou_thr,
                     int32_t* selected, int32_t* pop_order, int32_t* n_popped) {
  auto cmp = [](const Candidate a, const Candidate b) { return a.score < b.score; };
  std::priority_queue<Candidate, std::deque<Candidate>, decltype(cmp)> pq(cmp);
  for (int i = 0; i < n; ++i) {
    if (scores[i] > -std::numeric_limits<float>::infinity()) pq.emplace(Candidate{i, scores[i]});
  }
  std::vector<int> sel;
  int popped = 0;
  while (static_cast<int>(sel.size()) < max_out && !pq.empty()) {
    Candidate c = pq.top();
    pq.pop();
    if (pop_order) pop_order[popped] = c.box_index;
    ++popped;
    bool keep = true;
    for (int j = static_cast<int>(sel.size()) - 1; j >= 0; --j) {
      if (iou_f32(boxes, c.box_index, sel[j]) > iou_thr) {
        keep = false;
        break;
      }
    }
    if (keep) sel.push_back(c.box_index);
  }
  if (n_popped) *n_popped = popped;
  for (size_t k = 0; k < sel.size(); ++k) selected[k] = sel[k];
  return static_cast<int>(sel.size());
}

// Full pop order of the heap (no NMS) — used to test the device heap emulation on tied scores.
void oracle_heap_pop_order(const float* scores, int n, int32_t* order) {
  auto cmp = [](const Candidate a, const Candidate b) { return a.score < b.score; };
  std::priority_queue<Candidate, std::deque<Candidate>, decltype(cmp)> pq(cmp);
  for (int i = 0; i < n; ++i) pq.emplace(Candidate{i, scores[i]});
  int k = 0;
  while (!pq.empty()) {
    order[k++] = pq.top().box_index;
    pq.pop();
  }
}

// Pairwise IoU, same float32 routine as above (for property tests).
float oracle_iou(const float* boxes, int i, int j) { return iou_f32(boxes, i, j); }

}  // extern "C"
