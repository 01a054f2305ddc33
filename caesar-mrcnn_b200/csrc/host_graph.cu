// Host-only helpers of the detection post-processing (no device code): the merge graph of
// Analyzer.extract_det_masks (mrcnn/analyze.py:1258-1320) for a whole batch of frames at once.
// The reference builds one mrcnn/graph.py Graph per image (addEdge for every mergeable pair, in pair-loop order) and
// takes connectedComponents(): components in order of their smallest vertex, members in the pre-order of a recursive
// depth-first walk that follows each adjacency list in insertion order.  This file reproduces exactly that order.
#include <vector>

#include "common.cuh"
#include "mrcnn_b200.h"

// Pairs (i < j) inside every frame, frame-major, in the reference's double-loop order (analyze.py:1263-1266, 1335-1338).
extern "C" int mrcnn_host_all_pairs(int n_frames, const int32_t* counts, int32_t* pairs) {
  MRCNN_REQUIRE(n_frames >= 0 && (n_frames == 0 || counts), "host_all_pairs: bad frame list");
  int64_t base = 0, out = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int n = counts[f];
    MRCNN_REQUIRE(n >= 0, "host_all_pairs: negative frame size");
    MRCNN_REQUIRE(pairs || n < 2, "host_all_pairs: null output");
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        pairs[out++] = (int32_t)(base + i);
        pairs[out++] = (int32_t)(base + j);
      }
    base += n;
  }
  return MRCNN_OK;
}

// sklearn jaccard_score(average='binary') as the reference calls it (analyze.py:1281-1284): tp / (tp + fp + fn) in
// float64, 0.0 for an empty union
static inline double pair_iou(int32_t inter, int32_t area_a, int32_t area_b) {
  const int64_t uni = (int64_t)area_a + (int64_t)area_b - (int64_t)inter;
  return uni > 0 ? (double)inter / (double)uni : 0.0;
}

// Pair tests of both graph stages of extract_det_masks for a batch (pairs implicit: mrcnn_host_all_pairs order).
//  stage 0 (merge, analyze.py:1268-1290): flag = connected && same class && iou >= thr
//  stage 1 (selection, analyze.py:1340-1360): flag = connected && !(split_source_sidelobe && exactly one of the two is
//           'spurious' && iou < thr); additionally loses[v] = 1 if a linked mask has a strictly higher score, and
//           tie_frame[f] = 1 if two linked masks of frame f have equal scores (the clique order then decides)
extern "C" int mrcnn_host_pair_flags(int stage, int n_frames, const int32_t* counts, const int32_t* cls_or_spurious,
                                     const float* score, const int32_t* area, const int32_t* inter, const int32_t* touch,
                                     int use_iou, double iou_thr, uint8_t* flags, uint8_t* loses, uint8_t* tie_frame) {
  MRCNN_REQUIRE(stage == 0 || stage == 1, "host_pair_flags: stage");
  MRCNN_REQUIRE(n_frames >= 0 && (n_frames == 0 || counts), "host_pair_flags: bad frame list");
  int64_t base = 0, p = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int n = counts[f];
    MRCNN_REQUIRE(n >= 0, "host_pair_flags: negative frame size");
    if (stage == 1 && tie_frame) tie_frame[f] = 0;
    if (n >= 2) MRCNN_REQUIRE(cls_or_spurious && area && inter && touch && flags, "host_pair_flags: null pointer");
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j, ++p) {
        const int64_t a = base + i, b = base + j;
        bool on = touch[p] != 0;
        if (stage == 0) {
          on = on && cls_or_spurious[a] == cls_or_spurious[b] && pair_iou(inter[p], area[a], area[b]) >= iou_thr;
        } else {
          if (on && use_iou && (cls_or_spurious[a] != 0) != (cls_or_spurious[b] != 0) &&
              pair_iou(inter[p], area[a], area[b]) < iou_thr)
            on = false;
          if (on && score && loses) {
            if (score[a] < score[b]) loses[a] = 1;
            else if (score[b] < score[a]) loses[b] = 1;
            else if (score[a] == score[b] && tie_frame) tie_frame[f] = 1;
          }
        }
        flags[p] = on ? 1 : 0;
      }
    base += n;
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_host_merge_components(int n_frames, const int32_t* det_count, const int32_t* pairs,
                                           const uint8_t* mergeable, int64_t n_pairs, int32_t* members, int32_t* offsets,
                                           int32_t* frame_components, int32_t* n_components) {
  MRCNN_REQUIRE(n_frames >= 0 && n_pairs >= 0, "host_merge_components: negative count");
  MRCNN_REQUIRE(det_count && members && offsets && frame_components && n_components, "host_merge_components: null pointer");
  MRCNN_REQUIRE(n_pairs == 0 || (pairs && mergeable), "host_merge_components: null pair arrays");
  int64_t total = 0;
  for (int f = 0; f < n_frames; ++f) {
    MRCNN_REQUIRE(det_count[f] >= 0, "host_merge_components: negative frame size");
    total += det_count[f];
  }
  MRCNN_REQUIRE(total < (1ll << 31), "host_merge_components: too many masks");
  const int n = (int)total;
  // adjacency lists in insertion order: CSR built in two passes over the mergeable pairs (addEdge(v,w) appends w to
  // adj[v] and v to adj[w])
  std::vector<int32_t> deg(n + 1, 0);
  for (int64_t p = 0; p < n_pairs; ++p) {
    if (!mergeable[p]) continue;
    const int a = pairs[2 * p], b = pairs[2 * p + 1];
    MRCNN_REQUIRE(a >= 0 && a < n && b >= 0 && b < n, "host_merge_components: pair index out of range");
    ++deg[a + 1];
    ++deg[b + 1];
  }
  for (int v = 0; v < n; ++v) deg[v + 1] += deg[v];
  std::vector<int32_t> adj(deg[n]), fill(deg.begin(), deg.end() - 1);
  for (int64_t p = 0; p < n_pairs; ++p) {
    if (!mergeable[p]) continue;
    const int a = pairs[2 * p], b = pairs[2 * p + 1];
    adj[fill[a]++] = b;
    adj[fill[b]++] = a;
  }
  std::vector<uint8_t> seen(n, 0);
  std::vector<int32_t> stack_node, stack_next;
  int out = 0, comps = 0, base = 0;
  offsets[0] = 0;
  for (int f = 0; f < n_frames; ++f) {
    int in_frame = 0;
    for (int v = base; v < base + det_count[f]; ++v) {
      if (seen[v]) continue;
      // iterative form of the recursive walk: visit a vertex when it is first reached, resume its parent's list after
      seen[v] = 1;
      members[out++] = v;
      stack_node.assign(1, v);
      stack_next.assign(1, deg[v]);
      while (!stack_node.empty()) {
        const int u = stack_node.back();
        int& k = stack_next.back();
        bool descended = false;
        while (k < deg[u + 1]) {
          const int w = adj[k++];
          if (!seen[w]) {
            seen[w] = 1;
            members[out++] = w;
            stack_node.push_back(w);
            stack_next.push_back(deg[w]);
            descended = true;
            break;
          }
        }
        if (!descended) {
          stack_node.pop_back();
          stack_next.pop_back();
        }
      }
      offsets[++comps] = out;
      ++in_frame;
    }
    frame_components[f] = in_frame;
    base += det_count[f];
  }
  *n_components = comps;
  return MRCNN_OK;
}
