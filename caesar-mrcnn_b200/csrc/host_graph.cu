// Host-only helpers of the detection post-processing (no device code): the merge graph of
// Analyzer.extract_det_masks (mrcnn/analyze.py:1258-1320) for a whole batch of frames at once.
// The reference builds one mrcnn/graph.py Graph per image (addEdge for every mergeable pair, in pair-loop order) and
// takes connectedComponents(): components in order of their smallest vertex, members in the pre-order of a recursive
// depth-first walk that follows each adjacency list in insertion order.  This file reproduces exactly that order.
#include <vector>

#include "common.cuh"
#include "mrcnn_b200.h"

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
