// HOST-ONLY: the `vertexes` entry of the source catalogue (mrcnn/analyze.py:1908-1927 make_json_results and
// mrcnn/sfinder.py:885-910 merge_edge_sources): skimage.measure.find_contours(zero_padded_mask, 0.5) with the defaults
// fully_connected='low', positive_orientation='low', vertices flipped to (x, y) in image coordinates.
//
// scikit-image is a third-party dependency of the reference that is absent here; this follows the published 0.15
// algorithm (marching squares in raster order, oriented segments per square case, `_assemble_contours` joining through
// starts / ends dictionaries; the contour created first survives a join; output ordered by creation index) — the same
// restatement as oracle/contours.py, parity UNPINNED.
//
// Input is what the catalogue already holds on the host: each object's pixel list (np.argwhere(mask == 1) + image
// origin, produced on the device by planes_pixels_kernel).  For a binary mask every crossing sits at the middle of a
// square edge, so points are handled as integer doubled coordinates and every vertex is (pixel coordinate +- 0.5),
// exactly representable: the result equals the reference's float64 arithmetic bit for bit.
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <vector>
#include "host_pool.h"
#include "mrcnn_b200.h"

void mrcnn_set_error(const char* fmt, ...);

namespace {

struct Result {
  std::vector<double> verts;            // (x, y) pairs
  std::vector<int64_t> contour_off;     // [n_contours + 1] in vertices
  std::vector<int64_t> object_off;      // [n_objects + 1] in contours
};
thread_local Result g_res;

// _assemble_contours with the dictionaries replaced by direct-indexed tables over the doubled-coordinate grid of the
// object's padded bounding box (point key = 2*row * stride + 2*col) and the deques by singly linked point lists
// (append / prepend / concatenate are O(1)); same decisions in the same order as the Python original.
struct Assembler {
  int stride = 0;
  std::vector<int> starts, ends;          // point key -> contour id, -1 = absent
  std::vector<int> touched;               // keys to reset for the next object
  std::vector<int> pt, nxt;               // linked point nodes
  std::vector<int> first, last;           // per contour (index = creation order)
  std::vector<char> alive;

  void reset(int64_t PH, int64_t PW) {
    stride = (int)(2 * PW + 1);
    const size_t need = (size_t)(2 * PH + 1) * (size_t)stride;
    if (starts.size() < need) {
      starts.assign(need, -1);
      ends.assign(need, -1);
    } else {
      for (int k : touched) starts[(size_t)k] = ends[(size_t)k] = -1;
    }
    touched.clear();
    pt.clear(); nxt.clear(); first.clear(); last.clear(); alive.clear();
  }
  int node(int key) {
    pt.push_back(key);
    nxt.push_back(-1);
    return (int)pt.size() - 1;
  }
  void segment(int from, int to) {
    if (from == to) return;
    touched.push_back(from);
    touched.push_back(to);
    const int tail = starts[(size_t)to], head = ends[(size_t)from];
    if (tail >= 0 && head >= 0) {
      if (tail == head) {                       // close the contour
        const int n = node(to);
        nxt[(size_t)last[(size_t)head]] = n;
        last[(size_t)head] = n;
        starts[(size_t)to] = -1;
        ends[(size_t)from] = -1;
      } else {
        // head's points followed by tail's points; the contour created first keeps its identity
        const int tail_last = pt[(size_t)last[(size_t)tail]], head_first = pt[(size_t)first[(size_t)head]];
        nxt[(size_t)last[(size_t)head]] = first[(size_t)tail];
        if (tail > head) {                      // tail was created second: append it to head
          starts[(size_t)to] = -1;
          ends[(size_t)tail_last] = -1;
          alive[(size_t)tail] = 0;
          ends[(size_t)from] = -1;
          last[(size_t)head] = last[(size_t)tail];
          ends[(size_t)tail_last] = head;
        } else {                                // head was created second: prepend it to tail
          starts[(size_t)head_first] = -1;
          ends[(size_t)from] = -1;
          alive[(size_t)head] = 0;
          starts[(size_t)to] = -1;
          first[(size_t)tail] = first[(size_t)head];
          starts[(size_t)head_first] = tail;
        }
      }
    } else if (tail < 0 && head < 0) {
      const int id = (int)first.size();
      const int n1 = node(from), n2 = node(to);
      nxt[(size_t)n1] = n2;
      first.push_back(n1);
      last.push_back(n2);
      alive.push_back(1);
      starts[(size_t)from] = id;
      ends[(size_t)to] = id;
    } else if (tail >= 0) {                     // prepend to the contour that starts at `to`
      const int n = node(from);
      nxt[(size_t)n] = first[(size_t)tail];
      first[(size_t)tail] = n;
      starts[(size_t)to] = -1;
      starts[(size_t)from] = tail;
    } else {                                    // append to the contour that ends at `from`
      const int n = node(to);
      nxt[(size_t)last[(size_t)head]] = n;
      last[(size_t)head] = n;
      ends[(size_t)from] = -1;
      ends[(size_t)to] = head;
    }
  }
};
thread_local Assembler g_as;
thread_local std::vector<uint8_t> g_bm;

void object_contours(const int32_t* px, int64_t n, Result* out) {
  if (n <= 0) return;
  int32_t y0 = px[0], y1 = px[0], x0 = px[1], x1 = px[1];
  for (int64_t i = 1; i < n; ++i) {
    const int32_t y = px[2 * i], x = px[2 * i + 1];
    y0 = y < y0 ? y : y0; y1 = y > y1 ? y : y1;
    x0 = x < x0 ? x : x0; x1 = x > x1 ? x : x1;
  }
  const int64_t h = (int64_t)y1 - y0 + 1, w = (int64_t)x1 - x0 + 1;
  const int64_t PH = h + 2, PW = w + 2;           // zero border of one pixel, as the reference pads the whole mask
  std::vector<uint8_t>& bm = g_bm;
  bm.assign((size_t)(PH * PW), 0);
  for (int64_t i = 0; i < n; ++i) bm[(size_t)(((int64_t)px[2 * i] - y0 + 1) * PW + ((int64_t)px[2 * i + 1] - x0 + 1))] = 1;
  Assembler& as = g_as;
  as.reset(PH, PW);
  const int st = as.stride;
  for (int64_t r0 = 0; r0 + 1 < PH; ++r0) {
    const uint8_t* a = &bm[(size_t)(r0 * PW)];
    const uint8_t* b = a + PW;
    for (int64_t c0 = 0; c0 + 1 < PW; ++c0) {
      const int sq = a[c0] | (a[c0 + 1] << 1) | (b[c0] << 2) | (b[c0 + 1] << 3);
      if (sq == 0 || sq == 15) continue;
      const int top = (int)(2 * r0) * st + (int)(2 * c0 + 1), bottom = (int)(2 * r0 + 2) * st + (int)(2 * c0 + 1);
      const int left = (int)(2 * r0 + 1) * st + (int)(2 * c0), right = (int)(2 * r0 + 1) * st + (int)(2 * c0 + 2);
      switch (sq) {
        case 1: as.segment(top, left); break;
        case 2: as.segment(right, top); break;
        case 3: as.segment(right, left); break;
        case 4: as.segment(left, bottom); break;
        case 5: as.segment(top, bottom); break;
        case 6: as.segment(right, top); as.segment(left, bottom); break;      // fully_connected = 'low'
        case 7: as.segment(right, bottom); break;
        case 8: as.segment(bottom, right); break;
        case 9: as.segment(top, left); as.segment(bottom, right); break;      // fully_connected = 'low'
        case 10: as.segment(bottom, top); break;
        case 11: as.segment(bottom, left); break;
        case 12: as.segment(left, right); break;
        case 13: as.segment(top, right); break;
        case 14: as.segment(left, top); break;
      }
    }
  }
  for (size_t k = 0; k < as.first.size(); ++k) {
    if (!as.alive[k]) continue;
    for (int nd = as.first[k]; nd >= 0; nd = as.nxt[(size_t)nd]) {
      const int key = as.pt[(size_t)nd];
      const double r = (double)(key / st) * 0.5, c = (double)(key % st) * 0.5;
      out->verts.push_back(c - 1.0 + (double)x0);      // np.fliplr(verts) - 1 (+ origin, already part of the pixel list)
      out->verts.push_back(r - 1.0 + (double)y0);
    }
    out->contour_off.push_back((int64_t)(out->verts.size() / 2));
  }
}

}  // namespace

extern "C" int mrcnn_host_contours(const int32_t* pixels_yx, const int64_t* pixel_offsets, int n_objects, int64_t* n_vertices,
                                   int64_t* n_contours) {
  if (n_objects < 0 || (n_objects > 0 && (!pixels_yx || !pixel_offsets)) || !n_vertices || !n_contours) {
    mrcnn_set_error("host_contours: bad arguments");
    return MRCNN_STATUS_INVALID;
  }
  for (int o = 0; o < n_objects; ++o)
    if (pixel_offsets[o + 1] < pixel_offsets[o]) {
      mrcnn_set_error("host_contours: pixel_offsets must be non-decreasing");
      return MRCNN_STATUS_INVALID;
    }
  // objects are independent: chunks of 32 objects are handed out dynamically to the worker pool, each chunk fills its
  // own partial result (offsets relative to the chunk), and the partials are stitched together in object order
  constexpr int kChunk = 32;
  const int n_chunks = (n_objects + kChunk - 1) / kChunk;
  std::vector<Result> part((size_t)n_chunks);
  std::atomic<int> next(0);
  auto work = [&](int) {
    for (;;) {
      const int c = next.fetch_add(1, std::memory_order_relaxed);
      if (c >= n_chunks) break;
      Result& pr = part[(size_t)c];
      pr.contour_off.assign(1, 0);
      pr.object_off.assign(1, 0);
      const int o1 = (c + 1) * kChunk < n_objects ? (c + 1) * kChunk : n_objects;
      for (int o = c * kChunk; o < o1; ++o) {
        object_contours(pixels_yx + 2 * pixel_offsets[o], pixel_offsets[o + 1] - pixel_offsets[o], &pr);
        pr.object_off.push_back((int64_t)pr.contour_off.size() - 1);
      }
    }
  };
  int nt = mrcnn_host::default_threads();
  if (nt > n_chunks) nt = n_chunks;
  if (n_chunks > 0) mrcnn_host::Pool::get().run(nt, work);
  Result& r = g_res;
  r.verts.clear();
  r.contour_off.assign(1, 0);
  r.object_off.assign(1, 0);
  for (const Result& pr : part) {
    const int64_t v0 = (int64_t)(r.verts.size() / 2), c0 = (int64_t)r.contour_off.size() - 1;
    r.verts.insert(r.verts.end(), pr.verts.begin(), pr.verts.end());
    for (size_t k = 1; k < pr.contour_off.size(); ++k) r.contour_off.push_back(v0 + pr.contour_off[k]);
    for (size_t k = 1; k < pr.object_off.size(); ++k) r.object_off.push_back(c0 + pr.object_off[k]);
  }
  *n_vertices = (int64_t)(r.verts.size() / 2);
  *n_contours = (int64_t)r.contour_off.size() - 1;
  return MRCNN_STATUS_OK;
}

extern "C" int mrcnn_host_contours_fetch(double* vertices_xy, int64_t* contour_offsets, int64_t* object_offsets) {
  const Result& r = g_res;
  if (r.contour_off.empty()) {
    mrcnn_set_error("host_contours_fetch: nothing computed on this thread");
    return MRCNN_STATUS_INVALID;
  }
  if (vertices_xy && !r.verts.empty()) memcpy(vertices_xy, r.verts.data(), r.verts.size() * sizeof(double));
  if (contour_offsets) memcpy(contour_offsets, r.contour_off.data(), r.contour_off.size() * sizeof(int64_t));
  if (object_offsets) memcpy(object_offsets, r.object_off.data(), r.object_off.size() * sizeof(int64_t));
  return MRCNN_STATUS_OK;
}
