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
#include <deque>
#include <unordered_map>
#include <vector>
#include "mrcnn_b200.h"

void mrcnn_set_error(const char* fmt, ...);

namespace {

typedef uint64_t Pt;   // (2*row) << 32 | (2*col) in the zero-padded local frame of the object's bounding box
inline Pt mk(int64_t r2, int64_t c2) { return ((uint64_t)(uint32_t)r2 << 32) | (uint64_t)(uint32_t)c2; }

struct Result {
  std::vector<double> verts;            // (x, y) pairs
  std::vector<int64_t> contour_off;     // [n_contours + 1] in vertices
  std::vector<int64_t> object_off;      // [n_objects + 1] in contours
};
thread_local Result g_res;

struct Assembler {
  std::vector<std::deque<Pt>> contours;   // index = creation number - 1
  std::vector<char> alive;
  std::unordered_map<Pt, int> starts, ends;

  void segment(Pt from, Pt to) {
    if (from == to) return;
    auto ts = starts.find(to);
    auto he = ends.find(from);
    const bool has_tail = ts != starts.end(), has_head = he != ends.end();
    if (has_tail && has_head) {
      const int tail = ts->second, head = he->second;
      if (tail == head) {                       // close the contour
        contours[head].push_back(to);
        starts.erase(to);
        ends.erase(from);
      } else if (tail > head) {                 // tail was created second: append it to head
        std::deque<Pt>& h = contours[head];
        std::deque<Pt>& t = contours[tail];
        h.insert(h.end(), t.begin(), t.end());
        starts.erase(to);
        ends.erase(t.back());
        alive[tail] = 0;
        ends.erase(from);
        ends[h.back()] = head;
        t.clear();
      } else {                                  // head was created second: prepend it to tail
        std::deque<Pt>& h = contours[head];
        std::deque<Pt>& t = contours[tail];
        const Pt head_first = h.front();
        t.insert(t.begin(), h.begin(), h.end());
        starts.erase(head_first);
        ends.erase(from);
        alive[head] = 0;
        starts.erase(to);
        starts[t.front()] = tail;
        h.clear();
      }
    } else if (!has_tail && !has_head) {
      const int id = (int)contours.size();
      contours.emplace_back();
      contours.back().push_back(from);
      contours.back().push_back(to);
      alive.push_back(1);
      starts[from] = id;
      ends[to] = id;
    } else if (has_tail) {                      // prepend to the contour that starts at `to`
      const int tail = ts->second;
      contours[tail].push_front(from);
      starts.erase(to);
      starts[from] = tail;
    } else {                                    // append to the contour that ends at `from`
      const int head = he->second;
      contours[head].push_back(to);
      ends.erase(from);
      ends[to] = head;
    }
  }
};

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
  std::vector<uint8_t> bm((size_t)(PH * PW), 0);
  for (int64_t i = 0; i < n; ++i) bm[(size_t)(((int64_t)px[2 * i] - y0 + 1) * PW + ((int64_t)px[2 * i + 1] - x0 + 1))] = 1;
  Assembler as;
  for (int64_t r0 = 0; r0 + 1 < PH; ++r0) {
    const uint8_t* a = &bm[(size_t)(r0 * PW)];
    const uint8_t* b = a + PW;
    for (int64_t c0 = 0; c0 + 1 < PW; ++c0) {
      const int sq = a[c0] | (a[c0 + 1] << 1) | (b[c0] << 2) | (b[c0 + 1] << 3);
      if (sq == 0 || sq == 15) continue;
      const Pt top = mk(2 * r0, 2 * c0 + 1), bottom = mk(2 * r0 + 2, 2 * c0 + 1);
      const Pt left = mk(2 * r0 + 1, 2 * c0), right = mk(2 * r0 + 1, 2 * c0 + 2);
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
  for (size_t k = 0; k < as.contours.size(); ++k) {
    if (!as.alive[k]) continue;
    for (Pt p : as.contours[k]) {
      const double r = (double)(uint32_t)(p >> 32) * 0.5, c = (double)(uint32_t)(p & 0xffffffffu) * 0.5;
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
  Result& r = g_res;
  r.verts.clear();
  r.contour_off.assign(1, 0);
  r.object_off.assign(1, 0);
  for (int o = 0; o < n_objects; ++o) {
    const int64_t a = pixel_offsets[o], b = pixel_offsets[o + 1];
    if (b < a) {
      mrcnn_set_error("host_contours: pixel_offsets must be non-decreasing");
      return MRCNN_STATUS_INVALID;
    }
    object_contours(pixels_yx + 2 * a, b - a, &r);
    r.object_off.push_back((int64_t)r.contour_off.size() - 1);
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
