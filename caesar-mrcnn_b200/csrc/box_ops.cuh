// Bit-exact float32 box arithmetic shared by the ProposalLayer and DetectionLayer kernels.
// Every reference op is one IEEE round-to-nearest operation (explicit __f*_rn intrinsics, so the
// compiler can never contract a multiply-add), exp is "double exp, one rounding to float" — the
// convention the oracle uses (oracle/graph_layers.py: exp_f32).
#pragma once
#include <limits.h>
#include "common.cuh"

struct Box4 {
  float y1, x1, y2, x2;
};

// Box normalised for IoU: min/max corners + area (TF NonMaxSuppression IOU(), float32).
struct NBox {
  float ymin, xmin, ymax, xmax, area;
};

__device__ __forceinline__ float exp_f32_exact(float x) { return (float)exp((double)x); }

// mrcnn/model.py:287-308 apply_box_deltas_graph (deltas already multiplied by the std-dev)
__device__ __forceinline__ Box4 apply_box_deltas(Box4 b, float dy, float dx, float dh, float dw) {
  float h = __fsub_rn(b.y2, b.y1);
  float w = __fsub_rn(b.x2, b.x1);
  float cy = __fadd_rn(b.y1, __fmul_rn(0.5f, h));
  float cx = __fadd_rn(b.x1, __fmul_rn(0.5f, w));
  cy = __fadd_rn(cy, __fmul_rn(dy, h));
  cx = __fadd_rn(cx, __fmul_rn(dx, w));
  h = __fmul_rn(h, exp_f32_exact(dh));
  w = __fmul_rn(w, exp_f32_exact(dw));
  Box4 r;
  r.y1 = __fsub_rn(cy, __fmul_rn(0.5f, h));
  r.x1 = __fsub_rn(cx, __fmul_rn(0.5f, w));
  r.y2 = __fadd_rn(r.y1, h);
  r.x2 = __fadd_rn(r.x1, w);
  return r;
}

// mrcnn/model.py:311-326 clip_boxes_graph: max(min(v, hi), lo)
__device__ __forceinline__ Box4 clip_box(Box4 b, float wy1, float wx1, float wy2, float wx2) {
  Box4 r;
  r.y1 = fmaxf(fminf(b.y1, wy2), wy1);
  r.x1 = fmaxf(fminf(b.x1, wx2), wx1);
  r.y2 = fmaxf(fminf(b.y2, wy2), wy1);
  r.x2 = fmaxf(fminf(b.x2, wx2), wx1);
  return r;
}

__device__ __forceinline__ NBox normalise_box(Box4 b) {
  NBox n;
  n.ymin = fminf(b.y1, b.y2);
  n.xmin = fminf(b.x1, b.x2);
  n.ymax = fmaxf(b.y1, b.y2);
  n.xmax = fmaxf(b.x1, b.x2);
  n.area = __fmul_rn(__fsub_rn(n.ymax, n.ymin), __fsub_rn(n.xmax, n.xmin));
  return n;
}

// TF 1.13 IOUGreaterThanThreshold: IoU(a,b) > thr with area<=0 -> IoU 0, float32, strict >.
__device__ __forceinline__ bool iou_gt(const NBox& a, const NBox& b, float thr) {
  if (!(a.area > 0.f) || !(b.area > 0.f)) return 0.0f > thr;
  const float ih = fmaxf(__fsub_rn(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)), 0.0f);
  const float iw = fmaxf(__fsub_rn(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin)), 0.0f);
  const float inter = __fmul_rn(ih, iw);
  // disjoint boxes (the common case): 0 / (positive) = 0 exactly, no division needed
  if (inter == 0.0f) return 0.0f > thr;
  const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
  return __fdiv_rn(inter, uni) > thr;
}

// float -> uint32 key whose unsigned order equals the float order
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------
// Exact emulation of libstdc++'s std::priority_queue<Candidate, deque, score<> pop order
// (TF 1.13 NMS has no index tie-break; the order among equal scores is whatever
// std::push_heap / std::pop_heap produce — and with 6000 float32 scores per image at least one
// exact tie is the norm, not the exception).  The heap is kept 1-indexed in shared memory
// (h[1] = root, children of c at 2c / 2c+1, so a child pair is one aligned 16-byte load) and is
// popped LAZILY, 64 candidates at a time, only as far as the NMS actually consumes candidates.
// ---------------------------------------------------------------------------------------------
struct __align__(8) HeapEntry {
  float score;
  int id;
};

// priority_queue::emplace in input order: push_back + std::push_heap (only needed when the input is
// not already sorted non-increasing; pushing non-increasing scores never sifts up).
__device__ inline void heap_push_all_serial(HeapEntry* h, int n) {
  for (int i = 2; i <= n; ++i) {
    const HeapEntry v = h[i];
    int c = i;
    while (c > 1 && h[c >> 1].score < v.score) {
      h[c] = h[c >> 1];
      c >>= 1;
    }
    h[c] = v;
  }
}

__device__ __forceinline__ uint4 lds128(const void* p) {
  // volatile asm: the loads of one look-ahead step are issued back to back, never sunk behind the
  // data-dependent selects (the compiler otherwise predicates them and serialises the latencies)
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

// priority_queue::top + pop: std::pop_heap (__adjust_heap: the hole sinks to a leaf always taking
// the right child unless it is strictly smaller, then the former last element sifts up) + pop_back.
__device__ inline int heap_pop(HeapEntry* h, int& len_ref) {
  int len = len_ref;
  const int top_id = h[1].id;
  const HeapEntry v = h[len];
  len -= 1;
  len_ref = len;
  if (len == 0) return top_id;
  int c = 1;
  // three levels per step while all eight great-grandchildren exist (7 independent 16-byte loads)
  while (8 * c + 7 <= len) {
    const uint4 k0 = lds128(&h[2 * c]);
    const uint4 g0 = lds128(&h[4 * c]);
    const uint4 g1 = lds128(&h[4 * c + 2]);
    const uint4 q0 = lds128(&h[8 * c]);
    const uint4 q1 = lds128(&h[8 * c + 2]);
    const uint4 q2 = lds128(&h[8 * c + 4]);
    const uint4 q3 = lds128(&h[8 * c + 6]);
    const bool r1 = !(__uint_as_float(k0.z) < __uint_as_float(k0.x));
    const uint4 g = r1 ? g1 : g0;
    const bool r2 = !(__uint_as_float(g.z) < __uint_as_float(g.x));
    const uint4 qa = r2 ? q1 : q0, qb = r2 ? q3 : q2;
    const uint4 q = r1 ? qb : qa;
    const bool r3 = !(__uint_as_float(q.z) < __uint_as_float(q.x));
    const int c1 = 2 * c + (r1 ? 1 : 0);
    const int c2 = 2 * c1 + (r2 ? 1 : 0);
    *reinterpret_cast<uint2*>(&h[c]) = r1 ? make_uint2(k0.z, k0.w) : make_uint2(k0.x, k0.y);
    *reinterpret_cast<uint2*>(&h[c1]) = r2 ? make_uint2(g.z, g.w) : make_uint2(g.x, g.y);
    *reinterpret_cast<uint2*>(&h[c2]) = r3 ? make_uint2(q.z, q.w) : make_uint2(q.x, q.y);
    c = 2 * c2 + (r3 ? 1 : 0);
  }
  while (4 * c + 3 <= len) {
    const uint4 k0 = lds128(&h[2 * c]);
    const uint4 g0 = lds128(&h[4 * c]);
    const uint4 g1 = lds128(&h[4 * c + 2]);
    const bool r1 = !(__uint_as_float(k0.z) < __uint_as_float(k0.x));
    const uint4 g = r1 ? g1 : g0;
    const bool r2 = !(__uint_as_float(g.z) < __uint_as_float(g.x));
    const int c1 = 2 * c + (r1 ? 1 : 0);
    *reinterpret_cast<uint2*>(&h[c]) = r1 ? make_uint2(k0.z, k0.w) : make_uint2(k0.x, k0.y);
    *reinterpret_cast<uint2*>(&h[c1]) = r2 ? make_uint2(g.z, g.w) : make_uint2(g.x, g.y);
    c = 2 * c1 + (r2 ? 1 : 0);
  }
  while (2 * c + 1 <= len) {
    const uint4 kids = lds128(&h[2 * c]);
    const bool right = !(__uint_as_float(kids.z) < __uint_as_float(kids.x));
    *reinterpret_cast<uint2*>(&h[c]) = right ? make_uint2(kids.z, kids.w) : make_uint2(kids.x, kids.y);
    c = 2 * c + (right ? 1 : 0);
  }
  if (2 * c == len) {   // lone left child at the end of the array
    h[c] = h[2 * c];
    c = 2 * c;
  }
  while (c > 1 && h[c >> 1].score < v.score) {
    h[c] = h[c >> 1];
    c >>= 1;
  }
  h[c] = v;
  return top_id;
}

// ---------------------------------------------------------------------------------------------
// Block-wide greedy NMS over candidates taken in pop order, 64 at a time, LAZILY: a chunk's 64
// candidates are tested against the boxes selected so far (kept list in shared memory) only when
// the chunk is reached, so the work is ~64 x |selected| per chunk and stops with the NMS itself
// (TF stops at max_out picks) instead of pre-suppressing all n candidates for every pick.
//   (0) if a heap is given, the popper warp pops the NEXT chunk's 64 candidate ids (exact TF order under ties)
//       while the workers do (1)-(3) on the current one,
//   (1) workers: candidate k (15 threads each) vs the kept list -> suppressed flags, and the
//       64x64 intra-chunk suppression matrix,
//   (2) warp 0 gathers the alive mask; lane 0 resolves the chunk greedily,
//   (3) the newly kept boxes are appended to the kept list.
// Equivalent to TF's "pop best, suppress if IoU with any selected box > thr".
// boxes[id] = candidate box; order[p] = candidate id of pop position p; selected[] receives pop
// positions.  kept_box / kept_area: shared arrays of max_out entries.  Returns the number selected.
// Requires blockDim.x == 1024.
// ---------------------------------------------------------------------------------------------
struct NmsScratch {
  unsigned long long M[64];
  NBox chunk[64];
  unsigned char sup[64];
  unsigned long long kept_bits;
  int count;
  int heap_len;
};

// Warp 31 is the POPPER: while warps 0..30 (992 threads, named barrier 1) resolve chunk c, its lane 0 pops the
// candidate ids of chunk c+1 from the emulated heap (the pops do not depend on the NMS outcome), so the serial
// heap emulation and the NMS overlap instead of alternating; one __syncthreads per chunk hands over.  The chunk
// popped after the last one consumed is wasted work only.
constexpr int NMS_WORKERS = 992;
__device__ __forceinline__ void nms_workers_sync() { asm volatile("bar.sync 1, 992;" ::: "memory"); }

__device__ inline int block_nms(const Box4* boxes, uint16_t* order, int n, int max_out, float thr,
                                float4* kept_box, float* kept_area, uint16_t* selected, NmsScratch* sc,
                                HeapEntry* lazy_heap) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool worker = tid < NMS_WORKERS;
  auto pop_chunk = [&](int base) {     // thread NMS_WORKERS only
    int len = sc->heap_len;
    const int cnt = min(64, n - base);
    for (int k = 0; k < cnt; ++k) order[base + k] = (uint16_t)heap_pop(lazy_heap, len);
    sc->heap_len = len;
  };
  if (tid == 0) {
    sc->count = 0;
    sc->heap_len = n;
  }
  __syncthreads();
  if (lazy_heap != nullptr && tid == NMS_WORKERS && n > 0) pop_chunk(0);
  __syncthreads();
  int count = 0;
  for (int base = 0; base < n && count < max_out; base += 64) {
    const int n_in = min(64, n - base);
    if (!worker) {
      if (lazy_heap != nullptr && tid == NMS_WORKERS && base + 64 < n) pop_chunk(base + 64);
    } else {
      if (tid < 64) {
        NBox nb;
        if (tid < n_in) {
          nb = normalise_box(boxes[order[base + tid]]);
        } else {
          nb.ymin = nb.xmin = nb.ymax = nb.xmax = 0.f;
          nb.area = -1.f;
        }
        sc->chunk[tid] = nb;
        sc->M[tid] = 0ull;
        sc->sup[tid] = 0;
      }
      nms_workers_sync();
      if (tid < 64 * 15) {   // (1a) candidate k = tid/15 against the kept list, 15 threads striding over it
        const int k = tid / 15, sub = tid - k * 15;
        const NBox cb = sc->chunk[k];
        if (cb.area > 0.f) {
          for (int r = sub; r < count; r += 15) {
            const float4 kb = kept_box[r];
            const float ih = __fsub_rn(fminf(kb.z, cb.ymax), fmaxf(kb.x, cb.ymin));
            const float iw = __fsub_rn(fminf(kb.w, cb.xmax), fmaxf(kb.y, cb.xmin));
            if (ih > 0.f && iw > 0.f) {
              const float ka = kept_area[r];
              const float inter = __fmul_rn(ih, iw);
              if (ka > 0.f && inter != 0.0f) {
                const float uni = __fsub_rn(__fadd_rn(ka, cb.area), inter);
                if (__fdiv_rn(inter, uni) > thr) {
                  sc->sup[k] = 1;      // any of the 15 threads; same value, benign race
                  break;
                }
              }
            }
          }
        }
      }
      // (1b) intra-chunk suppression matrix
      for (int p = tid; p < 64 * 64; p += NMS_WORKERS) {
        const int i = p >> 6, j = p & 63;
        if (j > i && j < n_in && iou_gt(sc->chunk[i], sc->chunk[j], thr)) atomicOr(&sc->M[i], 1ull << j);
      }
      nms_workers_sync();
      if (tid < 32) {   // (2)
        const unsigned lo = __ballot_sync(0xffffffffu, lane < n_in && !sc->sup[lane]);
        const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < n_in && !sc->sup[lane + 32]);
        if (lane == 0) {
          unsigned long long alive = (unsigned long long)lo | ((unsigned long long)hi << 32);
          unsigned long long kept = 0ull;
          int c = sc->count;
          while (alive && c < max_out) {
            const int i = __ffsll((long long)alive) - 1;
            alive &= ~(1ull << i);
            kept |= 1ull << i;
            selected[c++] = (uint16_t)(base + i);
            alive &= ~sc->M[i];
          }
          sc->kept_bits = kept;
          sc->count = c;
        }
      }
      nms_workers_sync();
      const unsigned long long kept = sc->kept_bits;
      if (tid < 64 && ((kept >> tid) & 1ull)) {   // (3) append in selection order
        const int r = count + __popcll(kept & ((1ull << tid) - 1ull));
        const NBox nb = sc->chunk[tid];
        kept_box[r] = make_float4(nb.ymin, nb.xmin, nb.ymax, nb.xmax);
        kept_area[r] = nb.area;
      }
    }
    __syncthreads();          // chunk handover: next order[] popped, kept list and count final
    count = sc->count;
  }
  return count;
}
