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
  // two levels per step while all four grandchildren exist (3 independent 16-byte loads in flight)
  while (4 * c + 3 <= len) {
    const uint4 kids = *reinterpret_cast<const uint4*>(&h[2 * c]);
    const uint4 g0 = *reinterpret_cast<const uint4*>(&h[4 * c]);
    const uint4 g1 = *reinterpret_cast<const uint4*>(&h[4 * c + 2]);
    const bool right = !(__uint_as_float(kids.z) < __uint_as_float(kids.x));
    const uint4 g = right ? g1 : g0;
    const bool gright = !(__uint_as_float(g.z) < __uint_as_float(g.x));
    const int cc = 2 * c + (right ? 1 : 0);
    *reinterpret_cast<uint2*>(&h[c]) = right ? make_uint2(kids.z, kids.w) : make_uint2(kids.x, kids.y);
    *reinterpret_cast<uint2*>(&h[cc]) = gright ? make_uint2(g.z, g.w) : make_uint2(g.x, g.y);
    c = 2 * cc + (gright ? 1 : 0);
  }
  while (2 * c + 1 <= len) {
    const uint4 kids = *reinterpret_cast<const uint4*>(&h[2 * c]);
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
// Block-wide greedy NMS over candidates taken in pop order, 64 at a time:
//   (0) if a heap is given, thread 0 pops the next 64 candidate ids (exact TF order under ties),
//   (1) 64x64 intra-chunk suppression matrix in shared memory (all threads),
//   (2) one thread resolves the chunk greedily against the `removed` bitmap,
//   (3) all threads suppress later candidates against the boxes kept in this chunk: each thread
//       keeps up to NMS_JB of its candidates in registers per pass and streams the kept boxes.
// Equivalent to TF's "pop best, suppress if IoU with any selected box > thr".
// For later positions the heap has not been popped yet, so step (3) works on CANDIDATE IDS in
// sorted order (pos == id when no heap is given): `removed` is indexed by candidate id.
// boxes[id] = candidate box; order[p] = candidate id of pop position p; selected[] receives pop
// positions.  Returns the number selected (<= max_out).
// ---------------------------------------------------------------------------------------------
struct NmsScratch {
  unsigned long long M[64];
  NBox chunk[64];
  float4 kept_box[64];
  float kept_area[64];
  unsigned long long kept_bits;
  int count;
  int heap_len;
};

constexpr int NMS_JB = 3;

__device__ inline int block_nms(const Box4* boxes, uint16_t* order, int n, int max_out, float thr,
                                uint32_t* removed /* ceil(n/32)+2 words, by candidate id */, uint16_t* selected,
                                NmsScratch* sc, HeapEntry* lazy_heap) {
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int lane = tid & 31;
  const int nwords = (n + 31) / 32 + 2;
  for (int i = tid; i < nwords; i += nt) removed[i] = 0;
  if (tid == 0) {
    sc->count = 0;
    sc->heap_len = n;
  }
  __syncthreads();
  int count = 0;
  for (int base = 0; base < n && count < max_out; base += 64) {
    const int n_in = min(64, n - base);
    if (lazy_heap != nullptr) {
      if (tid == 0) {
        int len = sc->heap_len;
        for (int k = 0; k < n_in; ++k) order[base + k] = (uint16_t)heap_pop(lazy_heap, len);
        sc->heap_len = len;
      }
      __syncthreads();
    }
    int my_id = -1;
    if (tid < 64) {
      NBox nb;
      if (tid < n_in) {
        my_id = order[base + tid];
        nb = normalise_box(boxes[my_id]);
      } else {
        nb.ymin = nb.xmin = nb.ymax = nb.xmax = 0.f;
        nb.area = -1.f;
      }
      sc->chunk[tid] = nb;
      sc->M[tid] = 0ull;
    }
    __syncthreads();
    for (int p = tid; p < 64 * 64; p += nt) {
      const int i = p >> 6, j = p & 63;
      if (j > i && j < n_in && iou_gt(sc->chunk[i], sc->chunk[j], thr)) atomicOr(&sc->M[i], 1ull << j);
    }
    __syncthreads();
    if (tid < 32) {
      // alive mask of this chunk: bit k = candidate order[base+k] not yet removed
      unsigned long long alive = 0ull;
      for (int k = lane; k < n_in; k += 32) {
        const int id = order[base + k];
        if (!((removed[id >> 5] >> (id & 31)) & 1u)) alive |= 1ull << k;
      }
      alive |= __shfl_xor_sync(0xffffffffu, alive, 16);
      alive |= __shfl_xor_sync(0xffffffffu, alive, 8);
      alive |= __shfl_xor_sync(0xffffffffu, alive, 4);
      alive |= __shfl_xor_sync(0xffffffffu, alive, 2);
      alive |= __shfl_xor_sync(0xffffffffu, alive, 1);
      if (lane == 0) {
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
    __syncthreads();
    const unsigned long long kept = sc->kept_bits;
    count = sc->count;
    const int nk = __popcll(kept);
    if (tid < 64) {
      // the chunk's own candidates are done: mark them removed so step (3) of later chunks skips them
      if (my_id >= 0) atomicOr(&removed[my_id >> 5], 1u << (my_id & 31));
      if ((kept >> tid) & 1ull) {
        const int r = __popcll(kept & ((1ull << tid) - 1ull));
        const NBox nb = sc->chunk[tid];
        sc->kept_box[r] = make_float4(nb.ymin, nb.xmin, nb.ymax, nb.xmax);
        sc->kept_area[r] = nb.area;
      }
    }
    __syncthreads();
    if (nk == 0 || count >= max_out || base + 64 >= n) continue;
    // (3) suppress every not-yet-removed candidate id against the nk boxes kept in this chunk
    for (int q0 = 0; q0 < n; q0 += nt * NMS_JB) {
      NBox cand[NMS_JB];
      bool live[NMS_JB], live0[NMS_JB];
      bool any = false;
#pragma unroll
      for (int j = 0; j < NMS_JB; ++j) {
        const int q = q0 + j * nt + tid;
        live[j] = false;
        if (q < n && !((removed[q >> 5] >> (q & 31)) & 1u)) {
          cand[j] = normalise_box(boxes[q]);
          live[j] = cand[j].area > 0.f;      // zero-area boxes are never suppressed
        }
        live0[j] = live[j];
        any |= live[j];
      }
      if (__any_sync(0xffffffffu, any)) {
        for (int r = 0; r < nk; ++r) {
          const float4 kb = sc->kept_box[r];
          const float ka = sc->kept_area[r];
          if (!(ka > 0.f)) continue;
#pragma unroll
          for (int j = 0; j < NMS_JB; ++j) {
            if (live[j]) {
              const float ih = __fsub_rn(fminf(kb.z, cand[j].ymax), fmaxf(kb.x, cand[j].ymin));
              const float iw = __fsub_rn(fminf(kb.w, cand[j].xmax), fmaxf(kb.y, cand[j].xmin));
              if (ih > 0.f && iw > 0.f) {
                const float inter = __fmul_rn(ih, iw);
                if (inter != 0.0f) {
                  const float uni = __fsub_rn(__fadd_rn(ka, cand[j].area), inter);
                  if (__fdiv_rn(inter, uni) > thr) live[j] = false;   // suppressed
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NMS_JB; ++j) {
        const int qw = q0 + j * nt + (tid & ~31);          // first candidate id of this warp's word
        const unsigned bal = __ballot_sync(0xffffffffu, live0[j] && !live[j]);
        if (lane == 0 && bal) removed[qw >> 5] |= bal;      // this warp owns word qw>>5 in this pass
      }
    }
    __syncthreads();
  }
  return count;
}
