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
  float ih = fmaxf(__fsub_rn(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)), 0.0f);
  float iw = fmaxf(__fsub_rn(fminf(a.xmax, b.xmax), fmaxf(a.xmin, b.xmin)), 0.0f);
  float inter = __fmul_rn(ih, iw);
  float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
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
// std::push_heap / std::pop_heap produce).  Single thread, entries = (score, id) in shared memory.
// ---------------------------------------------------------------------------------------------
struct HeapEntry {
  float score;
  int id;
};

__device__ inline void heap_pop_order_serial(HeapEntry* heap, int n, bool presorted_desc,
                                             uint16_t* order) {
  // push phase: priority_queue::emplace -> push_back + std::push_heap
  if (!presorted_desc) {
    for (int i = 1; i < n; ++i) {
      HeapEntry v = heap[i];
      int hole = i;
      int parent = (hole - 1) / 2;
      while (hole > 0 && heap[parent].score < v.score) {
        heap[hole] = heap[parent];
        hole = parent;
        parent = (hole - 1) / 2;
      }
      heap[hole] = v;
    }
  }  // else: pushing non-increasing scores never sifts up -> array already is the heap
  int len = n;
  int out = 0;
  while (len > 0) {
    order[out++] = (uint16_t)heap[0].id;
    // std::pop_heap: value = last element, hole at the root sinks to a leaf, value sifts up
    HeapEntry v = heap[len - 1];
    len -= 1;
    if (len == 0) break;
    int hole = 0;
    int child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (heap[child].score < heap[child - 1].score) child--;
      heap[hole] = heap[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      heap[hole] = heap[child - 1];
      hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > 0 && heap[parent].score < v.score) {
      heap[hole] = heap[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    heap[hole] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Block-wide greedy NMS over candidates taken in `order` (pop order), 64 at a time:
//   (1) 64x64 intra-chunk suppression matrix in shared memory (all threads),
//   (2) one thread resolves the chunk greedily against the `removed` bitmap,
//   (3) all threads suppress later candidates against the boxes kept in this chunk.
// Equivalent to TF's "pop best, suppress if IoU with any selected box > thr".
// boxes: candidate boxes in shared memory indexed by candidate id; order[p] = candidate id of
// pop position p.  selected[] receives pop positions.  Returns the number selected (<= max_out).
// ---------------------------------------------------------------------------------------------
struct NmsScratch {
  unsigned long long M[64];
  NBox chunk[64];
  NBox kept[64];
  unsigned long long kept_bits;
  int count;
};

__device__ inline int block_nms(const Box4* boxes, const uint16_t* order, int n, int max_out,
                                float thr, uint32_t* removed /* ceil(n/32)+2 words */,
                                uint16_t* selected, NmsScratch* sc) {
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nwarps = nt >> 5;
  const int nwords = (n + 31) / 32 + 2;
  for (int i = tid; i < nwords; i += nt) removed[i] = 0;
  if (tid == 0) sc->count = 0;
  __syncthreads();
  int count = 0;
  for (int base = 0; base < n && count < max_out; base += 64) {
    const int n_in = min(64, n - base);
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
    }
    __syncthreads();
    for (int p = tid; p < 64 * 64; p += nt) {
      int i = p >> 6, j = p & 63;
      if (j > i && j < n_in && iou_gt(sc->chunk[i], sc->chunk[j], thr))
        atomicOr(&sc->M[i], 1ull << j);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long rem = (unsigned long long)removed[base >> 5] |
                               ((unsigned long long)removed[(base >> 5) + 1] << 32);
      unsigned long long alive = ~rem;
      if (n_in < 64) alive &= ((1ull << n_in) - 1ull);
      unsigned long long kept = 0ull;
      int c = sc->count;
      while (alive && c < max_out) {
        int i = __ffsll((long long)alive) - 1;
        alive &= ~(1ull << i);
        kept |= 1ull << i;
        selected[c++] = (uint16_t)(base + i);
        alive &= ~sc->M[i];
      }
      sc->kept_bits = kept;
      sc->count = c;
    }
    __syncthreads();
    const unsigned long long kept = sc->kept_bits;
    count = sc->count;
    const int nk = __popcll(kept);
    if (nk == 0 || count >= max_out || base + 64 >= n) {
      __syncthreads();
      continue;
    }
    if (tid < 64 && ((kept >> tid) & 1ull)) {
      int r = __popcll(kept & ((1ull << tid) - 1ull));
      sc->kept[r] = sc->chunk[tid];
    }
    __syncthreads();
    for (int q0 = base + 64 + warp * 32; q0 < n; q0 += nwarps * 32) {
      const int q = q0 + lane;
      bool sup = false;
      if (q < n && !((removed[q >> 5] >> (q & 31)) & 1u)) {
        NBox b = normalise_box(boxes[order[q]]);
        if (b.area > 0.f) {
          for (int r = 0; r < nk; ++r) {
            if (iou_gt(sc->kept[r], b, thr)) {
              sup = true;
              break;
            }
          }
        }
      }
      unsigned bal = __ballot_sync(0xffffffffu, sup);
      if (lane == 0 && bal) removed[q0 >> 5] |= bal;  // this warp owns word q0>>5 in this pass
    }
    __syncthreads();
  }
  return count;
}
