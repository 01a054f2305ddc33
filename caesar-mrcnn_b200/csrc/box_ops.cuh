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

__device__ __forceinline__ float key_to_float(uint32_t k) {       // inverse of float_to_key
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
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
// Pipelined exact heap pops (ProposalLayer).  std::pop_heap's bottom-up __adjust_heap (hole to a leaf,
// then the former last element v sifts up) leaves the same array as a TOP-DOWN sift of v that stops at the
// first level whose chosen child (the larger one, the right one on a tie) is strictly smaller than v: the
// chosen children along a path are non-increasing, so "sift up while parent < v" ends exactly there and
// restores every promotion below it.  A top-down pop never re-reads a level it has left, so pop j+1 can
// run two levels behind pop j: in a round every in-flight pop loads the children of its hole, decides,
// and fills the hole; lanes 0..7 of one warp each own one pop (a pop lasts <= 13 rounds for heaps below
// 8192 entries, a new one starts every second round).  What the sequential order fixes and the pipeline
// must re-establish:
//   * v of pop j is h[n-j] AFTER all older pops: an older in-flight pop that ends at that index rewrites
//     it.  A pop therefore re-reads its v source every round; the decisions taken with a stale v were all
//     "continue" (chosen child >= stale v) and stay valid iff the smallest child promoted so far is still
//     >= the new v (`xlast`, checked at every change; a violation raises `hazard`);
//   * a pop may STOP only when no older pop is in flight (its v is final then); otherwise it and every
//     younger pop freeze for the round while the older ones advance (lags only grow);
//   * the last 32 pops run sequentially (for tiny heaps a v source may still have children).
// On `hazard` (never observed; kept for rigour) the heap is rebuilt and popped sequentially from the start;
// entries are published to the consumers only once every older pop has ended hazard-free, so nothing
// already consumed can change.  tests/test_heap_pipeline_sim.py runs the same round-synchronous algorithm
// on the CPU against libstdc++'s std::pop_heap.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 lds64(const void* p) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ int ld_volatile_shared(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

struct PopperShared {
  int popped;    // order[0 .. popped) is final (written by the popper warp, read by the consumers)
  int stop;      // consumers are done: the popper may leave
};

constexpr int POP_TAIL = 32;

// Whole warp (32 lanes converged).  h: valid 1-indexed max-heap of n entries whose initial array is sorted (entry q+1 =
// {sorted_scores[q], q}); order[k] receives the id of the k-th pop.  Lane (j & 7) owns pop j; the pops in flight are
// [jnext-8, jnext), so "the oldest active pop" and "the oldest pop that wants to stop too early" are a rotate + find-first
// on the vote masks (uniform datapath, in the shadow of the loads).  The id of pop j+1 is the entry pop j puts into the
// root, so nobody has to read the root.
__device__ inline void popper_warp(HeapEntry* h, const int n, uint16_t* order, PopperShared* ps, const float* sorted_scores) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int npipe = n > POP_TAIL ? n - POP_TAIL : 0;
  const float INF = __int_as_float(0x7f800000);
  int c = 1, len = 0, vsrc = 1, vid = -1, pop = -1;
  float vs = 0.f, xlast = INF;
  bool active = false, hazard = false;
  int jnext = 0, since = 2, round = 0, published = 0;     // warp-uniform
  unsigned act = 0u;                                       // lanes with a pop in flight
  bool stopped = false;
  if (npipe > 0) {
    if (lane == 0) {                                       // pop 0
      order[0] = (uint16_t)h[1].id;
      len = n - 1;
      vsrc = n;
      pop = 0;
      active = true;
    }
  }
  // ---- fast mode: fixed schedule (pop j starts in round 2j on lane j & 7, a lane is free again after <= 13 rounds),
  // no freeze logic.  Valid while every pop that stops is the oldest one in flight (pops normally end in order: a pop
  // reaches its leaf after the older ones; the oldest is then a counter, not a search); the first stop that comes earlier — a v larger than a child high up,
  // which starts around pop 1200 of 6000 — is not committed: nobody stores in that round and the careful loop below
  // takes over from the same state.
  if (npipe > 0) {
    int r = 0;                                             // rounds done; lane 0 holds pop 0 (started "in round 0")
    int mystart = lane < 8 ? 2 * (lane == 0 ? 8 : lane) : INT_MAX;
    int jo = 0;                                            // oldest pop in flight (every pop below it has ended)
    bool handover = false;
    while (true) {
      if ((r & 15) == 0 && r > 0) {
        if (__any_sync(FULL, hazard)) break;
        const int fin = min(npipe, r >= 12 ? (r - 12) / 2 + 1 : 0);
        if (fin > published) {
          published = fin;
          if (lane == 0) {
            __threadfence_block();
            st_volatile_shared(&ps->popped, fin);
          }
        }
        if (__any_sync(FULL, ld_volatile_shared(&ps->stop) != 0)) {
          stopped = true;
          break;
        }
        if (r >= 2 * npipe + 16) break;                    // every pipelined pop has started and ended
      }
      if (r == mystart) {                                  // my next pop (pops lane, lane+8, ...)
        hazard |= active;
        pop = r >> 1;
        if (pop < npipe) {
          c = 1;
          len = n - pop - 1;
          vsrc = n - pop;
          vid = -1;
          xlast = INF;
          active = true;
        }
        mystart += 16;
      }
      const int l = 2 * c;
      const bool in = active && l <= len;
      const uint4 kids = lds128(&h[in ? l : 2]);
      const uint2 u = lds64(&h[vsrc]);
      if (active && (int)u.y != vid) {
        vs = __uint_as_float(u.x);
        vid = (int)u.y;
        hazard |= xlast < vs;
      }
      const float ls = __uint_as_float(kids.x), rs = __uint_as_float(kids.z);
      const bool right = l < len && !(rs < ls);
      const float xs = right ? rs : ls;
      const int xid = (int)(right ? kids.w : kids.y);
      const bool stop = !in || xs < vs;
      // in this mode pops end in order, so "the oldest pop in flight" is a warp-uniform counter jo: ONE vote per round says
      // who wants to stop; anybody but lane jo & 7 doing so is the early stop that hands over to the careful loop
      const unsigned sb = __ballot_sync(FULL, active && stop);
      const unsigned ob = 1u << (jo & 7);
      if (sb & ~ob) {
        handover = true;
        break;
      }
      jo += (sb & ob) ? 1 : 0;
      if (active) {
        const float ss = stop ? vs : xs;
        const int sid = stop ? vid : xid;
        *reinterpret_cast<uint2*>(&h[c]) = make_uint2(__float_as_uint(ss), (uint32_t)sid);
        if (c == 1) order[pop + 1] = (uint16_t)sid;
        xlast = stop ? xlast : xs;
        c = stop ? c : l + (right ? 1 : 0);
        active = !stop;
      }
      ++r;
      __syncwarp();
    }
    // state for the careful loop: pops [0, jnext) have started, the youngest one `since` rounds ago
    act = __ballot_sync(FULL, active);
    jnext = min(npipe, handover ? (r >> 1) + 1 : npipe);
    since = handover ? (r & 1) : 2;
    if (!handover) act = stopped ? 0u : act;
  }
  while (!stopped && (act != 0u || jnext < npipe)) {
    if ((++round & 15) == 0) {                             // checkpoint: publish, look for the consumers' stop flag
      if (__any_sync(FULL, hazard)) break;
      const unsigned rot = ((act | (act << 8)) >> (jnext & 7)) & 0xffu;
      const int fin = rot ? jnext - 8 + __ffs(rot) : jnext;      // oldest active pop + 1: every older pop has ended
      if (fin > published) {
        published = fin;
        if (lane == 0) {
          __threadfence_block();
          st_volatile_shared(&ps->popped, fin);
        }
      }
      if (__any_sync(FULL, ld_volatile_shared(&ps->stop) != 0)) {
        stopped = true;
        break;
      }
    }
    const int l = 2 * c;
    const bool in = active && l <= len;                    // the hole has at least a left child
    const uint4 kids = lds128(&h[in ? l : 2]);
    const uint2 u = lds64(&h[vsrc]);
    const unsigned rot = ((act | (act << 8)) >> (jnext & 7)) & 0xffu;
    const int minpop = jnext - 9 + __ffs(rot);             // oldest pop in flight (act != 0 whenever a lane is active)
    if (active && (int)u.y != vid) {                       // v source (re)read: first read, or an older pop ended there
      vs = __uint_as_float(u.x);
      vid = (int)u.y;
      hazard |= xlast < vs;
    }
    const float ls = __uint_as_float(kids.x), rs = __uint_as_float(kids.z);
    const bool right = l < len && !(rs < ls);
    const float xs = right ? rs : ls;
    const int xid = (int)(right ? kids.w : kids.y);
    const bool stop = !in || xs < vs;
    const unsigned ub = __ballot_sync(FULL, active && stop && pop != minpop);   // stop wanted, but v is not final yet
    int stall_pop = INT_MAX;
    if (ub != 0u) {
      const unsigned urot = ((ub | (ub << 8)) >> (jnext & 7)) & 0xffu;
      stall_pop = jnext - 9 + __ffs(urot);
    }
    if (active && pop < stall_pop) {                       // advance one level
      const float ss = stop ? vs : xs;
      const int sid = stop ? vid : xid;
      *reinterpret_cast<uint2*>(&h[c]) = make_uint2(__float_as_uint(ss), (uint32_t)sid);
      if (c == 1) order[pop + 1] = (uint16_t)sid;          // the new root is the next pop's result
      xlast = stop ? xlast : xs;
      c = stop ? c : l + (right ? 1 : 0);
      active = !stop;
    }
    act = __ballot_sync(FULL, active);
    if (ub == 0u) ++since;
    const int sl = jnext & 7;
    if (jnext < npipe && since >= 2 && !((act >> sl) & 1u)) {   // next pop: two levels behind the previous one
      if (lane == sl) {
        c = 1;
        len = n - jnext - 1;
        vsrc = n - jnext;
        vid = -1;
        xlast = INF;
        pop = jnext;
        active = true;
      }
      act |= 1u << sl;
      ++jnext;
      since = 0;
    }
    __syncwarp();
  }
  if (stopped) return;
  hazard = __any_sync(FULL, hazard);
  int k0 = npipe, hl = n - npipe;
  if (hazard) {                               // rebuild the initial heap (the sorted array) and pop it sequentially
    for (int q = lane; q < n; q += 32) {
      HeapEntry e;
      e.score = sorted_scores[q];
      e.id = q;
      h[q + 1] = e;
    }
    k0 = 0;
    hl = n;
  } else if (npipe > published) {             // every pipelined pop has ended hazard-free
    published = npipe;
    if (lane == 0) {
      __threadfence_block();
      st_volatile_shared(&ps->popped, npipe);
    }
  }
  __syncwarp();
  if (lane == 0) {
    for (int k = k0; k < n; ++k) {
      const int id = heap_pop(h, hl);
      if (k >= published) order[k] = (uint16_t)id;
      if ((k & 63) == 63 || k == n - 1) {
        if (k + 1 > published) {
          published = k + 1;
          __threadfence_block();
          st_volatile_shared(&ps->popped, k + 1);
        }
        if (ld_volatile_shared(&ps->stop) != 0) break;
      }
    }
  }
  __syncwarp();
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

// block_nms for the ProposalLayer (decode(q) -> Box4 yields the box of sorted candidate q and stores it for the output
// stage): same chunk resolution, but the popper warp runs `popper_warp` on its own and the
// 992 workers never wait for it at a block barrier: pop positions below `first_tie` are the sorted order itself
// (order[] pre-filled with the identity; the heap pops the unique maximum until the first equal pair reaches the
// root), later positions are consumed as soon as the popper has published them.  `heap` == nullptr: no ties at all.
template <typename Decode>
__device__ inline int block_nms_async(Decode decode, uint16_t* order, int n, int max_out, float thr,
                                      float4* kept_box, float* kept_area, uint16_t* selected, NmsScratch* sc,
                                      PopperShared* ps, HeapEntry* heap, const float* sorted_scores, int first_tie) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  if (tid == 0) {
    sc->count = 0;
    ps->popped = 0;
    ps->stop = 0;
  }
  __syncthreads();
  if (tid >= NMS_WORKERS) {
    if (heap != nullptr && n > 0) popper_warp(heap, n, order, ps, sorted_scores);
    return 0;                                  // the caller's __syncthreads() joins the warps again
  }
  int count = 0;
  for (int base = 0; base < n && count < max_out; base += 64) {
    const int n_in = min(64, n - base);
    if (heap != nullptr && base + n_in > first_tie) {
      if (tid == 0) {
        while (ld_volatile_shared(&ps->popped) < base + n_in) {
        }
        __threadfence_block();
      }
      nms_workers_sync();
    }
    if (tid < 64) {
      NBox nb;
      if (tid < n_in) {
        nb = normalise_box(decode((int)order[base + tid]));     // box of this candidate, computed (and stored) on first use
      } else {
        nb.ymin = nb.xmin = nb.ymax = nb.xmax = 0.f;
        nb.area = -1.f;
      }
      sc->chunk[tid] = nb;
      sc->M[tid] = 0ull;
      sc->sup[tid] = 0;
    }
    nms_workers_sync();
    if (tid < 64 * 15) {   // candidate k = tid/15 against the kept list, 15 threads striding over it
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
                sc->sup[k] = 1;
                break;
              }
            }
          }
        }
      }
    }
    for (int p = tid; p < 64 * 64; p += NMS_WORKERS) {   // intra-chunk suppression matrix
      const int i = p >> 6, j = p & 63;
      if (j > i && j < n_in && iou_gt(sc->chunk[i], sc->chunk[j], thr)) atomicOr(&sc->M[i], 1ull << j);
    }
    nms_workers_sync();
    if (tid < 32) {
      const unsigned lo = __ballot_sync(0xffffffffu, lane < n_in && !sc->sup[lane]);
      const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < n_in && !sc->sup[lane + 32]);
      if (lane == 0) {
        unsigned long long alive = (unsigned long long)lo | ((unsigned long long)hi << 32);
        unsigned long long kept = 0ull;
        int c = count;
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
    if (tid < 64 && ((kept >> tid) & 1ull)) {   // append in selection order
      const int r = count + __popcll(kept & ((1ull << tid) - 1ull));
      const NBox nb = sc->chunk[tid];
      kept_box[r] = make_float4(nb.ymin, nb.xmin, nb.ymax, nb.xmax);
      kept_area[r] = nb.area;
    }
    count = sc->count;
    nms_workers_sync();          // kept list complete; chunk scratch free for the next round
  }
  if (tid == 0) st_volatile_shared(&ps->stop, 1);
  return count;
}
