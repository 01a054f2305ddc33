// ProposalLayer — one fused pass per image (one 1024-thread CTA per image):
//   radix-select + bitonic sort top-k over the anchor fg scores (ties -> lower index first)
//   -> gather deltas/anchors, apply_box_deltas, clip to [0,1]
//   -> greedy NMS in shared memory in TF-1.13 pop order (closed form, see step 5) -> zero-padded rois.
// Replaces mrcnn/model.py:329-406 (ProposalLayer.call), :287-326 (apply_box_deltas_graph,
// clip_boxes_graph) and the per-image Python unrolling of utils.batch_slice (utils.py:872-906).
// Bit-exact contract: top-k indices, NMS keep indices and rois equal oracle/graph_layers.py
// proposal_layer() for identical float32 inputs.
#include <stdio.h>
#include <stdlib.h>
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

constexpr int PROP_THREADS = 1024;
constexpr int PROP_MAX_K = 6144;
constexpr int PROP_MAX_R = 2048;     // POST_NMS_ROIS_TRAINING = 2000 (mrcnn/config.py:98)

struct PropParams {
  const float* rpn_class;  // [B,A,2]
  const float* rpn_bbox;   // [B,A,4]
  const float* anchors;    // [A,4] (stride 0) or [B,A,4]
  long long anchor_bstride;
  int A, K, Kpad, R;
  float thr;
  float sd[4];
  float* rois;        // [B,R,4]
  int32_t* topk_idx;  // [B,K]   (required: caller or workspace)
  int32_t* keep_idx;  // [B,R] or null (index into the top-k order, -1 padded)
  int32_t* keep_cnt;  // [B] or null
  int r0_bytes;
  int kept_off;   // byte offset of the kept list inside region 0
  long long* phase_clocks;   // debug: [B][8] clock64() at the phase boundaries (null = off)
};

__device__ __forceinline__ int block_excl_scan_flag(bool flag, int* warp_tot, int& total) {
  // exclusive prefix count of `flag` over the block (blockDim multiple of 32, <= 1024)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  int pre = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < nw; ++w) {
    int v = warp_tot[w];
    if (w < warp) off += v;
    tot += v;
  }
  __syncthreads();
  total = tot;
  return off + pre;
}

// Bitonic sort (descending) of 8192 64-bit keys by 1024 threads, 8 consecutive keys per thread.  The network is the
// textbook one over the flat index i = 8*tid + r (partner i ^ j, direction by bit k of i); only WHERE the partner
// lives changes the mechanics: j < 8 same thread (registers), 8 <= j < 256 same warp (shuffles), j >= 256 another
// warp (shared memory, two barriers).  15 of the 91 stages touch shared memory instead of all of them.
__device__ __forceinline__ void cmpx(unsigned long long& a, unsigned long long& b, bool desc) {
  const bool sw = desc ? (a < b) : (a > b);
  const unsigned long long t = a;
  a = sw ? b : a;
  b = sw ? t : b;
}

template <int J> __device__ __forceinline__ void sort_intra(unsigned long long (&e)[8], int k, bool desc) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if ((r & J) == 0) {
      const bool d = k >= 8 ? desc : ((r & k) == 0);     // k < 8: the direction bit is inside the thread's 8 keys
      cmpx(e[r], e[r | J], d);
    }
  }
}

__device__ inline void sort_desc_regs(unsigned long long* buf) {
  const int tid = threadIdx.x;
  unsigned long long e[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) e[r] = buf[tid * 8 + r];
  for (int k = 2; k <= 8192; k <<= 1) {
    const bool desc = ((tid * 8) & k) == 0;          // k >= 8: same for the thread's 8 keys; k < 8 handled per key below
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 256) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[tid * 8 + r] = e[r];
        __syncthreads();
        const int pt = tid ^ (j >> 3);
        const bool lower = (tid & (j >> 3)) == 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const unsigned long long c = buf[pt * 8 + r];
          const bool keep_max = lower == desc;
          e[r] = keep_max ? (e[r] > c ? e[r] : c) : (e[r] < c ? e[r] : c);
        }
      } else if (j >= 8) {
        const int m = j >> 3;                          // partner lane distance 1..16
        const bool lower = (tid & m) == 0;
        const bool keep_max = lower == desc;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const unsigned long long c = __shfl_xor_sync(0xffffffffu, e[r], m);
          e[r] = keep_max ? (e[r] > c ? e[r] : c) : (e[r] < c ? e[r] : c);
        }
      } else if (j == 4) {
        sort_intra<4>(e, k, desc);
      } else if (j == 2) {
        sort_intra<2>(e, k, desc);
      } else {
        sort_intra<1>(e, k, desc);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) buf[tid * 8 + r] = e[r];
  __syncthreads();
}

// Pre-order "node, right, left" rank of heap index i (1-based) among indices below 2^16: the path from the root (bits of
// i below its leading one, 1 = right child) inverted so that right sorts first, left-aligned, with the depth as the
// tie-break that puts an ancestor before its right-most descendants.
__device__ __forceinline__ uint32_t pop_key(uint32_t i) {
  const int d = 31 - __clz(i);
  const uint32_t path = i - (1u << d);
  const uint32_t inv = ~path & ((1u << d) - 1u);
  return ((inv << (16 - d)) << 5) | (uint32_t)d;
}

__global__ void __launch_bounds__(PROP_THREADS, 1) proposal_kernel(PropParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int A = p.A, K = p.K, Kpad = p.Kpad, R = p.R;

  // ---- shared memory carve-up ---------------------------------------------------------------
  unsigned char* r0 = smem;                                              // phase A / phase B region
  Box4* boxes = reinterpret_cast<Box4*>(smem + p.r0_bytes);              // [K]
  float* s_scores = reinterpret_cast<float*>(boxes + K);                 // [K]
  NmsScratch* sc = reinterpret_cast<NmsScratch*>(s_scores + ((K + 3) & ~3));
  int* hist = reinterpret_cast<int*>(sc + 1);                            // [256]
  int* warp_tot = hist + 256;                                            // [32]
  int* misc = warp_tot + 32;                                             // [8]

  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(r0);
  const float* scores = p.rpn_class + (size_t)b * A * 2 + 1;             // fg score, stride 2

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 0] = clock64();
  // ---- 1. radix select: key of the K-th largest score ---------------------------------------
  // the sortable keys of all A scores are read from global memory once and cached in the (still unused) box array when
  // they fit (A <= 4 K: 16 368 anchors at 256^2, not the 261 888 of 1024^2): the other three radix passes and the
  // compaction then run out of shared memory
  uint32_t* kcache = ((size_t)A * 4 <= (size_t)K * sizeof(Box4)) ? reinterpret_cast<uint32_t*>(boxes) : nullptr;
  uint32_t prefix = 0, mask = 0;
  int need = K;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < A; i += nt) {
      uint32_t key;
      if (kcache != nullptr && pass > 0) {
        key = kcache[i];
      } else {
        key = float_to_key(scores[2 * (size_t)i]);
        if (kcache != nullptr) kcache[i] = key;
      }
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, bin = 0;
      for (int bb = 255; bb >= 0; --bb) {
        int h = hist[bb];
        if (cum + h >= need) {
          bin = bb;
          break;
        }
        cum += h;
      }
      misc[0] = bin;
      misc[1] = need - cum;
    }
    __syncthreads();
    prefix |= (uint32_t)misc[0] << shift;
    mask |= 255u << shift;
    need = misc[1];
    __syncthreads();
  }
  const uint32_t T = prefix;       // K-th largest key; `need` elements equal to T are taken,
  const int cnt_gt = K - need;     // lowest indices first (tf.nn.top_k tie rule)

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 1] = clock64();
  // ---- 2. compaction into the sort buffer -----------------------------------------------------
  if (tid == 0) misc[2] = 0;
  for (int i = K + tid; i < Kpad; i += nt) sortbuf[i] = 0ull;
  __syncthreads();
  int eq_base = 0;
  for (int base = 0; base < A; base += nt) {
    const int i = base + tid;
    const bool valid = i < A;
    const uint32_t key = valid ? (kcache != nullptr ? kcache[i] : float_to_key(scores[2 * (size_t)i])) : 0u;
    const bool gt = valid && key > T;
    const bool eq = valid && key == T;
    int tot_eq;
    const int rank = eq_base + block_excl_scan_flag(eq, warp_tot, tot_eq);
    const unsigned long long comp = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)i);
    if (eq && rank < need) sortbuf[cnt_gt + rank] = comp;
    if (gt) sortbuf[atomicAdd(&misc[2], 1)] = comp;
    eq_base += tot_eq;
  }
  __syncthreads();

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 2] = clock64();
  // ---- 3. bitonic sort, descending on (key, ~index) ------------------------------------------
  if (Kpad == 8 * PROP_THREADS && nt == PROP_THREADS) {
    sort_desc_regs(sortbuf);       // 8 keys per thread: registers / shuffles / shared memory by partner distance
  } else
  for (int k = 2; k <= Kpad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < Kpad; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = sortbuf[i], c = sortbuf[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) {
            sortbuf[i] = c;
            sortbuf[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 3] = clock64();
  // ---- 4. top-k indices and scores (the score is the sort key itself; boxes are decoded lazily, only for the
  //         candidates the NMS actually reaches: ~1 100 of 6 000 on the bench maps) ----------------------------------
  __syncthreads();                                  // the key cache in `boxes` is dead from here on
  const float* deltas = p.rpn_bbox + (size_t)b * A * 4;
  const float* anch = p.anchors + (size_t)b * p.anchor_bstride;
  int32_t* topk = p.topk_idx + (size_t)b * K;
  for (int q = tid; q < K; q += nt) {
    const unsigned long long e = sortbuf[q];
    topk[q] = (int32_t)(~(uint32_t)(e & 0xffffffffull));
    s_scores[q] = key_to_float((uint32_t)(e >> 32));
  }
  __syncthreads();
  auto decode = [&](int q) {                        // mrcnn/model.py:355-376 for candidate q of the sorted order
    const uint32_t idx = (uint32_t)topk[q];
    const float4 d = *reinterpret_cast<const float4*>(deltas + 4 * (size_t)idx);
    const float4 a = *reinterpret_cast<const float4*>(anch + 4 * (size_t)idx);
    Box4 bx = {a.x, a.y, a.z, a.w};
    bx = apply_box_deltas(bx, __fmul_rn(d.x, p.sd[0]), __fmul_rn(d.y, p.sd[1]), __fmul_rn(d.z, p.sd[2]), __fmul_rn(d.w, p.sd[3]));
    bx = clip_box(bx, 0.f, 0.f, 1.f, 1.f);
    boxes[q] = bx;
    return bx;
  };

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 4] = clock64();
  // ---- 5. pop order: identity unless scores tie, else popped lazily from the emulated heap ------
  HeapEntry* heap = reinterpret_cast<HeapEntry*>(r0);                  // 1-indexed: [K+2], 16-byte aligned
  uint16_t* order = reinterpret_cast<uint16_t*>(heap + ((K + 2 + 1) & ~1));   // [K]
  uint16_t* selected = order + ((K + 1) & ~1);                          // [R]
  float4* kept_box = reinterpret_cast<float4*>(r0 + p.kept_off);         // [R]
  float* kept_area = reinterpret_cast<float*>(kept_box + R);            // [R]
  // ---- 5. pop order in closed form ---------------------------------------------------------------------------------
  // TF 1.13 pops a std::priority_queue; equal scores come out in whatever order libstdc++'s heap gives them.  The
  // initial heap is the sorted array (entry of rank q at index q+1), every position of a heap holds the maximum of its
  // subtree, and an element only ever moves up its own ancestor chain — so two equal elements meet exactly once, as the
  // two children of their lowest common ancestor, where __adjust_heap takes the RIGHT child on a tie.  Hence: distinct
  // scores pop in sorted order, and a run of equal scores pops in "node, right subtree, left subtree" pre-order of the
  // members' heap indices (pop_key below).  This holds until an element is moved by pop_heap's "last element to the
  // hole" step before it is popped itself, i.e. for every run that ends at or below rank K/2; a run reaching beyond
  // (more than half of the candidates tied: a saturated plateau) is left to the heap emulation (popper warp), which the
  // workers then wait for from that run's first position on.  tests/test_heap_pipeline_sim.py checks the closed form
  // against libstdc++ on every tie pattern.
  if (tid == 0) misc[3] = K;
  __syncthreads();
  const int half = K / 2;
  for (int q = tid; q < K; q += nt) {
    const float sc = s_scores[q];
    int pos = q;
    if ((q > 0 && s_scores[q - 1] == sc) || (q + 1 < K && s_scores[q + 1] == sc)) {
      int s0 = q, e0 = q;
      while (s0 > 0 && s_scores[s0 - 1] == sc) --s0;
      while (e0 + 1 < K && s_scores[e0 + 1] == sc) ++e0;
      if (e0 + 1 > half) {
        atomicMin(&misc[3], s0);               // not covered: the heap decides from position s0 on
      } else {
        const uint32_t kq = pop_key((uint32_t)q + 1u);
        int r = 0;
        for (int m = s0; m <= e0; ++m) r += pop_key((uint32_t)m + 1u) < kq ? 1 : 0;
        pos = s0 + r;
      }
    }
    order[pos] = (uint16_t)q;
  }
  __syncthreads();
  const int first_tie = misc[3];               // positions below it are final in order[]
  const bool any_tie = first_tie < K;
  if (any_tie) {
    // pushing the (already sorted) scores in order never sifts up: the array IS the initial heap
    for (int q = tid; q < K; q += nt) {
      HeapEntry e;
      e.score = s_scores[q];
      e.id = q;
      heap[q + 1] = e;
    }
  }
  __syncthreads();

  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 5] = clock64();
  // ---- 6. NMS + output ------------------------------------------------------------------------
  PopperShared* ps = reinterpret_cast<PopperShared*>(misc + 4);
  block_nms_async(decode, order, K, R, p.thr, kept_box, kept_area, selected, sc, ps, any_tie ? heap : nullptr, s_scores, first_tie);
  __syncthreads();
  const int count = sc->count;
  float4* out = reinterpret_cast<float4*>(p.rois + (size_t)b * R * 4);
  for (int r = tid; r < R; r += nt) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int cand = -1;
    if (r < count) {
      cand = order[selected[r]];
      const Box4 bx = boxes[cand];
      v = make_float4(bx.y1, bx.x1, bx.y2, bx.x2);
    }
    out[r] = v;
    if (p.keep_idx) p.keep_idx[(size_t)b * R + r] = cand;
  }
  if (tid == 0 && p.keep_cnt) p.keep_cnt[b] = count;
  if (p.phase_clocks && tid == 0) p.phase_clocks[(size_t)b * 8 + 6] = clock64();
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t prop_smem_bytes(int K, int R, int* r0_bytes, int* kept_off) {
  const int Kpad = next_pow2(K);
  size_t a = (size_t)Kpad * 8;
  // phase B of region 0: heap (1-indexed) | order | selected | kept boxes | kept areas
  size_t bsz = (size_t)((K + 2 + 1) & ~1) * sizeof(HeapEntry) + (size_t)((K + 1) & ~1) * 2 + (size_t)((R + 1) & ~1) * 2;
  bsz = (bsz + 15) & ~(size_t)15;
  *kept_off = (int)bsz;
  bsz += (size_t)R * 16 + (size_t)R * 4 + 16;
  size_t r0 = a > bsz ? a : bsz;
  r0 = (r0 + 15) & ~(size_t)15;
  *r0_bytes = (int)r0;
  return r0 + (size_t)K * sizeof(Box4) + (size_t)((K + 3) & ~3) * 4 + sizeof(NmsScratch) + (256 + 32 + 8) * 4;
}

}  // namespace

extern "C" size_t mrcnn_proposal_workspace_bytes(int batch, int num_anchors, int pre_nms_limit) {
  int K = pre_nms_limit < num_anchors ? pre_nms_limit : num_anchors;
  return (size_t)batch * (size_t)(K > 0 ? K : 1) * sizeof(int32_t);
}

extern "C" int mrcnn_proposal_layer(const float* rpn_class, const float* rpn_bbox, const float* anchors,
                                    int anchors_batched, int batch, int num_anchors, int pre_nms_limit,
                                    int proposal_count, float nms_threshold, const float* bbox_std_dev,
                                    float* rpn_rois, int32_t* topk_idx, int32_t* keep_idx,
                                    int32_t* keep_count, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  MRCNN_REQUIRE(rpn_class && rpn_bbox && anchors && rpn_rois && bbox_std_dev, "proposal_layer: null pointer");
  MRCNN_REQUIRE(batch > 0 && num_anchors > 0, "proposal_layer: empty input (batch=%d anchors=%d)", batch, num_anchors);
  const int K = pre_nms_limit < num_anchors ? pre_nms_limit : num_anchors;
  MRCNN_REQUIRE(K >= 1 && K <= PROP_MAX_K, "proposal_layer: min(PRE_NMS_LIMIT, anchors)=%d outside [1,%d]", K, PROP_MAX_K);
  MRCNN_REQUIRE(proposal_count >= 1 && proposal_count <= PROP_MAX_R, "proposal_layer: proposal_count=%d outside [1,%d]", proposal_count, PROP_MAX_R);
  PropParams p;
  p.rpn_class = rpn_class;
  p.rpn_bbox = rpn_bbox;
  p.anchors = anchors;
  p.anchor_bstride = anchors_batched ? (long long)num_anchors * 4 : 0;
  p.A = num_anchors;
  p.K = K;
  p.Kpad = next_pow2(K);
  p.R = proposal_count;
  p.thr = nms_threshold;
  for (int i = 0; i < 4; ++i) p.sd[i] = bbox_std_dev[i];
  p.rois = rpn_rois;
  p.keep_idx = keep_idx;
  p.keep_cnt = keep_count;
  if (topk_idx) {
    p.topk_idx = topk_idx;
  } else {
    MRCNN_REQUIRE(workspace && workspace_bytes >= mrcnn_proposal_workspace_bytes(batch, num_anchors, pre_nms_limit),
                  "proposal_layer: workspace too small");
    p.topk_idx = static_cast<int32_t*>(workspace);
  }
  size_t smem = prop_smem_bytes(K, proposal_count, &p.r0_bytes, &p.kept_off);
  MRCNN_REQUIRE(smem <= 227 * 1024, "proposal_layer: shared memory %zu exceeds 227 KB", smem);
  MRCNN_CHECK_CUDA(cudaFuncSetAttribute(proposal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  p.phase_clocks = nullptr;
  const char* dbg = getenv("MRCNN_B200_PROPOSAL_CLOCKS");       // debug: per-phase SM clock counts of image 0 on stderr
  if (dbg && dbg[0] == '1') MRCNN_CHECK_CUDA(cudaMalloc((void**)&p.phase_clocks, (size_t)batch * 8 * sizeof(long long)));
  MRCNN_CHECK_CUDA(mrcnn_launch(proposal_kernel, dim3(batch), dim3(PROP_THREADS), smem, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  if (p.phase_clocks) {
    long long h[8];
    MRCNN_CHECK_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    MRCNN_CHECK_CUDA(cudaMemcpy(h, p.phase_clocks, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "proposal phases (cycles, image 0): select %lld compact %lld sort %lld decode %lld heap-init %lld nms %lld total %lld\n",
            h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[6] - h[0]);
    cudaFree(p.phase_clocks);
  }
  return MRCNN_OK;
}
