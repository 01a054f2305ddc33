// Source-mask post-processing on the device: the integer / bit work behind
// Analyzer.extract_det_masks and Analyzer.make_json_results (mrcnn/analyze.py:1162-1423, 1866-1942,
// 2142-2173) on masks that are still resident in HBM after unmold_detections.
//
// The reference keeps every detection as a full-frame [H,W] array and answers "are these two masks
// connected?" by labelling mask1, mask2 and mask1+mask2 with skimage (three flood fills per pair,
// O(N^2) pairs) and "how much do they overlap?" with sklearn's jaccard_score on the flattened frames.
// Here every mask is a bit-plane ([H][ceil(W/32)] uint32, bit k of word w = pixel x = 32 w + k, bits
// past W are zero) and the same answers are exact integer reductions over words:
//   * ncomp(m1 + m2) < ncomp(m1) + ncomp(m2)  <=>  some pixel of m1 coincides with or is 4-adjacent to
//     a pixel of m2 (a component of the union that absorbs one component of each is the only way to
//     lose a component), i.e. any(m1 & (m2 | m2<<1 | m2>>1 | up(m2) | down(m2)));
//   * jaccard = |m1 & m2| / (|m1| + |m2| - |m1 & m2|): popcounts, the division is done by the host in
//     float64 exactly as sklearn does.
// 4-connected labelling (skimage.measure.label(connectivity=1): labels numbered in raster order of
// each component's first pixel) is a union-find over pixels with the smallest raster index as the root,
// followed by a raster-order ranking of the roots.  Everything is bit-exact integer work; all kernels
// are HBM/L2-bound streaming passes over words.
// Also here: the 8-connected pixel-list adjacency test of the tile driver (SFinder.merge_edge_sources,
// mrcnn/sfinder.py:787-808).  The host-only merge-graph routine lives in host_graph.cu.
#include <climits>

#include "common.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGridY = 65535;   // per-plane / per-group work rides on gridDim.y; the launchers chunk beyond this

__device__ __forceinline__ int words_per_row(int W) { return (W + 31) >> 5; }

// ---- pack: [B,H,W,D] uint8 (the [H,W,N] layout of detect(), N padded to D) -> selected bit-planes ----
// General-shape variant (any depth / alignment; masks_pack4_kernel below is the fast path for depth % 4 == 0).
// grid (WW, H, B).  The CTA stages the 32-pixel x D-byte tile with coalesced loads, then each warp
// ballots one detection at a time (row stride D bytes: conflict-free for odd D/4, 2-way at worst).
__global__ void masks_pack_kernel(const uint8_t* __restrict__ masks, int H, int W, int D,
                                  const int32_t* __restrict__ plane_of, uint32_t* __restrict__ planes) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t s_tile[];  // [32][D]
  const int wx = blockIdx.x, y = blockIdx.y, b = blockIdx.z;
  const int WW = words_per_row(W);
  const int x0 = wx * 32;
  const int npix = min(32, W - x0);
  const size_t base = (((size_t)b * H + y) * W + x0) * D;
  const int nbytes = npix * D;
  if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(masks) & 3) == 0) {   // whole words: tile base and size are multiples of 4
    const uint32_t* src = reinterpret_cast<const uint32_t*>(masks + base);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_tile);
    for (int i = threadIdx.x; i < (nbytes >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s_tile[i] = masks[base + i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int d = warp; d < D; d += nwarps) {
    const int m = plane_of[b * D + d];
    if (m < 0) continue;
    const bool on = lane < npix && s_tile[lane * D + d] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) planes[((size_t)m * H + y) * WW + wx] = word;
  }
}

// Fast variant for depth % 4 == 0 and 16-byte aligned rows: 256 pixels (eight output words) of one image row per
// CTA, 16-byte loads (6 in flight per thread at depth 100), four detections per shared-memory word in the ballot
// phase (row stride depth/4 words: conflict-free when odd, as for depth 100); lanes 0-3 of a warp each collect the
// eight words of one detection and write them as one 32-byte run.  grid (ceil(W/256), H, B), 256 threads.
constexpr int kPackPx = 256;
__global__ void __launch_bounds__(256) masks_pack4_kernel(const uint8_t* __restrict__ masks, int H, int W, int D,
                                                          const int32_t* __restrict__ plane_of, uint32_t* __restrict__ planes) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t s_tile[];  // [kPackPx][D]
  const int y = blockIdx.y, b = blockIdx.z;
  const int WW = words_per_row(W);
  const int x0 = blockIdx.x * kPackPx;
  const int npix = min(kPackPx, W - x0);
  const size_t base = (((size_t)b * H + y) * W + x0) * D;
  const int nbytes = npix * D;
  {
    const uint4* src = reinterpret_cast<const uint4*>(masks + base);
    uint4* dst = reinterpret_cast<uint4*>(s_tile);
    const int n16 = nbytes >> 4;
    for (int i = threadIdx.x; i < n16; i += 256) dst[i] = __ldg(src + i);
    const uint32_t* src4 = reinterpret_cast<const uint32_t*>(masks + base);
    uint32_t* dst4 = reinterpret_cast<uint32_t*>(s_tile);
    for (int i = (n16 << 2) + threadIdx.x; i < (nbytes >> 2); i += 256) dst4[i] = __ldg(src4 + i);
  }
  __syncthreads();
  // ballot phase: warp w owns pixels 32w..32w+31; a lane reads the DQ words of its pixel (row stride DQ words:
  // conflict-free when odd) and every byte costs one test + one vote + one lane-0 store into the [D][8] staging
  // table, which then leaves as 32-byte runs per detection (no per-lane select chains)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int DQ = D >> 2;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s_tile);
  uint32_t* s_out = reinterpret_cast<uint32_t*>(s_tile + (size_t)kPackPx * D);   // [D][8]
  const int nwords = (npix + 31) >> 5;
  if (warp < nwords) {
    const int px = warp * 32 + lane;
    const uint32_t* row = s32 + (px < npix ? px : 0) * DQ;
#pragma unroll 4
    for (int w = 0; w < DQ; ++w) {
      const uint32_t v = px < npix ? row[w] : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t bits = __ballot_sync(0xffffffffu, (v & (0xffu << (8 * q))) != 0);
        if (lane == 0) s_out[(w * 4 + q) * 8 + warp] = bits;
      }
    }
  }
  __syncthreads();
  if (nwords == 8 && (WW & 3) == 0) {
    for (int i = threadIdx.x; i < D * 2; i += 256) {       // two 16-byte halves per detection
      const int d = i >> 1, h = i & 1;
      const int m = plane_of[b * D + d];
      if (m >= 0)
        reinterpret_cast<uint4*>(planes + ((size_t)m * H + y) * WW + blockIdx.x * 8)[h] =
            reinterpret_cast<const uint4*>(s_out + d * 8)[h];
    }
  } else {
    for (int i = threadIdx.x; i < D * 8; i += 256) {
      const int d = i >> 3, g = i & 7;
      const int m = plane_of[b * D + d];
      if (m >= 0 && g < nwords) planes[((size_t)m * H + y) * WW + blockIdx.x * 8 + g] = s_out[d * 8 + g];
    }
  }
}

// ---- area + bounding box (utils.extract_bboxes, mrcnn/utils.py:33-59): one CTA per plane ----
__global__ void planes_area_bbox_kernel(const uint32_t* __restrict__ planes, int H, int W, int32_t* __restrict__ area,
                                        int32_t* __restrict__ bbox) {
  pdl_prologue();
  const int m = blockIdx.x, WW = words_per_row(W);
  const uint32_t* pl = planes + (size_t)m * H * WW;
  int cnt = 0, y1 = INT_MAX, y2 = -1, x1 = INT_MAX, x2 = -1;
  for (int i = threadIdx.x; i < H * WW; i += blockDim.x) {
    const uint32_t w = pl[i];
    if (!w) continue;
    const int y = i / WW, wx = i - y * WW;
    cnt += __popc(w);
    y1 = min(y1, y);
    y2 = max(y2, y);
    x1 = min(x1, wx * 32 + __ffs(w) - 1);
    x2 = max(x2, wx * 32 + 31 - __clz(w));
  }
  __shared__ int s_cnt, s_y1, s_y2, s_x1, s_x2;
  if (threadIdx.x == 0) { s_cnt = 0; s_y1 = INT_MAX; s_x1 = INT_MAX; s_y2 = -1; s_x2 = -1; }
  __syncthreads();
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    y1 = min(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    x1 = min(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y2 = max(y2, __shfl_xor_sync(0xffffffffu, y2, o));
    x2 = max(x2, __shfl_xor_sync(0xffffffffu, x2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt, cnt);
    atomicMin(&s_y1, y1);
    atomicMin(&s_x1, x1);
    atomicMax(&s_y2, y2);
    atomicMax(&s_x2, x2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    area[m] = s_cnt;
    const bool any = s_cnt > 0;   // no pixels: the reference resets the box to zeros
    bbox[m * 4 + 0] = any ? s_y1 : 0;
    bbox[m * 4 + 1] = any ? s_x1 : 0;
    bbox[m * 4 + 2] = any ? s_y2 + 1 : 0;
    bbox[m * 4 + 3] = any ? s_x2 + 1 : 0;
  }
}

// ---- pair statistics: |a & b| and the 4-neighbourhood touch test, one CTA per pair ----
__global__ void planes_pair_stats_kernel(const uint32_t* __restrict__ planes, int H, int W, const int32_t* __restrict__ pairs,
                                         const int32_t* __restrict__ bbox, int32_t* __restrict__ inter,
                                         int32_t* __restrict__ touch) {
  pdl_prologue();
  const int p = blockIdx.x, WW = words_per_row(W);
  const int ia = pairs[2 * p], ib = pairs[2 * p + 1];
  int row_lo = 0, row_hi = H;
  if (bbox) {
    // (y1,x1,y2,x2), y2/x2 exclusive, zeros for an empty mask: masks whose boxes are more than one pixel apart can
    // neither intersect nor touch; otherwise only the rows of a's box can contribute
    const int ay1 = bbox[4 * ia], ax1 = bbox[4 * ia + 1], ay2 = bbox[4 * ia + 2], ax2 = bbox[4 * ia + 3];
    const int by1 = bbox[4 * ib], bx1 = bbox[4 * ib + 1], by2 = bbox[4 * ib + 2], bx2 = bbox[4 * ib + 3];
    if (ay2 < by1 || by2 < ay1 || ax2 < bx1 || bx2 < ax1 || ay2 <= ay1 || by2 <= by1) {
      if (threadIdx.x == 0) { inter[p] = 0; touch[p] = 0; }
      return;
    }
    row_lo = ay1;
    row_hi = ay2;
  }
  const uint32_t* A = planes + (size_t)ia * H * WW;
  const uint32_t* B = planes + (size_t)ib * H * WW;
  int cnt = 0;
  uint32_t hit = 0;
  for (int i = row_lo * WW + threadIdx.x; i < row_hi * WW; i += blockDim.x) {
    const uint32_t a = A[i];
    if (!a) continue;
    const int y = i / WW, wx = i - y * WW;
    const uint32_t b = B[i];
    cnt += __popc(a & b);
    uint32_t nb = b | (b << 1) | (b >> 1);
    if (wx > 0) nb |= B[i - 1] >> 31;
    if (wx + 1 < WW) nb |= B[i + 1] << 31;
    if (y > 0) nb |= B[i - WW];
    if (y + 1 < H) nb |= B[i + WW];
    hit |= a & nb;
  }
  __shared__ int s_cnt, s_hit;
  if (threadIdx.x == 0) { s_cnt = 0; s_hit = 0; }
  __syncthreads();
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const bool any = __any_sync(0xffffffffu, hit != 0);
  if ((threadIdx.x & 31) == 0) {
    if (cnt) atomicAdd(&s_cnt, cnt);
    if (any) atomicOr(&s_hit, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    inter[p] = s_cnt;
    touch[p] = s_hit;
  }
}

// ---- union of member planes (Analyzer.merge_masks folded over a graph component) ----
__global__ void planes_union_kernel(const uint32_t* __restrict__ planes, size_t words, const int32_t* __restrict__ members,
                                    const int32_t* __restrict__ offsets, uint32_t* __restrict__ out) {
  pdl_prologue();
  const int g = blockIdx.y;
  const int lo = offsets[g], hi = offsets[g + 1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t acc = 0;
    for (int k = lo; k < hi; ++k) acc |= planes[(size_t)members[k] * words + i];
    out[(size_t)g * words + i] = acc;
  }
}

// ---- 4-connected labelling ----
__device__ __forceinline__ int uf_find(const int32_t* parent, int p) {
  const volatile int32_t* par = parent;   // other threads lower entries concurrently: always re-read
  int q = par[p];
  while (q != p) {
    p = q;
    q = par[p];
  }
  return p;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
  // the smaller raster index becomes the root, so a component's root is its first pixel
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

// first pixel of the horizontal run that contains (y, x)
__device__ __forceinline__ int run_start(const uint32_t* row, int x) {
  int w = x >> 5;
  uint32_t zeros = ~row[w] & ((1u << (x & 31)) - 1u);
  while (true) {
    if (zeros) return w * 32 + 32 - __clz(zeros);
    if (w == 0) return 0;
    --w;
    zeros = ~row[w];
  }
}

// parent[p] = start of p's horizontal run (foreground) or -1 (background): grid (ceil(H*W/T), M)
__global__ void label_init_kernel(const uint32_t* __restrict__ planes, int H, int W, int32_t* __restrict__ parent) {
  pdl_prologue();
  const int m = blockIdx.y, WW = words_per_row(W);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int y = p / W, x = p - y * W;
  const uint32_t* row = planes + ((size_t)m * H + y) * WW;
  const bool on = (row[x >> 5] >> (x & 31)) & 1u;
  parent[(size_t)m * H * W + p] = on ? y * W + run_start(row, x) : -1;
}

// vertical links: the first pixel of every stretch where this row and the row above are both set
__global__ void label_merge_kernel(const uint32_t* __restrict__ planes, int H, int W, int32_t* __restrict__ parent) {
  pdl_prologue();
  const int m = blockIdx.y, WW = words_per_row(W);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // word index; row 0 has no row above
  if (i >= H * WW || i < WW) return;
  const uint32_t* pl = planes + (size_t)m * H * WW;
  const uint32_t both = pl[i] & pl[i - WW];
  if (!both) return;
  const int y = i / WW, wx = i - y * WW;
  const uint32_t carry = wx > 0 ? ((pl[i - 1] & pl[i - 1 - WW]) >> 31) : 0u;
  uint32_t starts = both & ~((both << 1) | carry);
  int32_t* par = parent + (size_t)m * H * W;
  while (starts) {
    const int x = wx * 32 + __ffs(starts) - 1;
    starts &= starts - 1;
    uf_union(par, y * W + x, (y - 1) * W + x);
  }
}

__global__ void label_flatten_kernel(int H, int W, int32_t* __restrict__ parent) {
  pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  int32_t* par = parent + (size_t)blockIdx.y * H * W;
  // roots never change here and non-roots only move closer to their root: no race that matters
  if (par[p] >= 0) par[p] = uf_find(par, p);
}

// rank of every root in raster order (one CTA per plane, 1024 threads), counts[m] = number of roots
__global__ void label_rank_kernel(const int32_t* __restrict__ parent, int H, int W, int32_t* __restrict__ rank,
                                  int32_t* __restrict__ counts) {
  pdl_prologue();
  const int m = blockIdx.x, n = H * W;
  const int32_t* par = parent + (size_t)m * n;
  int32_t* rk = rank + (size_t)m * n;
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int p = base + threadIdx.x;
    const bool root = p < n && par[p] == p;
    const uint32_t bal = __ballot_sync(0xffffffffu, root);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (root) rk[p] = before + __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_warp[w];
      s_carry += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[m] = s_carry;
}

__global__ void label_assign_kernel(int H, int W, const int32_t* __restrict__ rank, int32_t* __restrict__ labels) {
  pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const size_t off = (size_t)blockIdx.y * H * W;
  const int r = labels[off + p];
  // every pixel reads its root's rank from the separate rank array, so the in-place update is safe
  labels[off + p] = r < 0 ? 0 : rank[off + r] + 1;
}

// one warp per output word: component `comp[k]` of plane `src[k]` as a new bit-plane
__global__ void labels_select_kernel(const int32_t* __restrict__ labels, int H, int W, const int32_t* __restrict__ src,
                                     const int32_t* __restrict__ comp, uint32_t* __restrict__ out) {
  pdl_prologue();
  const int k = blockIdx.y, WW = words_per_row(W);
  const int word = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (word >= H * WW) return;
  const int lane = threadIdx.x & 31;
  const int y = word / WW, x = (word - y * WW) * 32 + lane;
  const int32_t* lab = labels + (size_t)src[k] * H * W;
  const bool on = x < W && lab[y * W + x] == comp[k];
  const uint32_t bits = __ballot_sync(0xffffffffu, on);
  if (lane == 0) out[(size_t)k * H * WW + word] = bits;
}

// ---- pixel lists (np.argwhere(mask == 1), row-major) : one CTA per plane ----
__global__ void planes_pixels_kernel(const uint32_t* __restrict__ planes, int H, int W, const int64_t* __restrict__ offsets,
                                     int y0, int x0, int32_t* __restrict__ out) {
  pdl_prologue();
  const int m = blockIdx.x, WW = words_per_row(W), n = H * WW;
  const uint32_t* pl = planes + (size_t)m * n;
  int32_t* dst = out + 2 * offsets[m];
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    uint32_t w = i < n ? pl[i] : 0u;
    const int c = __popc(w);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int pos = s_carry + incl - c;
    for (int k = 0; k < warp; ++k) pos += s_warp[k];
    if (w) {
      const int y = i / WW, xb = (i - y * WW) * 32;
      while (w) {
        const int x = xb + __ffs(w) - 1;
        w &= w - 1;
        dst[2 * pos] = y + y0;
        dst[2 * pos + 1] = x + x0;
        ++pos;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < nw; ++k) tot += s_warp[k];
      s_carry += tot;
    }
    __syncthreads();
  }
}

__global__ void planes_unpack_kernel(const uint32_t* __restrict__ planes, int H, int W, uint8_t* __restrict__ out) {
  pdl_prologue();
  const int m = blockIdx.y, WW = words_per_row(W);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int y = p / W, x = p - y * W;
  out[(size_t)m * H * W + p] = (planes[((size_t)m * H + y) * WW + (x >> 5)] >> (x & 31)) & 1u;
}

// ---- 8-connected adjacency of two pixel lists (SFinder.merge_edge_sources, mrcnn/sfinder.py:787-808) ----
// One CTA per candidate pair: the second list streams through shared memory in chunks, every thread walks its share
// of the first list against the chunk; the CTA stops at the first hit (checked once per chunk).
constexpr int kAdjChunk = 2048;
__global__ void __launch_bounds__(256) pixel_lists_adjacent_kernel(const int32_t* __restrict__ pixels,
                                                                   const int64_t* __restrict__ offsets,
                                                                   const int32_t* __restrict__ pairs, int32_t* __restrict__ out) {
  pdl_prologue();
  __shared__ int2 s_b[kAdjChunk];
  __shared__ int s_hit;
  const int p = blockIdx.x;
  const int64_t a0 = offsets[pairs[2 * p]], a1 = offsets[pairs[2 * p] + 1];
  const int64_t b0 = offsets[pairs[2 * p + 1]], b1 = offsets[pairs[2 * p + 1] + 1];
  const int2* px = reinterpret_cast<const int2*>(pixels);
  if (threadIdx.x == 0) s_hit = 0;
  __syncthreads();
  for (int64_t cb = b0; cb < b1; cb += kAdjChunk) {
    const int nb = (int)min((int64_t)kAdjChunk, b1 - cb);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_b[i] = px[cb + i];
    __syncthreads();
    bool hit = false;
    for (int64_t ia = a0 + threadIdx.x; ia < a1 && !hit; ia += blockDim.x) {
      const int2 a = px[ia];
      for (int i = 0; i < nb; ++i) {
        const int2 b = s_b[i];
        if (abs(a.x - b.x) <= 1 && abs(a.y - b.y) <= 1) { hit = true; break; }
      }
    }
    if (hit) s_hit = 1;
    __syncthreads();
    if (s_hit) break;
  }
  if (threadIdx.x == 0) out[p] = s_hit;
}

inline int check_frame(int H, int W, const char* who) {
  MRCNN_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 31), "%s: bad frame %dx%d", who, H, W);
  return MRCNN_OK;
}

}  // namespace

#define RC(x)                    \
  do {                           \
    int _rc = (x);               \
    if (_rc != MRCNN_OK) return _rc; \
  } while (0)

extern "C" size_t mrcnn_plane_words(int height, int width) {
  return height > 0 && width > 0 ? (size_t)height * ((width + 31) / 32) : 0;
}

extern "C" int mrcnn_masks_pack(const uint8_t* masks, int n_images, int height, int width, int depth,
                                const int32_t* plane_of, uint32_t* planes, void* stream) {
  MRCNN_REQUIRE(masks && plane_of && planes, "masks_pack: null pointer");
  RC(check_frame(height, width, "masks_pack"));
  MRCNN_REQUIRE(n_images > 0 && n_images <= 65535 && height <= 65535 && depth > 0 && depth * 32 <= 200 * 1024,
                "masks_pack: bad sizes (n_images %d, depth %d)", n_images, depth);
  const bool fast = (depth & 3) == 0 && ((size_t)width * depth) % 16 == 0 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(planes) & 15) == 0 && (size_t)(kPackPx + 32) * depth <= 48 * 1024;
  if (fast) {
    dim3 grid((width + kPackPx - 1) / kPackPx, height, n_images);
    MRCNN_CHECK_CUDA(mrcnn_launch(masks_pack4_kernel, grid, dim3(256), (size_t)(kPackPx + 32) * depth, (cudaStream_t)stream,
                                  masks, height, width, depth, plane_of, planes));
  } else {
    const size_t smem = (size_t)32 * depth;
    if (smem > 48 * 1024)
      MRCNN_CHECK_CUDA(cudaFuncSetAttribute(masks_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((width + 31) / 32, height, n_images);
    MRCNN_CHECK_CUDA(mrcnn_launch(masks_pack_kernel, grid, dim3(128), smem, (cudaStream_t)stream, masks, height, width, depth,
                                  plane_of, planes));
  }
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_planes_area_bbox(const uint32_t* planes, int n_planes, int height, int width, int32_t* area,
                                      int32_t* bbox, void* stream) {
  MRCNN_REQUIRE(n_planes >= 0, "planes_area_bbox: negative count");
  if (n_planes == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && area && bbox, "planes_area_bbox: null pointer");
  RC(check_frame(height, width, "planes_area_bbox"));
  MRCNN_CHECK_CUDA(mrcnn_launch(planes_area_bbox_kernel, dim3(n_planes), dim3(kThreads), 0, (cudaStream_t)stream, planes,
                                height, width, area, bbox));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_planes_pair_stats(const uint32_t* planes, int height, int width, const int32_t* pairs, int n_pairs,
                                       const int32_t* bbox, int32_t* inter, int32_t* touch, void* stream) {
  MRCNN_REQUIRE(n_pairs >= 0, "planes_pair_stats: negative count");
  if (n_pairs == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && pairs && inter && touch, "planes_pair_stats: null pointer");
  RC(check_frame(height, width, "planes_pair_stats"));
  MRCNN_CHECK_CUDA(mrcnn_launch(planes_pair_stats_kernel, dim3(n_pairs), dim3(kThreads), 0, (cudaStream_t)stream, planes,
                                height, width, pairs, bbox, inter, touch));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_planes_union(const uint32_t* planes, int height, int width, const int32_t* members,
                                  const int32_t* offsets, int n_groups, uint32_t* out, void* stream) {
  MRCNN_REQUIRE(n_groups >= 0, "planes_union: negative count");
  if (n_groups == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && members && offsets && out, "planes_union: null pointer");
  RC(check_frame(height, width, "planes_union"));
  const size_t words = mrcnn_plane_words(height, width);
  if (n_groups > kMaxGridY) {            // groups ride on gridDim.y: larger batches go in chunks
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGridY)
      RC(mrcnn_planes_union(planes, height, width, members, offsets + g0, n_groups - g0 < kMaxGridY ? n_groups - g0 : kMaxGridY,
                            out + (size_t)g0 * words, stream));
    return MRCNN_OK;
  }
  const int bx = (int)((words + kThreads - 1) / kThreads < 1024 ? (words + kThreads - 1) / kThreads : 1024);
  MRCNN_CHECK_CUDA(mrcnn_launch(planes_union_kernel, dim3(bx, n_groups), dim3(kThreads), 0, (cudaStream_t)stream, planes,
                                words, members, offsets, out));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" size_t mrcnn_planes_label_workspace_bytes(int n_planes, int height, int width) {
  return n_planes > 0 && height > 0 && width > 0 ? (size_t)n_planes * height * width * sizeof(int32_t) : 0;
}

extern "C" int mrcnn_planes_label(const uint32_t* planes, int n_planes, int height, int width, int32_t* labels,
                                  int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  MRCNN_REQUIRE(n_planes >= 0, "planes_label: negative count");
  if (n_planes == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && labels && counts && workspace, "planes_label: null pointer");
  RC(check_frame(height, width, "planes_label"));
  MRCNN_REQUIRE(workspace_bytes >= mrcnn_planes_label_workspace_bytes(n_planes, height, width),
                "planes_label: workspace too small");
  if (n_planes > kMaxGridY) {
    const size_t words = mrcnn_plane_words(height, width), npix = (size_t)height * width;
    for (int m0 = 0; m0 < n_planes; m0 += kMaxGridY) {
      const int cnt = n_planes - m0 < kMaxGridY ? n_planes - m0 : kMaxGridY;
      RC(mrcnn_planes_label(planes + (size_t)m0 * words, cnt, height, width, labels + (size_t)m0 * npix, counts + m0,
                            static_cast<int32_t*>(workspace) + (size_t)m0 * npix, (size_t)cnt * npix * sizeof(int32_t), stream));
    }
    return MRCNN_OK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = height * width, WW = (width + 31) / 32;
  dim3 gpix((n + kThreads - 1) / kThreads, n_planes), gword((height * WW + kThreads - 1) / kThreads, n_planes);
  int32_t* rank = static_cast<int32_t*>(workspace);
  MRCNN_CHECK_CUDA(mrcnn_launch(label_init_kernel, gpix, dim3(kThreads), 0, st, planes, height, width, labels));
  MRCNN_CHECK_CUDA(mrcnn_launch(label_merge_kernel, gword, dim3(kThreads), 0, st, planes, height, width, labels));
  MRCNN_CHECK_CUDA(mrcnn_launch(label_flatten_kernel, gpix, dim3(kThreads), 0, st, height, width, labels));
  MRCNN_CHECK_CUDA(mrcnn_launch(label_rank_kernel, dim3(n_planes), dim3(1024), 0, st, (const int32_t*)labels, height, width,
                                rank, counts));
  MRCNN_CHECK_CUDA(mrcnn_launch(label_assign_kernel, gpix, dim3(kThreads), 0, st, height, width, (const int32_t*)rank, labels));
  mrcnn_count_launch(5);
  return MRCNN_OK;
}

extern "C" int mrcnn_labels_select(const int32_t* labels, int height, int width, const int32_t* src, const int32_t* comp,
                                   int n_out, uint32_t* planes_out, void* stream) {
  MRCNN_REQUIRE(n_out >= 0, "labels_select: negative count");
  if (n_out == 0) return MRCNN_OK;
  MRCNN_REQUIRE(labels && src && comp && planes_out, "labels_select: null pointer");
  RC(check_frame(height, width, "labels_select"));
  const int words = (int)mrcnn_plane_words(height, width);
  if (n_out > kMaxGridY) {
    for (int k0 = 0; k0 < n_out; k0 += kMaxGridY)
      RC(mrcnn_labels_select(labels, height, width, src + k0, comp + k0, n_out - k0 < kMaxGridY ? n_out - k0 : kMaxGridY,
                             planes_out + (size_t)k0 * words, stream));
    return MRCNN_OK;
  }
  const int wpb = kThreads / 32;
  MRCNN_CHECK_CUDA(mrcnn_launch(labels_select_kernel, dim3((words + wpb - 1) / wpb, n_out), dim3(kThreads), 0,
                                (cudaStream_t)stream, labels, height, width, src, comp, planes_out));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_planes_pixels(const uint32_t* planes, int n_planes, int height, int width, const int64_t* offsets,
                                   int y_origin, int x_origin, int32_t* pixels, void* stream) {
  MRCNN_REQUIRE(n_planes >= 0, "planes_pixels: negative count");
  if (n_planes == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && offsets && pixels, "planes_pixels: null pointer");
  RC(check_frame(height, width, "planes_pixels"));
  MRCNN_CHECK_CUDA(mrcnn_launch(planes_pixels_kernel, dim3(n_planes), dim3(1024), 0, (cudaStream_t)stream, planes, height,
                                width, offsets, y_origin, x_origin, pixels));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_planes_unpack(const uint32_t* planes, int n_planes, int height, int width, uint8_t* out, void* stream) {
  MRCNN_REQUIRE(n_planes >= 0, "planes_unpack: negative count");
  if (n_planes == 0) return MRCNN_OK;
  MRCNN_REQUIRE(planes && out, "planes_unpack: null pointer");
  RC(check_frame(height, width, "planes_unpack"));
  const int n = height * width;
  if (n_planes > kMaxGridY) {
    const size_t words = mrcnn_plane_words(height, width);
    for (int m0 = 0; m0 < n_planes; m0 += kMaxGridY)
      RC(mrcnn_planes_unpack(planes + (size_t)m0 * words, n_planes - m0 < kMaxGridY ? n_planes - m0 : kMaxGridY, height, width,
                             out + (size_t)m0 * n, stream));
    return MRCNN_OK;
  }
  MRCNN_CHECK_CUDA(mrcnn_launch(planes_unpack_kernel, dim3((n + kThreads - 1) / kThreads, n_planes), dim3(kThreads), 0,
                                (cudaStream_t)stream, planes, height, width, out));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_pixel_lists_adjacent(const int32_t* pixels, const int64_t* offsets, const int32_t* pairs, int n_pairs,
                                          int32_t* adjacent, void* stream) {
  MRCNN_REQUIRE(n_pairs >= 0, "pixel_lists_adjacent: negative count");
  if (n_pairs == 0) return MRCNN_OK;
  MRCNN_REQUIRE(pixels && offsets && pairs && adjacent, "pixel_lists_adjacent: null pointer");
  MRCNN_CHECK_CUDA(mrcnn_launch(pixel_lists_adjacent_kernel, dim3(n_pairs), dim3(256), 0, (cudaStream_t)stream, pixels, offsets,
                                pairs, adjacent));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
