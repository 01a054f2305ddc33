// bf16 implicit-GEMM convolution for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle, OOB zero fill
// = SAME padding) -> shared memory ring -> tcgen05.mma (UMMA 128 x BLOCK_N x 16, fp32 accumulators in
// TMEM) -> tcgen05.ld epilogue (folded BN scale/shift, residual add incl. nearest-2x FPN upsample,
// ReLU, bf16/f32 store, 2x2-stride-2 transposed-conv scatter).
//
// Replaces every KL.Conv2D / TimeDistributed(Conv2D|Dense) / Conv2DTranspose on the detect path:
// mrcnn/model.py:99-210 (ResNet-101), :2003-2026 (FPN), :916-957 (RPN), :986-1039 (class head),
// :1042-1091 (mask head).  One CTA computes a 128 x BLOCK_N output tile:
//   warp 0   : TMA producer (one lane)      — A box (64ch, tw, th, nb) per filter tap, B box (64, BLOCK_N)
//   warp 1   : TMEM alloc + MMA issuer (one lane), tcgen05.commit releases ring slots
//   warps 2-9: epilogue; warp w owns TMEM lanes 32*(w%4)..+31 (= output rows) and half of the
//              tile's columns; per-channel scale/shift are staged in smem once per tile and the
//              TMEM loads are software-pipelined against the math/stores of the previous chunk
// PERSISTENT: one CTA per SM walks a static tile schedule (N fastest, so the CTAs that share an A
// tile run at the same time and re-read it from L2).  The smem ring runs across tile boundaries and
// the accumulator is double-buffered in TMEM (2 x BLOCK_N columns), so the TMA loads and MMAs of
// tile i+1 overlap the epilogue of tile i — layers with K = 64..256 (1-4 k-blocks per tile) are
// otherwise dominated by per-tile latency.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "conv_gemm.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;        // TMA warp + MMA warp + 8 epilogue warps
constexpr int EPI_THREADS = 256;

// EPI_TMA: the epilogue goes through shared memory: the residual tile is fetched with TMA loads and the
// bf16 output leaves with TMA stores ([128 rows x 32 channels] boxes, 64B swizzle), so global traffic is
// full-line and asynchronous instead of one 16-byte piece per thread per row (32 sectors per request).
// OCC2: two CTAs per SM (shared memory <= 113 KB, registers <= 102 per thread, 2 x 2 x BLOCK_N <= 512 TMEM columns): the
// ring is shorter and the epilogue rotates through two staging buffers instead of three, but the fill / drain phases of one
// CTA (first loads, accumulator read-out, the kernel's tail) run under the other CTA's MMAs, and the CTAs of the NEXT layer
// can become resident (programmatic dependent launch) while this layer's last tiles drain.  Layers with 1-16 k-blocks per
// tile are bound by exactly those phases.
template <int BLOCK_N, bool EPI_TMA, bool OCC2 = false> struct TileCfg {
  static constexpr int STAGES = OCC2 ? (EPI_TMA ? (BLOCK_N <= 32 ? 3 : 2) : (BLOCK_N <= 64 ? 4 : 3))
                                : EPI_TMA ? ((BLOCK_N <= 32) ? 8 : (BLOCK_N == 64 ? 7 : (BLOCK_N == 128 ? 5 : 3)))
                                          : ((BLOCK_N <= 64) ? 8 : (BLOCK_N == 128 ? 6 : 4));
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_CHUNK_BYTES = 32 * 32 * 2;                        // one warp's chunk: 32 rows x 32 bf16
  static constexpr int EPI_NBUF = OCC2 ? 2 : 3;                              // rotating staging buffers per epilogue warp
  static constexpr int STAGING_BYTES = EPI_TMA ? 8 * EPI_NBUF * EPI_CHUNK_BYTES : 0;
  // [acc][scale|shift][BLOCK_N] floats + (fused mask logits) W2 [256][4] floats + partials [128][8] floats
  static constexpr int EPI_BYTES = 2 * 2 * BLOCK_N * 4 + ((BLOCK_N == 256 && !EPI_TMA) ? (256 * 4 * 4 + 128 * 8 * 4) : 0);
  static constexpr int BAR_BYTES = 512;                                      // mbarriers + TMEM slot
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + BAR_BYTES + EPI_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;   // double-buffered accumulator (64..512, power of two)
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(!OCC2 || (SMEM_BYTES <= 115712 && BLOCK_N <= 128), "two CTAs per SM: 113 KB and 256 TMEM columns each");
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (error surfaces on the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 400000000u) {
      printf("conv_gemm: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// im2col-mode load (3x3 SAME convs whose output maps do not tile into 128-pixel rectangles): 128 consecutive
// output pixels in (n, h, w) order starting at base pixel (w, h, n) of the padded bounding box, shifted by the
// filter tap (off_w, off_h); pixels that fall into the padding are zero-filled by the TMA unit
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// one lane of a fully converged warp (the warp stays converged around it: the compiler can keep addresses,
// coordinates and descriptors in uniform registers for the uniform-datapath TMA / MMA instructions)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart):
// start address>>4 | LBO(ignored for swizzled K-major)=1 | SBO = 1024>>4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major, 128B-swizzled operand: start>>4 | LBO>>4 (distance between 64-element chunks along M/N) | SBO>>4 (distance
// between 8-row groups along K) | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// packed fp32 pairs (Blackwell FFMA2): two IEEE fp32 FMAs per issue slot
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float2 unpack_f32x2(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
// bf16x2 {hi, lo} = round-to-nearest-even of (max(hi,0), max(lo,0)): ReLU fused into the narrowing conversion
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// B_MN (train mode, data gradient): the B operand is read MN-major straight from the FORWARD layer's weights
// [Cout_f, KH, KW, Cin_f] — this GEMM's K runs over (flipped tap, Cout_f) and its N over Cin_f, which is the weights'
// contiguous dimension — so no transposed copy of the weights is ever made: a k-block is BLOCK_N/64 TMA boxes of
// 64 Cout_f rows x 64 Cin_f columns (the canonical MN-major SW128 layout, see wgrad_kernel).
template <int BLOCK_N, bool EPI_TMA, bool B_MN = false, bool OCC2 = false>
__global__ void __launch_bounds__(GEMM_THREADS, OCC2 ? 2 : 1) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                     const __grid_constant__ CUtensorMap tmap_b,
                                                                     const __grid_constant__ CUtensorMap tmap_out,
                                                                     const __grid_constant__ CUtensorMap tmap_res,
                                                                     const ConvGemmParams p) {
  using Cfg = TileCfg<BLOCK_N, EPI_TMA, OCC2>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment required by the 128B swizzle
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base_addr - raw_addr);
  constexpr int RING_BYTES = STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES;
  const uint32_t staging_base = base_addr + STAGES * Cfg::STAGE_BYTES;        // [8 warps][3][32 rows x 64 B]
  const uint32_t bar_base = base_addr + RING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto res_bar = [&](int ew, int b) { return bar_base + 8u * (2 * STAGES + 4 + ew * 3 + b); };   // per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + RING_BYTES + 8 * (2 * STAGES + 4 + 24));
  float* s_affine = reinterpret_cast<float*>(base_ptr + RING_BYTES + Cfg::BAR_BYTES);   // [2][2][BLOCK_N]

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.kh * p.kw * p.cin_blocks;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_nb;
  const int num_tiles = m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 8);   // one arrival per epilogue warp
    }
    if (EPI_TMA) {
      for (int b = 0; b < 24; ++b) mbar_init(res_bar(b / 3, b % 3), 1);
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
      if (p.residual != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_res) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, tensor-map prefetch) overlapped the tail of the previous kernel of the
  // stream; from here on its outputs are read (A operand, residual) and this layer's output is written
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: runs ahead across tile boundaries, bounded only by the smem ring.  The whole warp walks
    // the loop converged; one elected lane issues =====
    {
      // This one thread paces every layer with few MMAs per k-block (a 128x128x64 k-block is 256 tensor cycles): the
      // loop body is kept free of integer division / modulo — ring slot, phase, filter tap and channel block are
      // running counters — so that issuing a k-block costs tens, not hundreds, of dependent instructions.
      const uint32_t tx_bytes = (uint32_t)(p.tw * p.th * p.nb) * (BLOCK_K * 2) + (uint32_t)Cfg::B_BYTES;
      const int cin_blocks = p.cin_blocks, kw = p.kw, pad = p.pad;
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        if (p.tile_skip != nullptr && p.tile_skip[m_tile]) continue;
        int w0, h0, n0;
        if (p.im2col) {                       // base output pixel of this M tile, in padded-box coordinates
          const long long m0 = (long long)m_tile * BLOCK_M;
          const int hw = p.OH * p.OW;
          n0 = (int)(m0 / hw);
          const int rem = (int)(m0 - (long long)n0 * hw);
          h0 = rem / p.OW - pad;
          w0 = rem - (rem / p.OW) * p.OW - pad;
        } else if (p.flat) {
          w0 = m_tile * BLOCK_M; h0 = 0; n0 = 0;
        } else {
          const int tw_i = m_tile % p.tiles_w;
          const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
          const int tn_i = m_tile / (p.tiles_w * p.tiles_h);
          w0 = tw_i * p.tw - pad; h0 = th_i * p.th - pad; n0 = tn_i * p.nb;
        }
        const int b_row = n_tile * BLOCK_N;
        int cb = 0, sx = 0, r = 0;            // channel block, filter column, filter row of the current k-block
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t fb = full_bar(stage);
          if (elect_one()) {
            mbar_expect_tx(fb, tx_bytes);
            if (p.im2col)
              tma_load_im2col_4d(a_dst, &tmap_a, fb, cb * BLOCK_K, w0, h0, n0, (uint16_t)sx, (uint16_t)r);
            else
              tma_load_4d(a_dst, &tmap_a, fb, cb * BLOCK_K, w0 + sx, h0 + r, n0);
            if constexpr (B_MN) {
              const int ftap = (p.kh - 1 - r) * kw + (kw - 1 - sx);      // correlation with the flipped filter
#pragma unroll
              for (int nc = 0; nc < BLOCK_N / 64; ++nc)
                tma_load_2d(a_dst + Cfg::A_BYTES + nc * 8192, &tmap_b, fb, ftap * p.cout + b_row + nc * 64, cb * BLOCK_K);
            } else {
              tma_load_2d(a_dst + Cfg::A_BYTES, &tmap_b, fb, kb * BLOCK_K, b_row);
            }
          }
          __syncwarp();
          if (++cb == cin_blocks) {
            cb = 0;
            if (++sx == kw) {
              sx = 0;
              ++r;
            }
          }
          if (++stage == (uint32_t)STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: alternates between the two TMEM accumulators (warp converged, one elected lane issues) =====
    {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BLOCK_N, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (B_MN ? (1u << 16) : 0u) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                 ((uint32_t)(BLOCK_M >> 4) << 24);
      uint32_t stage = 0, phase = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (p.tile_skip != nullptr && p.tile_skip[tile / p.n_tiles]) continue;     // (tcount counts computed tiles only)
        const uint32_t acc = tcount & 1u;
        const uint32_t aph = (tcount >> 1) & 1u;
        ++tcount;
        mbar_wait(tmem_empty_bar(acc), aph ^ 1u);      // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t s = stage;
          mbar_wait(full_bar(s), phase);
          if (++stage == (uint32_t)STAGES) {
            stage = 0;
            phase ^= 1u;
          }
          tcgen05_fence_after();
          const uint32_t a_addr = base_addr + s * Cfg::STAGE_BYTES;
          const uint64_t da = make_sw128_desc(a_addr);
          const uint64_t db = make_sw128_desc(a_addr + Cfg::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the >>4 address field
              const uint64_t dbk = B_MN ? make_sw128_mn_desc(a_addr + Cfg::A_BYTES + k * 2048, 8192, 1024) : db + (uint64_t)(2 * k);
              umma_bf16(d_tmem, da + (uint64_t)(2 * k), dbk, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(s));          // frees the ring slot once these MMAs have read it
            if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));      // accumulator complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int ew = warp - 2;         // 0..7
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;        // which half of the tile's columns
    const int et = threadIdx.x - 64; // 0..255 within the epilogue group
    const int row = q * 32 + lane;   // output row inside the tile
    constexpr int COLS_PER_WARP = BLOCK_N >= 64 ? BLOCK_N / 2 : BLOCK_N;
    const bool has_cols = (BLOCK_N >= 64) || (half == 0);
    const int c_begin = (BLOCK_N >= 64) ? half * COLS_PER_WARP : 0;
    const int cout_store = p.out_mode >= 1 ? p.cout : p.out_ld;   // pitch padding is written as zeros
    const bool fused2 = !EPI_TMA && (BLOCK_N == 256) && (p.out_mode == 2);
    float* s_w2 = s_affine + 4 * BLOCK_N;          // [256][4] 1x1-conv weights (nc2 <= 4)
    float* s_part = s_w2 + 256 * 4;                // [128][4] partial logits of the upper column half
    if (fused2) {
      // column-pair layout: s_w2[(c/2)*8 + 2*j + (c&1)] = w2[j][c], so one FFMA2 feeds class j from two columns
      for (int i = et; i < 256 * 4; i += EPI_THREADS) {
        const int c = i >> 2, j = i & 3;
        s_w2[(c >> 1) * 8 + 2 * j + (c & 1)] = (j < p.nc2) ? __bfloat162float(p.w2[(size_t)j * p.cout + c]) : 0.f;
      }
    }
    if constexpr (EPI_TMA) {
      // ===== shared-memory epilogue (flat layers, bf16 out), one independent pipeline PER WARP: the warp's
      // [32 rows x 32 channels] chunks rotate through three 2 KB buffers: TMA-load the residual chunk g+1, update
      // chunk g in place (each lane owns one row: conflict-free 16-byte accesses under the 64B swizzle), TMA-store
      // it; the store of chunk g-1 is still draining.  Lane 0 issues the warp's bulk copies and owns its bulk
      // groups; only __syncwarp() is needed, no cross-warp barrier inside a tile.
      constexpr int NCH = COLS_PER_WARP / 32;                 // chunks per tile for this warp
      const bool elected = lane == 0;
      const bool has_res = p.residual != nullptr;
      constexpr uint32_t NBUF = Cfg::EPI_NBUF;
      const uint32_t warp_base = staging_base + (uint32_t)ew * NBUF * Cfg::EPI_CHUNK_BYTES;
      const uint32_t row_off = (uint32_t)lane * 64u;
      const uint32_t sw = (uint32_t)((lane >> 1) & 3);
      const int row0 = q * 32;                                // first tile row of this warp
      uint32_t g = 0, tcount = 0;
      if (has_cols && has_res && elected && (int)blockIdx.x < num_tiles) {   // residual of the very first chunk
        const int t0 = blockIdx.x;
        mbar_expect_tx(res_bar(ew, 0), Cfg::EPI_CHUNK_BYTES);
        tma_load_2d(warp_base, &tmap_res, res_bar(ew, 0), (t0 % p.n_tiles) * BLOCK_N + c_begin, (t0 / p.n_tiles) * BLOCK_M + row0);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        if (p.tile_skip != nullptr && p.tile_skip[m_tile]) continue;     // (never combined with a residual: host check)
        const uint32_t acc = tcount & 1u;
        const uint32_t aph = (tcount >> 1) & 1u;
        ++tcount;
        const int col_base = n_tile * BLOCK_N;
        float* t_scale = s_affine + acc * (2 * BLOCK_N);
        float* t_shift = t_scale + BLOCK_N;
        for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
          const int ch = col_base + c;
          t_scale[c] = ch < p.cout ? __ldg(p.scale + ch) : 0.f;
          t_shift[c] = ch < p.cout ? __ldg(p.shift + ch) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (has_cols) {
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + (uint32_t)c_begin;
          uint32_t vbuf[2][32];
          mbar_wait(tmem_full_bar(acc), aph);
          __syncwarp();
          tcgen05_fence_after();
          tmem_ld_32x32b_x32(t_addr, vbuf[0]);
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci, ++g) {
            const uint32_t b = g % NBUF;
            const uint32_t buf = warp_base + b * Cfg::EPI_CHUNK_BYTES;
            if (has_res) {
              if (elected) {
                // chunk g+1 (maybe of this CTA's next tile) -> buffer (g+1)%NBUF, last read by the store of chunk g+1-NBUF
                int nt = tile, nci = ci + 1;
                if (nci == NCH) { nt = tile + gridDim.x; nci = 0; }
                if (nt < num_tiles) {
                  bulk_wait_group_read<NBUF - 2>();
                  const uint32_t nb = (g + 1u) % NBUF;
                  mbar_expect_tx(res_bar(ew, nb), Cfg::EPI_CHUNK_BYTES);
                  tma_load_2d(warp_base + nb * Cfg::EPI_CHUNK_BYTES, &tmap_res, res_bar(ew, nb),
                              (nt % p.n_tiles) * BLOCK_N + c_begin + 32 * nci, (nt / p.n_tiles) * BLOCK_M + row0);
                }
              }
              mbar_wait(res_bar(ew, b), (g / NBUF) & 1u);
            } else {
              if (elected) bulk_wait_group_read<NBUF - 1>();   // the store of chunk g-NBUF has released this buffer
              __syncwarp();
            }
            tmem_ld_wait();                                     // chunk ci has landed
            if (ci + 1 < NCH) tmem_ld_32x32b_x32(t_addr + (uint32_t)(32 * (ci + 1)), vbuf[(ci + 1) & 1]);
            const uint32_t* v = vbuf[ci & 1];
            const int c = c_begin + 32 * ci;
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // 16-byte pieces = 8 channels
              const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
              const float4 s0 = *reinterpret_cast<const float4*>(t_scale + c + j * 8);
              const float4 s1 = *reinterpret_cast<const float4*>(t_scale + c + j * 8 + 4);
              const float4 t0 = *reinterpret_cast<const float4*>(t_shift + c + j * 8);
              const float4 t1 = *reinterpret_cast<const float4*>(t_shift + c + j * 8 + 4);
              const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
              float o[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fmaf(__uint_as_float(v[j * 8 + k]), sc[k], sh[k]);
              if (has_res) {
                const uint4 r4 = lds128(addr);
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  o[2 * k] += __uint_as_float(rw[k] << 16);
                  o[2 * k + 1] += __uint_as_float(rw[k] & 0xffff0000u);
                }
              }
              uint4 ov;
              if (p.relu) {
                ov = make_uint4(cvt_relu_bf16x2(o[0], o[1]), cvt_relu_bf16x2(o[2], o[3]), cvt_relu_bf16x2(o[4], o[5]),
                                cvt_relu_bf16x2(o[6], o[7]));
              } else {
                ov = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
              }
              sts128(addr, ov);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elected) {
              tma_store_2d(&tmap_out, buf, col_base + c, m_tile * BLOCK_M + row0);
              bulk_commit_group();
            }
          }
        } else {
          mbar_wait(tmem_full_bar(acc), aph);
          __syncwarp();
          tcgen05_fence_after();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
      }
      if (elected) bulk_wait_group_all();     // smem must outlive the last store; results visible at kernel end
    } else {
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      if (p.tile_skip != nullptr && p.tile_skip[m_tile]) continue;
      const uint32_t acc = tcount & 1u;
      const uint32_t aph = (tcount >> 1) & 1u;
      ++tcount;
      const int col_base = n_tile * BLOCK_N;
      int ch_base = col_base;          // channel index used for scale/shift and the store column
      int tap = 0;
      if (p.out_mode >= 1) {
        tap = col_base / p.cout;
        ch_base = col_base - tap * p.cout;
      }
      // stage this tile's per-channel affine in smem (overlaps the MMAs of the tile)
      float* t_scale = s_affine + acc * (2 * BLOCK_N);
      float* t_shift = t_scale + BLOCK_N;
      for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
        const int ch = ch_base + c;
        t_scale[c] = ch < p.cout ? __ldg(p.scale + ch) : 0.f;
        t_shift[c] = ch < p.cout ? __ldg(p.shift + ch) : 0.f;
      }
      int n, h, w;
      bool row_ok;
      if (p.flat) {
        // 1x1 stride-1 layers tile the flattened pixel index; decode (n,h,w) from the real geometry
        const long long m = (long long)m_tile * BLOCK_M + row;
        row_ok = m < p.M;
        const int hw = p.OH * p.OW;
        n = (int)(m / hw);
        const int rem = (int)(m - (long long)n * hw);
        h = rem / p.OW;
        w = rem - h * p.OW;
      } else {
        const int tw_i = m_tile % p.tiles_w;
        const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn_i = m_tile / (p.tiles_w * p.tiles_h);
        const int per_img = p.tw * p.th;
        const int ni = row / per_img;
        const int rem = row - ni * per_img;
        const int hi = rem / p.tw;
        const int wi = rem - hi * p.tw;
        n = tn_i * p.nb + ni; h = th_i * p.th + hi; w = tw_i * p.tw + wi;
        row_ok = (row < per_img * p.nb) && (n < p.N) && (h < p.OH) && (w < p.OW);
      }
      size_t out_off;
      if (p.out_mode >= 1) {
        const int oi = tap >> 1, oj = tap & 1;
        out_off = (((size_t)n * (2 * p.OH) + (2 * h + oi)) * (size_t)(2 * p.OW) + (2 * w + oj)) * (size_t)p.out_ld;
      } else {
        out_off = (((size_t)n * p.OH + h) * (size_t)p.OW + w) * (size_t)p.out_ld;
      }
      const __nv_bfloat16* res_row = nullptr;
      if (p.residual != nullptr && row_ok) {
        if (p.res_up2)
          res_row = p.residual + (((size_t)n * (p.OH >> 1) + (h >> 1)) * (size_t)(p.OW >> 1) + (w >> 1)) * (size_t)p.cout;
        else
          res_row = p.residual + (((size_t)n * p.OH + h) * (size_t)p.OW + w) * (size_t)p.cout;
      }
      // affine staged by all 256 epilogue threads -> visible to all of them
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (has_cols) {
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + (uint32_t)c_begin;
        uint32_t vbuf[2][32];
        uint64_t lacc[4] = {0ull, 0ull, 0ull, 0ull};   // fused 1x1: per class, (even-column, odd-column) partial sums
        uint4 rbuf[2][4];     // residual of the chunk, fetched one chunk ahead (scattered 16-byte loads)
        const bool res_vec = (res_row != nullptr);
        auto load_residual = [&](int ci, uint4* r) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ch = ch_base + c_begin + 32 * ci + g * 8;
            r[g] = (res_vec && ch + 8 <= p.cout) ? __ldg(reinterpret_cast<const uint4*>(res_row + ch)) : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        load_residual(0, rbuf[0]);                            // before the accumulator is even ready
        mbar_wait(tmem_full_bar(acc), aph);
        __syncwarp();
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(t_addr, vbuf[0]);
#pragma unroll
        for (int ci = 0; ci < COLS_PER_WARP / 32; ++ci) {
          tmem_ld_wait();                                   // chunk ci has landed
          if (ci + 1 < COLS_PER_WARP / 32) {
            tmem_ld_32x32b_x32(t_addr + (uint32_t)(32 * (ci + 1)), vbuf[(ci + 1) & 1]);
            load_residual(ci + 1, rbuf[(ci + 1) & 1]);
          }
          const uint32_t* v = vbuf[ci & 1];
          const uint4* rr4 = rbuf[ci & 1];
          const int c = c_begin + 32 * ci;                  // column inside the tile
          if (fused2) {
            // transposed-conv epilogue fused with the 1x1 conv: affine (FFMA2) -> ReLU + bf16 rounding in one
            // cvt -> 4 FFMA2 per column pair against smem-broadcast weights; nothing is stored per channel
            if (row_ok) {
              const bool unit_scale = p.unit_scale != 0;
#pragma unroll
              for (int g = 0; g < 8; ++g) {   // groups of 4 columns
                // per-channel affine in scalar fp32 (the accumulator registers of tcgen05.ld are not pair-aligned:
                // packing them for FFMA2 costs more moves than it saves); a unit BN scale (the transposed conv has
                // no BN) is a plain add of the bias: acc*1 + b == acc + b exactly
                const float4 t4 = *reinterpret_cast<const float4*>(t_shift + c + g * 4);
                float f0, f1, f2, f3;
                if (unit_scale) {
                  f0 = __uint_as_float(v[g * 4]) + t4.x;
                  f1 = __uint_as_float(v[g * 4 + 1]) + t4.y;
                  f2 = __uint_as_float(v[g * 4 + 2]) + t4.z;
                  f3 = __uint_as_float(v[g * 4 + 3]) + t4.w;
                } else {
                  const float4 s4 = *reinterpret_cast<const float4*>(t_scale + c + g * 4);
                  f0 = fmaf(__uint_as_float(v[g * 4]), s4.x, t4.x);
                  f1 = fmaf(__uint_as_float(v[g * 4 + 1]), s4.y, t4.y);
                  f2 = fmaf(__uint_as_float(v[g * 4 + 2]), s4.z, t4.z);
                  f3 = fmaf(__uint_as_float(v[g * 4 + 3]), s4.w, t4.w);
                }
                const uint32_t h01 = p.relu ? cvt_relu_bf16x2(f0, f1) : pack_bf16x2(f0, f1);
                const uint32_t h23 = p.relu ? cvt_relu_bf16x2(f2, f3) : pack_bf16x2(f2, f3);
                const uint64_t x01 = pack_f32x2(__uint_as_float(h01 << 16), __uint_as_float(h01 & 0xffff0000u));
                const uint64_t x23 = pack_f32x2(__uint_as_float(h23 << 16), __uint_as_float(h23 & 0xffff0000u));
                const float4* wp = reinterpret_cast<const float4*>(s_w2 + (size_t)((c + g * 4) >> 1) * 8);
                const float4 wa = wp[0], wb = wp[1], wc = wp[2], wd = wp[3];
                lacc[0] = ffma2(x01, pack_f32x2(wa.x, wa.y), lacc[0]);
                lacc[1] = ffma2(x01, pack_f32x2(wa.z, wa.w), lacc[1]);
                lacc[2] = ffma2(x01, pack_f32x2(wb.x, wb.y), lacc[2]);
                lacc[3] = ffma2(x01, pack_f32x2(wb.z, wb.w), lacc[3]);
                lacc[0] = ffma2(x23, pack_f32x2(wc.x, wc.y), lacc[0]);
                lacc[1] = ffma2(x23, pack_f32x2(wc.z, wc.w), lacc[1]);
                lacc[2] = ffma2(x23, pack_f32x2(wd.x, wd.y), lacc[2]);
                lacc[3] = ffma2(x23, pack_f32x2(wd.z, wd.w), lacc[3]);
              }
            }
          } else if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {   // groups of 8 columns
              const int ch = ch_base + c + g * 8;
              if (ch < cout_store) {
                float o[8];
                const float4 s0 = *reinterpret_cast<const float4*>(t_scale + c + g * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(t_scale + c + g * 8 + 4);
                const float4 t0 = *reinterpret_cast<const float4*>(t_shift + c + g * 8);
                const float4 t1 = *reinterpret_cast<const float4*>(t_shift + c + g * 8 + 4);
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = fmaf(__uint_as_float(v[g * 8 + k]), sc[k], sh[k]);
                if (res_row != nullptr) {
                  if (ch + 8 <= p.cout) {
                    const uint32_t rw[4] = {rr4[g].x, rr4[g].y, rr4[g].z, rr4[g].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      o[2 * k] += __uint_as_float(rw[k] << 16);
                      o[2 * k + 1] += __uint_as_float(rw[k] & 0xffff0000u);
                    }
                  } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                      if (ch + k < p.cout) o[k] += __bfloat162float(res_row[ch + k]);
                  }
                }
                if (p.relu) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
                }
                if (p.out_f32) {
                  float* dst = static_cast<float*>(p.out) + out_off + ch;
                  *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                  *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
                } else {
                  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + out_off + ch;
                  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                              pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
                }
              }
            }
          }
        }
        if (fused2) {
          // combine the two column halves of each row: half 1 -> smem, half 0 adds, bias, sigmoid, store
          float logit[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pr = unpack_f32x2(lacc[j]);
            logit[j] = pr.x + pr.y;
          }
          if (half == 1) *reinterpret_cast<float4*>(s_part + row * 4) = make_float4(logit[0], logit[1], logit[2], logit[3]);
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (half == 0 && row_ok) {
            const float4 pa = *reinterpret_cast<const float4*>(s_part + row * 4);
            const float other[4] = {pa.x, pa.y, pa.z, pa.w};
            float* dst = static_cast<float*>(p.out) + out_off;      // out_ld == nc2
            float sg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float z = logit[j] + other[j] + (j < p.nc2 ? __ldg(p.b2 + j) : 0.f);
              sg[j] = 1.0f / (1.0f + expf(-z));
            }
            if (p.nc2 == 4) {
              *reinterpret_cast<float4*>(dst) = make_float4(sg[0], sg[1], sg[2], sg[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < p.nc2) dst[j] = sg[j];
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");            // s_part reusable by the next tile
        }
      } else {
        mbar_wait(tmem_full_bar(acc), aph);
        __syncwarp();
        tcgen05_fence_after();
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
    }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// CUDA-core reference (tests only): same descriptor semantics, fp32 accumulation
// ---------------------------------------------------------------------------------------------
struct SimtParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
  int N, H, W, Cin, KH, KW, stride, pad, OH, OW, cout, cout_total, relu, res_up2, out_f32, out_mode, out_ld;
};

__global__ void conv_simt_kernel(SimtParams p) {
  const size_t total = (size_t)p.N * p.OH * p.OW * p.cout_total;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % p.cout_total);
    size_t pix = idx / p.cout_total;
    const int ow = (int)(pix % p.OW);
    pix /= p.OW;
    const int oh = (int)(pix % p.OH);
    const int n = (int)(pix / p.OH);
    const int K = p.KH * p.KW * p.Cin;
    const __nv_bfloat16* wrow = p.w + (size_t)col * K;
    float acc = 0.f;
    for (int r = 0; r < p.KH; ++r) {
      const int ih = oh * p.stride + r - p.pad;
      if (ih < 0 || ih >= p.H) continue;
      for (int s = 0; s < p.KW; ++s) {
        const int iw = ow * p.stride + s - p.pad;
        if (iw < 0 || iw >= p.W) continue;
        const __nv_bfloat16* xp = p.x + (((size_t)n * p.H + ih) * p.W + iw) * p.Cin;
        const __nv_bfloat16* wp = wrow + (size_t)(r * p.KW + s) * p.Cin;
        for (int c = 0; c < p.Cin; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
      }
    }
    int ch = col;
    size_t off;
    if (p.out_mode == 1) {
      const int tap = col / p.cout;
      ch = col - tap * p.cout;
      off = (((size_t)n * (2 * p.OH) + (2 * oh + (tap >> 1))) * (size_t)(2 * p.OW) + (2 * ow + (tap & 1))) * (size_t)p.out_ld + ch;
    } else {
      off = (((size_t)n * p.OH + oh) * (size_t)p.OW + ow) * (size_t)p.out_ld + ch;
    }
    float v = fmaf(acc, p.scale[ch], p.shift[ch]);
    if (p.residual) {
      size_t ro = p.res_up2 ? (((size_t)n * (p.OH >> 1) + (oh >> 1)) * (size_t)(p.OW >> 1) + (ow >> 1)) * (size_t)p.cout + ch
                            : (((size_t)n * p.OH + oh) * (size_t)p.OW + ow) * (size_t)p.cout + ch;
      v += __bfloat162float(p.residual[ro]);
    }
    if (p.relu) v = fmaxf(v, 0.f);
    if (p.out_f32)
      static_cast<float*>(p.out)[off] = v;
    else
      static_cast<__nv_bfloat16*>(p.out)[off] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(sym);
  });
  return fn;
}

// NHWC activation as (C, W, H, N); the base pixel traverses the W x H output positions of a SAME conv:
// lower corner = -pad, upper corner = pad - (k - 1); 64 channels x 128 pixels per load
int encode_map_im2col(CUtensorMap* map, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides_bytes, int pad,
                      int ksize, int pixels_per_column = BLOCK_M) {
  EncodeIm2colFn fn = get_encode_im2col_fn();
  if (!fn) {
    mrcnn_set_error("cuTensorMapEncodeIm2col unavailable (no CUDA driver?)");
    return MRCNN_ERR_CUDA;
  }
  const int lower[2] = {-pad, -pad};
  const int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides_bytes, lower, upper,
                  (cuuint32_t)BLOCK_K, (cuuint32_t)pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mrcnn_set_error("cuTensorMapEncodeIm2col failed (%d): dims [%llu,%llu,%llu,%llu]", (int)r, (unsigned long long)dims[0],
                    (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3]);
    return MRCNN_ERR_CUDA;
  }
  return MRCNN_OK;
}

int encode_map(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    mrcnn_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return MRCNN_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mrcnn_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r, rank,
                    (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                    (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return MRCNN_ERR_CUDA;
  }
  return MRCNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Weight gradient (train mode): dW[co, tap, ci] += sum_p dY[p, co] * X[p + tap, ci]
//
// GEMM with M = Cout, N = Cin (one filter tap per tile), K = pixels.  Both operands are read straight from the NHWC
// tensors, whose contiguous dimension is the GEMM's M (dY) or N (X): MN-major UMMA operands.  A TMA box of 64 pixels x
// 64 channels lands as 64 rows of 128 B (128B swizzle, 8-row atoms of 1024 B): exactly the canonical MN-major SW128
// layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units with SBO = 1024 B between 8-pixel groups and LBO = 8192 B
// between 64-channel chunks.  X is fetched per tap with the same im2col-mode loads as the forward pass (pixels that
// fall into the SAME padding arrive as zeros), 64 consecutive pixels per load.  K is split over CTAs (the pixel count
// is 10^3..10^5, the tile count 4..300) and partial tiles are accumulated with float32 red.add into dW, which has the
// layout of the parameter itself, [Cout, KH, KW, Cin].
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer   warps 2-5: TMEM -> red.global.add.f32
// ---------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;
constexpr int WG_BLOCK_K = 64;          // pixels per stage
constexpr int WG_CHUNK_BYTES = 64 * 128;   // 64 pixels x 64 channels bf16

struct WgradParams {
  int cout, cin, taps, kw, pad, OH, OW;
  int im2col;             // 3x3: X through im2col-mode loads
  int k_chunks;           // ceil(P / 64)
  int chunks_per_split;   // K partition
  int m_tiles, n_tiles, ksplit;
  float* out;             // [cout, taps*cin] float32
  int out_ld;
  const __nv_bfloat16* w; // optional: the layer's weights [cout, taps*cin] ...
  float* wdot;            // ... and wdot[co] += <w[co, :], dW tile> (gradient of a per-channel scale folded behind the conv)
};

template <int BLOCK_N> struct WgradCfg {
  static constexpr int A_BYTES = 2 * WG_CHUNK_BYTES;                 // M = 128 channels of dY
  static constexpr int B_BYTES = (BLOCK_N / 64) * WG_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 128 ? 6 : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                                                               const __grid_constant__ CUtensorMap tmap_x, const WgradParams p) {
  using Cfg = WgradCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base_addr - raw_addr);
  const uint32_t bar_base = base_addr + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 1));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile of this CTA: (split, tap, n_tile, m_tile), m fastest
  int t = blockIdx.x;
  const int m_tile = t % p.m_tiles; t /= p.m_tiles;
  const int n_tile = t % p.n_tiles; t /= p.n_tiles;
  const int tap = t % p.taps;
  const int split = t / p.taps;
  const int kc0 = split * p.chunks_per_split;
  const int kc1 = min(p.k_chunks, kc0 + p.chunks_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)BLOCK_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    const int r = tap / p.kw, sx = tap - r * p.kw;
    const int hw = p.OH * p.OW;
    uint32_t stage = 0, phase = 0;
    for (int kc = kc0; kc < kc1; ++kc) {
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t a_dst = base_addr + stage * Cfg::STAGE_BYTES;
      const uint32_t fb = full_bar(stage);
      if (elect_one()) {
        mbar_expect_tx(fb, (uint32_t)Cfg::STAGE_BYTES);
        const int p0 = kc * WG_BLOCK_K;
#pragma unroll
        for (int mc = 0; mc < 2; ++mc) tma_load_2d(a_dst + mc * WG_CHUNK_BYTES, &tmap_dy, fb, m_tile * BLOCK_M + mc * 64, p0);
        if (p.im2col) {
          const int n0 = p0 / hw, rem = p0 - n0 * hw;
          const int h0 = rem / p.OW - p.pad, w0 = rem - (rem / p.OW) * p.OW - p.pad;
#pragma unroll
          for (int nc = 0; nc < BLOCK_N / 64; ++nc)
            tma_load_im2col_4d(a_dst + Cfg::A_BYTES + nc * WG_CHUNK_BYTES, &tmap_x, fb, n_tile * BLOCK_N + nc * 64, w0, h0, n0,
                               (uint16_t)sx, (uint16_t)r);
        } else {
#pragma unroll
          for (int nc = 0; nc < BLOCK_N / 64; ++nc)
            tma_load_2d(a_dst + Cfg::A_BYTES + nc * WG_CHUNK_BYTES, &tmap_x, fb, n_tile * BLOCK_N + nc * 64, p0);
        }
      }
      __syncwarp();
      if (++stage == (uint32_t)STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // D = f32, A = B = bf16, both MN-major (bits 15 / 16), N = BLOCK_N, M = 128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
    uint32_t stage = 0, phase = 0;
    for (int kc = kc0; kc < kc1; ++kc) {
      mbar_wait(full_bar(stage), phase);
      tcgen05_fence_after();
      const uint32_t a_addr = base_addr + stage * Cfg::STAGE_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WG_BLOCK_K / UMMA_K; ++k) {
          // 16 pixels further along K = 16 rows of 128 B
          const uint64_t da = make_sw128_mn_desc(a_addr + k * 2048, WG_CHUNK_BYTES, 1024);
          const uint64_t db = make_sw128_mn_desc(a_addr + Cfg::A_BYTES + k * 2048, WG_CHUNK_BYTES, 1024);
          umma_bf16(tmem_base, da, db, idesc, (kc > kc0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (kc == kc1 - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == (uint32_t)STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (kc1 > kc0) {
    // epilogue: warp w may read TMEM lanes 32*(w % 4)..+31 = rows (output channels) of the tile
    const int q = warp & 3;
    const int co = m_tile * BLOCK_M + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
    const size_t roff = (size_t)co * p.out_ld + (size_t)tap * p.cin + (size_t)n_tile * BLOCK_N;
    float* orow = p.out + roff;
    float dot = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (co < p.cout) {       // cin is a multiple of 64: the 32 columns of a chunk are all real
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(orow + c0 + j), make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        if (p.wdot) {
          const uint4* wr = reinterpret_cast<const uint4*>(p.w + roff + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 ww = __ldg(wr + j);
            const uint32_t wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              dot += __uint_as_float(v[8 * j + 2 * k]) * __uint_as_float(wv[k] << 16);
              dot += __uint_as_float(v[8 * j + 2 * k + 1]) * __uint_as_float(wv[k] & 0xffff0000u);
            }
          }
        }
      }
    }
    if (p.wdot && co < p.cout) atomicAdd(p.wdot + co, dot);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N) : "memory");
  }
}

template <int BN> int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, cudaStream_t st) {
  using Cfg = WgradCfg<BN>;
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_done[dev]) {
    MRCNN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev] = true;
  }
  const unsigned grid = (unsigned)(p.m_tiles * p.n_tiles * p.taps * p.ksplit);
  MRCNN_CHECK_CUDA(mrcnn_launch(wgrad_kernel<BN>, dim3(grid), dim3(WG_THREADS), Cfg::SMEM_BYTES, st, tdy, tx, p));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

template <int BN, bool EPI, bool BMN = false, bool OCC2 = false> int launch_tile(const ConvPlan* plan, cudaStream_t st) {
  using Cfg = TileCfg<BN, EPI, OCC2>;
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_done[dev]) {
    MRCNN_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, EPI, BMN, OCC2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev] = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned tiles = plan->grid.x;
  const unsigned slots = (unsigned)num_sms * (OCC2 ? 2u : 1u);                   // persistent CTAs: one (OCC2: two) per SM
  const unsigned grid = tiles < slots ? tiles : slots;
  MRCNN_CHECK_CUDA(mrcnn_launch(conv_gemm_kernel<BN, EPI, BMN, OCC2>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, plan->tmap_a,
                                plan->tmap_b, plan->tmap_out, plan->tmap_res, plan->p));
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

int validate_desc(const mrcnn_conv_desc* d, int* OH, int* OW) {
  MRCNN_REQUIRE(d != nullptr, "conv2d: null descriptor");
  MRCNN_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "conv2d: empty tensor");
  MRCNN_REQUIRE(d->cin % 64 == 0, "conv2d: Cin=%d must be a multiple of 64", d->cin);
  const bool k1 = d->kh == 1 && d->kw == 1 && d->pad == 0 && (d->stride == 1 || d->stride == 2);
  const bool k3 = d->kh == 3 && d->kw == 3 && d->pad == 1 && d->stride == 1;
  MRCNN_REQUIRE(k1 || k3, "conv2d: unsupported filter %dx%d stride %d pad %d", d->kh, d->kw, d->stride, d->pad);
  *OH = (d->h + 2 * d->pad - d->kh) / d->stride + 1;
  *OW = (d->w + 2 * d->pad - d->kw) / d->stride + 1;
  if (d->stride == 2) MRCNN_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, "conv2d: stride 2 needs even H, W");
  if (d->residual_upsample2) MRCNN_REQUIRE(*OH % 2 == 0 && *OW % 2 == 0, "conv2d: upsampled residual needs even output size");
  MRCNN_REQUIRE(d->out_dtype == MRCNN_DTYPE_F32 || d->out_dtype == MRCNN_DTYPE_BF16, "conv2d: bad out_dtype");
  MRCNN_REQUIRE(d->out_mode == 0 || d->out_mode == 1, "conv2d: bad out_mode");
  const int ld = d->out_ld ? d->out_ld : d->cout;
  MRCNN_REQUIRE(ld % 8 == 0 && ld >= d->cout, "conv2d: out_ld=%d must be a multiple of 8 and >= cout", ld);
  if (d->out_mode == 1) MRCNN_REQUIRE(d->cout % 32 == 0, "conv2d: deconv mode needs cout %% 32 == 0");
  return MRCNN_OK;
}

}  // namespace

int conv_plan_create(const mrcnn_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                     const void* residual, void* out, int block_n, ConvPlan* plan) {
  return conv_plan_create_ex(d, x, w, scale, shift, residual, out, block_n, -1, plan);
}

bool conv_plan_epi_tma_eligible(const mrcnn_conv_desc* d) {
  const int ld = d->out_ld ? d->out_ld : d->cout;
  const bool flat = (d->kh == 1 && d->stride == 1) || d->kh == 3;     // 3x3 switches to im2col mode (flat M tiles)
  return flat && d->out_mode == 0 && d->out_dtype == MRCNN_DTYPE_BF16 && ld == d->cout && d->cout % 8 == 0 &&
         !d->residual_upsample2;
}

int conv_plan_create_ex(const mrcnn_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                        const void* residual, void* out, int block_n, int epi_tma, ConvPlan* plan) {
  int OH, OW;
  int rc = validate_desc(d, &OH, &OW);
  if (rc) return rc;
  MRCNN_REQUIRE(x && w && scale && shift && out && plan, "conv2d: null pointer");
  ConvGemmParams& p = plan->p;
  const int cout_total = d->out_mode == 1 ? 4 * d->cout : d->cout;
  if (block_n == 0) {
    // widest tile that still leaves >= 2 tiles per SM (fewer A re-reads, less per-tile overhead)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long m_tiles_est = ((long long)d->n * OH * OW + BLOCK_M - 1) / BLOCK_M;
    if (cout_total <= 32) block_n = 32;
    else if (cout_total <= 64) block_n = 64;
    else if (cout_total % 256 == 0 && m_tiles_est * (cout_total / 256) >= 2LL * sms) block_n = 256;
    else block_n = 128;
  }
  if (plan->b_mn) {       // tile width must divide the N extent (a box past it would read the next filter tap, not zeros)
    while (block_n > 64 && d->cout % block_n != 0) block_n >>= 1;
    if (block_n < 64) block_n = 64;
  }
  if (const char* fb = getenv("MRCNN_B200_BLOCK_N")) {   // development probe (tools/conv_probe.py): force the tile width
    const int v = atoi(fb);
    if ((v == 32 || v == 64 || v == 128 || v == 256) && !plan->b_mn && d->out_mode == 0) block_n = v;
  }
  plan->occ2 = 0;
  if (const char* o2 = getenv("MRCNN_B200_OCC2")) {     // "2": force the two-CTAs-per-SM variant (tests); the engine autotunes it
    if (o2[0] == '2' && !plan->b_mn && d->out_mode == 0) {
      plan->occ2 = 1;
      if (block_n > 128) block_n = 128;
    }
  }
  MRCNN_REQUIRE(block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256, "conv2d: block_n must be 32/64/128/256");
  if (d->out_mode == 1) MRCNN_REQUIRE(d->cout % block_n == 0, "conv2d: deconv cout %% block_n != 0");
  plan->block_n = block_n;

  // ---- A view: 4-D (C, W', H', N) where stride-2 1x1 convs see the sub-sampled view ----------
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4];
  // 1x1 stride 1 == plain [M, Cin] GEMM (rows need no (n,h,w) decode unless the residual is upsampled)
  const bool flat = (d->kh == 1 && d->stride == 1);
  p.flat = flat ? 1 : 0;
  p.OW = OW; p.OH = OH; p.N = d->n;
  p.M = (long long)d->n * OH * OW;
  int ext_w, ext_h, ext_n;   // extents the M tiles cover
  if (flat) {
    const unsigned long long M = (unsigned long long)p.M;
    dims[0] = d->cin; dims[1] = M; dims[2] = 1; dims[3] = 1;
    strides[0] = (cuuint64_t)d->cin * 2; strides[1] = strides[0] * M; strides[2] = strides[1];
    p.tw = 128; p.th = 1; p.nb = 1;
    MRCNN_REQUIRE(M < (1ull << 31), "conv2d: M too large");
    ext_w = (int)M; ext_h = 1; ext_n = 1;
  } else {
    dims[0] = d->cin; dims[1] = OW; dims[2] = OH; dims[3] = d->n;
    if (d->kh == 3) { dims[1] = d->w; dims[2] = d->h; }
    strides[0] = (cuuint64_t)d->cin * 2 * d->stride;
    strides[1] = (cuuint64_t)d->cin * 2 * d->w * d->stride;
    strides[2] = (cuuint64_t)d->cin * 2 * d->w * d->h;
    p.tw = OW < 128 ? OW : 128;
    p.th = 128 / p.tw; if (p.th > OH) p.th = OH;
    p.nb = (p.th == OH) ? 128 / (p.tw * p.th) : 1;
    if (p.nb < 1) p.nb = 1;
    if (p.nb > d->n) p.nb = d->n;
    ext_w = OW; ext_h = OH; ext_n = d->n;
  }
  p.tiles_w = ceil_div(ext_w, p.tw);
  p.tiles_h = ceil_div(ext_h, p.th);
  p.tiles_nb = ceil_div(ext_n, p.nb);
  // 3x3 convs whose maps do not tile into full 128-pixel rectangles (the 14x14 mask-head maps: 126 of 128 rows
  // per tile and 18 tile rows for 14 map rows = 77 % useful MMA rows) switch to TMA im2col mode, where an M tile
  // is 128 consecutive output pixels across row and image boundaries
  // epilogue through shared memory + TMA (flat bf16 layers): -1 = policy (env MRCNN_B200_EPI_TMA=0/1 overrides;
  // default on for layers with few k-blocks per tile, which are bound by the epilogue's global traffic)
  {
    const bool eligible = conv_plan_epi_tma_eligible(d);
    const char* env = getenv("MRCNN_B200_EPI_TMA");
    int want = epi_tma;
    if (want < 0 && env && (env[0] == '0' || env[0] == '1')) want = env[0] - '0';
    if (want < 0) want = (d->kh * d->kw * (d->cin / 64) <= 16) ? 1 : 0;
    plan->epi_tma = (want && eligible) ? 1 : 0;
  }
  p.im2col = 0;
  if (d->kh == 3) {
    const double useful = (double)p.M / ((double)p.tiles_w * p.tiles_h * p.tiles_nb * BLOCK_M);
    const char* env = getenv("MRCNN_B200_IM2COL");          // "0" never, "1" every 3x3 conv, default: when it pays
    const bool force_on = (env && env[0] == '1') || plan->epi_tma, force_off = env && env[0] == '0' && !plan->epi_tma;
    if (!force_off && (force_on || useful < 0.95)) {
      MRCNN_REQUIRE((unsigned long long)p.M < (1ull << 31), "conv2d: M too large");
      p.im2col = 1;
      p.flat = 1;                 // epilogue rows decode (n, h, w) from the flattened pixel index
      p.tw = 128; p.th = 1; p.nb = 1;
      p.tiles_w = (int)((p.M + BLOCK_M - 1) / BLOCK_M);
      p.tiles_h = 1; p.tiles_nb = 1;
    }
  }
  if (p.im2col) {
    rc = encode_map_im2col(&plan->tmap_a, x, dims, strides, d->pad, d->kh);
  } else {
    box[0] = 64; box[1] = p.tw; box[2] = p.th; box[3] = p.nb;
    rc = encode_map(&plan->tmap_a, x, 4, dims, strides, box);
  }
  if (rc) return rc;

  // ---- B: weights [cout_total, K] K-major; rows beyond cout_total are OOB -> zero fill --------
  const unsigned long long K = (unsigned long long)d->kh * d->kw * d->cin;
  if (plan->b_mn) {
    // data gradient: w is the FORWARD layer's [Cout_f = d->cin rows, KH*KW*Cin_f columns] matrix, read in 64 x 64 boxes
    MRCNN_REQUIRE(block_n >= 64 && d->cout % block_n == 0 && d->out_mode == 0, "conv2d (dgrad): Cin of the forward layer must be a "
                  "multiple of the tile width (%d)", block_n);
    const unsigned long long cols = (unsigned long long)d->kh * d->kw * d->cout;
    cuuint64_t bdims[2] = {cols, (cuuint64_t)d->cin};
    cuuint64_t bstr[1] = {cols * 2};
    cuuint32_t bbox[2] = {64, 64};
    rc = encode_map(&plan->tmap_b, w, 2, bdims, bstr, bbox);
    if (rc) return rc;
  } else {
  cuuint64_t bdims[2] = {K, (cuuint64_t)cout_total};
  cuuint64_t bstr[1] = {K * 2};
  cuuint32_t bbox[2] = {64, (cuuint32_t)block_n};
  rc = encode_map(&plan->tmap_b, w, 2, bdims, bstr, bbox);
  if (rc) return rc;
  }

  memset(&plan->tmap_out, 0, sizeof(CUtensorMap));
  memset(&plan->tmap_res, 0, sizeof(CUtensorMap));
  if (plan->epi_tma) {
    cuuint64_t odims[2] = {(cuuint64_t)d->cout, (cuuint64_t)p.M};
    cuuint64_t ostr[1] = {(cuuint64_t)d->cout * 2};
    cuuint32_t obox[2] = {32, 32};      // one epilogue warp's chunk: 32 channels x 32 rows
    rc = encode_map(&plan->tmap_out, out, 2, odims, ostr, obox, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (residual) {
      rc = encode_map(&plan->tmap_res, residual, 2, odims, ostr, obox, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  }

  p.kh = d->kh; p.kw = d->kw; p.pad = d->pad;
  p.cin_blocks = d->cin / 64;
  p.cout = d->cout;
  p.cout_total = cout_total;
  p.relu = d->relu;
  p.res_up2 = d->residual_upsample2;
  p.out_f32 = d->out_dtype == MRCNN_DTYPE_F32;
  p.out_mode = d->out_mode;
  p.out_ld = d->out_ld ? d->out_ld : d->cout;
  p.scale = scale;
  p.shift = shift;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out = out;
  p.w2 = nullptr;
  p.b2 = nullptr;
  p.nc2 = 0;
  p.unit_scale = 0;
  p.tile_skip = nullptr;
  if (residual) MRCNN_REQUIRE(d->cout % 8 == 0, "conv2d: residual needs cout %% 8 == 0");
  p.n_tiles = ceil_div(cout_total, block_n);
  const long long ctas = (long long)p.n_tiles * p.tiles_w * p.tiles_h * p.tiles_nb;
  MRCNN_REQUIRE(ctas < 2147483647LL, "conv2d: grid too large");
  plan->grid = dim3((unsigned)ctas, 1, 1);
  plan->flops = 2.0 * (double)d->n * OH * OW * (double)cout_total * (double)K;
  return MRCNN_OK;
}


int conv_plan_fuse_mask_logits(ConvPlan* plan, const void* w2, const float* b2, int nc2, void* out, int unit_scale) {
  MRCNN_REQUIRE(plan && w2 && b2 && out, "fuse_mask_logits: null pointer");
  MRCNN_REQUIRE(plan->block_n == 256 && plan->p.out_mode == 1 && plan->p.cout == 256 && !plan->epi_tma,
                "fuse_mask_logits: needs a 256-channel transposed-conv plan with 256-wide tiles");
  MRCNN_REQUIRE(nc2 >= 1 && nc2 <= 4, "fuse_mask_logits: nc2=%d outside [1,4]", nc2);
  plan->p.out_mode = 2;
  plan->p.w2 = static_cast<const __nv_bfloat16*>(w2);
  plan->p.b2 = b2;
  plan->p.nc2 = nc2;
  plan->p.unit_scale = unit_scale;
  plan->p.out = out;
  plan->p.out_ld = nc2;
  plan->flops += 2.0 * (double)plan->p.M * 4.0 * 256.0 * nc2;   // the fused 1x1 conv
  return MRCNN_OK;
}

int conv_plan_set_tile_skip(ConvPlan* plan, const unsigned char* flags) {
  MRCNN_REQUIRE(plan, "set_tile_skip: null plan");
  MRCNN_REQUIRE(plan->p.flat && plan->p.tw == 128 && plan->p.th == 1 && plan->p.nb == 1 && plan->p.residual == nullptr && !plan->b_mn,
                "set_tile_skip: needs a flat / im2col layer (M tiles = 128 consecutive rows) without residual");
  plan->p.tile_skip = flags;
  return MRCNN_OK;
}

int conv_plan_launch(const ConvPlan* plan, cudaStream_t st) {
  if (plan->b_mn) {
    if (plan->epi_tma) {
      switch (plan->block_n) {
        case 64: return launch_tile<64, true, true>(plan, st);
        case 128: return launch_tile<128, true, true>(plan, st);
        case 256: return launch_tile<256, true, true>(plan, st);
      }
    } else {
      switch (plan->block_n) {
        case 64: return launch_tile<64, false, true>(plan, st);
        case 128: return launch_tile<128, false, true>(plan, st);
        case 256: return launch_tile<256, false, true>(plan, st);
      }
    }
    mrcnn_set_error("conv2d (dgrad): bad block_n %d", plan->block_n);
    return MRCNN_ERR_INVALID;
  }
  if (plan->occ2 && plan->block_n <= 128 && plan->p.out_mode != 2) {     // two CTAs per SM (autotuned per layer)
    if (plan->epi_tma) {
      switch (plan->block_n) {
        case 32: return launch_tile<32, true, false, true>(plan, st);
        case 64: return launch_tile<64, true, false, true>(plan, st);
        case 128: return launch_tile<128, true, false, true>(plan, st);
      }
    } else {
      switch (plan->block_n) {
        case 32: return launch_tile<32, false, false, true>(plan, st);
        case 64: return launch_tile<64, false, false, true>(plan, st);
        case 128: return launch_tile<128, false, false, true>(plan, st);
      }
    }
  }
  if (plan->epi_tma) {
    switch (plan->block_n) {
      case 32: return launch_tile<32, true>(plan, st);
      case 64: return launch_tile<64, true>(plan, st);
      case 128: return launch_tile<128, true>(plan, st);
      case 256: return launch_tile<256, true>(plan, st);
    }
  } else {
    switch (plan->block_n) {
      case 32: return launch_tile<32, false>(plan, st);
      case 64: return launch_tile<64, false>(plan, st);
      case 128: return launch_tile<128, false>(plan, st);
      case 256: return launch_tile<256, false>(plan, st);
    }
  }
  mrcnn_set_error("conv2d: bad block_n %d", plan->block_n);
  return MRCNN_ERR_INVALID;
}

extern "C" int mrcnn_conv2d_bf16(const mrcnn_conv_desc* desc, const void* x, const void* w, const float* scale,
                                 const float* shift, const void* residual, void* out, void* stream) {
  ConvPlan plan;
  int rc = conv_plan_create(desc, x, w, scale, shift, residual, out, 0, &plan);
  if (rc) return rc;
  return conv_plan_launch(&plan, static_cast<cudaStream_t>(stream));
}

extern "C" int mrcnn_conv2d_bf16_simt(const mrcnn_conv_desc* d, const void* x, const void* w, const float* scale,
                                      const float* shift, const void* residual, void* out, void* stream) {
  int OH, OW;
  int rc = validate_desc(d, &OH, &OW);
  if (rc) return rc;
  SimtParams p;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.w = static_cast<const __nv_bfloat16*>(w);
  p.scale = scale; p.shift = shift;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out = out;

  p.N = d->n; p.H = d->h; p.W = d->w; p.Cin = d->cin; p.KH = d->kh; p.KW = d->kw; p.stride = d->stride; p.pad = d->pad;
  p.OH = OH; p.OW = OW; p.cout = d->cout; p.cout_total = d->out_mode == 1 ? 4 * d->cout : d->cout;
  p.relu = d->relu; p.res_up2 = d->residual_upsample2; p.out_f32 = d->out_dtype == MRCNN_DTYPE_F32;
  p.out_mode = d->out_mode; p.out_ld = d->out_ld ? d->out_ld : d->cout;
  const size_t total = (size_t)p.N * OH * OW * p.cout_total;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  conv_simt_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  return MRCNN_OK;
}

extern "C" int mrcnn_conv2d_wgrad_bf16(const mrcnn_conv_desc* d, const void* x, const void* dy, float* dw, const void* w,
                                       float* wdot, void* stream) {
  MRCNN_REQUIRE(d && x && dy && dw, "conv2d_wgrad: null pointer");
  MRCNN_REQUIRE((w == nullptr) == (wdot == nullptr), "conv2d_wgrad: w and wdot go together");
  const bool k1 = d->kh == 1 && d->kw == 1 && d->pad == 0 && d->stride == 1;
  const bool k3 = d->kh == 3 && d->kw == 3 && d->pad == 1 && d->stride == 1;
  MRCNN_REQUIRE(k1 || k3, "conv2d_wgrad: 1x1 stride 1 or 3x3 stride 1 pad 1 only (got %dx%d stride %d pad %d)", d->kh, d->kw,
                d->stride, d->pad);
  MRCNN_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "conv2d_wgrad: empty tensor");
  MRCNN_REQUIRE(d->cin % 64 == 0 && d->cout % 8 == 0, "conv2d_wgrad: Cin %% 64 and Cout %% 8 (got %d, %d)", d->cin, d->cout);
  const long long P = (long long)d->n * d->h * d->w;
  MRCNN_REQUIRE(P < (1ll << 31), "conv2d_wgrad: too many pixels");
  CUtensorMap tdy, tx;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->cout, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)d->cout * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)WG_BLOCK_K};
    int rc = encode_map(&tdy, dy, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (k3) {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->cin * 2, (cuuint64_t)d->cin * 2 * d->w, (cuuint64_t)d->cin * 2 * d->w * d->h};
    int rc = encode_map_im2col(&tx, x, dims, strides, d->pad, d->kh, WG_BLOCK_K);
    if (rc) return rc;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->cin, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)d->cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)WG_BLOCK_K};
    int rc = encode_map(&tx, x, 2, dims, strides, box);
    if (rc) return rc;
  }
  WgradParams p;
  p.cout = d->cout;
  p.cin = d->cin;
  p.taps = d->kh * d->kw;
  p.kw = d->kw;
  p.pad = d->pad;
  p.OH = d->h;
  p.OW = d->w;
  p.im2col = k3 ? 1 : 0;
  p.k_chunks = (int)((P + WG_BLOCK_K - 1) / WG_BLOCK_K);
  const int bn = d->cin % 128 == 0 ? 128 : 64;
  p.m_tiles = ceil_div(d->cout, BLOCK_M);
  p.n_tiles = d->cin / bn;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.m_tiles * p.n_tiles * p.taps;
  int ksplit = ceil_div(2 * sms, tiles);                     // about two waves of CTAs ...
  const int max_split = p.k_chunks >= 8 ? p.k_chunks / 4 : 1;   // ... of at least 4 k-chunks each
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  p.chunks_per_split = ceil_div(p.k_chunks, ksplit);
  p.ksplit = ceil_div(p.k_chunks, p.chunks_per_split);
  p.out = dw;
  p.out_ld = p.taps * d->cin;
  p.w = static_cast<const __nv_bfloat16*>(w);
  p.wdot = wdot;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return bn == 128 ? launch_wgrad<128>(tdy, tx, p, st) : launch_wgrad<64>(tdy, tx, p, st);
}

extern "C" int mrcnn_conv2d_dgrad_bf16(const mrcnn_conv_desc* fwd, const void* dy, const void* w, const float* ones,
                                       const float* zeros, void* dx, void* stream) {
  MRCNN_REQUIRE(fwd && dy && w && ones && zeros && dx, "conv2d_dgrad: null pointer");
  const bool k1 = fwd->kh == 1 && fwd->kw == 1 && fwd->pad == 0 && fwd->stride == 1;
  const bool k3 = fwd->kh == 3 && fwd->kw == 3 && fwd->pad == 1 && fwd->stride == 1;
  MRCNN_REQUIRE(k1 || k3, "conv2d_dgrad: 1x1 stride 1 or 3x3 stride 1 pad 1 only");
  MRCNN_REQUIRE(fwd->cin % 64 == 0 && fwd->cout % 64 == 0, "conv2d_dgrad: Cin and Cout must be multiples of 64");
  mrcnn_conv_desc d = *fwd;          // the data gradient is a stride-1 convolution of dy [N,H,W,Cout_f] giving dx [N,H,W,Cin_f]
  d.cin = fwd->cout;
  d.cout = fwd->cin;
  d.relu = 0;
  d.residual_upsample2 = 0;
  d.out_dtype = MRCNN_DTYPE_BF16;
  d.out_mode = 0;
  d.out_ld = 0;
  ConvPlan plan;
  plan.b_mn = 1;
  int rc = conv_plan_create(&d, dy, w, ones, zeros, nullptr, dx, 0, &plan);
  if (rc) return rc;
  return conv_plan_launch(&plan, static_cast<cudaStream_t>(stream));
}
