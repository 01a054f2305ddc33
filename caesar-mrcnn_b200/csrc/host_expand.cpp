// HOST side of the result path of detect(): the engine ships each image's detection masks over PCIe as pixel-major
// bits (unmold.cu: unmold_paint_bits_kernel, 16 bytes per pixel for 100 detections) and this routine expands them
// into the reference's contract, one dense [H, W, N] bool array per image (MaskRCNN.unmold_detections,
// mrcnn/model.py:2613-2619: `full_masks = np.stack(full_masks, axis=-1)`), N = that image's detection count.
// Multi-threaded (persistent pool, dynamic chunks); AVX-512BW path (mask register -> 64 bytes per instruction)
// selected at run time, portable 8-bits-per-lookup path otherwise.  No device work, no CUDA calls.
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "host_pool.h"
#include "mrcnn_b200.h"

void mrcnn_set_error(const char* fmt, ...);

namespace {

// ---- expansion kernels ------------------------------------------------------------------------
struct Lut {
  uint64_t v[256];
  Lut() {
    for (int b = 0; b < 256; ++b) {
      uint64_t x = 0;
      for (int k = 0; k < 8; ++k)
        if (b & (1 << k)) x |= (uint64_t)1 << (8 * k);
      v[b] = x;
    }
  }
};
const Lut g_lut;

// pixels [p0, p1) of one image: bits [npx][dw] words -> dst [npx][n] bytes
void expand_generic(const uint32_t* bits, int dw, int n, int64_t p0, int64_t p1, uint8_t* dst) {
  const int full = n >> 3, tail = n & 7;
  for (int64_t p = p0; p < p1; ++p) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(bits + p * dw);
    uint8_t* o = dst + p * n;
    for (int j = 0; j < full; ++j) memcpy(o + 8 * j, &g_lut.v[src[j]], 8);
    if (tail) memcpy(o + 8 * full, &g_lut.v[src[full]], tail);
  }
}

#if defined(__x86_64__)
// AVX-512BW: a 64-bit mask register becomes 64 bytes in one instruction.  The expanded bytes are first written to a small
// cache-resident staging buffer whose layout is congruent (mod 64) to the destination, and leave as full-line
// NON-TEMPORAL stores: the output (420 MB per 64-image batch) is write-only here, so the read-for-ownership traffic of
// ordinary stores (2x DRAM traffic) is what limits the multi-rank case.  Only the partial lines at the two ends of a
// chunk (shared with the neighbouring chunks, possibly written by another thread) use ordinary stores.
constexpr size_t kStage = 16384;

struct StageState {
  uint8_t* stage;        // 64-byte aligned staging buffer, congruent (mod 64) to the destination
  uint8_t* line;         // destination address of stage[0]
  size_t fill;           // bytes of the stage that are filled
  size_t first_valid;    // bytes of stage line 0 that belong to the previous chunk (not mine to write)
};

__attribute__((target("avx512f,avx512bw"))) void stage_flush(StageState& s, bool last) {
  const size_t lines = s.fill >> 6;
  size_t l = 0;
  if (lines && s.first_valid) {                          // shared first line: ordinary stores of my bytes only
    memcpy(s.line + s.first_valid, s.stage + s.first_valid, 64 - s.first_valid);
    s.first_valid = 0;
    l = 1;
  }
  for (; l < lines; ++l)
    _mm512_stream_si512(reinterpret_cast<__m512i*>(s.line + 64 * l), _mm512_load_si512(reinterpret_cast<const __m512i*>(s.stage + 64 * l)));
  const size_t rem = s.fill - 64 * lines;
  if (last) {
    if (rem > s.first_valid) memcpy(s.line + 64 * lines + s.first_valid, s.stage + 64 * lines + s.first_valid, rem - s.first_valid);
  } else if (lines) {
    memcpy(s.stage, s.stage + 64 * lines, 64);           // carry the partial line (plus harmless overrun bytes)
    s.line += 64 * lines;
    s.fill = rem;
  }
}

__attribute__((target("avx512f,avx512bw"))) void expand_avx512(const uint32_t* bits, int dw, int n, int64_t p0, int64_t p1,
                                                                uint8_t* dst) {
  alignas(64) uint8_t stage[kStage + 64 * 9];            // slack: a pixel's last 64-byte store may overrun its n bytes
  const __m512i one = _mm512_set1_epi8(1);
  const int vecs = (n + 63) >> 6;                        // 64-byte stores per pixel (<= 4 for n <= 256)
  const uintptr_t o = reinterpret_cast<uintptr_t>(dst + p0 * n);
  StageState st;
  st.stage = stage;
  st.first_valid = o & 63;
  st.line = reinterpret_cast<uint8_t*>(o - st.first_valid);
  st.fill = st.first_valid;
  for (int64_t p = p0; p < p1; ++p) {
    const uint32_t* src = bits + p * dw;
    uint8_t* w = stage + st.fill;
    for (int j = 0; j < vecs; ++j) {
      uint64_t k = src[2 * j];
      if (2 * j + 1 < dw) k |= (uint64_t)src[2 * j + 1] << 32;
      _mm512_storeu_si512(w + 64 * j, _mm512_maskz_mov_epi8((__mmask64)k, one));
    }
    st.fill += (size_t)n;
    if (st.fill >= kStage) stage_flush(st, false);
  }
  stage_flush(st, true);
  _mm_sfence();
}
#endif

typedef void (*ExpandFn)(const uint32_t*, int, int, int64_t, int64_t, uint8_t*);

ExpandFn pick_expand() {
#if defined(__x86_64__)
  const char* e = getenv("MRCNN_B200_HOST_SIMD");
  const bool allow = !(e && e[0] == '0');
  if (allow && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return expand_avx512;
#endif
  return expand_generic;
}

}  // namespace

extern "C" int mrcnn_host_threads(void) { return mrcnn_host::default_threads(); }

extern "C" int mrcnn_host_expand_mask_bits(const uint32_t* bits, int n_images, int64_t pixels_per_image, int words_per_pixel,
                                           const int32_t* counts, uint8_t* const* dst, int n_threads) {
  if (!bits || !counts || !dst || n_images < 0 || pixels_per_image < 0 || words_per_pixel < 1) {
    mrcnn_set_error("host_expand_mask_bits: bad arguments");
    return MRCNN_STATUS_INVALID;
  }
  for (int b = 0; b < n_images; ++b) {
    if (counts[b] < 0 || counts[b] > 32 * words_per_pixel || (counts[b] > 0 && !dst[b])) {
      mrcnn_set_error("host_expand_mask_bits: image %d: count %d outside [0,%d] or null destination", b, counts[b],
                      32 * words_per_pixel);
      return MRCNN_STATUS_INVALID;
    }
  }
  static const ExpandFn fn = pick_expand();
  const int64_t chunk = 2048;                                   // pixels per task (~200 KB of output at N = 100)
  const int64_t chunks_per_image = (pixels_per_image + chunk - 1) / chunk;
  const int64_t total = chunks_per_image * n_images;
  if (total == 0) return MRCNN_STATUS_OK;
  int nt = n_threads > 0 ? n_threads : mrcnn_host::default_threads();
  if ((int64_t)nt > total) nt = (int)total;
  std::atomic<int64_t> next(0);
  auto work = [&](int) {
    for (;;) {
      const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
      if (t >= total) break;
      const int b = (int)(t / chunks_per_image);
      const int n = counts[b];
      if (n == 0) continue;
      const int64_t p0 = (t % chunks_per_image) * chunk;
      const int64_t p1 = p0 + chunk < pixels_per_image ? p0 + chunk : pixels_per_image;
      fn(bits + (size_t)b * pixels_per_image * words_per_pixel, words_per_pixel, n, p0, p1, dst[b]);
    }
  };
  mrcnn_host::Pool::get().run(nt, work);
  return MRCNN_STATUS_OK;
}
