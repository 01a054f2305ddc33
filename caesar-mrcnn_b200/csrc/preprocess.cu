// Input preprocessing on the device (a1, a2 of the hot-path table):
//   zscale_params  : NaN -> nanmin fill value, astropy ZScaleInterval limits (1000 strided samples,
//                    sorted, <=5 k-sigma line-fit iterations with dilation), per-channel contrast
//                    -> (fill, vmin, range, zmax)              [mrcnn/utils.py:1090-1091, 1166-1172]
//   stretch_to_rgb8: clip((x-vmin)/range,0,1) / max -> round_half_even(255*v) -> uint8 RGB
//                                                               [mrcnn/utils.py:1182-1208]
//   resize_pad_mold: skimage<=0.15 bilinear resize (float64, cval 0, clip, uint8 truncation),
//                    centre zero-pad to S x S, minus MEAN_PIXEL  [mrcnn/utils.py:456-561, 957-978;
//                                                               mrcnn/model.py:2519-2556, 2964-2969]
// Float32/float64 op order follows oracle/host_ops.py exactly (no FMA contraction in this file).
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

constexpr int ZS_THREADS = 1024;
constexpr int ZS_NSAMPLES = 1000;

struct ZParams {
  const float* maps;
  int H, W;
  float contrast[3];
  float* params;  // [n,3,4]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum of up to 3 doubles (all threads get the result)
__device__ void block_sum3(double& a, double& b, double& c, double* red /*[32*3]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { red[warp * 3] = a; red[warp * 3 + 1] = b; red[warp * 3 + 2] = c; }
  __syncthreads();
  double sa = 0, sb = 0, sc = 0;
  for (int w = 0; w < nw; ++w) { sa += red[w * 3]; sb += red[w * 3 + 1]; sc += red[w * 3 + 2]; }
  a = sa; b = sb; c = sc;
}

// ZScaleInterval.__call__ after the limits, legacy numpy scalar casting (oracle: zscale_apply)
__device__ __forceinline__ float zscale_apply(float x, float vmin32, float rng32) {
  float z = __fsub_rn(x, vmin32);
  if (rng32 != 0.f) z = __fdiv_rn(z, rng32);
  return fminf(fmaxf(z, 0.f), 1.f);
}

__global__ void __launch_bounds__(ZS_THREADS, 1) zscale_params_kernel(ZParams p) {
  pdl_prologue();
  __shared__ float s_samp[1024];
  __shared__ unsigned char s_bad[1024];
  __shared__ unsigned char s_bad2[1024];
  __shared__ double s_red[32 * 3];
  __shared__ int s_wtot[32];
  __shared__ float s_fmin[32], s_fmax[32];
  const int img = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const size_t npx = (size_t)p.H * p.W;
  const float* x = p.maps + (size_t)img * npx;

  // ---- nanmin / max over non-NaN pixels ----------------------------------------------------
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = tid; i < npx; i += nt) {
    const float v = x[i];
    if (v == v) { mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { s_fmin[warp] = mn; s_fmax[warp] = mx; }
  __syncthreads();
  mn = s_fmin[0]; mx = s_fmax[0];
  for (int w = 1; w < nw; ++w) { mn = fminf(mn, s_fmin[w]); mx = fmaxf(mx, s_fmax[w]); }
  const float fill = mn;   // np.nanmin; NaN pixels become this value

  // ---- finite count -> stride -> strided sample of the first 1000 -----------------------------
  int cnt = 0;
  for (size_t i = tid; i < npx; i += nt) {
    float v = x[i];
    if (!(v == v)) v = fill;
    cnt += (fabsf(v) <= 3.402823466e38f) ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __syncthreads();
  if (lane == 0) s_wtot[warp] = cnt;
  __syncthreads();
  long long nfinite = 0;
  for (int w = 0; w < nw; ++w) nfinite += s_wtot[w];
  __syncthreads();
  const long long stride = max(1LL, (long long)((double)nfinite / (double)ZS_NSAMPLES));   // int(max(1.0, n/1000))
  const int ns = (int)min((long long)ZS_NSAMPLES, (nfinite + stride - 1) / stride);
  s_samp[tid] = INFINITY;
  __syncthreads();
  long long rank_base = 0;
  if (nfinite == (long long)npx) {
    // every pixel is finite after the NaN fill (the usual case): rank == index, sample k is pixel k*stride
    for (int k = tid; k < ns; k += nt) {
      float v = x[(size_t)k * (size_t)stride];
      if (!(v == v)) v = fill;
      s_samp[k] = v;
    }
    rank_base = (long long)ns * stride;          // skips the ranked scan below
  }
  for (size_t base = 0; base < npx && rank_base < (long long)ns * stride; base += nt) {
    const size_t i = base + tid;
    float v = 0.f;
    bool fin = false;
    if (i < npx) {
      v = x[i];
      if (!(v == v)) v = fill;
      fin = fabsf(v) <= 3.402823466e38f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, fin);
    const int pre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_wtot[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < nw; ++w) { const int t = s_wtot[w]; if (w < warp) off += t; tot += t; }
    __syncthreads();
    if (fin) {
      const long long rank = rank_base + off + pre;
      if (rank % stride == 0) {
        const long long k = rank / stride;
        if (k < ns) s_samp[k] = v;
      }
    }
    rank_base += tot;
  }
  __syncthreads();

  // ---- ascending bitonic sort of 1024 (padding = +inf) ---------------------------------------
  for (int k = 2; k <= 1024; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = tid, ixj = i ^ j;
      if (ixj > i) {
        const float a = s_samp[i], b = s_samp[ixj];
        const bool asc = (i & k) == 0;
        if (asc ? (a > b) : (a < b)) { s_samp[i] = b; s_samp[ixj] = a; }
      }
      __syncthreads();
    }
  }

  // ---- iterative line fit with 2.5-sigma rejection and dilation ------------------------------
  const int minpix = max(5, (int)(ns * 0.5));
  const int ngrow = max(1, (int)(ns * 0.01));
  const bool in = tid < ns;
  const double xi = (double)tid;
  const double yi = in ? (double)s_samp[tid] : 0.0;
  s_bad[tid] = 0;
  __syncthreads();
  int ngood = ns, last = ns + 1;
  double slope = 0.0;
  for (int it = 0; it < 5; ++it) {
    if (ngood >= last || ngood < minpix) break;
    const bool good = in && !s_bad[tid];
    double a = good ? 1.0 : 0.0, b = good ? xi : 0.0, c = good ? yi : 0.0;
    block_sum3(a, b, c, s_red);
    const double ng = a, xm = b / a, ym = c / a;
    double sxx = good ? (xi - xm) * (xi - xm) : 0.0, sxy = good ? (xi - xm) * (yi - ym) : 0.0, dummy = 0.0;
    block_sum3(sxx, sxy, dummy, s_red);
    slope = sxy / sxx;
    const double icpt = ym - slope * xm;
    const double flat = yi - (slope * xi + icpt);
    double f1 = good ? flat : 0.0, d2 = 0.0, d3 = 0.0;
    block_sum3(f1, d2, d3, s_red);
    const double fmean = f1 / ng;
    double v2 = good ? (flat - fmean) * (flat - fmean) : 0.0;
    d2 = 0.0; d3 = 0.0;
    block_sum3(v2, d2, d3, s_red);
    const double thr = 2.5 * sqrt(v2 / ng);
    if (in && (flat < -thr || flat > thr)) s_bad[tid] = 1;
    __syncthreads();
    // np.convolve(bad, ones(ngrow), 'same'): new[i] = any(bad[i - ngrow/2 .. i + (ngrow-1)/2])
    unsigned char nb = 0;
    if (in) {
      const int lo = max(0, tid - ngrow / 2), hi = min(ns - 1, tid + (ngrow - 1) / 2);
      for (int k = lo; k <= hi; ++k) nb |= s_bad[k];
    }
    s_bad2[tid] = nb;
    __syncthreads();
    s_bad[tid] = s_bad2[tid];
    last = ngood;
    ngood = __syncthreads_count(in && !nb);
  }

  if (tid == 0) {
    const float s0 = s_samp[0], s1 = s_samp[ns - 1];
    float med32;
    if (ns & 1) med32 = s_samp[ns / 2];
    else med32 = __fdiv_rn(__fadd_rn(s_samp[ns / 2 - 1], s_samp[ns / 2]), 2.0f);   // np.median on float32
    const int center = (ns - 1) / 2;
    for (int c = 0; c < 3; ++c) {
      double vmin = (double)s0, vmax = (double)s1;
      bool vmin_f32 = true, vmax_f32 = true;
      if (ngood >= minpix) {
        double sl = slope;
        if (p.contrast[c] > 0.f) sl = sl / (double)p.contrast[c];
        const double lo = (double)med32 - (double)(center - 1) * sl;
        const double hi = (double)med32 + (double)(ns - center) * sl;
        if (lo > vmin) { vmin = lo; vmin_f32 = false; }     // python max(vmin, lo)
        if (hi < vmax) { vmax = hi; vmax_f32 = false; }     // python min(vmax, hi)
      }
      const float vmin32 = (float)vmin;
      float rng32;
      if (vmin_f32 && vmax_f32) rng32 = __fsub_rn(s1, s0);  // float32 - float32 stays float32
      else rng32 = (float)(vmax - vmin);
      if (vmax - vmin == 0.0) rng32 = 0.f;
      float* o = p.params + ((size_t)img * 3 + c) * 4;
      o[0] = fill;
      o[1] = vmin32;
      o[2] = rng32;
      o[3] = zscale_apply(mx, vmin32, rng32);                // np.max of the stretched image (monotone)
    }
  }
}

__global__ void init_minmax_kernel(int32_t* minmax, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { minmax[2 * i] = 255; minmax[2 * i + 1] = 0; }
}

__device__ __forceinline__ int stretch_one(float v, float vmin, float rng, float zmax) {
  const float z = zscale_apply(v, vmin, rng);
  const float zn = __fdiv_rn(z, zmax);                   // normalize_img: data / data.max()
  const float q = rintf(__fmul_rn(zn, 255.0f));          // (x*255).round(), half to even
  return (int)q;
}

// 4 pixels per thread (one 16-byte load, three 4-byte stores) when the plane size allows it; identical channel
// parameters (equal zscale contrasts, the run.py default) are evaluated once and replicated.
__global__ void stretch_rgb8_kernel(const float* __restrict__ maps, const float* __restrict__ params, int H, int W,
                                    uint8_t* __restrict__ rgb, int32_t* __restrict__ minmax) {
  pdl_prologue();
  const int img = blockIdx.y;
  const size_t npx = (size_t)H * W;
  const float* pr = params + (size_t)img * 12;
  const float fill = pr[0];
  float vmin[3], rng[3], zmax[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { vmin[c] = pr[c * 4 + 1]; rng[c] = pr[c * 4 + 2]; zmax[c] = pr[c * 4 + 3]; }
  const bool same = vmin[0] == vmin[1] && vmin[0] == vmin[2] && rng[0] == rng[1] && rng[0] == rng[2] && zmax[0] == zmax[1] &&
                    zmax[0] == zmax[2];
  int lo = 255, hi = 0;
  const float* src = maps + (size_t)img * npx;
  uint8_t* dst = rgb + (size_t)img * npx * 3;
  if ((npx & 3) == 0) {
    for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < npx / 4; i4 += (size_t)gridDim.x * blockDim.x) {
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(src) + i4);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      uint32_t bytes[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = vv[k];
        if (!(v == v)) v = fill;
        int u[3];
        u[0] = stretch_one(v, vmin[0], rng[0], zmax[0]);
        if (same) {
          u[1] = u[0];
          u[2] = u[0];
        } else {
          u[1] = stretch_one(v, vmin[1], rng[1], zmax[1]);
          u[2] = stretch_one(v, vmin[2], rng[2], zmax[2]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          bytes[k * 3 + c] = (uint32_t)(u[c] & 255);
          lo = min(lo, u[c] & 255);
          hi = max(hi, u[c] & 255);
        }
      }
      uint32_t* o = reinterpret_cast<uint32_t*>(dst) + i4 * 3;     // 12 bytes per 4 pixels, 4-byte aligned
#pragma unroll
      for (int w = 0; w < 3; ++w)
        o[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | (bytes[4 * w + 3] << 24);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
      float v = src[i];
      if (!(v == v)) v = fill;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int u = stretch_one(v, vmin[c], rng[c], zmax[c]);
        dst[i * 3 + c] = (uint8_t)u;
        lo = min(lo, u & 255);
        hi = max(hi, u & 255);
      }
    }
  }
  // one pair of atomics per CTA (warp reduce -> shared -> thread 0): thousands of warps hammering the same two
  // words per image used to dominate this kernel
  __shared__ int s_lo[32], s_hi[32];
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = min(lo, s_lo[w]);
      hi = max(hi, s_hi[w]);
    }
    atomicMin(&minmax[2 * img], lo);
    atomicMax(&minmax[2 * img + 1], hi);
  }
}

struct MoldParams {
  const uint8_t* rgb;
  const int32_t* minmax;
  int H, W, out_h, out_w, SH, SW, top, left;
  float mean[3];
  float* molded;
};

__device__ __forceinline__ double px_or_zero(const uint8_t* img, int H, int W, int r, int c, int ch) {
  if (r < 0 || r >= H || c < 0 || c >= W) return 0.0;
  return (double)img[((size_t)r * W + c) * 3 + ch];
}

__global__ void resize_pad_mold_kernel(MoldParams p) {
  pdl_prologue();
  const int img = blockIdx.y;
  const size_t total = (size_t)p.SH * p.SW;
  const uint8_t* src = p.rgb + (size_t)img * p.H * p.W * 3;
  const bool resize = !(p.out_h == p.H && p.out_w == p.W);
  const double row_scale = (double)p.H / (double)p.out_h;
  const double col_scale = (double)p.W / (double)p.out_w;
  const double row_off = __dsub_rn(__dmul_rn(0.5, row_scale), 0.5);
  const double col_off = __dsub_rn(__dmul_rn(0.5, col_scale), 0.5);
  const double mn = (double)p.minmax[2 * img], mx = (double)p.minmax[2 * img + 1];
  const bool preserve_cval = !(mn <= 0.0 && 0.0 <= mx);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % p.SW), oy = (int)(i / p.SW);
    const int y = oy - p.top, x = ox - p.left;
    float out[3] = {0.f, 0.f, 0.f};
    if (y >= 0 && y < p.out_h && x >= 0 && x < p.out_w) {
      if (!resize) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = (float)src[((size_t)y * p.W + x) * 3 + c];
      } else {
        const double r = __dadd_rn(__dmul_rn(row_scale, (double)y), row_off);
        const double cc = __dadd_rn(__dmul_rn(col_scale, (double)x), col_off);
        const double fr = floor(r), fc = floor(cc);
        const int r0 = (int)fr, r1 = (int)ceil(r), c0 = (int)fc, c1 = (int)ceil(cc);
        const double dr = __dsub_rn(r, fr), dc = __dsub_rn(cc, fc);
        const double wr = __dsub_rn(1.0, dr), wc = __dsub_rn(1.0, dc);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double t = __dadd_rn(__dmul_rn(wc, px_or_zero(src, p.H, p.W, r0, c0, c)),
                                     __dmul_rn(dc, px_or_zero(src, p.H, p.W, r0, c1, c)));
          const double b = __dadd_rn(__dmul_rn(wc, px_or_zero(src, p.H, p.W, r1, c0, c)),
                                     __dmul_rn(dc, px_or_zero(src, p.H, p.W, r1, c1, c)));
          double v = __dadd_rn(__dmul_rn(wr, t), __dmul_rn(dr, b));
          if (!(preserve_cval && v == 0.0)) v = fmin(fmax(v, mn), mx);
          out[c] = (float)(int)v;                          // astype(uint8): truncation (0..255)
        }
      }
    }
    float* o = p.molded + ((size_t)img * total + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (float)((double)out[c] - (double)p.mean[c]);
  }
}


// skimage.transform.resize(order=1, mode='constant', cval=0, clip=True, anti_aliasing=False) for a float64 [H,W,C]
// image (scikit-image <= 0.15 affine warp; same arithmetic, op for op, as resize_pad_mold_kernel / unmold_paint_kernel)
struct ResizeParams {
  const double* src;
  double* dst;
  int H, W, C, out_h, out_w;
  double mn, mx;
};

__global__ void skimage_resize_kernel(ResizeParams p) {
  pdl_prologue();
  const size_t total = (size_t)p.out_h * p.out_w * p.C;
  const double row_scale = (double)p.H / (double)p.out_h;
  const double col_scale = (double)p.W / (double)p.out_w;
  const bool single = (p.out_h == 1 && p.out_w == 1);        // skimage uses a translation-only transform for a 1x1 output
  const double row_off = __dsub_rn(__dmul_rn(0.5, row_scale), 0.5);
  const double col_off = __dsub_rn(__dmul_rn(0.5, col_scale), 0.5);
  const bool preserve_cval = !(p.mn <= 0.0 && 0.0 <= p.mx);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % p.C);
    const size_t px = i / p.C;
    const int x = (int)(px % p.out_w), y = (int)(px / p.out_w);
    const double r = single ? __dsub_rn((double)p.H / 2.0, 0.5) : __dadd_rn(__dmul_rn(row_scale, (double)y), row_off);
    const double cc = single ? __dsub_rn((double)p.W / 2.0, 0.5) : __dadd_rn(__dmul_rn(col_scale, (double)x), col_off);
    const double fr = floor(r), fc = floor(cc);
    const int r0 = (int)fr, r1 = (int)ceil(r), c0 = (int)fc, c1 = (int)ceil(cc);
    const double dr = __dsub_rn(r, fr), dc = __dsub_rn(cc, fc);
    const double wr = __dsub_rn(1.0, dr), wc = __dsub_rn(1.0, dc);
    auto at = [&](int ri, int ci) -> double {
      if (ri < 0 || ri >= p.H || ci < 0 || ci >= p.W) return 0.0;
      return p.src[((size_t)ri * p.W + ci) * p.C + ch];
    };
    const double t = __dadd_rn(__dmul_rn(wc, at(r0, c0)), __dmul_rn(dc, at(r0, c1)));
    const double b = __dadd_rn(__dmul_rn(wc, at(r1, c0)), __dmul_rn(dc, at(r1, c1)));
    double v = __dadd_rn(__dmul_rn(wr, t), __dmul_rn(dr, b));
    if (!(preserve_cval && v == 0.0)) v = fmin(fmax(v, p.mn), p.mx);
    p.dst[i] = v;
  }
}

}  // namespace

extern "C" int mrcnn_zscale_params(const float* maps, int n_images, int height, int width, const float* contrasts3,
                                   float* params, void* stream) {
  MRCNN_REQUIRE(maps && params && contrasts3, "zscale_params: null pointer");
  MRCNN_REQUIRE(n_images > 0 && height > 0 && width > 0, "zscale_params: empty input");
  MRCNN_REQUIRE((long long)height * width >= 5, "zscale_params: image too small");
  ZParams p;
  p.maps = maps; p.H = height; p.W = width; p.params = params;
  for (int i = 0; i < 3; ++i) p.contrast[i] = contrasts3[i];
  MRCNN_CHECK_CUDA(mrcnn_launch(zscale_params_kernel, dim3(n_images), dim3(ZS_THREADS), 0, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_stretch_to_rgb8(const float* maps, const float* params, int n_images, int height, int width,
                                     uint8_t* rgb, int32_t* minmax, void* stream) {
  MRCNN_REQUIRE(maps && params && rgb && minmax, "stretch_to_rgb8: null pointer");
  MRCNN_REQUIRE(n_images > 0 && height > 0 && width > 0, "stretch_to_rgb8: empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MRCNN_CHECK_CUDA(mrcnn_launch(init_minmax_kernel, dim3(ceil_div(n_images, 128)), dim3(128), 0, st, minmax, n_images));
  const size_t npx = (size_t)height * width;
  int bx = (int)((npx / 4 + 255) / 256);          // 4 pixels per thread
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;                           // grid-stride beyond that: few CTAs per image, few atomics
  MRCNN_CHECK_CUDA(mrcnn_launch(stretch_rgb8_kernel, dim3(dim3(bx, n_images)), dim3(256), 0, st, maps, params, height, width, rgb, minmax));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(2);
  return MRCNN_OK;
}

extern "C" int mrcnn_resize_pad_mold(const uint8_t* rgb, const int32_t* minmax, int n_images, int height, int width,
                                     int out_h, int out_w, int square, int top, int left, const float* mean_pixel3,
                                     float* molded, void* stream) {
  MRCNN_REQUIRE(rgb && minmax && mean_pixel3 && molded, "resize_pad_mold: null pointer");
  MRCNN_REQUIRE(n_images > 0 && height > 0 && width > 0 && out_h > 0 && out_w > 0 && square > 0, "resize_pad_mold: empty input");
  MRCNN_REQUIRE(top >= 0 && left >= 0 && top + out_h <= square && left + out_w <= square,
                "resize_pad_mold: window (%d,%d)+(%d,%d) exceeds the %d frame", top, left, out_h, out_w, square);
  MoldParams p;
  p.rgb = rgb; p.minmax = minmax; p.H = height; p.W = width; p.out_h = out_h; p.out_w = out_w;
  p.SH = square; p.SW = square; p.top = top; p.left = left; p.molded = molded;
  for (int i = 0; i < 3; ++i) p.mean[i] = mean_pixel3[i];
  const size_t total = (size_t)square * square;
  int bx = (int)((total + 255) / 256);
  if (bx > 148 * 4) bx = 148 * 4;
  MRCNN_CHECK_CUDA(mrcnn_launch(resize_pad_mold_kernel, dim3(dim3(bx, n_images)), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_skimage_resize_f64(const double* image, int height, int width, int channels, int out_h, int out_w,
                                        double image_min, double image_max, double* out, void* stream) {
  MRCNN_REQUIRE(image && out, "skimage_resize: null pointer");
  MRCNN_REQUIRE(height > 0 && width > 0 && channels > 0 && out_h > 0 && out_w > 0, "skimage_resize: empty input / output");
  ResizeParams p;
  p.src = image; p.dst = out; p.H = height; p.W = width; p.C = channels; p.out_h = out_h; p.out_w = out_w;
  p.mn = image_min; p.mx = image_max;
  const size_t total = (size_t)out_h * out_w * channels;
  int bx = (int)((total + 255) / 256);
  if (bx > 148 * 8) bx = 148 * 8;
  MRCNN_CHECK_CUDA(mrcnn_launch(skimage_resize_kernel, dim3(bx), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
