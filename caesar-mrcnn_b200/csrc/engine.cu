// mrcnn_engine — keras_model.predict([molded_images, image_metas, anchors]) of mode='inference'
// (graph built at mrcnn/model.py:1935-2054, 2133-2159) as a static plan of kernel launches on one
// CUDA stream, plus by-name weight loading (MaskRCNN.load_weights, mrcnn/model.py:2197-2239) and
// unmold (mrcnn/model.py:2558-2621).  Layer names / kernel layouts follow SURVEY.md Appendix B.
//
// HBM layout: activations are NHWC bf16, one allocation per tensor (a 64-image batch at S=256
// needs ~12 GB of the 180 GB); weights are bf16 [Cout, KH*KW*Cin] (K-major GEMM-B operands) with
// the frozen BatchNorm folded into per-channel fp32 (scale, shift) applied in the GEMM epilogue.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

int launch_pyramid_roi_align(const void* const* feature_maps, const int* feat_h, const int* feat_w, int channels,
                             int dtype, const float* boxes, int box_stride, int batch, int num_boxes, int pool_size,
                             float image_area, void* pooled, int32_t* levels, cudaStream_t st, int valid_col = -1);

namespace {

enum { KIND_CONV = 0, KIND_BN = 1, KIND_DENSE = 2, KIND_DECONV = 3 };
enum { DT_F32 = 0, DT_BF16 = 1, DT_I32 = 2, DT_U8 = 3 };

struct LayerSpec {
  std::string name;
  int kind;
  int shape[4];
  int nweights;
  std::vector<std::vector<float>> host;  // per weight_index
  std::vector<bool> set;
  size_t count(int wi) const {
    if (kind == KIND_BN) return (size_t)shape[0];
    if (wi == 0) {
      size_t c = 1;
      for (int i = 0; i < 4; ++i) if (shape[i] > 0) c *= (size_t)shape[i];
      return c;
    }
    // bias length
    if (kind == KIND_DENSE) return (size_t)shape[1];
    if (kind == KIND_DECONV) return (size_t)shape[2];
    return (size_t)shape[3];
  }
};

struct Tensor {
  void* ptr = nullptr;
  size_t bytes = 0;
  int dtype = DT_F32;
  size_t elems = 0;
};

struct GemmW {   // device-side parameters of one GEMM layer
  __nv_bfloat16* w = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  int cout = 0, K = 0;
};

struct Step {
  std::string stage;
  std::function<int(cudaStream_t)> run;
  std::string kind;   // kernel family, for per-kernel timing ("conv_gemm", "roialign", ...)
  std::string label;  // output tensor / layer name
  double flops = 0;
};

uint16_t f2bf(float f) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

struct mrcnn_engine {
  mrcnn_engine_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<LayerSpec> layers;
  std::map<std::string, int> layer_index;
  std::map<std::string, Tensor> tensors;
  std::map<std::string, GemmW> gemm;
  std::vector<void*> allocs;
  std::vector<Step> steps;
  std::vector<ConvPlan*> plans;
  bool finalized = false;
  int num_anchors = 0;
  int feat[5] = {0, 0, 0, 0, 0};
  double flops = 0;
  // timing
  std::vector<std::string> stage_names;
  std::vector<cudaEvent_t> stage_events;  // stage_names.size() + 1
  bool use_graph = false;                 // replay the plan as a CUDA graph when not profiling (default: batch <= 16)
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_fresh = false;
  unsigned long long launches_per_predict = 0;
  bool autotune = true;
  std::map<std::string, std::pair<int, int>> tune_cache;   // layer key -> (BLOCK_N, epi_tma | occ2 << 1)
  bool tune_cache_dirty = false;
  std::string tune_cache_path;
  int profiling = 0;                      // 0 off, 1 events around every launch, 2 events at kernel-family boundaries
  std::vector<cudaEvent_t> step_events;   // steps.size() + 1, recorded when profiling == 1
  // profiling == 2: a "run" is a maximal sequence of consecutive launches of one family; kRunSets event sets are
  // cycled so the last kRunSets predicts can be read back after a timed loop without a sync inside it
  static constexpr int kRunSets = 8;
  std::vector<int> run_first;             // first step index of each run
  std::vector<cudaEvent_t> run_events[kRunSets];   // runs + 1 per set
  unsigned long long run_calls = 0;
  std::vector<std::string> kind_names;    // scratch for kernel_times()
  // unmold scratch
  void* unmold_ws = nullptr;
  size_t unmold_ws_bytes = 0;
  int32_t* d_windows = nullptr;
  // result slots: unmold outputs are double-buffered so the D2H copy of step k (copy stream) can
  // overlap the compute of step k+1 (main stream) in the asynchronous detect calls
  struct ResultSlot {
    void* masks = nullptr; size_t masks_bytes = 0;     // [B,H0,W0,D] uint8 (mask_format 0: device consumers, mrcnn.analyze)
    void* bits = nullptr; size_t bits_bytes = 0;       // [B,H0*W0,DW] uint32 pixel-major bits (mask_format 1: host results)
    int32_t* rois = nullptr; int32_t* class_ids = nullptr; float* scores = nullptr; int32_t* counts = nullptr;
    cudaEvent_t computed = nullptr, copied = nullptr;
    bool copy_pending = false;
    // dense share of the host results (mrcnn_engine_set_dense_output): masks of the first images expanded on the device
    void* dense = nullptr; size_t dense_bytes = 0;
    cudaEvent_t dense_t0 = nullptr, dense_t1 = nullptr;   // dense_t0 doubles as "everything but the dense share is on the host"
    int dense_images = 0;
  } slots[2];
  int cur_slot = 0;
  uint8_t* dense_host = nullptr;   // pending request for the next host-result fetch
  int dense_n = 0;
  cudaStream_t copy_stream = nullptr;
  // preprocessing scratch (detect_maps).  Host maps of asynchronous calls are uploaded on their own stream into one
  // of two buffers, so the H2D copy of step k+1 overlaps the compute of step k
  void* pre_maps = nullptr; size_t pre_maps_bytes = 0;
  void* up_maps[2] = {nullptr, nullptr}; size_t up_maps_bytes[2] = {0, 0};
  cudaEvent_t up_done[2] = {nullptr, nullptr}, up_consumed[2] = {nullptr, nullptr};
  bool up_used[2] = {false, false};
  int up_next = 0;
  cudaStream_t upload_stream = nullptr;
  void* pre_rgb = nullptr; size_t pre_rgb_bytes = 0;
  void* pre_small = nullptr; size_t pre_small_bytes = 0;

  int meta_size() const { return 12 + cfg.num_classes; }
};

namespace {

void add_layer(mrcnn_engine* e, const std::string& name, int kind, int a, int b = 0, int c = 0, int d = 0) {
  LayerSpec l;
  l.name = name;
  l.kind = kind;
  l.shape[0] = a; l.shape[1] = b; l.shape[2] = c; l.shape[3] = d;
  l.nweights = kind == KIND_BN ? 4 : 2;
  l.host.resize(l.nweights);
  l.set.assign(l.nweights, false);
  e->layer_index[name] = (int)e->layers.size();
  e->layers.push_back(l);
}

void build_layer_table(mrcnn_engine* e) {
  const int NC = e->cfg.num_classes, FC = e->cfg.fc_layers_size, PY = e->cfg.top_down_pyramid_size;
  const int apl = e->cfg.anchors_per_location;
  add_layer(e, "conv1", KIND_CONV, 7, 7, 3, 64);
  add_layer(e, "bn_conv1", KIND_BN, 64);
  int cin = 64;
  const int nblocks[4] = {3, 4, 23, 3};
  const int f[4][3] = {{64, 64, 256}, {128, 128, 512}, {256, 256, 1024}, {512, 512, 2048}};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < nblocks[s]; ++b) {
      const std::string blk(1, (char)('a' + b));
      const std::string base = "res" + std::to_string(s + 2) + blk + "_branch";
      const std::string bnb = "bn" + std::to_string(s + 2) + blk + "_branch";
      add_layer(e, base + "2a", KIND_CONV, 1, 1, cin, f[s][0]);
      add_layer(e, bnb + "2a", KIND_BN, f[s][0]);
      add_layer(e, base + "2b", KIND_CONV, 3, 3, f[s][0], f[s][1]);
      add_layer(e, bnb + "2b", KIND_BN, f[s][1]);
      add_layer(e, base + "2c", KIND_CONV, 1, 1, f[s][1], f[s][2]);
      add_layer(e, bnb + "2c", KIND_BN, f[s][2]);
      if (b == 0) {
        add_layer(e, base + "1", KIND_CONV, 1, 1, cin, f[s][2]);
        add_layer(e, bnb + "1", KIND_BN, f[s][2]);
      }
      cin = f[s][2];
    }
  }
  add_layer(e, "fpn_c5p5", KIND_CONV, 1, 1, 2048, PY);
  add_layer(e, "fpn_c4p4", KIND_CONV, 1, 1, 1024, PY);
  add_layer(e, "fpn_c3p3", KIND_CONV, 1, 1, 512, PY);
  add_layer(e, "fpn_c2p2", KIND_CONV, 1, 1, 256, PY);
  for (int l = 2; l <= 5; ++l) add_layer(e, "fpn_p" + std::to_string(l), KIND_CONV, 3, 3, PY, PY);
  add_layer(e, "rpn_conv_shared", KIND_CONV, 3, 3, PY, 512);
  add_layer(e, "rpn_class_raw", KIND_CONV, 1, 1, 512, 2 * apl);
  add_layer(e, "rpn_bbox_pred", KIND_CONV, 1, 1, 512, 4 * apl);
  add_layer(e, "mrcnn_class_conv1", KIND_CONV, e->cfg.pool_size, e->cfg.pool_size, PY, FC);
  add_layer(e, "mrcnn_class_bn1", KIND_BN, FC);
  add_layer(e, "mrcnn_class_conv2", KIND_CONV, 1, 1, FC, FC);
  add_layer(e, "mrcnn_class_bn2", KIND_BN, FC);
  add_layer(e, "mrcnn_class_logits", KIND_DENSE, FC, NC);
  add_layer(e, "mrcnn_bbox_fc", KIND_DENSE, FC, 4 * NC);
  for (int i = 1; i <= 4; ++i) {
    add_layer(e, "mrcnn_mask_conv" + std::to_string(i), KIND_CONV, 3, 3, PY, PY);
    add_layer(e, "mrcnn_mask_bn" + std::to_string(i), KIND_BN, PY);
  }
  add_layer(e, "mrcnn_mask_deconv", KIND_DECONV, 2, 2, PY, PY);
  add_layer(e, "mrcnn_mask", KIND_CONV, 1, 1, PY, NC);
}

int dev_alloc(mrcnn_engine* e, size_t bytes, void** out) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  MRCNN_CHECK_CUDA(cudaMalloc(&p, bytes));
  e->allocs.push_back(p);
  *out = p;
  return MRCNN_OK;
}

int new_tensor(mrcnn_engine* e, const std::string& name, int dtype, size_t elems, Tensor* out, bool zero = false) {
  Tensor t;
  t.dtype = dtype;
  t.elems = elems;
  const size_t es = dtype == DT_BF16 ? 2 : (dtype == DT_U8 ? 1 : 4);
  t.bytes = elems * es;
  int rc = dev_alloc(e, t.bytes, &t.ptr);
  if (rc) return rc;
  if (zero) MRCNN_CHECK_CUDA(cudaMemset(t.ptr, 0, t.bytes));
  e->tensors[name] = t;
  if (out) *out = t;
  return MRCNN_OK;
}

const LayerSpec* find_layer(const mrcnn_engine* e, const std::string& name) {
  auto it = e->layer_index.find(name);
  return it == e->layer_index.end() ? nullptr : &e->layers[it->second];
}

// (scale, shift) with y = acc*scale + shift: folded Keras BatchNormalization (eps 1e-3) + conv bias
void fold_affine(const LayerSpec* conv, const LayerSpec* bn, int cout, std::vector<float>* scale, std::vector<float>* shift) {
  scale->assign(cout, 1.0f);
  shift->assign(cout, 0.0f);
  const std::vector<float>* bias = (conv && conv->set[1]) ? &conv->host[1] : nullptr;
  for (int c = 0; c < cout; ++c) {
    const double b = bias ? (double)(*bias)[c] : 0.0;
    if (bn && bn->set[0] && bn->set[1] && bn->set[2] && bn->set[3]) {
      const double s = (double)bn->host[0][c] / sqrt((double)bn->host[3][c] + 1e-3);
      (*scale)[c] = (float)s;
      (*shift)[c] = (float)((b - (double)bn->host[2][c]) * s + (double)bn->host[1][c]);
    } else {
      (*shift)[c] = (float)b;
    }
  }
}

int upload_gemm(mrcnn_engine* e, const std::string& key, const std::vector<uint16_t>& w, int cout, int K,
                const std::vector<float>& scale, const std::vector<float>& shift) {
  GemmW g;
  g.cout = cout;
  g.K = K;
  int rc;
  // rows padded to a multiple of 128 so a B tile never reads past the allocation even without OOB fill
  const size_t rows = (size_t)((cout + 127) / 128) * 128;
  if ((rc = dev_alloc(e, rows * K * 2, (void**)&g.w))) return rc;
  MRCNN_CHECK_CUDA(cudaMemset(g.w, 0, rows * K * 2));
  MRCNN_CHECK_CUDA(cudaMemcpy(g.w, w.data(), (size_t)cout * K * 2, cudaMemcpyHostToDevice));
  const size_t cpad = rows;
  if ((rc = dev_alloc(e, cpad * 4, (void**)&g.scale))) return rc;
  if ((rc = dev_alloc(e, cpad * 4, (void**)&g.shift))) return rc;
  MRCNN_CHECK_CUDA(cudaMemset(g.scale, 0, cpad * 4));
  MRCNN_CHECK_CUDA(cudaMemset(g.shift, 0, cpad * 4));
  MRCNN_CHECK_CUDA(cudaMemcpy(g.scale, scale.data(), (size_t)cout * 4, cudaMemcpyHostToDevice));
  MRCNN_CHECK_CUDA(cudaMemcpy(g.shift, shift.data(), (size_t)cout * 4, cudaMemcpyHostToDevice));
  e->gemm[key] = g;
  return MRCNN_OK;
}

// Keras Conv2D kernel [kh,kw,cin,cout] -> [cout][(r*kw+s)*cin + c] (+ zero K padding to kpad)
void conv_to_gemm(const std::vector<float>& k, int kh, int kw, int cin, int cout, int kpad, std::vector<uint16_t>* out,
                  int row0 = 0, int rows_total = -1) {
  const int K = kh * kw * cin;
  if (rows_total < 0) rows_total = cout;
  if (out->empty()) out->assign((size_t)rows_total * kpad, 0);
  for (int t = 0; t < kh * kw; ++t)
    for (int c = 0; c < cin; ++c) {
      const float* src = k.empty() ? nullptr : &k[((size_t)t * cin + c) * cout];
      for (int o = 0; o < cout; ++o)
        (*out)[(size_t)(row0 + o) * kpad + (size_t)t * cin + c] = src ? f2bf(src[o]) : 0;
    }
  (void)K;
}

int build_weights(mrcnn_engine* e) {
  int rc;
  auto L = [&](const std::string& n) { return find_layer(e, n); };
  auto kernel = [&](const std::string& n) -> const std::vector<float>& { return L(n)->host[0]; };
  // plain conv (+ optional BN) layers
  for (const LayerSpec& l : e->layers) {
    if (l.kind != KIND_CONV) continue;
    if (l.name == "rpn_class_raw" || l.name == "rpn_bbox_pred") continue;
    const int kh = l.shape[0], kw = l.shape[1], cin = l.shape[2], cout = l.shape[3];
    int kpad = kh * kw * cin;
    if (l.name == "conv1") kpad = 192;
    std::vector<uint16_t> w;
    if (l.name == "conv1") {
      // stem im2col K order (elementwise.cu): k = r*24 + s*3 + c, 3 zero entries per filter row, zero tail
      w.assign((size_t)cout * kpad, 0);
      if (!l.host[0].empty())
        for (int r = 0; r < kh; ++r)
          for (int sx = 0; sx < kw; ++sx)
            for (int c = 0; c < cin; ++c)
              for (int o = 0; o < cout; ++o)
                w[(size_t)o * kpad + r * 24 + sx * cin + c] = f2bf(l.host[0][(((size_t)r * kw + sx) * cin + c) * cout + o]);
    } else {
      conv_to_gemm(l.host[0], kh, kw, cin, cout, kpad, &w);
    }
    // BN partner by naming convention
    std::string bn;
    if (l.name == "conv1") bn = "bn_conv1";
    else if (l.name.rfind("res", 0) == 0) bn = "bn" + l.name.substr(3);
    else if (l.name.rfind("mrcnn_class_conv", 0) == 0) bn = "mrcnn_class_bn" + l.name.substr(16);
    else if (l.name.rfind("mrcnn_mask_conv", 0) == 0) bn = "mrcnn_mask_bn" + l.name.substr(15);
    std::vector<float> sc, sh;
    fold_affine(&l, bn.empty() ? nullptr : L(bn), cout, &sc, &sh);
    if ((rc = upload_gemm(e, l.name, w, cout, kpad, sc, sh))) return rc;
  }
  {  // RPN head: [class_raw(2*apl) ; bbox_pred(4*apl)] x 512
    const LayerSpec* a = L("rpn_class_raw");
    const LayerSpec* b = L("rpn_bbox_pred");
    const int ca = a->shape[3], cb = b->shape[3], cin = a->shape[2];
    std::vector<uint16_t> w;
    conv_to_gemm(a->host[0], 1, 1, cin, ca, cin, &w, 0, ca + cb);
    conv_to_gemm(b->host[0], 1, 1, cin, cb, cin, &w, ca, ca + cb);
    std::vector<float> sc(ca + cb, 1.0f), sh(ca + cb, 0.0f);
    for (int i = 0; i < ca; ++i) sh[i] = a->set[1] ? a->host[1][i] : 0.f;
    for (int i = 0; i < cb; ++i) sh[ca + i] = b->set[1] ? b->host[1][i] : 0.f;
    if ((rc = upload_gemm(e, "rpn_head", w, ca + cb, cin, sc, sh))) return rc;
  }
  {  // class head: Dense [in,out] kernels -> [logits(NC) ; bbox(4NC)] x FC
    const LayerSpec* a = L("mrcnn_class_logits");
    const LayerSpec* b = L("mrcnn_bbox_fc");
    const int in = a->shape[0], ca = a->shape[1], cb = b->shape[1];
    std::vector<uint16_t> w((size_t)(ca + cb) * in, 0);
    for (int i = 0; i < in; ++i) {
      for (int o = 0; o < ca; ++o) w[(size_t)o * in + i] = a->host[0].empty() ? 0 : f2bf(a->host[0][(size_t)i * ca + o]);
      for (int o = 0; o < cb; ++o) w[(size_t)(ca + o) * in + i] = b->host[0].empty() ? 0 : f2bf(b->host[0][(size_t)i * cb + o]);
    }
    std::vector<float> sc(ca + cb, 1.0f), sh(ca + cb, 0.0f);
    for (int i = 0; i < ca; ++i) sh[i] = a->set[1] ? a->host[1][i] : 0.f;
    for (int i = 0; i < cb; ++i) sh[ca + i] = b->set[1] ? b->host[1][i] : 0.f;
    if ((rc = upload_gemm(e, "class_head", w, ca + cb, in, sc, sh))) return rc;
  }
  {  // Conv2DTranspose kernel [2,2,cout,cin] -> [(i*2+j)*cout + o][c]
    const LayerSpec* d = L("mrcnn_mask_deconv");
    const int cout = d->shape[2], cin = d->shape[3];
    std::vector<uint16_t> w((size_t)4 * cout * cin, 0);
    if (!d->host[0].empty())
      for (int t = 0; t < 4; ++t)
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c) w[((size_t)t * cout + o) * cin + c] = f2bf(d->host[0][((size_t)t * cout + o) * cin + c]);
    std::vector<float> sc((size_t)4 * cout, 1.0f), sh((size_t)4 * cout, 0.0f);
    for (int t = 0; t < 4; ++t)
      for (int i = 0; i < cout; ++i) sh[(size_t)t * cout + i] = d->set[1] ? d->host[1][i] : 0.f;
    // the kernel indexes scale/shift per output channel (first `cout` entries; same for the 4 taps)
    if ((rc = upload_gemm(e, "mrcnn_mask_deconv", w, 4 * cout, cin, sc, sh))) return rc;
    e->gemm["mrcnn_mask_deconv"].cout = cout;
  }
  (void)kernel;
  return MRCNN_OK;
}

// Tile-width autotuning at build time: every legal BLOCK_N is timed on the layer's real shape (the result
// is bit-identical for any width: the K order per output element does not depend on it) and the
// fastest one is kept.  Small layers are dominated by wave quantisation / per-tile latency, which no
// static heuristic predicts well.
int autotune_block_n(mrcnn_engine* e, const mrcnn_conv_desc* d, const void* x, const GemmW& g, const void* residual,
                     void* out, ConvPlan* plan) {
  const int cout_total = d->out_mode == 1 ? 4 * d->cout : d->cout;
  int cap = 32;
  while (cap < cout_total && cap < 256) cap <<= 1;
  cudaEvent_t e0, e1;
  MRCNN_CHECK_CUDA(cudaEventCreate(&e0));
  MRCNN_CHECK_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  int best_bn = plan->block_n, best_epi = plan->epi_tma, best_occ2 = plan->occ2;
  const int n_epi = conv_plan_epi_tma_eligible(d) ? 2 : 1;
  static const bool allow_occ2 = !(getenv("MRCNN_B200_OCC2") && getenv("MRCNN_B200_OCC2")[0] == '0');
  for (int variant = 0; variant < 2 * n_epi; ++variant) {
    const int epi = variant % n_epi, occ2 = variant / n_epi;
    if (occ2 && !allow_occ2) continue;
    for (int bn = 32; bn <= cap; bn <<= 1) {
      if (d->out_mode == 1 && d->cout % bn != 0) continue;
      if (occ2 && bn > 128) continue;                   // two CTAs per SM: 2 x 2 x BLOCK_N TMEM columns
      ConvPlan trial;
      if (conv_plan_create_ex(d, x, g.w, g.scale, g.shift, residual, out, bn, epi, &trial) != MRCNN_OK) continue;
      trial.occ2 = occ2;
      float tmin = 1e30f;
      for (int rep = 0; rep < 10; ++rep) {       // 2 warm-up launches, minimum of 8
        MRCNN_CHECK_CUDA(cudaEventRecord(e0, e->stream));
        int rc = conv_plan_launch(&trial, e->stream);
        if (rc) return rc;
        MRCNN_CHECK_CUDA(cudaEventRecord(e1, e->stream));
        MRCNN_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MRCNN_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < tmin) tmin = ms;
      }
      if (tmin < best * 0.97f) {   // keep the earlier candidate unless the new one is clearly faster
        best = tmin;
        best_bn = bn;
        best_epi = epi;
        best_occ2 = occ2;
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (best_bn != plan->block_n || best_epi != plan->epi_tma) {
    int rc = conv_plan_create_ex(d, x, g.w, g.scale, g.shift, residual, out, best_bn, best_epi, plan);
    if (rc) return rc;
  }
  plan->occ2 = best_occ2;
  return MRCNN_OK;
}

struct Act {   // NHWC bf16 activation
  __nv_bfloat16* p;
  int n, h, w, c;
};

int add_conv(mrcnn_engine* e, const std::string& stage, const std::string& wname, Act in, int k, int stride, int relu,
             const __nv_bfloat16* residual, int res_up2, const std::string& out_name, Act* out, int out_f32 = 0,
             int out_ld = 0, int out_mode = 0, void** raw_out = nullptr) {
  auto it = e->gemm.find(wname);
  if (it == e->gemm.end()) {
    mrcnn_set_error("engine: no GEMM weights for %s", wname.c_str());
    return MRCNN_ERR_NOTFOUND;
  }
  const GemmW& g = it->second;
  mrcnn_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.n = in.n; d.h = in.h; d.w = in.w; d.cin = in.c;
  d.kh = k; d.kw = k; d.stride = stride; d.pad = k == 3 ? 1 : 0;
  d.cout = g.cout;
  d.relu = relu;
  d.residual_upsample2 = res_up2;
  d.out_dtype = out_f32 ? MRCNN_DTYPE_F32 : MRCNN_DTYPE_BF16;
  d.out_mode = out_mode;
  d.out_ld = out_ld;
  const int oh = (in.h + 2 * d.pad - k) / stride + 1, ow = (in.w + 2 * d.pad - k) / stride + 1;
  const int ld = out_ld ? out_ld : g.cout;
  const int mult = out_mode == 1 ? 4 : 1;
  Tensor t;
  int rc = new_tensor(e, out_name, out_f32 ? DT_F32 : DT_BF16, (size_t)in.n * oh * ow * mult * ld, &t, /*zero=*/true);
  if (rc) return rc;
  ConvPlan* plan = new ConvPlan();
  e->plans.push_back(plan);
  rc = conv_plan_create(&d, in.p, g.w, g.scale, g.shift, residual, t.ptr, 0, plan);
  if (rc) return rc;
  if (e->autotune) {
    // optional on-disk cache of the choices (MRCNN_B200_AUTOTUNE_CACHE=<file>): a second process — e.g. the same
    // workload under ncu — builds exactly the same launch plan without re-timing (and without the timing launches)
    char key[160];
    snprintf(key, sizeof(key), "%s:%d:%d:%d:%d:%d:%d:%d:%d", out_name.c_str(), d.n, d.h, d.w, d.cin, d.cout, k, stride, out_mode);
    auto hit = e->tune_cache.find(key);
    if (hit != e->tune_cache.end()) {
      const int c_epi = hit->second.second & 1, c_occ2 = (hit->second.second >> 1) & 1;     // second field: epi_tma | occ2 << 1
      if (hit->second.first != plan->block_n || c_epi != plan->epi_tma) {
        rc = conv_plan_create_ex(&d, in.p, g.w, g.scale, g.shift, residual, t.ptr, hit->second.first, c_epi, plan);
        if (rc) return rc;
      }
      plan->occ2 = c_occ2;
    } else {
      rc = autotune_block_n(e, &d, in.p, g, residual, t.ptr, plan);
      if (rc) return rc;
      e->tune_cache[key] = std::make_pair(plan->block_n, plan->epi_tma | (plan->occ2 << 1));
      e->tune_cache_dirty = true;
    }
  }
  e->flops += plan->flops;
  e->steps.push_back({stage, [plan](cudaStream_t st) { return conv_plan_launch(plan, st); }, "conv_gemm", out_name, plan->flops});
  if (out) {
    out->p = static_cast<__nv_bfloat16*>(t.ptr);
    out->n = in.n;
    out->h = out_mode == 1 ? 2 * oh : oh;
    out->w = out_mode == 1 ? 2 * ow : ow;
    out->c = ld;
  }
  if (raw_out) *raw_out = t.ptr;
  return MRCNN_OK;
}

#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

int build_graph(mrcnn_engine* e) {
  const mrcnn_engine_config& c = e->cfg;
  const int B = c.batch_size, S = c.image_size, NC = c.num_classes, R = c.post_nms_rois, D = c.detection_max_instances;
  const int PY = c.top_down_pyramid_size, apl = c.anchors_per_location;
  Tensor t_img, t_meta;
  RC(new_tensor(e, "input_image", DT_F32, (size_t)B * S * S * 3, &t_img, true));
  RC(new_tensor(e, "input_image_meta", DT_F32, (size_t)B * e->meta_size(), &t_meta, true));

  // ---- stem --------------------------------------------------------------------------------
  Tensor t_col;
  RC(new_tensor(e, "conv1_im2col", DT_BF16, (size_t)B * (S / 2) * (S / 2) * 192, &t_col));
  {
    const float* img = static_cast<const float*>(t_img.ptr);
    __nv_bfloat16* col = static_cast<__nv_bfloat16*>(t_col.ptr);
    e->steps.push_back({"backbone", [=](cudaStream_t st) { return launch_stem_im2col(img, B, S, col, st); }, "stem_im2col"});
  }
  Act x = {static_cast<__nv_bfloat16*>(t_col.ptr), 1, 1, B * (S / 2) * (S / 2), 192};
  Act c1;
  RC(add_conv(e, "backbone", "conv1", x, 1, 1, 1, nullptr, 0, "conv1_out", &c1));
  c1.n = B; c1.h = S / 2; c1.w = S / 2; c1.c = 64;
  Tensor t_pool;
  RC(new_tensor(e, "C1", DT_BF16, (size_t)B * (S / 4) * (S / 4) * 64, &t_pool));
  {
    const __nv_bfloat16* src = c1.p;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(t_pool.ptr);
    e->steps.push_back({"backbone", [=](cudaStream_t st) { return launch_maxpool3x3s2(src, B, S / 2, S / 2, 64, dst, st); }, "maxpool"});
  }
  x = {static_cast<__nv_bfloat16*>(t_pool.ptr), B, S / 4, S / 4, 64};

  // ---- ResNet-101 stages 2..5 ----------------------------------------------------------------
  const int nblocks[4] = {3, 4, 23, 3};
  Act C[6];
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < nblocks[s]; ++b) {
      const std::string blk(1, (char)('a' + b));
      const std::string base = "res" + std::to_string(s + 2) + blk + "_branch";
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      Act y1, y2, sc, y3;
      RC(add_conv(e, "backbone", base + "2a", x, 1, stride, 1, nullptr, 0, base + "2a_out", &y1));
      if (b == 0) RC(add_conv(e, "backbone", base + "1", x, 1, stride, 0, nullptr, 0, base + "1_out", &sc));
      else sc = x;
      RC(add_conv(e, "backbone", base + "2b", y1, 3, 1, 1, nullptr, 0, base + "2b_out", &y2));
      const std::string oname = (b == nblocks[s] - 1) ? ("C" + std::to_string(s + 2)) : ("res" + std::to_string(s + 2) + blk + "_out");
      RC(add_conv(e, "backbone", base + "2c", y2, 1, 1, 1, sc.p, 0, oname, &y3));
      x = y3;
    }
    C[s + 2] = x;
  }

  // ---- FPN ---------------------------------------------------------------------------------------
  Act p5, p4, p3, p2, P[7];
  RC(add_conv(e, "fpn", "fpn_c5p5", C[5], 1, 1, 0, nullptr, 0, "fpn_p5_pre", &p5));
  RC(add_conv(e, "fpn", "fpn_c4p4", C[4], 1, 1, 0, p5.p, 1, "fpn_p4add", &p4));
  RC(add_conv(e, "fpn", "fpn_c3p3", C[3], 1, 1, 0, p4.p, 1, "fpn_p3add", &p3));
  RC(add_conv(e, "fpn", "fpn_c2p2", C[2], 1, 1, 0, p3.p, 1, "fpn_p2add", &p2));
  RC(add_conv(e, "fpn", "fpn_p2", p2, 3, 1, 0, nullptr, 0, "P2", &P[2]));
  RC(add_conv(e, "fpn", "fpn_p3", p3, 3, 1, 0, nullptr, 0, "P3", &P[3]));
  RC(add_conv(e, "fpn", "fpn_p4", p4, 3, 1, 0, nullptr, 0, "P4", &P[4]));
  RC(add_conv(e, "fpn", "fpn_p5", p5, 3, 1, 0, nullptr, 0, "P5", &P[5]));
  {
    Tensor t6;
    const int h6 = (P[5].h + 1) / 2, w6 = (P[5].w + 1) / 2;
    RC(new_tensor(e, "P6", DT_BF16, (size_t)B * h6 * w6 * PY, &t6));
    const __nv_bfloat16* src = P[5].p;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(t6.ptr);
    const int h5 = P[5].h, w5 = P[5].w;
    e->steps.push_back({"fpn", [=](cudaStream_t st) { return launch_subsample2(src, B, h5, w5, PY, dst, st); }, "subsample"});
    P[6] = {dst, B, h6, w6, PY};
  }
  for (int l = 2; l <= 6; ++l) e->feat[l - 2] = P[l].h;

  // ---- RPN ---------------------------------------------------------------------------------------
  int A = 0;
  for (int l = 2; l <= 6; ++l) A += P[l].h * P[l].w * apl;
  e->num_anchors = A;
  Tensor t_rc, t_rb, t_anch;
  RC(new_tensor(e, "rpn_class", DT_F32, (size_t)B * A * 2, &t_rc, true));
  RC(new_tensor(e, "rpn_bbox", DT_F32, (size_t)B * A * 4, &t_rb, true));
  RC(new_tensor(e, "anchors", DT_F32, (size_t)A * 4, &t_anch, true));
  int level_off = 0;
  for (int l = 2; l <= 6; ++l) {
    Act sh;
    void* head = nullptr;
    const std::string ls = std::to_string(l);
    RC(add_conv(e, "rpn", "rpn_conv_shared", P[l], 3, 1, 1, nullptr, 0, "rpn_shared_p" + ls, &sh));
    RC(add_conv(e, "rpn", "rpn_head", sh, 1, 1, 0, nullptr, 0, "rpn_head_p" + ls, nullptr, 1, 32, 0, &head));
    const int hw = P[l].h * P[l].w;
    const int off = level_off;
    float* rc_ = static_cast<float*>(t_rc.ptr);
    float* rb_ = static_cast<float*>(t_rb.ptr);
    const float* hd = static_cast<const float*>(head);
    e->steps.push_back({"rpn", [=](cudaStream_t st) { return launch_rpn_post(hd, 32, B, hw, apl, A, off, rc_, rb_, st); }, "rpn_post"});
    level_off += hw * apl;
  }

  // ---- ProposalLayer -----------------------------------------------------------------------------
  const int K = c.pre_nms_limit < A ? c.pre_nms_limit : A;
  Tensor t_rois, t_topk, t_keep, t_kcnt;
  RC(new_tensor(e, "rpn_rois", DT_F32, (size_t)B * R * 4, &t_rois, true));
  RC(new_tensor(e, "topk_idx", DT_I32, (size_t)B * K, &t_topk, true));
  RC(new_tensor(e, "keep_idx", DT_I32, (size_t)B * R, &t_keep, true));
  RC(new_tensor(e, "keep_count", DT_I32, (size_t)B, &t_kcnt, true));
  {
    const float* rc_ = static_cast<const float*>(t_rc.ptr);
    const float* rb_ = static_cast<const float*>(t_rb.ptr);
    const float* an = static_cast<const float*>(t_anch.ptr);
    float* rois = static_cast<float*>(t_rois.ptr);
    int32_t* tk = static_cast<int32_t*>(t_topk.ptr);
    int32_t* kp = static_cast<int32_t*>(t_keep.ptr);
    int32_t* kc = static_cast<int32_t*>(t_kcnt.ptr);
    const mrcnn_engine_config cc = c;
    e->steps.push_back({"proposal", [=](cudaStream_t st) {
      return mrcnn_proposal_layer(rc_, rb_, an, 0, B, A, cc.pre_nms_limit, R, cc.rpn_nms_threshold, cc.rpn_bbox_std_dev,
                                  rois, tk, kp, kc, nullptr, 0, st);
    }, "proposal"});
  }

  // ---- class head --------------------------------------------------------------------------------
  const int PS = c.pool_size, MP = c.mask_pool_size;
  Tensor t_pooled, t_lvl;
  RC(new_tensor(e, "pooled", DT_BF16, (size_t)B * R * PS * PS * PY, &t_pooled));
  RC(new_tensor(e, "roi_levels", DT_I32, (size_t)B * R, &t_lvl, true));
  const float image_area = (float)S * (float)S;
  struct FeatPack { const void* p[4]; int h[4]; int w[4]; };
  FeatPack fp;
  for (int l = 0; l < 4; ++l) { fp.p[l] = P[l + 2].p; fp.h[l] = P[l + 2].h; fp.w[l] = P[l + 2].w; }
  {
    const float* rois = static_cast<const float*>(t_rois.ptr);
    void* pooled = t_pooled.ptr;
    int32_t* lv = static_cast<int32_t*>(t_lvl.ptr);
    e->steps.push_back({"roialign", [=](cudaStream_t st) {
      return launch_pyramid_roi_align(fp.p, fp.h, fp.w, PY, MRCNN_DTYPE_BF16, rois, 4, B, R, PS, image_area, pooled, lv, st);
    }, "roialign"});
  }
  Act pooled = {static_cast<__nv_bfloat16*>(t_pooled.ptr), 1, 1, B * R, PS * PS * PY};
  Act fc1, fc2;
  void* head = nullptr;
  RC(add_conv(e, "class_head", "mrcnn_class_conv1", pooled, 1, 1, 1, nullptr, 0, "mrcnn_class_fc1", &fc1));
  RC(add_conv(e, "class_head", "mrcnn_class_conv2", fc1, 1, 1, 1, nullptr, 0, "mrcnn_class_fc2", &fc2));
  RC(add_conv(e, "class_head", "class_head", fc2, 1, 1, 0, nullptr, 0, "mrcnn_class_head_raw", nullptr, 1, 32, 0, &head));
  MRCNN_REQUIRE(5 * NC <= 32, "engine: NUM_CLASSES=%d too large for the fused class/bbox head (max 6)", NC);
  Tensor t_cls, t_bbox;
  RC(new_tensor(e, "mrcnn_class", DT_F32, (size_t)B * R * NC, &t_cls, true));
  RC(new_tensor(e, "mrcnn_bbox", DT_F32, (size_t)B * R * NC * 4, &t_bbox, true));
  {
    const float* hd = static_cast<const float*>(head);
    float* pc = static_cast<float*>(t_cls.ptr);
    float* pb = static_cast<float*>(t_bbox.ptr);
    e->steps.push_back({"class_head", [=](cudaStream_t st) { return launch_class_post(hd, 32, B * R, NC, pc, pb, st); }, "class_post"});
  }

  // ---- DetectionLayer ----------------------------------------------------------------------------
  Tensor t_det;
  RC(new_tensor(e, "detections", DT_F32, (size_t)B * D * 6, &t_det, true));
  {
    const float* rois = static_cast<const float*>(t_rois.ptr);
    const float* pc = static_cast<const float*>(t_cls.ptr);
    const float* pb = static_cast<const float*>(t_bbox.ptr);
    const float* meta = static_cast<const float*>(t_meta.ptr);
    float* det = static_cast<float*>(t_det.ptr);
    const mrcnn_engine_config cc = c;
    const int ms = e->meta_size();
    e->steps.push_back({"detection", [=](cudaStream_t st) {
      return mrcnn_detection_layer(rois, pc, pb, meta, ms, B, R, NC, D, cc.detection_min_confidence,
                                   cc.detection_nms_threshold, cc.bbox_std_dev, det, st);
    }, "detection"});
  }

  // ---- mask head ---------------------------------------------------------------------------------
  // The mask branch runs on the detections that exist: `detections` is zero-padded to DETECTION_MAX_INSTANCES rows per
  // image, the reference computes masks for the padding too and unmold_detections throws them away (mrcnn/model.py:
  // 2575-2577).  Here a flag per 128-row M tile (mask_tile_flags_kernel, from the class-id column) lets ROIAlign and the
  // five mask-head GEMMs skip tiles that hold padding only; mrcnn_mask rows of padded detections read as zeros.
  // Results of real detections are unchanged (rows of a GEMM are independent; a 3x3 tap never leaves its own ROI).
  // MRCNN_B200_SKIP_PADDED=0 computes everything, as the reference does.
  const bool skip_padded = !(getenv("MRCNN_B200_SKIP_PADDED") && getenv("MRCNN_B200_SKIP_PADDED")[0] == '0');
  Tensor t_pm, t_lvl_mask, t_skip;
  const int mask_rows = MP * MP, mask_tiles = (int)(((long long)B * D * mask_rows + 127) / 128);
  RC(new_tensor(e, "pooled_mask", DT_BF16, (size_t)B * D * MP * MP * PY, &t_pm, true));
  RC(new_tensor(e, "roi_levels_mask", DT_I32, (size_t)B * D, &t_lvl_mask, true));
  RC(new_tensor(e, "mask_tile_skip", DT_U8, (size_t)mask_tiles, &t_skip, true));
  const unsigned char* skip_flags = skip_padded ? static_cast<const unsigned char*>(t_skip.ptr) : nullptr;
  {
    const float* det = static_cast<const float*>(t_det.ptr);
    void* pm = t_pm.ptr;
    int32_t* lvm = static_cast<int32_t*>(t_lvl_mask.ptr);
    unsigned char* flags = static_cast<unsigned char*>(t_skip.ptr);
    if (skip_padded)
      e->steps.push_back({"roialign_mask", [=](cudaStream_t st) {
        return launch_mask_tile_flags(det, B * D, mask_rows, 128, mask_tiles, flags, st);
      }, "mask_tile_flags"});
    e->steps.push_back({"roialign_mask", [=](cudaStream_t st) {
      return launch_pyramid_roi_align(fp.p, fp.h, fp.w, PY, MRCNN_DTYPE_BF16, det, 6, B, D, MP, image_area, pm, lvm, st,
                                      skip_padded ? 4 : -1);
    }, "roialign"});
  }
  Act m = {static_cast<__nv_bfloat16*>(t_pm.ptr), B * D, MP, MP, PY};
  for (int i = 1; i <= 4; ++i) {
    Act o;
    RC(add_conv(e, "mask_head", "mrcnn_mask_conv" + std::to_string(i), m, 3, 1, 1, nullptr, 0,
                "mrcnn_mask_conv" + std::to_string(i) + "_out", &o));
    ConvPlan* cp = e->plans.back();
    if (skip_flags && cp->p.flat && cp->p.tw == 128 && cp->p.th == 1 && cp->p.nb == 1) RC(conv_plan_set_tile_skip(cp, skip_flags));
    m = o;
  }
  MRCNN_REQUIRE(NC <= 8, "engine: NUM_CLASSES=%d too large for the mask logits pitch (max 8)", NC);
  Tensor t_mask;
  RC(new_tensor(e, "mrcnn_mask", DT_F32, (size_t)B * D * 4 * MP * MP * NC, &t_mask, true));
  if (PY == 256 && NC <= 4) {
    // Conv2DTranspose(2x2, s2) + ReLU + Conv2D(1x1 -> NC) + sigmoid as ONE GEMM: the 1x1 conv runs in the
    // epilogue on the bf16-rounded deconv output, so the [B*D,28,28,256] tensor never touches HBM
    const GemmW& gd = e->gemm["mrcnn_mask_deconv"];
    const GemmW& gm = e->gemm["mrcnn_mask"];
    mrcnn_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.n = m.n; d.h = m.h; d.w = m.w; d.cin = m.c; d.kh = 1; d.kw = 1; d.stride = 1; d.pad = 0;
    d.cout = gd.cout; d.relu = 1; d.out_dtype = MRCNN_DTYPE_BF16; d.out_mode = 1;
    ConvPlan* plan = new ConvPlan();
    e->plans.push_back(plan);
    RC(conv_plan_create(&d, m.p, gd.w, gd.scale, gd.shift, nullptr, t_mask.ptr, 256, plan));
    RC(conv_plan_fuse_mask_logits(plan, gm.w, gm.shift, NC, t_mask.ptr, /*unit_scale=*/1));   // Conv2DTranspose has no BN
    if (skip_flags) RC(conv_plan_set_tile_skip(plan, skip_flags));
    e->flops += plan->flops;
    e->steps.push_back({"mask_head", [plan](cudaStream_t st) { return conv_plan_launch(plan, st); }, "conv_gemm",
                        "mrcnn_mask (deconv+1x1+sigmoid)", plan->flops});
  } else {
    Act dc;
    RC(add_conv(e, "mask_head", "mrcnn_mask_deconv", m, 1, 1, 1, nullptr, 0, "mrcnn_mask_deconv_out", &dc, 0, 0, 1));
    void* mlog = nullptr;
    RC(add_conv(e, "mask_head", "mrcnn_mask", dc, 1, 1, 0, nullptr, 0, "mrcnn_mask_logits", nullptr, 1, 8, 0, &mlog));
    const float* lg = static_cast<const float*>(mlog);
    float* out = static_cast<float*>(t_mask.ptr);
    const size_t M = (size_t)B * D * 4 * MP * MP;
    e->steps.push_back({"mask_head", [=](cudaStream_t st) { return launch_mask_post(lg, 8, M, NC, out, st); }, "mask_post"});
  }
  if (skip_flags && ((size_t)4 * MP * MP * NC) % 4 == 0) {
    const float* det = static_cast<const float*>(t_det.ptr);
    float* out = static_cast<float*>(t_mask.ptr);
    const size_t per_roi = (size_t)4 * MP * MP * NC;
    e->steps.push_back({"mask_head", [=](cudaStream_t st) { return launch_mask_zero_padded(det, B * D, per_roi, out, st); },
                        "mask_zero_padded"});
  }

  // ---- stage table for timing / run_stage ----------------------------------------------------------
  for (const Step& s : e->steps)
    if (e->stage_names.empty() || e->stage_names.back() != s.stage) e->stage_names.push_back(s.stage);
  e->stage_events.resize(e->stage_names.size() + 1);
  for (auto& ev : e->stage_events) MRCNN_CHECK_CUDA(cudaEventCreate(&ev));
  return MRCNN_OK;
}

int run_steps(mrcnn_engine* e, const char* only_stage, bool timed) {
  size_t si = 0, run_i = 0;
  std::string cur;
  for (const Step& s : e->steps) {
    if (only_stage && s.stage != only_stage) continue;
    if (timed && s.stage != cur) {
      while (si < e->stage_names.size() && e->stage_names[si] != s.stage) ++si;
      MRCNN_CHECK_CUDA(cudaEventRecord(e->stage_events[si], e->stream));
      cur = s.stage;
    }
    const size_t step_i = (size_t)(&s - &e->steps[0]);
    if (e->profiling == 1 && !only_stage) MRCNN_CHECK_CUDA(cudaEventRecord(e->step_events[step_i], e->stream));
    if (e->profiling == 2 && !only_stage && run_i < e->run_first.size() && (size_t)e->run_first[run_i] == step_i) {
      MRCNN_CHECK_CUDA(cudaEventRecord(e->run_events[e->run_calls % mrcnn_engine::kRunSets][run_i], e->stream));
      ++run_i;
    }
    int rc = s.run(e->stream);
    if (rc) return rc;
  }
  if (e->profiling == 1 && !only_stage) MRCNN_CHECK_CUDA(cudaEventRecord(e->step_events[e->steps.size()], e->stream));
  if (e->profiling == 2 && !only_stage) {
    MRCNN_CHECK_CUDA(cudaEventRecord(e->run_events[e->run_calls % mrcnn_engine::kRunSets][e->run_first.size()], e->stream));
    ++e->run_calls;
  }
  if (timed) MRCNN_CHECK_CUDA(cudaEventRecord(e->stage_events[e->stage_names.size()], e->stream));
  return MRCNN_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" int mrcnn_engine_create(const mrcnn_engine_config* cfg, int device, mrcnn_engine** out) {
  MRCNN_REQUIRE(cfg && out, "engine_create: null pointer");
  MRCNN_REQUIRE(cfg->batch_size >= 1, "engine_create: batch_size must be >= 1");
  // mrcnn/model.py:1944-1948: image size must be divisible by 2^6
  MRCNN_REQUIRE(cfg->image_size >= 64 && cfg->image_size % 64 == 0,
                "Image size must be dividable by 2 at least 6 times to avoid fractions when downscaling and upscaling."
                "For example, use 256, 320, 384, 448, 512, ... etc. (got %d)", cfg->image_size);
  MRCNN_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 6, "engine_create: num_classes must be in [1,6]");
  MRCNN_REQUIRE(cfg->top_down_pyramid_size % 64 == 0 && cfg->fc_layers_size % 64 == 0, "engine_create: pyramid / fc sizes must be multiples of 64");
  MRCNN_REQUIRE(cfg->anchors_per_location >= 1 && 6 * cfg->anchors_per_location <= 32, "engine_create: anchors_per_location must be in [1,5]");
  const int strides[5] = {4, 8, 16, 32, 64};
  for (int i = 0; i < 5; ++i)
    MRCNN_REQUIRE(cfg->backbone_strides[i] == strides[i], "engine_create: BACKBONE_STRIDES must be [4,8,16,32,64] (resnet101)");
  int ndev = 0;
  MRCNN_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  MRCNN_REQUIRE(device >= 0 && device < ndev, "engine_create: device %d not available (%d devices)", device, ndev);
  MRCNN_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MRCNN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    mrcnn_set_error("engine_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return MRCNN_ERR_UNSUPPORTED;
  }
  mrcnn_engine* e = new mrcnn_engine();
  e->cfg = *cfg;
  e->device = device;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    mrcnn_set_error("engine_create: cudaStreamCreate failed");
    return MRCNN_ERR_CUDA;
  }
  // small batches are launch-bound (~150 launches in ~2 ms at B = 1): replay the plan as a CUDA graph there
  // (measured: B=1 2.07 -> 1.96 ms, B=8 3.05 -> 2.94 ms, B=64 neutral); MRCNN_B200_GRAPH=0/1 overrides
  e->use_graph = cfg->batch_size <= 16;
  if (const char* g = getenv("MRCNN_B200_GRAPH")) e->use_graph = g[0] == '1';
  const char* at = getenv("MRCNN_B200_AUTOTUNE");
  e->autotune = !(at && at[0] == '0');
  if (const char* cp = getenv("MRCNN_B200_AUTOTUNE_CACHE")) {
    e->tune_cache_path = cp;
    if (FILE* f = fopen(cp, "r")) {
      char key[160];
      int bn, epi;
      while (fscanf(f, "%159s %d %d", key, &bn, &epi) == 3) e->tune_cache[key] = std::make_pair(bn, epi);
      fclose(f);
    }
  }
  build_layer_table(e);
  *out = e;
  return MRCNN_OK;
}

extern "C" void mrcnn_engine_destroy(mrcnn_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  for (void* p : e->allocs) cudaFree(p);
  for (ConvPlan* p : e->plans) {
    if (p->w2_table) cudaFree(p->w2_table);
    delete p;
  }
  for (auto ev : e->stage_events) cudaEventDestroy(ev);
  for (auto ev : e->step_events) cudaEventDestroy(ev);
  if (e->unmold_ws) cudaFree(e->unmold_ws);
  if (e->d_windows) cudaFree(e->d_windows);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  for (auto& sl : e->slots) {
    if (sl.masks) cudaFree(sl.masks);
    if (sl.bits) cudaFree(sl.bits);
    if (sl.dense) cudaFree(sl.dense);
    if (sl.dense_t0) cudaEventDestroy(sl.dense_t0);
    if (sl.dense_t1) cudaEventDestroy(sl.dense_t1);
    if (sl.computed) cudaEventDestroy(sl.computed);
    if (sl.copied) cudaEventDestroy(sl.copied);
  }
  if (e->pre_maps) cudaFree(e->pre_maps);
  if (e->upload_stream) { cudaStreamSynchronize(e->upload_stream); cudaStreamDestroy(e->upload_stream); }
  for (int k = 0; k < 2; ++k) {
    if (e->up_maps[k]) cudaFree(e->up_maps[k]);
    if (e->up_done[k]) cudaEventDestroy(e->up_done[k]);
    if (e->up_consumed[k]) cudaEventDestroy(e->up_consumed[k]);
  }
  if (e->pre_rgb) cudaFree(e->pre_rgb);
  if (e->pre_small) cudaFree(e->pre_small);
  cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int mrcnn_engine_num_layers(const mrcnn_engine* e) { return e ? (int)e->layers.size() : 0; }

extern "C" int mrcnn_engine_layer_info(const mrcnn_engine* e, int index, const char** name, int* kind, int* num_weights, int* shape4) {
  MRCNN_REQUIRE(e && index >= 0 && index < (int)e->layers.size(), "layer_info: bad index");
  const LayerSpec& l = e->layers[index];
  if (name) *name = l.name.c_str();
  if (kind) *kind = l.kind;
  if (num_weights) *num_weights = l.nweights;
  if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = l.shape[i];
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_set_weight(mrcnn_engine* e, const char* layer_name, int weight_index, const float* data, size_t count) {
  MRCNN_REQUIRE(e && layer_name && data, "set_weight: null pointer");
  MRCNN_REQUIRE(!e->finalized, "set_weight: engine already finalized");
  auto it = e->layer_index.find(layer_name);
  if (it == e->layer_index.end()) {
    mrcnn_set_error("set_weight: no layer named '%s' in the inference graph", layer_name);
    return MRCNN_ERR_NOTFOUND;
  }
  LayerSpec& l = e->layers[it->second];
  MRCNN_REQUIRE(weight_index >= 0 && weight_index < l.nweights, "set_weight: layer %s has %d weights", layer_name, l.nweights);
  const size_t expect = l.count(weight_index);
  MRCNN_REQUIRE(count == expect, "set_weight: layer %s weight %d has %zu elements, expected %zu (shape mismatch)", layer_name,
                weight_index, count, expect);
  l.host[weight_index].assign(data, data + count);
  l.set[weight_index] = true;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_finalize(mrcnn_engine* e, int allow_missing) {
  MRCNN_REQUIRE(e, "finalize: null engine");
  MRCNN_REQUIRE(!e->finalized, "finalize: already finalized");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  for (LayerSpec& l : e->layers)
    for (int wi = 0; wi < l.nweights; ++wi)
      if (!l.set[wi]) {
        if (!allow_missing) {
          mrcnn_set_error("finalize: layer '%s' weight %d was never loaded", l.name.c_str(), wi);
          return MRCNN_ERR_NOTFOUND;
        }
        if (l.kind == KIND_BN) {   // identity BN: gamma 1, beta 0, mean 0, var 1
          l.host[wi].assign(l.count(wi), (wi == 0 || wi == 3) ? 1.0f : 0.0f);
          l.set[wi] = true;
        } else {
          l.host[wi].assign(l.count(wi), 0.0f);
          l.set[wi] = true;
        }
      }
  int rc = build_weights(e);
  if (rc) return rc;
  rc = build_graph(e);
  if (rc == MRCNN_OK && e->tune_cache_dirty && !e->tune_cache_path.empty()) {
    if (FILE* f = fopen(e->tune_cache_path.c_str(), "w")) {
      for (const auto& kv : e->tune_cache) fprintf(f, "%s %d %d\n", kv.first.c_str(), kv.second.first, kv.second.second);
      fclose(f);
    }
    e->tune_cache_dirty = false;
  }
  if (rc) return rc;
  for (LayerSpec& l : e->layers) { l.host.clear(); l.host.shrink_to_fit(); l.host.resize(l.nweights); }
  MRCNN_CHECK_CUDA(cudaDeviceSynchronize());
  e->finalized = true;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_set_anchors(mrcnn_engine* e, const float* anchors, int num_anchors) {
  MRCNN_REQUIRE(e && anchors && e->finalized, "set_anchors: engine not finalized / null pointer");
  MRCNN_REQUIRE(num_anchors == e->num_anchors, "set_anchors: got %d anchors, the graph has %d", num_anchors, e->num_anchors);
  MRCNN_CHECK_CUDA(cudaMemcpy(e->tensors["anchors"].ptr, anchors, (size_t)num_anchors * 16, cudaMemcpyHostToDevice));
  return MRCNN_OK;
}

static int predict_internal(mrcnn_engine* e, const float* molded, cudaMemcpyKind molded_kind, const float* image_metas,
                            cudaMemcpyKind metas_kind) {
  MRCNN_REQUIRE(e && e->finalized, "predict: engine not finalized (load weights first)");
  MRCNN_REQUIRE(molded && image_metas, "predict: null input");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  const Tensor& ti = e->tensors["input_image"];
  const Tensor& tm = e->tensors["input_image_meta"];
  if (molded != ti.ptr) MRCNN_CHECK_CUDA(cudaMemcpyAsync(ti.ptr, molded, ti.bytes, molded_kind, e->stream));
  if (image_metas != tm.ptr) MRCNN_CHECK_CUDA(cudaMemcpyAsync(tm.ptr, image_metas, tm.bytes, metas_kind, e->stream));
  // The launch plan is static (fixed pointers, fixed geometry): outside profiling it is captured once into a CUDA
  // graph (programmatic-dependent-launch edges included) and replayed with one cudaGraphLaunch per predict.
  if (e->use_graph && e->profiling == 0) {
    if (!e->graph_exec) {
      cudaGraph_t graph = nullptr;
      MRCNN_CHECK_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      const unsigned long long n0 = mrcnn_kernel_launch_count();
      const int rc = run_steps(e, nullptr, false);
      e->launches_per_predict = mrcnn_kernel_launch_count() - n0;
      const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      MRCNN_CHECK_CUDA(ce);
      const cudaError_t ie = cudaGraphInstantiate(&e->graph_exec, graph, 0);
      e->graph_fresh = true;
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) {            // fall back to stream launches for good
        e->graph_exec = nullptr;
        e->use_graph = false;
        cudaGetLastError();
        return run_steps(e, nullptr, true);
      }
    }
    const bool first = e->graph_fresh;
    e->graph_fresh = false;
    MRCNN_CHECK_CUDA(cudaGraphLaunch(e->graph_exec, e->stream));
    if (!first) mrcnn_count_launch(e->launches_per_predict);     // the capture pass already counted its launches
    return MRCNN_OK;
  }
  return run_steps(e, nullptr, true);
}

extern "C" int mrcnn_engine_predict(mrcnn_engine* e, const float* molded, const float* image_metas, int inputs_on_host, int async) {
  const cudaMemcpyKind kind = inputs_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  int rc = predict_internal(e, molded, kind, image_metas, kind);
  if (rc) return rc;
  if (!async) MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_tensor(const mrcnn_engine* e, const char* name, void** device_ptr, size_t* bytes) {
  MRCNN_REQUIRE(e && name, "engine_tensor: null pointer");
  if (!strncmp(name, "unmold_masks", 12)) {   // [B,H0,W0,D] uint8 of result slot 0 / 1 (sized by the last detect call)
    const int slot = !strcmp(name + 12, "#1") ? 1 : 0;
    MRCNN_REQUIRE(name[12] == 0 || slot == 1, "engine_tensor: unknown tensor '%s'", name);
    MRCNN_REQUIRE(e->slots[slot].masks, "engine_tensor: '%s' has not been produced yet", name);
    if (device_ptr) *device_ptr = e->slots[slot].masks;
    if (bytes) *bytes = e->slots[slot].masks_bytes;
    return MRCNN_OK;
  }
  if (!strncmp(name, "unmold_mask_bits", 16)) {   // [B,H0*W0,DW] uint32 of result slot 0 / 1
    const int slot = !strcmp(name + 16, "#1") ? 1 : 0;
    MRCNN_REQUIRE(name[16] == 0 || slot == 1, "engine_tensor: unknown tensor '%s'", name);
    MRCNN_REQUIRE(e->slots[slot].bits, "engine_tensor: '%s' has not been produced yet", name);
    if (device_ptr) *device_ptr = e->slots[slot].bits;
    if (bytes) *bytes = e->slots[slot].bits_bytes;
    return MRCNN_OK;
  }
  auto it = e->tensors.find(name);
  if (it == e->tensors.end()) {
    mrcnn_set_error("engine_tensor: unknown tensor '%s'", name);
    return MRCNN_ERR_NOTFOUND;
  }
  if (device_ptr) *device_ptr = it->second.ptr;
  if (bytes) *bytes = it->second.bytes;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_read(const mrcnn_engine* e, const char* name, void* host_dst, size_t dst_bytes) {
  MRCNN_REQUIRE(e && name && host_dst, "engine_read: null pointer");
  auto it = e->tensors.find(name);
  if (it == e->tensors.end()) {
    mrcnn_set_error("engine_read: unknown tensor '%s'", name);
    return MRCNN_ERR_NOTFOUND;
  }
  const Tensor& t = it->second;
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  if (t.dtype == DT_BF16) {
    MRCNN_REQUIRE(dst_bytes >= t.elems * 4, "engine_read: %s needs %zu bytes (float32)", name, t.elems * 4);
    std::vector<uint16_t> tmp(t.elems);
    MRCNN_CHECK_CUDA(cudaMemcpy(tmp.data(), t.ptr, t.bytes, cudaMemcpyDeviceToHost));
    uint32_t* dst = static_cast<uint32_t*>(host_dst);
    for (size_t i = 0; i < t.elems; ++i) dst[i] = (uint32_t)tmp[i] << 16;
  } else {
    MRCNN_REQUIRE(dst_bytes >= t.bytes, "engine_read: %s needs %zu bytes", name, t.bytes);
    MRCNN_CHECK_CUDA(cudaMemcpy(host_dst, t.ptr, t.bytes, cudaMemcpyDeviceToHost));
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_write(mrcnn_engine* e, const char* name, const void* host_src, size_t src_bytes) {
  MRCNN_REQUIRE(e && name && host_src, "engine_write: null pointer");
  auto it = e->tensors.find(name);
  if (it == e->tensors.end()) {
    mrcnn_set_error("engine_write: unknown tensor '%s'", name);
    return MRCNN_ERR_NOTFOUND;
  }
  const Tensor& t = it->second;
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  if (t.dtype == DT_BF16) {   // source is float32, narrowed here
    MRCNN_REQUIRE(src_bytes == t.elems * 4, "engine_write: %s expects %zu float32 values", name, t.elems);
    const float* src = static_cast<const float*>(host_src);
    std::vector<uint16_t> tmp(t.elems);
    for (size_t i = 0; i < t.elems; ++i) tmp[i] = f2bf(src[i]);
    MRCNN_CHECK_CUDA(cudaMemcpy(t.ptr, tmp.data(), t.bytes, cudaMemcpyHostToDevice));
  } else {
    MRCNN_REQUIRE(src_bytes == t.bytes, "engine_write: %s expects %zu bytes", name, t.bytes);
    MRCNN_CHECK_CUDA(cudaMemcpy(t.ptr, host_src, t.bytes, cudaMemcpyHostToDevice));
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_run_stage(mrcnn_engine* e, const char* stage) {
  MRCNN_REQUIRE(e && e->finalized && stage, "run_stage: engine not finalized / null stage");
  bool known = false;
  for (const std::string& s : e->stage_names) known |= (s == stage);
  if (!known) {
    mrcnn_set_error("run_stage: unknown stage '%s'", stage);
    return MRCNN_ERR_NOTFOUND;
  }
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  int rc = run_steps(e, stage, false);
  if (rc) return rc;
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  return MRCNN_OK;
}

extern "C" void* mrcnn_engine_stream(const mrcnn_engine* e) { return e ? (void*)e->stream : nullptr; }

extern "C" int mrcnn_engine_stage_times(const mrcnn_engine* e, int max_stages, const char** names, float* ms) {
  MRCNN_REQUIRE(e && e->finalized, "stage_times: engine not finalized");
  const int n = (int)e->stage_names.size();
  for (int i = 0; i < n && i < max_stages; ++i) {
    if (names) names[i] = e->stage_names[i].c_str();
    if (ms) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, e->stage_events[i], e->stage_events[i + 1]) != cudaSuccess) t = -1.f;
      ms[i] = t;
    }
  }
  return n;
}

extern "C" double mrcnn_engine_flops(const mrcnn_engine* e) { return e ? e->flops : 0.0; }

extern "C" int mrcnn_engine_set_profiling(mrcnn_engine* e, int enable) {
  MRCNN_REQUIRE(e && e->finalized, "set_profiling: engine not finalized");
  MRCNN_REQUIRE(enable >= 0 && enable <= 2, "set_profiling: mode must be 0, 1 or 2");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  if (enable == 1 && e->step_events.empty()) {
    e->step_events.resize(e->steps.size() + 1);
    for (auto& ev : e->step_events) MRCNN_CHECK_CUDA(cudaEventCreate(&ev));
  }
  if (enable == 2) {
    if (e->run_first.empty()) {
      for (size_t i = 0; i < e->steps.size(); ++i)
        if (i == 0 || e->steps[i].kind != e->steps[i - 1].kind) e->run_first.push_back((int)i);
      for (auto& set : e->run_events) {
        set.resize(e->run_first.size() + 1);
        for (auto& ev : set) MRCNN_CHECK_CUDA(cudaEventCreate(&ev));
      }
    }
    e->run_calls = 0;
  }
  e->profiling = enable;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_step_info(mrcnn_engine* e, int index, const char** label, const char** kind, float* ms, double* flops) {
  MRCNN_REQUIRE(e && e->finalized, "step_info: engine not finalized");
  if (index < 0 || index >= (int)e->steps.size()) return MRCNN_ERR_NOTFOUND;
  const Step& s = e->steps[index];
  if (label) *label = s.label.empty() ? s.kind.c_str() : s.label.c_str();
  if (kind) *kind = s.kind.c_str();
  if (flops) *flops = s.flops;
  if (ms) {
    *ms = -1.f;
    if (!e->step_events.empty()) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, e->step_events[index], e->step_events[index + 1]) == cudaSuccess) *ms = t;
    }
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_kernel_times(mrcnn_engine* e, int max_kinds, const char** names, float* ms, int* launches) {
  MRCNN_REQUIRE(e && e->finalized && (!e->step_events.empty() || !e->run_first.empty()), "kernel_times: profiling was never enabled");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  e->kind_names.clear();
  std::vector<float> tot;
  std::vector<int> cnt;
  if (e->profiling == 2) {
    // average over the predicts still held by the event ring (the last kRunSets of them)
    const int nsets = (int)(e->run_calls < (unsigned long long)mrcnn_engine::kRunSets ? e->run_calls : mrcnn_engine::kRunSets);
    MRCNN_REQUIRE(nsets > 0, "kernel_times: no profiled predict has completed yet");
    const size_t nruns = e->run_first.size();
    for (size_t r = 0; r < nruns; ++r) {
      const size_t first = (size_t)e->run_first[r];
      const size_t last = r + 1 < nruns ? (size_t)e->run_first[r + 1] : e->steps.size();
      float sum = 0.f;
      for (int sidx = 0; sidx < nsets; ++sidx) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e->run_events[sidx][r], e->run_events[sidx][r + 1]) != cudaSuccess) {
          mrcnn_set_error("kernel_times: profiled predict not complete");
          return MRCNN_ERR_INVALID;
        }
        sum += t;
      }
      const std::string& kind = e->steps[first].kind;
      size_t k = 0;
      for (; k < e->kind_names.size(); ++k) if (e->kind_names[k] == kind) break;
      if (k == e->kind_names.size()) { e->kind_names.push_back(kind); tot.push_back(0.f); cnt.push_back(0); }
      tot[k] += sum / nsets;
      cnt[k] += (int)(last - first);
    }
    const int n = (int)e->kind_names.size();
    for (int k = 0; k < n && k < max_kinds; ++k) {
      if (names) names[k] = e->kind_names[k].c_str();
      if (ms) ms[k] = tot[k];
      if (launches) launches[k] = cnt[k];
    }
    return n;
  }
  MRCNN_REQUIRE(!e->step_events.empty(), "kernel_times: per-launch profiling was never enabled");
  for (size_t i = 0; i < e->steps.size(); ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e->step_events[i], e->step_events[i + 1]) != cudaSuccess) {
      mrcnn_set_error("kernel_times: no profiled predict has completed yet");
      return MRCNN_ERR_INVALID;
    }
    size_t k = 0;
    for (; k < e->kind_names.size(); ++k) if (e->kind_names[k] == e->steps[i].kind) break;
    if (k == e->kind_names.size()) { e->kind_names.push_back(e->steps[i].kind); tot.push_back(0.f); cnt.push_back(0); }
    tot[k] += t;
    cnt[k] += 1;
  }
  const int n = (int)e->kind_names.size();
  for (int k = 0; k < n && k < max_kinds; ++k) {
    if (names) names[k] = e->kind_names[k].c_str();
    if (ms) ms[k] = tot[k];
    if (launches) launches[k] = cnt[k];
  }
  return n;
}

// ---- unmold + result fetch (shared by detect_molded / detect_maps) ---------------------------------
static int ensure_scratch(mrcnn_engine* e, void** ptr, size_t* have, size_t need) {
  (void)e;
  if (*have >= need) return MRCNN_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  MRCNN_CHECK_CUDA(cudaMalloc(ptr, need));
  *have = need;
  return MRCNN_OK;
}

static int unmold_internal(mrcnn_engine* e, int slot, const int* orig_hw, const int32_t* windows_host, int mask_format) {
  const mrcnn_engine_config& c = e->cfg;
  const int B = c.batch_size, D = c.detection_max_instances;
  mrcnn_engine::ResultSlot& sl = e->slots[slot];
  RC(ensure_scratch(e, &e->unmold_ws, &e->unmold_ws_bytes, mrcnn_unmold_workspace_bytes(B, D)));
  if (sl.copy_pending) {                       // the slot's previous D2H must be complete before it is overwritten
    MRCNN_CHECK_CUDA(cudaEventSynchronize(sl.copied));
    sl.copy_pending = false;
  }
  const size_t npx = (size_t)orig_hw[0] * orig_hw[1];
  if (mask_format == 0) RC(ensure_scratch(e, &sl.masks, &sl.masks_bytes, (size_t)B * npx * D));
  else RC(ensure_scratch(e, &sl.bits, &sl.bits_bytes, (size_t)B * npx * mrcnn_mask_bits_words(D) * 4));
  if (!e->d_windows) MRCNN_CHECK_CUDA(cudaMalloc((void**)&e->d_windows, (size_t)B * 16));
  if (!sl.rois) {
    const std::string sfx = slot == 0 ? "" : "#1";
    Tensor t;
    RC(new_tensor(e, "unmold_rois" + sfx, DT_I32, (size_t)B * D * 4, &t, true)); sl.rois = static_cast<int32_t*>(t.ptr);
    RC(new_tensor(e, "unmold_class_ids" + sfx, DT_I32, (size_t)B * D, &t, true)); sl.class_ids = static_cast<int32_t*>(t.ptr);
    RC(new_tensor(e, "unmold_scores" + sfx, DT_F32, (size_t)B * D, &t, true)); sl.scores = static_cast<float*>(t.ptr);
    RC(new_tensor(e, "unmold_counts" + sfx, DT_I32, (size_t)B, &t, true)); sl.counts = static_cast<int32_t*>(t.ptr);
    MRCNN_CHECK_CUDA(cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming));
    MRCNN_CHECK_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
  }
  MRCNN_CHECK_CUDA(cudaMemcpyAsync(e->d_windows, windows_host, (size_t)B * 16, cudaMemcpyHostToDevice, e->stream));
  const int image_hw[2] = {c.image_size, c.image_size};
  const float* det = static_cast<const float*>(e->tensors["detections"].ptr);
  const float* mm = static_cast<const float*>(e->tensors["mrcnn_mask"].ptr);
  if (mask_format == 0)
    return mrcnn_unmold_detections(det, mm, B, D, 2 * c.mask_pool_size, 2 * c.mask_pool_size, c.num_classes, orig_hw, image_hw,
                                   e->d_windows, sl.rois, sl.class_ids, sl.scores, sl.counts, static_cast<uint8_t*>(sl.masks),
                                   e->unmold_ws, e->unmold_ws_bytes, e->stream);
  return mrcnn_unmold_detections_bits(det, mm, B, D, 2 * c.mask_pool_size, 2 * c.mask_pool_size, c.num_classes, orig_hw, image_hw,
                                      e->d_windows, sl.rois, sl.class_ids, sl.scores, sl.counts, static_cast<uint32_t*>(sl.bits),
                                      e->unmold_ws, e->unmold_ws_bytes, e->stream);
}

// D2H of one slot's results.  async: on the copy stream after the compute of this step (recorded event),
// so the main stream is free to start the next step; otherwise on the main stream.
static int fetch_internal(mrcnn_engine* e, int slot, bool async, const int* orig_hw, int32_t* rois_host,
                          int32_t* class_ids_host, float* scores_host, int32_t* counts_host, uint32_t* mask_bits_host) {
  const mrcnn_engine_config& c = e->cfg;
  const int B = c.batch_size, D = c.detection_max_instances;
  const size_t mbytes = (size_t)B * orig_hw[0] * orig_hw[1] * mrcnn_mask_bits_words(D) * 4;
  mrcnn_engine::ResultSlot& sl = e->slots[slot];
  cudaStream_t st = e->stream;
  if (async) {
    if (!e->copy_stream) MRCNN_CHECK_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    MRCNN_CHECK_CUDA(cudaEventRecord(sl.computed, e->stream));
    MRCNN_CHECK_CUDA(cudaStreamWaitEvent(e->copy_stream, sl.computed, 0));
    st = e->copy_stream;
  }
  if (rois_host) MRCNN_CHECK_CUDA(cudaMemcpyAsync(rois_host, sl.rois, (size_t)B * D * 16, cudaMemcpyDeviceToHost, st));
  if (class_ids_host) MRCNN_CHECK_CUDA(cudaMemcpyAsync(class_ids_host, sl.class_ids, (size_t)B * D * 4, cudaMemcpyDeviceToHost, st));
  if (scores_host) MRCNN_CHECK_CUDA(cudaMemcpyAsync(scores_host, sl.scores, (size_t)B * D * 4, cudaMemcpyDeviceToHost, st));
  if (counts_host) MRCNN_CHECK_CUDA(cudaMemcpyAsync(counts_host, sl.counts, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  if (mask_bits_host) MRCNN_CHECK_CUDA(cudaMemcpyAsync(mask_bits_host, sl.bits, mbytes, cudaMemcpyDeviceToHost, st));
  sl.dense_images = 0;
  if (mask_bits_host && e->dense_host && e->dense_n > 0) {
    // hybrid delivery: the masks of the first dense_n images leave as the reference's dense [H, W, N] bytes, expanded from
    // the bits by a small kernel and copied by the DMA engine, so the host's cores only expand the remaining images
    const int n = e->dense_n < B ? e->dense_n : B;
    const size_t npx = (size_t)orig_hw[0] * orig_hw[1];
    RC(ensure_scratch(e, &sl.dense, &sl.dense_bytes, (size_t)B * npx * D));   // sized once: the share varies from call to call
    if (!sl.dense_t0) {
      MRCNN_CHECK_CUDA(cudaEventCreate(&sl.dense_t0));
      MRCNN_CHECK_CUDA(cudaEventCreate(&sl.dense_t1));
    }
    MRCNN_CHECK_CUDA(cudaEventRecord(sl.dense_t0, st));
    RC(mrcnn_mask_bits_expand_device(static_cast<const uint32_t*>(sl.bits), sl.counts, n, (int64_t)npx, D,
                                     static_cast<uint8_t*>(sl.dense), st));
    MRCNN_CHECK_CUDA(cudaMemcpyAsync(e->dense_host, sl.dense, (size_t)n * npx * D, cudaMemcpyDeviceToHost, st));
    MRCNN_CHECK_CUDA(cudaEventRecord(sl.dense_t1, st));
    sl.dense_images = n;
  }
  e->dense_host = nullptr;
  e->dense_n = 0;
  if (async) {
    MRCNN_CHECK_CUDA(cudaEventRecord(sl.copied, e->copy_stream));
    sl.copy_pending = true;
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_wait(mrcnn_engine* e) {
  MRCNN_REQUIRE(e, "engine_wait: null engine");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  if (e->copy_stream) MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->copy_stream));
  for (auto& sl : e->slots) sl.copy_pending = false;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_wait_slot(mrcnn_engine* e, int slot) {
  MRCNN_REQUIRE(e && (slot == 0 || slot == 1), "engine_wait_slot: bad arguments");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  mrcnn_engine::ResultSlot& sl = e->slots[slot];
  if (sl.copy_pending) {
    MRCNN_CHECK_CUDA(cudaEventSynchronize(sl.copied));
    sl.copy_pending = false;
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_next_slot(const mrcnn_engine* e) { return e ? e->cur_slot : -1; }

extern "C" int mrcnn_engine_wait_slot_packed(mrcnn_engine* e, int slot) {
  MRCNN_REQUIRE(e && (slot == 0 || slot == 1), "engine_wait_slot_packed: bad arguments");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  mrcnn_engine::ResultSlot& sl = e->slots[slot];
  if (sl.copy_pending && sl.dense_images > 0) {
    MRCNN_CHECK_CUDA(cudaEventSynchronize(sl.dense_t0));     // boxes / ids / scores / counts / mask bits have arrived
    return MRCNN_OK;
  }
  return mrcnn_engine_wait_slot(e, slot);
}

extern "C" int mrcnn_engine_set_dense_output(mrcnn_engine* e, uint8_t* dense_host, int n_images) {
  MRCNN_REQUIRE(e, "engine_set_dense_output: null engine");
  MRCNN_REQUIRE(n_images >= 0 && (n_images == 0 || dense_host), "engine_set_dense_output: bad arguments");
  MRCNN_REQUIRE(n_images == 0 || e->cfg.detection_max_instances % 4 == 0,
                "engine_set_dense_output: DETECTION_MAX_INSTANCES must be a multiple of 4");
  e->dense_host = n_images ? dense_host : nullptr;
  e->dense_n = n_images;
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_dense_copy_ms(mrcnn_engine* e, int slot, int* n_images, float* ms) {
  MRCNN_REQUIRE(e && (slot == 0 || slot == 1) && n_images && ms, "engine_dense_copy_ms: bad arguments");
  mrcnn_engine::ResultSlot& sl = e->slots[slot];
  *n_images = sl.dense_images;
  *ms = 0.f;
  if (sl.dense_images > 0) {
    MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
    MRCNN_CHECK_CUDA(cudaEventSynchronize(sl.dense_t1));
    MRCNN_CHECK_CUDA(cudaEventElapsedTime(ms, sl.dense_t0, sl.dense_t1));
  }
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_detect_molded(mrcnn_engine* e, const float* molded, int molded_on_host,
                                          const float* metas_host, const int* orig_hw,
                                          const int32_t* windows_host, int32_t* rois_host,
                                          int32_t* class_ids_host, float* scores_host, int32_t* counts_host,
                                          uint32_t* mask_bits_host) {
  MRCNN_REQUIRE(e && e->finalized, "detect_molded: engine not finalized");
  MRCNN_REQUIRE(molded && metas_host && orig_hw && windows_host, "detect_molded: null pointer");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  RC(predict_internal(e, molded, molded_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, metas_host,
                      cudaMemcpyHostToDevice));
  RC(unmold_internal(e, 0, orig_hw, windows_host, 1));
  RC(fetch_internal(e, 0, false, orig_hw, rois_host, class_ids_host, scores_host, counts_host, mask_bits_host));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  return MRCNN_OK;
}

extern "C" int mrcnn_engine_detect_maps(mrcnn_engine* e, const float* maps, int maps_on_host, int map_h, int map_w,
                                        const float* contrasts3, const float* mean_pixel3, int out_h, int out_w,
                                        int top, int left, const float* metas_host, const int32_t* windows_host,
                                        int32_t* rois_host, int32_t* class_ids_host, float* scores_host,
                                        int32_t* counts_host, uint32_t* mask_bits_host, int mask_format, int async) {
  MRCNN_REQUIRE(e && e->finalized, "detect_maps: engine not finalized");
  MRCNN_REQUIRE(mask_format == 0 || mask_format == 1, "detect_maps: mask_format must be 0 (bytes, device only) or 1 (bits)");
  MRCNN_REQUIRE(mask_format == 1 || !mask_bits_host, "detect_maps: host masks are shipped as bits (mask_format 1)");
  MRCNN_REQUIRE(maps && contrasts3 && mean_pixel3 && metas_host && windows_host, "detect_maps: null pointer");
  MRCNN_REQUIRE(map_h > 0 && map_w > 0, "detect_maps: empty maps");
  MRCNN_CHECK_CUDA(cudaSetDevice(e->device));
  const mrcnn_engine_config& c = e->cfg;
  const int B = c.batch_size;
  const size_t npx = (size_t)map_h * map_w;
  RC(ensure_scratch(e, &e->pre_rgb, &e->pre_rgb_bytes, (size_t)B * npx * 3));
  RC(ensure_scratch(e, &e->pre_small, &e->pre_small_bytes, (size_t)B * (12 * 4 + 2 * 4)));
  const float* d_maps = maps;
  int up = -1;
  if (maps_on_host && async) {
    // upload on the side stream into buffer `up`; the main stream only waits for the copy's event.  The buffer is
    // recycled two calls later: the copy first waits until the preprocessing kernels that read it have run.
    up = e->up_next;
    e->up_next ^= 1;
    if (!e->upload_stream) MRCNN_CHECK_CUDA(cudaStreamCreateWithFlags(&e->upload_stream, cudaStreamNonBlocking));
    if (!e->up_done[up]) {
      MRCNN_CHECK_CUDA(cudaEventCreateWithFlags(&e->up_done[up], cudaEventDisableTiming));
      MRCNN_CHECK_CUDA(cudaEventCreateWithFlags(&e->up_consumed[up], cudaEventDisableTiming));
    }
    if (e->up_maps_bytes[up] < (size_t)B * npx * 4) {
      if (e->up_used[up]) MRCNN_CHECK_CUDA(cudaEventSynchronize(e->up_consumed[up]));
      RC(ensure_scratch(e, &e->up_maps[up], &e->up_maps_bytes[up], (size_t)B * npx * 4));
      e->up_used[up] = false;
    }
    if (e->up_used[up]) MRCNN_CHECK_CUDA(cudaStreamWaitEvent(e->upload_stream, e->up_consumed[up], 0));
    MRCNN_CHECK_CUDA(cudaMemcpyAsync(e->up_maps[up], maps, (size_t)B * npx * 4, cudaMemcpyHostToDevice, e->upload_stream));
    MRCNN_CHECK_CUDA(cudaEventRecord(e->up_done[up], e->upload_stream));
    MRCNN_CHECK_CUDA(cudaStreamWaitEvent(e->stream, e->up_done[up], 0));
    d_maps = static_cast<const float*>(e->up_maps[up]);
  } else if (maps_on_host) {
    RC(ensure_scratch(e, &e->pre_maps, &e->pre_maps_bytes, (size_t)B * npx * 4));
    MRCNN_CHECK_CUDA(cudaMemcpyAsync(e->pre_maps, maps, (size_t)B * npx * 4, cudaMemcpyHostToDevice, e->stream));
    d_maps = static_cast<const float*>(e->pre_maps);
  }
  float* d_params = static_cast<float*>(e->pre_small);
  int32_t* d_minmax = reinterpret_cast<int32_t*>(d_params + (size_t)B * 12);
  uint8_t* d_rgb = static_cast<uint8_t*>(e->pre_rgb);
  RC(mrcnn_zscale_params(d_maps, B, map_h, map_w, contrasts3, d_params, e->stream));
  RC(mrcnn_stretch_to_rgb8(d_maps, d_params, B, map_h, map_w, d_rgb, d_minmax, e->stream));
  if (up >= 0) {                     // the uploaded maps have been consumed: the buffer may be overwritten after this point
    MRCNN_CHECK_CUDA(cudaEventRecord(e->up_consumed[up], e->stream));
    e->up_used[up] = true;
  }
  float* d_img = static_cast<float*>(e->tensors["input_image"].ptr);
  RC(mrcnn_resize_pad_mold(d_rgb, d_minmax, B, map_h, map_w, out_h, out_w, c.image_size, top, left, mean_pixel3, d_img, e->stream));
  RC(predict_internal(e, d_img, cudaMemcpyDeviceToDevice, metas_host, cudaMemcpyHostToDevice));
  const int orig_hw[2] = {map_h, map_w};
  const bool any_out = rois_host || class_ids_host || scores_host || counts_host || mask_bits_host;
  const int slot = (async && any_out) ? e->cur_slot : 0;
  RC(unmold_internal(e, slot, orig_hw, windows_host, mask_format));
  if (async && any_out) {
    RC(fetch_internal(e, slot, true, orig_hw, rois_host, class_ids_host, scores_host, counts_host, mask_bits_host));
    e->cur_slot ^= 1;
    return MRCNN_OK;                 // caller collects with mrcnn_engine_wait()
  }
  if (async) return MRCNN_OK;        // device-only and asynchronous: results stay in the unmold_* tensors, no sync
  RC(fetch_internal(e, slot, false, orig_hw, rois_host, class_ids_host, scores_host, counts_host, mask_bits_host));
  MRCNN_CHECK_CUDA(cudaStreamSynchronize(e->stream));
  return MRCNN_OK;
}
