// Engine implementation: weight packing (finalize) and kernel sequencing of DIR's eval forward
// (models/dir.py:513-540 -> backbone :516, init_regressor :517, decoder :518 / :437-483).
#include "engine.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace dirb200 {

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  // static + dynamic above 48 KB needs the opt-in too (the joint-space kernels pair ~32 KB of each): ask for >= 64 KB
  if (bytes < 64 * 1024) bytes = 64 * 1024;
  int dev = 0;
  cudaGetDevice(&dev);
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = done[std::make_pair(func, dev)];
  if (cur >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DIRB200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

#define CK(expr)                                                                 \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
      return DIRB200_E_CUDA;                                                     \
    }                                                                            \
  } while (0)

Engine::~Engine() {
  if (nccl_comm && nccl_destroy) nccl_destroy(nccl_comm);
  if (side) cudaStreamDestroy(side);
  for (int i = 0; i < 2; ++i) {
    if (ev_fork[i]) cudaEventDestroy(ev_fork[i]);
    if (ev_join[i]) cudaEventDestroy(ev_join[i]);
  }
  for (void* p : owned) cudaFree(p);
  for (auto& r : prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
}

// ------------------------------------------------------------------------------------------------ finalize helpers
static const float* kDummy = reinterpret_cast<const float*>(0x100);

const float* Engine::W(const std::string& name, std::vector<int64_t> shape) {
  if (dry) {
    bool seen = false;
    for (const auto& r : required) seen = seen || r == name;
    if (!seen) required.push_back(name);
    return kDummy;
  }
  auto it = raw.find(name);
  if (it == raw.end()) {
    if (err.empty()) err = "missing state_dict key: " + name;
    return nullptr;
  }
  if (it->second.dtype != DIRB200_DTYPE_F32) {
    if (err.empty()) err = "key is not fp32: " + name;
    return nullptr;
  }
  if (!shape.empty() && it->second.shape != shape) {
    if (err.empty()) err = "shape mismatch for " + name;
    return nullptr;
  }
  return reinterpret_cast<const float*>(it->second.p);
}

float* Engine::dalloc(size_t nfloats) {
  if (dry) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, nfloats * sizeof(float)) != cudaSuccess) {
    if (err.empty()) err = "cudaMalloc failed in finalize";
    return nullptr;
  }
  owned.push_back(p);
  return reinterpret_cast<float*>(p);
}

float* Engine::copy_of(const std::string& name) {
  const float* src = W(name);
  if (dry || !src) return nullptr;
  int64_t n = raw[name].numel();
  float* d = dalloc(n);
  if (d) cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, fin_stream);
  return d;
}

// src is (rows, cols) row-major -> returns (cols, rows)
float* Engine::transposed(const std::string& name, int rows, int cols) {
  const float* src = W(name);
  if (dry || !src) return nullptr;
  if (raw[name].numel() != (int64_t)rows * cols) {
    if (err.empty()) err = "unexpected size for " + name;
    return nullptr;
  }
  float* d = dalloc((size_t)rows * cols);
  if (d) launch_transpose2d(src, d, rows, cols, fin_stream);
  return d;
}

// src is a Linear weight (N, K) -> k-pair interleaved K-major [K/2][N][2]
float* Engine::transposed_pairs(const std::string& name, int N, int K) {
  const float* src = W(name);
  if (dry || !src) return nullptr;
  if (raw[name].numel() != (int64_t)N * K) {
    if (err.empty()) err = "unexpected size for " + name;
    return nullptr;
  }
  float* d = dalloc((size_t)N * K);
  if (d) launch_transpose_pairs(src, d, N, K, fin_stream);
  return d;
}

// scale/shift = eval-BN folded with an optional preceding conv bias. Writes n values at offset `off` of
// buffers of `total` (allocated on first use when *scale == nullptr).
void Engine::fold(const std::string& conv_bias, const std::string& bn, float** scale, float** shift, int n, int off,
                  int total) {
  const float* cb = conv_bias.empty() ? nullptr : W(conv_bias);
  const float *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
  if (!bn.empty()) {
    g = W(bn + "weight");
    be = W(bn + "bias");
    mu = W(bn + "running_mean");
    var = W(bn + "running_var");
  }
  if (dry || !err.empty() || n <= 0) return;
  if (total == 0) total = n;
  if (!*scale) {
    *scale = dalloc(total);
    *shift = dalloc(total);
    if (*scale && *shift) {  // channels nobody folds into (zero padding) must produce exact zeros
      cudaMemsetAsync(*scale, 0, (size_t)total * sizeof(float), fin_stream);
      cudaMemsetAsync(*shift, 0, (size_t)total * sizeof(float), fin_stream);
    }
  }
  if (!*scale || !*shift || !err.empty()) return;
  launch_fold_affine(cb, g, be, mu, var, *scale + off, *shift + off, n, fin_stream);
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

ConvLayer Engine::make_conv(const std::string& wname, const std::string& bias, const std::string& bn, int stride,
                            int pad, int relu, int cin_pad, int cout_pad) {
  ConvLayer L;
  L.name = wname;
  const float* w = W(wname);
  L.stride = stride;
  L.pad = pad;
  L.relu = relu;
  if (dry) {
    fold(bias, bn, &L.scale, &L.shift, 0);
    return L;
  }
  if (!w) return L;
  const auto& sh = raw[wname].shape;
  if (sh.size() != 4) {
    if (err.empty()) err = "conv weight must be 4-D: " + wname;
    return L;
  }
  const int cout_src = (int)sh[0], cin_src = (int)sh[1];
  L.Cout = cout_pad > 0 ? cout_pad : cout_src;
  L.Cin = cin_pad > 0 ? cin_pad : cin_src;
  if (L.Cout < cout_src || L.Cin < cin_src) {
    if (err.empty()) err = "channel padding smaller than the tensor: " + wname;
    return L;
  }
  L.kh = (int)sh[2];
  L.kw = (int)sh[3];
  L.K = L.kh * L.kw * L.Cin;
  L.Kpad = round_up(L.K, 16);
  L.w32 = dalloc((size_t)L.Cout * L.Kpad);
  if (bf16()) L.w16 = reinterpret_cast<__nv_bfloat16*>(dalloc(((size_t)L.Cout * L.Kpad + 1) / 2));
  if (L.w32)
    launch_pack_conv_weight(w, L.w32, L.w16, L.Cout, L.Cin, L.kh, L.kw, L.Kpad, fin_stream, cout_src, cin_src);
  fold(bias, bn, &L.scale, &L.shift, cout_src, 0, L.Cout);
  if (bf16() && L.w16 && err.empty()) conv_tc_prepare_weights(L);
  if (!bf16() && err.empty()) prepare_tf32(L);
  return L;
}

// fp32 / tf32 configurations: tf32 hi/lo copies of the packed weights for conv_tf32.cu (tensor-core shaped layers only)
void Engine::prepare_tf32(ConvLayer& L) {
  if (dry || !L.w32 || L.Cin % 32 != 0 || L.Cout % 64 != 0 || L.K != L.Kpad) return;
  if (L.Cin == 2560) return;  // fusion.0 runs in its exact factored form (fusion.cu), never as a dense conv
  float* hi = dalloc((size_t)L.Cout * L.Kpad);
  float* lo = dalloc((size_t)L.Cout * L.Kpad);
  if (hi && lo) conv_tf32_prepare_weights(L, hi, lo, fin_stream);
}

void Engine::make_dual(ConvLayer& F, const ConvLayer& main, const ConvLayer& second) {
  F = ConvLayer();
  if (dry || !bf16() || !err.empty() || !main.w32 || !second.w32) return;
  const size_t K = (size_t)main.K + second.K;
  __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(dalloc(((size_t)main.Cout * K + 1) / 2));
  float* sc = dalloc(main.Cout);
  float* sh = dalloc(main.Cout);
  if (w16 && sc && sh) conv_tc_prepare_dual(F, main, second, w16, sc, sh, fin_stream);
}

PointMlp Engine::make_mlp(const std::string& p, int cin, int cmid, int cout) {
  PointMlp m{};
  m.w1t = transposed(p + "0.weight", cmid, cin);
  float *s = nullptr, *b = nullptr;
  fold(p + "0.bias", p + "1.", &s, &b, cmid);
  m.s1 = s;
  m.b1 = b;
  m.w2t = transposed(p + "3.weight", cout, cmid);
  m.b2 = copy_of(p + "3.bias");
  return m;
}

// cin_pad > 0: the block's input tensor carries zero-padded channels beyond the tensor's own (HRNet-W48: the 96-channel
// c2 lives in 128); conv1 / skip weights and the pre-activation affine are padded with zeros to match
ResidualBlock Engine::make_residual(const std::string& p, int cin_pad) {
  ResidualBlock r;
  r.c1 = make_conv(p + "conv1.conv.weight", p + "conv1.conv.bias", p + "bn2.", 1, 0, 1, cin_pad);
  r.c2 = make_conv(p + "conv2.conv.weight", p + "conv2.conv.bias", p + "bn3.", 1, 1, 1);
  r.c3 = make_conv(p + "conv3.conv.weight", p + "conv3.conv.bias", "", 1, 0, 0);
  if (!dry && bf16() && r.c3.w16) {
    r.c3.tc_bn_cap = 128;
    conv_tc_prepare_weights(r.c3);
  }
  r.skip = make_conv(p + "skip_layer.conv.weight", p + "skip_layer.conv.bias", "", 1, 0, 0, cin_pad);
  r.cin = r.c1.Cin;
  r.cout = r.c3.Cout;
  r.need_skip = r.cin != r.cout;  // hourglass.py:49-52
  if (r.need_skip) make_dual(r.c3skip, r.c3, r.skip);
  int cin_src = r.cin;  // pre-activation BN: padded channels keep scale = shift = 0 (relu(0) = 0)
  if (!dry && raw.count(p + "bn1.weight") && !raw[p + "bn1.weight"].shape.empty()) cin_src = (int)raw[p + "bn1.weight"].shape[0];
  fold("", p + "bn1.", &r.bn1s, &r.bn1b, cin_src, 0, r.cin);  // applied by conv1 (PRE kernels) or concat_preact
  return r;
}

ManoWeights Engine::make_mano(const std::string& p, bool left) {
  ManoWeights m{};
  m.comps = copy_of(p + "th_selected_comps");
  m.mean = copy_of(p + "th_hands_mean");
  m.shapedirs_t = transposed(p + "th_shapedirs", 2334, 10);
  m.posedirs_t = transposed(p + "th_posedirs", 2334, 135);
  m.v_template = copy_of(p + "th_v_template");
  m.jreg = copy_of(p + "th_J_regressor");
  m.skin_w = copy_of(p + "th_weights");
  m.tip2 = left ? 445 : 444;  // manolayer.py:249-252
  return m;
}

void Engine::build_stage(int s, const std::string& p) {
  StageWeights& st = stage[s];
  st.S = s == 0 ? 16 : 32;           // models/dir.py:395,401
  st.distance = s == 0 ? 1.f : 2.f;
  const char* sides[2] = {"left", "right"};
  for (int h = 0; h < 2; ++h) {
    std::string sd = sides[h];
    st.filters[h] = make_mlp(p + "img2joint_" + sd + ".filters.", 256, 128, 128);
    st.pos[h] = make_mlp(p + "pos_emb_" + sd + ".", 3, 128, 128);
    for (int l = 0; l < 4; ++l) {
      std::string q = p + "gcn_" + sd + ".gconv_layers." + std::to_string(l) + ".";
      st.gcn[l].W[h] = copy_of(q + "gconv.W");
      if (bf16()) {
        const float* gw = W(q + "gconv.W", {2, 21, 128, 128});
        float* pk = dalloc(gcn_tc_packed_bytes() / 4);
        if (!dry && gw && pk) launch_pack_gcn_weight_tc(gw, pk, fin_stream);
        st.gcn[l].Wtc[h] = pk;
      }
      const float* e1 = W(q + "gconv.e_1");
      float* A = dalloc(21 * 21);
      if (!dry && e1 && A) launch_gcn_adjacency(e1, A, fin_stream);
      st.gcn[l].A1[h] = A;
      float *sc = nullptr, *sh = nullptr;
      fold(q + "gconv.bias", q + "bn.", &sc, &sh, 128);
      st.gcn[l].scale[h] = sc;
      st.gcn[l].shift[h] = sh;
    }
    st.Wm[h] = copy_of(p + "regressor.mano_" + sd + ".weight");
    st.bm[h] = copy_of(p + "regressor.mano_" + sd + ".bias");
    mano[s + 1][h] = make_mano(p + "regressor.mano_layer_" + sd + ".", h == 0);
  }
  st.Wo = copy_of(p + "regressor.offset.weight");
  st.bo = copy_of(p + "regressor.offset.bias");
  st.gpos = make_mlp(p + "global_pos_emb.", 3, 128, 128);
  st.proj_feat = make_mlp(p + "proj_feat_emb.", 64, 64, 64);
  std::string q = p + "interaction.";
  st.ste.pos = copy_of(q + "spatial_pos_embed");
  for (int l = 0; l < 3; ++l) {  // blocks 1..3 only (mixSTE.py:197)
    std::string bq = q + "STEblocks." + std::to_string(l + 1) + ".";
    auto& B = st.ste.blk[l];
    B.n1w = copy_of(bq + "norm1.weight");
    B.n1b = copy_of(bq + "norm1.bias");
    B.qkv_t = transposed_pairs(bq + "attn.qkv.weight", 384, 128);
    B.qkv_b = copy_of(bq + "attn.qkv.bias");
    B.proj_t = transposed_pairs(bq + "attn.proj.weight", 128, 128);
    B.proj_b = copy_of(bq + "attn.proj.bias");
    B.n2w = copy_of(bq + "norm2.weight");
    B.n2b = copy_of(bq + "norm2.bias");
    B.fc1_t = transposed_pairs(bq + "mlp.fc1.weight", 256, 128);
    B.fc1_b = copy_of(bq + "mlp.fc1.bias");
    B.fc2_t = transposed_pairs(bq + "mlp.fc2.weight", 128, 256);
    B.fc2_b = copy_of(bq + "mlp.fc2.bias");
  }
  st.ste.snw = copy_of(q + "spatial_norm.weight");
  st.ste.snb = copy_of(q + "spatial_norm.bias");
  st.ste.hnw = copy_of(q + "head.0.weight");
  st.ste.hnb = copy_of(q + "head.0.bias");
  st.ste.head_t = transposed_pairs(q + "head.1.weight", 64, 128);
  st.ste.head_b = copy_of(q + "head.1.bias");
  if (bf16()) {  // operand tiles of the tcgen05 kernel, in consumption order
    float* pk = dalloc(ste_tc_packed_bytes() / 4);
    const char* names[4] = {"attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"};
    const std::vector<int64_t> shp[4] = {{384, 128}, {128, 128}, {256, 128}, {128, 256}};
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < 4; ++k) {
        const float* src = W(q + "STEblocks." + std::to_string(l + 1) + "." + names[k], shp[k]);
        if (!dry && src && pk) launch_pack_ste_tc(src, l, k, pk, fin_stream);
      }
    const float* hw = W(q + "head.1.weight", {64, 128});
    if (!dry && hw && pk) launch_pack_ste_tc(hw, 3, 0, pk, fin_stream);
    st.ste_packed = pk;
  }
  st.fusion0 = make_conv(p + "fusion.0.weight", p + "fusion.0.bias", p + "fusion.1.", 1, 1, 1);
  st.fusion3 = make_conv(p + "fusion.3.weight", p + "fusion.3.bias", "", 1, 0, 0);
  {
    const float* fw = W(p + "fusion.0.weight", {256, 2560, 3, 3});
    float* wp = dalloc((size_t)40 * 64 * 9 * 256);
    if (!dry && fw && wp) launch_pack_fusion_weight(fw, wp, fin_stream);
    st.fus_wp = wp;
    if (bf16()) {
      float* wt = dalloc(bone_coef_tc_packed_bytes() / 4);
      if (!dry && fw && wt) launch_pack_fusion_weight_tc(fw, wt, fin_stream);
      st.fus_wp_tc = wt;
    }
  }
}

// Two convs sharing their input, concatenated along Cout (attention_left|right, seg|dense first convs).
static ConvLayer concat_convs(Engine& e, const std::string& w0, const std::string& b0, const std::string& bn0,
                              const std::string& w1, const std::string& b1, const std::string& bn1, int pad, int relu) {
  ConvLayer L;
  L.name = w0;
  const float* a = e.W(w0);
  const float* b = e.W(w1);
  L.stride = 1;
  L.pad = pad;
  L.relu = relu;
  if (e.dry) {
    e.fold(b0, bn0, &L.scale, &L.shift, 0);
    e.fold(b1, bn1, &L.scale, &L.shift, 0);
    return L;
  }
  if (!a || !b) return L;
  const auto& sh = e.raw[w0].shape;
  int half = (int)sh[0];
  L.Cout = 2 * half;
  L.Cin = (int)sh[1];
  L.kh = (int)sh[2];
  L.kw = (int)sh[3];
  L.K = L.kh * L.kw * L.Cin;
  L.Kpad = round_up(L.K, 16);
  L.w32 = e.dalloc((size_t)L.Cout * L.Kpad);
  if (e.bf16()) L.w16 = reinterpret_cast<__nv_bfloat16*>(e.dalloc(((size_t)L.Cout * L.Kpad + 1) / 2));
  if (!L.w32) return L;
  size_t hoff = (size_t)half * L.Kpad;
  launch_pack_conv_weight(a, L.w32, L.w16, half, L.Cin, L.kh, L.kw, L.Kpad, e.fin_stream);
  launch_pack_conv_weight(b, L.w32 + hoff, L.w16 ? L.w16 + hoff : nullptr, half, L.Cin, L.kh, L.kw, L.Kpad,
                          e.fin_stream);
  e.fold(b0, bn0, &L.scale, &L.shift, half, 0, L.Cout);
  e.fold(b1, bn1, &L.scale, &L.shift, half, half, L.Cout);
  if (e.bf16() && L.w16 && e.err.empty()) conv_tc_prepare_weights(L);
  if (!e.bf16() && e.err.empty()) e.prepare_tf32(L);
  return L;
}

int Engine::build(cudaStream_t st) {
  fin_stream = st;
  required.clear();
  if (!dry && !side && !no_overlap) {
    if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) side = nullptr;
    for (int i = 0; i < 2 && side; ++i) {
      cudaEventCreateWithFlags(&ev_fork[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming);
    }
  }
  if (hrnet()) {
    build_hrnet(cfg.backbone);
  } else {
    // ---- backbone (models/backbone/resnet.py)
    stem = make_conv("backbone.conv1.weight", "", "backbone.bn1.", 2, 3, 1);
    if (!dry && bf16() && err.empty()) {
      __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(dalloc(64 * 224 / 2));
      if (wp) conv_tc_prepare_stem(stem, W("backbone.conv1.weight"), wp, st);
      float* sp = dalloc(stem_pool_weight_bytes() / 4);
      if (sp) launch_pack_stem_pool_weight(W("backbone.conv1.weight"), sp, st);
      stem_pool_w = sp;
    }
    if (!dry && !bf16() && err.empty()) {
      float* buf = dalloc(conv_tf32_stem_weight_floats());
      if (buf) conv_tf32_prepare_stem(stem, W("backbone.conv1.weight"), buf, st);
    }
    const int nblocks[4] = {3, 4, 6, 3};
    for (int l = 0; l < 4; ++l) {
      layers[l].clear();
      for (int b = 0; b < nblocks[l]; ++b) {
        std::string p = "backbone.layer" + std::to_string(l + 1) + "." + std::to_string(b) + ".";
        Bottleneck bk;
        int stride = (b == 0 && l > 0) ? 2 : 1;
        bk.c1 = make_conv(p + "conv1.weight", "", p + "bn1.", 1, 0, 1);
        bk.c2 = make_conv(p + "conv2.weight", "", p + "bn2.", stride, 1, 1);
        bk.c3 = make_conv(p + "conv3.weight", "", p + "bn3.", 1, 0, 1);
        if (!dry && bf16() && bk.c3.w16) {  // residual-adding layer: smaller n-tile leaves room for the residual ring
          bk.c3.tc_bn_cap = 128;  // (256 measured: -4 % images/s)
          if (const char* e = getenv("DIRB200_C3_BN")) {  // experiment switch: 64 / 128 / 256 only
            const int v = atoi(e);
            if (v == 64 || v == 128 || v == 256) bk.c3.tc_bn_cap = v;
          }
          conv_tc_prepare_weights(bk.c3);
        }
        bk.has_ds = (b == 0);
        if (bk.has_ds) {
          bk.ds = make_conv(p + "downsample.0.weight", "", p + "downsample.1.", stride, 0, 0);
          make_dual(bk.c3ds, bk.c3, bk.ds);
        }
        layers[l].push_back(bk);
      }
    }
  }
  // ---- init regressor (models/dir.py:218-305)
  const std::string ir = "init_regressor.";
  attn_conv = concat_convs(*this, ir + "attention_left.0.weight", ir + "attention_left.0.bias", ir + "attention_left.1.",
                           ir + "attention_right.0.weight", ir + "attention_right.0.bias", ir + "attention_right.1.", 1,
                           1);
  {
    const float* wl = W(ir + "attention_left.3.weight");
    const float* wr = W(ir + "attention_right.3.weight");
    const float* bl = W(ir + "attention_left.3.bias");
    const float* br = W(ir + "attention_right.3.bias");
    if (!dry && wl && wr && bl && br) {
      const int half = c4ch / 2;  // feat_dim // 2 (models/dir.py:227-241)
      attn_w = dalloc(2 * half);
      attn_b = dalloc(2);
      if (attn_w && attn_b) {
        cudaMemcpyAsync(attn_w, wl, (size_t)half * 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(attn_w + half, wr, (size_t)half * 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(attn_b, bl, 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(attn_b + 1, br, 4, cudaMemcpyDeviceToDevice, st);
      }
    }
  }
  init_Wm[0] = copy_of(ir + "mano_left.weight");
  init_bm[0] = copy_of(ir + "mano_left.bias");
  init_Wm[1] = copy_of(ir + "mano_right.weight");
  init_bm[1] = copy_of(ir + "mano_right.bias");
  init_Wo = copy_of(ir + "offset.weight");
  init_bo = copy_of(ir + "offset.bias");
  mano[0][0] = make_mano(ir + "mano_layer_left.", true);
  mano[0][1] = make_mano(ir + "mano_layer_right.", false);
  // ---- decoder (models/dir.py:389-483)
  for (const char* n : {"skip_layer4", "fusion_layer4", "enhance_layer4", "skip_layer3", "fusion_layer3",
                        "enhance_layer3"}) {
    std::string p = std::string("decoder.") + n + ".";
    // the only decoder input with padded channels: c2 of HRNet-W48 (96 -> 128), consumed by skip_layer3
    res[p] = make_residual(p, (hrnet() && std::string(n) == "skip_layer3" && c2ch % 64 == 0) ? c2ch : 0);
  }
  build_stage(0, "decoder.projecter_4.");
  build_stage(1, "decoder.projecter_3.");
  conv_final0 = make_conv("decoder.conv_final.0.weight", "", "decoder.conv_final.1.", 1, 1, 1);
  conv_final3 = make_conv("decoder.conv_final.3.weight", "decoder.conv_final.3.bias", "", 1, 0, 0);
  segdense0 = concat_convs(*this, "decoder.seg.0.weight", "decoder.seg.0.bias", "decoder.seg.1.",
                           "decoder.dense.0.weight", "decoder.dense.0.bias", "decoder.dense.1.", 1, 1);
  seg3_w = copy_of("decoder.seg.3.weight");
  seg3_b = copy_of("decoder.seg.3.bias");
  dense3_w = copy_of("decoder.dense.3.weight");
  dense3_b = copy_of("decoder.dense.3.bias");
  if (dry) return DIRB200_OK;
  if (!err.empty()) return err.rfind("missing", 0) == 0 ? DIRB200_E_MISSING : DIRB200_E_INVALID;
  CK(cudaGetLastError());
  return DIRB200_OK;
}

// ------------------------------------------------------------------------------------------------ HRNet extension
// HRNet-W32 (Sun et al., CVPR 2019; module / state_dict names of the authors' pose_hrnet.py) as a flat program over
// numbered buffers. Not in the reference: the oracle is the self-authored oracle/hrnet_oracle.py.
// Channel counts that are not multiples of 64 (the 32-channel branch) are zero-padded to 64 so that every conv runs on
// the tensor-core kernels; padded output channels get scale = shift = 0 and stay exactly zero through ReLU/residual/fuse.
void Engine::build_hrnet(int width) {
  hr_convs.clear();
  hr_ops.clear();
  hr_bufs.clear();
  hr_convs.reserve(512);  // the profiling hook keeps pointers into this vector: it must never reallocate
  const int C[4] = {width, 2 * width, 4 * width, 8 * width};
  auto pad64 = [](int c) { return (c + 63) / 64 * 64; };
  const std::string p = "backbone.";
  auto newbuf = [&](int H, int W_, int Cc) {
    hr_bufs.push_back(HrBuf{H, W_, Cc});
    return (int)hr_bufs.size() - 1;
  };
  // one conv + its BN: returns the id of the output buffer
  auto add_conv = [&](const std::string& w, const std::string& bn, int cout, int k, int stride, int relu, int in, int res) {
    const int padn = (k - 1) / 2;
    const int H = in < 0 ? 256 : hr_bufs[in].H, W_ = in < 0 ? 256 : hr_bufs[in].W;
    ConvLayer L = make_conv(p + w + ".weight", "", p + bn + ".", stride, padn, relu, in < 0 ? 0 : hr_bufs[in].C, pad64(cout));
    if (!dry && bf16() && in < 0 && err.empty() && L.w32) {  // 3 -> 64, 3x3 / s2 on the image: the TMA stem variant of conv_tc.cu
      __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(dalloc(64 * 3 * 32 / 2));
      if (wp) conv_tc_prepare_stem(L, W(p + w + ".weight"), wp, fin_stream);
    }
    if (!dry && bf16() && L.w16 && res >= 0) {
      L.tc_bn_cap = 128;  // residual-adding layer (see the ResNet bottlenecks)
      conv_tc_prepare_weights(L);
    }
    hr_convs.push_back(L);
    HrOp op{};
    op.kind = 0;
    op.layer = (int)hr_convs.size() - 1;
    op.in = in;
    op.res = res;
    op.out = newbuf((H + 2 * padn - k) / stride + 1, (W_ + 2 * padn - k) / stride + 1, pad64(cout));
    hr_ops.push_back(op);
    return op.out;
  };
  int x = add_conv("conv1", "bn1", 64, 3, 2, 1, -1, -1);
  x = add_conv("conv2", "bn2", 64, 3, 2, 1, x, -1);
  for (int b = 0; b < 4; ++b) {  // layer1: Bottleneck(64 -> 256) x4
    const std::string q = "layer1." + std::to_string(b) + ".";
    int t = add_conv(q + "conv1", q + "bn1", 64, 1, 1, 1, x, -1);
    t = add_conv(q + "conv2", q + "bn2", 64, 3, 1, 1, t, -1);
    const int idn = b == 0 ? add_conv(q + "downsample.0", q + "downsample.1", 256, 1, 1, 0, x, -1) : x;
    x = add_conv(q + "conv3", q + "bn3", 256, 1, 1, 1, t, idn);
  }
  std::vector<int> xs;
  xs.push_back(add_conv("transition1.0.0", "transition1.0.1", C[0], 3, 1, 1, x, -1));
  xs.push_back(add_conv("transition1.1.0.0", "transition1.1.0.1", C[1], 3, 2, 1, x, -1));
  const int stages[3][3] = {{2, 1, 2}, {3, 4, 3}, {4, 3, 4}};  // stage, modules, branches
  for (const auto& sg : stages) {
    const int stage = sg[0], nmod = sg[1], nbr = sg[2];
    if (stage > 2) {  // the new branch is made from the last branch of the previous stage
      const std::string t = "transition" + std::to_string(stage - 1) + "." + std::to_string(nbr - 1) + ".0.";
      xs.push_back(add_conv(t + "0", t + "1", C[nbr - 1], 3, 2, 1, xs.back(), -1));
    }
    for (int m = 0; m < nmod; ++m) {
      const std::string q = "stage" + std::to_string(stage) + "." + std::to_string(m) + ".";
      for (int br = 0; br < nbr; ++br)
        for (int k = 0; k < 4; ++k) {  // BasicBlock: conv-bn-relu, conv-bn, + identity, relu
          const std::string bq = q + "branches." + std::to_string(br) + "." + std::to_string(k) + ".";
          const int t = add_conv(bq + "conv1", bq + "bn1", C[br], 3, 1, 1, xs[br], -1);
          xs[br] = add_conv(bq + "conv2", bq + "bn2", C[br], 3, 1, 1, t, xs[br]);
        }
      std::vector<int> fused;
      for (int i = 0; i < nbr; ++i) {
        HrOp f{};
        f.kind = 1;
        for (int j = 0; j < nbr; ++j) {
          const std::string fq = q + "fuse_layers." + std::to_string(i) + "." + std::to_string(j) + ".";
          int t = xs[j], sh = 0;
          if (j > i) {  // 1x1 conv + BN at the low resolution; the nearest upsampling happens inside the fuse kernel
            t = add_conv(fq + "0", fq + "1", C[i], 1, 1, 0, xs[j], -1);
            sh = j - i;
          } else if (j < i) {
            for (int k = 0; k < i - j; ++k) {
              const bool last = k == i - j - 1;
              t = add_conv(fq + std::to_string(k) + ".0", fq + std::to_string(k) + ".1", last ? C[i] : C[j], 3, 2,
                           last ? 0 : 1, t, -1);
            }
          }
          f.term[f.nterm] = t;
          f.shift[f.nterm] = sh;
          ++f.nterm;
        }
        f.out = newbuf(hr_bufs[xs[i]].H, hr_bufs[xs[i]].W, hr_bufs[xs[i]].C);
        hr_ops.push_back(f);
        fused.push_back(f.out);
      }
      xs = fused;
    }
  }
  for (int i = 0; i < 4; ++i) hr_out[i] = xs[i];
  c1ch = pad64(C[0]);
  c2ch = pad64(C[1]);
  c3ch = pad64(C[2]);
  c4ch = pad64(C[3]);
  // c2 may be padded (W48: 96 -> 128; skip_layer3 is built with matching zero weights); c3 feeds skip_layer4 AND sits
  // in the middle of fusion_layer4's concat, c4 is InitRegressor's feat_dim: those two must be exact
  if (!dry && err.empty() && (c3ch != C[2] || c4ch != C[3])) err = "HRNet width must make c3 and c4 multiples of 64";
}

template <typename T>
int Engine::run_backbone_hrnet(const float* img, int B, Arena& ar, T** c1, T** c2, T** c3, T** c4, cudaStream_t st) {
  std::vector<T*> ptr(hr_bufs.size());
  for (size_t i = 0; i < hr_bufs.size(); ++i)
    ptr[i] = reinterpret_cast<T*>(ar.alloc((size_t)B * hr_bufs[i].H * hr_bufs[i].W * hr_bufs[i].C * sizeof(T)));
  float* u8_scratch = reinterpret_cast<float*>(ar.alloc((size_t)B * 3 * 256 * 256 * sizeof(float)));
  // first stem conv on the tensor cores (bf16): image -> padded NHWC4 operand (uint8 preprocessing fused in) -> TMA stem
  const bool tc_stem = sizeof(T) == 2 && !disable_tc && !hr_ops.empty() && hr_ops[0].kind == 0 && hr_ops[0].in < 0 &&
                       hr_convs[hr_ops[0].layer].tc_stem && conv_tc_supported(hr_convs[hr_ops[0].layer], B, 256, 256);
  __nv_bfloat16* stem_scratch =
      tc_stem ? reinterpret_cast<__nv_bfloat16*>(ar.alloc(conv_tc_stem_scratch_bytes(B, 256, 256))) : nullptr;
  if (!ar.base) return DIRB200_OK;
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  if (img_u8 && !tc_stem) {  // input pipeline (apps/eval.py:56-61) as its own pass
    launch_preprocess_u8(img_u8, u8_scratch, B, 256, 256, st);
    ++launches;
    img = u8_scratch;
  }
  for (const HrOp& op : hr_ops) {
    if (op.kind == 0) {
      const ConvLayer& L = hr_convs[op.layer];
      if (op.in < 0 && tc_stem) {
        int rc = launch_conv_tc_stem(L, img, img_u8, stem_scratch, reinterpret_cast<__nv_bfloat16*>(ptr[op.out]), B, 256, 256,
                                     st);
        if (rc && !sticky_rc) {
          sticky_rc = rc;
          err = "tcgen05 stem launch failed";
        }
        launches += 2;
        tc_launches += 1;
      } else if (op.in < 0)
        conv<T>(L, reinterpret_cast<const T*>(img), ptr[op.out], nullptr, B, 256, 256, st, /*in_nchw=*/true);
      else
        conv<T>(L, ptr[op.in], ptr[op.out], op.res >= 0 ? ptr[op.res] : nullptr, B, hr_bufs[op.in].H, hr_bufs[op.in].W, st);
    } else {
      const T* terms[4];
      for (int k = 0; k < op.nterm; ++k) terms[k] = ptr[op.term[k]];
      const HrBuf& o = hr_bufs[op.out];
      launch_fuse_sum_relu<T>(terms, op.shift, op.nterm, ptr[op.out], B, o.H, o.W, o.C, st);
      ++launches;
    }
  }
  if (c1) *c1 = ptr[hr_out[0]];
  *c2 = ptr[hr_out[1]];
  *c3 = ptr[hr_out[2]];
  *c4 = ptr[hr_out[3]];
  return DIRB200_OK;
}
template int Engine::run_backbone_hrnet<float>(const float*, int, Arena&, float**, float**, float**, float**, cudaStream_t);
template int Engine::run_backbone_hrnet<__nv_bfloat16>(const float*, int, Arena&, __nv_bfloat16**, __nv_bfloat16**,
                                                       __nv_bfloat16**, __nv_bfloat16**, cudaStream_t);

const ConvLayer* Engine::find_conv(const std::string& k) const {
  std::vector<const ConvLayer*> all = {&stem, &attn_conv, &conv_final0, &conv_final3, &segdense0};
  for (const auto& c : hr_convs) all.push_back(&c);
  for (int l = 0; l < 4; ++l)
    for (const auto& b : layers[l]) {
      all.push_back(&b.c1);
      all.push_back(&b.c2);
      all.push_back(&b.c3);
      if (b.has_ds) all.push_back(&b.ds);
    }
  for (const auto& kv : res) {
    all.push_back(&kv.second.c1);
    all.push_back(&kv.second.c2);
    all.push_back(&kv.second.c3);
    all.push_back(&kv.second.skip);
  }
  for (int s = 0; s < 2; ++s) {
    all.push_back(&stage[s].fusion0);
    all.push_back(&stage[s].fusion3);
  }
  for (const ConvLayer* c : all)
    if (c->name == k) return c;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ forward pieces
template <typename T>
void Engine::conv(const ConvLayer& L, const T* x, T* y, const T* resid, int B, int H, int W_, cudaStream_t st,
                  bool in_nchw) {
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W_ + 2 * L.pad - L.kw) / L.stride + 1;
  ++launches;
  Engine::ProfRec* pr = nullptr;
  if (prof_on && L.name.compare(0, prof_prefix.size(), prof_prefix) == 0) {
    if (prof_used == prof.size()) {
      ProfRec r;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      r.flops = 0;
      prof.push_back(r);
    }
    pr = &prof[prof_used++];
    pr->flops = 2.0 * B * Ho * Wo * (double)L.Cout * L.K;
    pr->bytes = sizeof(T) * ((double)B * H * W_ * L.Cin + (double)B * Ho * Wo * L.Cout * (resid ? 2 : 1) +
                             (double)L.Cout * L.K);
    pr->layer = &L;
    pr->tc = (sizeof(T) == 2 && !in_nchw && !disable_tc && conv_tc_supported(L, B, H, W_)) ? 1 : 0;
    if (sizeof(T) == 4 && !in_nchw && !fp32_simt && conv_tf32_supported(L, B, H, W_)) pr->tc = 1;
    cudaEventRecord(pr->a, st);
  }
  if (sizeof(T) == 4 && !in_nchw && !fp32_simt && conv_tf32_supported(L, B, H, W_)) {
    int rc = launch_conv_tf32(L, reinterpret_cast<const float*>(x), reinterpret_cast<float*>(y),
                              reinterpret_cast<const float*>(resid), B, H, W_, tf32_nsplit(), st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "tcgen05 tf32 conv launch failed for " + L.name;
    }
    ++tc_launches;
    if (pr) cudaEventRecord(pr->b, st);
    return;
  }
  if (sizeof(T) == 2 && !in_nchw && !disable_tc && !no_halo && conv_tc_supported(L, B, H, W_) &&
      conv_halo_supported(L, B, H, W_)) {
    int rc = launch_conv_halo(L, reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                              reinterpret_cast<const __nv_bfloat16*>(resid), B, H, W_, st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "halo conv launch failed for " + L.name;
    }
    ++tc_launches;
    if (pr) cudaEventRecord(pr->b, st);
    return;
  }
  if (sizeof(T) == 2 && !in_nchw && !disable_tc && conv_tc_supported(L, B, H, W_)) {
    int rc = launch_conv_tc(L, reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                            reinterpret_cast<const __nv_bfloat16*>(resid), B, H, W_, st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "tcgen05 conv launch failed for " + L.name;
    }
    ++tc_launches;
    if (pr) cudaEventRecord(pr->b, st);
    return;
  }
  ConvArgs a;
  a.x = x;
  a.w32 = L.w32;
  a.scale = L.scale;
  a.shift = L.shift;
  a.res = resid;
  a.y = y;
  a.B = B; a.H = H; a.W = W_; a.Cin = L.Cin; a.Ho = Ho; a.Wo = Wo; a.Cout = L.Cout;
  a.kh = L.kh; a.kw = L.kw; a.stride = L.stride; a.pad = L.pad; a.K = L.K; a.Kpad = L.Kpad;
  a.relu = L.relu;
  a.in_nchw = in_nchw ? 1 : 0;
  launch_conv_simt<T>(a, st);
  if (pr) cudaEventRecord(pr->b, st);
}

// Bottleneck tail (conv3 + identity, or the conv3 + downsample pair) and conv1 of the next block in one kernel; returns
// false if the shapes are not covered (then the caller runs them separately). t1_next receives conv1's output.
template <typename T>
bool Engine::conv_b2b(const Bottleneck& bk, const Bottleneck& nx, const T* t2, const T* x, T* out, T* t1_next, int B,
                      int Ho, int Wo, cudaStream_t st) {
  if (sizeof(T) != 2 || disable_tc || no_b2b) return false;
  const bool pair = bk.has_ds;
  if (pair && (disable_pair_fusion || bk.c3ds.wmap_bn == 0 || bk.ds.stride != 1)) return false;
  const ConvLayer& first = pair ? bk.c3ds : bk.c3;
  const int Ka = bk.c3.Cin, Kb = pair ? bk.ds.Cin : 0, M = B * Ho * Wo;
  if (!conv_b2b_supported(first, Ka, Kb, nx.c1, M, !pair)) return false;
  launches += 1;
  tc_launches += 1;
  Engine::ProfRec* pr = nullptr;
  if (prof_on && first.name.compare(0, prof_prefix.size(), prof_prefix) == 0) {
    if (prof_used == prof.size()) {
      ProfRec r;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      prof.push_back(r);
    }
    pr = &prof[prof_used++];
    pr->flops = 2.0 * M * ((double)first.Cout * first.K + (double)nx.c1.Cout * nx.c1.K);
    pr->bytes = 2.0 * ((double)M * (first.K + (pair ? 0 : first.Cout) + first.Cout + nx.c1.Cout) +
                       (double)first.Cout * first.K + (double)nx.c1.Cout * nx.c1.K);
    pr->layer = &first;
    pr->tc = 1;
    cudaEventRecord(pr->a, st);
  }
  int rc = launch_conv_b2b(first, reinterpret_cast<const __nv_bfloat16*>(t2), Ka,
                           pair ? reinterpret_cast<const __nv_bfloat16*>(x) : nullptr, Kb,
                           pair ? nullptr : reinterpret_cast<const __nv_bfloat16*>(x), nx.c1,
                           reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<__nv_bfloat16*>(t1_next), M, st);
  if (pr) cudaEventRecord(pr->b, st);
  if (rc && !sticky_rc) {
    sticky_rc = rc;
    err = "back-to-back conv launch failed for " + first.name;
  }
  return true;
}

// conv3 + (downsample | skip) as one K-concatenated tensor-core GEMM; returns false if the pair must run separately
template <typename T>
bool Engine::conv_pair(const ConvLayer& F, const ConvLayer& main, const ConvLayer& second, const T* x1, const T* x2,
                       T* y, int B, int Ho, int Wo, cudaStream_t st, const T* x2b, int C2b) {
  if (sizeof(T) != 2 || disable_tc || disable_pair_fusion || F.wmap_bn == 0) return false;
  ConvLayer probe = F;  // same geometry checks as a 1x1 stride-1 conv on the output grid
  probe.kh = probe.kw = 1;
  probe.stride = 1;
  probe.pad = 0;
  if (!conv_tc_supported(probe, B, Ho, Wo)) return false;
  ++launches;
  ++tc_launches;
  Engine::ProfRec* pr = nullptr;
  if (prof_on && F.name.compare(0, prof_prefix.size(), prof_prefix) == 0) {
    if (prof_used == prof.size()) {
      ProfRec r;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      prof.push_back(r);
    }
    pr = &prof[prof_used++];
    pr->flops = 2.0 * B * Ho * Wo * (double)F.Cout * F.K;
    pr->bytes = 2.0 * ((double)B * Ho * Wo * main.Cin + (double)B * Ho * Wo * second.stride * second.stride * second.Cin +
                       (double)B * Ho * Wo * F.Cout + (double)F.Cout * F.K);
    pr->layer = &F;
    pr->tc = 1;
    cudaEventRecord(pr->a, st);
  }
  int rc = launch_conv_tc_dual(F, reinterpret_cast<const __nv_bfloat16*>(x1), main.Cin,
                               reinterpret_cast<const __nv_bfloat16*>(x2), second.Cin, second.stride,
                               reinterpret_cast<__nv_bfloat16*>(y), B, Ho, Wo, st,
                               reinterpret_cast<const __nv_bfloat16*>(x2b), C2b);
  if (pr) cudaEventRecord(pr->b, st);
  if (rc && !sticky_rc) {
    sticky_rc = rc;
    err = "tcgen05 paired conv launch failed for " + F.name;
  }
  return true;
}

template <typename T>
static T* aalloc(Arena& ar, int64_t n) {
  return reinterpret_cast<T*>(ar.alloc((size_t)n * sizeof(T)));
}

template <typename T>
int Engine::run_backbone(const float* img, int B, int H, int W_, Arena& ar, T** c1, T** c2, T** c3, T** c4,
                         cudaStream_t st) {
  const int H2 = H / 2, W2 = W_ / 2, H4 = H / 4, W4 = W_ / 4;
  const bool tc_stem = sizeof(T) == 2 && !disable_tc && conv_tc_supported(stem, B, H, W_);
  __nv_bfloat16* stem_scratch =
      tc_stem ? reinterpret_cast<__nv_bfloat16*>(ar.alloc(conv_tc_stem_scratch_bytes(B, H, W_))) : nullptr;
  float* u8_scratch = tc_stem ? nullptr : aalloc<float>(ar, (int64_t)B * 3 * H * W_);  // fp32 image from uint8 frames
  const bool stem32 = sizeof(T) == 4 && !fp32_simt && conv_tf32_stem_supported(stem, H, W_);
  float* stem32_scratch = stem32 ? reinterpret_cast<float*>(ar.alloc(conv_tf32_stem_scratch_bytes(B, H, W_))) : nullptr;
  T* stem_out = aalloc<T>(ar, (int64_t)B * H2 * W2 * 64);
  T* pool_out = aalloc<T>(ar, (int64_t)B * H4 * W4 * 64);
  const int64_t rsz = (int64_t)B * H4 * W4 * 256;
  T* R[5];
  for (auto& r : R) r = aalloc<T>(ar, rsz);
  T* keep[4];
  keep[0] = nullptr;  // c1 lives in a rotating buffer unless requested
  if (c1) keep[0] = aalloc<T>(ar, rsz);
  keep[1] = aalloc<T>(ar, (int64_t)B * (H / 8) * (W_ / 8) * 512);
  keep[2] = aalloc<T>(ar, (int64_t)B * (H / 16) * (W_ / 16) * 1024);
  keep[3] = aalloc<T>(ar, (int64_t)B * (H / 32) * (W_ / 32) * 2048);
  if (!ar.base) return DIRB200_OK;
  if (ar.overflow) return DIRB200_E_WORKSPACE;

  if (img_u8 && !tc_stem) {  // input pipeline (apps/eval.py:56-61) as its own pass on the CUDA-core path
    launch_preprocess_u8(img_u8, u8_scratch, B, H, W_, st);
    ++launches;
    img = u8_scratch;
  }
  const bool fused_stem = tc_stem && stem_pool_w && !stem_split && H == 256 && W_ == 256;
  if (fused_stem) {  // conv1 + bn1 + ReLU + maxpool in one kernel: the conv map never reaches HBM
    launch_stem_pack(img, img_u8, stem_scratch, B, H, W_, st);
    int rc = launch_stem_pool(stem_scratch, stem_pool_w, stem.scale, stem.shift, reinterpret_cast<__nv_bfloat16*>(pool_out),
                              B, st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "stem_pool launch failed";
    }
    launches += 2;
    tc_launches += 1;
  } else if (tc_stem) {
    int rc = launch_conv_tc_stem(stem, img, img_u8, stem_scratch, reinterpret_cast<__nv_bfloat16*>(stem_out), B, H, W_,
                                 st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "tcgen05 stem launch failed";
    }
    launches += 2;
    tc_launches += 1;
  } else if (stem32) {  // fp32 / tf32 configurations: the stem on conv_tf32.cu's overlapping-view variant
    int rc = launch_conv_tf32_stem(stem, img, stem32_scratch, reinterpret_cast<float*>(stem_out), B, H, W_, tf32_nsplit(),
                                   st);
    if (rc && !sticky_rc) {
      sticky_rc = rc;
      err = "tcgen05 tf32 stem launch failed";
    }
    launches += 2;
    tc_launches += 1;
  } else {
    conv<T>(stem, reinterpret_cast<const T*>(img), stem_out, nullptr, B, H, W_, st, /*in_nchw=*/true);
  }
  if (!fused_stem) {
    launch_maxpool3x3s2<T>(stem_out, pool_out, B, H2, W2, 64, st);
    ++launches;
  }
  const T* x = pool_out;
  int xi = -1;  // index of the rotating buffer holding x (-1: none)
  int t1i = -1;  // index of the rotating buffer that already holds conv1's output of the coming block (-1: none)
  int h = H4, w = W4;
  for (int l = 0; l < 4; ++l) {
    for (size_t b = 0; b < layers[l].size(); ++b) {
      const Bottleneck& bk = layers[l][b];
      int idx[4], n = 0;
      if (t1i >= 0) idx[n++] = t1i;
      for (int i = 0; i < 5 && n < 4; ++i)
        if (i != xi && i != t1i) idx[n++] = i;
      T *t1 = R[idx[0]], *t2 = R[idx[1]], *dsb = R[idx[2]], *out = R[idx[3]];
      const bool last = b + 1 == layers[l].size();
      if (last && keep[l]) out = keep[l];
      const int ho = h / bk.c2.stride, wo = w / bk.c2.stride;
      if (t1i < 0) conv<T>(bk.c1, x, t1, nullptr, B, h, w, st);  // else: written by the previous block's tail kernel
      t1i = -1;
      conv<T>(bk.c2, t1, t2, nullptr, B, h, w, st);
      // tail of this block + conv1 of the next one as back-to-back GEMMs (conv_b2b.cu): `out` is not re-read
      const Bottleneck* nx = !last ? &layers[l][b + 1] : (l + 1 < 4 ? &layers[l + 1][0] : nullptr);
      if (nx && conv_b2b<T>(bk, *nx, t2, x, out, dsb, B, ho, wo, st)) {
        t1i = idx[2];
      } else if (!(bk.has_ds && conv_pair<T>(bk.c3ds, bk.c3, bk.ds, t2, x, out, B, ho, wo, st))) {
        const T* identity = x;
        if (bk.has_ds) {
          conv<T>(bk.ds, x, dsb, nullptr, B, h, w, st);
          identity = dsb;
        }
        conv<T>(bk.c3, t2, out, identity, B, ho, wo, st);
      }
      x = out;
      xi = (last && keep[l]) ? -1 : idx[3];
      h = ho;
      w = wo;
    }
  }
  if (c1) *c1 = keep[0];
  *c2 = keep[1];
  *c3 = keep[2];
  *c4 = keep[3];
  return DIRB200_OK;
}

template <typename T>
bool Engine::virtual_concat_ok(const ResidualBlock& r, int B, int H, int W_) const {
  if (sizeof(T) != 2 || disable_tc || disable_pair_fusion || !r.need_skip || r.c3skip.wmap_bn == 0) return false;
  ConvLayer probe = r.c3skip;
  probe.kh = probe.kw = 1;
  probe.stride = 1;
  probe.pad = 0;
  return conv_tc_supported(probe, B, H, W_);
}

// true if conv1 of `r` can apply the block's pre-activation (bn1 + ReLU, hourglass.py:60-61) to its own A tiles, reading
// the raw input from one source (c2 = 0) or from the two sources of a channel concat (c2 = channels of the second)
template <typename T>
bool Engine::preact_fold_ok(const ResidualBlock& r, int B, int H, int W_, int c2) const {
  return sizeof(T) == 2 && !disable_tc && !no_preact_fold && conv_tc_pre_supported(r.c1, B, H, W_, r.c1.Cin - c2, c2);
}

// conv1 of a Residual with the pre-activation folded in (PRE variant of conv_tc_kernel)
void Engine::conv_pre(const ResidualBlock& r, const __nv_bfloat16* x1, const __nv_bfloat16* x2, int c2, __nv_bfloat16* y,
                      int B, int H, int W_, cudaStream_t st) {
  const ConvLayer& L = r.c1;
  ++launches;
  ++tc_launches;
  Engine::ProfRec* pr = nullptr;
  if (prof_on && L.name.compare(0, prof_prefix.size(), prof_prefix) == 0) {
    if (prof_used == prof.size()) {
      ProfRec rec;
      cudaEventCreate(&rec.a);
      cudaEventCreate(&rec.b);
      rec.flops = 0;
      prof.push_back(rec);
    }
    pr = &prof[prof_used++];
    pr->flops = 2.0 * B * H * W_ * (double)L.Cout * L.K;
    pr->bytes = 2.0 * ((double)B * H * W_ * (L.Cin + L.Cout) + (double)L.Cout * L.K);
    pr->layer = &L;
    pr->tc = 1;
    cudaEventRecord(pr->a, st);
  }
  int rc = launch_conv_tc_pre(L, x1, L.Cin - c2, x2, c2, r.bn1s, r.bn1b, y, B, H, W_, st);
  if (pr) cudaEventRecord(pr->b, st);
  if (rc && !sticky_rc) {
    sticky_rc = rc;
    err = "tcgen05 pre-activated conv launch failed for " + L.name;
  }
}

// act == null: the caller checked preact_fold_ok(); conv1 reads the raw input and pre-activates it on the fly
template <typename T>
T* Engine::run_residual(const ResidualBlock& r, const T* rawx, const T* act, int B, int H, int W_, Arena& ar,
                        cudaStream_t st, const T* raw2, int c2) {
  const int64_t px = (int64_t)B * H * W_;
  T* t1 = aalloc<T>(ar, px * r.c1.Cout);
  T* t2 = aalloc<T>(ar, px * r.c2.Cout);
  T* sk = r.need_skip ? aalloc<T>(ar, px * r.cout) : nullptr;
  T* out = aalloc<T>(ar, px * r.cout);
  if (!ar.base || ar.overflow) return nullptr;
  if (act) {
    conv<T>(r.c1, act, t1, nullptr, B, H, W_, st);
  } else if (sizeof(T) == 2) {
    conv_pre(r, reinterpret_cast<const __nv_bfloat16*>(rawx), reinterpret_cast<const __nv_bfloat16*>(raw2), raw2 ? c2 : 0,
             reinterpret_cast<__nv_bfloat16*>(t1), B, H, W_, st);
  } else if (!sticky_rc) {
    sticky_rc = DIRB200_E_STATE;
    err = "folded pre-activation needs the tensor-core conv";
  }
  conv<T>(r.c2, t1, t2, nullptr, B, H, W_, st);
  if (r.need_skip && conv_pair<T>(r.c3skip, r.c3, r.skip, t2, rawx, out, B, H, W_, st, raw2, c2)) return out;
  if (raw2) {  // callers check virtual_concat_ok() first
    if (!sticky_rc) {
      sticky_rc = DIRB200_E_STATE;
      err = "virtual concat needs the tensor-core pair GEMM";
    }
    return out;
  }
  const T* resid = rawx;
  if (r.need_skip) {
    conv<T>(r.skip, rawx, sk, nullptr, B, H, W_, st);
    resid = sk;
  }
  conv<T>(r.c3, t2, out, resid, B, H, W_, st);
  return out;
}

template <typename T>
int Engine::run_init(const T* c4, int B, float* stage_rec, int rec_stride, float* para, int para_stride, Arena& ar,
                     cudaStream_t st) {
  const int F = c4ch;  // feat_dim: 2048 (ResNet-50) or 256 (HRNet-W32)
  T* attn_act = aalloc<T>(ar, (int64_t)B * 64 * F);
  float* attn = aalloc<float>(ar, (int64_t)B * 64 * 2);
  float* pooled = aalloc<float>(ar, (int64_t)B * 3 * F);
  if (!ar.base) return DIRB200_OK;
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  conv<T>(attn_conv, c4, attn_act, nullptr, B, 8, 8, st);
  launch_attn_logits<T>(attn_act, attn_w, attn_b, attn, B, 64, F / 2, st);
  launch_attn_pool<T>(c4, attn, pooled, B, 64, F, st);
  RegressArgs a{};
  for (int h = 0; h < 2; ++h) {
    a.in0[h] = VecSeg{pooled + h * F, F, 3 * F};
    a.in1[h] = VecSeg{nullptr, 0, 0};
    a.Wm[h] = init_Wm[h];
    a.bm[h] = init_bm[h];
    a.mano[h] = mano[0][h];
  }
  a.off0 = VecSeg{pooled + 2 * F, F, 3 * F};
  a.off1 = VecSeg{nullptr, 0, 0};
  a.Wo = init_Wo;
  a.bo = init_bo;
  a.stage_record = stage_rec;
  a.rec_stride = rec_stride;
  a.mano_para = para;
  a.para_stride = para_stride;
  a.do_proj_feat = 0;
  a.B = B;
  launch_regress_mano(a, st);
  launches += 3;
  return DIRB200_OK;
}

// SemGCN stack of stage s for both hands (SemGCN/p_gcn.py:63-73) + global_pos_emb (models/dir.py:103-110):
// x (B,2,21,128) -> y (B,2,21,128); gh0/gh1 = two (2,B,2,21,128) scratch planes. tc: tf32 tcgen05 GEMMs (bf16 config).
void Engine::run_gcn(int s, bool tc, const float* x, float* gh0, float* gh1, const float* prev_rec, int prev_stride,
                     float* y, int B, int skip_gpos, cudaStream_t st) {
  const StageWeights& sw = stage[s];
  auto agg_of = [&](int l) {
    GcnAgg g{};
    for (int h = 0; h < 2; ++h) {
      g.A1[h] = sw.gcn[l].A1[h];
      g.scale[h] = sw.gcn[l].scale[h];
      g.shift[h] = sw.gcn[l].shift[h];
    }
    return g;
  };
  float *hin = nullptr, *hout = gh0;
  for (int l = 0; l < 4; ++l) {
    GcnGemmArgs g{};
    g.x = l == 0 ? x : nullptr;
    g.hin = hin;
    if (l > 0) g.agg = agg_of(l - 1);
    g.hout = hout;
    for (int h = 0; h < 2; ++h) g.W[h] = sw.gcn[l].W[h];
    g.B = B;
    if (tc && sw.gcn[l].Wtc[0] && sw.gcn[l].Wtc[1] && !gcn_simt)
      launch_gcn_gemm_tc(g, sw.gcn[l].Wtc[0], sw.gcn[l].Wtc[1], st);
    else
      launch_gcn_gemm(g, st);
    hin = hout;
    hout = (hout == gh0) ? gh1 : gh0;
  }
  GcnFinishArgs f{};
  f.hin = hin;
  f.agg = agg_of(3);
  f.gpos = sw.gpos;
  f.prev_record = prev_rec;
  f.rec_stride = prev_stride;
  f.y = y;
  f.B = B;
  f.skip_gpos = skip_gpos;
  launch_gcn_finish(f, st);
}

// bone_proj x2 -> cat -> fusion conv3x3(2560->256)+BN+ReLU -> conv1x1 (models/dir.py:118-122,146-174,57-62) of stage s:
// uv from stage_rec, joint features jfeat (B,2,21,64) -> out NHWC (B,S,S,256). Scratch: bone (dense path only), coef, fus_mid.
template <typename T>
void Engine::run_bone_fusion(int s, const float* stage_rec, int rec_stride, const float* jfeat, int B, T* bone, float* coef,
                             T* fus_mid, T* out, cudaStream_t st) {
  const StageWeights& sw = stage[s];
  const int S = sw.S;
  if (dense_fusion) {
    launch_bone_raster<T>(stage_rec, rec_stride, jfeat, bone, B, S, sw.distance, st);
    ++launches;
    conv<T>(sw.fusion0, bone, fus_mid, nullptr, B, S, S, st);
  } else {
    const bool coef_tc = std::is_same<T, __nv_bfloat16>::value && sw.fus_wp_tc && !coef_simt;
    const bool fus_tc = std::is_same<T, __nv_bfloat16>::value && !fusion_simt && (S == 16 || S == 32);
    const int p_bf16 = coef_tc && fus_tc;  // both ends on the tensor cores: the coefficient tensor travels as bf16
    if (coef_tc)
      launch_bone_coef_tc(jfeat, sw.fus_wp_tc, coef, p_bf16, B, st);
    else
      launch_bone_coef(jfeat, sw.fus_wp, coef, B, st);
    if (fus_tc)
      launch_bone_fusion_tc(stage_rec, rec_stride, coef, p_bf16, sw.fusion0.scale, sw.fusion0.shift,
                            reinterpret_cast<__nv_bfloat16*>(fus_mid), B, S, sw.distance, st);
    else
      launch_bone_fusion<T>(stage_rec, rec_stride, coef, sw.fusion0.scale, sw.fusion0.shift, fus_mid, B, S, sw.distance,
                            st);
    launches += 2;
  }
  conv<T>(sw.fusion3, fus_mid, out, nullptr, B, S, S, st);
}

template <typename T>
int Engine::run_stage(int s, const T* img_feat, const float* prev_rec, int prev_stride, const float* prev_para,
                      int prev_para_stride, int B, float* stage_rec, int rec_stride, float* para, int para_stride,
                      T** img_feat_out, float** joint_feat_out, float* vis_nchw, Arena& ar, cudaStream_t st) {
  const StageWeights& sw = stage[s];
  const int S = sw.S;
  float* jf0 = aalloc<float>(ar, (int64_t)B * 42 * 128);
  float* jf1 = aalloc<float>(ar, (int64_t)B * 42 * 128);
  float* gh0 = aalloc<float>(ar, (int64_t)2 * B * 42 * 128);
  float* gh1 = aalloc<float>(ar, (int64_t)2 * B * 42 * 128);
  float* tok = aalloc<float>(ar, (int64_t)B * 42 * 64);
  float* jfeat = aalloc<float>(ar, (int64_t)B * 42 * 64);
  T* bone = dense_fusion ? aalloc<T>(ar, (int64_t)B * S * S * 2560) : nullptr;
  float* coef = dense_fusion ? nullptr : aalloc<float>(ar, (int64_t)B * 40 * 2 * 9 * 256);
  T* fus_mid = aalloc<T>(ar, (int64_t)B * S * S * 256);
  T* out = aalloc<T>(ar, (int64_t)B * S * S * 256);
  if (!ar.base) return DIRB200_OK;
  if (ar.overflow) return DIRB200_E_WORKSPACE;

  EmbedArgs e{};
  e.feat = img_feat;
  e.S = S;
  e.prev_record = prev_rec;
  e.rec_stride = prev_stride;
  for (int h = 0; h < 2; ++h) {
    e.filters[h] = sw.filters[h];
    e.pos[h] = sw.pos[h];
  }
  e.out = jf0;
  e.B = B;
  launch_joint_embed<T>(e, st);
  run_gcn(s, std::is_same<T, __nv_bfloat16>::value, jf0, gh0, gh1, prev_rec, prev_stride, jf1, B, 0, st);
  float* gin = jf1;
  if (std::is_same<T, __nv_bfloat16>::value && sw.ste_packed && !ste_simt)
    launch_ste_tc(gin, tok, sw.ste, sw.ste_packed, B, st);
  else
    launch_ste(gin, tok, sw.ste, B, st);
  RegressArgs a{};
  for (int h = 0; h < 2; ++h) {
    a.in0[h] = VecSeg{tok + h * 1344, 1344, 2688};
    a.in1[h] = VecSeg{prev_para + h * 64, 64, prev_para_stride};
    a.Wm[h] = sw.Wm[h];
    a.bm[h] = sw.bm[h];
    a.mano[h] = mano[s + 1][h];
  }
  a.off0 = VecSeg{tok, 2688, 2688};
  a.off1 = VecSeg{prev_rec + DIRB200_OFF_OFFSET, 3, prev_stride};
  a.Wo = sw.Wo;
  a.bo = sw.bo;
  a.stage_record = stage_rec;
  a.rec_stride = rec_stride;
  a.mano_para = para;
  a.para_stride = para_stride;
  a.do_proj_feat = 1;
  a.proj_feat = sw.proj_feat;
  a.joint_feat = jfeat;
  a.B = B;
  launch_regress_mano(a, st);
  launches += 8;
  const bool vis_on_side = vis_nchw && side && !no_overlap;
  if (vis_on_side) {  // proj_feat (aux output) depends only on this stage's uv + joint features: HBM-bound, so it runs
                      // on the side stream underneath the instruction-bound fusion kernels; forward() joins ev_join[1]
    cudaEventRecord(ev_fork[1], st);
    cudaStreamWaitEvent(side, ev_fork[1], 0);
    launch_bone_vis_nchw(stage_rec + DIRB200_OFF_UV_L, stage_rec + DIRB200_OFF_UV_R, rec_stride, jfeat, jfeat + 21 * 64,
                         42 * 64, vis_nchw, B, S, sw.distance, 1, side);
    cudaEventRecord(ev_join[1], side);
    ++launches;
  }
  run_bone_fusion<T>(s, stage_rec, rec_stride, jfeat, B, bone, coef, fus_mid, out, st);
  if (vis_nchw && !vis_on_side) {
    launch_bone_vis_nchw(stage_rec + DIRB200_OFF_UV_L, stage_rec + DIRB200_OFF_UV_R, rec_stride, jfeat, jfeat + 21 * 64,
                         42 * 64, vis_nchw, B, S, sw.distance, 1, st);
    ++launches;
  }
  *img_feat_out = out;
  if (joint_feat_out) *joint_feat_out = jfeat;
  return DIRB200_OK;
}

template <typename T>
int Engine::forward(const float* img, int B, Arena& ar, const dirb200_outputs* o, cudaStream_t st) {
  launches = 0;
  tc_launches = 0;
  sticky_rc = 0;
  const bool plan = ar.base == nullptr;
  T *c2 = nullptr, *c3 = nullptr, *c4 = nullptr;
  int rc = hrnet() ? run_backbone_hrnet<T>(img, B, ar, nullptr, &c2, &c3, &c4, st)
                   : run_backbone<T>(img, B, 256, 256, ar, nullptr, &c2, &c3, &c4, st);
  if (rc) return rc;
  float* rec = plan ? nullptr : o->record;
  float* para = plan ? nullptr : o->mano_para;
  const int RS = DIRB200_RECORD_FLOATS, PS = 3 * 128;
  rc = run_init<T>(c4, B, rec, RS, para, PS, ar, st);
  if (rc) return rc;

  auto concat = [&](const T* s0, int C0, int up, const T* s1, int C1, const ResidualBlock& r, int S, T** rawo,
                    T** acto) {
    const int64_t n = (int64_t)B * S * S * (C0 + C1);
    // bf16 configuration: the pre-activated copy is never written (conv1 of `r` pre-activates its A tiles, PRE kernels);
    // what remains of this pass is the upsample + concat into `raw`, and nothing at all for a single full-size source
    const bool fold = preact_fold_ok<T>(r, B, S, S, 0);
    *rawo = s1 ? aalloc<T>(ar, n) : nullptr;
    *acto = fold ? nullptr : aalloc<T>(ar, n);
    if (plan || ar.overflow || (fold && !s1)) return;
    launch_concat_preact<T>(s0, C0, up, s1, C1, r.bn1s, r.bn1b, *rawo, *acto, B, S, S, st);
    ++launches;
  };
  // fusion_layer{4,3}: Residual over cat(upsample2x(s0), s1) (models/dir.py:443-444,460-461). With the pre-activation
  // folded and the pair GEMM available only upsample2x(s0) is materialised; conv1 and the skip conv read s1 in place.
  auto up_residual = [&](const ResidualBlock& r, const T* s0, int C0, const T* s1, int C1, int S) -> T* {
    if (virtual_concat_ok<T>(r, B, S, S) && preact_fold_ok<T>(r, B, S, S, C1)) {
      T* up = aalloc<T>(ar, (int64_t)B * S * S * C0);
      if (!plan && !ar.overflow) {
        launch_concat_preact<T>(s0, C0, 1, (const T*)nullptr, 0, r.bn1s, r.bn1b, up, (T*)nullptr, B, S, S, st);
        ++launches;
      }
      return run_residual<T>(r, up, (const T*)nullptr, B, S, S, ar, st, s1, C1);
    }
    T *rw = nullptr, *ac = nullptr;
    concat(s0, C0, 1, s1, C1, r, S, &rw, &ac);
    return run_residual<T>(r, rw, ac, B, S, S, ar, st);
  };
  // ---- stage 1 @16x16 (models/dir.py:442-456)
  T *raw = nullptr, *act = nullptr;
  const ResidualBlock& skip4 = res["decoder.skip_layer4."];
  concat(c3, c3ch, 0, nullptr, 0, skip4, 16, &raw, &act);
  T* c3_skip = run_residual<T>(skip4, c3, act, B, 16, 16, ar, st);
  const ResidualBlock& fus4 = res["decoder.fusion_layer4."];
  T* fusion4 = up_residual(fus4, c4, c4ch, c3_skip, 256, 16);
  if (cfg.refine_stages == 1) {  // "1 refine iter": init regression + projecter_4 only (truncation after models/dir.py:456)
    T* img_feat1 = nullptr;
    rc = run_stage<T>(0, fusion4, rec, RS, para, PS, B, plan ? nullptr : rec + DIRB200_STAGE_FLOATS, RS,
                      plan ? nullptr : para + 128, PS, &img_feat1, nullptr, nullptr, ar, st);
    if (rc) return rc;
    if (plan) return DIRB200_OK;
    if (ar.overflow) return DIRB200_E_WORKSPACE;
    cudaMemset2DAsync(rec + 2 * DIRB200_STAGE_FLOATS, (size_t)RS * 4, 0, (size_t)DIRB200_STAGE_FLOATS * 4, B, st);
    cudaMemset2DAsync(para + 256, (size_t)PS * 4, 0, 128 * 4, B, st);
    last_forward_launches = launches;
    if (sticky_rc) return sticky_rc;
    CK(cudaPeekAtLastError());
    return DIRB200_OK;
  }
  // skip_layer3 (models/dir.py:459) needs only c2: fork it onto the side stream here, so its convs fill the SMs that
  // stage 1's small joint-space grids (SemGCN, mixSTE, MANO: 64-336 CTAs, latency-bound) leave idle
  const ResidualBlock& skip3 = res["decoder.skip_layer3."];
  const bool overlap = side && !no_overlap && !plan;
  cudaStream_t st3 = overlap ? side : st;
  if (overlap) {
    cudaEventRecord(ev_fork[0], st);
    cudaStreamWaitEvent(side, ev_fork[0], 0);
  }
  T *raw3 = nullptr, *act3 = nullptr;
  if (!preact_fold_ok<T>(skip3, B, 32, 32, 0)) {
    const int64_t n = (int64_t)B * 32 * 32 * c2ch;
    act3 = aalloc<T>(ar, n);
    if (!plan && !ar.overflow) {
      launch_concat_preact<T>(c2, c2ch, 0, (const T*)nullptr, 0, skip3.bn1s, skip3.bn1b, raw3, act3, B, 32, 32, st3);
      ++launches;
    }
  }
  T* c2_skip = run_residual<T>(skip3, c2, act3, B, 32, 32, ar, st3);
  if (overlap) cudaEventRecord(ev_join[0], side);
  // an early return must not leave work of this forward running on the side stream behind the caller's back
  auto bail = [&](int code) {
    if (side && !no_overlap && !plan) {
      cudaEventRecord(ev_join[0], side);
      cudaStreamWaitEvent(st, ev_join[0], 0);
    }
    return code;
  };
  T* img_feat1 = nullptr;
  rc = run_stage<T>(0, fusion4, rec, RS, para, PS, B, plan ? nullptr : rec + DIRB200_STAGE_FLOATS, RS,
                    plan ? nullptr : para + 128, PS, &img_feat1, nullptr, nullptr, ar, st);
  if (rc) return bail(rc);
  // enhance_layer{4,3}: the block input cat(fusion, img_feat) is only pre-activated (act); its raw copy is never
  // written: the skip half of the pair GEMM reads the two sources directly (third A operand of conv_tc_kernel)
  auto enhance = [&](const ResidualBlock& r, const T* a0, const T* a1, int S) -> T* {
    if (virtual_concat_ok<T>(r, B, S, S) && preact_fold_ok<T>(r, B, S, S, 256))  // neither concat nor pre-activation
      return run_residual<T>(r, a0, (const T*)nullptr, B, S, S, ar, st, a1, 256);  // is materialised
    if (virtual_concat_ok<T>(r, B, S, S)) {
      T* acto = aalloc<T>(ar, (int64_t)B * S * S * 512);
      if (!plan && !ar.overflow) {
        launch_concat_preact<T>(a0, 256, 0, a1, 256, r.bn1s, r.bn1b, (T*)nullptr, acto, B, S, S, st);
        ++launches;
      }
      return run_residual<T>(r, a0, acto, B, S, S, ar, st, a1, 256);
    }
    T *rw = nullptr, *ac = nullptr;
    concat(a0, 256, 0, a1, 256, r, S, &rw, &ac);
    return run_residual<T>(r, rw, ac, B, S, S, ar, st);
  };
  const ResidualBlock& enh4 = res["decoder.enhance_layer4."];
  T* enhance4 = enhance(enh4, fusion4, img_feat1, 16);
  // ---- stage 2 @32x32 (models/dir.py:459-471)
  if (overlap) cudaStreamWaitEvent(st, ev_join[0], 0);
  const ResidualBlock& fus3 = res["decoder.fusion_layer3."];
  T* fusion3 = up_residual(fus3, enhance4, 256, c2_skip, 256, 32);
  T* img_feat2 = nullptr;
  const bool aux = cfg.aux_outputs != 0;
  rc = run_stage<T>(1, fusion3, plan ? nullptr : rec + DIRB200_STAGE_FLOATS, RS, plan ? nullptr : para + 128, PS, B,
                    plan ? nullptr : rec + 2 * DIRB200_STAGE_FLOATS, RS, plan ? nullptr : para + 256, PS, &img_feat2,
                    nullptr, (aux && !plan) ? o->proj_feat : nullptr, ar, st);
  if (rc) return bail(rc);
  if (aux) {  // models/dir.py:470-476: only the seg/dense heads consume enhance_layer3
    const ResidualBlock& enh3 = res["decoder.enhance_layer3."];
    T* enhance3 = enhance(enh3, fusion3, img_feat2, 32);
    const int64_t px = (int64_t)B * 32 * 32;
    T* fmid = aalloc<T>(ar, px * 256);
    T* feat = aalloc<T>(ar, px * 256);
    T* sd = aalloc<T>(ar, px * 256);
    if (!plan && !ar.overflow) {
      conv<T>(conv_final0, enhance3, fmid, nullptr, B, 32, 32, st);
      conv<T>(conv_final3, fmid, feat, nullptr, B, 32, 32, st);
      conv<T>(segdense0, feat, sd, nullptr, B, 32, 32, st);
      launch_head3x2<T>(sd, seg3_w, seg3_b, dense3_w, dense3_b, o->seg, o->dense, B, 1024, st);
      launches += 1;
    }
  }
  if (plan) return DIRB200_OK;
  if (ar.overflow) return bail(DIRB200_E_WORKSPACE);
  if (aux && side && !no_overlap) cudaStreamWaitEvent(st, ev_join[1], 0);  // proj_feat rasteriser (run_stage)
  last_forward_launches = launches;
  if (sticky_rc) return bail(sticky_rc);
  CK(cudaPeekAtLastError());
  return DIRB200_OK;
}

template void Engine::conv<float>(const ConvLayer&, const float*, float*, const float*, int, int, int, cudaStream_t,
                                  bool);
template void Engine::conv<__nv_bfloat16>(const ConvLayer&, const __nv_bfloat16*, __nv_bfloat16*,
                                          const __nv_bfloat16*, int, int, int, cudaStream_t, bool);
template int Engine::forward<float>(const float*, int, Arena&, const dirb200_outputs*, cudaStream_t);
template int Engine::forward<__nv_bfloat16>(const float*, int, Arena&, const dirb200_outputs*, cudaStream_t);
template int Engine::run_backbone<float>(const float*, int, int, int, Arena&, float**, float**, float**, float**,
                                         cudaStream_t);
template int Engine::run_backbone<__nv_bfloat16>(const float*, int, int, int, Arena&, __nv_bfloat16**,
                                                 __nv_bfloat16**, __nv_bfloat16**, __nv_bfloat16**, cudaStream_t);
template bool Engine::preact_fold_ok<float>(const ResidualBlock&, int, int, int, int) const;
template bool Engine::preact_fold_ok<__nv_bfloat16>(const ResidualBlock&, int, int, int, int) const;
template float* Engine::run_residual<float>(const ResidualBlock&, const float*, const float*, int, int, int, Arena&,
                                            cudaStream_t, const float*, int);
template __nv_bfloat16* Engine::run_residual<__nv_bfloat16>(const ResidualBlock&, const __nv_bfloat16*,
                                                            const __nv_bfloat16*, int, int, int, Arena&, cudaStream_t,
                                                            const __nv_bfloat16*, int);
template void Engine::run_bone_fusion<float>(int, const float*, int, const float*, int, float*, float*, float*, float*,
                                             cudaStream_t);
template void Engine::run_bone_fusion<__nv_bfloat16>(int, const float*, int, const float*, int, __nv_bfloat16*, float*,
                                                     __nv_bfloat16*, __nv_bfloat16*, cudaStream_t);
template int Engine::run_init<float>(const float*, int, float*, int, float*, int, Arena&, cudaStream_t);
template int Engine::run_init<__nv_bfloat16>(const __nv_bfloat16*, int, float*, int, float*, int, Arena&,
                                             cudaStream_t);
template int Engine::run_stage<float>(int, const float*, const float*, int, const float*, int, int, float*, int, float*,
                                      int, float**, float**, float*, Arena&, cudaStream_t);
template int Engine::run_stage<__nv_bfloat16>(int, const __nv_bfloat16*, const float*, int, const float*, int, int,
                                              float*, int, float*, int, __nv_bfloat16**, float**, float*, Arena&,
                                              cudaStream_t);

}  // namespace dirb200
