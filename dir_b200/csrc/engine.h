// Engine: owns packed weights and sequences the kernels of DIR's eval forward.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/dirb200.h"
#include "kernels.h"

namespace dirb200 {

struct RawWeight {
  const void* p;
  int dtype;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct ConvLayer {
  std::string name;  // state_dict key of the weight (profiling hook / error messages)
  int Cin = 0, Cout = 0, kh = 1, kw = 1, stride = 1, pad = 0, K = 0, Kpad = 0, relu = 0;
  float* w32 = nullptr;
  __nv_bfloat16* w16 = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  CUtensorMap wmap;  // bf16 weights [Cout][Kpad], box 64(k) x BN rows, 128B swizzle (tensor-core path)
  int wmap_bn = 0;   // rows per weight box (0: no tensor map)
  CUtensorMap wmap2;  // same weights with wmap_bn/2-row boxes: operand halves of the 2-CTA (cta_group::2) kernel
  bool wmap2_ok = false;
  // fp32 / tf32 configurations (conv_tf32.cu): the weights split into tf32 hi + lo parts, fp32 [Cout][Kpad] each
  float* w32hi = nullptr;
  float* w32lo = nullptr;
  CUtensorMap wmap32hi, wmap32lo;  // box 32(k) x wmap32_bn rows, 128B swizzle
  int wmap32_bn = 0;               // rows per weight box (0: no tensor-core path for this layer)
  CUtensorMap wmap32hi64, wmap32lo64;  // the same weights as 64-row boxes (narrow tiles for small-M launches)
  bool wmap32_alt64 = false;
  bool tf32_stem = false;          // 7x7/s2/Cin=3 stem packed for conv_tf32.cu's stem variant
  int tc_bn_cap = 256;   // largest n-tile the tensor-core path may pick (128 for layers that add a residual)
  bool tc_stem = false;  // 7x7/s2/Cin=3 stem packed for the tensor-core stem variant
};

struct Bottleneck {
  ConvLayer c1, c2, c3, ds;
  bool has_ds = false;
  ConvLayer c3ds;  // tensor-core path: conv3 and the downsample conv as one K-concatenated GEMM (wmap_bn > 0 if built)
};

struct ResidualBlock {  // hourglass Residual
  int cin = 0, cout = 0;
  bool need_skip = true;
  float *bn1s = nullptr, *bn1b = nullptr;
  ConvLayer c1, c2, c3, skip;
  ConvLayer c3skip;  // tensor-core path: conv3 and the skip conv as one K-concatenated GEMM
};

struct StageWeights {
  int S = 16;
  float distance = 1.f;
  PointMlp filters[2], pos[2], gpos, proj_feat;
  struct Gcn {
    const float* W[2];
    const void* Wtc[2] = {nullptr, nullptr};  // bf16 configuration: tf32 operand tiles of gcn_gemm_tc_kernel
    const float* A1[2];
    const float* scale[2];
    const float* shift[2];
  } gcn[4];
  SteWeights ste;
  const void* ste_packed = nullptr;  // bf16 configuration: operand tiles of the tcgen05 STE kernel (ste_tc.cu)
  const float *Wm[2], *bm[2], *Wo, *bo;
  ConvLayer fusion0, fusion3;
  const float* fus_wp = nullptr;  // fusion.0 weights packed for the factored form [40][64][9][256]
  const void* fus_wp_tc = nullptr;  // bf16 configuration: the same weights as tf32 operand tiles (fusion.cu)
};

struct Arena {
  char* base = nullptr;
  size_t size = 0, off = 0;
  bool overflow = false;
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (!base) return nullptr;
    if (off > size) {
      overflow = true;
      return nullptr;
    }
    return base + a;
  }
};

struct Engine {
  dirb200_config cfg{};
  std::string err;
  std::map<std::string, RawWeight> raw;
  std::vector<std::string> required;
  bool finalized = false;
  bool dry = false;  // enumerate required keys without touching the GPU
  std::vector<void*> owned;
  cudaStream_t fin_stream = nullptr;
  int launches = 0;
  int last_forward_launches = 0;
  int tc_launches = 0;      // convs that went to the tcgen05 kernel since the last forward() start
  int sticky_rc = 0;        // first launch error inside a forward
  // Side stream for the two independent branches of the decoder (forked / joined with events relative to the
  // caller's stream, so a forward stays one asynchronous, graph-capturable unit): skip_layer3 (needs only c2) runs
  // under stage 1's latency-bound joint-space kernels, the HBM-bound proj_feat rasteriser under stage 2's fusion.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  bool no_overlap = false;    // DIRB200_NO_OVERLAP=1: everything on the caller's stream
  bool fusion_simt = false;   // DIRB200_FUSION_SIMT=1: CUDA-core sparse accumulate also in the bf16 configuration
  bool gcn_simt = false;      // DIRB200_GCN_SIMT=1: fp32 CUDA-core SemGCN GEMMs also in the bf16 configuration
  bool coef_simt = false;     // DIRB200_COEF_SIMT=1: fp32 CUDA-core bone_coef also in the bf16 configuration
  bool stem_split = false;    // DIRB200_STEM_SPLIT=1: stem conv and max-pool as two kernels (TMA implicit GEMM + pool)
  const void* stem_pool_w = nullptr;  // conv1 weights packed for stem_pool_kernel
  bool ste_simt = false;      // DIRB200_STE_SIMT=1: fp32 CUDA-core mixSTE also in the bf16 configuration
  bool dense_fusion = false;  // DIRB200_DENSE_FUSION=1: materialise bone_proj and run the dense 2560-ch conv
  bool disable_pair_fusion = false;  // DIRB200_NO_PAIR_FUSION=1: keep conv3 and skip/downsample as separate launches
  bool disable_tc = false;  // DIRB200_DISABLE_TC=1: force the CUDA-core conv in bf16 mode (debug A/B)
  bool no_b2b = false;          // DIRB200_NO_B2B=1: bottleneck tail and the next block's conv1 as separate launches
  bool no_preact_fold = false;  // DIRB200_NO_PREACT_FOLD=1: Residual pre-activations as a separate pass (concat_preact_kernel)
  bool no_halo = false;     // DIRB200_NO_HALO=1: 64-channel 3x3 convs on the per-tap kernel (conv_tc.cu) instead of conv_halo.cu
  bool fp32_simt = false;   // DIRB200_FP32_SIMT=1: fp32 configuration on the CUDA-core conv (the round-1 path; debug A/B)
  const ConvLayer* find_conv(const std::string& weight_key) const;
  void* nccl_comm = nullptr;
  void (*nccl_destroy)(void*) = nullptr;  // set with nccl_comm (capi.cu binds NCCL at run time)
  const unsigned char* img_u8 = nullptr;  // set by forward_u8: raw uint8 HWC BGR frames instead of the fp32 image
  // timing hook (bench.py roofline): CUDA events around every conv launch whose name starts with prof_prefix
  struct ProfRec {
    cudaEvent_t a, b;
    double flops;
    double bytes;  // compulsory activation+weight bytes of the launch (input + output (+ residual) + weights)
    const ConvLayer* layer;
    int tc;
  };
  bool prof_on = false;
  std::string prof_prefix;
  std::vector<ProfRec> prof;
  size_t prof_used = 0;

  // ---- HRNet extension (cfg.backbone = 32; parity unpinned, oracle/hrnet_oracle.py): the backbone as a flat program
  // over numbered NHWC buffers (no reuse: ~55 MB of bf16 activations per image)
  struct HrOp {
    int kind;                // 0: conv (+BN)(+residual)(+ReLU), 1: fuse = relu(sum of up to 4 nearest-upsampled terms)
    int layer;               // index into hr_convs
    int in, out, res;        // buffer ids; in = -1: the NCHW fp32 image; res = -1: none
    int nterm, term[4], shift[4];
  };
  struct HrBuf {
    int H, W, C;             // per image at a 256x256 input; C = packed channel count (multiple of 64, zero padded)
  };
  std::vector<ConvLayer> hr_convs;
  std::vector<HrOp> hr_ops;
  std::vector<HrBuf> hr_bufs;
  int hr_out[4] = {-1, -1, -1, -1};
  int c1ch = 256, c2ch = 512, c3ch = 1024, c4ch = 2048;  // packed channels of the pyramid the decoder sees
  bool hrnet() const { return cfg.backbone != 0; }
  void build_hrnet(int width);
  template <typename T>
  int run_backbone_hrnet(const float* img, int B, Arena& ar, T** c1, T** c2, T** c3, T** c4, cudaStream_t st);

  // ---- network
  ConvLayer stem;
  std::vector<Bottleneck> layers[4];
  ConvLayer attn_conv;  // both hands, Cout = 2048
  float *attn_w = nullptr, *attn_b = nullptr;
  const float *init_Wm[2] = {nullptr, nullptr}, *init_bm[2] = {nullptr, nullptr}, *init_Wo = nullptr, *init_bo = nullptr;
  ManoWeights mano[3][2];
  std::map<std::string, ResidualBlock> res;
  StageWeights stage[2];
  ConvLayer conv_final0, conv_final3, segdense0;
  float *seg3_w = nullptr, *seg3_b = nullptr, *dense3_w = nullptr, *dense3_b = nullptr;

  ~Engine();
  bool bf16() const { return cfg.precision == DIRB200_PRECISION_BF16; }
  int tf32_nsplit() const { return cfg.precision == DIRB200_PRECISION_TF32 ? 1 : 3; }  // MMAs per k-step, conv_tf32.cu
  size_t esize() const { return bf16() ? 2 : 4; }

  // finalize helpers
  int build(cudaStream_t st);
  const float* W(const std::string& name, std::vector<int64_t> shape = {});
  float* dalloc(size_t nfloats);
  float* copy_of(const std::string& name);
  float* transposed(const std::string& name, int rows, int cols);
  float* transposed_pairs(const std::string& name, int N, int K);
  void fold(const std::string& conv_bias, const std::string& bn, float** scale, float** shift, int n, int off = 0,
            int total = 0);
  // cin_pad / cout_pad > 0: pack with zero-padded input / output channels (HRNet's 32-channel branch lives in 64)
  ConvLayer make_conv(const std::string& wname, const std::string& bias, const std::string& bn, int stride, int pad,
                      int relu, int cin_pad = 0, int cout_pad = 0);
  void fuse_post_bn(ConvLayer& c, const std::string& conv_bias, const std::string& bn);
  void prepare_tf32(ConvLayer& L);
  PointMlp make_mlp(const std::string& prefix, int cin, int cmid, int cout);
  ResidualBlock make_residual(const std::string& prefix, int cin_pad = 0);
  ManoWeights make_mano(const std::string& prefix, bool left);
  void build_stage(int s, const std::string& p);

  // forward pieces (T = activation element type)
  template <typename T>
  void conv(const ConvLayer& L, const T* x, T* y, const T* res, int B, int H, int W_, cudaStream_t st,
            bool in_nchw = false);
  void make_dual(ConvLayer& F, const ConvLayer& main, const ConvLayer& second);
  template <typename T>
  bool conv_b2b(const Bottleneck& bk, const Bottleneck& nx, const T* t2, const T* x, T* out, T* t1_next, int B, int Ho,
                int Wo, cudaStream_t st);
  template <typename T>
  bool conv_pair(const ConvLayer& F, const ConvLayer& main, const ConvLayer& second, const T* x1, const T* x2, T* y,
                 int B, int Ho, int Wo, cudaStream_t st, const T* x2b = nullptr, int C2b = 0);
  template <typename T>
  int run_backbone(const float* img, int B, int H, int W_, Arena& ar, T** c1, T** c2, T** c3, T** c4, cudaStream_t st);
  template <typename T>
  // raw2 != null: the block input is the channel concat [raw (cin - c2 channels) | raw2 (c2 channels)], never built
  T* run_residual(const ResidualBlock& r, const T* raw, const T* act, int B, int H, int W_, Arena& ar, cudaStream_t st,
                  const T* raw2 = nullptr, int c2 = 0);
  template <typename T>
  bool preact_fold_ok(const ResidualBlock& r, int B, int H, int W_, int c2) const;
  void conv_pre(const ResidualBlock& r, const __nv_bfloat16* x1, const __nv_bfloat16* x2, int c2, __nv_bfloat16* y, int B,
                int H, int W_, cudaStream_t st);
  // true if run_residual can take the skip operand of `r` from two sources (tensor-core pair GEMM available)
  template <typename T>
  bool virtual_concat_ok(const ResidualBlock& r, int B, int H, int W_) const;
  template <typename T>
  int run_init(const T* c4, int B, float* stage_rec, int rec_stride, float* para, int para_stride, Arena& ar,
               cudaStream_t st);
  void run_gcn(int s, bool tc, const float* x, float* gh0, float* gh1, const float* prev_rec, int prev_stride, float* y,
               int B, int skip_gpos, cudaStream_t st);
  template <typename T>
  void run_bone_fusion(int s, const float* stage_rec, int rec_stride, const float* jfeat, int B, T* bone, float* coef,
                       T* fus_mid, T* out, cudaStream_t st);
  template <typename T>
  int run_stage(int s, const T* img_feat, const float* prev_rec, int prev_stride, const float* prev_para,
                int prev_para_stride, int B, float* stage_rec, int rec_stride, float* para, int para_stride,
                T** img_feat_out, float** joint_feat_out, float* vis_nchw, Arena& ar, cudaStream_t st);
  template <typename T>
  int forward(const float* img, int B, Arena& ar, const dirb200_outputs* out, cudaStream_t st);
};

// fp32-grade tensor-core conv (conv_tf32.cu): 3xTF32 with chunked round-to-nearest accumulation (nsplit 3) or plain TF32
bool conv_tf32_supported(const ConvLayer& L, int B, int H, int W);
int conv_tf32_prepare_weights(ConvLayer& L, float* hi, float* lo, cudaStream_t st);  // splits L.w32, builds the maps
int launch_conv_tf32(const ConvLayer& L, const float* x, float* y, const float* res, int B, int H, int W, int nsplit,
                     cudaStream_t st);
size_t conv_tf32_stem_scratch_bytes(int B, int H, int W);
size_t conv_tf32_stem_weight_floats();
int conv_tf32_prepare_stem(ConvLayer& L, const float* w_raw, float* buf, cudaStream_t st);
bool conv_tf32_stem_supported(const ConvLayer& L, int H, int W);
int launch_conv_tf32_stem(const ConvLayer& L, const float* img, float* scratch, float* y, int B, int H, int W, int nsplit,
                          cudaStream_t st);
// halo-tile 3x3 conv for the 64 -> 64 channel layers (conv_halo.cu): one input fetch per tile instead of one per tap
bool conv_halo_supported(const ConvLayer& L, int B, int H, int W);
int launch_conv_halo(const ConvLayer& L, const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* res, int B, int H,
                     int W, cudaStream_t st);
// back-to-back 1x1 convs across a bottleneck boundary (conv_b2b.cu)
bool conv_b2b_supported(const ConvLayer& first, int Ka, int Kb, const ConvLayer& next, int M, bool has_res);
int launch_conv_b2b(const ConvLayer& first, const __nv_bfloat16* xa, int Ka, const __nv_bfloat16* xb, int Kb,
                    const __nv_bfloat16* res, const ConvLayer& next, __nv_bfloat16* out, __nv_bfloat16* t1, int M,
                    cudaStream_t st);
// tensor-core conv (conv_tc.cu). Returns false if the shape is not supported (caller falls back to CUDA cores).
bool conv_tc_supported(const ConvLayer& L, int B, int H, int W);
int conv_tc_prepare_weights(ConvLayer& L);  // builds L.wmap
int launch_conv_tc(const ConvLayer& L, const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* res, int B,
                   int H, int W, cudaStream_t st);
int conv_tc_prepare_dual(ConvLayer& F, const ConvLayer& main, const ConvLayer& second, __nv_bfloat16* w16, float* scale,
                         float* shift, cudaStream_t st);
// 1x1 conv over relu(x * pre_scale + pre_shift), x = [x1 | x2] by channels, pre-activation applied in shared memory
bool conv_tc_pre_supported(const ConvLayer& L, int B, int H, int W, int C1, int C2);
int launch_conv_tc_pre(const ConvLayer& L, const __nv_bfloat16* x1, int C1, const __nv_bfloat16* x2, int C2,
                       const float* pre_scale, const float* pre_shift, __nv_bfloat16* y, int B, int H, int W,
                       cudaStream_t st);
int launch_conv_tc_dual(const ConvLayer& L, const __nv_bfloat16* x1, int C1, const __nv_bfloat16* x2, int C2, int stride2,
                        __nv_bfloat16* y, int B, int Ho, int Wo, cudaStream_t st, const __nv_bfloat16* x2b = nullptr,
                        int C2b = 0);
int conv_tc_prepare_stem(ConvLayer& L, const float* w_raw, __nv_bfloat16* w_packed, cudaStream_t st);
size_t conv_tc_stem_scratch_bytes(int B, int H, int W);
void launch_stem_pack(const float* img, const unsigned char* img_u8, __nv_bfloat16* scratch, int B, int H, int W,
                      cudaStream_t st);
// stem_pool.cu: conv1 + bn1 + ReLU + maxpool in one kernel, from the packed image
size_t stem_pool_weight_bytes();
void launch_pack_stem_pool_weight(const float* w, void* packed, cudaStream_t st);
int launch_stem_pool(const void* in, const void* packed_w, const float* scale, const float* shift, __nv_bfloat16* y, int B,
                     cudaStream_t st);
int launch_conv_tc_stem(const ConvLayer& L, const float* img, const unsigned char* img_u8, __nv_bfloat16* scratch,
                        __nv_bfloat16* y, int B, int H, int W, cudaStream_t st);

}  // namespace dirb200
