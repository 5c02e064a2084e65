// Host-side launch prototypes of every dirb200 kernel (all asynchronous on `st`).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dirb200 {

// ---------------------------------------------------------------- conv (implicit GEMM)
// y[m, n] = act( (sum_k A[m,k] W[n,k]) * scale[n] + shift[n] (+ res[m,n]) ),  m=(b,ho,wo), k=(ky,kx,ci)
struct ConvArgs {
  const void* x;       // NHWC activations (float or bf16), or NCHW fp32 when in_nchw
  const float* w32;    // [Cout][Kpad] fp32 (CUDA-core path)
  const float* scale;  // [Cout]
  const float* shift;  // [Cout]
  const void* res;     // optional residual, NHWC [M][Cout]
  void* y;             // NHWC [M][Cout]
  int B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad, K, Kpad, relu, in_nchw;
};
template <typename T>
void launch_conv_simt(const ConvArgs& a, cudaStream_t st);

// ---------------------------------------------------------------- elementwise / layout
template <typename T>
void launch_maxpool3x3s2(const T* x, T* y, int B, int H, int W, int C, cudaStream_t st);
// raw[..., 0:C0] = (up0 ? bilinear2x(src0) : src0), raw[..., C0:C0+C1] = src1 (src1 may be null, C1=0)
// act = relu(raw*bn_scale + bn_shift). raw or act may be null. Output spatial size Ho x Wo.
template <typename T>
void launch_concat_preact(const T* src0, int C0, int up0, const T* src1, int C1, const float* bn_scale,
                          const float* bn_shift, T* raw, T* act, int B, int Ho, int Wo, cudaStream_t st);
template <typename T>
void launch_nchw_to_nhwc(const float* x, T* y, int B, int C, int H, int W, cudaStream_t st);
template <typename T>
void launch_nhwc_to_nchw(const T* x, float* y, int B, int C, int H, int W, cudaStream_t st);
// out[b, k, p] (NCHW fp32, k<3) = sum_c w[k][c] * x[b, p, coff + c] + bias[k]   (1x1 conv to 3 channels)
template <typename T>
void launch_head3(const T* x, int Cx, int coff, int C, const float* w, const float* bias, float* out, int B, int HW,
                  cudaStream_t st);

// the two 128->3 heads over the halves of one 256-channel map in a single pass (seg: channels 0..127, dense: 128..255)
template <typename T>
void launch_head3x2(const T* x, const float* w0, const float* b0, const float* w1, const float* b1, float* out0,
                    float* out1, int B, int HW, cudaStream_t st);

// uint8 (B,H,W,3) BGR -> normalised fp32 NCHW RGB (apps/eval.py:56-61)
void launch_preprocess_u8(const unsigned char* img, float* out, int B, int H, int W, cudaStream_t st);

// ---------------------------------------------------------------- weight packing (finalize time)
void launch_fold_affine(const float* conv_bias, const float* gamma, const float* beta, const float* mean,
                        const float* var, float* scale, float* shift, int n, cudaStream_t st);
// src [Cout][Cin][kh][kw] fp32 -> dst32 [Cout][Kpad] (ky,kx,ci) fp32 and/or dst16 bf16 (same layout)
// Cout / Cin are the PACKED channel counts; Cout_src / Cin_src (0 = same) the tensor's own: the rest is zero padding
void launch_pack_conv_weight(const float* src, float* dst32, __nv_bfloat16* dst16, int Cout, int Cin, int kh, int kw,
                             int Kpad, cudaStream_t st, int Cout_src = 0, int Cin_src = 0);
// HRNet fuse layer: out = relu(sum_k nearest-upsample(terms[k], 2^shifts[k])); NHWC, C channels everywhere, up to 4 terms
template <typename T>
void launch_fuse_sum_relu(const T* const* terms, const int* shifts, int nterm, T* out, int B, int H, int W, int C,
                          cudaStream_t st);
void launch_transpose2d(const float* src, float* dst, int rows, int cols, cudaStream_t st);  // dst[c][r]=src[r][c]
void launch_transpose_pairs(const float* src /*[N][K]*/, float* dst /*[K/2][N][2]*/, int N, int K, cudaStream_t st);
void launch_gcn_adjacency(const float* e1, float* A /*21x21*/, cudaStream_t st);

// ---------------------------------------------------------------- init regressor
// attn[b][p][hand] = sigmoid(sum_c w[hand][c] * a[b][p][hand*C + c] + bias[hand]);  a: (B,P,2C)
template <typename T>
void launch_attn_logits(const T* a, const float* w /*2xC*/, const float* bias /*2*/, float* attn, int B, int P, int C,
                        cudaStream_t st);
// pooled[b][0][c] = sum_p f*attn_l/(sum attn_l+1e-8), [1] right, [2] = mean_p f.   f: (B,P,C) -> pooled (B,3,C)
template <typename T>
void launch_attn_pool(const T* f, const float* attn, float* pooled, int B, int P, int C, cudaStream_t st);

// ---------------------------------------------------------------- joint space
struct PointMlp {  // Conv1d(k=1) -> BN1d -> ReLU -> Conv1d(k=1), weights stored K-major (transposed)
  const float* w1t;  // [Cin][Cmid]
  const float* s1;   // [Cmid] folded BN scale
  const float* b1;   // [Cmid] folded BN shift (incl. conv bias)
  const float* w2t;  // [Cmid][Cout]
  const float* b2;   // [Cout]
};
struct EmbedArgs {
  const void* feat;  // NHWC (B,S,S,256)
  int S;
  const float* prev_record;  // (B, rec_stride): previous stage slice; reads joint xyz / uv
  int rec_stride;
  PointMlp filters[2];  // img2joint_{left,right}.filters 256->128->128
  PointMlp pos[2];      // pos_emb_{left,right} 3->128->128
  float* out;           // (B,2,21,128)
  int B;
  int skip_pos;         // 1: image feature only (ImgFeature2JointFeature.forward seam), no pos_emb term
};
template <typename T>
void launch_joint_embed(const EmbedArgs& a, cudaStream_t st);

// SemGCN stack, restructured for uniform work: every layer is 42 independent (B x 128)·(128 x 128) products per hand,
//   H[k][b][hand][j] = X[b][hand][j] · W[k][j]   (k = 0 self weight, k = 1 neighbour weight),
// and the graph aggregation + BN + ReLU of layer l,  X' [i] = relu(bn(H0[i] + sum_j A1[i][j] H1[j] + bias)),
// is applied on the fly when layer l+1 (or the final kernel) loads its operand.
struct GcnAgg {          // aggregation parameters of the layer that PRODUCED H
  const float* A1[2];    // per hand: (21,21) softmaxed adjacency
  const float* scale[2]; // folded BN scale (128)
  const float* shift[2]; // folded BN shift incl. gconv bias
};
struct GcnGemmArgs {
  const float* x;     // first layer: (B,2,21,128) features; otherwise nullptr
  const float* hin;   // later layers: H of the previous layer, [2][B][2][21][128]
  GcnAgg agg;         // how to turn hin into this layer's input
  float* hout;        // [2][B][2][21][128]
  const float* W[2];  // this layer's weights per hand: (2,21,128,128)
  int B;
};
void launch_gcn_gemm(const GcnGemmArgs& a, cudaStream_t st);
// tensor-core version (gcn_tc.cu, kind::tf32): per hand and layer a packed copy of gconv.W (gcn_tc_packed_bytes())
size_t gcn_tc_packed_bytes();
void launch_pack_gcn_weight_tc(const float* w /*(2,21,128,128)*/, void* wpk, cudaStream_t st);
void launch_gcn_gemm_tc(const GcnGemmArgs& a, const void* wpk_left, const void* wpk_right, cudaStream_t st);
struct GcnFinishArgs {  // tokens = relu(bn(agg(H))) + global_pos_emb(xyz/0.15 -/+ offset/2)   (models/dir.py:103-110)
  const float* hin;
  GcnAgg agg;
  PointMlp gpos;
  const float* prev_record;
  int rec_stride;
  float* y;  // (B,2,21,128) = (B,42,128) tokens
  int B;
  int skip_gpos;  // 1: y = relu(bn(agg(H))) only (ResSimplePGCN.forward seam)
};
void launch_gcn_finish(const GcnFinishArgs& a, cudaStream_t st);

struct SteWeights {
  const float* pos;  // (42,128)
  struct Block {
    // *_t weights are k-pair interleaved K-major: [K/2][N][2] (launch_transpose_pairs)
    const float *n1w, *n1b, *qkv_t, *qkv_b, *proj_t, *proj_b, *n2w, *n2b, *fc1_t, *fc1_b, *fc2_t, *fc2_b;
  } blk[3];
  const float *snw, *snb, *hnw, *hnb, *head_t, *head_b;
};
void launch_ste(const float* x, float* y, const SteWeights& w, int B, cudaStream_t st);
// tcgen05 version (ste_tc.cu, bf16 operands): `packed` = ste_tc_packed_bytes() of operand tiles filled at finalize by
// launch_pack_ste_tc(weight, block l in 0..2, which: 0 qkv 1 proj 2 fc1 3 fc2; l = 3: head.1.weight)
size_t ste_tc_packed_bytes();
void launch_pack_ste_tc(const float* src, int l, int which, void* packed, cudaStream_t st);
void launch_ste_tc(const float* x, float* y, const SteWeights& w, const void* packed, int B, cudaStream_t st);

struct ManoWeights {
  const float* comps;       // (45,45)
  const float* mean;        // (45)
  const float* shapedirs_t; // (10, 2334)
  const float* posedirs_t;  // (135, 2334)
  const float* v_template;  // (2334)
  const float* jreg;        // (16,778)
  const float* skin_w;      // (778,16)
  int tip2;                 // 444 right / 445 left
};
struct VecSeg {
  const float* p;
  int n;
  int stride;  // per-image stride in floats
};
struct RegressArgs {
  // para[hand] = Wm[hand] . concat(seg0[hand], seg1[hand]) + bm[hand]       (64 outputs)
  VecSeg in0[2], in1[2];
  const float* Wm[2];  // (64, n0+n1) row-major (PyTorch Linear layout)
  const float* bm[2];
  // offset = Wo . concat(off0, off1) + bo                                    (3 outputs)
  VecSeg off0, off1;
  const float* Wo;
  const float* bo;
  ManoWeights mano[2];
  float* stage_record;  // (B, rec_stride)
  int rec_stride;
  float* mano_para;  // (B, para_stride) -> [hand][64]
  int para_stride;
  // optional proj_feat_emb on the (21,64) token block of in0 (stage regressors only)
  int do_proj_feat;
  PointMlp proj_feat;
  float* joint_feat;  // (B,2,21,64)
  int B;
};
void launch_regress_mano(const RegressArgs& a, cudaStream_t st);
// MANO only: para (B,2,64) given
void launch_mano_only(const float* para, const ManoWeights mano[2], float* stage_record, int rec_stride, int B,
                      cudaStream_t st);

// bone rasterisation: out NHWC (B,S,S,2560) in T;  uv read from record, feat = joint_feat (B,2,21,64)
template <typename T>
void launch_bone_raster(const float* stage_record, int rec_stride, const float* joint_feat, T* out, int B, int S,
                        float distance, cudaStream_t st);
// exact factored bone_proj -> conv3x3(2560->256) -> BN -> ReLU (fusion.cu)
void launch_pack_fusion_weight(const float* w /*[256][2560][3][3]*/, float* wp /*[40][64][9][256]*/, cudaStream_t st);
void launch_bone_coef(const float* joint_feat, const float* wp, float* P /*(B,40,2,9,256)*/, int B, cudaStream_t st);
// tensor-core version (kind::tf32): wpk = bone_coef_tc_packed_bytes() filled by launch_pack_fusion_weight_tc
size_t bone_coef_tc_packed_bytes();
void launch_pack_fusion_weight_tc(const float* w /*fusion.0.weight*/, void* wpk, cudaStream_t st);
void launch_bone_coef_tc(const float* jf, const void* wpk, void* P, int p_bf16 /*P element type: 0 fp32, 1 bf16*/, int B,
                         cudaStream_t st);
// tensor-core version of the sparse accumulate (bf16 configuration; fusion.cu), S in {16, 32}
void launch_bone_fusion_tc(const float* stage_record, int rec_stride, const void* P, int p_bf16, const float* scale,
                           const float* shift, __nv_bfloat16* out, int B, int S, float distance, cudaStream_t st);
template <typename T>
void launch_bone_fusion(const float* stage_record, int rec_stride, const float* P, const float* scale,
                        const float* shift, T* out /*NHWC (B,S,S,256)*/, int B, int S, float distance, cudaStream_t st);
// vis = left + right, NCHW fp32 (B,1280,S,S); uv given per hand as (B,21,2) with image stride uv_stride
void launch_bone_vis_nchw(const float* uv_l, const float* uv_r, int uv_stride, const float* feat_l, const float* feat_r,
                          int feat_stride, float* out, int B, int S, float distance, int add_right, cudaStream_t st);

// eval metric (apps/eval.py:151-241) from the packed record (eval_metric.cu)
void launch_eval_metric(const float* record, const float* gt_verts, const float* gt_verts2d, const float* cam,
                        const float* jreg21, int B, int use_scale, float* joint_err, float* vert_err,
                        float* joint2d_err, float* vert2d_err, float* root_err, cudaStream_t st);

}  // namespace dirb200
