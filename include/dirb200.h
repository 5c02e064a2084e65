/*
 * dirb200.h — C ABI of the B200-native DIR inference hot path.
 *
 * The reference (PengfeiRen96/DIR) has no FFI/plugin layer: its only seam is the
 * nn.Module interface `DIR.forward(input, target, meta_info)` (models/dir.py:513-540)
 * plus its state_dict key set. This header is the boundary a host binds instead of
 * that module's PyTorch ops (the Python host in dir_b200/module.py binds it with
 * ctypes; INTEGRATION.md shows the stub). Each entry point cites what it replaces.
 *
 * Conventions: every function returns 0 on success or a negative DIRB200_E_* code and
 * never throws; dirb200_last_error() gives the message. All device buffers (weights in,
 * image in, workspace, outputs) are allocated and owned by the CALLER; the library only
 * allocates its own packed-weight copies inside dirb200_finalize_weights(). Every
 * compute call is asynchronous on the given CUDA stream, performs no host sync and is
 * CUDA-graph capturable. One handle per device; a handle is not thread-safe.
 * `stream` is a cudaStream_t passed as void*.
 */
#ifndef DIRB200_H_
#define DIRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIRB200_OK 0
#define DIRB200_E_INVALID (-1)   /* bad argument / unknown name / shape mismatch */
#define DIRB200_E_STATE (-2)     /* call order violated (e.g. forward before finalize) */
#define DIRB200_E_CUDA (-3)      /* a CUDA runtime/driver call failed */
#define DIRB200_E_MISSING (-4)   /* finalize: a required state_dict key was never set */
#define DIRB200_E_WORKSPACE (-5) /* workspace too small */

#define DIRB200_PRECISION_FP32 0 /* parity config: fp32 activations; convs on tcgen05 as error-compensated 3xTF32 with
                                    round-to-nearest fp32 accumulation (<= the error of an fp32 cuDNN/MKL conv) */
#define DIRB200_PRECISION_BF16 1 /* bf16 feature maps, tcgen05 bf16 MMA with fp32 accumulate; joint space fp32 */
#define DIRB200_PRECISION_TF32 2 /* fp32 activations, plain TF32 tcgen05 convs: the arithmetic PyTorch's cuDNN default
                                    (torch.backends.cudnn.allow_tf32 = True) gives the reference on this GPU */

#define DIRB200_DTYPE_F32 0
#define DIRB200_DTYPE_I64 1

/* floats per image in the packed output record: 3 stages x 4887 */
#define DIRB200_STAGE_FLOATS 4887
#define DIRB200_RECORD_FLOATS (3 * DIRB200_STAGE_FLOATS)
/* offsets inside one stage of the record */
#define DIRB200_OFF_MESH_L 0
#define DIRB200_OFF_MESH_R 2334
#define DIRB200_OFF_JOINT_L 4668
#define DIRB200_OFF_JOINT_R 4731
#define DIRB200_OFF_UV_L 4794
#define DIRB200_OFF_UV_R 4836
#define DIRB200_OFF_PROJ_L 4878
#define DIRB200_OFF_PROJ_R 4881
#define DIRB200_OFF_OFFSET 4884

typedef struct dirb200_handle dirb200_handle;

typedef struct dirb200_config {
  int precision;   /* DIRB200_PRECISION_* */
  int max_batch;   /* largest per-call batch the handle will be asked for */
  int aux_outputs; /* 1: also compute seg/dense/proj_feat (models/dir.py:474-482,536-540) */
  int device;      /* CUDA device ordinal */
  int refine_stages; /* 0 or 2: both refinement stages, i.e. the reference forward (stage_num = 3, models/dir.py:437-471);
                        1: stop after projecter_4 (init regression + one refinement, BASELINE.json configs[0] "1 refine
                        iter"; needs aux_outputs = 0, the stage-2 slice of the record is zero-filled). The reference has
                        no further stages, so larger values are rejected. */
  int backbone;    /* 0: ResNet-50, the reference's only backbone (models/dir.py:492). 32 / 48: HRNet-W32 / -W48 — an
                      EXTENSION with no counterpart in the reference (BASELINE.json configs 3-5): the
                      published HRNet backbone with the reference's decoder knobs inDim=[8w,4w,2w,w],
                      InitRegressor(feat_dim=8w) (models/dir.py:390,501). Parity unpinned: tested against the self-authored oracle/hrnet_oracle.py. */
} dirb200_config;

/* Caller-owned output buffers of one forward (device pointers). */
typedef struct dirb200_outputs {
  float* record;    /* (B, DIRB200_RECORD_FLOATS): per image, 3 stages x [mesh_l mesh_r joint_l joint_r uv_l uv_r
                       proj_l proj_r offset] == the 9 tensors per stage of outs_list[0..2] (models/dir.py:521-535) */
  float* mano_para; /* (B, 3, 2, 64): pd_mano_para_{left,right} per stage (models/dir.py:291-292,370-371) */
  float* seg;       /* (B,3,32,32) NCHW or NULL when aux_outputs=0 */
  float* dense;     /* (B,3,32,32) NCHW or NULL */
  float* proj_feat; /* (B,1280,32,32) NCHW or NULL */
} dirb200_outputs;

/* Lifetime. Replaces DIR.__init__ (models/dir.py:487-511) minus weight creation. */
int dirb200_create(const dirb200_config* cfg, dirb200_handle** out);
void dirb200_destroy(dirb200_handle* h);
const char* dirb200_last_error(const dirb200_handle* h); /* h may be NULL: last create() error */

/* Weights. Replaces nn.Module.load_state_dict (apps/eval.py:107-108): call set_weight once per
 * state_dict key (reference key names, fp32 or int64 device tensors, contiguous, PyTorch layout),
 * then finalize_weights, which folds BN, repacks to kernel layouts, precomputes the SemGCN softmax
 * adjacency and verifies key coverage (strict). Source tensors may be freed after finalize returns
 * and the stream is synchronised by the caller. */
int dirb200_set_weight(dirb200_handle* h, const char* name, const void* dev_ptr, int dtype, int ndim,
                       const int64_t* shape);
int dirb200_finalize_weights(dirb200_handle* h, void* stream);
int dirb200_num_required_keys(const dirb200_handle* h);
const char* dirb200_required_key(const dirb200_handle* h, int i);

/* Whole forward. Replaces DIR.forward's eval branch (models/dir.py:513-540).
 * img: (B,3,256,256) fp32 NCHW, ImageNet-normalised, device memory. */
int dirb200_workspace_bytes(const dirb200_handle* h, int batch, size_t* bytes);
int dirb200_forward(dirb200_handle* h, const float* img, int batch, void* workspace, size_t workspace_bytes,
                    const dirb200_outputs* out, void* stream);
/* Same forward fed with raw frames: img_bgr (B,256,256,3) uint8, HWC, BGR (what cv2 hands to apps/eval.py:56-61). The
 * reference's host-side preprocessing (BGR->RGB, /255, ImageNet mean/std, HWC->CHW) runs on the device, fused into
 * the stem operand packing in the bf16 configuration. 4x fewer host->device bytes than the fp32 image. */
int dirb200_forward_u8(dirb200_handle* h, const unsigned char* img_bgr, int batch, void* workspace,
                       size_t workspace_bytes, const dirb200_outputs* out, void* stream);
/* The preprocessing alone: (B,H,W,3) uint8 BGR -> (B,3,H,W) fp32 normalised RGB (apps/eval.py:56-61). */
int dirb200_preprocess_u8(dirb200_handle* h, const unsigned char* img_bgr, int batch, int height, int width,
                          float* out_nchw, void* stream);
/* number of kernels one dirb200_forward(batch) enqueues (for launch accounting) */
int dirb200_forward_launches(const dirb200_handle* h, int batch);

/* Kernel timing hook (bench.py's roofline line): CUDA events are recorded on the launch stream around every
 * conv launch whose weight key starts with `prefix` (e.g. "decoder.projecter_3.fusion.0"; "" = all convs;
 * NULL disables). profile_read synchronises those events, returns their summed duration, the number of launches
 * and their algorithmic FLOPs (2*M*N*K) since the last read, and resets. Not usable under graph capture. */
int dirb200_profile_layer(dirb200_handle* h, const char* prefix);
int dirb200_profile_read(dirb200_handle* h, float* total_ms, int* launches, double* total_flops);
/* Per-launch detail of the records accumulated since the last read, one line per conv launch:
 * "weight_key \t tensor_cores \t KhxKw \t sS \t Cin \t Cout \t ms \t flops \t compulsory_bytes". Does not reset. */
int dirb200_profile_dump(dirb200_handle* h, char* buf, size_t buf_bytes);

/* Per-seam entry points; used by the parity tests. All tensors fp32 device memory,
 * feature maps NCHW exactly as the reference module sees them; conversion to the internal NHWC /
 * bf16 layout happens inside, in `workspace`. */
/* ResNet.forward (models/backbone/resnet.py:243-255): img (B,3,H,W) -> c1..c4 NCHW fp32. HRNet handles:
 * 256x256 only; backbone = 32: c1..c4 = (B,64,64,64) [32 channels + 32 zero-padded], (B,64,32,32), (B,128,16,16),
 * (B,256,8,8); backbone = 48: (B,64,64,64) [48 + 16 zeros], (B,128,32,32) [96 + 32 zeros], (B,192,16,16), (B,384,8,8). */
int dirb200_backbone(dirb200_handle* h, const float* img, int batch, int height, int width, float* c1, float* c2,
                     float* c3, float* c4, void* workspace, size_t workspace_bytes, void* stream);
/* Residual.forward (models/backbone/hourglass.py:55-70); name = "decoder.enhance_layer4." etc. */
int dirb200_residual(dirb200_handle* h, const char* name, const float* x, int batch, int cin, int height, int width,
                     float* y, void* workspace, size_t workspace_bytes, void* stream);
/* InitRegressor.forward (models/dir.py:260-305): c4 (B,2048,8,8) -> stage-0 slice of record + mano_para */
int dirb200_init_regressor(dirb200_handle* h, const float* c4, int batch, float* stage_record /*(B,4887)*/,
                           float* mano_para /*(B,2,64)*/, void* workspace, size_t workspace_bytes, void* stream);
/* manopth ManoLayer.forward + projection_batch_xy (manopth/manopth/manolayer.py:110-270, utils/utils.py:47-63):
 * para (B,2,64) = [pose51|beta10|proj3] per hand -> stage slice of the record. `which` = 0 init_regressor,
 * 1 projecter_4.regressor, 2 projecter_3.regressor (six buffer copies live in the state_dict). */
int dirb200_mano(dirb200_handle* h, int which, const float* para, int batch, float* stage_record, void* stream);
/* Joint2BoneFeature.forward (models/dir.py:86-130); stage = 1 (projecter_4, S=16) or 2 (projecter_3, S=32).
 * img_feat (B,256,S,S); prev_record (B,4887) and prev_para (B,2,64) from the previous stage.
 * Outputs: stage_record (B,4887), mano_para (B,2,64), img_feat_out (B,256,S,S), joint_feat (B,2,21,64),
 * vis_img_feat (B,1280,S,S) or NULL. */
int dirb200_joint2bone(dirb200_handle* h, int stage, const float* img_feat, const float* prev_record,
                       const float* prev_para, int batch, float* stage_record, float* mano_para, float* img_feat_out,
                       float* joint_feat, float* vis_img_feat, void* workspace, size_t workspace_bytes, void* stream);
/* ImgFeature2JointFeature.forward (models/dir.py:197-200) of both hands of a stage: F.grid_sample (bilinear, zeros,
 * align_corners=False) + the `filters` point-MLP. img_feat (B,256,S,S); uv (B,21,2) in [-1,1]; out (B,21,128) = the
 * reference's (B,128*21) viewed (B,128,21) and permuted (models/dir.py:94-95). */
int dirb200_img2joint(dirb200_handle* h, int stage, const float* img_feat, const float* uv_left, const float* uv_right,
                      int batch, float* out_left, float* out_right, void* workspace, size_t workspace_bytes,
                      void* stream);
/* ResSimplePGCN.forward (SemGCN/p_gcn.py:63-73; gcn_left / gcn_right of a stage): four PGraphConv + BN + ReLU layers
 * (SemGCN/p_graph_conv.py:39-60). x, y (B,21,128). bf16 handles run the tf32 tcgen05 GEMMs (gcn_tc.cu). */
int dirb200_gcn(dirb200_handle* h, int stage, const float* x_left, const float* x_right, int batch, float* y_left,
                float* y_right, void* workspace, size_t workspace_bytes, void* stream);
/* STE.forward (transformer/mixSTE.py:194-205; `interaction` of a stage): x (B,42,128) -> y (B,42,64). The reference
 * adds the position embedding into x in place; here x is read-only. bf16 handles run the tcgen05 kernel (ste_tc.cu). */
int dirb200_ste(dirb200_handle* h, int stage, const float* x, int batch, float* y, void* stream);
/* RegressorOffset.forward (models/dir.py:339-381): joint features (B,21,64) per hand, previous MANO parameters (B,64)
 * per hand and previous offset (B,3) -> stage slice of the record (MANO layers + projection included) + mano_para. */
int dirb200_regressor_offset(dirb200_handle* h, int stage, const float* feat_left, const float* feat_right,
                             const float* para_left, const float* para_right, const float* offset, int batch,
                             float* stage_record, float* mano_para, void* workspace, size_t workspace_bytes, void* stream);
/* Joint2BoneFeature.bone_proj (models/dir.py:146-174): uv (B,21,2), feat (B,21,64) -> (B,1280,S,S) */
int dirb200_bone_proj(dirb200_handle* h, const float* uv, const float* feat, int batch, int size, float distance,
                      float* out, void* stream);

/* The image-space half of Joint2BoneFeature.forward (models/dir.py:118-122): bone_proj of both hands (:146-174), channel
 * concat, `fusion` = conv3x3(2560->256) + BN + ReLU + conv1x1 (:57-62). uv (B,21,2) and joint features (B,21,64) per
 * hand (the output of proj_feat_emb) -> img_feat (B,256,S,S) NCHW. Runs the exact factored form (fusion.cu): fp32
 * CUDA-core kernels on fp32 handles, the tcgen05 coefficient + accumulate kernels on bf16 handles. */
int dirb200_bone_fusion(dirb200_handle* h, int stage, const float* uv_left, const float* uv_right,
                        const float* feat_left, const float* feat_right, int batch, float* img_feat_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* One nn.Conv2d (+ its folded eval BatchNorm / bias, optional residual add, ReLU exactly as fused in the
 * forward) by the state_dict key of its weight, e.g. "backbone.layer2.0.conv2.weight"
 * (models/backbone/resnet.py:120-140). x (B,Cin,H,W), res/y (B,Cout,Ho,Wo) NCHW fp32. *used_tensor_cores is
 * set to 1 when the tcgen05 kernel ran (bf16 handles, tensor-core shaped layers), 0 for the CUDA-core kernel. */
int dirb200_conv_layer(dirb200_handle* h, const char* weight_key, const float* x, const float* res, int batch,
                       int height, int width, float* y, int* used_tensor_cores, void* workspace,
                       size_t workspace_bytes, void* stream);

/* The evaluation metric of apps/eval.py:151-241 (root_joint = 0) computed on the device from the packed record of
 * dirb200_forward: joints regressed from vertices with the 21x778 regressor of class Jr (apps/eval.py:22-44), wrist
 * alignment, optional |j9-j0| scale alignment (opt.scale), L2 errors in metres and re-projection errors in pixels.
 * record (B,RECORD); gt_verts (B,2,778,3); gt_verts2d (B,2,778,2); cam (B,3,3); jreg21 (2,21,778) [left,right].
 * Outputs: joint_err, joint2d_err (B,2,21); vert_err, vert2d_err (B,2,778); root_err (B). */
int dirb200_eval_metrics(dirb200_handle* h, const float* record, const float* gt_verts, const float* gt_verts2d,
                         const float* cam, const float* jreg21, int batch, int use_scale, float* joint_err,
                         float* vert_err, float* joint2d_err, float* vert2d_err, float* root_err, void* stream);

/* Multi-GPU: images shard over ranks with no exchange inside the forward; the only
 * collective is one all-gather of the per-image records over NVLink (the reference has no distributed
 * code at all). The library resolves NCCL at run time (dlopen of the libnccl the host process already
 * loaded); rank 0 creates the id, the host broadcasts the 128 bytes, every rank calls nccl_init.
 * send (B_local, RECORD) -> recv (world*B_local, RECORD), rank-major. */
int dirb200_nccl_unique_id(dirb200_handle* h, char id_out[128]);
int dirb200_nccl_init(dirb200_handle* h, const char id[128], int rank, int world);
int dirb200_allgather_records(dirb200_handle* h, const float* send, float* recv, int batch_local, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIRB200_H_ */
