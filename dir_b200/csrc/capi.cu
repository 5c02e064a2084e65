// extern "C" boundary (include/dirb200.h). No exceptions cross it; errors are codes + last_error().
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "engine.h"

using namespace dirb200;

struct dirb200_handle {
  Engine e;
};

static thread_local std::string g_create_err;

#define H_CHECK(h)              \
  if (!(h)) return DIRB200_E_INVALID; \
  Engine& e = (h)->e;           \
  e.err.clear();

static int fail(Engine& e, int code, const std::string& msg) {
  e.err = msg;
  return code;
}

extern "C" int dirb200_create(const dirb200_config* cfg, dirb200_handle** out) {
  if (!cfg || !out) {
    g_create_err = "null argument";
    return DIRB200_E_INVALID;
  }
  if (cfg->precision != DIRB200_PRECISION_FP32 && cfg->precision != DIRB200_PRECISION_BF16 &&
      cfg->precision != DIRB200_PRECISION_TF32) {
    g_create_err = "unknown precision";
    return DIRB200_E_INVALID;
  }
  if (cfg->refine_stages < 0 || cfg->refine_stages > 2) {
    g_create_err = "refine_stages must be 0 (= 2), 1 or 2: the reference defines two refinement stages (models/dir.py:437-471)";
    return DIRB200_E_INVALID;
  }
  if (cfg->backbone != 0 && cfg->backbone != 32 && cfg->backbone != 48) {
    g_create_err = "backbone must be 0 (ResNet-50, the reference) or 32 / 48 (HRNet-W32 / -W48 extension)";
    return DIRB200_E_INVALID;
  }
  if (cfg->refine_stages == 1 && cfg->aux_outputs) {
    g_create_err = "refine_stages = 1 needs aux_outputs = 0 (seg/dense/proj_feat hang off the second stage)";
    return DIRB200_E_INVALID;
  }
  if (cfg->max_batch <= 0) {
    g_create_err = "max_batch must be positive";
    return DIRB200_E_INVALID;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
    g_create_err = ce != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(ce)
                                     : "device ordinal out of range";
    return DIRB200_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10) {
    g_create_err = "dirb200 is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor);
    return DIRB200_E_CUDA;
  }
  dirb200_handle* h = new (std::nothrow) dirb200_handle();
  if (!h) {
    g_create_err = "out of host memory";
    return DIRB200_E_INVALID;
  }
  h->e.cfg = *cfg;
  {
    const char* dis = getenv("DIRB200_DISABLE_TC");
    h->e.disable_tc = dis && dis[0] == '1';
    const char* nh = getenv("DIRB200_NO_HALO");
    h->e.no_halo = nh && nh[0] == '1';
    const char* nb2b = getenv("DIRB200_NO_B2B");
    h->e.no_b2b = nb2b && nb2b[0] == '1';
    const char* npf = getenv("DIRB200_NO_PREACT_FOLD");
    h->e.no_preact_fold = npf && npf[0] == '1';
    const char* f32s = getenv("DIRB200_FP32_SIMT");
    h->e.fp32_simt = f32s && f32s[0] == '1';
    const char* nopair = getenv("DIRB200_NO_PAIR_FUSION");
    h->e.disable_pair_fusion = nopair && nopair[0] == '1';
    const char* dense = getenv("DIRB200_DENSE_FUSION");
    h->e.dense_fusion = dense && dense[0] == '1';
    const char* nov = getenv("DIRB200_NO_OVERLAP");
    h->e.no_overlap = nov && nov[0] == '1';
    const char* ssplit = getenv("DIRB200_STEM_SPLIT");
    h->e.stem_split = ssplit && ssplit[0] == '1';
    const char* fsimt = getenv("DIRB200_FUSION_SIMT");
    h->e.fusion_simt = fsimt && fsimt[0] == '1';
    const char* gsimt = getenv("DIRB200_GCN_SIMT");
    h->e.gcn_simt = gsimt && gsimt[0] == '1';
    const char* csimt = getenv("DIRB200_COEF_SIMT");
    h->e.coef_simt = csimt && csimt[0] == '1';
    const char* ssimt = getenv("DIRB200_STE_SIMT");
    h->e.ste_simt = ssimt && ssimt[0] == '1';
  }
  // required-key inventory (no GPU work)
  h->e.dry = true;
  h->e.build(nullptr);
  h->e.dry = false;
  h->e.err.clear();
  *out = h;
  return DIRB200_OK;
}

extern "C" void dirb200_destroy(dirb200_handle* h) { delete h; }

extern "C" const char* dirb200_last_error(const dirb200_handle* h) { return h ? h->e.err.c_str() : g_create_err.c_str(); }

extern "C" int dirb200_set_weight(dirb200_handle* h, const char* name, const void* dev_ptr, int dtype, int ndim,
                       const int64_t* shape) {
  H_CHECK(h);
  if (!name || !dev_ptr || ndim < 0 || (ndim > 0 && !shape)) return fail(e, DIRB200_E_INVALID, "bad set_weight argument");
  if (dtype != DIRB200_DTYPE_F32 && dtype != DIRB200_DTYPE_I64) return fail(e, DIRB200_E_INVALID, "bad dtype");
  RawWeight w;
  w.p = dev_ptr;
  w.dtype = dtype;
  w.shape.assign(shape, shape + ndim);
  e.raw[name] = w;
  e.finalized = false;
  return DIRB200_OK;
}

extern "C" int dirb200_num_required_keys(const dirb200_handle* h) { return h ? (int)h->e.required.size() : 0; }

extern "C" const char* dirb200_required_key(const dirb200_handle* h, int i) {
  if (!h || i < 0 || i >= (int)h->e.required.size()) return nullptr;
  return h->e.required[i].c_str();
}

extern "C" int dirb200_finalize_weights(dirb200_handle* h, void* stream) {
  H_CHECK(h);
  cudaSetDevice(e.cfg.device);
  std::vector<std::string> req = e.required;
  for (void* p : e.owned) cudaFree(p);
  e.owned.clear();
  int rc = e.build(reinterpret_cast<cudaStream_t>(stream));
  e.required = req;
  if (rc != DIRB200_OK) return rc;
  e.finalized = true;
  return DIRB200_OK;
}

extern "C" int dirb200_workspace_bytes(const dirb200_handle* hc, int batch, size_t* bytes) {
  dirb200_handle* h = const_cast<dirb200_handle*>(hc);
  H_CHECK(h);
  if (!bytes || batch <= 0) return fail(e, DIRB200_E_INVALID, "bad argument");
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  Arena ar;
  int rc = e.bf16() ? e.forward<__nv_bfloat16>(nullptr, batch, ar, nullptr, nullptr)
                    : e.forward<float>(nullptr, batch, ar, nullptr, nullptr);
  if (rc) return rc;
  // the seam entry points stage NCHW<->NHWC copies in the same workspace: leave generous head-room
  *bytes = ar.off + (size_t)batch * 4 * (2560 * 1024 + 3 * 256 * 256) + (1 << 20);
  return DIRB200_OK;
}

extern "C" int dirb200_forward(dirb200_handle* h, const float* img, int batch, void* workspace, size_t workspace_bytes,
                    const dirb200_outputs* out, void* stream) {
  H_CHECK(h);
  if (!img || !workspace || !out || !out->record || !out->mano_para || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad forward argument");
  if (batch > e.cfg.max_batch) return fail(e, DIRB200_E_INVALID, "batch exceeds max_batch");
  if (e.cfg.aux_outputs && (!out->seg || !out->dense || !out->proj_feat))
    return fail(e, DIRB200_E_INVALID, "aux_outputs=1 needs seg/dense/proj_feat buffers");
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? e.forward<__nv_bfloat16>(img, batch, ar, out, st) : e.forward<float>(img, batch, ar, out, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

extern "C" int dirb200_forward_u8(dirb200_handle* h, const unsigned char* img_bgr, int batch, void* workspace,
                                  size_t workspace_bytes, const dirb200_outputs* out, void* stream) {
  if (!h) return DIRB200_E_INVALID;
  if (!img_bgr) {
    h->e.err = "bad forward_u8 argument";
    return DIRB200_E_INVALID;
  }
  h->e.img_u8 = img_bgr;
  int rc = dirb200_forward(h, reinterpret_cast<const float*>(img_bgr), batch, workspace, workspace_bytes, out, stream);
  h->e.img_u8 = nullptr;
  return rc;
}

extern "C" int dirb200_preprocess_u8(dirb200_handle* h, const unsigned char* img_bgr, int batch, int height, int width,
                                     float* out_nchw, void* stream) {
  H_CHECK(h);
  if (!img_bgr || !out_nchw || batch <= 0 || height <= 0 || width <= 0) return fail(e, DIRB200_E_INVALID, "bad argument");
  launch_preprocess_u8(img_bgr, out_nchw, batch, height, width, reinterpret_cast<cudaStream_t>(stream));
  return DIRB200_OK;
}

extern "C" int dirb200_forward_launches(const dirb200_handle* h, int) { return h ? h->e.last_forward_launches : 0; }

extern "C" int dirb200_profile_layer(dirb200_handle* h, const char* prefix) {
  H_CHECK(h);
  e.prof_on = prefix != nullptr;
  e.prof_prefix = prefix ? prefix : "";
  e.prof_used = 0;
  return DIRB200_OK;
}

extern "C" int dirb200_profile_read(dirb200_handle* h, float* total_ms, int* launches, double* total_flops) {
  H_CHECK(h);
  float ms = 0.f;
  double fl = 0.0;
  for (size_t i = 0; i < e.prof_used; ++i) {
    if (cudaEventSynchronize(e.prof[i].b) != cudaSuccess) return fail(e, DIRB200_E_CUDA, "event sync failed");
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.prof[i].a, e.prof[i].b) != cudaSuccess)
      return fail(e, DIRB200_E_CUDA, "event elapsed failed");
    ms += t;
    fl += e.prof[i].flops;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = (int)e.prof_used;
  if (total_flops) *total_flops = fl;
  e.prof_used = 0;
  return DIRB200_OK;
}

extern "C" int dirb200_profile_dump(dirb200_handle* h, char* buf, size_t buf_bytes) {
  H_CHECK(h);
  if (!buf || buf_bytes == 0) return fail(e, DIRB200_E_INVALID, "bad buffer");
  std::string out;
  for (size_t i = 0; i < e.prof_used; ++i) {
    if (cudaEventSynchronize(e.prof[i].b) != cudaSuccess) return fail(e, DIRB200_E_CUDA, "event sync failed");
    float t = 0.f;
    cudaEventElapsedTime(&t, e.prof[i].a, e.prof[i].b);
    const ConvLayer* L = e.prof[i].layer;
    char line[512];
    snprintf(line, sizeof line, "%s\t%d\t%dx%d\ts%d\t%d\t%d\t%.6f\t%.6e\t%.6e\n", L->name.c_str(), e.prof[i].tc,
             L->kh, L->kw, L->stride, L->Cin, L->Cout, t, e.prof[i].flops, e.prof[i].bytes);
    out += line;
  }
  if (out.size() + 1 > buf_bytes) return fail(e, DIRB200_E_INVALID, "buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return DIRB200_OK;
}

// ------------------------------------------------------------------------------------------------ seams
template <typename T>
static int seam_backbone(Engine& e, const float* img, int B, int H, int W, float* c1, float* c2, float* c3, float* c4,
                         Arena& ar, cudaStream_t st) {
  T *t1 = nullptr, *t2 = nullptr, *t3 = nullptr, *t4 = nullptr;
  if (e.hrnet() && (H != 256 || W != 256)) return DIRB200_E_INVALID;
  int rc = e.hrnet() ? e.run_backbone_hrnet<T>(img, B, ar, &t1, &t2, &t3, &t4, st)
                     : e.run_backbone<T>(img, B, H, W, ar, &t1, &t2, &t3, &t4, st);
  if (rc) return rc;
  // packed channel counts (HRNet: the 32-channel branch is stored in 64 channels, the upper half exactly zero)
  launch_nhwc_to_nchw<T>(t1, c1, B, e.c1ch, H / 4, W / 4, st);
  launch_nhwc_to_nchw<T>(t2, c2, B, e.c2ch, H / 8, W / 8, st);
  launch_nhwc_to_nchw<T>(t3, c3, B, e.c3ch, H / 16, W / 16, st);
  launch_nhwc_to_nchw<T>(t4, c4, B, e.c4ch, H / 32, W / 32, st);
  return DIRB200_OK;
}

extern "C" int dirb200_backbone(dirb200_handle* h, const float* img, int batch, int height, int width, float* c1, float* c2,
                     float* c3, float* c4, void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (!img || !c1 || !c2 || !c3 || !c4 || batch <= 0 || height % 32 || width % 32)
    return fail(e, DIRB200_E_INVALID, "bad backbone argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_backbone<__nv_bfloat16>(e, img, batch, height, width, c1, c2, c3, c4, ar, st)
                    : seam_backbone<float>(e, img, batch, height, width, c1, c2, c3, c4, ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

template <typename T>
static int seam_residual(Engine& e, const ResidualBlock& r, const float* x, int B, int H, int W, float* y, Arena& ar,
                         cudaStream_t st) {
  const int64_t n = (int64_t)B * H * W * r.cin;
  const bool fold = e.preact_fold_ok<T>(r, B, H, W, 0);  // same path as the forward: conv1 pre-activates its own operand
  T* raw = reinterpret_cast<T*>(ar.alloc(n * sizeof(T)));
  T* act = fold ? nullptr : reinterpret_cast<T*>(ar.alloc(n * sizeof(T)));
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  launch_nchw_to_nhwc<T>(x, raw, B, r.cin, H, W, st);
  if (!fold) launch_concat_preact<T>(raw, r.cin, 0, nullptr, 0, r.bn1s, r.bn1b, nullptr, act, B, H, W, st);
  T* out = e.run_residual<T>(r, raw, act, B, H, W, ar, st);
  if (!out) return DIRB200_E_WORKSPACE;
  launch_nhwc_to_nchw<T>(out, y, B, r.cout, H, W, st);
  return DIRB200_OK;
}

extern "C" int dirb200_residual(dirb200_handle* h, const char* name, const float* x, int batch, int cin, int height, int width,
                     float* y, void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  auto it = name ? e.res.find(name) : e.res.end();
  if (it == e.res.end()) return fail(e, DIRB200_E_INVALID, "unknown Residual block name");
  if (it->second.cin != cin || !x || !y) return fail(e, DIRB200_E_INVALID, "bad residual argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_residual<__nv_bfloat16>(e, it->second, x, batch, height, width, y, ar, st)
                    : seam_residual<float>(e, it->second, x, batch, height, width, y, ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

template <typename T>
static int seam_init(Engine& e, const float* c4, int B, float* rec, float* para, Arena& ar, cudaStream_t st) {
  T* x = reinterpret_cast<T*>(ar.alloc((size_t)B * 64 * e.c4ch * sizeof(T)));
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  launch_nchw_to_nhwc<T>(c4, x, B, e.c4ch, 8, 8, st);
  return e.run_init<T>(x, B, rec, DIRB200_STAGE_FLOATS, para, 128, ar, st);
}

extern "C" int dirb200_init_regressor(dirb200_handle* h, const float* c4, int batch, float* stage_record, float* mano_para,
                           void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (!c4 || !stage_record || !mano_para || batch <= 0) return fail(e, DIRB200_E_INVALID, "bad argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_init<__nv_bfloat16>(e, c4, batch, stage_record, mano_para, ar, st)
                    : seam_init<float>(e, c4, batch, stage_record, mano_para, ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

extern "C" int dirb200_mano(dirb200_handle* h, int which, const float* para, int batch, float* stage_record, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (which < 0 || which > 2 || !para || !stage_record || batch <= 0) return fail(e, DIRB200_E_INVALID, "bad argument");
  launch_mano_only(para, e.mano[which], stage_record, DIRB200_STAGE_FLOATS, batch, reinterpret_cast<cudaStream_t>(stream));
  return DIRB200_OK;
}

template <typename T>
static int seam_stage(Engine& e, int s, const float* img_feat, const float* prev_rec, const float* prev_para, int B,
                      float* rec, float* para, float* img_feat_out, float* joint_feat, float* vis, Arena& ar,
                      cudaStream_t st) {
  const int S = e.stage[s].S;
  T* x = reinterpret_cast<T*>(ar.alloc((size_t)B * S * S * 256 * sizeof(T)));
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  launch_nchw_to_nhwc<T>(img_feat, x, B, 256, S, S, st);
  T* out = nullptr;
  float* jf = nullptr;
  int rc = e.run_stage<T>(s, x, prev_rec, DIRB200_STAGE_FLOATS, prev_para, 128, B, rec, DIRB200_STAGE_FLOATS, para, 128,
                          &out, &jf, vis, ar, st);
  if (rc) return rc;
  if (vis && e.side && !e.no_overlap) cudaStreamWaitEvent(st, e.ev_join[1], 0);  // rasteriser ran on the side stream
  launch_nhwc_to_nchw<T>(out, img_feat_out, B, 256, S, S, st);
  if (joint_feat)
    cudaMemcpyAsync(joint_feat, jf, (size_t)B * 42 * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  return DIRB200_OK;
}

extern "C" int dirb200_joint2bone(dirb200_handle* h, int stage, const float* img_feat, const float* prev_record,
                       const float* prev_para, int batch, float* stage_record, float* mano_para, float* img_feat_out,
                       float* joint_feat, float* vis_img_feat, void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !img_feat || !prev_record || !prev_para || !stage_record || !mano_para ||
      !img_feat_out || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_stage<__nv_bfloat16>(e, stage - 1, img_feat, prev_record, prev_para, batch, stage_record,
                                                mano_para, img_feat_out, joint_feat, vis_img_feat, ar, st)
                    : seam_stage<float>(e, stage - 1, img_feat, prev_record, prev_para, batch, stage_record, mano_para,
                                        img_feat_out, joint_feat, vis_img_feat, ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

extern "C" int dirb200_bone_proj(dirb200_handle* h, const float* uv, const float* feat, int batch, int size, float distance,
                      float* out, void* stream) {
  H_CHECK(h);
  if (!uv || !feat || !out || batch <= 0 || size <= 0) return fail(e, DIRB200_E_INVALID, "bad argument");
  launch_bone_vis_nchw(uv, nullptr, 42, feat, nullptr, 21 * 64, out, batch, size, distance, 0,
                       reinterpret_cast<cudaStream_t>(stream));
  return DIRB200_OK;
}

// ---- seams of the joint space: ImgFeature2JointFeature, ResSimplePGCN, STE, RegressorOffset.
// Inputs arrive in the reference's own tensor layouts and are packed into the record / token layouts the kernels read.
static void scatter_rows(float* dst, int dst_stride, int dst_off, const float* src, int n, int B, cudaStream_t st) {
  cudaMemcpy2DAsync(dst + dst_off, (size_t)dst_stride * 4, src, (size_t)n * 4, (size_t)n * 4, B, cudaMemcpyDeviceToDevice,
                    st);
}

template <typename T>
static int seam_img2joint(Engine& e, int s, const float* img_feat, const float* uv_l, const float* uv_r, int B, float* out_l,
                          float* out_r, Arena& ar, cudaStream_t st) {
  const StageWeights& sw = e.stage[s];
  const int S = sw.S;
  T* x = reinterpret_cast<T*>(ar.alloc((size_t)B * S * S * 256 * sizeof(T)));
  float* rec = reinterpret_cast<float*>(ar.alloc((size_t)B * DIRB200_STAGE_FLOATS * 4));
  float* out = reinterpret_cast<float*>(ar.alloc((size_t)B * 42 * 128 * 4));
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  launch_nchw_to_nhwc<T>(img_feat, x, B, 256, S, S, st);
  cudaMemsetAsync(rec, 0, (size_t)B * DIRB200_STAGE_FLOATS * 4, st);
  scatter_rows(rec, DIRB200_STAGE_FLOATS, DIRB200_OFF_UV_L, uv_l, 42, B, st);
  scatter_rows(rec, DIRB200_STAGE_FLOATS, DIRB200_OFF_UV_R, uv_r, 42, B, st);
  EmbedArgs a{};
  a.feat = x;
  a.S = S;
  a.prev_record = rec;
  a.rec_stride = DIRB200_STAGE_FLOATS;
  for (int h = 0; h < 2; ++h) {
    a.filters[h] = sw.filters[h];
    a.pos[h] = sw.pos[h];
  }
  a.out = out;
  a.B = B;
  a.skip_pos = 1;
  launch_joint_embed<T>(a, st);
  cudaMemcpy2DAsync(out_l, 21 * 128 * 4, out, 42 * 128 * 4, 21 * 128 * 4, B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpy2DAsync(out_r, 21 * 128 * 4, out + 21 * 128, 42 * 128 * 4, 21 * 128 * 4, B, cudaMemcpyDeviceToDevice, st);
  return DIRB200_OK;
}

extern "C" int dirb200_img2joint(dirb200_handle* h, int stage, const float* img_feat, const float* uv_left,
                                 const float* uv_right, int batch, float* out_left, float* out_right, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !img_feat || !uv_left || !uv_right || !out_left || !out_right || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad img2joint argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_img2joint<__nv_bfloat16>(e, stage - 1, img_feat, uv_left, uv_right, batch, out_left, out_right,
                                                    ar, st)
                    : seam_img2joint<float>(e, stage - 1, img_feat, uv_left, uv_right, batch, out_left, out_right, ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

extern "C" int dirb200_gcn(dirb200_handle* h, int stage, const float* x_left, const float* x_right, int batch,
                           float* y_left, float* y_right, void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !x_left || !x_right || !y_left || !y_right || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad gcn argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t tok = (size_t)batch * 42 * 128;
  float* x = reinterpret_cast<float*>(ar.alloc(tok * 4));
  float* y = reinterpret_cast<float*>(ar.alloc(tok * 4));
  float* gh0 = reinterpret_cast<float*>(ar.alloc(2 * tok * 4));
  float* gh1 = reinterpret_cast<float*>(ar.alloc(2 * tok * 4));
  if (ar.overflow) return fail(e, DIRB200_E_WORKSPACE, "workspace too small");
  cudaMemcpy2DAsync(x, 42 * 128 * 4, x_left, 21 * 128 * 4, 21 * 128 * 4, batch, cudaMemcpyDeviceToDevice, st);
  cudaMemcpy2DAsync(x + 21 * 128, 42 * 128 * 4, x_right, 21 * 128 * 4, 21 * 128 * 4, batch, cudaMemcpyDeviceToDevice, st);
  e.run_gcn(stage - 1, e.bf16(), x, gh0, gh1, nullptr, 0, y, batch, /*skip_gpos=*/1, st);
  cudaMemcpy2DAsync(y_left, 21 * 128 * 4, y, 42 * 128 * 4, 21 * 128 * 4, batch, cudaMemcpyDeviceToDevice, st);
  cudaMemcpy2DAsync(y_right, 21 * 128 * 4, y + 21 * 128, 42 * 128 * 4, 21 * 128 * 4, batch, cudaMemcpyDeviceToDevice, st);
  return DIRB200_OK;
}

extern "C" int dirb200_ste(dirb200_handle* h, int stage, const float* x, int batch, float* y, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !x || !y || batch <= 0) return fail(e, DIRB200_E_INVALID, "bad ste argument");
  const StageWeights& sw = e.stage[stage - 1];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (e.bf16() && sw.ste_packed && !e.ste_simt)
    launch_ste_tc(x, y, sw.ste, sw.ste_packed, batch, st);
  else
    launch_ste(x, y, sw.ste, batch, st);
  return DIRB200_OK;
}

extern "C" int dirb200_regressor_offset(dirb200_handle* h, int stage, const float* feat_left, const float* feat_right,
                                        const float* para_left, const float* para_right, const float* offset, int batch,
                                        float* stage_record, float* mano_para, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !feat_left || !feat_right || !para_left || !para_right || !offset || !stage_record ||
      !mano_para || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad regressor_offset argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = batch, s = stage - 1;
  float* tok = reinterpret_cast<float*>(ar.alloc((size_t)B * 2688 * 4));
  float* prev_para = reinterpret_cast<float*>(ar.alloc((size_t)B * 128 * 4));
  float* prev_rec = reinterpret_cast<float*>(ar.alloc((size_t)B * DIRB200_STAGE_FLOATS * 4));
  if (ar.overflow) return fail(e, DIRB200_E_WORKSPACE, "workspace too small");
  scatter_rows(tok, 2688, 0, feat_left, 1344, B, st);
  scatter_rows(tok, 2688, 1344, feat_right, 1344, B, st);
  scatter_rows(prev_para, 128, 0, para_left, 64, B, st);
  scatter_rows(prev_para, 128, 64, para_right, 64, B, st);
  cudaMemsetAsync(prev_rec, 0, (size_t)B * DIRB200_STAGE_FLOATS * 4, st);
  scatter_rows(prev_rec, DIRB200_STAGE_FLOATS, DIRB200_OFF_OFFSET, offset, 3, B, st);
  const StageWeights& sw = e.stage[s];
  RegressArgs a{};
  for (int hd = 0; hd < 2; ++hd) {  // models/dir.py:344-347: [feat | prev_para] per hand, [feat_l | feat_r | offset]
    a.in0[hd] = VecSeg{tok + hd * 1344, 1344, 2688};
    a.in1[hd] = VecSeg{prev_para + hd * 64, 64, 128};
    a.Wm[hd] = sw.Wm[hd];
    a.bm[hd] = sw.bm[hd];
    a.mano[hd] = e.mano[s + 1][hd];
  }
  a.off0 = VecSeg{tok, 2688, 2688};
  a.off1 = VecSeg{prev_rec + DIRB200_OFF_OFFSET, 3, DIRB200_STAGE_FLOATS};
  a.Wo = sw.Wo;
  a.bo = sw.bo;
  a.stage_record = stage_record;
  a.rec_stride = DIRB200_STAGE_FLOATS;
  a.mano_para = mano_para;
  a.para_stride = 128;
  a.do_proj_feat = 0;
  a.B = B;
  launch_regress_mano(a, st);
  return DIRB200_OK;
}

template <typename T>
static int seam_bone_fusion(Engine& e, int s, const float* uv_l, const float* uv_r, const float* feat_l, const float* feat_r,
                            int B, float* out_nchw, Arena& ar, cudaStream_t st) {
  const int S = e.stage[s].S;
  float* rec = reinterpret_cast<float*>(ar.alloc((size_t)B * DIRB200_STAGE_FLOATS * 4));
  float* jf = reinterpret_cast<float*>(ar.alloc((size_t)B * 42 * 64 * 4));
  T* bone = e.dense_fusion ? reinterpret_cast<T*>(ar.alloc((size_t)B * S * S * 2560 * sizeof(T))) : nullptr;
  float* coef = e.dense_fusion ? nullptr : reinterpret_cast<float*>(ar.alloc((size_t)B * 40 * 2 * 9 * 256 * 4));
  T* mid = reinterpret_cast<T*>(ar.alloc((size_t)B * S * S * 256 * sizeof(T)));
  T* out = reinterpret_cast<T*>(ar.alloc((size_t)B * S * S * 256 * sizeof(T)));
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  cudaMemsetAsync(rec, 0, (size_t)B * DIRB200_STAGE_FLOATS * 4, st);
  scatter_rows(rec, DIRB200_STAGE_FLOATS, DIRB200_OFF_UV_L, uv_l, 42, B, st);
  scatter_rows(rec, DIRB200_STAGE_FLOATS, DIRB200_OFF_UV_R, uv_r, 42, B, st);
  scatter_rows(jf, 42 * 64, 0, feat_l, 21 * 64, B, st);
  scatter_rows(jf, 42 * 64, 21 * 64, feat_r, 21 * 64, B, st);
  e.sticky_rc = 0;
  e.run_bone_fusion<T>(s, rec, DIRB200_STAGE_FLOATS, jf, B, bone, coef, mid, out, st);
  if (e.sticky_rc) return e.sticky_rc;
  launch_nhwc_to_nchw<T>(out, out_nchw, B, 256, S, S, st);
  return DIRB200_OK;
}

extern "C" int dirb200_bone_fusion(dirb200_handle* h, int stage, const float* uv_left, const float* uv_right,
                                   const float* feat_left, const float* feat_right, int batch, float* img_feat_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  if (stage < 1 || stage > 2 || !uv_left || !uv_right || !feat_left || !feat_right || !img_feat_out || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad bone_fusion argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_bone_fusion<__nv_bfloat16>(e, stage - 1, uv_left, uv_right, feat_left, feat_right, batch,
                                                      img_feat_out, ar, st)
                    : seam_bone_fusion<float>(e, stage - 1, uv_left, uv_right, feat_left, feat_right, batch, img_feat_out,
                                              ar, st);
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

template <typename T>
static int seam_conv(Engine& e, const ConvLayer& L, const float* x, const float* res, int B, int H, int W, float* y,
                     Arena& ar, cudaStream_t st) {
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W + 2 * L.pad - L.kw) / L.stride + 1;
  T* xi = reinterpret_cast<T*>(ar.alloc((size_t)B * H * W * L.Cin * sizeof(T)));
  T* yo = reinterpret_cast<T*>(ar.alloc((size_t)B * Ho * Wo * L.Cout * sizeof(T)));
  T* ri = res ? reinterpret_cast<T*>(ar.alloc((size_t)B * Ho * Wo * L.Cout * sizeof(T))) : nullptr;
  if (ar.overflow) return DIRB200_E_WORKSPACE;
  launch_nchw_to_nhwc<T>(x, xi, B, L.Cin, H, W, st);
  if (res) launch_nchw_to_nhwc<T>(res, ri, B, L.Cout, Ho, Wo, st);
  e.sticky_rc = 0;
  e.tc_launches = 0;
  e.conv<T>(L, xi, yo, ri, B, H, W, st);
  if (e.sticky_rc) return e.sticky_rc;
  launch_nhwc_to_nchw<T>(yo, y, B, L.Cout, Ho, Wo, st);
  return DIRB200_OK;
}

extern "C" int dirb200_conv_layer(dirb200_handle* h, const char* weight_key, const float* x, const float* res,
                                  int batch, int height, int width, float* y, int* used_tensor_cores, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  H_CHECK(h);
  if (!e.finalized) return fail(e, DIRB200_E_STATE, "finalize_weights first");
  const ConvLayer* L = weight_key ? e.find_conv(weight_key) : nullptr;
  if (!L) return fail(e, DIRB200_E_INVALID, "unknown conv weight key");
  if (!x || !y || batch <= 0 || L->Cin % 4 != 0) return fail(e, DIRB200_E_INVALID, "bad conv_layer argument");
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  ar.size = workspace_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = e.bf16() ? seam_conv<__nv_bfloat16>(e, *L, x, res, batch, height, width, y, ar, st)
                    : seam_conv<float>(e, *L, x, res, batch, height, width, y, ar, st);
  if (used_tensor_cores) *used_tensor_cores = e.tc_launches;
  if (rc == DIRB200_E_WORKSPACE) e.err = "workspace too small";
  return rc;
}

extern "C" int dirb200_eval_metrics(dirb200_handle* h, const float* record, const float* gt_verts,
                                    const float* gt_verts2d, const float* cam, const float* jreg21, int batch,
                                    int use_scale, float* joint_err, float* vert_err, float* joint2d_err,
                                    float* vert2d_err, float* root_err, void* stream) {
  H_CHECK(h);
  if (!record || !gt_verts || !gt_verts2d || !cam || !jreg21 || !joint_err || !vert_err || !joint2d_err || !vert2d_err ||
      !root_err || batch <= 0)
    return fail(e, DIRB200_E_INVALID, "bad eval_metrics argument");
  launch_eval_metric(record, gt_verts, gt_verts2d, cam, jreg21, batch, use_scale, joint_err, vert_err, joint2d_err,
                     vert2d_err, root_err, reinterpret_cast<cudaStream_t>(stream));
  return DIRB200_OK;
}

// ------------------------------------------------------------------------------------------------ NCCL (run-time bound)
namespace {
struct NcclId {
  char internal[128];
};
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
struct NcclApi {
  void* lib = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allgather allgather = nullptr;
  fn_destroy destroy = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
    get_id = (fn_get_id)dlsym(lib, "ncclGetUniqueId");
    init_rank = (fn_init_rank)dlsym(lib, "ncclCommInitRank");
    allgather = (fn_allgather)dlsym(lib, "ncclAllGather");
    destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
    if (!get_id || !init_rank || !allgather || !destroy) {
      err = "libnccl.so.2 lacks an expected symbol";
      return false;
    }
    return true;
  }
} g_nccl;
}  // namespace

extern "C" int dirb200_nccl_unique_id(dirb200_handle* h, char id_out[128]) {
  H_CHECK(h);
  if (!id_out) return fail(e, DIRB200_E_INVALID, "null id buffer");
  if (!g_nccl.load(e.err)) return DIRB200_E_STATE;
  NcclId id;
  if (g_nccl.get_id(&id) != 0) return fail(e, DIRB200_E_CUDA, "ncclGetUniqueId failed");
  memcpy(id_out, id.internal, 128);
  return DIRB200_OK;
}

extern "C" int dirb200_nccl_init(dirb200_handle* h, const char id[128], int rank, int world) {
  H_CHECK(h);
  if (!id || rank < 0 || rank >= world) return fail(e, DIRB200_E_INVALID, "bad nccl_init argument");
  if (!g_nccl.load(e.err)) return DIRB200_E_STATE;
  cudaSetDevice(e.cfg.device);
  NcclId nid;
  memcpy(nid.internal, id, 128);
  if (e.nccl_comm) {  // re-initialisation: the previous communicator is released first
    g_nccl.destroy(e.nccl_comm);
    e.nccl_comm = nullptr;
  }
  void* comm = nullptr;
  if (g_nccl.init_rank(&comm, world, nid, rank) != 0) return fail(e, DIRB200_E_CUDA, "ncclCommInitRank failed");
  e.nccl_comm = comm;
  e.nccl_destroy = [](void* c) { g_nccl.destroy(c); };
  return DIRB200_OK;
}

extern "C" int dirb200_allgather_records(dirb200_handle* h, const float* send, float* recv, int batch_local, void* stream) {
  H_CHECK(h);
  if (!e.nccl_comm) return fail(e, DIRB200_E_STATE, "nccl_init first");
  if (!send || !recv || batch_local <= 0) return fail(e, DIRB200_E_INVALID, "bad allgather argument");
  const size_t count = (size_t)batch_local * DIRB200_RECORD_FLOATS;
  if (g_nccl.allgather(send, recv, count, /*ncclFloat32=*/7, e.nccl_comm, reinterpret_cast<cudaStream_t>(stream)) != 0)
    return fail(e, DIRB200_E_CUDA, "ncclAllGather failed");
  return DIRB200_OK;
}

