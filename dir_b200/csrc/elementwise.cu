// HBM-bound helpers: pooling, bilinear-upsample + concat + pre-activation, layout conversion,
// tiny 1x1 heads, attention pooling, and the finalize-time weight packing kernels.
// All feature maps are NHWC; 4 channels per thread (16 B fp32 / 8 B bf16 accesses), coalesced along C.
#include "common.cuh"
#include "kernels.h"

namespace dirb200 {

namespace {

// ---------------------------------------------------------------- vector helpers: 8 channels per thread
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// ---------------------------------------------------------------- maxpool 3x3 s2 p1 (resnet.py:247)
// one thread = one output pixel x 8 channels; C/8 consecutive threads share a pixel (coalesced 16B accesses)
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  pdl_wait();
  const int C8 = C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = (unsigned)B * Ho * Wo * C8;
  if (idx >= total) return;
  const int c = (idx % C8) * 8;
  unsigned t = idx / C8;
  const int wo = t % Wo;
  t /= Wo;
  const int ho = t % Ho;
  const int b = t / Ho;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int hi = ho * 2 - 1 + dy;
    if (hi < 0 || hi >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int wi = wo * 2 - 1 + dx;
      if (wi < 0 || wi >= W) continue;
      float v[8];
      Vec8<T>::ld(x + ((size_t)(b * H + hi) * W + wi) * C + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
  }
  Vec8<T>::st(y + ((size_t)(b * Ho + ho) * Wo + wo) * C + c, m);
}

// one bilinear sample, same association as ATen's upsample_bilinear2d (columns inside rows); the fused multiply-adds
// are spelled out so that every kernel using it produces the same bits
__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float w0x, float wx, float w0y, float wy) {
  const float top = __fmaf_rn(w0x, a, __fmul_rn(wx, b));
  const float bot = __fmaf_rn(w0x, c, __fmul_rn(wx, d));
  return __fmaf_rn(w0y, top, __fmul_rn(wy, bot));
}

// ---------------------------------------------------------------- upsample(2x bilinear, align_corners=False) alone
// (models/dir.py:443,460: the first half of the decoder's channel concats, the only half that is materialised once the
// pre-activation lives in conv1 and the pair GEMM reads the second half in place). Output rows {2g-1, 2g} and columns
// {2h-1, 2h} all interpolate between input rows (g-1, g) and columns (h-1, h): one thread loads those four 8-channel
// vectors once and writes up to four output pixels, so the L2->SM traffic is 1x the output instead of 4x
// (the per-output-pixel version ran at 1.6 TB/s of HBM traffic, limited by exactly that).
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int Hi, int Wi,
                                                         int C, unsigned groups) {
  pdl_wait();
  const unsigned grp = blockIdx.x * blockDim.y + threadIdx.y;
  if (grp >= groups) return;
  const int gw = grp % (Wi + 1);
  const unsigned t = grp / (Wi + 1);
  const int gh = t % (Hi + 1);
  const int b = t / (Hi + 1);
  const int ra = max(gh - 1, 0), rb = min(gh, Hi - 1), ca = max(gw - 1, 0), cb = min(gw, Wi - 1);
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const T* base = x + (size_t)b * Hi * Wi * C;
  const T* p00 = base + ((size_t)ra * Wi + ca) * C;
  const T* p01 = base + ((size_t)ra * Wi + cb) * C;
  const T* p10 = base + ((size_t)rb * Wi + ca) * C;
  const T* p11 = base + ((size_t)rb * Wi + cb) * C;
  T* out = y + (size_t)b * Ho * Wo * C;
  for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
    float a[8], bq[8], cc[8], d[8];
    Vec8<T>::ld(p00 + c, a);
    Vec8<T>::ld(p01 + c, bq);
    Vec8<T>::ld(p10 + c, cc);
    Vec8<T>::ld(p11 + c, d);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int ho = 2 * gh - 1 + dy;
      if (ho < 0 || ho >= Ho) continue;
      // the reference's source coordinate; its rows (y0, y1) are (ra, rb), except at ho = 0 where y1 has weight 0
      const float sy = fmaxf((ho + 0.5f) * 0.5f - 0.5f, 0.f);
      const float wy = sy - (float)(int)sy;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int wo = 2 * gw - 1 + dx;
        if (wo < 0 || wo >= Wo) continue;
        const float sx = fmaxf((wo + 0.5f) * 0.5f - 0.5f, 0.f);
        const float wx = sx - (float)(int)sx;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = bilerp(a[i], bq[i], cc[i], d[i], 1.f - wx, wx, 1.f - wy, wy);
        Vec8<T>::st(out + ((size_t)ho * Wo + wo) * C + c, v);
      }
    }
  }
}

// ---------------------------------------------------------------- upsample(2x bilinear, align_corners=False) + concat + BN/ReLU
// models/dir.py:442-444,455,459-461,470 and hourglass.py:60-61 (bn1+relu1 of the consuming Residual).
// One CTA per output pixel (all index math once per CTA, 32-bit); threads stride over 8-channel vectors.
// (Several pixels per CTA or several vectors per thread were measured and are not faster: the kernel writes 2-5x
// more than it reads and sits at the ~3.9 TB/s every write-dominated kernel of this pipeline reaches.)
template <typename T>
__global__ void __launch_bounds__(128) concat_preact_kernel(const T* __restrict__ s0, int C0, int up0,
                                                            const T* __restrict__ s1, int C1,
                                                            const float* __restrict__ bns,
                                                            const float* __restrict__ bnb, T* __restrict__ raw,
                                                            T* __restrict__ act, int Ho, int Wo) {
  pdl_wait();
  const int C = C0 + C1;
  const unsigned pix = blockIdx.x;
  const int wo = pix % Wo;
  const unsigned t = pix / Wo;
  const int ho = t % Ho;
  const int b = t / Ho;
  const T *p00 = nullptr, *p01 = nullptr, *p10 = nullptr, *p11 = nullptr;
  float wy = 0.f, wx = 0.f;
  if (up0) {
    const int Hi = Ho >> 1, Wi = Wo >> 1;
    const float sy = fmaxf((ho + 0.5f) * 0.5f - 0.5f, 0.f);
    const float sx = fmaxf((wo + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    wy = sy - y0;
    wx = sx - x0;
    const T* base = s0 + (size_t)b * Hi * Wi * C0;
    p00 = base + ((size_t)y0 * Wi + x0) * C0;
    p01 = base + ((size_t)y0 * Wi + x1) * C0;
    p10 = base + ((size_t)y1 * Wi + x0) * C0;
    p11 = base + ((size_t)y1 * Wi + x1) * C0;
  } else {
    p00 = s0 + (size_t)pix * C0;
  }
  const T* q = s1 ? s1 + (size_t)pix * C1 : nullptr;
  const float w0y = 1.f - wy, w0x = 1.f - wx;
  for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
    float v[8];
    if (c < C0) {
      if (up0) {
        float a[8], bq[8], cc[8], d[8];
        Vec8<T>::ld(p00 + c, a);
        Vec8<T>::ld(p01 + c, bq);
        Vec8<T>::ld(p10 + c, cc);
        Vec8<T>::ld(p11 + c, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = bilerp(a[i], bq[i], cc[i], d[i], w0x, wx, w0y, wy);
      } else {
        Vec8<T>::ld(p00 + c, v);
      }
    } else {
      Vec8<T>::ld(q + (c - C0), v);
    }
    if (raw) Vec8<T>::st(raw + (size_t)pix * C + c, v);
    if (act) {
      if (raw && sizeof(T) == 2) {  // pre-activation must see the value the consumer of `raw` sees
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
      }
      const float4 sa = __ldg(reinterpret_cast<const float4*>(bns + c)), sb = __ldg(reinterpret_cast<const float4*>(bns + c) + 1);
      const float4 ha = __ldg(reinterpret_cast<const float4*>(bnb + c)), hb = __ldg(reinterpret_cast<const float4*>(bnb + c) + 1);
      v[0] = fmaxf(fmaf(v[0], sa.x, ha.x), 0.f); v[1] = fmaxf(fmaf(v[1], sa.y, ha.y), 0.f);
      v[2] = fmaxf(fmaf(v[2], sa.z, ha.z), 0.f); v[3] = fmaxf(fmaf(v[3], sa.w, ha.w), 0.f);
      v[4] = fmaxf(fmaf(v[4], sb.x, hb.x), 0.f); v[5] = fmaxf(fmaf(v[5], sb.y, hb.y), 0.f);
      v[6] = fmaxf(fmaf(v[6], sb.z, hb.z), 0.f); v[7] = fmaxf(fmaf(v[7], sb.w, hb.w), 0.f);
      Vec8<T>::st(act + (size_t)pix * C + c, v);
    }
  }
}

// ---------------------------------------------------------------- input pipeline (apps/eval.py:56-61): uint8 HWC BGR ->
// RGB, /255, ImageNet mean/std, NCHW fp32. Same op order as torchvision Normalize: (x/255 - mean) / std.
__device__ __forceinline__ float normalize_px(unsigned char v, int c) {
  const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
  const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
  return ((float)v / 255.f - mean) / sd;
}

__global__ void preprocess_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int B, int HW) {
  pdl_wait();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)B * HW) return;
  const unsigned b = idx / HW, p = idx - b * HW;
  const unsigned char* px = img + (size_t)idx * 3;  // B, G, R
  float* o = out + (size_t)b * 3 * HW + p;
  o[0] = normalize_px(px[2], 0);
  o[HW] = normalize_px(px[1], 1);
  o[2 * (size_t)HW] = normalize_px(px[0], 2);
}

// ---------------------------------------------------------------- layout conversion (seam entry points, aux outputs)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int C, int HW) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? x[((int64_t)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) ActIO<T>::st(y + ((int64_t)b * HW + p) * C + c, tile[threadIdx.x][i]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int B, int C, int HW) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? ActIO<T>::ld(x + ((int64_t)b * HW + p) * C + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[((int64_t)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------- 1x1 conv to 3 channels, NCHW fp32 out (seg/dense heads)
template <typename T>
__global__ void head3_kernel(const T* __restrict__ x, int Cx, int coff, int C, const float* __restrict__ w,
                             const float* __restrict__ bias, float* __restrict__ out, int B, int HW) {
  pdl_wait();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * HW) return;
  int b = warp / HW, p = warp - b * HW;
  const T* px = x + (int64_t)warp * Cx + coff;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = ActIO<T>::ld(px + c);
    a0 = fmaf(v, __ldg(w + c), a0);
    a1 = fmaf(v, __ldg(w + C + c), a1);
    a2 = fmaf(v, __ldg(w + 2 * C + c), a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) {
    out[((int64_t)b * 3 + 0) * HW + p] = a0 + bias[0];
    out[((int64_t)b * 3 + 1) * HW + p] = a1 + bias[1];
    out[((int64_t)b * 3 + 2) * HW + p] = a2 + bias[2];
  }
}

// Both 128->3 heads of the shared seg|dense feature map (models/dir.py:404-420,474-476) in one pass: a warp per
// pixel reads its 256 channels once (16 bytes per lane); lanes 0-15 hold the seg half, 16-31 the dense half.
template <typename T>
__global__ void head3x2_kernel(const T* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ b0,
                               const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ out0,
                               float* __restrict__ out1, int B, int HW) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * HW) return;
  const int b = warp / HW, p = warp - b * HW;
  float v[8];
  Vec8<T>::ld(x + (int64_t)warp * 256 + lane * 8, v);
  const float* w = (lane < 16 ? w0 : w1) + (lane & 15) * 8;
  float a[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(w + k * 128)), u1 = __ldg(reinterpret_cast<const float4*>(w + k * 128) + 1);
    a[k] = v[0] * u0.x + v[1] * u0.y + v[2] * u0.z + v[3] * u0.w + v[4] * u1.x + v[5] * u1.y + v[6] * u1.z + v[7] * u1.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);  // within each half-warp
  }
  if ((lane & 15) == 0) {
    float* out = lane ? out1 : out0;
    const float* bias = lane ? b1 : b0;
#pragma unroll
    for (int k = 0; k < 3; ++k) out[((int64_t)b * 3 + k) * HW + p] = a[k] + bias[k];
  }
}

// ---------------------------------------------------------------- finalize-time packing
__global__ void fold_affine_kernel(const float* cb, const float* g, const float* be, const float* mu, const float* var,
                                   float* scale, float* shift, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 1.f, h = 0.f;
  if (g) {
    s = g[i] / sqrtf(var[i] + kBnEps);
    h = be[i] - mu[i] * s;
  }
  if (cb) h += cb[i] * s;
  scale[i] = s;
  shift[i] = h;
}

// Cout/Cin = packed (possibly zero-padded) channel counts, Cout_src/Cin_src = the tensor's own
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, float* __restrict__ d32,
                                        __nv_bfloat16* __restrict__ d16, int Cout, int Cin, int kh, int kw, int Kpad,
                                        int Cout_src, int Cin_src) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Cout * Kpad) return;
  int k = (int)(idx % Kpad);
  int n = (int)(idx / Kpad);
  float v = 0.f;
  if (k < kh * kw * Cin && n < Cout_src) {
    int tap = k / Cin, ci = k - tap * Cin;
    int ky = tap / kw, kx = tap - ky * kw;
    if (ci < Cin_src) v = src[(((int64_t)n * Cin_src + ci) * kh + ky) * kw + kx];
  }
  if (d32) d32[idx] = v;
  if (d16) d16[idx] = __float2bfloat16_rn(v);
}

__global__ void transpose2d_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// Linear weight [N][K] (PyTorch) -> k-pair interleaved K-major [K/2][N][2]: (w[n][2p], w[n][2p+1]) adjacent, so a
// 128-bit shared-memory load yields two (k, k+1) operand pairs for the packed FFMA2 of ring_gemm.
__global__ void transpose_pairs_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  int e = idx & 1, n = (idx >> 1) % N, p = (idx >> 1) / N;
  dst[idx] = src[(size_t)n * K + 2 * p + e];
}

// A_1 = row softmax of the 40 learned logits placed at the row-major nonzeros of the symmetric
// 21-joint skeleton adjacency (SemGCN/p_graph_conv.py:43-50, SemGCN/utils.py:27-43,66-71).
__global__ void gcn_adjacency_kernel(const float* __restrict__ e1, float* __restrict__ A) {
  __shared__ unsigned char adj[21][21];
  const int i = threadIdx.x;
  if (i < 21)
    for (int j = 0; j < 21; ++j) adj[i][j] = 0;
  __syncthreads();
  if (i < 20) {  // edges: joint c=i+1 connects to its parent
    int c = i + 1;
    int p = (c % 4 == 1) ? 0 : c - 1;
    adj[p][c] = 1;
    adj[c][p] = 1;
  }
  __syncthreads();
  if (i >= 21) return;
  int start = 0;
  for (int r = 0; r < i; ++r)
    for (int j = 0; j < 21; ++j) start += adj[r][j];
  float mx = -INFINITY;
  int k = start;
  for (int j = 0; j < 21; ++j)
    if (adj[i][j]) mx = fmaxf(mx, e1[k++]);
  float sum = 0.f;
  k = start;
  for (int j = 0; j < 21; ++j)
    if (adj[i][j]) sum += expf(e1[k++] - mx);
  k = start;
  for (int j = 0; j < 21; ++j) A[i * 21 + j] = adj[i][j] ? expf(e1[k++] - mx) / sum : 0.f;
}

// ---------------------------------------------------------------- init regressor attention (models/dir.py:231-232,263-268)
template <typename T>
__global__ void attn_logits_kernel(const T* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ attn, int BP, int C) {
  pdl_wait();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= BP * 2) return;
  int hand = warp & 1, bp = warp >> 1;
  const T* pa = a + (int64_t)bp * 2 * C + hand * C;
  const float* pw = w + hand * C;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    float4 v = ActIO<T>::ld4(pa + c);
    float4 u = __ldg(reinterpret_cast<const float4*>(pw + c));
    s += v.x * u.x + v.y * u.y + v.z * u.z + v.w * u.w;
  }
  s = warp_sum(s);
  if (lane == 0) attn[bp * 2 + hand] = 1.f / (1.f + expf(-(s + bias[hand])));
}

// grid (B, C/512), 256 threads = 64 groups of 8 channels x 4 pixel quarters: every thread streams 16-byte vectors of
// P/4 pixels (independent loads), the quarters meet in shared memory. Accumulation order over the pixels differs from
// a sequential sum only in association (fp32).
template <typename T>
__global__ void __launch_bounds__(256) attn_pool_kernel(const T* __restrict__ f, const float* __restrict__ attn,
                                                        float* __restrict__ pooled, int P, int C, int chunk) {
  pdl_wait();
  extern __shared__ float sa[];  // [P][2] attention, then [4][3][512] partial sums
  float* part = sa + P * 2;
  const int b = blockIdx.x, c0 = blockIdx.y * chunk;  // chunk = channels per block: 512, or 256 for narrow maps
  for (int i = threadIdx.x; i < P * 2; i += blockDim.x) sa[i] = attn[(int64_t)b * P * 2 + i];
  __syncthreads();
  const int g = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int pq = P >> 2;
  float al[8], ar[8], am[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) al[i] = ar[i] = am[i] = 0.f;
  const T* base = f + ((int64_t)b * P + q * pq) * C + c0 + g * 8;
#pragma unroll 4
  for (int p = 0; p < (g * 8 < chunk ? pq : 0); ++p) {
    float v[8];
    Vec8<T>::ld(base + (int64_t)p * C, v);
    const float wl = sa[(q * pq + p) * 2], wr = sa[(q * pq + p) * 2 + 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      al[i] = fmaf(v[i], wl, al[i]);
      ar[i] = fmaf(v[i], wr, ar[i]);
      am[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    part[(q * 3 + 0) * 512 + g * 8 + i] = al[i];
    part[(q * 3 + 1) * 512 + g * 8 + i] = ar[i];
    part[(q * 3 + 2) * 512 + g * 8 + i] = am[i];
  }
  __syncthreads();
  float sl = 0.f, sr = 0.f;
  for (int p = 0; p < P; ++p) {
    sl += sa[p * 2];
    sr += sa[p * 2 + 1];
  }
  for (int c = threadIdx.x; c < chunk; c += blockDim.x) {
    float tl = 0.f, tr = 0.f, tm = 0.f;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      tl += part[(qq * 3 + 0) * 512 + c];
      tr += part[(qq * 3 + 1) * 512 + c];
      tm += part[(qq * 3 + 2) * 512 + c];
    }
    pooled[((int64_t)b * 3 + 0) * C + c0 + c] = tl / (sl + 1e-8f);
    pooled[((int64_t)b * 3 + 1) * C + c0 + c] = tr / (sr + 1e-8f);
    pooled[((int64_t)b * 3 + 2) * C + c0 + c] = tm / (float)P;
  }
}

}  // namespace

template <typename T>
void launch_maxpool3x3s2(const T* x, T* y, int B, int H, int W, int C, cudaStream_t st) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  launch_pdl(maxpool_kernel<T>, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, st, x, y, B, H, W, C, Ho, Wo);
}

template <typename T>
void launch_concat_preact(const T* s0, int C0, int up0, const T* s1, int C1, const float* bns, const float* bnb, T* raw,
                          T* act, int B, int Ho, int Wo, cudaStream_t st) {
  const int C = C0 + C1;
  if (up0 && !s1 && raw && !act && C0 % 8 == 0) {  // upsample alone: four outputs per loaded 2x2 neighbourhood
    const int Hi = Ho / 2, Wi = Wo / 2;
    const int tx = C0 >= 1024 ? 128 : (C0 >= 512 ? 64 : 32), ty = 256 / tx;
    const unsigned groups = (unsigned)B * (Hi + 1) * (Wi + 1);
    launch_pdl(upsample2x_kernel<T>, dim3((groups + ty - 1) / ty), dim3(tx, ty), 0, st, s0, raw, Hi, Wi, C0, groups);
    return;
  }
  const int threads = C >= 1024 ? 128 : (C >= 512 ? 64 : 32);
  launch_pdl(concat_preact_kernel<T>, dim3((unsigned)(B * Ho * Wo)), dim3(threads), 0, st, s0, C0, up0, s1, C1, bns, bnb,
             raw, act, Ho, Wo);
}

template <typename T>
void launch_nchw_to_nhwc(const float* x, T* y, int B, int C, int H, int W, cudaStream_t st) {
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<T><<<grid, dim3(32, 8), 0, st>>>(x, y, B, C, H * W);
}

template <typename T>
void launch_nhwc_to_nchw(const T* x, float* y, int B, int C, int H, int W, cudaStream_t st) {
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B);
  nhwc_to_nchw_kernel<T><<<grid, dim3(32, 8), 0, st>>>(x, y, B, C, H * W);
}

template <typename T>
void launch_head3(const T* x, int Cx, int coff, int C, const float* w, const float* bias, float* out, int B, int HW,
                  cudaStream_t st) {
  int warps = B * HW;
  launch_pdl(head3_kernel<T>, dim3(ceil_div(warps * 32, 256)), dim3(256), 0, st, x, Cx, coff, C, w, bias, out, B, HW);
}

template <typename T>
void launch_head3x2(const T* x, const float* w0, const float* b0, const float* w1, const float* b1, float* out0,
                    float* out1, int B, int HW, cudaStream_t st) {
  launch_pdl(head3x2_kernel<T>, dim3(ceil_div(B * HW * 32, 256)), dim3(256), 0, st, x, w0, b0, w1, b1, out0, out1, B, HW);
}

void launch_preprocess_u8(const unsigned char* img, float* out, int B, int H, int W, cudaStream_t st) {
  launch_pdl(preprocess_u8_kernel, dim3(ceil_div(B * H * W, 256)), dim3(256), 0, st, img, out, B, H * W);
}

void launch_fold_affine(const float* cb, const float* g, const float* be, const float* mu, const float* var,
                        float* scale, float* shift, int n, cudaStream_t st) {
  fold_affine_kernel<<<ceil_div(n, 256), 256, 0, st>>>(cb, g, be, mu, var, scale, shift, n);
}

void launch_pack_conv_weight(const float* src, float* d32, __nv_bfloat16* d16, int Cout, int Cin, int kh, int kw,
                             int Kpad, cudaStream_t st, int Cout_src, int Cin_src) {
  int64_t total = (int64_t)Cout * Kpad;
  pack_conv_weight_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(src, d32, d16, Cout, Cin, kh, kw, Kpad,
                                                                            Cout_src > 0 ? Cout_src : Cout,
                                                                            Cin_src > 0 ? Cin_src : Cin);
}

// HRNet fuse layer: out = relu(sum_k up_nearest(term_k, 2^shift_k)), NHWC, all terms with C channels
struct FuseArgs {
  const void* term[4];
  int shift[4];
  int nterm;
  void* out;
  int B, H, W, C;
};
template <typename T>
__global__ void __launch_bounds__(256) fuse_sum_relu_kernel(FuseArgs a) {
  pdl_wait();
  // one 8-channel vector (16 bytes of bf16) per thread, 32-bit index arithmetic: the first version (4 channels, 64-bit
  // div/mod per thread) ran the 64x64 fuse of HRNet's first branch at 1.4 TB/s
  const unsigned c8n = (unsigned)a.C >> 3;
  const unsigned total = (unsigned)a.B * a.H * a.W * c8n;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const unsigned c = (idx % c8n) * 8, pix = idx / c8n;
  const unsigned w = pix % (unsigned)a.W, t = pix / (unsigned)a.W;
  const unsigned h = t % (unsigned)a.H, b = t / (unsigned)a.H;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int k = 0; k < a.nterm; ++k) {  // same term order as before: the sums are bit-identical
    const int sh = a.shift[k];
    const unsigned Hk = (unsigned)a.H >> sh, Wk = (unsigned)a.W >> sh;
    const T* p = reinterpret_cast<const T*>(a.term[k]) + ((size_t)(b * Hk + (h >> sh)) * Wk + (w >> sh)) * a.C + c;
    float v[8];
    Vec8<T>::ld(p, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = fmaxf(s[i], 0.f);
  Vec8<T>::st(reinterpret_cast<T*>(a.out) + (size_t)idx * 8, s);
}

template <typename T>
void launch_fuse_sum_relu(const T* const* terms, const int* shifts, int nterm, T* out, int B, int H, int W, int C,
                          cudaStream_t st) {
  FuseArgs a{};
  for (int k = 0; k < nterm; ++k) {
    a.term[k] = terms[k];
    a.shift[k] = shifts[k];
  }
  a.nterm = nterm;
  a.out = out;
  a.B = B; a.H = H; a.W = W; a.C = C;
  const int64_t total = (int64_t)B * H * W * (C / 8);
  launch_pdl(fuse_sum_relu_kernel<T>, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, st, a);
}
template void launch_fuse_sum_relu<float>(const float* const*, const int*, int, float*, int, int, int, int, cudaStream_t);
template void launch_fuse_sum_relu<__nv_bfloat16>(const __nv_bfloat16* const*, const int*, int, __nv_bfloat16*, int, int,
                                                  int, int, cudaStream_t);

void launch_transpose2d(const float* src, float* dst, int rows, int cols, cudaStream_t st) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose2d_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, rows, cols);
}

void launch_transpose_pairs(const float* src, float* dst, int N, int K, cudaStream_t st) {
  transpose_pairs_kernel<<<ceil_div(N * K, 256), 256, 0, st>>>(src, dst, N, K);
}

void launch_gcn_adjacency(const float* e1, float* A, cudaStream_t st) { gcn_adjacency_kernel<<<1, 32, 0, st>>>(e1, A); }

template <typename T>
void launch_attn_logits(const T* a, const float* w, const float* bias, float* attn, int B, int P, int C,
                        cudaStream_t st) {
  int warps = B * P * 2;
  launch_pdl(attn_logits_kernel<T>, dim3(ceil_div(warps * 32, 256)), dim3(256), 0, st, a, w, bias, attn, B * P, C);
}

template <typename T>
void launch_attn_pool(const T* f, const float* attn, float* pooled, int B, int P, int C, cudaStream_t st) {
  // C is a multiple of 256 and P of 4 for every caller (2048 channels, 8x8 map; models/dir.py:263-268; 256 for HRNet-W32)
  const int chunk = C % 512 == 0 ? 512 : (C % 256 == 0 ? 256 : (C <= 512 ? C : 128));  // 384 (HRNet-W48): one block per image
  launch_pdl(attn_pool_kernel<T>, dim3(B, C / chunk), dim3(256), (P * 2 + 4 * 3 * 512) * sizeof(float), st, f, attn, pooled,
             P, C, chunk);
}

#define INST(T)                                                                                                      \
  template void launch_maxpool3x3s2<T>(const T*, T*, int, int, int, int, cudaStream_t);                              \
  template void launch_concat_preact<T>(const T*, int, int, const T*, int, const float*, const float*, T*, T*, int,  \
                                        int, int, cudaStream_t);                                                     \
  template void launch_nchw_to_nhwc<T>(const float*, T*, int, int, int, int, cudaStream_t);                          \
  template void launch_nhwc_to_nchw<T>(const T*, float*, int, int, int, int, cudaStream_t);                          \
  template void launch_head3<T>(const T*, int, int, int, const float*, const float*, float*, int, int, cudaStream_t); \
  template void launch_head3x2<T>(const T*, const float*, const float*, const float*, const float*, float*, float*,  \
                                  int, int, cudaStream_t);                                                           \
  template void launch_attn_logits<T>(const T*, const float*, const float*, float*, int, int, int, cudaStream_t);    \
  template void launch_attn_pool<T>(const T*, const float*, float*, int, int, int, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
#undef INST

}  // namespace dirb200
