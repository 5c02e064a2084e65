// HBM-bound helpers: pooling, bilinear-upsample + concat + pre-activation, layout conversion,
// tiny 1x1 heads, attention pooling, and the finalize-time weight packing kernels.
// All feature maps are NHWC; 4 channels per thread (16 B fp32 / 8 B bf16 accesses), coalesced along C.
#include "common.cuh"
#include "kernels.h"

namespace dirb200 {

namespace {

// ---------------------------------------------------------------- maxpool 3x3 s2 p1 (resnet.py:247)
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * Ho * Wo * C4;
  if (idx >= total) return;
  int c = (int)(idx % C4) * 4;
  int64_t t = idx / C4;
  int wo = (int)(t % Wo);
  t /= Wo;
  int ho = (int)(t % Ho);
  int b = (int)(t / Ho);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int hi = ho * 2 - 1 + dy;
    if (hi < 0 || hi >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int wi = wo * 2 - 1 + dx;
      if (wi < 0 || wi >= W) continue;
      float4 v = ActIO<T>::ld4(x + (((int64_t)b * H + hi) * W + wi) * C + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  ActIO<T>::st4(y + (((int64_t)b * Ho + ho) * Wo + wo) * C + c, m);
}

// ---------------------------------------------------------------- upsample(2x bilinear, align_corners=False) + concat + BN/ReLU
// models/dir.py:442-444,455,459-461,470 and hourglass.py:60-61 (bn1+relu1 of the consuming Residual)
template <typename T>
__global__ void concat_preact_kernel(const T* __restrict__ s0, int C0, int up0, const T* __restrict__ s1, int C1,
                                     const float* __restrict__ bns, const float* __restrict__ bnb, T* __restrict__ raw,
                                     T* __restrict__ act, int B, int Ho, int Wo) {
  const int C = C0 + C1;
  const int C4 = C >> 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)B * Ho * Wo * C4;
  if (idx >= total) return;
  int c = (int)(idx % C4) * 4;
  int64_t pix = idx / C4;
  int wo = (int)(pix % Wo);
  int64_t t = pix / Wo;
  int ho = (int)(t % Ho);
  int b = (int)(t / Ho);
  float4 v;
  if (c < C0) {
    if (up0) {
      const int Hi = Ho >> 1, Wi = Wo >> 1;
      float sy = fmaxf((ho + 0.5f) * 0.5f - 0.5f, 0.f);
      float sx = fmaxf((wo + 0.5f) * 0.5f - 0.5f, 0.f);
      int y0 = (int)sy, x0 = (int)sx;
      int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
      float wy = sy - y0, wx = sx - x0;
      const T* base = s0 + (int64_t)b * Hi * Wi * C0 + c;
      float4 v00 = ActIO<T>::ld4(base + ((int64_t)y0 * Wi + x0) * C0);
      float4 v01 = ActIO<T>::ld4(base + ((int64_t)y0 * Wi + x1) * C0);
      float4 v10 = ActIO<T>::ld4(base + ((int64_t)y1 * Wi + x0) * C0);
      float4 v11 = ActIO<T>::ld4(base + ((int64_t)y1 * Wi + x1) * C0);
      // same association as ATen's upsample_bilinear2d: rows first, then columns
      float w0y = 1.f - wy, w0x = 1.f - wx;
      v.x = w0y * (w0x * v00.x + wx * v01.x) + wy * (w0x * v10.x + wx * v11.x);
      v.y = w0y * (w0x * v00.y + wx * v01.y) + wy * (w0x * v10.y + wx * v11.y);
      v.z = w0y * (w0x * v00.z + wx * v01.z) + wy * (w0x * v10.z + wx * v11.z);
      v.w = w0y * (w0x * v00.w + wx * v01.w) + wy * (w0x * v10.w + wx * v11.w);
    } else {
      v = ActIO<T>::ld4(s0 + pix * C0 + c);
    }
  } else {
    v = ActIO<T>::ld4(s1 + pix * C1 + (c - C0));
  }
  if (raw) ActIO<T>::st4(raw + pix * C + c, v);
  if (act) {
    float4 s = __ldg(reinterpret_cast<const float4*>(bns + c));
    float4 h = __ldg(reinterpret_cast<const float4*>(bnb + c));
    if (raw && sizeof(T) == 2) {  // pre-activation must see the value the consumer of `raw` sees
      v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
      v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
    }
    v.x = fmaxf(fmaf(v.x, s.x, h.x), 0.f);
    v.y = fmaxf(fmaf(v.y, s.y, h.y), 0.f);
    v.z = fmaxf(fmaf(v.z, s.z, h.z), 0.f);
    v.w = fmaxf(fmaf(v.w, s.w, h.w), 0.f);
    ActIO<T>::st4(act + pix * C + c, v);
  }
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

__global__ void pack_conv_weight_kernel(const float* __restrict__ src, float* __restrict__ d32,
                                        __nv_bfloat16* __restrict__ d16, int Cout, int Cin, int kh, int kw, int Kpad) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Cout * Kpad) return;
  int k = (int)(idx % Kpad);
  int n = (int)(idx / Kpad);
  float v = 0.f;
  if (k < kh * kw * Cin) {
    int tap = k / Cin, ci = k - tap * Cin;
    int ky = tap / kw, kx = tap - ky * kw;
    v = src[(((int64_t)n * Cin + ci) * kh + ky) * kw + kx];
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

template <typename T>
__global__ void attn_pool_kernel(const T* __restrict__ f, const float* __restrict__ attn, float* __restrict__ pooled,
                                 int P, int C) {
  extern __shared__ float sa[];  // [P][2]
  int b = blockIdx.x;
  for (int i = threadIdx.x; i < P * 2; i += blockDim.x) sa[i] = attn[(int64_t)b * P * 2 + i];
  __syncthreads();
  float sl = 0.f, sr = 0.f;
  for (int p = 0; p < P; ++p) {
    sl += sa[p * 2];
    sr += sa[p * 2 + 1];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float al = 0.f, ar = 0.f, am = 0.f;
    for (int p = 0; p < P; ++p) {
      float v = ActIO<T>::ld(f + ((int64_t)b * P + p) * C + c);
      al = fmaf(v, sa[p * 2], al);
      ar = fmaf(v, sa[p * 2 + 1], ar);
      am += v;
    }
    pooled[((int64_t)b * 3 + 0) * C + c] = al / (sl + 1e-8f);
    pooled[((int64_t)b * 3 + 1) * C + c] = ar / (sr + 1e-8f);
    pooled[((int64_t)b * 3 + 2) * C + c] = am / (float)P;
  }
}

}  // namespace

template <typename T>
void launch_maxpool3x3s2(const T* x, T* y, int B, int H, int W, int C, cudaStream_t st) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  maxpool_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(x, y, B, H, W, C, Ho, Wo);
}

template <typename T>
void launch_concat_preact(const T* s0, int C0, int up0, const T* s1, int C1, const float* bns, const float* bnb, T* raw,
                          T* act, int B, int Ho, int Wo, cudaStream_t st) {
  int64_t total = (int64_t)B * Ho * Wo * ((C0 + C1) / 4);
  concat_preact_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(s0, C0, up0, s1, C1, bns, bnb, raw, act, B,
                                                                          Ho, Wo);
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
  head3_kernel<T><<<ceil_div(warps * 32, 256), 256, 0, st>>>(x, Cx, coff, C, w, bias, out, B, HW);
}

void launch_fold_affine(const float* cb, const float* g, const float* be, const float* mu, const float* var,
                        float* scale, float* shift, int n, cudaStream_t st) {
  fold_affine_kernel<<<ceil_div(n, 256), 256, 0, st>>>(cb, g, be, mu, var, scale, shift, n);
}

void launch_pack_conv_weight(const float* src, float* d32, __nv_bfloat16* d16, int Cout, int Cin, int kh, int kw,
                             int Kpad, cudaStream_t st) {
  int64_t total = (int64_t)Cout * Kpad;
  pack_conv_weight_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(src, d32, d16, Cout, Cin, kh, kw, Kpad);
}

void launch_transpose2d(const float* src, float* dst, int rows, int cols, cudaStream_t st) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose2d_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, rows, cols);
}

void launch_gcn_adjacency(const float* e1, float* A, cudaStream_t st) { gcn_adjacency_kernel<<<1, 32, 0, st>>>(e1, A); }

template <typename T>
void launch_attn_logits(const T* a, const float* w, const float* bias, float* attn, int B, int P, int C,
                        cudaStream_t st) {
  int warps = B * P * 2;
  attn_logits_kernel<T><<<ceil_div(warps * 32, 256), 256, 0, st>>>(a, w, bias, attn, B * P, C);
}

template <typename T>
void launch_attn_pool(const T* f, const float* attn, float* pooled, int B, int P, int C, cudaStream_t st) {
  attn_pool_kernel<T><<<B, 256, P * 2 * sizeof(float), st>>>(f, attn, pooled, P, C);
}

#define INST(T)                                                                                                      \
  template void launch_maxpool3x3s2<T>(const T*, T*, int, int, int, int, cudaStream_t);                              \
  template void launch_concat_preact<T>(const T*, int, int, const T*, int, const float*, const float*, T*, T*, int,  \
                                        int, int, cudaStream_t);                                                     \
  template void launch_nchw_to_nhwc<T>(const float*, T*, int, int, int, int, cudaStream_t);                          \
  template void launch_nhwc_to_nchw<T>(const T*, float*, int, int, int, int, cudaStream_t);                          \
  template void launch_head3<T>(const T*, int, int, int, const float*, const float*, float*, int, int, cudaStream_t); \
  template void launch_attn_logits<T>(const T*, const float*, const float*, float*, int, int, int, cudaStream_t);    \
  template void launch_attn_pool<T>(const T*, const float*, float*, int, int, int, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
#undef INST

}  // namespace dirb200
