// Shared helpers for the dirb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dirb200 {

constexpr float kBnEps = 1e-5f;

template <typename T>
struct ActIO;  // feature-map element type: float (fp32 mode) or __nv_bfloat16 (bf16 mode)

template <>
struct ActIO<float> {
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};

template <>
struct ActIO<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Register-tiled CUDA-core GEMM for the small joint-space matrices:
//   C[r][n] = sum_k A[r][k] * Bt[k][n],  A in shared memory (row-major, lda floats, 16B-aligned rows),
//   Bt = K-major weights in global memory (row k at Bt + k*ldb, 16B-aligned, K % 4 == 0).
// Work item = TM rows x 4 columns (row group rg, column group cg). With consecutive threads on consecutive column
// groups the weight loads are coalesced LDG.128 and the A reads are warp-broadcast LDS.128:
// 16*TM FMAs per TM LDS + 4 LDG. Rows >= M are clamped on load (their accumulators are garbage, never stored).
template <int TM>
__device__ __forceinline__ void smem_gemm_item(const float* __restrict__ A, int lda, int M, int K,
                                               const float* __restrict__ Bt, int ldb, int cg, int rg,
                                               float (&acc)[TM][4]) {
#pragma unroll
  for (int r = 0; r < TM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  const float* arow[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) arow[r] = A + (size_t)min(rg * TM + r, M - 1) * lda;
  const float* bp = Bt + cg * 4;
  // weights come straight from L2 (~600 cycles): keep two k-steps (8 x LDG.128) in flight ahead of the FMAs
  float4 bq[2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      bq[u][q] = (u * 4 < K) ? __ldg(reinterpret_cast<const float4*>(bp + (size_t)(u * 4 + q) * ldb))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < K; k += 8) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kk = k + u * 4;
      if (kk >= K) break;
      const float4 b0 = bq[u][0], b1 = bq[u][1], b2 = bq[u][2], b3 = bq[u][3];
      if (kk + 8 < K) {
#pragma unroll
        for (int q = 0; q < 4; ++q) bq[u][q] = __ldg(reinterpret_cast<const float4*>(bp + (size_t)(kk + 8 + q) * ldb));
      }
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(arow[r] + kk);
        acc[r][0] = fmaf(a.x, b0.x, acc[r][0]); acc[r][1] = fmaf(a.x, b0.y, acc[r][1]);
        acc[r][2] = fmaf(a.x, b0.z, acc[r][2]); acc[r][3] = fmaf(a.x, b0.w, acc[r][3]);
        acc[r][0] = fmaf(a.y, b1.x, acc[r][0]); acc[r][1] = fmaf(a.y, b1.y, acc[r][1]);
        acc[r][2] = fmaf(a.y, b1.z, acc[r][2]); acc[r][3] = fmaf(a.y, b1.w, acc[r][3]);
        acc[r][0] = fmaf(a.z, b2.x, acc[r][0]); acc[r][1] = fmaf(a.z, b2.y, acc[r][1]);
        acc[r][2] = fmaf(a.z, b2.z, acc[r][2]); acc[r][3] = fmaf(a.z, b2.w, acc[r][3]);
        acc[r][0] = fmaf(a.w, b3.x, acc[r][0]); acc[r][1] = fmaf(a.w, b3.y, acc[r][1]);
        acc[r][2] = fmaf(a.w, b3.z, acc[r][2]); acc[r][3] = fmaf(a.w, b3.w, acc[r][3]);
      }
    }
  }
}

// All items of an (M x N) product, strided over the CTA; epi(row, col, value) for every valid element.
template <int TM, typename Epi>
__device__ __forceinline__ void smem_gemm(const float* __restrict__ A, int lda, int M, int K,
                                          const float* __restrict__ Bt, int ldb, int N, int nthreads, Epi epi) {
  const int ncg = N >> 2, nrg = (M + TM - 1) / TM;
  for (int item = threadIdx.x; item < ncg * nrg; item += nthreads) {
    const int cg = item % ncg, rg = item / ncg;
    float acc[TM][4];
    smem_gemm_item<TM>(A, lda, M, K, Bt, ldb, cg, rg, acc);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int row = rg * TM + r;
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 4; ++j) epi(row, cg * 4 + j, acc[r][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-cooperative GEMM with the weights STREAMED through shared memory by bulk async copies (cp.async.bulk +
// mbarrier), for the latency-bound joint-space products:
//   C[M x NC] = A[M x K] * Bt[K x NC],   A in smem, Bt rows in global (row k at Bt + k*ldb, NC in {64,128} columns).
// Slabs of 32 k-rows (NC*128 B) are double-buffered; warp 0 issues one 16B-aligned bulk copy per row two slabs ahead,
// every thread owns one TM x 4 register tile. All threads of the CTA must call this (it contains __syncthreads).
struct WStream {
  float* wbuf;     // 2 x 32 x 128 floats (32 KB), 128B-aligned
  uint64_t* bar;   // 2 mbarriers (initialised to count 1)
  uint32_t it;     // running slab counter (buffer = it & 1, parity = (it >> 1) & 1)
};

__device__ __forceinline__ uint32_t cta_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void wstream_init(WStream& ws, float* wbuf, uint64_t* bar) {
  ws.wbuf = wbuf;
  ws.bar = bar;
  ws.it = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cta_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cta_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ void wstream_issue(const WStream& ws, uint32_t g, const float* __restrict__ Bt, int ldb,
                                              int k0, int NC) {
  // called by warp 0 only: slab g <- rows k0 .. k0+31 (NC floats each)
  const int lane = threadIdx.x & 31;
  const uint32_t buf = g & 1;
  const uint32_t bar = cta_smem_u32(&ws.bar[buf]);
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(32 * NC * 4))
                 : "memory");
  __syncwarp();
  const uint32_t dst = cta_smem_u32(ws.wbuf + buf * (32 * 128) + lane * NC);
  const float* src = Bt + (size_t)(k0 + lane) * ldb;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"((uint32_t)(NC * 4)), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void wstream_wait(const WStream& ws, uint32_t g) {
  const uint32_t bar = cta_smem_u32(&ws.bar[g & 1]);
  const uint32_t parity = (g >> 1) & 1;
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

template <int TM, typename Epi>
__device__ __forceinline__ void cta_gemm(const float* __restrict__ A, int lda, int M, int K,
                                         const float* __restrict__ Bt, int ldb, int NC, WStream& ws, Epi epi) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ncg = NC >> 2, nrg = (M + TM - 1) / TM;
  const int cg = tid % ncg, rg = tid / ncg;
  const bool active = tid < ncg * nrg;
  const int nslab = K >> 5;
  const uint32_t g0 = ws.it;
  if (warp == 0) {
    wstream_issue(ws, g0, Bt, ldb, 0, NC);
    if (nslab > 1) wstream_issue(ws, g0 + 1, Bt, ldb, 32, NC);
  }
  float acc[TM][4];
#pragma unroll
  for (int r = 0; r < TM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  const float* arow[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) arow[r] = A + (size_t)min(rg * TM + r, M - 1) * lda;
  for (int sl = 0; sl < nslab; ++sl) {
    const uint32_t g = g0 + sl;
    wstream_wait(ws, g);
    if (active) {
      const float* wb = ws.wbuf + (g & 1) * (32 * 128) + cg * 4;
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(wb + (kk + 0) * NC);
        const float4 b1 = *reinterpret_cast<const float4*>(wb + (kk + 1) * NC);
        const float4 b2 = *reinterpret_cast<const float4*>(wb + (kk + 2) * NC);
        const float4 b3 = *reinterpret_cast<const float4*>(wb + (kk + 3) * NC);
#pragma unroll
        for (int r = 0; r < TM; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(arow[r] + sl * 32 + kk);
          acc[r][0] = fmaf(a.x, b0.x, acc[r][0]); acc[r][1] = fmaf(a.x, b0.y, acc[r][1]);
          acc[r][2] = fmaf(a.x, b0.z, acc[r][2]); acc[r][3] = fmaf(a.x, b0.w, acc[r][3]);
          acc[r][0] = fmaf(a.y, b1.x, acc[r][0]); acc[r][1] = fmaf(a.y, b1.y, acc[r][1]);
          acc[r][2] = fmaf(a.y, b1.z, acc[r][2]); acc[r][3] = fmaf(a.y, b1.w, acc[r][3]);
          acc[r][0] = fmaf(a.z, b2.x, acc[r][0]); acc[r][1] = fmaf(a.z, b2.y, acc[r][1]);
          acc[r][2] = fmaf(a.z, b2.z, acc[r][2]); acc[r][3] = fmaf(a.z, b2.w, acc[r][3]);
          acc[r][0] = fmaf(a.w, b3.x, acc[r][0]); acc[r][1] = fmaf(a.w, b3.y, acc[r][1]);
          acc[r][2] = fmaf(a.w, b3.z, acc[r][2]); acc[r][3] = fmaf(a.w, b3.w, acc[r][3]);
        }
      }
      if (sl == nslab - 1) {
#pragma unroll
        for (int r = 0; r < TM; ++r) {
          const int row = rg * TM + r;
          if (row < M) epi(r, row, cg * 4, acc[r]);
        }
      }
    }
    __syncthreads();  // slab buffer (g & 1) fully consumed (and, on the last slab, epi() results visible)
    if (warp == 0 && sl + 2 < nslab) wstream_issue(ws, g + 2, Bt, ldb, (sl + 2) * 32, NC);
  }
  ws.it = g0 + nslab;
}

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch. Every kernel of the forward is launched with the programmatic-serialization
// attribute and calls pdl_wait() before its first access to memory another kernel of the forward wrote or will
// read (everything before the wait touches only registers, smem, TMEM and finalize-time weights). Because EVERY
// kernel waits, completion of kernel N implies completion of all its predecessors, so workspace reuse stays
// race-free. No kernel triggers early: measured on B200 (profiles/pdl_r1.txt) the implicit trigger at CTA exit
// gives +3.9 % images/s, an explicit griddepcontrol.launch_dependents at kernel entry costs 3 % instead.
// DIRB200_PDL=0 launches without the attribute (the wait is then a no-op).
#ifdef DIRB200_NO_PDL_INSTR
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_wait() {
#ifndef DIRB200_PDL_NO_WAIT
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#ifndef DIRB200_PDL_NO_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

bool pdl_enabled();  // engine.cu
// Opt a kernel in to more than 48 KB of (static + dynamic) shared memory. The attribute is per (function, device), so
// it is remembered per pair (a process-wide "done" flag would leave a second device of the process without it).
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);  // engine.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (smem > 0) {
    const cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem);
    if (e != cudaSuccess) return e;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace dirb200
