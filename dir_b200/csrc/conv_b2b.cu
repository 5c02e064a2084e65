// Back-to-back 1x1 convolutions across a ResNet bottleneck boundary (models/backbone/resnet.py:120-140 of the
// torchvision-style Bottleneck the reference uses): the tail of block b
//     out = relu(bn3(conv3(t2)) + identity)                    (or the K-concatenated conv3 + downsample pair of block 0)
// and the head of block b+1
//     t1' = relu(bn1'(conv1'(out)))
// in ONE kernel. The 128-pixel x 256-channel `out` tile is staged in shared memory for its TMA store anyway, in exactly
// the K-major SWIZZLE_128B layout a TMA load would produce; the second GEMM reads it from there as its A operand, so the
// 268 MB (B=128, layer1) re-read of `out` by conv1' never happens. bf16 operands, fp32 accumulation in TMEM, same MMA
// order over K as the stand-alone kernels (conv_tc.cu), so `out` and t1' carry the same values.
//
// Per CTA (persistent over 128-pixel tiles, 320 threads):
//   warp 0    : TMA producer. Both weight matrices once (W3 [256][K1], W1' [N2][256]: 64-96 KB resident), then one A1
//               tile (t2, plus x for the pair) per tile; a single tile-sized stage is enough because GEMM1 is the first
//               ~10 % of a tile's timeline and the stage is free for the next tile's load as soon as it has been read.
//   warp 1    : MMA issuer. GEMM1: K1/16 MMAs of M128 x N256 into TMEM columns 0-255; GEMM2: one k-block of 4 MMAs
//               (M128 x N2) per staged 64-channel chunk of `out`, as soon as that chunk is ready, into columns 256...
//   warps 2-9 : two epilogue groups of 4 warps taking alternate 64-column chunks of GEMM1: tcgen05.ld -> scale/shift
//               (+ identity) -> ReLU -> bf16 -> swizzled staging -> TMA store of `out` AND hand-over of the chunk to the
//               MMA warp; then the (small) epilogue of GEMM2.
// The identity tile arrives by TMA through a ring of four (N2 = 64) or two (N2 = 128) chunks, a tile / half a tile ahead. (Reading it with per-thread 16-byte global
// loads, one row per thread, cost 5 400 cycles per tile: every such warp instruction touches 32 different lines.)
#include <cuda.h>

#include <tuple>

#include "common.cuh"
#include "engine.h"
#include "tc_common.cuh"
#include "tma_host.h"

namespace dirb200 {

namespace {

using namespace tc;

constexpr int B2B_THREADS = 320;
constexpr int CHUNK = 128 * 128;  // one 64-channel x 128-pixel bf16 box
constexpr int N1 = 256;           // channels of `out` (= K of the second GEMM)
constexpr int MAX_RES_SLOTS = 4;  // identity ring (16 KB chunks). slot = chunk % slots with 2 or 4 slots, so that every
                                  // slot is consumed by ONE epilogue group (chunk parity): a waiter that does not see
                                  // every phase of an mbarrier can pass a parity wait one phase early

struct B2bArgs {
  const float *scale1, *shift1, *scale2, *shift2;
  int res_slots;             // 0: no identity; 2 or 4: identity [M][256] through tmR and a ring of that many chunks
  int nk1, nk1a;             // k-blocks of GEMM1, of which the first nk1a come from tmA1a and the rest from tmA1b
  int w1_rows;               // rows per box of tmW1 (the layer's weight map may hold 128- or 256-row boxes)
  int tiles, relu1, relu2;
};

struct B2bBars {
  uint64_t w, a1_full, a1_empty, acc1_full, acc1_empty, a2_ready[4], g2_done, acc2_empty, res_full[MAX_RES_SLOTS], res_empty[MAX_RES_SLOTS];
  uint32_t tmem_ptr;
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// acc (64 fp32 columns of one TMEM lane) -> scale/shift (+ identity) (+ ReLU) -> 8 x 16 bytes of bf16, written to the
// swizzled staging row of this thread
template <bool HAS_RES>
__device__ __forceinline__ void finish_chunk(const float* v, uint32_t s_scale, uint32_t s_shift, uint32_t rrow, int relu,
                                             uint32_t orow, uint32_t swz) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 s0 = lds_f4(s_scale + q * 32), s1 = lds_f4(s_scale + q * 32 + 16);
    const float4 h0 = lds_f4(s_shift + q * 32), h1 = lds_f4(s_shift + q * 32 + 16);
    float o[8];
    o[0] = fmaf(v[q * 8 + 0], s0.x, h0.x);
    o[1] = fmaf(v[q * 8 + 1], s0.y, h0.y);
    o[2] = fmaf(v[q * 8 + 2], s0.z, h0.z);
    o[3] = fmaf(v[q * 8 + 3], s0.w, h0.w);
    o[4] = fmaf(v[q * 8 + 4], s1.x, h1.x);
    o[5] = fmaf(v[q * 8 + 5], s1.y, h1.y);
    o[6] = fmaf(v[q * 8 + 6], s1.z, h1.z);
    o[7] = fmaf(v[q * 8 + 7], s1.w, h1.w);
    if (HAS_RES) {
      const uint4 u = lds128(rrow + (((uint32_t)q ^ swz) << 4));
      const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(hr[e]);
        o[2 * e] += f.x;
        o[2 * e + 1] += f.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    uint4 pk;
    __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
    sts128(orow + (((uint32_t)q ^ swz) << 4), pk);
  }
}

template <int N2>
__global__ void __launch_bounds__(B2B_THREADS, 1)
conv1x1_b2b_kernel(const __grid_constant__ CUtensorMap tmA1a, const __grid_constant__ CUtensorMap tmA1b,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                   const __grid_constant__ CUtensorMap tmR, const B2bArgs a) {
  constexpr int NCH2 = N2 / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk1 = a.nk1;
  uint8_t* sW1 = smem;                        // nk1 x [256 rows][128 B]
  uint8_t* sW2 = sW1 + nk1 * 2 * CHUNK;       // 4 x [N2 rows][128 B]
  uint8_t* sA1 = sW2 + 4 * N2 * 128;          // nk1 x 16 KB
  uint8_t* sOut = sA1 + nk1 * CHUNK;          // 4 x 16 KB: the `out` tile = A operand of GEMM2
  uint8_t* sRes = sOut + 4 * CHUNK;           // res_slots x 16 KB identity ring
  float* s_aff = reinterpret_cast<float*>(sRes + a.res_slots * CHUNK);  // scale1[256] shift1[256] scale2[N2] shift2[N2]
  B2bBars* bars = reinterpret_cast<B2bBars*>(s_aff + 2 * N1 + 2 * N2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA1a);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmY1);
    tma_prefetch_desc(&tmY2);
    mbar_init(&bars->w, 1);
    mbar_init(&bars->a1_full, 1);
    mbar_init(&bars->a1_empty, 1);
    mbar_init(&bars->acc1_full, 1);
    mbar_init(&bars->acc1_empty, 256);
    for (int c = 0; c < 4; ++c) mbar_init(&bars->a2_ready[c], 1);
    mbar_init(&bars->g2_done, 1);
    mbar_init(&bars->acc2_empty, 128 * NCH2);
    for (int c = 0; c < MAX_RES_SLOTS; ++c) {
      mbar_init(&bars->res_full[c], 1);
      mbar_init(&bars->res_empty[c], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < N1; j += B2B_THREADS) {  // finalize-time constants: may be read before the PDL wait
    s_aff[j] = a.scale1[j];
    s_aff[N1 + j] = a.shift1[j];
  }
  for (int j = threadIdx.x; j < N2; j += B2B_THREADS) {
    s_aff[2 * N1 + j] = a.scale2[j];
    s_aff[2 * N1 + N2 + j] = a.shift2[j];
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0 && (int)blockIdx.x < a.tiles) {  // ===================== TMA producer
      mbar_expect_tx(&bars->w, (uint32_t)(nk1 * 2 * CHUNK + 4 * N2 * 128));
      for (int kb = 0; kb < nk1; ++kb)
        for (int r0 = 0; r0 < N1; r0 += a.w1_rows)
          tma_load_2d(&tmW1, &bars->w, sW1 + kb * 2 * CHUNK + r0 * 128, kb * 64, r0);
      for (int c = 0; c < 4; ++c) tma_load_2d(&tmW2, &bars->w, sW2 + c * N2 * 128, c * 64, 0);
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bars->a1_empty, (i & 1) ^ 1);
        mbar_expect_tx(&bars->a1_full, (uint32_t)(nk1 * CHUNK));
        for (int kb = 0; kb < nk1; ++kb) {
          if (kb < a.nk1a) tma_load_2d(&tmA1a, &bars->a1_full, sA1 + kb * CHUNK, kb * 64, tile * 128);
          else tma_load_2d(&tmA1b, &bars->a1_full, sA1 + kb * CHUNK, (kb - a.nk1a) * 64, tile * 128);
        }
        if (a.res_slots) {  // identity chunks of this tile; a slot is handed back by the epilogue group that consumed it
          for (int c = 0; c < 4; ++c) {
            const uint32_t slot = c & (a.res_slots - 1), use = a.res_slots == 4 ? i : 2 * i + (c >> 1);
            mbar_wait(&bars->res_empty[slot], (use & 1) ^ 1);
            mbar_expect_tx(&bars->res_full[slot], CHUNK);
            tma_load_2d(&tmR, &bars->res_full[slot], sRes + slot * CHUNK, c * 64, tile * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (int)blockIdx.x < a.tiles) {  // ===================== MMA issuer
      constexpr uint32_t ID1 = idesc(N1, 1u), ID2 = idesc(N2, 1u);
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + N1;
      mbar_wait(&bars->w, 0);
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++i) {
        mbar_wait(&bars->acc1_empty, (i & 1) ^ 1);
        mbar_wait(&bars->a1_full, i & 1);
        fence_after();
        for (int kb = 0; kb < nk1; ++kb) {
          const uint64_t da = desc128(s32(sA1 + kb * CHUNK)), db = desc128(s32(sW1 + kb * 2 * CHUNK));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma(acc1, da + 2 * k, db + 2 * k, ID1, (kb | k) ? 1u : 0u);
        }
        umma_commit(&bars->a1_empty);
        umma_commit(&bars->acc1_full);
        mbar_wait(&bars->acc2_empty, (i & 1) ^ 1);
        for (int c = 0; c < 4; ++c) {
          mbar_wait(&bars->a2_ready[c], i & 1);  // chunk c of `out` staged (and fenced to the async proxy)
          fence_after();
          const uint64_t da = desc128(s32(sOut + c * CHUNK)), db = desc128(s32(sW2 + c * N2 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma(acc2, da + 2 * k, db + 2 * k, ID2, (c | k) ? 1u : 0u);
        }
        umma_commit(&bars->g2_done);  // accumulator 2 complete AND the staged `out` tile no longer needed by the MMAs
      }
    }
  } else {
    // ===================== epilogues
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t s_scale1 = s32(s_aff), s_shift1 = s32(s_aff + N1);
    const uint32_t s_scale2 = s32(s_aff + 2 * N1), s_shift2 = s32(s_aff + 2 * N1 + N2);
    const bool has_res = a.res_slots != 0;
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++i) {
      mbar_wait(&bars->acc1_full, i & 1);
      fence_after();
      // the previous tile's stores have read their staging buffers (they had a whole tile's time to)
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = eg + 2 * cc;
        float v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + c * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (cc == 1) {  // this thread has read all it needs from accumulator 1
          fence_before();
          mbar_arrive(&bars->acc1_empty);
        }
        group_barrier(eg);  // wait_group.read above (first chunk) ordered before any write of the group
        const uint32_t orow = s32(sOut + c * CHUNK + row * 128);
        const uint32_t slot = c & (a.res_slots - 1), use = a.res_slots == 4 ? i : 2 * i + cc;
        if (has_res) {
          mbar_wait(&bars->res_full[slot], use & 1);
          finish_chunk<true>(v, s_scale1 + c * 256, s_shift1 + c * 256, s32(sRes + slot * CHUNK + row * 128), a.relu1, orow,
                             swz);
        } else {
          finish_chunk<false>(v, s_scale1 + c * 256, s_shift1 + c * 256, 0u, a.relu1, orow, swz);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> TMA store and MMA (async proxy)
        group_barrier(eg);  // chunk staged; every thread of the group is also past its identity reads
        if (et == 0) {
          tma_store_2d(&tmY1, sOut + c * CHUNK, c * 64, tile * 128);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          mbar_arrive(&bars->a2_ready[c]);
          if (has_res) mbar_arrive(&bars->res_empty[slot]);
        }
      }
      // ---- GEMM2 epilogue; every epilogue thread passes this wait: it also says the staged `out` tile may be overwritten
      mbar_wait(&bars->g2_done, i & 1);
      fence_after();
      if (eg < NCH2) {
        float v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + N1 + eg * 64;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        fence_before();
        mbar_arrive(&bars->acc2_empty);
        // t1' is staged in chunk `eg` of the `out` staging: the MMAs are done with it (g2_done) and this group's store of
        // that chunk, the older of its two, has been read (the newer one may still be in flight)
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        group_barrier(eg);
        const uint32_t orow = s32(sOut + eg * CHUNK + row * 128);
        finish_chunk<false>(v, s_scale2 + eg * 256, s_shift2 + eg * 256, 0u, a.relu2, orow, swz);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_barrier(eg);
        if (et == 0) {
          tma_store_2d(&tmY2, sOut + eg * CHUNK, eg * 64, tile * 128);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int N2>
constexpr int b2b_smem(int nk1, int res_slots) {
  return 1024 + nk1 * 2 * CHUNK + 4 * N2 * 128 + nk1 * CHUNK + 4 * CHUNK + res_slots * CHUNK + (2 * N1 + 2 * N2) * 4 + 256;
}
constexpr int b2b_res_slots(int n2, bool has_res) { return !has_res ? 0 : (n2 == 64 ? 4 : 2); }

// [rows][cols] bf16 row-major viewed as 64-column x 128-row boxes with 128B swizzle
bool rowmajor_map(const void* ptr, int rows, int cols, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int> Key;
  static thread_local tma::MapCache<Key> cache;
  Key k(ptr, rows, cols);
  if (cache.find(k, out)) return true;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  if (tma::get_encode()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  cache.put(k, *out);
  return true;
}

template <int N2>
int launch_b2b(const CUtensorMap& tmA1a, const CUtensorMap& tmA1b, const ConvLayer& first, const ConvLayer& next,
               const CUtensorMap& tmY1, const CUtensorMap& tmY2, const CUtensorMap& tmR, const B2bArgs& a, cudaStream_t st) {
  const int smem = b2b_smem<N2>(a.nk1, a.res_slots);
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(conv1x1_b2b_kernel<N2>), 227 * 1024) != cudaSuccess)
    return DIRB200_E_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.tiles < tma::num_sms() ? a.tiles : tma::num_sms());
  cfg.blockDim = dim3(B2B_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, conv1x1_b2b_kernel<N2>, tmA1a, tmA1b, first.wmap, next.wmap, tmY1, tmY2, tmR, a) != cudaSuccess)
    return DIRB200_E_CUDA;
  return DIRB200_OK;
}

}  // namespace

// `first`: a 1x1 conv (or K-concatenated 1x1 pair) with 256 output channels over K1 = Ka (+ Kb) in {64, 128} input
// channels; `next`: the 1x1 stride-1 conv (256 -> 64 or 128) that consumes its output; M output pixels.
bool conv_b2b_supported(const ConvLayer& first, int Ka, int Kb, const ConvLayer& next, int M, bool has_res) {
  if (!tma::get_encode() || !first.w16 || !next.w16) return false;
  if (first.Cout != N1 || (first.wmap_bn != 128 && first.wmap_bn != 256) || first.K != Ka + Kb || first.K != first.Kpad)
    return false;
  if (Ka % 64 || Kb % 64 || Ka <= 0 || (Ka + Kb != 64 && Ka + Kb != 128)) return false;
  if (next.kh != 1 || next.kw != 1 || next.stride != 1 || next.pad != 0 || next.Cin != N1 || next.K != next.Kpad) return false;
  if ((next.Cout != 64 && next.Cout != 128) || next.wmap_bn != next.Cout) return false;
  const int rs = b2b_res_slots(next.Cout, has_res);
  const int smem = next.Cout == 64 ? b2b_smem<64>((Ka + Kb) / 64, rs) : b2b_smem<128>((Ka + Kb) / 64, rs);
  return smem <= 227 * 1024 && M > 0 && M % 128 == 0;
}

// out[M][256] = act1(xa[M][Ka] | xb[M][Kb] . W1^T * scale1 + shift1 (+ res));  t1[M][N2] = act2(out . W2^T * scale2 + shift2)
int launch_conv_b2b(const ConvLayer& first, const __nv_bfloat16* xa, int Ka, const __nv_bfloat16* xb, int Kb,
                    const __nv_bfloat16* res, const ConvLayer& next, __nv_bfloat16* out, __nv_bfloat16* t1, int M,
                    cudaStream_t st) {
  CUtensorMap tmA1a, tmA1b, tmY1, tmY2, tmR;
  if (!rowmajor_map(xa, M, Ka, &tmA1a) || !rowmajor_map(out, M, N1, &tmY1) || !rowmajor_map(t1, M, next.Cout, &tmY2))
    return DIRB200_E_CUDA;
  if (!xb || Kb == 0) tmA1b = tmA1a;
  else if (!rowmajor_map(xb, M, Kb, &tmA1b)) return DIRB200_E_CUDA;
  if (!res) tmR = tmY1;
  else if (!rowmajor_map(res, M, N1, &tmR)) return DIRB200_E_CUDA;
  B2bArgs a{};
  a.scale1 = first.scale;
  a.shift1 = first.shift;
  a.scale2 = next.scale;
  a.shift2 = next.shift;
  a.res_slots = b2b_res_slots(next.Cout, res != nullptr);
  a.nk1 = (Ka + Kb) / 64;
  a.nk1a = Ka / 64;
  a.w1_rows = first.wmap_bn;
  a.tiles = M / 128;
  a.relu1 = first.relu;
  a.relu2 = next.relu;
  if (next.Cout == 64) return launch_b2b<64>(tmA1a, tmA1b, first, next, tmY1, tmY2, tmR, a, st);
  return launch_b2b<128>(tmA1a, tmA1b, first, next, tmY1, tmY2, tmR, a, st);
}

}  // namespace dirb200
