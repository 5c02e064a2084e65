// ResNet stem in one kernel (bf16 configuration): conv1 7x7/s2 (3->64) + bn1 + ReLU + maxpool 3x3/s2
// (models/backbone/resnet.py:176-179,243-247), from the zero-padded NHWC4 bf16 image written by stem_pack_kernel.
//
// Why a dedicated kernel: as an implicit GEMM fed by TMA the stem was bound by the TMA request rate (one 64-byte
// request per (output pixel, kernel row): 2.9 cycles each, 222 us at B=128) and wrote the 268 MB conv map only for
// the max-pool to read it again (80 us). Here
//  * the A operand needs NO im2col and no per-window loads: a kernel row of an output pixel is the 64-byte window
//    [2*wo, 2*wo+8) px of one padded input row, and consecutive windows are 16 bytes apart — which is exactly the
//    canonical NO-swizzle K-major UMMA layout with LBO = 16 B, SBO = 128 B (row i, 16-byte K-chunk c at 16*(i+c)):
//    the tensor core reads overlapping windows straight out of the raw input row in shared memory. One bulk copy
//    (2112 B) per input row per unit replaces 128 TMA requests per output row and kernel row;
//  * a unit = (image, 8 pooled rows) = 17 conv rows (one halo row recomputed); conv rows go TMEM -> bn/ReLU -> bf16
//    registers; a thread keeps its pixel's last two rows, so the vertical 3-max never touches memory; every second
//    row that maximum goes through a double-buffered smem row for the horizontal stride-2 3-max and the pooled row
//    is written to HBM: the conv map never leaves the SM (HBM traffic 268+268+67 MB -> 67 MB).
// Roles (320 threads): warps 0-7 epilogue + pooling (two threads per output pixel), warp 8 input-row producer,
// warp 9 TMEM alloc + MMA issue. 14 tcgen05.mma (M=128, N=64, K=16) per conv row.
#include "../../include/dirb200.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dirb200 {

namespace {

using namespace tc;

constexpr int SP_THREADS = 320;
constexpr int ROW_BYTES = 264 * 8;   // one padded NHWC4 bf16 input row (W + 8 = 264 px)
constexpr int IN_RING = 16;          // input-row slots (7 live per conv row + prefetch)
constexpr int W_BYTES = 7 * 4096;    // weights: 7 kernel rows x [64 n][32 k] bf16, no-swizzle core-matrix layout
constexpr int CONV_ROW = 128 * 128;  // one conv output row in smem: 128 px x 64 ch bf16
constexpr int CRING = 2;             // double-buffered row of vertical maxima
constexpr int OFF_IN = 0;
constexpr int OFF_W = OFF_IN + IN_RING * ROW_BYTES;  // 33792
constexpr int OFF_C = OFF_W + W_BYTES;               // 62464
constexpr int OFF_AFF = OFF_C + CRING * CONV_ROW;    // bn1 scale | shift (2 x 64 floats)
constexpr int OFF_BARS = OFF_AFF + 512;
constexpr int SP_SMEM = 1024 + OFF_BARS + 512;       // ~104 KB: two CTAs per SM hide each other's epilogue latency
constexpr int ACC_BUFS = 4;

struct SpBars {
  uint64_t full[IN_RING], empty[IN_RING], wbar, tfull[ACC_BUFS], tempty[ACC_BUFS];
  uint32_t tmem_ptr;
};

// no-swizzle K-major descriptor: core matrices of 8 rows x 16 B; lbo = byte step between the two K-chunks of one
// MMA, sbo = byte step between 8-row groups (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE)
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

struct UnitGeom {
  int b, q;       // image, group of 8 pooled rows
  int ho0;        // first conv row of the unit (16q - 1, may be -1)
  int y_lo, y_hi; // input rows loaded for the unit (inclusive)
};
__device__ __forceinline__ UnitGeom unit_geom(int u) {
  UnitGeom g;
  g.b = u >> 3;
  g.q = u & 7;
  g.ho0 = 16 * g.q - 1;
  const int first_real = g.ho0 < 0 ? 0 : g.ho0;
  g.y_lo = max(0, 2 * first_real - 3);
  g.y_hi = min(255, 2 * (g.ho0 + 16) + 3);
  return g;
}

__global__ void __launch_bounds__(SP_THREADS, 2)
stem_pool_kernel(const uint8_t* __restrict__ in /*[B][256][264][4] bf16*/, const uint8_t* __restrict__ wpk,
                 const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SpBars* bars = reinterpret_cast<SpBars*>(smem + OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = B * 8;

  if (threadIdx.x == 0) {
    for (int i = 0; i < IN_RING; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    mbar_init(&bars->wbar, 1);
    for (int i = 0; i < ACC_BUFS; ++i) {
      mbar_init(&bars->tfull[i], 1);
      mbar_init(&bars->tempty[i], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_ptr;

  if (warp == 8) {
    // ===================================================== producer: weights once, then the input rows of each unit
    if (lane == 0) {
      mbar_expect_tx(&bars->wbar, W_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       s32(smem + OFF_W)),
                   "l"(wpk), "r"(W_BYTES), "r"(s32(&bars->wbar))
                   : "memory");
      pdl_wait();  // the image rows are written by the preceding stem_pack kernel
      uint32_t idx = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const UnitGeom g = unit_geom(u);
        for (int yy = g.y_lo; yy <= g.y_hi; ++yy, ++idx) {
          const int slot = idx % IN_RING;
          if (idx >= IN_RING) mbar_wait(&bars->empty[slot], ((idx / IN_RING) - 1) & 1);
          mbar_expect_tx(&bars->full[slot], ROW_BYTES);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s32(smem + OFF_IN + slot * ROW_BYTES)),
                       "l"(in + ((size_t)g.b * 256 + yy) * ROW_BYTES), "r"(ROW_BYTES), "r"(s32(&bars->full[slot]))
                       : "memory");
        }
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t sb = s32(smem);
      mbar_wait(&bars->wbar, 0);
      uint32_t base_idx = 0, nrow = 0;  // running input-row index of the unit's y_lo; conv rows issued so far
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const UnitGeom g = unit_geom(u);
        for (int j = 0; j < 17; ++j) {
          const int ho = g.ho0 + j;
          if (ho < 0) continue;  // the halo row above the image: the epilogue writes zeros
          const uint32_t buf = nrow % ACC_BUFS;
          mbar_wait(&bars->tempty[buf], ((nrow / ACC_BUFS) & 1) ^ 1);
          fence_after();
          uint32_t acc = 0;
          for (int ky = 0; ky < 7; ++ky) {
            const int yy = 2 * ho - 3 + ky;
            if (yy < 0 || yy > 255) continue;  // zero padding in y: no contribution
            const uint32_t idx = base_idx + (uint32_t)(yy - g.y_lo);
            const uint32_t slot = idx % IN_RING;
            mbar_wait(&bars->full[slot], (idx / IN_RING) & 1);
            fence_after();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t da = desc_nosw(sb + OFF_IN + slot * ROW_BYTES + ks * 32, 16, 128);
              const uint64_t db = desc_nosw(sb + OFF_W + ky * 4096 + ks * 256, 128, 512);
              umma(tmem + buf * 64, da, db, idesc(64), acc);
              acc = 1;
            }
          }
          umma_commit(&bars->tfull[buf]);
          ++nrow;
          // input rows no later conv row needs: 2ho-3, 2ho-2 (and everything left after the unit's last row)
          const int rel_hi = j == 16 ? g.y_hi : 2 * ho - 2;
          for (int yy = max(g.y_lo, 2 * ho - 3); yy <= rel_hi; ++yy) {
            const uint32_t idx = base_idx + (uint32_t)(yy - g.y_lo);
            umma_commit(&bars->empty[idx % IN_RING]);
          }
        }
        base_idx += (uint32_t)(g.y_hi - g.y_lo + 1);
      }
    }
  } else {
    // ===================================================== epilogue + pooling: thread = (pixel, channel half)
    const int px = threadIdx.x & 127, half = threadIdx.x >> 7;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* aff = reinterpret_cast<float*>(smem + OFF_AFF);
    if (threadIdx.x < 64) {
      aff[threadIdx.x] = __ldg(scale + threadIdx.x);
      aff[64 + threadIdx.x] = __ldg(shift + threadIdx.x);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float* sc = aff + 32 * half;
    const float* sh = aff + 64 + 32 * half;
    uint8_t* vbuf = smem + OFF_C;  // 2 x [128 px][128 B]: vertical 3-max of the conv rows of one pooled row
    const int ppx = threadIdx.x >> 2, quarter = threadIdx.x & 3;  // pooling role: pooled pixel, 16-channel quarter
    uint32_t nrow = 0, npool = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const UnitGeom g = unit_geom(u);
      uint4 prev1[4], prev2[4];  // this pixel's 32 channels of conv rows j-1 and j-2 (post-ReLU: 0 == -inf padding)
#pragma unroll
      for (int c = 0; c < 4; ++c) prev1[c] = prev2[c] = make_uint4(0u, 0u, 0u, 0u);
      for (int j = 0; j < 17; ++j) {
        const int ho = g.ho0 + j;
        uint4 cur[4];
        if (ho < 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) cur[c] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          const uint32_t buf = nrow % ACC_BUFS;
          mbar_wait(&bars->tfull[buf], (nrow / ACC_BUFS) & 1);
          fence_after();
          float v[32];
          tmem_ld32(trow + buf * 64 + 32 * half, v);
          tmem_ld_wait();
          fence_before();
          mbar_arrive(&bars->tempty[buf]);
          ++nrow;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 s0 = *reinterpret_cast<const float4*>(sc + 8 * c), s1 = *reinterpret_cast<const float4*>(sc + 8 * c + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(sh + 8 * c), h1 = *reinterpret_cast<const float4*>(sh + 8 * c + 4);
            const float* vv = v + 8 * c;
            __nv_bfloat162 q0 = __floats2bfloat162_rn(fmaxf(fmaf(vv[0], s0.x, h0.x), 0.f), fmaxf(fmaf(vv[1], s0.y, h0.y), 0.f));
            __nv_bfloat162 q1 = __floats2bfloat162_rn(fmaxf(fmaf(vv[2], s0.z, h0.z), 0.f), fmaxf(fmaf(vv[3], s0.w, h0.w), 0.f));
            __nv_bfloat162 q2 = __floats2bfloat162_rn(fmaxf(fmaf(vv[4], s1.x, h1.x), 0.f), fmaxf(fmaf(vv[5], s1.y, h1.y), 0.f));
            __nv_bfloat162 q3 = __floats2bfloat162_rn(fmaxf(fmaf(vv[6], s1.z, h1.z), 0.f), fmaxf(fmaf(vv[7], s1.w, h1.w), 0.f));
            cur[c] = make_uint4(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1),
                                *reinterpret_cast<uint32_t*>(&q2), *reinterpret_cast<uint32_t*>(&q3));
          }
        }
        if (j >= 2 && (j & 1) == 0) {  // conv rows j-2, j-1, j = 2po-1, 2po, 2po+1: vertical max stays in registers
          const int po = 8 * g.q + (j >> 1) - 1;
          uint8_t* vb = vbuf + (npool & 1) * CONV_ROW;
          uint8_t* vrow = vb + px * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c)  // 16-byte chunk (4*half + c) of this pixel, XOR-swizzled against bank conflicts
            *reinterpret_cast<uint4*>(vrow + (((4 * half + c) ^ (px & 7)) << 4)) =
                max_bf16x8(max_bf16x8(prev2[c], prev1[c]), cur[c]);
          asm volatile("bar.sync 1, 256;" ::: "memory");  // (the buffer written two pooled rows ago is free again)
          uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int p = 2 * ppx + dx;
            if (p < 0) continue;
            const uint8_t* pp = vb + p * 128;
            m0 = max_bf16x8(m0, *reinterpret_cast<const uint4*>(pp + (((2 * quarter) ^ (p & 7)) << 4)));
            m1 = max_bf16x8(m1, *reinterpret_cast<const uint4*>(pp + (((2 * quarter + 1) ^ (p & 7)) << 4)));
          }
          uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)g.b * 64 + po) * 64 + ppx) * 64 + 16 * quarter);
          dst[0] = m0;
          dst[1] = m1;
          ++npool;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          prev2[c] = prev1[c];
          prev1[c] = cur[c];
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// conv1.weight [64][3][7][7] fp32 -> per kernel row ky a [64 n][32 k] bf16 tile (k = kx*4 + c, zero for kx = 7 or
// c = 3) in the no-swizzle K-major core-matrix layout: (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2 bytes
__global__ void pack_stem_pool_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 7 * 64 * 32) return;
  const int k = idx & 31, n = (idx >> 5) & 63, ky = idx >> 11;
  const int kx = k >> 2, c = k & 3;
  const float v = (kx < 7 && c < 3) ? w[((n * 3 + c) * 7 + ky) * 7 + kx] : 0.f;
  out[ky * 2048 + (n >> 3) * 256 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
}

}  // namespace

size_t stem_pool_weight_bytes() { return W_BYTES; }

void launch_pack_stem_pool_weight(const float* w, void* packed, cudaStream_t st) {
  pack_stem_pool_weight_kernel<<<(7 * 64 * 32 + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(packed));
}

// in: the padded NHWC4 bf16 image (B,256,264,4) of stem_pack_kernel; y: (B,64,64,64) NHWC bf16
int launch_stem_pool(const void* in, const void* packed_w, const float* scale, const float* shift, __nv_bfloat16* y, int B,
                     cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = B * 8;
  launch_pdl(stem_pool_kernel, dim3(units < 2 * sms ? units : 2 * sms), dim3(SP_THREADS), SP_SMEM, st,
             reinterpret_cast<const uint8_t*>(in), reinterpret_cast<const uint8_t*>(packed_w), scale, shift, y, B);
  return DIRB200_OK;
}

}  // namespace dirb200
