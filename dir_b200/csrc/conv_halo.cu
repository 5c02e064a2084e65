// 3x3 / stride 1 / 64 -> 64 channel convolution with the input halo loaded ONCE per tile (bf16, tcgen05).
//
// conv_tc.cu fetches one 16 KB A tile per filter tap: nine L2->SM transfers of (almost) the same pixels. For the
// 64-channel 3x3 layers (ResNet layer1 conv2; most of HRNet) that makes the SM's TMA ingest, not the tensor pipe, the
// limit: ncu shows 615 MB L2->SM for 94 MB of DRAM traffic and 18 % tensor pipe (profiles/conv_ncu_r2_bf16.txt). Here a
// tile = 128 consecutive output pixels = R = 128/W whole image rows; its (R+2) x W input pixels arrive by ONE TMA box
// (zero-filled above / below the image) and all nine taps are start-address offsets into shared memory.
// To make a tap a pure address offset the operand must be linear in the pixel index, which the 128B-swizzled layout is
// not. Four warps therefore re-lay the halo CHUNK-major -- plane[8-channel chunk][pixel][16 B] -- which is the
// canonical no-swizzle K-major UMMA layout (core matrix = 8 pixels x 16 B contiguous: SBO = 128 B; the two K-chunks of
// an MMA are one chunk plane apart: LBO = plane stride). Operand row r of tap (ky,kx) is then pixel r + ky*W + kx - 1.
// Horizontal zero padding without a padded pitch: three copies of the planes -- kx = 0 reads a copy whose last image
// column is zero (row x = 0 would otherwise wrap to the previous row's last pixel), kx = 2 a copy whose first column is
// zero, kx = 1 the plain copy; one slack pixel of zeros before and after each plane catches the two corner reads.
// Roles (448 threads, persistent): warp 0 TMA producer (weights once: 9 x [64 n][64 k] swizzled tiles = 72 KB resident;
// one raw halo box per tile), warp 1 MMA issuer (36 x M128 N64 K16 per tile, two TMEM accumulators), warps 2-9 re-layout,
// warps 10-13 epilogue (scale/shift (+residual) (+ReLU) -> bf16 -> swizzled staging -> TMA store).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <tuple>

#include "common.cuh"
#include "engine.h"
#include "tc_common.cuh"
#include "tma_host.h"

namespace dirb200 {

namespace {

using namespace tc;

constexpr int HALO_THREADS = 448;  // producer, MMA issuer, 8 re-layout warps, 4 epilogue warps
constexpr int RELAYOUT_THREADS = 256;
constexpr int W_BYTES = 9 * 64 * 128;       // resident weights: 9 taps x [64 n][64 k] bf16
constexpr int RAW_BYTES = 4 * 64 * 128;     // largest halo: (2+2) rows x 64 px x 128 B
constexpr int LBO = 265 * 16;               // chunk-plane stride: (256 + 2 slack + 7 pad) pixels; 265 % 8 == 1 keeps the
                                            // re-layout's 16-byte stores of one pixel's 8 chunks on distinct banks
constexpr int PLANE_BYTES = 8 * LBO;        // 8 chunks of 8 channels
constexpr int OUT_BYTES = 128 * 128;

struct HaloArgs {
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;  // optional residual [M][64]
  int W, R, tiles_per_img, total_tiles, relu;
};

struct HaloBars {
  uint64_t wbar, raw_full, raw_empty, planes_full, planes_empty, tfull[2], tempty[2];
  uint32_t tmem_ptr;
};

// the 128B-swizzled regions (weights, output staging) must sit on 1024-byte boundaries: the swizzle is a function of
// address bits 7..9, and the epilogue's manual XOR assumes they equal the row index
constexpr int OFF_W = 0;
constexpr int OFF_OUT = OFF_W + W_BYTES;
constexpr int OFF_RAW = OFF_OUT + OUT_BYTES;
constexpr int OFF_PLANES = OFF_RAW + RAW_BYTES;
constexpr int OFF_AFF = OFF_PLANES + 3 * PLANE_BYTES;
static_assert(OFF_OUT % 1024 == 0 && OFF_RAW % 128 == 0 && OFF_PLANES % 16 == 0, "smem carve-up alignment");
constexpr int OFF_BARS = OFF_AFF + 2 * 64 * 4;
constexpr int HALO_SMEM = 1024 + OFF_BARS + 128;

__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3x3_c64_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ CUtensorMap tmY, const HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem + OFF_W;
  uint8_t* sRaw = smem + OFF_RAW;
  uint8_t* sPl = smem + OFF_PLANES;  // [left | mid | right]
  uint8_t* sOut = smem + OFF_OUT;
  float* s_scale = reinterpret_cast<float*>(smem + OFF_AFF);
  float* s_shift = s_scale + 64;
  HaloBars* bars = reinterpret_cast<HaloBars*>(smem + OFF_BARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = a.W, R = a.R, Np = (R + 2) * W;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
    mbar_init(&bars->wbar, 1);
    mbar_init(&bars->raw_full, 1);
    mbar_init(&bars->raw_empty, RELAYOUT_THREADS);
    mbar_init(&bars->planes_full, RELAYOUT_THREADS);
    mbar_init(&bars->planes_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tfull[i], 1);
      mbar_init(&bars->tempty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the slack pixels and the padding of the planes must read as zero: clear everything once
  for (int i = threadIdx.x; i < 3 * PLANE_BYTES / 16; i += HALO_THREADS)
    reinterpret_cast<uint4*>(sPl)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = a.scale[threadIdx.x];
    s_shift[threadIdx.x] = a.shift[threadIdx.x];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer
      if ((int)blockIdx.x < a.total_tiles) {
        mbar_expect_tx(&bars->wbar, W_BYTES);
        for (int t = 0; t < 9; ++t) tma_load_2d(&tmW, &bars->wbar, sW + t * 8192, t * 64, 0);
      }
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++i) {
        const int b = tile / a.tiles_per_img, y0 = (tile - b * a.tiles_per_img) * R;
        mbar_wait(&bars->raw_empty, (i & 1) ^ 1);
        mbar_expect_tx(&bars->raw_full, (uint32_t)Np * 128);
        tma_load_4d(&tmX, &bars->raw_full, sRaw, 0, 0, y0 - 1, b);  // rows y0-1 .. y0+R: out-of-image rows arrive as zeros
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================== MMA issuer
      constexpr uint32_t ID = idesc(64, 1u);
      mbar_wait(&bars->wbar, 0);
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++i) {
        const uint32_t buf = i & 1;
        mbar_wait(&bars->tempty[buf], ((i >> 1) & 1) ^ 1);
        mbar_wait(&bars->planes_full, i & 1);
        fence_after();
        const uint32_t d = tmem_base + buf * 64;
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t - ky * 3;
          const uint32_t abase = s32(sPl + kx * PLANE_BYTES) + (uint32_t)(ky * W + kx) * 16;
          const uint64_t db = desc128(s32(sW + t * 8192));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma(d, desc_nosw(abase + ks * 2 * LBO, LBO, 128), db + 2 * ks, ID, acc);
            acc = 1;
          }
        }
        umma_commit(&bars->planes_empty);  // the planes may be overwritten once these MMAs have read them
        umma_commit(&bars->tfull[buf]);
      }
    }
  } else if (warp < 10) {
    // ===================== re-layout: raw [pixel][8 chunks x 16 B] -> three chunk-major planes
    const int t = threadIdx.x - 64;
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++i) {
      mbar_wait(&bars->raw_full, i & 1);
      mbar_wait(&bars->planes_empty, (i & 1) ^ 1);  // previous tile's MMAs are done with the planes
      const uint4* src = reinterpret_cast<const uint4*>(sRaw);
      const uint4 zero = make_uint4(0, 0, 0, 0);
      for (int idx = t; idx < Np * 8; idx += RELAYOUT_THREADS) {
        const int p = idx >> 3, c = idx & 7;
        const uint4 v = src[idx];
        const int x = p & (W - 1);  // W is a power of two
        const uint32_t off = (uint32_t)c * LBO + (uint32_t)(p + 1) * 16;
        *reinterpret_cast<uint4*>(sPl + off) = x == W - 1 ? zero : v;                    // kx = 0 copy
        *reinterpret_cast<uint4*>(sPl + PLANE_BYTES + off) = v;                          // kx = 1
        *reinterpret_cast<uint4*>(sPl + 2 * PLANE_BYTES + off) = x == 0 ? zero : v;      // kx = 2
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bars->raw_empty);
      mbar_arrive(&bars->planes_full);
    }
  } else {
    // ===================== epilogue
    const int et = threadIdx.x - 320;  // 0..127
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++i) {
      const uint32_t buf = i & 1;
      const int64_t m = (int64_t)tile * 128 + row;
      uint4 rres[8];
      if (a.res) {
#pragma unroll
        for (int q = 0; q < 8; ++q) rres[q] = __ldg(reinterpret_cast<const uint4*>(a.res + m * 64) + q);
      }
      mbar_wait(&bars->tfull[buf], (i >> 1) & 1);
      fence_after();
      float v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + buf * 64;
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tmem_ld_wait();
      fence_before();
      mbar_arrive(&bars->tempty[buf]);
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has read the staging
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(v[q * 8 + e], s_scale[q * 8 + e], s_shift[q * 8 + e]);
        if (a.res) {
          const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&rres[q]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(hr[e]);
            o[2 * e] += f.x;
            o[2 * e + 1] += f.y;
          }
        }
        if (a.relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        uint4 pk;
        __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        *reinterpret_cast<uint4*>(sOut + row * 128 + (((uint32_t)q ^ swz) << 4)) = pk;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        tma_store_2d(&tmY, sOut, 0, tile * 128);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

}  // namespace

bool conv_halo_supported(const ConvLayer& L, int B, int H, int W) {
  if (!L.w16 || L.wmap_bn != 64 || !tma::get_encode()) return false;
  if (L.kh != 3 || L.kw != 3 || L.stride != 1 || L.pad != 1 || L.Cin != 64 || L.Cout != 64) return false;
  if (W != 16 && W != 32 && W != 64) return false;
  const int R = 128 / W;
  return H >= R && H % R == 0 && B > 0;
}

int launch_conv_halo(const ConvLayer& L, const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* res, int B, int H,
                     int W, cudaStream_t st) {
  const int R = 128 / W;
  typedef std::tuple<const void*, int, int, int, int> Key;
  static thread_local tma::MapCache<Key> xcache, ycache;
  CUtensorMap tmX, tmY;
  Key kx(x, B, H, W, R);
  if (!xcache.find(kx, &tmX)) {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)(R + 2), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (tma::get_encode()(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DIRB200_E_CUDA;
    xcache.put(kx, tmX);
  }
  const int M = B * H * W;
  Key ky(y, M, 0, 0, 0);
  if (!ycache.find(ky, &tmY)) {
    cuuint64_t dims[2] = {64, (cuuint64_t)M};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    if (tma::get_encode()(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DIRB200_E_CUDA;
    ycache.put(ky, tmY);
  }
  HaloArgs a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.res = res;
  a.W = W;
  a.R = R;
  a.tiles_per_img = H / R;
  a.total_tiles = B * a.tiles_per_img;
  a.relu = L.relu;
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(conv3x3_c64_halo_kernel), HALO_SMEM) != cudaSuccess)
    return DIRB200_E_CUDA;
  const int grid = a.total_tiles < tma::num_sms() ? a.total_tiles : tma::num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = HALO_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, conv3x3_c64_halo_kernel, tmX, L.wmap, tmY, a) != cudaSuccess) return DIRB200_E_CUDA;
  return DIRB200_OK;
}

}  // namespace dirb200
