// fp32-grade implicit-GEMM convolution on the tcgen05 tensor cores (precision = fp32): error-compensated 3xTF32 with
// the accumulation promoted to round-to-nearest fp32 registers every few MMAs. NHWC fp32 activations, weights
// [Cout][(ky,kx,ci)] pre-split into tf32 hi / lo parts at finalize. Replaces conv_simt.cu for every tensor-core shaped
// layer of the parity configuration (models/backbone/resnet.py:120-140, models/backbone/hourglass.py:55-70,
// models/dir.py:57-62,227-241,404-420).
//
// Why it is built this way (measured on B200, profiles/mma_probe_r2.txt):
//  * kind::tf32 TRUNCATES its fp32 operands to 10 mantissa bits, so the hardware itself forms A_hi = trunc(A) from the
//    raw activation tile; only A_lo = rn_tf32(A - trunc(A)) has to be produced in software (an elementwise pass over
//    the 16 KB tile by four "splitter" warps, position for position, so the 128B swizzle never has to be decoded).
//  * D = A_hi*W_hi + (A_lo*W_hi + A_hi*W_lo); A_lo*W_lo (2^-22) is dropped. W_hi and W_lo tiles sit back to back in shared
//    memory, so ONE N = 2*BN instruction computes A_hi*[W_hi | W_lo] (main term into columns 0..BN-1, first cross term into
//    BN..2BN-1) and a second N = BN instruction adds A_lo*W_hi to the cross columns: 2 instructions per k-step instead
//    of 3, A is read from shared memory once instead of twice, and the wide instruction is the efficient one (an M=128 MMA
//    costs 107 cycles at N=128 but only 171 at N=256, profiles/mma_probe_r2.txt). The cross terms are 2^-11 of the main
//    term and live in their own columns, so their rounding never touches the main sum.
//  * the TMEM accumulator rounds TOWARD ZERO on every MMA: a biased 2^-24 relative loss per instruction, which over the
//    2304 k-steps of the 18432-deep attention convolution would reach 1e-4. The main accumulator is therefore drained
//    every CHUNK_KB k-blocks (8 accumulating MMAs) into fp32 registers of the epilogue warps with ordinary round-to-nearest
//    adds (main + cross columns together) while the MMA warp continues into the other TMEM buffer (ping-pong), bounding
//    the bias at ~4e-7.
// Roles per CTA (448 threads, persistent over tiles of 128 pixels x BN channels, k-block = 32 channels of one tap):
//   warp 0    TMA producer: A box {32 ch, wbox*s, hbox*s, nbox} (im2col, padding and stride are TMA coordinates),
//             W_hi and W_lo boxes {32 k, BN rows}; `stages`-deep mbarrier ring
//   warp 1    TMEM allocation (4 x BN columns) + single-thread tcgen05.mma.kind::tf32 issue (4 wide + 4 narrow MMAs per k-block)
//   warps 2-5 splitter: A_lo tile of each stage
//   warps 6-13 accumulate/epilogue, two groups of four warps owning one half of the BN columns each: tcgen05.ld of each
//             finished chunk -> += registers; at the end of the tile add the cross-term accumulator, apply scale/shift
//             (+residual, fetched in batches of eight 16-byte loads per thread) (+ReLU) and store fp32 rows
// Stem variant (7x7 stride 2, Cin = 3; models/backbone/resnet.py:176): the image is repacked to a zero-padded NHWC4
//   fp32 buffer; a k-block is one kernel row (8 px x 4 ch = 32 floats = 128 B) and the A box comes from an
//   OVERLAPPING-stride TMA view of that buffer (consecutive output pixels are 2 px = 32 B apart).
// nsplit = 1 runs the same pipeline as plain TF32 (one MMA per k-step, what PyTorch's cuDNN default does on this GPU); the
// splitter warps then round the A tile to tf32 in place (nearest-even), because the MMA's own truncation is one-sided.
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

constexpr int BM = 128;
constexpr int KB = 32;          // channels per k-block: one 128-byte fp32 row = one swizzle atom
constexpr int NUM_THREADS = 448;  // producer + MMA + 4 splitter + 8 accumulate/epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int MAX_STAGES = 4;
constexpr int CHUNK_KB = 2;     // k-blocks accumulated in TMEM before promotion (2 x 4 k-steps = 8 MMAs)
constexpr int A_TILE = BM * 128;
constexpr int OUT_STAGE = BM * 128;  // epilogue staging of one 32-column fp32 box (128B-swizzled), one per epilogue group

struct T32Args {
  const float* scale;
  const float* shift;
  const float* res;  // optional residual [M][Cout]
  float* y;          // [M][Cout]
  int M, Cout, Ho, Wo;
  int stride, pad, kw;
  int taps, cblocks;  // K loop = taps * cblocks k-blocks
  int relu;
  int m_tiles, n_tiles, stages, raster_m;
  int nsplit;  // 3: error-compensated 3xTF32, 1: plain TF32
  int stem;    // 1: k-block = kernel row ky of the 7x7 stem, A from the overlapping view of the padded NHWC4 image
};

template <int BN>
struct Cfg32 {
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;  // A, A_lo, W_hi, W_lo
  static constexpr uint32_t TMEM_COLS = 4 * BN;                // 2 ping-pong buffers of [main BN | cross BN] columns
  static int stages() {
    int s = (192 * 1024) / STAGE_BYTES;
    return s > MAX_STAGES ? MAX_STAGES : s;
  }
  static int smem_bytes(int stages) { return 1024 + stages * STAGE_BYTES + 2 * OUT_STAGE + 2 * BN * 4 + 256; }
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                 const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmY, const T32Args a) {
  using Cfg = Cfg32<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = a.stages;
  uint8_t* s_out = smem + stages * Cfg::STAGE_BYTES;  // 2 x 16 KB output staging (one per epilogue group)
  float* s_scale = reinterpret_cast<float*>(s_out + 2 * OUT_STAGE);  // affine of the current n-tile
  float* s_shift = s_scale + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* full_bar = bars;                      // TMA bytes of a stage have landed
  uint64_t* empty_bar = bars + MAX_STAGES;        // the MMAs reading a stage have completed
  uint64_t* split_bar = bars + 2 * MAX_STAGES;    // A_lo of a stage is written (128 arrivals)
  uint64_t* main_full = bars + 3 * MAX_STAGES;    // [2] a chunk's accumulator is complete
  uint64_t* main_empty = main_full + 2;           // [2] ... and has been drained (256 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(main_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.n_tiles;
  const int nkb = a.taps * a.cblocks;
  const bool comp = a.nsplit == 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBhi);
    tma_prefetch_desc(&tmBlo);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&main_full[i], 1);
      mbar_init(&main_empty[i], EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_ptr_smem)),
                 "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above overlaps the previous kernel's tail

  auto stage_A = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_Alo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + A_TILE; };
  auto stage_Bhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * A_TILE; };
  auto stage_Blo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * A_TILE + Cfg::B_TILE; };
  auto tile_m0 = [&](int tile) { return (a.raster_m ? tile % a.m_tiles : tile / a.n_tiles) * BM; };
  auto tile_n0 = [&](int tile) { return (a.raster_m ? tile / a.m_tiles : tile % a.n_tiles) * BN; };

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = A_TILE + Cfg::B_TILE * (comp ? 2 : 1);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = tile_m0(tile), n0 = tile_n0(tile);
        const int wo0 = m0 % a.Wo;
        const int ho0 = (m0 / a.Wo) % a.Ho;
        const int b0 = m0 / (a.Wo * a.Ho);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], tx);
          if (a.stem) {  // kernel row ky = kb: window {8 px x 4 ch} per output pixel, rows 2*ho + ky - 3
            tma_load_4d(&tmA, &full_bar[s], stage_A(s), 0, wo0, ho0 * 2 + kb - 3, b0);
          } else {
            const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            tma_load_4d(&tmA, &full_bar[s], stage_A(s), cb * KB, wo0 * a.stride + kx - a.pad,
                        ho0 * a.stride + ky - a.pad, b0);
          }
          tma_load_2d(&tmBhi, &full_bar[s], stage_Bhi(s), kb * KB, n0);
          if (comp) tma_load_2d(&tmBlo, &full_bar[s], stage_Blo(s), kb * KB, n0);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================== MMA issuer
      constexpr uint32_t ID = idesc(BN, 2u), ID_WIDE = idesc(2 * BN, 2u);
      int s = 0;
      uint32_t ph = 0, g = 0;  // g: running chunk counter (ping-pong of the [main | cross] accumulator pair)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          const int kc = kb % CHUNK_KB;
          if (kc == 0) {  // a new chunk: its accumulator buffer must have been drained
            mbar_wait(&main_empty[g & 1], ((g >> 1) & 1) ^ 1);
            fence_after();
          }
          const uint32_t d_main = tmem_base + (g & 1) * 2 * BN, d_cross = d_main + BN;
          mbar_wait(&full_bar[s], ph);
          if (!comp) mbar_wait(&split_bar[s], ph);  // plain TF32: the splitter rounds the A tile in place first
          fence_after();
          const uint64_t dA = desc128(s32(stage_A(s))), dAlo = desc128(s32(stage_Alo(s)));
          const uint64_t dB = desc128(s32(stage_Bhi(s)));  // W_hi rows 0..BN-1, W_lo rows BN..2BN-1 (adjacent tiles)
          if (comp) {
            // A_hi * [W_hi | W_lo]: needs only TMA data, so it goes first and covers the splitter's latency
#pragma unroll
            for (int k = 0; k < 4; ++k)  // +32 B per K=8 step inside the swizzle atom
              umma_tf32(d_main, dA + 2 * k, dB + 2 * k, ID_WIDE, (kc | k) ? 1u : 0u);
            mbar_wait(&split_bar[s], ph);
            fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(d_cross, dAlo + 2 * k, dB + 2 * k, ID, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(d_main, dA + 2 * k, dB + 2 * k, ID, (kc | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have consumed it
          if (kc == CHUNK_KB - 1 || kb == nkb - 1) {
            umma_commit(&main_full[g & 1]);
            ++g;
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== splitter. 3xTF32: A_lo[i] = rn_tf32(A[i] - trunc_tf32(A[i])), same byte position in its own
    // tile. Plain TF32: A[i] = rn_tf32(A[i]) in place — the MMA would TRUNCATE, a one-sided 2^-11 relative error per
    // operand that compounds through ~60 layers into a 5x larger drift than round-to-nearest (what cuDNN's TF32 does).
    {
      const int tid = threadIdx.x - 64;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          const uint4* src = reinterpret_cast<const uint4*>(stage_A(s));
          uint4* dst = reinterpret_cast<uint4*>(comp ? stage_Alo(s) : stage_A(s));
#pragma unroll
          for (int i = 0; i < A_TILE / 16 / 128; ++i) {
            uint4 v = src[tid + i * 128];
            uint32_t* e = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (comp) {
                const float x = __uint_as_float(e[j]);
                const float lo = x - __uint_as_float(e[j] & 0xFFFFE000u);  // exact; the MMA truncates x the same way
                e[j] = (__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u;      // round to tf32 (the MMA would truncate)
              } else {
                e[j] = (e[j] + 0xFFFu + ((e[j] >> 13) & 1u)) & 0xFFFFE000u;  // round-to-nearest-even to tf32
              }
            }
            dst[tid + i * 128] = v;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
          mbar_arrive(&split_bar[s]);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== accumulate + epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; warps 6-9 own the first
    // half of the tile's columns, warps 10-13 the second
    constexpr int HN = BN / 2;
    const int et = threadIdx.x - 192;  // 0..255
    const int half = et >> 7;
    const int lane_base = (warp & 3) * 32;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const int nchunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;
    uint32_t g = 0;
    int cur_n0 = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = tile_m0(tile), n0 = tile_n0(tile);
      if (n0 != cur_n0) {  // (re)stage the per-channel affine of this n-tile
        if (cur_n0 >= 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done with the previous one
        if (et < BN) {
          s_scale[et] = __ldg(a.scale + n0 + et);
          s_shift[et] = __ldg(a.shift + n0 + et);
        }
        cur_n0 = n0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      float acc[HN];
#pragma unroll
      for (int j = 0; j < HN; ++j) acc[j] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++g) {
        mbar_wait(&main_full[g & 1], (g >> 1) & 1);
        fence_after();
        const uint32_t taddr = tmem_base + lane_addr + (g & 1) * 2 * BN + half * HN;
#pragma unroll
        for (int j = 0; j < HN / 32; ++j) {
          float v[32];
          tmem_ld32(taddr + j * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
        }
        if (comp) {  // the cross terms of the same chunk
#pragma unroll
          for (int j = 0; j < HN / 32; ++j) {
            float v[32];
            tmem_ld32(taddr + BN + j * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
          }
        }
        fence_before();
        mbar_arrive(&main_empty[g & 1]);
      }
      // epilogue: 32 columns at a time through the group's 128B-swizzled staging buffer, written out by one TMA store
      // (whole 128-byte lines; rows beyond M are clipped by the tensor map)
      const int row = lane_base + lane;
      const int m = m0 + row;
      const bool live = m < a.M;
      const float* rrow = (a.res && live) ? a.res + (size_t)m * a.Cout + n0 + half * HN : nullptr;
      const float* sc = s_scale + half * HN;
      const float* sh = s_shift + half * HN;
      uint8_t* stg = s_out + half * OUT_STAGE + row * 128;
      const uint32_t swz = (uint32_t)(row & 7);
      const int gt = et & 127;  // thread index inside the group
#pragma unroll
      for (int jb = 0; jb < HN; jb += 32) {
        float4 r[8];
        if (rrow) {  // eight 16-byte loads in flight per thread
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = __ldg(reinterpret_cast<const float4*>(rrow + jb + q * 4));
        }
        if (gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has read the buffer
        asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = jb + q * 4;
          const float4 s4 = *reinterpret_cast<const float4*>(sc + j);
          const float4 h4 = *reinterpret_cast<const float4*>(sh + j);
          float4 o;
          o.x = fmaf(acc[j + 0], s4.x, h4.x);
          o.y = fmaf(acc[j + 1], s4.y, h4.y);
          o.z = fmaf(acc[j + 2], s4.z, h4.z);
          o.w = fmaf(acc[j + 3], s4.w, h4.w);
          if (rrow) {
            o.x += r[q].x;
            o.y += r[q].y;
            o.z += r[q].z;
            o.w += r[q].w;
          }
          if (a.relu) {
            o.x = fmaxf(o.x, 0.f);
            o.y = fmaxf(o.y, 0.f);
            o.z = fmaxf(o.z, 0.f);
            o.w = fmaxf(o.w, 0.f);
          }
          *reinterpret_cast<float4*>(stg + (((uint32_t)q ^ swz) << 4)) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> TMA (async proxy) reads
        asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
        if (gt == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmY)),
                       "r"(s32(s_out + half * OUT_STAGE)), "r"(n0 + half * HN + jb), "r"(m0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if ((et & 127) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output bytes written
  }
  fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
}

// w [n] -> hi = rn_tf32(w), lo = rn_tf32(w - hi)   (round-to-nearest-even on the 13 dropped mantissa bits)
__device__ __forceinline__ float rn_tf32(float x) {
  const uint32_t b = __float_as_uint(x);
  return __uint_as_float((b + 0xFFFu + ((b >> 13) & 1u)) & 0xFFFFE000u);
}
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i];
  const float h = rn_tf32(x);
  hi[i] = h;
  lo[i] = rn_tf32(x - h);
}

// NCHW fp32 image -> zero-padded NHWC4 fp32: out[b][h][w + 3][c] (c = 3 is zero), row pitch (W + 8) pixels
__global__ void stem_pack32_kernel(const float* __restrict__ img, float4* __restrict__ out, int B, int H, int W) {
  pdl_wait();
  const int Wp = W + 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * Wp) return;
  const int wp = (int)(idx % Wp);
  const int64_t t = idx / Wp;
  const int h = (int)(t % H), b = (int)(t / H);
  const int w = wp - 3;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (w >= 0 && w < W) {
    const float* p = img + ((int64_t)b * 3 * H + h) * W + w;
    v.x = __ldg(p);
    v.y = __ldg(p + (int64_t)H * W);
    v.z = __ldg(p + 2 * (int64_t)H * W);
  }
  out[idx] = v;
}

// stem weights [64][3][7][7] -> [64][7 * 32] with k = ky*32 + kx*4 + c (zero for kx = 7 and c = 3)
__global__ void stem_pack_weight32_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 224) return;
  const int k = idx % 224, n = idx / 224;
  const int ky = k / 32, r = k % 32, kx = r / 4, c = r % 4;
  out[idx] = (kx < 7 && c < 3) ? w[((n * 3 + c) * 7 + ky) * 7 + kx] : 0.f;
}

int pick_bn32(int Cout) { return Cout % 128 == 0 ? 128 : 64; }

bool weight_maps32(float* hi, float* lo, int Kpad, int Cout, int bn, CUtensorMap* mhi, CUtensorMap* mlo) {
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 4};
  cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    CUresult r = tma::get_encode()(i ? mlo : mhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i ? lo : hi, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  return true;
}

bool act_map32_cached(const float* x, int B, int H, int W, int C, int stride, const tma::Boxes& bx, const char* name,
                      CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int> Key;
  static thread_local tma::MapCache<Key> cache;
  Key key(x, B, H, W, C, stride, bx.wbox, bx.hbox, bx.nbox);
  if (cache.find(key, out)) return true;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)KB, (cuuint32_t)(bx.wbox * stride), (cuuint32_t)(bx.hbox * stride), (cuuint32_t)bx.nbox};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = tma::get_encode()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "dirb200: cuTensorMapEncodeTiled(fp32 A) failed: %d (layer %s)\n", (int)r, name);
    return false;
  }
  cache.put(key, *out);
  return true;
}

// fp32 [rows][cols] row-major as 32-column x 128-row boxes with 128B swizzle (epilogue store)
bool out_map32_cached(const float* y, int rows, int cols, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int> Key;
  static thread_local tma::MapCache<Key> cache;
  Key key(y, rows, cols);
  if (cache.find(key, out)) return true;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)BM};
  cuuint32_t es[2] = {1, 1};
  if (tma::get_encode()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(y), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  cache.put(key, *out);
  return true;
}

template <int BN>
int launch_t32(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, T32Args a, cudaStream_t st) {
  using Cfg = Cfg32<BN>;
  CUtensorMap tmY;
  if (!out_map32_cached(a.y, a.M, a.Cout, &tmY)) return DIRB200_E_CUDA;
  a.stages = Cfg::stages();
  const int smem = Cfg::smem_bytes(a.stages);
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(conv_tf32_kernel<BN>), smem) != cudaSuccess) return DIRB200_E_CUDA;
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < tma::num_sms() ? tiles : tma::num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, conv_tf32_kernel<BN>, tmA, tmBhi, tmBlo, tmY, a) != cudaSuccess) return DIRB200_E_CUDA;
  return DIRB200_OK;
}

}  // namespace

bool conv_tf32_supported(const ConvLayer& L, int B, int H, int W) {
  if (!L.w32hi || L.wmap32_bn == 0 || !tma::get_encode()) return false;
  if (L.Cin % KB != 0 || L.Cout % 64 != 0 || L.K != L.Kpad) return false;
  if (L.stride != 1 && L.stride != 2) return false;
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W + 2 * L.pad - L.kw) / L.stride + 1;
  if (!tma::is_pow2(Ho) || !tma::is_pow2(Wo)) return false;
  if (Wo > BM && Wo % BM != 0) return false;
  return true;
}

// Splits L.w32 into the tf32 hi / lo copies (caller-allocated, Cout*Kpad floats each) and builds their tensor maps.
int conv_tf32_prepare_weights(ConvLayer& L, float* hi, float* lo, cudaStream_t st) {
  L.wmap32_bn = 0;
  L.w32hi = L.w32lo = nullptr;
  tma::EncodeTiledFn enc = tma::get_encode();
  if (!enc || !L.w32 || !hi || !lo || L.Cin % KB != 0 || L.Cout % 64 != 0 || L.K != L.Kpad) return 0;
  const int64_t n = (int64_t)L.Cout * L.Kpad;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.w32, hi, lo, n);
  const int bn = pick_bn32(L.Cout);
  if (!weight_maps32(hi, lo, L.Kpad, L.Cout, bn, &L.wmap32hi, &L.wmap32lo)) return DIRB200_E_CUDA;
  // 64-row boxes as well: small-M launches (deep layers at small batch) use the narrower tile to fill the SMs
  L.wmap32_alt64 = bn == 128 && weight_maps32(hi, lo, L.Kpad, L.Cout, 64, &L.wmap32hi64, &L.wmap32lo64);
  L.w32hi = hi;
  L.w32lo = lo;
  L.wmap32_bn = bn;
  return DIRB200_OK;
}

int launch_conv_tf32(const ConvLayer& L, const float* x, float* y, const float* res, int B, int H, int W, int nsplit,
                     cudaStream_t st) {
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W + 2 * L.pad - L.kw) / L.stride + 1;
  const tma::Boxes bx = tma::pick_boxes(Ho, Wo, BM);
  CUtensorMap tmA;
  if (!act_map32_cached(x, B, H, W, L.Cin, L.stride, bx, L.name.c_str(), &tmA)) return DIRB200_E_CUDA;
  T32Args a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.res = res;
  a.y = y;
  a.M = B * Ho * Wo;
  a.Cout = L.Cout;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = L.stride;
  a.pad = L.pad;
  a.kw = L.kw;
  a.taps = L.kh * L.kw;
  a.cblocks = L.Cin / KB;
  a.relu = L.relu;
  a.m_tiles = (a.M + BM - 1) / BM;
  a.n_tiles = L.Cout / L.wmap32_bn;
  a.raster_m = (double)L.Cout * L.K > (double)B * H * W * L.Cin ? 1 : 0;
  a.nsplit = nsplit == 1 ? 1 : 3;
  if (L.wmap32_bn == 128 && !(L.wmap32_alt64 && a.m_tiles * a.n_tiles * 4 < tma::num_sms() * 3))
    return launch_t32<128>(tmA, L.wmap32hi, L.wmap32lo, a, st);
  a.n_tiles = L.Cout / 64;
  if (L.wmap32_bn == 128) return launch_t32<64>(tmA, L.wmap32hi64, L.wmap32lo64, a, st);
  return launch_t32<64>(tmA, L.wmap32hi, L.wmap32lo, a, st);
}

// ---- stem (7x7 s2, 3 -> 64) on the same kernel
size_t conv_tf32_stem_scratch_bytes(int B, int H, int W) { return (size_t)B * H * (W + 8) * 16; }
size_t conv_tf32_stem_weight_floats() { return 3 * 64 * 224; }  // packed + hi + lo

// `buf` = conv_tf32_stem_weight_floats() floats (caller-allocated): packed weights, then their hi / lo split
int conv_tf32_prepare_stem(ConvLayer& L, const float* w_raw, float* buf, cudaStream_t st) {
  L.tf32_stem = false;
  if (!tma::get_encode() || !buf || L.Cin != 3 || L.Cout != 64 || L.kh != 7 || L.kw != 7 || L.stride != 2 || L.pad != 3)
    return 0;
  float *packed = buf, *hi = buf + 64 * 224, *lo = buf + 2 * 64 * 224;
  stem_pack_weight32_kernel<<<(64 * 224 + 255) / 256, 256, 0, st>>>(w_raw, packed);
  split_tf32_kernel<<<(64 * 224 + 255) / 256, 256, 0, st>>>(packed, hi, lo, 64 * 224);
  if (!weight_maps32(hi, lo, 224, 64, 64, &L.wmap32hi, &L.wmap32lo)) return DIRB200_E_CUDA;
  L.tf32_stem = true;
  return DIRB200_OK;
}

bool conv_tf32_stem_supported(const ConvLayer& L, int H, int W) { return L.tf32_stem && H % 2 == 0 && W == 256; }

int launch_conv_tf32_stem(const ConvLayer& L, const float* img, float* scratch, float* y, int B, int H, int W, int nsplit,
                          cudaStream_t st) {
  const int Wp = W + 8, Ho = H / 2, Wo = W / 2;
  const int64_t n = (int64_t)B * H * Wp;
  launch_pdl(stem_pack32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, img,
             reinterpret_cast<float4*>(scratch), B, H, W);
  typedef std::tuple<const void*, int, int, int> Key;
  static thread_local tma::MapCache<Key> cache;
  Key key(scratch, B, H, W);
  CUtensorMap tm;
  if (!cache.find(key, &tm)) {
    // overlapping view of the padded NHWC4 buffer: the 8-pixel window of output pixel wo starts at padded pixel 2*wo
    cuuint64_t dims[4] = {32, (cuuint64_t)Wo, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {32, (cuuint64_t)Wp * 16, (cuuint64_t)H * Wp * 16};
    cuuint32_t box[4] = {32, (cuuint32_t)BM, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = tma::get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, scratch, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "dirb200: cuTensorMapEncodeTiled(fp32 stem A, overlapping strides) failed: %d\n", (int)r);
      return DIRB200_E_CUDA;
    }
    cache.put(key, tm);
  }
  T32Args a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.res = nullptr;
  a.y = y;
  a.M = B * Ho * Wo;
  a.Cout = 64;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = 2;
  a.pad = 3;
  a.kw = 7;
  a.taps = 7;
  a.cblocks = 1;
  a.relu = L.relu;
  a.m_tiles = (a.M + BM - 1) / BM;
  a.n_tiles = 1;
  a.nsplit = nsplit == 1 ? 1 : 3;
  a.stem = 1;
  return launch_t32<64>(tm, L.wmap32hi, L.wmap32lo, a, st);
}

}  // namespace dirb200
