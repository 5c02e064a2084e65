// mixSTE blocks 1..3 + head (transformer/mixSTE.py:76-97,129-141,194-205) on the tcgen05 tensor cores — the bf16
// configuration's version of ste_kernel (joint.cu keeps the fp32 CUDA-core one for the fp32 configuration).
//
// One CTA = two images. Image s owns token rows 64s .. 64s+41 of a 128-row tile (rows 42..63 of each slot are
// padding), so every warp of the row-per-thread epilogues belongs to exactly one image. All seven contractions of a
// block run as M=128 tcgen05.mma (bf16 operands, fp32 accumulation in TMEM):
//   QKV  (128x384x128)  A = LN1(x)      B = Wqkv          D = ACC[0:384]
//   S_h  (128x128x32)   A = Q[:,h]      B = K[:,h]        D = S buffers (two heads in flight); only the two diagonal
//                                                              64x64 blocks (own image) are read back
//   O_h  (128x32x128)   A = P_h         B = V_h^T         D = O[32h:32h+32]   (P_h is block diagonal: zero across images)
//   proj (128x128x128)  A = O           B = Wproj         D = X  (accumulate: the residual add is the MMA itself)
//   fc1  (128x256x128)  A = LN2(x)      B = Wfc1          D = ACC[0:256]
//   fc2  (128x128x256)  A = gelu(h)     B = Wfc2          D = X  (accumulate)
//   head (128x64x128)   A = LN(x)       B = Whead         D = ACC[0:64]
// The residual stream X lives in TMEM (128 fp32 columns) for the whole kernel; LayerNorm / softmax / GELU run on
// 8 warps with two threads per token row (tcgen05.ld 32x32b -> registers -> bf16 -> 128B-swizzled K-major smem, the
// canonical UMMA operand layout). Weights are packed at finalize time into the exact smem image of their operand
// tiles ([128 n][64 k] bf16, 16 KB), in consumption order, so a dedicated producer warp streams them with one
// cp.async.bulk per tile through a 4-deep mbarrier ring, starting before the previous kernel has finished (PDL).
// Roles: warps 0-7 row epilogues (two threads per token row: column halves; TMEM lane quarter = warp % 4),
// warp 8 weight producer, warp 9 TMEM alloc + single-thread MMA issue.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dirb200 {

namespace {

using namespace tc;

constexpr int NT = 42;
constexpr int TILE = 128 * 128;  // bytes of one [128 rows][64 bf16] operand tile
constexpr int RING = 4;
constexpr int TC_THREADS = 320;  // warps 0-7 row epilogues, warp 8 weight producer, warp 9 MMA issue
constexpr int OFF_H = 0;             // 2 tiles: LN output (A of QKV/fc1/head), attention output O, P of odd heads
constexpr int OFF_Q = 2 * TILE;      // 2 tiles: Q (pre-scaled)      \ the 4 tiles of gelu(fc1) alias Q and K
constexpr int OFF_K = 4 * TILE;      // 2 tiles: K                   /
constexpr int OFF_V = 6 * TILE;      // 2 tiles: V^T [channel][token]
constexpr int OFF_P = 8 * TILE;      // 2 tiles: P of even heads
constexpr int OFF_RING = 10 * TILE;  // RING weight tiles
constexpr int OFF_BAR = OFF_RING + RING * TILE;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 128;
constexpr int TILES_PER_BLOCK = 16;  // qkv 6, proj 2, fc1 4, fc2 4
constexpr int NUM_WTILES = 3 * TILES_PER_BLOCK + 2;
constexpr float QK_SCALE = 0.17677669529663688110f;  // 32^-0.5 (mixSTE.py:59), folded into Wq / bq
// TMEM columns
constexpr uint32_t COL_X = 0, COL_ACC = 128, COL_S0 = 128, COL_S1 = 256, COL_O = 384;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32 consecutive K-columns (c0 = 0 or 32 inside a 64-column tile) of operand row `row`: four swizzled 16-byte chunks
__device__ __forceinline__ void store_row32(uint8_t* tile, int row, int c0, const float* v) {
  uint8_t* rp = tile + row * 128;
  const int sw = row & 7;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack2(v[8 * q + 0], v[8 * q + 1]);
    u.y = pack2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack2(v[8 * q + 4], v[8 * q + 5]);
    u.w = pack2(v[8 * q + 6], v[8 * q + 7]);
    *reinterpret_cast<uint4*>(rp + ((((c0 >> 3) + q) ^ sw) << 4)) = u;
  }
}

// erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16 rounding of the result) — ~14 instructions
// instead of erff's ~40; the fp32 configuration (joint.cu) keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = 1.f - p * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// Row statistics are shared by the two threads of a row (column halves) through `red` and a 256-thread named barrier.
__device__ __forceinline__ float pair_sum(float v, float* red, uint32_t& k) {
  float* r = red + (k & 1) * 256;
  r[threadIdx.x] = v;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  ++k;
  return v + r[threadIdx.x ^ 128];
}
// LayerNorm over the 128 channels of a row (two-pass, like ATen); this thread holds columns 64*half .. +63.
__device__ __forceinline__ void ln_stats(const float (&x)[64], float eps, float* red, uint32_t& k, float& mu, float& rs) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x[i];
  mu = pair_sum(s, red, k) * (1.f / 128.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) q = fmaf(x[i] - mu, x[i] - mu, q);
  rs = rsqrtf(pair_sum(q, red, k) * (1.f / 128.f) + eps);
}
// -> bf16 A operand: this thread's 64 columns are exactly row `row` of tile `half`
__device__ __forceinline__ void layernorm_to_tile(const float (&x)[64], const float* __restrict__ g,
                                                  const float* __restrict__ b, float eps, uint8_t* tile, int row,
                                                  float* red, uint32_t& k) {
  float mu, rs;
  ln_stats(x, eps, red, k, mu, rs);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + 32 * c + i));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + 32 * c + i));
      o[i + 0] = fmaf((x[32 * c + i + 0] - mu) * rs, gg.x, bb.x);
      o[i + 1] = fmaf((x[32 * c + i + 1] - mu) * rs, gg.y, bb.y);
      o[i + 2] = fmaf((x[32 * c + i + 2] - mu) * rs, gg.z, bb.z);
      o[i + 3] = fmaf((x[32 * c + i + 3] - mu) * rs, gg.w, bb.w);
    }
    store_row32(tile, row, c * 32, o);
  }
}
// same LayerNorm, result back into x (the shared spatial_norm, mixSTE.py:200)
__device__ __forceinline__ void layernorm_inplace(float (&x)[64], const float* __restrict__ g,
                                                  const float* __restrict__ b, float eps, float* red, uint32_t& k) {
  float mu, rs;
  ln_stats(x, eps, red, k, mu, rs);
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + i));
    x[i + 0] = fmaf((x[i + 0] - mu) * rs, gg.x, bb.x);
    x[i + 1] = fmaf((x[i + 1] - mu) * rs, gg.y, bb.y);
    x[i + 2] = fmaf((x[i + 2] - mu) * rs, gg.z, bb.z);
    x[i + 3] = fmaf((x[i + 3] - mu) * rs, gg.w, bb.w);
  }
}
__device__ __forceinline__ void add_bias32(float* v, const float* __restrict__ b, float sc) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + i));
    v[i + 0] = fmaf(bb.x, sc, v[i + 0]);
    v[i + 1] = fmaf(bb.y, sc, v[i + 1]);
    v[i + 2] = fmaf(bb.z, sc, v[i + 2]);
    v[i + 3] = fmaf(bb.w, sc, v[i + 3]);
  }
}

struct Bars {
  uint64_t full[RING], empty[RING], ready, done;
  uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
ste_tc_kernel(const float* __restrict__ xin, float* __restrict__ yout, SteWeights w, const uint8_t* __restrict__ wpk,
              int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    mbar_init(&bars->ready, 256);
    mbar_init(&bars->done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {  // the whole TMEM: X 128 + ACC 384 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_ptr;

  if (warp == 8) {
    // ===================================================== weight producer (finalize-time data only: no pdl_wait)
    if (lane == 0) {
      for (int t = 0; t < NUM_WTILES; ++t) {
        const int slot = t % RING;
        if (t >= RING) mbar_wait(&bars->empty[slot], ((t / RING) - 1) & 1);
        const uint32_t bytes = t < 3 * TILES_PER_BLOCK ? TILE : TILE / 2;
        const uint8_t* src = wpk + (t < 3 * TILES_PER_BLOCK ? (size_t)t * TILE
                                                             : (size_t)3 * TILES_PER_BLOCK * TILE + (t - 3 * TILES_PER_BLOCK) * (TILE / 2));
        mbar_expect_tx(&bars->full[slot], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(smem + OFF_RING + slot * TILE)),
                     "l"(src), "r"(bytes), "r"(s32(&bars->full[slot]))
                     : "memory");
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t t = 0, ph = 0;
      const uint32_t sb = s32(smem);
      auto wait_ready = [&]() {
        mbar_wait(&bars->ready, ph & 1);
        fence_after();
      };
      auto commit_done = [&]() {
        umma_commit(&bars->done);
        ++ph;
      };
      // D[:, dcol : dcol+N] (+)= A(tiles at a_off, nkb k-blocks) * Wtile^T, weight tiles taken from the ring in order
      auto gemm_w = [&](uint32_t dcol, int a_off, int nkb, int n, bool acc_first) {
        for (int kb = 0; kb < nkb; ++kb, ++t) {
          const int slot = t % RING;
          mbar_wait(&bars->full[slot], (t / RING) & 1);
          fence_after();
          const uint64_t da = desc128(sb + a_off + kb * TILE), db = desc128(sb + OFF_RING + slot * TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma(tmem + dcol, da + 2 * k, db + 2 * k, idesc(n), (acc_first || kb || k) ? 1u : 0u);
          umma_commit(&bars->empty[slot]);
        }
      };
      auto scores = [&](int h, uint32_t dcol) {  // S_h = Q[:, 32h:32h+32] K[:, 32h:32h+32]^T
        const uint32_t off = (h >> 1) * TILE + (h & 1) * 64;
        const uint64_t da = desc128(sb + OFF_Q + off), db = desc128(sb + OFF_K + off);
        umma(tmem + dcol, da, db, idesc(128), 0u);
        umma(tmem + dcol, da + 2, db + 2, idesc(128), 1u);
      };
      auto pv = [&](int h, int p_off) {  // O[:, 32h:32h+32] = P_h V[:, 32h:32h+32]
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = desc128(sb + p_off + kb * TILE), db = desc128(sb + OFF_V + kb * TILE + h * 32 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma(tmem + COL_O + 32 * h, da + 2 * k, db + 2 * k, idesc(32), (kb || k) ? 1u : 0u);
        }
      };
      for (int l = 0; l < 3; ++l) {
        wait_ready();  // H = LN1(x)
        for (int nc = 0; nc < 3; ++nc) gemm_w(COL_ACC + 128 * nc, OFF_H, 2, 128, false);
        commit_done();
        wait_ready();  // Q, K, V^T in smem
        scores(0, COL_S0);
        scores(1, COL_S1);
        commit_done();
        wait_ready();  // P0 (OFF_P), P1 (OFF_H); S buffers drained
        pv(0, OFF_P);
        pv(1, OFF_H);
        scores(2, COL_S0);
        scores(3, COL_S1);
        commit_done();
        wait_ready();  // P2, P3
        pv(2, OFF_P);
        pv(3, OFF_H);
        commit_done();
        wait_ready();  // H = O (bf16)
        gemm_w(COL_X, OFF_H, 2, 128, true);  // x += O Wproj^T
        commit_done();
        wait_ready();  // X = x + proj_b, H = LN2(x)
        for (int nc = 0; nc < 2; ++nc) gemm_w(COL_ACC + 128 * nc, OFF_H, 2, 128, false);
        commit_done();
        wait_ready();  // gelu(fc1) in the Q/K tiles
        gemm_w(COL_X, OFF_Q, 4, 128, true);  // x += h Wfc2^T
        commit_done();
      }
      wait_ready();  // H = head LN(x)
      gemm_w(COL_ACC, OFF_H, 2, 64, false);
      commit_done();
    }
  } else {
    // ===================================================== row epilogues: two threads per token row (column halves)
    pdl_wait();
    const int row = threadIdx.x & 127, half = threadIdx.x >> 7;  // warps 0-3: half 0, warps 4-7: half 1
    const int slot = row >> 6, tok = row & 63;
    const int img = blockIdx.x * 2 + slot;
    const bool valid = tok < NT && img < B;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // a warp may touch TMEM lanes 32*(warp%4)..+31
    uint8_t* Hh = smem + OFF_H + half * TILE;       // this thread's half of a 128-column A operand = tile `half`
    float* red = reinterpret_cast<float*>(smem + OFF_P);  // LN exchange scratch: P is live only inside attention
    uint32_t ph = 0, rk = 0;
    auto signal_ready = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fence_before();
      mbar_arrive(&bars->ready);
    };
    auto wait_done = [&]() {
      mbar_wait(&bars->done, ph & 1);
      ++ph;
      fence_after();
    };
    const uint32_t xcol = COL_X + 64 * half;
    float x[64];
    {
      const float* src = xin + ((size_t)img * NT + tok) * 128 + 64 * half;
      const float* pos = w.pos + tok * 128 + 64 * half;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
          const float4 p = __ldg(reinterpret_cast<const float4*>(pos + i));
          v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        }
        x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
      }
      tmem_st32(trow + xcol, x);
      tmem_st32(trow + xcol + 32, x + 32);
      tmem_st_wait();
      layernorm_to_tile(x, w.blk[0].n1w + 64 * half, w.blk[0].n1b + 64 * half, 1e-6f, Hh, row, red, rk);
    }
    for (int l = 0; l < 3; ++l) {
      const SteWeights::Block& Bk = w.blk[l];
      signal_ready();
      wait_done();  // ---- QKV accumulators ready: this thread takes columns 64*half..+63 of each of Q, K, V
      {
        float v[32];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {  // Q (pre-scaled) and K rows, +bias -> K-major operand tiles
          const int qk = c >> 1, cc = c & 1;  // qk 0: Q, 1: K
          tmem_ld32(trow + COL_ACC + 128 * qk + 64 * half + 32 * cc, v);
          tmem_ld_wait();
          add_bias32(v, Bk.qkv_b + 128 * qk + 64 * half + 32 * cc, qk ? 1.f : QK_SCALE);
          store_row32(smem + (qk ? OFF_K : OFF_Q) + half * TILE, row, cc * 32, v);
        }
        __nv_bfloat16* vt = reinterpret_cast<__nv_bfloat16*>(smem + OFF_V + slot * TILE);
        const int cch = tok >> 3, e = tok & 7;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // V^T[channel][token]: this thread's token is one K column
          const int ch0 = 64 * half + 32 * c;
          tmem_ld32(trow + COL_ACC + 256 + ch0, v);
          tmem_ld_wait();
          add_bias32(v, Bk.qkv_b + 256 + ch0, 1.f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ch = ch0 + i;
            vt[ch * 64 + ((cch ^ (ch & 7)) << 3) + e] = __float2bfloat16_rn(v[i]);
          }
        }
      }
      signal_ready();
#pragma unroll 1
      for (int hp = 0; hp < 2; ++hp) {
        wait_done();  // ---- S of heads 2hp (half 0) and 2hp+1 (half 1) ready, the previous pair's P consumed
        {
          float sv[64];
          const uint32_t scol = (half ? COL_S1 : COL_S0) + 64 * slot;  // own image's 64 key columns
          tmem_ld32(trow + scol, sv);
          tmem_ld32(trow + scol + 32, sv + 32);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < NT; ++j) mx = fmaxf(mx, sv[j]);
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            sv[j] = __expf(sv[j] - mx);
            sum += sv[j];
          }
          const float inv = __fdividef(1.f, sum);
#pragma unroll
          for (int j = 0; j < 64; ++j) sv[j] = j < NT ? sv[j] * inv : 0.f;
          uint8_t* P = smem + (half ? OFF_H : OFF_P);
          store_row32(P + slot * TILE, row, 0, sv);
          store_row32(P + slot * TILE, row, 32, sv + 32);
          float z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0.f;
          store_row32(P + (slot ^ 1) * TILE, row, 0, z);  // the other image's keys: exact zeros
          store_row32(P + (slot ^ 1) * TILE, row, 32, z);
        }
        signal_ready();
      }
      wait_done();  // ---- O complete
      {
        float v[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(trow + COL_O + 64 * half + 32 * c, v);
          tmem_ld_wait();
          store_row32(Hh, row, c * 32, v);
        }
      }
      signal_ready();
      wait_done();  // ---- X = x + O Wproj^T
      tmem_ld32(trow + xcol, x);
      tmem_ld32(trow + xcol + 32, x + 32);
      tmem_ld_wait();
      add_bias32(x, Bk.proj_b + 64 * half, 1.f);
      add_bias32(x + 32, Bk.proj_b + 64 * half + 32, 1.f);
      tmem_st32(trow + xcol, x);
      tmem_st32(trow + xcol + 32, x + 32);
      tmem_st_wait();
      layernorm_to_tile(x, Bk.n2w + 64 * half, Bk.n2b + 64 * half, 1e-6f, Hh, row, red, rk);
      signal_ready();
      wait_done();  // ---- fc1 accumulators ready: this thread takes hidden columns 128*half..+127
      {
        float v[32];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = 128 * half + 32 * c;
          tmem_ld32(trow + COL_ACC + col, v);
          tmem_ld_wait();
          add_bias32(v, Bk.fc1_b + col, 1.f);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
          store_row32(smem + OFF_Q + (col >> 6) * TILE, row, col & 32, v);
        }
      }
      signal_ready();
      wait_done();  // ---- X = x + h Wfc2^T
      tmem_ld32(trow + xcol, x);
      tmem_ld32(trow + xcol + 32, x + 32);
      tmem_ld_wait();
      add_bias32(x, Bk.fc2_b + 64 * half, 1.f);
      add_bias32(x + 32, Bk.fc2_b + 64 * half + 32, 1.f);
      layernorm_inplace(x, w.snw + 64 * half, w.snb + 64 * half, 1e-6f, red, rk);
      if (l < 2) {
        tmem_st32(trow + xcol, x);
        tmem_st32(trow + xcol + 32, x + 32);
        tmem_st_wait();
        layernorm_to_tile(x, w.blk[l + 1].n1w + 64 * half, w.blk[l + 1].n1b + 64 * half, 1e-6f, Hh, row, red, rk);
      } else {
        layernorm_to_tile(x, w.hnw + 64 * half, w.hnb + 64 * half, 1e-5f, Hh, row, red, rk);
      }
    }
    signal_ready();
    wait_done();  // ---- head accumulators: 32 of the 64 output channels per thread
    {
      float v[32];
      tmem_ld32(trow + COL_ACC + 32 * half, v);
      tmem_ld_wait();
      if (valid) {
        add_bias32(v, w.head_b + 32 * half, 1.f);
        float* dst = yout + ((size_t)img * NT + tok) * 64 + 32 * half;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Linear weight [N][K] fp32 (PyTorch layout = K-major) -> bf16 operand tiles of `rt` rows x 64 k in 128B-swizzled
// smem image order, tile index = (n / rt) * (K / 64) + k / 64; rows below `scaled_rows` are multiplied by QK_SCALE.
__global__ void pack_ste_tiles_kernel(const float* __restrict__ src, int N, int K, int rt, int scaled_rows,
                                      __nv_bfloat16* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const int n = idx / K, k = idx - n * K;
  const int nc = n / rt, r = n - nc * rt, kb = k >> 6, c = k & 63;
  const size_t tile = (size_t)nc * (K >> 6) + kb;
  const size_t off = tile * rt * 64 + (size_t)r * 64 + (((c >> 3) ^ (r & 7)) << 3) + (c & 7);
  dst[off] = __float2bfloat16_rn(src[idx] * (n < scaled_rows ? QK_SCALE : 1.f));
}

}  // namespace

size_t ste_tc_packed_bytes() { return (size_t)3 * TILES_PER_BLOCK * TILE + 2 * (TILE / 2); }

// block l in 0..2, which: 0 qkv (384x128), 1 proj (128x128), 2 fc1 (256x128), 3 fc2 (128x256); l = 3: head (64x128)
void launch_pack_ste_tc(const float* src, int l, int which, void* packed, cudaStream_t st) {
  static const int tile0[4] = {0, 6, 8, 12};
  static const int Ns[4] = {384, 128, 256, 128}, Ks[4] = {128, 128, 128, 256};
  uint8_t* base = reinterpret_cast<uint8_t*>(packed);
  if (l == 3) {
    pack_ste_tiles_kernel<<<ceil_div(64 * 128, 256), 256, 0, st>>>(
        src, 64, 128, 64, 0, reinterpret_cast<__nv_bfloat16*>(base + (size_t)3 * TILES_PER_BLOCK * TILE));
    return;
  }
  pack_ste_tiles_kernel<<<ceil_div(Ns[which] * Ks[which], 256), 256, 0, st>>>(
      src, Ns[which], Ks[which], 128, which == 0 ? 128 : 0,
      reinterpret_cast<__nv_bfloat16*>(base + ((size_t)l * TILES_PER_BLOCK + tile0[which]) * TILE));
}

void launch_ste_tc(const float* x, float* y, const SteWeights& w, const void* packed, int B, cudaStream_t st) {
  launch_pdl(ste_tc_kernel, dim3((B + 1) / 2), dim3(TC_THREADS), SMEM_BYTES, st, x, y, w,
             reinterpret_cast<const uint8_t*>(packed), B);
}

}  // namespace dirb200
