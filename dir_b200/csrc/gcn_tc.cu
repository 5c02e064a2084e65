// SemGCN layer GEMMs on the tensor cores (bf16 configuration): the 84 uniform products per layer
//   H_k[b, hand, j, :] = X'[b, hand, j, :] . W[hand][k][j]        (k in {0,1}, j in 0..20; SemGCN/p_graph_conv.py:39-60)
// of gcn_gemm_kernel (joint.cu) as tcgen05.mma kind::tf32 (fp32 operands read as tf32, fp32 accumulation).
// CTA = (j, hand, 128 images). The A operand X' = relu(bn(H0[j] + sum_j' A1[j][j'] H1[j'])) of the previous layer
// (or the embedded features for layer 0) is aggregated once, one warp per image row, straight into the K-major
// 128B-swizzled smem tile; the two 64 KB weight tiles arrive by cp.async.bulk from their finalize-time smem image,
// issued before griddepcontrol.wait; 2 x 16 MMAs (M=128, N=128, K=8) accumulate in TMEM; the epilogue transposes each
// warp's 32x64 patch through padded smem so that a store instruction writes two 256-byte row segments.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dirb200 {

namespace {

using namespace tc;

constexpr int NJ = 21;
constexpr int G_ATILE = 128 * 128;  // bytes of one k-tile ([128 rows][32 fp32])
constexpr int G_PITCH = 68;
constexpr int G_OFF_A = 0, G_OFF_W = 4 * G_ATILE;  // A: 4 k-tiles; W: 2 (k = 0, 1) x 4 k-tiles
constexpr int G_OFF_BAR = G_OFF_W + 8 * G_ATILE;
constexpr int G_SMEM = 1024 + G_OFF_BAR + 64;
constexpr int G_THREADS = 320;

struct GBars {
  uint64_t wfull, a_ready, done;
  uint32_t tmem_ptr;
};

// CTA = (joint j, hand, 128 images): the operand is aggregated ONCE and multiplied by both W[0][j] and W[1][j]
// (two 128-column accumulators). Warps 0-7: staging (a warp per image row, lane = 16-byte chunk) and epilogue (warps 0-3 drain H_0,
// warps 4-7 H_1; TMEM lane quarter = warp % 4); warp 8: weight tiles; warp 9: TMEM + MMA issue.
__global__ void __launch_bounds__(G_THREADS, 1) gcn_gemm_tc_kernel(GcnGemmArgs a, const uint8_t* wpk0, const uint8_t* wpk1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GBars* bars = reinterpret_cast<GBars*>(smem + G_OFF_BAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, hand = blockIdx.y, b0 = blockIdx.z * 128;
  if (threadIdx.x == 0) {
    mbar_init(&bars->wfull, 1);
    mbar_init(&bars->a_ready, 256);
    mbar_init(&bars->done, 1);
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

  if (warp == 8) {  // ---- both weight tiles of joint j (finalize-time data: issued before griddepcontrol.wait)
    if (lane == 0) {
      const uint8_t* base = hand ? wpk1 : wpk0;
      mbar_expect_tx(&bars->wfull, 8 * G_ATILE);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(smem + G_OFF_W + k * 4 * G_ATILE)),
                     "l"(base + (size_t)(k * NJ + j) * (4 * G_ATILE)), "r"(4 * G_ATILE), "r"(s32(&bars->wfull))
                     : "memory");
    }
  } else if (warp == 9) {  // ---- MMA issuer
    if (lane == 0) {
      const uint32_t sb = s32(smem);
      mbar_wait(&bars->wfull, 0);
      mbar_wait(&bars->a_ready, 0);
      fence_after();
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          const uint64_t da = desc128(sb + G_OFF_A + kt * G_ATILE);
          const uint64_t db = desc128(sb + G_OFF_W + (k * 4 + kt) * G_ATILE);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(tmem + 128 * k, da + 2 * ks, db + 2 * ks, idesc(128, 2u), (kt | ks) ? 1u : 0u);
        }
      umma_commit(&bars->done);
    }
  } else {  // ---- operand staging: thread = (image row, channel half), then the epilogue
    pdl_wait();
    // warp w stages rows w, w+8, ...: lane = 16-byte channel chunk, so every load instruction reads one contiguous
    // 512-byte row (a thread per row touched 32 lines per instruction and made the gather LSU-bound)
    const int c4 = lane;
    const uint32_t koff = (uint32_t)(c4 >> 3) * G_ATILE;
    if (a.x) {
#pragma unroll 4
      for (int r = warp; r < 128; r += 8) {
        const int b = b0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < a.B) v = __ldg(reinterpret_cast<const float4*>(a.x + ((size_t)(b * 2 + hand) * NJ + j) * 128) + c4);
        *reinterpret_cast<float4*>(smem + G_OFF_A + koff + r * 128 + (((c4 & 7) ^ (r & 7)) << 4)) = v;
      }
    } else {
      // X'[j] = relu(bn(H0[j] + sum_j' A1[j][j'] H1[j'])): ascending j', like the dense A1 @ h1 of the reference.
      // A joint has at most 5 neighbours (the wrist); unused slots carry weight 0 on the joint itself (finite data).
      const size_t plane = (size_t)a.B * 2 * NJ * 128;
      const float* A1 = a.agg.A1[hand] + j * NJ;
      int nbr[5] = {j, j, j, j, j};
      float wgt[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      {
        int nn = 0;
        for (int jj = 0; jj < NJ; ++jj) {
          const float aw = __ldg(A1 + jj);
          if (aw != 0.f && nn < 5) {
#pragma unroll
            for (int e = 0; e < 5; ++e)
              if (e == nn) {
                nbr[e] = jj;
                wgt[e] = aw;
              }
            ++nn;
          }
        }
      }
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.agg.scale[hand]) + c4);
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(a.agg.shift[hand]) + c4);
      const bool wide = wgt[2] != 0.f;  // only the wrist row has more than 2 neighbours
#pragma unroll 4
      for (int r = warp; r < 128; r += 8) {
        const int b = b0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < a.B) {
          const float* base = a.hin + ((size_t)(b * 2 + hand) * NJ) * 128;
          v = __ldg(reinterpret_cast<const float4*>(base + (size_t)j * 128) + c4);
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(base + plane + (size_t)nbr[0] * 128) + c4);
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(base + plane + (size_t)nbr[1] * 128) + c4);
          v.x = fmaf(wgt[0], h0.x, v.x); v.y = fmaf(wgt[0], h0.y, v.y); v.z = fmaf(wgt[0], h0.z, v.z); v.w = fmaf(wgt[0], h0.w, v.w);
          v.x = fmaf(wgt[1], h1.x, v.x); v.y = fmaf(wgt[1], h1.y, v.y); v.z = fmaf(wgt[1], h1.z, v.z); v.w = fmaf(wgt[1], h1.w, v.w);
          if (wide) {
#pragma unroll
            for (int e = 2; e < 5; ++e) {
              const float4 h = __ldg(reinterpret_cast<const float4*>(base + plane + (size_t)nbr[e] * 128) + c4);
              v.x = fmaf(wgt[e], h.x, v.x); v.y = fmaf(wgt[e], h.y, v.y); v.z = fmaf(wgt[e], h.z, v.z); v.w = fmaf(wgt[e], h.w, v.w);
            }
          }
          v = make_float4(fmaxf(fmaf(v.x, s4.x, h4.x), 0.f), fmaxf(fmaf(v.y, s4.y, h4.y), 0.f),
                          fmaxf(fmaf(v.z, s4.z, h4.z), 0.f), fmaxf(fmaf(v.w, s4.w, h4.w), 0.f));
        }
        *reinterpret_cast<float4*>(smem + G_OFF_A + koff + r * 128 + (((c4 & 7) ^ (r & 7)) << 4)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(&bars->a_ready);
    mbar_wait(&bars->done, 0);  // both products complete: A and W are dead, the transpose patches alias them
    fence_after();
    const int k = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128 * k;
    float* patch = reinterpret_cast<float*>(smem) + warp * (32 * G_PITCH);
    float* out = a.hout + (size_t)k * a.B * 2 * NJ * 128;
    const int prow_l = lane >> 4, pcol = (lane & 15) * 4;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[64];
      tmem_ld32(trow + 64 * c, v);
      tmem_ld32(trow + 64 * c + 32, v + 32);
      tmem_ld_wait();
      float4* mine = reinterpret_cast<float4*>(patch + lane * G_PITCH);
#pragma unroll
      for (int i = 0; i < 16; ++i) mine[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      __syncwarp();
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int rr = 2 * it + prow_l, bb = b0 + (warp & 3) * 32 + rr;
        if (bb < a.B)
          *reinterpret_cast<float4*>(out + ((size_t)(bb * 2 + hand) * NJ + j) * 128 + 64 * c + pcol) =
              *reinterpret_cast<const float4*>(patch + rr * G_PITCH + pcol);
      }
      __syncwarp();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// gconv.W (2, 21, 128 in, 128 out) fp32 -> per (k, j): 4 k-tiles x [128 n][32 c] fp32, 128B-swizzled smem image
__global__ void pack_gcn_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ wpk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * NJ * 128 * 128) return;
  const int n = idx & 127, c = (idx >> 7) & 127, g = idx >> 14;  // source element w[g][c][n]
  const int kt = c >> 5, cc = c & 31;
  wpk[(size_t)g * (4 * 128 * 32) + kt * (128 * 32) + n * 32 + ((((cc >> 2) ^ (n & 7)) << 2) | (cc & 3))] = w[idx];
}

}  // namespace

size_t gcn_tc_packed_bytes() { return (size_t)2 * NJ * 4 * G_ATILE; }

void launch_pack_gcn_weight_tc(const float* w, void* wpk, cudaStream_t st) {
  pack_gcn_weight_tc_kernel<<<ceil_div(2 * NJ * 128 * 128, 256), 256, 0, st>>>(w, reinterpret_cast<float*>(wpk));
}

void launch_gcn_gemm_tc(const GcnGemmArgs& a, const void* wpk_left, const void* wpk_right, cudaStream_t st) {
  launch_pdl(gcn_gemm_tc_kernel, dim3(NJ, 2, ceil_div(a.B, 128)), dim3(G_THREADS), G_SMEM, st, a,
             reinterpret_cast<const uint8_t*>(wpk_left), reinterpret_cast<const uint8_t*>(wpk_right));
}

}  // namespace dirb200
