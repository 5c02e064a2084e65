// Exact factored form of  bone_proj -> Conv2d(2560->256, 3x3) -> BN -> ReLU   (models/dir.py:57-60,120-122,146-174).
//
// bone_proj's 1280-channel map of one hand is, per bone g, a rank-2 field:
//     in[b, g*64+c, p] = mask_g(p) * ( wa_g(p) * fa[c] + wb_g(p) * fb[c] ),   fa/fb = 64-d features of the bone's joints
// so the 3x3 convolution over those 2x1280 channels factors EXACTLY (only the fp32 summation order changes):
//     out[b, n, q] = bias[n] + sum_{hand,g,tap} mask(p) * ( wa(p) * Pa[b,hand,g,tap,n] + wb(p) * Pb[...] ),  p = q + tap - 1
//     P{a,b}[b,hand,g,tap,n] = sum_c W[n, hand*1280+g*64+c, tap] * f{a,b}[c]
// Step 1 (bone_coef_kernel) is a small grouped GEMM (B*2 x 64 x 2304 per bone); step 2 (bone_fusion_kernel) walks
// only the non-zero (pixel, bone) pairs (~3-7 % of the map). 12.2 GFLOP/img of dense conv at 32x32 become
// ~0.06 GFLOP/img, the (B,2560,S,S) tensor is never materialised, and the arithmetic stays fp32.
// The dense tcgen05 path (bone_raster + conv_tc) remains available (DIRB200_DENSE_FUSION=1) and is tested against this.
#include "../../include/dirb200.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dirb200 {

namespace {

constexpr int NJ = 21;

struct BoneGeom {
  float ax, ay, bx, by, dx, dy;
};

// Conservative reject: a pixel centre farther than `distance` (+ slack for rounding) from the bone's bounding box
// cannot pass the capsule test, which measures the distance to the segment. NaN geometry never rejects here.
__device__ __forceinline__ bool bone_bbox_reject(const BoneGeom& g, float px, float py, float distance) {
  const float m = distance + 0.01f;
  return px < fminf(g.ax, g.bx) - m || px > fmaxf(g.ax, g.bx) + m || py < fminf(g.ay, g.by) - m ||
         py > fmaxf(g.ay, g.by) + m;
}

// identical arithmetic to joint.cu::bone_weights (reference op order, no FMA contraction in the mask)
__device__ __forceinline__ bool bone_weights(const BoneGeom& g, float px, float py, float distance, float& wa,
                                             float& wb) {
  float s = __fadd_rn(__fmul_rn(__fsub_rn(g.ax, px), g.dx), __fmul_rn(__fsub_rn(g.ay, py), g.dy));
  float t = __fadd_rn(__fmul_rn(__fsub_rn(px, g.bx), g.dx), __fmul_rn(__fsub_rn(py, g.by), g.dy));
  float h = fmaxf(fmaxf(s, t), 0.f);
  float c = __fsub_rn(__fmul_rn(__fsub_rn(px, g.ax), g.dy), __fmul_rn(__fsub_rn(py, g.ay), g.dx));
  float dist = hypotf(h, c);
  if (!(dist < distance)) return false;
  float ex = __fadd_rn(__fsub_rn(px, g.ax), 1e-6f), ey = __fadd_rn(__fsub_rn(py, g.ay), 1e-6f);
  float da = sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
  ex = __fadd_rn(__fsub_rn(px, g.bx), 1e-6f);
  ey = __fadd_rn(__fsub_rn(py, g.by), 1e-6f);
  float db = sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
  float sum = __fadd_rn(da, db);
  wa = 1.f - da / sum;
  wb = 1.f - db / sum;
  return true;
}

__device__ __forceinline__ BoneGeom make_bone(const float* uv, int bone, int S) {
  int pa = (bone % 4 == 0) ? 0 : bone, ch = bone + 1;
  BoneGeom g;
  g.ax = (uv[pa * 2] + 1.f) / 2.f * S;
  g.ay = (uv[pa * 2 + 1] + 1.f) / 2.f * S;
  g.bx = (uv[ch * 2] + 1.f) / 2.f * S;
  g.by = (uv[ch * 2 + 1] + 1.f) / 2.f * S;
  float ex = g.bx - g.ax, ey = g.by - g.ay;
  float len = hypotf(ex, ey);
  g.dx = ex / len;
  g.dy = ey / len;
  return g;
}

// fusion.0.weight [256][2560][3][3] -> Wp[hb][c][tap][n]  (hb = hand*20 + bone)
__global__ void pack_fusion_weight_kernel(const float* __restrict__ w, float* __restrict__ wp) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)40 * 64 * 9 * 256) return;
  int n = (int)(idx % 256);
  int64_t t = idx / 256;
  int tap = (int)(t % 9);
  t /= 9;
  int c = (int)(t % 64);
  int hb = (int)(t / 64);
  wp[idx] = w[((int64_t)n * 2560 + hb * 64 + c) * 9 + tap];
}

// P[b][hb][role][tap][n] = sum_c Wp[hb][c][tap][n] * jf[b][hand][joint(hb,role)][c]
// grid (40, 9, ceil(2B/64)), 256 threads (n); 64 rows (b,role) per CTA.
__global__ void __launch_bounds__(256) bone_coef_kernel(const float* __restrict__ jf, const float* __restrict__ wp,
                                                        float* __restrict__ P, int B) {
  pdl_wait();
  extern __shared__ __align__(128) float dyn_smem[];  // weight stream: 2 x 32 x 128 floats + 2 mbarriers
  WStream ws;
  wstream_init(ws, dyn_smem, reinterpret_cast<uint64_t*>(dyn_smem + 2 * 32 * 128));
  __shared__ __align__(16) float xs[64][68];
  const int hb = blockIdx.x, tap = blockIdx.y, r0 = blockIdx.z * 64;
  const int hand = hb / 20, bone = hb % 20;
  const int pa = (bone % 4 == 0) ? 0 : bone, ch = bone + 1;
  const int rows = min(64, 2 * B - r0);
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    int r = i >> 6, c = i & 63;
    float v = 0.f;
    if (r < rows) {
      int ri = r0 + r, b = ri >> 1, role = ri & 1;
      v = jf[((int64_t)(b * 2 + hand) * NJ + (role ? ch : pa)) * 64 + c];
    }
    xs[r][c] = v;
  }
  __syncthreads();
  // (rows x 64)·(64 x 256) as two 128-column blocks on the weight-streaming CTA GEMM (8 row groups x 32 col groups)
  for (int cb = 0; cb < 256; cb += 128) {
    const float* Bt = wp + ((int64_t)hb * 64 * 9 + tap) * 256 + cb;  // row c at Bt + c*9*256
    cta_gemm<8>(&xs[0][0], 68, rows, 64, Bt, 9 * 256, 128, ws, [&](int, int r, int c0, float (&v)[4]) {
      int ri = r0 + r, b = ri >> 1, role = ri & 1;
      *reinterpret_cast<float4*>(P + ((((int64_t)b * 40 + hb) * 2 + role) * 9 + tap) * 256 + cb + c0) =
          make_float4(v[0], v[1], v[2], v[3]);
    });
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bone_coef on the tensor cores (bf16 configuration): the same grouped GEMM as tcgen05.mma kind::tf32 (fp32 operands
// read as tf32, fp32 accumulation), i.e. 8x the operand precision of the bf16 dense conv it replaces.
// CTA = (bone hb, 128 rows (b, role)): the 128x64 feature tile is staged once (K-major, 128B swizzle, 2 k-tiles of
// 32 floats), then the 9 taps stream through a 2-deep ring of 64 KB weight tiles ([256 n][64 c], packed at finalize
// in smem-image order: one cp.async.bulk each) into two alternating 256-column TMEM accumulators while 4 warps drain
// the previous tap to P. Roles: warps 0-3 staging + epilogue, warp 4 weight producer, warp 5 TMEM + MMA issue.
constexpr int CT_WTILE = 2 * 256 * 128;  // bytes of one (hb, tap) weight tile
constexpr int CT_ATILE = 128 * 128;      // bytes of one k-tile of the feature operand
constexpr int CT_PITCH = 68;             // floats per row of an epilogue transpose patch (64 + 4: conflict-free)
constexpr int CT_OFF_A = 0, CT_OFF_W = 2 * CT_ATILE, CT_OFF_STAGE = CT_OFF_W + 2 * CT_WTILE;
constexpr int CT_OFF_BAR = CT_OFF_STAGE + 4 * 32 * CT_PITCH * 4;
constexpr int CT_SMEM = 1024 + CT_OFF_BAR + 128;

struct CoefBars {
  uint64_t full[2], empty[2], acc_full[2], acc_empty[2], a_ready;
  uint32_t tmem_ptr;
};

template <typename PT>  // PT = float, or __nv_bfloat16 when the consumer is bone_fusion_tc_kernel (which rounds P to bf16 anyway)
__global__ void __launch_bounds__(192, 1)
bone_coef_tc_kernel(const float* __restrict__ jf, const uint8_t* __restrict__ wpk, PT* __restrict__ P, int B) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  CoefBars* bars = reinterpret_cast<CoefBars*>(smem + CT_OFF_BAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hb = blockIdx.x, r0 = blockIdx.y * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], 128);
    }
    mbar_init(&bars->a_ready, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_ptr;

  if (warp == 4) {  // ---- weight producer (finalize-time data: starts before the previous kernel has finished)
    if (lane == 0) {
      for (int tap = 0; tap < 9; ++tap) {
        const int slot = tap & 1;
        if (tap >= 2) mbar_wait(&bars->empty[slot], ((tap >> 1) - 1) & 1);
        mbar_expect_tx(&bars->full[slot], CT_WTILE);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(smem + CT_OFF_W + slot * CT_WTILE)),
                     "l"(wpk + ((size_t)hb * 9 + tap) * CT_WTILE), "r"(CT_WTILE), "r"(s32(&bars->full[slot]))
                     : "memory");
      }
    }
  } else if (warp == 5) {  // ---- MMA issuer
    if (lane == 0) {
      const uint32_t sb = s32(smem);
      mbar_wait(&bars->a_ready, 0);
      fence_after();
      for (int tap = 0; tap < 9; ++tap) {
        const int slot = tap & 1;
        mbar_wait(&bars->full[slot], (tap >> 1) & 1);
        mbar_wait(&bars->acc_empty[slot], ((tap >> 1) & 1) ^ 1);
        fence_after();
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          const uint64_t da = desc128(sb + CT_OFF_A + kt * CT_ATILE);
          const uint64_t db = desc128(sb + CT_OFF_W + slot * CT_WTILE + kt * (CT_WTILE / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(tmem + slot * 256, da + 2 * k, db + 2 * k, idesc(256, 2u), (kt | k) ? 1u : 0u);
        }
        umma_commit(&bars->empty[slot]);
        umma_commit(&bars->acc_full[slot]);
      }
    }
  } else {  // ---- stage the feature rows, then drain the accumulators: thread = row (b, role)
    pdl_wait();
    const int r = threadIdx.x, ri = r0 + r;
    const int hand = hb / 20, bone = hb % 20;
    const int pa = (bone % 4 == 0) ? 0 : bone, ch = bone + 1;
    const bool valid = ri < 2 * B;
    const int b = ri >> 1, role = ri & 1;
    {
      const float4* src = reinterpret_cast<const float4*>(jf + ((int64_t)(b * 2 + hand) * NJ + (role ? ch : pa)) * 64);
      uint8_t* arow = smem + CT_OFF_A + r * 128;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 v = valid ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(arow + (c4 >> 3) * CT_ATILE + (((c4 & 7) ^ (r & 7)) << 4)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(&bars->a_ready);
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // TMEM hands a thread one row; a row-per-lane global store would touch 32 lines per instruction (LSU-bound:
    // measured 82 us). Each warp transposes 32 rows x 64 columns through its own padded smem patch instead, so one
    // store instruction writes two 256-byte row segments.
    float* patch = reinterpret_cast<float*>(smem + CT_OFF_STAGE) + warp * (32 * CT_PITCH);
    const int prow_l = lane >> 4, pcol = (lane & 15) * 4;
    for (int tap = 0; tap < 9; ++tap) {
      const int slot = tap & 1;
      mbar_wait(&bars->acc_full[slot], (tap >> 1) & 1);
      fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[64];
        tmem_ld32(trow + slot * 256 + 64 * c, v);
        tmem_ld32(trow + slot * 256 + 64 * c + 32, v + 32);
        tmem_ld_wait();
        if constexpr (sizeof(PT) == 4) {
          float4* mine = reinterpret_cast<float4*>(patch + lane * CT_PITCH);
#pragma unroll
          for (int i = 0; i < 16; ++i) mine[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
#pragma unroll 4
          for (int it = 0; it < 16; ++it) {
            const int rr = 2 * it + prow_l, rj = r0 + warp * 32 + rr;
            if (rj < 2 * B) {
              const float4 o = *reinterpret_cast<const float4*>(patch + rr * CT_PITCH + pcol);
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(P) +
                                         ((((int64_t)(rj >> 1) * 40 + hb) * 2 + (rj & 1)) * 9 + tap) * 256 + 64 * c + pcol) = o;
            }
          }
        } else {  // bf16: a row of the patch is 64 values = 128 bytes (+16 pad); a store instruction writes 4 rows
          uint4* mine = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(patch) + lane * 144);
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * g8], v[8 * g8 + 1]), h1 = __floats2bfloat162_rn(v[8 * g8 + 2], v[8 * g8 + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * g8 + 4], v[8 * g8 + 5]), h3 = __floats2bfloat162_rn(v[8 * g8 + 6], v[8 * g8 + 7]);
            mine[g8] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                  *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = 4 * it + (lane >> 3), rj = r0 + warp * 32 + rr, ch = lane & 7;
            if (rj < 2 * B) {
              const uint4 o = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(patch) + rr * 144 + ch * 16);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P) +
                                        ((((int64_t)(rj >> 1) * 40 + hb) * 2 + (rj & 1)) * 9 + tap) * 256 + 64 * c + ch * 8) = o;
            }
          }
        }
        __syncwarp();
      }
      fence_before();
      mbar_arrive(&bars->acc_empty[slot]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// fusion.0.weight [256][2560][3][3] -> per (hb, tap): 2 k-tiles x [256 n][32 c] fp32 in 128B-swizzled smem-image order
__global__ void pack_fusion_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ wpk) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)40 * 9 * 256 * 64) return;
  const int c = (int)(idx & 63);
  const int n = (int)((idx >> 6) & 255);
  const int t = (int)(idx >> 14);
  const int tap = t % 9, hb = t / 9;
  const int kt = c >> 5, cc = c & 31;
  const int64_t off = (((int64_t)hb * 9 + tap) * 2 + kt) * 256 * 32 + n * 32 + ((((cc >> 2) ^ (n & 7)) << 2) | (cc & 3));
  wpk[off] = w[((int64_t)n * 2560 + hb * 64 + c) * 9 + tap];
}

struct Entry {
  int key;  // xs | (ky*40 + hb) << 8
  float wa, wb;
};

// One CTA per (image, output row). 256 threads = 256 output channels.
template <typename T, int S>
__global__ void __launch_bounds__(256) bone_fusion_kernel(const float* __restrict__ rec, int rec_stride,
                                                          const float* __restrict__ P, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, T* __restrict__ out,
                                                          float distance) {
  pdl_wait();
  constexpr int LOG2S = S == 32 ? 5 : 4;
  static_assert(S == 16 || S == 32, "feature map sizes of projecter_4 / projecter_3 (models/dir.py:395,401)");
  extern __shared__ __align__(16) uint8_t smraw[];
  float* s_out = reinterpret_cast<float*>(smraw);                  // [S][256]
  Entry* list = reinterpret_cast<Entry*>(s_out + S * 256);         // [3*S*40]
  int* counts = reinterpret_cast<int*>(list + 3 * S * 40);         // [nchunks + 1]
  int* gstart = counts + (3 * S * 40 + 31) / 32 + 1;                // [<= 121] run starts of (ky, bone) groups
  __shared__ BoneGeom geo[40];
  __shared__ float uv[84];
  const int b = blockIdx.x, y = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 84) uv[tid] = rec[(int64_t)b * rec_stride + DIRB200_OFF_UV_L + tid];
  for (int i = tid; i < S * 256; i += 256) s_out[i] = 0.f;
  __syncthreads();
  if (tid < 40) geo[tid] = make_bone(uv + (tid / 20) * 42, tid % 20, S);
  __syncthreads();
  // ---- deterministic compaction of the non-zero (source pixel, bone) pairs of rows y-1, y, y+1
  // test index t = (ky*40 + hb)*S + xs (bone-major inside a source row, so one bone's pixels are consecutive and
  // share their 6 coefficient vectors), processed in chunks of 32 (one warp-ballot each)
  const int ntests = 3 * S * 40, nchunks = (ntests + 31) / 32;
  for (int ck = warp; ck < nchunks; ck += 8) {
    int t = ck * 32 + lane;
    bool hit = false;
    if (t < ntests) {
      int xs = t & (S - 1), g = t >> LOG2S, ky = g / 40, hb = g - ky * 40;
      int ys = y + ky - 1;
      float wa, wb;
      if (ys >= 0 && ys < S && !bone_bbox_reject(geo[hb], xs + 0.5f, ys + 0.5f, distance))
        hit = bone_weights(geo[hb], xs + 0.5f, ys + 0.5f, distance, wa, wb);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) counts[ck] = __popc(m);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of counts[0..nchunks) -> counts (in place), total in counts[nchunks]
    int carry = 0;
    for (int base = 0; base < nchunks; base += 32) {
      int i = base + lane;
      int v = i < nchunks ? counts[i] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (i < nchunks) counts[i] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) counts[nchunks] = carry;
  }
  __syncthreads();
  for (int ck = warp; ck < nchunks; ck += 8) {
    int t = ck * 32 + lane;
    bool hit = false;
    float wa = 0.f, wb = 0.f;
    int key = 0;
    if (t < ntests) {
      int xs = t & (S - 1), g = t >> LOG2S, ky = g / 40, hb = g - ky * 40;
      int ys = y + ky - 1;
      if (ys >= 0 && ys < S && !bone_bbox_reject(geo[hb], xs + 0.5f, ys + 0.5f, distance))
        hit = bone_weights(geo[hb], xs + 0.5f, ys + 0.5f, distance, wa, wb);
      key = xs | (g << 8);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      int pos = counts[ck] + __popc(m & ((1u << lane) - 1));
      list[pos].key = key;
      list[pos].wa = wa;
      list[pos].wb = wb;
    }
  }
  __syncthreads();
  // ---- accumulate: thread n owns column n of s_out (no conflicts, fixed order => bit-reproducible)
  const int n = tid, nent = counts[nchunks];
  __syncthreads();  // everyone has read nent before counts[nchunks] is reused for the group count
  const float* Pb = P + (int64_t)b * 40 * 2 * 9 * 256 + n;
  // Entries arrive grouped by (ky, bone) with ascending xs. Group table (start index of every (ky, bone) run):
  if (warp == 0) {
    int count = 0;
    for (int base = 0; base < nent; base += 32) {
      const int e = base + lane;
      const bool flag = e < nent && (e == 0 || (list[e].key >> 8) != (list[e - 1].key >> 8));
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      if (flag) gstart[count + __popc(m & ((1u << lane) - 1))] = e;
      count += __popc(m);
    }
    if (lane == 0) {
      gstart[count] = nent;
      counts[nchunks] = count;  // number of groups (nent is gstart[count])
    }
  }
  __syncthreads();
  const int ngroups = counts[nchunks];
  // Four groups at a time: their 24 coefficient loads (L2) are issued together, then a 3-wide register window
  // (outputs xs-1, xs, xs+1) slides along each bone's pixels so shared memory is touched about once per pixel.
  for (int g0 = 0; g0 < ngroups; g0 += 4) {
    float pa[4][3], pb[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gi = min(g0 + u, ngroups - 1);
      const int g = list[gstart[gi]].key >> 8;
      const int ky = g / 40, hb = g - ky * 40;
      // source pixel (ys, xs) feeds output (y, xs - kx + 1) through tap (ky, kx), ky = ys - y + 1
      const float* qa = Pb + ((int64_t)(hb * 2 + 0) * 9 + ky * 3) * 256;
      const float* qb = Pb + ((int64_t)(hb * 2 + 1) * 9 + ky * 3) * 256;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        pa[u][kx] = __ldg(qa + kx * 256);
        pb[u][kx] = __ldg(qb + kx * 256);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (g0 + u >= ngroups) break;
      const int e0 = gstart[g0 + u], e1 = gstart[g0 + u + 1];
      int last_xs = -100;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
      for (int e = e0; e < e1; ++e) {
        const Entry en = list[e];
        const int xs = en.key & 0xff;
        if (xs == last_xs + 1) {  // slide by one pixel: retire the left output
          if (last_xs - 1 >= 0) s_out[(last_xs - 1) * 256 + n] += w0;
          w0 = w1;
          w1 = w2;
          w2 = 0.f;
        } else if (last_xs >= 0) {  // gap: flush the whole window
          if (last_xs - 1 >= 0) s_out[(last_xs - 1) * 256 + n] += w0;
          s_out[last_xs * 256 + n] += w1;
          if (last_xs + 1 < S) s_out[(last_xs + 1) * 256 + n] += w2;
          w0 = w1 = w2 = 0.f;
        }
        last_xs = xs;
        w0 += fmaf(en.wa, pa[u][2], en.wb * pb[u][2]);  // kx = 2 -> x = xs - 1
        w1 += fmaf(en.wa, pa[u][1], en.wb * pb[u][1]);  // kx = 1 -> x = xs
        w2 += fmaf(en.wa, pa[u][0], en.wb * pb[u][0]);  // kx = 0 -> x = xs + 1
      }
      if (last_xs >= 0) {
        if (last_xs - 1 >= 0) s_out[(last_xs - 1) * 256 + n] += w0;
        s_out[last_xs * 256 + n] += w1;
        if (last_xs + 1 < S) s_out[(last_xs + 1) * 256 + n] += w2;
      }
    }
  }
  __syncthreads();
  // ---- BN + ReLU epilogue, NHWC store (256 consecutive channels per pixel: fully coalesced)
  const float sc = scale[n], sh = shift[n];
  T* orow = out + ((int64_t)b * S + y) * S * 256 + n;
  for (int x = 0; x < S; ++x) ActIO<T>::st(orow + x * 256, fmaxf(fmaf(s_out[x * 256 + n], sc, sh), 0.f));
}

// ---------------------------------------------------------------------------------------------------------------
// bone_fusion on the tensor cores (bf16 configuration). The sparse accumulate
//     out[b, y, x, n] = sum_{bone, tap} mask(p) * ( wa(p) Pa[b,bone,tap,n] + wb(p) Pb[b,bone,tap,n] ),  p = (y,x) + tap - 1
// is a GEMM per block of output rows:  out[m, n] = sum_k A[m, k] B[k, n],  m = (row, x) (128 output pixels per CTA),
// k = (bone, endpoint role, tap):  A[m, k] = w_role(source pixel of m under tap) if that pixel lies on the bone else 0,
// B[k, n] = P[b, bone, role, tap, n]. Every A element has exactly one possible source pixel, so A is written (never
// accumulated) by the thread that evaluates that (pixel, bone) capsule test — no compaction, no lists, no atomics,
// and the per-entry work is done ONCE instead of once per output channel (the CUDA-core kernel repeats it in 256
// threads). K walks the image's ACTIVE bones (bounding box meets the row band) three at a time: 54 columns of a
// 64-column bf16 chunk, double-buffered: 8 warps build chunk c+1 (capsule tests -> A; P rows -> B, bf16) while one
// thread issues the 4 tcgen05.mma (M=128, N=256, K=16) of chunk c. Epilogue: TMEM -> bn/ReLU -> bf16 -> per-warp
// padded patch -> coalesced stores (the 128 x 256 tile is one contiguous 64 KB block of the NHWC map).
constexpr int FT_A = 128 * 128;  // bytes of one A chunk  [128 m][64 k] bf16
constexpr int FT_B = 256 * 128;  // bytes of one B chunk  [256 n][64 k] bf16
constexpr int FT_OFF_A = 0, FT_OFF_B = 2 * FT_A, FT_OFF_BAR = FT_OFF_B + 2 * FT_B;
constexpr int FT_SMEM = 1024 + FT_OFF_BAR + 256;
constexpr int FT_THREADS = 288;
constexpr int FT_PATCH = 144;  // bytes per patch row: 64 bf16 + 16 pad

struct FtBars {
  uint64_t ready[2], free_[2], done;
  uint32_t tmem_ptr;
  int nact;
  int act[40];
};

template <int S, typename PT>
__global__ void __launch_bounds__(FT_THREADS, 2)
bone_fusion_tc_kernel(const float* __restrict__ rec, int rec_stride, const PT* __restrict__ P,
                      const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                      float distance) {
  using namespace tc;
  constexpr int R = 128 / S;            // output rows per CTA
  constexpr int NPX = (R + 2) * S;      // source pixels a CTA looks at
  constexpr int LOG2S = S == 32 ? 5 : 4;
  static_assert(S == 16 || S == 32, "feature map sizes of projecter_4 / projecter_3 (models/dir.py:395,401)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  FtBars* bars = reinterpret_cast<FtBars*>(smem + FT_OFF_BAR);
  __shared__ BoneGeom geo[40];
  __shared__ float uv[84];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x, y0 = blockIdx.y * R;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->ready[i], 256);
      mbar_init(&bars->free_[i], 1);
    }
    mbar_init(&bars->done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&bars->tmem_ptr)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero both B chunks once: columns 54..63 are never written and must multiply A's zeros with finite values
  for (int i = tid; i < 2 * FT_B / 16; i += FT_THREADS) reinterpret_cast<uint4*>(smem + FT_OFF_B)[i] = make_uint4(0u, 0u, 0u, 0u);
  pdl_wait();
  if (tid < 84) uv[tid] = rec[(int64_t)b * rec_stride + DIRB200_OFF_UV_L + tid];
  __syncthreads();
  if (tid < 40) geo[tid] = make_bone(uv + (tid / 20) * 42, tid % 20, S);
  __syncthreads();
  if (warp == 0) {  // active bones in ascending order: bounding box (+ distance) meets source rows y0-1 .. y0+R
    int n = 0;
    for (int base = 0; base < 40; base += 32) {
      const int hb = base + lane;
      bool on = false;
      if (hb < 40) {
        const BoneGeom g = geo[hb];
        const float m = distance + 0.01f;
        const float lo = fminf(g.ay, g.by) - m, hi = fmaxf(g.ay, g.by) + m;
        // pixel centres of the band: y0 - 0.5 .. y0 + R + 0.5; NaN geometry (collapsed bone) compares false -> skipped,
        // like the capsule test that can never pass for it
        on = hi >= (float)y0 - 0.5f && lo <= (float)(y0 + R) + 0.5f;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, on);
      if (on) bars->act[n + __popc(mask & ((1u << lane) - 1))] = hb;
      n += __popc(mask);
    }
    if (lane == 0) bars->nact = n;
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_ptr;
  const int nact = bars->nact, nchunks = (nact + 2) / 3;

  if (warp == 8) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t sb = s32(smem);
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&bars->ready[buf], (c >> 1) & 1);
        fence_after();
        const uint64_t da = desc128(sb + FT_OFF_A + buf * FT_A), db = desc128(sb + FT_OFF_B + buf * FT_B);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem, da + 2 * k, db + 2 * k, idesc(256), (c | k) ? 1u : 0u);
        umma_commit(&bars->free_[buf]);
      }
      umma_commit(&bars->done);
    }
  } else {
    // ===================================================== chunk builders (8 warps), then the epilogue
    const PT* Pimg = P + (int64_t)b * 40 * 2 * 9 * 256;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      uint8_t* At = smem + FT_OFF_A + buf * FT_A;
      uint8_t* Bt = smem + FT_OFF_B + buf * FT_B;
      if (c >= 2) mbar_wait(&bars->free_[buf], ((c >> 1) - 1) & 1);
      for (int i = tid; i < FT_A / 16; i += 256) reinterpret_cast<uint4*>(At)[i] = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- A: capsule tests of this chunk's (up to) three bones over the band's source pixels
      for (int id = tid; id < 3 * NPX; id += 256) {
        const int s = id / NPX, pid = id - s * NPX;
        if (3 * c + s >= nact) break;
        const int hb = bars->act[3 * c + s];
        const int ys = y0 - 1 + (pid >> LOG2S), xs = pid & (S - 1);
        if (ys < 0 || ys >= S) continue;
        float wa, wb;
        if (bone_bbox_reject(geo[hb], xs + 0.5f, ys + 0.5f, distance)) continue;
        if (!bone_weights(geo[hb], xs + 0.5f, ys + 0.5f, distance, wa, wb)) continue;
        const __nv_bfloat16 ha = __float2bfloat16_rn(wa), hbq = __float2bfloat16_rn(wb);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yl = ys - ky + 1 - y0;  // output row (local) fed through tap row ky
          if (yl < 0 || yl >= R) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int x = xs - kx + 1;
            if (x < 0 || x >= S) continue;
            const int m = yl * S + x, col = s * 18 + ky * 3 + kx;  // role a; role b = col + 9
            __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(At + m * 128);
            row[(((col >> 3) ^ (m & 7)) << 3) | (col & 7)] = ha;
            row[((((col + 9) >> 3) ^ (m & 7)) << 3) | ((col + 9) & 7)] = hbq;
          }
        }
      }
      // ---- B: thread n gathers its channel of the chunk's 54 coefficient vectors (coalesced rows of P). All loads
      // are issued before the first store: one L2 round trip per chunk instead of three.
      {
        const int n = tid;
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(Bt + n * 128);
        PT v[54];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const bool on = 3 * c + s < nact;
          const PT* src = Pimg + (int64_t)(on ? bars->act[3 * c + s] : 0) * (2 * 9 * 256) + n;
#pragma unroll
          for (int t = 0; t < 18; ++t) v[s * 18 + t] = on ? __ldg(src + t * 256) : PT(0.f);  // t = role * 9 + tap
        }
#pragma unroll
        for (int col = 0; col < 54; ++col) {
          if constexpr (sizeof(PT) == 4) row[(((col >> 3) ^ (n & 7)) << 3) | (col & 7)] = __float2bfloat16_rn(v[col]);
          else row[(((col >> 3) ^ (n & 7)) << 3) | (col & 7)] = v[col];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bars->ready[buf]);
    }
    mbar_wait(&bars->done, 0);
    fence_after();
    // ---- epilogue: warp = (TMEM lane quarter, 128-column half)
    const int q = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + 128 * half;
    uint8_t* patch = smem + warp * (32 * FT_PATCH);  // aliases the A/B chunks: every MMA has completed
    __nv_bfloat16* obase = out + ((int64_t)b * S * S + (int64_t)y0 * S) * 256;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int col0 = 128 * half + 64 * cc;
      float v[64];
      if (nchunks > 0) {
        tmem_ld32(trow + 64 * cc, v);
        tmem_ld32(trow + 64 * cc + 32, v + 32);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = 0.f;
      }
      uint4* mine = reinterpret_cast<uint4*>(patch + lane * FT_PATCH);
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 8 * g8 + 2 * e;
          const float a0 = fmaxf(fmaf(v[i], __ldg(scale + col0 + i), __ldg(shift + col0 + i)), 0.f);
          const float a1 = fmaxf(fmaf(v[i + 1], __ldg(scale + col0 + i + 1), __ldg(shift + col0 + i + 1)), 0.f);
          __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
          w4[e] = *reinterpret_cast<uint32_t*>(&hh);
        }
        mine[g8] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {  // 4 rows x 128 bytes per store instruction
        const int rr = 4 * it + (lane >> 3), ch = lane & 7;
        const uint4 o = *reinterpret_cast<const uint4*>(patch + rr * FT_PATCH + ch * 16);
        *reinterpret_cast<uint4*>(obase + (int64_t)(q * 32 + rr) * 256 + col0 + ch * 8) = o;
      }
      __syncwarp();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

}  // namespace

void launch_pack_fusion_weight(const float* w, float* wp, cudaStream_t st) {
  int64_t total = (int64_t)40 * 64 * 9 * 256;
  pack_fusion_weight_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(w, wp);
}

size_t bone_coef_tc_packed_bytes() { return (size_t)40 * 9 * CT_WTILE; }

void launch_pack_fusion_weight_tc(const float* w, void* wpk, cudaStream_t st) {
  int64_t total = (int64_t)40 * 9 * 256 * 64;
  pack_fusion_weight_tc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(w, reinterpret_cast<float*>(wpk));
}

void launch_bone_coef_tc(const float* jf, const void* wpk, void* P, int p_bf16, int B, cudaStream_t st) {
  if (p_bf16)
    launch_pdl(bone_coef_tc_kernel<__nv_bfloat16>, dim3(40, ceil_div(2 * B, 128)), dim3(192), CT_SMEM, st, jf,
               reinterpret_cast<const uint8_t*>(wpk), reinterpret_cast<__nv_bfloat16*>(P), B);
  else
    launch_pdl(bone_coef_tc_kernel<float>, dim3(40, ceil_div(2 * B, 128)), dim3(192), CT_SMEM, st, jf,
               reinterpret_cast<const uint8_t*>(wpk), reinterpret_cast<float*>(P), B);
}

void launch_bone_coef(const float* jf, const float* wp, float* P, int B, cudaStream_t st) {
  launch_pdl(bone_coef_kernel, dim3(dim3(40, 9, ceil_div(2 * B, 64))), dim3(256), 2 * 32 * 128 * 4 + 16, st, jf, wp, P, B);
}

template <typename T>
void launch_bone_fusion(const float* rec, int rec_stride, const float* P, const float* scale, const float* shift, T* out,
                        int B, int S, float distance, cudaStream_t st) {
  const int nchunks = (3 * S * 40 + 31) / 32;
  const size_t smem = (size_t)S * 256 * 4 + (size_t)3 * S * 40 * sizeof(Entry) + (nchunks + 1) * 4 + 128 * 4;
  if (S == 32) {
    launch_pdl(bone_fusion_kernel<T, 32>, dim3(dim3(B, S)), dim3(256), smem, st, rec, rec_stride, P, scale, shift, out, distance);
  } else {
    launch_pdl(bone_fusion_kernel<T, 16>, dim3(dim3(B, S)), dim3(256), smem, st, rec, rec_stride, P, scale, shift, out, distance);
  }
}
template <typename PT>
static void launch_bone_fusion_tc_t(const float* rec, int rec_stride, const PT* P, const float* scale, const float* shift,
                                    __nv_bfloat16* out, int B, int S, float distance, cudaStream_t st) {
  if (S == 32) {
    launch_pdl(bone_fusion_tc_kernel<32, PT>, dim3(B, 8), dim3(FT_THREADS), FT_SMEM, st, rec, rec_stride, P, scale, shift,
               out, distance);
  } else {
    launch_pdl(bone_fusion_tc_kernel<16, PT>, dim3(B, 2), dim3(FT_THREADS), FT_SMEM, st, rec, rec_stride, P, scale, shift,
               out, distance);
  }
}

void launch_bone_fusion_tc(const float* rec, int rec_stride, const void* P, int p_bf16, const float* scale,
                           const float* shift, __nv_bfloat16* out, int B, int S, float distance, cudaStream_t st) {
  if (p_bf16)
    launch_bone_fusion_tc_t(rec, rec_stride, reinterpret_cast<const __nv_bfloat16*>(P), scale, shift, out, B, S, distance, st);
  else
    launch_bone_fusion_tc_t(rec, rec_stride, reinterpret_cast<const float*>(P), scale, shift, out, B, S, distance, st);
}

template void launch_bone_fusion<float>(const float*, int, const float*, const float*, const float*, float*, int, int,
                                        float, cudaStream_t);
template void launch_bone_fusion<__nv_bfloat16>(const float*, int, const float*, const float*, const float*,
                                                __nv_bfloat16*, int, int, float, cudaStream_t);

}  // namespace dirb200
