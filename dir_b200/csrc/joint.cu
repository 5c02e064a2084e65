// Joint-space kernels of one refinement stage (models/dir.py:86-130): everything between the
// feature map and the fusion conv. All arithmetic fp32; feature maps may be fp32 or bf16 NHWC.
//   joint_embed : F.grid_sample gather + img2joint filters + pos_emb          (dir.py:94-101,197-200)
//   gcn_layer   : PGraphConv + BN1d + ReLU (+ global_pos_emb on the last one) (p_graph_conv.py:39-60, p_gcn.py:20-27, dir.py:106-110)
//   ste         : mixSTE blocks 1..3 + head                                    (mixSTE.py:194-205)
//   bone_raster : bone_proj rasterisation                                      (dir.py:132-174)
#include "../../include/dirb200.h"
#include "common.cuh"
#include "kernels.h"

namespace dirb200 {

namespace {

constexpr int NJ = 21;

// =================================================================== joint_embed
template <typename T>
__global__ void __launch_bounds__(128) joint_embed_kernel(EmbedArgs a) {
  pdl_wait();
  extern __shared__ __align__(128) float dyn_smem[];  // weight stream: 2 x 32 x 128 floats + 2 mbarriers
  WStream ws;
  wstream_init(ws, dyn_smem, reinterpret_cast<uint64_t*>(dyn_smem + 2 * 32 * 128));
  __shared__ __align__(16) float samp[NJ][256];
  __shared__ __align__(16) float hid[NJ][128];
  __shared__ float xyz[NJ][3];
  __shared__ float uv[NJ][2];
  const int b = blockIdx.x, hand = blockIdx.y, n = threadIdx.x;
  const float* rec = a.prev_record + (int64_t)b * a.rec_stride;
  if (n < NJ * 3) xyz[n / 3][n % 3] = rec[DIRB200_OFF_JOINT_L + hand * 63 + n] / 0.15f;
  if (n < NJ * 2) uv[n / 2][n % 2] = rec[DIRB200_OFF_UV_L + hand * 42 + n];
  __syncthreads();

  // ---- bilinear gather, zeros padding, align_corners=False
  const int S = a.S;
  const T* feat = reinterpret_cast<const T*>(a.feat) + (int64_t)b * S * S * 256;
  for (int j = 0; j < NJ; ++j) {
    float ix = ((uv[j][0] + 1.f) * S - 1.f) * 0.5f;
    float iy = ((uv[j][1] + 1.f) * S - 1.f) * 0.5f;
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy;
    float wx1 = ix - fx, wy1 = iy - fy;
    float wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        int xi = x0 + dx, yi = y0 + dy;
        if (xi < 0 || xi >= S || yi < 0 || yi >= S) continue;
        float w = (dx ? wx1 : wx0) * (dy ? wy1 : wy0);
        const T* p = feat + ((int64_t)yi * S + xi) * 256;
        v0 = fmaf(ActIO<T>::ld(p + n), w, v0);
        v1 = fmaf(ActIO<T>::ld(p + n + 128), w, v1);
      }
    samp[j][n] = v0;
    samp[j][n + 128] = v1;
  }
  __syncthreads();

  // Three chained (21 x K)·(K x 128) products on the weight-streaming CTA GEMM: 3 row groups of 7 joints x 32
  // column groups = 96 register tiles.
  float outv[7][4];
  {  // ---- filters: 256 -> 128 (BN, ReLU) -> 128
    const PointMlp& f = a.filters[hand];
    cta_gemm<7>(&samp[0][0], 256, NJ, 256, f.w1t, 128, 128, ws, [&](int, int row, int c0, float (&v)[4]) {
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(f.s1 + c0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(f.b1 + c0));
      *reinterpret_cast<float4*>(&hid[row][c0]) =
          make_float4(fmaxf(fmaf(v[0], s1.x, b1.x), 0.f), fmaxf(fmaf(v[1], s1.y, b1.y), 0.f),
                      fmaxf(fmaf(v[2], s1.z, b1.z), 0.f), fmaxf(fmaf(v[3], s1.w, b1.w), 0.f));
    });
    cta_gemm<7>(&hid[0][0], 128, NJ, 128, f.w2t, 128, 128, ws, [&](int r, int, int c0, float (&v)[4]) {
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(f.b2 + c0));
      outv[r][0] = v[0] + b2.x; outv[r][1] = v[1] + b2.y; outv[r][2] = v[2] + b2.z; outv[r][3] = v[3] + b2.w;
    });
  }
  if (a.skip_pos) {  // seam: ImgFeature2JointFeature.forward alone (models/dir.py:197-200)
    float* out = a.out + ((int64_t)(b * 2 + hand) * NJ) * 128;
    const int cg = n % 32, rg = n / 32;
    if (rg < 3) {
#pragma unroll
      for (int r = 0; r < 7; ++r)
        *reinterpret_cast<float4*>(out + (rg * 7 + r) * 128 + cg * 4) =
            make_float4(outv[r][0], outv[r][1], outv[r][2], outv[r][3]);
    }
    return;
  }
  {  // ---- pos_emb: 3 -> 128 (BN, ReLU) -> 128
    const PointMlp& f = a.pos[hand];
    float w0 = f.w1t[n], w1 = f.w1t[128 + n], w2 = f.w1t[256 + n];
    float s1 = f.s1[n], b1 = f.b1[n];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float h = fmaf(xyz[j][2], w2, fmaf(xyz[j][1], w1, xyz[j][0] * w0));
      hid[j][n] = fmaxf(fmaf(h, s1, b1), 0.f);
    }
    __syncthreads();
    float* out = a.out + ((int64_t)(b * 2 + hand) * NJ) * 128;
    cta_gemm<7>(&hid[0][0], 128, NJ, 128, f.w2t, 128, 128, ws, [&](int r, int row, int c0, float (&v)[4]) {
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(f.b2 + c0));
      *reinterpret_cast<float4*>(out + row * 128 + c0) =
          make_float4(outv[r][0] + (v[0] + b2.x), outv[r][1] + (v[1] + b2.y), outv[r][2] + (v[2] + b2.z),
                      outv[r][3] + (v[3] + b2.w));
    });
  }
}

// =================================================================== SemGCN (see kernels.h: GcnGemmArgs)
constexpr int GBT = 32;  // images per CTA

// value of the aggregated + normalised layer output X'[b][hand][i][c4*4 .. +3]
__device__ __forceinline__ float4 gcn_aggregate(const float* __restrict__ hin, const GcnAgg& g, int B, int b, int hand,
                                                int i, int c4) {
  const size_t plane = (size_t)B * 2 * NJ * 128;  // H1 offset
  const float* base = hin + ((size_t)(b * 2 + hand) * NJ) * 128 + c4 * 4;
  float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)i * 128));
  const float* A1 = g.A1[hand] + i * NJ;
  for (int j = 0; j < NJ; ++j) {  // ascending j, like the dense A1 @ h1 of the reference
    const float aw = __ldg(A1 + j);
    if (aw == 0.f) continue;
    const float4 h = __ldg(reinterpret_cast<const float4*>(base + plane + (size_t)j * 128));
    v.x = fmaf(aw, h.x, v.x); v.y = fmaf(aw, h.y, v.y); v.z = fmaf(aw, h.z, v.z); v.w = fmaf(aw, h.w, v.w);
  }
  const float4 sc = __ldg(reinterpret_cast<const float4*>(g.scale[hand] + c4 * 4));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(g.shift[hand] + c4 * 4));
  return make_float4(fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f),
                     fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
}

// grid (42 = k*21 + j, 2 hands, ceil(B/32)); 256 threads = 8 row groups (4 images) x 32 column groups
__global__ void __launch_bounds__(256) gcn_gemm_kernel(GcnGemmArgs a) {
  pdl_wait();
  extern __shared__ __align__(128) float dyn_smem[];  // weight stream: 2 x 32 x 128 floats + 2 mbarriers
  WStream ws;
  wstream_init(ws, dyn_smem, reinterpret_cast<uint64_t*>(dyn_smem + 2 * 32 * 128));
  __shared__ __align__(16) float xs[GBT][128];
  const int k = blockIdx.x / NJ, j = blockIdx.x % NJ, hand = blockIdx.y, b0 = blockIdx.z * GBT, tid = threadIdx.x;
  const int nb = min(GBT, a.B - b0);
  for (int e = tid; e < GBT * 32; e += 256) {
    const int bb = e >> 5, c4 = e & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bb < nb) {
      if (a.x)
        v = __ldg(reinterpret_cast<const float4*>(a.x + ((size_t)((b0 + bb) * 2 + hand) * NJ + j) * 128) + c4);
      else
        v = gcn_aggregate(a.hin, a.agg, a.B, b0 + bb, hand, j, c4);
    }
    *reinterpret_cast<float4*>(&xs[bb][c4 * 4]) = v;
  }
  __syncthreads();
  const float* Wkj = a.W[hand] + (size_t)(k * NJ + j) * 128 * 128;
  float* out = a.hout + (size_t)k * a.B * 2 * NJ * 128;
  cta_gemm<4>(&xs[0][0], 128, nb, 128, Wkj, 128, 128, ws, [&](int, int row, int c0, float (&v)[4]) {
    *reinterpret_cast<float4*>(out + ((size_t)((b0 + row) * 2 + hand) * NJ + j) * 128 + c0) =
        make_float4(v[0], v[1], v[2], v[3]);
  });
}

// grid (21 joints, 2 hands, ceil(B/32)); 256 threads
__global__ void __launch_bounds__(256) gcn_finish_kernel(GcnFinishArgs a) {
  pdl_wait();
  extern __shared__ __align__(128) float dyn_smem[];
  WStream ws;
  wstream_init(ws, dyn_smem, reinterpret_cast<uint64_t*>(dyn_smem + 2 * 32 * 128));
  __shared__ __align__(16) float xs[GBT][128];
  const int i = blockIdx.x, hand = blockIdx.y, b0 = blockIdx.z * GBT, tid = threadIdx.x;
  const int nb = min(GBT, a.B - b0);
  if (a.skip_gpos) {  // seam: the SemGCN stack alone (SemGCN/p_gcn.py:63-73)
    for (int e = tid; e < nb * 32; e += 256) {
      const int row = e >> 5, c4 = e & 31;
      *reinterpret_cast<float4*>(a.y + ((size_t)((b0 + row) * 2 + hand) * NJ + i) * 128 + c4 * 4) =
          gcn_aggregate(a.hin, a.agg, a.B, b0 + row, hand, i, c4);
    }
    return;
  }
  const PointMlp& g = a.gpos;
  const float sgn = hand == 0 ? -1.f : 1.f;
  for (int e = tid; e < GBT * 128; e += 256) {  // hidden layer of global_pos_emb: 3 -> 128, BN, ReLU
    const int bb = e >> 7, kk = e & 127;
    float h = 0.f;
    if (bb < nb) {
      const float* rec = a.prev_record + (size_t)(b0 + bb) * a.rec_stride;
      const float* p = rec + DIRB200_OFF_JOINT_L + hand * 63 + i * 3;
      const float* off = rec + DIRB200_OFF_OFFSET;
      const float px = p[0] / 0.15f + sgn * (off[0] / 2.f);
      const float py = p[1] / 0.15f + sgn * (off[1] / 2.f);
      const float pz = p[2] / 0.15f + sgn * (off[2] / 2.f);
      h = fmaxf(fmaf(fmaf(pz, g.w1t[256 + kk], fmaf(py, g.w1t[128 + kk], px * g.w1t[kk])), g.s1[kk], g.b1[kk]), 0.f);
    }
    xs[bb][kk] = h;
  }
  __syncthreads();
  cta_gemm<4>(&xs[0][0], 128, nb, 128, g.w2t, 128, 128, ws, [&](int, int row, int c0, float (&v)[4]) {
    const float4 x = gcn_aggregate(a.hin, a.agg, a.B, b0 + row, hand, i, c0 >> 2);
    const float4 b2 = __ldg(reinterpret_cast<const float4*>(g.b2 + c0));
    *reinterpret_cast<float4*>(a.y + ((size_t)((b0 + row) * 2 + hand) * NJ + i) * 128 + c0) =
        make_float4(x.x + (v[0] + b2.x), x.y + (v[1] + b2.y), x.z + (v[2] + b2.z), x.w + (v[3] + b2.w));
  });
}

// =================================================================== STE
// One CTA per image: the whole 3-block transformer + head with the token state in shared memory.
// 8 consumer warps do the math; a 9th warp is a dedicated weight producer that streams every Linear's K-major
// weights (100 slabs of 32 x 128 floats per image) through a 4-deep smem ring with cp.async.bulk + full/empty
// mbarriers, running ahead across layer boundaries, so no GEMM ever starts on a cold L2 fetch.
constexpr int NT = 42;
constexpr int STE_CONSUMERS = 512;  // two groups of 256: group g multiplies the slabs with (slab index & 1) == g
constexpr int STE_GROUP = 256;
constexpr int STE_THREADS = STE_CONSUMERS + 32;
constexpr int SC_LD = 44;
constexpr int RING_D = 4;
constexpr int SLAB_FLOATS = 32 * 128;
constexpr int STE_NSEG = 22;
constexpr int QLD = 388;  // qkv row pitch: 388 % 32 = 4 banks of skew per token, so 128-bit K-row reads by 8
                          // consecutive lanes cover all 32 banks (a pitch of 384 makes them 8-way conflicted)
constexpr int STE_SMEM_FLOATS = RING_D * SLAB_FLOATS + NT * 128 * 2 + NT * QLD + 16 * 48 + NT * 128;
constexpr int STE_SMEM_BYTES = STE_SMEM_FLOATS * 4 + 2 * RING_D * 8 + STE_NSEG * 24 + 64;

struct WSeg {
  const float* ptr;  // first row of this 128-(or 64-)column block, K-major
  int ldb;           // floats between consecutive k rows
  int nslab;         // K / 32
  int nc;            // columns in the block (128 or 64)
};
struct WRing {
  float* buf;
  uint64_t* full;
  uint64_t* empty;
  uint32_t it;  // slabs consumed so far
};

__device__ __forceinline__ void ste_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = cta_smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// producer warp: walk the schedule, one 32-row slab per ring slot
__device__ __forceinline__ void ste_producer(const WSeg* segs, int nseg, float* ring, uint64_t* full, uint64_t* empty) {
  const int lane = threadIdx.x & 31;
  uint32_t g = 0;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const WSeg sg = segs[sgi];
    for (int sl = 0; sl < sg.nslab; ++sl, ++g) {
      const uint32_t buf = g % RING_D;
      if (g >= RING_D) mbar_wait_parity(&empty[buf], ((g / RING_D) - 1) & 1);
      const uint32_t bar = cta_smem_u32(&full[buf]);
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(32 * sg.nc * 4))
                     : "memory");
      __syncwarp();
      if (lane < 16) {  // slab = 16 k-pair rows of (nc x 2) floats
        const uint32_t dst = cta_smem_u32(ring + buf * SLAB_FLOATS + lane * sg.nc * 2);
        const float* src = sg.ptr + (size_t)(sl * 16 + lane) * sg.ldb;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"((uint32_t)(sg.nc * 8)), "r"(bar)
                     : "memory");
      }
    }
  }
}

// consumer side of one (M x K)·(K x NC) product; the next K/32 slabs of the ring must hold its weights.
// 512 consumers = 2 groups x (TM x 4 register tiles); group g takes the slabs of its parity, the two partial sums
// meet in `part` (group 1 -> smem, group 0 adds and runs the epilogue). Inner product on packed FFMA2: the pair
// (even k, odd k) of one output rides in one 64-bit register, operands come pre-paired from smem.
template <int TM, typename Epi>
__device__ __forceinline__ void ring_gemm(const float* __restrict__ A, int lda, int M, int K, int NC, WRing& rg,
                                          float* __restrict__ part, Epi epi) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> 8, gt = tid & 255;
  const int ncg = NC >> 2, nrg = (M + TM - 1) / TM;
  const int cg = gt % ncg, rgi = gt / ncg;
  const bool active = gt < ncg * nrg;
  const int nslab = K >> 5;
  float2 acc[TM][4];
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = make_float2(0.f, 0.f);
  const float* arow[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) arow[r] = A + (size_t)min(rgi * TM + r, M - 1) * lda;
  for (int sl = 0; sl < nslab; ++sl) {
    const uint32_t g = rg.it + sl;
    if ((int)(g & 1) != grp) continue;
    const uint32_t buf = g % RING_D;
    mbar_wait_parity(&rg.full[buf], (g / RING_D) & 1);
    if (active) {
      const float* wb = rg.buf + buf * SLAB_FLOATS + cg * 8;  // row p: [n][2] pairs
#pragma unroll
      for (int kp = 0; kp < 16; kp += 2) {  // two k-pairs = 4 k per step
        const float4 b0 = *reinterpret_cast<const float4*>(wb + (kp + 0) * NC * 2);      // (k0,k1) for n, n+1
        const float4 b1 = *reinterpret_cast<const float4*>(wb + (kp + 0) * NC * 2 + 4);  // (k0,k1) for n+2, n+3
        const float4 b2 = *reinterpret_cast<const float4*>(wb + (kp + 1) * NC * 2);      // (k2,k3) for n, n+1
        const float4 b3 = *reinterpret_cast<const float4*>(wb + (kp + 1) * NC * 2 + 4);
#pragma unroll
        for (int r = 0; r < TM; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(arow[r] + sl * 32 + kp * 2);
          const float2 a01 = make_float2(a.x, a.y), a23 = make_float2(a.z, a.w);
          acc[r][0] = __ffma2_rn(a01, make_float2(b0.x, b0.y), acc[r][0]);
          acc[r][1] = __ffma2_rn(a01, make_float2(b0.z, b0.w), acc[r][1]);
          acc[r][2] = __ffma2_rn(a01, make_float2(b1.x, b1.y), acc[r][2]);
          acc[r][3] = __ffma2_rn(a01, make_float2(b1.z, b1.w), acc[r][3]);
          acc[r][0] = __ffma2_rn(a23, make_float2(b2.x, b2.y), acc[r][0]);
          acc[r][1] = __ffma2_rn(a23, make_float2(b2.z, b2.w), acc[r][1]);
          acc[r][2] = __ffma2_rn(a23, make_float2(b3.x, b3.y), acc[r][2]);
          acc[r][3] = __ffma2_rn(a23, make_float2(b3.z, b3.w), acc[r][3]);
        }
      }
    }
    __syncwarp();
    if (lane == 0)  // this warp is done with the slot
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cta_smem_u32(&rg.empty[buf])) : "memory");
  }
  if (grp == 1 && active) {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int row = rgi * TM + r;
      if (row < M)
        *reinterpret_cast<float4*>(part + row * 128 + cg * 4) =
            make_float4(acc[r][0].x + acc[r][0].y, acc[r][1].x + acc[r][1].y, acc[r][2].x + acc[r][2].y,
                        acc[r][3].x + acc[r][3].y);
    }
  }
  ste_bar();
  if (grp == 0 && active) {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int row = rgi * TM + r;
      if (row < M) {
        const float4 p = *reinterpret_cast<const float4*>(part + row * 128 + cg * 4);
        float v[4] = {(acc[r][0].x + acc[r][0].y) + p.x, (acc[r][1].x + acc[r][1].y) + p.y,
                      (acc[r][2].x + acc[r][2].y) + p.z, (acc[r][3].x + acc[r][3].y) + p.w};
        epi(r, row, cg * 4, v);
      }
    }
  }
  rg.it += nslab;
}

// out[r][n] = sum_k in[r][k] * Wt[k][n] + bias[n]; optional GELU; optional residual accumulate into out.
// 42 token rows = 7 row groups of 6; N is walked in 128-column blocks (the producer's schedule order).
template <int MODE>  // 0: store, 1: store GELU, 2: out += result
__device__ __forceinline__ void ste_linear(const float* __restrict__ in, int ldin, int K,
                                           const float* __restrict__ bias, int N, float* __restrict__ out, int ldout,
                                           WRing& rg, float* __restrict__ part) {
  for (int cb = 0; cb < N; cb += 128) {
    const int NC = min(128, N - cb);
    if (cb) ste_bar();  // `part` of the previous column block has been consumed
    ring_gemm<6>(in, ldin, NT, K, NC, rg, part, [&](int, int row, int c0, float (&v)[4]) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + cb + c0));
      float4 o = make_float4(v[0] + bb.x, v[1] + bb.y, v[2] + bb.z, v[3] + bb.w);
      if (MODE == 1) {
        o.x = 0.5f * o.x * (1.f + erff(o.x * 0.70710678118654752440f));
        o.y = 0.5f * o.y * (1.f + erff(o.y * 0.70710678118654752440f));
        o.z = 0.5f * o.z * (1.f + erff(o.z * 0.70710678118654752440f));
        o.w = 0.5f * o.w * (1.f + erff(o.w * 0.70710678118654752440f));
      }
      float4* po = reinterpret_cast<float4*>(out + row * ldout + cb + c0);
      if (MODE == 2) {
        const float4 p = *po;
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      *po = o;
    });
  }
}

// LayerNorm over 128 channels, one warp per row (two-pass like ATen); consumer warps only
__device__ __forceinline__ void ste_layernorm(const float* __restrict__ in, float* __restrict__ out,
                                              const float* __restrict__ w, const float* __restrict__ b, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < NT; r += STE_CONSUMERS / 32) {
    float4 v = *reinterpret_cast<const float4*>(in + r * 128 + lane * 4);
    float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
    float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
    float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
    float rs = rsqrtf(var + eps);
    float4 g = __ldg(reinterpret_cast<const float4*>(w + lane * 4));
    float4 be = __ldg(reinterpret_cast<const float4*>(b + lane * 4));
    float4 o = make_float4(dx * rs * g.x + be.x, dy * rs * g.y + be.y, dz * rs * g.z + be.z, dw * rs * g.w + be.w);
    *reinterpret_cast<float4*>(out + r * 128 + lane * 4) = o;
  }
}

__global__ void __launch_bounds__(STE_THREADS) ste_kernel(const float* __restrict__ xin, float* __restrict__ yout,
                                                          SteWeights w) {
  extern __shared__ __align__(128) float sm[];
  float* ring = sm;                        // [4][32][128] weight slabs
  float* x = ring + RING_D * SLAB_FLOATS;  // [42][128] residual stream
  float* h = x + NT * 128;                 // [42][128] LN output / attention output
  float* big = h + NT * 128;               // [42][388] qkv, later [42][256] MLP hidden
  float* sc = big + NT * QLD;              // [16 warps][48] one attention probability row per warp
  float* part = sc + 16 * 48;              // [42][128] partial sums of consumer group 1
  uint64_t* full = reinterpret_cast<uint64_t*>(part + NT * 128);
  uint64_t* empty = full + RING_D;
  WSeg* segs = reinterpret_cast<WSeg*>(empty + RING_D);
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < RING_D; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cta_smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cta_smem_u32(&empty[i])), "r"(STE_GROUP / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int n = 0;
    for (int l = 0; l < 3; ++l) {  // consumption order of ste_linear calls below
      const SteWeights::Block& B = w.blk[l];
      // k-pair interleaved weights [K/2][N][2]: a row (one k-pair) is 2N floats, column block cb starts at 2*cb
      for (int cb = 0; cb < 384; cb += 128) segs[n++] = WSeg{B.qkv_t + 2 * cb, 2 * 384, 4, 128};
      segs[n++] = WSeg{B.proj_t, 2 * 128, 4, 128};
      for (int cb = 0; cb < 256; cb += 128) segs[n++] = WSeg{B.fc1_t + 2 * cb, 2 * 256, 4, 128};
      segs[n++] = WSeg{B.fc2_t, 2 * 128, 8, 128};
    }
    segs[n++] = WSeg{w.head_t, 2 * 64, 4, 64};
  }
  __syncthreads();
  if (tid >= STE_CONSUMERS) {  // producer warp: weights only, so it starts before the previous kernel has finished
    ste_producer(segs, STE_NSEG, ring, full, empty);
    return;
  }
  pdl_wait();
  WRing rg{ring, full, empty, 0};
  for (int i = tid; i < NT * 128; i += STE_CONSUMERS) x[i] = xin[(int64_t)b * NT * 128 + i] + w.pos[i];
  ste_bar();
  for (int l = 0; l < 3; ++l) {
    const SteWeights::Block& B = w.blk[l];
    ste_layernorm(x, h, B.n1w, B.n1b, 1e-6f);
    ste_bar();
    ste_linear<0>(h, 128, 128, B.qkv_b, 384, big, QLD, rg, part);
    ste_bar();
    {  // attention, one warp per (head, query token): scores -> softmax -> P·V without leaving the warp
      const int warp = tid >> 5, lane = tid & 31;
      float* prow = sc + warp * 48;
      for (int r = warp; r < 4 * NT; r += STE_CONSUMERS / 32) {
        const int hd = r / NT, i = r - hd * NT;
        const float* q = big + i * QLD + hd * 32;
        const int j1 = lane + 32;
        const float* k0 = big + lane * QLD + 128 + hd * 32;
        const float* k1 = big + min(j1, NT - 1) * QLD + 128 + hd * 32;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(q + d);
          const float4 a4 = *reinterpret_cast<const float4*>(k0 + d);
          const float4 b4 = *reinterpret_cast<const float4*>(k1 + d);
          s0 = fmaf(q4.x, a4.x, s0); s0 = fmaf(q4.y, a4.y, s0); s0 = fmaf(q4.z, a4.z, s0); s0 = fmaf(q4.w, a4.w, s0);
          s1 = fmaf(q4.x, b4.x, s1); s1 = fmaf(q4.y, b4.y, s1); s1 = fmaf(q4.z, b4.z, s1); s1 = fmaf(q4.w, b4.w, s1);
        }
        s0 *= 0.17677669529663688110f;  // 32^-0.5 (mixSTE.py:59)
        s1 = (j1 < NT) ? s1 * 0.17677669529663688110f : -INFINITY;
        const float mx = warp_max(fmaxf(s0, s1));
        const float e0 = expf(s0 - mx), e1 = (j1 < NT) ? expf(s1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        prow[lane] = e0 * inv;
        if (j1 < NT) prow[j1] = e1 * inv;
        __syncwarp();
        const float* v = big + 256 + hd * 32 + lane;  // lane = channel of this head
        float o = 0.f;
#pragma unroll 6
        for (int j = 0; j < NT; ++j) o = fmaf(prow[j], v[j * QLD], o);
        h[i * 128 + hd * 32 + lane] = o;
        __syncwarp();
      }
    }
    ste_bar();
    ste_linear<2>(h, 128, 128, B.proj_b, 128, x, 128, rg, part);
    ste_bar();
    ste_layernorm(x, h, B.n2w, B.n2b, 1e-6f);
    ste_bar();
    ste_linear<1>(h, 128, 128, B.fc1_b, 256, big, 256, rg, part);
    ste_bar();
    ste_linear<2>(big, 256, 256, B.fc2_b, 128, x, 128, rg, part);
    ste_bar();
    ste_layernorm(x, x, w.snw, w.snb, 1e-6f);  // shared spatial_norm (mixSTE.py:200), in place (row-local)
    ste_bar();
  }
  ste_layernorm(x, h, w.hnw, w.hnb, 1e-5f);
  ste_bar();
  ste_linear<0>(h, 128, 128, w.head_b, 64, big, 64, rg, part);
  ste_bar();
  for (int i = tid; i < NT * 64; i += STE_CONSUMERS) yout[(int64_t)b * NT * 64 + i] = big[i];
}

// =================================================================== bone rasterisation
struct BoneGeom {
  float ax, ay, bx, by, dx, dy;
};

// Per (pixel, bone): capsule test + endpoint weights, op-for-op as the reference (no FMA contraction in
// the mask so that boundary pixels agree with the PyTorch eager kernels).
__device__ __forceinline__ bool bone_weights(const BoneGeom& g, float px, float py, float distance, float& wa,
                                             float& wb) {
  float s = __fadd_rn(__fmul_rn(__fsub_rn(g.ax, px), g.dx), __fmul_rn(__fsub_rn(g.ay, py), g.dy));
  float t = __fadd_rn(__fmul_rn(__fsub_rn(px, g.bx), g.dx), __fmul_rn(__fsub_rn(py, g.by), g.dy));
  float h = fmaxf(fmaxf(s, t), 0.f);
  float c = __fsub_rn(__fmul_rn(__fsub_rn(px, g.ax), g.dy), __fmul_rn(__fsub_rn(py, g.ay), g.dx));
  float dist = hypotf(h, c);
  if (!(dist < distance)) return false;  // NaN (degenerate bone) -> false
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

__device__ __forceinline__ BoneGeom make_bone(const float* uv /*21x2*/, int bone, int S) {
  int pa = (bone % 4 == 0) ? 0 : bone;  // parents [0,1,2,3,0,5,6,7,...], child = bone+1
  int ch = bone + 1;
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

template <typename T>
__global__ void __launch_bounds__(256) bone_raster_kernel(const float* __restrict__ rec, int rec_stride,
                                                          const float* __restrict__ jf, T* __restrict__ out, int S,
                                                          float distance) {
  pdl_wait();
  __shared__ BoneGeom geo[40];
  __shared__ float uv[84];
  __shared__ __align__(16) float feat[2 * NJ * 64];
  const int b = blockIdx.x, y = blockIdx.y, tid = threadIdx.x;
  if (tid < 84) uv[tid] = rec[(int64_t)b * rec_stride + DIRB200_OFF_UV_L + tid];
  for (int i = tid; i < 2 * NJ * 64; i += 256) feat[i] = jf[(int64_t)b * 2 * NJ * 64 + i];
  __syncthreads();
  if (tid < 40) geo[tid] = make_bone(uv + (tid / 20) * 42, tid % 20, S);
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const float py = y + 0.5f;
  T* orow = out + ((int64_t)b * S + y) * S * 2560;
  for (int item = warp; item < S * 40; item += 8) {
    int x = item / 40, hb = item % 40;
    float wa, wb;
    float2 v = make_float2(0.f, 0.f);
    if (bone_weights(geo[hb], x + 0.5f, py, distance, wa, wb)) {
      int hand = hb / 20, bone = hb % 20;
      int pa = (bone % 4 == 0) ? 0 : bone, ch = bone + 1;
      float2 fa = *reinterpret_cast<const float2*>(&feat[(hand * NJ + pa) * 64 + lane * 2]);
      float2 fb = *reinterpret_cast<const float2*>(&feat[(hand * NJ + ch) * 64 + lane * 2]);
      v.x = __fadd_rn(__fmul_rn(fa.x, wa), __fmul_rn(fb.x, wb));
      v.y = __fadd_rn(__fmul_rn(fa.y, wa), __fmul_rn(fb.y, wb));
    }
    T* o = orow + (int64_t)x * 2560 + hb * 64 + lane * 2;
    if (sizeof(T) == 4) {
      *reinterpret_cast<float2*>(o) = v;
    } else {
      *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v.x, v.y);
    }
  }
}

// vis = left (+ right), NCHW fp32: out[b][bone*64+c][y][x]
__global__ void __launch_bounds__(256) bone_vis_kernel(const float* __restrict__ uv_l, const float* __restrict__ uv_r,
                                                       int uv_stride, const float* __restrict__ f_l,
                                                       const float* __restrict__ f_r, int f_stride,
                                                       float* __restrict__ out, int S, float distance, int add_right) {
  pdl_wait();
  __shared__ float uv[2][42];
  __shared__ float fe[2][2][64];  // [hand][a/b][c]
  __shared__ BoneGeom geo[2];
  const int b = blockIdx.x, bone = blockIdx.y, tid = threadIdx.x;
  const int pa = (bone % 4 == 0) ? 0 : bone, ch = bone + 1;
  if (tid < 42) uv[0][tid] = uv_l[(int64_t)b * uv_stride + tid];
  if (tid >= 64 && tid < 106 && add_right) uv[1][tid - 64] = uv_r[(int64_t)b * uv_stride + tid - 64];
  if (tid >= 128) {
    int c = (tid - 128) & 63, ab = (tid - 128) >> 6;
    fe[0][ab][c] = f_l[(int64_t)b * f_stride + (ab ? ch : pa) * 64 + c];
    fe[1][ab][c] = add_right ? f_r[(int64_t)b * f_stride + (ab ? ch : pa) * 64 + c] : 0.f;
  }
  __syncthreads();
  if (tid == 0) geo[0] = make_bone(uv[0], bone, S);
  if (tid == 32 && add_right) geo[1] = make_bone(uv[1], bone, S);
  __syncthreads();
  float* ob = out + ((int64_t)b * 1280 + bone * 64) * S * S;
  for (int p = tid; p < S * S; p += 256) {
    int y = p / S, x = p % S;
    float wal = 0.f, wbl = 0.f, war = 0.f, wbr = 0.f;
    bool ml = bone_weights(geo[0], x + 0.5f, y + 0.5f, distance, wal, wbl);
    bool mr = add_right ? bone_weights(geo[1], x + 0.5f, y + 0.5f, distance, war, wbr) : false;
    for (int c = 0; c < 64; ++c) {
      float vl = ml ? __fadd_rn(__fmul_rn(fe[0][0][c], wal), __fmul_rn(fe[0][1][c], wbl)) : 0.f;
      float vr = mr ? __fadd_rn(__fmul_rn(fe[1][0][c], war), __fmul_rn(fe[1][1][c], wbr)) : 0.f;
      ob[(int64_t)c * S * S + p] = vl + vr;
    }
  }
}

}  // namespace

template <typename T>
void launch_joint_embed(const EmbedArgs& a, cudaStream_t st) {
  launch_pdl(joint_embed_kernel<T>, dim3(dim3(a.B, 2)), dim3(128), 2 * 32 * 128 * 4 + 16, st, a);
}
template void launch_joint_embed<float>(const EmbedArgs&, cudaStream_t);
template void launch_joint_embed<__nv_bfloat16>(const EmbedArgs&, cudaStream_t);

void launch_gcn_gemm(const GcnGemmArgs& a, cudaStream_t st) {
  launch_pdl(gcn_gemm_kernel, dim3(dim3(2 * NJ, 2, ceil_div(a.B, GBT))), dim3(256), 2 * 32 * 128 * 4 + 16, st, a);
}

void launch_gcn_finish(const GcnFinishArgs& a, cudaStream_t st) {
  launch_pdl(gcn_finish_kernel, dim3(dim3(NJ, 2, ceil_div(a.B, GBT))), dim3(256), 2 * 32 * 128 * 4 + 16, st, a);
}

void launch_ste(const float* x, float* y, const SteWeights& w, int B, cudaStream_t st) {
  const int smem = STE_SMEM_BYTES;
  launch_pdl(ste_kernel, dim3(B), dim3(STE_THREADS), smem, st, x, y, w);
}

template <typename T>
void launch_bone_raster(const float* rec, int rec_stride, const float* jf, T* out, int B, int S, float distance,
                        cudaStream_t st) {
  launch_pdl(bone_raster_kernel<T>, dim3(dim3(B, S)), dim3(256), 0, st, rec, rec_stride, jf, out, S, distance);
}
template void launch_bone_raster<float>(const float*, int, const float*, float*, int, int, float, cudaStream_t);
template void launch_bone_raster<__nv_bfloat16>(const float*, int, const float*, __nv_bfloat16*, int, int, float,
                                                cudaStream_t);

void launch_bone_vis_nchw(const float* uv_l, const float* uv_r, int uv_stride, const float* f_l, const float* f_r,
                          int f_stride, float* out, int B, int S, float distance, int add_right, cudaStream_t st) {
  launch_pdl(bone_vis_kernel, dim3(dim3(B, 20)), dim3(256), 0, st, uv_l, uv_r, uv_stride, f_l, f_r, f_stride, out, S, distance, add_right);
}

}  // namespace dirb200
