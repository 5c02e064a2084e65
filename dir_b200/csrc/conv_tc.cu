// tcgen05 / TMA implicit-GEMM convolution for sm_100a: bf16 operands, fp32 accumulation in TMEM,
// fused affine (+residual) (+ReLU) epilogue, persistent CTAs. NHWC activations, weights [Cout][(ky,kx,ci)].
//
//   D[m, n] = sum_k A[m, k] * W[n, k],  m = (b, ho, wo) (128 consecutive output pixels per tile),
//                                        k = (ky, kx, ci) walked in blocks of 64 channels of one tap.
//
// A operand: one 4-D TMA box per (tap, 64-channel block): {64 ch, wbox*stride, hbox*stride, nbox} of the NHWC
//   input with elementStrides {1, stride, stride, 1}; the tap offset (ky-pad, kx-pad) is a coordinate shift and
//   TMA zero-fills out-of-bounds pixels, so im2col and padding cost no instructions.
// B operand: 2-D TMA box {64 k, BN rows} of the packed weights. Both land K-major with 128B swizzle, i.e. the
//   canonical UMMA SWIZZLE_128B K-major layout (8-row groups 1024 B apart).
// Stem variant (7x7 stride 2, Cin=3; models/backbone/resnet.py:176): the image is first repacked to a
//   zero-padded NHWC4 bf16 buffer; a k-block is one kernel row (8 px x 4 ch = 32 elements = 64 B, SWIZZLE_64B)
//   and the A box is taken from an OVERLAPPING-stride view of that buffer (consecutive wo are 2 px = 16 B apart),
//   so the 7x7 window gather is again pure TMA.
// Persistent schedule (grid = min(tiles, SMs)), roles per CTA (320 threads):
//   warp 0   : TMA producer over a `stages`-deep smem ring (runs ahead across tiles)
//   warp 1   : TMEM alloc (2 accumulator buffers of BN fp32 columns) + single-thread tcgen05.mma issue
//   warps 2-9: two epilogue groups of 4 warps taking alternate 64-column chunks; the epilogue of tile i overlaps the
//              main loop of tile i+1: tcgen05.ld -> affine/residual/ReLU -> bf16 -> 128B-swizzled smem staging ->
//              TMA store (coalesced, asynchronous); residual tiles arrive by TMA through a ring that the epilogue group
//              clocks itself (whoever frees a slot issues the load of the chunk that uses it next).
//   warps 10-15 (PRE kernels only): apply the consuming Residual's bn1 + ReLU to every A tile in place, between its
//              TMA arrival and the MMA (one warp per smem stage).
// Variants selected per layer at launch:
//   CG = 2       : clusters of two CTAs share every MMA (tcgen05.mma.cta_group::2, M = 256); each CTA loads its own A
//                  tile and half of the weight tile (every layer with BN >= 128 and an even number of M tiles)
//   b_resident   : single-n-tile layers whose whole weight matrix fits next to the A ring load it once per CTA
//   kb2 / kb2a   : K-concatenated second (and third) 1x1 operand: conv3 + downsample/skip as one GEMM, the skip operand
//                  optionally read from the two sources of a channel concat that is never materialised
//   stem         : legacy TMA stem (DIRB200_STEM_SPLIT=1); the default stem is stem_pool.cu
//   PRE = 1      : 1x1 conv over relu(x * s[c] + h[c]) with x optionally the channel concat of two tensors (hourglass
//                  Residual conv1: neither the pre-activated tensor nor the concat exists in HBM)
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>

#include "common.cuh"
#include "engine.h"

namespace dirb200 {

namespace {

constexpr int BM = 128;
constexpr int NUM_THREADS = 320;  // producer warp + MMA warp + 2 epilogue groups of 4 warps
constexpr int MAX_STAGES = 8;
constexpr int CHUNK_BYTES = BM * 128;  // one 64-column bf16 epilogue box: 16 KB
constexpr int PRE_WARPS = 6;           // transform warps of the PRE kernels (16 warps = 4 per scheduler: same 128-register cap as 14)
constexpr int RES_BUFS = 4;            // residual ring: chunks are fetched RES_BUFS chunks (1-2 tiles) ahead of their use

struct TcArgs {
  const float* scale;
  const float* shift;
  int M, Cout, Ho, Wo;
  int stride, pad, kw;
  int taps, cblocks;  // K loop = taps * cblocks k-blocks
  int relu, has_res, stem;
  int m_tiles, n_tiles, stages;
  int raster_m;      // 1: consecutive tiles walk m first (weights larger than activations: keep an n-tile's weights hot)
  int kb2, stride2;  // K-concatenated second operand: kb2 extra 1x1 k-blocks read from tmA2 at spatial stride2
  int kb2a;          // ... of which the first kb2a come from tmA2 and the rest from tmA3 (a channel concat never built)
  int b_resident;    // 1: the whole weight matrix (one n-tile, all k-blocks) is loaded once per CTA and stays in smem
  // PRE kernels (1x1 conv with the consuming Residual's pre-activation folded into the A operand, hourglass.py:60-61):
  const float* pre_scale;  // per INPUT channel: A <- relu(A * pre_scale + pre_shift) in shared memory, before the MMA
  const float* pre_shift;
  int pre_cb1;             // the first pre_cb1 channel blocks come from tmA, the rest from tmA2 (virtual channel concat)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 2-CTA variants (cute::SM100_TMA_2SM_LOAD_*): the data lands in THIS CTA's smem, the transaction bytes are counted
// on the leader CTA's barrier (peer bit of the shared::cluster address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
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
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {  // {bf16(max(lo,0)), bf16(max(hi,0))}, RN
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// UMMA shared-memory descriptor, K-major, rows of SW bytes (SW = 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B),
// 8-row groups 8*SW bytes apart (cute::UMMA::SmemDescriptor / make_umma_desc<Major::K>).
template <int SW>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);         // start address, 16-byte units
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)((8 * SW) >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)(SW == 128 ? 2 : 4) << 61;         // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// cta_group::2: one instruction drives the tensor cores of both SMs of the pair (M = 256: rows 0-127 from the leader's
// smem / into the leader's TMEM, rows 128-255 the peer's; each CTA supplies half of the N rows of B).
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void epi_barrier(int group) {  // named barrier of one 128-thread epilogue group
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

template <int BN, int SW, int CG = 1>
struct TcCfg {
  static constexpr int A_STAGE = BM * SW;
  static constexpr int B_STAGE = (BN / CG) * SW;  // 2-CTA mode: each CTA of the pair holds half of the N rows
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int KSTEPS = SW / 32;  // tcgen05.mma K=16 bf16 = 32 B per step
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, K-major both, N=BN, M=128 per CTA
  static constexpr uint32_t IDESC =
      (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static constexpr int NCH = BN / 64;
  // bslots = weight tiles held in smem: one per stage, or (resident mode) one per k-block
  static int smem_bytes(int stages, int has_res, int bslots, int pre = 0) {
    return 1024 + stages * A_STAGE + bslots * B_STAGE + 2 * CHUNK_BYTES + (has_res ? RES_BUFS * CHUNK_BYTES : 0) +
           4 * BN * 4 + 256 + (pre ? 128 : 0);
  }
};

// PRE = 1 (1x1 convs only): PRE_WARPS more warps (10-15) apply relu(x * pre_scale[c] + pre_shift[c]) to every A tile in place
// between its TMA arrival and the MMA: the pre-activated copy of the input never exists in HBM. A tiles then complete on
// a CTA-local barrier (afull_bar) that the transform warps wait on; the MMA warp waits for the weights (full_bar) and for
// the transform warps of BOTH CTAs of the pair (xf_bar on the leader, remote arrivals).
template <int BN, int SW, int CG, int PRE = 0>
__global__ void __launch_bounds__(NUM_THREADS + PRE * PRE_WARPS * 32, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3, const TcArgs a) {
  using Cfg = TcCfg<BN, SW, CG>;
  // CG == 2: launched as clusters of two CTAs that share every MMA (tile = 256 output pixels x BN: this CTA owns rows
  // 128*rank..+127 and loads its own A tile plus HALF of the weight tile, so the L2->SM bytes per FLOP drop by a third;
  // the 3x3 / wide layers are bound by exactly that fabric). Only rank 0 issues tcgen05.mma.
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = a.stages;
  uint8_t* sA = smem;
  uint8_t* sB = sA + stages * Cfg::A_STAGE;
  const int bres = a.b_resident;  // weights resident: B slot = k-block (loaded once), stages carry only A
  const int nkb_all = a.taps * a.cblocks + a.kb2;
  uint8_t* sOut = sB + (bres ? nkb_all : stages) * Cfg::B_STAGE;  // 2 x 16 KB output staging (128B-swizzled boxes)
  uint8_t* sRes = sOut + 2 * CHUNK_BYTES;                 // RES_BUFS x 16 KB residual ring (only if has_res)
  float* s_affine = reinterpret_cast<float*>(sRes + (a.has_res ? RES_BUFS * CHUNK_BYTES : 0));  // [2 groups][2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_affine + 4 * BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint64_t* res_empty = res_full + RES_BUFS;
  uint64_t* bres_bar = res_empty + RES_BUFS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint64_t* afull_bar = bres_bar + 2;           // PRE only (the extra 128 bytes of smem_bytes(.., pre = 1))
  uint64_t* xf_bar = afull_bar + MAX_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_units = a.m_tiles / CG;  // scheduling unit = CG adjacent M tiles (one per CTA of the pair)
  const int total_tiles = m_units * a.n_tiles;
  const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;
  const int nkb1 = a.taps * a.cblocks;
  const int nkb = nkb1 + a.kb2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256 * CG);  // CG == 2: both CTAs' epilogues arrive on the leader's barrier
    }
    for (int i = 0; i < RES_BUFS; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 1);
    }
    mbar_init(bres_bar, 1);
    if constexpr (PRE) {
      for (int s = 0; s < MAX_STAGES; ++s) {
        mbar_init(&afull_bar[s], 1);
        mbar_init(&xf_bar[s], CG);  // one arrival per CTA of the pair (the transform warp that owns the k-block)
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: two accumulator buffers of BN fp32 columns x 128 lanes
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer
      int s = 0;
      uint32_t ph = 0;
      if (bres && first_tile < total_tiles) {  // weights: finalize-time data, one load per CTA for all its tiles
        mbar_expect_tx(bres_bar, (uint32_t)nkb * Cfg::B_STAGE);
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&tmB, bres_bar, sB + kb * Cfg::B_STAGE, kb * 64, 0);
      }
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int m0 = ((a.raster_m ? tile % m_units : tile / a.n_tiles) * CG + (int)rank) * BM;
        const int n0 = (a.raster_m ? tile / m_units : tile % a.n_tiles) * BN;
        const int nb0 = n0 + (int)rank * (BN / CG);  // first weight row this CTA loads
        const int wo0 = m0 % a.Wo;
        const int ho0 = (m0 / a.Wo) % a.Ho;
        const int b0 = m0 / (a.Wo * a.Ho);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if constexpr (PRE) {  // A -> this CTA's afull barrier (transform warps); weights -> the MMA warp's full barrier
            mbar_expect_tx(&afull_bar[s], Cfg::A_STAGE);
            const bool first = kb < a.pre_cb1;
            tma_load_4d(first ? &tmA : &tmA2, &afull_bar[s], sA + s * Cfg::A_STAGE, (first ? kb : kb - a.pre_cb1) * 64,
                        wo0, ho0, b0);
            if constexpr (CG == 2) {
              if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::B_STAGE);
              tma_load_2d_2sm(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 64, nb0);
            } else {
              mbar_expect_tx(&full_bar[s], Cfg::B_STAGE);
              tma_load_2d(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 64, n0);
            }
          } else if constexpr (CG == 2) {  // both CTAs' bytes are counted on the leader's full barrier
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
            if (kb >= nkb1) {
              const int k2 = kb - nkb1;
              tma_load_4d_2sm(k2 < a.kb2a ? &tmA2 : &tmA3, &full_bar[s], sA + s * Cfg::A_STAGE,
                              (k2 < a.kb2a ? k2 : k2 - a.kb2a) * 64, wo0 * a.stride2, ho0 * a.stride2, b0);
            } else {
              const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
              const int ky = tap / a.kw, kx = tap - ky * a.kw;
              tma_load_4d_2sm(&tmA, &full_bar[s], sA + s * Cfg::A_STAGE, cb * 64, wo0 * a.stride + kx - a.pad,
                              ho0 * a.stride + ky - a.pad, b0);
            }
            tma_load_2d_2sm(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 64, nb0);
          } else if (bres) {  // A only (1-CTA, non-stem layers)
            mbar_expect_tx(&full_bar[s], Cfg::A_STAGE);
            if (kb >= nkb1) {
              const int k2 = kb - nkb1;
              tma_load_4d(k2 < a.kb2a ? &tmA2 : &tmA3, &full_bar[s], sA + s * Cfg::A_STAGE,
                          (k2 < a.kb2a ? k2 : k2 - a.kb2a) * 64, wo0 * a.stride2, ho0 * a.stride2, b0);
            } else {
              const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
              const int ky = tap / a.kw, kx = tap - ky * a.kw;
              tma_load_4d(&tmA, &full_bar[s], sA + s * Cfg::A_STAGE, cb * 64, wo0 * a.stride + kx - a.pad,
                          ho0 * a.stride + ky - a.pad, b0);
            }
          } else {
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          if (a.stem) {  // k-block = kernel row ky; overlapped view: {32 elems, wo (16 B apart), h, n}
            tma_load_4d(&tmA, &full_bar[s], sA + s * Cfg::A_STAGE, 0, wo0, ho0 * 2 + kb - a.pad, b0);
            tma_load_2d(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 32, n0);
          } else if (kb >= nkb1) {  // second operand of a K-concatenated pair (1x1, own stride): skip / downsample conv
            const int k2 = kb - nkb1;
            tma_load_4d(k2 < a.kb2a ? &tmA2 : &tmA3, &full_bar[s], sA + s * Cfg::A_STAGE,
                        (k2 < a.kb2a ? k2 : k2 - a.kb2a) * 64, wo0 * a.stride2, ho0 * a.stride2, b0);
            tma_load_2d(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 64, n0);
          } else {
            const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            tma_load_4d(&tmA, &full_bar[s], sA + s * Cfg::A_STAGE, cb * 64, wo0 * a.stride + kx - a.pad,
                        ho0 * a.stride + ky - a.pad, b0);
            tma_load_2d(&tmB, &full_bar[s], sB + s * Cfg::B_STAGE, kb * 64, n0);
          }
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ===================== MMA issuer (the leader CTA issues for the pair)
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      if (bres && first_tile < total_tiles) mbar_wait(bres_bar, 0);
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++i) {
        const int buf = i & 1;
        mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);  // epilogue drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          if constexpr (PRE) mbar_wait(&xf_bar[s], ph);  // A tiles of the pair transformed in place
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc<SW>(smem_u32(sA + s * Cfg::A_STAGE));
          const uint64_t db = umma_desc<SW>(smem_u32(sB + (bres ? kb : s) * Cfg::B_STAGE));
#pragma unroll
          for (int k = 0; k < Cfg::KSTEPS; ++k) {  // +32 B per K=16 step inside the swizzle atom
            if constexpr (CG == 2) umma_bf16_2sm(d, da + 2 * k, db + 2 * k, Cfg::IDESC, (kb | k) ? 1u : 0u);
            else umma_bf16(d, da + 2 * k, db + 2 * k, Cfg::IDESC, (kb | k) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs) once these MMAs have consumed it
          if constexpr (CG == 2) umma_commit_2sm(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CG == 2) umma_commit_2sm(&tmem_full[buf]);
        else umma_commit(&tmem_full[buf]);
      }
    }
  } else if (PRE && warp >= 10) {
    // ===================== pre-activation of the A operand, in place in the 128B-swizzled stage. Warp w owns smem stage
    // w (the host caps `stages` at PRE_WARPS), i.e. every stages-th k-block of this CTA's stream, so that many
    // wait -> load -> store -> fence -> arrive chains run concurrently. (A warp must see EVERY phase of a barrier it
    // waits on: handing k-blocks out round-robin over a different number of warps lets a warp reach a stage one phase
    // early, where a parity wait passes at once.) lane = 16-byte chunk column q (8 channels: scale/shift in registers
    // for the k-block) x rows rr + 4 j: one warp instruction touches 4 full 128-byte rows (conflict-free under the swizzle).
    const int w = warp - 10;
    const int q = lane & 7, rr = lane >> 3;
    const uint32_t coff0 = (uint32_t)((q ^ rr) << 4), coff1 = (uint32_t)((q ^ (rr + 4)) << 4);  // row & 7 = rr + 4 (j & 1)
    const int my_tiles = first_tile < total_tiles ? (total_tiles - first_tile + tile_step - 1) / tile_step : 0;
    const int total_kb = my_tiles * nkb;
    const int s = w;
    int kb = w;
    uint32_t ph = 0;
    while (kb >= nkb) kb -= nkb;
    for (int g = w; w < stages && g < total_kb; g += stages) {
      const float4* ps = reinterpret_cast<const float4*>(a.pre_scale + kb * 64 + q * 8);
      const float4* pb = reinterpret_cast<const float4*>(a.pre_shift + kb * 64 + q * 8);
      const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), h0 = __ldg(pb), h1 = __ldg(pb + 1);
      mbar_wait(&afull_bar[s], ph);
      const uint32_t base = smem_u32(sA + s * Cfg::A_STAGE) + (uint32_t)rr * 128u;
#pragma unroll 1
      for (int jb = 0; jb < 32; jb += 8) {
        uint4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = lds128(base + (jb + j) * 512 + ((j & 1) ? coff1 : coff0));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[j]);
          const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
          const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
          uint4 o;
          o.x = pack_relu_bf16x2(fmaf(f0.x, s0.x, h0.x), fmaf(f0.y, s0.y, h0.y));
          o.y = pack_relu_bf16x2(fmaf(f1.x, s0.z, h0.z), fmaf(f1.y, s0.w, h0.w));
          o.z = pack_relu_bf16x2(fmaf(f2.x, s1.x, h1.x), fmaf(f2.y, s1.y, h1.y));
          o.w = pack_relu_bf16x2(fmaf(f3.x, s1.z, h1.z), fmaf(f3.y, s1.w, h1.w));
          sts128(base + (jb + j) * 512 + ((j & 1) ? coff1 : coff0), o);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> the MMA's async proxy
      __syncwarp();
      if (lane == 0) {
        // same publication pattern as the epilogue's tmem_empty hand-back: proxy fence, then a (remote) arrive.
        // (An explicit .release.cluster here compiles to MEMBAR.ALL.GPU per k-block and tripled the kernel time.)
        if constexpr (CG == 2) mbar_arrive_leader(&xf_bar[s]);
        else mbar_arrive(&xf_bar[s]);
      }
      ph ^= 1;
      kb += stages;
      while (kb >= nkb) kb -= nkb;
    }
  } else {
    // ===================== epilogue: two groups of 4 warps take alternate 64-column chunks (global chunk parity),
    // each with its own staging buffer, named barrier and bulk-store groups, so one group's TMEM->smem->TMA chain
    // overlaps the other's. Warp w may touch TMEM lanes 32*(w%4) .. +31.
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    float* s_scale = s_affine + eg * 2 * BN;
    float* s_shift = s_scale + BN;
    uint8_t* sOutG = sOut + eg * CHUNK_BYTES;
    int i = 0;
    int cur_n0 = -1;
    // Residual chunks arrive by TMA in a ring of RES_BUFS slots, slot = chunk index % RES_BUFS. The ring is clocked by its
    // consumer: the elected thread of the group that has finished with a slot issues the load of the chunk that uses it
    // next (RES_BUFS chunks ahead, same chunk parity = same group). Issuing them from the producer warp instead blocked
    // it on the epilogue's progress before it could fetch the next tile's operands (BN = 256: one tile fills the ring).
    auto load_res = [&](uint32_t j) {  // chunk j of this CTA's chunk stream
      const int tile_j = first_tile + (int)(j / Cfg::NCH) * tile_step;
      if (tile_j >= total_tiles) return;
      const int c_j = (int)(j % Cfg::NCH);
      const int m0_j = ((a.raster_m ? tile_j % m_units : tile_j / a.n_tiles) * CG + (int)rank) * BM;
      const int n0_j = (a.raster_m ? tile_j / m_units : tile_j % a.n_tiles) * BN;
      const uint32_t rb = j % RES_BUFS;
      mbar_expect_tx(&res_full[rb], CHUNK_BYTES);
      tma_load_2d(&tmR, &res_full[rb], sRes + rb * CHUNK_BYTES, n0_j + c_j * 64, m0_j);
    };
    if (a.has_res && et == 0) {  // this group's first RES_BUFS / 2 chunks
      for (uint32_t j = (uint32_t)eg; j < RES_BUFS; j += 2) load_res(j);
    }
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++i) {
      const int m0 = ((a.raster_m ? tile % m_units : tile / a.n_tiles) * CG + (int)rank) * BM;
      const int n0 = (a.raster_m ? tile / m_units : tile % a.n_tiles) * BN;
      const int buf = i & 1;
      if (n0 != cur_n0) {  // (re)stage the per-channel affine of this n-tile; the group's readers are past (d)
        for (int j = et; j < BN; j += 128) {
          s_scale[j] = a.scale[n0 + j];
          s_shift[j] = a.shift[n0 + j];
        }
        cur_n0 = n0;
        epi_barrier(eg);
      }
      mbar_wait(&tmem_full[buf], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < Cfg::NCH; ++c) {
        const uint32_t rchunk = (uint32_t)i * Cfg::NCH + c;  // global chunk index of this CTA
        if ((int)(rchunk & 1) != eg) continue;
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + buf * BN + c * 64;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t sc = smem_u32(s_scale + c * 64);
        const uint32_t sh = smem_u32(s_shift + c * 64);
        uint4 packed[8];
        if (a.has_res) {
          const uint32_t rb = rchunk % RES_BUFS;
          mbar_wait(&res_full[rb], (rchunk / RES_BUFS) & 1);
          const uint32_t rrow = smem_u32(sRes + rb * CHUNK_BYTES + row * 128);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint4 u = lds128(rrow + ((q ^ swz) << 4));
            const __nv_bfloat162* hres = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float4 s0 = lds_f4(sc + q * 32), s1 = lds_f4(sc + q * 32 + 16);
            const float4 h0 = lds_f4(sh + q * 32), h1 = lds_f4(sh + q * 32 + 16);
            float v[8];
            v[0] = fmaf(__uint_as_float(r[q * 8 + 0]), s0.x, h0.x);
            v[1] = fmaf(__uint_as_float(r[q * 8 + 1]), s0.y, h0.y);
            v[2] = fmaf(__uint_as_float(r[q * 8 + 2]), s0.z, h0.z);
            v[3] = fmaf(__uint_as_float(r[q * 8 + 3]), s0.w, h0.w);
            v[4] = fmaf(__uint_as_float(r[q * 8 + 4]), s1.x, h1.x);
            v[5] = fmaf(__uint_as_float(r[q * 8 + 5]), s1.y, h1.y);
            v[6] = fmaf(__uint_as_float(r[q * 8 + 6]), s1.z, h1.z);
            v[7] = fmaf(__uint_as_float(r[q * 8 + 7]), s1.w, h1.w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hres[e]);
              v[2 * e] += f.x;
              v[2 * e + 1] += f.y;
            }
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&packed[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 s0 = lds_f4(sc + q * 32), s1 = lds_f4(sc + q * 32 + 16);
            const float4 h0 = lds_f4(sh + q * 32), h1 = lds_f4(sh + q * 32 + 16);
            float v[8];
            v[0] = fmaf(__uint_as_float(r[q * 8 + 0]), s0.x, h0.x);
            v[1] = fmaf(__uint_as_float(r[q * 8 + 1]), s0.y, h0.y);
            v[2] = fmaf(__uint_as_float(r[q * 8 + 2]), s0.z, h0.z);
            v[3] = fmaf(__uint_as_float(r[q * 8 + 3]), s0.w, h0.w);
            v[4] = fmaf(__uint_as_float(r[q * 8 + 4]), s1.x, h1.x);
            v[5] = fmaf(__uint_as_float(r[q * 8 + 5]), s1.y, h1.y);
            v[6] = fmaf(__uint_as_float(r[q * 8 + 6]), s1.z, h1.z);
            v[7] = fmaf(__uint_as_float(r[q * 8 + 7]), s1.w, h1.w);
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&packed[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          }
        }
        // (a) this group's previous TMA store must have finished reading its staging buffer
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        epi_barrier(eg);  // (b)
        const uint32_t orow = smem_u32(sOutG + row * 128);
#pragma unroll
        for (int q = 0; q < 8; ++q) sts128(orow + ((q ^ swz) << 4), packed[q]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (c) generic-proxy writes -> async proxy
        epi_barrier(eg);  // (d) staging complete; every thread of the group is also done reading its residual slot
        if (et == 0) {
          tma_store_2d(&tmY, sOutG, n0 + c * 64, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (a.has_res) load_res(rchunk + RES_BUFS);  // every thread of the group is past its reads of this slot
        }
      }
      // this thread has read everything it needs from accumulator buffer `buf` (256 arrivals hand it back to the MMA warp)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty[buf]);
      else mbar_arrive(&tmem_empty[buf]);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output bytes written
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA may exit while the pair's MMAs / remote arrivals are in flight
  else __syncthreads();
  if (warp == 1) {
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS)
                   : "memory");
  }
}

// NCHW fp32 image -> zero-padded NHWC4 bf16: out[b][h][w + 3][c], row pitch (W + 8) pixels (stem operand)
__global__ void stem_pack_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  pdl_wait();
  const int Wp = W + 8;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * Wp) return;
  int wp = (int)(idx % Wp);
  int64_t t = idx / Wp;
  int h = (int)(t % H);
  int b = (int)(t / H);
  int w = wp - 3;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (w >= 0 && w < W) {
    const float* p = img + ((int64_t)b * 3 * H + h) * W + w;
    v0 = __ldg(p);
    v1 = __ldg(p + (int64_t)H * W);
    v2 = __ldg(p + 2 * (int64_t)H * W);
  }
  __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), c = __floats2bfloat162_rn(v2, 0.f);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(out + idx * 4) = u;
}

// uint8 HWC BGR image -> the same padded NHWC4 bf16 operand, with the reference's preprocessing fused in
// (apps/eval.py:56-61: BGR->RGB, /255, ImageNet mean/std)
__global__ void stem_pack_u8_kernel(const unsigned char* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H,
                                    int W) {
  pdl_wait();
  const int Wp = W + 8;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * Wp) return;
  int wp = (int)(idx % Wp);
  int64_t t = idx / Wp;
  int h = (int)(t % H);
  int b = (int)(t / H);
  int w = wp - 3;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (w >= 0 && w < W) {
    const unsigned char* p = img + (((int64_t)b * H + h) * W + w) * 3;  // B, G, R
    v0 = ((float)p[2] / 255.f - 0.485f) / 0.229f;
    v1 = ((float)p[1] / 255.f - 0.456f) / 0.224f;
    v2 = ((float)p[0] / 255.f - 0.406f) / 0.225f;
  }
  __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), c = __floats2bfloat162_rn(v2, 0.f);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(out + idx * 4) = u;
}

// stem weights [64][3][ks][ks] fp32 -> [64][ks*32] bf16 with k = ky*32 + kx*4 + c (zero elsewhere); ks = 7 or 3
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int ks) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int kk = ks * 32;
  if (idx >= 64 * kk) return;
  int k = idx % kk, n = idx / kk;
  int ky = k / 32, r = k % 32, kx = r / 4, c = r % 4;
  float v = (kx < ks && c < 3) ? w[((n * 3 + c) * ks + ky) * ks + kx] : 0.f;
  out[idx] = __float2bfloat16_rn(v);
}

// dst[n][k] = k < K1 ? w1[n][k] * s1[n] : w2[n][k - K1] * s2[n]  (bf16); shift[n] = h1[n] + h2[n]; scale[n] = 1
__global__ void pack_dual_weight_kernel(const float* __restrict__ w1, int K1, const float* __restrict__ s1,
                                        const float* __restrict__ h1, const float* __restrict__ w2, int K2,
                                        const float* __restrict__ s2, const float* __restrict__ h2,
                                        __nv_bfloat16* __restrict__ dst, float* __restrict__ scale,
                                        float* __restrict__ shift, int Cout) {
  const int K = K1 + K2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Cout * K) return;
  const int k = (int)(idx % K), n = (int)(idx / K);
  const float v = k < K1 ? w1[(size_t)n * K1 + k] * s1[n] : w2[(size_t)n * K2 + (k - K1)] * s2[n];
  dst[idx] = __float2bfloat16_rn(v);
  if (k == 0) {
    scale[n] = 1.f;
    shift[n] = h1[n] + h2[n];
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int pick_bn(int Cout) { return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64); }

struct Boxes {
  int wbox, hbox, nbox;
};
Boxes pick_boxes(int Ho, int Wo) {
  Boxes b;
  b.wbox = Wo < BM ? Wo : BM;
  b.hbox = Ho < BM / b.wbox ? Ho : BM / b.wbox;
  b.nbox = BM / (b.wbox * b.hbox);
  return b;
}

// [rows][cols] bf16 row-major viewed as 64-column x 128-row boxes with 128B swizzle (epilogue store / residual load)
bool make_rowmajor_map(CUtensorMap* tm, const void* ptr, int rows, int cols) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BM};
  cuuint32_t es[2] = {1, 1};
  return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool cg2_enabled() {
  static const bool on = [] {
    const char* e = getenv("DIRB200_TC_CG2");
    return !(e && e[0] == '0');
  }();
  return on;
}

bool resident_enabled() {
  static const bool on = [] {
    const char* e = getenv("DIRB200_TC_RESIDENT");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int BN, int SW, int CG, int PRE = 0>
int launch_tc_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                 const CUtensorMap& tmA2, const CUtensorMap& tmA3, TcArgs a, cudaStream_t st) {
  using Cfg = TcCfg<BN, SW, CG>;
  const int nkb = a.taps * a.cblocks + a.kb2;
  // resident weights: one n-tile whose whole K extent fits next to a useful A ring (layer1: <= 72 KB); the TMA engine
  // then only fetches activations (it retires ~one 128-byte box row per 3-5 cycles, the limit of these layers)
  a.b_resident = (CG == 1 && SW == 128 && !PRE && !a.stem && a.n_tiles == 1 && nkb * Cfg::B_STAGE <= 80 * 1024 &&
                  resident_enabled())
                     ? 1 : 0;
  int stages = MAX_STAGES;  // the ring runs ahead across tiles, so short K loops still want every stage that fits
  while (stages > 1 && Cfg::smem_bytes(stages, a.has_res, a.b_resident ? nkb : stages, PRE) > 227 * 1024) --stages;
  if (PRE && stages > PRE_WARPS) stages = PRE_WARPS;  // one transform warp per stage
  a.stages = stages;
  const int smem = Cfg::smem_bytes(stages, a.has_res, a.b_resident ? nkb : stages, PRE);
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(conv_tc_kernel<BN, SW, CG, PRE>), 227 * 1024) != cudaSuccess)
    return DIRB200_E_CUDA;
  const int units = (a.m_tiles / CG) * a.n_tiles;        // scheduling units (one per CTA, or per CTA pair)
  const int slots = num_sms() / CG;
  const int grid = (units < slots ? units : slots) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS + PRE * PRE_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (a.kb2a == 0) a.kb2a = a.kb2;  // no third operand
  if (cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SW, CG, PRE>, tmA, tmB, tmY, tmR, tmA2, tmA3, a) != cudaSuccess)
    return DIRB200_E_CUDA;
  return DIRB200_OK;
}

// tmB2 = the weight map with BN/2-row boxes (null: 1-CTA kernel only)
template <int BN, int SW>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmB2, const CUtensorMap& tmY,
              const CUtensorMap& tmR, const CUtensorMap& tmA2, TcArgs a, cudaStream_t st,
              const CUtensorMap* tmA3 = nullptr) {
  const CUtensorMap& t3 = tmA3 ? *tmA3 : tmA2;
  if constexpr (BN >= 128 && SW == 128) {
    if (tmB2 && cg2_enabled() && !a.stem && a.m_tiles % 2 == 0)
      return launch_tc_cg<BN, SW, 2>(tmA, *tmB2, tmY, tmR, tmA2, t3, a, st);
  }
  return launch_tc_cg<BN, SW, 1>(tmA, tmB, tmY, tmR, tmA2, t3, a, st);
}

// Tensor maps are cached per (pointer, geometry). Lookups hand out COPIES: an insertion may evict (clear) the cache, so a
// pointer into it could dangle while the same launch is still collecting its other maps.
template <typename Key>
struct MapCache {
  std::map<Key, CUtensorMap> m;
  bool find(const Key& k, CUtensorMap* out) const {
    auto it = m.find(k);
    if (it == m.end()) return false;
    *out = it->second;
    return true;
  }
  void put(const Key& k, const CUtensorMap& v) {
    if (m.size() > 8192) m.clear();
    m.emplace(k, v);
  }
};

bool rowmajor_map_cached(const void* ptr, int rows, int cols, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int> Key;
  static thread_local MapCache<Key> cache;
  Key k(ptr, rows, cols);
  if (cache.find(k, out)) return true;
  if (!make_rowmajor_map(out, ptr, rows, cols)) return false;
  cache.put(k, *out);
  return true;
}

}  // namespace

bool conv_tc_supported(const ConvLayer& L, int B, int H, int W) {
  if (!L.w16 || L.wmap_bn == 0 || !get_encode()) return false;
  if (L.tc_stem) return H % 2 == 0 && W == 256;  // one tile = one 128-pixel output row
  if (L.Cin % 64 != 0 || L.Cout % 64 != 0 || L.K != L.Kpad) return false;
  if (L.stride != 1 && L.stride != 2) return false;
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W + 2 * L.pad - L.kw) / L.stride + 1;
  if (!is_pow2(Ho) || !is_pow2(Wo)) return false;
  if (Wo > BM && Wo % BM != 0) return false;
  return true;
}

int conv_tc_prepare_weights(ConvLayer& L) {
  L.wmap_bn = 0;
  EncodeTiledFn enc = get_encode();
  if (!enc || L.Cin % 64 != 0 || L.Cout % 64 != 0 || L.K != L.Kpad) return 0;
  int bn = pick_bn(L.Cout);
  if (bn > L.tc_bn_cap) bn = L.tc_bn_cap;
  cuuint64_t dims[2] = {(cuuint64_t)L.Kpad, (cuuint64_t)L.Cout};
  cuuint64_t strides[1] = {(cuuint64_t)L.Kpad * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&L.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w16, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return DIRB200_E_CUDA;
  L.wmap_bn = bn;
  L.wmap2_ok = false;
  if (bn >= 128) {  // 2-CTA mode: each CTA of a pair loads half of the N rows of a weight tile
    cuuint32_t box2[2] = {64, (cuuint32_t)(bn / 2)};
    L.wmap2_ok = enc(&L.wmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w16, dims, strides, box2, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  return DIRB200_OK;
}

// Build the K-concatenated layer of a (main 1x1, second 1x1) pair; `w16`, `scale`, `shift` are caller-allocated.
int conv_tc_prepare_dual(ConvLayer& F, const ConvLayer& main, const ConvLayer& second, __nv_bfloat16* w16, float* scale,
                         float* shift, cudaStream_t st) {
  F = ConvLayer();
  if (!get_encode() || main.kh != 1 || second.kh != 1 || main.Cout != second.Cout || main.Cin % 64 || second.Cin % 64 ||
      main.Cout % 64)
    return 0;
  F.name = main.name + "+" + second.name;
  F.Cin = main.Cin + second.Cin;
  F.Cout = main.Cout;
  F.K = F.Kpad = F.Cin;
  F.relu = main.relu;
  F.w16 = w16;
  F.scale = scale;
  F.shift = shift;
  F.tc_bn_cap = 256;
  const int64_t total = (int64_t)F.Cout * F.K;
  pack_dual_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(main.w32, main.K, main.scale, main.shift,
                                                                           second.w32, second.K, second.scale,
                                                                           second.shift, w16, scale, shift, F.Cout);
  return conv_tc_prepare_weights(F);
}

// Stem (7x7 s2, 3->64): pack weights as [64][7*32] bf16 (k = ky*32 + kx*4 + c) into `w_packed` (caller-allocated,
// 64*224 bf16) and build the SWIZZLE_64B weight map.
int conv_tc_prepare_stem(ConvLayer& L, const float* w_raw, __nv_bfloat16* w_packed, cudaStream_t st) {
  L.tc_stem = false;
  EncodeTiledFn enc = get_encode();
  // 7x7 / pad 3 (ResNet, resnet.py:176) or 3x3 / pad 1 (HRNet's first stem conv): one k-block per kernel row
  const bool k7 = L.kh == 7 && L.kw == 7 && L.pad == 3, k3 = L.kh == 3 && L.kw == 3 && L.pad == 1;
  if (!enc || L.Cin != 3 || L.Cout != 64 || !(k7 || k3) || L.stride != 2) return 0;
  const int kk = L.kh * 32;
  stem_pack_weight_kernel<<<(64 * kk + 255) / 256, 256, 0, st>>>(w_raw, w_packed, L.kh);
  cuuint64_t dims[2] = {(cuuint64_t)kk, 64};
  cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
  cuuint32_t box[2] = {32, 64};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&L.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return DIRB200_E_CUDA;
  L.w16 = w_packed;
  L.wmap_bn = 64;
  L.tc_stem = true;
  return DIRB200_OK;
}

size_t conv_tc_stem_scratch_bytes(int B, int H, int W) { return (size_t)B * H * (W + 8) * 4 * 2; }

// image (fp32 NCHW, or raw uint8 HWC BGR frames with the reference's preprocessing fused in) -> zero-padded NHWC4 bf16
void launch_stem_pack(const float* img, const unsigned char* img_u8, __nv_bfloat16* scratch, int B, int H, int W,
                      cudaStream_t st) {
  const int64_t n = (int64_t)B * H * (W + 8);
  if (img_u8)
    launch_pdl(stem_pack_u8_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, img_u8, scratch, B, H, W);
  else
    launch_pdl(stem_pack_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, img, scratch, B, H, W);
}

int launch_conv_tc_stem(const ConvLayer& L, const float* img, const unsigned char* img_u8, __nv_bfloat16* scratch,
                        __nv_bfloat16* y, int B, int H, int W, cudaStream_t st) {
  const int Wp = W + 8, Ho = H / 2, Wo = W / 2;
  launch_stem_pack(img, img_u8, scratch, B, H, W, st);
  typedef std::tuple<const void*, int, int, int, int> Key;
  static thread_local MapCache<Key> cache;
  Key key(scratch, B, H, W, L.pad);
  CUtensorMap tm;
  if (!cache.find(key, &tm)) {
    // overlapping view of the padded NHWC4 buffer (pixel w at index w + 3): the 8-pixel window of output column wo starts
    // at pixel 2*wo - pad, i.e. the view's base is (3 - pad) pixels into the buffer (0 for the 7x7, 16 bytes for the 3x3)
    cuuint64_t dims[4] = {32, (cuuint64_t)Wo, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {16, (cuuint64_t)Wp * 8, (cuuint64_t)H * Wp * 8};
    cuuint32_t box[4] = {32, (cuuint32_t)BM, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, scratch + (3 - L.pad) * 4, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "dirb200: cuTensorMapEncodeTiled(stem A, overlapping strides) failed: %d\n", (int)r);
      return DIRB200_E_CUDA;
    }
    cache.put(key, tm);
  }
  const int M = B * Ho * Wo;
  CUtensorMap tmY;
  if (!rowmajor_map_cached(y, M, 64, &tmY)) return DIRB200_E_CUDA;
  TcArgs a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.M = M;
  a.Cout = 64;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = 2;
  a.pad = L.pad;
  a.kw = L.kw;
  a.taps = L.kh;  // k-blocks = kernel rows
  a.cblocks = 1;
  a.relu = L.relu;
  a.has_res = 0;
  a.stem = 1;
  a.m_tiles = (M + BM - 1) / BM;
  a.n_tiles = 1;
  return launch_tc<64, 64>(tm, L.wmap, nullptr, tmY, tmY, tm, a, st);
}

namespace {
// activation tensor map {C, W, H, N} with box {64, wbox*s, hbox*s, nbox}, cached per (pointer, geometry)
bool act_map_cached(const __nv_bfloat16* x, int B, int H, int W, int C, int stride, const Boxes& bx, const char* name,
                    CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int> Key;
  static thread_local MapCache<Key> cache;
  Key key(x, B, H, W, C, stride, bx.wbox, bx.hbox, bx.nbox);
  if (cache.find(key, out)) return true;
  CUtensorMap& t = *out;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(bx.wbox * stride), (cuuint32_t)(bx.hbox * stride), (cuuint32_t)bx.nbox};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = get_encode()(&t, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "dirb200: cuTensorMapEncodeTiled(A) failed: %d (layer %s)\n", (int)r, name);
    return false;
  }
  cache.put(key, t);
  return true;
}
}  // namespace

// K-concatenated pair: y = act( [x1 | x2(strided)] . [W1 | W2]^T + shift ) — a 1x1 conv over x1 (spatial Ho x Wo) fused
// with a 1x1 stride-`stride2` conv over x2 (the ResNet downsample / hourglass skip branch). L holds the concatenated,
// BN-scale-folded weights [Cout][C1 + C2].
int launch_conv_tc_dual(const ConvLayer& L, const __nv_bfloat16* x1, int C1, const __nv_bfloat16* x2, int C2,
                        int stride2, __nv_bfloat16* y, int B, int Ho, int Wo, cudaStream_t st, const __nv_bfloat16* x2b,
                        int C2b) {
  // x2b != null: the second operand is the channel concat [x2 (C2 - C2b channels) | x2b (C2b channels)], read from its
  // two sources (the concatenated tensor is never materialised)
  const Boxes bx = pick_boxes(Ho, Wo);
  const int C2a = x2b ? C2 - C2b : C2;
  if (C2a % 64 || (x2b && C2b % 64)) return DIRB200_E_CUDA;
  CUtensorMap tmA, tmA2, tmA3s, tmY;
  const int M = B * Ho * Wo;
  if (!act_map_cached(x1, B, Ho, Wo, C1, 1, bx, L.name.c_str(), &tmA) ||
      !act_map_cached(x2, B, Ho * stride2, Wo * stride2, C2a, stride2, bx, L.name.c_str(), &tmA2) ||
      (x2b && !act_map_cached(x2b, B, Ho * stride2, Wo * stride2, C2b, stride2, bx, L.name.c_str(), &tmA3s)) ||
      !rowmajor_map_cached(y, M, L.Cout, &tmY))
    return DIRB200_E_CUDA;
  const CUtensorMap* tmA3 = x2b ? &tmA3s : nullptr;
  TcArgs a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.M = M;
  a.Cout = L.Cout;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = 1;
  a.pad = 0;
  a.kw = 1;
  a.taps = 1;
  a.cblocks = C1 / 64;
  a.kb2 = C2 / 64;
  a.kb2a = C2a / 64;
  a.stride2 = stride2;
  a.relu = L.relu;
  a.m_tiles = (M + BM - 1) / BM;
  a.n_tiles = L.Cout / L.wmap_bn;
  switch (L.wmap_bn) {
    case 256: return launch_tc<256, 128>(tmA, L.wmap, L.wmap2_ok ? &L.wmap2 : nullptr, tmY, tmY, tmA2, a, st, tmA3);
    case 128: return launch_tc<128, 128>(tmA, L.wmap, L.wmap2_ok ? &L.wmap2 : nullptr, tmY, tmY, tmA2, a, st, tmA3);
    default: return launch_tc<64, 128>(tmA, L.wmap, nullptr, tmY, tmY, tmA2, a, st, tmA3);
  }
}

bool conv_tc_pre_supported(const ConvLayer& L, int B, int H, int W, int C1, int C2) {
  return L.kh == 1 && L.kw == 1 && L.stride == 1 && L.pad == 0 && L.wmap_bn == 128 &&
         C1 % 64 == 0 && C2 % 64 == 0 && C1 + C2 == L.Cin && conv_tc_supported(L, B, H, W);
}

// 1x1 conv over relu(bn(x)) with x = [x1 (C1 channels) | x2 (C2 channels, may be null)]: the pre-activation of the
// consuming Residual (models/backbone/hourglass.py:60-61) is applied to the A tiles in shared memory (PRE kernels), so
// neither the pre-activated tensor nor the channel concat is ever written.
int launch_conv_tc_pre(const ConvLayer& L, const __nv_bfloat16* x1, int C1, const __nv_bfloat16* x2, int C2,
                       const float* pre_scale, const float* pre_shift, __nv_bfloat16* y, int B, int H, int W,
                       cudaStream_t st) {
  const Boxes bx = pick_boxes(H, W);
  CUtensorMap tmA, tmA2, tmY;
  const int M = B * H * W;
  if (!act_map_cached(x1, B, H, W, C1, 1, bx, L.name.c_str(), &tmA) || !rowmajor_map_cached(y, M, L.Cout, &tmY))
    return DIRB200_E_CUDA;
  if (!x2 || C2 == 0) tmA2 = tmA;
  else if (!act_map_cached(x2, B, H, W, C2, 1, bx, L.name.c_str(), &tmA2)) return DIRB200_E_CUDA;
  TcArgs a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.M = M;
  a.Cout = L.Cout;
  a.Ho = H;
  a.Wo = W;
  a.stride = 1;
  a.pad = 0;
  a.kw = 1;
  a.taps = 1;
  a.cblocks = L.Cin / 64;
  a.relu = L.relu;
  a.m_tiles = (M + BM - 1) / BM;
  a.n_tiles = L.Cout / L.wmap_bn;
  a.pre_scale = pre_scale;
  a.pre_shift = pre_shift;
  a.pre_cb1 = C1 / 64;
  const bool cg2 = L.wmap2_ok && cg2_enabled() && a.m_tiles % 2 == 0;
  return cg2 ? launch_tc_cg<128, 128, 2, 1>(tmA, L.wmap2, tmY, tmY, tmA2, tmA2, a, st)
             : launch_tc_cg<128, 128, 1, 1>(tmA, L.wmap, tmY, tmY, tmA2, tmA2, a, st);
}

int launch_conv_tc(const ConvLayer& L, const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* res, int B,
                   int H, int W, cudaStream_t st) {
  const int Ho = (H + 2 * L.pad - L.kh) / L.stride + 1, Wo = (W + 2 * L.pad - L.kw) / L.stride + 1;
  const Boxes bx = pick_boxes(Ho, Wo);
  CUtensorMap tmA, tmY, tmR;
  const int M = B * Ho * Wo;
  if (!act_map_cached(x, B, H, W, L.Cin, L.stride, bx, L.name.c_str(), &tmA) ||
      !rowmajor_map_cached(y, M, L.Cout, &tmY))
    return DIRB200_E_CUDA;
  if (!res) tmR = tmY;
  else if (!rowmajor_map_cached(res, M, L.Cout, &tmR)) return DIRB200_E_CUDA;
  TcArgs a{};
  a.scale = L.scale;
  a.shift = L.shift;
  a.M = M;
  a.Cout = L.Cout;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = L.stride;
  a.pad = L.pad;
  a.kw = L.kw;
  a.taps = L.kh * L.kw;
  a.cblocks = L.Cin / 64;
  a.relu = L.relu;
  a.has_res = res ? 1 : 0;
  a.stem = 0;
  a.m_tiles = (M + BM - 1) / BM;
  a.n_tiles = L.Cout / L.wmap_bn;
  a.raster_m = (double)L.Cout * L.K > (double)B * H * W * L.Cin ? 1 : 0;
  switch (L.wmap_bn) {
    case 256: return launch_tc<256, 128>(tmA, L.wmap, L.wmap2_ok ? &L.wmap2 : nullptr, tmY, tmR, tmA, a, st);
    case 128: return launch_tc<128, 128>(tmA, L.wmap, L.wmap2_ok ? &L.wmap2 : nullptr, tmY, tmR, tmA, a, st);
    default: return launch_tc<64, 128>(tmA, L.wmap, nullptr, tmY, tmR, tmA, a, st);
  }
}

}  // namespace dirb200
