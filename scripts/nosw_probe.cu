// Probe: tcgen05.mma kind::f16 with a NO-swizzle K-major A operand whose two K-chunks are `lbo` bytes apart (chunk-major
// halo planes of conv_halo.cu) and a 128B-swizzled B operand. A[r][k] = (k == kk) ? r + 1 : 0, B[n][k] = (k == kk) ? n + 1 : 0
// -> D[r][n] must be (r+1)(n+1). Prints the rows that come back for a few settings.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include "../dir_b200/csrc/tc_common.cuh"
using namespace dirb200::tc;

__global__ void __launch_bounds__(128, 1) probe(int lbo, int kk, int start_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;            // [64 n][128 B] swizzle-128
  uint8_t* sA = smem + 8192;     // planes: chunk c at c*lbo, row r at r*16
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (8192 + 2 * lbo + 4096 + 4096) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {  // A rows: this thread's row r
    const int r = threadIdx.x;
    const int c = kk / 8, e = kk % 8;
    reinterpret_cast<__nv_bfloat16*>(sA + start_off + c * lbo + r * 16)[e] = __float2bfloat16_rn((float)(r + 1));
    if (r < 64) {
      const int chunk = (kk * 2) / 16, off = (kk * 2) % 16;
      *reinterpret_cast<__nv_bfloat16*>(sB + r * 128 + ((chunk ^ (r & 7)) << 4) + off) = __float2bfloat16_rn((float)(r + 1));
    }
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    umma(tmem, desc_nosw(s32(sA + start_off), lbo, 128), desc128(s32(sB)), idesc(64, 1u), 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x * 2] = v[0];
  out[threadIdx.x * 2 + 1] = v[1];
  fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
  float* d; cudaMalloc(&d, 256 * 4);
  float h[256];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int lbos[] = {16, 128, 2048, 4096, 4240, 4224};
  for (int lbo : lbos)
    for (int kk : {0, 3, 8, 15})
      for (int so : {0, 16, 1040}) {
        probe<<<1, 128, 90 * 1024>>>(lbo, kk, so, d);
        cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int r = 0; r < 128; ++r)
          if (h[2 * r] != (float)(r + 1) || h[2 * r + 1] != 2.f * (r + 1)) { if (first < 0) first = r; ++bad; }
        printf("lbo %5d kk %2d start+%4d: %s  bad rows %3d (first %d: got %.0f,%.0f)  row0=%.0f row1=%.0f row8=%.0f row127=%.0f %s\n", lbo, kk, so,
               bad ? "WRONG" : "ok   ", bad, first, first >= 0 ? h[2 * first] : 0.f, first >= 0 ? h[2 * first + 1] : 0.f, h[0], h[2], h[16],
               h[254], e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
