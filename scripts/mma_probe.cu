// Hardware probe (sm_100a): numerical semantics and issue rate of tcgen05.mma kind::tf32 / kind::f16.
// Answers the questions the fp32-parity convolution (conv_tf32.cu) is designed around:
//   * how the fp32 accumulator in TMEM is rounded when an MMA adds to it (nearest-even or toward zero),
//   * how the products of ONE instruction are aligned/added before they reach the accumulator,
//   * whether kind::tf32 truncates or rounds the low 13 mantissa bits of its fp32 operands,
//   * MMA flop/clk/SM for tf32 and bf16 (M=128, N=256), issued back to back by one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_probe scripts/mma_probe.cu
// Output: one line per experiment; profiles/mma_probe_r2.txt holds the run this round's design used.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../dir_b200/csrc/tc_common.cuh"

using namespace dirb200::tc;

struct Phase {
  float a[16];
  float b[16];
  int repeat;      // number of identical MMAs issued
  int accumulate;  // 0: the first MMA of the phase overwrites D
};

constexpr int MAXP = 8;
struct Program {
  Phase ph[MAXP];
  int nph;
  int bf16;  // 1: kind::f16 with bf16 operands (K=16), 0: kind::tf32 (K=8)
};

// one operand row = 128 bytes, 128B swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
__device__ void fill_row(uint8_t* tile, int row, const float* vals, int bf16) {
  uint8_t* base = tile + row * 128;
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(base + c * 16) = make_uint4(0, 0, 0, 0);
  if (bf16) {
    for (int k = 0; k < 16; ++k) {
      const int chunk = (k * 2) / 16, off = (k * 2) % 16;
      *reinterpret_cast<__nv_bfloat16*>(base + ((chunk ^ (row & 7)) << 4) + off) = __float2bfloat16_rn(vals[k]);
    }
  } else {
    for (int k = 0; k < 8; ++k) {
      const int chunk = (k * 4) / 16, off = (k * 4) % 16;
      *reinterpret_cast<float*>(base + ((chunk ^ (row & 7)) << 4) + off) = vals[k];
    }
  }
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const Program prog, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 128 rows x 128 B
  uint8_t* sB = smem + 128 * 128;  // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t id = idesc(64, prog.bf16 ? 1u : 2u);
  uint32_t parity = 0;
  for (int p = 0; p < prog.nph; ++p) {
    fill_row(sA, threadIdx.x, prog.ph[p].a, prog.bf16);
    if (threadIdx.x < 64) fill_row(sB, threadIdx.x, prog.ph[p].b, prog.bf16);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      fence_after();
      const uint64_t da = desc128(s32(sA)), db = desc128(s32(sB));
      for (int r = 0; r < prog.ph[p].repeat; ++r) {
        const uint32_t acc = (r > 0 || prog.ph[p].accumulate) ? 1u : 0u;
        if (prog.bf16) umma(tmem, da, db, id, acc);
        else umma_tf32(tmem, da, db, id, acc);
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    fence_after();
    __syncthreads();
  }
  if (warp == 0) {
    float v[32];
    tmem_ld32(tmem, v);
    tmem_ld_wait();
    if (threadIdx.x == 0) out[0] = v[0];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

// issue rate: `n` MMAs (M=128, N=256, one 32-byte k-step) back to back into one accumulator
__global__ void __launch_bounds__(128, 1) rate_kernel(int bf16, int n, long long* cycles, int N = 256) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;  // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t id = idesc(N, bf16 ? 1u : 2u);
    const uint64_t da = desc128(s32(sA)), db = desc128(s32(sB));
    const long long t0 = clock64();
    for (int r = 0; r < n; ++r) {
      const uint64_t k = (uint64_t)(2 * (r & 3));  // walk the four k-steps of the 128-byte atom like a real mainloop
      if (bf16) umma(tmem, da + k, db + k, id, r ? 1u : 0u);
      else umma_tf32(tmem, da + k, db + k, id, r ? 1u : 0u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static float run(const Program& p, float* dout) {
  probe_kernel<<<1, 128, 128 * 128 + 64 * 128 + 1024>>>(p, dout);
  float h = 0;
  cudaError_t e = cudaMemcpy(&h, dout, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return h;
}

static Phase phase(std::vector<float> a, std::vector<float> b, int repeat, int acc) {
  Phase ph;
  memset(&ph, 0, sizeof ph);
  for (size_t i = 0; i < a.size(); ++i) ph.a[i] = a[i];
  for (size_t i = 0; i < b.size(); ++i) ph.b[i] = b[i];
  ph.repeat = repeat;
  ph.accumulate = acc;
  return ph;
}

int main() {
  float* dout;
  cudaMalloc(&dout, 64);
  const float two24 = 16777216.f, two12 = 4096.f;
  for (int bf16 = 0; bf16 < 2; ++bf16) {
    const char* kind = bf16 ? "kind::f16(bf16)" : "kind::tf32";
    Program p;
    memset(&p, 0, sizeof p);
    p.bf16 = bf16;
    // T1: accumulator = 2^24, then 16 MMAs adding exactly 3 each (ulp = 2): nearest -> +64, toward zero -> +32
    p.nph = 2;
    p.ph[0] = phase({two12}, {two12}, 1, 0);
    p.ph[1] = phase({3.f}, {1.f}, 16, 1);
    float v = run(p, dout);
    printf("%s T1 acc=2^24 then 16 x (+3):      D - 2^24 = %.1f   (exact 48; round-nearest 64; toward-zero 32)\n", kind,
           (double)v - (double)two24);
    // T1b: adding exactly 1 (half ulp, tie): nearest-even keeps 2^24, round-up/away would grow
    p.ph[1] = phase({1.f}, {1.f}, 16, 1);
    v = run(p, dout);
    printf("%s T1b acc=2^24 then 16 x (+1):     D - 2^24 = %.1f   (exact 16; nearest-even 0; toward-zero 0)\n", kind,
           (double)v - (double)two24);
    // T2: one instruction, products {2^24, 1 x 7}: exact 2^24+7
    p.nph = 1;
    p.ph[0] = phase({two12, 1, 1, 1, 1, 1, 1, 1}, {two12, 1, 1, 1, 1, 1, 1, 1}, 1, 0);
    v = run(p, dout);
    printf("%s T2 one MMA {2^24,1,1,1,1,1,1,1}:  D - 2^24 = %.1f   (exact 7; nearest 8; toward-zero 6; per-product "
           "truncation at ulp(max) 0)\n", kind, (double)v - (double)two24);
    // T3: one instruction, products {2^24, 0.75 x 4}: exact 2^24+3 (each addend < half ulp)
    p.ph[0] = phase({two12, 0.75f, 0.75f, 0.75f, 0.75f}, {two12, 1, 1, 1, 1}, 1, 0);
    v = run(p, dout);
    printf("%s T3 one MMA {2^24,.75,.75,.75,.75}: D - 2^24 = %.1f  (exact 3; sum-then-round-nearest 4; toward-zero 2)\n",
           kind, (double)v - (double)two24);
    // T3b: guard bits of the product alignment: {2^24, 2^-k x 1}, accumulate many times is not needed: add 2^24 * (1 + ...)
    for (int g = 1; g <= 6; ++g) {
      const float small = ldexpf(1.f, -g);  // 0.5, 0.25, ...
      std::vector<float> a = {two12}, b = {two12};
      const int cnt = bf16 ? 15 : 7;
      for (int i = 0; i < cnt; ++i) {
        a.push_back(small);
        b.push_back(1.f);
      }
      // second phase adds 2^24 more so the sum of the small terms (cnt * small) matters only if they survived alignment
      p.nph = 1;
      p.ph[0] = phase(a, b, 1, 0);
      v = run(p, dout);
      printf("%s T3b one MMA {2^24, %d x 2^-%d}:    D - 2^24 = %.1f   (exact %.4f)\n", kind, cnt, g,
             (double)v - (double)two24, cnt * (double)small);
    }
    if (!bf16) {
      // T4: tf32 operand conversion: a = 1 + 2^-11 + 2^-12 (between tf32 values 1 and 1+2^-10), b = 2^10
      p.nph = 1;
      p.ph[0] = phase({1.f + ldexpf(1.f, -11) + ldexpf(1.f, -12)}, {1024.f}, 1, 0);
      v = run(p, dout);
      printf("%s T4 a = 1+2^-11+2^-12, b = 1024:    D = %.4f   (truncate 1024; round-nearest 1025; full fp32 1024.75)\n",
             kind, v);
      p.ph[0] = phase({1.f + ldexpf(1.f, -11)}, {1024.f}, 1, 0);
      v = run(p, dout);
      printf("%s T4b a = 1+2^-11 (tie), b = 1024:    D = %.4f   (truncate/nearest-even 1024; round-half-away 1025)\n",
             kind, v);
      p.ph[0] = phase({-(1.f + ldexpf(1.f, -11) + ldexpf(1.f, -12))}, {1024.f}, 1, 0);
      v = run(p, dout);
      printf("%s T4c a = -(1+2^-11+2^-12), b = 1024: D = %.4f\n", kind, v);
    }
    // T5: products of one instruction are exact before summation? a*b with 2 x 11-bit (tf32) / 8-bit (bf16) mantissas
    {
      const float a = bf16 ? 1.f + ldexpf(1.f, -7) : 1.f + ldexpf(1.f, -10);
      p.nph = 1;
      p.ph[0] = phase({a}, {a}, 1, 0);
      v = run(p, dout);
      printf("%s T5 a*a, a = 1+2^-%d:               D - 1 = %.10g   (exact %.10g)\n", kind, bf16 ? 7 : 10, (double)v - 1.0,
             (double)a * a - 1.0);
    }
    // T6: long accumulation of a constant: sum of n x c with c = 1 + 2^-10 (exactly representable in both kinds? tf32 yes,
    // bf16 no -> use 1 + 2^-7): compares against fp32 sequential round-nearest and toward-zero models on the host
    {
      const float c = bf16 ? 1.f + ldexpf(1.f, -7) : 1.f + ldexpf(1.f, -10);
      const int kper = bf16 ? 16 : 8;
      std::vector<float> a(kper, c), b(kper, c);
      const int reps = 4096;
      p.nph = 1;
      p.ph[0] = phase(a, b, reps, 0);
      v = run(p, dout);
      const double prod = (double)c * c, step = prod * kper;
      float rn = 0.f, rz = 0.f;
      for (int i = 0; i < reps; ++i) {
        rn = (float)((double)rn + step);  // double add then round-to-nearest fp32
        double t = (double)rz + step;
        float f = (float)t;
        if ((double)f > t) f = nextafterf(f, 0.f);
        rz = f;
      }
      printf("%s T6 %d x MMA of %d x %.10g:  D = %.2f   exact %.2f   fp32-RN-per-MMA %.2f   fp32-RZ-per-MMA %.2f\n", kind, reps,
             kper, prod, v, step * reps, rn, rz);
    }
  }
  // issue rate
  long long* dc;
  cudaMalloc(&dc, 8 * 148);
  for (int bf16 = 0; bf16 < 2; ++bf16) {
    for (int n : {256, 4096}) {
      cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      rate_kernel<<<1, 128, (128 + 256) * 128 + 1024>>>(bf16, n, dc);
      long long h = 0;
      cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      const double flop = 2.0 * 128 * 256 * (bf16 ? 16 : 8) * n;
      printf("rate %s: %d MMAs (M128 N256) in %lld clk = %.1f clk/MMA = %.0f flop/clk/SM (x148 SMs x 1.965 GHz = %.0f "
             "TFLOP/s)\n", bf16 ? "bf16" : "tf32", n, h, (double)h / n, flop / h, flop / h * 148 * 1.965e9 / 1e12);
    }
  }
  for (int N : {64, 128, 256}) {  // does the instruction time scale with N? (conv_tc<64,...>, conv_halo issue N = 64 MMAs)
    rate_kernel<<<1, 128, (128 + 256) * 128 + 1024>>>(1, 4096, dc, N);
    long long h = 0;
    cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
    printf("rate bf16 N=%d: %.1f clk/MMA (M128 K16) = %.0f flop/clk/SM\n", N, (double)h / 4096, 2.0 * 128 * N * 16 * 4096 / h);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("final status: %s\n", cudaGetErrorString(e));
  return 0;
}
