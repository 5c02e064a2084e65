// Regression heads + manopth MANO layer + orthographic projection, one CTA per (image, hand).
//   Linear heads          : models/dir.py:268-270 (init) and :339-351 (RegressorOffset)
//   proj_feat_emb         : models/dir.py:118-119
//   ManoLayer.forward     : manopth/manopth/manolayer.py:110-270 (6D robust root, 45 PCA comps,
//                           quaternion Rodrigues, 3-level FK, LBS, tips, reorder, centre on joint 0)
//   projection_batch_xy   : utils/utils.py:47-63
// Replaces ~320 ATen launches (and the per-sample torch.det host sync, manopth/manopth/rot6d.py:50) per call.
#include "../../include/dirb200.h"
#include "common.cuh"
#include "kernels.h"

namespace dirb200 {

namespace {

constexpr int NV = 778;
constexpr int NV3 = 2334;
constexpr int THREADS = 256;
constexpr int VMAX = 2752;  // >= 2691 (offset head input), >= 2048 (init head input)

struct ManoSmem {
  float para[64];
  float aa[48];
  float R[16][9];      // [0] = root, [1..15] = joint rotations
  float pose_map[136];
  float J[16][3];
  float GR[16][9];
  float Gt[16][3];
  float At[16][3];
  float tips[5][3];
  float vs[NV3];       // v_shaped, then v_posed
};

__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
  float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-8f);
  x /= n; y /= n; z /= n;
}

__device__ __forceinline__ void matmul3(const float* A, const float* B, float* C) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// The whole MANO layer for one (image, hand). s.para must be filled; all threads of the CTA participate.
__device__ void mano_forward(ManoSmem& s, const ManoWeights& w, float* __restrict__ rec /*stage record of image*/,
                             int hand) {
  const int tid = threadIdx.x;
  const float* pose = s.para;         // [0:6] 6D root, [6:51] PCA coeffs
  const float* beta = s.para + 51;    // 10
  const float* proj = s.para + 61;    // scale, tx, ty

  // 1. PCA -> axis-angle (manolayer.py:124-133)
  if (tid < 45) {
    float a = 0.f;
    for (int k = 0; k < 45; ++k) a = fmaf(pose[6 + k], __ldg(w.comps + k * 45 + tid), a);
    s.aa[tid] = __ldg(w.mean + tid) + a;
  }
  __syncthreads();
  // 2. rotations
  if (tid < 15) {  // rodrigues_layer.py:15-54
    float ax = s.aa[tid * 3], ay = s.aa[tid * 3 + 1], az = s.aa[tid * 3 + 2];
    float ex = ax + 1e-8f, ey = ay + 1e-8f, ez = az + 1e-8f;
    float theta = sqrtf(ex * ex + ey * ey + ez * ez);
    float nx = ax / theta, ny = ay / theta, nz = az / theta;
    float half = theta * 0.5f;
    float sn, cs;
    sincosf(half, &sn, &cs);
    float qw = cs, qx = sn * nx, qy = sn * ny, qz = sn * nz;
    float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= qn; qx /= qn; qy /= qn; qz /= qn;
    float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    float wx = qw * qx, wy = qw * qy, wz = qw * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
    float* R = s.R[tid + 1];
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
#pragma unroll
    for (int k = 0; k < 9; ++k) s.pose_map[tid * 9 + k] = R[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
  } else if (tid == 32) {  // rot6d.py:26-51 (robust), columns x,y,z
    float x0 = pose[0], x1 = pose[1], x2 = pose[2], y0 = pose[3], y1 = pose[4], y2 = pose[5];
    normalize3(x0, x1, x2);
    normalize3(y0, y1, y2);
    float m0 = x0 + y0, m1 = x1 + y1, m2 = x2 + y2;
    normalize3(m0, m1, m2);
    float o0 = x0 - y0, o1 = x1 - y1, o2 = x2 - y2;
    normalize3(o0, o1, o2);
    x0 = m0 + o0; x1 = m1 + o1; x2 = m2 + o2;
    normalize3(x0, x1, x2);
    y0 = m0 - o0; y1 = m1 - o1; y2 = m2 - o2;
    normalize3(y0, y1, y2);
    float z0 = x1 * y2 - x2 * y1, z1 = x2 * y0 - x0 * y2, z2 = x0 * y1 - x1 * y0;
    normalize3(z0, z1, z2);
    float* R = s.R[0];
    R[0] = x0; R[1] = y0; R[2] = z0;
    R[3] = x1; R[4] = y1; R[5] = z1;
    R[6] = x2; R[7] = y2; R[8] = z2;
  }
  // 3. v_shaped = template + shapedirs . beta (manolayer.py:173-176)
  for (int i = tid; i < NV3; i += THREADS) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) v = fmaf(__ldg(w.shapedirs_t + k * NV3 + i), beta[k], v);
    s.vs[i] = v + __ldg(w.v_template + i);
  }
  __syncthreads();
  // 4. J = J_regressor . v_shaped (manolayer.py:177): 48 dot products of length 778, one warp each
  {
    // warp w owns joints 2w, 2w+1 (all three coordinates): six concurrent dot products, each regressor weight
    // loaded once for its three coordinates (per-output sums unchanged)
    const int warp = tid >> 5, lane = tid & 31;
    const int ja = 2 * warp, jb = 2 * warp + 1;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (int v = lane; v < NV; v += 32) {
      const float wa = __ldg(w.jreg + ja * NV + v), wb = __ldg(w.jreg + jb * NV + v);
      const float x = s.vs[v * 3], y = s.vs[v * 3 + 1], z = s.vs[v * 3 + 2];
      a0 = fmaf(wa, x, a0); a1 = fmaf(wa, y, a1); a2 = fmaf(wa, z, a2);
      b0 = fmaf(wb, x, b0); b1 = fmaf(wb, y, b1); b2 = fmaf(wb, z, b2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    b0 = warp_sum(b0); b1 = warp_sum(b1); b2 = warp_sum(b2);
    if (lane == 0) {
      s.J[ja][0] = a0; s.J[ja][1] = a1; s.J[ja][2] = a2;
      s.J[jb][0] = b0; s.J[jb][1] = b1; s.J[jb][2] = b2;
    }
  }
  __syncthreads();
  // 5. v_posed = v_shaped + posedirs . pose_map (manolayer.py:180-181); each thread owns elements tid + 256*e of vs.
  //    The 1.26 MB posedirs matrix streams from L2: 10 independent accumulators x 5 k-steps = 50 loads in flight.
  {
    constexpr int NE = (NV3 + THREADS - 1) / THREADS;  // 10
    float pv[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) pv[e] = 0.f;
    const float* pd = w.posedirs_t + tid;
    for (int k = 0; k < 135; k += 5) {  // 135 = 27 x 5: 50 independent loads in flight per thread
      float m[5], l[5][NE];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = s.pose_map[k + q];
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int e = 0; e < NE; ++e)
          l[q][e] = (tid + e * THREADS < NV3) ? __ldg(pd + (size_t)(k + q) * NV3 + e * THREADS) : 0.f;
#pragma unroll
      for (int q = 0; q < 5; ++q)  // ascending k: the per-element summation order is unchanged
#pragma unroll
        for (int e = 0; e < NE; ++e) pv[e] = fmaf(l[q][e], m[q], pv[e]);
    }
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (tid + e * THREADS < NV3) s.vs[tid + e * THREADS] += pv[e];
  }
  // 6. forward kinematics: root, then one thread per finger chain (manolayer.py:186-227)
  if (tid < 5) {
    const float* PR = s.R[0];
    float pt[3] = {s.J[0][0], s.J[0][1], s.J[0][2]};
    float PRl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) PRl[k] = PR[k];
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) s.GR[0][k] = PRl[k];
      s.Gt[0][0] = pt[0]; s.Gt[0][1] = pt[1]; s.Gt[0][2] = pt[2];
    }
    int par = 0;
    for (int l = 0; l < 3; ++l) {
      int j = 1 + tid * 3 + l;
      float rel[3] = {s.J[j][0] - s.J[par][0], s.J[j][1] - s.J[par][1], s.J[j][2] - s.J[par][2]};
      float nt[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) nt[r] = PRl[r * 3] * rel[0] + PRl[r * 3 + 1] * rel[1] + PRl[r * 3 + 2] * rel[2] + pt[r];
      float NR[9];
      matmul3(PRl, s.R[j], NR);
#pragma unroll
      for (int k = 0; k < 9; ++k) { s.GR[j][k] = NR[k]; PRl[k] = NR[k]; }
#pragma unroll
      for (int r = 0; r < 3; ++r) { s.Gt[j][r] = nt[r]; pt[r] = nt[r]; }
      par = j;
    }
  }
  __syncthreads();
  // 7. A_j = G_j - [0 | G_R J_j] (manolayer.py:229-231)
  if (tid < 48) {
    int j = tid / 3, r = tid % 3;
    s.At[j][r] = s.Gt[j][r] - (s.GR[j][r * 3] * s.J[j][0] + s.GR[j][r * 3 + 1] * s.J[j][1] + s.GR[j][r * 3 + 2] * s.J[j][2]);
  }
  __syncthreads();
  // 8. LBS, centred on joint 0 (= Gt[0] after the reorder), + write mesh (manolayer.py:233-244,261-265)
  const float cx = s.Gt[0][0], cy = s.Gt[0][1], cz = s.Gt[0][2];
  float* mesh = rec + (hand ? DIRB200_OFF_MESH_R : DIRB200_OFF_MESH_L);
  for (int v = tid; v < NV; v += THREADS) {
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = 0.f;
    const float4* wp = reinterpret_cast<const float4*>(w.skin_w + v * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 w4 = __ldg(wp + q);
      float ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int j = q * 4 + e;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          T[r * 4 + 0] = fmaf(ww[e], s.GR[j][r * 3 + 0], T[r * 4 + 0]);
          T[r * 4 + 1] = fmaf(ww[e], s.GR[j][r * 3 + 1], T[r * 4 + 1]);
          T[r * 4 + 2] = fmaf(ww[e], s.GR[j][r * 3 + 2], T[r * 4 + 2]);
          T[r * 4 + 3] = fmaf(ww[e], s.At[j][r], T[r * 4 + 3]);
        }
      }
    }
    float px = s.vs[v * 3], py = s.vs[v * 3 + 1], pz = s.vs[v * 3 + 2];
    float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
    float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
    float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
    int tip = -1;
    if (v == 745) tip = 0;
    else if (v == 317) tip = 1;
    else if (v == w.tip2) tip = 2;
    else if (v == 556) tip = 3;
    else if (v == 673) tip = 4;
    if (tip >= 0) { s.tips[tip][0] = ox; s.tips[tip][1] = oy; s.tips[tip][2] = oz; }
    mesh[v * 3 + 0] = ox - cx;
    mesh[v * 3 + 1] = oy - cy;
    mesh[v * 3 + 2] = oz - cz;
  }
  __syncthreads();
  // 9. joints: [16 FK joints | 5 tips] reordered (manolayer.py:253-259), centred, projected (utils.py:47-63)
  if (tid < 21) {
    const int reorder[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
    int src = reorder[tid];
    float jx, jy, jz;
    if (src < 16) { jx = s.Gt[src][0]; jy = s.Gt[src][1]; jz = s.Gt[src][2]; }
    else { jx = s.tips[src - 16][0]; jy = s.tips[src - 16][1]; jz = s.tips[src - 16][2]; }
    jx -= cx; jy -= cy; jz -= cz;
    float* jo = rec + (hand ? DIRB200_OFF_JOINT_R : DIRB200_OFF_JOINT_L) + tid * 3;
    jo[0] = jx; jo[1] = jy; jo[2] = jz;
    float* uvo = rec + (hand ? DIRB200_OFF_UV_R : DIRB200_OFF_UV_L) + tid * 2;
    uvo[0] = proj[0] * jx + proj[1];
    uvo[1] = proj[0] * jy + proj[2];
  }
  if (tid >= 32 && tid < 35) rec[(hand ? DIRB200_OFF_PROJ_R : DIRB200_OFF_PROJ_L) + tid - 32] = proj[tid - 32];
}

__device__ __forceinline__ void load_vec(float* dst, const VecSeg& a, const VecSeg& b, int img) {
  for (int i = threadIdx.x; i < a.n; i += THREADS) dst[i] = a.p[(int64_t)img * a.stride + i];
  if (b.p)
    for (int i = threadIdx.x; i < b.n; i += THREADS) dst[a.n + i] = b.p[(int64_t)img * b.stride + i];
}

// out[o] = W[o,:] . vec + bias[o], one warp per output, coalesced weight rows
__device__ __forceinline__ void warp_linear(const float* __restrict__ W, const float* __restrict__ bias,
                                            const float* vec, int K, int nout, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < nout; o += THREADS / 32) {
    const float* wr = W + (int64_t)o * K;
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), vec[k], a);
    a = warp_sum(a);
    if (lane == 0) out[o] = a + bias[o];
  }
}

// The 64-output MANO head: warp w owns outputs w, w+8, ..., w+56 and runs the eight dot products CONCURRENTLY (eight
// independent coalesced weight-row loads in flight per lane instead of one); per-output sums are unchanged.
__device__ __forceinline__ void warp_linear64(const float* __restrict__ W, const float* __restrict__ bias,
                                              const float* vec, int K, float* out) {
  constexpr int NW = THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = 0.f;
  const float* wr = W + (int64_t)warp * K;
  const int64_t ostride = (int64_t)NW * K;
  for (int k = lane; k < K; k += 32) {
    const float x = vec[k];
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __ldg(wr + u * ostride + k);
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = fmaf(wv[u], x, a[u]);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float r = warp_sum(a[u]);
    if (lane == 0) out[warp + u * NW] = r + bias[warp + u * NW];
  }
}

__global__ void __launch_bounds__(THREADS) regress_mano_kernel(RegressArgs a) {
  pdl_wait();
  __shared__ ManoSmem s;
  __shared__ __align__(16) float vin[VMAX];
  __shared__ float hid[21 * 64];
  const int b = blockIdx.x, hand = blockIdx.y, tid = threadIdx.x;
  float* rec = a.stage_record + (int64_t)b * a.rec_stride;

  load_vec(vin, a.in0[hand], a.in1[hand], b);
  __syncthreads();
  const int K = a.in0[hand].n + (a.in1[hand].p ? a.in1[hand].n : 0);
  static_assert(THREADS == 256, "warp_linear64 assumes 8 warps x 8 outputs");
  warp_linear64(a.Wm[hand], a.bm[hand], vin, K, s.para);
  __syncthreads();
  if (tid < 64) a.mano_para[(int64_t)b * a.para_stride + hand * 64 + tid] = s.para[tid];

  if (a.do_proj_feat) {  // proj_feat_emb on the 21x64 token block (first 1344 floats of vin)
    const PointMlp& f = a.proj_feat;
    for (int item = tid; item < 21 * 64; item += THREADS) {
      int n = item & 63, j = item >> 6;
      float acc = 0.f;
      for (int k = 0; k < 64; ++k) acc = fmaf(vin[j * 64 + k], __ldg(f.w1t + k * 64 + n), acc);
      hid[item] = fmaxf(fmaf(acc, f.s1[n], f.b1[n]), 0.f);
    }
    __syncthreads();
    for (int item = tid; item < 21 * 64; item += THREADS) {
      int n = item & 63, j = item >> 6;
      float acc = 0.f;
      for (int k = 0; k < 64; ++k) acc = fmaf(hid[j * 64 + k], __ldg(f.w2t + k * 64 + n), acc);
      a.joint_feat[((int64_t)(b * 2 + hand) * 21 + j) * 64 + n] = acc + f.b2[n];
    }
  }
  __syncthreads();
  if (hand == 0) {  // offset head (block-uniform branch)
    load_vec(vin, a.off0, a.off1, b);
    __syncthreads();
    const int Ko = a.off0.n + (a.off1.p ? a.off1.n : 0);
    // three outputs over K ~ 2.7k: all 8 warps split K (one warp per output left 5 warps idle behind an 84-iteration
    // load chain); partial sums meet in smem (fp32, association differs from the sequential sum only)
    {
      const int warp = tid >> 5, lane = tid & 31;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      for (int k = tid; k < Ko; k += THREADS) {
        const float x = vin[k];
        o0 = fmaf(__ldg(a.Wo + k), x, o0);
        o1 = fmaf(__ldg(a.Wo + Ko + k), x, o1);
        o2 = fmaf(__ldg(a.Wo + 2 * (int64_t)Ko + k), x, o2);
      }
      o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2);
      if (lane == 0) { hid[warp * 3] = o0; hid[warp * 3 + 1] = o1; hid[warp * 3 + 2] = o2; }
      __syncthreads();
      if (tid < 3) {
        float r = 0.f;
#pragma unroll
        for (int wv = 0; wv < THREADS / 32; ++wv) r += hid[wv * 3 + tid];
        rec[DIRB200_OFF_OFFSET + tid] = r + a.bo[tid];
      }
    }
  }
  mano_forward(s, a.mano[hand], rec, hand);
}

__global__ void __launch_bounds__(THREADS) mano_only_kernel(const float* __restrict__ para, ManoWeights m0,
                                                            ManoWeights m1, float* __restrict__ stage_record,
                                                            int rec_stride) {
  __shared__ ManoSmem s;
  const int b = blockIdx.x, hand = blockIdx.y;
  if (threadIdx.x < 64) s.para[threadIdx.x] = para[((int64_t)b * 2 + hand) * 64 + threadIdx.x];
  __syncthreads();
  mano_forward(s, hand ? m1 : m0, stage_record + (int64_t)b * rec_stride, hand);
}

}  // namespace

void launch_regress_mano(const RegressArgs& a, cudaStream_t st) {
  launch_pdl(regress_mano_kernel, dim3(dim3(a.B, 2)), dim3(THREADS), 0, st, a);
}

void launch_mano_only(const float* para, const ManoWeights mano[2], float* stage_record, int rec_stride, int B,
                      cudaStream_t st) {
  mano_only_kernel<<<dim3(B, 2), THREADS, 0, st>>>(para, mano[0], mano[1], stage_record, rec_stride);
}

}  // namespace dirb200
