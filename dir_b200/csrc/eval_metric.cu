// Next-row N2: the evaluation metric of apps/eval.py:151-241 on the device, straight from the packed
// output record: J-regress joints from vertices (class Jr, :22-44), wrist-root alignment, bone-length scale
// alignment (|j9 - j0| ratio), per-joint / per-vertex L2 errors, pinhole re-projection errors (xyz2uvd :80-83) and
// the relative-root error. One CTA per (image, hand); no host round trip per batch.
#include "../../include/dirb200.h"
#include "common.cuh"
#include "kernels.h"

namespace dirb200 {

namespace {

constexpr int NV = 778;

__device__ __forceinline__ void project(const float* cam, float x, float y, float z, float& u, float& v) {
  const float p0 = x * cam[0] + y * cam[1] + z * cam[2];
  const float p1 = x * cam[3] + y * cam[4] + z * cam[5];
  const float p2 = x * cam[6] + y * cam[7] + z * cam[8];
  u = p0 / p2;
  v = p1 / p2;
}

__global__ void __launch_bounds__(256) eval_metric_kernel(const float* __restrict__ record, const float* __restrict__ gt_verts,
                                                          const float* __restrict__ gt_verts2d,
                                                          const float* __restrict__ cam_all,
                                                          const float* __restrict__ jreg21, int use_scale,
                                                          float* __restrict__ joint_err, float* __restrict__ vert_err,
                                                          float* __restrict__ joint2d_err,
                                                          float* __restrict__ vert2d_err, float* __restrict__ root_err) {
  __shared__ float vg[NV * 3], vp[NV * 3];
  __shared__ float jg[21][3], jp[21][3], other_root[3], cam[9], sc_s;
  const int b = blockIdx.x, hand = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* rec = record + (size_t)b * DIRB200_RECORD_FLOATS + 2 * DIRB200_STAGE_FLOATS;  // last stage (eval.py:170-172)
  const float* pv = rec + (hand ? DIRB200_OFF_MESH_R : DIRB200_OFF_MESH_L);
  const float* gv = gt_verts + ((size_t)b * 2 + hand) * NV * 3;
  for (int i = tid; i < NV * 3; i += 256) {
    vg[i] = gv[i];
    vp[i] = pv[i];
  }
  if (tid < 9) cam[tid] = cam_all[(size_t)b * 9 + tid];
  __syncthreads();
  const float* J = jreg21 + (size_t)hand * 21 * NV;
  // 126 regressed coordinates (+3 for the other hand's GT wrist, needed by the relative-root error in the hand-0 CTA)
  const int nout = 126 + (hand == 0 ? 3 : 0);
  for (int o = warp; o < nout; o += 8) {
    float a = 0.f;
    if (o < 126) {
      const int which = o / 63, j = (o % 63) / 3, c = o % 3;
      const float* src = which ? vp : vg;
      for (int v = lane; v < NV; v += 32) a = fmaf(__ldg(J + j * NV + v), src[v * 3 + c], a);
    } else {
      const int c = o - 126;
      const float* Jr0 = jreg21 + (size_t)21 * NV;  // right hand, joint 0
      const float* gr = gt_verts + ((size_t)b * 2 + 1) * NV * 3;
      for (int v = lane; v < NV; v += 32) a = fmaf(__ldg(Jr0 + v), __ldg(gr + v * 3 + c), a);
    }
    a = warp_sum(a);
    if (lane == 0) {
      if (o < 63) jg[o / 3][o % 3] = a;
      else if (o < 126) jp[(o - 63) / 3][o % 3] = a;
      else other_root[o - 126] = a;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float dg[3], dp[3];
    for (int c = 0; c < 3; ++c) {
      dg[c] = jg[9][c] - jg[0][c];
      dp[c] = jp[9][c] - jp[0][c];
    }
    const float lg = sqrtf(dg[0] * dg[0] + dg[1] * dg[1] + dg[2] * dg[2]);
    const float lp = sqrtf(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    sc_s = use_scale ? lg / lp : 1.f;
    if (hand == 0) {  // root error (eval.py:156,231-233): |(root_right_gt - root_left_gt) - pd_offset * 0.15|
      const float* off = rec + DIRB200_OFF_OFFSET;
      float e = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float d = (other_root[c] - jg[0][c]) - off[c] * 0.15f;
        e += d * d;
      }
      root_err[b] = sqrtf(e);
    }
  }
  __syncthreads();
  const float sc = sc_s;
  const float rgx = jg[0][0], rgy = jg[0][1], rgz = jg[0][2];
  const float rpx = jp[0][0], rpy = jp[0][1], rpz = jp[0][2];
  const size_t base = (size_t)b * 2 + hand;
  if (tid < 21) {
    const float gx = jg[tid][0] - rgx, gy = jg[tid][1] - rgy, gz = jg[tid][2] - rgz;
    const float px = (jp[tid][0] - rpx) * sc, py = (jp[tid][1] - rpy) * sc, pz = (jp[tid][2] - rpz) * sc;
    joint_err[base * 21 + tid] = sqrtf((px - gx) * (px - gx) + (py - gy) * (py - gy) + (pz - gz) * (pz - gz));
    float u0, v0, u1, v1;
    project(cam, jg[tid][0], jg[tid][1], jg[tid][2], u0, v0);
    project(cam, px + rgx, py + rgy, pz + rgz, u1, v1);
    joint2d_err[base * 21 + tid] = sqrtf((u1 - u0) * (u1 - u0) + (v1 - v0) * (v1 - v0));
  }
  const float* g2 = gt_verts2d + base * NV * 2;
  for (int v = tid; v < NV; v += 256) {
    const float gx = vg[v * 3] - rgx, gy = vg[v * 3 + 1] - rgy, gz = vg[v * 3 + 2] - rgz;
    const float px = (vp[v * 3] - rpx) * sc, py = (vp[v * 3 + 1] - rpy) * sc, pz = (vp[v * 3 + 2] - rpz) * sc;
    vert_err[base * NV + v] = sqrtf((px - gx) * (px - gx) + (py - gy) * (py - gy) + (pz - gz) * (pz - gz));
    float u1, v1;
    project(cam, px + rgx, py + rgy, pz + rgz, u1, v1);
    const float du = u1 - g2[v * 2], dv = v1 - g2[v * 2 + 1];
    vert2d_err[base * NV + v] = sqrtf(du * du + dv * dv);
  }
}

}  // namespace

void launch_eval_metric(const float* record, const float* gt_verts, const float* gt_verts2d, const float* cam,
                        const float* jreg21, int B, int use_scale, float* joint_err, float* vert_err,
                        float* joint2d_err, float* vert2d_err, float* root_err, cudaStream_t st) {
  eval_metric_kernel<<<dim3(B, 2), 256, 0, st>>>(record, gt_verts, gt_verts2d, cam, jreg21, use_scale, joint_err, vert_err,
                                                 joint2d_err, vert2d_err, root_err);
}

}  // namespace dirb200
