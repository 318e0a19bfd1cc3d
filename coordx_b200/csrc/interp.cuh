// Feature-interpolated per-point fusion: the A-operand producer of the rendering path (ray_rendering.py:133-145,
// get_pts_pred with split=True).  For every query point p with fractional grid indices idx[p, c] along each axis c:
//     f_c[p, :] = (1 - r_c) * feat_c[floor(idx_c), :] + r_c * feat_c[floor(idx_c) + 1, :],   r_c = idx_c - floor(idx_c)
//     h[p, :]   = fuse_c f_c[p, :]                                   (sum | prod, modules.py:336-341)
// The reference materialises two gathered [C, P, Hd] tensors, their lerp and the stacked result before the fusion
// (about ten passes over [P, Hd] fp32); here the per-axis tables stay L2-resident and only h [P, Hd] is written,
// 16 bytes per thread.  h then goes through the trunk as a 1 x P grid (coordx_trunk_forward).
#pragma once
#include "common.cuh"

namespace coordx {

struct InterpArgs {
  int C, F, op;
  const float* feat[COORDX_MAX_BRANCHES];   // [n_c, F]
  int64_t n[COORDX_MAX_BRANCHES];
  const float* idx;                         // [P, C] fractional indices
  int64_t P;
  float* h;                                 // [P, F]
};

__global__ void __launch_bounds__(256) interp_fuse_kernel(const InterpArgs a) {
  const int f4 = a.F / 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.P * f4) return;
  const int64_t p = t / f4;
  const int k = (int)(t - p * f4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < a.C; ++c) {
    const float x = a.idx[p * a.C + c];
    int64_t lo = (int64_t)x;                              // .long() of the reference: truncation
    const float r = x - (float)lo;
    lo = min(max(lo, (int64_t)0), a.n[c] - 1);
    const int64_t hi = min(lo + 1, a.n[c] - 1);           // (the reference indexes lo + 1 unguarded: callers keep idx < n_c - 1)
    const float4 u = __ldg(reinterpret_cast<const float4*>(a.feat[c] + lo * a.F + k));
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.feat[c] + hi * a.F + k));
    const float w0 = 1.f - r;
    const float4 f = make_float4(u.x * w0 + r * v.x, u.y * w0 + r * v.y, u.z * w0 + r * v.z, u.w * w0 + r * v.w);
    if (c == 0) {
      acc = f;
    } else if (a.op == COORDX_FUSE_SUM) {
      acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
    } else {
      acc.x *= f.x; acc.y *= f.y; acc.z *= f.z; acc.w *= f.w;
    }
  }
  *reinterpret_cast<float4*>(a.h + p * a.F + k) = acc;
}

}  // namespace coordx
