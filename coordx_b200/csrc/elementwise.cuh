// Fusion (implicit outer-product broadcast), loss and optimiser kernels of the FP32 path.
#pragma once
#include "common.cuh"

namespace coordx {

struct FuseArgs {
  int C;
  const float* feat[COORDX_MAX_BRANCHES];   // [n_c, F]
  int64_t n[COORDX_MAX_BRANCHES];
  int F, Hd, gs;                            // F = Hd * gs (group-sum of modules.py:349-351)
  int op, fba, act;
  float S;
  const float* code;                        // nullable [F]
  int64_t row0, nrows;                      // chunk = rows [row0, row0+nrows) of axis 0
  int64_t inner;                            // prod_{c>0} n_c
};

__device__ __forceinline__ void decode_point(const FuseArgs& a, int64_t q, int64_t* ic) {
  int64_t rem = q % a.inner;
  ic[0] = a.row0 + q / a.inner;
  for (int c = a.C - 1; c >= 1; --c) {
    ic[c] = rem % a.n[c];
    rem /= a.n[c];
  }
}

// H0[q, j] = sum_s code[jj] * maybe_act( fuse_c feat_c[i_c(q), jj] ),  jj = j*gs + s
// SplitFCBlock.forward_fusion, modules.py:336-351 - for one bounded chunk of whole axis-0 rows.
__global__ void __launch_bounds__(256) fuse_fwd_kernel(FuseArgs a, int64_t npts, float* __restrict__ H0) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npts * a.Hd) return;
  const int j = (int)(idx % a.Hd);
  const int64_t q = idx / a.Hd;
  int64_t ic[COORDX_MAX_BRANCHES];
  decode_point(a, q, ic);
  float out = 0.f;
  for (int s = 0; s < a.gs; ++s) {
    const int jj = j * a.gs + s;
    float v = a.feat[0][ic[0] * a.F + jj];
    for (int c = 1; c < a.C; ++c) {
      const float f = a.feat[c][ic[c] * a.F + jj];
      v = a.op == COORDX_FUSE_PROD ? v * f : v + f;
    }
    if (a.fba) v = act_fwd(a.act, a.S, v);
    if (a.code) v *= a.code[jj];
    out += v;
  }
  H0[idx] = out;
}

// the same for the plain case (group size 1, no activation before the fusion, no code, F a multiple of 4): four columns
// per thread - one index decode and 16-byte accesses instead of four of each (the scalar kernel moved 0.57 TB/s)
__global__ void __launch_bounds__(256) fuse_fwd_vec4_kernel(FuseArgs a, int64_t npts, float* __restrict__ H0) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over npts * Hd / 4
  const int hd4 = a.Hd >> 2;
  if (idx >= npts * hd4) return;
  const int j4 = (int)(idx % hd4);
  const int64_t q = idx / hd4;
  int64_t ic[COORDX_MAX_BRANCHES];
  decode_point(a, q, ic);
  float4 v = __ldg(reinterpret_cast<const float4*>(a.feat[0] + ic[0] * a.F) + j4);
  for (int c = 1; c < a.C; ++c) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(a.feat[c] + ic[c] * a.F) + j4);
    if (a.op == COORDX_FUSE_PROD) { v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w; }
    else { v.x += f.x; v.y += f.y; v.z += f.z; v.w += f.w; }
  }
  reinterpret_cast<float4*>(H0)[idx] = v;
}

static bool fuse_fwd_vec4_ok(const FuseArgs& a, const float* H0) {
  if (a.gs != 1 || a.fba || a.code || (a.F & 3) || (a.Hd & 3) || a.F != a.Hd) return false;
  uintptr_t al = reinterpret_cast<uintptr_t>(H0);
  for (int c = 0; c < a.C; ++c) al |= reinterpret_cast<uintptr_t>(a.feat[c]);
  return (al & 15) == 0;
}

// dfeat_b[i, jj] += sum over the chunk's points with index i on axis b of
//   dH0[q, jj/gs] * code[jj] * act'(fused) * prod_{c != b} feat_c      (prod)   | 1 (sum)
// also dcode[jj] (branch 0 pass only) when learn_code.
__global__ void __launch_bounds__(128)
fuse_bwd_kernel(FuseArgs a, int b, const float* __restrict__ dH0, float* __restrict__ dfeat, float* __restrict__ dcode,
                int64_t T, int64_t t_per_block) {
  const int jj = blockIdx.y * 128 + threadIdx.x;
  if (jj >= a.F) return;
  const int64_t ib_local = blockIdx.x;                              // index on axis b within the chunk
  const int64_t ib = (b == 0 ? a.row0 : 0) + ib_local;
  const int64_t t0 = (int64_t)blockIdx.z * t_per_block, t1 = min(T, t0 + t_per_block);
  const float fb = a.feat[b][ib * a.F + jj];
  const float cd = a.code ? a.code[jj] : 1.f;
  float acc = 0.f, acc_code = 0.f;
  if (a.C == 2 && a.gs == 1 && !a.fba && !a.code && !dcode) {
    // plain two-axis case: no index decode (the general loop below spends its time on 64-bit divisions: 0.6 TB/s)
    //   b = 0: points (ib, t), t = index on axis 1;   b = 1: points (t, ib), t = row of the chunk
    const float* dh = b == 0 ? dH0 + (ib_local * a.inner + t0) * a.Hd + jj : dH0 + (t0 * a.inner + ib_local) * a.Hd + jj;
    const int64_t dstep = b == 0 ? a.Hd : a.inner * a.Hd;
    const float* ot = b == 0 ? a.feat[1] + t0 * a.F + jj : a.feat[0] + (a.row0 + t0) * a.F + jj;
    if (a.op == COORDX_FUSE_PROD) {
      for (int64_t t = t0; t < t1; ++t, dh += dstep, ot += a.F) acc = fmaf(__ldg(dh), __ldg(ot), acc);
    } else {
      for (int64_t t = t0; t < t1; ++t, dh += dstep) acc += __ldg(dh);
    }
    atomicAdd(&dfeat[ib * a.F + jj], acc);
    return;
  }
  for (int64_t t = t0; t < t1; ++t) {
    // decode the other axes' indices from t (row-major over axes != b, axis 0 counts nrows)
    int64_t rem = t, q = 0, mul = 1;
    float other = a.op == COORDX_FUSE_PROD ? 1.f : 0.f;
    for (int c = a.C - 1; c >= 0; --c) {
      const int64_t cnt = (c == 0) ? a.nrows : a.n[c];
      int64_t i_local;
      if (c == b) {
        i_local = ib_local;
      } else {
        i_local = rem % cnt;
        rem /= cnt;
        const float f = a.feat[c][((c == 0 ? a.row0 : 0) + i_local) * a.F + jj];
        other = a.op == COORDX_FUSE_PROD ? other * f : other + f;
      }
      q += i_local * mul;
      mul *= cnt;
    }
    float d = dH0[q * a.Hd + jj / a.gs];
    const float fused = a.op == COORDX_FUSE_PROD ? fb * other : fb + other;
    if (dcode) acc_code += d * (a.fba ? act_fwd(a.act, a.S, fused) : fused);
    d *= cd;
    if (a.fba) d *= act_grad(a.act, a.S, fused);
    acc += a.op == COORDX_FUSE_PROD ? d * other : d;
  }
  atomicAdd(&dfeat[ib * a.F + jj], acc);
  if (dcode) atomicAdd(&dcode[jj], acc_code);
}

__global__ void __launch_bounds__(256)
mul_act_grad_kernel(const float* __restrict__ d, const float* __restrict__ Z, float* __restrict__ out, int64_t n, int act, float S) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = d[i] * act_grad(act, S, Z[i]);
}

// loss_functions.py:8-12 (image_mse) and :56-60 (occupancy_3d)
__global__ void __launch_bounds__(256)
loss_kernel(int kind, const float* __restrict__ out, const float* __restrict__ gt, int64_t n, float inv_count,
            float* __restrict__ loss_accum, float* __restrict__ dout) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = out[i], z = gt[i];
    float l, d;
    if (kind == COORDX_LOSS_MSE) {
      const float e = x - z;
      l = e * e;
      d = 2.f * e;
    } else {
      const float ex = expf(-fabsf(x));
      l = fmaxf(x, 0.f) - x * z + logf(1.f + ex);
      const float sig = x >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      d = sig - z;
    }
    s += l;
    if (dout) dout[i] = d * inv_count;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss_accum, t * inv_count);
  }
}

// Batch producer of the video / volume path (dataio.py:575-600, Implicit3DWrapper.__getitem__): the reference gathers the
// ground truth of a training batch on the CPU, data[coord_idx[:, None] * n_xy + coord_idy[None, :]], from a [T * H * W, C]
// array (943 MB for the 300 x 512 x 512 RGB video) and ships it to the GPU every iteration.  Here the array lives on the
// device: out[(i * m + j), :] = data[idz[i] * stride_z + idx[j], :]   (idz == nullptr: one level, row idx[j]).
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ data, int width, const int64_t* __restrict__ idz, int64_t nz, int64_t stride_z,
                   const int64_t* __restrict__ idx, int64_t m, float* __restrict__ out) {
  const int64_t total = nz * m * width;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / width;
    const int c = (int)(e - row * width);
    const int64_t i = row / m, j = row - i * m;
    const int64_t src = (idz ? __ldg(idz + i) * stride_z : 0) + __ldg(idx + j);
    out[e] = __ldg(data + src * width + c);
  }
}

// torch.optim.Adam (training.py:17,146), defaults: no weight decay, no amsgrad
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float step_size, float beta1, float beta2, float eps, float inv_sqrt_bc2, const float* __restrict__ gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gs = gscale ? *gscale : 1.f;
  const float gi = g[i] * gs;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] -= step_size * (mi / denom);
}

// ---- the whole optimiser step of training.py:140-146 over a LIST of tensors, state on the device -----------------
// (clip_grad_norm_ coefficient, Adam for every parameter tensor, step counter and ExponentialLR decay): three small
// launches per iteration, replayable from a CUDA graph because nothing depends on a host-side step count.
struct OptimTensors {
  int n;
  float* p[COORDX_OPT_MAX_TENSORS];
  const float* g[COORDX_OPT_MAX_TENSORS];
  float* m[COORDX_OPT_MAX_TENSORS];
  float* v[COORDX_OPT_MAX_TENSORS];
  int64_t numel[COORDX_OPT_MAX_TENSORS];
};

// state[0] = lr, state[1] = step (as float, exact up to 2^24), state[2] = sum of squared gradients, state[3] = clip coef
__global__ void __launch_bounds__(256) optim_sqnorm_kernel(OptimTensors t, float* __restrict__ state) {
  const int k = blockIdx.y;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < t.numel[k]; i += (int64_t)gridDim.x * 256) {
    const float gi = t.g[k][i];
    s += gi * gi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < 8; ++w) r += red[w];
    if (r != 0.f) atomicAdd(&state[2], r);
  }
}

__global__ void __launch_bounds__(256)
optim_adam_kernel(OptimTensors t, const float* __restrict__ state, float beta1, float beta2, float eps, float max_norm) {
  const int k = blockIdx.y;
  const float lr = state[0], step = state[1] + 1.f;
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  const float coef = max_norm > 0.f ? fminf(max_norm / (sqrtf(state[2]) + 1e-6f), 1.f) : 1.f;
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < t.numel[k]; i += (int64_t)gridDim.x * 256) {
    const float gi = t.g[k][i] * coef;
    const float mi = beta1 * t.m[k][i] + (1.f - beta1) * gi;
    const float vi = beta2 * t.v[k][i] + (1.f - beta2) * gi * gi;
    t.m[k][i] = mi;
    t.v[k][i] = vi;
    t.p[k][i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}

__global__ void optim_advance_kernel(float* __restrict__ state, float lr_gamma) {
  state[0] *= lr_gamma;
  state[1] += 1.f;
  state[2] = 0.f;
}

}  // namespace coordx
