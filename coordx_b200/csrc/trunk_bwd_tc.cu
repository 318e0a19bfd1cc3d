// tcgen05 / TMEM fused BACKWARD of fusion + trunk for hidden = 256 (one Hd x Hd trunk layer +
// the Hd -> O layer): what autograd does for SplitFCBlock.forward_fusion (modules.py:329-354)
// in training.py:137-138, without ever materialising the [P, Hd] activations.
//
// Per 128-point tile (same tiling as the forward: rowvec(o) (+|*) 128 rows of the inner axis):
//   G1  Z^T = W_h H^T             recompute the pre-activation, TRANSPOSED: lanes = output feature n
//   E1  dZ^T = (dout Wf) * act'(Z^T) -> 16-bit tile in smem; db[n], dWf[n] are per-thread running sums
//   G3  dW_h += dZ^T H            accumulated in TMEM over ALL tiles of the CTA, flushed once
//   G2  dH  = dZ W_h              (two 128-column pieces)
//   E2  d rowvec(o) = sum_p dH[p,:] * inner[p,:]   (prod)  |  sum_p dH[p,:]   (sum)
// dW alone is 256x256 fp32 = all 512 TMEM columns, so a CTA owns ONE HALF h of the trunk's
// output features (n_out in [128h, 128h+128)): Z^T_h is 128 columns, dW_h 256, a dH piece 128.
// The partial dH of the two halves add up linearly in the d-rowvec atomics.  The gradient of
// the INNER axis' features needs the other reduction direction (over tiles, not inside a
// tile); it is obtained by running the same kernel a second time with the two axes' roles
// swapped and G3 switched off (DO_DW = false).
//
// Why Z is computed transposed.  With points on the TMEM lanes (first version) every bias / final-weight
// gradient was a cross-lane reduction: 62 shuffles per 32 columns per warp, and E1 (4 warps, sin AND cos per
// element) was the longest stage of a fully serialised G1 -> E1 -> G3/G2 -> E2 chain (12.2 k clk per tile for
// 3.1 k clk of tensor work, profiles/r01a_train_launches.txt).  With the output feature on the lane, bias'[n],
// wfs[n], db[n] and dWf[n] are per-thread registers, E1 runs on 8 warps (two per lane quadrant, 64 points
// each) and hands the dZ tile to the tensor core in two 64-point chunks so G3 starts half an E1 early.
//
// The same shared-memory tiles serve as K-major and as MN-major UMMA operands:
//   W_h   [128 n_out x 256 k_in] : A of G1 (K-major, M = n_out)    and B of G2 (MN-major, N = k_in)
//   H     [128 pts   x 256 k_in] : B of G1 (K-major, N = pts)      and B of G3 (MN-major, N = k_in)
//   dZ^T_h[128 n_out x 128 pts ] : A of G3 (K-major, M = n_out)    and A of G2 (MN-major, M = pts)
//
// CTA = 21 warps: 0-7 E1 (quadrant = warp % 4, point half = warp / 4) | 8-15 E2 (quadrant, column half) |
// 16-19 producers | 20 MMA issuer.
#include <stdlib.h>

#include "trunk_bwd_common.cuh"

namespace coordx {
namespace tcb {

// W fp32 [256(n), 256(k)] -> 4 K-chunks of [256 rows x 128 B] SWIZZLE_128B; epilogue constants
template <typename T>
__global__ void __launch_bounds__(256)
prep_weights_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ Wf,
                    const float* __restrict__ bf, int O, int act, float S, uint8_t* __restrict__ w_img,
                    void* __restrict__ /*unused*/) {
  const int idx = blockIdx.x * 256 + threadIdx.x;          // over 256*256
  const int n = idx >> 8, k = idx & 255;
  const uint32_t off = (k >> 6) * W_CHUNK_BYTES + ptx::sw128_offset(n, k & 63);
  *reinterpret_cast<T*>(w_img + off) = to_T<T>(W[idx]);
}

// inner feature table fp32 [n, 256] -> per 128-row block: 4 K-chunks [128 rows x 128 B] SWIZZLE_128B
template <typename T>
__global__ void __launch_bounds__(256)
prep_inner_kernel(const float* __restrict__ feat, int64_t n, uint8_t* __restrict__ img, int64_t n_padded) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;   // over n_padded*256
  if (idx >= n_padded * HD) return;
  const int64_t j = idx >> 8;
  const int k = (int)(idx & 255);
  const float v = j < n ? feat[idx] : 0.f;
  const size_t off = (size_t)(j >> 7) * INNER_BYTES + (k >> 6) * A_CHUNK_BYTES + ptx::sw128_offset((uint32_t)(j & 127), k & 63);
  *reinterpret_cast<T*>(img + off) = to_T<T>(v);
}


// fp32 table [n, 256] -> 16-bit rows, natural order (rowvec source of the dH-pass producers)
template <typename T>
__global__ void __launch_bounds__(256)
prep_rows16_kernel(const float* __restrict__ feat, int64_t n, uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx < n * HD) reinterpret_cast<T*>(img)[idx] = to_T<T>(feat[idx]);
}

// rowvec(o) for the outer indices o of a slab, 16-bit rows: t0[r0 + o / n1] (op) t1[o % n1]   (t1 == nullptr: t0[r0 + o])
template <typename T>
__global__ void __launch_bounds__(256)
prep_rowvec16_kernel(const float* __restrict__ t0, const float* __restrict__ t1, int64_t r0, int64_t n1, int64_t n_rows, int op,
                     uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n_rows * HD) return;
  const int64_t o = idx >> 8;
  const int k = (int)(idx & 255);
  float v;
  if (t1) {
    const float a = t0[(r0 + o / n1) * HD + k], b = t1[(o % n1) * HD + k];
    v = op == COORDX_FUSE_PROD ? a * b : a + b;
  } else {
    v = t0[(r0 + o) * HD + k];
  }
  reinterpret_cast<T*>(img)[idx] = to_T<T>(v);
}

// sine: bias' = S b, wfs = S Wf  (dz = dout . wfs * cos(S z + bias'));  relu: bias' = b, wfs = Wf
__global__ void __launch_bounds__(256)
prep_bwd_consts_kernel(const float* __restrict__ b, const float* __restrict__ Wf, int O, int act, float S, float* __restrict__ consts) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float sc = act == COORDX_ACT_SINE ? S : 1.f;
  if (i < HD) consts[i] = sc * b[i];
  if (i < MAX_O * HD) consts[HD + i] = (i >> 8) < O ? sc * Wf[i] : 0.f;
}

// the same constants plus a unit loss scale in ONE launch (bf16 operands: no scale needed, see prep_scale_kernel)
__global__ void __launch_bounds__(256)
prep_bwd_consts_unit_kernel(const float* __restrict__ b, const float* __restrict__ Wf, int O, int act, float S, float* __restrict__ consts) {
  const float sc = act == COORDX_ACT_SINE ? S : 1.f;
  for (int i = threadIdx.x; i < MAX_O * HD; i += 256) {
    if (i < HD) consts[i] = sc * b[i];
    consts[HD + i] = (i >> 8) < O ? sc * Wf[i] : 0.f;
  }
  if (threadIdx.x < 2) consts[CONST_SCALE + threadIdx.x] = 1.f;
}

// max |x| over a tensor, as the bit pattern of a non-negative float (atomicMax on unsigned keeps the order)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// Power-of-two loss scale for 16-bit dZ (fp16 operands only; bf16 has fp32's exponent range and uses 1):
// dZ = dout . wfs * act' is packed to fp16 before the gradient GEMMs; with mean losses dout ~ 2 e / P, which is below
// fp16's normal range from ~512^2 points on.  The scale puts max|dout| * max|wfs| at 2^12; every accumulator that
// holds scaled values is multiplied by the exact inverse when it is flushed.
__global__ void __launch_bounds__(256)
prep_scale_kernel(float* __restrict__ consts, int O, int enable) {
  __shared__ float red[8];
  float m = 0.f;
  for (int i = threadIdx.x; i < O * HD; i += 256) m = fmaxf(m, fabsf(consts[HD + i]));
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    const float maxd = __uint_as_float(reinterpret_cast<const unsigned*>(consts)[CONST_SCALE + 2]);
    const float prod = maxd * m;
    int e = 0;
    if (enable && prod > 0.f && isfinite(prod)) e = min(max((int)floorf(log2f(4096.f / prod)), -60), 60);
    consts[CONST_SCALE] = exp2f((float)e);
    consts[CONST_SCALE + 1] = exp2f((float)-e);
  }
}

// column sums over the 32 lanes of a warp: on return lane l holds sum over lanes of v[l]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? v[i + n / 2] : v[i];
      const float send = up ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]);
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void unpack8<__half>(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}


// barrier block (byte offsets from SM_BAR), shared by the two kernels
struct Bars {
  uint32_t full, empty;      // [4] H chunk (dH pass) / [2] H stage (dW pass)
  uint32_t zfull, zempty;    // [2] Z^T accumulators
  uint32_t dzfull;           // [2] 64-point chunks of dZ^T written by E1 (128 each)
  uint32_t dzempty;          // tensor core done with dZ^T (commit)
  uint32_t dhfull, dhempty;  // dH accumulator (dH pass)
  uint32_t dsfull, dsempty;  // dout staging
  uint32_t w, inner, dwfull, tmem_slot, e2_done;
  uint32_t rv;               // [2] rowvec ring (dH pass, copy engine)
  __device__ explicit Bars(uint32_t b0)
      : full(b0), empty(b0 + 32), zfull(b0 + 64), zempty(b0 + 80), dzfull(b0 + 96), dzempty(b0 + 112), dhfull(b0 + 120),
        dhempty(b0 + 128), dsfull(b0 + 136), dsempty(b0 + 144), w(b0 + 152), inner(b0 + 160), dwfull(b0 + 168),
        tmem_slot(b0 + 176), e2_done(b0 + 184), rv(b0 + 192) {}
};

// ---------------------------------------------------------------------------------
// E1 (8 warps, both passes): Z^T_h (TMEM, lanes = output feature n) -> dZ^T_h (16-bit tile in shared memory).
// thread = output feature n = 128 h + row; warp (q, hh) covers the 64 points [64 hh, 64 hh + 64) of the tile.
// bias'[n], wfs[o][n] are registers; db[n] and dWf[o][n] are per-thread running sums (no cross-lane reduction).
// Only the dW pass needs sin (for dWf); the dH passes evaluate cos alone: half the MUFU work.
// ---------------------------------------------------------------------------------
template <typename T, int ACT, int O, bool DO_DW, bool DZ_TMEM = false>
__device__ __forceinline__ void e1_stage(const BwdParams& p, const Bars& B, uint8_t* smem_gen, uint32_t tmem_z0, uint32_t sm_dz,
                                         int h, int warp, int lane, int my_tiles, int64_t tile0 = 0) {
  const int q = warp & 3, hh = warp >> 2;
  const int row = q * 32 + lane;
  const float bias_n = p.consts[h * 128 + row];                 // sine: S b[n]; relu: b[n]
  const float gscale = p.consts[CONST_SCALE], ginv = p.consts[CONST_SCALE + 1];   // power-of-two loss scale (fp16 operands)
  float wfs_n[O];                                               // sine: S Wf[o][n]; relu: Wf[o][n]   (times the loss scale)
#pragma unroll
  for (int oo = 0; oo < O; ++oo) wfs_n[oo] = p.consts[HD + oo * HD + h * 128 + row] * gscale;
  float db_acc = 0.f;
  float dwf_acc[O];
#pragma unroll
  for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = 0.f;
  const float* ds = reinterpret_cast<const float*>(smem_gen + SM_DOUT) + 64 * hh * O;   // this warp's 64 points
  uint8_t* dzrow = smem_gen + sm_dz + hh * CH;                  // 64 points = one 128-byte row of this chunk
  for (int t = 0; t < my_tiles; ++t) {
    const uint32_t zb = t & 1, zuse = t >> 1;
    if (warp == 0 && lane == 0) trace_stamp(p, 1, t, 0);
    ptx::mbar_wait(B.dsfull, t & 1);
    ptx::mbar_wait(B.zfull + 8 * zb, zuse & 1);
    if (warp == 0 && lane == 0) trace_stamp(p, 1, t, 1);
    ptx::tc_fence_after();
    // The whole dZ^T row of this thread (64 points = 32 packed registers) is computed BEFORE the wait for the
    // shared-memory tile: the tensor core is still reading dZ(t-1) (G3 / G2 of the previous tile) while this
    // tile's sin/cos run; only the 8 stores at the end need the tile.  (With the wait in front of the first store
    // E1 and the second GEMM strictly alternated: 5.4 k clk per tile, profiles/r01c_bwd_dh_timeline.txt.)
    uint32_t pk[32];
#pragma unroll
    for (int b = 0; b < 4; ++b) {                  // 16 columns (points) per TMEM load
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(tmem_z0 + zb * 128 + ((uint32_t)(q * 32) << 16) + hh * 64 + b * 16, r);
      ptx::tmem_ld_wait();
      if (b == 3 && !DZ_TMEM) {                    // every Z^T column of this warp is in registers
        ptx::tc_fence_before();
        ptx::mbar_arrive(B.zempty + 8 * zb);
      }
      float4 d4[4];                                // O == 1: dout of the batch's 16 points, 4 broadcast LDS.128
      if (O == 1) {
#pragma unroll
        for (int v = 0; v < 4; ++v) d4[v] = *reinterpret_cast<const float4*>(ds + b * 16 + v * 4);
      }
#pragma unroll
      for (int e2 = 0; e2 < 8; ++e2) {
        float dz2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = e2 * 2 + u;
          const float* dp = ds + (b * 16 + i) * O;              // warp-uniform shared-memory address: broadcast
          float gw, d0 = 0.f;
          if (O == 1) {
            const float4 dv = d4[i >> 2];
            d0 = (i & 3) == 0 ? dv.x : ((i & 3) == 1 ? dv.y : ((i & 3) == 2 ? dv.z : dv.w));
            gw = d0 * wfs_n[0];
          } else {
            gw = 0.f;
#pragma unroll
            for (int oo = 0; oo < O; ++oo) gw = fmaf(dp[oo], wfs_n[oo], gw);
          }
          const float z = __uint_as_float(r[i]);
          float av;                                // activation value (for dWf); the derivative is folded into dz
          if (ACT == COORDX_ACT_SINE) {
            const float th = fmaf(z, p.S, bias_n);
            dz2[u] = gw * __cosf(th);
            av = DO_DW ? __sinf(th) : 0.f;
          } else {
            const float pre = z + bias_n;
            dz2[u] = pre > 0.f ? gw : 0.f;
            av = fmaxf(pre, 0.f);
          }
          if (DO_DW) {
            db_acc += dz2[u];
            if (O == 1) {
              dwf_acc[0] = fmaf(d0, av, dwf_acc[0]);
            } else {
#pragma unroll
              for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = fmaf(dp[oo], av, dwf_acc[oo]);
            }
          }
        }
        pk[b * 8 + e2] = Pack<T>::pack(dz2[0], dz2[1]);
      }
    }
    ptx::mbar_arrive(B.dsempty);                   // last read of the staged dout
    if (DZ_TMEM) {
      // dW pass: dZ^T goes back into TENSOR MEMORY, over the first 32 of this warp's own 64 Z^T columns (all of them
      // are in registers by now), as the A operand of G3: no shared-memory tile, no proxy fence, and G3 reads only
      // its B operand from shared memory (8 KB per K-step: full rate instead of 68 %).  Safe without extra barriers:
      // the tensor pipe is in order, G1(t+2) - the next writer of this buffer - is issued after G3(t).
      if (p.dz_img) {                              // the tile image for the drv / din passes: shared memory, then one bulk store
        ptx::mbar_wait(B.dzempty, (t & 1) ^ 1);    // the copy engine has read the previous image
#pragma unroll
        for (int pc = 0; pc < 8; ++pc)
          *reinterpret_cast<uint4*>(dzrow + ptx::sw128_offset(row, pc * 8)) = make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
        ptx::fence_proxy_async();
      }
      ptx::tmem_st_32x32b_x32(tmem_z0 + zb * 128 + ((uint32_t)(q * 32) << 16) + hh * 64, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(B.dzfull + 8 * hh);
      if (warp == 0 && lane == 0) trace_stamp(p, 1, t, 3);
      continue;
    }
    ptx::mbar_wait(B.dzempty, (t & 1) ^ 1);        // the tensor core is done with dZ(t-1)
    if (warp == 0 && lane == 0) trace_stamp(p, 1, t, 2);
#pragma unroll
    for (int pc = 0; pc < 8; ++pc)                 // 8 points = one swizzled 16-byte piece of row n
      *reinterpret_cast<uint4*>(dzrow + ptx::sw128_offset(row, pc * 8)) = make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
    ptx::fence_proxy_async();
    ptx::mbar_arrive(B.dzfull + 8 * hh);
    if (warp == 0 && lane == 0) trace_stamp(p, 1, t, 3);
  }
  if (DO_DW && my_tiles > 0) {
    atomicAdd(p.db + h * 128 + row, db_acc * ginv);
#pragma unroll
    for (int oo = 0; oo < O; ++oo) atomicAdd(p.dWf + oo * HD + h * 128 + row, dwf_acc[oo]);
  }
}

// one element of a 16-bit operand tile as fp32
template <typename T>
__device__ __forceinline__ float ld_el(const uint8_t* p);
template <>
__device__ __forceinline__ float ld_el<__nv_bfloat16>(const uint8_t* p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(p)) << 16);
}
template <>
__device__ __forceinline__ float ld_el<__half>(const uint8_t* p) { return __half2float(*reinterpret_cast<const __half*>(p)); }

// =================================================================================
// dW pass:  G1 Z^T = W_h H^T  ->  E1 (sin + cos; db, dWf)  ->  G3 dW_h += dZ^T H     (dW, db, dWf)
// TMEM: Z^T x2 (cols 0, 128) | dW_h (cols 256..511).  H double-buffered in shared memory, the inner block lives in
// the producers' registers (as in the forward kernel), so G1(t+1) overlaps E1(t) and G3(t-1).
// =================================================================================
template <typename T, int OP, int ACT, int O>
__global__ void __launch_bounds__(DW_THREADS, 1)
trunk_bwd_dw_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();          // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Bars B(smem_base + SM_BAR);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SM_BAR + 176);
  constexpr int W_PROD = 8, W_MMA = 16;

  const int h = blockIdx.x & 1;                    // half of the trunk's output features
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * cta / n_cta;
  const int64_t t_end = total * (cta + 1) / n_cta;
  const int my_tiles = (int)(t_end - t_begin);

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(B.full + 8 * c, 8);         // 8 producer warps
        ptx::mbar_init(B.empty + 8 * c, 1);
        ptx::mbar_init(B.zfull + 8 * c, 1);
        ptx::mbar_init(B.zempty + 8 * c, 256);
        ptx::mbar_init(B.dzfull + 8 * c, 128);
      }
      ptx::mbar_init(B.dzempty, 1);                // the bulk store has read the dZ^T image
      ptx::mbar_init(B.dsfull, 128);
      ptx::mbar_init(B.dsempty, 256);
      ptx::mbar_init(B.w, 1);
      ptx::mbar_init(B.dwfull, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(B.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tZ = tmem_base, tDW = tmem_base + 256;

  if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (my_tiles > 0 && ptx::elect_one()) {       // (elect, not lane == 0: see ptx::elect_one)
      ptx::mbar_arrive_expect_tx(B.w, 4 * CH);
      for (int kc = 0; kc < 4; ++kc)
        ptx::bulk_g2s(smem_base + SM_W + kc * CH, p.w_img + (size_t)kc * W_CHUNK_BYTES + (size_t)h * CH, CH, B.w);
      ptx::mbar_wait(B.w, 0);
      constexpr uint32_t id_g1 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 0, 0);   // A = W_h (M = n_out), B = H (N = points), both K-major
      constexpr uint32_t id_g3 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 256, 0, 1);   // A = dZ^T (M = n_out) K-major, B = H (N = k_in) MN-major
      auto issue_g1 = [&](int t) {
        const uint32_t zb = t & 1, use = t >> 1;
        // (no wait for the Z^T buffer: its last reader is G3(t-2), earlier in this thread's in-order tensor pipe)
        ptx::mbar_wait(B.full + 8 * zb, use & 1);
        ptx::tc_fence_after();
        const uint32_t hs = smem_base + (zb ? SM_H1 : SM_H);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint32_t off = (k >> 2) * CH + (k & 3) * 32;
          ptx::mma_f16_ss(tZ + zb * 128, ptx::umma_desc_k_sw128(smem_base + SM_W + off), ptx::umma_desc_k_sw128(hs + off), id_g1, k != 0);
        }
        ptx::mma_commit(B.zfull + 8 * zb);
      };
      auto issue_g3 = [&](int t) {
        const uint32_t hs = smem_base + ((t & 1) ? SM_H1 : SM_H);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {              // K = 128 points, 16 per step; starts on the first 64-point chunk of dZ^T
          ptx::mbar_wait(B.dzfull + 8 * c, t & 1);
          ptx::tc_fence_after();
          if (p.dz_img) ptx::bulk_s2g(p.dz_img + ((size_t)(t_begin + t) * 2 + h) * (2 * CH) + c * CH, smem_base + SM_DZ_DW + c * CH, CH);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ks = c * 4 + k;              // A = dZ^T in tensor memory: 8 columns (16 points) per K-step
            ptx::mma_f16_ts(tDW, tZ + (t & 1) * 128 + c * 64 + k * 8, ptx::umma_desc_mn_sw128(hs + ks * 2048, CH), id_g3, (t | ks) != 0);
          }
        }
        ptx::mma_commit(B.empty + 8 * (t & 1));    // H stage free for the producers
        if (p.dz_img) {
          ptx::bulk_commit_group();
          ptx::bulk_wait_group_read0();
          ptx::mbar_arrive(B.dzempty);
        }
      };
      // G1(t+1) is issued AHEAD of G3(t): E1 is the longest stage, so its input must never queue behind G3 in the
      // in-order tensor pipe (polling "whichever is ready first" was measured 18 % slower for exactly that reason)
      issue_g1(0);
      for (int t = 0; t < my_tiles; ++t) {
        if (t + 1 < my_tiles) issue_g1(t + 1);
        issue_g3(t);
      }
      ptx::mma_commit(B.dwfull);
      if (p.dz_img) ptx::bulk_wait_group0();
    }
    __syncwarp();
  } else if (warp >= W_PROD) {
    // ============================ producers: H tile (inner block piece in registers) ============================
    // thread (g = lane, rg = warp - 8) owns the 16-byte piece g (k = 8g .. 8g+7) of the 16 rows 16 rg .. 16 rg + 15
    const int g = lane, rg = warp - W_PROD;
    const int ptid = threadIdx.x - W_PROD * 32;    // 0..255; the first 128 also stage dout
    uint32_t in[16][4];
    int64_t cur_jb = -1;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tt = t_begin + t;
      const int64_t jb = tt / n_o;
      const int64_t o_local = tt - jb * n_o;
      uint32_t rv[4];
      {
        float4 x0, x1;
        load_rowvec8<OP>(p, p.o_begin + o_local, 8 * g, x0, x1);      // L1 hit: prefetched during the previous tile
        rv[0] = Pack<T>::pack(x0.x, x0.y);
        rv[1] = Pack<T>::pack(x0.z, x0.w);
        rv[2] = Pack<T>::pack(x1.x, x1.y);
        rv[3] = Pack<T>::pack(x1.z, x1.w);
      }
      if (t + 1 < my_tiles) {                      // next tile's rowvec -> L1 while this tile is produced (no registers held)
        const int64_t tn = tt + 1, jn = tn / n_o;
        const int64_t on = p.o_begin + (tn - jn * n_o);
        const int64_t i0 = p.outer1 ? on / p.n_outer1 : on;
        ptx::prefetch_l1(p.outer0 + i0 * HD + 8 * g);
        if (p.outer1) ptx::prefetch_l1(p.outer1 + (on - i0 * p.n_outer1) * HD + 8 * g);
      }
      float dreg[O];
#pragma unroll
      for (int oo = 0; oo < O; ++oo) dreg[oo] = 0.f;
      if (ptid < TILE_M) {
        const int64_t j = jb * TILE_M + ptid;
        if (j < p.n_inner) {
#pragma unroll
          for (int oo = 0; oo < O; ++oo) dreg[oo] = __ldg(p.dout + (o_local * p.stride_o + j * p.stride_j) * O + oo);
        }
      }
      if (jb != cur_jb) {
        uint32_t chain = 0;                        // batches of 4 rows chained through an opaque zero: bounds the loads in flight
#pragma unroll
        for (int r4 = 0; r4 < 16; r4 += 4) {
          float4 v0[4], v1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t j = jb * TILE_M + rg * 16 + r4 + u;
            v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            v1[u] = v0[u];
            if (j < p.n_inner) {
              const float* src = p.inner_f32 + j * HD + 8 * g + chain;
              v0[u] = *reinterpret_cast<const float4*>(src);
              v1[u] = *reinterpret_cast<const float4*>(src + 4);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            in[r4 + u][0] = Pack<T>::pack(v0[u].x, v0[u].y);
            in[r4 + u][1] = Pack<T>::pack(v0[u].z, v0[u].w);
            in[r4 + u][2] = Pack<T>::pack(v1[u].x, v1[u].y);
            in[r4 + u][3] = Pack<T>::pack(v1[u].z, v1[u].w);
          }
          asm volatile("and.b32 %0, %1, 0;" : "=r"(chain) : "r"(in[r4 + 3][3]));
        }
        cur_jb = jb;
      }
      const uint32_t st = t & 1, use = t >> 1;
      ptx::mbar_wait(B.empty + 8 * st, (use & 1) ^ 1);
      uint8_t* dst = smem_gen + (st ? SM_H1 : SM_H) + (uint32_t)(g >> 3) * CH + (uint32_t)rg * 2048u;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        uint4 w;
        w.x = fuse2<T, OP>(in[r][0], rv[0]);
        w.y = fuse2<T, OP>(in[r][1], rv[1]);
        w.z = fuse2<T, OP>(in[r][2], rv[2]);
        w.w = fuse2<T, OP>(in[r][3], rv[3]);
        *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + (((g & 7) ^ (r & 7)) << 4)) = w;
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(B.full + 8 * st);
      if (ptid < TILE_M) {
        ptx::mbar_wait(B.dsempty, (t & 1) ^ 1);    // E1(t-1) has read its dout
        float* ds = reinterpret_cast<float*>(smem_gen + SM_DOUT) + ptid * O;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) ds[oo] = dreg[oo];
        ptx::mbar_arrive(B.dsfull);
      }
    }
  } else {
    e1_stage<T, ACT, O, true, true>(p, B, smem_gen, tZ, SM_DZ_DW, h, warp, lane, my_tiles, t_begin);
    // ---- flush dW_h (accumulated over every tile of this CTA) with one atomic per element
    if (my_tiles > 0) {
      ptx::mbar_wait(B.dwfull, 0);
      ptx::tc_fence_after();
      const int q = warp & 3, ch = warp >> 2;
      const float ginv_flush = p.consts[CONST_SCALE + 1];
      float* drow = p.dW + (size_t)(h * 128 + q * 32 + lane) * HD + ch * 128;
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tDW + ((uint32_t)(q * 32) << 16) + ch * 128 + b * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(drow + b * 32 + i, __uint_as_float(r[i]) * ginv_flush);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================
// dH pass:  G1 Z^T = W_h H^T  ->  E1 (cos only)  ->  G2 dH = dZ W_h (256 columns)  ->  E2 d rowvec(o)
// TMEM: Z^T x2 (cols 0, 128) | dH (cols 256..511).  Run twice: with the grid's own roles (gradient of the outer
// axes' features) and with the roles swapped (gradient of the last axis' features).
// =================================================================================
template <typename T, int OP, int ACT, int O>
__global__ void __launch_bounds__(DH_THREADS, 1)
trunk_bwd_dh_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Bars B(smem_base + SM_BAR);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SM_BAR + 176);
  volatile int* e2_done = reinterpret_cast<volatile int*>(smem_gen + SM_BAR + 184);   // tiles fully reduced by E2 (x8 warps)
  constexpr int W_E2 = 8, W_PROD = 16, W_MMA = 20;

  const int h = blockIdx.x & 1;
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * cta / n_cta;
  const int64_t t_end = total * (cta + 1) / n_cta;
  const int my_tiles = (int)(t_end - t_begin);

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) {
        ptx::mbar_init(B.full + 8 * c, 128);
        ptx::mbar_init(B.empty + 8 * c, 1);
      }
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(B.zfull + 8 * c, 1);
        ptx::mbar_init(B.zempty + 8 * c, 256);
        ptx::mbar_init(B.dzfull + 8 * c, 128);
      }
      ptx::mbar_init(B.dzempty, 1);
      ptx::mbar_init(B.dhfull, 1);
      ptx::mbar_init(B.dhempty, 256);
      ptx::mbar_init(B.dsfull, 128);
      ptx::mbar_init(B.dsempty, 256);
      ptx::mbar_init(B.w, 1);
      ptx::mbar_init(B.inner, 1);
      ptx::mbar_init(B.rv, 1);
      ptx::mbar_init(B.rv + 8, 1);
      *e2_done = 0;
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(B.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tZ = tmem_base, tDH = tmem_base + 256;

  if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (my_tiles > 0 && ptx::elect_one()) {       // (elect, not lane == 0: see ptx::elect_one)
      ptx::mbar_arrive_expect_tx(B.w, 4 * CH);
      for (int kc = 0; kc < 4; ++kc)
        ptx::bulk_g2s(smem_base + SM_W + kc * CH, p.w_img + (size_t)kc * W_CHUNK_BYTES + (size_t)h * CH, CH, B.w);
      ptx::mbar_wait(B.w, 0);
      constexpr uint32_t id_g1 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 0, 0);
      constexpr uint32_t id_g2 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 1, 1);   // A = W_h^T (M = k_in) MN-major, B = dZ^T (N = pts) MN-major
      auto issue_g1 = [&](int t) {
        const uint32_t zb = t & 1, use = t >> 1;
        ptx::mbar_wait(B.zempty + 8 * zb, (use & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          ptx::mbar_wait(B.full + 8 * kc, t & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tZ + zb * 128, ptx::umma_desc_k_sw128(smem_base + SM_W + kc * CH + k * 32),
                            ptx::umma_desc_k_sw128(smem_base + SM_H + kc * CH + k * 32), id_g1, (kc | k) != 0);
          ptx::mma_commit(B.empty + 8 * kc);
        }
        ptx::mma_commit(B.zfull + 8 * zb);
      };
      auto issue_g2 = [&](int t) {
        trace_stamp(p, 0, t, 0);
        ptx::mbar_wait(B.dzfull, t & 1);
        ptx::mbar_wait(B.dzfull + 8, t & 1);
        trace_stamp(p, 0, t, 1);
        ptx::mbar_wait(B.dhempty, (t & 1) ^ 1);
        trace_stamp(p, 0, t, 2);
        ptx::tc_fence_after();
        if (OP == COORDX_FUSE_PROD) {
          // d rowvec(o)[k] = sum_p inner[p,k] sum_n dZ[p,n] W[n,k] = sum_n W[n,k] M^T[k,n] with M^T = inner^T dZ:
          // the tensor core contracts over the POINTS (A = inner block, MN-major, two M-blocks of 128 input
          // features; B = dZ^T tile, K-major; K = 128 points = 16 tile rows per step) and E2 is left with a dot
          // product against W[:,k], which is constant for the whole kernel and lives in its registers
          constexpr uint32_t id_g2p = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 1, 0);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t a = ptx::umma_desc_mn_sw128(smem_base + SM_INNER + blk * 2 * CH + ks * 2048, CH);
              const uint64_t b = ptx::umma_desc_k_sw128(smem_base + SM_DZ_DH + (ks >> 2) * CH + (ks & 3) * 32);
              ptx::mma_f16_ss(tDH + blk * 128, a, b, id_g2p, ks != 0);
            }
          }
        } else {
          // dH^T[k_in, pts] = W_h^T dZ^T, two M-blocks of 128 input features; K = 128 n_out = 16 tile rows per step
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t a = ptx::umma_desc_mn_sw128(smem_base + SM_W + blk * 2 * CH + ks * 2048, CH);
              const uint64_t b = ptx::umma_desc_mn_sw128(smem_base + SM_DZ_DH + ks * 2048, CH);
              ptx::mma_f16_ss(tDH + blk * 128, a, b, id_g2, ks != 0);
            }
          }
        }
        ptx::mma_commit(B.dhfull);
        ptx::mma_commit(B.dzempty);
        trace_stamp(p, 0, t, 3);
      };
      issue_g1(0);
      for (int t = 0; t < my_tiles; ++t) {
        // G1(t+1) is issued ahead of G2(t) - except across a change of inner block: there the producers wait for
        // E2(t) before they overwrite the inner block, so H(t+1) cannot exist before G2(t) has been issued
        const bool ahead = t + 1 < my_tiles && (t_begin + t + 1) / n_o == (t_begin + t) / n_o;
        if (ahead) issue_g1(t + 1);
        issue_g2(t);
        if (!ahead && t + 1 < my_tiles) issue_g1(t + 1);
      }
    }
    __syncwarp();
  } else if (warp >= W_PROD) {
    // ============================ producers: H tile (inner block in shared memory, E2 needs it too) ============================
    const int ptid = threadIdx.x - W_PROD * 32;
    const int g8 = ptid & 7, r0 = ptid >> 3;
    const uint32_t piece_off = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((g8 ^ (r0 & 7)) << 4);
    uint32_t inner_phase = 0;
    int64_t cur_jb = -1;
    uint32_t rv[NKC][4];
    // rowvec(o) through the copy engine, two tiles ahead (16-bit rows of p.outer_img): a global load here sat on the
    // producers' critical path for ~2.3 k clk per tile (the LSU queue is full of this role's own LDS/STS traffic)
    const bool rv_tma = O <= 2 && p.outer_img != nullptr && p.outer1 == nullptr;
    const int n_o32 = (int)n_o;
    int64_t jb = my_tiles > 0 ? t_begin / n_o : 0;                       // (no 64-bit division inside the loop)
    int o_local = my_tiles > 0 ? (int)(t_begin - jb * n_o) : 0;
    if (rv_tma && ptid == 0) {
      int on = o_local;
      for (int u = 0; u < 2 && u < my_tiles; ++u) {
        ptx::mbar_arrive_expect_tx(B.rv + 8 * u, 512);
        ptx::bulk_g2s(smem_base + SM_RV + u * 512, p.outer_img + (size_t)(p.o_begin + on) * 512, 512, B.rv + 8 * u);
        if (++on == n_o32) on = 0;
      }
    }
    for (int t = 0; t < my_tiles; ++t) {
      if (t > 0 && ++o_local == n_o32) { o_local = 0; ++jb; }
      float dreg[O];                               // dout of this thread's point: issued first, consumed after the H tile is written
      {
        const int64_t j = jb * TILE_M + ptid;
#pragma unroll
        for (int oo = 0; oo < O; ++oo)
          dreg[oo] = j < p.n_inner ? __ldg(p.dout + ((int64_t)o_local * p.stride_o + j * p.stride_j) * O + oo) : 0.f;
      }
      if (rv_tma) {
        ptx::mbar_wait(B.rv + 8 * (t & 1), (t >> 1) & 1);
        const uint4* src = reinterpret_cast<const uint4*>(smem_gen + SM_RV + (t & 1) * 512 + g8 * 16);
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const uint4 v = src[kc * 8];
          rv[kc][0] = v.x; rv[kc][1] = v.y; rv[kc][2] = v.z; rv[kc][3] = v.w;
        }
        ptx::named_bar_sync(2, 128);               // every producer has read the slot
        if (ptid == 0 && t + 2 < my_tiles) {
          int on = o_local + 2;
          if (on >= n_o32) on -= n_o32;
          if (on >= n_o32) on -= n_o32;            // n_o == 1
          ptx::mbar_arrive_expect_tx(B.rv + 8 * (t & 1), 512);
          ptx::bulk_g2s(smem_base + SM_RV + (t & 1) * 512, p.outer_img + (size_t)(p.o_begin + on) * 512, 512, B.rv + 8 * (t & 1));
        }
      } else {
        // all 8 (16 with a second outer axis) loads are issued back to back: one L2 round trip per tile, not four
        // (the branch on outer1 inside a per-chunk helper serialised them: 2.9 k clk of the producer's 5.3 k period)
        const int64_t o = p.o_begin + o_local;
        int64_t i0 = o, i1 = 0;
        if (p.outer1) { i0 = o / p.n_outer1; i1 = o - i0 * p.n_outer1; }
        const float4* s0 = reinterpret_cast<const float4*>(p.outer0 + i0 * HD + g8 * 8);
        float4 xa[NKC], xb[NKC];
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) { xa[kc] = __ldg(s0 + kc * (KCHUNK / 4)); xb[kc] = __ldg(s0 + kc * (KCHUNK / 4) + 1); }
        if (p.outer1) {
          const float4* s1 = reinterpret_cast<const float4*>(p.outer1 + i1 * HD + g8 * 8);
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            const float4 ya = __ldg(s1 + kc * (KCHUNK / 4)), yb = __ldg(s1 + kc * (KCHUNK / 4) + 1);
            if (OP == COORDX_FUSE_PROD) {
              xa[kc].x *= ya.x; xa[kc].y *= ya.y; xa[kc].z *= ya.z; xa[kc].w *= ya.w;
              xb[kc].x *= yb.x; xb[kc].y *= yb.y; xb[kc].z *= yb.z; xb[kc].w *= yb.w;
            } else {
              xa[kc].x += ya.x; xa[kc].y += ya.y; xa[kc].z += ya.z; xa[kc].w += ya.w;
              xb[kc].x += yb.x; xb[kc].y += yb.y; xb[kc].z += yb.z; xb[kc].w += yb.w;
            }
          }
        }
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          rv[kc][0] = Pack<T>::pack(xa[kc].x, xa[kc].y);
          rv[kc][1] = Pack<T>::pack(xa[kc].z, xa[kc].w);
          rv[kc][2] = Pack<T>::pack(xb[kc].x, xb[kc].y);
          rv[kc][3] = Pack<T>::pack(xb[kc].z, xb[kc].w);
        }
        if (t + 1 < my_tiles) {                    // next tile's rowvec -> L1 while this tile is produced
          int on32 = o_local + 1;
          if (on32 == n_o32) on32 = 0;
          const int64_t on = p.o_begin + on32;
          const int64_t i0n = p.outer1 ? on / p.n_outer1 : on;
          ptx::prefetch_l1(p.outer0 + i0n * HD + ptid * 2);          // 128 threads x 8 B cover the 1 KB row
          if (p.outer1) ptx::prefetch_l1(p.outer1 + (on - i0n * p.n_outer1) * HD + ptid * 2);
        }
      }
      if (jb != cur_jb) {
        ptx::named_bar_sync(1, 128);
        if (ptid == 0) {
          if (OP == COORDX_FUSE_PROD) {            // E2 multiplies dH by the inner block: wait until it is done with it
            while (*e2_done < 8 * t) {
            }
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive_expect_tx(B.inner, INNER_BYTES);
          for (uint32_t c = 0; c < INNER_BYTES; c += CH)
            ptx::bulk_g2s(smem_base + SM_INNER + c, p.inner_img + (size_t)jb * INNER_BYTES + c, CH, B.inner);
        }
        ptx::mbar_wait(B.inner, inner_phase);
        inner_phase ^= 1;
        cur_jb = jb;
      }
      if (ptid == 0) trace_stamp(p, 3, t, 0);
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) ptx::mbar_wait(B.empty + 8 * kc, (t & 1) ^ 1);
      if (ptid == 0) trace_stamp(p, 3, t, 1);
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
        const uint8_t* src = smem_gen + SM_INNER + kc * CH + piece_off;
        uint8_t* dst = smem_gen + SM_H + kc * CH + piece_off;
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(src + i * 2048);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 w;
          w.x = fuse2<T, OP>(v[i].x, rv[kc][0]);
          w.y = fuse2<T, OP>(v[i].y, rv[kc][1]);
          w.z = fuse2<T, OP>(v[i].z, rv[kc][2]);
          w.w = fuse2<T, OP>(v[i].w, rv[kc][3]);
          *reinterpret_cast<uint4*>(dst + i * 2048) = w;
        }
      }
      // ONE generic->async proxy fence per tile (a fence per chunk cost ~4 x 200 clk of MEMBAR), and no global load
      // of this thread is outstanding while it runs (dout is fetched after it)
      ptx::fence_proxy_async();
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) ptx::mbar_arrive(B.full + 8 * kc);
      if (ptid == 0) trace_stamp(p, 3, t, 2);
      ptx::mbar_wait(B.dsempty, (t & 1) ^ 1);      // E1(t-1) has read its dout
      {
        float* ds = reinterpret_cast<float*>(smem_gen + SM_DOUT) + ptid * O;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) ds[oo] = dreg[oo];
      }
      ptx::mbar_arrive(B.dsfull);
      if (ptid == 0) trace_stamp(p, 3, t, 3);
    }
  } else if (warp < W_E2) {
    if (OP == COORDX_FUSE_PROD) ptx::setmaxnreg_dec<72>();       // hand 8 x 256 registers to E2 (W column in registers)
    e1_stage<T, ACT, O, false>(p, B, smem_gen, tZ, SM_DZ_DH, h, warp, lane, my_tiles);
  } else {
    if (OP == COORDX_FUSE_PROD) ptx::setmaxnreg_inc<88>();
    // ============================ E2: dH^T -> d rowvec ============================
    // dH is produced transposed too (lanes = input feature k): d rowvec(o)[k] = sum_p dH^T[k,p] * inner[p,k] is a
    // per-thread sum over the 128 point columns - one atomic per thread per tile, no shuffles.
    // warp (q, blk): k = 128 blk + 32 q + lane
    const int q = warp & 3, blk = (warp - W_E2) >> 2;
    const int k = blk * 128 + q * 32 + lane;
    // prod: W[128 h + n][k] for the 128 output features of this CTA, 16-bit pairs (n = 2i, 2i+1): 64 registers
    uint32_t wreg[64];
    if (OP == COORDX_FUSE_PROD) {
      uint32_t chain = 0;                          // batches of 8 loads chained through an opaque zero (bounds the loads in flight)
#pragma unroll
      for (int i4 = 0; i4 < 64; i4 += 4) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p.w_f32 + (size_t)(h * 128 + 2 * i4 + u) * HD + k + chain);
#pragma unroll
        for (int u = 0; u < 4; ++u) wreg[i4 + u] = Pack<T>::pack(v[2 * u], v[2 * u + 1]);
        asm volatile("and.b32 %0, %1, 0;" : "=r"(chain) : "r"(wreg[i4 + 3]));
      }
    }
    // (no 64-bit division inside the loop: it is a call, and the 64 live W registers would be spilled around it)
    int o_local = my_tiles > 0 ? (int)(t_begin - (t_begin / n_o) * n_o) : 0;
    const int n_o32 = (int)n_o;
    for (int t = 0; t < my_tiles; ++t) {
      if (warp == W_E2 && lane == 0) trace_stamp(p, 2, t, 0);
      ptx::mbar_wait(B.dhfull, t & 1);
      if (warp == W_E2 && lane == 0) trace_stamp(p, 2, t, 1);
      ptx::tc_fence_after();
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};         // independent partial sums: no 128-long dependent FFMA chain
#pragma unroll
      for (int b = 0; b < 16; ++b) {                // 8 columns per load: (prod) n = 8 b .. 8 b + 7, (sum) points
        uint32_t r[8];
        ptx::tmem_ld_32x32b_x8(tDH + blk * 128 + ((uint32_t)(q * 32) << 16) + b * 8, r);
        ptx::tmem_ld_wait();
        if (b == 15) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(B.dhempty);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (OP == COORDX_FUSE_PROD) {
            // opaque to the optimiser on purpose: the unpack is loop-invariant and would otherwise be hoisted out of
            // the tile loop as 128 fp32 registers (spilled)
            uint32_t wp = wreg[b * 4 + (i >> 1)];
            asm volatile("" : "+r"(wp));
            const float w = Pack<T>::unpack(wp, i & 1);
            acc4[i & 3] = fmaf(__uint_as_float(r[i]), w, acc4[i & 3]);
          } else {
            acc4[i & 3] += __uint_as_float(r[i]);
          }
        }
      }
      atomicAdd(p.d_rv + (int64_t)o_local * HD + k, ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3])) * p.consts[CONST_SCALE + 1]);
      if (OP == COORDX_FUSE_PROD) {
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<int*>(e2_done), 1);
      }
      if (warp == W_E2 && lane == 0) trace_stamp(p, 2, t, 2);
      if (++o_local == n_o32) o_local = 0;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================
// d-rowvec pass from STORED dZ^T tiles (same tiling as the dW pass, which wrote them):
//   G2' M^T = inner^T dZ (prod) | dH^T = W_h^T dZ^T (sum)  ->  E2 d rowvec(o)
// No G1, no E1, no H tile: the copy engine streams the 32 KB tiles (4-deep ring), the accumulator is double-buffered
// in TMEM (2 x 256 columns), E2 is the only arithmetic.  64 KB of global reads per point tile: HBM-bound.
// CTA = 10 warps: 0-7 E2 | 8 tile loader | 9 MMA issuer.
// =================================================================================
constexpr int DRV_THREADS = 10 * 32, DRV_STAGES = 4;
constexpr uint32_t SMR_OPER = 0;                        // inner block (prod) / W_h (sum): 4 chunks
constexpr uint32_t SMR_DZ = 4 * CH;                     // DRV_STAGES x 2 chunks
constexpr uint32_t SMR_BAR = SMR_DZ + DRV_STAGES * 2 * CH;
constexpr uint32_t SMR_TOTAL = SMR_BAR + 256;

template <typename T, int OP>
__global__ void __launch_bounds__(DRV_THREADS, 1)
trunk_bwd_drv_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SMR_BAR;
  const uint32_t b_dzfull = bar0, b_dzempty = bar0 + 8 * DRV_STAGES, b_dhfull = bar0 + 16 * DRV_STAGES, b_dhempty = b_dhfull + 16,
                 b_oper = b_dhfull + 32, b_slot = b_dhfull + 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SMR_BAR + 16 * DRV_STAGES + 40);
  constexpr int W_LOAD = 8, W_MMA = 9;

  const int h = blockIdx.x & 1;
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * cta / n_cta;
  const int64_t t_end = total * (cta + 1) / n_cta;
  const int my_tiles = (int)(t_end - t_begin);

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < DRV_STAGES; ++s) {
        ptx::mbar_init(b_dzfull + 8 * s, 1);
        ptx::mbar_init(b_dzempty + 8 * s, 1);
      }
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(b_dhfull + 8 * c, 1);
        ptx::mbar_init(b_dhempty + 8 * c, 256);
      }
      ptx::mbar_init(b_oper, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(b_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == W_LOAD) {
    // ============================ loader: operand block + the ring of dZ^T tiles ============================
    if (lane == 0) {
      int64_t cur_jb = -1;
      uint32_t oper_loads = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int64_t tt = t_begin + t;
        const int64_t jb = tt / n_o;
        if (OP == COORDX_FUSE_PROD ? jb != cur_jb : t == 0) {
          // the previous tile's MMAs read the operand block: its commit on the ring's "empty" barrier covers them
          if (t > 0) ptx::mbar_wait(b_dzempty + 8 * ((t - 1) % DRV_STAGES), ((t - 1) / DRV_STAGES) & 1);
          ptx::mbar_arrive_expect_tx(b_oper, 4 * CH);
          for (int kc = 0; kc < 4; ++kc) {
            if (OP == COORDX_FUSE_PROD)
              ptx::bulk_g2s(smem_base + SMR_OPER + kc * CH, p.inner_img + (size_t)jb * INNER_BYTES + (size_t)kc * CH, CH, b_oper);
            else
              ptx::bulk_g2s(smem_base + SMR_OPER + kc * CH, p.w_img + (size_t)kc * W_CHUNK_BYTES + (size_t)h * CH, CH, b_oper);
          }
          ++oper_loads;
          cur_jb = jb;
        }
        const int s = t % DRV_STAGES;
        ptx::mbar_wait(b_dzempty + 8 * s, ((t / DRV_STAGES) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(b_dzfull + 8 * s, 2 * CH);
        ptx::bulk_g2s(smem_base + SMR_DZ + s * 2 * CH, p.dz_img + ((size_t)tt * 2 + h) * (2 * CH), 2 * CH, b_dzfull + 8 * s);
      }
      (void)oper_loads;
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (ptx::elect_one()) {
      constexpr uint32_t id_g2 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 1, 1);    // sum: A = W_h^T MN-major, B = dZ^T MN-major
      constexpr uint32_t id_g2p = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 1, 0);   // prod: A = inner^T MN-major, B = dZ^T K-major
      int64_t cur_jb = -1;
      uint32_t oper_phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int64_t jb = (t_begin + t) / n_o;
        if (OP == COORDX_FUSE_PROD ? jb != cur_jb : t == 0) {
          ptx::mbar_wait(b_oper, oper_phase);
          oper_phase ^= 1;
          cur_jb = jb;
        }
        const int s = t % DRV_STAGES, ab = t & 1;
        ptx::mbar_wait(b_dzfull + 8 * s, (t / DRV_STAGES) & 1);
        ptx::mbar_wait(b_dhempty + 8 * ab, ((t >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t dz = smem_base + SMR_DZ + s * 2 * CH;
        const uint32_t acc = tmem_base + ab * 256;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = ptx::umma_desc_mn_sw128(smem_base + SMR_OPER + blk * 2 * CH + ks * 2048, CH);
            if (OP == COORDX_FUSE_PROD)
              ptx::mma_f16_ss(acc + blk * 128, a, ptx::umma_desc_k_sw128(dz + (ks >> 2) * CH + (ks & 3) * 32), id_g2p, ks != 0);
            else
              ptx::mma_f16_ss(acc + blk * 128, a, ptx::umma_desc_mn_sw128(dz + ks * 2048, CH), id_g2, ks != 0);
          }
        }
        ptx::mma_commit(b_dhfull + 8 * ab);
        ptx::mma_commit(b_dzempty + 8 * s);
      }
    }
    __syncwarp();
  } else {
    // ============================ E2: accumulator -> d rowvec (as in the dH kernel) ============================
    const int q = warp & 3, blk = warp >> 2;
    const int k = blk * 128 + q * 32 + lane;
    uint32_t wreg[64];
    if (OP == COORDX_FUSE_PROD) {
      uint32_t chain = 0;
#pragma unroll
      for (int i4 = 0; i4 < 64; i4 += 4) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p.w_f32 + (size_t)(h * 128 + 2 * i4 + u) * HD + k + chain);
#pragma unroll
        for (int u = 0; u < 4; ++u) wreg[i4 + u] = Pack<T>::pack(v[2 * u], v[2 * u + 1]);
        asm volatile("and.b32 %0, %1, 0;" : "=r"(chain) : "r"(wreg[i4 + 3]));
      }
    }
    int o_local = my_tiles > 0 ? (int)(t_begin - (t_begin / n_o) * n_o) : 0;
    const int n_o32 = (int)n_o;
    for (int t = 0; t < my_tiles; ++t) {
      const int ab = t & 1;
      ptx::mbar_wait(b_dhfull + 8 * ab, (t >> 1) & 1);
      ptx::tc_fence_after();
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
      // 4 loads of 32 columns, double-buffered: a load per 8 columns with its own wait made E2 latency-bound
      // (16 serialised TMEM round trips = 2.3 k clk per tile for 1 k clk of tensor work)
      const uint32_t tacc = tmem_base + ab * 256 + blk * 128 + ((uint32_t)(q * 32) << 16);
      uint32_t ra[32], rb[32];
      ptx::tmem_ld_32x32b_x32(tacc, ra);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint32_t(&r)[32] = (b & 1) ? rb : ra;
        ptx::tmem_ld_wait();
        if (b < 3) ptx::tmem_ld_32x32b_x32(tacc + (b + 1) * 32, (b & 1) ? ra : rb);
        if (b == 3) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(b_dhempty + 8 * ab);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (OP == COORDX_FUSE_PROD) {
            uint32_t wp = wreg[b * 16 + (i >> 1)];
            asm volatile("" : "+r"(wp));             // keep the unpack inside the loop (see the dH kernel)
            acc4[i & 3] = fmaf(__uint_as_float(r[i]), Pack<T>::unpack(wp, i & 1), acc4[i & 3]);
          } else {
            acc4[i & 3] += __uint_as_float(r[i]);
          }
        }
      }
      atomicAdd(p.d_rv + (int64_t)o_local * HD + k, (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
      if (++o_local == n_o32) o_local = 0;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================
// inner-axis gradient pass from the STORED dZ^T tiles (tiling of the dW pass):
//   d inner[j,k] = sum_o rowvec(o)[k] * sum_n dZ[(o,j),n] W[n,k]
// For a fixed block of 128 inner rows the sum over the outer index o runs over CONSECUTIVE tiles of a CTA, so the
// tensor core accumulates it in TMEM and the accumulator is flushed with atomics only when the inner block changes
// (once or twice per CTA).  Computed transposed,  acc^T[k,j] += (W_h (.) rowvec(o))^T[k,n] . dZ^T_tile[n,j]
// (two M-blocks of 128 input features k, N = 128 j, K = 128 n of this CTA's half), because then
//   * the per-tile operand (W_h scaled by rowvec(o)) has k on the TMEM LANE: thread k keeps its W column in 64
//     registers (16-bit pairs), scales it by ONE scalar rowvec(o)[k] and writes it straight into tensor memory as the
//     A operand (tcgen05.st) - no shared-memory tile, no proxy fence;
//   * the MMA reads only the streamed dZ^T tile from shared memory (4 KB per K-step: the operand-fetch limit of
//     ~70 B/clk is not hit, unlike the SS form with 12 KB per step which ran at 68 %).
// No G1, no E1, no per-tile epilogue; the copy engine streams the dZ^T tiles (4-deep ring) and the rowvec rows.
// CTA = 10 warps: 0-7 scalers (+ accumulator flush; warp = (lane quadrant, M-block)) | 8 tile loader | 9 MMA issuer.
// TMEM: acc^T 2 x 128 columns | A operand 2 buffers x 2 M-blocks x 64 columns.
// =================================================================================
#ifndef COORDX_DIN_EXP
#define COORDX_DIN_EXP 0          // timing experiments only: 1 = no MMAs, 2 = no tile loads
#endif
constexpr int DIN_THREADS = 10 * 32, DIN_DZ_STAGES = 4;
constexpr uint32_t SMI_DZ = 0;                          // DIN_DZ_STAGES x 2 chunks
constexpr uint32_t SMI_BAR = SMI_DZ + DIN_DZ_STAGES * 2 * CH;
constexpr uint32_t SMI_RV = SMI_BAR + 256;              // 2 x 512 B rowvec slots
constexpr uint32_t SMI_TOTAL = SMI_RV + 1024;

template <typename T, int OP>
__global__ void __launch_bounds__(DIN_THREADS, 1)
trunk_bwd_din_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SMI_BAR;
  const uint32_t b_afull = bar0, b_aempty = bar0 + 16, b_dzfull = bar0 + 32, b_dzempty = bar0 + 64, b_accfull = bar0 + 96,
                 b_accempty = bar0 + 104, b_rv = bar0 + 112, b_slot = bar0 + 128;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SMI_BAR + 128);
  constexpr int W_LOAD = 8, W_MMA = 9;

  const int h = blockIdx.x & 1;
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * cta / n_cta;
  const int64_t t_end = total * (cta + 1) / n_cta;
  const int my_tiles = (int)(t_end - t_begin);
  const int n_o32 = (int)n_o;
  const int64_t jb0 = my_tiles > 0 ? t_begin / n_o : 0;
  const int o0 = my_tiles > 0 ? (int)(t_begin - jb0 * n_o) : 0;

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(b_afull + 8 * c, 256);
        ptx::mbar_init(b_aempty + 8 * c, 1);
        ptx::mbar_init(b_rv + 8 * c, 1);
      }
      for (int c = 0; c < DIN_DZ_STAGES; ++c) {
        ptx::mbar_init(b_dzfull + 8 * c, 1);
        ptx::mbar_init(b_dzempty + 8 * c, 1);
      }
      ptx::mbar_init(b_accfull, 1);
      ptx::mbar_init(b_accempty, 256);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(b_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tacc = *tmem_slot_gen, tA = tacc + 256;

  if (warp == W_LOAD) {
    // ============================ loader: dZ^T tiles ============================
    if (my_tiles > 0 && ptx::elect_one()) {
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t % DIN_DZ_STAGES;
        ptx::mbar_wait(b_dzempty + 8 * s, ((t / DIN_DZ_STAGES) & 1) ^ 1);
#if COORDX_DIN_EXP & 2
        if (t >= DIN_DZ_STAGES) { ptx::mbar_arrive(b_dzfull + 8 * s); continue; }
#endif
        ptx::mbar_arrive_expect_tx(b_dzfull + 8 * s, 2 * CH);
        ptx::bulk_g2s(smem_base + SMI_DZ + s * 2 * CH, p.dz_img + ((size_t)(t_begin + t) * 2 + h) * (2 * CH), 2 * CH, b_dzfull + 8 * s);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (my_tiles > 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 0, 1);   // A in tensor memory, B = dZ^T (N = j) MN-major
      int o_local = o0;
      bool fresh = true;
      uint32_t flushes = 0;
      for (int t = 0; t < my_tiles; ++t) {
        if (t > 0 && ++o_local == n_o32) {           // next block of inner rows: hand the accumulator to the flush
          o_local = 0;
          ptx::mma_commit(b_accfull);
          ptx::mbar_wait(b_accempty, flushes & 1);
          ++flushes;
          ptx::tc_fence_after();
          fresh = true;
        }
        const int s = t % DIN_DZ_STAGES, ab = t & 1;
        ptx::mbar_wait(b_dzfull + 8 * s, (t / DIN_DZ_STAGES) & 1);
        ptx::mbar_wait(b_afull + 8 * ab, (t >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t dz = smem_base + SMI_DZ + s * 2 * CH;
#if !(COORDX_DIN_EXP & 1)
#pragma unroll
#endif
        for (int blk = 0; blk < ((COORDX_DIN_EXP & 1) ? 0 : 2); ++blk) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            ptx::mma_f16_ts(tacc + blk * 128, tA + ab * 128 + blk * 64 + ks * 8, ptx::umma_desc_mn_sw128(dz + ks * 2048, CH), idesc,
                            !(fresh && ks == 0));
        }
        fresh = false;
        ptx::mma_commit(b_aempty + 8 * ab);
        ptx::mma_commit(b_dzempty + 8 * s);
      }
      ptx::mma_commit(b_accfull);
    }
    __syncwarp();
  } else {
    // ============================ scalers: A operand = (W_h (.) rowvec(o))^T in tensor memory; accumulator flush ============================
    const int q = warp & 3, blk = warp >> 2;
    const int k = blk * 128 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t wreg[64];                               // W[128 h + n][k], 16-bit pairs (n = 2i, 2i+1)
    {
      uint32_t chain = 0;
#pragma unroll
      for (int i4 = 0; i4 < 64; i4 += 4) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p.w_f32 + (size_t)(h * 128 + 2 * i4 + u) * HD + k + chain);
#pragma unroll
        for (int u = 0; u < 4; ++u) wreg[i4 + u] = Pack<T>::pack(v[2 * u], v[2 * u + 1]);
        asm volatile("and.b32 %0, %1, 0;" : "=r"(chain) : "r"(wreg[i4 + 3]));
      }
    }
    auto flush = [&](int64_t jb, uint32_t nflush) {
      ptx::mbar_wait(b_accfull, nflush & 1);
      ptx::tc_fence_after();
      float* dcol = p.d_inner + jb * TILE_M * HD + k;            // acc^T[k, j]: lanes = consecutive k (coalesced atomics)
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tacc + blk * 128 + lane_base + b * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (jb * TILE_M + b * 32 + i < p.n_inner) atomicAdd(dcol + (size_t)(b * 32 + i) * HD, __uint_as_float(r[i]));
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(b_accempty);
    };
    int64_t jb = jb0;
    int o_local = o0;
    uint32_t nflush = 0;
    if (OP == COORDX_FUSE_PROD && threadIdx.x == 0) {          // rowvec ring prologue
      int on = o0;
      for (int u = 0; u < 2 && u < my_tiles; ++u) {
        ptx::mbar_arrive_expect_tx(b_rv + 8 * u, 512);
        ptx::bulk_g2s(smem_base + SMI_RV + u * 512, p.outer_img + (size_t)on * 512, 512, b_rv + 8 * u);
        if (++on == n_o32) on = 0;
      }
    }
    for (int t = 0; t < my_tiles; ++t) {
      if (t > 0 && ++o_local == n_o32) {
        o_local = 0;
        flush(jb, nflush++);
        ++jb;
      }
      uint32_t rv2 = Pack<T>::pack(1.f, 1.f);        // sum fusion: dH is not scaled
      if (OP == COORDX_FUSE_PROD) {
        ptx::mbar_wait(b_rv + 8 * (t & 1), (t >> 1) & 1);
        const uint32_t e = *reinterpret_cast<const uint16_t*>(smem_gen + SMI_RV + (t & 1) * 512 + k * 2);
        rv2 = e | (e << 16);
        ptx::named_bar_sync(2, 256);                 // every scaler has read the slot
        if (threadIdx.x == 0 && t + 2 < my_tiles) {
          int on = o_local + 2;
          if (on >= n_o32) on -= n_o32;
          if (on >= n_o32) on -= n_o32;
          ptx::mbar_arrive_expect_tx(b_rv + 8 * (t & 1), 512);
          ptx::bulk_g2s(smem_base + SMI_RV + (t & 1) * 512, p.outer_img + (size_t)on * 512, 512, b_rv + 8 * (t & 1));
        }
      }
      const uint32_t ab = t & 1;
      uint32_t sc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) sc[i] = fuse2<T, COORDX_FUSE_PROD>(wreg[i], rv2);
      ptx::mbar_wait(b_aempty + 8 * ab, ((t >> 1) & 1) ^ 1);     // the tensor core is done with this A buffer (tile t-2)
      ptx::tc_fence_after();
      ptx::tmem_st_32x32b_x32(tA + ab * 128 + blk * 64 + lane_base, sc);
#pragma unroll
      for (int i = 0; i < 32; ++i) sc[i] = fuse2<T, COORDX_FUSE_PROD>(wreg[32 + i], rv2);
      ptx::tmem_st_32x32b_x32(tA + ab * 128 + blk * 64 + 32 + lane_base, sc);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(b_afull + 8 * ab);
    }
    if (my_tiles > 0) flush(jb, nflush);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tacc, 512);
  }
}

// ---------------------------------------------------------------------------------
// small helpers for the 3-axis case
// ---------------------------------------------------------------------------------
// pair table (axis0 x axis1 fused) as the INNER operand of the role-swapped pass:
// row (i0 - r0) * n1 + i1  ->  fuse(t0[i0], t1[i1]), 16-bit, blocked + swizzled like prep_inner_kernel
template <typename T>
__global__ void __launch_bounds__(256)
prep_pair_inner_kernel(const float* __restrict__ t0, const float* __restrict__ t1, int64_t r0, int64_t n1, int64_t n_rows,
                       int64_t n_padded, int op, uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n_padded * HD) return;
  const int64_t j = idx >> 8;
  const int k = (int)(idx & 255);
  float v = 0.f;
  if (j < n_rows) {
    const int64_t i0 = r0 + j / n1, i1 = j % n1;
    const float a = t0[i0 * HD + k], b = t1[i1 * HD + k];
    v = op == COORDX_FUSE_PROD ? a * b : a + b;
  }
  const size_t off = (size_t)(j >> 7) * INNER_BYTES + (k >> 6) * A_CHUNK_BYTES + ptx::sw128_offset((uint32_t)(j & 127), k & 63);
  *reinterpret_cast<T*>(img + off) = to_T<T>(v);
}

// d rowvec[(i0,i1)] -> d t0[i0] (which == 0) or d t1[i1] (which == 1); one thread per (row, k), no atomics
__global__ void __launch_bounds__(256)
pair_scatter_kernel(const float* __restrict__ drv, const float* __restrict__ t0, const float* __restrict__ t1, int64_t r0,
                    int64_t rows0, int64_t n1, int op, int which, float* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t cnt = which == 0 ? rows0 : n1;
  if (idx >= cnt * HD) return;
  const int64_t i = idx >> 8;
  const int k = (int)(idx & 255);
  float acc = 0.f;
  if (which == 0) {
    for (int64_t i1 = 0; i1 < n1; ++i1) {
      const float g = drv[(i * n1 + i1) * HD + k];
      acc += op == COORDX_FUSE_PROD ? g * t1[i1 * HD + k] : g;
    }
    dst[(r0 + i) * HD + k] += acc;
  } else {
    for (int64_t i0 = 0; i0 < rows0; ++i0) {
      const float g = drv[(i0 * n1 + i) * HD + k];
      acc += op == COORDX_FUSE_PROD ? g * t0[(r0 + i0) * HD + k] : g;
    }
    dst[i * HD + k] += acc;
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <typename T, int OP, int ACT, int O, bool DO_DW>
static int launch_bwd_inst(const BwdParams& bp, int grid, cudaStream_t st) {
  static AttrOnce once;
  if (DO_DW) {
    auto kern = trunk_bwd_dw_kernel<T, OP, ACT, O>;
    CX_TRY(once.set(kern, (int)SM_TOTAL_B));
    kern<<<grid, DW_THREADS, SM_TOTAL_B, st>>>(bp);
  } else {
    auto kern = trunk_bwd_dh_kernel<T, OP, ACT, O>;
    CX_TRY(once.set(kern, (int)SM_TOTAL_B));
    kern<<<grid, DH_THREADS, SM_TOTAL_B, st>>>(bp);
  }
  CX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int OP, int ACT, bool DO_DW>
static int launch_bwd_o(const BwdParams& bp, int O, int grid, cudaStream_t st) {
  switch (O) {
    case 1: return launch_bwd_inst<T, OP, ACT, 1, DO_DW>(bp, grid, st);
    case 2: return launch_bwd_inst<T, OP, ACT, 2, DO_DW>(bp, grid, st);
    case 3: return launch_bwd_inst<T, OP, ACT, 3, DO_DW>(bp, grid, st);
    case 4: return launch_bwd_inst<T, OP, ACT, 4, DO_DW>(bp, grid, st);
  }
  return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk backward: out width %s%lld > 4", "", O);
}

template <typename T, bool DO_DW>
static int launch_bwd_t(const BwdParams& bp, int op, int act, int O, int grid, cudaStream_t st) {
  if (op == COORDX_FUSE_PROD)
    return act == COORDX_ACT_SINE ? launch_bwd_o<T, COORDX_FUSE_PROD, COORDX_ACT_SINE, DO_DW>(bp, O, grid, st)
                                  : launch_bwd_o<T, COORDX_FUSE_PROD, COORDX_ACT_RELU, DO_DW>(bp, O, grid, st);
  return act == COORDX_ACT_SINE ? launch_bwd_o<T, COORDX_FUSE_SUM, COORDX_ACT_SINE, DO_DW>(bp, O, grid, st)
                                : launch_bwd_o<T, COORDX_FUSE_SUM, COORDX_ACT_RELU, DO_DW>(bp, O, grid, st);
}

template <typename T>
static int launch_drv(const BwdParams& bp, int op, int grid, cudaStream_t st) {
  static AttrOnce once_p, once_s;
  CX_TRY(once_p.set(trunk_bwd_drv_kernel<T, COORDX_FUSE_PROD>, (int)SMR_TOTAL));
  CX_TRY(once_s.set(trunk_bwd_drv_kernel<T, COORDX_FUSE_SUM>, (int)SMR_TOTAL));
  if (op == COORDX_FUSE_PROD)
    trunk_bwd_drv_kernel<T, COORDX_FUSE_PROD><<<grid, DRV_THREADS, SMR_TOTAL, st>>>(bp);
  else
    trunk_bwd_drv_kernel<T, COORDX_FUSE_SUM><<<grid, DRV_THREADS, SMR_TOTAL, st>>>(bp);
  CX_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_din(const BwdParams& bp, int op, int grid, cudaStream_t st) {
  static AttrOnce once_p, once_s;
  CX_TRY(once_p.set(trunk_bwd_din_kernel<T, COORDX_FUSE_PROD>, (int)SMI_TOTAL));
  CX_TRY(once_s.set(trunk_bwd_din_kernel<T, COORDX_FUSE_SUM>, (int)SMI_TOTAL));
  if (op == COORDX_FUSE_PROD)
    trunk_bwd_din_kernel<T, COORDX_FUSE_PROD><<<grid, DIN_THREADS, SMI_TOTAL, st>>>(bp);
  else
    trunk_bwd_din_kernel<T, COORDX_FUSE_SUM><<<grid, DIN_THREADS, SMI_TOTAL, st>>>(bp);
  CX_LAUNCH_CHECK();
  return 0;
}

// dZ^T tiles kept between the dW pass and the feature-gradient pass (64 KB per 128-point tile).  The caller decides
// by the workspace it provides: coordx_trunk_workspace_bytes(.., backward = 1) asks for the tiles unless
// COORDX_BWD_DZ_REUSE=0 or they would need more than COORDX_BWD_DZ_MAX_GB (default: 24 GB, at most a quarter of the
// device's memory); backward = 2 is the size WITHOUT them (the recompute passes), which trunk_backward_tc accepts
// as well - the host glue falls back to it when the large allocation fails.
static bool dz_reuse_wanted(int64_t tiles) {
  const char* e = getenv("COORDX_BWD_DZ_REUSE");
  if (e && e[0] == '0') return false;
  const char* m = getenv("COORDX_BWD_DZ_MAX_GB");
  double max_bytes = (m ? atof(m) : 24.0) * 1e9;
  if (!m) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) max_bytes = fmin(max_bytes, 0.25 * (double)total_b);
    else (void)cudaGetLastError();
  }
  return (double)tiles * (4.0 * CH) <= max_bytes;
}

// COORDX_BWD_V1=1: the first-generation kernels (dW pass with W_h in shared memory, separate drv and din passes): A/B timing only
static bool bwd_v1() {
  const char* e = getenv("COORDX_BWD_V1");
  return e && e[0] == '1';
}

static bool bwd_recompute_forced() {
  const char* e = getenv("COORDX_BWD_DZ_REUSE");
  return e && e[0] == '0';
}

static long long* g_bwd_trace = nullptr;
static int g_bwd_trace_pass = 1;      // 0 = dW pass, 1 = dH pass A, 2 = dH pass B

// COORDX_BWD_FUSED=1: both passes of the stored-dZ path in ONE launch (trunk_bwd_fused_kernel) when the grid has at
// least FUSED_MIN_TILES_PER_SM tiles per SM.  Opt-in: measured on 2048^2 (profiles/r02f_fused_bwd_sweep.txt) the dW
// CTAs keep their stand-alone rate (2.28 us per tile and pair) but a feature-gradient CTA needs 1.9-2.1 us per tile next
// to them - no faster than HBM-fed - so the best split (1.45 ms) does not beat the two launches (1.01 + 0.43 ms).
// COORDX_BWD_FUSED_PAIRS / _CHUNK / _GROUPS override the role split (default: 3/8 of the SMs as dW pairs), the
// hand-over chunk and the number of CTA groups; a negative chunk makes the consumers idle (producer-only timing).
constexpr int FUSED_MIN_TILES_PER_SM = 16;
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && e[0] ? atoi(e) : dflt;
}
static bool bwd_fused_wanted(int64_t tiles) {
  return env_int("COORDX_BWD_FUSED", 0) != 0 && tiles >= (int64_t)FUSED_MIN_TILES_PER_SM * num_sms() && num_sms() >= 8;
}

struct BwdPlan {
  size_t w_img_off, epi_off, consts_off, innerA_off, innerB_off, drv_off, rows_off, flags_off, dz_off, total;
  bool reuse;
  int nbA, nbB;
  int64_t n_innerA, n_innerB, n_oA;
};

static BwdPlan plan_bwd(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1, bool reuse) {
  const int C = net->n_branches;
  BwdPlan pl{};
  pl.n_innerA = n[C - 1];
  pl.n_oA = (r1 - r0) * (C == 3 ? n[1] : 1);
  pl.n_innerB = pl.n_oA;                           // role swap: the (slab of the) outer index becomes the inner axis
  pl.nbA = (int)cdiv(pl.n_innerA, TILE_M);
  pl.nbB = (int)cdiv(pl.n_innerB, TILE_M);
  size_t off = 0;
  pl.w_img_off = off; off += align_up(W_BYTES, 1024);
  pl.epi_off = off;
  pl.consts_off = off; off += align_up(CONST_FLOATS * 4, 1024);
  pl.innerA_off = off; off += (size_t)pl.nbA * INNER_BYTES;
  pl.innerB_off = off; off += (size_t)pl.nbB * INNER_BYTES;
  pl.drv_off = off; off += C == 3 ? align_up((size_t)pl.n_oA * HD * 4, 1024) : 0;
  pl.rows_off = off; off += align_up((size_t)max(pl.n_innerA, pl.n_oA) * HD * 2, 1024);   // 16-bit rowvec rows (din pass / pass B)
  pl.reuse = reuse;
  pl.flags_off = off; off += reuse ? align_up((size_t)pl.nbA * pl.n_oA * 4, 1024) : 0;    // per-tile hand-over counters (fused kernel)
  pl.dz_off = off; off += pl.reuse ? (size_t)pl.nbA * pl.n_oA * (4 * CH) : 0;
  pl.total = off;
  return pl;
}

template <typename T>
static int trunk_backward_tc_t(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                               const float* dout, coordx_grads* grads, float* const* dfeat, int precision, void* ws, size_t ws_bytes,
                               cudaStream_t st) {
  const int C = net->n_branches, L = net->n_pre, O = net->layer_out[L + 1];
  BwdPlan pl = plan_bwd(net, n, r0, r1, true);     // the stored-dZ path if the caller's workspace holds the tiles
  if (ws_bytes < pl.total || bwd_recompute_forced()) pl = plan_bwd(net, n, r0, r1, false);
  if (!ws || ws_bytes < pl.total) return set_err(COORDX_ERR_WORKSPACE, "tcgen05 trunk backward: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)pl.total);
  uint8_t* base = (uint8_t*)ws;
  uint8_t* w_img = base + pl.w_img_off;
  float* consts = (float*)(base + pl.consts_off);
  uint8_t* innerA = base + pl.innerA_off;
  uint8_t* innerB = base + pl.innerB_off;
  float* drv = (float*)(base + pl.drv_off);
  const float S = 30.f * net->act_scale;
  const int64_t n1 = C == 3 ? n[1] : 1;
  const int64_t n_last = n[C - 1];

  const bool v1 = bwd_v1();
  // the 16-bit weight / inner-block images feed the first-generation and the recompute kernels only: the stored-dZ path
  // (dw2 + dfeat) converts the fp32 tables itself
  const bool need_images = v1 || !pl.reuse;
  if (need_images) {
    prep_weights_kernel<T><<<HD * HD / 256, 256, 0, st>>>(net->layer_w[L], net->layer_b[L], net->layer_w[L + 1], net->layer_b[L + 1],
                                                         O, net->act, S, w_img, nullptr);
    CX_LAUNCH_CHECK();
  }
  const bool fp16 = sizeof(T) == 2 && Pack<T>::kFmt == 0;
  const int64_t n_pts = (r1 - r0) * n1 * n_last;
  if (fp16 && !v1) {                               // max|dout| for the loss scale (bf16 needs none)
    prep_bwd_consts_kernel<<<MAX_O * HD / 256, 256, 0, st>>>(net->layer_b[L], net->layer_w[L + 1], O, net->act, S, consts);
    CX_LAUNCH_CHECK();
    CX_CUDA(cudaMemsetAsync(consts + CONST_SCALE + 2, 0, 4, st));
    absmax_kernel<<<(unsigned)min(cdiv(n_pts * O, 256 * 8), (int64_t)(4 * num_sms())), 256, 0, st>>>(dout, n_pts * O,
                                                                                               (unsigned*)(consts + CONST_SCALE + 2));
    CX_LAUNCH_CHECK();
    prep_scale_kernel<<<1, 256, 0, st>>>(consts, O, 1);
    CX_LAUNCH_CHECK();
  } else {
    prep_bwd_consts_unit_kernel<<<1, 256, 0, st>>>(net->layer_b[L], net->layer_w[L + 1], O, net->act, S, consts);
    CX_LAUNCH_CHECK();
  }
  // ---------------- pass W: tiles = (outer index, 128 rows of the last axis): dW, db, dWf
  // ---------------- pass A: same tiles: d(outer features)
  if (need_images) {
    prep_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbA * TILE_M * HD, 256), 256, 0, st>>>(feats[C - 1], n_last, innerA, (int64_t)pl.nbA * TILE_M);
    CX_LAUNCH_CHECK();
  }
  BwdParams a{};
  a.w_img = w_img; a.inner_img = innerA; a.inner_f32 = feats[C - 1]; a.w_f32 = net->layer_w[L]; a.consts = consts;
  a.outer0 = feats[0]; a.outer1 = C == 3 ? feats[1] : nullptr; a.n_outer1 = n1;
  a.o_begin = r0 * n1; a.o_end = r1 * n1; a.n_inner = n_last; a.nb = pl.nbA;
  a.dout = dout; a.stride_o = n_last; a.stride_j = 1;
  a.dW = grads->layer_w[L]; a.db = grads->layer_b[L]; a.dWf = grads->layer_w[L + 1];
  a.S = S;
  if (C == 3) {
    CX_CUDA(cudaMemsetAsync(drv, 0, (size_t)pl.n_oA * HD * 4, st));
    a.d_rv = drv;
  } else {
    a.d_rv = dfeat[0] + r0 * HD;
  }
  {
    const int64_t total = (int64_t)pl.nbA * pl.n_oA;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), total);
    a.trace = g_bwd_trace_pass == 0 ? g_bwd_trace : nullptr;
    a.dz_img = pl.reuse ? base + pl.dz_off : nullptr;
    if (v1) {
      CX_TRY((launch_bwd_t<T, true>(a, net->fusion_op, net->act, O, grid, st)));
    } else {
      // rowvec(o) as 16-bit rows (slab-local index): the dW pass' producers and the dfeat pass read them
      prep_rowvec16_kernel<T><<<(unsigned)cdiv(pl.n_oA * HD, 256), 256, 0, st>>>(feats[0], C == 3 ? feats[1] : nullptr, r0, n1, pl.n_oA,
                                                                                net->fusion_op, base + pl.rows_off);
      CX_LAUNCH_CHECK();
      a.outer_img = base + pl.rows_off;
      a.dz_no_wf = dz_without_wf<T>(net->act, O) ? 1 : 0;     // what the dW pass stores (and the dfeat / fused kernels expect)
      if (pl.reuse && bwd_fused_wanted(total)) {
        // ---------------- dW pass and feature-gradient pass side by side in one launch
        const int sms = num_sms();
        int pairs = env_int("COORDX_BWD_FUSED_PAIRS", (sms * 3 + 4) / 8);
        pairs = max(1, min(pairs, (sms - 1) / 2));
        a.flags = (uint32_t*)(base + pl.flags_off);
        a.fused_pairs = pairs;
        a.fused_chunk = env_int("COORDX_BWD_FUSED_CHUNK", 4);
        if (a.fused_chunk == 0) a.fused_chunk = 1;
        const int cons = sms - 2 * pairs;          // every group needs a pair and a consumer
        a.fused_groups = max(1, min(env_int("COORDX_BWD_FUSED_GROUPS", max(1, cons / 2)), min(pairs, cons)));
        a.d_inner = dfeat[C - 1];
        CX_CUDA(cudaMemsetAsync(a.flags, 0, (size_t)total * 4, st));
        if (env_int("COORDX_BWD_POISON", 0))       // debugging aid: a tile consumed before it was stored turns every gradient into NaN
          CX_CUDA(cudaMemsetAsync(base + pl.dz_off, 0xFF, (size_t)total * (4 * CH), st));
        CX_TRY(launch_bwd_fused(a, precision, net->fusion_op, net->act, O, sms, st));
        if (C == 3) {
          pair_scatter_kernel<<<(unsigned)cdiv((r1 - r0) * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 0, dfeat[0]);
          CX_LAUNCH_CHECK();
          pair_scatter_kernel<<<(unsigned)cdiv(n1 * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 1, dfeat[1]);
          CX_LAUNCH_CHECK();
        }
        return 0;
      }
      CX_TRY(launch_dw2(a, precision, net->fusion_op, net->act, O, grid, st));
      if (!pl.reuse) a.outer_img = nullptr;        // (the recompute passes index their row image differently)
    }
    a.trace = g_bwd_trace_pass == 1 ? g_bwd_trace : nullptr;
    if (pl.reuse && !v1) {
      // ---------------- dfeat pass: BOTH feature gradients from the stored dZ^T tiles, one GEMM per tile
      a.d_inner = dfeat[C - 1];
      CX_TRY(launch_dfeat(a, precision, net->fusion_op, (int)min((int64_t)num_sms(), total), st));
    } else if (pl.reuse) {
      CX_TRY((launch_drv<T>(a, net->fusion_op, grid, st)));
    } else {
      CX_TRY((launch_bwd_t<T, false>(a, net->fusion_op, net->act, O, grid, st)));
    }
  }
  if (C == 3) {
    pair_scatter_kernel<<<(unsigned)cdiv((r1 - r0) * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 0, dfeat[0]);
    CX_LAUNCH_CHECK();
    pair_scatter_kernel<<<(unsigned)cdiv(n1 * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 1, dfeat[1]);
    CX_LAUNCH_CHECK();
  }
  if (pl.reuse && !v1) return 0;
  if (pl.reuse) {
    // ---------------- din pass: gradient of the last axis' features from the stored dZ^T tiles (same tiling)
    prep_rowvec16_kernel<T><<<(unsigned)cdiv(pl.n_oA * HD, 256), 256, 0, st>>>(feats[0], C == 3 ? feats[1] : nullptr, r0, n1, pl.n_oA,
                                                                              net->fusion_op, base + pl.rows_off);
    CX_LAUNCH_CHECK();
    a.outer_img = base + pl.rows_off;
    a.d_inner = dfeat[C - 1];
    a.trace = nullptr;
    const int64_t total = (int64_t)pl.nbA * pl.n_oA;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), total);
    return launch_din<T>(a, net->fusion_op, grid, st);
  }
  // ---------------- pass B: roles swapped: tiles = (index on the last axis, 128 outer rows): d(last-axis features)
  if (C == 3)
    prep_pair_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbB * TILE_M * HD, 256), 256, 0, st>>>(feats[0], feats[1], r0, n1, pl.n_innerB, (int64_t)pl.nbB * TILE_M, net->fusion_op, innerB);
  else
    prep_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbB * TILE_M * HD, 256), 256, 0, st>>>(feats[0] + r0 * HD, pl.n_innerB, innerB, (int64_t)pl.nbB * TILE_M);
  CX_LAUNCH_CHECK();
  prep_rows16_kernel<T><<<(unsigned)cdiv(n_last * HD, 256), 256, 0, st>>>(feats[C - 1], n_last, base + pl.rows_off);
  CX_LAUNCH_CHECK();
  BwdParams b{};
  b.outer_img = base + pl.rows_off;
  b.w_img = w_img; b.inner_img = innerB; b.w_f32 = net->layer_w[L]; b.consts = consts;
  b.outer0 = feats[C - 1]; b.outer1 = nullptr; b.n_outer1 = 1;
  b.o_begin = 0; b.o_end = n_last; b.n_inner = pl.n_innerB; b.nb = pl.nbB;
  b.dout = dout; b.stride_o = 1; b.stride_j = n_last;
  b.d_rv = dfeat[C - 1];
  b.S = S;
  b.trace = g_bwd_trace_pass == 2 ? g_bwd_trace : nullptr;
  {
    const int64_t total = (int64_t)pl.nbB * n_last;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), total);
    CX_TRY((launch_bwd_t<T, false>(b, net->fusion_op, net->act, O, grid, st)));
  }
  return 0;
}

}  // namespace tcb
}  // namespace coordx

namespace coordx {
namespace tc {
int profile_set_bwd_trace(void* dev_buf, int pass) {
#if !COORDX_TRACE
  if (dev_buf) return set_err(COORDX_ERR_UNSUPPORTED, "pipeline tracing is a build option%s (scripts/build_trace_lib.sh, -DCOORDX_TRACE=1)");
#endif
  tcb::g_bwd_trace = (long long*)dev_buf;
  tcb::g_bwd_trace_pass = pass;
  return 0;
}
size_t bwd_workspace_bytes(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1, bool minimal) {
  const tcb::BwdPlan small = tcb::plan_bwd(net, n, r0, r1, false);
  if (minimal || !tcb::dz_reuse_wanted((int64_t)small.nbA * small.n_oA)) return small.total;
  return tcb::plan_bwd(net, n, r0, r1, true).total;
}
int trunk_backward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                      const float* dout, coordx_grads* grads, float* const* dfeat, int precision, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  if (precision == COORDX_PREC_BF16)
    return tcb::trunk_backward_tc_t<__nv_bfloat16>(net, feats, n, r0, r1, dout, grads, dfeat, precision, ws, ws_bytes, st);
  return tcb::trunk_backward_tc_t<__half>(net, feats, n, r0, r1, dout, grads, dfeat, precision, ws, ws_bytes, st);
}
}  // namespace tc
}  // namespace coordx
