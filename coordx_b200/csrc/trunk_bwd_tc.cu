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
#include "tc_common.cuh"

namespace coordx {
namespace tcb {
using namespace tc;

constexpr uint32_t W_CHUNK_BYTES = HD * 128;        // 32 KB: [256 rows x 128 B]
constexpr uint32_t W_BYTES = NKC * W_CHUNK_BYTES;   // 128 KB
constexpr uint32_t A_CHUNK_BYTES = TILE_M * 128;    // 16 KB
constexpr uint32_t INNER_BYTES = NKC * A_CHUNK_BYTES;  // 64 KB per 128-row block

template <typename T>
__device__ __forceinline__ T to_T(float v);
template <>
__device__ __forceinline__ __nv_bfloat16 to_T<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half to_T<__half>(float v) { return __float2half_rn(v); }

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


constexpr uint32_t CH = 16384;                    // one [128 rows x 128 B] SWIZZLE_128B chunk
constexpr uint32_t SM_W = 0;                      // 4 chunks (64 KB)
constexpr uint32_t SM_H = 4 * CH;                 // 4 chunks (64 KB)
constexpr uint32_t SM_DZ = 8 * CH;                // 2 chunks (32 KB)
constexpr uint32_t SM_INNER = 10 * CH;            // 4 chunks (64 KB)
constexpr uint32_t SM_BAR = 14 * CH;
constexpr uint32_t SM_DOUT = SM_BAR + 256;        // float dout[128 points][MAX_O] of the current tile (staged by the producers)
constexpr uint32_t SM_TOTAL_B = SM_DOUT + TILE_M * MAX_O * 4;     // 231680 bytes <= 227 KB
constexpr int BWD_THREADS = 21 * 32;
constexpr int W_E2 = 8, W_PROD = 16, W_MMA = 20;  // first warp of each role

struct BwdParams {
  const uint8_t* w_img;      // full W image of the forward prep: 4 K-chunks of [256 rows x 128 B]
  const uint8_t* inner_img;  // inner axis table, per 128-row block 4 K-chunks [128 x 128 B]
  const float* consts;       // [256] bias' then [MAX_O][256] wfs (see prep_bwd_consts_kernel)
  const float* outer0;
  const float* outer1;
  int64_t n_outer1;
  int64_t o_begin, o_end;
  int64_t n_inner;
  int nb;
  const float* dout;         // element (o_local, j, oo) at (o_local*stride_o + j*stride_j)*O + oo
  int64_t stride_o, stride_j;
  float* d_rv;               // [n_o, 256] +=   (gradient w.r.t. rowvec(o))
  float* dW;                 // [256, 256] +=   (DO_DW)
  float* db;                 // [256] +=
  float* dWf;                // [O, 256] +=
  float S;
};

// sine: bias' = S b, wfs = S Wf  (dz = dout . wfs * cos(S z + bias'));  relu: bias' = b, wfs = Wf
__global__ void __launch_bounds__(256)
prep_bwd_consts_kernel(const float* __restrict__ b, const float* __restrict__ Wf, int O, int act, float S, float* __restrict__ consts) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float sc = act == COORDX_ACT_SINE ? S : 1.f;
  if (i < HD) consts[i] = sc * b[i];
  if (i < MAX_O * HD) consts[HD + i] = (i >> 8) < O ? sc * Wf[i] : 0.f;
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

template <typename T, int OP, int ACT, int O, bool DO_DW>
__global__ void __launch_bounds__(BWD_THREADS, 1)
trunk_bwd_tc_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();          // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem_gen = smem_raw;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SM_BAR;
  const uint32_t bar_full = bar0;                  // [4] H chunk written        (128 producers)
  const uint32_t bar_empty = bar0 + 32;            // [4] H chunk no longer read (tcgen05.commit)
  const uint32_t bar_zfull = bar0 + 64;
  const uint32_t bar_zempty = bar0 + 72;           // E1 has read Z^T            (256)
  const uint32_t bar_dzfull = bar0 + 152;          // [2] E1 has written the 64-point chunk of dZ^T (128 each)
  const uint32_t bar_dsfull = bar0 + 168;          // producers have staged dout of the tile (128)
  const uint32_t bar_dzempty = bar0 + 88;          // G2/G3 have read dZ         (commit)
  const uint32_t bar_dhfull = bar0 + 96;
  const uint32_t bar_dhempty = bar0 + 104;         // E2 has read the dH piece   (256)
  const uint32_t bar_w = bar0 + 112;
  const uint32_t bar_inner = bar0 + 120;
  const uint32_t bar_dwfull = bar0 + 128;
  const uint32_t tmem_slot = bar0 + 136;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SM_BAR + 136);
  volatile int* e2_done = reinterpret_cast<volatile int*>(smem_gen + SM_BAR + 144);   // tiles fully reduced by E2 (x8 warps)

  const int h = blockIdx.x & 1;                    // half of the trunk's output features
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * cta / n_cta;
  const int64_t t_end = total * (cta + 1) / n_cta;
  const int my_tiles = (int)(t_end - t_begin);

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) {
        ptx::mbar_init(bar_full + 8 * c, 128);
        ptx::mbar_init(bar_empty + 8 * c, 1);
      }
      ptx::mbar_init(bar_zfull, 1);
      ptx::mbar_init(bar_zempty, 256);
      ptx::mbar_init(bar_dzfull, 128);
      ptx::mbar_init(bar_dzfull + 8, 128);
      ptx::mbar_init(bar_dsfull, 128);
      ptx::mbar_init(bar_dzempty, 1);
      ptx::mbar_init(bar_dhfull, 1);
      ptx::mbar_init(bar_dhempty, 256);
      ptx::mbar_init(bar_w, 1);
      ptx::mbar_init(bar_inner, 1);
      ptx::mbar_init(bar_dwfull, 1);
      *e2_done = 0;
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tZ = tmem_base, tDH = tmem_base + 128, tDW = tmem_base + 256;

  if (warp == W_MMA) {
    // ============================ MMA issuer ============================
    if (lane == 0 && my_tiles > 0) {
      ptx::mbar_arrive_expect_tx(bar_w, 4 * CH);
      for (int kc = 0; kc < 4; ++kc)
        ptx::bulk_g2s(smem_base + SM_W + kc * CH, p.w_img + (size_t)kc * W_CHUNK_BYTES + (size_t)h * CH, CH, bar_w);
      ptx::mbar_wait(bar_w, 0);
      constexpr uint32_t id_g1 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 0, 0);
      constexpr uint32_t id_g2 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 128, 1, 1);   // A = dZ (M = pts) MN-major, B = W_h (N = k_in) MN-major
      constexpr uint32_t id_g3 = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 256, 0, 1);   // A = dZ^T (M = n_out) K-major, B = H (N = k_in) MN-major
      uint32_t dh_uses = 0;
      auto issue_g2 = [&](int piece) {
        ptx::mbar_wait(bar_dhempty, (dh_uses & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {           // K = 128 n_out of this half = 16 rows of the dZ^T tile per step
          const uint64_t a = ptx::umma_desc_mn_sw128(smem_base + SM_DZ + ks * 2048, CH);
          const uint64_t b = ptx::umma_desc_mn_sw128(smem_base + SM_W + piece * 2 * CH + ks * 2048, CH);
          ptx::mma_f16_ss(tDH, a, b, id_g2, ks != 0);
        }
        ptx::mma_commit(bar_dhfull);
        ++dh_uses;
      };
      for (int t = 0; t < my_tiles; ++t) {
        // G2b(t-1), the second dH piece of the previous tile, is software-pipelined behind G1(t) -
        // except across a change of inner block: there the producers wait for E2(t-1) before they
        // overwrite the inner block, so G2b(t-1) must not wait for their H(t)
        const bool defer = t > 0 && (t_begin + t) / n_o == (t_begin + t - 1) / n_o;
        if (t > 0 && !defer) {
          issue_g2(1);
          ptx::mma_commit(bar_dzempty);
        }
        // ---- G1(t): Z^T_h = W_h H^T   (A = W_h, M = n_out; B = H, N = points)
        ptx::mbar_wait(bar_zempty, (t & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          ptx::mbar_wait(bar_full + 8 * kc, t & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(tZ, ptx::umma_desc_k_sw128(smem_base + SM_W + kc * CH + k * 32),
                            ptx::umma_desc_k_sw128(smem_base + SM_H + kc * CH + k * 32), id_g1, (kc | k) != 0);
          if (!DO_DW) ptx::mma_commit(bar_empty + 8 * kc);
        }
        ptx::mma_commit(bar_zfull);
        if (defer) {
          issue_g2(1);
          ptx::mma_commit(bar_dzempty);
        }
        // ---- G3(t): dW_h += dZ^T_h H   (K = 128 points, 16 per step; starts on the first 64-point chunk of dZ^T)
        if (DO_DW) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            ptx::mbar_wait(bar_dzfull + 8 * c, t & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int ks = c * 4 + k;
              const uint64_t a = ptx::umma_desc_k_sw128(smem_base + SM_DZ + c * CH + k * 32);
              const uint64_t b = ptx::umma_desc_mn_sw128(smem_base + SM_H + ks * 2048, CH);
              ptx::mma_f16_ss(tDW, a, b, id_g3, (t | ks) != 0);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) ptx::mma_commit(bar_empty + 8 * c);
        } else {
          ptx::mbar_wait(bar_dzfull, t & 1);
          ptx::mbar_wait(bar_dzfull + 8, t & 1);
          ptx::tc_fence_after();
        }
        // ---- G2a(t)
        issue_g2(0);
      }
      issue_g2(1);
      ptx::mma_commit(bar_dzempty);
      ptx::mma_commit(bar_dwfull);
    }
    __syncwarp();
  } else if (warp >= W_PROD) {
    // ============================ producers: H tile ============================
    const int ptid = threadIdx.x - W_PROD * 32;
    const int g8 = ptid & 7, r0 = ptid >> 3;
    const uint32_t piece_off = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((g8 ^ (r0 & 7)) << 4);
    uint32_t inner_phase = 0;
    int64_t cur_jb = -1;
    uint32_t rv[NKC][4];
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tt = t_begin + t;
      const int64_t jb = tt / n_o;
      const int64_t o = p.o_begin + (tt - jb * n_o);
      {
        const int64_t i0 = o / p.n_outer1, i1 = o - i0 * p.n_outer1;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const int k0 = kc * KCHUNK + g8 * 8;
          float4 x0 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + k0);
          float4 x1 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + k0 + 4);
          if (p.outer1) {
            const float4 y0 = *reinterpret_cast<const float4*>(p.outer1 + i1 * HD + k0);
            const float4 y1 = *reinterpret_cast<const float4*>(p.outer1 + i1 * HD + k0 + 4);
            if (OP == COORDX_FUSE_PROD) {
              x0.x *= y0.x; x0.y *= y0.y; x0.z *= y0.z; x0.w *= y0.w;
              x1.x *= y1.x; x1.y *= y1.y; x1.z *= y1.z; x1.w *= y1.w;
            } else {
              x0.x += y0.x; x0.y += y0.y; x0.z += y0.z; x0.w += y0.w;
              x1.x += y1.x; x1.y += y1.y; x1.z += y1.z; x1.w += y1.w;
            }
          }
          rv[kc][0] = Pack<T>::pack(x0.x, x0.y);
          rv[kc][1] = Pack<T>::pack(x0.z, x0.w);
          rv[kc][2] = Pack<T>::pack(x1.x, x1.y);
          rv[kc][3] = Pack<T>::pack(x1.z, x1.w);
        }
      }
      // dout of this thread's point (row ptid of the tile): loaded now, staged for E1 after the H chunks
      float dreg[O];
      {
        const int64_t j = jb * TILE_M + ptid;
        const int64_t o_local = tt - jb * n_o;
#pragma unroll
        for (int oo = 0; oo < O; ++oo)
          dreg[oo] = j < p.n_inner ? __ldg(p.dout + (o_local * p.stride_o + j * p.stride_j) * O + oo) : 0.f;
      }
      if (jb != cur_jb) {
        ptx::named_bar_sync(1, 128);
        if (ptid == 0) {
          if (OP == COORDX_FUSE_PROD) {            // E2 multiplies dH by the inner block: wait until it is done with it
            while (*e2_done < 8 * t) {
            }
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive_expect_tx(bar_inner, INNER_BYTES);
          for (uint32_t c = 0; c < INNER_BYTES; c += CH)
            ptx::bulk_g2s(smem_base + SM_INNER + c, p.inner_img + (size_t)jb * INNER_BYTES + c, CH, bar_inner);
        }
        ptx::mbar_wait(bar_inner, inner_phase);
        inner_phase ^= 1;
        cur_jb = jb;
      }
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
        ptx::mbar_wait(bar_empty + 8 * kc, (t & 1) ^ 1);
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
        ptx::fence_proxy_async();
        ptx::mbar_arrive(bar_full + 8 * kc);
      }
      // E1(t-1) has read its dout (it arrives on zempty after its last read); G1(t) runs meanwhile
      ptx::mbar_wait(bar_zempty, (t & 1) ^ 1);
      {
        float* ds = reinterpret_cast<float*>(smem_gen + SM_DOUT) + ptid * O;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) ds[oo] = dreg[oo];
      }
      ptx::mbar_arrive(bar_dsfull);
    }
  } else if (warp < W_E2) {
    // ============================ E1: Z^T_h -> dZ^T_h (+ db, dWf per-thread sums) ============================
    // thread = output feature n = 128 h + row; warp (q, hh) covers the 64 points [64 hh, 64 hh + 64) of the tile
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;
    const float bias_n = p.consts[h * 128 + row];                 // sine: S b[n]; relu: b[n]
    float wfs_n[O];                                               // sine: S Wf[o][n]; relu: Wf[o][n]
#pragma unroll
    for (int oo = 0; oo < O; ++oo) wfs_n[oo] = p.consts[HD + oo * HD + h * 128 + row];
    float db_acc = 0.f;
    float dwf_acc[O];
#pragma unroll
    for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = 0.f;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tt = t_begin + t;
      const int64_t jb = tt / n_o;
      const int64_t o_local = tt - jb * n_o;
      const float* ds = reinterpret_cast<const float*>(smem_gen + SM_DOUT) + 64 * hh * O;   // this warp's 64 points
      ptx::mbar_wait(bar_zfull, t & 1);
      ptx::tc_fence_after();
      ptx::mbar_wait(bar_dsfull, t & 1);
      uint8_t* dzrow = smem_gen + SM_DZ + hh * CH;                // 64 points = one 128-byte row of this chunk
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        // d0[i] = dout[p_i, 0]  (O == 1)   |   g[i] = sum_o dout[p_i, o] * wfs[o][n]  (O > 1)
        // (loads in flight while the TMEM load completes)
        float g[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float* dp = ds + (b * 32 + i) * O;                  // warp-uniform shared-memory address: broadcast
          if (O == 1) {
            g[i] = dp[0];
          } else {
            float acc = 0.f;
#pragma unroll
            for (int oo = 0; oo < O; ++oo) acc = fmaf(dp[oo], wfs_n[oo], acc);
            g[i] = acc;
          }
        }
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tZ + ((uint32_t)(q * 32) << 16) + hh * 64 + b * 32, r);
        ptx::tmem_ld_wait();
        if (b == 1) {                              // every Z^T column of this warp is in registers
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_zempty);
        }
        float dzv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float z = __uint_as_float(r[i]);
          const float gw = O == 1 ? g[i] * wfs_n[0] : g[i];
          float av;                                // activation value (for dWf), derivative folded into dzv
          if (ACT == COORDX_ACT_SINE) {
            const float th = fmaf(z, p.S, bias_n);
            dzv[i] = gw * __cosf(th);
            av = DO_DW ? __sinf(th) : 0.f;
          } else {
            const float pre = z + bias_n;
            dzv[i] = pre > 0.f ? gw : 0.f;
            av = fmaxf(pre, 0.f);
          }
          if (DO_DW) {
            db_acc += dzv[i];
            if (O == 1) {
              dwf_acc[0] = fmaf(g[i], av, dwf_acc[0]);
            } else {
              const float* dp = ds + (b * 32 + i) * O;
#pragma unroll
              for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = fmaf(dp[oo], av, dwf_acc[oo]);
            }
          }
        }
        // points 32b .. 32b+31 of row n -> 4 swizzled 16-byte pieces (the tensor core must be done with dZ(t-1))
        if (b == 0) ptx::mbar_wait(bar_dzempty, (t & 1) ^ 1);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          uint4 w;
          w.x = Pack<T>::pack(dzv[pc * 8 + 0], dzv[pc * 8 + 1]);
          w.y = Pack<T>::pack(dzv[pc * 8 + 2], dzv[pc * 8 + 3]);
          w.z = Pack<T>::pack(dzv[pc * 8 + 4], dzv[pc * 8 + 5]);
          w.w = Pack<T>::pack(dzv[pc * 8 + 6], dzv[pc * 8 + 7]);
          *reinterpret_cast<uint4*>(dzrow + ptx::sw128_offset(row, b * 32 + pc * 8)) = w;
        }
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(bar_dzfull + 8 * hh);
    }
    if (DO_DW && my_tiles > 0) {
      atomicAdd(p.db + h * 128 + row, db_acc);
#pragma unroll
      for (int oo = 0; oo < O; ++oo) atomicAdd(p.dWf + oo * HD + h * 128 + row, dwf_acc[oo]);
    }
  } else {
    // ============================ E2: dH pieces -> d rowvec ============================
    // warp (q, eh): lane quadrant q (32 points), columns [64 eh, 64 eh + 64) of every 128-column dH piece
    const int q = warp & 3, eh = (warp - W_E2) >> 2;
    const int row = q * 32 + lane;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tt = t_begin + t;
      const int64_t jb = tt / n_o;
      const int64_t o_local = tt - jb * n_o;
      float* dst = p.d_rv + o_local * HD;
#pragma unroll 1
      for (int piece = 0; piece < 2; ++piece) {
        const uint32_t u = 2u * (uint32_t)t + (uint32_t)piece;
        ptx::mbar_wait(bar_dhfull, u & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int b = 2 * eh; b < 2 * eh + 2; ++b) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tDH + ((uint32_t)(q * 32) << 16) + b * 32, r);
          ptx::tmem_ld_wait();
          if (b == 2 * eh + 1) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_dhempty);
          }
          const int k0 = piece * 128 + b * 32;     // k_in of column 0 of this batch
          float v[32];
          if (OP == COORDX_FUSE_PROD) {
            const uint8_t* irow = smem_gen + SM_INNER + (k0 >> 6) * CH;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              const uint4 qv = *reinterpret_cast<const uint4*>(irow + ptx::sw128_offset(row, (k0 & 63) + pc * 8));
              float f[8];
              unpack8<T>(qv, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[pc * 8 + e] = __uint_as_float(r[pc * 8 + e]) * f[e];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          }
          const float s = warp_transpose_reduce(v, lane);
          atomicAdd(dst + k0 + lane, s);
        }
      }
      if (OP == COORDX_FUSE_PROD) {
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<int*>(e2_done), 1);
      }
    }
  }

  // ---- flush dW_h (accumulated over every tile of this CTA) with one atomic per element
  if (DO_DW && warp < 8 && my_tiles > 0) {
    ptx::mbar_wait(bar_dwfull, 0);
    ptx::tc_fence_after();
    const int q = warp & 3, ch = warp >> 2;
    float* drow = p.dW + (size_t)(h * 128 + q * 32 + lane) * HD + ch * 128;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tDW + ((uint32_t)(q * 32) << 16) + ch * 128 + b * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(drow + b * 32 + i, __uint_as_float(r[i]));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
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
  auto kern = trunk_bwd_tc_kernel<T, OP, ACT, O, DO_DW>;
  static bool attr_done = false;
  if (!attr_done) {
    CX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL_B));
    attr_done = true;
  }
  kern<<<grid, BWD_THREADS, SM_TOTAL_B, st>>>(bp);
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

struct BwdPlan {
  size_t w_img_off, epi_off, consts_off, innerA_off, innerB_off, drv_off, total;
  int nbA, nbB;
  int64_t n_innerA, n_innerB, n_oA;
};

static BwdPlan plan_bwd(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1) {
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
  pl.consts_off = off; off += align_up((1 + MAX_O) * HD * 4, 1024);
  pl.innerA_off = off; off += (size_t)pl.nbA * INNER_BYTES;
  pl.innerB_off = off; off += (size_t)pl.nbB * INNER_BYTES;
  pl.drv_off = off; off += C == 3 ? align_up((size_t)pl.n_oA * HD * 4, 1024) : 0;
  pl.total = off;
  return pl;
}

template <typename T>
static int trunk_backward_tc_t(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                               const float* dout, coordx_grads* grads, float* const* dfeat, void* ws, size_t ws_bytes,
                               cudaStream_t st) {
  const int C = net->n_branches, L = net->n_pre, O = net->layer_out[L + 1];
  const BwdPlan pl = plan_bwd(net, n, r0, r1);
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

  prep_weights_kernel<T><<<HD * HD / 256, 256, 0, st>>>(net->layer_w[L], net->layer_b[L], net->layer_w[L + 1], net->layer_b[L + 1],
                                                       O, net->act, S, w_img, nullptr);
  CX_LAUNCH_CHECK();
  prep_bwd_consts_kernel<<<MAX_O * HD / 256, 256, 0, st>>>(net->layer_b[L], net->layer_w[L + 1], O, net->act, S, consts);
  CX_LAUNCH_CHECK();
  // ---------------- pass A: tiles = (outer index, 128 rows of the last axis): dW, db, dWf, d(outer features)
  prep_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbA * TILE_M * HD, 256), 256, 0, st>>>(feats[C - 1], n_last, innerA, (int64_t)pl.nbA * TILE_M);
  CX_LAUNCH_CHECK();
  BwdParams a{};
  a.w_img = w_img; a.inner_img = innerA; a.consts = consts;
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
    CX_TRY((launch_bwd_t<T, true>(a, net->fusion_op, net->act, O, grid, st)));
  }
  if (C == 3) {
    pair_scatter_kernel<<<(unsigned)cdiv((r1 - r0) * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 0, dfeat[0]);
    CX_LAUNCH_CHECK();
    pair_scatter_kernel<<<(unsigned)cdiv(n1 * HD, 256), 256, 0, st>>>(drv, feats[0], feats[1], r0, r1 - r0, n1, net->fusion_op, 1, dfeat[1]);
    CX_LAUNCH_CHECK();
  }
  // ---------------- pass B: roles swapped: tiles = (index on the last axis, 128 outer rows): d(last-axis features)
  if (C == 3)
    prep_pair_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbB * TILE_M * HD, 256), 256, 0, st>>>(feats[0], feats[1], r0, n1, pl.n_innerB, (int64_t)pl.nbB * TILE_M, net->fusion_op, innerB);
  else
    prep_inner_kernel<T><<<(unsigned)cdiv((int64_t)pl.nbB * TILE_M * HD, 256), 256, 0, st>>>(feats[0] + r0 * HD, pl.n_innerB, innerB, (int64_t)pl.nbB * TILE_M);
  CX_LAUNCH_CHECK();
  BwdParams b{};
  b.w_img = w_img; b.inner_img = innerB; b.consts = consts;
  b.outer0 = feats[C - 1]; b.outer1 = nullptr; b.n_outer1 = 1;
  b.o_begin = 0; b.o_end = n_last; b.n_inner = pl.n_innerB; b.nb = pl.nbB;
  b.dout = dout; b.stride_o = 1; b.stride_j = n_last;
  b.d_rv = dfeat[C - 1];
  b.S = S;
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
size_t bwd_workspace_bytes(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1) { return tcb::plan_bwd(net, n, r0, r1).total; }
int trunk_backward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                      const float* dout, coordx_grads* grads, float* const* dfeat, int precision, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  if (precision == COORDX_PREC_BF16)
    return tcb::trunk_backward_tc_t<__nv_bfloat16>(net, feats, n, r0, r1, dout, grads, dfeat, ws, ws_bytes, st);
  return tcb::trunk_backward_tc_t<__half>(net, feats, n, r0, r1, dout, grads, dfeat, ws, ws_bytes, st);
}
}  // namespace tc
}  // namespace coordx
