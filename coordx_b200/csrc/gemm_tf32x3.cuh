// FP32-accurate GEMM on the tcgen05 tensor cores: 3xTF32 split.
//
//   Y[M,N] = act( X[M,K] . W[N,K]^T + bias )          (the NT form of BatchLinear.forward, modules.py:26-27)
//
// Every operand element v is split into hi = rn_tf32(v) (10-bit mantissa, exactly representable) and lo = v - hi
// (exact in fp32, |lo| <= 2^-11 |v|); the product is accumulated as hi*hi + (hi*lo + lo*hi) in fp32 TMEM
// accumulators (kind::tf32 reads the lo operands with 11 significant bits: error <= 2^-22 |v|; the dropped lo*lo term
// is <= 2^-22 of the product, zero-mean).  The tensor core TRUNCATES its fp32 accumulator after every instruction
// (measured: with all three products in one accumulator, K = 1024 = 384 truncations of the full-size sum, the error
// was 7x an FFMA GEMM's), so the two cross terms go to a SECOND accumulator (columns 256..511): its truncations are
// 2^-11 smaller, the main accumulator sees K/8 truncations instead of 3K/8, and the epilogue adds the two in fp32.
// ~10x the FFMA kernel's throughput on the large per-axis GEMMs: the 262 144 x 1024 x 1024 layers of the video net's
// xy branch, the U/V tables and the output GEMM of the separable sum path.
//
// One 128 x BN output tile per CTA (BN = 256 for the large GEMMs; BN = 64 for the per-axis branch layers, M = a few
// thousand rows and N = 256, so that they spread over 4x more SMs), K in chunks of 32 (= one 128-byte SWIZZLE_128B
// row of fp32).  No operand image
// in global memory and no TMA descriptor: 8 producer warps read X and W rows with coalesced 16-byte loads (one chunk
// ahead, in registers), split hi/lo and store the four K-major tiles (A_hi, A_lo, B_hi, B_lo: 96 KB per stage, 2
// stages); one elected thread issues 12 kind::tf32 MMAs (M128 x N256 x K8) per chunk; the producer warps drain the
// accumulator at the end (bias, activation, optional pre-activation output, strided columns).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace coordx {
namespace g3 {

constexpr int BM = 128, BK = 32;
constexpr uint32_t A_T = BM * 128;                       // 16 KB: [rows x 128 B] SWIZZLE_128B
constexpr int STAGES = 2;
constexpr int THREADS = 9 * 32;                          // 8 producer / epilogue warps + the MMA issuer
template <int BN>
struct Cfg {
  static constexpr uint32_t B_T = BN * 128;              // 32 KB (BN 256) / 8 KB (BN 64)
  static constexpr uint32_t STAGE = 2 * A_T + 2 * B_T;   // A_hi | A_lo | B_hi | B_lo
  static constexpr uint32_t SM_BAR = STAGES * STAGE;
  static constexpr uint32_t SM_TOTAL = SM_BAR + 128;
  static constexpr uint32_t TMEM_COLS = 2 * BN;          // main + cross-term accumulator (power of two >= 32)
};

struct Args {
  const float* A;      // X [M, K], row stride lda (16-byte aligned rows)
  const float* B;      // W [N, K], row stride ldb
  float* Y;            // activated output (nullable)
  float* Z;            // pre-activation output (nullable)
  const float* bias;   // [N] (nullable), scaled by bscale
  const float* bias1;  // one value added to every element (nullable)
  int64_t lda, ldb, ldc, cs;     // Y/Z element (m, n) at m * ldc + n * cs
  int M, N, K;
  float bscale;
  int act;             // COORDX_ACT_* or -1
  float S;
  const float* zprev;  // dX form (nullable): the result is multiplied by act'(zprev[m * ldz + n]) and stored to Y
  int64_t ldz;
  // weight-gradient form (dw_tensor): the K range is split over gridDim.z CTAs (`ksplits`, every split non-empty) and
  // the plain product is ADDED to Y with atomics (no bias, activation or Z output)
  int ksplits;
  int atomic;
};

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_tf32_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float act_fwd_inl(int act, float S, float z) { return act == COORDX_ACT_SINE ? sin_inl(S * z) : fmaxf(z, 0.f); }
__device__ __forceinline__ float act_grad_inl(int act, float S, float z) {
  return act == COORDX_ACT_SINE ? S * cos_inl(S * z) : (z > 0.f ? 1.f : 0.f);
}

__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ void split_store(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float4& v) {
  float4 h, l;
  h.x = rn_tf32(v.x); h.y = rn_tf32(v.y); h.z = rn_tf32(v.z); h.w = rn_tf32(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const Args g) {
  constexpr uint32_t B_T = Cfg<BN>::B_T, STAGE = Cfg<BN>::STAGE, SM_BAR = Cfg<BN>::SM_BAR, TMEM_COLS = Cfg<BN>::TMEM_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SM_BAR;
  const uint32_t b_full = bar0, b_empty = bar0 + 16, b_acc = bar0 + 32, b_slot = bar0 + 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SM_BAR + 40);
  constexpr int W_MMA = 8;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int nk = g.K / BK, kc0 = 0;                      // this CTA's K chunks: [kc0, kc0 + nk)
  if (g.ksplits > 1) {
    const int per = (nk + g.ksplits - 1) / g.ksplits;
    kc0 = (int)blockIdx.z * per;
    nk = min(per, nk - kc0);
  }

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(b_full + 8 * s, 8);
        ptx::mbar_init(b_empty + 8 * s, 1);
      }
      ptx::mbar_init(b_acc, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(b_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tacc = *tmem_slot_gen;

  if (warp == W_MMA) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(2 /* TF32 */, BM, BN, 0, 0);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc & 1;
        ptx::mbar_wait(b_full + 8 * s, (kc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t st = smem_base + s * STAGE;
        const uint32_t a_hi = st, a_lo = st + A_T, b_hi = st + 2 * A_T, b_lo = st + 2 * A_T + B_T;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {             // 8 tf32 = 32 bytes per K-step
          const uint64_t dah = ptx::umma_desc_k_sw128(a_hi + ks * 32), dal = ptx::umma_desc_k_sw128(a_lo + ks * 32);
          const uint64_t dbh = ptx::umma_desc_k_sw128(b_hi + ks * 32), dbl = ptx::umma_desc_k_sw128(b_lo + ks * 32);
          mma_tf32_ss(tacc + BN, dal, dbh, idesc, (kc | ks) != 0);      // cross terms: their own accumulator
          mma_tf32_ss(tacc + BN, dah, dbl, idesc, 1);
          mma_tf32_ss(tacc, dah, dbh, idesc, (kc | ks) != 0);
        }
        ptx::mma_commit(b_empty + 8 * s);
      }
      ptx::mma_commit(b_acc);
    }
    __syncwarp();
  } else {
    // ============================ producers ============================
    const int tid = threadIdx.x;                     // 0..255
    const int piece = tid & 7, rbase = tid >> 3;     // 16-byte piece of the 128-byte row; rows rbase + 32 i
    // two register sets: the global loads of chunk kc + 2 are issued right after chunk kc has been written to shared
    // memory and are not needed before chunk kc + 1 has been written too - a whole chunk period to cover the L2 latency
    float4 ra[2][4], rb[2][BN / 32];
    auto gload = [&](int set, int kc) {
      const int k0 = (kc0 + kc) * BK + piece * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + rbase + 32 * i;
        ra[set][i] = m < g.M ? __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)m * g.lda + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        const int n = n0 + rbase + 32 * i;
        rb[set][i] = n < g.N ? __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)n * g.ldb + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto produce = [&](int set, int kc) {
      const int s = kc & 1;
      ptx::mbar_wait(b_empty + 8 * s, ((kc >> 1) & 1) ^ 1);
      uint8_t* st = smem_gen + s * STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = rbase + 32 * i;
        split_store(st, st + A_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), ra[set][i]);
      }
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        const int row = rbase + 32 * i;
        split_store(st + 2 * A_T, st + 2 * A_T + B_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), rb[set][i]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(b_full + 8 * s);
      if (kc + 2 < nk) gload(set, kc + 2);
    };
    gload(0, 0);
    if (nk > 1) gload(1, 1);
    for (int kc = 0; kc < nk; kc += 2) {
      produce(0, kc);
      if (kc + 1 < nk) produce(1, kc + 1);
    }
    // ============================ epilogue ============================
    ptx::mbar_wait(b_acc, 0);
    ptx::tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const float add1 = g.bias1 ? __ldg(g.bias1) : 0.f;
    const bool vec = g.cs == 1 && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.Y) | reinterpret_cast<uintptr_t>(g.Z)) & 15) == 0;
    // Rolled loop over groups of 8 columns (static indices inside a group): 8 independent sinf / cosf chains are enough
    // to fill the issue slots of the 2 warps per scheduler, and the epilogue stays a few hundred instructions - these
    // kernels mostly start with a cold instruction cache (short launches between L2-flushing phases).
    const int ncol = BN / 2;                             // this warp: lane quadrant q, column half `half`
    const bool zvec = g.zprev && (g.ldz & 3) == 0 && (reinterpret_cast<uintptr_t>(g.zprev) & 15) == 0;
#pragma unroll 1
    for (int c8 = 0; c8 < ncol; c8 += 8) {
      uint32_t r[8], r2[8];
      ptx::tmem_ld_32x32b_x8(tacc + ((uint32_t)(q * 32) << 16) + half * ncol + c8, r);
      ptx::tmem_ld_32x32b_x8(tacc + ((uint32_t)(q * 32) << 16) + BN + half * ncol + c8, r2);
      ptx::tmem_ld_wait();
      const int nb = n0 + half * ncol + c8;
      if (m < g.M && nb < g.N) {
        const bool full = nb + 8 <= g.N;
        float v[8];
        if (g.atomic) {                                 // split-K weight gradient: Y += partial product
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (nb + i < g.N) atomicAdd(&g.Y[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs], __uint_as_float(r[i]) + __uint_as_float(r2[i]));
          continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = (__uint_as_float(r[i]) + __uint_as_float(r2[i])) + add1 + ((g.bias && nb + i < g.N) ? __ldg(g.bias + nb + i) * g.bscale : 0.f);
        if (g.zprev) {                                  // dX = (dZ W) * act'(Zprev)
          const float* zp = g.zprev + (int64_t)m * g.ldz + nb;
          float z[8];
          if (zvec && full) {
            const float4 z0 = __ldg(reinterpret_cast<const float4*>(zp)), z1 = __ldg(reinterpret_cast<const float4*>(zp) + 1);
            z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = nb + i < g.N ? __ldg(zp + i) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= act_grad_inl(g.act, g.S, z[i]);
        }
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = (g.act >= 0 && !g.zprev && g.Y) ? act_fwd_inl(g.act, g.S, v[i]) : v[i];
        if (vec && full) {
          if (g.Z) {
            float4* zo = reinterpret_cast<float4*>(g.Z + (int64_t)m * g.ldc + nb);
            zo[0] = make_float4(v[0], v[1], v[2], v[3]);
            zo[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (g.Y) {
            float4* yo = reinterpret_cast<float4*>(g.Y + (int64_t)m * g.ldc + nb);
            yo[0] = make_float4(y[0], y[1], y[2], y[3]);
            yo[1] = make_float4(y[4], y[5], y[6], y[7]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {                  // (static indices: the arrays stay in registers)
            if (nb + i < g.N) {
              if (g.Z) g.Z[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs] = v[i];
              if (g.Y) g.Y[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs] = y[i];
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tacc, TMEM_COLS);
  }
}

// The 128 x 256 tile as a CTA PAIR (cluster of two CTAs along M): one cta_group::2 MMA M256 x N256 x K8 per step, every
// CTA supplying its own 128 rows of A and HALF of the W tile (128 of the 256 output columns).  The single-CTA kernel is
// bound by shared memory (profiles/r01f_gemm_tf32x3_full.txt: 36 KB of operand fetch per K8 step at the ~70 B/clk
// fetch limit plus the 96 KB of split stores per chunk); per CTA the pair fetches a third less and its producers split and
// store half of W.  Each CTA keeps the accumulator rows of its own 128 rows (all 256 columns) and drains them itself.
struct PairCfg {
  static constexpr int BN = 256;
  static constexpr uint32_t B_T = 128 * 128;             // 16 KB: this CTA's 128 columns of the W tile
  static constexpr uint32_t STAGE = 2 * A_T + 2 * B_T;   // A_hi | A_lo | B_hi | B_lo = 64 KB
  static constexpr int STAGES = 3;
  static constexpr uint32_t SM_BAR = STAGES * STAGE;
  static constexpr uint32_t SM_TOTAL = SM_BAR + 128;
  static constexpr uint32_t TMEM_COLS = 512;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tf32x3_pair_kernel(const Args g) {
  constexpr int BN = PairCfg::BN, STAGES = PairCfg::STAGES;
  constexpr uint32_t B_T = PairCfg::B_T, STAGE = PairCfg::STAGE, SM_BAR = PairCfg::SM_BAR, TMEM_COLS = PairCfg::TMEM_COLS;
  const uint32_t rank = ptx::cluster_ctarank();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SM_BAR;
  const uint32_t b_full = bar0, b_empty = bar0 + 32, b_acc = bar0 + 64, b_slot = bar0 + 72;      // full: leader only (16 warps)
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SM_BAR + 72);
  constexpr int W_MMA = 8;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;     // (pairs along x, like the forward trunk)
  int nk = g.K / BK, kc0 = 0;                      // this CTA's K chunks: [kc0, kc0 + nk)
  if (g.ksplits > 1) {
    const int per = (nk + g.ksplits - 1) / g.ksplits;
    kc0 = (int)blockIdx.z * per;
    nk = min(per, nk - kc0);
  }

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(b_full + 8 * s, 16);        // 8 producer warps of each CTA
        ptx::mbar_init(b_empty + 8 * s, 1);
      }
      ptx::mbar_init(b_acc, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc_2cta(b_slot, TMEM_COLS);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                             // the peer's barriers and tensor memory exist
  ptx::tc_fence_after();
  const uint32_t tacc = *tmem_slot_gen;

  if (warp == W_MMA) {
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(2 /* TF32 */, 2 * BM, BN, 0, 0);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % STAGES;
        ptx::mbar_wait(b_full + 8 * s, (kc / STAGES) & 1);
        ptx::tc_fence_after();
        const uint32_t st = smem_base + s * STAGE;
        const uint32_t a_hi = st, a_lo = st + A_T, b_hi = st + 2 * A_T, b_lo = st + 2 * A_T + B_T;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {             // 8 tf32 = 32 bytes per K-step
          const uint64_t dah = ptx::umma_desc_k_sw128(a_hi + ks * 32), dal = ptx::umma_desc_k_sw128(a_lo + ks * 32);
          const uint64_t dbh = ptx::umma_desc_k_sw128(b_hi + ks * 32), dbl = ptx::umma_desc_k_sw128(b_lo + ks * 32);
          mma_tf32_ss_2cta(tacc + BN, dal, dbh, idesc, (kc | ks) != 0);      // cross terms: their own accumulator
          mma_tf32_ss_2cta(tacc + BN, dah, dbl, idesc, 1);
          mma_tf32_ss_2cta(tacc, dah, dbh, idesc, (kc | ks) != 0);
        }
        ptx::mma_commit_2cta(b_empty + 8 * s, 3);
      }
      ptx::mma_commit_2cta(b_acc, 3);
    }
    __syncwarp();
  } else {
    // ============================ producers ============================
    const int tid = threadIdx.x;                     // 0..255
    const int piece = tid & 7, rbase = tid >> 3;     // 16-byte piece of the 128-byte row; rows rbase + 32 i
    // two register sets: the global loads of chunk kc + 2 are issued right after chunk kc has been written to shared
    // memory and are not needed before chunk kc + 1 has been written too - a whole chunk period to cover the L2 latency
    const uint32_t full_remote = ptx::mapa(b_full, 0);
    const int nb0 = n0 + (int)rank * 128;            // this CTA's half of the W tile
    float4 ra[2][4], rb[2][4];
    auto gload = [&](int set, int kc) {
      const int k0 = (kc0 + kc) * BK + piece * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + rbase + 32 * i;
        ra[set][i] = m < g.M ? __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)m * g.lda + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = nb0 + rbase + 32 * i;
        rb[set][i] = n < g.N ? __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)n * g.ldb + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto produce = [&](int set, int kc) {
      const int s = kc % STAGES;
      ptx::mbar_wait(b_empty + 8 * s, ((kc / STAGES) & 1) ^ 1);
      uint8_t* st = smem_gen + s * STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = rbase + 32 * i;
        split_store(st, st + A_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), ra[set][i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = rbase + 32 * i;
        split_store(st + 2 * A_T, st + 2 * A_T + B_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), rb[set][i]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(full_remote + 8 * s);
      if (kc + 2 < nk) gload(set, kc + 2);
    };
    gload(0, 0);
    if (nk > 1) gload(1, 1);
    for (int kc = 0; kc < nk; kc += 2) {
      produce(0, kc);
      if (kc + 1 < nk) produce(1, kc + 1);
    }
    // ============================ epilogue ============================
    ptx::mbar_wait(b_acc, 0);
    ptx::tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const float add1 = g.bias1 ? __ldg(g.bias1) : 0.f;
    const bool vec = g.cs == 1 && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.Y) | reinterpret_cast<uintptr_t>(g.Z)) & 15) == 0;
    // Rolled loop over groups of 8 columns (static indices inside a group): 8 independent sinf / cosf chains are enough
    // to fill the issue slots of the 2 warps per scheduler, and the epilogue stays a few hundred instructions - these
    // kernels mostly start with a cold instruction cache (short launches between L2-flushing phases).
    const int ncol = BN / 2;                             // this warp: lane quadrant q, column half `half`
    const bool zvec = g.zprev && (g.ldz & 3) == 0 && (reinterpret_cast<uintptr_t>(g.zprev) & 15) == 0;
#pragma unroll 1
    for (int c8 = 0; c8 < ncol; c8 += 8) {
      uint32_t r[8], r2[8];
      ptx::tmem_ld_32x32b_x8(tacc + ((uint32_t)(q * 32) << 16) + half * ncol + c8, r);
      ptx::tmem_ld_32x32b_x8(tacc + ((uint32_t)(q * 32) << 16) + BN + half * ncol + c8, r2);
      ptx::tmem_ld_wait();
      const int nb = n0 + half * ncol + c8;
      if (m < g.M && nb < g.N) {
        const bool full = nb + 8 <= g.N;
        float v[8];
        if (g.atomic) {                                 // split-K weight gradient: Y += partial product
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (nb + i < g.N) atomicAdd(&g.Y[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs], __uint_as_float(r[i]) + __uint_as_float(r2[i]));
          continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = (__uint_as_float(r[i]) + __uint_as_float(r2[i])) + add1 + ((g.bias && nb + i < g.N) ? __ldg(g.bias + nb + i) * g.bscale : 0.f);
        if (g.zprev) {                                  // dX = (dZ W) * act'(Zprev)
          const float* zp = g.zprev + (int64_t)m * g.ldz + nb;
          float z[8];
          if (zvec && full) {
            const float4 z0 = __ldg(reinterpret_cast<const float4*>(zp)), z1 = __ldg(reinterpret_cast<const float4*>(zp) + 1);
            z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = nb + i < g.N ? __ldg(zp + i) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= act_grad_inl(g.act, g.S, z[i]);
        }
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = (g.act >= 0 && !g.zprev && g.Y) ? act_fwd_inl(g.act, g.S, v[i]) : v[i];
        if (vec && full) {
          if (g.Z) {
            float4* zo = reinterpret_cast<float4*>(g.Z + (int64_t)m * g.ldc + nb);
            zo[0] = make_float4(v[0], v[1], v[2], v[3]);
            zo[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (g.Y) {
            float4* yo = reinterpret_cast<float4*>(g.Y + (int64_t)m * g.ldc + nb);
            yo[0] = make_float4(y[0], y[1], y[2], y[3]);
            yo[1] = make_float4(y[4], y[5], y[6], y[7]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {                  // (static indices: the arrays stay in registers)
            if (nb + i < g.N) {
              if (g.Z) g.Z[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs] = v[i];
              if (g.Y) g.Y[(int64_t)m * g.ldc + (int64_t)(nb + i) * g.cs] = y[i];
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                             // both CTAs are done with each other's barriers / operands
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tacc, TMEM_COLS);
  }
}


// worth it (and legal) for this problem?  K a multiple of 32, 16-byte aligned operand rows, and enough work: the
// 128 x 256 tile for GEMMs that fill the chip with it, the 128 x 64 tile for the per-axis layers (M >= 256 rows).
// Returns the N tile width to use, 0 = take the FFMA kernel.  COORDX_FP32_GEMM=ffma forces the FFMA kernel.
static int tile_n(const Args& a) {
  const char* e = getenv("COORDX_FP32_GEMM");
  if (e && strcmp(e, "ffma") == 0) return 0;
  if (a.K < BK || a.K % BK != 0 || a.N < 64) return 0;
  // ReLU layers stay on the FFMA kernel: its pre-activations differ from the reference's by ~1e-7 relative, the 3xTF32
  // ones by ~2e-6, and every pre-activation that changes sign flips a ReLU gate (a 1/P-sized jump in the gradients)
  if (a.act == COORDX_ACT_RELU && !a.zprev) return 0;
  if (a.act > COORDX_ACT_RELU) return 0;          // sigmoid / tanh / selu / softplus / elu: FFMA kernels only
  if ((a.lda & 3) || (a.ldb & 3)) return 0;
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B)) & 15) return 0;
  const int64_t work = (int64_t)a.M * a.N * a.K;
  if (a.N >= 128 && cdiv(a.M, BM) * cdiv(a.N, 256) >= 32 && work >= ((int64_t)1 << 28)) return 256;
  if (a.M >= 96 && a.K >= 128 && work >= ((int64_t)1 << 24)) return 64;
  return 0;
}
static bool usable(const Args& a) { return tile_n(a) != 0; }

// COORDX_G3_PAIR=0: single-CTA kernel for the 128 x 256 tile as well (A/B)
static bool pair_wanted() {
  const char* e = getenv("COORDX_G3_PAIR");
  return !(e && e[0] == '0');
}

static int launch_pair(const Args& a, cudaStream_t st) {
  static AttrOnce once;
  CX_TRY(once.set(gemm_tf32x3_pair_kernel, (int)PairCfg::SM_TOTAL));
  const unsigned my = (unsigned)cdiv(a.M, BM);
  dim3 grid((my + 1) & ~1u, (unsigned)cdiv(a.N, 256), (unsigned)max(a.ksplits, 1));      // row tiles in pairs along x (an odd last one pads with zeros)
  gemm_tf32x3_pair_kernel<<<grid, THREADS, PairCfg::SM_TOTAL, st>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

template <int BN>
static int launch_bn(const Args& a, cudaStream_t st) {
  if (BN == 256 && pair_wanted() && a.M > BM) return launch_pair(a, st);
  static AttrOnce once;
  CX_TRY(once.set(gemm_tf32x3_kernel<BN>, (int)Cfg<BN>::SM_TOTAL));
  dim3 grid((unsigned)cdiv(a.N, BN), (unsigned)cdiv(a.M, BM), (unsigned)max(a.ksplits, 1));
  gemm_tf32x3_kernel<BN><<<grid, THREADS, Cfg<BN>::SM_TOTAL, st>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

static int launch(const Args& a, cudaStream_t st) {
  return tile_n(a) == 256 ? launch_bn<256>(a, st) : launch_bn<64>(a, st);
}

}  // namespace g3
}  // namespace coordx
