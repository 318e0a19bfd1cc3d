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
// One 128 x 256 output tile per CTA, K in chunks of 32 (= one 128-byte SWIZZLE_128B row of fp32).  No operand image
// in global memory and no TMA descriptor: 8 producer warps read X and W rows with coalesced 16-byte loads (one chunk
// ahead, in registers), split hi/lo and store the four K-major tiles (A_hi, A_lo, B_hi, B_lo: 96 KB per stage, 2
// stages); one elected thread issues 12 kind::tf32 MMAs (M128 x N256 x K8) per chunk; the producer warps drain the
// accumulator at the end (bias, activation, optional pre-activation output, strided columns).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace coordx {
namespace g3 {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr uint32_t A_T = BM * 128, B_T = BN * 128;       // 16 KB, 32 KB: [rows x 128 B] SWIZZLE_128B
constexpr uint32_t STAGE = 2 * A_T + 2 * B_T;            // A_hi | A_lo | B_hi | B_lo
constexpr int STAGES = 2;
constexpr uint32_t SM_BAR = STAGES * STAGE;
constexpr uint32_t SM_TOTAL = SM_BAR + 128;
constexpr int THREADS = 9 * 32;                          // 8 producer / epilogue warps + the MMA issuer

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

__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const Args g) {
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
  const int nk = g.K / BK;

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
    ptx::tmem_alloc(b_slot, 512);
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
          mma_tf32_ss(tacc + 256, dal, dbh, idesc, (kc | ks) != 0);      // cross terms: their own accumulator
          mma_tf32_ss(tacc + 256, dah, dbl, idesc, 1);
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
    float4 ra[4], rb[8];
    auto gload = [&](int kc) {
      const int k0 = kc * BK + piece * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + rbase + 32 * i;
        ra[i] = m < g.M ? __ldg(reinterpret_cast<const float4*>(g.A + (int64_t)m * g.lda + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + rbase + 32 * i;
        rb[i] = n < g.N ? __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)n * g.ldb + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    gload(0);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc & 1;
      ptx::mbar_wait(b_empty + 8 * s, ((kc >> 1) & 1) ^ 1);
      uint8_t* st = smem_gen + s * STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = rbase + 32 * i;
        split_store(st, st + A_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), ra[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = rbase + 32 * i;
        split_store(st + 2 * A_T, st + 2 * A_T + B_T, (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), rb[i]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(b_full + 8 * s);
      if (kc + 1 < nk) gload(kc + 1);
    }
    // ============================ epilogue ============================
    ptx::mbar_wait(b_acc, 0);
    ptx::tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const float add1 = g.bias1 ? __ldg(g.bias1) : 0.f;
    const bool vec = g.cs == 1 && (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.Y) | reinterpret_cast<uintptr_t>(g.Z)) & 15) == 0;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      uint32_t r[32], r2[32];
      ptx::tmem_ld_32x32b_x32(tacc + ((uint32_t)(q * 32) << 16) + half * 128 + b * 32, r);
      ptx::tmem_ld_32x32b_x32(tacc + ((uint32_t)(q * 32) << 16) + 256 + half * 128 + b * 32, r2);
      ptx::tmem_ld_wait();
      const int nb = n0 + half * 128 + b * 32;
      if (m < g.M && nb < g.N) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = nb + i;
          v[i] = (__uint_as_float(r[i]) + __uint_as_float(r2[i])) + add1 + ((g.bias && n < g.N) ? __ldg(g.bias + n) * g.bscale : 0.f);
        }
        if (vec && nb + 32 <= g.N) {
          if (g.Z) {
            float4* z = reinterpret_cast<float4*>(g.Z + (int64_t)m * g.ldc + nb);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (g.Y) {
            if (g.act >= 0) {
#pragma unroll 1
              for (int i = 0; i < 32; ++i) v[i] = act_fwd(g.act, g.S, v[i]);
            }
            float4* y = reinterpret_cast<float4*>(g.Y + (int64_t)m * g.ldc + nb);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < 32; ++i) {
            const int n = nb + i;
            if (n >= g.N) break;
            if (g.Z) g.Z[(int64_t)m * g.ldc + (int64_t)n * g.cs] = v[i];
            if (g.Y) g.Y[(int64_t)m * g.ldc + (int64_t)n * g.cs] = g.act < 0 ? v[i] : act_fwd(g.act, g.S, v[i]);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tacc, 512);
  }
}

// worth it (and legal) for this problem?  Large enough to fill the chip with 128 x 256 tiles, K a multiple of 32,
// 16-byte aligned operand rows.  COORDX_FP32_GEMM=ffma forces the FFMA kernel.
static bool usable(const Args& a) {
  const char* e = getenv("COORDX_FP32_GEMM");
  if (e && strcmp(e, "ffma") == 0) return false;
  if (a.K < BK || a.K % BK != 0 || a.N < 128) return false;
  if ((a.lda & 3) || (a.ldb & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B)) & 15) return false;
  const int64_t tiles = cdiv(a.M, BM) * cdiv(a.N, BN);
  return tiles >= 32 && (int64_t)a.M * a.N * a.K >= ((int64_t)1 << 28);
}

static int launch(const Args& a, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    CX_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    attr_done = true;
  }
  dim3 grid((unsigned)cdiv(a.N, BN), (unsigned)cdiv(a.M, BM));
  gemm_tf32x3_kernel<<<grid, THREADS, SM_TOTAL, st>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

}  // namespace g3
}  // namespace coordx
