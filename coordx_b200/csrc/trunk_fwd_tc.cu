// tcgen05 / TMEM fused fusion + trunk FORWARD, hidden = 256, one Hd x Hd trunk layer followed by
// the Hd -> O layer (approx_layers = 2, num_hidden_layers = 3: SURVEY 8a configs 1, 2, 4, 5).
//
// Replaces SplitFCBlock.forward_fusion (modules.py:329-354).  The broadcast
// h = hs[0] (+|*) hs[1] ... (modules.py:336-341) is generated 128 points at a time straight into the
// UMMA operand layout in shared memory and never reaches HBM; the trunk BatchLinear
// (modules.py:26-27) runs on the 5th-gen tensor cores with the fp32 accumulator in TMEM; bias,
// sin(30 x) / ReLU (modules.py:39,172) and the final Hd -> O BatchLinear are applied in registers by
// the epilogue warps.
//
// Why a CTA PAIR (cta_group::2).  A single-CTA M128 x N256 x K16 MMA reads 4 KB of A and 8 KB of B
// from shared memory every 128 cycles (96 B/clk of the SM's 128 B/clk) - with the A-tile producer's
// own traffic on top, the first version of this kernel was shared-memory bound at 53 % tensor-pipe
// utilisation (profiles/r01a_fwd_trunk_full.txt).  Here two CTAs of a cluster issue ONE M256 x N256
// MMA: each CTA supplies its own 128-point A tile and HALF of W (its 128 output features), so the
// tensor core pulls 64 B/clk per SM, W takes 64 KB of shared memory instead of 128, and the freed
// space is an 8-deep A ring.  The producers keep the inner-axis feature block in REGISTERS
// (128 x 256 16-bit values = 64 registers per producer thread) for hundreds of tiles and only
// write the fused A tile (32 B/clk): total 96 B/clk.
//
// Work decomposition: the LAST axis is the "inner" axis; a tile = 128 consecutive inner samples x one
// combination `o` of the outer axes.  A-tile row p is  A[p, :] = rowvec(o) (+|*) inner[jb*128 + p, :]
// with rowvec(o) = outer0[o] (C = 2) or outer0[i0] (+|*) outer1[i1] (C = 3).  Tiles are ordered
// (jb major, o minor) and split contiguously over the clusters; inside a cluster the two CTAs take
// alternate tiles, so both keep the same inner block most of the time.
//
// CTA = 20 warps (5 warpgroups, registers re-balanced with setmaxnreg), 1 CTA / SM:
//   warps 0-3   epilogue group 0: columns   0..127 of every accumulator, lane quadrant = warp % 4
//   warps 4-7   epilogue group 1: columns 128..255 of every accumulator
//   warps 8-15  producers: inner block piece (registers) x rowvec piece (registers) -> A ring (smem, SW128)
//   warp 16     TMEM alloc; in the LEADER CTA one lane issues tcgen05.mma.cta_group::2 for the pair
//   warps 17-19 idle (complete the warpgroup that gives its registers away)
// Shared memory: W half 64 KB (resident) | A ring 2 x 64 KB (whole tiles, 4 K-chunks of 64 each).
//
// This file is compiled EPI_COPIES times (-DCOORDX_FWD_COPY=k, coordx_b200/build.py): the epilogue reads the trunk
// bias and the final layer as constant-bank operands with compile-time offsets (a runtime slot index turns the
// 16-byte uniform loads into twice as many 8-byte ones and costs 10 % of the kernel, measured), so every copy of the
// kernel owns one `__constant__` image.  Launches take the copies round-robin: forwards of different nets on
// different streams, or in different captured graphs, do not overwrite each other's constants.  Copy 0 also holds
// the host side.
#include <type_traits>
#include <mutex>

#include "tc_common.cuh"

#ifndef COORDX_FWD_COPY
#define COORDX_FWD_COPY 0
#endif
#define CX_CAT2(a, b) a##b
#define CX_CAT(a, b) CX_CAT2(a, b)
#define FWD_COPY_NS CX_CAT(fwd_copy, COORDX_FWD_COPY)

namespace coordx {
namespace tc {

constexpr uint32_t CHUNK_BYTES = TILE_M * 128;      // 16 KB: [128 rows x 128 B]
constexpr int A_STAGES = 2;                         // whole 128 x 256 tiles (4 chunks each)
constexpr uint32_t SMEM_W = 0;                                      // 4 chunks: this CTA's 128 output features
constexpr uint32_t TILE_BYTES = NKC * CHUNK_BYTES;                  // 64 KB
constexpr uint32_t SMEM_A = SMEM_W + NKC * CHUNK_BYTES;             // A ring
constexpr uint32_t SMEM_PB = SMEM_A + A_STAGES * TILE_BYTES;        // float pb[256]: S b / 2pi + 6144 (polynomial sine path)
constexpr uint32_t SMEM_PWF = SMEM_PB + HD * 4;                     // __half2 pwf[MAX_O][128]: final-layer weights as fp16 pairs
constexpr int XCH_DEPTH = 4;
constexpr uint32_t SMEM_XCH = SMEM_PWF + MAX_O * (HD / 2) * 4;          // float xch[XCH_DEPTH][MAX_O][128]
constexpr uint32_t SMEM_BAR = SMEM_XCH + XCH_DEPTH * MAX_O * TILE_M * 4;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 256 + 4 * XCH_DEPTH * 8;     // + xfull[4 quadrants][XCH_DEPTH]
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack to 1024-align the base by hand
constexpr int NUM_THREADS = 20 * 32;
#ifndef COORDX_FWD_MUFU_COLS
#define COORDX_FWD_MUFU_COLS 32
#endif
// of the 16 column pairs of a 32-column batch, how many evaluate sin on the FMA pipe (packed fp32x2 minimax polynomial,
// max error 2.5e-4) instead of the MUFU: the epilogue is MUFU-bound (one sin per 512 tensor FLOPs = the tcgen05 time of
// the tile) with issue slots to spare
#ifndef COORDX_FWD_POLY_PAIRS
#define COORDX_FWD_POLY_PAIRS 4
#endif
#ifndef COORDX_FWD_EPI0_BATCHES
#define COORDX_FWD_EPI0_BATCHES 3
#endif
__host__ __device__ constexpr bool fwd_poly_pair(int pr) {
  return (pr * COORDX_FWD_POLY_PAIRS) / 16 != ((pr + 1) * COORDX_FWD_POLY_PAIRS) / 16;
}
// launch: 640 threads x 96 registers; then 256 x 96 (epilogue, unchanged) + 256 x 128 (producers) + 128 x 32 (MMA warpgroup) = 640 x 96
constexpr int REGS_PROD = 128, REGS_MMA = 32;

// epilogue constants (trunk bias pre-scaled, final layer): warp-uniform, compile-time offsets after
// unrolling => FFMA operands straight from the constant bank
struct EpiConst {
  float bias[HD];            // sine: S*b ; relu: b
  float bias_rev[HD];        // sine: S*b / 2 pi (angle in revolutions, for the columns that take the polynomial)
  float wf[MAX_O][HD];       // final layer weight
  float bf[MAX_O];
};

// ---------------------------------------------------------------------------------
struct FwdParams {
  const float* W;           // trunk weight [256 (n_out), 256 (k_in)] fp32
  const float* inner;       // inner-axis feature table [n_inner, 256] fp32
  const float* outer0;      // [N0, 256] fp32
  const float* outer1;      // [N1, 256] fp32 or nullptr
  int64_t n_outer1;         // N1 (1 when outer1 == nullptr)
  int64_t o_begin, o_end;   // slab of the flattened outer index
  int64_t n_inner;
  int nb;                   // inner blocks = ceil(n_inner / 128)
  float* out;               // [(o_end - o_begin) * n_inner, O]
  float S;
  // loss fused into the last layer (LOSS instantiation; training.py:133-135 + loss_functions.py:8-12 / 56-60): the epilogue
  // thread that completes out[p, :] also reads gt[p, :], writes d loss / d out and keeps a running loss sum
  const float* gt;          // [same layout as out]
  float* dout;              // [same layout as out]   d loss / d out, mean reduction (inv_count)
  float* loss_acc;          // [1] += loss
  float inv_count;
  int loss_kind;            // COORDX_LOSS_MSE | COORDX_LOSS_BCE
  long long* trace;         // optional [4 roles][TRACE_ITERS][4] SM-clock stamps written by CTA 0 (profiling hook)
};
constexpr int TRACE_ITERS = 96;

namespace FWD_COPY_NS {
__constant__ EpiConst c_epi;     // this copy's image of the epilogue constants

// sin(2 pi t) for the "polynomial" share of the epilogue's columns (SINE only).  The MUFU unit evaluates
// 16 sin/clk/SM - exactly one 128 x 256 tile per 2048 clk, the tcgen05 time of the same tile - so an
// all-MUFU epilogue can never run ahead of the tensor core.  NM of every 32 columns use MUFU.SIN, the rest:
//   u  = z * S/2pi + (S b/2pi + 6144)     fp32 FFMA; 6144 = 1.5 * 2^12 makes ulp(u) = 2^-11, so the low 11
//                                         mantissa bits of u are round(frac(t) * 2^11): free range reduction
//   x  = 0x3C00 | (bits 9..0)             fp16 in [1, 2) encoding the position inside the half period
//   |sin| = cos(pi (x - 1.5))             even polynomial of degree 6 in fp16x2 (HFMA2: two columns per op)
//   sign  = bit 10                        XORed into the fp16 sign bits
//   acc  += sin * wf                      FHFMA (fp16 x fp16 + fp32)
// Error of this path: 6.5e-4 rms / 2.2e-3 max per activation (oracle-checked in tests), below the
// ~1.3e-3 rad the bf16 operands already put on the sine's argument; output parity stays well inside 2e-3.
__device__ __forceinline__ uint64_t f2x(float v) { return ptx::pack_f32x2(v, v); }
__device__ __forceinline__ uint32_t h2_as_u32(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b, float c) {      // c + a.lo * b.lo
  float r;
  asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t b, float c) {      // c + a.hi * b.hi
  float r;
  asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}"
      : "=f"(r) : "r"(a), "r"(b), "f"(c));
  return r;
}
constexpr float POLY_MAGIC = 6144.f;
constexpr float POLY_C0 = 0.99999325f, POLY_C1 = -4.93393467f, POLY_C2 = 4.0412481f, POLY_C3 = -1.22203176f;

__device__ __forceinline__ void trace_stamp(const FwdParams& p, int role, int it, int slot) {
#if COORDX_TRACE
  if (p.trace != nullptr && blockIdx.x == 0 && it < TRACE_ITERS) p.trace[(role * TRACE_ITERS + it) * 4 + slot] = clock64();
#endif
}

// LOSS: 0 = plain forward, 1 + COORDX_LOSS_* = that loss fused into the last layer (a compile-time kind: with the
// kind as a run-time branch the unused BCE code alone cost the 2048^2 step 13-26 us)
template <typename T, int OP, int ACT, int O, int NM, int LOSS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
trunk_fwd_tc_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;     // same offset in both CTAs of the pair
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t bar0 = smem_base + SMEM_BAR;
  const uint32_t bar_full = bar0;              // [A_STAGES]  leader only: 16 producer warps (8 per CTA) have written the tile
  const uint32_t bar_empty = bar0 + 64;        // [A_STAGES]  each CTA: the MMAs reading the tile have completed (multicast commit)
  const uint32_t bar_tfull = bar0 + 128;       // [2]         each CTA: accumulator complete (multicast commit)
  const uint32_t bar_tempty = bar0 + 144;      // [2]         leader only: 16 epilogue warps (8 per CTA) have drained the accumulator
  const uint32_t bar_xfull = bar0 + 256;       // [4][XCH_DEPTH] group 1's warp of quadrant q has deposited its partial sums (32)
  const uint32_t tmem_slot = bar0 + 160;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SMEM_BAR + 160);

  // tiles of this cluster; CTA `rank` takes tiles c_begin + 2*it + rank
  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t n_cl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  const int64_t c_begin = total * cl / n_cl;
  const int64_t c_end = total * (cl + 1) / n_cl;
  const int n_iter = (int)((c_end - c_begin + 1) >> 1);

  if (warp == 16) {
    if (lane == 0) {
      for (int s = 0; s < A_STAGES; ++s) {
        ptx::mbar_init(bar_full + 8 * s, 16);
        ptx::mbar_init(bar_empty + 8 * s, 1);
      }
      for (int i = 0; i < 4 * XCH_DEPTH; ++i) ptx::mbar_init(bar_xfull + 8 * i, 32);
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(bar_tfull + 8 * a, 1);
        ptx::mbar_init(bar_tempty + 8 * a, 16);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc_2cta(tmem_slot, 512);
    ptx::tmem_relinquish_2cta();
  }
  // ---- this CTA's half of W: fp32 [128 rows of n_out, 256] -> bf16/fp16, 4 K-chunks [128 x 128 B] SWIZZLE_128B
  {
    // SINE: the activation scale S = 30 * act_scale is folded into the 16-bit weights (W' = S W), so the
    // accumulator already holds S z and the epilogue adds the pre-scaled bias with one packed FADD2
    const float ws = ACT == COORDX_ACT_SINE ? p.S : 1.f;
    const float* Wh = p.W + (size_t)rank * 128 * HD;
    for (int pc = threadIdx.x; pc < 128 * 32; pc += NUM_THREADS) {
      const int n = pc >> 5, g = pc & 31;                  // row, 16-byte piece (8 elements) of the 256-wide row
      const float4 a = *reinterpret_cast<const float4*>(Wh + n * HD + g * 8);
      const float4 b = *reinterpret_cast<const float4*>(Wh + n * HD + g * 8 + 4);
      uint4 w;
      w.x = Pack<T>::pack(ws * a.x, ws * a.y);
      w.y = Pack<T>::pack(ws * a.z, ws * a.w);
      w.z = Pack<T>::pack(ws * b.x, ws * b.y);
      w.w = Pack<T>::pack(ws * b.z, ws * b.w);
      *reinterpret_cast<uint4*>(smem_gen + SMEM_W + (g >> 3) * CHUNK_BYTES + ptx::sw128_offset(n, (g & 7) * 8)) = w;
    }
    ptx::fence_proxy_async();
  }
  if (NM < 32) {
    float* pb = reinterpret_cast<float*>(smem_gen + SMEM_PB);
    __half2* pwf = reinterpret_cast<__half2*>(smem_gen + SMEM_PWF);
    for (int i = threadIdx.x; i < HD; i += NUM_THREADS) pb[i] = c_epi.bias[i] * 0.15915494309189535f + POLY_MAGIC;
    for (int i = threadIdx.x; i < MAX_O * (HD / 2); i += NUM_THREADS) {
      const int o = i / (HD / 2), k2 = i % (HD / 2);
      pwf[i] = __floats2half2_rn(c_epi.wf[o][2 * k2], c_epi.wf[o][2 * k2 + 1]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                                     // peer's barriers, W half and TMEM allocation are ready
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp >= 16) {
    ptx::setmaxnreg_dec<REGS_MMA>();
    if (warp == 16 && rank == 0 && n_iter > 0 && ptx::elect_one()) {     // (elect, not lane == 0: see ptx::elect_one)
      // ============================ MMA issuer (leader CTA, one thread) ============================
      constexpr uint32_t idesc = ptx::umma_idesc_f16(Pack<T>::kFmt, 2 * TILE_M, HD);
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t a = it & 1, use = it >> 1;          // accumulator == A stage == iteration parity
        trace_stamp(p, 0, it, 0);
        ptx::mbar_wait(bar_tempty + 8 * a, (use & 1) ^ 1);
        trace_stamp(p, 0, it, 1);
        ptx::mbar_wait(bar_full + 8 * a, use & 1);
        trace_stamp(p, 0, it, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * HD;
        const uint32_t a_addr = smem_base + SMEM_A + a * TILE_BYTES;
        const uint32_t b_addr = smem_base + SMEM_W;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;
          ptx::mma_f16_ss_2cta(d_tmem, ptx::umma_desc_k_sw128(a_addr + off), ptx::umma_desc_k_sw128(b_addr + off), idesc, k != 0);
        }
        ptx::mma_commit_2cta(bar_empty + 8 * a, 3);
        trace_stamp(p, 0, it, 3);
        ptx::mma_commit_2cta(bar_tfull + 8 * a, 3);
      }
    }
  } else if (warp >= 8) {
    // ============================ producers ============================
    ptx::setmaxnreg_inc<REGS_PROD>();
    // thread (g = lane, rg = warp - 8) owns the 16-byte piece g (k = 8g .. 8g+7) of the 16 rows
    // 16*rg .. 16*rg+15: its 8 rowvec values are 4 packed registers loaded straight from global
    // memory (no broadcast through shared memory), and a warp's stores cover whole 128-byte lines.
    const int g = lane, rg = warp - 8;
    const uint32_t full_remote = ptx::mapa(bar_full, 0);
    // byte offset of (row 16*rg, piece g) inside a tile stage, before the per-row swizzle
    uint8_t* const a_base = smem_gen + SMEM_A + (uint32_t)(g >> 3) * CHUNK_BYTES + (uint32_t)rg * 2048u;
    uint32_t in[16][4];                           // inner block rows x this piece, 16-bit pairs
    int cur_jb = -1;
    // tile -> (jb, o) without divisions in the loop: tiles advance by 2
    int64_t t0 = c_begin + rank;
    if (t0 >= c_end) t0 = c_end - 1;
    int64_t t_next = c_begin + rank;
    const int n_o32 = (int)n_o;
    int jb = n_iter > 0 ? (int)(t0 / n_o) : 0;
    int ol = n_iter > 0 ? (int)(t0 - (int64_t)jb * n_o) : 0;
    auto load_rv = [&](int64_t o, float4& x0, float4& x1) {
      int64_t i0 = o, i1 = 0;
      if (p.outer1) { i0 = o / p.n_outer1; i1 = o - i0 * p.n_outer1; }
      x0 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + 8 * g);
      x1 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + 8 * g + 4);
      if (p.outer1) {
        const float4 y0 = *reinterpret_cast<const float4*>(p.outer1 + i1 * HD + 8 * g);
        const float4 y1 = *reinterpret_cast<const float4*>(p.outer1 + i1 * HD + 8 * g + 4);
        if (OP == COORDX_FUSE_PROD) {
          x0.x *= y0.x; x0.y *= y0.y; x0.z *= y0.z; x0.w *= y0.w;
          x1.x *= y1.x; x1.y *= y1.y; x1.z *= y1.z; x1.w *= y1.w;
        } else {
          x0.x += y0.x; x0.y += y0.y; x0.z += y0.z; x0.w += y0.w;
          x1.x += y1.x; x1.y += y1.y; x1.z += y1.z; x1.w += y1.w;
        }
      }
    };

    for (int it = 0; it < n_iter; ++it) {
      // rowvec piece of this tile (L2-resident table; the ~1 us of latency is slack: a tile needs ~600 of the
      // producers' 2048 cycles) - not prefetched, the registers are needed for the inner block
      uint32_t rv[4];
      {
        float4 x0, x1;
        load_rv(p.o_begin + ol, x0, x1);
        rv[0] = Pack<T>::pack(x0.x, x0.y);
        rv[1] = Pack<T>::pack(x0.z, x0.w);
        rv[2] = Pack<T>::pack(x1.x, x1.y);
        rv[3] = Pack<T>::pack(x1.z, x1.w);
      }
      if (jb != cur_jb) {
        // (rare: once per ~n_o tiles) reload the inner block piece.  Batches of 4 rows are chained through an
        // opaque zero so that ptxas cannot hoist all 64 loads at once (it would spill the block to local memory).
        uint32_t chain = 0;
#pragma unroll
        for (int r4 = 0; r4 < 16; r4 += 4) {
          float4 v0[4], v1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t j = (int64_t)jb * TILE_M + rg * 16 + r4 + u;
            v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            v1[u] = v0[u];
            if (j < p.n_inner) {
              const float* src = p.inner + j * HD + 8 * g + chain;
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
      // next tile of this CTA; past the end of the range (odd range, rank 1, last iteration) the previous
      // tile is produced again and the epilogue discards it.
      t_next += 2;
      if (t_next < c_end) {
        ol += 2;
        while (ol >= n_o32) { ol -= n_o32; ++jb; }
      }
      const uint32_t st = it & 1, use = it >> 1;
      if (threadIdx.x == 8 * 32) trace_stamp(p, 3, it, 0);
      ptx::mbar_wait(bar_empty + 8 * st, (use & 1) ^ 1);
      if (threadIdx.x == 8 * 32) trace_stamp(p, 3, it, 1);
      uint8_t* dst = a_base + st * TILE_BYTES;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        uint4 w;
        w.x = fuse2<T, OP>(in[r][0], rv[0]);
        w.y = fuse2<T, OP>(in[r][1], rv[1]);
        w.z = fuse2<T, OP>(in[r][2], rv[2]);
        w.w = fuse2<T, OP>(in[r][3], rv[3]);
        // row 16*rg + r: 8-row group (r >> 3) is 1024 bytes, row (r & 7) 128 bytes, piece XOR-swizzled with the row
        *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + (((g & 7) ^ (r & 7)) << 4)) = w;
      }
      ptx::fence_proxy_async();             // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(full_remote + 8 * st);
      if (threadIdx.x == 8 * 32) trace_stamp(p, 3, it, 2);
    }
  } else {
    // ============================ epilogue ============================
    // all 8 warps work on EVERY tile: group e = warp >> 2 takes columns [128 e, 128 e + 128) of the
    // accumulator, so one tile is drained in half the time and the accumulator goes back to the tensor
    // core before the other one is complete (two groups that each owned an accumulator ran in phase
    // and left the MMA idle ~40 % of the time: profiles/r01b_fwd_timeline.txt)
    const int e = warp >> 2;                      // column half
    const int q = warp & 3;                       // TMEM lane quadrant
    const int row = q * 32 + lane;
    float* xch = reinterpret_cast<float*>(smem_gen + SMEM_XCH);     // [XCH_DEPTH][MAX_O][128] partial outputs of group 1
    const uint32_t tempty_remote0 = ptx::mapa(bar_tempty, 0);
    // (the column half is a compile-time constant inside the body so that bias / final-layer weights stay
    //  direct constant-bank operands)
    auto run = [&](auto EC) {
    constexpr int E = decltype(EC)::value;
    // group 0 also finalises and stores every tile: it takes fewer of the eight 32-column batches
    constexpr int NB = E == 0 ? COORDX_FWD_EPI0_BATCHES : 8 - COORDX_FWD_EPI0_BATCHES, CB0 = E == 0 ? 0 : COORDX_FWD_EPI0_BATCHES;
    int64_t t_cur = c_begin + rank;
    const int n_o32 = (int)n_o;
    int jb_e = (int)(t_cur / n_o), ol_e = (int)(t_cur - (int64_t)jb_e * n_o);
    // The two column halves of a row meet in shared memory: group 1 deposits its partial sums of tile `it` into
    // slot it % 4 and moves on; group 0 picks them up ONE TILE LATER (after its own tile it+1), by which time they
    // are there, so neither warp ever waits for its partner.  The asymmetry is deliberate: it keeps the two warps
    // of an SMSP out of phase - one keeps the MUFU busy while the other does its scalar end-of-tile work
    // (alternating the storer role balanced the groups, put them back in phase and was 9 % slower).  Slot reuse
    // distance 4 >> the <= 1.25-tile skew the accumulator hand-off allows between any two epilogue warps.
    float pend[O];                               // group 0: own partial sums of the previous tile
    float pend_gt[O];                            // LOSS: the targets of that row, loaded when the row was parked
    float loss_part = 0.f;                       // LOSS: this thread's running loss sum
    float* pend_dst = nullptr;
    int pend_it = -1;
    auto finalize = [&]() {
      if (E == 0 && pend_it >= 0) {
        ptx::mbar_wait(bar_xfull + 8 * (q * XCH_DEPTH + (pend_it & (XCH_DEPTH - 1))), (pend_it / XCH_DEPTH) & 1);
        if (pend_dst != nullptr) {
          const float* xb = xch + (pend_it & (XCH_DEPTH - 1)) * (MAX_O * TILE_M);
#pragma unroll
          for (int oo = 0; oo < O; ++oo) {
            const float x = pend[oo] + xb[oo * TILE_M + row] + c_epi.bf[oo];
            pend_dst[oo] = x;
            if (LOSS) {                          // same expressions as loss_kernel (elementwise.cuh)
              const float z = pend_gt[oo];
              float l, d;
              if (LOSS == 1 + COORDX_LOSS_MSE) {
                const float e = x - z;
                l = e * e;
                d = 2.f * e;
              } else {
                const float ex = expf(-fabsf(x));
                l = fmaxf(x, 0.f) - x * z + logf(1.f + ex);
                const float sig = x >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
                d = sig - z;
              }
              loss_part += l;
              p.dout[(pend_dst - p.out) + oo] = d * p.inv_count;
            }
          }
        }
      }
    };
    for (int it = 0; it < n_iter; ++it) {
      const uint32_t a = it & 1, use = it >> 1;
      const uint32_t tempty_remote = tempty_remote0 + 8 * a;
      if (q == 0 && lane == 0) trace_stamp(p, 1 + E, it, 0);
      ptx::mbar_wait(bar_tfull + 8 * a, use & 1);
      if (q == 0 && lane == 0) trace_stamp(p, 1 + E, it, 1);
      ptx::tc_fence_after();
      const bool valid = t_cur < c_end;
      const int64_t jb = jb_e, o_local = ol_e;     // tile -> (inner block, outer index), advanced without divisions
      const uint32_t taddr = tmem_base + a * HD + CB0 * 32 + ((uint32_t)(q * 32) << 16);
      // acc2: packed pair (even-column sum, odd-column sum) fed by FFMA2; acc: the polynomial path's FHFMAs
      uint64_t acc2[O];
      float acc[O];
#pragma unroll
      for (int oo = 0; oo < O; ++oo) { acc2[oo] = 0ull; acc[oo] = 0.f; }
      // TMEM -> registers double-buffered: the load of column batch cb+1 is in flight while
      // batch cb goes through bias + sin/relu + the Hd -> O dot product.  Bias and final-layer
      // weights are direct constant-bank operands of the FFMAs (an LDC per value is far slower).
      uint32_t ra[32], rb[32];
      auto consume = [&](const uint32_t (&r)[32], int cbl) {
        const int cb = CB0 + cbl;                 // 32-column batch inside the 256-wide accumulator
        constexpr int NMU = ACT == COORDX_ACT_SINE ? NM : 32;
#pragma unroll
        for (int i = 0; i < NMU; i += 2) {
          // MUFU path, two columns per packed fp32x2 instruction: FADD2 (bias pair from uniform registers),
          // 2 x (FMUL.RZ + MUFU.SIN), FFMA2 into the packed accumulator = 3.25 issue slots per activation
          const int k = cb * 32 + i;
          const uint64_t zz = ptx::pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          uint64_t hh;
          if (ACT == COORDX_ACT_SINE && fwd_poly_pair(i >> 1)) {
            // sin(2 pi t), t = (S z + S b) / 2 pi: r = t - rint(t) in [-1/2, 1/2] by the 1.5 * 2^23 trick, odd polynomial in r
            const uint64_t t2 = ptx::fma_f32x2(zz, f2x(0.15915494309189535f), ptx::pack_f32x2(c_epi.bias_rev[k], c_epi.bias_rev[k + 1]));
            const uint64_t m2 = ptx::add_f32x2(t2, f2x(12582912.f));
            const uint64_t n2 = ptx::add_f32x2(m2, f2x(-12582912.f));
            const uint64_t rr = ptx::fma_f32x2(n2, f2x(-1.f), t2);
            const uint64_t q2 = ptx::mul_f32x2(rr, rr);
            uint64_t ps = ptx::fma_f32x2(q2, f2x(-56.08980178833008f), f2x(77.93164825439453f));
            ps = ptx::fma_f32x2(ps, q2, f2x(-41.093868255615234f));
            ps = ptx::fma_f32x2(ps, q2, f2x(6.27863883972168f));
            hh = ptx::mul_f32x2(ps, rr);
          } else {
            float t0, t1;
            ptx::unpack_f32x2(ptx::add_f32x2(zz, ptx::pack_f32x2(c_epi.bias[k], c_epi.bias[k + 1])), t0, t1);
            float h0, h1;
            if (ACT == COORDX_ACT_SINE) {
              h0 = __sinf(t0);
              h1 = __sinf(t1);
            } else {
              h0 = fmaxf(t0, 0.f);
              h1 = fmaxf(t1, 0.f);
            }
            hh = ptx::pack_f32x2(h0, h1);
          }
#pragma unroll
          for (int oo = 0; oo < O; ++oo)
            acc2[oo] = ptx::fma_f32x2(hh, ptx::pack_f32x2(c_epi.wf[oo][k], c_epi.wf[oo][k + 1]), acc2[oo]);
        }
        if (NMU < 32) {
          // polynomial path, two columns per fp16x2 operation
          const float Sp = 0.15915494309189535f;      // the accumulator holds S z (S is folded into the weights)
          const __half2 c0 = __float2half2_rn(POLY_C0), c1 = __float2half2_rn(POLY_C1), c2 = __float2half2_rn(POLY_C2),
                        c3 = __float2half2_rn(POLY_C3), mh = __float2half2_rn(-1.5f);
#pragma unroll
          for (int i = NMU; i < 32; i += 4) {
            const int k = cb * 32 + i;
            const float4 b4 = *reinterpret_cast<const float4*>(smem_gen + SMEM_PB + k * 4);            // warp-uniform: broadcast
            uint2 w2[O];
#pragma unroll
            for (int oo = 0; oo < O; ++oo) w2[oo] = *reinterpret_cast<const uint2*>(smem_gen + SMEM_PWF + (oo * (HD / 2) + k / 2) * 4);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float u0 = fmaf(__uint_as_float(r[i + 2 * j]), Sp, bb[2 * j]);
              const float u1 = fmaf(__uint_as_float(r[i + 2 * j + 1]), Sp, bb[2 * j + 1]);
              const uint32_t w = __byte_perm(__float_as_uint(u0), __float_as_uint(u1), 0x5410);   // {u1.lo16, u0.lo16}
              const __half2 v = __hadd2(u32_as_h2((w & 0x03FF03FFu) | 0x3C003C00u), mh);
              const __half2 v2 = __hmul2(v, v);
              __half2 q = __hfma2(c3, v2, c2);
              q = __hfma2(q, v2, c1);
              q = __hfma2(q, v2, c0);
              const uint32_t hs = h2_as_u32(q) ^ ((w << 5) & 0x80008000u);
#pragma unroll
              for (int oo = 0; oo < O; ++oo) {
                const uint32_t wf2 = j == 0 ? w2[oo].x : w2[oo].y;
                acc[oo] = fhfma_lo(hs, wf2, acc[oo]);
                acc[oo] = fhfma_hi(hs, wf2, acc[oo]);
              }
            }
          }
        }
      };
      ptx::tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        ptx::tmem_ld_wait();
        if (b + 1 < NB) {
          if (b & 1) ptx::tmem_ld_32x32b_x32(taddr + (b + 1) * 32, ra);
          else ptx::tmem_ld_32x32b_x32(taddr + (b + 1) * 32, rb);
        } else {
          // the last column batch is in registers: hand the accumulator back to the tensor core now, one batch of
          // the tile's epilogue work early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty_remote);
        }
        if (b & 1) consume(rb, b);
        else consume(ra, b);
      }
      if (q == 0 && lane == 0) trace_stamp(p, 1 + E, it, 2);
#pragma unroll
      for (int oo = 0; oo < O; ++oo) {
        float lo, hi;
        ptx::unpack_f32x2(acc2[oo], lo, hi);
        acc[oo] += lo + hi;
      }
      if (E == 1) {
        float* xb = xch + (it & (XCH_DEPTH - 1)) * (MAX_O * TILE_M);
#pragma unroll
        for (int oo = 0; oo < O; ++oo) xb[oo * TILE_M + row] = acc[oo];
        ptx::mbar_arrive(bar_xfull + 8 * (q * XCH_DEPTH + (it & (XCH_DEPTH - 1))));
      } else {
        finalize();                              // tile it - 1
        const int64_t j = jb * TILE_M + row;
        pend_dst = (valid && j < p.n_inner) ? p.out + (o_local * p.n_inner + j) * O : nullptr;
        pend_it = it;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) pend[oo] = acc[oo];
        if (LOSS && pend_dst != nullptr) {
#pragma unroll
          for (int oo = 0; oo < O; ++oo) pend_gt[oo] = __ldg(p.gt + (pend_dst - p.out) + oo);
        }
      }
      if (q == 0 && lane == 0) trace_stamp(p, 1 + E, it, 3);
      t_cur += 2;
      ol_e += 2;
      while (ol_e >= n_o32) { ol_e -= n_o32; ++jb_e; }
    }
    finalize();                                  // the last tile
    if (LOSS && E == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      if (lane == 0) atomicAdd(p.loss_acc, loss_part * p.inv_count);
    }
    };
    if (e == 0) run(std::integral_constant<int, 0>{}); else run(std::integral_constant<int, 1>{});
  }

  // both CTAs are done with each other's shared memory / TMEM before anything is torn down
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 16) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------

// columns per 32 that take MUFU.SIN; the polynomial share is only used with bf16 operands, whose own
// rounding already dominates the error (fp16 operands keep the exact path)
template <typename T> struct MufuCols { static constexpr int value = 32; };
template <> struct MufuCols<__nv_bfloat16> { static constexpr int value = COORDX_FWD_MUFU_COLS; };

template <typename T, int OP, int ACT, int O>
static int launch_fwd_inst(const FwdParams& fp, int grid, cudaStream_t st) {
  constexpr int NM = ACT == COORDX_ACT_SINE ? MufuCols<T>::value : 32;
  static AttrOnce once, once_mse, once_bce;
  if (fp.gt != nullptr && fp.loss_kind == COORDX_LOSS_MSE) {
    auto kern = trunk_fwd_tc_kernel<T, OP, ACT, O, NM, 1 + COORDX_LOSS_MSE>;
    CX_TRY(once_mse.set(kern, (int)SMEM_ALLOC));
    kern<<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(fp);
  } else if (fp.gt != nullptr) {
    auto kern = trunk_fwd_tc_kernel<T, OP, ACT, O, NM, 1 + COORDX_LOSS_BCE>;
    CX_TRY(once_bce.set(kern, (int)SMEM_ALLOC));
    kern<<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(fp);
  } else {
    auto kern = trunk_fwd_tc_kernel<T, OP, ACT, O, NM, 0>;
    CX_TRY(once.set(kern, (int)SMEM_ALLOC));
    kern<<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(fp);
  }
  CX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int OP, int ACT>
static int launch_fwd_o(const FwdParams& fp, int O, int grid, cudaStream_t st) {
  switch (O) {
    case 1: return launch_fwd_inst<T, OP, ACT, 1>(fp, grid, st);
    case 2: return launch_fwd_inst<T, OP, ACT, 2>(fp, grid, st);
    case 3: return launch_fwd_inst<T, OP, ACT, 3>(fp, grid, st);
    case 4: return launch_fwd_inst<T, OP, ACT, 4>(fp, grid, st);
  }
  return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk: out width %s%lld > 4", "", O);
}

template <typename T>
static int launch_fwd_t(const FwdParams& fp, int op, int act, int O, int grid, cudaStream_t st) {
  if (op == COORDX_FUSE_PROD)
    return act == COORDX_ACT_SINE ? launch_fwd_o<T, COORDX_FUSE_PROD, COORDX_ACT_SINE>(fp, O, grid, st)
                                  : launch_fwd_o<T, COORDX_FUSE_PROD, COORDX_ACT_RELU>(fp, O, grid, st);
  return act == COORDX_ACT_SINE ? launch_fwd_o<T, COORDX_FUSE_SUM, COORDX_ACT_SINE>(fp, O, grid, st)
                                : launch_fwd_o<T, COORDX_FUSE_SUM, COORDX_ACT_RELU>(fp, O, grid, st);
}

}  // namespace FWD_COPY_NS

// entry of this copy: load its constant image from `epi_dev` (device memory, stream-ordered), launch its kernel
#define FWD_COPY_ENTRY CX_CAT(trunk_fwd_launch_copy, COORDX_FWD_COPY)
// (ev_start: optional event recorded between the two - the profiling hook times the kernel alone)
int FWD_COPY_ENTRY(const FwdParams& fp, const EpiConst* epi_dev, int precision, int op, int act, int O, int grid, cudaStream_t st,
                   cudaEvent_t ev_start) {
  CX_CUDA(cudaMemcpyToSymbolAsync(FWD_COPY_NS::c_epi, epi_dev, sizeof(EpiConst), 0, cudaMemcpyDeviceToDevice, st));
  if (ev_start) CX_CUDA(cudaEventRecord(ev_start, st));
  if (precision == COORDX_PREC_BF16) return FWD_COPY_NS::launch_fwd_t<__nv_bfloat16>(fp, op, act, O, grid, st);
  return FWD_COPY_NS::launch_fwd_t<__half>(fp, op, act, O, grid, st);
}

#if COORDX_FWD_COPY == 0
// =================================================================================
// host side (copy 0 only)
// =================================================================================
constexpr int EPI_COPIES = 4;            // keep in step with coordx_b200/build.py
int trunk_fwd_launch_copy1(const FwdParams&, const EpiConst*, int, int, int, int, int, cudaStream_t, cudaEvent_t);
int trunk_fwd_launch_copy2(const FwdParams&, const EpiConst*, int, int, int, int, int, cudaStream_t, cudaEvent_t);
int trunk_fwd_launch_copy3(const FwdParams&, const EpiConst*, int, int, int, int, int, cudaStream_t, cudaEvent_t);

static __global__ void __launch_bounds__(256)
prep_epi_kernel(const float* __restrict__ b, const float* __restrict__ Wf, const float* __restrict__ bf, int O, int act,
                float S, EpiConst* __restrict__ epi) {
  const int idx = blockIdx.x * 256 + threadIdx.x;          // over MAX_O * 256
  if (idx < HD) {
    epi->bias[idx] = act == COORDX_ACT_SINE ? S * b[idx] : b[idx];
    epi->bias_rev[idx] = act == COORDX_ACT_SINE ? S * b[idx] * 0.15915494309189535f : 0.f;
  }
  if (idx < MAX_O * HD) {
    const int o = idx >> 8;
    epi->wf[o][idx & 255] = o < O ? Wf[idx] : 0.f;
  }
  if (idx < MAX_O) epi->bf[idx] = idx < O ? bf[idx] : 0.f;
}


struct FwdPlan {
  size_t epi_off, total;
  int nb;
};

static FwdPlan plan_fwd(int64_t n_inner) {
  FwdPlan pl{};
  pl.nb = (int)cdiv(n_inner, TILE_M);
  size_t off = 0;
  pl.epi_off = off; off += align_up(sizeof(EpiConst), 1024);
  pl.total = off;
  return pl;
}

// optional device timing of the main kernel alone (bench.py roofline)
static bool g_profile = false;
static cudaEvent_t g_prof_ev[2] = {nullptr, nullptr};
static long long* g_trace = nullptr;
void profile_enable(bool on) { g_profile = on; }
int profile_set_trace(void* dev_buf) {
#if !COORDX_TRACE
  if (dev_buf) return set_err(COORDX_ERR_UNSUPPORTED, "pipeline tracing is a build option%s (scripts/build_trace_lib.sh, -DCOORDX_TRACE=1)");
#endif
  g_trace = (long long*)dev_buf;
  return 0;
}
// milliseconds of the last profiled trunk kernel (synchronises on its end event)
int profile_last_ms(float* ms_host) {
  if (!ms_host || !g_prof_ev[0]) return set_err(COORDX_ERR_ARG, "no profiled launch%s");
  CX_CUDA(cudaEventSynchronize(g_prof_ev[1]));
  CX_CUDA(cudaEventElapsedTime(ms_host, g_prof_ev[0], g_prof_ev[1]));
  return 0;
}
size_t fwd_workspace_bytes(int64_t n_inner) { return plan_fwd(n_inner).total; }

// Who used a copy's constant image last (per device): an eager launch on ANOTHER stream waits for that kernel before
// it overwrites the image.  Captured launches are ordered by their graph; they neither wait on nor record these
// events (an event recorded inside a capture belongs to the capture).
struct CopyState {
  cudaEvent_t done = nullptr;
  cudaStream_t last = nullptr;
  bool recorded = false;
};
static CopyState g_copy_state[64][EPI_COPIES];
static std::mutex g_copy_mutex;
static unsigned g_next_copy = 0;

// feats: C tables [n_c, 256] fp32.  The last axis is the inner axis; axes 0..C-2 are outer.
int trunk_forward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                     int64_t row_end, float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st,
                     const FwdLoss* loss) {
  const int C = net->n_branches;
  const int L = net->n_pre;                       // trunk = layers L (Hd x Hd) and L+1 (Hd -> O)
  const int O = net->layer_out[L + 1];
  const int64_t n_inner = n[C - 1];
  const FwdPlan pl = plan_fwd(n_inner);
  if (ws_bytes < pl.total || ws == nullptr) return set_err(COORDX_ERR_WORKSPACE, "tcgen05 trunk: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)pl.total);
  EpiConst* epi = (EpiConst*)((uint8_t*)ws + pl.epi_off);
  const float S = 30.f * net->act_scale;

  FwdParams fp{};
  fp.W = net->layer_w[L];
  fp.inner = feats[C - 1];
  fp.outer0 = feats[0];
  fp.outer1 = C == 3 ? feats[1] : nullptr;
  fp.n_outer1 = C == 3 ? n[1] : 1;
  fp.o_begin = row_begin * fp.n_outer1;
  fp.o_end = row_end * fp.n_outer1;
  fp.n_inner = n_inner;
  fp.nb = pl.nb;
  fp.out = out;
  fp.S = S;
  fp.trace = g_trace;
  if (loss != nullptr) {
    fp.gt = loss->gt; fp.dout = loss->dout; fp.loss_acc = loss->loss_acc; fp.inv_count = loss->inv_count; fp.loss_kind = loss->kind;
  }
  const int64_t total = (int64_t)pl.nb * (fp.o_end - fp.o_begin);
  if (total <= 0) return 0;
  const int grid = 2 * (int)min((int64_t)(num_sms() / 2), cdiv(total, 2));

  prep_epi_kernel<<<MAX_O * HD / 256, 256, 0, st>>>(net->layer_b[L], net->layer_w[L + 1], net->layer_b[L + 1], O, net->act, S, epi);
  CX_LAUNCH_CHECK();

  int dev = 0;
  CX_CUDA(cudaGetDevice(&dev));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CX_CUDA(cudaStreamIsCapturing(st, &cap));
  const bool eager = cap == cudaStreamCaptureStatusNone && dev >= 0 && dev < 64;
  std::lock_guard<std::mutex> lock(g_copy_mutex);
  const int copy = (int)(g_next_copy++ % EPI_COPIES);
  CopyState* cs = eager ? &g_copy_state[dev][copy] : nullptr;
  if (cs && cs->recorded && cs->last != st) CX_CUDA(cudaStreamWaitEvent(st, cs->done, 0));
  if (g_profile) {
    if (!g_prof_ev[0]) {
      CX_CUDA(cudaEventCreate(&g_prof_ev[0]));
      CX_CUDA(cudaEventCreate(&g_prof_ev[1]));
    }
  }
  cudaEvent_t ev0 = g_profile ? g_prof_ev[0] : nullptr;
  int rc;
  switch (copy) {
    case 0: rc = trunk_fwd_launch_copy0(fp, epi, precision, net->fusion_op, net->act, O, grid, st, ev0); break;
    case 1: rc = trunk_fwd_launch_copy1(fp, epi, precision, net->fusion_op, net->act, O, grid, st, ev0); break;
    case 2: rc = trunk_fwd_launch_copy2(fp, epi, precision, net->fusion_op, net->act, O, grid, st, ev0); break;
    default: rc = trunk_fwd_launch_copy3(fp, epi, precision, net->fusion_op, net->act, O, grid, st, ev0); break;
  }
  CX_TRY(rc);
  if (g_profile) CX_CUDA(cudaEventRecord(g_prof_ev[1], st));
  if (cs) {
    if (!cs->done) CX_CUDA(cudaEventCreateWithFlags(&cs->done, cudaEventDisableTiming));
    CX_CUDA(cudaEventRecord(cs->done, st));
    cs->last = st;
    cs->recorded = true;
  }
  return 0;
}

bool tc_supported(const coordx_net* net, int precision) {
  if (precision != COORDX_PREC_BF16 && precision != COORDX_PREC_FP16) return false;
  const int C = net->n_branches, L = net->n_pre;
  if (C < 2 || C > 3) return false;
  if (net->hidden != HD || net->fuse_width != HD) return false;
  if (net->n_layers - net->n_pre != 2) return false;
  if (net->layer_in[L] != HD || net->layer_out[L] != HD || net->layer_in[L + 1] != HD) return false;
  if (net->layer_out[L + 1] < 1 || net->layer_out[L + 1] > MAX_O) return false;
  if (net->fusion_before_act || net->has_code) return false;
  if (net->act != COORDX_ACT_SINE && net->act != COORDX_ACT_RELU) return false;
  return true;
}
#endif  // COORDX_FWD_COPY == 0

}  // namespace tc
}  // namespace coordx
