// Declarations shared by the two translation units of the tcgen05 trunk backward
// (trunk_bwd_tc.cu: recompute / fallback passes and the host side; trunk_bwd_tc2.cu: the dW pass with W_h and
// dZ^T in tensor memory and the merged feature-gradient pass).
#pragma once
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

constexpr uint32_t CH = 16384;                    // one [128 rows x 128 B] SWIZZLE_128B chunk
constexpr uint32_t SM_W = 0;                      // 4 chunks (64 KB): this CTA's 128 output features of W
// dH pass:  H 4 chunks | dZ^T 2 chunks | inner block 4 chunks          dW pass:  H stage 0 | H stage 1 | dZ^T 2 chunks
constexpr uint32_t SM_H = 4 * CH;
constexpr uint32_t SM_DZ_DH = 8 * CH, SM_INNER = 10 * CH;
constexpr uint32_t SM_H1 = 8 * CH, SM_DZ_DW = 12 * CH;
constexpr uint32_t SM_BAR = 14 * CH;
constexpr uint32_t SM_DOUT = SM_BAR + 256;        // float dout[128 points][MAX_O] of the current tile (staged by the producers)
constexpr uint32_t SM_TOTAL_B = SM_DOUT + TILE_M * MAX_O * 4;     // 231680 bytes <= 227 KB
constexpr uint32_t SM_RV = SM_DOUT + 1024;        // O <= 2: two 512-byte rowvec slots in the unused tail of the dout staging area
// dH pass: 0-7 E1 | 8-15 E2 | 16-19 producers | 20 MMA.   dW pass: 0-7 E1 | 8-15 producers | 16 MMA.
constexpr int DH_THREADS = 21 * 32, DW_THREADS = 17 * 32;

struct BwdParams {
  const uint8_t* w_img;      // full W image of the forward prep: 4 K-chunks of [256 rows x 128 B]
  const uint8_t* inner_img;  // inner axis table, per 128-row block 4 K-chunks [128 x 128 B]   (dH pass)
  const float* inner_f32;    // the same table as fp32 [n_inner, 256]                           (dW pass)
  const float* w_f32;        // trunk weight [256 n_out, 256 k_in] fp32                          (dH pass, E2)
  long long* trace;          // optional [4 roles][TRACE_ITERS][4] SM-clock stamps written by CTA 0 (profiling hook)
  const uint8_t* outer_img;  // optional: the outer0 table as 16-bit rows [n, 256] (512 B each): dH-pass producers fetch rowvec(o) with
                             // the copy engine two tiles ahead instead of global loads on their critical path (outer1 == nullptr only)
  float* d_inner;            // [n_inner, 256] +=  gradient w.r.t. the inner axis' features (din pass)
  uint8_t* dz_img;           // optional: dZ^T_h tiles, 32 KB per (tile, h), written by the dW pass and read by the drv pass
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
  uint32_t* flags;           // fused kernel: [tiles] stores completed per tile (4 = both halves x both sub-tiles), zeroed by the host
  int fused_pairs;           // fused kernel: number of dW CTA pairs (the other CTAs of the grid run the feature-gradient pass)
  int fused_chunk;           // fused kernel: tiles per hand-over chunk
  int fused_groups;          // fused kernel: groups of CTAs that sweep a range of tiles together
  int dz_no_wf;              // the stored dZ^T tiles lack the factor wfs[0][n] (dz_without_wf): the feature-gradient pass folds it into W^T
};

// consts[] layout (floats): [256] bias' | [MAX_O][256] wfs | loss scale, 1 / loss scale, max|dout| bits, pad
constexpr int CONST_SCALE = (1 + MAX_O) * HD;
constexpr int CONST_FLOATS = CONST_SCALE + 8;

constexpr int TRACE_ITERS = 96;
__device__ __forceinline__ void trace_stamp(const BwdParams& p, int role, int it, int slot) {
#if COORDX_TRACE
  if (p.trace != nullptr && blockIdx.x == 0 && it < TRACE_ITERS) p.trace[(role * TRACE_ITERS + it) * 4 + slot] = clock64();
#endif
}


// rowvec(o) piece: 8 fp32 values of outer0[i0] (+|*) outer1[i1] starting at column k0
template <int OP>
__device__ __forceinline__ void load_rowvec8(const BwdParams& p, int64_t o, int k0, float4& x0, float4& x1) {
  int64_t i0 = o, i1 = 0;
  if (p.outer1) { i0 = o / p.n_outer1; i1 = o - i0 * p.n_outer1; }
  x0 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + k0);
  x1 = *reinterpret_cast<const float4*>(p.outer0 + i0 * HD + k0 + 4);
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
}

// One output (O = 1), sine, bf16 operands: dZ[p, n] = dout[p] * wfs[n] * cos(..) is stored and fed to the GEMMs WITHOUT
// its per-feature factor wfs[n] = S Wf[n] - one packed multiply less per pair in E1, the loop that bounds the dW pass.
// The factor comes back for free: dW_h and db are scaled per row when they are flushed, and the feature-gradient pass
// multiplies it into its W^T operand when it loads it (dH = dZ W = dZ' (diag(wfs) W)).  fp16 operands keep the factor
// (their power-of-two loss scale rides on wfs).
template <typename T>
__host__ __device__ constexpr bool dz_without_wf(int act, int O) { return act == COORDX_ACT_SINE && O == 1 && Pack<T>::kFmt == 1; }

// ---- entry points of trunk_bwd_tc2.cu
int launch_dw2(const BwdParams& bp, int precision, int op, int act, int O, int grid, cudaStream_t st);
int launch_dfeat(const BwdParams& bp, int precision, int op, int n_cta, cudaStream_t st);
int launch_bwd_fused(const BwdParams& bp, int precision, int op, int act, int O, int grid, cudaStream_t st);

}  // namespace tcb
}  // namespace coordx
