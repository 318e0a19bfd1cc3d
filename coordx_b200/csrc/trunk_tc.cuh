// tcgen05 / TMEM fused fusion+trunk FORWARD for hidden = 256, one Hd x Hd trunk layer
// followed by the Hd -> O layer (the default approx_layers=2, num_hidden_layers=3 net;
// SURVEY 8a configs 1,2,4,5).
//
// Replaces SplitFCBlock.forward_fusion (modules.py:329-354): the broadcast
// h = hs[0] (+|*) hs[1] ... (modules.py:336-341) is generated 128 points at a time
// straight into the UMMA operand layout in shared memory and never reaches HBM; the
// trunk BatchLinear (modules.py:26-27) runs on the 5th-gen tensor cores with the fp32
// accumulator in TMEM; bias, sin(30 x) / ReLU (modules.py:39,172) and the final
// Hd -> O BatchLinear are applied in registers by the epilogue warps.
//
// Work decomposition: the LAST axis is the "inner" axis; a tile = 128 consecutive
// inner samples x one combination `o` of the outer axes.  A-tile row p is
//     A[p, :] = rowvec(o) (+|*) inner[jb*128 + p, :]
// where rowvec(o) is outer0[o] (C=2) or outer0[i0] (+|*) outer1[i1] (C=3).  Tiles are
// ordered (jb major, o minor) and split contiguously over a persistent grid, so a CTA
// keeps one 128-row inner block (64 KB, 16-bit, pre-swizzled by the prep kernel)
// resident in shared memory for hundreds of tiles and only fetches 1 KB of rowvec per
// tile.
//
// CTA = 13 warps, 1 CTA / SM:
//   warps 0-3  epilogue group 0 (TMEM accumulator 0, even tiles), lane quadrant = warp % 4
//   warps 4-7  epilogue group 1 (TMEM accumulator 1, odd tiles)
//   warps 8-11 producers: inner block (smem) x rowvec (registers) -> A ring (smem, SW128)
//   warp 12    one lane issues tcgen05.mma M128 N256 K16 x16 per tile; owns TMEM alloc
// Shared memory: W 128 KB (resident) | inner block 64 KB | A ring 2 x 16 KB (K chunks of 64).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace coordx {
namespace tc {

constexpr int HD = 256;
constexpr int TILE_M = 128;
constexpr int KCHUNK = 64;                          // one SWIZZLE_128B row of 16-bit elements
constexpr int NKC = HD / KCHUNK;                    // 4
constexpr uint32_t W_CHUNK_BYTES = HD * 128;        // 32 KB: [256 rows x 128 B]
constexpr uint32_t W_BYTES = NKC * W_CHUNK_BYTES;   // 128 KB
constexpr uint32_t A_CHUNK_BYTES = TILE_M * 128;    // 16 KB
constexpr uint32_t INNER_BYTES = NKC * A_CHUNK_BYTES;  // 64 KB per 128-row block
constexpr int A_STAGES = 2;
constexpr uint32_t SMEM_W = 0;
constexpr uint32_t SMEM_INNER = SMEM_W + W_BYTES;
constexpr uint32_t SMEM_A = SMEM_INNER + INNER_BYTES;
constexpr uint32_t SMEM_BAR = SMEM_A + A_STAGES * A_CHUNK_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 128;
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack to 1024-align the base by hand
constexpr int NUM_THREADS = 13 * 32;
constexpr int MAX_O = 4;

// epilogue constants (trunk bias pre-scaled, final layer) - warp-uniform, compile-time
// offsets after unrolling => FFMA with constant-bank operands, no extra instructions
struct EpiConst {
  float bias[HD];            // sine: S*b ; relu: b
  float wf[MAX_O][HD];       // final layer weight
  float bf[MAX_O];
};
__constant__ EpiConst c_epi;

// ---------------------------------------------------------------------------------
// prep kernels (tiny; run every call because weights change every training step)
// ---------------------------------------------------------------------------------
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
                    EpiConst* __restrict__ epi) {
  const int idx = blockIdx.x * 256 + threadIdx.x;          // over 256*256
  const int n = idx >> 8, k = idx & 255;
  const uint32_t off = (k >> 6) * W_CHUNK_BYTES + ptx::sw128_offset(n, k & 63);
  *reinterpret_cast<T*>(w_img + off) = to_T<T>(W[idx]);
  if (idx < HD) epi->bias[idx] = act == COORDX_ACT_SINE ? S * b[idx] : b[idx];
  if (idx < MAX_O * HD) {
    const int o = idx >> 8;
    epi->wf[o][idx & 255] = o < O ? Wf[idx] : 0.f;
  }
  if (idx < MAX_O) epi->bf[idx] = idx < O ? bf[idx] : 0.f;
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

// ---------------------------------------------------------------------------------
struct FwdParams {
  const uint8_t* w_img;
  const uint8_t* inner_img;
  const float* outer0;      // [N0, 256] fp32
  const float* outer1;      // [N1, 256] fp32 or nullptr
  int64_t n_outer1;         // N1 (1 when outer1 == nullptr)
  int64_t o_begin, o_end;   // slab of the flattened outer index
  int64_t n_inner;
  int nb;                   // inner blocks = ceil(n_inner / 128)
  float* out;               // [(o_end - o_begin) * n_inner, O]
  float S;
};

template <typename T>
struct Pack;
template <>
struct Pack<__nv_bfloat16> {
  using V2 = __nv_bfloat162;
  static constexpr uint32_t kFmt = 1;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <>
struct Pack<__half> {
  using V2 = __half2;
  static constexpr uint32_t kFmt = 0;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

template <typename T, int OP>
__device__ __forceinline__ uint32_t fuse2(uint32_t x, uint32_t y) {
  using V2 = typename Pack<T>::V2;
  V2 a = *reinterpret_cast<V2*>(&x), b = *reinterpret_cast<V2*>(&y);
  V2 r = OP == COORDX_FUSE_PROD ? __hmul2(a, b) : __hadd2(a, b);
  return *reinterpret_cast<uint32_t*>(&r);
}

template <typename T, int OP, int ACT, int O>
__global__ void __launch_bounds__(NUM_THREADS, 1)
trunk_fwd_tc_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + SMEM_BAR;
  const uint32_t bar_full = bar0;            // [A_STAGES]
  const uint32_t bar_empty = bar0 + 16;      // [A_STAGES]
  const uint32_t bar_tfull = bar0 + 32;      // [2]
  const uint32_t bar_tempty = bar0 + 48;     // [2]
  const uint32_t bar_w = bar0 + 64;
  const uint32_t bar_inner = bar0 + 72;
  const uint32_t tmem_slot = bar0 + 80;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SMEM_BAR + 80);

  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int64_t t_begin = total * blockIdx.x / gridDim.x;
  const int64_t t_end = total * (blockIdx.x + 1) / gridDim.x;
  const int my_tiles = (int)(t_end - t_begin);

  if (warp == 12) {
    if (lane == 0) {
      for (int s = 0; s < A_STAGES; ++s) {
        ptx::mbar_init(bar_full + 8 * s, 128);
        ptx::mbar_init(bar_empty + 8 * s, 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(bar_tfull + 8 * a, 1);
        ptx::mbar_init(bar_tempty + 8 * a, 128);
      }
      ptx::mbar_init(bar_w, 1);
      ptx::mbar_init(bar_inner, 1);
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

  if (warp == 12) {
    // ============================ MMA issuer ============================
    if (lane == 0 && my_tiles > 0) {
      ptx::mbar_arrive_expect_tx(bar_w, W_BYTES);
      for (uint32_t c = 0; c < W_BYTES; c += 16384)
        ptx::bulk_g2s(smem_base + SMEM_W + c, p.w_img + c, 16384, bar_w);
      ptx::mbar_wait(bar_w, 0);
      constexpr uint32_t idesc = ptx::umma_idesc_f16(Pack<T>::kFmt, TILE_M, HD);
      uint32_t chunk_it = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const uint32_t a = lt & 1, use = lt >> 1;
        ptx::mbar_wait(bar_tempty + 8 * a, (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * HD;
#pragma unroll 1
        for (int kc = 0; kc < NKC; ++kc, ++chunk_it) {
          const uint32_t s = chunk_it & 1, ph = (chunk_it >> 1) & 1;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + SMEM_A + s * A_CHUNK_BYTES;
          const uint32_t b_addr = smem_base + SMEM_W + kc * W_CHUNK_BYTES;
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) {
            ptx::mma_f16_ss(d_tmem, ptx::umma_desc_k_sw128(a_addr + k * 32), ptx::umma_desc_k_sw128(b_addr + k * 32),
                            idesc, (kc | k) != 0);
          }
          ptx::mma_commit(bar_empty + 8 * s);
        }
        ptx::mma_commit(bar_tfull + 8 * a);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ============================ producers ============================
    const int ptid = threadIdx.x - 8 * 32;        // 0..127
    const int g8 = ptid & 7;                      // 16-byte piece (8 elements) inside the 128-byte row
    const int r0 = ptid >> 3;                     // rows r0 + 16*i
    const uint32_t piece_off = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((g8 ^ (r0 & 7)) << 4);
    uint32_t chunk_it = 0, inner_phase = 0;
    int64_t cur_jb = -1;
    uint32_t rv[NKC][4];

    auto load_rowvec = [&](int64_t o) {
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
    };

    for (int lt = 0; lt < my_tiles; ++lt) {
      const int64_t t = t_begin + lt;
      const int64_t jb = t / n_o;
      const int64_t o = p.o_begin + (t - jb * n_o);
      load_rowvec(o);
      if (jb != cur_jb) {
        // every producer has finished reading the previous inner block before it is overwritten
        ptx::named_bar_sync(1, 128);
        if (ptid == 0) {
          ptx::fence_proxy_async();
          ptx::mbar_arrive_expect_tx(bar_inner, INNER_BYTES);
          for (uint32_t c = 0; c < INNER_BYTES; c += 16384)
            ptx::bulk_g2s(smem_base + SMEM_INNER + c, p.inner_img + (size_t)jb * INNER_BYTES + c, 16384, bar_inner);
        }
        ptx::mbar_wait(bar_inner, inner_phase);
        inner_phase ^= 1;
        cur_jb = jb;
      }
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc, ++chunk_it) {
        const uint32_t s = chunk_it & 1, ph = (chunk_it >> 1) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint8_t* src = smem_gen + SMEM_INNER + kc * A_CHUNK_BYTES + piece_off;
        uint8_t* dst = smem_gen + SMEM_A + s * A_CHUNK_BYTES + piece_off;
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
        ptx::fence_proxy_async();           // generic-proxy stores -> visible to the tensor core (async proxy)
        ptx::mbar_arrive(bar_full + 8 * s);
      }
    }
  } else {
    // ============================ epilogue ============================
    const int e = warp >> 2;                      // accumulator / tile parity
    const int q = warp & 3;                       // TMEM lane quadrant
    const int row = q * 32 + lane;
    for (int lt = e; lt < my_tiles; lt += 2) {
      const uint32_t use = lt >> 1;
      ptx::mbar_wait(bar_tfull + 8 * e, use & 1);
      ptx::tc_fence_after();
      const int64_t t = t_begin + lt;
      const int64_t jb = t / n_o;
      const int64_t o_local = t - jb * n_o;
      const uint32_t taddr = tmem_base + e * HD + ((uint32_t)(q * 32) << 16);
      float acc[O];
#pragma unroll
      for (int oo = 0; oo < O; ++oo) acc[oo] = 0.f;
      // TMEM -> registers double-buffered: the load of column batch cb+1 is in flight while
      // batch cb goes through bias + sin/relu + the Hd -> O dot product
      uint32_t ra[32], rb[32];
      auto consume = [&](const uint32_t (&r)[32], int cb) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = cb * 32 + i;
          const float z = __uint_as_float(r[i]);
          float h;
          if (ACT == COORDX_ACT_SINE)
            h = __sinf(fmaf(z, p.S, c_epi.bias[k]));
          else
            h = fmaxf(z + c_epi.bias[k], 0.f);
#pragma unroll
          for (int oo = 0; oo < O; ++oo) acc[oo] = fmaf(h, c_epi.wf[oo][k], acc[oo]);
        }
      };
      ptx::tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll
      for (int cb = 0; cb < HD / 32; cb += 2) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + (cb + 1) * 32, rb);
        consume(ra, cb);
        ptx::tmem_ld_wait();
        if (cb + 2 < HD / 32) ptx::tmem_ld_32x32b_x32(taddr + (cb + 2) * 32, ra);
        consume(rb, cb + 1);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * e);
      const int64_t j = jb * TILE_M + row;
      if (j < p.n_inner) {
        float* dst = p.out + (o_local * p.n_inner + j) * O;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) dst[oo] = acc[oo] + c_epi.bf[oo];
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
struct FwdPlan {
  size_t w_img_off, epi_off, inner_img_off, total;
  int64_t n_inner_padded;
  int nb;
};

static FwdPlan plan_fwd(int64_t n_inner) {
  FwdPlan pl{};
  pl.nb = (int)cdiv(n_inner, TILE_M);
  pl.n_inner_padded = (int64_t)pl.nb * TILE_M;
  size_t off = 0;
  pl.w_img_off = off; off += align_up(W_BYTES, 1024);
  pl.epi_off = off; off += align_up(sizeof(EpiConst), 1024);
  pl.inner_img_off = off; off += (size_t)pl.nb * INNER_BYTES;
  pl.total = off;
  return pl;
}

// optional device timing of the main kernel alone (bench.py roofline)
static bool g_profile = false;
static cudaEvent_t g_prof_ev[2] = {nullptr, nullptr};

template <typename T, int OP, int ACT, int O>
static int launch_fwd_inst(const FwdParams& fp, int grid, cudaStream_t st) {
  auto kern = trunk_fwd_tc_kernel<T, OP, ACT, O>;
  static bool attr_done = false;
  if (!attr_done) {
    CX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
    attr_done = true;
  }
  if (g_profile) {
    if (!g_prof_ev[0]) {
      CX_CUDA(cudaEventCreate(&g_prof_ev[0]));
      CX_CUDA(cudaEventCreate(&g_prof_ev[1]));
    }
    CX_CUDA(cudaEventRecord(g_prof_ev[0], st));
  }
  kern<<<grid, NUM_THREADS, SMEM_ALLOC, st>>>(fp);
  CX_LAUNCH_CHECK();
  if (g_profile) CX_CUDA(cudaEventRecord(g_prof_ev[1], st));
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

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// feats: C tables [n_c, 256] fp32.  The last axis is the inner axis; axes 0..C-2 are outer.
static int trunk_forward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                            int64_t row_end, float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int C = net->n_branches;
  const int L = net->n_pre;                       // trunk = layers L (Hd x Hd) and L+1 (Hd -> O)
  const int O = net->layer_out[L + 1];
  const int64_t n_inner = n[C - 1];
  const FwdPlan pl = plan_fwd(n_inner);
  if (ws_bytes < pl.total || ws == nullptr) return set_err(COORDX_ERR_WORKSPACE, "tcgen05 trunk: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)pl.total);
  uint8_t* base = (uint8_t*)ws;
  uint8_t* w_img = base + pl.w_img_off;
  EpiConst* epi = (EpiConst*)(base + pl.epi_off);
  uint8_t* inner_img = base + pl.inner_img_off;
  const float S = 30.f * net->act_scale;

  if (precision == COORDX_PREC_BF16) {
    prep_weights_kernel<__nv_bfloat16><<<HD * HD / 256, 256, 0, st>>>(net->layer_w[L], net->layer_b[L], net->layer_w[L + 1],
                                                                     net->layer_b[L + 1], O, net->act, S, w_img, epi);
    prep_inner_kernel<__nv_bfloat16><<<(unsigned)cdiv(pl.n_inner_padded * HD, 256), 256, 0, st>>>(feats[C - 1], n_inner, inner_img, pl.n_inner_padded);
  } else {
    prep_weights_kernel<__half><<<HD * HD / 256, 256, 0, st>>>(net->layer_w[L], net->layer_b[L], net->layer_w[L + 1],
                                                              net->layer_b[L + 1], O, net->act, S, w_img, epi);
    prep_inner_kernel<__half><<<(unsigned)cdiv(pl.n_inner_padded * HD, 256), 256, 0, st>>>(feats[C - 1], n_inner, inner_img, pl.n_inner_padded);
  }
  CX_LAUNCH_CHECK();
  CX_CUDA(cudaMemcpyToSymbolAsync(c_epi, epi, sizeof(EpiConst), 0, cudaMemcpyDeviceToDevice, st));

  FwdParams fp{};
  fp.w_img = w_img;
  fp.inner_img = inner_img;
  fp.outer0 = feats[0];
  fp.outer1 = C == 3 ? feats[1] : nullptr;
  fp.n_outer1 = C == 3 ? n[1] : 1;
  fp.o_begin = row_begin * fp.n_outer1;
  fp.o_end = row_end * fp.n_outer1;
  fp.n_inner = n_inner;
  fp.nb = pl.nb;
  fp.out = out;
  fp.S = S;
  const int64_t total = (int64_t)pl.nb * (fp.o_end - fp.o_begin);
  if (total <= 0) return 0;
  const int grid = (int)min((int64_t)num_sms(), total);
  if (precision == COORDX_PREC_BF16) return launch_fwd_t<__nv_bfloat16>(fp, net->fusion_op, net->act, O, grid, st);
  return launch_fwd_t<__half>(fp, net->fusion_op, net->act, O, grid, st);
}

static bool tc_supported(const coordx_net* net, int precision) {
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

}  // namespace tc
}  // namespace coordx
