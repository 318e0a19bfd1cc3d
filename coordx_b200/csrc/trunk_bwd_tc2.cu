// tcgen05 / TMEM trunk backward, second translation unit: the two kernels of the stored-dZ path
// (what autograd does for SplitFCBlock.forward_fusion, modules.py:329-354, in training.py:137-138).
//
//   trunk_bwd_dw2_kernel    G1 Z^T = W_h H^T -> E1 (sin + cos; db, dWf) -> G3 dW_h += dZ^T H, dZ^T tiles -> HBM
//   trunk_bwd_dfeat_kernel  dH^T = W^T dZ^T from the stored tiles, ONE GEMM per tile serving BOTH feature gradients:
//                           d rowvec(o)[k] = sum_p dH[p,k] inner[p,k]   and   d inner[j,k] += rowvec(o)[k] dH[(o,j),k]
//
// dw2 differs from the first dW kernel (trunk_bwd_tc.cu) in where the operands live.  There, W_h (64 KB) sat in shared
// memory next to two 64 KB H stages, and the chain  G3(t-1) -> H stage free -> producers -> G1(t+1) -> E1(t+1)  left
// E1 - the longest stage, sin AND cos of every element - idle ~45 % of the time (profiles/r01e_bwd_dw_full.txt:
// XU 46 %, tensor 44 %).  Here W_h is a TMEM-resident A operand (128 columns, written once), the Z^T double buffer
// works on 64-point sub-tiles (2 x 64 columns), dZ^T is written back over the consumed Z^T columns as before, and the
// 64 KB of shared memory that W_h occupied is a THIRD H stage: H(t+1) can be produced while G3(t-1) is still pending.
// The bulk store of the dZ^T image moved from the MMA-issuing thread to a warp of its own.
//
// dfeat replaces the drv and din kernels: both streamed the same 2.1 GB (2048^2) of dZ^T tiles and each ran its own
// GEMM.  With the output feature k on the TMEM lane, thread k owns dH^T[k, p]: the reduction over the points of a tile
// (d rowvec) is a per-thread dot product with its column of the inner block, and the reduction over the outer index
// (d inner) is a per-thread accumulator per point that lives in registers while the CTA sweeps o for a fixed block
// of inner rows.  W^T is a TMEM-resident A operand (256 columns), so the tensor core fetches only the streamed dZ^T
// tile from shared memory.
#include <type_traits>

#include "trunk_bwd_common.cuh"

namespace coordx {
namespace tcb {

// =================================================================================
// the tiles of one CTA
// =================================================================================
// A CTA works through every `stride`-th chunk of G consecutive tiles of a range [rb, re), starting with chunk `first`.
//  * stand-alone passes: one chunk = the CTA's whole range;
//  * fused kernel: the CTAs form groups of P dW pairs and C feature-gradient CTAs that sweep the group's range front
//    to back TOGETHER: chunk m of the range is computed by pair m mod P and consumed by feature-gradient CTA m mod C,
//    a few tens of microseconds later, out of the L2.  Neighbouring chunks share the block of inner rows (jb) almost
//    always, so the dW producers keep their inner block in registers and the feature-gradient CTAs their d(inner)
//    accumulators, as in the stand-alone passes.  Every role of a CTA derives the same sequence on its own.
struct TileRuns {
  int64_t rb, re;
  int G, m, stride;
  __device__ static TileRuns make(int64_t rb, int64_t re, int G, int first, int stride) {
    TileRuns r;
    r.rb = rb; r.re = re; r.G = G; r.m = first; r.stride = stride;
    return r;
  }
  __device__ static TileRuns single(int64_t b, int64_t e) {
    return make(b, e, (int)(e - b > 0 ? e - b : 1), 0, 1);
  }
  __device__ bool next(int64_t& t0, int& cnt) {
    t0 = rb + (int64_t)m * G;
    if (t0 >= re) return false;
    cnt = (int)(re - t0 < G ? re - t0 : G);
    m += stride;
    return true;
  }
  __device__ int count() const {                   // tiles in all chunks from m on
    const int64_t len = re - rb;
    if (len <= 0) return 0;
    const int n_chunks = (int)((len + G - 1) / G);
    if (m >= n_chunks) return 0;
    const int mine = (n_chunks - 1 - m) / stride + 1;
    const int last = m + (mine - 1) * stride;       // my last chunk: ragged if it is the range's last
    return mine * G - (last == n_chunks - 1 ? (int)((int64_t)n_chunks * G - len) : 0);
  }
};

// walks the tiles of a TileRuns one by one: (jb, o) of the current tile and of the one after it
struct TileWalk {
  TileRuns it;
  int64_t t0, n_o;
  int cnt, idx;
  int64_t t, jb;
  int o;
  bool valid;
  __device__ TileWalk(const TileRuns& r, int64_t n_o_) : it(r), n_o(n_o_), cnt(0), idx(0), t(0), jb(0), o(0) {
    valid = it.next(t0, cnt);
    if (valid) { t = t0; jb = t0 / n_o; o = (int)(t0 - jb * n_o); }
  }
  __device__ void advance() {
    if (++idx < cnt) {
      ++t;
      if (++o == (int)n_o) { o = 0; ++jb; }
    } else {
      idx = 0;
      valid = it.next(t0, cnt);
      if (valid) { t = t0; jb = t0 / n_o; o = (int)(t0 - jb * n_o); }
    }
  }
};

// =================================================================================
// dW pass, second version
// =================================================================================
constexpr int D2_STAGES = 3;
constexpr uint32_t D2_H = 0;                               // 3 stages x 4 chunks [128 points x 128 B]
constexpr uint32_t D2_IMG = D2_STAGES * 4 * CH;            // 2 x 16 KB: dZ^T image of a 64-point sub-tile
constexpr uint32_t D2_BAR = D2_IMG + 2 * CH;
constexpr uint32_t D2_DOUT = D2_BAR + 256;                 // dout staging: D stages x 128 points x O floats
// 0-15 E1 (lane quadrant = warp & 3, 16-point column group = warp >> 2) | 16-23 producers | 24 MMA | 25 image store | 26-27 idle
// (warps come in groups of four for setmaxnreg).  Launched with 72 registers per thread; E1 gives registers away
// (64), the producers - 64 registers of inner block each - take them (104), the MMA / store group keeps 32.
constexpr int DW2_THREADS = 28 * 32;
// tuning knobs (A/B builds: scripts/build_variant_bwd2.sh)
#ifndef COORDX_DW2_NPOLY
#define COORDX_DW2_NPOLY 0          // of the 8 (even, odd) pairs of a thread's 16 values, how many take the FMA-pipe polynomial
#endif
constexpr int D2_E1_ARRIVALS = 16;      // one elected arrive per E1 warp
// CTA-pair form: how the leader learns that the PEER's producers / E1 warps are done with a (sub-)tile.
// 1: the peer's otherwise idle warps 24 / 26 wait on the peer's own barriers and forward ONE remote arrive each, so
//    producers and E1 warps signal exactly as in the single-CTA form; 0: every producer / E1 warp arrives remotely.
#ifndef COORDX_DW3_FORWARD
#define COORDX_DW3_FORWARD 0        // measured: 1035 us (forwarded) vs 1024 us (direct) per 2048^2, profiles/r02g_dw_pair.txt
#endif
constexpr bool D3_FWD = COORDX_DW3_FORWARD != 0;
__host__ __device__ constexpr bool d2_poly_pair(int e2) {
  return (e2 * COORDX_DW2_NPOLY) / 8 != ((e2 + 1) * COORDX_DW2_NPOLY) / 8;
}
constexpr int D2_REGS_E1 = 64, D2_REGS_PROD = 104, D2_REGS_MMA = 32;
__host__ __device__ constexpr int d2_dout_stages(int O) { return O <= 2 ? 2 : 1; }
__host__ __device__ constexpr uint32_t d2_smem_bytes(int O) { return D2_DOUT + d2_dout_stages(O) * TILE_M * O * 4; }

struct Bars2 {
  uint32_t full, empty;        // [3] H stage
  uint32_t zfull;              // [2] Z^T sub-tile buffer (G1 commit)
  uint32_t dzfull;             // [2] dZ^T written by E1 (512 threads): TMEM operand + image chunk
  uint32_t imgempty;           // [2] the bulk store has read the image chunk
  uint32_t dsfull, dsempty;    // [2] dout staging
  uint32_t wfull, dwfull, tmem_slot;
  uint32_t dzpair;             // [2] CTA-pair form, leader only: dZ^T written by the E1 warps of the peer (forwarded) / of both CTAs
  uint32_t fullpeer;           // [3] CTA-pair form, leader only: the peer's H stage is written (forwarded)
  __device__ explicit Bars2(uint32_t b0)
      : full(b0), empty(b0 + 24), zfull(b0 + 48), dzfull(b0 + 64), imgempty(b0 + 80), dsfull(b0 + 96), dsempty(b0 + 112),
        wfull(b0 + 128), dwfull(b0 + 136), tmem_slot(b0 + 144), dzpair(b0 + 160), fullpeer(b0 + 176) {}
};

// sin(2 pi r) and cos(2 pi r) on r in [-1/2, 1/2] (r = fractional part of the angle in revolutions), two values per
// packed fp32x2 instruction.  Minimax fits: max error 2.5e-4 (sin, feeds only the dWf sums) and 4.0e-5 (cos, feeds dZ,
// which is rounded to 16 bits right after).  E1 evaluates HALF of its elements this way and half on the MUFU unit:
// sin AND cos of every element on the MUFU (16 results / clk / SM) was the floor of the whole pass.  The share is
// COORDX_DW2_NPOLY of every 8 pairs: scripts/micro/micro_e1.cu times this arithmetic alone (profiles/r02b_micro_e1.txt:
// 1062 clk per 16 values x 4 warps with 0 / 8, 870 with 2 / 8, 919 with 4 / 8, 1181 with 8 / 8 - MUFU-bound below 2,
// issue-bound above; scalar FFMAs with immediate coefficients instead of the packed ones are 10 % slower).
// Inside the kernel the balance moved once the pipeline-trace stamps were compiled out (they sat in this loop): the pass
// takes 901 us with 0 / 8, 904 with 1 / 8, 922 with 2 / 8, 957 with 4 / 8 (profiles/r02l_poly_share_ab.txt), so the
// default share is now 0 - the polynomial path stays for the A/B.
__device__ __forceinline__ uint64_t f2x(float v) { return ptx::pack_f32x2(v, v); }
__device__ __forceinline__ void sincos_rev_x2(uint64_t t2, uint64_t& sn, uint64_t& cs) {
  const uint64_t m2 = ptx::add_f32x2(t2, f2x(12582912.f));            // 1.5 * 2^23: rounds t to the nearest integer
  const uint64_t n2 = ptx::add_f32x2(m2, f2x(-12582912.f));
  const uint64_t rr = ptx::fma_f32x2(n2, f2x(-1.f), t2);              // r = t - rint(t)
  const uint64_t q2 = ptx::mul_f32x2(rr, rr);
  uint64_t ps = ptx::fma_f32x2(q2, f2x(-56.08980178833008f), f2x(77.93164825439453f));
  ps = ptx::fma_f32x2(ps, q2, f2x(-41.093868255615234f));
  ps = ptx::fma_f32x2(ps, q2, f2x(6.27863883972168f));
  sn = ptx::mul_f32x2(ps, rr);
  uint64_t pc = ptx::fma_f32x2(q2, f2x(45.64854431152344f), f2x(-82.40443420410156f));
  pc = ptx::fma_f32x2(pc, q2, f2x(64.67355346679688f));
  pc = ptx::fma_f32x2(pc, q2, f2x(-19.731046676635742f));
  cs = ptx::fma_f32x2(pc, q2, f2x(0.9999598264694214f));
}
// One CTA's share of the dW pass: half h of the output features, the tiles of `runs`.  p.flags != nullptr (fused
// kernel): the store warp counts every completed sub-tile store into flags[tile] for the feature-gradient CTAs.
//
// PAIR = true (trunk_bwd_dw3_kernel): the two CTAs that own the two halves of the output features are a cluster and
// work on the SAME tile at the same time; both GEMMs are cta_group::2 instructions issued by the leader (h = 0):
//   G1  Z^T [256 n x 64 points] : A = W_h in each CTA's tensor memory, B (N = points) 32 points from each CTA;
//   G3  dW  [256 n x 256 k]     : A = dZ^T in each CTA's tensor memory, B (N = k) 128 columns from each CTA.
// Every tensor core then fetches HALF of each B operand from its own shared memory (32 B/clk instead of 64 B/clk,
// which is the measured fetch limit of SS/TS MMAs, section 3.2 of DESIGN.md) and receives the other half from the
// peer.  A CTA's H stage holds [all 128 points x its 128 k] (chunks 0-1, the G3 operand) and [its 2 x 32 points x
// 256 k] (chunks 2-3 re-organised as [4 k-chunks][2 sub-tiles][32 rows], the G1 operand): 64 KB as before.
template <typename T, int OP, int ACT, int O, bool PAIR>
__device__ __forceinline__ void dw2_body(const BwdParams& p, uint8_t* smem_raw, const int h, const TileRuns runs) {
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();          // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Bars2 B(smem_base + D2_BAR);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + D2_BAR + 144);
  constexpr int W_PROD = 16, W_MMA = 24, W_STORE = 25;
  constexpr int DST = d2_dout_stages(O);

  const int64_t n_o = p.o_end - p.o_begin;
  const int my_tiles = runs.count();
  const int n_sub = 2 * my_tiles;

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int c = 0; c < D2_STAGES; ++c) {
        ptx::mbar_init(B.full + 8 * c, PAIR && !D3_FWD ? 16 : 8);   // 8 producer warps (of each CTA)
        ptx::mbar_init(B.fullpeer + 8 * c, 1);
        ptx::mbar_init(B.empty + 8 * c, 1);
      }
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(B.zfull + 8 * c, 1);
        ptx::mbar_init(B.dzfull + 8 * c, D2_E1_ARRIVALS);
        ptx::mbar_init(B.dzpair + 8 * c, D3_FWD ? 1 : 2 * D2_E1_ARRIVALS);
        ptx::mbar_init(B.imgempty + 8 * c, 1);
        ptx::mbar_init(B.dsfull + 8 * c, 128);
        ptx::mbar_init(B.dsempty + 8 * c, D2_E1_ARRIVALS);
      }
      ptx::mbar_init(B.wfull, PAIR ? 1024 : 512);
      ptx::mbar_init(B.dwfull, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    if (PAIR) {
      ptx::tmem_alloc_2cta(B.tmem_slot, 512);
      ptx::tmem_relinquish_2cta();
    } else {
      ptx::tmem_alloc(B.tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();                   // the peer's barriers and tensor memory exist
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tZ = tmem_base, tW = tmem_base + 128, tDW = tmem_base + 256;
  auto mma_ts = [](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (PAIR) ptx::mma_f16_ts_2cta(d, a, b, idesc, acc); else ptx::mma_f16_ts(d, a, b, idesc, acc);
  };
  auto commit = [](uint32_t bar) {
    if (PAIR) ptx::mma_commit_2cta(bar, 3); else ptx::mma_commit(bar);
  };

  if (warp >= W_MMA) {
    ptx::setmaxnreg_dec<D2_REGS_MMA>();
    if (warp == W_MMA) {
      // ============================ MMA issuer ============================
      if ((!PAIR || h == 0) && my_tiles > 0 && ptx::elect_one()) {     // (elect, not lane == 0: see ptx::elect_one)
        constexpr uint32_t MM = PAIR ? 256 : 128;
        constexpr uint32_t id_g1 = ptx::umma_idesc_f16(Pack<T>::kFmt, MM, 64, 0, 0);    // A = W_h (TMEM, M = n_out), B = H (N = 64 points) K-major
        constexpr uint32_t id_g3 = ptx::umma_idesc_f16(Pack<T>::kFmt, MM, 256, 0, 1);   // A = dZ^T (TMEM, M = n_out), B = H (N = k_in) MN-major
        ptx::mbar_wait(B.wfull, 0);
        ptx::tc_fence_after();
        int st1 = 0, use1 = 0;                     // H stage / use count of the tile G1 works on
        int st3 = 0;                               // ... of the tile G3 works on
        // (a G1 of two N = 32 halves with a commit each took 1.0 k clk instead of 0.54 k: with the A operand in tensor
        // memory an N = 32 MMA costs the same ~32 clk as an N = 64 one)
        auto issue_g1 = [&](int s) {
          const uint32_t zb = s & 1, hh = s & 1;
          if (hh == 0) {
            ptx::mbar_wait(B.full + 8 * st1, use1 & 1);
            if (PAIR && D3_FWD) ptx::mbar_wait(B.fullpeer + 8 * st1, use1 & 1);
            ptx::tc_fence_after();
          }
          // (no wait for the Z^T buffer: its last reader G3(s-2) is earlier in this thread's in-order tensor pipe)
          // rows 64 hh .. 64 hh + 63 of every chunk | pair: this CTA's 32 rows of sub-tile hh, k-chunks 8 KB apart
          const uint32_t hs = PAIR ? smem_base + D2_H + st1 * 4 * CH + 2 * CH + hh * 4096 : smem_base + D2_H + st1 * 4 * CH + hh * 8192;
          constexpr uint32_t KC = PAIR ? 8192 : CH;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            mma_ts(tZ + zb * 64, tW + k * 8, ptx::umma_desc_k_sw128(hs + (k >> 2) * KC + (k & 3) * 32), id_g1, k != 0);
          commit(B.zfull + 8 * zb);
          if (hh == 1 && ++st1 == D2_STAGES) { st1 = 0; ++use1; }
        };
        auto issue_g3 = [&](int s) {
          const uint32_t zb = s & 1, hh = s & 1;
          if (!PAIR || D3_FWD) ptx::mbar_wait(B.dzfull + 8 * zb, (s >> 1) & 1);
          if (PAIR) ptx::mbar_wait(B.dzpair + 8 * zb, (s >> 1) & 1);
          trace_stamp(p, 0, s, 2);
          ptx::tc_fence_after();
          const uint32_t hs = smem_base + D2_H + st3 * 4 * CH;         // (pair: chunks 0-1 = this CTA's 128 k)
#pragma unroll
          for (int j = 0; j < 4; ++j)              // K = 64 points, 16 per step: E1 column group j wrote its dZ^T pairs over its first 8 columns
            mma_ts(tDW, tZ + zb * 64 + j * 16, ptx::umma_desc_mn_sw128(hs + (hh * 4 + j) * 2048, CH), id_g3, (s | j) != 0);
          if (hh == 1) {
            commit(B.empty + 8 * st3);             // H stage free for the producers
            if (++st3 == D2_STAGES) st3 = 0;
          }
        };
        // G1(s+1) is issued AHEAD of G3(s): E1 is the longest stage, its input must never queue behind G3
        issue_g1(0);
        for (int s = 0; s < n_sub; ++s) {
          trace_stamp(p, 0, s, 0);
          if (s + 1 < n_sub) issue_g1(s + 1);
          trace_stamp(p, 0, s, 1);
          issue_g3(s);
          trace_stamp(p, 0, s, 3);
        }
        commit(B.dwfull);
      }
      if (PAIR && D3_FWD && h == 1 && my_tiles > 0 && ptx::elect_one()) {
        // peer CTA: forward "dZ^T of sub-tile s is in this CTA's tensor memory" to the leader's MMA thread
        const uint32_t remote = ptx::mapa(B.dzpair, 0);
        for (int s = 0; s < n_sub; ++s) {
          ptx::mbar_wait(B.dzfull + 8 * (s & 1), (s >> 1) & 1);
          ptx::tc_fence_after();
          ptx::tc_fence_before();
          ptx::mbar_arrive_cluster(remote + 8 * (s & 1));
        }
      }
      __syncwarp();
    } else if (PAIR && D3_FWD && warp == W_STORE + 1) {
      if (h == 1 && my_tiles > 0 && ptx::elect_one()) {
        // peer CTA: forward "H stage written by this CTA's producers"
        const uint32_t remote = ptx::mapa(B.fullpeer, 0);
        int st = 0, use = 0;
        for (int t = 0; t < my_tiles; ++t) {
          ptx::mbar_wait(B.full + 8 * st, use & 1);
          ptx::mbar_arrive_cluster(remote + 8 * st);
          if (++st == D2_STAGES) { st = 0; ++use; }
        }
      }
      __syncwarp();
    } else if (warp == W_STORE) {
      // ============================ dZ^T image -> HBM (one 16 KB bulk store per sub-tile) ============================
      if (p.dz_img != nullptr && my_tiles > 0 && ptx::elect_one()) {
        TileWalk w(runs, n_o);
        int64_t t_prev = 0;                        // tile of the store issued last
        for (int s = 0; s < n_sub; ++s) {
          const int hh = s & 1;
          ptx::mbar_wait(B.dzfull + 8 * (s & 1), (s >> 1) & 1);
          ptx::bulk_s2g(p.dz_img + (((size_t)w.t * 2 + h) * 2 + hh) * CH, smem_base + D2_IMG + (s & 1) * CH, CH);
          ptx::bulk_commit_group();
          // hand the chunk back as soon as the copy engine has READ it (this thread has nothing else to do): signalled
          // one store later, E1(s+2) found its chunk free only after every warp had finished sub-tile s+1
          ptx::bulk_wait_group_read<0>();
          ptx::mbar_arrive(B.imgempty + 8 * (s & 1));
          if (p.flags != nullptr && s > 0) {       // store s - 1 has landed: publish it (one store of latency behind)
            ptx::bulk_wait_group<1>();
            ptx::publish_add(p.flags + t_prev);
          }
          t_prev = w.t;
          if (hh == 1) w.advance();
        }
        ptx::bulk_wait_group0();
        if (p.flags != nullptr) ptx::publish_add(p.flags + t_prev);
      }
      __syncwarp();
    }
  } else if (warp >= W_PROD) {
    ptx::setmaxnreg_inc<D2_REGS_PROD>();
    // ============================ producers: H tile (inner block piece in registers) ============================
    // thread (g = lane, rg = warp - 16) owns the 16-byte piece g (k = 8g .. 8g+7) of the 16 rows 16 rg .. 16 rg + 15.
    // rowvec(o) comes as the 16-bit row image (prep_rowvec16_kernel), one 16-byte load per thread, fetched ONE TILE
    // AHEAD: with 227 KB of shared memory there is no L1 to prefetch into, and an L2 round trip plus three 64-bit
    // divisions per tile put ~2 k clk between two H tiles (profiles/r02a_dw2_timeline.txt).
    const int g = lane, rg = warp - W_PROD;
    const int ptid = threadIdx.x - W_PROD * 32;    // 0..255; the first 128 also stage dout
    uint32_t in[16][4];
    int64_t cur_jb = -1;
    int st = 0, use = 0, dsg = 0, dsuse = 0;
    TileWalk w(runs, n_o);
    const uint4* rv_rows = reinterpret_cast<const uint4*>(p.outer_img) + g;      // row o at + o * 32 (512 B per row)
    uint4 rv_next = my_tiles > 0 ? __ldg(rv_rows + (size_t)w.o * 32) : make_uint4(0, 0, 0, 0);
    for (int t = 0; t < my_tiles; ++t) {
      const uint4 rv = rv_next;
      const int64_t jb = w.jb;
      const int o_local = w.o;
      w.advance();
      if (t + 1 < my_tiles) rv_next = __ldg(rv_rows + (size_t)w.o * 32);         // in flight while this tile is written
      float dreg[O];
#pragma unroll
      for (int oo = 0; oo < O; ++oo) dreg[oo] = 0.f;
      if (ptid < TILE_M) {
        const int64_t j = jb * TILE_M + ptid;
        if (j < p.n_inner) {
#pragma unroll
          for (int oo = 0; oo < O; ++oo) dreg[oo] = __ldg(p.dout + ((int64_t)o_local * p.stride_o + j * p.stride_j) * O + oo);
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
      if (ptid == 0) trace_stamp(p, 3, t, 0);
      ptx::mbar_wait(B.empty + 8 * st, (use & 1) ^ 1);
      if (ptid == 0) trace_stamp(p, 3, t, 1);
      if (!PAIR) {
        uint8_t* dst = smem_gen + D2_H + (uint32_t)st * 4 * CH + (uint32_t)(g >> 3) * CH + (uint32_t)rg * 2048u;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          uint4 w;
          w.x = fuse2<T, OP>(in[r][0], rv.x);
          w.y = fuse2<T, OP>(in[r][1], rv.y);
          w.z = fuse2<T, OP>(in[r][2], rv.z);
          w.w = fuse2<T, OP>(in[r][3], rv.w);
          *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + (((g & 7) ^ (r & 7)) << 4)) = w;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(B.full + 8 * st);
      } else {
        // G3 operand (chunks 0-1): every row, the pieces of this CTA's 128 k (half of the lanes);
        // G1 operand (chunks 2-3): every piece, the rows of this CTA's 32-point blocks (half of the warps)
        uint8_t* stage = smem_gen + D2_H + (uint32_t)st * 4 * CH;
        const bool mine_k = (g >> 4) == h;
        const bool mine_p = ((rg >> 1) & 1) == h;
        // two passes of the single-CTA loop (one base address live at a time: with both, the inner block spilled)
        if (mine_k) {
          uint8_t* dst = stage + (uint32_t)((g >> 3) & 1) * CH + (uint32_t)rg * 2048u;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            uint4 w;
            w.x = fuse2<T, OP>(in[r][0], rv.x);
            w.y = fuse2<T, OP>(in[r][1], rv.y);
            w.z = fuse2<T, OP>(in[r][2], rv.z);
            w.w = fuse2<T, OP>(in[r][3], rv.w);
            *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + (((g & 7) ^ (r & 7)) << 4)) = w;
          }
        }
        if (mine_p) {                              // (warp-uniform)
          uint8_t* dst = stage + 2 * CH + (uint32_t)(g >> 3) * 8192u + (uint32_t)(rg >> 2) * 4096u + (uint32_t)(rg & 1) * 2048u;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            uint4 w;
            w.x = fuse2<T, OP>(in[r][0], rv.x);
            w.y = fuse2<T, OP>(in[r][1], rv.y);
            w.z = fuse2<T, OP>(in[r][2], rv.z);
            w.w = fuse2<T, OP>(in[r][3], rv.w);
            *reinterpret_cast<uint4*>(dst + (r >> 3) * 1024 + (r & 7) * 128 + (((g & 7) ^ (r & 7)) << 4)) = w;
          }
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (D3_FWD) ptx::mbar_arrive(B.full + 8 * st); else ptx::mbar_arrive_cluster(ptx::mapa(B.full, 0) + 8 * st);
        }
      }
      if (ptid == 0) trace_stamp(p, 3, t, 2);
      if (++st == D2_STAGES) { st = 0; ++use; }
      if (ptid < TILE_M) {
        ptx::mbar_wait(B.dsempty + 8 * dsg, (dsuse & 1) ^ 1);    // E1 has read the dout that lived in this slot
        float* ds = reinterpret_cast<float*>(smem_gen + D2_DOUT) + (dsg * TILE_M + ptid) * O;
#pragma unroll
        for (int oo = 0; oo < O; ++oo) ds[oo] = dreg[oo];
        ptx::mbar_arrive(B.dsfull + 8 * dsg);
      }
      if (++dsg == DST) { dsg = 0; ++dsuse; }
      if (ptid == 0) trace_stamp(p, 3, t, 3);
    }
  } else {
    ptx::setmaxnreg_dec<D2_REGS_E1>();
    // ============================ E1 (16 warps): Z^T sub-tile -> dZ^T (TMEM operand of G3 + image chunk) ============================
    // thread = output feature n = 128 h + row; warp (q, pq) covers the 16 points [16 pq, 16 pq + 16) of the 64-point sub-tile.
    // Four E1 warps per scheduler: with two (32 points each) a scheduler issued one sin / cos chain at a time and
    // reached 75 % of the MUFU rate inside the batch (1039 -> 979 us per 2048^2).
    // bias'[n], wfs[o][n] are registers; db[n] and dWf[o][n] are per-thread running sums (no cross-lane reduction).
    const int q = warp & 3, pq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    {
      // W_h row n (this thread's TMEM lane), k in [64 pq, 64 pq + 64): the K-major A operand of G1, 16-bit pairs
      const float* wrow = p.w_f32 + (size_t)(h * 128 + row) * HD + pq * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t pkw[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + c * 32 + i * 4));
          pkw[2 * i] = Pack<T>::pack(v.x, v.y);
          pkw[2 * i + 1] = Pack<T>::pack(v.z, v.w);
        }
        ptx::tmem_st_32x32b_x16(tW + lane_base + pq * 32 + c * 16, pkw);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa(B.wfull, 0)); else ptx::mbar_arrive(B.wfull);
    }
    const float gscale = p.consts[CONST_SCALE], ginv = p.consts[CONST_SCALE + 1];   // power-of-two loss scale (fp16 operands)
    const float bias_n = p.consts[h * 128 + row];                 // sine: S b[n]; relu: b[n]
    float wfs_n[O];                                               // sine: S Wf[o][n]; relu: Wf[o][n]   (times the loss scale)
#pragma unroll
    for (int oo = 0; oo < O; ++oo) wfs_n[oo] = p.consts[HD + oo * HD + h * 128 + row] * gscale;
    constexpr bool PACKED = ACT == COORDX_ACT_SINE && O == 1;     // the headline case: packed fp32x2 arithmetic
    constexpr bool NOWF = dz_without_wf<T>(ACT, O);               // dZ^T without its factor wfs[n] (trunk_bwd_common.cuh)
    static_assert(!NOWF || PACKED, "dz_without_wf implies the packed path");
    float db_acc = 0.f;
    float dwf_acc[O];
#pragma unroll
    for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = 0.f;
    uint64_t db2 = 0ull, dwf2 = 0ull;                             // PACKED: (even point, odd point) partial sums
    const uint64_t S2 = f2x(p.S), B2 = f2x(bias_n);               // angle in radians (MUFU half)
    const uint64_t S2r = f2x(p.S * 0.15915494309189535f), B2r = f2x(bias_n * 0.15915494309189535f);   // in revolutions (polynomial half)
    const uint64_t W2 = f2x(wfs_n[0]);
    int dsg = 0, dsuse = 0;
    const bool tr = q == 0 && lane == 0 && (pq == 0 || pq == 3);
    const int trole = pq == 0 ? 1 : 2;
    const uint32_t zaddr0 = tZ + pq * 16 + lane_base;
    const uint32_t img_a0 = smem_base + D2_IMG + ptx::sw128_offset(row, pq * 16);
    // one sub-tile; the half hh (= Z^T buffer = image chunk) is a compile-time constant so that the tensor-memory,
    // image and barrier addresses are loop invariants (as run-time values they were ~25 of the loop's ~170 instructions)
    auto e1_sub = [&](auto HHC, const int s) {
      constexpr int hh = decltype(HHC)::value;
      constexpr uint32_t zb = hh;
      if (tr) trace_stamp(p, trole, s, 0);
      if (hh == 0) ptx::mbar_wait(B.dsfull + 8 * dsg, dsuse & 1);
      ptx::mbar_wait(B.zfull + 8 * zb, (s >> 1) & 1);
      if (tr) trace_stamp(p, trole, s, 1);
      ptx::tc_fence_after();
      const float* ds = reinterpret_cast<const float*>(smem_gen + D2_DOUT) + (dsg * TILE_M + hh * 64 + pq * 16) * O;
      const uint32_t zaddr = zaddr0 + zb * 64;
      uint32_t pk[8];
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(zaddr, r);
      if (PACKED) {
        uint64_t d2[8];                            // dout of the 16 points as (even, odd) pairs: 4 broadcast LDS.128
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ds + v * 4);
          d2[2 * v] = t.x;
          d2[2 * v + 1] = t.y;
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) {
          const uint64_t zz = ptx::pack_f32x2(__uint_as_float(r[2 * e2]), __uint_as_float(r[2 * e2 + 1]));
          uint64_t sn, cs;
          if (d2_poly_pair(e2)) {                  // polynomial share
            sincos_rev_x2(ptx::fma_f32x2(zz, S2r, B2r), sn, cs);
          } else {                                 // MUFU half
            float th0, th1;
            ptx::unpack_f32x2(ptx::fma_f32x2(zz, S2, B2), th0, th1);
            sn = ptx::pack_f32x2(__sinf(th0), __sinf(th1));
            cs = ptx::pack_f32x2(__cosf(th0), __cosf(th1));
          }
          const uint64_t dz = NOWF ? ptx::mul_f32x2(d2[e2], cs) : ptx::mul_f32x2(ptx::mul_f32x2(d2[e2], W2), cs);
          db2 = ptx::add_f32x2(db2, dz);
          dwf2 = ptx::fma_f32x2(d2[e2], sn, dwf2);
          float dz0, dz1;
          ptx::unpack_f32x2(dz, dz0, dz1);
          pk[e2] = Pack<T>::pack(dz0, dz1);
        }
      } else {
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) {
          float dz2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = e2 * 2 + u;
            const float* dp = ds + i * O;          // warp-uniform shared-memory address: broadcast
            float gw = 0.f;
#pragma unroll
            for (int oo = 0; oo < O; ++oo) gw = fmaf(dp[oo], wfs_n[oo], gw);
            const float z = __uint_as_float(r[i]);
            float av;                              // activation value (for dWf); the derivative is folded into dz
            if (ACT == COORDX_ACT_SINE) {
              const float th = fmaf(z, p.S, bias_n);
              dz2[u] = gw * __cosf(th);
              av = __sinf(th);
            } else {
              const float pre = z + bias_n;
              dz2[u] = pre > 0.f ? gw : 0.f;
              av = fmaxf(pre, 0.f);
            }
            db_acc += dz2[u];
#pragma unroll
            for (int oo = 0; oo < O; ++oo) dwf_acc[oo] = fmaf(dp[oo], av, dwf_acc[oo]);
          }
          pk[e2] = Pack<T>::pack(dz2[0], dz2[1]);
        }
      }
      if (tr) trace_stamp(p, trole, s, 2);
      if (hh == 1) {                               // last read of this tile's staged dout
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(B.dsempty + 8 * dsg);
        if (++dsg == DST) { dsg = 0; ++dsuse; }
      }
      // dZ^T goes back into TENSOR MEMORY, over the first 8 of this warp's own 16 Z^T columns (all of them are in
      // registers by now), as the A operand of G3.  No extra barrier: G1(s+2), the next writer of this buffer, is
      // issued after G3(s) in the in-order tensor pipe.
      ptx::tmem_st_32x32b_x8(zaddr, pk);
      if (p.dz_img) {                              // the image for the dfeat pass: shared memory, then one bulk store per sub-tile
        ptx::mbar_wait(B.imgempty + 8 * zb, ((s >> 1) & 1) ^ 1);
        // 2 x 8 points = two swizzled 16-byte pieces of row n (piece index 2 pq and 2 pq + 1: the addresses differ in bit 4)
        ptx::sts128(img_a0 + zb * CH, pk[0], pk[1], pk[2], pk[3]);
        ptx::sts128((img_a0 ^ 16u) + zb * CH, pk[4], pk[5], pk[6], pk[7]);
        ptx::fence_proxy_async();
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(B.dzfull + 8 * zb);
        if (PAIR && !D3_FWD) ptx::mbar_arrive_cluster(ptx::mapa(B.dzpair, 0) + 8 * zb);
      }
      if (tr) trace_stamp(p, trole, s, 3);
    };
    for (int t = 0; t < my_tiles; ++t) {
      e1_sub(std::integral_constant<int, 0>{}, 2 * t);
      e1_sub(std::integral_constant<int, 1>{}, 2 * t + 1);
    }
    if (my_tiles > 0) {
      if (PACKED) {
        float a0, a1;
        ptx::unpack_f32x2(db2, a0, a1);
        db_acc = a0 + a1;
        ptx::unpack_f32x2(dwf2, a0, a1);
        dwf_acc[0] = a0 + a1;
      }
      const float rscale = NOWF ? ginv * wfs_n[0] : ginv;         // (NOWF: the row factor the tiles were stored without)
      atomicAdd(p.db + h * 128 + row, db_acc * rscale);
#pragma unroll
      for (int oo = 0; oo < O; ++oo) atomicAdd(p.dWf + oo * HD + h * 128 + row, dwf_acc[oo]);
      // ---- flush dW_h (accumulated over every tile of this CTA) with one atomic per element
      ptx::mbar_wait(B.dwfull, 0);
      ptx::tc_fence_after();
      float* drow = p.dW + (size_t)(h * 128 + row) * HD + pq * 64;
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        uint32_t rr[16];
        ptx::tmem_ld_32x32b_x16(tDW + lane_base + pq * 64 + b * 16, rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(drow + b * 16 + i, __uint_as_float(rr[i]) * rscale);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();                   // both CTAs are done with each other's barriers / operands
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_2cta(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================
// merged feature-gradient pass from the STORED dZ^T tiles
// =================================================================================
constexpr int DF_STAGES = 5;                               // ring of 64-point sub-tiles: 2 x 16 KB each (both halves of n)
constexpr uint32_t DF_DZ = 0;
constexpr uint32_t DF_INNER = DF_DZ + DF_STAGES * 2 * CH;  // 64 KB: per-thread column slice of the inner block (prod)
constexpr uint32_t DF_BAR = DF_INNER + 4 * CH;
constexpr uint32_t DF_TOTAL = DF_BAR + 256;
// 0-15 E2 (group = warp >> 3 owns sub-tile hh = group) | 16 loader | 17 MMA | 18-19 idle (warps re-balance registers in
// groups of four).  A scheduler's register file holds 16 K registers: with 5 warps on it the launch gets 96 per thread,
// which left E2 (64 accumulators + a double-buffered 2 x 16-column tensor-memory load) one load in flight at a time;
// the loader / MMA group gives its share away (32) and the E2 warps take it (112).  Fused kernel: 72 at launch, E2 96.
#ifndef COORDX_DF_SETREG
#define COORDX_DF_SETREG 0
#endif
#ifndef COORDX_DF_DOUBLE
#define COORDX_DF_DOUBLE 0          // tensor-memory loads of E2 double-buffered (needs the 112 registers of COORDX_DF_SETREG)
#endif
constexpr int DF_THREADS = COORDX_DF_SETREG ? 20 * 32 : 18 * 32;
constexpr int DF_REGS_E2 = 112, DF_REGS_E2_FUSED = 96, DF_REGS_REST = 32;

// rowvec(o)[k] is fetched ONE TILE AHEAD as its raw 16-bit pattern and only turned into a float behind the
// accumulator wait of the tile that uses it (an opaque move pins the first use there): with the conversion next to the
// load the E2 warps sat out an L2 round trip (~1 k clk under this pass' HBM load) at the top of every tile.
__device__ __forceinline__ uint32_t pin_after_waits(uint32_t v) {
  uint32_t r;
  asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

template <typename T, int OP, bool FUSED>
__device__ __forceinline__ void dfeat_body(const BwdParams& p, uint8_t* smem_raw, const TileRuns runs) {
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + DF_BAR;
  const uint32_t b_dzfull = bar0, b_dzempty = bar0 + 8 * DF_STAGES, b_accfull = bar0 + 16 * DF_STAGES, b_accempty = b_accfull + 16,
                 b_wfull = b_accfull + 32, b_slot = b_accfull + 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + DF_BAR + 16 * DF_STAGES + 40);
  constexpr int W_LOAD = 16, W_MMA = 17;

  const int64_t n_o = p.o_end - p.o_begin;

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < DF_STAGES; ++s) {
        ptx::mbar_init(b_dzfull + 8 * s, 1);
        ptx::mbar_init(b_dzempty + 8 * s, 1);
      }
      for (int c = 0; c < 2; ++c) {
        ptx::mbar_init(b_accfull + 8 * c, 1);
        ptx::mbar_init(b_accempty + 8 * c, 256);
      }
      ptx::mbar_init(b_wfull, 512);
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
  const uint32_t tA = tmem_base, tAcc = tmem_base + 256;

  if (warp >= 16) {
    if (FUSED || COORDX_DF_SETREG) ptx::setmaxnreg_dec<DF_REGS_REST>();
  if (warp == W_LOAD) {
    // ============================ loader: ring of dZ^T sub-tiles (h = 0 and h = 1 chunks) ============================
    if (ptx::elect_one()) {
      TileRuns it = runs;
      int64_t t0;
      int cnt;
      int stg = 0, use = 0, sl = 0;
      while (it.next(t0, cnt)) {
        for (int s = 0; s < 2 * cnt; ++s) {
          const int64_t t = t0 + (s >> 1);
          const int hh = s & 1;
          if (FUSED && hh == 0) {                  // both halves, both sub-tiles of tile t stored by the dW CTAs?
            while (ptx::ld_acquire_gpu(p.flags + t) < 4u) __nanosleep(64);
            ptx::fence_proxy_async_all();          // the copy engine reads what the acquire made visible
          }
          trace_stamp(p, 3, sl, 0);
          ptx::mbar_wait(b_dzempty + 8 * stg, (use & 1) ^ 1);
          trace_stamp(p, 3, sl, 1);
          ++sl;
          ptx::mbar_arrive_expect_tx(b_dzfull + 8 * stg, 2 * CH);
          const uint8_t* src = p.dz_img + ((size_t)t * 4 + hh) * CH;
          ptx::bulk_g2s(smem_base + DF_DZ + stg * 2 * CH, src, CH, b_dzfull + 8 * stg);               // h = 0
          ptx::bulk_g2s(smem_base + DF_DZ + stg * 2 * CH + CH, src + 2 * CH, CH, b_dzfull + 8 * stg);  // h = 1
          if (++stg == DF_STAGES) { stg = 0; ++use; }
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ============================ MMA issuer: dH^T[k, p] = sum_n W^T[k, n] dZ^T[n, p] ============================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(Pack<T>::kFmt, 128, 64, 0, 1);   // A in tensor memory, B = dZ^T (N = 64 points) MN-major
      ptx::mbar_wait(b_wfull, 0);
      ptx::tc_fence_after();
      TileRuns it = runs;
      int64_t t0;
      int cnt;
      int stg = 0, use = 0, s = 0;                 // s: sub-tiles issued so far (accumulator hh = s & 1, its use count s >> 1)
      while (it.next(t0, cnt)) {
        for (int ss = 0; ss < 2 * cnt; ++ss, ++s) {
          const int hh = s & 1;
          trace_stamp(p, 0, s, 0);
          ptx::mbar_wait(b_dzfull + 8 * stg, use & 1);
          trace_stamp(p, 0, s, 1);
          ptx::mbar_wait(b_accempty + 8 * hh, ((s >> 1) & 1) ^ 1);
          trace_stamp(p, 0, s, 2);
          ptx::tc_fence_after();
          const uint32_t dz = smem_base + DF_DZ + stg * 2 * CH;
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)        // K = 256 output features n: 8 steps in the h = 0 chunk, 8 in the h = 1 chunk
              ptx::mma_f16_ts(tAcc + hh * 128 + blk * 64, tA + blk * 128 + ks * 8,
                              ptx::umma_desc_mn_sw128(dz + (ks >> 3) * CH + (ks & 7) * 2048, CH), idesc, ks != 0);
          }
          ptx::mma_commit(b_accfull + 8 * hh);
          ptx::mma_commit(b_dzempty + 8 * stg);
          trace_stamp(p, 0, s, 3);
          if (++stg == DF_STAGES) { stg = 0; ++use; }
        }
      }
    }
    __syncwarp();
  }
  } else {
    if (FUSED || COORDX_DF_SETREG) ptx::setmaxnreg_inc<FUSED ? DF_REGS_E2_FUSED : DF_REGS_E2>();
    // ============================ E2: dH^T sub-tile -> d rowvec(o) (atomics per tile) and d inner (registers) ============================
    const int g = warp >> 3, q = warp & 3, blk = (warp >> 2) & 1;
    const int k = blk * 128 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    {
      // W^T: lane k, n in [128 g, 128 g + 128) as 16-bit pairs: columns 64 g .. 64 g + 63 of this M-block's A operand
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pkw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = g * 128 + c * 32 + 2 * i;
          const float f0 = p.dz_no_wf ? p.consts[HD + n] : 1.f, f1 = p.dz_no_wf ? p.consts[HD + n + 1] : 1.f;
          pkw[i] = Pack<T>::pack(f0 * __ldg(p.w_f32 + (size_t)n * HD + k), f1 * __ldg(p.w_f32 + (size_t)(n + 1) * HD + k));
        }
        ptx::tmem_st_32x32b_x16(tA + lane_base + blk * 128 + g * 64 + c * 16, pkw);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(b_wfull);
    }
    const float ginv = p.consts[CONST_SCALE + 1];
    uint4* my_inner = reinterpret_cast<uint4*>(smem_gen + DF_INNER) + threadIdx.x;     // slot i at [i * 512 + thread]: conflict-free
    // d(inner) sums of this thread's 64 points as 32 (even point, odd point) pairs: all E2 arithmetic is packed fp32x2
    uint64_t din2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) din2[i] = 0ull;
    auto load_inner = [&](int64_t jb) {            // this thread's column k of the 64 inner rows of its sub-tile, 16-bit
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t j = jb * TILE_M + g * 64 + c * 8 + i;
          v[i] = j < p.n_inner ? __ldg(p.inner_f32 + j * HD + k) : 0.f;
        }
        my_inner[c * 512] = make_uint4(Pack<T>::pack(v[0], v[1]), Pack<T>::pack(v[2], v[3]), Pack<T>::pack(v[4], v[5]),
                                       Pack<T>::pack(v[6], v[7]));
      }
    };
    auto flush_din = [&](int64_t jb) {
      float* dcol = p.d_inner + (jb * TILE_M + g * 64) * HD + k;        // lanes = consecutive k: coalesced atomics
#pragma unroll
      for (int i2 = 0; i2 < 32; ++i2) {
        float d0, d1;
        ptx::unpack_f32x2(din2[i2], d0, d1);
        if (jb * TILE_M + g * 64 + 2 * i2 < p.n_inner) atomicAdd(dcol + (size_t)(2 * i2) * HD, d0 * ginv);
        if (jb * TILE_M + g * 64 + 2 * i2 + 1 < p.n_inner) atomicAdd(dcol + (size_t)(2 * i2 + 1) * HD, d1 * ginv);
        din2[i2] = 0ull;
      }
    };
    auto load_rv = [&](int o_local) -> uint32_t {  // rowvec(o)[k]: the raw 16-bit value the forward multiplied with
      if (OP != COORDX_FUSE_PROD) return 0u;
      return __ldg(reinterpret_cast<const uint16_t*>(p.outer_img) + (size_t)o_local * HD + k);
    };
    TileWalk w(runs, n_o);
    int64_t cur_jb = -1;                           // block of inner rows whose 16-bit column is staged / whose d(inner) sums are live
    bool dirty = false;
    uint32_t rv_next = w.valid ? load_rv(w.o) : 0u;
    const uint32_t tacc = tAcc + g * 128 + blk * 64 + lane_base;
    const bool tr = (warp & 7) == 0 && lane == 0;
    for (int tc = 0; w.valid; ++tc) {              // tc: tiles done so far (accumulator phase)
      const uint32_t rv_raw = rv_next;
      const int64_t jb = w.jb;
      const int o_local = w.o;
      w.advance();
      if (w.valid) rv_next = load_rv(w.o);                             // in flight during this tile
      if (jb != cur_jb) {                          // new block of inner rows (stand-alone pass: the sweep over o is complete)
        if (dirty) flush_din(cur_jb);
        dirty = false;
        if (OP == COORDX_FUSE_PROD) load_inner(jb);
        cur_jb = jb;
      }
      if (tr) trace_stamp(p, 1 + g, tc, 0);
      ptx::mbar_wait(b_accfull + 8 * g, tc & 1);
      if (tr) trace_stamp(p, 1 + g, tc, 1);
      ptx::tc_fence_after();
      // The 64 columns come in four batches of 16; batch b + 1 is in flight while batch b is consumed (one batch at a
      // time exposed the tensor-memory latency four times per tile: the pass was bound by E2, not by HBM -
      // profiles/r02h_dfeat_stage_timing.txt).
      uint64_t acc2 = 0ull;
      constexpr bool DBL = COORDX_DF_DOUBLE && !FUSED;
      uint32_t r[2][16];
      if (DBL) ptx::tmem_ld_32x32b_x16(tacc, r[0]);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (!DBL) ptx::tmem_ld_32x32b_x16(tacc + b * 16, r[b & 1]);
        uint4 iv[2];
        if (OP == COORDX_FUSE_PROD) {
          iv[0] = my_inner[(2 * b) * 512];
          iv[1] = my_inner[(2 * b + 1) * 512];
        }
        ptx::tmem_ld_wait();
        if (b < 3) {
          if (DBL) ptx::tmem_ld_32x32b_x16(tacc + (b + 1) * 16, r[(b + 1) & 1]);
        } else {
          ptx::tc_fence_before();
          ptx::mbar_arrive(b_accempty + 8 * g);
          if (tr) trace_stamp(p, 1 + g, tc, 2);
        }
        const float rv = OP == COORDX_FUSE_PROD ? Pack<T>::unpack(pin_after_waits(rv_raw), 0) : 1.f;
        const uint64_t rv2 = ptx::pack_f32x2(rv, rv);
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
          const uint64_t dh2 = ptx::pack_f32x2(__uint_as_float(r[b & 1][2 * i2]), __uint_as_float(r[b & 1][2 * i2 + 1]));
          if (OP == COORDX_FUSE_PROD) {
            const uint4 v = iv[i2 >> 2];
            const uint32_t w16 = (i2 & 3) == 0 ? v.x : ((i2 & 3) == 1 ? v.y : ((i2 & 3) == 2 ? v.z : v.w));
            acc2 = ptx::fma_f32x2(dh2, ptx::pack_f32x2(Pack<T>::unpack(w16, 0), Pack<T>::unpack(w16, 1)), acc2);
            din2[b * 8 + i2] = ptx::fma_f32x2(rv2, dh2, din2[b * 8 + i2]);
          } else {
            acc2 = ptx::add_f32x2(acc2, dh2);
            din2[b * 8 + i2] = ptx::add_f32x2(din2[b * 8 + i2], dh2);
          }
        }
      }
      float a0, a1;
      ptx::unpack_f32x2(acc2, a0, a1);
      atomicAdd(p.d_rv + (int64_t)o_local * HD + k, (a0 + a1) * ginv);
      if (tr) trace_stamp(p, 1 + g, tc, 3);
      dirty = true;
    }
    if (dirty) flush_din(cur_jb);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================
// feature-gradient pass as CTA pairs
// =================================================================================
// The pipeline trace of the single-CTA pass (scripts/dfeat_trace.py, profiles/r02i_dfeat_trace.txt) shows what bounds
// it: neither HBM (the loader waits ~1.1 k clk per sub-tile for a free stage) nor E2 (it waits 0.7-1.4 k clk per tile
// for its accumulator) but the tensor pipe, whose M128 x N64 x K16 MMAs take 44.5 clk instead of 32: every MMA fetches
// 2 KB of dZ^T from shared memory - 64 B/clk at full rate, the ~70 B/clk fetch limit of section 3.2 in DESIGN.md
// next to the 23 B/clk the copy engine writes - and the tile is fetched TWICE, once per 128-row block of W^T.
// Here the two row blocks belong to the two CTAs of a cluster: one cta_group::2 MMA M256 (k) x N128 (the points of a
// whole tile) x K16 (n) per step, A = this CTA's half of W^T in its tensor memory (128 columns), B = dZ^T with each
// CTA supplying ONE 64-point sub-tile image (so a CTA also loads only half of every tile from HBM).  Operand fetch per
// SM: 2 KB per 64 clk.  The accumulator of a tile is [128 k x 128 points] = 128 columns: three of them rotate.
#ifndef COORDX_DP_AHEAD
#define COORDX_DP_AHEAD 3           // tiles the MMA issue may run ahead of completed tiles (3 = unlimited: one per accumulator;
                                    // measured per 2048^2: 3: 382.5 us, 2: 380.7 us, 1: 453 us - the queued tiles hold ring stages)
#endif
constexpr int DP_STAGES = 6;                               // ring of sub-tile images (2 x 16 KB: both halves of n)
constexpr int DP_ACC = 3;
constexpr uint32_t DP_DZ = 0;
constexpr uint32_t DP_INNER = DP_DZ + DP_STAGES * 2 * CH;  // 32 KB: per-thread column slice (32 points) of the inner block
constexpr uint32_t DP_BAR = DP_INNER + 2 * CH;
constexpr uint32_t DP_TOTAL = DP_BAR + 256;
constexpr int DP_THREADS = 18 * 32;                        // 0-15 E2 (lane quadrant = warp & 3, 32-point column group = warp >> 2) | 16 loader | 17 MMA (leader) / forwarder (peer)

template <typename T, int OP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DP_THREADS, 1)
trunk_bwd_dfeat_pair_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();    // = the 128-row block of W^T (features k) this CTA owns
  const uint32_t bar0 = smem_base + DP_BAR;
  const uint32_t b_dzfull = bar0;                            // [6] this CTA's sub-tile image has landed (complete_tx)
  const uint32_t b_dzpeer = bar0 + 8 * DP_STAGES;            // [6] leader only: the peer's image has landed (forwarded)
  const uint32_t b_dzempty = bar0 + 16 * DP_STAGES;          // [6] each CTA: the MMAs reading the stage are complete (multicast commit)
  const uint32_t b_accfull = bar0 + 24 * DP_STAGES;          // [3] each CTA: accumulator complete (multicast commit)
  const uint32_t b_accempty = b_accfull + 8 * DP_ACC;        // [3] leader only: 32 E2 warps (16 per CTA) have drained it
  const uint32_t b_wfull = b_accempty + 8 * DP_ACC;          // leader only: W^T of both CTAs is in tensor memory
  const uint32_t b_slot = b_wfull + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + DP_BAR + 24 * DP_STAGES + 16 * DP_ACC + 8);
  constexpr int W_LOAD = 16, W_MMA = 17;

  const int64_t n_o = p.o_end - p.o_begin;
  const int64_t total = (int64_t)p.nb * n_o;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const TileRuns runs = TileRuns::single(total * pair / n_pairs, total * (pair + 1) / n_pairs);
  const int my_tiles = runs.count();

  if (warp == W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < DP_STAGES; ++s) {
        ptx::mbar_init(b_dzfull + 8 * s, 1);
        ptx::mbar_init(b_dzpeer + 8 * s, 1);
        ptx::mbar_init(b_dzempty + 8 * s, 1);
      }
      for (int c = 0; c < DP_ACC; ++c) {
        ptx::mbar_init(b_accfull + 8 * c, 1);
        ptx::mbar_init(b_accempty + 8 * c, 32);
      }
      ptx::mbar_init(b_wfull, 1024);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc_2cta(b_slot, 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                             // the peer's barriers and tensor memory exist
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tA = tmem_base, tAcc = tmem_base + 128;

  if (warp == W_LOAD) {
    // ============================ loader: this CTA's sub-tile (hh = rank) of every tile ============================
    if (ptx::elect_one()) {
      TileWalk w(runs, n_o);
      int stg = 0, use = 0;
      for (int t = 0; t < my_tiles; ++t) {
        ptx::mbar_wait(b_dzempty + 8 * stg, (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(b_dzfull + 8 * stg, 2 * CH);
        const uint8_t* src = p.dz_img + ((size_t)w.t * 4 + rank) * CH;
        ptx::bulk_g2s(smem_base + DP_DZ + stg * 2 * CH, src, CH, b_dzfull + 8 * stg);               // h = 0
        ptx::bulk_g2s(smem_base + DP_DZ + stg * 2 * CH + CH, src + 2 * CH, CH, b_dzfull + 8 * stg);  // h = 1
        if (++stg == DP_STAGES) { stg = 0; ++use; }
        w.advance();
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    if (rank == 0) {
      // ============================ MMA issuer (leader): dH^T[k, p] = sum_n W^T[k, n] dZ^T[n, p] ============================
      if (ptx::elect_one()) {
        constexpr uint32_t idesc = ptx::umma_idesc_f16(Pack<T>::kFmt, 256, 128, 0, 1);   // A in tensor memory, B = dZ^T (N = points) MN-major
        ptx::mbar_wait(b_wfull, 0);
        ptx::tc_fence_after();
        int stg = 0, use = 0, acc = 0, ause = 0;
        for (int t = 0; t < my_tiles; ++t) {
          trace_stamp(p, 0, t, 0);
          if (COORDX_DP_AHEAD < DP_ACC && t >= COORDX_DP_AHEAD) {
            // do not run more than COORDX_DP_AHEAD tiles ahead of the tensor pipe: every queued tile holds its stage of the
            // ring, i.e. costs the loader a stage of prefetch depth
            const int td = t - COORDX_DP_AHEAD;    // tile whose completion is awaited: accumulator td % 3, its use td / 3
            ptx::mbar_wait(b_accfull + 8 * (td % DP_ACC), (td / DP_ACC) & 1);
          }
          ptx::mbar_wait(b_dzfull + 8 * stg, use & 1);
          ptx::mbar_wait(b_dzpeer + 8 * stg, use & 1);
          trace_stamp(p, 0, t, 1);
          ptx::mbar_wait(b_accempty + 8 * acc, (ause & 1) ^ 1);
          trace_stamp(p, 0, t, 2);
          ptx::tc_fence_after();
          const uint32_t dz = smem_base + DP_DZ + stg * 2 * CH;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)          // K = 256 output features n: 8 steps in the h = 0 chunk, 8 in the h = 1 chunk
            ptx::mma_f16_ts_2cta(tAcc + acc * 128, tA + ks * 8, ptx::umma_desc_mn_sw128(dz + (ks >> 3) * CH + (ks & 7) * 2048, CH), idesc, ks != 0);
          ptx::mma_commit_2cta(b_accfull + 8 * acc, 3);
          ptx::mma_commit_2cta(b_dzempty + 8 * stg, 3);
          trace_stamp(p, 0, t, 3);
          if (++stg == DP_STAGES) { stg = 0; ++use; }
          if (++acc == DP_ACC) { acc = 0; ++ause; }
        }
      }
    } else {
      // ============================ peer: forward "my image of stage s has landed" to the leader ============================
      if (ptx::elect_one()) {
        const uint32_t remote = ptx::mapa(b_dzpeer, 0);
        int stg = 0, use = 0;
        for (int t = 0; t < my_tiles; ++t) {
          ptx::mbar_wait(b_dzfull + 8 * stg, use & 1);
          ptx::mbar_arrive_cluster(remote + 8 * stg);
          if (++stg == DP_STAGES) { stg = 0; ++use; }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================ E2: dH^T tile -> d rowvec(o) (atomics per tile) and d inner (registers) ============================
    const int q = warp & 3, cg = warp >> 2;
    const int k = rank * 128 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    {
      // W^T: lane k, n in [64 cg, 64 cg + 64) as 16-bit pairs: columns 32 cg .. 32 cg + 31 of the A operand
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t pkw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = cg * 64 + c * 32 + 2 * i;
          const float f0 = p.dz_no_wf ? p.consts[HD + n] : 1.f, f1 = p.dz_no_wf ? p.consts[HD + n + 1] : 1.f;
          pkw[i] = Pack<T>::pack(f0 * __ldg(p.w_f32 + (size_t)n * HD + k), f1 * __ldg(p.w_f32 + (size_t)(n + 1) * HD + k));
        }
        ptx::tmem_st_32x32b_x16(tA + lane_base + cg * 32 + c * 16, pkw);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(ptx::mapa(b_wfull, 0));
    }
    const float ginv = p.consts[CONST_SCALE + 1];
    const uint32_t accempty_remote = ptx::mapa(b_accempty, 0);
    uint4* my_inner = reinterpret_cast<uint4*>(smem_gen + DP_INNER) + threadIdx.x;     // slot i at [i * 512 + thread]: conflict-free
    uint64_t din2[16];                             // d(inner) sums of this thread's 32 points as (even, odd) pairs
#pragma unroll
    for (int i = 0; i < 16; ++i) din2[i] = 0ull;
    auto load_inner = [&](int64_t jb) {            // this thread's column k of its 32 inner rows, 16-bit
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t j = jb * TILE_M + cg * 32 + c * 8 + i;
          v[i] = j < p.n_inner ? __ldg(p.inner_f32 + j * HD + k) : 0.f;
        }
        my_inner[c * 512] = make_uint4(Pack<T>::pack(v[0], v[1]), Pack<T>::pack(v[2], v[3]), Pack<T>::pack(v[4], v[5]),
                                       Pack<T>::pack(v[6], v[7]));
      }
    };
    auto flush_din = [&](int64_t jb) {
      float* dcol = p.d_inner + (jb * TILE_M + cg * 32) * HD + k;       // lanes = consecutive k: coalesced atomics
#pragma unroll
      for (int i2 = 0; i2 < 16; ++i2) {
        float d0, d1;
        ptx::unpack_f32x2(din2[i2], d0, d1);
        if (jb * TILE_M + cg * 32 + 2 * i2 < p.n_inner) atomicAdd(dcol + (size_t)(2 * i2) * HD, d0 * ginv);
        if (jb * TILE_M + cg * 32 + 2 * i2 + 1 < p.n_inner) atomicAdd(dcol + (size_t)(2 * i2 + 1) * HD, d1 * ginv);
        din2[i2] = 0ull;
      }
    };
    auto load_rv = [&](int o_local) -> uint32_t {  // rowvec(o)[k]: the raw 16-bit value the forward multiplied with
      if (OP != COORDX_FUSE_PROD) return 0u;
      return __ldg(reinterpret_cast<const uint16_t*>(p.outer_img) + (size_t)o_local * HD + k);
    };
    TileWalk w(runs, n_o);
    int64_t cur_jb = -1;
    bool dirty = false;
    uint32_t rv_next = w.valid ? load_rv(w.o) : 0u;
    const bool tr = (warp == 0 || warp == 12) && lane == 0;
    const int trole = warp == 0 ? 1 : 2;
    int acc = 0, ause = 0;
    for (int tc = 0; w.valid; ++tc) {
      const uint32_t rv_raw = rv_next;
      const int64_t jb = w.jb;
      const int o_local = w.o;
      w.advance();
      if (w.valid) rv_next = load_rv(w.o);                             // in flight during this tile
      if (jb != cur_jb) {                          // new block of inner rows: the sweep over o is complete
        if (dirty) flush_din(cur_jb);
        dirty = false;
        if (OP == COORDX_FUSE_PROD) load_inner(jb);
        cur_jb = jb;
      }
      if (tr) trace_stamp(p, trole, tc, 0);
      ptx::mbar_wait(b_accfull + 8 * acc, ause & 1);
      if (tr) trace_stamp(p, trole, tc, 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tAcc + acc * 128 + cg * 32 + lane_base;
      uint64_t acc2 = 0ull;
      uint32_t r[2][16];
      ptx::tmem_ld_32x32b_x16(tacc, r[0]);
      ptx::tmem_ld_32x32b_x16(tacc + 16, r[1]);
      uint4 iv[4];
      if (OP == COORDX_FUSE_PROD) {
#pragma unroll
        for (int c = 0; c < 4; ++c) iv[c] = my_inner[c * 512];
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(accempty_remote + 8 * acc);
      if (tr) trace_stamp(p, trole, tc, 2);
      const float rv = OP == COORDX_FUSE_PROD ? Pack<T>::unpack(pin_after_waits(rv_raw), 0) : 1.f;
      const uint64_t rv2 = ptx::pack_f32x2(rv, rv);
#pragma unroll
      for (int i2 = 0; i2 < 16; ++i2) {
        const uint64_t dh2 = ptx::pack_f32x2(__uint_as_float(r[i2 >> 3][2 * (i2 & 7)]), __uint_as_float(r[i2 >> 3][2 * (i2 & 7) + 1]));
        if (OP == COORDX_FUSE_PROD) {
          const uint4 v = iv[i2 >> 2];
          const uint32_t w16 = (i2 & 3) == 0 ? v.x : ((i2 & 3) == 1 ? v.y : ((i2 & 3) == 2 ? v.z : v.w));
          acc2 = ptx::fma_f32x2(dh2, ptx::pack_f32x2(Pack<T>::unpack(w16, 0), Pack<T>::unpack(w16, 1)), acc2);
          din2[i2] = ptx::fma_f32x2(rv2, dh2, din2[i2]);
        } else {
          acc2 = ptx::add_f32x2(acc2, dh2);
          din2[i2] = ptx::add_f32x2(din2[i2], dh2);
        }
      }
      float a0, a1;
      ptx::unpack_f32x2(acc2, a0, a1);
      atomicAdd(p.d_rv + (int64_t)o_local * HD + k, (a0 + a1) * ginv);
      if (tr) trace_stamp(p, trole, tc, 3);
      dirty = true;
      if (++acc == DP_ACC) { acc = 0; ++ause; }
    }
    if (dirty) flush_din(cur_jb);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                             // both CTAs are done with each other's barriers / operands
  if (warp == W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

// =================================================================================
// kernels: the two passes on their own, and both in ONE launch
// =================================================================================
template <typename T, int OP, int ACT, int O>
__global__ void __launch_bounds__(DW2_THREADS, 1)
trunk_bwd_dw2_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int h = blockIdx.x & 1;                    // half of the trunk's output features
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t total = (int64_t)p.nb * (p.o_end - p.o_begin);
  dw2_body<T, OP, ACT, O, false>(p, smem_raw, h, TileRuns::single(total * cta / n_cta, total * (cta + 1) / n_cta));
}

// the dW pass as CTA pairs (see dw2_body, PAIR): blocks 2 i and 2 i + 1 are one cluster
template <typename T, int OP, int ACT, int O>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DW2_THREADS, 1)
trunk_bwd_dw3_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int h = (int)ptx::cluster_ctarank();
  const int n_cta = gridDim.x >> 1, cta = blockIdx.x >> 1;
  const int64_t total = (int64_t)p.nb * (p.o_end - p.o_begin);
  dw2_body<T, OP, ACT, O, true>(p, smem_raw, h, TileRuns::single(total * cta / n_cta, total * (cta + 1) / n_cta));
}

template <typename T, int OP>
__global__ void __launch_bounds__(DF_THREADS, 1)
trunk_bwd_dfeat_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int n_cta = gridDim.x, cta = blockIdx.x;
  const int64_t total = (int64_t)p.nb * (p.o_end - p.o_begin);
  dfeat_body<T, OP, false>(p, smem_raw, TileRuns::single(total * cta / n_cta, total * (cta + 1) / n_cta));
}

// Fused backward: CTAs 0 .. 2 p.fused_pairs - 1 run the dW pass (pair = blockIdx / 2, its two CTAs the two halves of the
// output features), the remaining CTAs run the feature-gradient pass BEHIND them on the tiles the pairs have already
// stored.  Run back to back, the dW pass is bound by the MUFU / tensor pipes (HBM idle but for 2 GB of writes) and the
// feature-gradient pass by HBM (2.1 GB read at ~4.8 TB/s, tensor pipe 30 %, profiles/r02b_bwd_dfeat_full.txt); side
// by side the dZ^T tiles are consumed out of the L2 shortly after they were written.  The CTAs form p.fused_groups
// groups (pairs and consumers dealt out evenly); a group sweeps its own range of tiles - sized by its number of
// pairs - as described at TileRuns.  One CTA per SM and a grid of at most the SM count keep every CTA resident; the
// dW CTAs have the lower block indices and never wait for anybody, so a consumer's spin always ends.
template <typename T, int OP, int ACT, int O>
__global__ void __launch_bounds__(DW2_THREADS, 1)
trunk_bwd_fused_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t total = (int64_t)p.nb * (p.o_end - p.o_begin);
  const int n_pairs = p.fused_pairs, n_prod = 2 * n_pairs, n_cons = (int)gridDim.x - n_prod, NG = p.fused_groups;
  const bool prod = (int)blockIdx.x < n_prod;
  const int me = prod ? (int)blockIdx.x >> 1 : (int)blockIdx.x - n_prod;       // pair index / consumer index
  const int n_mine = prod ? n_pairs : n_cons;
  int g = (int)(((int64_t)me * NG) / n_mine);      // group g holds members [n g / NG, n (g + 1) / NG) of either kind
  while (g + 1 < NG && (int64_t)n_mine * (g + 1) / NG <= me) ++g;
  while (g > 0 && (int64_t)n_mine * g / NG > me) --g;
  const int p0 = (int)((int64_t)n_pairs * g / NG), p1 = (int)((int64_t)n_pairs * (g + 1) / NG);
  const int c0 = (int)((int64_t)n_cons * g / NG), c1 = (int)((int64_t)n_cons * (g + 1) / NG);
  const int64_t rb = total * p0 / n_pairs, re = total * p1 / n_pairs;
  const int G = p.fused_chunk < 0 ? -p.fused_chunk : p.fused_chunk;      // (< 0: timing aid, the consumers do nothing)
  if (prod) {
    dw2_body<T, OP, ACT, O, false>(p, smem_raw, blockIdx.x & 1, TileRuns::make(rb, re, G, me - p0, p1 - p0));
  } else {
    dfeat_body<T, OP, true>(p, smem_raw, TileRuns::make(rb, c1 > c0 && p.fused_chunk > 0 ? re : rb, G, me - c0, c1 - c0));
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <typename T, int OP, int ACT, int O>
static int launch_fused_inst(const BwdParams& bp, int grid, cudaStream_t st) {
  static AttrOnce once;
  auto kern = trunk_bwd_fused_kernel<T, OP, ACT, O>;
  constexpr uint32_t smem = d2_smem_bytes(O) > DF_TOTAL ? d2_smem_bytes(O) : DF_TOTAL;
  CX_TRY(once.set(kern, (int)smem));
  kern<<<grid, DW2_THREADS, smem, st>>>(bp);
  CX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int OP, int ACT>
static int launch_fused_o(const BwdParams& bp, int O, int grid, cudaStream_t st) {
  switch (O) {
    case 1: return launch_fused_inst<T, OP, ACT, 1>(bp, grid, st);
    case 2: return launch_fused_inst<T, OP, ACT, 2>(bp, grid, st);
    case 3: return launch_fused_inst<T, OP, ACT, 3>(bp, grid, st);
    case 4: return launch_fused_inst<T, OP, ACT, 4>(bp, grid, st);
  }
  return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk backward: out width %s%lld > 4", "", O);
}

template <typename T>
static int launch_fused_t(const BwdParams& bp, int op, int act, int O, int grid, cudaStream_t st) {
  if (op == COORDX_FUSE_PROD)
    return act == COORDX_ACT_SINE ? launch_fused_o<T, COORDX_FUSE_PROD, COORDX_ACT_SINE>(bp, O, grid, st)
                                  : launch_fused_o<T, COORDX_FUSE_PROD, COORDX_ACT_RELU>(bp, O, grid, st);
  return act == COORDX_ACT_SINE ? launch_fused_o<T, COORDX_FUSE_SUM, COORDX_ACT_SINE>(bp, O, grid, st)
                                : launch_fused_o<T, COORDX_FUSE_SUM, COORDX_ACT_RELU>(bp, O, grid, st);
}

int launch_bwd_fused(const BwdParams& bp, int precision, int op, int act, int O, int grid, cudaStream_t st) {
  if (precision == COORDX_PREC_BF16) return launch_fused_t<__nv_bfloat16>(bp, op, act, O, grid, st);
  return launch_fused_t<__half>(bp, op, act, O, grid, st);
}

// COORDX_BWD_PAIR=1 selects the CTA-pair form of the dW pass (trunk_bwd_dw3_kernel).  Opt-in: it halves the tensor
// cores' operand fetch and takes the E1 warps' wait for Z^T from ~550 to ~180 clk per sub-tile, but E1 itself - the
// sin / cos arithmetic plus its tail, ~1.5 k clk per sub-tile - is what bounds the pass, and the pair's producers
// write two operand layouts per tile: 1024-1035 us against 970 us for the single-CTA form (profiles/r02g_dw_pair.txt).
static bool dw_pair_wanted() {
  const char* e = getenv("COORDX_BWD_PAIR");
  return e && e[0] == '1';
}

template <typename T, int OP, int ACT, int O>
static int launch_dw2_inst(const BwdParams& bp, int grid, cudaStream_t st) {
  static AttrOnce once, once3;
  if (dw_pair_wanted()) {
    auto kern = trunk_bwd_dw3_kernel<T, OP, ACT, O>;
    CX_TRY(once3.set(kern, (int)d2_smem_bytes(O)));
    kern<<<grid, DW2_THREADS, d2_smem_bytes(O), st>>>(bp);
  } else {
    auto kern = trunk_bwd_dw2_kernel<T, OP, ACT, O>;
    CX_TRY(once.set(kern, (int)d2_smem_bytes(O)));
    kern<<<grid, DW2_THREADS, d2_smem_bytes(O), st>>>(bp);
  }
  CX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int OP, int ACT>
static int launch_dw2_o(const BwdParams& bp, int O, int grid, cudaStream_t st) {
  switch (O) {
    case 1: return launch_dw2_inst<T, OP, ACT, 1>(bp, grid, st);
    case 2: return launch_dw2_inst<T, OP, ACT, 2>(bp, grid, st);
    case 3: return launch_dw2_inst<T, OP, ACT, 3>(bp, grid, st);
    case 4: return launch_dw2_inst<T, OP, ACT, 4>(bp, grid, st);
  }
  return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk backward: out width %s%lld > 4", "", O);
}

template <typename T>
static int launch_dw2_t(const BwdParams& bp, int op, int act, int O, int grid, cudaStream_t st) {
  if (op == COORDX_FUSE_PROD)
    return act == COORDX_ACT_SINE ? launch_dw2_o<T, COORDX_FUSE_PROD, COORDX_ACT_SINE>(bp, O, grid, st)
                                  : launch_dw2_o<T, COORDX_FUSE_PROD, COORDX_ACT_RELU>(bp, O, grid, st);
  return act == COORDX_ACT_SINE ? launch_dw2_o<T, COORDX_FUSE_SUM, COORDX_ACT_SINE>(bp, O, grid, st)
                                : launch_dw2_o<T, COORDX_FUSE_SUM, COORDX_ACT_RELU>(bp, O, grid, st);
}

int launch_dw2(const BwdParams& bp, int precision, int op, int act, int O, int grid, cudaStream_t st) {
  if (precision == COORDX_PREC_BF16) return launch_dw2_t<__nv_bfloat16>(bp, op, act, O, grid, st);
  return launch_dw2_t<__half>(bp, op, act, O, grid, st);
}

// COORDX_BWD_DFEAT_PAIR=0 selects the single-CTA form of the feature-gradient pass (A/B timing)
static bool dfeat_pair_wanted() {
  const char* e = getenv("COORDX_BWD_DFEAT_PAIR");
  return !(e && e[0] == '0');
}

template <typename T, int OP>
static int launch_dfeat_inst(const BwdParams& bp, int n_cta, cudaStream_t st) {
  static AttrOnce once, once_pair;
  if (dfeat_pair_wanted() && n_cta >= 2) {
    auto kern = trunk_bwd_dfeat_pair_kernel<T, OP>;
    CX_TRY(once_pair.set(kern, (int)DP_TOTAL));
    kern<<<n_cta & ~1, DP_THREADS, DP_TOTAL, st>>>(bp);
  } else {
    auto kern = trunk_bwd_dfeat_kernel<T, OP>;
    CX_TRY(once.set(kern, (int)DF_TOTAL));
    kern<<<n_cta, DF_THREADS, DF_TOTAL, st>>>(bp);
  }
  CX_LAUNCH_CHECK();
  return 0;
}

int launch_dfeat(const BwdParams& bp, int precision, int op, int n_cta, cudaStream_t st) {
  if (precision == COORDX_PREC_BF16)
    return op == COORDX_FUSE_PROD ? launch_dfeat_inst<__nv_bfloat16, COORDX_FUSE_PROD>(bp, n_cta, st)
                                  : launch_dfeat_inst<__nv_bfloat16, COORDX_FUSE_SUM>(bp, n_cta, st);
  return op == COORDX_FUSE_PROD ? launch_dfeat_inst<__half, COORDX_FUSE_PROD>(bp, n_cta, st)
                                : launch_dfeat_inst<__half, COORDX_FUSE_SUM>(bp, n_cta, st);
}

}  // namespace tcb
}  // namespace coordx
