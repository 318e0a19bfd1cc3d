// Fused per-axis branch MLP forward (FP32, the 1e-5 parity arithmetic): SplitFCBlock.forward_channel
// (modules.py:311-327) for ALL branches in one launch.  One CTA pushes 16 axis samples through
// coord_linears[c] and net[0..n_pre-1]; the activations never leave shared memory, the weights are
// streamed through L1 (each thread walks two rows of W), and only the final [16, 256] feature rows are
// written.  Covers the 256-wide nets (pre_width = every pre-fusion layer = 256); anything else takes the
// generic per-layer GEMM path in coordx_abi.cu.
#pragma once
#include "common.cuh"

namespace coordx {
namespace bf {

constexpr int WIDTH = 256;
constexpr int ROWS = 16;
constexpr int THREADS = 128;
constexpr int MAX_F = 256;                    // widest first-layer input (positional encodings: 21, 170)
constexpr int KC = 16;                        // k-chunk of W staged through shared memory
constexpr int WS_LD = KC + 1;                 // padded row stride (words): conflict-free column walks
constexpr int SMEM_BYTES = (2 * ROWS * WIDTH + 2 * WIDTH * WS_LD + ROWS * MAX_F) * 4;

struct BranchArgs {
  int n_branches, n_pre, act, last_noact;
  float S, inv_c;
  const float* coords[COORDX_MAX_BRANCHES];   // [n_c, f_c]
  float* feat[COORDX_MAX_BRANCHES];           // [n_c, 256]
  int64_t n[COORDX_MAX_BRANCHES];
  int f[COORDX_MAX_BRANCHES];
  int blk0[COORDX_MAX_BRANCHES + 1];          // first block index of each branch
  const float* first_w[COORDX_MAX_BRANCHES];
  const float* first_b[COORDX_MAX_BRANCHES];
  const float* w[COORDX_MAX_LAYERS];
  const float* b[COORDX_MAX_LAYERS];
  // training mode (nullable): pre-activation / activation of every layer, [n_c, 256] each, kept for the backward
  // pass (layer 0 = coord_linears[c], layer l + 1 = net[l]); the last layer's activation is `feat` itself
  float* zsave[COORDX_MAX_BRANCHES][COORDX_MAX_LAYERS + 1];
  float* asave[COORDX_MAX_BRANCHES][COORDX_MAX_LAYERS + 1];
};

// coalesced copy of the CTA's [16, 256] activation rows to a [n, 256] global array
// (dst is CTA-uniform; sync_after: the rows are overwritten in place by the activation pass that follows)
__device__ __forceinline__ void store_rows(float* dst, const float* h, int64_t r0, int64_t n, int tid, bool sync_after) {
  if (!dst) return;
  for (int i = tid; i < 16 * 256 / 4; i += 128) {
    const int r = i / (256 / 4);
    if (r0 + r < n) reinterpret_cast<float4*>(dst + (r0 + r) * 256)[i % (256 / 4)] = reinterpret_cast<const float4*>(h + r * 256)[i % (256 / 4)];
  }
  if (sync_after) __syncthreads();
}

__device__ __forceinline__ float act_exact(int act, float S, float z) { return act_fwd(act, S, z); }

constexpr int ROWS_ = 16, WIDTH_ = 256, THREADS_ = 128;
__device__ __noinline__ void apply_act(float* h, int act, float S, int tid) {
#pragma unroll 1
  for (int i = tid; i < ROWS_ * WIDTH_; i += THREADS_) h[i] = act_exact(act, S, h[i]);
}

// Thread (cg = tid % 64, rg = tid / 64) owns the 8 x 4 register tile rows 8 rg .. 8 rg + 7 x columns
// cg, cg + 64, cg + 128, cg + 192.  W ([256 out, 256 in] row-major) is streamed in k-chunks of 16 through a
// double-buffered, padded shared-memory tile with coalesced 64-byte row segments; per 4 k a warp issues
// 16 conflict-free LDS (weights) + 8 broadcast LDS.128 (activations) for 128 FFMAs.
__global__ void __launch_bounds__(THREADS)
branch_fused_fwd_kernel(const BranchArgs a) {
  extern __shared__ __align__(16) float smem[];
  float (*hbuf)[ROWS][WIDTH] = reinterpret_cast<float (*)[ROWS][WIDTH]>(smem);                  // [2] ping-pong activations
  float (*wsb)[WIDTH * WS_LD] = reinterpret_cast<float (*)[WIDTH * WS_LD]>(smem + 2 * ROWS * WIDTH);   // [2] W chunk [j][kk]
  float (*xs)[MAX_F] = reinterpret_cast<float (*)[MAX_F]>(smem + 2 * ROWS * WIDTH + 2 * WIDTH * WS_LD);
  int c = 0;
  while (c + 1 < a.n_branches && (int)blockIdx.x >= a.blk0[c + 1]) ++c;
  const int64_t r0 = (int64_t)(blockIdx.x - a.blk0[c]) * ROWS;
  const int64_t n = a.n[c];
  const int f = a.f[c];
  const int tid = threadIdx.x;
  const int cg = tid & 63, rg = tid >> 6;

  for (int i = tid; i < ROWS * f; i += THREADS) {
    const int r = i / f, k = i - r * f;
    xs[r][k] = (r0 + r < n) ? a.coords[c][(r0 + r) * f + k] : 0.f;
  }
  __syncthreads();

  float acc[8][4];
  // ---- coord_linears[c]: bias scale 1, always followed by the activation (modules.py:312-313)
  {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const float bv = a.first_b[c][cg + 64 * cc];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r][cc] = bv;
    }
    for (int k = 0; k < f; ++k) {
      float u[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) u[cc] = a.first_w[c][(size_t)(cg + 64 * cc) * f + k];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float x = xs[rg * 8 + r][k];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[r][cc] = fmaf(x, u[cc], acc[r][cc]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) hbuf[0][rg * 8 + r][cg + 64 * cc] = acc[r][cc];
  }
  __syncthreads();
  store_rows(a.zsave[c][0], &hbuf[0][0][0], r0, n, tid, true);
  // the activation runs as a ROLLED pass over shared memory: the accurate sinf is ~60 instructions, and 32 inlined
  // copies per layer made the kernel 160 KB of code - 40 % of its stalls were instruction-cache misses
  apply_act(&hbuf[0][0][0], a.act, a.S, tid);
  __syncthreads();
  if (a.n_pre > 0) store_rows(a.asave[c][0], &hbuf[0][0][0], r0, n, tid, false);

  // ---- net[0 .. n_pre-1]: bias / C (modules.py:224-229); the last one without activation when fusion_before_act
  int cur = 0;
  for (int l = 0; l < a.n_pre; ++l) {
    const float* W = a.w[l];
    // chunk loader: 256 rows x 16 k = 1024 float4; thread t takes float4 #(t + 128 i): row = idx / 4, quarter = idx % 4
    float4 stage[8];
    auto gload = [&](int k0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + THREADS * i;
        stage[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(idx >> 2) * WIDTH + k0 + (idx & 3) * 4));
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + THREADS * i;
        float* d = &wsb[buf][(idx >> 2) * WS_LD + (idx & 3) * 4];
        d[0] = stage[i].x; d[1] = stage[i].y; d[2] = stage[i].z; d[3] = stage[i].w;
      }
    };
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) acc[r][cc] = 0.f;
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < WIDTH; k0 += KC) {
      const bool more = k0 + KC < WIDTH;
      if (more) gload(k0 + KC);
#pragma unroll
      for (int k4 = 0; k4 < KC; k4 += 4) {
        float4 h[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) h[r] = *reinterpret_cast<const float4*>(&hbuf[cur][rg * 8 + r][k0 + k4]);   // broadcast
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float u[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) u[cc] = wsb[buf][(cg + 64 * cc) * WS_LD + k4 + kk];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float hv = kk == 0 ? h[r].x : (kk == 1 ? h[r].y : (kk == 2 ? h[r].z : h[r].w));
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[r][cc] = fmaf(hv, u[cc], acc[r][cc]);
          }
        }
      }
      if (more) {
        sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
    const bool last = l == a.n_pre - 1;
    const bool noact = last && a.last_noact;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int j = cg + 64 * cc;
      const float bv = a.b[l][j] * a.inv_c;
#pragma unroll
      for (int r = 0; r < 8; ++r) hbuf[cur ^ 1][rg * 8 + r][j] = acc[r][cc] + bv;
    }
    __syncthreads();
    cur ^= 1;
    store_rows(a.zsave[c][l + 1], &hbuf[cur][0][0], r0, n, tid, true);
    if (!noact) {
      apply_act(&hbuf[cur][0][0], a.act, a.S, tid);
      __syncthreads();
    }
    if (!last) store_rows(a.asave[c][l + 1], &hbuf[cur][0][0], r0, n, tid, false);
    if (last) {                                   // coalesced write of the [16, 256] feature rows
      for (int i = tid; i < ROWS * WIDTH / 4; i += THREADS) {
        const int r = i / (WIDTH / 4);
        if (r0 + r < n)
          reinterpret_cast<float4*>(a.feat[c] + (r0 + r) * WIDTH)[i % (WIDTH / 4)] = reinterpret_cast<const float4*>(&hbuf[cur][r][0])[i % (WIDTH / 4)];
      }
    }
  }
  if (a.n_pre == 0) {                       // approx_layers = 0: the first layer's output is the feature
    for (int i = tid; i < ROWS * WIDTH; i += THREADS) {
      const int r = i / WIDTH, k = i % WIDTH;
      if (r0 + r < n) a.feat[c][(r0 + r) * WIDTH + k] = hbuf[0][r][k];
    }
  }
}

static bool supported(const coordx_net* net) {
  if (net->pre_width != WIDTH || net->fuse_width != WIDTH) return false;
  for (int k = 0; k < net->n_pre; ++k)
    if (net->layer_in[k] != WIDTH || net->layer_out[k] != WIDTH) return false;
  for (int c = 0; c < net->n_branches; ++c)
    if (net->branch_in[c] < 1 || net->branch_in[c] > MAX_F) return false;
  return true;
}

// branches [c_begin, c_end) in one launch
static int launch(const coordx_net* net, int c_begin, int c_end, const float* const* coords, const int64_t* n,
                  float* const* feat, cudaStream_t st, float* const* zsave = nullptr, float* const* asave = nullptr) {
  BranchArgs a{};
  a.n_branches = c_end - c_begin;
  a.n_pre = net->n_pre;
  a.act = net->act;
  a.last_noact = (net->fusion_before_act || net->n_pre == net->n_layers) ? 1 : 0;
  a.S = 30.f * net->act_scale;
  a.inv_c = 1.f / net->n_branches;
  int blocks = 0;
  for (int c = c_begin; c < c_end; ++c) {
    const int i = c - c_begin;
    a.coords[i] = coords[i];
    a.feat[i] = feat[i];
    a.n[i] = n[i];
    a.f[i] = net->branch_in[c];
    a.first_w[i] = net->first_w[c];
    a.first_b[i] = net->first_b[c];
    a.blk0[i] = blocks;
    blocks += (int)cdiv(n[i], ROWS);
    if (zsave && c_end - c_begin == 1)                 // training mode: one branch per call
      for (int l = 0; l <= net->n_pre; ++l) { a.zsave[i][l] = zsave[l]; a.asave[i][l] = asave[l]; }
  }
  a.blk0[a.n_branches] = blocks;
  for (int k = 0; k < net->n_pre; ++k) { a.w[k] = net->layer_w[k]; a.b[k] = net->layer_b[k]; }
  if (blocks == 0) return 0;
  static AttrOnce once;
  CX_TRY(once.set(branch_fused_fwd_kernel, (int)SMEM_BYTES));
  branch_fused_fwd_kernel<<<blocks, THREADS, SMEM_BYTES, st>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

}  // namespace bf
}  // namespace coordx
