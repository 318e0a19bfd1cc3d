// C ABI of coordx_b200 (see include/coordx_b200.h).  Host-side orchestration of the
// CUDA kernels; no torch types, no allocation, no synchronisation.
#include <stdlib.h>

#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_f32.cuh"
#include "branch_fused.cuh"
#include "sum_sep.cuh"
#include "interp.cuh"
#include "tc_common.cuh"

namespace coordx {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

struct Layer {
  const float* W;
  const float* b;
  float bscale;
  int in, out;
  int act;          // COORDX_ACT_* or -1 (identity)
  float* dW;
  float* db;
};

// Y_l = act_l(X_l W_l^T + b_l * bscale_l), chained.  If Zs/As are given, every layer's
// pre-activation / output is kept (As[nl-1] == Yfinal unless Yfinal is null).
static int mlp_forward(const Layer* L, int nl, const float* X, int64_t ldx, int M, float* Yfinal, float* const* Zs,
                       float* const* As, float* tmp0, float* tmp1, float S, cudaStream_t st) {
  const float* cur = X;
  int64_t ld = ldx;
  for (int l = 0; l < nl; ++l) {
    float* y = (l == nl - 1 && Yfinal) ? Yfinal : (As ? As[l] : ((l & 1) ? tmp1 : tmp0));
    float* z = Zs ? Zs[l] : nullptr;
    CX_TRY(linear_forward(cur, ld, L[l].W, L[l].b, L[l].bscale, M, L[l].out, L[l].in, y, z, L[l].out, L[l].act, S, st));
    cur = y;
    ld = L[l].out;
  }
  return 0;
}

// dX[M,K] = dZ[M,N] W[N,K] (* act'(Zprev)) on the tensor cores (3xTF32) when the shape pays: W is transposed into `wt`
// first so that both operands are K-major.  Returns 1 when it launched, 0 when the caller should take the FFMA kernel.
static int dx_tensor(const float* dZ, int64_t lddz, const float* W, int M, int N, int K, float* dX, int64_t lddx,
                     const float* Zprev, int64_t ldz, int act, float S, float* wt, cudaStream_t st, int* launched) {
  *launched = 0;
  if (!wt) return 0;
  g3::Args t{};
  t.A = dZ; t.lda = lddz; t.B = wt; t.ldb = N; t.M = M; t.N = K; t.K = N;
  t.Y = dX; t.ldc = lddx; t.cs = 1; t.act = Zprev ? act : -1; t.S = S; t.zprev = Zprev; t.ldz = ldz;
  if (!g3::usable(t)) return 0;
  dim3 grid((unsigned)cdiv(K, 32), (unsigned)cdiv(N, 32));
  transpose_kernel<<<grid, 256, 0, st>>>(W, N, K, wt);
  CX_LAUNCH_CHECK();
  CX_TRY(g3::launch(t, st));
  *launched = 1;
  return 0;
}

// Side stream of a backward chain: the weight- and bias-gradient kernels of a layer only CONSUME dz_l, the chain
// itself continues with dX.  On one stream a 3-layer branch backward is 3 x (dW, colsum, dX) small launch-latency-bound
// GEMMs in a row (250 us per 2048-sample axis, measured inside the captured training step); forked, the critical path
// is the dX chain alone.  Plain event fork/join: legal inside a CUDA-graph capture.
struct Fork {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main[COORDX_MAX_LAYERS + 2] = {};
  cudaEvent_t ev_side[COORDX_MAX_LAYERS + 2] = {};
};

// (one set per device and slot, created on first use - i.e. during the warm-up iterations that precede a graph
//  capture; calls are expected from one host thread per device, like every other entry point of this library)
static int get_fork(int slot, Fork** out) {
  constexpr int MAX_DEV = 16;
  static Fork forks[MAX_DEV][COORDX_MAX_BRANCHES];
  int dev = 0;
  CX_CUDA(cudaGetDevice(&dev));
  Fork& f = forks[dev % MAX_DEV][slot % COORDX_MAX_BRANCHES];
  if (!f.side) {
    CX_CUDA(cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking));
    for (int i = 0; i < COORDX_MAX_LAYERS + 2; ++i) {
      CX_CUDA(cudaEventCreateWithFlags(&f.ev_main[i], cudaEventDisableTiming));
      CX_CUDA(cudaEventCreateWithFlags(&f.ev_side[i], cudaEventDisableTiming));
    }
  }
  *out = &f;
  return 0;
}

// Backward of mlp_forward.  dY = gradient w.r.t. the chain output.  d0/d1: scratch of
// M * max width floats.  dX (nullable) [M, L[0].in].  fork (nullable): run dW / db on its side stream.
static int mlp_backward(const Layer* L, int nl, const float* X, int64_t ldx, int M, float* const* Zs, float* const* As,
                        const float* dY, float* d0, float* d1, float* dX, float S, cudaStream_t st, Fork* fork = nullptr,
                        float* wt = nullptr, float* tw = nullptr, size_t tw_floats = 0) {
  const float* dz = dY;
  float* bufs[2] = {d0, d1};
  int which = 0;
  if (L[nl - 1].act >= 0) {
    const int64_t cnt = (int64_t)M * L[nl - 1].out;
    mul_act_grad_kernel<<<(unsigned)cdiv(cnt, 256), 256, 0, st>>>(dY, Zs[nl - 1], bufs[which], cnt, L[nl - 1].act, S);
    CX_LAUNCH_CHECK();
    dz = bufs[which];
    which ^= 1;
  }
  int side_pending = -1;            // index of the last ev_side recorded and not yet waited for by `st`
  for (int l = nl - 1; l >= 0; --l) {
    const float* in = l == 0 ? X : As[l - 1];
    const int64_t ldin = l == 0 ? ldx : L[l - 1].out;
    cudaStream_t gs = st;
    if (fork && (L[l].dW || L[l].db)) {
      CX_CUDA(cudaEventRecord(fork->ev_main[l], st));                  // dz_l is ready (and the grads are zero-filled)
      CX_CUDA(cudaStreamWaitEvent(fork->side, fork->ev_main[l], 0));
      gs = fork->side;
    }
    if (L[l].dW) {
      int on_tc = 0;
      CX_TRY(dw_tensor(dz, L[l].out, in, ldin, M, L[l].out, L[l].in, L[l].dW, tw, tw_floats, gs, &on_tc));
      if (!on_tc) CX_TRY(linear_backward_dw(dz, L[l].out, in, ldin, M, L[l].out, L[l].in, L[l].dW, gs));
    }
    if (L[l].db) CX_TRY(bias_backward(dz, L[l].out, M, L[l].out, L[l].bscale, L[l].db, gs));
    // the dX below overwrites the buffer that held dz_{l+1}: the side stream must be done reading it
    if (side_pending >= 0) {
      CX_CUDA(cudaStreamWaitEvent(st, fork->ev_side[side_pending], 0));
      side_pending = -1;
    }
    if (gs != st) {
      CX_CUDA(cudaEventRecord(fork->ev_side[l], gs));
      side_pending = l;
    }
    if (l > 0) {
      const bool has_act = L[l - 1].act >= 0;
      int on_tc = 0;
      CX_TRY(dx_tensor(dz, L[l].out, L[l].W, M, L[l].out, L[l].in, bufs[which], L[l].in, has_act ? Zs[l - 1] : nullptr,
                       L[l - 1].out, L[l - 1].act, S, wt, st, &on_tc));
      if (!on_tc)
        CX_TRY(linear_backward_dx(dz, L[l].out, L[l].W, M, L[l].out, L[l].in, bufs[which], L[l].in,
                                  has_act ? Zs[l - 1] : nullptr, L[l - 1].out, L[l - 1].act, S, st));
      dz = bufs[which];
      which ^= 1;
    } else if (dX) {
      int on_tc = 0;                                // (the input gradient of the chain: the fused features of the fp32 trunk)
      CX_TRY(dx_tensor(dz, L[l].out, L[l].W, M, L[l].out, L[l].in, dX, L[l].in, nullptr, 0, -1, S, wt, st, &on_tc));
      if (!on_tc) CX_TRY(linear_backward_dx(dz, L[l].out, L[l].W, M, L[l].out, L[l].in, dX, L[l].in, nullptr, 0, -1, S, st));
    }
  }
  if (side_pending >= 0) CX_CUDA(cudaStreamWaitEvent(st, fork->ev_side[side_pending], 0));     // join
  return 0;
}

static int check_net(const coordx_net* net) {
  if (!net) return set_err(COORDX_ERR_ARG, "net is null%s");
  if (net->n_branches < 1 || net->n_branches > COORDX_MAX_BRANCHES) return set_err(COORDX_ERR_ARG, "n_branches out of range%s (%lld)", "", net->n_branches);
  if (net->n_layers < 1 || net->n_layers > COORDX_MAX_LAYERS || net->n_pre < 0 || net->n_pre > net->n_layers)
    return set_err(COORDX_ERR_ARG, "bad layer counts%s (n_pre %lld, n_layers %lld)", "", net->n_pre, net->n_layers);
  if (net->fusion_op != COORDX_FUSE_SUM && net->fusion_op != COORDX_FUSE_PROD)
    return set_err(COORDX_ERR_ARG, "unknown fusion operator%s (%lld)", "", net->fusion_op);   // the reference silently skips the fusion (modules.py:336-341)
  if (net->act < COORDX_ACT_SINE || net->act > COORDX_ACT_ELU) return set_err(COORDX_ERR_ARG, "unknown activation%s (%lld)", "", net->act);
  if (net->hidden <= 0 || net->fuse_width % net->hidden != 0) return set_err(COORDX_ERR_ARG, "fuse_width must be a multiple of hidden%s");
  return 0;
}

// layers of branch `b`: coord_linears[b] then net[0..n_pre-1]
static int branch_layers(const coordx_net* net, int b, const coordx_grads* g, Layer* L) {
  int nl = 0;
  L[nl++] = Layer{net->first_w[b], net->first_b[b], 1.f, net->branch_in[b], net->pre_width, net->act,
                  g ? g->first_w[b] : nullptr, g ? g->first_b[b] : nullptr};
  for (int k = 0; k < net->n_pre; ++k) {
    const bool last = (k == net->n_pre - 1);
    const bool noact = last && (net->fusion_before_act || k == net->n_layers - 1);
    L[nl++] = Layer{net->layer_w[k], net->layer_b[k], 1.f / net->n_branches, net->layer_in[k], net->layer_out[k],
                    noact ? -1 : net->act, g ? g->layer_w[k] : nullptr, g ? g->layer_b[k] : nullptr};
  }
  return nl;
}

static int trunk_layers(const coordx_net* net, const coordx_grads* g, Layer* L) {
  int nl = 0;
  for (int k = net->n_pre; k < net->n_layers; ++k)
    L[nl++] = Layer{net->layer_w[k], net->layer_b[k], 1.f, net->layer_in[k], net->layer_out[k],
                    k == net->n_layers - 1 ? -1 : net->act, g ? g->layer_w[k] : nullptr, g ? g->layer_b[k] : nullptr};
  return nl;
}

struct ChainWs {
  float* Z[COORDX_MAX_LAYERS + 1];
  float* A[COORDX_MAX_LAYERS + 1];
  float* d0;
  float* d1;
  float* tmp0;
  float* tmp1;
  float* wt;        // backward: one transposed weight matrix (the dX GEMM as X . W^T^T on the tensor cores)
  float* tw;        // backward: transposed dZ / input chunks of the weight-gradient GEMM (dw_tensor)
  size_t tw_floats;
};

static void carve_chain(Arena& ar, const Layer* L, int nl, int64_t M, bool backward, ChainWs& w) {
  int maxw = 1;
  for (int l = 0; l < nl; ++l) maxw = max(maxw, max(L[l].in, L[l].out));
  if (backward) {
    for (int l = 0; l < nl; ++l) {
      w.Z[l] = ar.get<float>((size_t)M * L[l].out);
      w.A[l] = ar.get<float>((size_t)M * L[l].out);
    }
    w.d0 = ar.get<float>((size_t)M * maxw);
    w.d1 = ar.get<float>((size_t)M * maxw);
    w.wt = ar.get<float>((size_t)maxw * maxw);
    w.tw_floats = M >= 128 && maxw >= 64 ? (size_t)2 * maxw * ((min(M, (int64_t)dw_rows(2 * maxw)) + 31) / 32 * 32) : 0;
    w.tw = w.tw_floats ? ar.get<float>(w.tw_floats) : nullptr;
    w.tmp0 = w.tmp1 = nullptr;
  } else {
    w.tmp0 = ar.get<float>((size_t)M * maxw);
    w.tmp1 = ar.get<float>((size_t)M * maxw);
    w.wt = nullptr;
    w.tw = nullptr;
    w.tw_floats = 0;
  }
}

// ------------------------------------------------------------------ trunk planning (fp32 path)
// The [P, Hd] fused tensor is built chunk by chunk inside the workspace.  A chunk is a box of the grid that is
// CONTIGUOUS in the row-major output: a single index on every axis before `axis`, `step` consecutive samples of
// `axis`, and all of the axes after it - whole axis-0 rows for image-sized grids (axis = 0), pieces of one row when
// a single row is already longer than the target (a 1 x P scattered "grid", a 3-D grid with a large n1 * n2).
struct TrunkPlan {
  int axis;             // the axis that is cut into pieces
  int64_t step;         // samples of `axis` per chunk
  int64_t chunk_pts;    // points of a full chunk: the workspace is sized for this many
};

static TrunkPlan plan_trunk(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1) {
  const int C = net->n_branches;
  const int64_t target = 32768;      // points per chunk: [chunk, 256] fp32 = 32 MB, stays L2-resident
  TrunkPlan pl{};
  int64_t below = 1;                 // prod of the extents of the axes after `axis`
  pl.axis = C - 1;
  for (int c = C - 1; c >= 0; --c) {
    const int64_t ext = c == 0 ? r1 - r0 : n[c];
    pl.axis = c;
    if (below * ext > target || c == 0) break;
    below *= ext;
  }
  const int64_t ext_a = pl.axis == 0 ? r1 - r0 : n[pl.axis];
  pl.step = max((int64_t)1, min(ext_a, target / max(below, (int64_t)1)));
  pl.chunk_pts = pl.step * below;
  return pl;
}

// one chunk of the plan as a sub-grid of its own: offset table pointers, per-axis counts, position in the output
struct TrunkChunk {
  const float* feats[COORDX_MAX_BRANCHES];
  float* dfeat[COORDX_MAX_BRANCHES];
  int64_t n[COORDX_MAX_BRANCHES];
  int64_t first_pt;     // index of the chunk's first point inside the slab's output
  int64_t pts;
};

// enumerate the chunks in output order; fn(const TrunkChunk&) returns 0 or an error code
template <typename Fn>
static int for_each_chunk(const coordx_net* net, const TrunkPlan& pl, const float* const* feats, float* const* dfeat,
                          const int64_t* n, int64_t r0, int64_t r1, Fn fn) {
  const int C = net->n_branches, F = net->fuse_width;
  int64_t idx[COORDX_MAX_BRANCHES] = {};          // current index on the axes before pl.axis (axis 0 relative to r0)
  auto ext = [&](int c) { return c == 0 ? r1 - r0 : n[c]; };
  for (;;) {
    for (int64_t i0 = 0; i0 < ext(pl.axis); i0 += pl.step) {
      TrunkChunk ch{};
      int64_t lin = 0;
      ch.pts = 1;
      for (int c = 0; c < C; ++c) {
        const int64_t start = c < pl.axis ? idx[c] : (c == pl.axis ? i0 : 0);
        const int64_t cnt = c < pl.axis ? 1 : (c == pl.axis ? min(pl.step, ext(c) - i0) : ext(c));
        const int64_t abs0 = start + (c == 0 ? r0 : 0);
        ch.feats[c] = feats[c] + abs0 * F;
        ch.dfeat[c] = dfeat ? dfeat[c] + abs0 * F : nullptr;
        ch.n[c] = cnt;
        ch.pts *= cnt;
        lin = lin * ext(c) + start;
      }
      ch.first_pt = lin;
      CX_TRY(fn(ch));
    }
    int c = pl.axis - 1;                           // odometer over the axes before pl.axis
    for (; c >= 0; --c) {
      if (++idx[c] < ext(c)) break;
      idx[c] = 0;
    }
    if (c < 0) break;
  }
  return 0;
}

static FuseArgs make_fuse_args(const coordx_net* net, const TrunkChunk& ch) {
  FuseArgs a{};
  a.C = net->n_branches;
  a.inner = 1;
  for (int c = 0; c < a.C; ++c) {
    a.feat[c] = ch.feats[c];
    a.n[c] = ch.n[c];
    if (c > 0) a.inner *= ch.n[c];
  }
  a.F = net->fuse_width;
  a.Hd = net->hidden;
  a.gs = net->fuse_width / net->hidden;
  a.op = net->fusion_op;
  a.fba = net->fusion_before_act;
  a.act = net->act;
  a.S = 30.f * net->act_scale;
  a.code = net->has_code ? net->code : nullptr;
  a.row0 = 0;
  a.nrows = ch.n[0];
  return a;
}

}  // namespace coordx

using namespace coordx;

extern "C" {

int coordx_abi_version(void) { return COORDX_ABI_VERSION; }
const char* coordx_last_error(void) { return g_err; }

unsigned long long coordx_launch_count(void) { return g_launches; }

int coordx_profile_enable(int on) {
  tc::profile_enable(on != 0);
  return 0;
}

// device buffer of 4 x 96 x 4 int64 that CTA 0 of the forward trunk kernel fills with SM-clock stamps of its
// pipeline events (roles: MMA issuer, epilogue 0/1, producer warp 8); nullptr switches it off
int coordx_profile_set_trace(void* dev_buf) {
  return tc::profile_set_trace(dev_buf);
}
// same for the backward trunk kernels; pass: 0 = dW pass, 1 = dH pass (outer axes), 2 = dH pass (last axis)
int coordx_profile_set_bwd_trace(void* dev_buf, int pass) {
  return tc::profile_set_bwd_trace(dev_buf, pass);
}

// milliseconds of the last profiled trunk kernel (synchronises on its end event)
int coordx_profile_last_ms(float* ms_host) {
  return tc::profile_last_ms(ms_host);
}

int coordx_trunk_tc_supported(const coordx_net* net, int precision) {
  if (check_net(net) != 0) return 0;
  return tc::tc_supported(net, precision) ? 1 : 0;
}

// -------------------------------------------------------------------------- branch
// inference over many rows (the scattered-point path: one "axis sample" per query point): rows are pushed through the
// per-layer GEMM chain in chunks of this many, so the workspace is bounded and the 3xTF32 tensor-core GEMM takes the
// 256-wide layers (the fused FFMA kernel below is built for a few thousand axis samples)
constexpr int64_t BRANCH_CHUNK = 65536;
constexpr int64_t BRANCH_FUSED_MAX = 255;     // from 256 rows on, the 128 x 64-tile 3xTF32 GEMM per layer beats the fused FFMA kernel

size_t coordx_branch_workspace_bytes(const coordx_net* net, int branch, int64_t n, int backward) {
  if (check_net(net) != 0 || branch < 0 || branch >= net->n_branches) return 0;
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = branch_layers(net, branch, nullptr, L);
  Arena ar(nullptr, 0);
  ChainWs w{};
  carve_chain(ar, L, nl, backward ? n : min(n, BRANCH_CHUNK), backward != 0, w);
  return ar.off + 256;
}

int coordx_branch_forward(const coordx_net* net, int branch, const float* coords, int64_t n, float* feat, void* ws,
                          size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (branch < 0 || branch >= net->n_branches) return set_err(COORDX_ERR_ARG, "branch out of range%s (%lld)", "", branch);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (bf::supported(net) && n <= BRANCH_FUSED_MAX) return bf::launch(net, branch, branch + 1, &coords, &n, &feat, st);
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = branch_layers(net, branch, nullptr, L);
  Arena ar(ws, ws_bytes);
  ChainWs w{};
  carve_chain(ar, L, nl, min(n, BRANCH_CHUNK), false, w);
  if (!ar.ok() || (nl > 1 && !w.tmp0)) return set_err(COORDX_ERR_WORKSPACE, "branch forward: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const int f = net->branch_in[branch];
  for (int64_t r0 = 0; r0 < n; r0 += BRANCH_CHUNK) {
    const int rows = (int)min(BRANCH_CHUNK, n - r0);
    CX_TRY(mlp_forward(L, nl, coords + r0 * f, f, rows, feat + r0 * L[nl - 1].out, nullptr, nullptr, w.tmp0, w.tmp1,
                       30.f * net->act_scale, st));
  }
  return 0;
}

int coordx_branch_backward(const coordx_net* net, int branch, const float* coords, int64_t n, const float* dfeat,
                           coordx_grads* grads, float* dcoords, void* ws, size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (branch < 0 || branch >= net->n_branches) return set_err(COORDX_ERR_ARG, "branch out of range%s (%lld)", "", branch);
  if (!grads) return set_err(COORDX_ERR_ARG, "grads is null%s");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = branch_layers(net, branch, grads, L);
  Arena ar(ws, ws_bytes);
  ChainWs w{};
  carve_chain(ar, L, nl, n, true, w);
  if (!ar.ok() || !ws) return set_err(COORDX_ERR_WORKSPACE, "branch backward: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const float S = 30.f * net->act_scale;
  CX_TRY(mlp_forward(L, nl, coords, net->branch_in[branch], (int)n, nullptr, w.Z, w.A, nullptr, nullptr, S, st));
  Fork* fork = nullptr;
  CX_TRY(get_fork(branch, &fork));
  return mlp_backward(L, nl, coords, net->branch_in[branch], (int)n, w.Z, w.A, dfeat, w.d0, w.d1, dcoords, S, st, fork, w.wt, w.tw, w.tw_floats);
}

// Training-mode pair: the forward keeps every layer's pre-activation / activation in `save` (the backward
// workspace layout of carve_chain), the backward starts from them instead of recomputing the branch.
int coordx_branch_forward_save(const coordx_net* net, int branch, const float* coords, int64_t n, float* feat, void* save,
                               size_t save_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (branch < 0 || branch >= net->n_branches) return set_err(COORDX_ERR_ARG, "branch out of range%s (%lld)", "", branch);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = branch_layers(net, branch, nullptr, L);
  Arena ar(save, save_bytes);
  ChainWs w{};
  carve_chain(ar, L, nl, n, true, w);
  if (!ar.ok() || !save) return set_err(COORDX_ERR_WORKSPACE, "branch forward (save): buffer too small%s (%lld < %lld)", "", (long long)save_bytes, (long long)ar.off);
  if (bf::supported(net) && n <= BRANCH_FUSED_MAX) return bf::launch(net, branch, branch + 1, &coords, &n, &feat, st, w.Z, w.A);
  return mlp_forward(L, nl, coords, net->branch_in[branch], (int)n, feat, w.Z, w.A, nullptr, nullptr, 30.f * net->act_scale, st);
}

int coordx_branch_backward_saved(const coordx_net* net, int branch, const float* coords, int64_t n, const float* dfeat,
                                 coordx_grads* grads, float* dcoords, void* save, size_t save_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (branch < 0 || branch >= net->n_branches) return set_err(COORDX_ERR_ARG, "branch out of range%s (%lld)", "", branch);
  if (!grads) return set_err(COORDX_ERR_ARG, "grads is null%s");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = branch_layers(net, branch, grads, L);
  Arena ar(save, save_bytes);
  ChainWs w{};
  carve_chain(ar, L, nl, n, true, w);
  if (!ar.ok() || !save) return set_err(COORDX_ERR_WORKSPACE, "branch backward (saved): buffer too small%s (%lld < %lld)", "", (long long)save_bytes, (long long)ar.off);
  Fork* fork = nullptr;
  CX_TRY(get_fork(branch, &fork));
  return mlp_backward(L, nl, coords, net->branch_in[branch], (int)n, w.Z, w.A, dfeat, w.d0, w.d1, dcoords,
                      30.f * net->act_scale, st, fork, w.wt, w.tw, w.tw_floats);
}

// -------------------------------------------------------------------------- trunk
size_t coordx_trunk_workspace_bytes(const coordx_net* net, const int64_t* n, int64_t row_begin, int64_t row_end,
                                    int precision, int backward) {
  if (check_net(net) != 0 || !n) return 0;
  if (sep::enabled() && sep::supported(net) && (!backward || sep::backward_fits(net, n, row_begin, row_end)))
    return sep::workspace_bytes(net, n, row_begin, row_end, backward != 0);
  if (precision != COORDX_PREC_FP32) {
    if (!tc::tc_supported(net, precision)) return 0;
    if (backward) return tc::bwd_workspace_bytes(net, n, row_begin, row_end, backward == 2) + 1024;
    return tc::fwd_workspace_bytes(n[net->n_branches - 1]) + 1024;
  }
  const TrunkPlan pl = plan_trunk(net, n, row_begin, row_end);
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = trunk_layers(net, nullptr, L);
  Arena ar(nullptr, 0);
  ar.get<float>((size_t)pl.chunk_pts * net->hidden);     // H0
  if (backward) ar.get<float>((size_t)pl.chunk_pts * net->hidden);   // dH0
  ChainWs w{};
  if (nl > 0) carve_chain(ar, L, nl, pl.chunk_pts, backward != 0, w);
  return ar.off + 256;
}

int coordx_trunk_forward(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                         int64_t row_end, float* out, int precision, void* ws, size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (!feats || !n) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (row_begin < 0 || row_end > n[0] || row_begin > row_end) return set_err(COORDX_ERR_ARG, "bad slab%s [%lld, %lld)", "", row_begin, row_end);
  {                                                        // empty grid (an axis without samples): nothing to do
    int64_t pts = row_end - row_begin;
    for (int c = 1; c < net->n_branches; ++c) pts *= n[c];
    if (pts == 0) return 0;
  }
  if (!out) return set_err(COORDX_ERR_ARG, "null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  // sum-fused sine trunk: separable evaluation over the per-axis tables, exact to fp32 rounding, any precision
  if (sep::enabled() && sep::supported(net)) return sep::forward(net, feats, n, row_begin, row_end, out, ws, ws_bytes, st);
  if (precision != COORDX_PREC_FP32) {
    if (!tc::tc_supported(net, precision)) return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk does not cover this architecture%s");
    return tc::trunk_forward_tc(net, feats, n, row_begin, row_end, out, precision, ws, ws_bytes, st);
  }
  const TrunkPlan pl = plan_trunk(net, n, row_begin, row_end);
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = trunk_layers(net, nullptr, L);
  Arena ar(ws, ws_bytes);
  float* H0 = ar.get<float>((size_t)pl.chunk_pts * net->hidden);
  ChainWs w{};
  if (nl > 0) carve_chain(ar, L, nl, pl.chunk_pts, false, w);
  if (!ws || !ar.ok()) return set_err(COORDX_ERR_WORKSPACE, "trunk forward: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const float S = 30.f * net->act_scale;
  const int ow = nl > 0 ? L[nl - 1].out : net->hidden;
  return for_each_chunk(net, pl, feats, nullptr, n, row_begin, row_end, [&](const TrunkChunk& ch) -> int {
    const FuseArgs fa = make_fuse_args(net, ch);
    const int64_t pts = ch.pts;
    float* o = out + ch.first_pt * ow;
    float* h0 = nl > 0 ? H0 : o;            // approx_layers == num_hidden_layers+1: the fused value is the output
    if (fuse_fwd_vec4_ok(fa, h0))
      fuse_fwd_vec4_kernel<<<(unsigned)cdiv(pts * (net->hidden / 4), 256), 256, 0, st>>>(fa, pts, h0);
    else
      fuse_fwd_kernel<<<(unsigned)cdiv(pts * net->hidden, 256), 256, 0, st>>>(fa, pts, h0);
    CX_LAUNCH_CHECK();
    if (nl > 0) CX_TRY(mlp_forward(L, nl, H0, net->hidden, (int)pts, o, nullptr, nullptr, w.tmp0, w.tmp1, S, st));
    return 0;
  });
}

int coordx_trunk_backward(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                          int64_t row_end, const float* dout, coordx_grads* grads, float* const* dfeat, int precision,
                          void* ws, size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (!feats || !n || !grads || !dfeat) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (row_begin < 0 || row_end > n[0] || row_begin > row_end) return set_err(COORDX_ERR_ARG, "bad slab%s [%lld, %lld)", "", row_begin, row_end);
  {
    int64_t pts = row_end - row_begin;
    for (int c = 1; c < net->n_branches; ++c) pts *= n[c];
    if (pts == 0) return 0;
  }
  if (!dout) return set_err(COORDX_ERR_ARG, "null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  if (sep::enabled() && sep::supported(net) && sep::backward_fits(net, n, row_begin, row_end))
    return sep::backward(net, feats, n, row_begin, row_end, dout, grads, dfeat, ws, ws_bytes, st);
  if (precision != COORDX_PREC_FP32) {
    if (!tc::tc_supported(net, precision)) return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk does not cover this architecture%s");
    const int Lf = net->n_layers - 1;
    int64_t pts = row_end - row_begin;
    for (int c = 1; c < net->n_branches; ++c) pts *= n[c];
    // gradient of the final bias: column sums of dout
    CX_TRY(bias_backward(dout, net->layer_out[Lf], (int)pts, net->layer_out[Lf], 1.f, grads->layer_b[Lf], st));
    return tc::trunk_backward_tc(net, feats, n, row_begin, row_end, dout, grads, dfeat, precision, ws, ws_bytes, st);
  }
  const TrunkPlan pl = plan_trunk(net, n, row_begin, row_end);
  Layer L[COORDX_MAX_LAYERS + 1];
  const int nl = trunk_layers(net, grads, L);
  Arena ar(ws, ws_bytes);
  float* H0 = ar.get<float>((size_t)pl.chunk_pts * net->hidden);
  float* dH0 = ar.get<float>((size_t)pl.chunk_pts * net->hidden);
  ChainWs w{};
  if (nl > 0) carve_chain(ar, L, nl, pl.chunk_pts, true, w);
  if (!ws || !ar.ok()) return set_err(COORDX_ERR_WORKSPACE, "trunk backward: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const float S = 30.f * net->act_scale;
  const int ow = nl > 0 ? L[nl - 1].out : net->hidden;
  const int C = net->n_branches;
  return for_each_chunk(net, pl, feats, dfeat, n, row_begin, row_end, [&](const TrunkChunk& ch) -> int {
    const FuseArgs fa = make_fuse_args(net, ch);
    const int64_t pts = ch.pts;
    const float* dY = dout + ch.first_pt * ow;
    const float* dh = dY;
    if (nl > 0) {
      if (fuse_fwd_vec4_ok(fa, H0))
        fuse_fwd_vec4_kernel<<<(unsigned)cdiv(pts * (net->hidden / 4), 256), 256, 0, st>>>(fa, pts, H0);
      else
        fuse_fwd_kernel<<<(unsigned)cdiv(pts * net->hidden, 256), 256, 0, st>>>(fa, pts, H0);
      CX_LAUNCH_CHECK();
      CX_TRY(mlp_forward(L, nl, H0, net->hidden, (int)pts, nullptr, w.Z, w.A, nullptr, nullptr, S, st));
      CX_TRY(mlp_backward(L, nl, H0, net->hidden, (int)pts, w.Z, w.A, dY, w.d0, w.d1, dH0, S, st, nullptr, w.wt, w.tw, w.tw_floats));
      dh = dH0;
    }
    for (int b = 0; b < C; ++b) {
      const int64_t cnt_b = ch.n[b];
      const int64_t T = pts / cnt_b;
      // spread the reduction over the chip: ~4 waves of 128-thread blocks
      const int64_t base_blocks = cdiv(net->fuse_width, 128) * cnt_b;
      int64_t splits = max((int64_t)1, min(cdiv(T, 64), cdiv((int64_t)148 * 16, base_blocks)));
      const int64_t tpb = cdiv(T, splits);
      splits = cdiv(T, tpb);
      if (splits > 65535) return set_err(COORDX_ERR_UNSUPPORTED, "axis too long for fuse_bwd grid%s (%lld)", "", cnt_b);
      // (gridDim.x carries the samples of axis b: up to 2^31 - 1)
      dim3 grid((unsigned)cnt_b, (unsigned)cdiv(net->fuse_width, 128), (unsigned)splits);
      fuse_bwd_kernel<<<grid, 128, 0, st>>>(fa, b, dh, ch.dfeat[b], (b == 0 && net->has_code) ? grads->code : nullptr, T, tpb);
      CX_LAUNCH_CHECK();
    }
    return 0;
  });
}

// -------------------------------------------------------------------------- loss / optimiser
int coordx_loss_forward_backward(int kind, const float* out, const float* gt, int64_t numel, float inv_count,
                                 float* loss_accum, float* dout, void* stream) {
  if (kind != COORDX_LOSS_MSE && kind != COORDX_LOSS_BCE) return set_err(COORDX_ERR_ARG, "unknown loss%s (%lld)", "", kind);
  if (!out || !gt || !loss_accum) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (numel <= 0) return 0;
  const int blocks = (int)min((int64_t)148 * 8, cdiv(numel, 256));
  loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kind, out, gt, numel, inv_count, loss_accum, dout);
  CX_LAUNCH_CHECK();
  return 0;
}

// coordx_trunk_forward with the loss of training.py:133-135 (loss_functions.py:8-12 image_mse, :56-60 occupancy_3d) fused
// into the last layer: the tcgen05 trunk evaluates it in its epilogue (the thread that completes out[p, :] reads
// gt[p, :], writes d loss / d out and keeps a running sum: no pass over [P, O] besides the kernel's own store);
// every other path (separable sum evaluation, FP32) runs its forward followed by the loss kernel - same results.
int coordx_trunk_forward_loss(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                              int64_t row_end, float* out, int loss_kind, const float* gt, float inv_count,
                              float* loss_accum, float* dout, int precision, void* ws, size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (loss_kind != COORDX_LOSS_MSE && loss_kind != COORDX_LOSS_BCE) return set_err(COORDX_ERR_ARG, "unknown loss%s (%lld)", "", loss_kind);
  if (!feats || !n || !gt || !loss_accum || !dout || !out) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (row_begin < 0 || row_end > n[0] || row_begin > row_end) return set_err(COORDX_ERR_ARG, "bad slab%s [%lld, %lld)", "", row_begin, row_end);
  int64_t pts = row_end - row_begin;
  for (int c = 1; c < net->n_branches; ++c) pts *= n[c];
  if (pts == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_path = !(sep::enabled() && sep::supported(net)) && precision != COORDX_PREC_FP32;
  if (tc_path) {
    if (!tc::tc_supported(net, precision)) return set_err(COORDX_ERR_UNSUPPORTED, "tcgen05 trunk does not cover this architecture%s");
    const tc::FwdLoss fl{loss_kind, gt, inv_count, loss_accum, dout};
    return tc::trunk_forward_tc(net, feats, n, row_begin, row_end, out, precision, ws, ws_bytes, st, &fl);
  }
  CX_TRY(coordx_trunk_forward(net, feats, n, row_begin, row_end, out, precision, ws, ws_bytes, stream));
  const int ow = net->layer_out[net->n_layers - 1];
  return coordx_loss_forward_backward(loss_kind, out, gt, pts * ow, inv_count, loss_accum, dout, stream);
}

// ground-truth gather of a training batch (dataio.py:575-600), device-resident data; see gather_rows_kernel
int coordx_gather_rows(const float* data, int width, const int64_t* idz, int64_t nz, int64_t stride_z, const int64_t* idx,
                       int64_t m, float* out, void* stream) {
  if (!data || !idx || !out || width < 1 || nz < 0 || m < 0 || (idz && stride_z < 0)) return set_err(COORDX_ERR_ARG, "bad gather argument%s");
  if (!idz) nz = 1;
  const int64_t total = nz * m * width;
  if (total == 0) return 0;
  const int blocks = (int)min((int64_t)148 * 16, cdiv(total, 256));
  gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, width, idz, nz, stride_z, idx, m, out);
  CX_LAUNCH_CHECK();
  return 0;
}

int coordx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                     float beta1, float beta2, float eps, int64_t step_host, const float* grad_scale_dev, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step_host < 1) return set_err(COORDX_ERR_ARG, "bad adam argument%s");
  if (numel <= 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step_host);
  const double bc2 = 1.0 - pow((double)beta2, (double)step_host);
  adam_kernel<<<(unsigned)cdiv(numel, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, numel, (float)(lr / bc1),
                                                                          beta1, beta2, eps, (float)(1.0 / sqrt(bc2)), grad_scale_dev);
  CX_LAUNCH_CHECK();
  return 0;
}

int coordx_optim_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, const int64_t* numel, float* state, float beta1, float beta2, float eps,
                      float max_grad_norm, float lr_gamma, void* stream) {
  if (n_tensors < 0 || n_tensors > COORDX_OPT_MAX_TENSORS) return set_err(COORDX_ERR_ARG, "optim_step: too many tensors%s (%lld)", "", n_tensors);
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !state) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (n_tensors == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  OptimTensors t{};
  t.n = n_tensors;
  int64_t biggest = 1;
  for (int k = 0; k < n_tensors; ++k) {
    if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || numel[k] < 0) return set_err(COORDX_ERR_ARG, "optim_step: bad tensor%s (%lld)", "", k);
    t.p[k] = params[k]; t.g[k] = grads[k]; t.m[k] = exp_avg[k]; t.v[k] = exp_avg_sq[k]; t.numel[k] = numel[k];
    biggest = max(biggest, numel[k]);
  }
  dim3 grid((unsigned)min((int64_t)64, cdiv(biggest, 256 * 4)), (unsigned)n_tensors);
  if (max_grad_norm > 0.f) {
    optim_sqnorm_kernel<<<grid, 256, 0, st>>>(t, state);
    CX_LAUNCH_CHECK();
  }
  optim_adam_kernel<<<grid, 256, 0, st>>>(t, state, beta1, beta2, eps, max_grad_norm);
  CX_LAUNCH_CHECK();
  optim_advance_kernel<<<1, 1, 0, st>>>(state, lr_gamma);
  CX_LAUNCH_CHECK();
  return 0;
}

// -------------------------------------------------------------------------- feature-interpolated points
int coordx_interp_fuse(int n_axes, const float* const* feats, const int64_t* n, int feat_width, int fusion_op,
                       const float* idx, int64_t n_points, float* h, void* stream) {
  if (n_axes < 1 || n_axes > COORDX_MAX_BRANCHES) return set_err(COORDX_ERR_ARG, "interp_fuse: bad axis count%s (%lld)", "", n_axes);
  if (!feats || !n || !idx || !h) return set_err(COORDX_ERR_ARG, "null argument%s");
  if (feat_width <= 0 || (feat_width & 3)) return set_err(COORDX_ERR_ARG, "interp_fuse: feature width must be a multiple of 4%s (%lld)", "", feat_width);
  if (fusion_op != COORDX_FUSE_SUM && fusion_op != COORDX_FUSE_PROD) return set_err(COORDX_ERR_ARG, "unknown fusion operator%s (%lld)", "", fusion_op);
  if (n_points <= 0) return 0;
  InterpArgs a{};
  a.C = n_axes; a.F = feat_width; a.op = fusion_op; a.idx = idx; a.P = n_points; a.h = h;
  for (int c = 0; c < n_axes; ++c) {
    if (!feats[c] || n[c] < 1) return set_err(COORDX_ERR_ARG, "interp_fuse: bad feature table%s (%lld)", "", c);
    if (reinterpret_cast<uintptr_t>(feats[c]) & 15) return set_err(COORDX_ERR_ARG, "interp_fuse: feature table not 16-byte aligned%s (%lld)", "", c);
    a.feat[c] = feats[c]; a.n[c] = n[c];
  }
  if (reinterpret_cast<uintptr_t>(h) & 15) return set_err(COORDX_ERR_ARG, "interp_fuse: output not 16-byte aligned%s");
  const int64_t threads = n_points * (feat_width / 4);
  interp_fuse_kernel<<<(unsigned)cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

// -------------------------------------------------------------------------- host-buffer inference
size_t coordx_infer_grid_host_workspace_bytes(const coordx_net* net, const int64_t* n, int precision) {
  if (check_net(net) != 0 || !n) return 0;
  size_t total = 0;
  int64_t P = 1;
  for (int c = 0; c < net->n_branches; ++c) {
    P *= n[c];
    total += align_up((size_t)n[c] * net->branch_in[c] * 4, 256);       // coords
    total += align_up((size_t)n[c] * net->fuse_width * 4, 256);         // features
  }
  size_t bws = 0;
  for (int c = 0; c < net->n_branches; ++c) bws = max(bws, coordx_branch_workspace_bytes(net, c, n[c], 0));
  const size_t tws = coordx_trunk_workspace_bytes(net, n, 0, n[0], precision, 0);
  total += align_up(max((size_t)net->n_branches * align_up(bws, 256), tws), 256);
  total += align_up((size_t)P * net->layer_out[net->n_layers - 1] * 4, 256);   // out
  return total + 1024;
}

int coordx_infer_grid_host(const coordx_net* net, const float* const* coords_host, const int64_t* n, float* out_host,
                           int precision, void* ws, size_t ws_bytes, void* stream) {
  CX_TRY(check_net(net));
  if (!coords_host || !n || !out_host || !ws) return set_err(COORDX_ERR_ARG, "null argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = net->n_branches;
  Arena ar(ws, ws_bytes);
  float* dcoords[COORDX_MAX_BRANCHES];
  float* dfeat[COORDX_MAX_BRANCHES];
  int64_t P = 1;
  for (int c = 0; c < C; ++c) {
    P *= n[c];
    dcoords[c] = ar.get<float>((size_t)n[c] * net->branch_in[c]);
    dfeat[c] = ar.get<float>((size_t)n[c] * net->fuse_width);
  }
  size_t bws = 0;
  for (int c = 0; c < C; ++c) bws = max(bws, coordx_branch_workspace_bytes(net, c, n[c], 0));
  const size_t tws = coordx_trunk_workspace_bytes(net, n, 0, n[0], precision, 0);
  const size_t sws = max((size_t)net->n_branches * align_up(bws, 256), tws);
  char* scratch = ar.get<char>(sws);
  const int ow = net->layer_out[net->n_layers - 1];
  float* dout = ar.get<float>((size_t)P * ow);
  if (!ar.ok() || !dout) return set_err(COORDX_ERR_WORKSPACE, "infer_grid_host: workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  for (int c = 0; c < C; ++c)
    CX_CUDA(cudaMemcpyAsync(dcoords[c], coords_host[c], (size_t)n[c] * net->branch_in[c] * 4, cudaMemcpyHostToDevice, st));
  bool small = true;
  for (int c = 0; c < C; ++c) small = small && n[c] <= BRANCH_FUSED_MAX;
  if (bf::supported(net) && small) {                 // every branch in ONE launch (two CTAs per SM run side by side)
    CX_TRY(bf::launch(net, 0, C, dcoords, n, dfeat, st));
  } else if (C > 1 && (size_t)C * align_up(bws, 256) <= sws) {
    // per-layer chains: the branches are independent, one internal side stream each (fork / join with events),
    // every branch with its own slice of the scratch area
    Fork* fk[COORDX_MAX_BRANCHES];
    for (int c = 1; c < C; ++c) CX_TRY(get_fork(c, &fk[c]));
    CX_CUDA(cudaEventRecord(fk[1]->ev_main[0], st));                  // the coordinates are on the device
    for (int c = 1; c < C; ++c) {
      CX_CUDA(cudaStreamWaitEvent(fk[c]->side, fk[1]->ev_main[0], 0));
      CX_TRY(coordx_branch_forward(net, c, dcoords[c], n[c], dfeat[c], scratch + c * align_up(bws, 256), bws, fk[c]->side));
      CX_CUDA(cudaEventRecord(fk[c]->ev_side[0], fk[c]->side));
    }
    CX_TRY(coordx_branch_forward(net, 0, dcoords[0], n[0], dfeat[0], scratch, bws, stream));
    for (int c = 1; c < C; ++c) CX_CUDA(cudaStreamWaitEvent(st, fk[c]->ev_side[0], 0));
  } else {
    for (int c = 0; c < C; ++c) CX_TRY(coordx_branch_forward(net, c, dcoords[c], n[c], dfeat[c], scratch, sws, stream));
  }
  const float* feats[COORDX_MAX_BRANCHES];
  for (int c = 0; c < C; ++c) feats[c] = dfeat[c];
  // Pinned (page-locked) result buffer: it is addressable from the device, so the trunk kernel stores the [P, O]
  // result straight into host memory - posted PCIe writes that overlap the computation (16.8 MB per 2048^2 grid in
  // 0.41 ms = 41 GB/s of the link's measured 56 GB/s), no device copy of the result and no device->host memcpy.
  {
    cudaPointerAttributes pa{};
    const bool direct_ok = getenv("COORDX_E2E_DIRECT") == nullptr || getenv("COORDX_E2E_DIRECT")[0] != '0';
    if (direct_ok && cudaPointerGetAttributes(&pa, out_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
      return coordx_trunk_forward(net, feats, n, 0, n[0], (float*)pa.devicePointer, precision, scratch, sws, stream);
    }
    cudaGetLastError();                              // pageable memory: not an error, take the staged path
  }
  // The [P, O] result is the only large transfer (16 MB for 2048^2: ~0.4 ms of PCIe, as long as the whole
  // computation).  The grid is evaluated in axis-0 slabs and each slab's device->host copy runs on a second
  // stream while the next slab is computed.
  static cudaStream_t copy_stream = nullptr;
  static cudaEvent_t ev_slab[8], ev_done = nullptr;
  if (!copy_stream) {
    CX_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) CX_CUDA(cudaEventCreateWithFlags(&ev_slab[i], cudaEventDisableTiming));
    CX_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
  }
  const int64_t inner = P / n[0];
  const size_t out_bytes = (size_t)P * ow * 4;
  int slabs = out_bytes >= ((size_t)4 << 20) ? 4 : 1;
  if (slabs > n[0]) slabs = (int)n[0];
  for (int i = 0; i < slabs; ++i) {
    const int64_t ra = n[0] * i / slabs, rb = n[0] * (i + 1) / slabs;
    float* dslab = dout + ra * inner * ow;
    CX_TRY(coordx_trunk_forward(net, feats, n, ra, rb, dslab, precision, scratch, sws, stream));
    if (slabs == 1) {
      CX_CUDA(cudaMemcpyAsync(out_host, dout, out_bytes, cudaMemcpyDeviceToHost, st));
    } else {
      CX_CUDA(cudaEventRecord(ev_slab[i], st));
      CX_CUDA(cudaStreamWaitEvent(copy_stream, ev_slab[i], 0));
      CX_CUDA(cudaMemcpyAsync(out_host + ra * inner * ow, dslab, (size_t)(rb - ra) * inner * ow * 4, cudaMemcpyDeviceToHost, copy_stream));
    }
  }
  if (slabs > 1) {                                  // the caller synchronises on `stream` only
    CX_CUDA(cudaEventRecord(ev_done, copy_stream));
    CX_CUDA(cudaStreamWaitEvent(st, ev_done, 0));
  }
  return 0;
}

}  // extern "C"
