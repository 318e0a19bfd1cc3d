// Separable evaluation of the SUM-fused SINE trunk: no per-point Hd x Hd GEMM at all.
//
// SplitFCBlock.forward_fusion (modules.py:329-354) with fusion_operator='sum' computes, for grid point (i, j),
//     out[i,j,o] = bf[o] + sum_n Wf[o,n] * sin(S * (W (a_i + b_j) + b)[n])
// where a_i / b_j are the branch features of axis 0 / of the remaining axes.  The layer is linear BEFORE the
// sine, so Z = U_i + V_j with the per-axis tables U = A W^T + b  [N_0, Nh] and V = B W^T  [inner, Nh], and
//     sin(S(U+V)) = sin(SU) cos(SV) + cos(SU) sin(SV)
// turns the whole trunk into ONE small GEMM over the per-axis tables,
//     out[(i,o), j] = [Wf sin(SU) | Wf cos(SU) | bf](i,o) . [cos(SV) | sin(SV) | 1](j)        K = 2 Nh + 1
// i.e. 2(2Nh+1)O FLOP per point instead of 2 Nh Hd + 2 Nh O (1026 vs 131 584 at Hd = 256; 12 294 vs 2 103 296 for
// the 1024-wide video net), and it holds for any hidden width.  The same identity gives the backward pass:
// with dZ[(i,j),n] = S sum_o dout[i,j,o] Wf[o,n] cos(S(U+V)) the trunk's gradients only need the two reductions
//     RA[i,n] = sum_j dZ = S sum_o Wf[o,n] (cos(SU) (D_o cosSV) - sin(SU) (D_o sinSV))[i,n]       D_o = dout[:,:,o]
//     RB[j,n] = sum_i dZ = S sum_o Wf[o,n] (cos(SV) (D_o^T cosSU) - sin(SV) (D_o^T sinSU))[j,n]
// (two GEMMs over the grid of dout with K = inner resp. N_0), then dW = RA^T A + RB^T B, db = colsum RA,
// dA = RA W, dB = RB W, dWf[o,n] = sum_i (sin(SU) (D_o cosSV) + cos(SU) (D_o sinSV))[i,n].
// More than two axes: axes 1.. are flattened into one inner table (their features summed by fuse_fwd_kernel,
// its gradient scattered back by fuse_bwd_kernel).
//
// Everything here is fp32 (FFMA GEMMs of gemm_f32.cuh, accurate sincos): the results agree with the direct
// evaluation to ~1e-7, so this path serves every `precision`.  COORDX_SUM_PATH=direct switches it off.
#pragma once
#include <stdlib.h>

#include "elementwise.cuh"
#include "gemm_f32.cuh"

namespace coordx {
namespace sep {

inline bool enabled() {                       // read on every call so tests can compare the two paths in one process
  const char* e = getenv("COORDX_SUM_PATH");
  return !(e && strcmp(e, "direct") == 0);
}

inline bool supported(const coordx_net* net) {
  if (net->fusion_op != COORDX_FUSE_SUM || net->act != COORDX_ACT_SINE) return false;
  if (net->fusion_before_act || net->has_code || net->fuse_width != net->hidden) return false;
  if (net->n_branches < 2 || net->n_layers - net->n_pre != 2) return false;   // one hidden trunk layer + the final linear
  const int Lh = net->n_pre, Lf = net->n_layers - 1;
  return net->layer_in[Lh] == net->hidden && net->layer_in[Lf] == net->layer_out[Lh] && net->layer_out[Lf] <= 16;
}

struct Plan {
  int C, Kin, Nh, O, K1;
  int64_t R, inner, inner2;     // slab rows of axis 0 | prod_{c>0} n_c | prod_{c>1} n_c
  int64_t rb, jb;               // forward chunking: rows of axis 0 / flattened inner points per chunk
};

static Plan make_plan(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1) {
  Plan p{};
  p.C = net->n_branches;
  p.Kin = net->hidden;
  p.Nh = net->layer_out[net->n_pre];
  p.O = net->layer_out[net->n_layers - 1];
  p.K1 = 2 * p.Nh + 4;              // table row stride: [2Nh values | bias or 1 | 3 unused], 16-byte aligned rows
  p.R = r1 - r0;
  p.inner = 1;
  p.inner2 = 1;
  for (int c = 1; c < p.C; ++c) p.inner *= n[c];
  for (int c = 2; c < p.C; ++c) p.inner2 *= n[c];
  p.rb = min(p.R, (int64_t)8192);
  const int64_t rows1 = max((int64_t)1, (int64_t)32768 / max(p.inner2, (int64_t)1));   // whole axis-1 rows per chunk
  p.jb = min(p.inner, rows1 * p.inner2);
  return p;
}

static __device__ __noinline__ void sincos_accurate(float x, float* s, float* c) { sincosf(x, s, c); }

// T[r, 0:Nh] = cos(S X[r,:]),  T[r, Nh:2Nh] = sin(S X[r,:]),  T[r, 2Nh] = 1 when ld > 2Nh
__global__ void __launch_bounds__(256)
table_cs_kernel(const float* __restrict__ X, float S, int64_t rows, int Nh, int64_t ld, float* __restrict__ T) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * Nh) return;
  const int64_t r = idx / Nh;
  const int n = (int)(idx % Nh);
  float s, c;
  sincos_accurate(S * X[idx], &s, &c);
  float* row = T + r * ld;
  row[n] = c;
  row[Nh + n] = s;
  if (n == 0 && ld > 2 * Nh) row[2 * Nh] = 1.f;
}

// TU[(i O + o), :] = [ Wf[o,:] sin(S U[i,:]) | Wf[o,:] cos(S U[i,:]) | bf[o] ]
__global__ void __launch_bounds__(256)
table_outer_kernel(const float* __restrict__ U, const float* __restrict__ Wf, const float* __restrict__ bf, float S,
                   int64_t rows, int Nh, int O, float* __restrict__ TU) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * Nh) return;
  const int64_t i = idx / Nh;
  const int n = (int)(idx % Nh);
  const int K1 = 2 * Nh + 4;
  float s, c;
  sincos_accurate(S * U[idx], &s, &c);
  for (int o = 0; o < O; ++o) {
    const float w = Wf[(int64_t)o * Nh + n];
    float* row = TU + (i * O + o) * K1;
    row[n] = w * s;
    row[Nh + n] = w * c;
    if (n == 0) row[2 * Nh] = bf ? bf[o] : 0.f;
  }
}

// T [M, 2Nh] = [cos | sin] of one axis' table, G [O][M, 2Nh] = D_o (or D_o^T) times the OTHER axis' [cos | sin]:
//   R[m,n]      = S sum_o Wf[o,n] (cos G_o[m,n] - sin G_o[m,Nh+n])       (row sums of dZ)
//   G_o[m,n]   <- sin G_o[m,n] + cos G_o[m,Nh+n]   when want_q            (its column sums are dWf[o,:])
__global__ void __launch_bounds__(256)
combine_kernel(const float* __restrict__ T, float* __restrict__ G, const float* __restrict__ Wf, float S, int64_t M, int Nh,
               int O, float* __restrict__ Rout, int want_q) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= M * Nh) return;
  const int64_t m = idx / Nh;
  const int n = (int)(idx % Nh);
  const float c = T[m * 2 * Nh + n], s = T[m * 2 * Nh + Nh + n];
  float r = 0.f;
  for (int o = 0; o < O; ++o) {
    float* g = G + ((int64_t)o * M + m) * 2 * Nh;
    const float gc = g[n], gs = g[Nh + n];
    r = fmaf(Wf[(int64_t)o * Nh + n], c * gc - s * gs, r);
    if (want_q) g[n] = s * gc + c * gs;
  }
  Rout[idx] = S * r;
}

struct Bufs {
  float *U, *TU, *Fin, *V, *TV;                        // forward
  float *TUr, *TVr, *G, *Hm, *RA, *RB, *dFin;          // backward
  float *wt, *tw;                                      // backward: W_h transposed; operand chunks of the tensor-core weight gradient
  size_t tw_floats;
  float *Do, *DoT, *TVrT, *TUrT;                       // backward, tensor-core reductions of dout: one channel of dout [R, Jpad], its
  int64_t Jpad, Rpad;                                  // transpose [J, Rpad], and the transposed [cos | sin] tables [2Nh, Jpad], [2Nh, Rpad]
};

// the two reductions of dout over the grid run on the tensor cores when the tables are small enough to transpose
static bool reductions_on_tc(const Plan& p) {
  const int64_t jp = (p.inner + 31) / 32 * 32, rp = (p.R + 31) / 32 * 32;
  return p.R >= 96 && p.inner >= 128 && p.Nh >= 64 && (int64_t)2 * p.Nh * (jp + rp) + p.R * jp + p.inner * rp <= ((int64_t)96 << 20);
}

static void carve(Arena& ar, const Plan& p, bool backward, Bufs& b) {
  b = Bufs{};
  if (!backward) {
    b.U = ar.get<float>((size_t)p.rb * p.Nh);
    b.TU = ar.get<float>((size_t)p.rb * p.O * p.K1);
    if (p.C > 2) b.Fin = ar.get<float>((size_t)p.jb * p.Kin);
    b.V = ar.get<float>((size_t)p.jb * p.Nh);
    b.TV = ar.get<float>((size_t)p.jb * p.K1);
  } else {
    b.U = ar.get<float>((size_t)p.R * p.Nh);
    b.TUr = ar.get<float>((size_t)p.R * 2 * p.Nh);
    b.G = ar.get<float>((size_t)p.O * p.R * 2 * p.Nh);
    b.RA = ar.get<float>((size_t)p.R * p.Nh);
    if (p.C > 2) b.Fin = ar.get<float>((size_t)p.inner * p.Kin);
    if (p.C > 2) b.dFin = ar.get<float>((size_t)p.inner * p.Kin);
    b.V = ar.get<float>((size_t)p.inner * p.Nh);
    b.TVr = ar.get<float>((size_t)p.inner * 2 * p.Nh);
    b.Hm = ar.get<float>((size_t)p.O * p.inner * 2 * p.Nh);
    b.RB = ar.get<float>((size_t)p.inner * p.Nh);
    b.wt = ar.get<float>((size_t)p.Nh * p.Kin);
    const int64_t rows = min(max(p.R, p.inner), (int64_t)DW_ROWS);
    b.tw_floats = (size_t)(p.Nh + p.Kin) * (size_t)((rows + 31) / 32 * 32);
    b.tw = ar.get<float>(b.tw_floats);
    if (reductions_on_tc(p)) {
      b.Jpad = (p.inner + 31) / 32 * 32;
      b.Rpad = (p.R + 31) / 32 * 32;
      b.Do = ar.get<float>((size_t)p.R * b.Jpad);
      b.DoT = ar.get<float>((size_t)p.inner * b.Rpad);
      b.TVrT = ar.get<float>((size_t)2 * p.Nh * b.Jpad);
      b.TUrT = ar.get<float>((size_t)2 * p.Nh * b.Rpad);
    }
  }
}

static size_t workspace_bytes(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1, bool backward) {
  const Plan p = make_plan(net, n, r0, r1);
  Arena ar(nullptr, 0);
  Bufs b;
  carve(ar, p, backward, b);
  return ar.off + 256;
}

// the flattened inner axis' feature rows [j0, j0+cnt): feats[1] itself for two axes, else the sum over axes 1..
static int inner_features(const coordx_net* net, const float* const* feats, const int64_t* n, const Plan& p, int64_t j0,
                          int64_t cnt, float* Fin, const float** out, cudaStream_t st) {
  if (p.C == 2) {
    *out = feats[1] + j0 * p.Kin;
    return 0;
  }
  FuseArgs a{};
  a.C = p.C - 1;
  for (int c = 1; c < p.C; ++c) { a.feat[c - 1] = feats[c]; a.n[c - 1] = n[c]; }
  a.F = a.Hd = p.Kin;
  a.gs = 1;
  a.op = COORDX_FUSE_SUM;
  a.inner = p.inner2;
  a.row0 = j0 / p.inner2;
  a.nrows = cnt / p.inner2;
  fuse_fwd_kernel<<<(unsigned)cdiv(cnt * p.Kin, 256), 256, 0, st>>>(a, cnt, Fin);
  CX_LAUNCH_CHECK();
  *out = Fin;
  return 0;
}

static int forward(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1, float* out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  const Plan p = make_plan(net, n, r0, r1);
  Arena ar(ws, ws_bytes);
  Bufs b;
  carve(ar, p, false, b);
  if (!ws || !ar.ok()) return set_err(COORDX_ERR_WORKSPACE, "trunk forward (separable sum): workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const int Lh = net->n_pre, Lf = net->n_layers - 1;
  const float S = 30.f * net->act_scale;
  for (int64_t ra = 0; ra < p.R; ra += p.rb) {
    const int64_t rows = min(p.rb, p.R - ra);
    CX_TRY(linear_forward(feats[0] + (r0 + ra) * p.Kin, p.Kin, net->layer_w[Lh], net->layer_b[Lh], 1.f, (int)rows, p.Nh, p.Kin,
                          b.U, nullptr, p.Nh, -1, S, st));
    table_outer_kernel<<<(unsigned)cdiv(rows * p.Nh, 256), 256, 0, st>>>(b.U, net->layer_w[Lf], net->layer_b[Lf], S, rows, p.Nh, p.O, b.TU);
    CX_LAUNCH_CHECK();
    for (int64_t j0 = 0; j0 < p.inner; j0 += p.jb) {
      const int64_t cnt = min(p.jb, p.inner - j0);
      const float* fin = nullptr;
      CX_TRY(inner_features(net, feats, n, p, j0, cnt, b.Fin, &fin, st));
      CX_TRY(linear_forward(fin, p.Kin, net->layer_w[Lh], nullptr, 0.f, (int)cnt, p.Nh, p.Kin, b.V, nullptr, p.Nh, -1, S, st));
      table_cs_kernel<<<(unsigned)cdiv(cnt * p.Nh, 256), 256, 0, st>>>(b.V, S, cnt, p.Nh, p.K1, b.TV);
      CX_LAUNCH_CHECK();
      for (int o = 0; o < p.O; ++o) {      // out[(i, j), o] = TU[(i,o), :] . TV[j, :]
        {                                    // large grids: 3xTF32 on the tensor cores, K = 2Nh, bf[o] added in the epilogue
          g3::Args t{};
          t.A = b.TU + (int64_t)o * p.K1; t.lda = (int64_t)p.O * p.K1;
          t.B = b.TV; t.ldb = p.K1;
          t.M = (int)rows; t.N = (int)cnt; t.K = 2 * p.Nh;
          t.Y = out + (ra * p.inner + j0) * p.O + o; t.ldc = p.inner * p.O; t.cs = p.O;
          t.bias1 = net->layer_b[Lf] ? net->layer_b[Lf] + o : nullptr; t.act = -1;
          if (g3::usable(t)) { CX_TRY(g3::launch(t, st)); continue; }
        }
        GemmArgs g{};
        g.A = b.TU + (int64_t)o * p.K1; g.lda = (int64_t)p.O * p.K1;
        g.B = b.TV; g.ldb = p.K1;
        g.M = (int)rows; g.N = (int)cnt; g.K = 2 * p.Nh + 1;
        g.Y = out + (ra * p.inner + j0) * p.O + o; g.ldc = p.inner * p.O; g.cs = p.O;
        g.act = -1;
        CX_TRY((launch_sgemm<true, true, EPI_LINEAR>(g, st)));
      }
    }
  }
  return 0;
}

// rough bound on the backward tables (fp32 words); beyond it the caller takes the direct path
static bool backward_fits(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1) {
  const Plan p = make_plan(net, n, r0, r1);
  const double words = (double)(p.R + p.inner) * p.Nh * (6.0 + 2.0 * p.O);
  return words * 4.0 < 16e9 && p.R < (1 << 30) && p.inner < (1 << 30);
}

// dst[r, j] = src[(r * J + j) * stride] for j < J, 0 for J <= j < jpad   (one channel of the interleaved dout, rows padded)
__global__ void __launch_bounds__(256)
channel_pad_kernel(const float* __restrict__ src, int stride, int R, int J, int jpad, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)R * jpad) return;
  const int64_t r = i / jpad;
  const int j = (int)(i - r * jpad);
  dst[i] = j < J ? __ldg(src + (r * J + j) * stride) : 0.f;
}

static int backward(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                    const float* dout, coordx_grads* grads, float* const* dfeat, void* ws, size_t ws_bytes, cudaStream_t st) {
  const Plan p = make_plan(net, n, r0, r1);
  Arena ar(ws, ws_bytes);
  Bufs b;
  carve(ar, p, true, b);
  if (!ws || !ar.ok()) return set_err(COORDX_ERR_WORKSPACE, "trunk backward (separable sum): workspace too small%s (%lld < %lld)", "", (long long)ws_bytes, (long long)ar.off);
  const int Lh = net->n_pre, Lf = net->n_layers - 1;
  const float S = 30.f * net->act_scale;
  const int R = (int)p.R, J = (int)p.inner, Nh = p.Nh, N2 = 2 * p.Nh;
  const float* f0 = feats[0] + r0 * p.Kin;
  const float* fin = nullptr;
  CX_TRY(inner_features(net, feats, n, p, 0, p.inner, b.Fin, &fin, st));
  // per-axis pre-activation tables and their [cos | sin]
  CX_TRY(linear_forward(f0, p.Kin, net->layer_w[Lh], net->layer_b[Lh], 1.f, R, Nh, p.Kin, b.U, nullptr, Nh, -1, S, st));
  CX_TRY(linear_forward(fin, p.Kin, net->layer_w[Lh], nullptr, 0.f, J, Nh, p.Kin, b.V, nullptr, Nh, -1, S, st));
  table_cs_kernel<<<(unsigned)cdiv((int64_t)R * Nh, 256), 256, 0, st>>>(b.U, S, R, Nh, N2, b.TUr);
  CX_LAUNCH_CHECK();
  table_cs_kernel<<<(unsigned)cdiv((int64_t)J * Nh, 256), 256, 0, st>>>(b.V, S, J, Nh, N2, b.TVr);
  CX_LAUNCH_CHECK();
  // the two reductions of dout over the grid
  const char* ffma_env = getenv("COORDX_FP32_GEMM");
  const bool red_tc = b.Do != nullptr && !(ffma_env && strcmp(ffma_env, "ffma") == 0);
  if (red_tc) {
    // tensor cores (3xTF32): G_o = D_o TVr and H_o = D_o^T TUr as NT GEMMs over zero-padded, transposed operands
    const int jp = (int)b.Jpad, rp = (int)b.Rpad;
    transpose_pad_kernel<<<dim3((unsigned)cdiv(N2, 32), (unsigned)(jp / 32)), 256, 0, st>>>(b.TVr, N2, J, N2, jp, b.TVrT);
    CX_LAUNCH_CHECK();
    transpose_pad_kernel<<<dim3((unsigned)cdiv(N2, 32), (unsigned)(rp / 32)), 256, 0, st>>>(b.TUr, N2, R, N2, rp, b.TUrT);
    CX_LAUNCH_CHECK();
    for (int o = 0; o < p.O; ++o) {
      channel_pad_kernel<<<(unsigned)cdiv((int64_t)R * jp, 256), 256, 0, st>>>(dout + o, p.O, R, J, jp, b.Do);
      CX_LAUNCH_CHECK();
      transpose_pad_kernel<<<dim3((unsigned)cdiv(J, 32), (unsigned)(rp / 32)), 256, 0, st>>>(b.Do, jp, R, J, rp, b.DoT);
      CX_LAUNCH_CHECK();
      g3::Args t{};
      t.A = b.Do; t.lda = jp; t.B = b.TVrT; t.ldb = jp; t.M = R; t.N = N2; t.K = jp;
      t.Y = b.G + (int64_t)o * R * N2; t.ldc = N2; t.cs = 1; t.act = -1;
      CX_TRY((N2 >= 256 && cdiv(R, 128) * cdiv(N2, 256) >= 32) ? g3::launch_bn<256>(t, st) : g3::launch_bn<64>(t, st));
      g3::Args u{};
      u.A = b.DoT; u.lda = rp; u.B = b.TUrT; u.ldb = rp; u.M = J; u.N = N2; u.K = rp;
      u.Y = b.Hm + (int64_t)o * J * N2; u.ldc = N2; u.cs = 1; u.act = -1;
      CX_TRY((N2 >= 256 && cdiv(J, 128) * cdiv(N2, 256) >= 32) ? g3::launch_bn<256>(u, st) : g3::launch_bn<64>(u, st));
    }
  }
  for (int o = 0; o < (red_tc ? 0 : p.O); ++o) {
    GemmArgs g{};                                   // G_o [R, 2Nh] = D_o [R, J] . TVr [J, 2Nh]
    g.A = dout + o; g.lda = p.inner * p.O; g.sa = p.O;
    g.B = b.TVr; g.ldb = N2;
    g.M = R; g.N = N2; g.K = J;
    g.Y = b.G + (int64_t)o * R * N2; g.ldc = N2; g.act = -1;
    CX_TRY((launch_sgemm<true, false, EPI_DX>(g, st)));
    GemmArgs h{};                                   // H_o [J, 2Nh] = D_o^T [J, R] . TUr [R, 2Nh]
    h.A = dout + o; h.lda = p.inner * p.O; h.sa = p.O;
    h.B = b.TUr; h.ldb = N2;
    h.M = J; h.N = N2; h.K = R;
    h.Y = b.Hm + (int64_t)o * J * N2; h.ldc = N2; h.act = -1;
    CX_TRY((launch_sgemm<false, false, EPI_DX>(h, st)));
  }
  combine_kernel<<<(unsigned)cdiv((int64_t)R * Nh, 256), 256, 0, st>>>(b.TUr, b.G, net->layer_w[Lf], S, R, Nh, p.O, b.RA, 1);
  CX_LAUNCH_CHECK();
  combine_kernel<<<(unsigned)cdiv((int64_t)J * Nh, 256), 256, 0, st>>>(b.TVr, b.Hm, net->layer_w[Lf], S, J, Nh, p.O, b.RB, 0);
  CX_LAUNCH_CHECK();
  // final layer: dWf[o,:] = column sums of Q_o (left in G_o[:, 0:Nh]);  dbf = column sums of dout
  for (int o = 0; o < p.O; ++o)
    if (grads->layer_w[Lf]) CX_TRY(bias_backward(b.G + (int64_t)o * R * N2, N2, R, Nh, 1.f, grads->layer_w[Lf] + (int64_t)o * Nh, st));
  if (grads->layer_b[Lf]) CX_TRY(bias_backward(dout, p.O, (int)min(p.R * p.inner, (int64_t)INT32_MAX), p.O, 1.f, grads->layer_b[Lf], st));
  // hidden trunk layer
  if (grads->layer_b[Lh]) CX_TRY(bias_backward(b.RA, Nh, R, Nh, 1.f, grads->layer_b[Lh], st));
  // (weight gradients and feature gradients on the tensor cores where the shape pays - 3xTF32 with an atomic epilogue,
  //  dw_tensor / nn_tensor_acc in gemm_f32.cuh - else the FFMA kernels)
  if (grads->layer_w[Lh]) {
    int tc = 0;
    CX_TRY(dw_tensor(b.RA, Nh, f0, p.Kin, R, Nh, p.Kin, grads->layer_w[Lh], b.tw, b.tw_floats, st, &tc));
    if (!tc) CX_TRY(linear_backward_dw(b.RA, Nh, f0, p.Kin, R, Nh, p.Kin, grads->layer_w[Lh], st));
    CX_TRY(dw_tensor(b.RB, Nh, fin, p.Kin, J, Nh, p.Kin, grads->layer_w[Lh], b.tw, b.tw_floats, st, &tc));
    if (!tc) CX_TRY(linear_backward_dw(b.RB, Nh, fin, p.Kin, J, Nh, p.Kin, grads->layer_w[Lh], st));
  }
  // branch features: dA += RA W, dB += RB W
  if (dfeat[0] || (p.C == 2 && dfeat[1])) {
    dim3 tg((unsigned)cdiv(p.Kin, 32), (unsigned)cdiv(Nh, 32));
    transpose_kernel<<<tg, 256, 0, st>>>(net->layer_w[Lh], Nh, p.Kin, b.wt);      // W_h [Nh, Kin] -> [Kin, Nh]
    CX_LAUNCH_CHECK();
  }
  if (dfeat[0]) {
    int tc = 0;
    CX_TRY(nn_tensor_acc(b.RA, Nh, b.wt, R, p.Kin, Nh, dfeat[0] + r0 * p.Kin, p.Kin, st, &tc));
    if (!tc) {
      GemmArgs g{};
      g.A = b.RA; g.lda = Nh; g.B = net->layer_w[Lh]; g.ldb = p.Kin; g.M = R; g.N = p.Kin; g.K = Nh;
      g.Y = dfeat[0] + r0 * p.Kin; g.ldc = p.Kin;
      CX_TRY((launch_sgemm<true, false, EPI_ATOMIC>(g, st)));
    }
  }
  if (p.C == 2) {
    if (dfeat[1]) {
      int tc = 0;
      CX_TRY(nn_tensor_acc(b.RB, Nh, b.wt, J, p.Kin, Nh, dfeat[1], p.Kin, st, &tc));
      if (!tc) {
        GemmArgs g{};
        g.A = b.RB; g.lda = Nh; g.B = net->layer_w[Lh]; g.ldb = p.Kin; g.M = J; g.N = p.Kin; g.K = Nh;
        g.Y = dfeat[1]; g.ldc = p.Kin;
        CX_TRY((launch_sgemm<true, false, EPI_ATOMIC>(g, st)));
      }
    }
  } else {
    CX_TRY(linear_backward_dx(b.RB, Nh, net->layer_w[Lh], J, Nh, p.Kin, b.dFin, p.Kin, nullptr, 0, -1, S, st));
    FuseArgs a{};
    a.C = p.C - 1;
    for (int c = 1; c < p.C; ++c) { a.feat[c - 1] = feats[c]; a.n[c - 1] = n[c]; }
    a.F = a.Hd = p.Kin;
    a.gs = 1;
    a.op = COORDX_FUSE_SUM;
    a.inner = p.inner2;
    a.row0 = 0;
    a.nrows = n[1];
    for (int c = 1; c < p.C; ++c) {
      if (!dfeat[c]) continue;
      const int64_t cnt_b = n[c];
      const int64_t T = p.inner / cnt_b;
      const int64_t base_blocks = cdiv(p.Kin, 128) * cnt_b;
      int64_t splits = max((int64_t)1, min(cdiv(T, 64), cdiv((int64_t)148 * 16, base_blocks)));
      const int64_t tpb = cdiv(T, splits);
      splits = cdiv(T, tpb);
      dim3 grid((unsigned)cnt_b, (unsigned)cdiv(p.Kin, 128), (unsigned)splits);
      fuse_bwd_kernel<<<grid, 128, 0, st>>>(a, c - 1, b.dFin, dfeat[c], nullptr, T, tpb);
      CX_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // namespace sep
}  // namespace coordx
