// FP32 (FFMA) GEMM family for the 1e-5-relative parity path and for every
// configuration the tcgen05 kernels do not cover.  One register-tiled kernel,
// three operand layouts:
//   NT  Y[M,N]  = X[M,K] . W[N,K]^T (+bias, act)     BatchLinear.forward, modules.py:26-27 (+:39)
//   NN  dX[M,K] = dZ[M,N] . W[N,K]  (* act'(Zprev))  autograd of the above (training.py:138)
//   TN  dW[N,K]+= dZ[M,N]^T . X[M,K]  (split over M, atomics)
#pragma once
#include "common.cuh"
#include "gemm_tf32x3.cuh"

namespace coordx {

enum { EPI_LINEAR = 0, EPI_DX = 1, EPI_ATOMIC = 2 };

struct GemmArgs {
  const float* A;
  const float* B;
  int M, N, K;          // C[M,N] = sum_k A(m,k) * B(k,n)
  int64_t lda, ldb;     // leading dimension of the row-major storage of A / B
  float* Y;             // EPI_LINEAR: activated output (nullable) | EPI_DX: output | EPI_ATOMIC: += target
  float* Z;             // EPI_LINEAR: pre-activation output (nullable)
  int64_t ldc;
  const float* bias;    // EPI_LINEAR (nullable)
  float bscale;
  int act;              // COORDX_ACT_* or -1 (identity)
  float S;
  const float* Zprev;   // EPI_DX: multiply by act'(Zprev[m,n]) when non-null
  int64_t ldz;
  int kchunk;           // EPI_ATOMIC: K range per blockIdx.z
  int64_t sa;           // element stride of A along its contiguous index (0 = 1): an interleaved channel of [.., O]
  int64_t cs;           // EPI_LINEAR: column stride of Y / Z (0 = 1): writes one channel of an interleaved [.., O]
};

// AK : A(m,k) = A[m*lda + k]  (k contiguous)   else A[k*lda + m]
// BKc: B(k,n) = B[n*ldb + k]  (k contiguous)   else B[k*ldb + n]
// BK: k-depth of a shared-memory stage.  8 for the 128 x 128 tile (static shared memory); 32 for the 64 x 64 tile of
// the small per-axis GEMMs, which are latency-bound (one CTA per SM): 4x fewer barriers and 16 instead of 4 global
// loads in flight per thread.  K-contiguous operands are walked as BK/8 sub-tiles of 8 k x rows, so a warp always
// covers 4 rows x 8 k: whole 32-byte sectors from global memory, conflict-free transposed stores to shared memory.
template <int BM, int BN, int TM, int TN, bool AK, bool BKc, int EPI, int BK = 8>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int RM = TM / 4, RN = TN / 4;      // 4-wide register groups, strided over the tile
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int kbeg = 0, kend = g.K;
  if (EPI == EPI_ATOMIC) {
    kbeg = blockIdx.z * g.kchunk;
    kend = min(g.K, kbeg + g.kchunk);
    if (kbeg >= kend) return;
  }

  const int64_t sa = g.sa ? g.sa : 1, cs = g.cs ? g.cs : 1;
  constexpr int LA = (BM * BK + NT - 1) / NT, LB = (BN * BK + NT - 1) / NT;
  float ra[LA], rb[LB];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT;
      int r, kk;
      if (AK) { r = (e / 8) % BM; kk = (e / (8 * BM)) * 8 + e % 8; } else { kk = e / BM; r = e % BM; }
      int m = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (e < BM * BK && m < g.M && k < kend) v = AK ? g.A[(int64_t)m * g.lda + k * sa] : g.A[(int64_t)k * g.lda + m * sa];
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT;
      int c, kk;
      if (BKc) { c = (e / 8) % BN; kk = (e / (8 * BN)) * 8 + e % 8; } else { kk = e / BN; c = e % BN; }
      int n = n0 + c, k = k0 + kk;
      float v = 0.f;
      if (e < BN * BK && n < g.N && k < kend) v = BKc ? g.B[(int64_t)n * g.ldb + k] : g.B[(int64_t)k * g.ldb + n];
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT;
      int r, kk;
      if (AK) { r = (e / 8) % BM; kk = (e / (8 * BM)) * 8 + e % 8; } else { kk = e / BM; r = e % BM; }
      if (e < BM * BK) As[buf][kk][r] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT;
      int c, kk;
      if (BKc) { c = (e / 8) % BN; kk = (e / (8 * BN)) * 8 + e % 8; } else { kk = e / BN; c = e % BN; }
      if (e < BN * BK) Bs[buf][kk][c] = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  gload(kbeg);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][r * (BM / RM) + ty * 4]);
        a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][r * (BN / RN) + tx * 4]);
        b[r * 4 + 0] = v.x; b[r * 4 + 1] = v.y; b[r * 4 + 2] = v.z; b[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / RM) + ty * 4 + (i % 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j / 4) * (BN / RN) + tx * 4 + (j % 4);
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_LINEAR) {
        if (g.bias) v += g.bias[n] * g.bscale;
        if (g.Z) g.Z[(int64_t)m * g.ldc + n * cs] = v;
        if (g.Y) g.Y[(int64_t)m * g.ldc + n * cs] = g.act < 0 ? v : act_fwd(g.act, g.S, v);
      } else if (EPI == EPI_DX) {
        if (g.Zprev) v *= act_grad(g.act, g.S, g.Zprev[(int64_t)m * g.ldz + n]);
        g.Y[(int64_t)m * g.ldc + n] = v;
      } else {
        atomicAdd(&g.Y[(int64_t)m * g.ldc + n], v);
      }
    }
  }
}

template <bool AK, bool BKc, int EPI>
static int launch_sgemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  int splits = 1;
  GemmArgs a = g;
  // 128x128 tiles only when they still give every SM a CTA; small problems (the per-axis
  // branch layers: M = N_c <= a few thousand rows) get 4x more 64x64 CTAs instead
  // (the split-K weight-gradient GEMMs take the big tile only when their K range still yields >= 148 CTAs)
  const bool big = (g.M > 64 && g.N > 64) &&
                   (EPI == EPI_ATOMIC ? cdiv(g.M, 128) * cdiv(g.N, 128) * cdiv(g.K, 256) >= 148 : cdiv(g.M, 128) * cdiv(g.N, 128) >= 148);
  const int BM = big ? 128 : 64, BN = big ? 128 : 64;
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), 1);
  if (EPI == EPI_ATOMIC) {
    // split the (long) reduction so the small [N,K] output still fills the chip
    int64_t tiles = (int64_t)grid.x * grid.y;
    splits = (int)max((int64_t)1, min(cdiv(g.K, big ? 256 : 128), cdiv(148 * 4, tiles)));
    a.kchunk = (int)(cdiv(cdiv(g.K, splits), 8) * 8);
    splits = (int)cdiv(g.K, a.kchunk);
    grid.z = splits;
  }
  const int krange = EPI == EPI_ATOMIC ? a.kchunk : g.K;
  if (big)
    sgemm_kernel<128, 128, 8, 8, AK, BKc, EPI><<<grid, 256, 0, st>>>(a);
  else if (krange >= 64)
    sgemm_kernel<64, 64, 4, 4, AK, BKc, EPI, 32><<<grid, 256, 0, st>>>(a);
  else
    sgemm_kernel<64, 64, 4, 4, AK, BKc, EPI><<<grid, 256, 0, st>>>(a);
  CX_LAUNCH_CHECK();
  return 0;
}

// Narrow output (N <= 8: the Hd -> O final layer): one warp per row.
template <int NMAX>
__global__ void __launch_bounds__(256)
linear_narrow_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, const float* __restrict__ bias,
                     float bscale, int M, int N, int K, float* __restrict__ Y, float* __restrict__ Z, int64_t ldc,
                     int act, float S) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  const float* x = X + (int64_t)warp * ldx;
  for (int k = lane; k < K; k += 32) {
    float xv = x[k];
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) acc[n] = fmaf(xv, W[(int64_t)n * K + k], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && n < N) {
      if (bias) v += bias[n] * bscale;
      if (Z) Z[(int64_t)warp * ldc + n] = v;
      if (Y) Y[(int64_t)warp * ldc + n] = act < 0 ? v : act_fwd(act, S, v);
    }
  }
}

// Y = act(X W^T + b*bscale); Z (pre-activation) optional.  X [M,K] (ld ldx), W [N,K].
// Tiny reduction (K <= 4: coord_linears on raw coordinates, f_c = 1..3): an elementwise kernel, one thread per output,
// coalesced along n - a 64 x 64-tile GEMM spends its time on k-steps that are all zero padding.
__global__ void __launch_bounds__(256)
linear_smallk_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, const float* __restrict__ bias,
                     float bscale, int64_t M, int N, int K, float* __restrict__ Y, float* __restrict__ Z, int64_t ldc, int act, float S) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int64_t m = idx / N;
  const int n = (int)(idx - m * N);
  float v = bias ? __ldg(bias + n) * bscale : 0.f;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(X + m * ldx + k), __ldg(W + (int64_t)n * K + k), acc);
  v += acc;                                            // (x . w) + b: the reference's matmul-then-add order
  if (Z) Z[m * ldc + n] = v;
  if (Y) Y[m * ldc + n] = act < 0 ? v : act_fwd(act, S, v);
}

static int linear_forward(const float* X, int64_t ldx, const float* W, const float* bias, float bscale, int M, int N,
                          int K, float* Y, float* Z, int64_t ldc, int act, float S, cudaStream_t st) {
  if (M <= 0) return 0;
  if (K <= 4 && N > 8) {
    linear_smallk_kernel<<<(unsigned)cdiv((int64_t)M * N, 256), 256, 0, st>>>(X, ldx, W, bias, bscale, M, N, K, Y, Z, ldc, act, S);
    CX_LAUNCH_CHECK();
    return 0;
  }
  if (N <= 8) {
    int blocks = (int)cdiv((int64_t)M * 32, 256);
    linear_narrow_kernel<8><<<blocks, 256, 0, st>>>(X, ldx, W, bias, bscale, M, N, K, Y, Z, ldc, act, S);
    CX_LAUNCH_CHECK();
    return 0;
  }
  {                                                   // large layers: fp32-accurate 3xTF32 on the tensor cores
    g3::Args t{};
    t.A = X; t.lda = ldx; t.B = W; t.ldb = K; t.M = M; t.N = N; t.K = K;
    t.Y = Y; t.Z = Z; t.ldc = ldc; t.cs = 1; t.bias = bias; t.bscale = bscale; t.act = act; t.S = S;
    if (g3::usable(t)) return g3::launch(t, st);
  }
  GemmArgs g{};
  g.A = X; g.lda = ldx; g.B = W; g.ldb = K; g.M = M; g.N = N; g.K = K;
  g.Y = Y; g.Z = Z; g.ldc = ldc; g.bias = bias; g.bscale = bscale; g.act = act; g.S = S;
  return launch_sgemm<true, true, EPI_LINEAR>(g, st);
}

// Narrow input gradient (K <= 8: the coordinates of a branch, f_c = 1..3 without positional encoding): one warp per row.
// dX[m, k] = sum_n dZ[m, n] W[n, k]
template <int KMAX>
__global__ void __launch_bounds__(256)
dx_narrow_kernel(const float* __restrict__ dZ, int64_t lddz, const float* __restrict__ W, int M, int N, int K,
                 float* __restrict__ dX, int64_t lddx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  const float* dz = dZ + (int64_t)warp * lddz;
  for (int n = lane; n < N; n += 32) {
    const float d = dz[n];
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) acc[k] = fmaf(d, W[(int64_t)n * K + k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && k < K) dX[(int64_t)warp * lddx + k] = v;
  }
}

// dX[M,K] = dZ[M,N] W[N,K], optionally * act'(Zprev[M,K])
static int linear_backward_dx(const float* dZ, int64_t lddz, const float* W, int M, int N, int K, float* dX, int64_t lddx,
                              const float* Zprev, int64_t ldz, int act, float S, cudaStream_t st) {
  if (K <= 8 && !Zprev && M > 0) {
    dx_narrow_kernel<8><<<(unsigned)cdiv((int64_t)M * 32, 256), 256, 0, st>>>(dZ, lddz, W, M, N, K, dX, lddx);
    CX_LAUNCH_CHECK();
    return 0;
  }
  GemmArgs g{};
  g.A = dZ; g.lda = lddz; g.B = W; g.ldb = K; g.M = M; g.N = K; g.K = N;
  g.Y = dX; g.ldc = lddx; g.Zprev = Zprev; g.ldz = ldz; g.act = act; g.S = S;
  return launch_sgemm<true, false, EPI_DX>(g, st);
}

// dW[N,K] += dZ[M,N]^T X[M,K]
static int linear_backward_dw(const float* dZ, int64_t lddz, const float* X, int64_t ldx, int M, int N, int K, float* dW,
                              cudaStream_t st) {
  GemmArgs g{};
  g.A = dZ; g.lda = lddz; g.B = X; g.ldb = ldx; g.M = N; g.N = K; g.K = M;
  g.Y = dW; g.ldc = K;
  return launch_sgemm<false, false, EPI_ATOMIC>(g, st);
}

// Wt[c, r] = W[r, c]
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ Wt) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows && c0 + tx < cols) t[i][tx] = W[(size_t)(r0 + i) * cols + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) Wt[(size_t)(c0 + i) * rows + r0 + tx] = t[tx][i];
}

// dst[c, m] = src[m, c] for m < rows, 0 for rows <= m < mpad   (dst [cols, mpad]): operands of the TN weight-gradient GEMM
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols, int mpad,
                                                            float* __restrict__ dst) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) t[i][tx] = (r0 + i < rows && c0 + tx < cols) ? src[(size_t)(r0 + i) * ld + c0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < mpad) dst[(size_t)(c0 + i) * mpad + r0 + tx] = t[tx][i];
}

// dW[N,K] += dZ[M,N]^T X[M,K] on the tensor cores (3xTF32): the contraction runs over the ROWS, so both operands are
// transposed into `tw` first ([N, Mpad] and [K, Mpad], rows padded with zeros to a multiple of 32, DW_ROWS rows at a
// time) and the NT kernel runs with its K range split over the CTAs, adding into dW with atomics.  The FFMA split-K
// kernel it replaces took 25-48 us per 2048 x 256 x 256 layer and 140 us per 1992 x 1024 x 1024 layer of the video net
// - and competed for SMs with the latency-bound dX chain.  Returns 1 in *launched when it ran.
// rows per pass: as many as a 64 MB scratch area holds for the two transposed operands ([N + K, rows] fp32), between 4096
// and 32768 - the fp32 trunk feeds 32768-point chunks, and eight 4096-row passes per chunk were launch-latency-bound
constexpr int DW_ROWS = 4096;                      // (the separable path sizes its scratch with this minimum)
static int dw_rows(int n_plus_k) {
  const int64_t r = ((int64_t)16 << 20) / max(n_plus_k, 1);
  return (int)min((int64_t)32768, max((int64_t)DW_ROWS, r / 32 * 32));
}
static int dw_tensor(const float* dZ, int64_t lddz, const float* X, int64_t ldx, int M, int N, int K, float* dW, float* tw,
                     size_t tw_floats, cudaStream_t st, int* launched) {
  *launched = 0;
  const char* e = getenv("COORDX_FP32_GEMM");
  if (!tw || (e && strcmp(e, "ffma") == 0) || N < 64 || K < 64 || M < 128) return 0;
  // the largest pass the caller's scratch area allows (it was sized with dw_rows(2 * max width) or DW_ROWS)
  int pass = dw_rows(N + K);
  while (pass > DW_ROWS && tw_floats < (size_t)(N + K) * ((min(M, pass) + 31) / 32 * 32)) pass /= 2;
  const int rows_max = min(M, pass), pad_max = (rows_max + 31) / 32 * 32;
  if (tw_floats < (size_t)(N + K) * pad_max) return 0;
  float* tz = tw;
  float* tx = tw + (size_t)N * pad_max;
  for (int m0 = 0; m0 < M; m0 += pass) {
    const int rows = min(pass, M - m0), mpad = (rows + 31) / 32 * 32;
    transpose_pad_kernel<<<dim3((unsigned)cdiv(N, 32), (unsigned)(mpad / 32)), 256, 0, st>>>(dZ + (size_t)m0 * lddz, lddz, rows, N, mpad, tz);
    CX_LAUNCH_CHECK();
    transpose_pad_kernel<<<dim3((unsigned)cdiv(K, 32), (unsigned)(mpad / 32)), 256, 0, st>>>(X + (size_t)m0 * ldx, ldx, rows, K, mpad, tx);
    CX_LAUNCH_CHECK();
    g3::Args t{};
    t.A = tz; t.lda = mpad; t.B = tx; t.ldb = mpad; t.M = N; t.N = K; t.K = mpad;
    t.Y = dW; t.ldc = K; t.cs = 1; t.act = -1; t.atomic = 1;
    const bool wide = K >= 256 && cdiv(N, 128) * cdiv(K, 256) >= 32;
    const int64_t tiles = cdiv(N, 128) * cdiv(K, wide ? 256 : 64);
    const int nk = mpad / 32;
    int splits = (int)max((int64_t)1, min((int64_t)nk, cdiv(148, tiles)));
    const int per = (nk + splits - 1) / splits;
    t.ksplits = (nk + per - 1) / per;                  // every split non-empty
    CX_TRY(wide ? g3::launch_bn<256>(t, st) : g3::launch_bn<64>(t, st));
  }
  *launched = 1;
  return 0;
}

// Y[M,N] += A[M,K] W[K,N] on the tensor cores (3xTF32, atomic epilogue): W is given TRANSPOSED (`wt` [N,K], made once by
// the caller with transpose_kernel), so this is the NT kernel.  Returns 1 in *launched when the shape pays.
static int nn_tensor_acc(const float* A, int64_t lda, const float* wt, int M, int N, int K, float* Y, int64_t ldy, cudaStream_t st,
                         int* launched) {
  *launched = 0;
  const char* e = getenv("COORDX_FP32_GEMM");
  if (!wt || (e && strcmp(e, "ffma") == 0) || M < 96 || N < 64 || K < 128 || K % 32 != 0 || (lda & 3)) return 0;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(wt)) & 15) return 0;
  g3::Args t{};
  t.A = A; t.lda = lda; t.B = wt; t.ldb = K; t.M = M; t.N = N; t.K = K;
  t.Y = Y; t.ldc = ldy; t.cs = 1; t.act = -1; t.atomic = 1; t.ksplits = 1;
  const bool wide = N >= 256 && cdiv(M, 128) * cdiv(N, 256) >= 32;
  CX_TRY(wide ? g3::launch_bn<256>(t, st) : g3::launch_bn<64>(t, st));
  *launched = 1;
  return 0;
}

// db[n] += scale * sum_m dZ[m,n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dZ, int64_t ld, int M, int N, float scale, float* __restrict__ db, int rows_per_block) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = r0 + (threadIdx.x >> 5); m < r1; m += 8) s += dZ[(int64_t)m * ld + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    if (n < N) atomicAdd(&db[n], t * scale);
  }
}

// narrow matrices (N <= 8, contiguous rows): every thread walks the flat array with a stride that is a
// multiple of N, so its column is fixed; one shared-memory atomic per thread, one global atomic per column per CTA
__global__ void __launch_bounds__(256)
colsum_narrow_kernel(const float* __restrict__ dZ, int64_t total, int N, float scale, float* __restrict__ db) {
  __shared__ float red[8];
  if (threadIdx.x < 8) red[threadIdx.x] = 0.f;
  __syncthreads();
  const int active = (256 / N) * N;               // threads per block that take part: a multiple of N
  if ((int)threadIdx.x < active) {
    const int64_t stride = (int64_t)gridDim.x * active;       // multiple of N: a thread's column never changes
    const int64_t i0 = (int64_t)blockIdx.x * active + threadIdx.x;
    float s = 0.f;
    for (int64_t i = i0; i < total; i += stride) s += dZ[i];
    if ((N & (N - 1)) == 0) {
      // N = 1, 2, 4, 8: all 256 threads take part and a thread's column is lane % N - fold the warp with
      // shuffles first (256 same-address shared-memory atomics per CTA cost 19 us on a 4 M-element dout)
      for (int o = 16; o >= N; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((int)(threadIdx.x & 31) < N) atomicAdd(&red[threadIdx.x & 31], s);
    } else {
      atomicAdd(&red[i0 % N], s);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < N) atomicAdd(&db[threadIdx.x], red[threadIdx.x] * scale);
}

static int bias_backward(const float* dZ, int64_t ld, int M, int N, float scale, float* db, cudaStream_t st) {
  if (M <= 0) return 0;
  if (N <= 8 && ld == N) {
    const int64_t total = (int64_t)M * N;
    const int blocks = (int)min((int64_t)148 * 8, cdiv(total, 256 * 4));
    colsum_narrow_kernel<<<max(blocks, 1), 256, 0, st>>>(dZ, total, N, scale, db);
    CX_LAUNCH_CHECK();
    return 0;
  }
  // ~4 CTAs per SM: rows per block from the number of 32-column strips
  const int64_t strips = cdiv(N, 32);
  int rpb = (int)max((int64_t)32, cdiv((int64_t)M * strips, (int64_t)148 * 4));
  dim3 grid((unsigned)strips, (unsigned)cdiv(M, rpb));
  colsum_kernel<<<grid, 256, 0, st>>>(dZ, ld, M, N, scale, db, rpb);
  CX_LAUNCH_CHECK();
  return 0;
}

}  // namespace coordx
