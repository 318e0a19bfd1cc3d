// Shared helpers for the coordx_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/coordx_b200.h"

namespace coordx {

extern thread_local char g_err[512];
extern unsigned long long g_launches;     // kernels launched by this library (bench.py's gpu_launches)

inline int set_err(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

#define CX_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      snprintf(coordx::g_err, sizeof(coordx::g_err), "%s failed: %s (%s:%d)", #expr,       \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                \
      return COORDX_ERR_CUDA;                                                              \
    }                                                                                      \
  } while (0)

#define CX_LAUNCH_CHECK()          \
  do {                             \
    ++coordx::g_launches;          \
    CX_CUDA(cudaGetLastError());   \
  } while (0)

#define CX_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// cudaFuncSetAttribute is per DEVICE: one flag per (kernel instantiation, device ordinal).  Use as
//   static AttrOnce once;  CX_TRY(once.set(kernel, bytes));
struct AttrOnce {
  bool done[64] = {};
  template <typename K>
  int set(K kern, int bytes) {
    int dev = 0;
    CX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
      CX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return 0;
  }
};

// SM count of the CURRENT device (cached per device ordinal)
static inline int num_sms() {
  static int cache[64] = {};
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* get(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (base == nullptr) { off += bytes; return nullptr; }     // sizing pass
    if (off + bytes > cap) { off += bytes; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

// accurate sin / cos as real function calls: ~60 instructions each, and the register-tiled epilogues would
// otherwise inline dozens of copies (a 160 KB kernel spends 40 % of its stalls on instruction-cache misses)
static __device__ __noinline__ float sin_accurate(float x) { return sinf(x); }
static __device__ __noinline__ float cos_accurate(float x) { return cosf(x); }
// Inlinable accurate sin / cos: the libdevice sinf / cosf carry their Payne-Hanek slow path (|x| > 105615, ~120
// instructions and a local-memory array) into every inlined copy, which is what made unrolled epilogues too large.
// Here only the fast path is inline - 3-term Cody-Waite reduction by pi/2 and the single-precision minimax
// polynomials on [-pi/4, pi/4] (max error ~1.5 ulp, the same class as sinf) - and the rare huge argument takes the
// out-of-line libdevice call.  25 instructions per value, independent chains interleave freely.
__device__ __forceinline__ void sincos_reduce(float x, float& r, int& q) {
  const float n = rintf(x * 0.63661977236758134f);
  r = fmaf(n, -1.57079601287841796875f, x);            // pi/2 split into three parts (24 + 24 + rest bits)
  r = fmaf(n, -3.1391647326017846353352069854736e-7f, r);
  r = fmaf(n, -5.3903025299577648829435417064e-15f, r);
  q = (int)n;
}
__device__ __forceinline__ float sin_poly(float r) {
  const float r2 = r * r;
  float p = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  p = fmaf(p, r2, -1.6666654611e-1f);
  return fmaf(p * r2, r, r);
}
__device__ __forceinline__ float cos_poly(float r) {
  const float r2 = r * r;
  float p = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  p = fmaf(p, r2, 4.166664568298827e-2f);
  return fmaf(p * r2, r2, fmaf(r2, -0.5f, 1.0f));
}
__device__ __forceinline__ float sin_inl(float x) {
  if (fabsf(x) > 48000.f) return sin_accurate(x);
  float r;
  int q;
  sincos_reduce(x, r, q);
  const float v = (q & 1) ? cos_poly(r) : sin_poly(r);
  return (q & 2) ? -v : v;
}
__device__ __forceinline__ float cos_inl(float x) {
  if (fabsf(x) > 48000.f) return cos_accurate(x);
  float r;
  int q;
  sincos_reduce(x, r, q);
  const float v = (q & 1) ? sin_poly(r) : cos_poly(r);
  return ((q + 1) & 2) ? -v : v;
}

// the remaining entries of the reference's nonlinearity table (modules.py:170-176: nn.Sigmoid, nn.Tanh, nn.SELU,
// nn.Softplus(beta 1, threshold 20), nn.ELU(alpha 1)): FP32 path only, out of line so that the sine / relu code
// of every kernel stays as it was
static __device__ __noinline__ float act_other_fwd(int act, float z) {
  switch (act) {
    case COORDX_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case COORDX_ACT_TANH: return tanhf(z);
    case COORDX_ACT_SELU: return 1.0507009873554804934193349852946f * (z > 0.f ? z : 1.6732632423543772848170429916717f * expm1f(z));
    case COORDX_ACT_SOFTPLUS: return z > 20.f ? z : log1pf(expf(z));
    case COORDX_ACT_ELU: return z > 0.f ? z : expm1f(z);
  }
  return z;
}
static __device__ __noinline__ float act_other_grad(int act, float z) {
  switch (act) {
    case COORDX_ACT_SIGMOID: {
      const float s = 1.f / (1.f + expf(-z));
      return s * (1.f - s);
    }
    case COORDX_ACT_TANH: {
      const float t = tanhf(z);
      return 1.f - t * t;
    }
    case COORDX_ACT_SELU: return 1.0507009873554804934193349852946f * (z > 0.f ? 1.f : 1.6732632423543772848170429916717f * expf(z));
    case COORDX_ACT_SOFTPLUS: return z > 20.f ? 1.f : 1.f / (1.f + expf(-z));
    case COORDX_ACT_ELU: return z > 0.f ? 1.f : expf(z);
  }
  return 1.f;
}

__device__ __forceinline__ float act_fwd(int act, float S, float z) {
  if (act == COORDX_ACT_SINE) return sin_inl(S * z);
  if (act == COORDX_ACT_RELU) return fmaxf(z, 0.f);
  return act_other_fwd(act, z);
}
__device__ __forceinline__ float act_grad(int act, float S, float z) {
  if (act == COORDX_ACT_SINE) return S * cos_inl(S * z);
  if (act == COORDX_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return act_other_grad(act, z);
}

}  // namespace coordx
