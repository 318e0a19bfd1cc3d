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
__device__ __forceinline__ float act_fwd(int act, float S, float z) {
  return act == COORDX_ACT_SINE ? sin_accurate(S * z) : fmaxf(z, 0.f);
}
__device__ __forceinline__ float act_grad(int act, float S, float z) {
  return act == COORDX_ACT_SINE ? S * cos_accurate(S * z) : (z > 0.f ? 1.f : 0.f);
}

}  // namespace coordx
