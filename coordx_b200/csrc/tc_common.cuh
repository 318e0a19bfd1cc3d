// Shared pieces of the tcgen05 trunk kernels (forward: trunk_fwd_tc.cu, backward: trunk_bwd_tc.cu).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace coordx {
namespace tc {

// Pipeline tracing (SM-clock stamps written by CTA 0 of the trunk kernels; scripts/*_trace.py) is a BUILD option:
// the run-time "is a trace buffer set?" checks alone cost the dW pass 6 % (963 -> 909 us per 2048^2,
// profiles/r02i_trace_cost.txt).  scripts/build_trace_lib.sh builds coordx_b200/_build/libcoordx_trace.so with
// -DCOORDX_TRACE=1; the product library refuses coordx_profile_set_trace / _set_bwd_trace with a buffer.
#ifndef COORDX_TRACE
#define COORDX_TRACE 0
#endif
constexpr int HD = 256;
constexpr int TILE_M = 128;
constexpr int KCHUNK = 64;                          // one SWIZZLE_128B row of 16-bit elements
constexpr int NKC = HD / KCHUNK;                    // 4
constexpr int MAX_O = 4;

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
  static __device__ __forceinline__ float unpack(uint32_t v, int hi) {       // element 0 (low) or 1 (high) as fp32
    return __uint_as_float(hi ? (v & 0xffff0000u) : (v << 16));
  }
};
template <>
struct Pack<__half> {
  using V2 = __half2;
  static constexpr uint32_t kFmt = 0;
  // fp16 saturates instead of overflowing (cvt.satfinite): an unnormalised ReLU feature beyond 65504 becomes the
  // largest finite value, not an inf that the product fusion and the MMA would turn into NaN
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    uint32_t v;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(b), "f"(a));
    return v;
  }
  static __device__ __forceinline__ float unpack(uint32_t v, int hi) {
    const __half2 h = *reinterpret_cast<const __half2*>(&v);
    return hi ? __high2float(h) : __low2float(h);
  }
};

template <typename T, int OP>
__device__ __forceinline__ uint32_t fuse2(uint32_t x, uint32_t y) {
  using V2 = typename Pack<T>::V2;
  V2 a = *reinterpret_cast<V2*>(&x), b = *reinterpret_cast<V2*>(&y);
  V2 r = OP == COORDX_FUSE_PROD ? __hmul2(a, b) : __hadd2(a, b);
  if (Pack<T>::kFmt == 0) {     // fp16: clamp the fused value to the finite range as well (bf16 has fp32's exponent)
    const uint32_t hi = 0x7bff7bffu, lo = 0xfbfffbffu;      // +-65504 in both halves
    asm("min.f16x2 %0, %0, %1;" : "+r"(*reinterpret_cast<uint32_t*>(&r)) : "r"(hi));
    asm("max.f16x2 %0, %0, %1;" : "+r"(*reinterpret_cast<uint32_t*>(&r)) : "r"(lo));
  }
  return *reinterpret_cast<uint32_t*>(&r);
}

// ---- entry points of the two tcgen05 translation units (called from coordx_abi.cu)
bool tc_supported(const coordx_net* net, int precision);
size_t fwd_workspace_bytes(int64_t n_inner);
struct FwdLoss {            // loss fused into the last layer (coordx_trunk_forward_loss)
  int kind;
  const float* gt;
  float inv_count;
  float* loss_acc;
  float* dout;
};
int trunk_forward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t row_begin,
                     int64_t row_end, float* out, int precision, void* ws, size_t ws_bytes, cudaStream_t st,
                     const FwdLoss* loss = nullptr);
void profile_enable(bool on);
int profile_set_trace(void* dev_buf);
int profile_set_bwd_trace(void* dev_buf, int pass);
int profile_last_ms(float* ms_host);

size_t bwd_workspace_bytes(const coordx_net* net, const int64_t* n, int64_t r0, int64_t r1, bool minimal);
int trunk_backward_tc(const coordx_net* net, const float* const* feats, const int64_t* n, int64_t r0, int64_t r1,
                      const float* dout, coordx_grads* grads, float* const* dfeat, int precision, void* ws,
                      size_t ws_bytes, cudaStream_t st);

}  // namespace tc
}  // namespace coordx
