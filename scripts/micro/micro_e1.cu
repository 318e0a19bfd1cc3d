// Micro-benchmark of the dW pass' E1 arithmetic (sin + cos of 16 values per thread per 64-point sub-tile) in
// isolation: 16 warps per SM, 64 registers per thread, no tensor memory / barriers.  Variants: how many of the 8
// (even, odd) pairs take the FMA-pipe polynomial instead of the MUFU, packed fp32x2 vs scalar arithmetic with
// immediate coefficients.  Prints SM clocks per sub-tile (per warp pass over its 16 values).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o micro_e1 micro_e1.cu && ./micro_e1
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../coordx_b200/csrc/ptx_sm100.cuh"
using namespace coordx;

__device__ __forceinline__ uint64_t f2x(float v) { return ptx::pack_f32x2(v, v); }
__device__ __forceinline__ void sincos_rev_x2(uint64_t t2, uint64_t& sn, uint64_t& cs) {
  const uint64_t m2 = ptx::add_f32x2(t2, f2x(12582912.f));
  const uint64_t n2 = ptx::add_f32x2(m2, f2x(-12582912.f));
  const uint64_t rr = ptx::fma_f32x2(n2, f2x(-1.f), t2);
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
// scalar: coefficients are immediates of the FFMAs
__device__ __forceinline__ void sincos_rev(float t, float& sn, float& cs) {
  const float m = t + 12582912.f;
  const float n = m - 12582912.f;
  const float r = t - n;
  const float q = r * r;
  float ps = fmaf(q, -56.08980178833008f, 77.93164825439453f);
  ps = fmaf(ps, q, -41.093868255615234f);
  ps = fmaf(ps, q, 6.27863883972168f);
  sn = ps * r;
  float pc = fmaf(q, 45.64854431152344f, -82.40443420410156f);
  pc = fmaf(pc, q, 64.67355346679688f);
  pc = fmaf(pc, q, -19.731046676635742f);
  cs = fmaf(pc, q, 0.9999598264694214f);
}

template <int NPOLY, bool PACKED>
__global__ void __launch_bounds__(512, 2) e1_kernel(const float* __restrict__ zin, float* __restrict__ out, int iters, long long* clk, float S, float bias) {
  float z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] = zin[(threadIdx.x * 16 + i) & 4095];
  float d[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = zin[(i * 7 + blockIdx.x) & 4095];
  uint64_t db2 = 0, dwf2 = 0;
  float dbs = 0.f, dwfs = 0.f;
  uint32_t x = 0;
  const uint64_t S2 = f2x(S), B2 = f2x(bias), S2r = f2x(S * 0.15915494309189535f), B2r = f2x(bias * 0.15915494309189535f), W2 = f2x(0.37f);
  const float Sr = S * 0.15915494309189535f, Br = bias * 0.15915494309189535f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[8];
#pragma unroll
    for (int e2 = 0; e2 < 8; ++e2) {
      const float pert = __int_as_float(x & 0x3f000000u);      // every value depends on the previous iteration: nothing hoists
      const float z0 = z[2 * e2] + pert, z1 = z[2 * e2 + 1] - pert;
      // which pairs take the polynomial: spread them over the 8 pairs
      const bool poly = ((e2 * NPOLY) / 8) != (((e2 + 1) * NPOLY) / 8);
      if (PACKED) {
        const uint64_t zz = ptx::pack_f32x2(z0, z1);
        const uint64_t d2 = ptx::pack_f32x2(d[2 * e2], d[2 * e2 + 1]);
        uint64_t sn, cs;
        if (poly) {
          sincos_rev_x2(ptx::fma_f32x2(zz, S2r, B2r), sn, cs);
        } else {
          float th0, th1;
          ptx::unpack_f32x2(ptx::fma_f32x2(zz, S2, B2), th0, th1);
          sn = ptx::pack_f32x2(__sinf(th0), __sinf(th1));
          cs = ptx::pack_f32x2(__cosf(th0), __cosf(th1));
        }
        const uint64_t dz = ptx::mul_f32x2(ptx::mul_f32x2(d2, W2), cs);
        db2 = ptx::add_f32x2(db2, dz);
        dwf2 = ptx::fma_f32x2(d2, sn, dwf2);
        float dz0, dz1;
        ptx::unpack_f32x2(dz, dz0, dz1);
        __nv_bfloat162 v = __floats2bfloat162_rn(dz0, dz1);
        pk[e2] = *reinterpret_cast<uint32_t*>(&v);
      } else {
        float dzv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float zv = u ? z1 : z0;
          float sn, cs;
          if (poly) {
            sincos_rev(fmaf(zv, Sr, Br), sn, cs);
          } else {
            const float th = fmaf(zv, S, bias);
            sn = __sinf(th);
            cs = __cosf(th);
          }
          const float gw = d[2 * e2 + u] * 0.37f;
          dzv[u] = gw * cs;
          dbs += dzv[u];
          dwfs = fmaf(d[2 * e2 + u], sn, dwfs);
        }
        __nv_bfloat162 v = __floats2bfloat162_rn(dzv[0], dzv[1]);
        pk[e2] = *reinterpret_cast<uint32_t*>(&v);
      }
    }
#pragma unroll
    for (int e2 = 0; e2 < 8; ++e2) x ^= pk[e2];
  }
  const long long t1 = clock64();
  float a0, a1;
  ptx::unpack_f32x2(db2, a0, a1);
  float b0, b1;
  ptx::unpack_f32x2(dwf2, b0, b1);
  out[blockIdx.x * 512 + threadIdx.x] = a0 + a1 + b0 + b1 + dbs + dwfs + __uint_as_float(x);
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int NPOLY, bool PACKED>
void run(const float* zin, float* out, long long* clk) {
  const int iters = 2000;
  e1_kernel<NPOLY, PACKED><<<148, 512>>>(zin, out, 10, clk, 30.f, 0.1f);
  e1_kernel<NPOLY, PACKED><<<148, 512>>>(zin, out, iters, clk, 30.f, 0.1f);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("NPOLY %d/8 %s: %.1f clk per sub-tile pass (16 values/thread, 4 warps/scheduler)  [%s]\n", NPOLY, PACKED ? "packed" : "scalar",
         (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* zin; float* out; long long* clk;
  cudaMalloc(&zin, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) >> 8) / 16777216.f - 0.5f;
  cudaMemcpy(zin, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, true>(zin, out, clk); run<2, true>(zin, out, clk); run<3, true>(zin, out, clk); run<4, true>(zin, out, clk);
  run<5, true>(zin, out, clk); run<6, true>(zin, out, clk); run<8, true>(zin, out, clk);
  run<0, false>(zin, out, clk); run<2, false>(zin, out, clk); run<3, false>(zin, out, clk); run<4, false>(zin, out, clk);
  run<5, false>(zin, out, clk); run<6, false>(zin, out, clk); run<8, false>(zin, out, clk);
  return 0;
}
