// Probe (tools only): per-SM throughput of the instructions the fused filter kernels lean on (sm_100a).
// One CTA of 16 warps per SM, ITER x UNROLL independent ops per thread; reports warp-instructions / cycle / SM.
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int ITER = 2000, U = 8;

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed) {
  float a[U];
  uint32_t h[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { a[u] = seed + 0.001f * (threadIdx.x + u); h[u] = 0x3c003800u + threadIdx.x + u; }
  __shared__ unsigned short sm[4096];
  sm[threadIdx.x] = (unsigned short)threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (OP == 0) { a[u] = fmaf(a[u], 1.0001f, 0.5f); }                                                    // FFMA
      if (OP == 1) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[u])); }                                 // MUFU.TANH
      if (OP == 2) { asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[u])); }                               // MUFU.TANH.F16x2
      if (OP == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u])); }                              // MUFU.EX2
      if (OP == 4) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[u]) : "f"(a[u]), "f"(a[(u + 1) % U])); a[u] += __uint_as_float(h[u]); }  // F2FP + FADD
      if (OP == 5) { asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.f32.f16 %0, lo;\n\t}" : "=f"(a[u]) : "r"(h[u])); h[u] += __float_as_uint(a[u]); }  // h2f + IADD
      if (OP == 6) { asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.f16 %0, lo, hi, %0;\n\t}" : "+f"(a[u]) : "r"(h[u])); }  // FHFMA
      if (OP == 7) { unsigned short v = *reinterpret_cast<volatile unsigned short*>(&sm[(threadIdx.x + u * 32 + it) & 4095]); h[u] += v; }   // LDS.U16 + IADD
      if (OP == 8) { asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(h[u]) : "r"(0x3c003c00u)); }            // HMUL2
      if (OP == 9) { a[u] = a[u] + __uint_as_float(h[u]); }                                                  // FADD (for subtracting the helper op)
      if (OP == 10) { h[u] += __float_as_uint(a[u]); }                                                       // IADD
      if (OP == 11) { asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(*reinterpret_cast<unsigned short*>(&h[u])) : "f"(a[u])); a[u] += __uint_as_float(h[u]); }  // F2F scalar + FADD
      if (OP == 12) { h[u] = ((__float_as_uint(a[u]) + 0x1000u) >> 13) + h[u]; }                             // integer rounding pieces (IADD, SHF, IADD)
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) s += a[u] + __uint_as_float(h[u]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 512 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  k<OP><<<148, 512>>>(out, cyc, 0.3f);
  CK(cudaDeviceSynchronize());
  k<OP><<<148, 512>>>(out, cyc, 0.3f);
  CK(cudaDeviceSynchronize());
  long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  const double winst = 16.0 * ITER * U;
  printf("%-44s %7.3f warp-inst/clk/SM   (%5.1f cycles per warp-inst per SMSP)\n", name, winst / mean, mean / (winst / 4));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA");
  run<9>("FADD (helper)");
  run<10>("IADD (helper)");
  run<1>("MUFU.TANH f32");
  run<2>("MUFU.TANH f16x2");
  run<3>("MUFU.EX2");
  run<4>("F2FP pack f32,f32->f16x2 (+FADD)");
  run<11>("F2F f32->f16 scalar (+FADD)");
  run<5>("cvt f16->f32 (+IADD)");
  run<6>("FHFMA (fma.rn.f32.f16)");
  run<7>("LDS.U16 (+IADD)");
  run<8>("HMUL2");
  run<12>("int round pieces (IADD+SHF+IADD)");
  return 0;
}
