// Shared device/host helpers for the flashmd B200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/fmd_b200.h"

#define FMD_PI_F 3.14159265358979323846f

extern "C" void fmd_set_error(const char* msg);

#define FMD_FAIL(msg)       \
  do {                      \
    fmd_set_error(msg);     \
    return FMD_ERR_INVALID; \
  } while (0)

#define FMD_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) FMD_FAIL(msg); \
  } while (0)

// Launch check without synchronising (the caller owns the stream).
#define FMD_CHECK_LAUNCH()                                  \
  do {                                                      \
    cudaError_t _e = cudaGetLastError();                    \
    if (_e != cudaSuccess) {                                \
      fmd_set_error(cudaGetErrorString(_e));                \
      return FMD_ERR_CUDA;                                  \
    }                                                       \
  } while (0)

#define FMD_CUDA(call)                       \
  do {                                       \
    cudaError_t _e = (call);                 \
    if (_e != cudaSuccess) {                 \
      fmd_set_error(cudaGetErrorString(_e)); \
      return FMD_ERR_CUDA;                   \
    }                                        \
  } while (0)

static inline int fmd_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (B200: 148); cached.
int fmd_num_sms();

namespace fmd {

// 0.5*(cos(pi d / rc) + 1) for d < rc, else 0   (reference models/cutoff.py:137-145)
__device__ __forceinline__ float cosine_cutoff(float d, float rc) {
  float c = 0.5f * (cosf(d * FMD_PI_F / rc) + 1.0f);
  return d < rc ? c : 0.0f;
}
// derivative of the above w.r.t. d
__device__ __forceinline__ float cosine_cutoff_grad(float d, float rc) {
  float s = -0.5f * (FMD_PI_F / rc) * sinf(d * FMD_PI_F / rc);
  return d < rc ? s : 0.0f;
}

// tanh via exp with the clamp of the reference's FP16 kernels (kernels/cfconv_kernels.py:449-454)
__device__ __forceinline__ float tanh_clamped(float x) {
  float xc = fminf(fmaxf(x, -10.0f), 10.0f);
  float e = __expf(2.0f * xc);
  return (e - 1.0f) / (e + 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// load 4 consecutive elements as float4 (16B for f32, 8B for f16); pointer must be aligned accordingly
__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4(const __half* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  __half2 a = *reinterpret_cast<__half2*>(&u.x);
  __half2 b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
// streaming variants (read-once data: bypass L1 allocation)
__device__ __forceinline__ float4 load4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 load4_stream(const __half* p) {
  uint2 u;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
  __half2 a = *reinterpret_cast<__half2*>(&u.x);
  __half2 b = *reinterpret_cast<__half2*>(&u.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

}  // namespace fmd
