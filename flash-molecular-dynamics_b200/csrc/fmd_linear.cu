// Generic dense layer  Y = epi(pro(X) @ W + b) [* (1 - aux^2)] [+ res]  with TRUE fp32 FMA
// accumulation (no TF32): the fp32 parity path (1e-5 vs the reference's --disable_optim path) and
// the building block of every operator-level drop-in (fused_tanh_linear, fused_linear_tanh_fp16,
// linear_fp16*, their backward GEMMs, nn.Linear).  128x128x16 CTA tile, 256 threads, 8x8 register
// micro-tile split 4+4 in both directions so shared-memory reads are conflict-free float4s.
// The W16A16 filter network of the fused step runs on tensor cores instead (fmd_filter_tc.cu).
#include "fmd_common.cuh"

using namespace fmd;

namespace {

constexpr int BM = 128, BN = 128, BK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == FMD_ACT_TANH) return tanhf(v);
  if (act == FMD_ACT_TANH_CLAMPED) return tanh_clamped(v);
  return v;
}

template <typename TX, typename TW, typename TY>
__global__ void __launch_bounds__(256)
linear_kernel(const TX* __restrict__ X, const TW* __restrict__ W, const TW* __restrict__ bias, TY* __restrict__ Y,
              int M, int N, int K, const int32_t* __restrict__ m_dev, int pro_act, int x_round_f16, int epi_act,
              const void* __restrict__ aux, int auxdt, const float* __restrict__ res) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  if (m_dev) M = min(M, *m_dev);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= M) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // A loader: row = tid % 128, k segment = (tid / 128) * 8
  const int a_row = tid & 127, a_k0 = (tid >> 7) * 8;
  // B loader: k = tid / 16, n segment = (tid % 16) * 8
  const int b_k = tid >> 4, b_n0 = (tid & 15) * 8;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- load A tile (transposed into As[k][m])
    {
      const int gm = m0 + a_row;
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int gk = k0 + a_k0 + u;
        float f = 0.f;
        if (gm < M && gk < K) {
          f = to_f32<TX>(X[(size_t)gm * K + gk]);
          if (x_round_f16) f = __half2float(__float2half_rn(f));
          f = apply_act(f, pro_act);
        }
        v[u] = f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) As[a_k0 + u][a_row] = v[u];
    }
    // ---- load B tile
    {
      const int gk = k0 + b_k;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int gn = n0 + b_n0 + u;
        float f = 0.f;
        if (gk < K && gn < N) f = to_f32<TW>(W[(size_t)gk * N + gn]);
        Bs[b_k][b_n0 + u] = f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += to_f32<TW>(bias[gn]);
      v = apply_act(v, epi_act);
      const size_t o = (size_t)gm * N + gn;
      if (aux) {
        const float t = auxdt == FMD_F16 ? __half2float(reinterpret_cast<const __half*>(aux)[o])
                                         : reinterpret_cast<const float*>(aux)[o];
        v *= (1.0f - t * t);
      }
      if (res) v += res[o];
      Y[o] = from_f32<TY>(v);
    }
  }
}

template <typename TX, typename TW, typename TY>
void launch(const void* X, const void* W, const void* bias, void* Y, int M, int N, int K, const int32_t* m_dev,
            int pro_act, int x_round, int epi_act, const void* aux, int auxdt, const float* res, cudaStream_t st) {
  dim3 grid(fmd_div_up(M, BM), fmd_div_up(N, BN));
  linear_kernel<TX, TW, TY><<<grid, 256, 0, st>>>((const TX*)X, (const TW*)W, (const TW*)bias, (TY*)Y, M, N, K, m_dev,
                                                  pro_act, x_round, epi_act, aux, auxdt, res);
}

}  // namespace

extern "C" int fmd_linear(const void* X, int xdt, const void* W, int wdt, const void* bias, void* Y, int ydt, int M,
                          int N, int K, const int32_t* m_dev, int pro_act, int x_round_f16, int epi_act,
                          const void* aux, int auxdt, const float* res, void* stream) {
  FMD_REQUIRE(X && W && Y && M >= 0 && N > 0 && K > 0, "fmd_linear: bad arguments");
  FMD_REQUIRE((xdt | wdt | ydt | auxdt) >= 0 && xdt <= 1 && wdt <= 1 && ydt <= 1 && auxdt <= 1, "fmd_linear: bad dtype");
  if (M == 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = xdt * 4 + wdt * 2 + ydt;
#define FMD_L(TX, TW, TY) \
  launch<TX, TW, TY>(X, W, bias, Y, M, N, K, m_dev, pro_act, x_round_f16, epi_act, aux, auxdt, res, st)
  switch (key) {
    case 0: FMD_L(float, float, float); break;
    case 1: FMD_L(float, float, __half); break;
    case 2: FMD_L(float, __half, float); break;
    case 3: FMD_L(float, __half, __half); break;
    case 4: FMD_L(__half, float, float); break;
    case 5: FMD_L(__half, float, __half); break;
    case 6: FMD_L(__half, __half, float); break;
    case 7: FMD_L(__half, __half, __half); break;
    default: FMD_FAIL("fmd_linear: bad dtype combination");
  }
#undef FMD_L
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
