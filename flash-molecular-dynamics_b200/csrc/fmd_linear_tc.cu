// Node-level dense layers on the sm_100a tensor cores (tcgen05.mma kind::tf32, fp32 TMEM accumulator):
//   Y[M,N] = epi( pro(X)[M,K] @ W[K,N] + bias ) [* (1 - aux^2)] [+ res]        K, N in {32..128}
// Same contract as fmd_linear; TF32 operands are what the reference's GPU path uses for these layers
// (nn.Linear under torch.set_float32_matmul_precision("high"), scripts/nvt_langevin.py:38; tl.dot default
// in fused_tanh_linear, kernels/cfconv_kernels.py:1758-1843).  fp16 inputs are exact in TF32, so the
// W16A16 output network (fp16 operands, fp32 accumulate: models/gptq.py:266-306) is reproduced exactly.
// The fp32 parity path (1e-5) never comes here: it stays on the true-fp32 FMA kernel of fmd_linear.cu.
//
// One persistent CTA per SM, two 128-thread groups that run the same code on two row tiles concurrently (their
// load / MMA / epilogue phases overlap; the weights are shared).  The weight matrix is staged once per CTA (transposed on the
// fly into the K-major B operand); per 128-row tile the CTA stages pro(X) as the K-major A operand (round-to-nearest TF32),
// one thread issues K/8 MMAs (both operands K-major), and thread r drains row r of the accumulator through the epilogue.
#include "fmd_tc.cuh"

using namespace fmd;
using namespace fmd::tc;

namespace {

constexpr int LT_TILE = 128;
constexpr int LT_GROUPS = 2;                  // two 128-thread groups per CTA work on two row tiles concurrently
constexpr int LT_THREADS = LT_TILE * LT_GROUPS;
constexpr uint32_t LO_A = 0;                  // per group: up to 4 K-blocks x [128 rows][128 B] = 64 KB
constexpr uint32_t LO_B = 2 * 64 * 1024;      // shared: up to 4 K-blocks x [128 n][128 B]     = 64 KB
constexpr uint32_t LO_BIAS = 3 * 64 * 1024;   // 2 x 128 floats (the chain kernel double-buffers the bias row)
constexpr uint32_t LO_BAR = LO_BIAS + 1024;
constexpr uint32_t LT_SMEM = LO_BAR + 32;
constexpr uint32_t LT_SMEM_ALLOC = LT_SMEM + 1024;

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// round-to-nearest (ties away) to TF32 with two full-rate integer ops (cvt.rna.tf32.f32 is a slow-pipe
// instruction and was the top stall of this kernel); identical for finite inputs
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// both tanh flavours map to the hardware tanh.approx.f32 (|err| <= 2^-11, the TF32 operand rounding level of
// this kernel); the exact tanhf / clamped-exp variants live in the fp32 kernel (fmd_linear.cu)
__device__ __forceinline__ float act(float v, int a) { return a == FMD_ACT_NONE ? v : tanh_approx(v); }

template <typename TX>
__device__ __forceinline__ float4 load_x4(const TX* p);
template <>
__device__ __forceinline__ float4 load_x4<float>(const float* p) { return load4(p); }
template <>
__device__ __forceinline__ float4 load_x4<__half>(const __half* p) { return load4(p); }

template <typename TX, typename TW, typename TY>
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const TX* __restrict__ X, const TW* __restrict__ W, const TW* __restrict__ bias, TY* __restrict__ Y,
                 int M, int N, int K, const int32_t* __restrict__ m_dev, int pro_act, int x_round_f16, int epi_act,
                 const void* __restrict__ aux, int auxdt, const float* __restrict__ res, int w_nk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid_all = threadIdx.x, warp_all = tid_all >> 5;
  const int group = tid_all >> 7;               // 0 / 1
  const int tid = tid_all & (LT_TILE - 1);      // thread index inside the group == accumulator row
  const int warp = tid >> 5;                    // == warp_all % 4: the TMEM lane quarter this warp may read
  float* sBias = reinterpret_cast<float*>(smem + LO_BIAS);
  const uint32_t bar = sbase + LO_BAR + 8u * group;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + LO_BAR + 16);
  uint8_t* sA = smem + LO_A + group * (64 * 1024);
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(LT_TILE) : "memory"); };
  if (m_dev) M = min(M, *m_dev);
  const int n_tiles = (M + LT_TILE - 1) / LT_TILE;
  const int kc = K / 4;  // 16-byte chunks per X row

  // ---- weights -> K-major B operand [N rows][K], K-block kb (32 floats) at kb * N * 128 bytes
  if (w_nk) {
    // W given as [N][K] (nn.Linear.weight layout): 16-byte chunks, coalesced, 8 in flight
    const int kc_shift = (K == 128) ? 5 : 4;
    const int total = N << kc_shift;
    for (int base = 0; base < total; base += LT_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * LT_THREADS + tid_all;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) v[u] = load4(W + (size_t)idx * 4);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * LT_THREADS + tid_all;
        const int n = idx >> kc_shift, c = idx & (kc - 1);
        float4 t = v[u];
        t.x = to_tf32(t.x); t.y = to_tf32(t.y); t.z = to_tf32(t.z); t.w = to_tf32(t.w);
        if (idx < total) *reinterpret_cast<float4*>(smem + LO_B + (c >> 3) * (N * 128) + sw128_off(n, c & 7)) = t;
      }
    }
  } else {
    // W given as [K][N]: task = (n, 4 consecutive k), the scalar loads are coalesced across threads (n fastest)
    const int n_shift = (N == 128) ? 7 : 6;
    const int total = (K >> 2) * N;
    for (int base = 0; base < total; base += LT_THREADS * 4) {
      float w[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * LT_THREADS + tid_all;
        const int n = idx & (N - 1), kq = idx >> n_shift;
#pragma unroll
        for (int q = 0; q < 4; ++q) w[u][q] = idx < total ? to_f32<TW>(W[(size_t)(kq * 4 + q) * N + n]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * LT_THREADS + tid_all;
        const int n = idx & (N - 1), kq = idx >> n_shift;
        if (idx < total)
          *reinterpret_cast<float4*>(smem + LO_B + (kq >> 3) * (N * 128) + sw128_off(n, kq & 7)) =
              make_float4(to_tf32(w[u][0]), to_tf32(w[u][1]), to_tf32(w[u][2]), to_tf32(w[u][3]));
      }
    }
  }
  if (tid_all < N) sBias[tid_all] = bias ? to_f32<TW>(bias[tid_all]) : 0.f;
  if (tid_all == 0) {
    mbar_init(sbase + LO_BAR, 1);
    mbar_init(sbase + LO_BAR + 8, 1);
    fence_mbar_init();
  }
  if (warp_all == 0) {
    __syncwarp();
    tmem_alloc(sbase + LO_BAR + 16, 256);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem = tmem_base + group * 128;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc = idesc_tf32(128, N, 0, 0);
  const uint64_t dA = smem_desc_sw128(sbase + LO_A + group * (64 * 1024), 16, 1024);
  const uint64_t dB = smem_desc_sw128(sbase + LO_B, 16, 1024);
  const uint32_t b_kblock = (uint32_t)(N * 128 / 16);

  uint32_t it = 0;
  for (int tile = blockIdx.x * LT_GROUPS + group; tile < n_tiles; tile += gridDim.x * LT_GROUPS, ++it) {
    const int m0 = tile * LT_TILE;
    // ---- stage pro(X) tile: [128 rows][K] fp32, K-block kb (32 floats) at kb * 16 KB; 8 loads in flight per thread
    {
      const int kc_shift = (K == 128) ? 5 : 4;
      const int total = LT_TILE << kc_shift;
      for (int base = 0; base < total; base += LT_TILE * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = base + u * LT_TILE + tid;
          const int r = idx >> kc_shift, c = idx & (kc - 1);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < M) v[u] = load_x4<TX>(X + (size_t)(m0 + r) * K + c * 4);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = base + u * LT_TILE + tid;
          const int r = idx >> kc_shift, c = idx & (kc - 1);
          float4 t = v[u];
          if (x_round_f16) {
            t.x = __half2float(__float2half_rn(t.x)); t.y = __half2float(__float2half_rn(t.y));
            t.z = __half2float(__float2half_rn(t.z)); t.w = __half2float(__float2half_rn(t.w));
          }
          if (pro_act) { t.x = act(t.x, pro_act); t.y = act(t.y, pro_act); t.z = act(t.z, pro_act); t.w = act(t.w, pro_act); }
          t.x = to_tf32(t.x); t.y = to_tf32(t.y); t.z = to_tf32(t.z); t.w = to_tf32(t.w);
          *reinterpret_cast<float4*>(sA + (c >> 3) * (128 * 128) + sw128_off(r, c & 7)) = t;
        }
      }
    }
    fence_async_smem();
    fence_before_sync();  // previous tile's accumulator reads are complete
    group_sync();
    if (tid == 0) {
      fence_after_sync();
      for (int k = 0; k < K / 8; ++k)
        mma_tf32(tmem, dA + (uint64_t)((k >> 2) * (128 * 128 / 16) + (k & 3) * 2), dB + (uint64_t)((k >> 2) * b_kblock + (k & 3) * 2), idesc,
                 k > 0);
      mma_commit(bar);
    }
    mbar_wait(bar, it & 1u);
    fence_after_sync();
    // ---- epilogue: thread r owns output row m0 + r
    const int m = m0 + tid;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t rr[32];
      tmem_ld32(tmem + lane_sel + c0, rr);
      tmem_ld_wait();
      if (m < M) {
        const size_t o = (size_t)m * N + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __uint_as_float(rr[q * 4 + u]) + sBias[c0 + q * 4 + u];
          if (epi_act) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = act(v[u], epi_act);
          }
          if (aux) {
            float4 t;
            if (auxdt == FMD_F16) t = load4(reinterpret_cast<const __half*>(aux) + o + q * 4);
            else t = load4(reinterpret_cast<const float*>(aux) + o + q * 4);
            v[0] *= 1.f - t.x * t.x; v[1] *= 1.f - t.y * t.y; v[2] *= 1.f - t.z * t.z; v[3] *= 1.f - t.w * t.w;
          }
          if (res) {
            const float4 t = load4(res + o + q * 4);
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          }
          store4(Y + o + q * 4, make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp_all == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// Chain of up to FMD_MAX_CHAIN dense layers on the same rows: the output tile of stage s (after its epilogue) is
// written back into the A-operand buffer IN PLACE and feeds stage s+1 without leaving the SM; the stage weights
// are streamed through the one shared weight buffer (they are L2-resident).  Same two-group structure.
struct ChainArgs {
  fmd_dense_stage st[FMD_MAX_CHAIN];
  int n_stages;
};

__device__ __forceinline__ float4 load4_dt(const void* p, size_t idx, int dt) {
  return dt == FMD_F16 ? load4(reinterpret_cast<const __half*>(p) + idx) : load4(reinterpret_cast<const float*>(p) + idx);
}
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// The chain kernel runs ONE iteration per CTA (269 row tiles on 148 SMs) with long dependent per-thread instruction
// streams (ncu: IPC 0.9 per SM with 8 warps, the loads-in-flight depth makes no difference), so thread-level parallelism is
// what it lacks: 256 threads per tile (two warps per TMEM lane quarter, each draining half of the accumulator columns),
// 512 per CTA.
constexpr int CH_GROUP = 512;
constexpr int CH_THREADS = CH_GROUP * LT_GROUPS;

template <int tpc>
__global__ void __launch_bounds__(CH_THREADS, 1)
linear_chain_tc_kernel(const void* __restrict__ X, int xdt, int M, int K0, int pro_act, int x_round_f16,
                       const __grid_constant__ ChainArgs ca) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid_all = threadIdx.x, warp_all = tid_all >> 5;
  const int group = tid_all / CH_GROUP;
  const int tid = tid_all & (CH_GROUP - 1);      // thread inside the group
  const int warp = tid >> 5;                     // TMEM lane quarter warp & 3, accumulator column slice warp >> 2
  const int row = (warp & 3) * 32 + (tid & 31);  // accumulator row drained by this thread
  constexpr int CSL = CH_GROUP / 128;            // column slices (warps per lane quarter)
  const int cslice = warp >> 2;
  float* sBias = reinterpret_cast<float*>(smem + LO_BIAS);
  const uint32_t bar = sbase + LO_BAR + 8u * group;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + LO_BAR + 16);
  uint8_t* sA = smem + LO_A + group * (64 * 1024);
  const int n_tiles = (M + LT_TILE - 1) / LT_TILE;
  if (tid_all == 0) {
    mbar_init(sbase + LO_BAR, 1);
    mbar_init(sbase + LO_BAR + 8, 1);
    fence_mbar_init();
  }
  if (warp_all == 0) {
    __syncwarp();
    tmem_alloc(sbase + LO_BAR + 16, 256);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem = tmem_base + group * 128;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const uint64_t dA = smem_desc_sw128(sbase + LO_A + group * (64 * 1024), 16, 1024);
  const uint64_t dB = smem_desc_sw128(sbase + LO_B, 16, 1024);
  uint32_t it = 0;
  // Stage weights [N][K] -> K-major TF32 B operand in the one shared weight buffer.  fp32 weights are copied with
  // cp.async (16-byte chunks straight to their swizzled position, no registers held, issued as soon as the previous
  // stage's MMAs have released the buffer and complete under the epilogue) and rounded to TF32 in place by the thread
  // that copied them; fp16 weights (the W16A16 output network) go through registers.
  auto w_issue = [&](const fmd_dense_stage& S, int K, float* bias_dst) {
    if (S.wdt != FMD_F32) return;
    const int kc = K >> 2, kc_shift = (K == 128) ? 5 : 4;
    const int N = S.N, total = N << kc_shift;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = u * CH_THREADS + tid_all;
      const int n = idx >> kc_shift, c = idx & (kc - 1);
      if (idx < total)
        cp_async16(sbase + LO_B + (uint32_t)((c >> 3) * (N * 128)) + sw128_off(n, c & 7),
                   reinterpret_cast<const float*>(S.W) + (size_t)idx * 4);
    }
    if (tid_all < N) {
      if (S.bias) cp_async4(smem_u32(bias_dst + tid_all), reinterpret_cast<const float*>(S.bias) + tid_all);
      else bias_dst[tid_all] = 0.f;
    }
    cp_async_commit();
  };
  auto w_finish = [&](const fmd_dense_stage& S, int K, float* bias_dst) {
    const int kc = K >> 2, kc_shift = (K == 128) ? 5 : 4;
    const int N = S.N, total = N << kc_shift;
    if (S.wdt == FMD_F32) {
      cp_async_wait_all();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = u * CH_THREADS + tid_all;
        const int n = idx >> kc_shift, c = idx & (kc - 1);
        if (idx < total) {
          float4* slot = reinterpret_cast<float4*>(smem + LO_B + (c >> 3) * (N * 128) + sw128_off(n, c & 7));
          float4 t = *slot;
          *slot = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
        }
      }
      return;
    }
    float4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = u * CH_THREADS + tid_all;
      wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) wv[u] = load4(reinterpret_cast<const __half*>(S.W) + (size_t)idx * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = u * CH_THREADS + tid_all;
      const int n = idx >> kc_shift, c = idx & (kc - 1);
      if (idx < total) *reinterpret_cast<float4*>(smem + LO_B + (c >> 3) * (N * 128) + sw128_off(n, c & 7)) = wv[u];
    }
    if (tid_all < N) bias_dst[tid_all] = S.bias ? __half2float(reinterpret_cast<const __half*>(S.bias)[tid_all]) : 0.f;
  };
  // tpc = row tiles per CTA and pass: 2 (both groups busy) when there are more tiles than SMs, else 1 - the second group
  // then only helps to stage the weights, and twice as many SMs work (small systems: 54 beads x 128 molecules = 54 tiles)
  const bool live = group < tpc;
  int pass = 0;
  for (int tile0 = blockIdx.x * tpc; tile0 < n_tiles; tile0 += gridDim.x * tpc, ++pass) {
    const int m0 = (tile0 + group) * LT_TILE;   // may lie beyond M for the second group: rows are then all invalid
    if (pass > 0) {        // a previous tile pass: its epilogue is done with sA, its MMAs have released the weight buffer
      fence_before_sync();
      __syncthreads();
    }
    // ---- weights of stage 0 are in flight while the pro(X) tile is staged
    w_issue(ca.st[0], K0, sBias);
    if (live) {
      const int kc = K0 >> 2, kc_shift = (K0 == 128) ? 5 : 4;
      const int total = LT_TILE << kc_shift;
      for (int base = 0; base < total; base += CH_GROUP * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * CH_GROUP + tid;
          const int r = idx >> kc_shift, c = idx & (kc - 1);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r < M) v[u] = load4_dt(X, (size_t)(m0 + r) * K0 + c * 4, xdt);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * CH_GROUP + tid;
          const int r = idx >> kc_shift, c = idx & (kc - 1);
          float4 t = v[u];
          if (x_round_f16) { t.x = round_h(t.x); t.y = round_h(t.y); t.z = round_h(t.z); t.w = round_h(t.w); }
          if (pro_act) { t.x = act(t.x, pro_act); t.y = act(t.y, pro_act); t.z = act(t.z, pro_act); t.w = act(t.w, pro_act); }
          t.x = to_tf32(t.x); t.y = to_tf32(t.y); t.z = to_tf32(t.z); t.w = to_tf32(t.w);
          *reinterpret_cast<float4*>(sA + (c >> 3) * (128 * 128) + sw128_off(r, c & 7)) = t;
        }
      }
    }
    int K = K0;
    for (int s = 0; s < ca.n_stages; ++s) {
      const fmd_dense_stage& S = ca.st[s];
      const int N = S.N;
      const float* sBiasS = sBias + (s & 1) * 128;
      const bool feed = s + 1 < ca.n_stages;
      const bool post = S.aux || S.res || S.Y;
      w_finish(S, K, sBias + (s & 1) * 128);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();   // operand tiles (A: staging / previous epilogue, B: weights) complete and visible
      if (tid == 0 && live) {
        fence_after_sync();
        const uint32_t idesc = idesc_tf32(128, N, 0, 0);
        const uint32_t b_kblock = (uint32_t)(N * 128 / 16);
        for (int k = 0; k < K / 8; ++k)
          mma_tf32(tmem, dA + (uint64_t)((k >> 2) * (128 * 128 / 16) + (k & 3) * 2),
                   dB + (uint64_t)((k >> 2) * b_kblock + (k & 3) * 2), idesc, k > 0);
        mma_commit(bar);
      }
      // coalesced epilogue phase: chunk kk of this thread is 16 bytes at row r0 + kk * rows_per_k, column block c; the
      // row step is a multiple of 8, so both the swizzled shared-memory slot and the global offset advance by constants
      const int nc_shift = (N == 128) ? 5 : 4;
      const int rows_per_k = CH_GROUP >> nc_shift;                 // 16 | 32
      const int nb = LT_TILE / (rows_per_k * 2);                   // batches of two chunks per thread
      const int r0 = tid >> nc_shift, c4 = tid & ((N >> 2) - 1);
      uint8_t* const slot0 = sA + (c4 >> 3) * (128 * 128) + sw128_off(r0, c4 & 7);
      const uint32_t slot_step = (uint32_t)(rows_per_k >> 3) * 1024u;
      const size_t o0 = (size_t)(m0 + r0) * N + c4 * 4, o_step = (size_t)rows_per_k * N;
      const int rlim = M - m0 - r0;                                // chunk kk holds a valid row iff kk * rows_per_k < rlim
      const bool has_aux = S.aux != nullptr, has_res = S.res != nullptr, has_y = S.Y != nullptr;
      const bool aux_h = S.auxdt == FMD_F16, y_h = S.ydt == FMD_F16, rnd = S.round_f16 != 0;
      float4 avA[2], rvA[2], avB[2], rvB[2];
      auto p2_issue = [&](int b, float4 (&av)[2], float4 (&rv)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = b * 2 + u;
          const size_t o = o0 + (size_t)kk * o_step;
          const bool ok = kk * rows_per_k < rlim;
          av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_aux && ok) av[u] = aux_h ? load4(reinterpret_cast<const __half*>(S.aux) + o) : load4(reinterpret_cast<const float*>(S.aux) + o);
          if (has_res && ok) rv[u] = load4(S.res + o);
        }
      };
      auto p2_apply = [&](int b, const float4 (&av)[2], const float4 (&rv)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = b * 2 + u;
          float4* slot = reinterpret_cast<float4*>(slot0 + (uint32_t)kk * slot_step);
          float4 v = *slot;
          v.x = fmaf(v.x, -av[u].x * av[u].x, v.x) + rv[u].x;     // v * (1 - aux^2) + res
          v.y = fmaf(v.y, -av[u].y * av[u].y, v.y) + rv[u].y;
          v.z = fmaf(v.z, -av[u].z * av[u].z, v.z) + rv[u].z;
          v.w = fmaf(v.w, -av[u].w * av[u].w, v.w) + rv[u].w;
          if (has_y && kk * rows_per_k < rlim) {
            const size_t o = o0 + (size_t)kk * o_step;
            if (y_h) store4(reinterpret_cast<__half*>(S.Y) + o, v);
            else store4(reinterpret_cast<float*>(S.Y) + o, v);
          }
          if (feed) {
            if (rnd) { v.x = round_h(v.x); v.y = round_h(v.y); v.z = round_h(v.z); v.w = round_h(v.w); }
            *slot = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          }
        }
      };
      if (live) {
        mbar_wait(bar, it & 1u);
        ++it;
      }
      fence_after_sync();
      // This kernel is one pass per CTA, bound by exposed load latencies rather than by throughput: the next stage's
      // weights and the first aux / residual chunks of this stage's coalesced epilogue phase leave L2 now
      if (feed) {
        fence_before_sync();
        __syncthreads();   // the MMAs of BOTH groups have completed: the weight buffer takes the next stage
        w_issue(ca.st[s + 1], N, sBias + ((s + 1) & 1) * 128);
      }
      if (!live) {
        K = N;
        continue;
      }
      if (post) p2_issue(0, avA, rvA);   // in flight during the accumulator drain
      // ---- epilogue of stage s, two phases:
      //  1. thread r drains accumulator row r: + bias, activation -> row r of the A-operand buffer, in place (the MMA
      //     that read the buffer has completed);
      //  2. the group walks the tile in row-major order, consecutive threads on consecutive 16-byte chunks of a row, so
      //     the aux / residual loads and the output stores are fully coalesced (row-per-thread global access made
      //     this epilogue 5x slower than the GEMM itself); the tile is rewritten as the TF32 operand of the next stage.
      const int cw = N / CSL;                      // columns per slice (multiple of 16)
      for (int c0 = cslice * cw; c0 < (cslice + 1) * cw; c0 += 16) {
        uint32_t rr[16];
        tmem_ld16(tmem + lane_sel + c0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
          const float4 bq = *reinterpret_cast<const float4*>(sBiasS + c0 + q * 4);
          v[0] = __uint_as_float(rr[q * 4]) + bq.x; v[1] = __uint_as_float(rr[q * 4 + 1]) + bq.y;
          v[2] = __uint_as_float(rr[q * 4 + 2]) + bq.z; v[3] = __uint_as_float(rr[q * 4 + 3]) + bq.w;
          if (S.epi_act) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = act(v[u], S.epi_act);
          }
          if (!post) {   // nothing else to apply: write the final operand form directly
            if (S.round_f16) { v[0] = round_h(v[0]); v[1] = round_h(v[1]); v[2] = round_h(v[2]); v[3] = round_h(v[3]); }
            v[0] = to_tf32(v[0]); v[1] = to_tf32(v[1]); v[2] = to_tf32(v[2]); v[3] = to_tf32(v[3]);
          }
          const int c = (c0 >> 2) + q;
          *reinterpret_cast<float4*>(sA + (c >> 3) * (128 * 128) + sw128_off(row, c & 7)) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
      if (post) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(CH_GROUP) : "memory");
        // double-buffered: the loads of batch b + 1 are in flight while batch b is applied
        for (int b = 0; b < nb; b += 2) {
          if (b + 1 < nb) p2_issue(b + 1, avB, rvB);
          p2_apply(b, avA, rvA);
          if (b + 2 < nb) p2_issue(b + 2, avA, rvA);
          if (b + 1 < nb) p2_apply(b + 1, avB, rvB);
        }
      }
      K = N;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp_all == 0) tmem_dealloc(tmem_base, 256);
}

template <typename TX, typename TW, typename TY>
int launch_tc(const void* X, const void* W, const void* bias, void* Y, int M, int N, int K, const int32_t* m_dev,
              int pro_act, int x_round, int epi_act, const void* aux, int auxdt, const float* res, int w_nk, cudaStream_t st) {
  auto kern = linear_tc_kernel<TX, TW, TY>;
  static bool attr_done = false;
  if (!attr_done) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM_ALLOC));
    attr_done = true;
  }
  const int pairs = fmd_div_up(fmd_div_up(M, LT_TILE), LT_GROUPS);
  const int grid = pairs < fmd_num_sms() ? pairs : fmd_num_sms();
  kern<<<grid, LT_THREADS, LT_SMEM_ALLOC, st>>>((const TX*)X, (const TW*)W, (const TW*)bias, (TY*)Y, M, N, K, m_dev,
                                             pro_act, x_round, epi_act, aux, auxdt, res, w_nk);
  return FMD_OK;
}

}  // namespace

extern "C" int fmd_linear_tc(const void* X, int xdt, const void* W, int wdt, const void* bias, void* Y, int ydt, int M,
                             int N, int K, const int32_t* m_dev, int pro_act, int x_round_f16, int epi_act,
                             const void* aux, int auxdt, const float* res, int w_is_nk, void* stream) {
  FMD_REQUIRE(X && W && Y && M >= 0, "fmd_linear_tc: bad arguments");
  FMD_REQUIRE((K == 64 || K == 128) && (N == 64 || N == 128),
              "fmd_linear_tc: needs K, N in {64, 128} (use fmd_linear otherwise)");
  FMD_REQUIRE((xdt | wdt | ydt | auxdt) >= 0 && xdt <= 1 && wdt <= 1 && ydt <= 1 && auxdt <= 1, "fmd_linear_tc: bad dtype");
  if (M == 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = xdt * 4 + wdt * 2 + ydt;
  int rc = FMD_OK;
#define FMD_LT(TX, TW, TY) \
  rc = launch_tc<TX, TW, TY>(X, W, bias, Y, M, N, K, m_dev, pro_act, x_round_f16, epi_act, aux, auxdt, res, w_is_nk, st)
  switch (key) {
    case 0: FMD_LT(float, float, float); break;
    case 1: FMD_LT(float, float, __half); break;
    case 2: FMD_LT(float, __half, float); break;
    case 3: FMD_LT(float, __half, __half); break;
    case 4: FMD_LT(__half, float, float); break;
    case 5: FMD_LT(__half, float, __half); break;
    case 6: FMD_LT(__half, __half, float); break;
    case 7: FMD_LT(__half, __half, __half); break;
    default: FMD_FAIL("fmd_linear_tc: bad dtype combination");
  }
#undef FMD_LT
  if (rc != FMD_OK) return rc;
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_linear_chain_tc(const void* X, int xdt, int M, int K, int pro_act, int x_round_f16,
                                   const fmd_dense_stage* stages, int n_stages, void* stream) {
  FMD_REQUIRE(X && stages && M >= 0 && n_stages >= 1 && n_stages <= FMD_MAX_CHAIN, "fmd_linear_chain_tc: bad arguments");
  FMD_REQUIRE(K == 64 || K == 128, "fmd_linear_chain_tc: K must be 64 or 128");
  ChainArgs ca;
  ca.n_stages = n_stages;
  for (int s = 0; s < n_stages; ++s) {
    ca.st[s] = stages[s];
    FMD_REQUIRE(stages[s].W && (stages[s].N == 64 || stages[s].N == 128), "fmd_linear_chain_tc: stage needs W and N in {64,128}");
    FMD_REQUIRE(stages[s].Y || s + 1 < n_stages, "fmd_linear_chain_tc: the last stage must store its output");
  }
  if (M == 0) return FMD_OK;
  static bool attr_done = false;
  if (!attr_done) {
    FMD_CUDA(cudaFuncSetAttribute(linear_chain_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM_ALLOC));
    FMD_CUDA(cudaFuncSetAttribute(linear_chain_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM_ALLOC));
    attr_done = true;
  }
  const int n_tiles = fmd_div_up(M, LT_TILE);
  const int tpc = n_tiles <= fmd_num_sms() ? 1 : LT_GROUPS;
  const int want = fmd_div_up(n_tiles, tpc);
  const int grid = want < fmd_num_sms() ? want : fmd_num_sms();
  if (tpc == 1)
    linear_chain_tc_kernel<1><<<grid, CH_THREADS, LT_SMEM_ALLOC, (cudaStream_t)stream>>>(X, xdt, M, K, pro_act, x_round_f16, ca);
  else
    linear_chain_tc_kernel<2><<<grid, CH_THREADS, LT_SMEM_ALLOC, (cudaStream_t)stream>>>(X, xdt, M, K, pro_act, x_round_f16, ca);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
