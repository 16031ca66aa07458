// Fused filter-network (x) CFConv kernels on the sm_100a tensor cores (tcgen05.mma, TMEM accumulators).
//
// The W16A16 path of the reference materialises, per interaction block, t = tanh(rbf Wf0^T + b) and
// W = t Wf1^T as fp16 [E,F] tensors in HBM (models/gptq.py:92-130, kernels/cfconv_kernels.py:644-952)
// and then streams W through the CSR segment reduce (kernels/csr_kernels.py:625-724).  Here one kernel
// per 128-edge tile recomputes the radial basis from d_e, runs BOTH filter GEMMs on the tensor cores
// and reduces the messages in the accumulator epilogue, so no [E,F] or [E,R] tensor ever reaches HBM.
//
// Everything is computed TRANSPOSED so that a TMEM lane (== a thread) owns one FEATURE and the 128
// accumulator columns are the tile's EDGES:
//     D1[j,e] = sum_k Wf0[j,k] rbf[e,k]      A = Wf0  [128 x 64]  K-major (resident in smem)
//                                            B = rbf  [128e x 64] K-major (written by thread-per-edge)
//     t[j,e]  = tanh(D1 + b_j) -> fp16       thread j writes ITS row of t^T, e-contiguous
//     D2[f,e] = sum_j Wf1[f,j] t[e,j]        A = Wf1  [128 x 128] K-major (resident)
//                                            B = t^T  [128j x 128e] MN-major
//     m[i,f]  = sum_{e in seg(i)} D2[f,e] * x[dst_e,f] * C(d_e)
// The segment reduce over e is then a serial in-thread sum (deterministic), the bias is a per-thread
// scalar and every gather x[dst_e, :] is one coalesced 512-byte row read across the CTA.
//
// Edges are sorted by their segment owner (fmd_nl_fill order); a segment that straddles tiles is
// completed by a tiny fix-up kernel from per-tile "head" partial sums in a fixed order: no atomics.
#include "fmd_filter_shared.cuh"

using namespace fmd;
using namespace fmd::tc;
using namespace fmd::filt;

namespace {

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t OFF_WF0 = 0;                      // 16 KB
constexpr uint32_t OFF_WF1 = OFF_WF0 + 128 * 128;    // 32 KB (2 K-blocks of [128][64])
constexpr uint32_t OFF_RBF = OFF_WF1 + 2 * 128 * 128;  // 16 KB
constexpr uint32_t OFF_TT = OFF_RBF + 128 * 128;     // 32 KB (2 MN-blocks of [128 j][64 e])
constexpr uint32_t OFF_META = OFF_TT + 2 * 128 * 128;  // 2 KB
constexpr uint32_t OFF_BIAS = OFF_META + TILE * 16;  // 512 B
constexpr uint32_t OFF_CEN = OFF_BIAS + NF * 4;      // 256 B
constexpr uint32_t OFF_BAR = OFF_CEN + RP * 4;       // mbarriers (8 B each at +0, +8, +24) + tmem slot (+16)
constexpr uint32_t OFF_RED = OFF_BAR + 32;           // 2 KB: per-warp partial sums of the cut-off term (bwd)
constexpr uint32_t FWD_SMEM = OFF_RED + 4 * TILE * 4;
constexpr uint32_t FWD_SMEM_ALLOC = FWD_SMEM + 1024;  // alignment slack

template <bool kDump>
__global__ void __launch_bounds__(TILE, 2)
filter_cfconv_fwd_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                         const int32_t* __restrict__ edge_nbr, const int32_t* __restrict__ seg_ptr, int capacity,
                         const int32_t* __restrict__ n_edges_dev, const __half* __restrict__ wf0,
                         const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                         const float* __restrict__ centers, int R, float gamma, float rc,
                         const float* __restrict__ x, float* __restrict__ out, float* __restrict__ part,
                         __half* __restrict__ dbg_t, __half* __restrict__ dbg_w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  EdgeMeta* sMeta = reinterpret_cast<EdgeMeta*>(smem + OFF_META);
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sCen = reinterpret_cast<float*>(smem + OFF_CEN);
  const uint32_t bar1 = sbase + OFF_BAR, bar2 = sbase + OFF_BAR + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 16);

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;

  // ---- one-time setup: weights -> smem (swizzled), barriers, TMEM
  load_weight_kmajor(smem + OFF_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + OFF_WF1, wf1, NF, NF / 8);
  sBias[tid] = bf0 ? __half2float(bf0[tid]) : 0.f;
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + OFF_BAR + 16, 256);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_d1 = tmem, tm_d2 = tmem + 128;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
  constexpr uint32_t IDESC2 = idesc_f16(128, 128, 0, 1);
  const uint64_t dA1 = smem_desc_sw128(sbase + OFF_WF0, 16, 1024);
  const uint64_t dB1 = smem_desc_sw128(sbase + OFF_RBF, 16, 1024);
  const uint64_t dA2 = smem_desc_sw128(sbase + OFF_WF1, 16, 1024);
  const uint64_t dB2 = smem_desc_sw128(sbase + OFF_TT, 128 * 128, 1024);

  const float bias = sBias[tid];
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int e_base = tile * TILE;
    const int n_valid = min(TILE, E - e_base);
    const uint32_t parity = it & 1u;
    __syncthreads();  // previous tile: every thread is done with sMeta
    // ---- S0: thread-per-edge: metadata + radial basis row
    {
      const bool valid = tid < n_valid;
      EdgeMeta m;
      m.owner = 0; m.nbr = 0; m.cut = 0.f; m.dist = 0.f;
      if (valid) {
        m.dist = dist[e_base + tid];
        m.owner = edge_owner[e_base + tid];
        m.nbr = edge_nbr[e_base + tid];
        m.cut = cosine_cutoff(m.dist, rc);
      }
      sMeta[tid] = m;
      write_rbf_row(smem + OFF_RBF, sCen, tid, m.dist, m.cut, R, gamma, valid);
    }
    fence_async_smem();
    __syncthreads();
    // ---- S1: D1 = Wf0 . rbf^T
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < RP / 16; ++s) mma_f16(tm_d1, dA1 + 2 * s, dB1 + 2 * s, IDESC1, s > 0);
      mma_commit(bar1);
    }
    // ---- S2: t = tanh(D1 + b) -> fp16 -> row `tid` of t^T (MN-major B operand of the second GEMM)
    mbar_wait(bar1, parity);
    fence_after_sync();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tm_d1 + lane_sel + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float t0 = tanh_approx(__uint_as_float(r[q * 8 + 2 * u]) + bias);
          const float t1 = tanh_approx(__uint_as_float(r[q * 8 + 2 * u + 1]) + bias);
          p[u] = pack_half2(t0, t1);
          if (kDump && dbg_t) {
            const int e0 = c * 32 + q * 8 + 2 * u;
            if (e0 < n_valid) dbg_t[(size_t)(e_base + e0) * NF + tid] = __float2half_rn(t0);
            if (e0 + 1 < n_valid) dbg_t[(size_t)(e_base + e0 + 1) * NF + tid] = __float2half_rn(t1);
          }
        }
        const int chunk = c * 4 + q;  // 16-byte chunk = 8 consecutive edges
        *reinterpret_cast<uint4*>(smem + OFF_TT + (chunk >> 3) * (128 * 128) + sw128_off(tid, chunk & 7)) =
            make_uint4(p[0], p[1], p[2], p[3]);
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- S3: D2 = Wf1 . t^T
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < NF / 16; ++s)
        mma_f16(tm_d2, dA2 + (uint64_t)((s >> 2) * (128 * 128 / 16) + (s & 3) * 2), dB2 + (uint64_t)(s * (2048 / 16)),
                IDESC2, s > 0);
      mma_commit(bar2);
    }
    // ---- S4: m[owner, tid] = sum_e D2[tid, e] * x[nbr_e, tid] * C_e
    mbar_wait(bar2, parity);
    fence_after_sync();
    int cur = sMeta[0].owner;
    float acc = 0.f;
    auto flush = [&](int node, float v) {
      if (__ldg(&seg_ptr[node]) >= e_base) out[(size_t)node * NF + tid] = v;  // segment starts in this tile
      else part[(size_t)tile * NF + tid] = v;                                 // head partial of a straddling segment
    };
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      if (c * 32 >= n_valid) break;
      uint32_t r[32];
      tmem_ld32(tm_d2 + lane_sel + c * 32, r);
      float xv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const EdgeMeta m = sMeta[c * 32 + i];
        xv[i] = __ldg(&x[(size_t)m.nbr * NF + tid]) * m.cut;
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int e = c * 32 + i;
        if (e < n_valid) {
          const int owner = sMeta[e].owner;
          if (owner != cur) {
            flush(cur, acc);
            cur = owner;
            acc = 0.f;
          }
          acc = fmaf(__uint_as_float(r[i]), xv[i], acc);
          if (kDump && dbg_w) dbg_w[(size_t)(e_base + e) * NF + tid] = __float2half_rn(__uint_as_float(r[i]));
        }
      }
    }
    flush(cur, acc);
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------
// Backward of one interaction block's edge part (weights frozen, only d/d(dist) is needed):
//   gW0[e,f] = a[nbr_e,f] * g_m[owner_e,f]                       (reference fused_grad_filter_out without C)
//   D3[j,e]  = sum_f Wf1[f,j] gW0[e,f]                           A = Wf1 read MN-major (same smem as the fwd A)
//   g_t[e,j] = C_e * D3[j,e] * (1 - t[e,j]^2)  -> fp16           t recomputed (D1, as in the forward)
//   D4[e,k]  = sum_j g_t[e,j] Wf0[j,k]                           A = g_t^T MN-major, B = Wf0 read MN-major
//   g_d[e]  += sum_k D4[e,k] d(rbf_k)/dd  +  C'(d_e) * sum_j t[e,j] D3[j,e]
// The last term is the exact d(cutoff)/d(distance) contribution the reference's Triton backward drops
// (kernels/csr_kernels.py:912): sum_f a W g_m == sum_j t_j D3_j by linearity, so W itself is not needed.
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}

template <bool kExact>
__global__ void __launch_bounds__(TILE, 2)
filter_cfconv_bwd_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                         const int32_t* __restrict__ edge_nbr, int capacity, const int32_t* __restrict__ n_edges_dev,
                         const __half* __restrict__ wf0, const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                         const float* __restrict__ centers, int R, float gamma, float rc, const float* __restrict__ a,
                         const float* __restrict__ g_m, float* __restrict__ g_d, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  EdgeMeta* sMeta = reinterpret_cast<EdgeMeta*>(smem + OFF_META);
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sCen = reinterpret_cast<float*>(smem + OFF_CEN);
  float* sRed = reinterpret_cast<float*>(smem + OFF_RED);  // [4 warps][128 edges]
  const uint32_t bar1 = sbase + OFF_BAR, bar3 = sbase + OFF_BAR + 8, bar4 = sbase + OFF_BAR + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 16);

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;

  load_weight_kmajor(smem + OFF_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + OFF_WF1, wf1, NF, NF / 8);
  sBias[tid] = bf0 ? __half2float(bf0[tid]) : 0.f;
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar3, 1);
    mbar_init(bar4, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + OFF_BAR + 16, 256);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_d13 = tmem, tm_d4 = tmem + 128;  // D3 reuses D1's columns (D1 is drained before MMA3)
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
  constexpr uint32_t IDESC3 = idesc_f16(128, 128, 1, 1);
  constexpr uint32_t IDESC4 = idesc_f16(128, 64, 1, 1);
  const uint64_t dA1 = smem_desc_sw128(sbase + OFF_WF0, 16, 1024);
  const uint64_t dB1 = smem_desc_sw128(sbase + OFF_RBF, 16, 1024);
  const uint64_t dA3 = smem_desc_sw128(sbase + OFF_WF1, 128 * 128, 1024);  // Wf1 [f][j] as MN-major A (M = j)
  const uint64_t dB3 = smem_desc_sw128(sbase + OFF_TT, 128 * 128, 1024);   // gW0^T [f][e] MN-major
  const uint64_t dA4 = smem_desc_sw128(sbase + OFF_TT, 128 * 128, 1024);   // g_t^T [j][e] MN-major A (M = e)
  const uint64_t dB4 = smem_desc_sw128(sbase + OFF_WF0, 16, 1024);         // Wf0 [j][k] as MN-major B (N = k)

  const float bias = sBias[tid];
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int e_base = tile * TILE;
    const int n_valid = min(TILE, E - e_base);
    const uint32_t parity = it & 1u;
    __syncthreads();
    // ---- S0: thread-per-edge metadata + radial basis row
    float my_d = 0.f, my_cut = 0.f;
    {
      const bool valid = tid < n_valid;
      EdgeMeta m;
      m.owner = 0; m.nbr = 0; m.cut = 0.f; m.dist = 0.f;
      if (valid) {
        m.dist = dist[e_base + tid];
        m.owner = edge_owner[e_base + tid];
        m.nbr = edge_nbr[e_base + tid];
        m.cut = cosine_cutoff(m.dist, rc);
      }
      my_d = m.dist;
      my_cut = m.cut;
      sMeta[tid] = m;
      write_rbf_row(smem + OFF_RBF, sCen, tid, m.dist, m.cut, R, gamma, valid);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < RP / 16; ++s) mma_f16(tm_d13, dA1 + 2 * s, dB1 + 2 * s, IDESC1, s > 0);
      mma_commit(bar1);
    }
    // ---- G: thread f writes row f of gW0^T (overlaps the first GEMM)
    {
      int cur = -1;
      float gm = 0.f;
#pragma unroll 2
      for (int q = 0; q < 16; ++q) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const EdgeMeta m = sMeta[q * 8 + u];
          if (m.owner != cur) {
            cur = m.owner;
            gm = __ldg(&g_m[(size_t)cur * NF + tid]);
          }
          v[u] = __ldg(&a[(size_t)m.nbr * NF + tid]) * gm;
          if (q * 8 + u >= n_valid) v[u] = 0.f;
        }
        *reinterpret_cast<uint4*>(smem + OFF_TT + (q >> 3) * (128 * 128) + sw128_off(tid, q & 7)) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
      }
    }
    // ---- S2: t = tanh(D1 + b), kept in registers as packed fp16 (the value the forward used)
    uint32_t tp[64];
    mbar_wait(bar1, parity);
    fence_after_sync();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tm_d13 + lane_sel + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        tp[c * 16 + i] = pack_half2(tanh_approx(__uint_as_float(r[2 * i]) + bias),
                                    tanh_approx(__uint_as_float(r[2 * i + 1]) + bias));
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- S3: D3 = Wf1^T . gW0^T
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < NF / 16; ++s)
        mma_f16(tm_d13, dA3 + (uint64_t)(s * (2048 / 16)), dB3 + (uint64_t)(s * (2048 / 16)), IDESC3, s > 0);
      mma_commit(bar3);
    }
    // ---- S4: g_t = C_e D3 (1 - t^2) -> fp16 row of g_t^T; cut-off term partial sums
    mbar_wait(bar3, parity);
    fence_after_sync();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tm_d13 + lane_sel + c * 32, r);
      tmem_ld_wait();
      float uu[32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = q * 8 + 2 * u;
          const float2 tt = unpack_half2(tp[c * 16 + q * 4 + u]);
          const float d0 = __uint_as_float(r[i]), d1 = __uint_as_float(r[i + 1]);
          const float c0 = sMeta[c * 32 + i].cut, c1 = sMeta[c * 32 + i + 1].cut;
          uu[i] = tt.x * d0;
          uu[i + 1] = tt.y * d1;
          p[u] = pack_half2(c0 * d0 * (1.f - tt.x * tt.x), c1 * d1 * (1.f - tt.y * tt.y));
        }
        const int chunk = c * 4 + q;
        *reinterpret_cast<uint4*>(smem + OFF_TT + (chunk >> 3) * (128 * 128) + sw128_off(tid, chunk & 7)) =
            make_uint4(p[0], p[1], p[2], p[3]);
      }
      if (kExact) {
        // transposed warp reduction: lane l ends with sum over the warp's 32 features of uu[l]
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
          const bool up = (lane & w) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float send = up ? uu[i] : uu[i + w];
            const float keep = up ? uu[i + w] : uu[i];
            uu[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
          }
        }
        sRed[warp * TILE + c * 32 + lane] = uu[0];
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- S5: D4 = g_t . Wf0
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < NF / 16; ++s)
        mma_f16(tm_d4, dA4 + (uint64_t)(s * (2048 / 16)), dB4 + (uint64_t)(s * (2048 / 16)), IDESC4, s > 0);
      mma_commit(bar4);
    }
    // ---- S6: thread-per-edge: g_d[e] += sum_k D4[e,k] drbf_k/dd + C'(d) * cut-term
    mbar_wait(bar4, parity);
    fence_after_sync();
    {
      const float dcut = cosine_cutoff_grad(my_d, rc);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tm_d4 + lane_sel + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = c * 32 + i;
          const float diff = my_d - sCen[k];
          const float ex = __expf(gamma * diff * diff);
          const float drbf = ex * (2.0f * gamma * diff * my_cut + dcut);
          if (k < R) acc = fmaf(__uint_as_float(r[i]), drbf, acc);
        }
      }
      if (kExact) acc += dcut * (sRed[tid] + sRed[TILE + tid] + sRed[2 * TILE + tid] + sRed[3 * TILE + tid]);
      if (tid < n_valid) g_d[e_base + tid] = accumulate ? g_d[e_base + tid] + acc : acc;
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// out[i] = 0 for empty segments; out[i] += head partials of the later tiles a straddling segment touches,
// in tile order (deterministic).  One warp per node, float4 per lane (NF = 128).
__global__ void __launch_bounds__(256)
cfconv_fixup_kernel(const int32_t* __restrict__ seg_ptr, int n_nodes, int capacity, const float* __restrict__ part,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (node >= n_nodes) return;
  const int E = min(capacity, __ldg(&seg_ptr[n_nodes]));
  const int s0 = min(__ldg(&seg_ptr[node]), E), s1 = min(__ldg(&seg_ptr[node + 1]), E);
  float4* o = reinterpret_cast<float4*>(out + (size_t)node * NF) + lane;
  if (s1 <= s0) {
    *o = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int t0 = s0 / TILE, t1 = (s1 - 1) / TILE;
  if (t1 == t0) return;
  float4 v = *o;
  for (int t = t0 + 1; t <= t1; ++t) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(part + (size_t)t * NF) + lane);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  }
  *o = v;
}

}  // namespace

void fmd_cfconv_fixup_launch(const int32_t* seg_ptr, int n_nodes, int capacity, const float* part, float* out,
                             cudaStream_t st) {
  cfconv_fixup_kernel<<<fmd_div_up((long long)n_nodes * 32, 256), 256, 0, st>>>(seg_ptr, n_nodes, capacity, part, out);
}

extern "C" int fmd_filter_cfconv_fwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                                     const int32_t* seg_ptr, int n_nodes, int capacity, const int32_t* n_edges_dev,
                                     const void* wf0_h, const void* bf0_h, const void* wf1_h, const float* centers,
                                     int num_rbf, float gamma, float rc, const float* x, int n_feat, float* out,
                                     float* part, void* dbg_t, void* dbg_w, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && seg_ptr && wf0_h && wf1_h && centers && x && out && part,
              "fmd_filter_cfconv_fwd: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf <= RP, "fmd_filter_cfconv_fwd: needs F == 128 and num_rbf <= 64");
  if (n_nodes <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool dump = dbg_t || dbg_w;
  auto kern = dump ? filter_cfconv_fwd_kernel<true> : filter_cfconv_fwd_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[dump]) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_ALLOC));
    attr_done[dump] = true;
  }
  if (capacity > 0) {
    const int max_tiles = fmd_div_up(capacity, TILE);
    const int grid = max_tiles < 2 * fmd_num_sms() ? max_tiles : 2 * fmd_num_sms();
    kern<<<grid, TILE, FWD_SMEM_ALLOC, st>>>(dist, edge_owner, edge_nbr, seg_ptr, capacity, n_edges_dev,
                                             (const __half*)wf0_h, (const __half*)bf0_h, (const __half*)wf1_h, centers,
                                             num_rbf, gamma, rc, x, out, part, (__half*)dbg_t, (__half*)dbg_w);
    FMD_CHECK_LAUNCH();
  }
  cfconv_fixup_kernel<<<fmd_div_up((long long)n_nodes * 32, 256), 256, 0, st>>>(seg_ptr, n_nodes, capacity, part, out);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_filter_cfconv_bwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr, int capacity,
                                     const int32_t* n_edges_dev, const void* wf0_h, const void* bf0_h,
                                     const void* wf1_h, const float* centers, int num_rbf, float gamma, float rc,
                                     const float* a, const float* g_m, int n_feat, float* g_d, int accumulate,
                                     int exact_cutoff_grad, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && wf0_h && wf1_h && centers && a && g_m && g_d,
              "fmd_filter_cfconv_bwd: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf <= RP, "fmd_filter_cfconv_bwd: needs F == 128 and num_rbf <= 64");
  if (capacity <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ex = exact_cutoff_grad ? 1 : 0;
  auto kern = ex ? filter_cfconv_bwd_kernel<true> : filter_cfconv_bwd_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[ex]) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_ALLOC));
    attr_done[ex] = true;
  }
  const int max_tiles = fmd_div_up(capacity, TILE);
  const int grid = max_tiles < 2 * fmd_num_sms() ? max_tiles : 2 * fmd_num_sms();
  kern<<<grid, TILE, FWD_SMEM_ALLOC, st>>>(dist, edge_owner, edge_nbr, capacity, n_edges_dev, (const __half*)wf0_h,
                                           (const __half*)bf0_h, (const __half*)wf1_h, centers, num_rbf, gamma, rc, a,
                                           g_m, g_d, accumulate);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
