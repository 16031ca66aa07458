// fp32-ACCURATE dense layers on the sm_100a tensor cores ("3xTF32"):
//   Y[M,N] = epi( X[M,K] @ W[K,N] + bias ) [* (1 - aux^2)] [+ res]          K, N <= 128, fp32 in / fp32 out
// Every fp32 operand is split into two TF32 numbers, x = hi + lo (hi = rn_tf32(x), lo = rn_tf32(x - hi): 22
// significant bits), and the product is accumulated in fp32 TMEM accumulators as hi*hi and, SEPARATELY, lo*hi + hi*lo
// (the lo*lo term is below fp32 rounding).  The tensor core truncates (round-toward-zero) every time it adds into
// an accumulator, a coherent bias of about half an ulp per MMA: keeping the two small-term MMAs of every k-step out
// of the main accumulator leaves it K/8 truncations instead of 3K/8; the two accumulators are added (RN) in the
// epilogue.  This is the GEMM of the fp32 PARITY path (1e-5 against the reference's
// --disable_optim fp32 path): it replaces the SIMT FMA kernel of fmd_linear.cu for the edge-level filter-network
// layers (reference models/mlp.py:41-57 and its autograd) and the node-level layers, which were 85 % of that step.
//
// The kernel is HBM-bound by construction (an [E,128] fp32 activation is read once and written once), so the
// structure is a streaming pipeline, one persistent CTA per SM, 13 warps:
//   P  (4 warps)  coalesced 16-byte global loads of a [128 rows x 32 k] block one block ahead, hi/lo split,
//                 written as two K-major swizzle-128B operand images into an n-stage shared-memory ring;
//   M  (1 thread) tcgen05.mma kind::tf32, 3 MMAs per k-step of 8, accumulators double-buffered in TMEM;
//   E  (8 warps)  TMEM -> registers (row per thread) -> per-warp shared staging -> row-contiguous global stores with
//                 bias / exact tanh / (1 - aux^2) / residual applied in the coalesced phase (aux and residual loads
//                 are coalesced too and several are in flight per thread).
// Both weight images (hi, lo) stay resident in shared memory for the life of the CTA.
#include "fmd_tc.cuh"

using namespace fmd;
using namespace fmd::tc;

namespace {

constexpr int X3_TILE = 128;
constexpr int X3_EPI_WARPS = 8, X3_PROD_WARPS = 8;
constexpr int X3_PREFETCH = 4;                                         // register ring: blocks in flight per producer thread
constexpr int X3_THREADS = (X3_EPI_WARPS + X3_PROD_WARPS + 1) * 32;   // 544
constexpr int X3_PROD_THREADS = X3_PROD_WARPS * 32;
constexpr uint32_t X3_HALF_STAGE = X3_TILE * 128;                      // one operand image of a stage: 16 KB
constexpr uint32_t X3_STAGE = 2 * X3_HALF_STAGE;                       // hi + lo
constexpr int X3_STG_LD = 20;                                          // floats per staging row (16 + pad)
constexpr uint32_t X3_STG_WARP = 32 * X3_STG_LD * 4;                   // 2560 B per epilogue warp
constexpr uint32_t X3_STG = X3_EPI_WARPS * X3_STG_WARP;
constexpr int X3_MAX_STAGES = 8;
constexpr uint32_t X3_TAIL = 512 + 256;                                // bias + barriers / tmem slot
constexpr uint32_t X3_SMEM_MAX = 227 * 1024;

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// round-to-nearest to TF32 (10 explicit mantissa bits) with integer ops; exact for finite inputs
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float act_exact(float v, int a) {
  if (a == FMD_ACT_TANH) return tanhf(v);
  if (a == FMD_ACT_TANH_CLAMPED) return tanh_clamped(v);
  return v;
}
__device__ __forceinline__ float2 load2_stream(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// one block of the A operand held in registers between its global loads and its shared-memory stores
template <int VEC>
struct ABlock {
  static constexpr int CH = 16 / (VEC * 4);              // pieces per 16-byte chunk: 1 (float4) or 2 (float2)
  static constexpr int NU = 1024 * CH / X3_PROD_THREADS; // pieces per producer thread per stage
  float v[NU][VEC];
  __device__ __forceinline__ void load(const float* __restrict__ X, int M, int K, int m0, int kb, int p) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int idx = u * X3_PROD_THREADS + p;
      const int r = idx / (8 * CH), c = idx % (8 * CH);  // row, piece inside the 128-byte row block
      const int row = m0 + r, col = kb * 32 + c * VEC;
      const bool ok = row < M && col < K;
      if constexpr (VEC == 4) {
        const float4 t = ok ? load4_stream(X + (size_t)row * K + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
      } else {
        const float2 t = ok ? load2_stream(X + (size_t)row * K + col) : make_float2(0.f, 0.f);
        v[u][0] = t.x; v[u][1] = t.y;
      }
    }
  }
  __device__ __forceinline__ void store(uint8_t* sHi, int p) const {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int idx = u * X3_PROD_THREADS + p;
      const int r = idx / (8 * CH), c = idx % (8 * CH);
      float hi[VEC], lo[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        hi[q] = rn_tf32(v[u][q]);
        lo[q] = rn_tf32(v[u][q] - hi[q]);
      }
      const uint32_t off = sw128_off(r, c / CH) + (uint32_t)(c % CH) * 8u;
      if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(sHi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sHi + X3_HALF_STAGE + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      } else {
        *reinterpret_cast<float2*>(sHi + off) = make_float2(hi[0], hi[1]);
        *reinterpret_cast<float2*>(sHi + X3_HALF_STAGE + off) = make_float2(lo[0], lo[1]);
      }
    }
  }
};

template <int VEC>
__global__ void __launch_bounds__(X3_THREADS, 1)
linear_x3_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                 float* __restrict__ Y, int M, int N, int K, const int32_t* __restrict__ m_dev, int epi_act,
                 const float* __restrict__ aux, const float* __restrict__ res, int nstage) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kb = (K + 31) >> 5, K8 = (K + 7) >> 3, Npad = (N + 15) & ~15;
  const uint32_t b_block = (uint32_t)Npad * 128u;          // one K-block of one weight image
  const uint32_t b_bytes = (uint32_t)Kb * b_block;         // one weight image
  const uint32_t off_a = 2u * b_bytes;
  const uint32_t off_stg = off_a + (uint32_t)nstage * X3_STAGE;
  const uint32_t off_bias = off_stg + X3_STG;
  const uint32_t off_bar = off_bias + 512u;
  float* sBias = reinterpret_cast<float*>(smem + off_bias);
  const uint32_t bar_full = sbase + off_bar;               // [nstage]  producers -> MMA
  const uint32_t bar_empty = bar_full + 8u * X3_MAX_STAGES; // [nstage]  MMA -> producers
  const uint32_t bar_accf = bar_empty + 8u * X3_MAX_STAGES; // [2]       MMA -> epilogue
  const uint32_t bar_acce = bar_accf + 16u;                // [2]       epilogue -> MMA
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + off_bar + 16 * X3_MAX_STAGES + 32);
  if (m_dev) M = min(M, *m_dev);
  const int n_tiles = (M + X3_TILE - 1) / X3_TILE;

  // ---- weights: zero both images (padding rows / columns), then scatter W[k][n] -> K-major [n][k] hi / lo
  for (uint32_t i = tid; i < (2u * b_bytes) >> 4; i += X3_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 128) sBias[tid] = (bias && tid < N) ? bias[tid] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(bar_full + 8u * s, X3_PROD_THREADS);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8u * b, 1);
      mbar_init(bar_acce + 8u * b, X3_EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  for (int idx = tid; idx < K * N; idx += X3_THREADS) {
    const int k = idx / N, n = idx - k * N;
    const float w = __ldg(W + idx);
    const float hi = rn_tf32(w), lo = rn_tf32(w - hi);
    const uint32_t off = (uint32_t)(k >> 5) * b_block + sw128_off(n, (k & 31) >> 2) + (uint32_t)(k & 3) * 4u;
    *reinterpret_cast<float*>(smem + off) = hi;
    *reinterpret_cast<float*>(smem + b_bytes + off) = lo;
  }
  if (warp == X3_EPI_WARPS + X3_PROD_WARPS) {
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < X3_EPI_WARPS) {
    // =========================== E: epilogue ===========================
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    float* stg = reinterpret_cast<float*>(smem + off_stg + (uint32_t)warp * X3_STG_WARP);
    const bool vec4 = (N & 3) == 0;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const uint32_t b = t & 1u;
      mbar_wait_guard(bar_accf + 8u * b, (t >> 1) & 1u);
      fence_after_sync();
      const int row0 = tile * X3_TILE + quarter * 32;
      if (half * 16 >= Npad) {              // no columns for this warp (N <= 16): release the accumulator at once
        fence_before_sync();
        mbar_arrive(bar_acce + 8u * b);
      }
      for (int c0 = half * 16; c0 < Npad; c0 += 32) {
        uint32_t rr[16], rs[16];
        tmem_ld16(tmem_base + b * 256u + lane_sel + (uint32_t)c0, rr);           // hi*hi
        tmem_ld16(tmem_base + b * 256u + 128u + lane_sel + (uint32_t)c0, rs);    // lo*hi + hi*lo
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) rr[q] = __float_as_uint(__uint_as_float(rr[q]) + __uint_as_float(rs[q]));
        if (c0 + 32 >= Npad) {            // this warp's last read of the accumulator
          fence_before_sync();
          mbar_arrive(bar_acce + 8u * b);
        }
        __syncwarp();                      // the previous chunk has been read out of the staging rows
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(stg + lane * X3_STG_LD + q * 4) = make_uint4(rr[q * 4], rr[q * 4 + 1], rr[q * 4 + 2], rr[q * 4 + 3]);
        __syncwarp();
        if (vec4) {
          // lane -> (row 8i + lane/4, 16-byte chunk lane%4): a warp store covers 8 rows x 64 contiguous bytes
          const int rsub = lane >> 2, cc = (lane & 3) * 4;
          const int col = c0 + cc;
          const bool col_ok = col < N;
          float4 av[4], rv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = row0 + i * 8 + rsub;
            const bool ok = col_ok && row < M;
            const size_t o = (size_t)row * N + col;
            av[i] = (aux && ok) ? load4_stream(aux + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            rv[i] = (res && ok) ? load4_stream(res + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const float4 bv = *reinterpret_cast<const float4*>(sBias + (col_ok ? col : 0));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = row0 + i * 8 + rsub;
            float4 v = *reinterpret_cast<const float4*>(stg + (i * 8 + rsub) * X3_STG_LD + cc);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (epi_act) {
              v.x = act_exact(v.x, epi_act); v.y = act_exact(v.y, epi_act);
              v.z = act_exact(v.z, epi_act); v.w = act_exact(v.w, epi_act);
            }
            v.x = fmaf(v.x, -av[i].x * av[i].x, v.x) + rv[i].x;     // v * (1 - aux^2) + res
            v.y = fmaf(v.y, -av[i].y * av[i].y, v.y) + rv[i].y;
            v.z = fmaf(v.z, -av[i].z * av[i].z, v.z) + rv[i].z;
            v.w = fmaf(v.w, -av[i].w * av[i].w, v.w) + rv[i].w;
            if (col_ok && row < M) store4(Y + (size_t)row * N + col, v);
          }
        } else {
          // any N: lane -> (row 2i + lane/16, column lane%16): 64 contiguous bytes per row
          const int rsub = lane >> 4, cc = lane & 15;
          const int col = c0 + cc;
          const bool col_ok = col < N;
          const float bv = sBias[col_ok ? col : 0];
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int row = row0 + i * 2 + rsub;
            if (!(col_ok && row < M)) continue;
            const size_t o = (size_t)row * N + col;
            float v = stg[(i * 2 + rsub) * X3_STG_LD + cc] + bv;
            v = act_exact(v, epi_act);
            if (aux) { const float a = aux[o]; v = fmaf(v, -a * a, v); }
            if (res) v += res[o];
            Y[o] = v;
          }
        }
      }
    }
  } else if (warp < X3_EPI_WARPS + X3_PROD_WARPS) {
    // =========================== P: producers ===========================
    const int p = tid - X3_EPI_WARPS * 32;
    // register ring of X3_PREFETCH blocks: the loads of block j + X3_PREFETCH - 1 are issued before block j is
    // split and stored, so that each thread keeps (X3_PREFETCH - 1) blocks of global loads in flight (the kernel is
    // bound by bytes in flight against the HBM latency, not by issue slots)
    ABlock<VEC> blk[X3_PREFETCH];
    int tile = blockIdx.x, kb = 0;           // block j
    int ptile = blockIdx.x, pkb = 0;         // block being prefetched
    auto advance = [&](int& t_, int& k_) {
      if (++k_ == Kb) { k_ = 0; t_ += gridDim.x; }
    };
#pragma unroll
    for (int d = 0; d < X3_PREFETCH - 1; ++d) {
      if (ptile < n_tiles) blk[d].load(X, M, K, ptile * X3_TILE, pkb, p);
      advance(ptile, pkb);
    }
    uint32_t j = 0;
    while (tile < n_tiles) {
#pragma unroll
      for (int d = 0; d < X3_PREFETCH; ++d) {
        if (tile < n_tiles) {
          if (ptile < n_tiles) blk[(d + X3_PREFETCH - 1) % X3_PREFETCH].load(X, M, K, ptile * X3_TILE, pkb, p);
          advance(ptile, pkb);
          const uint32_t s = j % (uint32_t)nstage, ph = (j / (uint32_t)nstage) & 1u;
          mbar_wait_guard(bar_empty + 8u * s, ph ^ 1u);
          blk[d].store(smem + off_a + s * X3_STAGE, p);
          fence_async_smem();
          mbar_arrive(bar_full + 8u * s);
          advance(tile, kb);
          ++j;
        }
      }
    }
  } else if (lane == 0) {
    // =========================== M: MMA issuer ===========================
    const uint32_t idesc = idesc_tf32(128, Npad);
    const uint64_t dB_hi = smem_desc_sw128(sbase, 16, 1024);
    const uint64_t dB_lo = smem_desc_sw128(sbase + b_bytes, 16, 1024);
    uint32_t j = 0, t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const uint32_t b = t & 1u;
      mbar_wait_guard(bar_acce + 8u * b, ((t >> 1) & 1u) ^ 1u);
      fence_after_sync();
      const uint32_t d = tmem_base + b * 256u, dsm = d + 128u;
      for (int kb = 0; kb < Kb; ++kb, ++j) {
        const uint32_t s = j % (uint32_t)nstage, ph = (j / (uint32_t)nstage) & 1u;
        mbar_wait_guard(bar_full + 8u * s, ph);
        fence_after_sync();
        const uint64_t dA_hi = smem_desc_sw128(sbase + off_a + s * X3_STAGE, 16, 1024);
        const uint64_t dA_lo = smem_desc_sw128(sbase + off_a + s * X3_STAGE + X3_HALF_STAGE, 16, 1024);
        const uint64_t kblk = (uint64_t)((uint32_t)kb * (b_block >> 4));
        const int ksteps = min(4, K8 - kb * 4);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ko = (uint64_t)(ks * 2);       // 32 bytes per k-step of 8 floats
          mma_tf32(dsm, dA_lo + ko, dB_hi + kblk + ko, idesc, (kb | ks) != 0);
          mma_tf32(dsm, dA_hi + ko, dB_lo + kblk + ko, idesc, 1);
          mma_tf32(d, dA_hi + ko, dB_hi + kblk + ko, idesc, (kb | ks) != 0);
        }
        mma_commit(bar_empty + 8u * s);
      }
      mma_commit(bar_accf + 8u * b);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == X3_EPI_WARPS + X3_PROD_WARPS) tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int fmd_linear_x3(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K,
                             const int32_t* m_dev, int epi_act, const float* aux, const float* res, void* stream) {
  FMD_REQUIRE(X && W && Y && M >= 0, "fmd_linear_x3: bad arguments");
  FMD_REQUIRE(K >= 2 && K <= 128 && (K & 1) == 0 && N >= 1 && N <= 128,
              "fmd_linear_x3: needs even K <= 128 and N <= 128 (use fmd_linear otherwise)");
  if (M == 0) return FMD_OK;
  const int Kb = (K + 31) >> 5, Npad = (N + 15) & ~15;
  const uint32_t b_bytes = 2u * (uint32_t)Kb * (uint32_t)Npad * 128u;
  const uint32_t fixed = b_bytes + X3_STG + X3_TAIL + 1024u;
  int nstage = (int)((X3_SMEM_MAX - fixed) / X3_STAGE);
  if (nstage > X3_MAX_STAGES) nstage = X3_MAX_STAGES;
  FMD_REQUIRE(nstage >= 2, "fmd_linear_x3: shared memory budget");
  const uint32_t smem = fixed + (uint32_t)nstage * X3_STAGE;
  static bool attr_done = false;
  if (!attr_done) {
    FMD_CUDA(cudaFuncSetAttribute(linear_x3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X3_SMEM_MAX));
    FMD_CUDA(cudaFuncSetAttribute(linear_x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X3_SMEM_MAX));
    attr_done = true;
  }
  const int tiles = fmd_div_up(M, X3_TILE);
  const int grid = tiles < fmd_num_sms() ? tiles : fmd_num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if ((K & 3) == 0)
    linear_x3_kernel<4><<<grid, X3_THREADS, smem, st>>>(X, W, bias, Y, M, N, K, m_dev, epi_act, aux, res, nstage);
  else
    linear_x3_kernel<2><<<grid, X3_THREADS, smem, st>>>(X, W, bias, Y, M, N, K, m_dev, epi_act, aux, res, nstage);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
