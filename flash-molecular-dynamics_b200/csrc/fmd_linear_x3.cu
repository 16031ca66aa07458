// fp32-ACCURATE dense layers on the sm_100a tensor cores ("BF16x3": fp32 emulation with bf16 slices):
//   Y[M,N] = epi( X[M,K] @ W[K,N] + bias ) [* (1 - aux^2)] [+ res]          K, N <= 128, fp32 in / fp32 out
// Every fp32 operand is split into three bf16 slices, x = s1 + s2 + s3 (8 significant bits each: s1 = rn(x),
// s2 = rn(x - s1), s3 = rn(x - s1 - s2); 24 bits in all), and the product is accumulated in fp32 TMEM as
//   acc0 += s1*s1                                      (leading term)
//   acc1 += s1*s2 + s2*s1 + s1*s3 + s2*s2 + s3*s1      (2^-8 and 2^-16 terms; s2*s3, s3*s2, s3*s3 are below fp32 rounding)
// and acc0 + acc1 is formed (round-to-nearest) in the epilogue.  Why slices of 8 bits and not the usual 3xTF32 split:
// the tensor core TRUNCATES (round-toward-zero) whenever it adds into an accumulator.  With 11-bit TF32 slices the
// 22-bit products do not fit the 24-bit accumulator next to a running sum, so every MMA truncates: a coherent bias of
// half an ulp per k-step (measured: energies of a 269-bead molecule off by 2e-5 relative, 3-10x the fp32 FMA kernel).
// Products of 8-bit slices have 16 bits; a sum of K <= 128 of them spans at most 16 + 7 = 23 bits, so the leading
// accumulator is EXACT for operands of comparable magnitude (bits are only dropped from products > 2^8 below the
// running sum), and the truncation in acc1 is scaled down by 2^-8.  Same tensor time as 3xTF32 (6 bf16 MMAs of K = 16
// per 16 columns against 3 tf32 MMAs of K = 8 per 8 columns at half the rate), less shared memory (6 instead of 8
// bytes per element).  This is the GEMM of the fp32 PARITY path (1e-5 against the reference's --disable_optim fp32
// path): the edge-level filter-network layers (reference models/mlp.py:41-57 and its autograd) and the node-level
// layers, which were 85 % of that step on the SIMT FMA kernel of fmd_linear.cu.
//
// The kernel is HBM-bound by construction (an [E,128] fp32 activation is read once and written once), so the
// structure is a streaming pipeline, one persistent CTA per SM, 17 warps:
//   P  (8 warps)  coalesced 16-byte global loads of a [128 rows x 64 k] block, a register ring keeps the loads of the
//                 next blocks in flight; slices written as three K-major swizzle-128B bf16 operand images into an
//                 n-stage shared-memory ring;
//   M  (1 thread) tcgen05.mma kind::f16 (bf16), 6 MMAs per k-step of 16, accumulator pairs double-buffered in TMEM;
//   E  (8 warps)  TMEM -> registers (row per thread) -> per-warp shared staging -> row-contiguous global stores with
//                 bias / exact tanh / (1 - aux^2) / residual applied in the coalesced phase (aux and residual loads
//                 are coalesced too and several are in flight per thread).
// The three weight images stay resident in shared memory for the life of the CTA.
#include <cuda_bf16.h>

#include "fmd_tc.cuh"

using namespace fmd;
using namespace fmd::tc;

namespace {

constexpr int X3_TILE = 128;
constexpr int X3_EPI_WARPS = 8, X3_PROD_WARPS = 8;
constexpr int X3_PREFETCH = 4;                                         // register ring: blocks in flight per producer thread
constexpr int X3_THREADS = (X3_EPI_WARPS + X3_PROD_WARPS + 1) * 32;   // 544
constexpr int X3_PROD_THREADS = X3_PROD_WARPS * 32;
constexpr int X3_KBLK = 64;                                            // k-columns per stage (128 bytes of bf16 per row)
constexpr uint32_t X3_IMG = X3_TILE * 128;                             // one slice image of a stage: 16 KB
constexpr uint32_t X3_STAGE = 3 * X3_IMG;                              // s1, s2, s3
constexpr uint32_t X3_STG_WARP = 32 * 32 * 4;                          // [32 rows][32 columns] fp32 per epilogue warp
constexpr uint32_t X3_STG = X3_EPI_WARPS * X3_STG_WARP;
constexpr int X3_MAX_STAGES = 8;
constexpr uint32_t X3_TAIL = 512 + 256;                                // bias + barriers / tmem slot
constexpr uint32_t X3_SMEM_MAX = 227 * 1024;

// kind::f16 instruction descriptor with BF16 A/B (format code 1), fp32 D, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// x -> three bf16 slices, returned in the HIGH half of each word, x == s1 + s2 + s3 EXACTLY: s1 = x rounded to 8
// significant bits (integer add of half an ulp + mask: cvt.rn.bf16.f32 runs on the slow conversion pipe and made the
// producers the bottleneck), the residual (<= 16 significant bits) truncated to its leading 8 bits, and what is
// left (<= 8 bits).  |s2| <= 2^-9 |x|, |s3| < 2^-16 |x|, so the dropped products s2*s3, s3*s2, s3*s3 are < 2^-24.
__device__ __forceinline__ void split3(float x, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
  s1 = (__float_as_uint(x) + 0x8000u) & 0xFFFF0000u;
  const float r1 = x - __uint_as_float(s1);
  s2 = __float_as_uint(r1) & 0xFFFF0000u;
  s3 = __float_as_uint(r1 - __uint_as_float(s2));
}
// two slices (high halves of lo, hi) -> one word of two bf16, element `lo` at the lower address
__device__ __forceinline__ uint32_t pack_hi16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x7632); }
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

// one [64 rows x 64 k] half-block of the A operand held in registers between its global loads and its shared-memory
// stores (the unit of the producers' register ring)
constexpr int X3_HROWS = X3_TILE / 2;
template <int VEC>
struct ABlock {
  static constexpr int PPR = X3_KBLK / VEC;                               // pieces per row: 16 (float4) or 32 (float2)
  static constexpr int NU = X3_HROWS * PPR / X3_PROD_THREADS;             // pieces per producer thread: 4 or 8
  float v[NU][VEC];
  __device__ __forceinline__ void load(const float* __restrict__ X, int M, int K, int m0, int kb, int h, int p) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int idx = u * X3_PROD_THREADS + p;
      const int r = h * X3_HROWS + idx / PPR, c = idx % PPR;
      const int row = m0 + r, col = kb * X3_KBLK + c * VEC;
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
  __device__ __forceinline__ void store(uint8_t* sImg, int h, int p) const {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int idx = u * X3_PROD_THREADS + p;
      const int r = h * X3_HROWS + idx / PPR, c = idx % PPR;
      uint32_t a[VEC], b[VEC], d[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) split3(v[u][q], a[q], b[q], d[q]);
      // 16-byte swizzle chunk = 8 consecutive k; this piece covers VEC of them
      const uint32_t off = sw128_off(r, (c * VEC) >> 3) + (uint32_t)((c * VEC) & 7) * 2u;
      if constexpr (VEC == 4) {
        *reinterpret_cast<uint2*>(sImg + off) = make_uint2(pack_hi16(a[0], a[1]), pack_hi16(a[2], a[3]));
        *reinterpret_cast<uint2*>(sImg + X3_IMG + off) = make_uint2(pack_hi16(b[0], b[1]), pack_hi16(b[2], b[3]));
        *reinterpret_cast<uint2*>(sImg + 2 * X3_IMG + off) = make_uint2(pack_hi16(d[0], d[1]), pack_hi16(d[2], d[3]));
      } else {
        *reinterpret_cast<uint32_t*>(sImg + off) = pack_hi16(a[0], a[1]);
        *reinterpret_cast<uint32_t*>(sImg + X3_IMG + off) = pack_hi16(b[0], b[1]);
        *reinterpret_cast<uint32_t*>(sImg + 2 * X3_IMG + off) = pack_hi16(d[0], d[1]);
      }
    }
  }
};

// epilogue mode "rbf backward": instead of storing Y = X @ W (= grad_rbf [E,R]), row e is contracted at once with
// d rbf_k / d d at d_e:  g_d[e] (+)= sum_k Y[e,k] * exp(gamma (d_e - mu_k)^2) * (2 gamma (d_e - mu_k) C(d_e) + C'(d_e))
// (reference: backward of FusedDistanceGaussianRBFCutoffFunction, kernels/cfconv_kernels.py:1679-1735).
struct X3Rbf {
  const float* dist;      // [M]; nullptr = normal store epilogue
  const float* centers;   // [N]
  float gamma, rc;
  float* g_d;             // [M]
  int accumulate;
};

template <int VEC>
__global__ void __launch_bounds__(X3_THREADS, 1)
linear_x3_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                 float* __restrict__ Y, int M, int N, int K, const int32_t* __restrict__ m_dev, int epi_act,
                 const float* __restrict__ aux, const float* __restrict__ res, int nstage, const X3Rbf rb) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kb = (K + X3_KBLK - 1) / X3_KBLK, K16 = (K + 15) >> 4, Npad = (N + 15) & ~15;
  const uint32_t b_block = (uint32_t)Npad * 128u;          // one K-block of one weight image
  const uint32_t b_bytes = (uint32_t)Kb * b_block;         // one weight image
  const uint32_t off_a = 3u * b_bytes;
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

  // ---- weights: zero the three images (padding rows / columns), then scatter W[k][n] -> K-major [n][k] slices
  for (uint32_t i = tid; i < (3u * b_bytes) >> 4; i += X3_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 128) sBias[tid] = tid < N ? (rb.dist ? rb.centers[tid] : (bias ? bias[tid] : 0.f)) : 0.f;
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
    uint32_t s1, s2, s3;
    split3(__ldg(W + idx), s1, s2, s3);
    const uint32_t off = (uint32_t)(k / X3_KBLK) * b_block + sw128_off(n, (k % X3_KBLK) >> 3) + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<uint16_t*>(smem + off) = (uint16_t)(s1 >> 16);
    *reinterpret_cast<uint16_t*>(smem + b_bytes + off) = (uint16_t)(s2 >> 16);
    *reinterpret_cast<uint16_t*>(smem + 2u * b_bytes + off) = (uint16_t)(s3 >> 16);
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
    // warp = (TMEM lane quarter, column half): quarter q reads accumulator rows 32q..32q+31, half h the 32-column
    // chunks h, h + 2, ...
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    float* stg = reinterpret_cast<float*>(smem + off_stg + (uint32_t)warp * X3_STG_WARP);
    const bool vec4 = (N & 3) == 0;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const uint32_t b = t & 1u;
      const int row0 = tile * X3_TILE + quarter * 32;
      float rb_d = 0.f, rb_c = 0.f, rb_dc = 0.f, rb_acc = 0.f;
      if (rb.dist) {
        rb_d = row0 + lane < M ? rb.dist[row0 + lane] : 0.f;
        rb_c = cosine_cutoff(rb_d, rb.rc);
        rb_dc = cosine_cutoff_grad(rb_d, rb.rc);
      }
      mbar_wait_guard(bar_accf + 8u * b, (t >> 1) & 1u);
      fence_after_sync();
      if (half * 32 >= Npad) {              // no columns for this warp (N <= 32): release the accumulator at once
        fence_before_sync();
        mbar_arrive(bar_acce + 8u * b);
      }
      for (int c0 = half * 32; c0 < Npad; c0 += 64) {
        const bool last = c0 + 64 >= Npad;  // this warp's last read of the accumulator
        // aux / residual of the first 16 rows of the chunk: issued before the TMEM reads so that they overlap them
        const int rsub = lane >> 3, cc = (lane & 7) * 4;
        const int col = c0 + cc;
        const bool col_ok = col < N;
        float4 av[4], rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_aux = [&](int i0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = row0 + (i0 + i) * 4 + rsub;
            const bool ok = col_ok && row < M;
            const size_t o = (size_t)row * N + col;
            if (aux && ok) av[i] = load4_stream(aux + o);
            if (res && ok) rv[i] = load4_stream(res + o);
          }
        };
        if (vec4 && !rb.dist && (aux || res)) load_aux(0);
        if (!rb.dist) __syncwarp();         // the previous chunk has been read out of the staging rows
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {    // two 16-column reads: main + lower-order accumulator
          uint32_t rr[16], rs[16];
          tmem_ld16(tmem_base + b * 256u + lane_sel + (uint32_t)(c0 + hc * 16), rr);           // s1*s1
          tmem_ld16(tmem_base + b * 256u + 128u + lane_sel + (uint32_t)(c0 + hc * 16), rs);    // lower-order terms
          tmem_ld_wait();
          if (last && hc == 1) {
            fence_before_sync();
            mbar_arrive(bar_acce + 8u * b);
          }
          float y[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) y[q] = __uint_as_float(rr[q]) + __uint_as_float(rs[q]);
          if (rb.dist) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int k = c0 + hc * 16 + q;
              const float diff = rb_d - sBias[k < N ? k : 0];
              const float ex = expf(rb.gamma * diff * diff);
              const float drbf = ex * fmaf(2.0f * rb.gamma * diff, rb_c, rb_dc);
              if (k < N) rb_acc = fmaf(y[q], drbf, rb_acc);
            }
          } else {
            // row `lane`, 16-byte chunk j stored at chunk position j ^ (lane & 7): conflict-free writes (row per lane)
            // and conflict-free row-contiguous reads below
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(stg + lane * 32 + (((hc * 4 + q) ^ (lane & 7)) << 2)) =
                  make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          }
        }
        if (rb.dist) continue;
        __syncwarp();
        if (vec4) {
          // lane -> (row 4i + lane/8, 16-byte chunk lane%8): a warp store covers 4 rows x 128 contiguous bytes
          const float4 bv = *reinterpret_cast<const float4*>(sBias + (col_ok ? col : 0));
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 4) {
            if (i0 && (aux || res)) load_aux(i0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = (i0 + i) * 4 + rsub;
              const int row = row0 + rl;
              float4 v = *reinterpret_cast<const float4*>(stg + rl * 32 + (((lane & 7) ^ (rl & 7)) << 2));
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
              if (epi_act) {
                v.x = act_exact(v.x, epi_act); v.y = act_exact(v.y, epi_act);
                v.z = act_exact(v.z, epi_act); v.w = act_exact(v.w, epi_act);
              }
              if (aux || res) {
                v.x = fmaf(v.x, -av[i].x * av[i].x, v.x) + rv[i].x;     // v * (1 - aux^2) + res
                v.y = fmaf(v.y, -av[i].y * av[i].y, v.y) + rv[i].y;
                v.z = fmaf(v.z, -av[i].z * av[i].z, v.z) + rv[i].z;
                v.w = fmaf(v.w, -av[i].w * av[i].w, v.w) + rv[i].w;
              }
              if (col_ok && row < M) store4(Y + (size_t)row * N + col, v);
            }
          }
        } else {
          // any N: lane -> column c0 + lane of one row per instruction (128 contiguous bytes)
          const int colS = c0 + lane;
          const bool okS = colS < N;
          const float bvS = sBias[okS ? colS : 0];
#pragma unroll 4
          for (int i = 0; i < 32; ++i) {
            const int row = row0 + i;
            if (!(okS && row < M)) continue;
            const size_t o = (size_t)row * N + colS;
            float v = stg[i * 32 + ((((lane >> 2) ^ (i & 7)) << 2) | (lane & 3))] + bvS;
            v = act_exact(v, epi_act);
            if (aux) { const float a = aux[o]; v = fmaf(v, -a * a, v); }
            if (res) v += res[o];
            Y[o] = v;
          }
        }
      }
      if (rb.dist) {
        // the two column halves of a row live in warps w and w + 4: the upper half hands its partial sum over through
        // its own staging row (double-buffered by tile parity), fixed order of the final addition
        float* slot = reinterpret_cast<float*>(smem + off_stg + (uint32_t)(quarter + 4) * X3_STG_WARP) + (t & 1u) * 32 + lane;
        if (half == 1) *slot = rb_acc;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(64) : "memory");
        if (half == 0 && row0 + lane < M) {
          const float v = rb_acc + *slot;
          float* o = rb.g_d + row0 + lane;
          *o = rb.accumulate ? *o + v : v;
        }
      }
    }
  } else if (warp < X3_EPI_WARPS + X3_PROD_WARPS) {
    // =========================== P: producers ===========================
    const int p = tid - X3_EPI_WARPS * 32;
    // register ring of X3_PREFETCH half-blocks: the loads of half-block j + X3_PREFETCH - 1 are issued before half-block
    // j is split and stored, so that each thread keeps (X3_PREFETCH - 1) x 4 16-byte loads in flight (the kernel is
    // bound by bytes in flight against the HBM latency, not by issue slots)
    ABlock<VEC> blk[X3_PREFETCH];
    int tile = blockIdx.x, kb = 0, h = 0;            // half-block j
    int ptile = blockIdx.x, pkb = 0, ph_ = 0;        // half-block being prefetched
    auto advance = [&](int& t_, int& k_, int& h_) {
      if (++h_ == 2) {
        h_ = 0;
        if (++k_ == Kb) { k_ = 0; t_ += gridDim.x; }
      }
    };
#pragma unroll
    for (int d = 0; d < X3_PREFETCH - 1; ++d) {
      if (ptile < n_tiles) blk[d].load(X, M, K, ptile * X3_TILE, pkb, ph_, p);
      advance(ptile, pkb, ph_);
    }
    // The generic->async proxy fence is executed by the WRITING threads before they arrive (the documented pattern).  A
    // variant with one fence by the MMA thread after it had acquired the stage passed every parity test and was 6 %
    // faster (the producers' fence waits for their prefetched loads too), but long trajectories then showed rare
    // run-to-run differences in their statistics, so it is not used.
    uint32_t s = 0, ph = 0;                          // ring slot / phase of the stage being filled
    while (tile < n_tiles) {
#pragma unroll
      for (int d = 0; d < X3_PREFETCH; ++d) {
        if (tile < n_tiles) {
          if (ptile < n_tiles) blk[(d + X3_PREFETCH - 1) % X3_PREFETCH].load(X, M, K, ptile * X3_TILE, pkb, ph_, p);
          advance(ptile, pkb, ph_);
          if (h == 0) mbar_wait_guard(bar_empty + 8u * s, ph ^ 1u);
          blk[d].store(smem + off_a + s * X3_STAGE, h, p);
          if (h == 1) {
            fence_async_smem();
            mbar_arrive(bar_full + 8u * s);
            if (++s == (uint32_t)nstage) { s = 0; ph ^= 1u; }
          }
          advance(tile, kb, h);
        }
      }
    }
  } else if (lane == 0) {
    // =========================== M: MMA issuer ===========================
    const uint32_t idesc = idesc_bf16(128, Npad);
    const uint64_t dB1 = smem_desc_sw128(sbase, 16, 1024);
    const uint64_t dB2 = smem_desc_sw128(sbase + b_bytes, 16, 1024);
    const uint64_t dB3 = smem_desc_sw128(sbase + 2u * b_bytes, 16, 1024);
    uint32_t s = 0, ph = 0, t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const uint32_t b = t & 1u;
      mbar_wait_guard(bar_acce + 8u * b, ((t >> 1) & 1u) ^ 1u);
      fence_after_sync();
      const uint32_t d0 = tmem_base + b * 256u, d1 = d0 + 128u;
      for (int kb = 0; kb < Kb; ++kb) {
        mbar_wait_guard(bar_full + 8u * s, ph);
        fence_after_sync();
        const uint32_t a_addr = sbase + off_a + s * X3_STAGE;
        const uint64_t dA1 = smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t dA2 = smem_desc_sw128(a_addr + X3_IMG, 16, 1024);
        const uint64_t dA3 = smem_desc_sw128(a_addr + 2u * X3_IMG, 16, 1024);
        const uint64_t kblk = (uint64_t)((uint32_t)kb * (b_block >> 4));
        const int ksteps = min(X3_KBLK / 16, K16 - kb * (X3_KBLK / 16));
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ka = (uint64_t)(ks * 2), kw = kblk + ka;    // 32 bytes per k-step of 16 bf16
          const uint32_t acc = (kb | ks) != 0;
          mma_f16(d1, dA3 + ka, dB1 + kw, idesc, acc);               // smallest terms first
          mma_f16(d1, dA1 + ka, dB3 + kw, idesc, 1);
          mma_f16(d1, dA2 + ka, dB2 + kw, idesc, 1);
          mma_f16(d1, dA2 + ka, dB1 + kw, idesc, 1);
          mma_f16(d1, dA1 + ka, dB2 + kw, idesc, 1);
          mma_f16(d0, dA1 + ka, dB1 + kw, idesc, acc);
        }
        mma_commit(bar_empty + 8u * s);
        if (++s == (uint32_t)nstage) { s = 0; ph ^= 1u; }
      }
      mma_commit(bar_accf + 8u * b);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == X3_EPI_WARPS + X3_PROD_WARPS) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static int launch_x3(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, const int32_t* m_dev,
                     int epi_act, const float* aux, const float* res, const X3Rbf& rb, void* stream) {
  if (M == 0) return FMD_OK;
  const int Kb = (K + X3_KBLK - 1) / X3_KBLK, Npad = (N + 15) & ~15;
  const uint32_t b_bytes = 3u * (uint32_t)Kb * (uint32_t)Npad * 128u;
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
    linear_x3_kernel<4><<<grid, X3_THREADS, smem, st>>>(X, W, bias, Y, M, N, K, m_dev, epi_act, aux, res, nstage, rb);
  else
    linear_x3_kernel<2><<<grid, X3_THREADS, smem, st>>>(X, W, bias, Y, M, N, K, m_dev, epi_act, aux, res, nstage, rb);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_linear_x3(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K,
                             const int32_t* m_dev, int epi_act, const float* aux, const float* res, void* stream) {
  FMD_REQUIRE(X && W && Y && M >= 0, "fmd_linear_x3: bad arguments");
  FMD_REQUIRE(K >= 2 && K <= 128 && (K & 1) == 0 && N >= 1 && N <= 128,
              "fmd_linear_x3: needs even K <= 128 and N <= 128 (use fmd_linear otherwise)");
  X3Rbf rb = {nullptr, nullptr, 0.f, 0.f, nullptr, 0};
  return launch_x3(X, W, bias, Y, M, N, K, m_dev, epi_act, aux, res, rb, stream);
}

extern "C" int fmd_linear_x3_rbf_bwd(const float* X, const float* W, int M, int num_rbf, int K, const int32_t* m_dev,
                                     const float* dist, const float* centers, float gamma, float rc, float* g_d,
                                     int accumulate, void* stream) {
  FMD_REQUIRE(X && W && dist && centers && g_d && M >= 0, "fmd_linear_x3_rbf_bwd: bad arguments");
  FMD_REQUIRE(K >= 2 && K <= 128 && (K & 1) == 0 && num_rbf >= 1 && num_rbf <= 128,
              "fmd_linear_x3_rbf_bwd: needs even K <= 128 and num_rbf <= 128");
  X3Rbf rb = {dist, centers, gamma, rc, g_d, accumulate};
  return launch_x3(X, W, nullptr, nullptr, M, num_rbf, K, m_dev, 0, nullptr, nullptr, rb, stream);
}
