// Warp-specialised, software-pipelined version of the fused filter-network (x) CFConv forward kernel
// (see fmd_filter_tc.cu for the math and the transposed formulation).  One persistent CTA per SM,
// 17 warps in five roles connected by mbarrier rings, so the phases of consecutive 128-edge tiles overlap:
//
//   P  warps 0-3    thread-per-edge: metadata prefetch, radial basis row -> sRbf[2], sMeta[4]
//   M  warp  4      one thread issues tcgen05.mma:  D1[s] = Wf0 . rbf^T ,  D2[s] = Wf1 . t^T
//   T  warps 5-8    thread-per-feature: D1[s] -> tanh -> fp16 row of t^T -> sTT[2]
//   E0 warps 9-12   thread-per-feature epilogue of even tiles: D2[0] * x[nbr] * C -> segment sums
//   E1 warps 13-16  same for odd tiles (D2[1]); two groups hide the gather latency of x
//
// TMEM: D1[2] + D2[2] = 512 columns.  smem: weights 48 KB + sRbf 2x16 KB + sTT 2x32 KB + meta.
#include "fmd_filter_shared.cuh"

using namespace fmd;
using namespace fmd::tc;
using namespace fmd::filt;

namespace {

constexpr int NTHREADS = 17 * 32;
constexpr int BWD_THREADS = 19 * 32;   // backward: three MMA-issuer warps (the register allocation granularity is 20 warps anyway)

// Optional timeline trace (tools only; scripts/trace_roles.py): when set, CTA 0 records clock64() stamps
// {wait start, work start, end} per role and tile into trace[role][tile < 64][3].
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int role, int tile, int k, bool leader) {
  if (g_trace != nullptr && blockIdx.x == 0 && leader && tile < 64)
    g_trace[(role * 64 + tile) * 3 + k] = (unsigned long long)clock64();
}
// gathered element = *(base + byte_off): ONE 32x32+64 multiply-add forms the 64-bit address (the compiler's own sequence for
// `base + element_offset` was IADD3 + IMAD.X + LEA + LEA.HI.X per element, 4 of the 5 instructions of every gathered
// element and 20 % of all instructions of the forward kernel)
__device__ __forceinline__ float ldg_byte_off(const float* base, uint32_t byte_off) {
  float v;
  // address formation and load in ONE asm block: the 64-bit address is a transient register (with a separate mad.wide
  // the compiler hoisted all 64 addresses of a gather batch ahead of the loads and spilled)
  asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 1, %2;\n\tld.global.nc.f32 %0, [a];\n\t}"
               : "=f"(v) : "r"(byte_off), "l"(base));
  return v;
}

constexpr int META_STAGES = 4;

constexpr uint32_t O_WF0 = 0;                                // 16 KB
constexpr uint32_t O_WF1 = O_WF0 + 128 * 128;                // 32 KB
constexpr uint32_t O_RBF = O_WF1 + 2 * 128 * 128;            // 2 x 16 KB
constexpr uint32_t O_TT = O_RBF + 2 * 128 * 128;             // 2 x 32 KB
constexpr uint32_t O_XOFF = O_TT + 2 * 2 * 128 * 128;        // 4 x 512 B: BYTE offset nbr*NF*4 of the gathered row
constexpr uint32_t O_CUT = O_XOFF + META_STAGES * TILE * 4;  // 4 x 512 B: C(d_e)
constexpr uint32_t O_OWN = O_CUT + META_STAGES * TILE * 4;   // 4 x 512 B: segment owner per edge
constexpr uint32_t O_HEAD = O_OWN + META_STAGES * TILE * 4;  // 4 x 32 B: {prev_owner, -, -, -, boundary mask[4]}
constexpr uint32_t O_BIAS = O_HEAD + META_STAGES * 32;
constexpr uint32_t O_CEN = O_BIAS + NF * 4;
constexpr uint32_t O_BAR = O_CEN + RP * 4;                   // 24 mbarriers + tmem slot
constexpr uint32_t SMEM2 = O_BAR + 24 * 8 + 16;
constexpr uint32_t SMEM2_ALLOC = SMEM2 + 1024;

// barrier indices
enum { B_RBF_FULL = 0, B_RBF_EMPTY = 2, B_D1_FULL = 4, B_D1_EMPTY = 6, B_TT_FULL = 8, B_TT_EMPTY = 10,
       B_D2_FULL = 12, B_D2_EMPTY = 14, B_META_EMPTY = 16, B_META_FULL = 20, B_COUNT = 24 };

__global__ void __launch_bounds__(NTHREADS, 1)
filter_cfconv_fwd2_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                          const int32_t* __restrict__ edge_nbr, int capacity,
                          const int32_t* __restrict__ n_edges_dev, const __half* __restrict__ wf0,
                          const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                          const float* __restrict__ centers, int R, float gamma, float rc,
                          const float* __restrict__ x, float* __restrict__ out, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sBias = reinterpret_cast<float*>(smem + O_BIAS);
  float* sCen = reinterpret_cast<float*>(smem + O_CEN);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + O_BAR + 24 * 8);
  auto bar = [&](int i) { return sbase + O_BAR + 8u * (uint32_t)i; };

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- one-time setup
  load_weight_kmajor(smem + O_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + O_WF1, wf1, NF, NF / 8);
  if (tid < NF) sBias[tid] = bf0 ? __half2float(bf0[tid]) : 0.f;
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_RBF_FULL + i), 128);
      mbar_init(bar(B_RBF_EMPTY + i), 1);
      mbar_init(bar(B_D1_FULL + i), 1);
      mbar_init(bar(B_D1_EMPTY + i), 128);
      mbar_init(bar(B_TT_FULL + i), 128);
      mbar_init(bar(B_TT_EMPTY + i), 1);
      mbar_init(bar(B_D2_FULL + i), 1);
      mbar_init(bar(B_D2_EMPTY + i), 128);
    }
    for (int i = 0; i < META_STAGES; ++i) {
      mbar_init(bar(B_META_EMPTY + i), 256);   // consumed by the tanh warps (cut) and one epilogue group
      mbar_init(bar(B_META_FULL + i), 128);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + O_BAR + 24 * 8, 512);
  }
  fence_async_smem();  // weights were written through the generic proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =========================================================== P: producers (thread = edge row)
    const float g2 = gamma * 1.4426950408889634f;
    const float pi_over_rc = FMD_PI_F / rc;
    const RbfRecurrence rrec = make_rbf_recurrence(sCen, R, g2);
    int tile = blockIdx.x;
    float d_n = 0.f;
    int own_n = 0, nbr_n = 0, prev_n = -1;
    auto prefetch = [&](int t) {
      const int e = t * TILE + tid;
      d_n = 0.f; own_n = -1; nbr_n = 0; prev_n = -1;
      if (e < E) {
        d_n = __ldg(&dist[e]);
        own_n = __ldg(&edge_owner[e]);
        nbr_n = __ldg(&edge_nbr[e]);
        if (e > 0) prev_n = __ldg(&edge_owner[e - 1]);
      }
    };
    if (n_my > 0) prefetch(tile);
    for (int i = 0; i < n_my; ++i, tile += gridDim.x) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const float d = d_n;
      const int own = own_n, nb = nbr_n, prev = prev_n;
      if (i + 1 < n_my) prefetch(tile + gridDim.x);
      const bool valid = own >= 0;
      const float cut = valid ? cosine_cutoff_fast(d, pi_over_rc, rc) : 0.f;
      // segment boundary inside the tile: this edge starts a new owner's run (tile-local edge 0 is the "head")
      const uint32_t bmask = __ballot_sync(0xffffffffu, valid && tid > 0 && own != prev);
      trace_stamp(0, i, 0, tid == 0);
      mbar_wait_guard(bar(B_META_EMPTY + ms), mph ^ 1);
      mbar_wait_guard(bar(B_RBF_EMPTY + s), ph ^ 1);
      trace_stamp(0, i, 1, tid == 0);
      reinterpret_cast<uint32_t*>(smem + O_XOFF + ms * TILE * 4)[tid] = (uint32_t)nb * (uint32_t)(NF * 4);
      reinterpret_cast<__half*>(smem + O_CUT + ms * TILE * 4)[tid] = __float2half_rn(cut);   // fp16: consumed as half2 pairs
      reinterpret_cast<int*>(smem + O_OWN + ms * TILE * 4)[tid] = own;
      if (tid == 0) *reinterpret_cast<int*>(smem + O_HEAD + ms * 32) = prev;
      if (lane == 0) reinterpret_cast<uint32_t*>(smem + O_HEAD + ms * 32 + 16)[warp] = bmask;
      mbar_arrive(bar(B_META_FULL + ms));
      write_rbf_row_rec(smem + O_RBF + s * (128 * 128), sCen, tid, d, cut, g2, rrec);
      fence_async_smem();
      mbar_arrive(bar(B_RBF_FULL + s));
      trace_stamp(0, i, 2, tid == 0);
    }
  } else if (warp == 4) {
    // =========================================================== M: MMA issuer (one thread)
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
      constexpr uint32_t IDESC2 = idesc_f16(128, 128, 0, 1);
      const uint64_t dA1 = smem_desc_sw128(sbase + O_WF0, 16, 1024);
      const uint64_t dA2 = smem_desc_sw128(sbase + O_WF1, 16, 1024);
      auto issue1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        trace_stamp(4, i, 0, true);
        mbar_wait_guard(bar(B_RBF_FULL + s), ph);
        mbar_wait_guard(bar(B_D1_EMPTY + s), ph ^ 1);
        trace_stamp(4, i, 1, true);
        fence_after_sync();
        const uint64_t dB1 = smem_desc_sw128(sbase + O_RBF + s * (128 * 128), 16, 1024);
#pragma unroll
        for (int k = 0; k < RP / 16; ++k) mma_f16(tmem + s * 128, dA1 + 2 * k, dB1 + 2 * k, IDESC1, k > 0);
        mma_commit(bar(B_RBF_EMPTY + s));
        mma_commit(bar(B_D1_FULL + s));
      };
      auto issue2 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        trace_stamp(5, i, 0, true);
        mbar_wait_guard(bar(B_TT_FULL + s), ph);
        mbar_wait_guard(bar(B_D2_EMPTY + s), ph ^ 1);
        trace_stamp(5, i, 1, true);
        fence_after_sync();
        const uint64_t dB2 = smem_desc_sw128(sbase + O_TT + s * (2 * 128 * 128), 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16(tmem + 256 + s * 128, dA2 + (uint64_t)((k >> 2) * (128 * 128 / 16) + (k & 3) * 2),
                  dB2 + (uint64_t)(k * (2048 / 16)), IDESC2, k > 0);
        mma_commit(bar(B_TT_EMPTY + s));
        mma_commit(bar(B_D2_FULL + s));
      };
      issue1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) issue1(i + 1);
        issue2(i);
      }
    }
  } else if (warp < 9) {
    // =========================================================== T: tanh (thread = feature j)
    const int j = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const float bias = sBias[j];
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1;
      trace_stamp(1, i, 0, j == 0);
      mbar_wait_guard(bar(B_D1_FULL + s), ph);
      mbar_wait_guard(bar(B_TT_EMPTY + s), ph ^ 1);
      trace_stamp(1, i, 1, j == 0);
      fence_after_sync();
      uint8_t* sTT = smem + O_TT + s * (2 * 128 * 128);
      const uint4* sCutH = reinterpret_cast<const uint4*>(smem + O_CUT + ms * TILE * 4);   // 8 fp16 cut-offs per load
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + s * 128 + lane_sel + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // t * C(d_e): D2 is linear in t, so the cut-off rides through the second GEMM for free.  tanh on packed fp16
          // pairs: one MUFU operation and one HMUL2 per two values (the result is an fp16 operand anyway)
          const uint4 cq = sCutH[c * 4 + q];
          const uint32_t cc[4] = {cq.x, cq.y, cq.z, cq.w};
          uint32_t p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            p[u] = hmul2_u32(tanh_approx_h2(pack_half2(__uint_as_float(r[q * 8 + 2 * u]) + bias,
                                                       __uint_as_float(r[q * 8 + 2 * u + 1]) + bias)), cc[u]);
          const int chunk = c * 4 + q;
          *reinterpret_cast<uint4*>(sTT + (chunk >> 3) * (128 * 128) + sw128_off(j, chunk & 7)) =
              make_uint4(p[0], p[1], p[2], p[3]);
        }
      }
      fence_before_sync();
      mbar_arrive(bar(B_D1_EMPTY + s));
      mbar_arrive(bar(B_META_EMPTY + ms));
      fence_async_smem();
      mbar_arrive(bar(B_TT_FULL + s));
      trace_stamp(1, i, 2, j == 0);
    }
  } else {
    // =========================================================== E0 / E1: epilogue (thread = feature f)
    const int g = warp < 13 ? 0 : 1;
    const int f = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const float* __restrict__ xf = x + f;
    for (int i = g; i < n_my; i += 2) {
      const int ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const int tile = blockIdx.x + i * gridDim.x;
      const uint4* sOff4 = reinterpret_cast<const uint4*>(smem + O_XOFF + ms * TILE * 4);
      const int* sOwn = reinterpret_cast<const int*>(smem + O_OWN + ms * TILE * 4);
      const uint32_t* sMask = reinterpret_cast<const uint32_t*>(smem + O_HEAD + ms * 32 + 16);
      // gather of x rows for 16 edges: independent of the GEMMs, so the first chunk is in flight before D2 is ready
      auto gather16 = [&](float (&xv)[16], int c16) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 o = sOff4[c16 * 4 + u];
          xv[4 * u] = ldg_byte_off(xf, o.x);
          xv[4 * u + 1] = ldg_byte_off(xf, o.y);
          xv[4 * u + 2] = ldg_byte_off(xf, o.z);
          xv[4 * u + 3] = ldg_byte_off(xf, o.w);
        }
      };
      trace_stamp(2 + g, i, 0, f == 0);
      mbar_wait_guard(bar(B_META_FULL + ms), mph);
      float xa[16], xb[16];
      gather16(xa, 0);
      int cur = sOwn[0];
      bool head = *reinterpret_cast<const int*>(smem + O_HEAD + ms * 32) == cur;  // run began in an earlier tile
      float acc = 0.f;
      auto flush = [&](int next_owner) {
        if (head) part[(size_t)tile * NF + f] = acc;
        else out[(size_t)cur * NF + f] = acc;
        head = false;
        cur = next_owner;
        acc = 0.f;
      };
      auto consume16 = [&](const uint32_t (&r)[16], const float (&xv)[16], int c16) {
        const uint32_t bits = (sMask[c16 >> 1] >> ((c16 & 1) * 16)) & 0xffffu;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (((bits >> (8 * q)) & 0xffu) == 0u) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fmaf(__uint_as_float(r[q * 8 + u]), xv[q * 8 + u], acc);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if ((bits >> (8 * q + u)) & 1u) flush(sOwn[c16 * 16 + q * 8 + u]);
              acc = fmaf(__uint_as_float(r[q * 8 + u]), xv[q * 8 + u], acc);
            }
          }
        }
      };
      mbar_wait_guard(bar(B_D2_FULL + g), ph);
      trace_stamp(2 + g, i, 1, f == 0);
      fence_after_sync();
      const uint32_t d2 = tmem + 256 + g * 128 + lane_sel;
#pragma unroll 1
      for (int c16 = 0; c16 < 8; c16 += 2) {
        uint32_t r[16];
        tmem_ld16(d2 + c16 * 16, r);
        gather16(xb, c16 + 1);
        tmem_ld_wait();
        consume16(r, xa, c16);
        tmem_ld16(d2 + (c16 + 1) * 16, r);
        if (c16 + 2 < 8) gather16(xa, c16 + 2);
        tmem_ld_wait();
        consume16(r, xb, c16 + 1);
      }
      fence_before_sync();
      mbar_arrive(bar(B_D2_EMPTY + g));
      mbar_arrive(bar(B_META_EMPTY + ms));
      flush(0);
      trace_stamp(2 + g, i, 2, f == 0);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// =================================================================================================
// Backward (see fmd_filter_tc.cu for the math), warp-specialised:
//   PE warps 0-3   thread-per-edge: produce(i) = metadata + rbf row; e4(i-2) = D4 -> g_d
//   G0 warps 4-7   even tiles: thread-per-feature row f of gW0^T = a[nbr,f] * g_m[owner,f]; then, thread-per-edge,
//                  D1^T -> t row (stash)
//   G1 warps 8-11  same for odd tiles
//   T  warps 12-15 thread-per-EDGE: D3^T -> g_t row (in place over t) + in-thread cut-off term sum
//   M  warps 16-18 one MMA-issuer thread each: D13[s] = Wf0.rbf^T | D13[s] = Wf1^T.gW0^T | D4[q] = g_t.Wf0
constexpr uint32_t BO_WF0 = 0;
constexpr uint32_t BO_WF1 = BO_WF0 + 128 * 128;
constexpr uint32_t BO_RBF = BO_WF1 + 2 * 128 * 128;          // 2 x 16 KB
constexpr uint32_t BO_OP = BO_RBF + 2 * 128 * 128;           // 2 x 32 KB: gW0^T (B operand of MMA3)
constexpr uint32_t BO_ST = BO_OP + 2 * 2 * 128 * 128;        // 2 x 32 KB: t^T stash, overwritten in place by g_t^T (A operand of MMA4)
constexpr uint32_t BO_META = BO_ST + 2 * 2 * 128 * 128;      // 4 x 1 KB
constexpr uint32_t BO_OWN = BO_META + META_STAGES * TILE * 8;
constexpr uint32_t BO_HEAD = BO_OWN + META_STAGES * TILE * 4;  // 4 x 16 B boundary masks
constexpr uint32_t BO_RED = BO_HEAD + META_STAGES * 16;      // 4 x [128] floats: cut-off term sum per edge
constexpr uint32_t BO_BIAS = BO_RED + 4 * TILE * 4;
constexpr uint32_t BO_CEN = BO_BIAS + NF * 4;
constexpr uint32_t BO_BAR = BO_CEN + RP * 4;
constexpr uint32_t BSMEM = BO_BAR + 40 * 8 + 16;
constexpr uint32_t BSMEM_ALLOC = BSMEM + 1024;
static_assert(BSMEM_ALLOC <= 232448, "backward kernel exceeds the 227 KB shared-memory limit");

enum { C_RBF_FULL = 0, C_RBF_EMPTY = 2, C_D1_FULL = 4, C_D1_EMPTY = 6, C_GW_FULL = 8, C_D3_FULL = 10, C_D3_EMPTY = 12,
       C_GT_FULL = 14, C_OP_EMPTY = 16, C_D4_FULL = 18, C_D4_EMPTY = 22, C_META_FULL = 26, C_META_EMPTY = 30,
       C_ST_EMPTY = 34, C_ST_FULL = 36, C_COUNT = 38 };
constexpr int D4_STAGES = 4;   // D4 (64 columns) is quadruple-buffered so that e4 may lag 4 tiles behind produce

__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }

template <bool kExact>
__global__ void __launch_bounds__(BWD_THREADS, 1)
filter_cfconv_bwd2_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                          const int32_t* __restrict__ edge_nbr, int capacity,
                          const int32_t* __restrict__ n_edges_dev, const __half* __restrict__ wf0,
                          const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                          const float* __restrict__ centers, int R, float gamma, float rc,
                          const float* __restrict__ a, const float* __restrict__ g_m, float* __restrict__ g_d,
                          int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sBias = reinterpret_cast<float*>(smem + BO_BIAS);
  float* sCen = reinterpret_cast<float*>(smem + BO_CEN);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + BO_BAR + 40 * 8);
  auto bar = [&](int i) { return sbase + BO_BAR + 8u * (uint32_t)i; };

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  load_weight_kmajor(smem + BO_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + BO_WF1, wf1, NF, NF / 8);
  if (tid < NF) sBias[tid] = bf0 ? __half2float(bf0[tid]) : 0.f;
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(C_RBF_FULL + i), 128);
      mbar_init(bar(C_RBF_EMPTY + i), 1);
      mbar_init(bar(C_D1_FULL + i), 1);
      mbar_init(bar(C_D1_EMPTY + i), 128);
      mbar_init(bar(C_GW_FULL + i), 128);
      mbar_init(bar(C_D3_FULL + i), 1);
      mbar_init(bar(C_D3_EMPTY + i), 128);
      mbar_init(bar(C_GT_FULL + i), 128);
      mbar_init(bar(C_OP_EMPTY + i), 1);
      mbar_init(bar(C_ST_EMPTY + i), 1);
      mbar_init(bar(C_ST_FULL + i), 128);
    }
    for (int i = 0; i < D4_STAGES; ++i) {
      mbar_init(bar(C_D4_FULL + i), 1);
      mbar_init(bar(C_D4_EMPTY + i), 128);
    }
    for (int i = 0; i < META_STAGES; ++i) {
      mbar_init(bar(C_META_FULL + i), 128);
      mbar_init(bar(C_META_EMPTY + i), 256);  // the G group and T both consume the metadata
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + BO_BAR + 40 * 8, 512);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: D13[s] at s*128 (s = tile & 1), D4[q] at 256 + q*64 (q = tile & 3)

  if (warp < 4) {
    // =========================================================== PE: produce(i), then e4(i-2)
    const float g2 = gamma * 1.4426950408889634f;
    const float pi_over_rc = FMD_PI_F / rc;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const RbfRecurrence rrec = make_rbf_recurrence(sCen, R, g2);
    int tile = blockIdx.x;
    float d_n = 0.f;
    int own_n = -1, nbr_n = 0, prev_n = -1;
    auto prefetch = [&](int t) {
      const int e = t * TILE + tid;
      d_n = 0.f; own_n = -1; nbr_n = 0; prev_n = -1;
      if (e < E) {
        d_n = __ldg(&dist[e]);
        own_n = __ldg(&edge_owner[e]);
        nbr_n = __ldg(&edge_nbr[e]);
        if (e > 0) prev_n = __ldg(&edge_owner[e - 1]);
      }
    };
    float dq[D4_STAGES] = {0.f, 0.f, 0.f, 0.f}, cq[D4_STAGES] = {0.f, 0.f, 0.f, 0.f};  // (dist, cut) of tiles i-1 .. i-4
    auto e4 = [&](int i, float d, float cut) {
      const int s = i & (D4_STAGES - 1);
      const uint32_t ph = (i / D4_STAGES) & 1;
      const int e = (blockIdx.x + i * gridDim.x) * TILE + tid;
      float prev_gd = 0.f;
      if (accumulate && e < E) prev_gd = g_d[e];
      trace_stamp(1, i, 0, tid == 0);
      mbar_wait_guard(bar(C_D4_FULL + s), ph);
      trace_stamp(1, i, 1, tid == 0);
      fence_after_sync();
      const float dcut = d < rc ? -0.5f * pi_over_rc * __sinf(d * pi_over_rc) : 0.f;
      const float two_g = 2.0f * gamma;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + 256 + s * 64 + lane_sel + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float diff = d - sCen[c * 32 + u];
          const float ex = ex2_approx(g2 * diff * diff);
          // d/dd [exp(gamma diff^2) C(d)] ; columns k >= R hold exact zeros (zero-padded Wf0)
          acc = fmaf(__uint_as_float(r[u]), ex * fmaf(two_g * diff, cut, dcut), acc);
        }
      }
      if (kExact) {
        acc += dcut * reinterpret_cast<const float*>(smem + BO_RED + s * (TILE * 4))[tid];
      }
      fence_before_sync();
      mbar_arrive(bar(C_D4_EMPTY + s));
      if (e < E) g_d[e] = prev_gd + acc;
      trace_stamp(1, i, 2, tid == 0);
    };
    if (n_my > 0) prefetch(tile);
    for (int i = 0; i < n_my; ++i, tile += gridDim.x) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const float d = d_n;
      const int own = own_n, nb = nbr_n, prev = prev_n;
      if (i + 1 < n_my) prefetch(tile + gridDim.x);
      const bool valid = own >= 0;
      const float cut = valid ? cosine_cutoff_fast(d, pi_over_rc, rc) : 0.f;
      const uint32_t bmask = __ballot_sync(0xffffffffu, valid && tid > 0 && own != prev);
      trace_stamp(0, i, 0, tid == 0);
      mbar_wait_guard(bar(C_META_EMPTY + ms), mph ^ 1);
      mbar_wait_guard(bar(C_RBF_EMPTY + s), ph ^ 1);
      trace_stamp(0, i, 1, tid == 0);
      reinterpret_cast<uint2*>(smem + BO_META + ms * TILE * 8)[tid] =
          make_uint2((uint32_t)nb * (uint32_t)(NF * 4), __float_as_uint(cut));   // byte offset of the gathered row
      reinterpret_cast<int*>(smem + BO_OWN + ms * TILE * 4)[tid] = valid ? own : 0;
      if (lane == 0) reinterpret_cast<uint32_t*>(smem + BO_HEAD + ms * 16)[warp] = bmask;
      write_rbf_row_rec(smem + BO_RBF + s * (128 * 128), sCen, tid, d, cut, g2, rrec);
      fence_async_smem();
      mbar_arrive(bar(C_RBF_FULL + s));
      mbar_arrive(bar(C_META_FULL + ms));
      trace_stamp(0, i, 2, tid == 0);
      if (i >= D4_STAGES) e4(i - D4_STAGES, dq[D4_STAGES - 1], cq[D4_STAGES - 1]);
#pragma unroll
      for (int k = D4_STAGES - 1; k > 0; --k) {
        dq[k] = dq[k - 1];
        cq[k] = cq[k - 1];
      }
      dq[0] = d;
      cq[0] = cut;
    }
    // drain: tiles n_my-4 .. n_my-1 sit in dq[3] .. dq[0]
#pragma unroll
    for (int k = D4_STAGES - 1; k >= 0; --k)
      if (n_my - 1 - k >= 0) e4(n_my - 1 - k, dq[k], cq[k]);
  } else if (warp < 12) {
    // =========================================================== G0 / G1: gW0^T rows (thread = feature f)
    const int g = warp < 8 ? 0 : 1;
    const int f = (warp & 3) * 32 + lane;
    const float* __restrict__ af = a + f;
    const float* __restrict__ gmf = g_m + f;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    // tanh phase of the same tile (D1 -> t -> stash), done by the gather group: these warps idle ~60 % of a tile
    // period waiting for buffers, while the dedicated T warps are the longest role (timeline trace)
    // In the backward everything except the gather is computed with a TMEM lane (= thread) per EDGE:
    // D1^T[e,j], D3^T[e,j], D4[e,k].  Thread e then owns row e of t / g_t (K-major A operand of the last GEMM,
    // written in place), C(d_e) is a per-thread scalar and the exact cut-off term sum_j t_j D3_j is an in-thread
    // FMA chain - no cross-lane reduction at all (the feature-per-lane variant spent 1/3 of its time in shuffles).
    auto phaseA = [&](int i) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      uint8_t* sT = smem + BO_ST + s * (2 * 128 * 128);
      trace_stamp(4, i, 0, f == 0);
      mbar_wait_guard(bar(C_D1_FULL + s), ph);
      mbar_wait_guard(bar(C_ST_EMPTY + s), ph ^ 1);   // MMA4(i-2) has consumed g_t from sT[s]
      trace_stamp(4, i, 1, f == 0);
      fence_after_sync();
      const float4* sBias4 = reinterpret_cast<const float4*>(sBias);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + s * 128 + lane_sel + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ba = sBias4[c * 8 + q * 2], bb = sBias4[c * 8 + q * 2 + 1];
          const float bj[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            p[u] = tanh_approx_h2(pack_half2(__uint_as_float(r[q * 8 + 2 * u]) + bj[2 * u],
                                             __uint_as_float(r[q * 8 + 2 * u + 1]) + bj[2 * u + 1]));
          const int chunk = c * 4 + q;   // 8 consecutive features j of edge row f (= this thread's edge)
          *reinterpret_cast<uint4*>(sT + (chunk >> 3) * (128 * 128) + sw128_off(f, chunk & 7)) =
              make_uint4(p[0], p[1], p[2], p[3]);
        }
      }
      fence_before_sync();
      mbar_arrive(bar(C_D1_EMPTY + s));
      mbar_arrive(bar(C_ST_FULL + s));
      trace_stamp(4, i, 2, f == 0);
    };
    for (int i = g; i < n_my; i += 2) {
      const int ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const uint4* sMeta2 = reinterpret_cast<const uint4*>(smem + BO_META + ms * TILE * 8);
      const int* sOwn = reinterpret_cast<const int*>(smem + BO_OWN + ms * TILE * 4);
      const uint32_t* sMask = reinterpret_cast<const uint32_t*>(smem + BO_HEAD + ms * 16);
      uint8_t* sOp = smem + BO_OP + g * (2 * 128 * 128);
      trace_stamp(2 + g, i, 0, f == 0);
      mbar_wait_guard(bar(C_META_FULL + ms), mph);
      float gm = __ldg(gmf + (size_t)sOwn[0] * NF);
      // Two half-tiles of 64 edges: all 64 row gathers of a half are in flight at once (the latency of the L2
      // gathers is paid twice per tile instead of once per 16 rows: 8.6k -> ~4k cycles in the timeline trace); the
      // first half is issued before the operand buffer is even free.
      float av[64];
      auto gather64 = [&](int h) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const uint4 m = sMeta2[h * 32 + u];
          av[2 * u] = ldg_byte_off(af, m.x);
          av[2 * u + 1] = ldg_byte_off(af, m.z);
        }
      };
      auto emit64 = [&](int h) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t bits = sMask[h * 2 + w];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* v = av + w * 32 + q * 8;
            if (((bits >> (8 * q)) & 0xffu) == 0u) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] *= gm;
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if ((bits >> (8 * q + u)) & 1u) gm = __ldg(gmf + (size_t)sOwn[h * 64 + w * 32 + q * 8 + u] * NF);
                v[u] *= gm;
              }
            }
            const int chunk = h * 8 + w * 4 + q;
            *reinterpret_cast<uint4*>(sOp + (chunk >> 3) * (128 * 128) + sw128_off(f, chunk & 7)) =
                make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
          }
        }
      };
      gather64(0);
      mbar_wait_guard(bar(C_OP_EMPTY + g), ph ^ 1);
      trace_stamp(2 + g, i, 1, f == 0);
      emit64(0);
      gather64(1);
      emit64(1);
      fence_async_smem();
      mbar_arrive(bar(C_GW_FULL + g));
      mbar_arrive(bar(C_META_EMPTY + ms));
      trace_stamp(2 + g, i, 2, f == 0);
      phaseA(i);
    }
  } else if (warp < 16) {
    // =========================================================== T (thread = feature j)
    const int j = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    auto phaseB = [&](int i) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1;
      // thread j here is EDGE row j of the tile.  g_t overwrites t IN PLACE (same thread, same address) and the
      // buffer then is the K-major A operand of MMA4; the gW0 buffer sOp[s] is free as soon as MMA3 has read it.
      uint8_t* sT = smem + BO_ST + s * (2 * 128 * 128);
      const int q4 = i & (D4_STAGES - 1);
      float* red = reinterpret_cast<float*>(smem + BO_RED + q4 * (TILE * 4));
      trace_stamp(5, i, 0, j == 0);
      mbar_wait_guard(bar(C_ST_FULL + s), ph);     // t of this tile written by the gather/tanh group
      mbar_wait_guard(bar(C_D3_FULL + s), ph);
      if (kExact) mbar_wait_guard(bar(C_D4_EMPTY + q4), ((i / D4_STAGES) & 1) ^ 1);  // e4(i-4) has read sRed[q4]
      trace_stamp(5, i, 1, j == 0);
      fence_after_sync();
      const float cut = __uint_as_float(reinterpret_cast<const uint2*>(smem + BO_META + ms * TILE * 8)[j].y);
      float usum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + s * 128 + lane_sel + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = c * 4 + q;
          uint4* slot = reinterpret_cast<uint4*>(sT + (chunk >> 3) * (128 * 128) + sw128_off(j, chunk & 7));
          const uint4 tq = *slot;
          const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
          uint32_t p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = q * 8 + 2 * u;
            const float2 tt = unpack_h2(tw[u]);
            const float d0 = __uint_as_float(r[k]), d1 = __uint_as_float(r[k + 1]);
            if (kExact) {
              usum = fmaf(tt.x, d0, usum);
              usum = fmaf(tt.y, d1, usum);
            }
            p[u] = pack_half2(cut * d0 * fmaf(-tt.x, tt.x, 1.f), cut * d1 * fmaf(-tt.y, tt.y, 1.f));
          }
          *slot = make_uint4(p[0], p[1], p[2], p[3]);
        }
      }
      if (kExact) red[j] = usum;
      fence_before_sync();
      mbar_arrive(bar(C_D3_EMPTY + s));
      fence_async_smem();
      mbar_arrive(bar(C_GT_FULL + s));
      mbar_arrive(bar(C_META_EMPTY + ms));
      trace_stamp(5, i, 2, j == 0);
    };
    for (int i = 0; i < n_my; ++i) phaseB(i);
  } else {
    // =========================================================== M: MMA issuer
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);   // D1^T[e,j]: A = rbf tile (K-major), B = Wf0 (K-major)
      constexpr uint32_t IDESC3 = idesc_f16(128, 128, 1, 1);   // D3^T[e,j]: A = gW0^T [f][e] (MN-major), B = Wf1 [f][j] (MN-major)
      constexpr uint32_t IDESC4 = idesc_f16(128, 64, 0, 1);    // D4[e,k]:   A = g_t [e][j] (K-major),   B = Wf0 [j][k] (MN-major)
      const uint64_t dW0k = smem_desc_sw128(sbase + BO_WF0, 16, 1024);          // Wf0 as K-major operand (rows j)
      const uint64_t dW1mn = smem_desc_sw128(sbase + BO_WF1, 128 * 128, 1024);  // Wf1 [f][j] read MN-major (N = j)
      const uint64_t dB4 = smem_desc_sw128(sbase + BO_WF0, 16, 1024);           // Wf0 [j][k] read MN-major (N = k)
      auto issue1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        trace_stamp(6, i, 0, true);
        mbar_wait_guard(bar(C_RBF_FULL + s), ph);
        mbar_wait_guard(bar(C_D3_EMPTY + s), ph ^ 1);  // D13[s] drained by phaseB(i-2)
        trace_stamp(6, i, 1, true);
        fence_after_sync();
        const uint64_t dB1 = smem_desc_sw128(sbase + BO_RBF + s * (128 * 128), 16, 1024);
#pragma unroll
        for (int k = 0; k < RP / 16; ++k) mma_f16(tmem + s * 128, dB1 + 2 * k, dW0k + 2 * k, IDESC1, k > 0);
        mma_commit(bar(C_RBF_EMPTY + s));
        mma_commit(bar(C_D1_FULL + s));
      };
      auto issue3 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        trace_stamp(7, i, 0, true);
        mbar_wait_guard(bar(C_GW_FULL + s), ph);
        trace_stamp(7, i, 2, true);
        mbar_wait_guard(bar(C_D1_EMPTY + s), ph);  // phaseA(i) has drained D1 from D13[s]
        trace_stamp(7, i, 1, true);
        fence_after_sync();
        const uint64_t dB3 = smem_desc_sw128(sbase + BO_OP + s * (2 * 128 * 128), 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16(tmem + s * 128, dB3 + (uint64_t)(k * (2048 / 16)), dW1mn + (uint64_t)(k * (2048 / 16)), IDESC3, k > 0);
        mma_commit(bar(C_OP_EMPTY + s));
        mma_commit(bar(C_D3_FULL + s));
      };
      auto issue4 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int q4 = i & (D4_STAGES - 1);
        trace_stamp(8, i, 0, true);
        mbar_wait_guard(bar(C_GT_FULL + s), ph);
        mbar_wait_guard(bar(C_D4_EMPTY + q4), ((i / D4_STAGES) & 1) ^ 1);
        trace_stamp(8, i, 1, true);
        fence_after_sync();
        const uint64_t dA4 = smem_desc_sw128(sbase + BO_ST + s * (2 * 128 * 128), 16, 1024);   // g_t [e][j], K-major
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16(tmem + 256 + q4 * 64, dA4 + (uint64_t)((k >> 2) * (128 * 128 / 16) + (k & 3) * 2),
                  dB4 + (uint64_t)(k * (2048 / 16)), IDESC4, k > 0);
        mma_commit(bar(C_ST_EMPTY + s));
        mma_commit(bar(C_D4_FULL + q4));
      };
      // One issuer warp per GEMM: an issuer blocks in program order on its mbarriers, and with a single issuer
      // MMA3(i) queued behind MMA1(i+1) (which waits for the previous tile's phase B) - 3.8k idle cycles per tile
      // in the timeline trace.  tcgen05.commit tracks the MMAs of the issuing thread only, which is what we want.
      if (warp == 16) {
        for (int i = 0; i < n_my; ++i) issue1(i);
      } else if (warp == 17) {
        for (int i = 0; i < n_my; ++i) issue3(i);
      } else {
        for (int i = 0; i < n_my; ++i) issue4(i);
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

// defined in fmd_filter_tc.cu
void fmd_cfconv_fixup_launch(const int32_t* seg_ptr, int n_nodes, int capacity, const float* part, float* out,
                             cudaStream_t st);

extern "C" int fmd_filter_cfconv_fwd2(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                                      const int32_t* seg_ptr, int n_nodes, int capacity, const int32_t* n_edges_dev,
                                      const void* wf0_h, const void* bf0_h, const void* wf1_h, const float* centers,
                                      int num_rbf, float gamma, float rc, const float* x, int n_feat, float* out,
                                      float* part, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && seg_ptr && wf0_h && wf1_h && centers && x && out && part,
              "fmd_filter_cfconv_fwd2: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf <= RP, "fmd_filter_cfconv_fwd2: needs F == 128 and num_rbf <= 64");
  FMD_REQUIRE(n_nodes < (1 << 23), "fmd_filter_cfconv_fwd2: gathered rows are addressed with 32-bit byte offsets (< 2^23 nodes)");
  if (n_nodes <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done = false;
  if (!attr_done) {
    FMD_CUDA(cudaFuncSetAttribute(filter_cfconv_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SMEM2_ALLOC));
    attr_done = true;
  }
  if (capacity > 0) {
    const int max_tiles = fmd_div_up(capacity, TILE);
    const int grid = max_tiles < fmd_num_sms() ? max_tiles : fmd_num_sms();
    filter_cfconv_fwd2_kernel<<<grid, NTHREADS, SMEM2_ALLOC, st>>>(
        dist, edge_owner, edge_nbr, capacity, n_edges_dev, (const __half*)wf0_h, (const __half*)bf0_h,
        (const __half*)wf1_h, centers, num_rbf, gamma, rc, x, out, part);
    FMD_CHECK_LAUNCH();
  }
  fmd_cfconv_fixup_launch(seg_ptr, n_nodes, capacity, part, out, st);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_filter_cfconv_bwd2(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                                      int capacity, const int32_t* n_edges_dev, const void* wf0_h, const void* bf0_h,
                                      const void* wf1_h, const float* centers, int num_rbf, float gamma, float rc,
                                      const float* a, const float* g_m, int n_feat, float* g_d, int accumulate,
                                      int exact_cutoff_grad, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && wf0_h && wf1_h && centers && a && g_m && g_d,
              "fmd_filter_cfconv_bwd2: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf <= RP, "fmd_filter_cfconv_bwd2: needs F == 128 and num_rbf <= 64");
  if (capacity <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ex = exact_cutoff_grad ? 1 : 0;
  auto kern = ex ? filter_cfconv_bwd2_kernel<true> : filter_cfconv_bwd2_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[ex]) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BSMEM_ALLOC));
    attr_done[ex] = true;
  }
  const int max_tiles = fmd_div_up(capacity, TILE);
  const int grid = max_tiles < fmd_num_sms() ? max_tiles : fmd_num_sms();
  kern<<<grid, BWD_THREADS, BSMEM_ALLOC, st>>>(dist, edge_owner, edge_nbr, capacity, n_edges_dev, (const __half*)wf0_h,
                                            (const __half*)bf0_h, (const __half*)wf1_h, centers, num_rbf, gamma, rc, a,
                                            g_m, g_d, accumulate);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

// tools only: enable/disable the role timeline trace of the pipelined kernels (buffer: 9 * 64 * 3 uint64)
extern "C" int fmd_debug_set_trace(void* device_buffer) {
  unsigned long long* p = (unsigned long long*)device_buffer;
  FMD_CUDA(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  return FMD_OK;
}
