// Fused backward of one interaction block's edge part on the W16A16 path, over UNDIRECTED PAIRS p = (own < nbr):
//
//   g[p] (+)= d/dd_p  sum_f W(d_p)[f] * C(d_p) * ( g_m[own,f] a[nbr,f] + g_m[nbr,f] a[own,f] )
//
// i.e. the sum of the directed-edge gradients g_d[e] + g_d[rev e] of the reference's backward: the filter depends on
// the distance alone (W(e) == W(rev e)), and the forces only need that sum, so the kernel runs over half the tiles.
//
// replaces the edge part of FusedCSRCFConvFunction.backward (kernels/csr_kernels.py:857-912: fused_grad_filter_out,
// kernels/cfconv_kernels.py:178-337), the backward GEMMs of the fp16 filter network (LinearFP16ToFP16Function /
// FusedLinearTanhFP16Function, kernels/cfconv_kernels.py:963-1226, 1329-1434) and the first half of
// FusedDistanceGaussianRBFCutoffFunction.backward (:1679-1735) by ONE tcgen05 kernel; gW, g_t, g_rbf never reach HBM.
//
// Per 128-edge tile, a TMEM lane (= one thread) per EDGE everywhere:
//   D1^T[e,j] = rbf[e,:] . Wf0[j,:] + b_j          (bias through a constant-1 column)      MMA1
//   t[e,j]    = tanh(D1^T)                           row e of the stash, fp16
//   gW0[e,f]  = a[nbr,f] g_m[own,f] + g_m[nbr,f] a[own,f]   row e, fp16 (e = pair index in the tile)
//   D3^T[e,j] = sum_f gW0[e,f] Wf1[f,j]                                                     MMA3
//   g_t[e,j]  = C(d_e) D3^T (1 - t^2)                in place over t (K-major A operand)
//   D4[e,k]   = sum_j g_t[e,j] Wf0[j,k]                                                     MMA4
//   g_d[e]   += sum_k D4[e,k] d rbf_k/dd  +  C'(d_e) sum_j t[e,j] D3^T[e,j]                 (second term: exact mode)
// The last sum equals sum_f a W g_m by linearity, so W is never needed in the backward.
//
// One persistent CTA per SM, 23 warps (every SIMT role is one serial dependency chain per tile, so the roles - not the
// issue slots - bound the tile period: five groups of four warps instead of four):
//   P  warps 0-3    metadata + radial-basis row
//   G  warps 4-7    the rows a[nbr,:] and g_m[nbr,:] (fp16, 256 B each) are staged with cp.async through a ring of four
//                   32-feature slots, one tile ahead, slot by slot into what the warp has just consumed (each warp stages
//                   and reads only ITS 32 rows: cp.async groups, no barrier); thread e reads ITS row with 16-byte
//                   shared-memory loads, combines it with the rows of its owner (broadcast 16-byte global loads) with
//                   packed HMUL2 / HFMA2 and writes the fp16 row with tcgen05.st into TENSOR MEMORY, where MMA3 reads it
//                   as its A operand.  (The kernel is bound by the shared-memory / L1TEX pipe - ncu:
//                   77 % of peak with the operand multiplied in place in shared memory and read from there by the MMA;
//                   per-lane row loads straight from global cost as much in L1TEX: 32 sectors per instruction.)
//   A  warps 8-11   D1^T -> tanh -> t row (stash)
//   B  warps 12-15  D3^T -> g_t row (in place over t) + in-thread cut-off term sum
//   E4 warps 16-19  D4 -> g_d (lags up to 4 tiles: D4 is quadruple-buffered)
//   M  warps 20-22  one MMA-issuer thread per GEMM (an issuer blocks in program order on its mbarriers)
// Software arrivals on the mbarriers are one per warp (lane 0 after __syncwarp).
// TMEM: D13[2] (D1^T then D3^T of the same tile) at s*128, D4[2] (64 columns) at 256 + q*64, gW0[2] (fp16 pairs, 64 columns)
// at 384 + s*64.
#include "fmd_filter_shared.cuh"

using namespace fmd;
using namespace fmd::tc;
using namespace fmd::filt;

namespace {

constexpr int BWD_THREADS = 23 * 32;   // 6 warps on three SM sub-partitions: 16384 / (6 * 32) -> 80 registers per thread
constexpr int META_STAGES = 4;
constexpr int D4_STAGES = 2;
constexpr uint32_t TM_D4 = 256, TM_GW = 384;   // tensor-memory column bases

constexpr uint32_t BO_WF0 = 0;
constexpr uint32_t BO_WF1 = BO_WF0 + 128 * 128;
constexpr uint32_t BO_RBF = BO_WF1 + 2 * 128 * 128;          // 2 x 16 KB
constexpr uint32_t BO_AS = BO_RBF + 2 * 128 * 128;           // 4 x 16 KB ring: slot q = 32 features of a[nbr] | g_m[nbr], [row][128 B] swizzled
constexpr uint32_t BO_ST = BO_AS + 4 * TILE * 128;           // 2 x 32 KB: t stash, overwritten in place by g_t (A operand of MMA4)
constexpr uint32_t BO_META = BO_ST + 2 * 2 * 128 * 128;      // 4 x 1 KB: {byte offset nbr * 256, C(d_e)}
constexpr uint32_t BO_OWN = BO_META + META_STAGES * TILE * 8;  // 4 x 512 B
constexpr uint32_t BO_RED = BO_OWN + META_STAGES * TILE * 4;   // 4 x [128] floats: cut-off term sum per edge
constexpr uint32_t BO_CEN = BO_RED + 4 * TILE * 4;
constexpr uint32_t BO_BAR = BO_CEN + RP * 4;
constexpr uint32_t BSMEM = BO_BAR + 48 * 8 + 16;
constexpr uint32_t BSMEM_ALLOC = BSMEM + 1024;
static_assert(BSMEM_ALLOC <= 232448, "backward kernel exceeds the 227 KB shared-memory limit");

enum { C_RBF_FULL = 0, C_RBF_EMPTY = 2, C_D1_FULL = 4, C_D1_EMPTY = 6, C_GW_FULL = 8, C_D3_FULL = 10, C_D3_EMPTY = 12,
       C_GT_FULL = 14, C_OP_EMPTY = 16, C_D4_FULL = 18, C_D4_EMPTY = 22, C_META_FULL = 26, C_META_EMPTY = 30,
       C_ST_EMPTY = 34, C_ST_FULL = 36,
       // second halves (columns 64-127 of the tile's 128 filter columns): D1 drained / t written / D3 ready
       C_D1E_HI = 38, C_STF_HI = 40, C_D3F_HI = 42, C_COUNT = 44 };

__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Role timeline (tools only, compiled out of the production instantiation): CTA 0 records clock64() stamps
// {wait start, work start, end} per role and tile into trace[role][tile < 64][3]  (scripts/trace_roles.py).
template <bool kTrace>
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int role, int tile, int k, bool leader) {
  if (kTrace) {
    if (blockIdx.x == 0 && leader && tile < 64) trace[(role * 64 + tile) * 3 + k] = (unsigned long long)clock64();
  }
}

// One 32-column chunk of  D3^T -> g_t  for this thread's edge row: g_t = D3 * C (1 - t^2) written in place over t (the
// K-major A operand of MMA4), and (exact mode) the running sum_j t_j D3_j of the cut-off term.
template <bool kExact>
__device__ __forceinline__ void gt_chunk(uint32_t d3, uint8_t* sT, uint32_t x7, uint32_t cut2, uint32_t ncut2, float& usum,
                                         int c) {
  uint32_t r[32];
  tmem_ld32(d3 + c * 32, r);
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = c * 4 + q;
    uint4* slot = reinterpret_cast<uint4*>(sT + (chunk >> 3) * (128 * 128) + ((((uint32_t)chunk & 7u) << 4) ^ x7));
    const uint4 tq = *slot;
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    uint32_t p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = q * 8 + 2 * u;
      // D3 pair rounded to fp16 once (it is an fp16 operand after the next multiply anyway)
      const uint32_t dh = pack_half2(__uint_as_float(r[k]), __uint_as_float(r[k + 1]));
      if (kExact) {   // sum_j t_j D3_j: fp16 x fp16 products are exact in fp32, fp32 accumulate (FHFMA)
        asm("{\n\t.reg .b16 tl, th, dl, dh;\n\tmov.b32 {tl, th}, %1;\n\tmov.b32 {dl, dh}, %2;\n\t"
            "fma.rn.f32.f16 %0, tl, dl, %0;\n\tfma.rn.f32.f16 %0, th, dh, %0;\n\t}" : "+f"(usum) : "r"(tw[u]), "r"(dh));
      }
      // the factor C (1 - t^2) as cut - (t cut) t in packed half arithmetic
      uint32_t wgt;
      asm("{\n\t.reg .b32 tc;\n\tmul.rn.f16x2 tc, %1, %3;\n\tfma.rn.f16x2 %0, tc, %1, %2;\n\t}"
          : "=r"(wgt) : "r"(tw[u]), "r"(cut2), "r"(ncut2));
      p[u] = hmul2_u32(dh, wgt);
    }
    *slot = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

template <bool kExact, bool kTrace>
__global__ void __maxnreg__(80)
filter_cfconv_bwd_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                          const int32_t* __restrict__ edge_nbr, int capacity,
                          const int32_t* __restrict__ n_edges_dev, const __half* __restrict__ wf0,
                          const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                          const float* __restrict__ centers, int R, float gamma, float rc,
                          const __half* __restrict__ a, const __half* __restrict__ g_m, float* __restrict__ g_d,
                          int accumulate, unsigned long long* __restrict__ trace) {
#define TR(role, tile, k, leader) trace_stamp<kTrace>(trace, role, tile, k, leader)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sCen = reinterpret_cast<float*>(smem + BO_CEN);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + BO_BAR + 48 * 8);
  auto bar = [&](int i) { return sbase + BO_BAR + 8u * (uint32_t)i; };

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  load_weight_kmajor(smem + BO_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + BO_WF1, wf1, NF, NF / 8);
  for (int idx = tid; idx < 2 * 128 * 128 / 16; idx += BWD_THREADS)
    reinterpret_cast<uint4*>(smem + BO_RBF)[idx] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  __syncthreads();
  // bias column: Wf0p[j, R] = b_j, multiplied by the constant-1 column R of every radial-basis row (MMA4 then produces a
  // meaningless column D4[:, R], which e4 never reads)
  if (tid < NF)
    *reinterpret_cast<__half*>(smem + BO_WF0 + sw128_off(tid, R >> 3) + (R & 7) * 2) = bf0 ? bf0[tid] : __float2half(0.f);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(C_RBF_FULL + i), 4);
      mbar_init(bar(C_RBF_EMPTY + i), 1);
      mbar_init(bar(C_D1_FULL + i), 1);
      mbar_init(bar(C_D1_EMPTY + i), 4);
      mbar_init(bar(C_GW_FULL + i), 4);
      mbar_init(bar(C_D3_FULL + i), 1);
      mbar_init(bar(C_D3_EMPTY + i), 4);
      mbar_init(bar(C_GT_FULL + i), 4);
      mbar_init(bar(C_OP_EMPTY + i), 1);
      mbar_init(bar(C_ST_EMPTY + i), 1);
      mbar_init(bar(C_ST_FULL + i), 4);
      mbar_init(bar(C_D1E_HI + i), 4);
      mbar_init(bar(C_STF_HI + i), 4);
      mbar_init(bar(C_D3F_HI + i), 1);
    }
    for (int i = 0; i < D4_STAGES; ++i) {
      mbar_init(bar(C_D4_FULL + i), 1);
      mbar_init(bar(C_D4_EMPTY + i), 4);
    }
    for (int i = 0; i < META_STAGES; ++i) {
      mbar_init(bar(C_META_FULL + i), 4);
      mbar_init(bar(C_META_EMPTY + i), 8);    // G (row offsets, owners) and B (cut-off) consume the metadata
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + BO_BAR + 48 * 8, 512);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =========================================================== P: metadata + radial-basis row (thread = edge row)
    const float g2 = gamma * 1.4426950408889634f;
    const float pi_over_rc = FMD_PI_F / rc;
    const RbfRecurrence rrec = make_rbf_recurrence(sCen, R, g2);
    const int nchunks = (R + 8) >> 3;          // chunks holding columns 0 .. R (radial basis + bias column)
    int tile = blockIdx.x;
    float d_n = 0.f;
    int own_n = -1, nbr_n = 0;
    auto prefetch = [&](int t) {
      const int e = t * TILE + tid;
      d_n = 0.f; own_n = -1; nbr_n = 0;
      if (e < E) {
        d_n = __ldg(&dist[e]);
        own_n = __ldg(&edge_owner[e]);
        nbr_n = __ldg(&edge_nbr[e]);
      }
    };
    if (n_my > 0) prefetch(tile);
    for (int i = 0; i < n_my; ++i, tile += gridDim.x) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const float d = d_n;
      const int own = own_n, nb = nbr_n;
      if (i + 1 < n_my) prefetch(tile + gridDim.x);
      const bool valid = own >= 0;
      const float cut = valid ? cosine_cutoff_fast(d, pi_over_rc, rc) : 0.f;
      TR(0, i, 0, tid == 0);
      mbar_wait_guard(bar(C_META_EMPTY + ms), mph ^ 1);
      reinterpret_cast<uint2*>(smem + BO_META + ms * TILE * 8)[tid] =
          make_uint2((uint32_t)nb * (uint32_t)(NF * 2), __float_as_uint(cut));   // byte offset of the gathered fp16 row
      reinterpret_cast<int*>(smem + BO_OWN + ms * TILE * 4)[tid] = valid ? own : 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C_META_FULL + ms));
      mbar_wait_guard(bar(C_RBF_EMPTY + s), ph ^ 1);
      TR(0, i, 1, tid == 0);
      {
        uint8_t* sRbf = smem + BO_RBF + s * (128 * 128);
#pragma unroll 2
        for (int c = 0; c < nchunks; ++c) {
          float vals[8];
          if (rrec.uniform) {
            const float x = d - sCen[c * 8];
            float v = ex2_approx(g2 * x * x) * cut;
            float q = ex2_approx(fmaf(x, rrec.a, rrec.b));
            vals[0] = v;
#pragma unroll
            for (int u = 1; u < 8; ++u) {
              v *= q;
              q *= rrec.cstep;
              vals[u] = v;
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float x = d - sCen[c * 8 + u];
              vals[u] = ex2_approx(g2 * x * x) * cut;
            }
          }
          *reinterpret_cast<uint4*>(sRbf + sw128_off(tid, c)) =
              make_uint4(pack_half2(vals[0], vals[1]), pack_half2(vals[2], vals[3]), pack_half2(vals[4], vals[5]), pack_half2(vals[6], vals[7]));
        }
        *reinterpret_cast<unsigned short*>(sRbf + sw128_off(tid, R >> 3) + (R & 7) * 2) = 0x3C00u;   // 1.0: bias column
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C_RBF_FULL + s));
      TR(0, i, 2, tid == 0);
    }
  } else if (warp < 8) {
    // =========================================================== G: gW0 rows (thread = pair row e of the tile)
    const int w = warp & 3;
    const int e = w * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
    const uint32_t e7 = (uint32_t)(e & 7);
    // copy instruction: 4 rows x {a | g_m} x four 16-byte chunks (64 + 64 bytes per row and slot)
    const int crow = lane >> 3, cseg = (lane >> 2) & 1, cch = lane & 3;
    const uint8_t* csrc = reinterpret_cast<const uint8_t*>(cseg ? g_m : a) + cch * 16;
    auto issue_slot = [&](int i, int q) {
      const int ms = i & (META_STAGES - 1);
      const uint2* sMeta = reinterpret_cast<const uint2*>(smem + BO_META + ms * TILE * 8) + w * 32 + crow;
      const uint32_t slot = sbase + BO_AS + (uint32_t)(q * (TILE * 128));
      // all row offsets first: the copies are `asm volatile` (ordered), a shared-memory load between two of them would
      // put its latency into every copy
      uint32_t off[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) off[k] = sMeta[4 * k].x;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = w * 32 + 4 * k + crow;
        cp_async16(slot + (uint32_t)(row * 128) + ((((uint32_t)(cseg * 4 + cch)) ^ ((uint32_t)row & 7u)) << 4),
                   csrc + off[k] + q * 64);
      }
    };
    if (n_my > 0) {
      mbar_wait_guard(bar(C_META_FULL + 0), 0);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        issue_slot(0, q);
        cp_async_commit();
      }
    }
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const bool has_next = i + 1 < n_my;
      TR(2, i, 0, e == 0);
      if (has_next) mbar_wait_guard(bar(C_META_FULL + ((i + 1) & (META_STAGES - 1))), ((i + 1) / META_STAGES) & 1);
      if (i >= 2) mbar_wait_guard(bar(C_OP_EMPTY + s), ((i - 2) >> 1) & 1);   // MMA3(i-2) has read gW0[s] (tensor memory)
      fence_after_sync();
      TR(2, i, 1, e == 0);
      const int own = reinterpret_cast<const int*>(smem + BO_OWN + ms * TILE * 4)[e];
      const uint4* gm_own = reinterpret_cast<const uint4*>(g_m + (size_t)own * NF);
      const uint4* a_own = reinterpret_cast<const uint4*>(a + (size_t)own * NF);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        // The owner rows (broadcast global loads, 64 + 64 bytes per slot) are a first touch in L1 for every slot: their L2
        // round trip was ~37 % of this role's time (ncu: long scoreboard on the first multiply).  They are issued first,
        // and the refill copies of the PREVIOUS slot (8 cp.async, ~55 issue cycles each) go out under that latency.
        uint4 gi[4], ai[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          gi[c] = __ldg(gm_own + q * 4 + c);
          ai[c] = __ldg(a_own + q * 4 + c);
        }
        if (q > 0) {
          if (has_next) issue_slot(i + 1, q - 1);   // the slot consumed last takes the same features of the next tile
          cp_async_commit();
        }
        cp_async_wait<3>();          // slot q of this tile (committed 4 groups ago) has landed ...
        __syncwarp();                // ... for every lane of this warp (the rows are private to the warp)
        const uint8_t* row = smem + BO_AS + q * (TILE * 128) + e * 128;
        uint32_t r[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 aj = *reinterpret_cast<const uint4*>(row + ((((uint32_t)c) ^ e7) << 4));
          const uint4 gj = *reinterpret_cast<const uint4*>(row + ((((uint32_t)(4 + c)) ^ e7) << 4));
          r[4 * c + 0] = hfma2_u32(aj.x, gi[c].x, hmul2_u32(gj.x, ai[c].x));
          r[4 * c + 1] = hfma2_u32(aj.y, gi[c].y, hmul2_u32(gj.y, ai[c].y));
          r[4 * c + 2] = hfma2_u32(aj.z, gi[c].z, hmul2_u32(gj.z, ai[c].z));
          r[4 * c + 3] = hfma2_u32(aj.w, gi[c].w, hmul2_u32(gj.w, ai[c].w));
        }
        tmem_st16(tmem + TM_GW + s * 64 + lane_sel + q * 16, r);
      }
      if (has_next) issue_slot(i + 1, 3);
      cp_async_commit();
      tmem_st_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(C_GW_FULL + s));
        mbar_arrive(bar(C_META_EMPTY + ms));
      }
      TR(2, i, 2, e == 0);
    }
    cp_async_wait<0>();
  } else if (warp < 12) {
    // =========================================================== A: D1^T -> tanh -> t row (thread = edge row e)
    const int e = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t rowoff = (uint32_t)((e >> 3) * 1024 + (e & 7) * 128), x7 = (uint32_t)(e & 7) << 4;
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      uint8_t* sT = smem + BO_ST + s * (2 * 128 * 128) + rowoff;
      TR(4, i, 0, e == 0);
      mbar_wait_guard(bar(C_D1_FULL + s), ph);
      mbar_wait_guard(bar(C_ST_EMPTY + s), ph ^ 1);   // MMA4(i-2) has consumed g_t from sT[s]
      TR(4, i, 1, e == 0);
      fence_after_sync();
      const uint32_t d1 = tmem + s * 128 + lane_sel;
      auto process = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            p[u] = pack_half2(tanh_approx(__uint_as_float(r[q * 8 + 2 * u])), tanh_approx(__uint_as_float(r[q * 8 + 2 * u + 1])));
          const int chunk = c * 4 + q;   // 8 consecutive features j of this thread's edge row
          *reinterpret_cast<uint4*>(sT + (chunk >> 3) * (128 * 128) + ((((uint32_t)chunk & 7u) << 4) ^ x7)) =
              make_uint4(p[0], p[1], p[2], p[3]);
        }
      };
      // The accumulator is handed on in two 64-column halves: MMA3 overwrites D1 with D3 half by half and B starts on the
      // first half while this role still works on the second (D1 and D3 share the buffer, so the chain
      // MMA1 -> A -> MMA3 -> B per buffer was the critical path of the kernel)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t ra[32];
        tmem_ld32(d1 + c * 32, ra);
        tmem_ld_wait();
        process(ra, c);
        if (c & 1) {
          fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar((c == 1 ? C_D1_EMPTY : C_D1E_HI) + s));
            mbar_arrive(bar((c == 1 ? C_ST_FULL : C_STF_HI) + s));
          }
        }
      }
      TR(4, i, 2, e == 0);
    }
  } else if (warp < 16) {
    // =========================================================== B: D3^T -> g_t row (thread = edge row j of the tile)
    const int j = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t rowoff = (uint32_t)((j >> 3) * 1024 + (j & 7) * 128), x7 = (uint32_t)(j & 7) << 4;
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1;
      // g_t overwrites t IN PLACE (same thread, same address) and the buffer then is the K-major A operand of MMA4
      uint8_t* sT = smem + BO_ST + s * (2 * 128 * 128) + rowoff;
      const int q4 = i & (D4_STAGES - 1);
      float* red = reinterpret_cast<float*>(smem + BO_RED + q4 * (TILE * 4));
      TR(5, i, 0, j == 0);
      mbar_wait_guard(bar(C_ST_FULL + s), ph);     // t of this tile written by A
      mbar_wait_guard(bar(C_D3_FULL + s), ph);
      if (kExact) mbar_wait_guard(bar(C_D4_EMPTY + q4), ((i / D4_STAGES) & 1) ^ 1);  // e4(i-4) has read sRed[q4]
      TR(5, i, 1, j == 0);
      fence_after_sync();
      const float cut = __uint_as_float(reinterpret_cast<const uint2*>(smem + BO_META + ms * TILE * 8)[j].y);
      const uint32_t cut2 = pack_half2(cut, cut), ncut2 = pack_half2(-cut, -cut);
      float usum = 0.f;
      const uint32_t d3 = tmem + s * 128 + lane_sel;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        if (c == 2) {
          mbar_wait_guard(bar(C_STF_HI + s), ph);
          mbar_wait_guard(bar(C_D3F_HI + s), ph);
          fence_after_sync();
        }
        gt_chunk<kExact>(d3, sT, x7, cut2, ncut2, usum, c);
      }
      if (kExact) red[j] = usum;
      fence_before_sync();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(C_D3_EMPTY + s));
        mbar_arrive(bar(C_GT_FULL + s));
        mbar_arrive(bar(C_META_EMPTY + ms));
      }
      TR(5, i, 2, j == 0);
    }
  } else if (warp < 20) {
    // =========================================================== E4: D4 -> g_d (thread = edge row)
    const int row = (warp & 3) * 32 + lane;
    const float g2 = gamma * 1.4426950408889634f;
    const float pi_over_rc = FMD_PI_F / rc;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const RbfRecurrence rrec = make_rbf_recurrence(sCen, R, g2);
    int tile = blockIdx.x;
    float d_n = 0.f, gd_n = 0.f;
    auto prefetch = [&](int t) {
      const int e = t * TILE + row;
      d_n = 0.f; gd_n = 0.f;
      if (e < E) {
        d_n = __ldg(&dist[e]);
        if (accumulate) gd_n = g_d[e];
      }
    };
    if (n_my > 0) prefetch(tile);
    for (int i = 0; i < n_my; ++i, tile += gridDim.x) {
      const int s = i & (D4_STAGES - 1);
      const uint32_t ph = (i / D4_STAGES) & 1;
      const int e = tile * TILE + row;
      const float d = d_n, prev_gd = gd_n;
      if (i + 1 < n_my) prefetch(tile + gridDim.x);
      const float cut = e < E ? cosine_cutoff_fast(d, pi_over_rc, rc) : 0.f;
      TR(1, i, 0, row == 0);
      mbar_wait_guard(bar(C_D4_FULL + s), ph);
      TR(1, i, 1, row == 0);
      fence_after_sync();
      const float dcut = d < rc ? -0.5f * pi_over_rc * __sinf(d * pi_over_rc) : 0.f;
      const float two_g_cut = 2.0f * gamma * cut;
      float acc = 0.f;
      // d/dd [exp(gamma (d - mu_k)^2) C(d)] = ex_k w_k,  w_k = 2 gamma (d - mu_k) C + C',  k < R only (column R of D4 belongs
      // to the bias).  Equally spaced centres: ex_k by the multiplicative recurrence of the producers (restarted every 8
      // columns), w_k by a running difference: 5 FP32 operations per column, 2 MUFU per 8 columns.
      const float wstep = -two_g_cut * rrec.delta;
#pragma unroll 1
      for (int c = 0; c < 64; c += 16) {
        if (c >= R) break;
        uint32_t r[16];
        tmem_ld16(tmem + TM_D4 + s * 64 + lane_sel + c, r);
        tmem_ld_wait();
        if (rrec.uniform) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k0 = c + 8 * h;
            if (k0 < R) {
              const float x0 = d - sCen[k0];
              float ex = ex2_approx(g2 * x0 * x0);
              float q = ex2_approx(fmaf(x0, rrec.a, rrec.b));
              float wk = fmaf(two_g_cut, x0, dcut);
              if (k0 + 8 <= R) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  acc = fmaf(__uint_as_float(r[8 * h + u]) * ex, wk, acc);
                  ex *= q;
                  q *= rrec.cstep;
                  wk += wstep;
                }
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  if (k0 + u < R) acc = fmaf(__uint_as_float(r[8 * h + u]) * ex, wk, acc);
                  ex *= q;
                  q *= rrec.cstep;
                  wk += wstep;
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            if (c + u < R) {
              const float diff = d - sCen[c + u];
              const float ex = ex2_approx(g2 * diff * diff);
              acc = fmaf(__uint_as_float(r[u]), ex * fmaf(two_g_cut, diff, dcut), acc);
            }
          }
        }
      }
      if (kExact) acc += dcut * reinterpret_cast<const float*>(smem + BO_RED + s * (TILE * 4))[row];
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C_D4_EMPTY + s));
      if (e < E) g_d[e] = prev_gd + acc;
      TR(1, i, 2, row == 0);
    }
  } else {
    // =========================================================== M: MMA issuers
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);   // D1^T[e,j]: A = rbf tile (K-major), B = Wf0 (K-major)
      constexpr uint32_t IDESC3H = idesc_f16(128, 64, 0, 1);   // D3^T[e,j]: A = gW0 [e][f] (K-major, TMEM), B = Wf1 [f][j] (MN-major)
      constexpr uint32_t IDESC4 = idesc_f16(128, 64, 0, 1);    // D4[e,k]:   A = g_t [e][j] (K-major),   B = Wf0 [j][k] (MN-major)
      const uint64_t dW0k = smem_desc_sw128(sbase + BO_WF0, 16, 1024);          // Wf0 as K-major operand (rows j)
      const uint64_t dW1mn = smem_desc_sw128(sbase + BO_WF1, 128 * 128, 1024);  // Wf1 [f][j] read MN-major (N = j)
      const uint64_t dB4 = smem_desc_sw128(sbase + BO_WF0, 16, 1024);           // Wf0 [j][k] read MN-major (N = k)
      auto issue1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        TR(6, i, 0, true);
        mbar_wait_guard(bar(C_RBF_FULL + s), ph);
        mbar_wait_guard(bar(C_D3_EMPTY + s), ph ^ 1);  // D13[s] drained by B(i-2)
        TR(6, i, 1, true);
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
        TR(7, i, 0, true);
        mbar_wait_guard(bar(C_GW_FULL + s), ph);
        mbar_wait_guard(bar(C_D1_EMPTY + s), ph);  // A(i) has drained D1 from D13[s]
        TR(7, i, 1, true);
        fence_after_sync();
        // A = gW0 [e][f] in tensor memory (8 columns per K = 16 step); two N = 64 halves (filter columns 0-63, 64-127)
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16_ts(tmem + s * 128, tmem + TM_GW + s * 64 + k * 8, dW1mn + (uint64_t)(k * (2048 / 16)), IDESC3H, k > 0);
        mma_commit(bar(C_D3_FULL + s));
        mbar_wait_guard(bar(C_D1E_HI + s), ph);    // A(i) has drained the second half of D1
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16_ts(tmem + s * 128 + 64, tmem + TM_GW + s * 64 + k * 8,
                     dW1mn + (uint64_t)(128 * 128 / 16 + k * (2048 / 16)), IDESC3H, k > 0);
        mma_commit(bar(C_OP_EMPTY + s));
        mma_commit(bar(C_D3F_HI + s));
      };
      auto issue4 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int q4 = i & (D4_STAGES - 1);
        TR(8, i, 0, true);
        mbar_wait_guard(bar(C_GT_FULL + s), ph);
        mbar_wait_guard(bar(C_D4_EMPTY + q4), ((i / D4_STAGES) & 1) ^ 1);
        TR(8, i, 1, true);
        fence_after_sync();
        const uint64_t dA4 = smem_desc_sw128(sbase + BO_ST + s * (2 * 128 * 128), 16, 1024);   // g_t [e][j], K-major
#pragma unroll
        for (int k = 0; k < NF / 16; ++k)
          mma_f16(tmem + TM_D4 + q4 * 64, dA4 + (uint64_t)((k >> 2) * (128 * 128 / 16) + (k & 3) * 2),
                  dB4 + (uint64_t)(k * (2048 / 16)), IDESC4, k > 0);
        mma_commit(bar(C_ST_EMPTY + s));
        mma_commit(bar(C_D4_FULL + q4));
      };
      if (warp == 20) {
        for (int i = 0; i < n_my; ++i) issue1(i);
      } else if (warp == 21) {
        for (int i = 0; i < n_my; ++i) issue3(i);
      } else {
        for (int i = 0; i < n_my; ++i) issue4(i);
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
#undef TR
}

}  // namespace

// tools only (scripts/trace_roles.py): device buffer of 9 * 64 * 3 uint64 that receives the role timeline of CTA 0; NULL = off
static unsigned long long* g_bwd_trace = nullptr;
extern "C" int fmd_debug_set_trace_bwd(void* device_buffer) {
  g_bwd_trace = (unsigned long long*)device_buffer;
  return FMD_OK;
}

extern "C" int fmd_filter_cfconv_bwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                                      int capacity, const int32_t* n_edges_dev, const void* wf0_h, const void* bf0_h,
                                      const void* wf1_h, const float* centers, int num_rbf, float gamma, float rc,
                                      const void* a_h, const void* g_m_h, int n_feat, float* g_d, int accumulate,
                                      int exact_cutoff_grad, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && wf0_h && wf1_h && centers && a_h && g_m_h && g_d,
              "fmd_filter_cfconv_bwd: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf < RP,
              "fmd_filter_cfconv_bwd: needs F == 128 and num_rbf <= 63 (one padded column carries the bias)");
  if (capacity <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ex = exact_cutoff_grad ? 1 : 0, tr = g_bwd_trace != nullptr ? 1 : 0;
  auto kern = filter_cfconv_bwd_kernel<false, false>;
  if (ex && !tr) kern = filter_cfconv_bwd_kernel<true, false>;
  if (!ex && tr) kern = filter_cfconv_bwd_kernel<false, true>;
  if (ex && tr) kern = filter_cfconv_bwd_kernel<true, true>;
  static bool attr_done[4] = {false, false, false, false};
  if (!attr_done[ex * 2 + tr]) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BSMEM_ALLOC));
    attr_done[ex * 2 + tr] = true;
  }
  const int max_tiles = fmd_div_up(capacity, TILE);
  const int grid = max_tiles < fmd_num_sms() ? max_tiles : fmd_num_sms();
  kern<<<grid, BWD_THREADS, BSMEM_ALLOC, st>>>(dist, edge_owner, edge_nbr, capacity, n_edges_dev, (const __half*)wf0_h,
                                               (const __half*)bf0_h, (const __half*)wf1_h, centers, num_rbf, gamma, rc,
                                               (const __half*)a_h, (const __half*)g_m_h, g_d, accumulate, g_bwd_trace);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
