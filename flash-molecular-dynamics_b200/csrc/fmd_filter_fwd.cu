// Fused filter network (x) CFConv, forward direction (also used for grad_x on the symmetric edge list):
//
//   out[i,f] = sum_{e in seg(i)}  W_e[f] * x[nbr_e,f] ,   W_e = Wf1 . ( tanh(Wf0 . rbf(d_e) + b) * C(d_e) )
//
// replaces, per interaction block, the reference's filter-network GEMMs (models/gptq.py:92-130,
// kernels/cfconv_kernels.py:644-952) and the CSR CFConv (kernels/csr_kernels.py:625-855); no [E,F] tensor exists.
//
// Per 128-edge tile everything is computed TRANSPOSED, so a TMEM lane (= one thread) owns a feature and the 128
// accumulator columns are the tile's edges: the segment sum over edges is an in-thread serial sum (deterministic,
// atomic-free) and `tanh`'s fp16 result row is the MN-major B operand of the second GEMM.
//
// One persistent CTA per SM, 17 warps in five roles connected by mbarrier rings:
//   P  warps 0-3    thread = edge: metadata prefetch, cut-off, radial-basis row (+ a constant-1 column that carries the
//                   bias through the first GEMM) -> sRbf[2], sMeta[4]
//   M  warp  16     one thread issues tcgen05.mma:  D1[s] = [Wf0 | b] . [rbf | 1]^T ,  D2[s] = Wf1 . t^T
//   T  warps 4-7    thread = feature: D1[s] -> tanh.approx.f16x2 -> * C(d_e) -> fp16 row of t^T -> sTT[2]
//   E0 warps 8-11   thread = feature, even tiles: D2[0] (rounded to fp16 like the reference's filter tensor) times the
//                   gathered fp16 row element, fp32 accumulate (one FHFMA per element), segment sums
//   E1 warps 12-15  same for odd tiles
// The gathered rows x[nbr_e,:] (fp16, 256 B) never go through registers: every epilogue warp prefetches ITS OWN 64-byte
// feature slice of the rows of its group's NEXT tile with cp.async (16 B per lane, 8 rows per instruction) into a
// private slice of a one-tile ring, chunk by chunk into the slots it has just consumed (cp.async groups: no barrier, no
// other warp involved); a full tile period hides the L2 latency.  (The per-thread LDG version of this kernel was bound by
// that latency; dedicated loader warps with per-chunk full / empty mbarriers were slower: an mbarrier wait costs ~100
// cycles of the waiting warp's in-order stream even when the phase has completed.)
//
// TMEM: D1[2] + D2[2] = 512 columns.  smem: weights 48 KB + sRbf 2x16 KB + sTT 2x32 KB + x ring 2x34 KB + meta.
#include "fmd_filter_shared.cuh"

using namespace fmd;
using namespace fmd::tc;
using namespace fmd::filt;

namespace {

constexpr int NTHREADS = 17 * 32;   // at most 5 warps per SM sub-partition: 16384 / (5 * 32) -> 96 registers per thread
constexpr int META_STAGES = 4;
constexpr int XPITCH = 272;                                   // bytes per staged row: 256 + 16 (bank spread of the 8-row copies)

constexpr uint32_t O_WF0 = 0;                                // 16 KB
constexpr uint32_t O_WF1 = O_WF0 + 128 * 128;                // 32 KB
constexpr uint32_t O_RBF = O_WF1 + 2 * 128 * 128;            // 2 x 16 KB
constexpr uint32_t O_TT = O_RBF + 2 * 128 * 128;             // 2 x 32 KB
constexpr uint32_t O_XS = O_TT + 2 * 2 * 128 * 128;          // 2 x 34 KB: staged x rows, one tile per epilogue group
constexpr uint32_t O_XOFF = O_XS + 2 * TILE * XPITCH;        // 4 x 512 B: BYTE offset nbr * 256 of the gathered row
constexpr uint32_t O_CUT = O_XOFF + META_STAGES * TILE * 4;  // 4 x 256 B: C(d_e) as fp16
constexpr uint32_t O_OWN = O_CUT + META_STAGES * TILE * 2;   // 4 x 512 B: segment owner per edge
constexpr uint32_t O_HEAD = O_OWN + META_STAGES * TILE * 4;  // 4 x 32 B: {prev_owner, -, -, -, boundary mask[4]}
constexpr uint32_t O_CEN = O_HEAD + META_STAGES * 32;
constexpr uint32_t O_BAR = O_CEN + RP * 4;                   // 24 mbarriers + tmem slot
constexpr uint32_t SMEM3 = O_BAR + 24 * 8 + 16;
constexpr uint32_t SMEM3_ALLOC = SMEM3 + 1024;
static_assert(SMEM3_ALLOC <= 232448, "forward kernel exceeds the 227 KB shared-memory limit");

enum { B_RBF_FULL = 0, B_RBF_EMPTY = 2, B_D1_FULL = 4, B_D1_EMPTY = 6, B_TT_FULL = 8, B_TT_EMPTY = 10,
       B_D2_FULL = 12, B_D2_EMPTY = 14, B_META_EMPTY = 16, B_META_FULL = 20, B_COUNT = 24 };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc += w.lo * x  /  acc += w.hi * x : fp16 x fp16 product (exact in fp32), fp32 accumulate, ONE instruction (FHFMA)
__device__ __forceinline__ void fhfma_lo(float& acc, uint32_t w2, unsigned short xh) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.f16 %0, lo, %2, %0;\n\t}" : "+f"(acc) : "r"(w2), "h"(xh));
}
__device__ __forceinline__ void fhfma_hi(float& acc, uint32_t w2, unsigned short xh) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.f16 %0, hi, %2, %0;\n\t}" : "+f"(acc) : "r"(w2), "h"(xh));
}

// column u of a chunk whose only segment boundary is at column n: (u < n ? pre : post) += w * x, predicated
__device__ __forceinline__ void fhfma_sel_lo(float& pre, float& post, uint32_t w2, unsigned short xh, int n, int u) {
  asm("{\n\t.reg .pred p;\n\t.reg .b16 lo, hi;\n\tsetp.gt.s32 p, %4, %5;\n\tmov.b32 {lo, hi}, %2;\n\t"
      "@p fma.rn.f32.f16 %0, lo, %3, %0;\n\t@!p fma.rn.f32.f16 %1, lo, %3, %1;\n\t}"
      : "+f"(pre), "+f"(post) : "r"(w2), "h"(xh), "r"(n), "r"(u));
}
__device__ __forceinline__ void fhfma_sel_hi(float& pre, float& post, uint32_t w2, unsigned short xh, int n, int u) {
  asm("{\n\t.reg .pred p;\n\t.reg .b16 lo, hi;\n\tsetp.gt.s32 p, %4, %5;\n\tmov.b32 {lo, hi}, %2;\n\t"
      "@p fma.rn.f32.f16 %0, hi, %3, %0;\n\t@!p fma.rn.f32.f16 %1, hi, %3, %1;\n\t}"
      : "+f"(pre), "+f"(post) : "r"(w2), "h"(xh), "r"(n), "r"(u));
}

// radial-basis row of one edge, chunks 0 .. nchunks-1 of 8 columns (see write_rbf_row_rec); column `R` is then set to 1
// (bias column).  Columns beyond carry finite values that multiply zero-padded weight columns.
__device__ __forceinline__ void write_rbf_row_bias(uint8_t* sRbf, const float* sCen, int row, float d, float cut, float g2,
                                                   const RbfRecurrence& rr, int nchunks, int R) {
#pragma unroll 2
  for (int c = 0; c < nchunks; ++c) {
    float vals[8];
    if (rr.uniform) {
      const float x = d - sCen[c * 8];
      float v = ex2_approx(g2 * x * x) * cut;
      float q = ex2_approx(fmaf(x, rr.a, rr.b));
      vals[0] = v;
#pragma unroll
      for (int u = 1; u < 8; ++u) {
        v *= q;
        q *= rr.cstep;
        vals[u] = v;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float x = d - sCen[c * 8 + u];
        vals[u] = ex2_approx(g2 * x * x) * cut;
      }
    }
    *reinterpret_cast<uint4*>(sRbf + sw128_off(row, c)) =
        make_uint4(pack_half2(vals[0], vals[1]), pack_half2(vals[2], vals[3]), pack_half2(vals[4], vals[5]), pack_half2(vals[6], vals[7]));
  }
  *reinterpret_cast<unsigned short*>(sRbf + sw128_off(row, R >> 3) + (R & 7) * 2) = 0x3C00u;   // 1.0 (fp16)
}

// Role timeline (tools only, compiled out of the production instantiation): CTA 0 records clock64() stamps
// {wait start, work start, end} per role and tile into trace[role][tile < 64][3]  (scripts/trace_roles_fwd.py).
template <bool kTrace>
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int role, int tile, int k, bool leader) {
  if (kTrace) {
    if (blockIdx.x == 0 && leader && tile < 64) trace[(role * 64 + tile) * 3 + k] = (unsigned long long)clock64();
  }
}

template <bool kTrace>
__global__ void __maxnreg__(96)
filter_cfconv_fwd_kernel(const float* __restrict__ dist, const int32_t* __restrict__ edge_owner,
                          const int32_t* __restrict__ edge_nbr, int capacity,
                          const int32_t* __restrict__ n_edges_dev, const __half* __restrict__ wf0,
                          const __half* __restrict__ bf0, const __half* __restrict__ wf1,
                          const float* __restrict__ centers, int R, float gamma, float rc,
                          const __half* __restrict__ x, float* __restrict__ out, float* __restrict__ part,
                          unsigned long long* __restrict__ trace) {
#define TR(role, tile, k, leader) trace_stamp<kTrace>(trace, role, tile, k, leader)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sCen = reinterpret_cast<float*>(smem + O_CEN);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + O_BAR + B_COUNT * 8);
  auto bar = [&](int i) { return sbase + O_BAR + 8u * (uint32_t)i; };

  const int E = min(capacity, n_edges_dev ? *n_edges_dev : capacity);
  const int n_tiles = (E + TILE - 1) / TILE;
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- one-time setup
  load_weight_kmajor(smem + O_WF0, wf0, NF, RP / 8);
  load_weight_kmajor(smem + O_WF1, wf1, NF, NF / 8);
  // both radial-basis stages start as zeros (the chunks beyond the bias column are never written again)
  for (int idx = tid; idx < 2 * 128 * 128 / 16; idx += NTHREADS)
    reinterpret_cast<uint4*>(smem + O_RBF)[idx] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < RP) sCen[tid] = tid < R ? centers[tid] : 0.f;
  __syncthreads();
  // bias column: Wf0p[j, R] = b_j, multiplied by the constant-1 column R of every radial-basis row
  if (tid < NF)
    *reinterpret_cast<__half*>(smem + O_WF0 + sw128_off(tid, R >> 3) + (R & 7) * 2) = bf0 ? bf0[tid] : __float2half(0.f);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      // software arrivals are ONE per warp (lane 0 after __syncwarp): every arrival wakes all the warps parked in
      // try_wait, and with one arrival per thread the wake-up / re-check loop was 22 % of all issued instructions
      mbar_init(bar(B_RBF_FULL + i), 4);
      mbar_init(bar(B_RBF_EMPTY + i), 1);
      mbar_init(bar(B_D1_FULL + i), 1);
      mbar_init(bar(B_D1_EMPTY + i), 4);
      mbar_init(bar(B_TT_FULL + i), 4);
      mbar_init(bar(B_TT_EMPTY + i), 1);
      mbar_init(bar(B_D2_FULL + i), 1);
      mbar_init(bar(B_D2_EMPTY + i), 4);
    }
    for (int i = 0; i < META_STAGES; ++i) {
      mbar_init(bar(B_META_EMPTY + i), 8);     // consumed by the tanh warps (cut) and one epilogue group
      mbar_init(bar(B_META_FULL + i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(sbase + O_BAR + B_COUNT * 8, 512);
  }
  fence_async_smem();  // weights / zeros were written through the generic proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =========================================================== P: producers (thread = edge row)
    const float g2 = gamma * 1.4426950408889634f;
    const float pi_over_rc = FMD_PI_F / rc;
    const RbfRecurrence rrec = make_rbf_recurrence(sCen, R, g2);
    const int nchunks = (R + 8) >> 3;          // chunks holding columns 0 .. R (radial basis + bias column)
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
      TR(0, i, 0, tid == 0);
      mbar_wait_guard(bar(B_META_EMPTY + ms), mph ^ 1);
      reinterpret_cast<uint32_t*>(smem + O_XOFF + ms * TILE * 4)[tid] = (uint32_t)nb * (uint32_t)(NF * 2);
      reinterpret_cast<__half*>(smem + O_CUT + ms * TILE * 2)[tid] = __float2half_rn(cut);
      reinterpret_cast<int*>(smem + O_OWN + ms * TILE * 4)[tid] = own;
      if (tid == 0) *reinterpret_cast<int*>(smem + O_HEAD + ms * 32) = prev;
      if (lane == 0) reinterpret_cast<uint32_t*>(smem + O_HEAD + ms * 32 + 16)[warp] = bmask;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_META_FULL + ms));
      mbar_wait_guard(bar(B_RBF_EMPTY + s), ph ^ 1);
      TR(0, i, 1, tid == 0);
      write_rbf_row_bias(smem + O_RBF + s * (128 * 128), sCen, tid, d, cut, g2, rrec, nchunks, R);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_RBF_FULL + s));
      TR(0, i, 2, tid == 0);
    }
  } else if (warp == 16) {
    // =========================================================== M: MMA issuer (one thread)
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t IDESC1 = idesc_f16(128, 128, 0, 0);
      constexpr uint32_t IDESC2 = idesc_f16(128, 128, 0, 1);
      const uint64_t dA1 = smem_desc_sw128(sbase + O_WF0, 16, 1024);
      const uint64_t dA2 = smem_desc_sw128(sbase + O_WF1, 16, 1024);
      auto issue1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        TR(7, i, 0, true);
        mbar_wait_guard(bar(B_RBF_FULL + s), ph);
        TR(7, i, 1, true);
        mbar_wait_guard(bar(B_D1_EMPTY + s), ph ^ 1);
        TR(7, i, 2, true);
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
        TR(8, i, 0, true);
        mbar_wait_guard(bar(B_TT_FULL + s), ph);
        TR(8, i, 1, true);
        mbar_wait_guard(bar(B_D2_EMPTY + s), ph ^ 1);
        TR(8, i, 2, true);
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
  } else if (warp < 8) {
    // =========================================================== T: tanh (thread = feature j)
    const int j = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t jrow = (uint32_t)((j >> 3) * 1024 + (j & 7) * 128), jx = (uint32_t)(j & 7);
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1, ms = i & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1;
      TR(1, i, 0, j == 0);
      mbar_wait2_guard(bar(B_D1_FULL + s), ph, bar(B_TT_EMPTY + s), ph ^ 1);
      TR(1, i, 1, j == 0);
      fence_after_sync();
      uint8_t* sTT = smem + O_TT + s * (2 * 128 * 128) + jrow;
      const uint4* sCutH = reinterpret_cast<const uint4*>(smem + O_CUT + ms * TILE * 2);   // 8 fp16 cut-offs per load
      const uint32_t d1 = tmem + s * 128 + lane_sel;
      // t * C(d_e): D2 is linear in t, so the cut-off rides through the second GEMM for free.  tanh.approx.f32 per value,
      // one F2FP pack and one HMUL2 per two values (the bias arrived through the GEMM; the packed-fp16 tanh is no cheaper:
      // it is issued as two MUFU operations plus a pack before and a PRMT after).  The MUFU pipe issues one warp
      // instruction every 8 cycles per sub-partition, so the TMEM load of the next 32 columns is in flight while the
      // current 32 go through it.
      auto process = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 cq = sCutH[c * 4 + q];
          const uint32_t cc[4] = {cq.x, cq.y, cq.z, cq.w};
          uint32_t p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            p[u] = hmul2_u32(pack_half2(tanh_approx(__uint_as_float(r[q * 8 + 2 * u])),
                                        tanh_approx(__uint_as_float(r[q * 8 + 2 * u + 1]))), cc[u]);
          const int chunk = c * 4 + q;
          *reinterpret_cast<uint4*>(sTT + (chunk >> 3) * (128 * 128) + ((((uint32_t)chunk & 7u) ^ jx) << 4)) =
              make_uint4(p[0], p[1], p[2], p[3]);
        }
      };
      uint32_t ra[32], rb[32];
      tmem_ld32(d1, ra);
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        tmem_ld_wait();
        tmem_ld32(d1 + (c + 1) * 32, rb);
        process(ra, c);
        tmem_ld_wait();
        if (c + 2 < 4) tmem_ld32(d1 + (c + 2) * 32, ra);
        process(rb, c + 1);
      }
      fence_before_sync();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_D1_EMPTY + s));
        mbar_arrive(bar(B_META_EMPTY + ms));
        mbar_arrive(bar(B_TT_FULL + s));
      }
      TR(1, i, 2, j == 0);
    }
  } else {
    // =========================================================== E0 / E1: epilogue (thread = feature f)
    const int g = warp < 12 ? 0 : 1;
    const int wq = warp & 3;
    const int f = wq * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    uint8_t* xs = smem + O_XS + g * (TILE * XPITCH);                 // this group's staged tile
    const uint8_t* xsf = xs + f * 2;                                  // element (e, f) at xsf + e * XPITCH
    const uint32_t xs_dst = smem_u32(xs) + (uint32_t)((lane >> 2) * XPITCH + wq * 64 + (lane & 3) * 16);
    const uint8_t* xsrc = reinterpret_cast<const uint8_t*>(x) + wq * 64 + (lane & 3) * 16;
    // prefetch of this warp's 64-byte slice of rows [c16*16, c16*16+16) of the tile whose metadata sits in stage `ms`
    auto stage16 = [&](int ms, int c16) {
      const uint32_t* sOff = reinterpret_cast<const uint32_t*>(smem + O_XOFF + ms * TILE * 4) + c16 * 16 + (lane >> 2);
      const uint32_t o0 = sOff[0], o1 = sOff[8];     // both offsets before the (ordered, asm volatile) copies
      cp_async16(xs_dst + (uint32_t)(c16 * 16 * XPITCH), xsrc + o0);
      cp_async16(xs_dst + (uint32_t)((c16 * 16 + 8) * XPITCH), xsrc + o1);
    };
    if (g < n_my) {
      mbar_wait_guard(bar(B_META_FULL + g), 0);
#pragma unroll 1
      for (int c16 = 0; c16 < 8; ++c16) {
        stage16(g, c16);
        cp_async_commit();
      }
    }
    for (int i = g; i < n_my; i += 2) {
      const int ms = i & (META_STAGES - 1), ms2 = (i + 2) & (META_STAGES - 1);
      const uint32_t ph = (i >> 1) & 1, mph = (i / META_STAGES) & 1;
      const int tile = blockIdx.x + i * gridDim.x;
      const bool has_next = i + 2 < n_my;
      const int* sOwn = reinterpret_cast<const int*>(smem + O_OWN + ms * TILE * 4);
      const uint32_t* sMask = reinterpret_cast<const uint32_t*>(smem + O_HEAD + ms * 32 + 16);
      TR(2 + g, i, 0, f == 0);
      if (has_next) mbar_wait2_guard(bar(B_META_FULL + ms), mph, bar(B_META_FULL + ms2), ((i + 2) / META_STAGES) & 1);
      else mbar_wait_guard(bar(B_META_FULL + ms), mph);
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
      mbar_wait_guard(bar(B_D2_FULL + g), ph);
      TR(2 + g, i, 1, f == 0);
      fence_after_sync();
      const uint32_t d2 = tmem + 256 + g * 128 + lane_sel;
      // software pipeline over the eight 16-column chunks: the TMEM load and the staged-row reads of chunk c+1 are in
      // flight while chunk c is multiplied
      uint32_t r[16];
      unsigned short xa[16], xb[16];
      auto ldx = [&](unsigned short (&xv)[16], int c16) {
        const uint8_t* xp = xsf + c16 * 16 * XPITCH;
#pragma unroll
        for (int u = 0; u < 16; ++u) xv[u] = *reinterpret_cast<const unsigned short*>(xp + u * XPITCH);
      };
      auto consume = [&](const uint32_t (&w2)[8], const unsigned short (&xv)[16], int c16) {
        const uint32_t bits = (sMask[c16 >> 1] >> ((c16 & 1) * 16)) & 0xffffu;
        if (bits == 0u) {
          float acc1 = 0.f;        // two independent chains (even / odd columns)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            fhfma_lo(acc, w2[u], xv[2 * u]);
            fhfma_hi(acc1, w2[u], xv[2 * u + 1]);
          }
          acc += acc1;
        } else if ((bits & (bits - 1u)) == 0u) {
          // exactly one segment boundary in the chunk (the common case: degrees > 16), at column n: columns < n finish
          // the running sum, columns >= n start the next one.  Predicated, branch-free (the per-column test-and-branch
          // form of this path cost ~500 cycles per boundary chunk: 32 warp-uniform branches)
          const int n = __ffs(bits) - 1;
          float post = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            fhfma_sel_lo(acc, post, w2[u], xv[2 * u], n, 2 * u);
            fhfma_sel_hi(acc, post, w2[u], xv[2 * u + 1], n, 2 * u + 1);
          }
          flush(sOwn[c16 * 16 + n]);
          acc = post;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if ((bits >> (2 * u)) & 1u) flush(sOwn[c16 * 16 + 2 * u]);
            fhfma_lo(acc, w2[u], xv[2 * u]);
            if ((bits >> (2 * u + 1)) & 1u) flush(sOwn[c16 * 16 + 2 * u + 1]);
            fhfma_hi(acc, w2[u], xv[2 * u + 1]);
          }
        }
        // the slots just consumed take the same rows of this group's next tile
        if (has_next) stage16(ms2, c16);
        cp_async_commit();
      };
      tmem_ld16(d2, r);
      cp_async_wait<7>();          // the copies of chunk 0 (committed 8 groups ago) have landed ...
      __syncwarp();                // ... for every lane of this warp (the slice is private to the warp)
      ldx(xa, 0);
#pragma unroll 1
      for (int c16 = 0; c16 < 8; c16 += 2) {
        uint32_t w2[8];
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) w2[u] = pack_half2(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1]));
        tmem_ld16(d2 + (c16 + 1) * 16, r);
        cp_async_wait<6>();        // one group fewer has been committed since: chunk c16 + 1 has landed
        __syncwarp();
        ldx(xb, c16 + 1);
        consume(w2, xa, c16);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) w2[u] = pack_half2(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1]));
        if (c16 + 2 < 8) {
          tmem_ld16(d2 + (c16 + 2) * 16, r);
          cp_async_wait<6>();
          __syncwarp();
          ldx(xa, c16 + 2);
        }
        consume(w2, xb, c16 + 1);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_D2_EMPTY + g));
        mbar_arrive(bar(B_META_EMPTY + ms));
      }
      flush(0);
      TR(2 + g, i, 2, f == 0);
    }
    cp_async_wait<0>();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
#undef TR
}

// out[i] = 0 for empty segments; out[i] += head partials of the later tiles a straddling segment touches,
// in tile order (deterministic).  Eight lanes per node, four float4 per lane (NF = 128): the kernel is a chain of dependent
// loads (segment bounds -> rows), so a warp works on four nodes at once.
__global__ void __launch_bounds__(256)
cfconv_fixup_kernel(const int32_t* __restrict__ seg_ptr, int n_nodes, int capacity, const float* __restrict__ part,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 7;
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (node >= n_nodes) return;
  const int E = min(capacity, __ldg(&seg_ptr[n_nodes]));
  const int s0 = min(__ldg(&seg_ptr[node]), E), s1 = min(__ldg(&seg_ptr[node + 1]), E);
  float4* o = reinterpret_cast<float4*>(out + (size_t)node * NF) + lane;
  if (s1 <= s0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[8 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int t0 = s0 / TILE, t1 = (s1 - 1) / TILE;
  if (t1 == t0) return;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = o[8 * k];
  for (int t = t0 + 1; t <= t1; ++t) {
    const float4* pr = reinterpret_cast<const float4*>(part + (size_t)t * NF) + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 p = __ldg(pr + 8 * k);
      v[k].x += p.x; v[k].y += p.y; v[k].z += p.z; v[k].w += p.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) o[8 * k] = v[k];
}

}  // namespace

// tools only (scripts/trace_roles_fwd.py): device buffer of 9 * 64 * 3 uint64 that receives the role timeline of CTA 0; NULL = off
static unsigned long long* g_fwd_trace = nullptr;
extern "C" int fmd_debug_set_trace_fwd(void* device_buffer) {
  g_fwd_trace = (unsigned long long*)device_buffer;
  return FMD_OK;
}
extern "C" int fmd_filter_cfconv_fwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                                      const int32_t* seg_ptr, int n_nodes, int capacity, const int32_t* n_edges_dev,
                                      const void* wf0_h, const void* bf0_h, const void* wf1_h, const float* centers,
                                      int num_rbf, float gamma, float rc, const void* x_h, int n_feat, float* out,
                                      float* part, void* stream) {
  FMD_REQUIRE(dist && edge_owner && edge_nbr && seg_ptr && wf0_h && wf1_h && centers && x_h && out && part,
              "fmd_filter_cfconv_fwd: null argument");
  FMD_REQUIRE(n_feat == NF && num_rbf > 0 && num_rbf < RP,
              "fmd_filter_cfconv_fwd: needs F == 128 and num_rbf <= 63 (one padded column carries the bias)");
  FMD_REQUIRE(n_nodes < (1 << 24), "fmd_filter_cfconv_fwd: gathered rows are addressed with 32-bit byte offsets (< 2^24 nodes)");
  if (n_nodes <= 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int tr = g_fwd_trace != nullptr ? 1 : 0;
  auto kern = tr ? filter_cfconv_fwd_kernel<true> : filter_cfconv_fwd_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[tr]) {
    FMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM3_ALLOC));
    attr_done[tr] = true;
  }
  if (capacity > 0) {
    const int max_tiles = fmd_div_up(capacity, TILE);
    const int grid = max_tiles < fmd_num_sms() ? max_tiles : fmd_num_sms();
    kern<<<grid, NTHREADS, SMEM3_ALLOC, st>>>(dist, edge_owner, edge_nbr, capacity, n_edges_dev, (const __half*)wf0_h,
                                              (const __half*)bf0_h, (const __half*)wf1_h, centers, num_rbf, gamma, rc,
                                              (const __half*)x_h, out, part, g_fwd_trace);
    FMD_CHECK_LAUNCH();
  }
  cfconv_fixup_kernel<<<fmd_div_up((long long)n_nodes * 8, 256), 256, 0, st>>>(seg_ptr, n_nodes, capacity, part, out);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
