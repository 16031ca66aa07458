// Pieces shared by the fused filter-network kernels (v1 single-stage, v2 warp-specialised pipeline).
#pragma once
#include "fmd_tc.cuh"

namespace fmd {
namespace filt {
using namespace fmd::tc;

constexpr int TILE = 128;  // edges per tile == threads per CTA == features
constexpr int NF = 128;    // filters / hidden width handled by this kernel
constexpr int RP = 64;     // num_rbf padded to a multiple of 16 (MMA K step)

struct __align__(16) EdgeMeta {
  int32_t owner;  // segment owner (edge_src)
  int32_t nbr;    // gathered node (edge_dst)
  float cut;      // C(d_e)
  float dist;
};

// copy a [rows][ncols16 * 8 halves] fp16 row-major global matrix into K-major swizzle-128B blocks of
// [rows][64 halves] (block kb holds columns 64*kb .. 64*kb+63, blocks are rows*128 bytes apart)
__device__ __forceinline__ void load_weight_kmajor(uint8_t* dst, const __half* __restrict__ src, int rows, int ncols16) {
  const int total = rows * ncols16;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx / ncols16, c = idx - r * ncols16;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + idx);
    const int kb = c >> 3, cc = c & 7;
    *reinterpret_cast<uint4*>(dst + kb * rows * 128 + sw128_off(r, cc)) = v;
  }
}

// thread e: radial basis of its edge -> row e of the K-major B operand (fp16, the reference's in-kernel
// cast kernels/cfconv_kernels.py:701); rbf_k = exp(gamma (d-mu_k)^2) * C(d)  (radial_basis/gaussian.py:83-102)
__device__ __forceinline__ void write_rbf_row(uint8_t* sRbf, const float* sCen, int row, float d, float cut, int R,
                                              float gamma, bool valid) {
#pragma unroll
  for (int c = 0; c < RP / 8; ++c) {
    uint32_t p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k0 = c * 8 + 2 * u;
      const float d0 = d - sCen[k0], d1 = d - sCen[k0 + 1];
      float v0 = __expf(gamma * d0 * d0) * cut, v1 = __expf(gamma * d1 * d1) * cut;
      if (!valid || k0 >= R) v0 = 0.f;
      if (!valid || k0 + 1 >= R) v1 = 0.f;
      p[u] = pack_half2(v0, v1);
    }
    *reinterpret_cast<uint4*>(sRbf + sw128_off(row, c)) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}


// Same, without per-element predicates: columns k >= R multiply zero-padded weight columns and rows of
// invalid edges are never consumed, so any FINITE value is fine there (sCen[k >= R] = 0, d = 0 for
// invalid rows).  g2 = gamma * log2(e): exp(gamma x^2) = ex2(g2 x^2), one MUFU per value.
__device__ __forceinline__ void write_rbf_row_fast(uint8_t* sRbf, const float* sCen, int row, float d, float cut,
                                                   float g2) {
#pragma unroll 2
  for (int c = 0; c < RP / 8; ++c) {
    const float4 ca = *reinterpret_cast<const float4*>(sCen + c * 8);
    const float4 cb = *reinterpret_cast<const float4*>(sCen + c * 8 + 4);
    const float mu[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    uint32_t p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d0 = d - mu[2 * u], d1 = d - mu[2 * u + 1];
      const float v0 = ex2_approx(g2 * d0 * d0) * cut, v1 = ex2_approx(g2 * d1 * d1) * cut;
      p[u] = pack_half2(v0, v1);
    }
    *reinterpret_cast<uint4*>(sRbf + sw128_off(row, c)) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

// Equally spaced centres (GaussianBasis: linspace(0, rc, R), radial_basis/gaussian.py:64-81): within a block of 8
// columns the Gaussian follows a two-term multiplicative recurrence,
//   v_{k+1} = v_k q_k,  q_{k+1} = q_k c,   q_k = exp(gamma (delta^2 - 2 delta (d - mu_k))),  c = exp(2 gamma delta^2),
// i.e. 2 MUFU + 14 FMUL per 8 values instead of 8 MUFU + 32 FP32 ops (the producers' share of the issue slots was
// 13 % of the forward kernel).  Every block restarts from a directly evaluated value, so a block whose first value
// underflows in fp32 only holds values below fp16 resolution (they grow by at most exp(7 x - 24.5) over 7 steps from
// < 2^-126).  Relative error <= 7 * 2^-24 + the ex2.approx error, far below the fp16 rounding of the operand.
struct RbfRecurrence {
  float a;      // -2 delta g2
  float b;      // g2 delta^2
  float cstep;  // ex2(2 g2 delta^2)
  float delta;  // centre spacing
  int uniform;  // centres equally spaced (else the direct evaluation is used)
};
__device__ __forceinline__ RbfRecurrence make_rbf_recurrence(const float* sCen, int R, float g2) {
  RbfRecurrence rr;
  const float delta = R > 1 ? sCen[1] - sCen[0] : 0.f;
  int ok = R > 8 && delta > 0.f;
  for (int k = 1; k + 1 < R; ++k) ok &= fabsf((sCen[k + 1] - sCen[k]) - delta) <= 1e-4f * delta;
  rr.a = -2.0f * delta * g2;
  rr.b = g2 * delta * delta;
  rr.cstep = ex2_approx(2.0f * g2 * delta * delta);
  rr.delta = delta;
  rr.uniform = ok;
  return rr;
}
__device__ __forceinline__ void write_rbf_row_rec(uint8_t* sRbf, const float* sCen, int row, float d, float cut, float g2,
                                                  const RbfRecurrence& rr) {
  if (!rr.uniform) {
    write_rbf_row_fast(sRbf, sCen, row, d, cut, g2);
    return;
  }
#pragma unroll 2
  for (int c = 0; c < RP / 8; ++c) {
    const float x = d - sCen[c * 8];
    float v = ex2_approx(g2 * x * x) * cut;
    float q = ex2_approx(fmaf(x, rr.a, rr.b));
    float vals[8];
    vals[0] = v;
#pragma unroll
    for (int u = 1; u < 8; ++u) {
      v *= q;
      q *= rr.cstep;
      vals[u] = v;
    }
    *reinterpret_cast<uint4*>(sRbf + sw128_off(row, c)) =
        make_uint4(pack_half2(vals[0], vals[1]), pack_half2(vals[2], vals[3]), pack_half2(vals[4], vals[5]), pack_half2(vals[6], vals[7]));
  }
}

// 0.5 (cos(pi d / rc) + 1) [d < rc] with the fast cosine (abs error ~1e-6, far below fp16 resolution)
__device__ __forceinline__ float cosine_cutoff_fast(float d, float pi_over_rc, float rc) {
  const float c = 0.5f * (__cosf(d * pi_over_rc) + 1.0f);
  return d < rc ? c : 0.0f;
}

}  // namespace filt
}  // namespace fmd
