// Classical prior terms, owner-computes form: eight lanes per bead walk the bead's incident terms
// (static topology, CSR built once at set-up) and accumulate the bead's force and its share of the
// term energies.  No atomics, deterministic; replaces the term-parallel atomicAdd kernel on the step
// path (which serialised ~36 k repulsion pairs per molecule onto one energy address).
// The kernel is bound by the latency of its dependent load chains (incidence record -> term indices ->
// positions -> parameters), not by arithmetic or bandwidth: measured on the benchmark system, a warp per
// bead at 48 registers took 83 us; 32 registers (64 resident warps per SM) 67 us; four beads per warp 53 us
// (17 us without the repulsion pairs); shared-memory staging of the positions made it slower.
//
// Math per term (reference file:line):
//   bonds      k (d - x0)^2 + V0          prior/harmonic.py:122-123, geometry/internal_coordinates.py:73-101
//   repulsion  (sigma / d)^6              prior/repulsion.py:119-122
//   angles     k (cos(theta) - x0)^2 + V0 prior/harmonic.py:122-123, internal_coordinates.py:140-170
//   dihedrals  v0 + sum_n k1_n sin(n phi) + k2_n cos(n phi)   prior/fourier_series.py:154-192, :174-223
// and the remaining prior classes of the reference (a model that contains one stays on the fused step):
//   polynomial bonds      V0 + sum_{n=1..4} k_n d^n                  prior/polynomial.py:13-186
//   polynomial angles     V0 + sum_{n=1..6} k_n cos^n                (QuarticAngles)
//   restricted bending    a c^4 + b c^3 + c c^2 + d c + k / sin^2 + V0   prior/restricted_bending.py:13-238
//   raw-angle harmonic    k (theta - x0)^2                            prior/harmonic.py:267-300
//   (shifted) harmonic impropers  k (x - x0)^2, x = phi or (phi < 0 ? phi + 2 pi : phi) - pi   prior/harmonic.py:230-265, 327-405
// Angle-like terms share ONE table with a per-term form code, improper-like terms a second one.
// A term's energy is split evenly between its beads (1/2, 1/3, 1/4) so that the per-molecule sum
// (fmd_segment_sum) reproduces scatter(y, mapping_batch) of the reference.
#include "fmd_common.cuh"

using namespace fmd;

namespace {

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 ld3(const float* __restrict__ p, int i) {
  return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

constexpr int WARPS = 8;
#ifndef FMD_PRIOR_LPB
#define FMD_PRIOR_LPB 8
#endif
constexpr int LPB = FMD_PRIOR_LPB;     // lanes per bead

// PACKED: two-body records are 8 bytes {other | kind << 28, parameter id} and the parameters (k, x0, V0 | sigma) come
// from a small deduplicated float4 table (a few hundred type pairs: L1-resident) - the 16-byte records were 147 MB of
// HBM traffic per step at 128 x 269 beads, by far the largest stream outside the two fused edge kernels.
#ifndef FMD_PRIOR_MIN_CTAS
#define FMD_PRIOR_MIN_CTAS 8
#endif
template <bool PACKED>
__global__ void __launch_bounds__(WARPS * 32, FMD_PRIOR_MIN_CTAS)
prior_csr_kernel(const float* __restrict__ pos, int n_nodes, const int32_t* __restrict__ pair_ptr,
                 const void* __restrict__ pair_ent_v, const float4* __restrict__ pair_tab,
                 const int32_t* __restrict__ mb_ptr,
                 const int32_t* __restrict__ mb_ent, const int32_t* __restrict__ ang_map, int n_ang,
                 const float4* __restrict__ ang_par,
                 const int32_t* __restrict__ dih_map, int n_dih, const float* __restrict__ dih_k1,
                 const float* __restrict__ dih_k2, const float* __restrict__ dih_v0, int n_degs,
                 const int32_t* __restrict__ imp_map, int n_imp, const float4* __restrict__ imp_par,
                 float* __restrict__ e_atom, float* __restrict__ forces, int accumulate_forces) {
  // LPB lanes per bead: the kernel is bound by the latency of its dependent load chains (record -> index -> position), so
  // a warp works on 32 / LPB beads at once instead of one
  const int lane = threadIdx.x & (LPB - 1);
  const int a_raw = (blockIdx.x * (WARPS * 32) + threadIdx.x) / LPB;
  const bool valid = a_raw < n_nodes;          // no early exit: the lane reductions below use the full-warp mask
  const int a = valid ? a_raw : n_nodes - 1;
  const V3 pa = ld3(pos, a);
  V3 f = {0.f, 0.f, 0.f};
  float e = 0.f;
  // ---- two-body terms: entry = {other | kind << 28, p0, p1, p2}
  if (pair_ptr && valid) {
    const int p1 = __ldg(&pair_ptr[a + 1]);
#pragma unroll 8   // independent record -> position load chains in flight (the loop is latency-bound on the record stream)
    for (int p = __ldg(&pair_ptr[a]) + lane; p < p1; p += LPB) {
      int head, tab_id = 0;
      float q0, q1, q2, q3 = 0.f;
      if (PACKED) {
        const int2 ent = __ldg(reinterpret_cast<const int2*>(pair_ent_v) + p);
        const float4 q = __ldg(&pair_tab[2 * ent.y]);      // two float4 per parameter id (the second: polynomial terms only)
        head = ent.x; tab_id = ent.y; q0 = q.x; q1 = q.y; q2 = q.z; q3 = q.w;
      } else {
        const int4 ent = __ldg(reinterpret_cast<const int4*>(pair_ent_v) + p);
        head = ent.x; q0 = __int_as_float(ent.y); q1 = __int_as_float(ent.z); q2 = __int_as_float(ent.w);
      }
      const int other = head & 0x0FFFFFFF, kind = (int)((unsigned)head >> 28);
      const V3 dr = sub(ld3(pos, other), pa);
      const float d = sqrtf(dot(dr, dr));
      float et, dEdd;
      if (kind == FMD_PRIOR_BONDS) {
        et = q0 * (d - q1) * (d - q1) + q2;
        dEdd = 2.0f * q0 * (d - q1);
      } else if (kind == FMD_PRIOR_POLY_BONDS) {
        // V0 + k1 d + k2 d^2 + k3 d^3 + k4 d^4 (PACKED only: V0 sits in the second float4 of the table row)
        const float v0 = PACKED ? __ldg(&pair_tab[2 * tab_id + 1]).x : 0.f;
        et = v0 + d * (q0 + d * (q1 + d * (q2 + d * q3)));
        dEdd = q0 + d * (2.0f * q1 + d * (3.0f * q2 + d * 4.0f * q3));
      } else {
        const float sg = q0 / d, rr = sg * sg;
        et = rr * rr * rr;
        dEdd = -6.0f * et / d;
      }
      // d = |x_other - x_a|  =>  dE/dx_a = -dEdd * dr/d ;  force = +dEdd * dr/d
      f = add(f, mul(dr, dEdd / d));
      e += 0.5f * et;
    }
  }
  // ---- three- and four-body terms: entry = term | role << 28 | is_dihedral << 30
  if (mb_ptr && valid) {
    const int q1 = __ldg(&mb_ptr[a + 1]);
    for (int q = __ldg(&mb_ptr[a]) + lane; q < q1; q += LPB) {
      const int ent = __ldg(&mb_ent[q]);
      const int t = ent & 0x0FFFFFFF, role = (ent >> 28) & 3, table = (int)((unsigned)ent >> 30);
      if (table == 0) {
        const int i = __ldg(&ang_map[t]), j = __ldg(&ang_map[n_ang + t]), k_ = __ldg(&ang_map[2 * n_ang + t]);
        const V3 d1 = sub(ld3(pos, i), ld3(pos, j)), d2 = sub(ld3(pos, k_), ld3(pos, j));
        const float n1 = sqrtf(dot(d1, d1)), n2 = sqrtf(dot(d2, d2));
        const float inv = 1.0f / (n1 * n2);
        const float c = dot(d1, d2) * inv;
        // per-term record: {p0..p3}, {p4, p5, V0, form}
        const float4 pa4 = __ldg(&ang_par[2 * t]), pb4 = __ldg(&ang_par[2 * t + 1]);
        const int form = __float_as_int(pb4.w);
        float et, dEdc;
        if (form == FMD_ANGLE_HARMONIC_COS) {
          et = pa4.x * (c - pa4.y) * (c - pa4.y);
          dEdc = 2.0f * pa4.x * (c - pa4.y);
        } else if (form == FMD_ANGLE_POLY_COS) {
          et = c * (pa4.x + c * (pa4.y + c * (pa4.z + c * (pa4.w + c * (pb4.x + c * pb4.y)))));
          dEdc = pa4.x + c * (2.0f * pa4.y + c * (3.0f * pa4.z + c * (4.0f * pa4.w + c * (5.0f * pb4.x + c * 6.0f * pb4.y))));
        } else if (form == FMD_ANGLE_RESTRICTED) {
          const float s2 = fmaxf(1.0f - c * c, 1e-12f);
          et = c * (pa4.w + c * (pa4.z + c * (pa4.y + c * pa4.x))) + pb4.x / s2;
          dEdc = pa4.w + c * (2.0f * pa4.z + c * (3.0f * pa4.y + c * 4.0f * pa4.x)) + 2.0f * pb4.x * c / (s2 * s2);
        } else {   // FMD_ANGLE_HARMONIC_RAW: k (theta - x0)^2, d theta / d cos = -1 / sin
          const float cc = fminf(fmaxf(c, -1.0f), 1.0f);
          const float th = acosf(cc), sn = sqrtf(fmaxf(1.0f - cc * cc, 1e-12f));
          et = pa4.x * (th - pa4.y) * (th - pa4.y);
          dEdc = -2.0f * pa4.x * (th - pa4.y) / sn;
        }
        et += pb4.z;
        const V3 gi = mul(sub(mul(d2, inv), mul(d1, c / (n1 * n1))), dEdc);
        const V3 gk = mul(sub(mul(d1, inv), mul(d2, c / (n2 * n2))), dEdc);
        const V3 g = role == 0 ? gi : (role == 2 ? gk : mul(add(gi, gk), -1.f));
        f = sub(f, g);
        e += et * (1.0f / 3.0f);
      } else {
        const int32_t* __restrict__ map = table == 1 ? dih_map : imp_map;
        const int nt = table == 1 ? n_dih : n_imp;
        const int i = __ldg(&map[t]), j = __ldg(&map[nt + t]), k_ = __ldg(&map[2 * nt + t]), l = __ldg(&map[3 * nt + t]);
        const V3 b1 = sub(ld3(pos, j), ld3(pos, i)), b2 = sub(ld3(pos, k_), ld3(pos, j)),
                 b3 = sub(ld3(pos, l), ld3(pos, k_));
        const V3 m = cross(b1, b2), n = cross(b2, b3);
        const float b2sq = dot(b2, b2), nb2 = sqrtf(b2sq);
        const float ys = nb2 * dot(b1, n), xc = dot(m, n);      // phi = atan2(ys, xc)
        float dEdphi = 0.f, et;
        if (table == 1) {
          // Fourier series in phi without a trigonometric call: (cos phi, sin phi) = (xc, ys) / |(xc, ys)|, the higher
          // harmonics by the angle-addition recurrence (the reference evaluates sin / cos of n * atan2(ys, xc))
          const float inv_r = rsqrtf(fmaxf(xc * xc + ys * ys, 1e-30f));
          const float c1 = xc * inv_r, s1 = ys * inv_r;
          float c = c1, sn = s1;
          et = dih_v0 ? __ldg(&dih_v0[t]) : 0.f;
          for (int d = 0; d < n_degs; ++d) {
            const float k1 = __ldg(&dih_k1[(size_t)t * n_degs + d]), k2 = __ldg(&dih_k2[(size_t)t * n_degs + d]);
            et += k1 * sn + k2 * c;
            dEdphi += (float)(d + 1) * (k1 * c - k2 * sn);
            const float cn = c * c1 - sn * s1;
            sn = sn * c1 + c * s1;
            c = cn;
          }
        } else {
          const float phi = atan2f(ys, xc);
          const float4 pp = __ldg(&imp_par[t]);     // {k, x0, V0, form}
          float x = phi;
          if (__float_as_int(pp.w) == FMD_IMPROPER_SHIFTED) x = (phi < 0.f ? phi + 2.0f * FMD_PI_F : phi) - FMD_PI_F;
          et = pp.x * (x - pp.y) * (x - pp.y) + pp.z;
          dEdphi = 2.0f * pp.x * (x - pp.y);
        }
        const V3 gi = mul(m, -nb2 / dot(m, m));
        const V3 gl = mul(n, nb2 / dot(n, n));
        const float s_ = dot(b1, b2) / b2sq, t_ = dot(b3, b2) / b2sq;
        V3 g;
        if (role == 0) g = gi;
        else if (role == 3) g = gl;
        else if (role == 1) g = add(mul(gi, -1.f - s_), mul(gl, t_));
        else g = add(mul(gl, -1.f - t_), mul(gi, s_));
        f = sub(f, mul(g, dEdphi));
        e += et * 0.25f;
      }
    }
  }
#pragma unroll
  for (int o = LPB / 2; o > 0; o >>= 1) {      // fixed butterfly order within the bead's lanes: deterministic
    f.x += __shfl_xor_sync(0xffffffffu, f.x, o);
    f.y += __shfl_xor_sync(0xffffffffu, f.y, o);
    f.z += __shfl_xor_sync(0xffffffffu, f.z, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if (lane == 0 && valid) {
    if (accumulate_forces) {
      forces[3 * a + 0] += f.x;
      forces[3 * a + 1] += f.y;
      forces[3 * a + 2] += f.z;
    } else {
      forces[3 * a + 0] = f.x;
      forces[3 * a + 1] = f.y;
      forces[3 * a + 2] = f.z;
    }
    e_atom[a] = e;
  }
}

}  // namespace

extern "C" int fmd_priors_csr(const float* pos, int n_nodes, const int32_t* pair_ptr, const void* pair_ent,
                              const float* pair_tab, const int32_t* mb_ptr, const int32_t* mb_ent, const int32_t* ang_map, int n_ang,
                              const float* ang_par, const int32_t* dih_map, int n_dih, const float* dih_k1,
                              const float* dih_k2, const float* dih_v0, int n_degs, const int32_t* imp_map, int n_imp,
                              const float* imp_par, float* e_atom, float* forces, int accumulate_forces, void* stream) {
  FMD_REQUIRE(pos && e_atom && forces, "fmd_priors_csr: bad arguments");
  FMD_REQUIRE(!pair_ptr || pair_ent, "fmd_priors_csr: pair_ptr without pair entries");
  FMD_REQUIRE(!mb_ptr || mb_ent, "fmd_priors_csr: mb_ptr without entries");
  FMD_REQUIRE(n_ang == 0 || (ang_map && ang_par), "fmd_priors_csr: missing angle tables");
  FMD_REQUIRE(n_dih == 0 || (dih_map && dih_k1 && dih_k2), "fmd_priors_csr: missing dihedral tables");
  FMD_REQUIRE(n_imp == 0 || (imp_map && imp_par), "fmd_priors_csr: missing improper tables");
  if (n_nodes <= 0) return FMD_OK;
  const int grid = fmd_div_up(n_nodes, WARPS * 32 / LPB);
  if (pair_tab)
    prior_csr_kernel<true><<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(
        pos, n_nodes, pair_ptr, pair_ent, (const float4*)pair_tab, mb_ptr, mb_ent, ang_map, n_ang, (const float4*)ang_par,
        dih_map, n_dih, dih_k1, dih_k2, dih_v0, n_degs, imp_map, n_imp, (const float4*)imp_par, e_atom, forces,
        accumulate_forces);
  else
    prior_csr_kernel<false><<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(
        pos, n_nodes, pair_ptr, pair_ent, nullptr, mb_ptr, mb_ent, ang_map, n_ang, (const float4*)ang_par,
        dih_map, n_dih, dih_k1, dih_k2, dih_v0, n_degs, imp_map, n_imp, (const float4*)imp_par, e_atom, forces,
        accumulate_forces);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
