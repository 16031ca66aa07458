// Node-level helpers, classical prior terms with analytic forces, BAOAB integrator with
// counter-based Philox noise, replica-exchange decision/swap.
#include "fmd_common.cuh"

using namespace fmd;

namespace {

// ---------------------------------------------------------------- embedding gather
template <typename IdxT>
__global__ void __launch_bounds__(256)
embedding_kernel(const float* __restrict__ table, const IdxT* __restrict__ types, int n_nodes, int F4,
                 float4* __restrict__ out) {
  const long long total = (long long)n_nodes * F4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / F4;
    const int f = (int)(i - n * F4);
    out[i] = __ldg(reinterpret_cast<const float4*>(table) + (long long)types[n] * F4 + f);
  }
}

// ---------------------------------------------------------------- per-molecule sum (deterministic)
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ v, const int32_t* __restrict__ mol_ptr, float* __restrict__ out,
                   int accumulate) {
  __shared__ float ws[8];
  const int b = blockIdx.x;
  const int lo = mol_ptr[b], hi = mol_ptr[b + 1];
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) s += v[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    out[b] = accumulate ? out[b] + t : t;
  }
}

// ---------------------------------------------------------------- priors
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 ld3(const float* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void atomic_sub3(float* f, int i, V3 g) {  // forces -= grad
  atomicAdd(&f[3 * i + 0], -g.x);
  atomicAdd(&f[3 * i + 1], -g.y);
  atomicAdd(&f[3 * i + 2], -g.z);
}

__global__ void __launch_bounds__(256)
prior_kernel(int kind, const float* __restrict__ pos, const int32_t* __restrict__ mapping,
             const int32_t* __restrict__ mapping_batch, int n_terms, const float* __restrict__ p0,
             const float* __restrict__ p1, const float* __restrict__ p2, int n_degs, float* __restrict__ energy,
             float* __restrict__ forces) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x) {
    float e = 0.f;
    if (kind == FMD_PRIOR_BONDS || kind == FMD_PRIOR_REPULSION) {
      const int i = mapping[t], j = mapping[n_terms + t];
      const V3 dr = sub(ld3(pos, j), ld3(pos, i));
      const float d = sqrtf(dot(dr, dr));
      float dEdd;
      if (kind == FMD_PRIOR_BONDS) {
        const float k = p0[t], x0 = p1[t];
        e = k * (d - x0) * (d - x0) + (p2 ? p2[t] : 0.f);
        dEdd = 2.0f * k * (d - x0);
      } else {
        const float sg = p0[t] / d;
        const float rr = sg * sg;
        e = rr * rr * rr;
        dEdd = -6.0f * e / d;
      }
      const V3 g = mul(dr, dEdd / d);  // dE/dr_j
      atomic_sub3(forces, j, g);
      atomic_sub3(forces, i, mul(g, -1.f));
    } else if (kind == FMD_PRIOR_ANGLES) {
      const int i = mapping[t], j = mapping[n_terms + t], k_ = mapping[2 * n_terms + t];
      const V3 d1 = sub(ld3(pos, i), ld3(pos, j)), d2 = sub(ld3(pos, k_), ld3(pos, j));
      const float n1 = sqrtf(dot(d1, d1)), n2 = sqrtf(dot(d2, d2));
      const float inv = 1.0f / (n1 * n2);
      const float c = dot(d1, d2) * inv;
      const float k = p0[t], x0 = p1[t];
      e = k * (c - x0) * (c - x0) + (p2 ? p2[t] : 0.f);
      const float dEdc = 2.0f * k * (c - x0);
      const V3 gi = mul(sub(mul(d2, inv), mul(d1, c / (n1 * n1))), dEdc);
      const V3 gk = mul(sub(mul(d1, inv), mul(d2, c / (n2 * n2))), dEdc);
      atomic_sub3(forces, i, gi);
      atomic_sub3(forces, k_, gk);
      atomic_sub3(forces, j, mul(add(gi, gk), -1.f));
    } else {  // dihedrals
      const int i = mapping[t], j = mapping[n_terms + t], k_ = mapping[2 * n_terms + t], l = mapping[3 * n_terms + t];
      const V3 b1 = sub(ld3(pos, j), ld3(pos, i)), b2 = sub(ld3(pos, k_), ld3(pos, j)), b3 = sub(ld3(pos, l), ld3(pos, k_));
      const V3 m = cross(b1, b2), n = cross(b2, b3);
      const float b2sq = dot(b2, b2), nb2 = sqrtf(b2sq);
      const float phi = atan2f(nb2 * dot(b1, n), dot(m, n));
      float dEdphi = 0.f;
      e = p2 ? p2[t] : 0.f;
      for (int q = 0; q < n_degs; ++q) {
        float s, c;
        sincosf((float)(q + 1) * phi, &s, &c);
        const float k1 = p0[(size_t)t * n_degs + q], k2 = p1[(size_t)t * n_degs + q];
        e += k1 * s + k2 * c;
        dEdphi += (float)(q + 1) * (k1 * c - k2 * s);
      }
      const V3 gi = mul(m, -nb2 / dot(m, m));
      const V3 gl = mul(n, nb2 / dot(n, n));
      const float s_ = dot(b1, b2) / b2sq, t_ = dot(b3, b2) / b2sq;
      const V3 gj = add(mul(gi, -1.f - s_), mul(gl, t_));
      const V3 gk = add(mul(gl, -1.f - t_), mul(gi, s_));
      atomic_sub3(forces, i, mul(gi, dEdphi));
      atomic_sub3(forces, j, mul(gj, dEdphi));
      atomic_sub3(forces, k_, mul(gk, dEdphi));
      atomic_sub3(forces, l, mul(gl, dEdphi));
    }
    atomicAdd(&energy[mapping_batch[t]], e);
  }
}

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}
// (0,1] uniform from 32 bits
__device__ __forceinline__ float u01_open0(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// [0,1) uniform from 32 bits (torch.rand convention)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ void normal3(uint64_t seed, uint64_t step, uint64_t node, float& z0, float& z1, float& z2) {
  uint32_t r[4];
  philox4x32_10(seed, node, step, r);
  const float u1 = u01_open0(r[0]), u2 = u01(r[1]), u3 = u01_open0(r[2]), u4 = u01(r[3]);
  const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = ra * c;
  z1 = ra * s;
  sincospif(2.0f * u4, &s, &c);
  z2 = rb * c;
}

__global__ void __launch_bounds__(256)
baoab_pre_kernel(float* __restrict__ pos, float* __restrict__ vel, const float* __restrict__ forces,
                 const float* __restrict__ inv_mass, const float* __restrict__ noise_std,
                 const float* __restrict__ noise, uint64_t seed, uint64_t step, const uint64_t* __restrict__ step_dev,
                 uint64_t node_offset, int n_nodes, float dt, float vscale, float noisescale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  if (step_dev) step += *step_dev;
  float z[3];
  if (noise) {
    z[0] = noise[3 * i]; z[1] = noise[3 * i + 1]; z[2] = noise[3 * i + 2];
  } else {
    normal3(seed, step, node_offset + (uint64_t)i, z[0], z[1], z[2]);
  }
  const float im = inv_mass[i], ns = noise_std[i];
  const float hdt = 0.5f * dt;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // B: v += dt/2 F/m ; A: x += dt/2 v ; O: v = v*vscale + noisescale*std*z ; A: x += dt/2 v
    float v = vel[3 * i + d] + hdt * forces[3 * i + d] * im;
    float x = pos[3 * i + d] + v * hdt;
    v = v * vscale + noisescale * (ns * z[d]);
    x = x + v * hdt;
    vel[3 * i + d] = v;
    pos[3 * i + d] = x;
  }
}

// Overdamped Langevin (Brownian) step: x += F D dt + sqrt(2 D dt) xi, dtau = D dt per bead
__global__ void __launch_bounds__(256)
overdamped_kernel(float* __restrict__ pos, const float* __restrict__ forces, const float* __restrict__ dtau,
                  const float* __restrict__ noise, uint64_t seed, uint64_t step, const uint64_t* __restrict__ step_dev,
                  uint64_t node_offset, int n_nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  if (step_dev) step += *step_dev;
  float z[3];
  if (noise) {
    z[0] = noise[3 * i]; z[1] = noise[3 * i + 1]; z[2] = noise[3 * i + 2];
  } else {
    normal3(seed, step, node_offset + (uint64_t)i, z[0], z[1], z[2]);
  }
  const float dt_ = dtau[i], amp = sqrtf(2.0f * dt_);
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[3 * i + d] = pos[3 * i + d] + forces[3 * i + d] * dt_ + amp * z[d];
}

__global__ void __launch_bounds__(256)
baoab_post_kernel(float* __restrict__ vel, const float* __restrict__ forces, const float* __restrict__ inv_mass,
                  int n_nodes, float dt, uint64_t* __restrict__ step_counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && step_counter) *step_counter += 1;      // the step's Philox counter (read by fmd_baoab_pre earlier in the step)
  if (i >= n_nodes) return;
  const float s = 0.5f * dt * inv_mass[i];
#pragma unroll
  for (int d = 0; d < 3; ++d) vel[3 * i + d] += s * forces[3 * i + d];
}

// kinetic energy per molecule, deterministic: one CTA per molecule
__global__ void __launch_bounds__(256)
kinetic_kernel(const float* __restrict__ vel, const float* __restrict__ inv_mass, const int32_t* __restrict__ mol_ptr,
               float* __restrict__ ke) {
  __shared__ float ws[8];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int i = mol_ptr[b] + threadIdx.x; i < mol_ptr[b + 1]; i += 256) {
    const float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
    s += (vx * vx + vy * vy + vz * vz) / inv_mass[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    ke[b] = 0.5f * t;
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(uint64_t seed, uint64_t step, int n_nodes, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  normal3(seed, step, (uint64_t)i, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

// ---------------------------------------------------------------- replica exchange
__global__ void __launch_bounds__(256)
pt_decide_kernel(const float* __restrict__ energy, const float* __restrict__ beta, const int32_t* __restrict__ pa,
                 const int32_t* __restrict__ pb, int n_pairs, const float* __restrict__ uniforms, uint64_t seed,
                 uint64_t exchange_index, int32_t* __restrict__ accept) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int a = pa[p], b = pb[p];
  const float prob = expf((energy[a] - energy[b]) * (beta[a] - beta[b]));
  float u;
  if (uniforms) {
    u = uniforms[p];
  } else {
    uint32_t r[4];
    philox4x32_10(seed ^ 0x5bd1e9955bd1e995ull, (uint64_t)p, exchange_index, r);
    u = u01(r[0]);
  }
  accept[p] = (u < prob) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
pt_swap_kernel(float* __restrict__ pos, float* __restrict__ vel, const float* __restrict__ beta,
               const int32_t* __restrict__ pa, const int32_t* __restrict__ pb, const int32_t* __restrict__ accept,
               int n_atoms) {
  const int p = blockIdx.x;
  if (!accept[p]) return;
  const int a = pa[p], b = pb[p];
  // reference parallel_tempering.py:465-477: v[a] <- v[b]*sqrt(beta_a/beta_b), v[b] <- v[a]*sqrt(beta_b/beta_a)
  const float s_ab = sqrtf(beta[a] / beta[b]), s_ba = sqrtf(beta[b] / beta[a]);
  const size_t oa = (size_t)a * n_atoms * 3, ob = (size_t)b * n_atoms * 3;
  for (int i = threadIdx.x; i < n_atoms * 3; i += blockDim.x) {
    const float xa = pos[oa + i], xb = pos[ob + i];
    pos[oa + i] = xb;
    pos[ob + i] = xa;
    const float va = vel[oa + i], vb = vel[ob + i];
    vel[oa + i] = vb * s_ab;
    vel[ob + i] = va * s_ba;
  }
}

}  // namespace

// =============================================================================== C ABI

extern "C" int fmd_embedding(const float* table, const void* types, int idx_bytes, int n_nodes, int n_feat, float* out,
                             void* stream) {
  FMD_REQUIRE(table && types && out, "fmd_embedding: bad arguments");
  FMD_REQUIRE(n_feat > 0 && n_feat % 4 == 0, "fmd_embedding: n_feat must be a positive multiple of 4");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_embedding: idx_bytes must be 4 or 8");
  if (n_nodes == 0) return FMD_OK;
  const int F4 = n_feat / 4;
  const int grid = min(fmd_div_up((long long)n_nodes * F4, 256), fmd_num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 4)
    embedding_kernel<int32_t><<<grid, 256, 0, st>>>(table, (const int32_t*)types, n_nodes, F4, (float4*)out);
  else
    embedding_kernel<int64_t><<<grid, 256, 0, st>>>(table, (const int64_t*)types, n_nodes, F4, (float4*)out);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_segment_sum(const float* e_atom, const int32_t* mol_ptr, int n_mols, float* out, int accumulate,
                               void* stream) {
  FMD_REQUIRE(e_atom && mol_ptr && out, "fmd_segment_sum: bad arguments");
  if (n_mols == 0) return FMD_OK;
  segment_sum_kernel<<<n_mols, 256, 0, (cudaStream_t)stream>>>(e_atom, mol_ptr, out, accumulate);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_prior_energy_forces(int kind, const float* pos, const int32_t* mapping,
                                       const int32_t* mapping_batch, int n_terms, const float* p0, const float* p1,
                                       const float* p2, int n_degs, float* energy, float* forces, void* stream) {
  FMD_REQUIRE(kind >= 0 && kind <= 3, "fmd_prior_energy_forces: unknown prior kind");
  FMD_REQUIRE(pos && mapping && mapping_batch && p0 && energy && forces, "fmd_prior_energy_forces: bad arguments");
  FMD_REQUIRE(kind == FMD_PRIOR_REPULSION || p1, "fmd_prior_energy_forces: missing parameter vector");
  if (n_terms == 0) return FMD_OK;
  const int grid = min(fmd_div_up(n_terms, 256), fmd_num_sms() * 16);
  prior_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, pos, mapping, mapping_batch, n_terms, p0, p1, p2, n_degs,
                                                       energy, forces);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

// Last output layer (hidden -> 1, no bias) and the first step of its backward in one pass over y:
//   e_atom[i] = sum_k y[i,k] w[k] ;  g_y[i,k] = w[k] (1 - y[i,k]^2)      (dE/d e_atom = 1)
// One warp per node; replaces a [N,K]x[K,1] and a [N,1]x[1,K] launch of the generic dense kernel.
template <typename T>
__global__ void __launch_bounds__(256)
out_head_kernel(const T* __restrict__ y, const T* __restrict__ w, int n_nodes, int K, float* __restrict__ e_atom,
                T* __restrict__ g_y) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_nodes) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float yv = to_f32<T>(y[(size_t)i * K + k]), wv = to_f32<T>(w[k]);
    acc = fmaf(yv, wv, acc);
    if (g_y) g_y[(size_t)i * K + k] = from_f32<T>(wv * (1.0f - yv * yv));
  }
  acc = warp_sum(acc);
  if (lane == 0) e_atom[i] = acc;
}

extern "C" int fmd_out_head(const void* y, const void* w, int dt, int n_nodes, int n_hidden, float* e_atom, void* g_y,
                            void* stream) {
  FMD_REQUIRE(y && w && e_atom && n_hidden > 0, "fmd_out_head: bad arguments");
  FMD_REQUIRE(dt == FMD_F32 || dt == FMD_F16, "fmd_out_head: bad dtype");
  if (n_nodes <= 0) return FMD_OK;
  const int grid = fmd_div_up((long long)n_nodes * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == FMD_F32)
    out_head_kernel<float><<<grid, 256, 0, st>>>((const float*)y, (const float*)w, n_nodes, n_hidden, e_atom, (float*)g_y);
  else
    out_head_kernel<__half><<<grid, 256, 0, st>>>((const __half*)y, (const __half*)w, n_nodes, n_hidden, e_atom,
                                                   (__half*)g_y);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

__global__ void increment_u64_kernel(uint64_t* p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *p += 1;
}

extern "C" int fmd_increment_u64(uint64_t* counter, void* stream) {
  FMD_REQUIRE(counter, "fmd_increment_u64: bad arguments");
  increment_u64_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_baoab_pre(float* pos, float* vel, const float* forces, const float* inv_mass, const float* noise_std,
                             const float* noise, uint64_t seed, uint64_t step, const uint64_t* step_dev, uint64_t node_offset, int n_nodes,
                             float dt, float vscale, float noisescale, void* stream) {
  FMD_REQUIRE(pos && vel && forces && inv_mass && noise_std, "fmd_baoab_pre: bad arguments");
  if (n_nodes == 0) return FMD_OK;
  baoab_pre_kernel<<<fmd_div_up(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(pos, vel, forces, inv_mass, noise_std,
                                                                                noise, seed, step, step_dev, node_offset, n_nodes, dt,
                                                                                vscale, noisescale);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_overdamped_step(float* pos, const float* forces, const float* dtau, const float* noise, uint64_t seed,
                                   uint64_t step, const uint64_t* step_dev, uint64_t node_offset, int n_nodes,
                                   void* stream) {
  FMD_REQUIRE(pos && forces && dtau, "fmd_overdamped_step: bad arguments");
  if (n_nodes == 0) return FMD_OK;
  overdamped_kernel<<<fmd_div_up(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(pos, forces, dtau, noise, seed, step,
                                                                                 step_dev, node_offset, n_nodes);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_baoab_post(float* vel, const float* forces, const float* inv_mass, int n_nodes, float dt,
                              const int32_t* mol_ptr, int n_mols, float* ke, uint64_t* step_counter, void* stream) {
  FMD_REQUIRE(vel && forces && inv_mass, "fmd_baoab_post: bad arguments");
  FMD_REQUIRE(!ke || mol_ptr, "fmd_baoab_post: ke requires mol_ptr");
  if (n_nodes == 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  baoab_post_kernel<<<fmd_div_up(n_nodes, 256), 256, 0, st>>>(vel, forces, inv_mass, n_nodes, dt, step_counter);
  if (ke && n_mols > 0) kinetic_kernel<<<n_mols, 256, 0, st>>>(vel, inv_mass, mol_ptr, ke);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_philox_normal(uint64_t seed, uint64_t step, int n_nodes, float* out, void* stream) {
  FMD_REQUIRE(out, "fmd_philox_normal: bad arguments");
  if (n_nodes == 0) return FMD_OK;
  philox_normal_kernel<<<fmd_div_up(n_nodes, 256), 256, 0, (cudaStream_t)stream>>>(seed, step, n_nodes, out);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_pt_decide(const float* energy, const float* beta, const int32_t* pair_a, const int32_t* pair_b,
                             int n_pairs, const float* uniforms, uint64_t seed, uint64_t exchange_index,
                             int32_t* accept, void* stream) {
  FMD_REQUIRE(energy && beta && pair_a && pair_b && accept, "fmd_pt_decide: bad arguments");
  if (n_pairs == 0) return FMD_OK;
  pt_decide_kernel<<<fmd_div_up(n_pairs, 256), 256, 0, (cudaStream_t)stream>>>(energy, beta, pair_a, pair_b, n_pairs,
                                                                               uniforms, seed, exchange_index, accept);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_pt_swap(float* pos, float* vel, const float* beta, const int32_t* pair_a, const int32_t* pair_b,
                           const int32_t* accept, int n_pairs, int n_atoms, void* stream) {
  FMD_REQUIRE(pos && vel && beta && pair_a && pair_b && accept, "fmd_pt_swap: bad arguments");
  if (n_pairs == 0 || n_atoms == 0) return FMD_OK;
  pt_swap_kernel<<<n_pairs, 256, 0, (cudaStream_t)stream>>>(pos, vel, beta, pair_a, pair_b, accept, n_atoms);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
