// Edge-level kernels: distance + Gaussian RBF + cosine cutoff (fwd/bwd), edge-gradient -> forces,
// CFConv CSR segment reduce (both directions), filter gradient.
//
// All of these are HBM-bound streaming kernels (see DESIGN.md for algorithmic bytes):
//   * edge rows (filter [E,F], rbf [E,R]) are read/written exactly once, fully coalesced;
//   * node rows (x [N,F], pos [N,3]) are gathered and stay L2-resident (N*F*4 = 17.6 MB at cfg2);
//   * no atomics on the step path; reductions use fixed lane assignment + xor-shuffle trees, so
//     results are bitwise reproducible run to run.
#include <cstdlib>

#include "fmd_common.cuh"

using namespace fmd;

namespace {

__device__ __forceinline__ int edge_count(int n_edges, const int32_t* n_edges_dev) {
  return n_edges_dev ? min(n_edges, *n_edges_dev) : n_edges;
}

// ---------------------------------------------------------------- dist + rbf + cutoff (fwd)
constexpr int RBF_EDGES = 64;  // edges per CTA iteration

template <typename IdxT>
__global__ void __launch_bounds__(256)
dist_rbf_fwd_kernel(const float* __restrict__ pos, const IdxT* __restrict__ src, const IdxT* __restrict__ dst,
                    int n_edges, const int32_t* __restrict__ n_edges_dev, const float* __restrict__ centers, int R,
                    float gamma, float rc, float* __restrict__ dist, float* __restrict__ rbf) {
  __shared__ float sd[RBF_EDGES], sc[RBF_EDGES];
  extern __shared__ float s_centers[];
  const int E = edge_count(n_edges, n_edges_dev);
  for (int k = threadIdx.x; k < R; k += blockDim.x) s_centers[k] = centers[k];
  for (long long e0 = (long long)blockIdx.x * RBF_EDGES; e0 < E; e0 += (long long)gridDim.x * RBF_EDGES) {
    __syncthreads();
    if (threadIdx.x < RBF_EDGES) {
      const long long e = e0 + threadIdx.x;
      float d = 0.f, c = 0.f;
      if (e < E) {
        const long long s = (long long)src[e], t = (long long)dst[e];
        const float dx = pos[3 * t + 0] - pos[3 * s + 0];
        const float dy = pos[3 * t + 1] - pos[3 * s + 1];
        const float dz = pos[3 * t + 2] - pos[3 * s + 2];
        d = sqrtf(dx * dx + dy * dy + dz * dz);
        c = cosine_cutoff(d, rc);
        if (dist) dist[e] = d;
      }
      sd[threadIdx.x] = d;
      sc[threadIdx.x] = c;
    }
    __syncthreads();
    if (rbf) {
      const int ne = (int)min((long long)RBF_EDGES, (long long)E - e0);
      const int total = ne * R;
      float* out = rbf + e0 * R;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int el = i / R, k = i - el * R;
        const float diff = sd[el] - s_centers[k];
        out[i] = expf(gamma * diff * diff) * sc[el];
      }
    }
  }
}

// ---------------------------------------------------------------- rbf backward: g_d
__global__ void __launch_bounds__(256)
rbf_bwd_kernel(const float* __restrict__ dist, const float* __restrict__ grad_rbf, const float* __restrict__ grad_dist,
               int n_edges, const int32_t* __restrict__ n_edges_dev, const float* __restrict__ centers, int R,
               float gamma, float rc, float* __restrict__ g_d, int accumulate) {
  const int E = edge_count(n_edges, n_edges_dev);
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int e = wid; e < E; e += nw) {
    const float d = dist[e];
    const float C = cosine_cutoff(d, rc), dC = cosine_cutoff_grad(d, rc);
    float acc = 0.f;
    for (int k = lane; k < R; k += 32) {
      const float diff = d - centers[k];
      const float ex = expf(gamma * diff * diff);
      const float drbf = 2.0f * gamma * diff * ex * C + ex * dC;
      acc += grad_rbf[(size_t)e * R + k] * drbf;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (grad_dist) acc += grad_dist[e];
      g_d[e] = accumulate ? g_d[e] + acc : acc;
    }
  }
}

// ---------------------------------------------------------------- g_d -> grad_pos (atomic, generic lists)
template <typename IdxT>
__global__ void __launch_bounds__(256)
edge_grad_to_pos_atomic_kernel(const float* __restrict__ pos, const IdxT* __restrict__ src, const IdxT* __restrict__ dst,
                               const float* __restrict__ dist, const float* __restrict__ g_d, int n_edges,
                               const int32_t* __restrict__ n_edges_dev, float* __restrict__ grad_pos) {
  const int E = edge_count(n_edges, n_edges_dev);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const long long s = (long long)src[e], t = (long long)dst[e];
    const float inv = 1.0f / fmaxf(dist[e], 1e-8f);
    const float g = g_d[e] * inv;
    const float gx = g * (pos[3 * t + 0] - pos[3 * s + 0]);
    const float gy = g * (pos[3 * t + 1] - pos[3 * s + 1]);
    const float gz = g * (pos[3 * t + 2] - pos[3 * s + 2]);
    atomicAdd(&grad_pos[3 * t + 0], gx);
    atomicAdd(&grad_pos[3 * t + 1], gy);
    atomicAdd(&grad_pos[3 * t + 2], gz);
    atomicAdd(&grad_pos[3 * s + 0], -gx);
    atomicAdd(&grad_pos[3 * s + 1], -gy);
    atomicAdd(&grad_pos[3 * s + 2], -gz);
  }
}

// ---------------------------------------------------------------- g_d -> forces (CSR, deterministic)
__global__ void __launch_bounds__(256)
edge_grad_to_forces_csr_kernel(const float* __restrict__ pos, const int32_t* __restrict__ seg_ptr,
                               const int32_t* __restrict__ dst, const int32_t* __restrict__ rev,
                               const float* __restrict__ dist, const float* __restrict__ g_d, int n_nodes, int n_edges,
                               float sign, float* __restrict__ out, int accumulate, int pair_mode) {
  // eight lanes per node (four nodes per warp): the kernel is a chain of dependent loads (segment bounds -> edge record ->
  // pair gradient / partner position), so more nodes in flight per warp is what it needs; every lane runs the same number of
  // iterations of the outer loop (uniform bound), the lane reduction uses the full-warp mask
  const int lane = threadIdx.x & 7;
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int ng = (gridDim.x * blockDim.x) >> 3;
  const int n_iter = (n_nodes + ng - 1) / ng;
  for (int it = 0; it < n_iter; ++it) {
    const int i_raw = gid + it * ng;
    const bool valid = i_raw < n_nodes;
    const int i = valid ? i_raw : n_nodes - 1;
    const int a = min(seg_ptr[i], n_edges), b = valid ? min(seg_ptr[i + 1], n_edges) : a;
    const float px = pos[3 * i + 0], py = pos[3 * i + 1], pz = pos[3 * i + 2];
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll 4
    for (int e = a + lane; e < b; e += 8) {
      const int j = dst[e];
      const int r = rev[e];
      float g;
      if (pair_mode) {            // rev = pair index of the edge, g_d = the pair's (already symmetrised) gradient
        g = r >= 0 ? g_d[r] : 0.f;
      } else {
        g = g_d[e];
        if (r >= 0) g += g_d[r];
      }
      g *= 1.0f / fmaxf(dist[e], 1e-8f);
      fx += g * (pos[3 * j + 0] - px);
      fy += g * (pos[3 * j + 1] - py);
      fz += g * (pos[3 * j + 2] - pz);
    }
#pragma unroll
    for (int o_ = 4; o_ > 0; o_ >>= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, o_);
      fy += __shfl_xor_sync(0xffffffffu, fy, o_);
      fz += __shfl_xor_sync(0xffffffffu, fz, o_);
    }
    if (lane == 0 && valid) {
      float* o = out + 3 * (size_t)i;
      if (accumulate) {
        o[0] += sign * fx;
        o[1] += sign * fy;
        o[2] += sign * fz;
      } else {
        o[0] = sign * fx;
        o[1] = sign * fy;
        o[2] = sign * fz;
      }
    }
  }
}

// ---------------------------------------------------------------- CFConv CSR segment reduce
// One warp per destination node; lane l owns features [4l, 4l+4) of every 128-wide feature chunk.
// Per batch of 32 edges the lanes cooperatively load index / distance (coalesced), compute the
// cutoff once per edge, then the warp streams the filter rows (coalesced 8B/16B per lane, L1
// no-allocate) against the gathered x rows (16B per lane, L1/L2 resident), 4 edges in flight.
template <typename WT, typename IdxT, bool PERM>
__global__ void __launch_bounds__(256)
cfconv_csr_kernel(const float* __restrict__ x, const WT* __restrict__ filt, const float* __restrict__ dist,
                  const IdxT* __restrict__ gather, const IdxT* __restrict__ seg_ptr, const IdxT* __restrict__ perm,
                  int n_nodes, long long n_edges, int F, float rc, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < n_nodes; i += nw) {
    const long long a = min((long long)seg_ptr[i], n_edges), b = min((long long)seg_ptr[i + 1], n_edges);
    for (int fc = 0; fc < F; fc += 128) {
      // every lane runs the loop (warp shuffles inside); lanes past F only skip the loads/stores
      const int f0 = fc + lane * 4;
      const bool fa = f0 < F;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long base = a; base < b; base += 32) {
        const long long p = base + lane;
        long long e_l = 0, j_l = 0;
        float c_l = 0.f;
        if (p < b) {
          e_l = PERM ? (long long)perm[p] : p;
          j_l = (long long)gather[e_l];
          c_l = cosine_cutoff(dist[e_l], rc);
        }
        const int cnt = (int)min((long long)32, b - base);
        // lanes >= cnt hold c=0 and (e,j) of a valid row (lane 0's) so the unrolled loads stay in bounds
        const long long e0v = __shfl_sync(0xffffffffu, e_l, 0), j0v = __shfl_sync(0xffffffffu, j_l, 0);
        if (lane >= cnt) { e_l = e0v; j_l = j0v; }
        for (int k = 0; k < cnt; k += 4) {
          long long e[4], j[4];
          float c[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = min(k + u, 31);
            e[u] = __shfl_sync(0xffffffffu, e_l, kk);
            j[u] = __shfl_sync(0xffffffffu, j_l, kk);
            c[u] = (k + u < 32) ? __shfl_sync(0xffffffffu, c_l, kk) : 0.f;
          }
          float4 w[4], xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            w[u] = fa ? load4_stream(filt + e[u] * F + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
            xv[u] = fa ? load4(x + j[u] * F + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc.x = fmaf(xv[u].x * w[u].x, c[u], acc.x);
            acc.y = fmaf(xv[u].y * w[u].y, c[u], acc.y);
            acc.z = fmaf(xv[u].z * w[u].z, c[u], acc.z);
            acc.w = fmaf(xv[u].w * w[u].w, c[u], acc.w);
          }
        }
      }
      if (fa) store4(out + (size_t)i * F + f0, acc);
    }
  }
}

// ---------------------------------------------------------------- CFConv CSR segment reduce, F = 128, list order
// Specialisation for the layout the engine produces (F = 128, int32 CSR, no permutation): a HALF-warp per edge, lane
// owns 8 features (one 16-byte load of an fp16 filter row / two of an fp32 row, two 16-byte loads of the gathered x
// row), so one warp-wide load instruction moves two edges and U slots keep 2U edges in flight per warp -- the generic
// kernel above has 4 edges (1 KB of fp16 filter) in flight per warp and sat at 48 % of the HBM roofline, latency-bound.
// The two half-warp partial sums are combined once per node in a fixed order (deterministic).
__device__ __forceinline__ void load8_stream(const float* p, float4& a, float4& b) {
  a = load4_stream(p);
  b = load4_stream(p + 4);
}
__device__ __forceinline__ void load8_stream(const __half* p, float4& a, float4& b) {
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  const float2 f2 = __half22float2(*reinterpret_cast<__half2*>(&u.z)), f3 = __half22float2(*reinterpret_cast<__half2*>(&u.w));
  a = make_float4(f0.x, f0.y, f1.x, f1.y);
  b = make_float4(f2.x, f2.y, f3.x, f3.y);
}

template <typename WT, int U>
__global__ void __launch_bounds__(256)
cfconv_csr128_kernel(const float* __restrict__ x, const WT* __restrict__ filt, const float* __restrict__ dist,
                     const int32_t* __restrict__ gather, const int32_t* __restrict__ seg_ptr, int n_nodes, int n_edges,
                     float rc, float* __restrict__ out) {
  constexpr int F = 128;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int h = lane >> 4, f0 = (lane & 15) * 8;
  for (int i = wid; i < n_nodes; i += nw) {
    const int a = min(seg_ptr[i], n_edges), b = min(seg_ptr[i + 1], n_edges);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int base = a; base < b; base += 32) {
      const int p = base + lane;
      int j_l = 0;
      float c_l = 0.f;
      if (p < b) {
        j_l = gather[p];
        c_l = cosine_cutoff(dist[p], rc);
      }
      const int cnt = min(32, b - base);
      for (int k = 0; k < cnt; k += 2 * U) {
        int j[U], e[U];
        float c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kk = k + 2 * u + h;
          const bool valid = kk < cnt;
          const int kc = valid ? kk : 0;          // past the end: edge `base` again with weight 0 (loads stay in bounds)
          j[u] = __shfl_sync(0xffffffffu, j_l, kc);
          const float cv = __shfl_sync(0xffffffffu, c_l, kc);
          c[u] = valid ? cv : 0.f;
          e[u] = base + kc;
        }
        float4 w0[U], w1[U], x0[U], x1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const WT* wp = filt + (size_t)e[u] * F + f0;
          load8_stream(wp, w0[u], w1[u]);
          const float* xp = x + (size_t)j[u] * F + f0;
          x0[u] = load4(xp);
          x1[u] = load4(xp + 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[0] = fmaf(x0[u].x * w0[u].x, c[u], acc[0]);
          acc[1] = fmaf(x0[u].y * w0[u].y, c[u], acc[1]);
          acc[2] = fmaf(x0[u].z * w0[u].z, c[u], acc[2]);
          acc[3] = fmaf(x0[u].w * w0[u].w, c[u], acc[3]);
          acc[4] = fmaf(x1[u].x * w1[u].x, c[u], acc[4]);
          acc[5] = fmaf(x1[u].y * w1[u].y, c[u], acc[5]);
          acc[6] = fmaf(x1[u].z * w1[u].z, c[u], acc[6]);
          acc[7] = fmaf(x1[u].w * w1[u].w, c[u], acc[7]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 16);
    if (h == 0) {
      float* o = out + (size_t)i * F + f0;
      store4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
      store4(o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
  }
}

// ---------------------------------------------------------------- filter gradient (+ exact cutoff term)
// One warp per batch of 32 consecutive edges: the lanes load the batch's indices / distances coalesced, then the warp
// walks the batch 4 edges at a time with all 8 (12 with the filter rows) 16-byte loads per lane issued before the
// first use -- one edge per iteration left the kernel latency-bound at 20 % of the HBM roofline.  The per-edge dot
// products of the exact cut-off term are reduced with a fixed xor tree and written back coalesced (lane = edge).
template <typename YT, typename WT, typename IdxT>
__global__ void __launch_bounds__(256)
grad_filter_kernel(const float* __restrict__ x, const float* __restrict__ g_out, const float* __restrict__ dist,
                   const IdxT* __restrict__ src, const IdxT* __restrict__ dst, int n_edges,
                   const int32_t* __restrict__ n_edges_dev, int F, float rc, YT* __restrict__ g_filt,
                   const WT* __restrict__ filt, float* __restrict__ g_dcut, int accumulate_dcut) {
  const int E = edge_count(n_edges, n_edges_dev);
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const bool want_dot = g_dcut && filt;
  for (long long base = (long long)wid * 32; base < E; base += (long long)nw * 32) {
    const int cnt = (int)min((long long)32, (long long)E - base);
    long long s_l = 0, t_l = 0;
    float d_l = 0.f, c_l = 0.f;
    if (lane < cnt) {
      s_l = (long long)src[base + lane];
      t_l = (long long)dst[base + lane];
      d_l = dist[base + lane];
      c_l = cosine_cutoff(d_l, rc);
    }
    float dot_l = 0.f;   // lane k ends up with the dot product of edge base + k
    for (int fc = 0; fc < F; fc += 128) {
      const int f0 = fc + lane * 4;
      const bool fa = f0 < F;
      for (int k = 0; k < cnt; k += 4) {
        long long s[4], t[4];
        float c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int kk = min(k + u, cnt - 1);      // tail: repeat the last valid edge (its stores are masked)
          s[u] = __shfl_sync(0xffffffffu, s_l, kk);
          t[u] = __shfl_sync(0xffffffffu, t_l, kk);
          c[u] = __shfl_sync(0xffffffffu, c_l, kk);
        }
        float4 xv[4], gv[4], wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xv[u] = fa ? load4(x + s[u] * F + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
          gv[u] = fa ? load4(g_out + t[u] * F + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[u] = (fa && want_dot) ? load4_stream(filt + (size_t)(base + min(k + u, cnt - 1)) * F + f0)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 pr = make_float4(xv[u].x * gv[u].x, xv[u].y * gv[u].y, xv[u].z * gv[u].z, xv[u].w * gv[u].w);
          if (g_filt && fa && k + u < cnt)
            store4(g_filt + (size_t)(base + k + u) * F + f0, make_float4(pr.x * c[u], pr.y * c[u], pr.z * c[u], pr.w * c[u]));
          if (want_dot) {
            const float dsum = warp_sum(pr.x * wv[u].x + pr.y * wv[u].y + pr.z * wv[u].z + pr.w * wv[u].w);
            if (lane == k + u) dot_l += dsum;
          }
        }
      }
    }
    if (want_dot && lane < cnt) {
      const float v = cosine_cutoff_grad(d_l, rc) * dot_l;
      g_dcut[base + lane] = accumulate_dcut ? g_dcut[base + lane] + v : v;
    }
  }
}

// ---------------------------------------------------------------- rbf backward, tiled (contiguous [E,R] rows)
// CTA = 128 consecutive edges: the [128, R] block of grad_rbf is one contiguous span, loaded with coalesced 16-byte
// loads into shared memory (row stride R + 1: conflict-free), then thread e reduces its own row against
// d rbf_k / d d computed in registers.  Replaces the warp-per-edge kernel (200-byte rows, one edge in flight per warp).
constexpr int RB_EDGES = 128;
__global__ void __launch_bounds__(RB_EDGES)
rbf_bwd_tiled_kernel(const float* __restrict__ dist, const float* __restrict__ grad_rbf, const float* __restrict__ grad_dist,
                     int n_edges, const int32_t* __restrict__ n_edges_dev, const float* __restrict__ centers, int R,
                     float gamma, float rc, float* __restrict__ g_d, int accumulate) {
  extern __shared__ float rb_smem[];
  float* s_c = rb_smem;                       // [R]
  float* s_g = rb_smem + ((R + 3) & ~3);      // [RB_EDGES][R + 1]
  const int E = edge_count(n_edges, n_edges_dev);
  const int ld = R + 1;
  for (int k = threadIdx.x; k < R; k += RB_EDGES) s_c[k] = centers[k];
  for (long long e0 = (long long)blockIdx.x * RB_EDGES; e0 < E; e0 += (long long)gridDim.x * RB_EDGES) {
    const int ne = (int)min((long long)RB_EDGES, (long long)E - e0);
    const int total = ne * R;                 // floats of this block; e0 * R * 4 bytes is 16-byte aligned (RB_EDGES % 4 == 0)
    const float* gsrc = grad_rbf + e0 * R;
    __syncthreads();
    for (int i4 = threadIdx.x * 4; i4 < total; i4 += RB_EDGES * 4) {
      float v[4];
      if (i4 + 3 < total) {
        const float4 t = load4_stream(gsrc + i4);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = i4 + q < total ? gsrc[i4 + q] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i4 + q;
        if (i < total) { const int el = i / R; s_g[el * ld + (i - el * R)] = v[q]; }
      }
    }
    __syncthreads();
    const int el = threadIdx.x;
    if (el < ne) {
      const float d = dist[e0 + el];
      const float C = cosine_cutoff(d, rc), dC = cosine_cutoff_grad(d, rc);
      float acc = 0.f;
      for (int k = 0; k < R; ++k) {
        const float diff = d - s_c[k];
        const float ex = expf(gamma * diff * diff);
        acc = fmaf(s_g[el * ld + k], 2.0f * gamma * diff * ex * C + ex * dC, acc);
      }
      if (grad_dist) acc += grad_dist[e0 + el];
      g_d[e0 + el] = accumulate ? g_d[e0 + el] + acc : acc;
    }
  }
}

}  // namespace

// =============================================================================== C ABI

extern "C" int fmd_dist_rbf_cutoff_fwd(const float* pos, const void* edge_src, const void* edge_dst, int idx_bytes,
                                       int n_edges, const int32_t* n_edges_dev, const float* centers, int num_rbf,
                                       float gamma, float rc, float* dist, float* rbf, void* stream) {
  FMD_REQUIRE(pos && edge_src && edge_dst && (dist || rbf), "fmd_dist_rbf_cutoff_fwd: bad arguments");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_dist_rbf_cutoff_fwd: idx_bytes must be 4 or 8");
  FMD_REQUIRE(!rbf || (centers && num_rbf > 0 && num_rbf <= 4096), "fmd_dist_rbf_cutoff_fwd: bad rbf parameters");
  if (n_edges == 0) return FMD_OK;
  const int grid = min(fmd_div_up(n_edges, RBF_EDGES), fmd_num_sms() * 16);
  const size_t smem = sizeof(float) * (size_t)(num_rbf > 0 ? num_rbf : 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 4)
    dist_rbf_fwd_kernel<int32_t><<<grid, 256, smem, st>>>(pos, (const int32_t*)edge_src, (const int32_t*)edge_dst,
                                                          n_edges, n_edges_dev, centers, rbf ? num_rbf : 0, gamma, rc,
                                                          dist, rbf);
  else
    dist_rbf_fwd_kernel<int64_t><<<grid, 256, smem, st>>>(pos, (const int64_t*)edge_src, (const int64_t*)edge_dst,
                                                          n_edges, n_edges_dev, centers, rbf ? num_rbf : 0, gamma, rc,
                                                          dist, rbf);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_rbf_bwd(const float* dist, const float* grad_rbf, const float* grad_dist, int n_edges,
                           const int32_t* n_edges_dev, const float* centers, int num_rbf, float gamma, float rc,
                           float* g_d, int accumulate, void* stream) {
  FMD_REQUIRE(dist && grad_rbf && centers && g_d && num_rbf > 0, "fmd_rbf_bwd: bad arguments");
  if (n_edges == 0) return FMD_OK;
  if (num_rbf <= 96) {
    const size_t smem = sizeof(float) * (size_t)(((num_rbf + 3) & ~3) + RB_EDGES * (num_rbf + 1));
    static bool attr_done = false;
    if (!attr_done) {
      FMD_CUDA(cudaFuncSetAttribute(rbf_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_done = true;
    }
    const int grid = min(fmd_div_up(n_edges, RB_EDGES), fmd_num_sms() * 8);
    rbf_bwd_tiled_kernel<<<grid, RB_EDGES, smem, (cudaStream_t)stream>>>(dist, grad_rbf, grad_dist, n_edges, n_edges_dev,
                                                                     centers, num_rbf, gamma, rc, g_d, accumulate);
  } else {
    const int grid = min(fmd_div_up((long long)n_edges * 32, 256), fmd_num_sms() * 16);
    rbf_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dist, grad_rbf, grad_dist, n_edges, n_edges_dev, centers,
                                                           num_rbf, gamma, rc, g_d, accumulate);
  }
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_edge_grad_to_pos_atomic(const float* pos, const void* edge_src, const void* edge_dst, int idx_bytes,
                                           const float* dist, const float* g_d, int n_edges,
                                           const int32_t* n_edges_dev, float* grad_pos, void* stream) {
  FMD_REQUIRE(pos && edge_src && edge_dst && dist && g_d && grad_pos, "fmd_edge_grad_to_pos_atomic: bad arguments");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_edge_grad_to_pos_atomic: idx_bytes must be 4 or 8");
  if (n_edges == 0) return FMD_OK;
  const int grid = min(fmd_div_up(n_edges, 256), fmd_num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 4)
    edge_grad_to_pos_atomic_kernel<int32_t><<<grid, 256, 0, st>>>(pos, (const int32_t*)edge_src,
                                                                  (const int32_t*)edge_dst, dist, g_d, n_edges,
                                                                  n_edges_dev, grad_pos);
  else
    edge_grad_to_pos_atomic_kernel<int64_t><<<grid, 256, 0, st>>>(pos, (const int64_t*)edge_src,
                                                                  (const int64_t*)edge_dst, dist, g_d, n_edges,
                                                                  n_edges_dev, grad_pos);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_edge_grad_to_forces_csr(const float* pos, const int32_t* seg_ptr, const int32_t* edge_dst,
                                           const int32_t* rev, const float* dist, const float* g_d, int n_nodes,
                                           int n_edges, float sign, float* out, int accumulate, int pair_mode,
                                           void* stream) {
  FMD_REQUIRE(pos && seg_ptr && edge_dst && rev && dist && g_d && out, "fmd_edge_grad_to_forces_csr: bad arguments");
  if (n_nodes == 0) return FMD_OK;
  const int grid = min(fmd_div_up((long long)n_nodes * 8, 256), fmd_num_sms() * 16);
  edge_grad_to_forces_csr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, seg_ptr, edge_dst, rev, dist, g_d,
                                                                         n_nodes, n_edges, sign, out, accumulate, pair_mode);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

template <typename WT, typename IdxT>
static void launch_cfconv(const float* x, const void* filt, const float* dist, const void* gather, const void* seg_ptr,
                          const void* perm, int n_nodes, int n_edges, int F, float rc, float* out, cudaStream_t st) {
  const int grid = min(fmd_div_up((long long)n_nodes * 32, 256), fmd_num_sms() * 8);
  if (perm)
    cfconv_csr_kernel<WT, IdxT, true><<<grid, 256, 0, st>>>(x, (const WT*)filt, dist, (const IdxT*)gather,
                                                            (const IdxT*)seg_ptr, (const IdxT*)perm, n_nodes, n_edges, F,
                                                            rc, out);
  else
    cfconv_csr_kernel<WT, IdxT, false><<<grid, 256, 0, st>>>(x, (const WT*)filt, dist, (const IdxT*)gather,
                                                             (const IdxT*)seg_ptr, nullptr, n_nodes, n_edges, F, rc, out);
}

extern "C" int fmd_cfconv_csr(const float* x, const void* filt, int wdt, const float* dist, const void* gather,
                              const void* seg_ptr, const void* perm, int idx_bytes, int n_nodes, int n_edges, int n_feat,
                              float rc, float* out, void* stream) {
  FMD_REQUIRE(x && filt && dist && gather && seg_ptr && out, "fmd_cfconv_csr: bad arguments");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_cfconv_csr: idx_bytes must be 4 or 8");
  FMD_REQUIRE(wdt == FMD_F32 || wdt == FMD_F16, "fmd_cfconv_csr: bad filter dtype");
  FMD_REQUIRE(n_feat > 0 && n_feat % 4 == 0, "fmd_cfconv_csr: n_feat must be a positive multiple of 4");
  if (n_nodes == 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_feat == 128 && idx_bytes == 4 && !perm) {
    static int u_sel = -1;   // edges in flight per warp = 2U; FMD_CSR128_U=4|8 overrides the default (tuning knob)
    if (u_sel < 0) {
      const char* ev = getenv("FMD_CSR128_U");
      u_sel = ev ? atoi(ev) : 0;
    }
    const int grid = min(fmd_div_up((long long)n_nodes * 32, 256), fmd_num_sms() * 8);
#define FMD_CSR128(WT, U)                                                                                          \
  cfconv_csr128_kernel<WT, U><<<grid, 256, 0, st>>>(x, (const WT*)filt, dist, (const int32_t*)gather,                \
                                                    (const int32_t*)seg_ptr, n_nodes, n_edges, rc, out)
    if (wdt == FMD_F32) {
      if (u_sel == 8) FMD_CSR128(float, 8); else FMD_CSR128(float, 4);
    } else {
      if (u_sel == 4) FMD_CSR128(__half, 4); else FMD_CSR128(__half, 8);
    }
#undef FMD_CSR128
    FMD_CHECK_LAUNCH();
    return FMD_OK;
  }
  if (wdt == FMD_F32) {
    if (idx_bytes == 4) launch_cfconv<float, int32_t>(x, filt, dist, gather, seg_ptr, perm, n_nodes, n_edges, n_feat, rc, out, st);
    else launch_cfconv<float, int64_t>(x, filt, dist, gather, seg_ptr, perm, n_nodes, n_edges, n_feat, rc, out, st);
  } else {
    if (idx_bytes == 4) launch_cfconv<__half, int32_t>(x, filt, dist, gather, seg_ptr, perm, n_nodes, n_edges, n_feat, rc, out, st);
    else launch_cfconv<__half, int64_t>(x, filt, dist, gather, seg_ptr, perm, n_nodes, n_edges, n_feat, rc, out, st);
  }
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

template <typename YT, typename WT, typename IdxT>
static void launch_grad_filter(const float* x, const float* g_out, const float* dist, const void* src, const void* dst,
                               int n_edges, const int32_t* n_edges_dev, int F, float rc, void* g_filt, const void* filt,
                               float* g_dcut, int acc, cudaStream_t st) {
  const int grid = min(fmd_div_up(n_edges, 256), fmd_num_sms() * 8);   // a warp owns 32 consecutive edges
  grad_filter_kernel<YT, WT, IdxT><<<grid, 256, 0, st>>>(x, g_out, dist, (const IdxT*)src, (const IdxT*)dst, n_edges,
                                                         n_edges_dev, F, rc, (YT*)g_filt, (const WT*)filt, g_dcut, acc);
}

extern "C" int fmd_cfconv_grad_filter(const float* x, const float* g_out, const float* dist, const void* edge_src,
                                      const void* edge_dst, int idx_bytes, int n_edges, const int32_t* n_edges_dev,
                                      int n_feat, float rc, void* g_filt, int ydt, const void* filt, int wdt,
                                      float* g_dcut, int accumulate_dcut, void* stream) {
  FMD_REQUIRE(x && g_out && dist && edge_src && edge_dst, "fmd_cfconv_grad_filter: bad arguments");
  FMD_REQUIRE(g_filt || (g_dcut && filt), "fmd_cfconv_grad_filter: nothing to compute");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_cfconv_grad_filter: idx_bytes must be 4 or 8");
  FMD_REQUIRE(n_feat > 0 && n_feat % 4 == 0, "fmd_cfconv_grad_filter: n_feat must be a positive multiple of 4");
  if (n_edges == 0) return FMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define FMD_GF(YT, WT)                                                                                              \
  do {                                                                                                              \
    if (idx_bytes == 4)                                                                                             \
      launch_grad_filter<YT, WT, int32_t>(x, g_out, dist, edge_src, edge_dst, n_edges, n_edges_dev, n_feat, rc,     \
                                          g_filt, filt, g_dcut, accumulate_dcut, st);                              \
    else                                                                                                            \
      launch_grad_filter<YT, WT, int64_t>(x, g_out, dist, edge_src, edge_dst, n_edges, n_edges_dev, n_feat, rc,     \
                                          g_filt, filt, g_dcut, accumulate_dcut, st);                              \
  } while (0)
  if (ydt == FMD_F32 && wdt == FMD_F32) FMD_GF(float, float);
  else if (ydt == FMD_F32 && wdt == FMD_F16) FMD_GF(float, __half);
  else if (ydt == FMD_F16 && wdt == FMD_F32) FMD_GF(__half, float);
  else if (ydt == FMD_F16 && wdt == FMD_F16) FMD_GF(__half, __half);
  else FMD_FAIL("fmd_cfconv_grad_filter: bad dtype");
#undef FMD_GF
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}
