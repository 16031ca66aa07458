// Per-step radius graph emitting sorted CSR, exclusive scan, reverse-edge map, generic CSR build.
//
// Layout: nodes of one molecule are contiguous; edges never cross molecules
// (reference neighbor_list/neighbor_list.py:50-52).  One CTA = one molecule x 8 centres, one
// warp per centre; the molecule's coordinates are staged in shared memory as SoA tiles and every
// lane tests one candidate per iteration, so hits are compacted with ballot/popc and written in
// ascending neighbour order (the order torch_cluster's radius kernel produces) without atomics.
#include "fmd_common.cuh"

namespace {

constexpr int NL_WARPS = 8;
constexpr int NL_TILE = 1024;

// Pair outputs (optional, int32 lists only): the undirected pairs are the edges with dst > src in list order.  The count pass
// also writes deg_hi[c] = number of such edges of centre c; the fill pass writes pair p = pair_ptr[c] + (rank among them).
struct NlPairs {
  int32_t* deg_hi;          // count pass
  const int32_t* pair_ptr;  // fill pass
  int32_t* own;
  int32_t* nbr;
  float* dist;
  int capacity;
};

template <bool FILL, typename IdxT>
__global__ void __launch_bounds__(NL_WARPS * 32)
nl_kernel(const float* __restrict__ pos, const int32_t* __restrict__ mol_ptr, float rc2, int max_hits,
          int32_t* __restrict__ deg, const int32_t* __restrict__ seg_ptr, int capacity,
          IdxT* __restrict__ edge_src, IdxT* __restrict__ edge_dst, float* __restrict__ dist, const NlPairs pr) {
  __shared__ float sx[NL_TILE], sy[NL_TILE], sz[NL_TILE];
  const int b = blockIdx.x;
  const int lo = mol_ptr[b], hi = mol_ptr[b + 1];
  const int n = hi - lo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_local = blockIdx.y * NL_WARPS + warp;
  if (blockIdx.y * NL_WARPS >= n) return;  // whole CTA idle (uniform)
  const bool active = c_local < n;
  const int c = lo + c_local;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (active) {
    cx = pos[3 * c + 0];
    cy = pos[3 * c + 1];
    cz = pos[3 * c + 2];
  }
  int hits = 0;     // hits incl. self so far (for the max_num_neighbors cap)
  int emitted = 0;  // edges emitted so far
  int emitted_hi = 0;  // ... of them with neighbour index > centre index (the undirected pairs listed under this centre)
  int out_base = 0, pair_base = 0;
  if (FILL && active) out_base = seg_ptr[c];
  if (FILL && active && pr.pair_ptr) pair_base = pr.pair_ptr[c];
  const bool want_pairs = FILL ? pr.pair_ptr != nullptr : pr.deg_hi != nullptr;
  for (int t0 = 0; t0 < n; t0 += NL_TILE) {
    const int tn = min(NL_TILE, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += blockDim.x) {
      const float* p = pos + 3 * (size_t)(lo + t0 + i);
      sx[i] = p[0];
      sy[i] = p[1];
      sz[i] = p[2];
    }
    __syncthreads();
    if (!active || hits >= max_hits) continue;
    for (int j0 = 0; j0 < tn; j0 += 32) {
      const int jl = j0 + lane;
      bool hit = false;
      float d2 = 0.f;
      if (jl < tn) {
        // accumulation order of torch_cluster's radius kernel: dist += (x_n[d]-x_c[d])^2, d = x,y,z
        const float dx = sx[jl] - cx, dy = sy[jl] - cy, dz = sz[jl] - cz;
        d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        hit = d2 < rc2;
      }
      const unsigned hmask = __ballot_sync(0xffffffffu, hit);
      const unsigned lt = (1u << lane) - 1u;
      const int rank = hits + __popc(hmask & lt);
      const bool emit = hit && (rank < max_hits) && (t0 + jl != c_local);
      const unsigned emask = __ballot_sync(0xffffffffu, emit);
      if (FILL && emit) {
        const int o = out_base + emitted + __popc(emask & lt);
        if (o < capacity) {
          edge_src[o] = (IdxT)c;
          edge_dst[o] = (IdxT)(lo + t0 + jl);
          if (dist) dist[o] = sqrtf(d2);
        }
      }
      if (want_pairs) {
        const bool hi_edge = emit && (t0 + jl > c_local);
        const unsigned pmask = __ballot_sync(0xffffffffu, hi_edge);
        if (FILL && hi_edge) {
          const int p = pair_base + emitted_hi + __popc(pmask & lt);
          if (p < pr.capacity) {
            pr.own[p] = c;
            pr.nbr[p] = lo + t0 + jl;
            pr.dist[p] = sqrtf(d2);
          }
        }
        emitted_hi += __popc(pmask);
      }
      hits += __popc(hmask);
      emitted += __popc(emask);
      if (hits >= max_hits) break;  // warp-uniform
    }
  }
  if (!FILL && active && lane == 0) {
    deg[c] = emitted;
    if (want_pairs) pr.deg_hi[c] = emitted_hi;
  }
}

// two exclusive scans in one launch (CTA b scans array b: degrees -> seg_ptr, pair counts -> pair_ptr), 4096 items per pass;
// out[n] = total.  n is a few 10^4: one CTA per array is faster than the three-kernel scan (three launches per array).
__global__ void __launch_bounds__(1024)
scan2_kernel(const int32_t* __restrict__ in0, int32_t* __restrict__ out0, const int32_t* __restrict__ in1,
             int32_t* __restrict__ out1, int n, int32_t* __restrict__ max_total0) {
  const int32_t* in = blockIdx.x == 0 ? in0 : in1;
  int32_t* out = blockIdx.x == 0 ? out0 : out1;
  if (!in || !out) return;
  __shared__ int ws[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    const int i = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i + k < n) ? in[i + k] : 0;
    const int tsum = v[0] + v[1] + v[2] + v[3];
    int x = tsum;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = ws[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      ws[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + (w > 0 ? ws[w - 1] : 0) + x - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i + k < n) out[i + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = carry_s;
    // sticky high-water mark of the edge count (capacity overflows between two host checks are not missed)
    if (blockIdx.x == 0 && max_total0 && carry_s > *max_total0) *max_total0 = carry_s;
  }
}

// ---------------------------------------------------------------- exclusive scan (3 phases)
constexpr int SCAN_ITEMS = 1024;

__global__ void __launch_bounds__(256) scan_block_sums(const int32_t* __restrict__ in, int n, int32_t* __restrict__ sums) {
  __shared__ int ws[8];
  const int base = blockIdx.x * SCAN_ITEMS;
  int s = 0;
  for (int i = threadIdx.x; i < SCAN_ITEMS; i += 256) {
    const int g = base + i;
    if (g < n) s += in[g];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    sums[blockIdx.x] = t;
  }
}

// single CTA: in-place exclusive scan of the block sums; writes the grand total to out_total
__global__ void __launch_bounds__(1024) scan_of_sums(int32_t* __restrict__ sums, int nb, int32_t* __restrict__ out_total) {
  __shared__ int ws[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int x = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = ws[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      ws[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const int warp_off = w > 0 ? ws[w - 1] : 0;
    const int carry = carry_s;
    if (i < nb) sums[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_total = carry_s;
}

__global__ void __launch_bounds__(256) scan_apply(const int32_t* __restrict__ in, int n, const int32_t* __restrict__ sums,
                                                  int32_t* __restrict__ out) {
  // each thread owns 4 consecutive items
  __shared__ int ws[8];
  const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? in[base + k] : 0;
  const int tsum = v[0] + v[1] + v[2] + v[3];
  int x = tsum;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) ws[w] = x;
  __syncthreads();
  int woff = 0;
  for (int k = 0; k < w; ++k) woff += ws[k];
  int run = sums[blockIdx.x] + woff + x - tsum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

// ---------------------------------------------------------------- reverse-edge map
template <typename IdxT>
__global__ void __launch_bounds__(256)
nl_reverse_kernel(const int32_t* __restrict__ seg_ptr, const IdxT* __restrict__ src, const IdxT* __restrict__ dst,
                  int n_nodes, int capacity, IdxT* __restrict__ rev, const int32_t* __restrict__ pair_ptr,
                  int32_t* __restrict__ pidx) {
  const int E = min(seg_ptr[n_nodes], capacity);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int s = (int)src[e], t = (int)dst[e];
    if (pidx && t > s) {
      // fused step (pidx given: the list is symmetric, the engine refuses lists that max_num_neighbors can truncate): this
      // edge is a pair's first direction, its pair index needs no search; rev of BOTH directions is written by the thread
      // of the mirror edge below, so only half of the threads walk a binary search
      const int first_hi = min(seg_ptr[s + 1], E) - (pair_ptr[s + 1] - pair_ptr[s]);
      pidx[e] = pair_ptr[s] + (e - first_hi);
      continue;
    }
    int lo = seg_ptr[t], hi = min(seg_ptr[t + 1], E);
    int found = -1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = (int)dst[mid];
      if (v < s) lo = mid + 1;
      else if (v > s) hi = mid;
      else { found = mid; break; }
    }
    rev[e] = (IdxT)found;
    if (pidx) {
      // t < s: the pair is listed under the smaller bead t, as edge `found` of t's segment (the pairs of t are the tail of
      // its segment)
      int p = -1;
      if (found >= 0) {
        rev[found] = (IdxT)e;
        const int first_hi = min(seg_ptr[t + 1], E) - (pair_ptr[t + 1] - pair_ptr[t]);
        p = pair_ptr[t] + (found - first_hi);
      }
      pidx[e] = p;
    }
  }
}

// ---------------------------------------------------------------- generic CSR build
template <typename IdxT>
__global__ void __launch_bounds__(256) csr_hist(const IdxT* __restrict__ keys, int n, int num_nodes, int32_t* __restrict__ counts) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const long long k = (long long)keys[e];
    if (k >= 0 && k < num_nodes) atomicAdd(&counts[k], 1);
  }
}
template <typename IdxT>
__global__ void __launch_bounds__(256) csr_fill(const IdxT* __restrict__ keys, int n, int num_nodes, const int32_t* __restrict__ ptr32,
                                                int32_t* __restrict__ cursor, int32_t* __restrict__ tmp) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const long long k = (long long)keys[e];
    if (k >= 0 && k < num_nodes) {
      const int slot = atomicAdd(&cursor[k], 1);
      tmp[ptr32[k] + slot] = e;
    }
  }
}
// warp per segment: rank sort (edge ids inside a segment are distinct) -> stable order
template <typename IdxT>
__global__ void __launch_bounds__(256) csr_rank_sort(const int32_t* __restrict__ ptr32, const int32_t* __restrict__ tmp, int num_nodes,
                                                     IdxT* __restrict__ perm, IdxT* __restrict__ ptr_out) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < num_nodes; i += nw) {
    const int a = ptr32[i], b = ptr32[i + 1];
    if (lane == 0) ptr_out[i] = (IdxT)a;
    if (i == num_nodes - 1 && lane == 0) ptr_out[num_nodes] = (IdxT)b;
    for (int p = a + lane; p < b; p += 32) {
      const int v = tmp[p];
      int r = 0;
      for (int q = a; q < b; ++q) r += (tmp[q] < v);
      perm[a + r] = (IdxT)v;
    }
  }
}
template <typename IdxT>
__global__ void write_last_ptr(const int32_t* __restrict__ ptr32, int num_nodes, IdxT* __restrict__ ptr_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ptr_out[num_nodes] = (IdxT)ptr32[num_nodes];
}

}  // namespace

extern "C" int fmd_exclusive_scan_i32(const int32_t* in, int32_t* out, int n, void* workspace, void* stream) {
  FMD_REQUIRE(n >= 0 && in && out && workspace, "fmd_exclusive_scan_i32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* sums = (int32_t*)workspace;
  const int nb = n > 0 ? fmd_div_up(n, SCAN_ITEMS) : 0;
  if (nb > 0) scan_block_sums<<<nb, 256, 0, st>>>(in, n, sums);
  scan_of_sums<<<1, 1024, 0, st>>>(sums, nb, out + n);
  if (nb > 0) scan_apply<<<nb, 256, 0, st>>>(in, n, sums, out);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_nl_count(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size,
                            float rc, int max_num_neighbors, int32_t* deg, void* stream) {
  FMD_REQUIRE(pos && mol_ptr && deg && n_mols >= 0 && max_mol_size >= 0, "fmd_nl_count: bad arguments");
  if (n_mols == 0 || n_nodes == 0) return FMD_OK;
  dim3 grid(n_mols, fmd_div_up(max_mol_size, NL_WARPS));
  FMD_REQUIRE(grid.y <= 65535, "fmd_nl_count: molecule too large");
  nl_kernel<false, int32_t><<<grid, NL_WARPS * 32, 0, (cudaStream_t)stream>>>(
      pos, mol_ptr, rc * rc, max_num_neighbors + 1, deg, nullptr, 0, nullptr, nullptr, nullptr, NlPairs{});
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_nl_fill(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size,
                           float rc, int max_num_neighbors, const int32_t* seg_ptr, int capacity, void* edge_src,
                           void* edge_dst, int idx_bytes, float* dist, void* stream) {
  FMD_REQUIRE(pos && mol_ptr && seg_ptr && edge_src && edge_dst, "fmd_nl_fill: bad arguments");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_nl_fill: idx_bytes must be 4 or 8");
  if (n_mols == 0 || n_nodes == 0) return FMD_OK;
  dim3 grid(n_mols, fmd_div_up(max_mol_size, NL_WARPS));
  FMD_REQUIRE(grid.y <= 65535, "fmd_nl_fill: molecule too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 4)
    nl_kernel<true, int32_t><<<grid, NL_WARPS * 32, 0, st>>>(pos, mol_ptr, rc * rc, max_num_neighbors + 1, nullptr,
                                                             seg_ptr, capacity, (int32_t*)edge_src,
                                                             (int32_t*)edge_dst, dist, NlPairs{});
  else
    nl_kernel<true, int64_t><<<grid, NL_WARPS * 32, 0, st>>>(pos, mol_ptr, rc * rc, max_num_neighbors + 1, nullptr,
                                                             seg_ptr, capacity, (int64_t*)edge_src,
                                                             (int64_t*)edge_dst, dist, NlPairs{});
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_nl_reverse(const int32_t* seg_ptr, const void* edge_src, const void* edge_dst, int idx_bytes,
                              int n_nodes, int capacity, void* rev, void* stream) {
  FMD_REQUIRE(seg_ptr && edge_src && edge_dst && rev, "fmd_nl_reverse: bad arguments");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_nl_reverse: idx_bytes must be 4 or 8");
  if (capacity == 0) return FMD_OK;
  const int grid = min(fmd_div_up(capacity, 256), fmd_num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bytes == 4)
    nl_reverse_kernel<int32_t><<<grid, 256, 0, st>>>(seg_ptr, (const int32_t*)edge_src, (const int32_t*)edge_dst,
                                                     n_nodes, capacity, (int32_t*)rev, nullptr, nullptr);
  else
    nl_reverse_kernel<int64_t><<<grid, 256, 0, st>>>(seg_ptr, (const int64_t*)edge_src, (const int64_t*)edge_dst,
                                                     n_nodes, capacity, (int64_t*)rev, nullptr, nullptr);
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_nl_step(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size, float rc,
                           int max_num_neighbors, int32_t* deg, int32_t* seg_ptr, int capacity, int32_t* edge_src,
                           int32_t* edge_dst, float* dist, int32_t* rev, int32_t* pair_cnt, int32_t* pair_ptr,
                           int pair_capacity, int32_t* pair_own, int32_t* pair_nbr, float* pair_dist, int32_t* pidx,
                           int32_t* max_edges, void* stream) {
  FMD_REQUIRE(pos && mol_ptr && deg && seg_ptr && edge_src && edge_dst && dist && rev, "fmd_nl_step: null argument");
  const bool pairs = pair_cnt != nullptr;
  FMD_REQUIRE(!pairs || (pair_ptr && pair_own && pair_nbr && pair_dist && pidx), "fmd_nl_step: incomplete pair outputs");
  if (n_mols == 0 || n_nodes == 0) return FMD_OK;
  dim3 grid(n_mols, fmd_div_up(max_mol_size, NL_WARPS));
  FMD_REQUIRE(grid.y <= 65535, "fmd_nl_step: molecule too large");
  cudaStream_t st = (cudaStream_t)stream;
  const float rc2 = rc * rc;
  NlPairs pc{};
  pc.deg_hi = pairs ? pair_cnt : nullptr;
  nl_kernel<false, int32_t><<<grid, NL_WARPS * 32, 0, st>>>(pos, mol_ptr, rc2, max_num_neighbors + 1, deg, nullptr, 0,
                                                            nullptr, nullptr, nullptr, pc);
  scan2_kernel<<<pairs ? 2 : 1, 1024, 0, st>>>(deg, seg_ptr, pair_cnt, pair_ptr, n_nodes, max_edges);
  NlPairs pf{};
  if (pairs) {
    pf.pair_ptr = pair_ptr;
    pf.own = pair_own;
    pf.nbr = pair_nbr;
    pf.dist = pair_dist;
    pf.capacity = pair_capacity;
  }
  nl_kernel<true, int32_t><<<grid, NL_WARPS * 32, 0, st>>>(pos, mol_ptr, rc2, max_num_neighbors + 1, nullptr, seg_ptr,
                                                           capacity, edge_src, edge_dst, dist, pf);
  if (capacity > 0) {
    const int g = min(fmd_div_up(capacity, 256), fmd_num_sms() * 8);
    nl_reverse_kernel<int32_t><<<g, 256, 0, st>>>(seg_ptr, edge_src, edge_dst, n_nodes, capacity, rev,
                                                  pairs ? pair_ptr : nullptr, pairs ? pidx : nullptr);
  }
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

template <typename IdxT>
static int build_csr_impl(const IdxT* keys, int n_edges, int num_nodes, IdxT* ptr, IdxT* perm, void* workspace,
                          cudaStream_t st) {
  int32_t* counts = (int32_t*)workspace;            // [num_nodes]
  int32_t* cursor = counts + num_nodes;             // [num_nodes]
  int32_t* ptr32 = cursor + num_nodes;              // [num_nodes+1]
  int32_t* tmp = ptr32 + num_nodes + 1;             // [n_edges]
  int32_t* scan_ws = tmp + n_edges;                 // [num_nodes/1024 + 2]
  FMD_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * (size_t)num_nodes, st));
  const int grid_e = max(1, min(fmd_div_up(n_edges, 256), fmd_num_sms() * 8));
  if (n_edges > 0) csr_hist<IdxT><<<grid_e, 256, 0, st>>>(keys, n_edges, num_nodes, counts);
  int rc = fmd_exclusive_scan_i32(counts, ptr32, num_nodes, scan_ws, st);
  if (rc != FMD_OK) return rc;
  if (n_edges > 0) csr_fill<IdxT><<<grid_e, 256, 0, st>>>(keys, n_edges, num_nodes, ptr32, cursor, tmp);
  if (num_nodes > 0) {
    const int grid_n = max(1, min(fmd_div_up((long long)num_nodes * 32, 256), fmd_num_sms() * 16));
    csr_rank_sort<IdxT><<<grid_n, 256, 0, st>>>(ptr32, tmp, num_nodes, perm, ptr);
  } else {
    write_last_ptr<IdxT><<<1, 32, 0, st>>>(ptr32, num_nodes, ptr);
  }
  FMD_CHECK_LAUNCH();
  return FMD_OK;
}

extern "C" int fmd_build_csr(const void* keys, int idx_bytes, int n_edges, int num_nodes, void* ptr, void* perm,
                             void* workspace, void* stream) {
  FMD_REQUIRE(ptr && workspace && n_edges >= 0 && num_nodes >= 0, "fmd_build_csr: bad arguments");
  FMD_REQUIRE(n_edges == 0 || (keys && perm), "fmd_build_csr: null keys/perm");
  FMD_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "fmd_build_csr: idx_bytes must be 4 or 8");
  if (idx_bytes == 4)
    return build_csr_impl<int32_t>((const int32_t*)keys, n_edges, num_nodes, (int32_t*)ptr, (int32_t*)perm, workspace,
                                   (cudaStream_t)stream);
  return build_csr_impl<int64_t>((const int64_t*)keys, n_edges, num_nodes, (int64_t*)ptr, (int64_t*)perm, workspace,
                                 (cudaStream_t)stream);
}
