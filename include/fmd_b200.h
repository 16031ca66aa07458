/*
 * fmd_b200.h — C ABI of libfmd_b200.so: hand-written sm_100a kernels for the CGSchNet
 * force-field + Langevin step (the hot path of UNITES-Lab/flash-molecular-dynamics).
 *
 * Conventions (all entry points):
 *   - plain pointers to DEVICE memory + sizes; no C++/torch types cross the boundary;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never allocates or
 *     frees device memory and keeps no mutable global state besides the last error string;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the stream or the device;
 *   - return 0 (FMD_OK) or a negative code; fmd_last_error() gives the text;
 *   - `idx_bytes` selects the integer width of index arrays: 8 (int64, the reference's dtype)
 *     or 4 (int32, used by the fused step);
 *   - `wdt` / `xdt` / `ydt` select element types: FMD_F32 or FMD_F16;
 *   - edge counts may live on the device: `n_edges_dev` (int32*, nullable) overrides the host
 *     value `n_edges` (then `n_edges` is the buffer capacity) so a whole MD step can be captured
 *     in a CUDA graph with no host round trip.
 *
 * "replaces:" cites the reference interface (paths under /root/reference/src/flashmd/).
 */
#ifndef FMD_B200_H
#define FMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMD_OK 0
#define FMD_ERR_INVALID (-1)
#define FMD_ERR_CUDA (-2)
#define FMD_ERR_UNSUPPORTED (-3)

#define FMD_F32 0
#define FMD_F16 1

/* epilogue / prologue flags of fmd_linear */
#define FMD_ACT_NONE 0
#define FMD_ACT_TANH 1          /* tanhf */
#define FMD_ACT_TANH_CLAMPED 2  /* (e^{2x}-1)/(e^{2x}+1), x clamped to +-10: kernels/cfconv_kernels.py:449-454 */

const char* fmd_last_error(void);
int fmd_version(void);
/* number of SMs of the current device (148 on B200) */
int fmd_sm_count(void);

/* ---------------------------------------------------------------- neighbour list + CSR ----- */

/* replaces: torch_cluster.radius_graph(x, r, batch, loop=False, max_num_neighbors,
 * flow="target_to_source") as called by neighbor_list/torch_impl.py:216-224, fused with
 * build_csr_index / build_src_csr_index (kernels/csr_kernels.py:88-169, 229-294).
 * mol_ptr [n_mols+1] int32: first node of each molecule (nodes of a molecule are contiguous).
 * Step 1: per-centre neighbour counts, deg [n_nodes] int32 (self excluded; at most
 * max_num_neighbors+1 hits incl. self are considered, in ascending neighbour index).
 * max_mol_size = largest mol_ptr[b+1]-mol_ptr[b] (host-known; sizes the grid). */
int fmd_nl_count(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size,
                 float rc, int max_num_neighbors, int32_t* deg, void* stream);

/* exclusive prefix sum: out[0..n] (n+1 values, out[n] = total). workspace: >= 4*(n/1024+2) bytes. */
int fmd_exclusive_scan_i32(const int32_t* in, int32_t* out, int n, void* workspace, void* stream);

/* Step 2: emit the edge list, centre-major ("src" ascending), neighbour ascending inside a
 * centre — exactly radius_graph's order — plus the edge length ||pos[dst]-pos[src]||.
 * seg_ptr [n_nodes+1] int32 = exclusive scan of deg (== src_ptr; for a symmetric list also dst_ptr).
 * Edges beyond `capacity` are dropped (seg_ptr[n_nodes] still holds the true count).
 * edge_src/edge_dst: idx_bytes wide. dist may be NULL. */
int fmd_nl_fill(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size,
                float rc, int max_num_neighbors, const int32_t* seg_ptr, int capacity, void* edge_src,
                void* edge_dst, int idx_bytes, float* dist, void* stream);

/* Step 3: rev[e] = index of the edge (dst_e -> src_e) (binary search in dst_e's sorted segment),
 * -1 if absent. For radius_graph output this IS csr_perm of build_csr_index(edge_dst) (stable
 * order) and dst_ptr == seg_ptr. rev: idx_bytes wide. */
int fmd_nl_reverse(const int32_t* seg_ptr, const void* edge_src, const void* edge_dst, int idx_bytes,
                   int n_nodes, int capacity, void* rev, void* stream);

/* replaces: build_csr_index(edge_dst, num_nodes) / build_src_csr_index(edge_src, num_nodes)
 * (kernels/csr_kernels.py:88, :229) for ARBITRARY key arrays. ptr [num_nodes+1], perm [n_edges],
 * both idx_bytes wide; perm is the STABLE counting-sort permutation (deterministic, unlike the
 * reference's atomic-cursor fill). workspace: >= 4*(3*num_nodes + n_edges + num_nodes/1024 + 16) bytes. */
int fmd_build_csr(const void* keys, int idx_bytes, int n_edges, int num_nodes, void* ptr, void* perm,
                  void* workspace, void* stream);

/* ---------------------------------------------------------------- edge features ------------- */

/* replaces: fused_distance_gaussian_rbf_cutoff (kernels/cfconv_kernels.py:1470-1677), forward.
 * dist [E] f32, rbf [E,R] f32 = exp(gamma (d-mu_k)^2) * 0.5(cos(pi d/rc)+1) [d<rc]. */
int fmd_dist_rbf_cutoff_fwd(const float* pos, const void* edge_src, const void* edge_dst, int idx_bytes,
                            int n_edges, const int32_t* n_edges_dev, const float* centers, int num_rbf,
                            float gamma, float rc, float* dist, float* rbf, void* stream);

/* replaces: FusedDistanceGaussianRBFCutoffFunction.backward (kernels/cfconv_kernels.py:1679-1735),
 * first half: g_d[e] (+)= grad_dist[e] + sum_k grad_rbf[e,k] * d rbf_k / d d.
 * grad_dist may be NULL. accumulate != 0 adds into g_d. */
int fmd_rbf_bwd(const float* dist, const float* grad_rbf, const float* grad_dist, int n_edges,
                const int32_t* n_edges_dev, const float* centers, int num_rbf, float gamma, float rc,
                float* g_d, int accumulate, void* stream);

/* replaces: FusedDistanceGaussianRBFCutoffFunction.backward (kernels/cfconv_kernels.py:1696-1733),
 * second half, generic edge lists (atomic adds, like the reference's index_add_):
 * grad_pos[dst] += g_d * u, grad_pos[src] -= g_d * u, u = (pos[dst]-pos[src]) / max(d, 1e-8).
 * grad_pos [n_nodes,3] must be zero-initialised by the caller. */
int fmd_edge_grad_to_pos_atomic(const float* pos, const void* edge_src, const void* edge_dst, int idx_bytes,
                                const float* dist, const float* g_d, int n_edges,
                                const int32_t* n_edges_dev, float* grad_pos, void* stream);

/* replaces: the same second half (kernels/cfconv_kernels.py:1696-1733), deterministic / atomic-free for the
 * sorted symmetric list of fmd_nl_fill:
 * out[i] = sign * sum_{e in seg(i)} (g_d[e] + g_d[rev[e]]) * u_e    (sign=+1 gives FORCES -dE/dx).
 * pair_mode != 0: `rev` is the edge -> pair map (pidx) of fmd_nl_step and g_d the per-PAIR gradient, already summed over
 * both directions (fmd_filter_cfconv_bwd): out[i] = sign * sum_e g_d[rev[e]] * u_e.
 * accumulate != 0 adds into out. int32 indices. */
int fmd_edge_grad_to_forces_csr(const float* pos, const int32_t* seg_ptr, const int32_t* edge_dst,
                                const int32_t* rev, const float* dist, const float* g_d, int n_nodes,
                                int n_edges, float sign, float* out, int accumulate, int pair_mode, void* stream);

/* ---------------------------------------------------------------- CFConv -------------------- */

/* replaces: fused_csr_cfconv (kernels/csr_kernels.py:625-810) and fused_src_csr_grad_x (:302-482).
 * out[i,:] = sum_{p in [seg_ptr[i], seg_ptr[i+1])} x[gather[e],:] * filt[e,:] * C(dist[e]),
 *            e = perm ? perm[p] : p.
 * forward of the reference: gather=edge_src, seg_ptr=dst_ptr, perm=csr_perm;
 * grad_x of the reference:  gather=edge_dst, seg_ptr=src_ptr, perm=src_perm, x=grad_out;
 * fused step (sorted symmetric list): gather=edge_dst, seg_ptr, perm=NULL for both directions.
 * x [n_rows_x,F] f32, filt [E,F] wdt, out [n_nodes,F] f32. F % 4 == 0. Warp per destination,
 * deterministic, atomic-free. seg_ptr/gather/perm are idx_bytes wide. Segment bounds are clamped to
 * n_edges (the number of rows of filt / capacity of the edge buffers). */
int fmd_cfconv_csr(const float* x, const void* filt, int wdt, const float* dist, const void* gather,
                   const void* seg_ptr, const void* perm, int idx_bytes, int n_nodes, int n_edges, int n_feat,
                   float rc, float* out, void* stream);

/* replaces: fused_grad_filter_out (kernels/cfconv_kernels.py:178-337), plus the exact
 * d(cutoff)/d(distance) term the reference's Triton backward drops (csr_kernels.py:912).
 * g_filt[e,:] = x[src_e,:] * g_out[dst_e,:] * C(d_e)          (ydt = FMD_F32 | FMD_F16)
 * g_dcut[e]  (+)= C'(d_e) * sum_f x[src_e,f] * g_out[dst_e,f] * filt[e,f]   (if g_dcut && filt) */
int fmd_cfconv_grad_filter(const float* x, const float* g_out, const float* dist, const void* edge_src,
                           const void* edge_dst, int idx_bytes, int n_edges, const int32_t* n_edges_dev,
                           int n_feat, float rc, void* g_filt, int ydt, const void* filt, int wdt,
                           float* g_dcut, int accumulate_dcut, void* stream);

/* replaces: the per-step call of torch_cluster.radius_graph (neighbor_list/torch_impl.py:216-224) + build_csr_index /
 * build_src_csr_index (kernels/csr_kernels.py:88, 229) of the fused step, in FOUR launches: count (degrees and, if pair
 * outputs are given, the number of neighbours with a larger index), one launch that scans both count arrays, fill (edge
 * list + the undirected pair list), reverse map (+ pidx).  Edge list, seg_ptr, dist, rev identical to fmd_nl_count /
 * fmd_exclusive_scan_i32 / fmd_nl_fill / fmd_nl_reverse (int32 indices).
 * Undirected pair list (no reference counterpart: the reference differentiates every directed edge): pairs = the edges with
 * dst > src in list order.  The gradient of the energy with respect to a distance is needed only as the SUM over the two
 * directions of a pair (the filter depends on the distance alone: W(e) == W(rev e)), so the fused backward kernel runs over
 * pairs - half the tiles.  pair_cnt [n_nodes], pair_ptr [n_nodes+1], pair_own / pair_nbr / pair_dist [pair_capacity],
 * pidx [capacity]: pair index of every directed edge (both directions of a pair map to it).
 * seg_ptr[n_nodes] / pair_ptr[n_nodes] hold the live edge / pair counts on the device.  pair_cnt == NULL: no pair list.
 * max_edges (nullable, int32[1]): sticky high-water mark, max_edges[0] = max(max_edges[0], edge count of this call): a host
 * that checks the capacity only now and then (the kernels clamp, they never write out of bounds) does not miss an overflow. */
int fmd_nl_step(const float* pos, const int32_t* mol_ptr, int n_mols, int n_nodes, int max_mol_size, float rc,
                int max_num_neighbors, int32_t* deg, int32_t* seg_ptr, int capacity, int32_t* edge_src,
                int32_t* edge_dst, float* dist, int32_t* rev, int32_t* pair_cnt, int32_t* pair_ptr, int pair_capacity,
                int32_t* pair_own, int32_t* pair_nbr, float* pair_dist, int32_t* pidx, int32_t* max_edges, void* stream);

/* ---------------------------------------------------------------- fused filter network (x) CFConv (tensor cores) */

/* replaces, for the W16A16 path, the chain  GPTQW16A16FilterNetwork.forward (models/gptq.py:92-130:
 * fused_linear_tanh_fp16 kernels/cfconv_kernels.py:644-760 + linear_fp16_to_fp16 :896-952) ->
 * fused_csr_cfconv (kernels/csr_kernels.py:625-810)  [and, with x = grad_out, fused_src_csr_grad_x :302-482]
 * by ONE tcgen05 kernel per call: per 128-edge tile rbf is recomputed from dist, both filter GEMMs run on
 * the tensor cores (fp16 operands, fp32 TMEM accumulators) and the messages are segment-reduced in the
 * epilogue; t, W and rbf never reach HBM.
 *   out[i,:] = sum_{e in [seg_ptr[i], seg_ptr[i+1])} (tanh(rbf_e Wf0^T + bf0) Wf1^T) * x[edge_nbr[e],:] * C(dist[e])
 * Edge list: the sorted symmetric list of fmd_nl_fill (edge_owner = its edge_src, edge_nbr = its edge_dst,
 * int32). wf0_h [128,64] fp16 = Wf0 [out,in] zero-padded from num_rbf to 64 columns; bf0_h [128] fp16 or
 * NULL; wf1_h [128,128] fp16 [out,in]. n_feat must be 128, num_rbf <= 63 (one padded column of the radial-basis
 * operand is a constant 1 that carries the bias through the first GEMM).
 * x_h [n_nodes,128] fp16: the gathered operand; its rows are staged through shared memory with cp.async (every epilogue
 * warp prefetches its own 64-byte feature slice one tile ahead). The filter value is rounded to fp16 like the
 * reference's [E,F] fp16 filter tensor and multiplied with one mixed-precision FMA per element (fp32 accumulate).
 * out [n_nodes,128] f32. part: scratch, >= ceil(capacity/128)*128 floats. Deterministic, atomic-free.
 * One persistent CTA per SM, producer / MMA-issuer / tanh / epilogue warps connected by mbarrier rings, D1 and D2
 * double-buffered in TMEM (512 columns). */
int fmd_filter_cfconv_fwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr,
                          const int32_t* seg_ptr, int n_nodes, int capacity, const int32_t* n_edges_dev,
                          const void* wf0_h, const void* bf0_h, const void* wf1_h, const float* centers, int num_rbf,
                          float gamma, float rc, const void* x_h, int n_feat, float* out, float* part, void* stream);

/* replaces, for the W16A16 path, the edge part of FusedCSRCFConvFunction.backward + the filter network's
 * backward + the fused-RBF backward: fused_grad_filter_out (kernels/cfconv_kernels.py:178-337),
 * LinearFP16ToFP16Function / FusedLinearTanhFP16Function backward GEMMs (:963-1226, :1329-1434) and
 * FusedDistanceGaussianRBFCutoffFunction.backward (:1679-1735, first half), in ONE tcgen05 kernel:
 *   g_d[e] (+)= d/d(dist_e) sum_f g_m[owner_e,f] * W_e,f(dist_e) * a[nbr_e,f] * C(dist_e)
 * through the radial basis (always) and through C (only when exact_cutoff_grad != 0; 0 reproduces the
 * reference's Triton backward, kernels/csr_kernels.py:912). t is recomputed on the tensor cores, g_W, g_t
 * are fp16 tensor-core operands (as in the reference) and never reach HBM. Same edge list / weight
 * layout as fmd_filter_cfconv_fwd. a_h, g_m_h [n_nodes,128] fp16: the rows a[edge_nbr[e],:] are copied by cp.async
 * straight into the swizzled K-major operand buffer of the g_W GEMM, one tile ahead, and multiplied in place by
 * g_m[edge_owner[e],:] with packed fp16 arithmetic. */
int fmd_filter_cfconv_bwd(const float* dist, const int32_t* edge_owner, const int32_t* edge_nbr, int capacity,
                          const int32_t* n_edges_dev, const void* wf0_h, const void* bf0_h, const void* wf1_h,
                          const float* centers, int num_rbf, float gamma, float rc, const void* a_h, const void* g_m_h,
                          int n_feat, float* g_d, int accumulate, int exact_cutoff_grad, void* stream);

/* tools only (scripts/trace_roles.py): when device_buffer != NULL, the next launches of the forward / backward fused
 * kernel run a traced instantiation in which CTA 0 records clock64() stamps {wait start, work start, end} per warp role
 * and tile into uint64 trace[9 roles][64 tiles][3]; NULL switches back to the production instantiation (which contains
 * no trace code). */
int fmd_debug_set_trace_fwd(void* device_buffer);
int fmd_debug_set_trace_bwd(void* device_buffer);

/* ---------------------------------------------------------------- dense layers -------------- */

/* replaces: fused_tanh_linear (kernels/cfconv_kernels.py:1758-1941), fused_linear_tanh_fp16 (:644-760),
 * linear_fp16 / linear_fp16_to_fp16 (:766-952), their backward GEMMs (:963-1226) and nn.Linear.
 *   Y[M,N] = epi( pro(X)[M,K] @ W[K,N] + bias[N] ),  then optionally  Y *= (1 - aux^2),  Y += res
 * X: xdt, W: wdt (row-major [K,N], i.e. nn.Linear.weight.t()), bias: wdt or NULL, Y: ydt,
 * aux [M,N] (auxdt) or NULL (tanh-backward fusion), res [M,N] f32 or NULL (residual add).
 * pro_act: FMD_ACT_* applied to X before the product; x_round_f16 != 0 rounds X to fp16 first
 * (the reference's in-kernel cast, cfconv_kernels.py:701). fp32 accumulation, true fp32 FMA
 * (no TF32). m_dev (int32*, nullable) overrides M (then M = capacity). */
int fmd_linear(const void* X, int xdt, const void* W, int wdt, const void* bias, void* Y, int ydt, int M,
               int N, int K, const int32_t* m_dev, int pro_act, int x_round_f16, int epi_act,
               const void* aux, int auxdt, const float* res, void* stream);

/* Same contract as fmd_linear for K, N in {64, 128}, computed on the tensor cores (tcgen05.mma
 * kind::tf32, fp32 accumulate in TMEM): operands are rounded to TF32 (round-to-nearest), which is what the
 * reference's GPU path does for these layers (nn.Linear under set_float32_matmul_precision("high"),
 * scripts/nvt_langevin.py:38; tl.dot default in fused_tanh_linear kernels/cfconv_kernels.py:1758-1843);
 * fp16 operands are exact. Used by the W16A16 step for the node-level layers; the fp32 parity path
 * keeps fmd_linear. w_is_nk != 0: W is given as [N,K] (the nn.Linear.weight layout) instead of [K,N],
 * which lets the kernel stage it with 16-byte loads. */
int fmd_linear_tc(const void* X, int xdt, const void* W, int wdt, const void* bias, void* Y, int ydt, int M, int N,
                  int K, const int32_t* m_dev, int pro_act, int x_round_f16, int epi_act, const void* aux, int auxdt,
                  const float* res, int w_is_nk, void* stream);

/* fp32-ACCURATE variant of fmd_linear on the tensor cores (fp32 emulation with three bf16 slices per operand,
 * x = s1 + s2 + s3 exactly, six slice products accumulated in two fp32 TMEM accumulators; csrc/fmd_linear_x3.cu explains
 * why 8-bit slices and not 3xTF32: the tensor core truncates when it accumulates, and only 16-bit products sum exactly).
 * replaces: the dense layers of the reference's fp32 path (--disable_optim: nn.Linear / MLP, models/mlp.py:41-57,
 * models/schnet.py:534-548,644,719, and their autograd) for the 1e-5 parity path: the edge-level filter-network layers
 * [E,R]x[R,F], [E,F]x[F,F] and their backward products.  fp32 in / fp32 out, even K <= 128, N <= 128, W row-major
 * [K,N]; epi_act / aux / res / m_dev as fmd_linear (exact tanhf for FMD_ACT_TANH).  Streaming pipeline, HBM-bound. */
int fmd_linear_x3(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K,
                  const int32_t* m_dev, int epi_act, const float* aux, const float* res, void* stream);

/* fmd_linear_x3 with the "rbf backward" epilogue: g_rbf = X[M,K] @ W[K,num_rbf] is never stored; row e is contracted
 * at once with d rbf_k / d d at d_e:
 *   g_d[e] (+)= sum_k g_rbf[e,k] * exp(gamma (d_e - mu_k)^2) * (2 gamma (d_e - mu_k) C(d_e) + C'(d_e))
 * replaces: the last backward GEMM of the filter network (autograd of models/mlp.py:41-57) followed by the backward of
 * FusedDistanceGaussianRBFCutoffFunction (kernels/cfconv_kernels.py:1679-1735, first half: grad_rbf -> grad_dist).
 * Saves the [E,R] write + read and one launch per interaction block on the fp32 parity path. */
int fmd_linear_x3_rbf_bwd(const float* X, const float* W, int M, int num_rbf, int K, const int32_t* m_dev,
                          const float* dist, const float* centers, float gamma, float rc, float* g_d, int accumulate,
                          void* stream);

/* One dense layer of fmd_linear_chain_tc. W is [N,K] (the nn.Linear.weight layout), K = N of the previous stage. */
#define FMD_MAX_CHAIN 4
typedef struct {
  const void* W;     /* [N,K], wdt */
  const void* bias;  /* [N] wdt, or NULL */
  int wdt;           /* FMD_F32 | FMD_F16 */
  int N;             /* 64 | 128 */
  int epi_act;       /* FMD_ACT_* applied to X W^T + bias */
  const void* aux;   /* [M,N] auxdt or NULL: result *= (1 - aux^2)  (tanh-backward fusion) */
  int auxdt;
  const float* res;  /* [M,N] f32 or NULL: result += res */
  void* Y;           /* [M,N] ydt or NULL: store this stage's result (the last stage must store) */
  int ydt;
  int round_f16;     /* round the result to fp16 before it feeds the next stage (W16A16 output network) */
} fmd_dense_stage;

/* replaces: runs of consecutive node-level layers of the reference (CFConv.lin2 -> tanh -> InteractionBlock.lin
 * -> residual -> next CFConv.lin1, models/schnet.py:534-548,644,719; the output MLP, models/gptq.py:266-306; and
 * the matching backward runs) by ONE launch: stage s+1 consumes the output tile of stage s from shared memory
 * (tf32 tensor-core GEMMs as fmd_linear_tc, same numerics). X [M,K] xdt; pro_act / x_round_f16 as fmd_linear. */
int fmd_linear_chain_tc(const void* X, int xdt, int M, int K, int pro_act, int x_round_f16,
                        const fmd_dense_stage* stages, int n_stages, void* stream);

/* ---------------------------------------------------------------- node-level helpers -------- */

/* replaces: torch.nn.Embedding (models/schnet.py:203). out[i,:] = table[types[i],:]. types: idx_bytes. */
int fmd_embedding(const float* table, const void* types, int idx_bytes, int n_nodes, int n_feat, float* out,
                  void* stream);

/* replaces: the last layer of the output MLP (Linear(hidden, 1, bias=False), models/schnet.py:829-834 /
 * models/gptq.py:304) and the first step of its backward, in one pass over y [n_nodes, n_hidden] (dt = FMD_F32 |
 * FMD_F16, also the type of w [n_hidden] and g_y):  e_atom[i] = sum_k y[i,k] w[k] ;  g_y[i,k] = w[k] (1 - y[i,k]^2)
 * (the gradient of sum(e_atom) w.r.t. the pre-activation of the last hidden layer; g_y nullable). */
int fmd_out_head(const void* y, const void* w, int dt, int n_nodes, int n_hidden, float* e_atom, void* g_y, void* stream);

/* replaces: scatter(energy, batch, reduce="sum") (models/schnet.py:355-357) for sorted `batch`:
 * out[b] (+)= sum_{i in [mol_ptr[b], mol_ptr[b+1])} e_atom[i]; deterministic block reduction. */
int fmd_segment_sum(const float* e_atom, const int32_t* mol_ptr, int n_mols, float* out, int accumulate,
                    void* stream);

/* ---------------------------------------------------------------- priors -------------------- */

#define FMD_PRIOR_BONDS 0      /* k (d - x0)^2 + V0            prior/harmonic.py:122-123 + internal_coordinates.py:73-101 */
#define FMD_PRIOR_ANGLES 1     /* k (cos(theta) - x0)^2 + V0   prior/harmonic.py:122-123 + internal_coordinates.py:140-170 */
#define FMD_PRIOR_DIHEDRALS 2  /* v0 + sum_n k1_n sin(n phi) + k2_n cos(n phi)   prior/fourier_series.py:154-192 */
#define FMD_PRIOR_REPULSION 3  /* (sigma / d)^6                 prior/repulsion.py:119-122 */
#define FMD_PRIOR_POLY_BONDS 4 /* V0 + sum_{n=1..4} k_n d^n     prior/polynomial.py:13-186 (fmd_priors_csr only) */
/* per-term form codes of the angle-like and improper-like tables of fmd_priors_csr */
#define FMD_ANGLE_HARMONIC_COS 0   /* k (cos - x0)^2                           HarmonicAngles / GeneralAngles */
#define FMD_ANGLE_POLY_COS 1       /* sum_{n=1..6} k_n cos^n                  QuarticAngles (prior/polynomial.py) */
#define FMD_ANGLE_RESTRICTED 2     /* a c^4 + b c^3 + c c^2 + d c + k/sin^2   prior/restricted_bending.py:13-238 */
#define FMD_ANGLE_HARMONIC_RAW 3   /* k (theta - x0)^2                        HarmonicAnglesRaw (prior/harmonic.py:267-300) */
#define FMD_IMPROPER_HARMONIC 0    /* k (phi - x0)^2                          HarmonicImpropers (prior/harmonic.py:230-265) */
#define FMD_IMPROPER_SHIFTED 1     /* k (x - x0)^2, x = (phi < 0 ? phi + 2 pi : phi) - pi   (prior/harmonic.py:327-405) */

/* replaces: prior.forward + its torch.autograd.grad (models/gradients.py:265) for one condensed
 * prior term class. mapping [order, n_terms] int32 (row-major, rows = roles), mapping_batch
 * [n_terms] int32, params: kind-specific flat f32 vectors
 *   BONDS/ANGLES: p0=k[n_terms] p1=x0[n_terms] p2=V0[n_terms]|NULL
 *   DIHEDRALS:    p0=k1[n_terms,n_degs] p1=k2[n_terms,n_degs] p2=v0[n_terms]
 *   REPULSION:    p0=sigma[n_terms]
 * energy [n_mols] and forces [n_nodes,3] are ACCUMULATED into (atomicAdd, like index_add_). */
int fmd_prior_energy_forces(int kind, const float* pos, const int32_t* mapping, const int32_t* mapping_batch,
                            int n_terms, const float* p0, const float* p1, const float* p2, int n_degs,
                            float* energy, float* forces, void* stream);

/* Same physics, owner-computes form used by the fused step: one warp per bead walks the bead's incident
 * terms (CSR built ONCE from the static prior topology) - no atomics, deterministic, one launch for all
 * prior classes. replaces: SumOut over the condensed priors + their autograd (models/gradients.py:72-152,
 * :265; simulation/specialize_prior.py:112-207).
 *   pair_ptr [n_nodes+1], pair_ent [n_pair_inc] of 16-byte records {other | kind<<28, p0, p1, p2}
 *     (kind FMD_PRIOR_BONDS: k, x0, V0; FMD_PRIOR_REPULSION: sigma, -, -), every pair listed under BOTH beads;
 *     with pair_tab != NULL the records are 8 bytes {other | kind<<28, id} and the parameters are the 8 floats
 *     pair_tab[8 id ..] (deduplicated table: halves the largest HBM stream outside the edge kernels):
 *     (p0, p1, p2, p3 | p4, -, -, -); FMD_PRIOR_POLY_BONDS (packed records only): k1..k4 | V0;
 *   mb_ptr [n_nodes+1], mb_ent [n_mb_inc] = term | role<<28 | table<<30 (0 angle-like, 1 Fourier dihedral,
 *     2 improper-like), every term under each of its beads;
 *   ang_map [3,n_ang] + ang_par [n_ang,8] = {p0..p5, V0, form as int bits}: HARMONIC_COS (k, x0), POLY_COS (k1..k6),
 *     RESTRICTED (a, b, c, d, k), HARMONIC_RAW (k, x0) - several angle prior classes share the one table;
 *   dih_map [4,n_dih] + k1, k2 [n_dih,n_degs], v0 [n_dih] | NULL;
 *   imp_map [4,n_imp] + imp_par [n_imp,4] = {k, x0, V0, form}.
 * Outputs: e_atom [n_nodes] = the bead's share of its terms' energies (sum per molecule with
 * fmd_segment_sum), forces [n_nodes,3] written (accumulate_forces == 0) or added to. Any group may be NULL. */
int fmd_priors_csr(const float* pos, int n_nodes, const int32_t* pair_ptr, const void* pair_ent, const float* pair_tab,
                   const int32_t* mb_ptr, const int32_t* mb_ent, const int32_t* ang_map, int n_ang, const float* ang_par,
                   const int32_t* dih_map, int n_dih, const float* dih_k1, const float* dih_k2, const float* dih_v0,
                   int n_degs, const int32_t* imp_map, int n_imp, const float* imp_par, float* e_atom, float* forces,
                   int accumulate_forces, void* stream);

/* ---------------------------------------------------------------- integrator ---------------- */

/* replaces: LangevinSimulation.timestep B-A-O-A (simulation/langevin.py:137-157).
 * In place on pos/vel [n_nodes,3]. noise: external N(0,1) [n_nodes,3] or NULL -> counter-based
 * Philox4x32-10 keyed by (seed), counter (step [+ *step_dev when non-NULL], node_offset + node) + Box-Muller:
 * node_offset = global index of this shard's first bead, so that a replica batch sharded over several GPUs draws
 * exactly the noise of the unsharded run (and no two shards share a stream).
 * inv_mass = 1/m [n_nodes]; noise_std = sqrt(1/(beta m)) [n_nodes] (beta_mass_ratio, :211-215). */
int fmd_baoab_pre(float* pos, float* vel, const float* forces, const float* inv_mass, const float* noise_std,
                  const float* noise, uint64_t seed, uint64_t step, const uint64_t* step_dev, uint64_t node_offset,
                  int n_nodes, float dt, float vscale, float noisescale, void* stream);

/* replaces: OverdampedSimulation.timestep (simulation/langevin.py:361-414): x += F dtau + sqrt(2 dtau) xi in place,
 * dtau [n_nodes] = D dt with the reference's D = 1 / (beta friction) per bead. Noise as in fmd_baoab_pre. */
int fmd_overdamped_step(float* pos, const float* forces, const float* dtau, const float* noise, uint64_t seed,
                        uint64_t step, const uint64_t* step_dev, uint64_t node_offset, int n_nodes, void* stream);

/* *counter += 1 on the stream (keeps the Philox step counter on the device so a captured CUDA
 * graph of the whole step can be replayed; fmd_baoab_pre adds *step_dev to `step`). */
int fmd_increment_u64(uint64_t* counter, void* stream);

/* replaces: final B half-kick (simulation/langevin.py:169) (+ optional kinetic energy per molecule:
 * ke[b] = 0.5 sum m v^2, langevin.py:266-270, when ke != NULL; mol_ptr then required).
 * step_counter (nullable): *step_counter += 1, the device-side Philox step counter of a graph-replayed step. */
int fmd_baoab_post(float* vel, const float* forces, const float* inv_mass, int n_nodes, float dt,
                   const int32_t* mol_ptr, int n_mols, float* ke, uint64_t* step_counter, void* stream);

/* standard-normal stream used by fmd_baoab_pre when noise == NULL, exposed for tests. out [n_nodes,3]. */
int fmd_philox_normal(uint64_t seed, uint64_t step, int n_nodes, float* out, void* stream);

/* ---------------------------------------------------------------- replica exchange ---------- */

/* replaces: PTSimulation._detect_exchange (simulation/parallel_tempering.py:368-413).
 * accept[p] = u_p < exp((E[a_p]-E[b_p]) (beta[a_p]-beta[b_p])), u_p from `uniforms` or, when NULL,
 * from Philox keyed by (seed, exchange_index, p) so every rank derives the same decisions. */
int fmd_pt_decide(const float* energy, const float* beta, const int32_t* pair_a, const int32_t* pair_b,
                  int n_pairs, const float* uniforms, uint64_t seed, uint64_t exchange_index, int32_t* accept,
                  void* stream);

/* replaces: PTSimulation._perform_exchange (:415-481) for pairs resident on this device:
 * swap positions, swap velocities scaled by sqrt(beta_old/beta_new). n_atoms beads per sim. */
int fmd_pt_swap(float* pos, float* vel, const float* beta, const int32_t* pair_a, const int32_t* pair_b,
                const int32_t* accept, int n_pairs, int n_atoms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FMD_B200_H */
