"""Operator-level drop-ins for `flashmd.kernels` (reference: kernels/__init__.py:47-75).

Same names, argument order, dtypes, allocation and error behaviour as the reference's Triton
wrappers (callee allocates and returns the outputs, current stream, contiguity asserted, empty
edge lists return zeros), but every function launches hand-written sm_100a kernels through the
C ABI of libfmd_b200.so.  No Triton, no CPU fallback.

Weights are frozen in simulations (reference simulation/base.py:357-358), so the autograd
Functions return gradients for activations only; weight gradients raise NotImplementedError.
"""
from __future__ import annotations

import torch

from .. import _lib as L

# ------------------------------------------------------------------------------------------ helpers


def _st():
    return L.stream_ptr()


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("flashmd.kernels: CUDA tensors required (no CPU fallback is provided)")


def _linear(x, w_kn, bias=None, out_dtype=torch.float32, pro_act=L.ACT_NONE, x_round_f16=False,
            epi_act=L.ACT_NONE, aux=None, res=None, m_dev=None, out=None):
    """Y = epi(pro(X) @ W[K,N] + b) [* (1-aux^2)] [+ res]  — thin wrapper over fmd_linear."""
    assert x.is_contiguous() and w_kn.is_contiguous()
    M, K = x.shape
    K2, N = w_kn.shape
    assert K == K2, f"shape mismatch {x.shape} @ {w_kn.shape}"
    if bias is not None:
        assert bias.dtype == w_kn.dtype and bias.numel() == N
    y = out if out is not None else torch.empty((M, N), dtype=out_dtype, device=x.device)
    L.call("fmd_linear", L.ptr(x), L.dt_code(x), L.ptr(w_kn), L.dt_code(w_kn), L.ptr(bias), L.ptr(y), L.dt_code(y),
           M, N, K, L.ptr(m_dev), pro_act, int(bool(x_round_f16)), epi_act, L.ptr(aux),
           L.dt_code(aux) if aux is not None else 0, L.ptr(res), _st())
    return y


# ------------------------------------------------------------------------------------------ CSR


def _build_csr(keys: torch.Tensor, num_nodes: int):
    _check_cuda(keys)
    assert keys.is_contiguous() and keys.dim() == 1
    E = keys.numel()
    ptr = torch.empty(num_nodes + 1, dtype=keys.dtype, device=keys.device)
    perm = torch.empty(E, dtype=keys.dtype, device=keys.device)
    ws = torch.empty(4 * (3 * num_nodes + E + num_nodes // 1024 + 16), dtype=torch.uint8, device=keys.device)
    L.call("fmd_build_csr", L.ptr(keys), L.idx_bytes(keys), E, num_nodes, L.ptr(ptr), L.ptr(perm) if E else None,
           L.ptr(ws), _st())
    return ptr, perm


def build_csr_index(edge_dst: torch.Tensor, num_nodes: int, edge_src: torch.Tensor = None):
    """(dst_ptr [N+1], csr_perm [E]) — reference kernels/csr_kernels.py:88-169.  `csr_perm` is the
    stable counting-sort order (the reference's is the same set per segment, order run-dependent)."""
    return _build_csr(edge_dst, num_nodes)


def build_src_csr_index(edge_src: torch.Tensor, num_nodes: int):
    """(src_ptr [N+1], src_perm [E]) — reference kernels/csr_kernels.py:229-294."""
    return _build_csr(edge_src, num_nodes)


# ------------------------------------------------------------------------------------------ CFConv


def _cfconv(x, filter_out, edge_weight, gather, seg_ptr, perm, num_nodes, cutoff_upper):
    _check_cuda(x, filter_out, edge_weight, gather, seg_ptr)
    assert x.is_contiguous() and filter_out.is_contiguous() and edge_weight.is_contiguous()
    assert gather.is_contiguous() and seg_ptr.is_contiguous()
    assert x.dtype == torch.float32
    F = x.shape[1]
    out = torch.zeros((num_nodes, F), dtype=torch.float32, device=x.device)
    if filter_out.shape[0] == 0:
        return out
    L.call("fmd_cfconv_csr", L.ptr(x), L.ptr(filter_out), L.dt_code(filter_out), L.ptr(edge_weight), L.ptr(gather),
           L.ptr(seg_ptr), L.ptr(perm), L.idx_bytes(gather), num_nodes, filter_out.shape[0], F, float(cutoff_upper),
           L.ptr(out), _st())
    return out


def fused_csr_cfconv(x, filter_out, edge_weight, edge_src, dst_ptr, csr_perm, num_nodes, cutoff_upper):
    """out[i] = sum_{e -> i} x[src_e] * filter_out[e] * C(d_e) — reference kernels/csr_kernels.py:727-810."""
    return _cfconv(x, filter_out, edge_weight, edge_src, dst_ptr, csr_perm, num_nodes, cutoff_upper)


def fused_src_csr_grad_x(grad_output, filter_out, edge_weight, edge_dst, src_ptr, src_perm, num_nodes, cutoff_upper):
    """grad_x[s] = sum_{e from s} grad_output[dst_e] * filter_out[e] * C(d_e) — reference csr_kernels.py:399-482."""
    return _cfconv(grad_output, filter_out, edge_weight, edge_dst, src_ptr, src_perm, num_nodes, cutoff_upper)


def fused_cutoff_gather_multiply_scatter(x, filter_out, edge_weight, edge_src, edge_dst, num_nodes, cutoff_upper):
    """Same result as the reference's atomic kernel (kernels/cfconv_kernels.py:93-175) but
    deterministic: a CSR is built on the fly and the segment-reduce kernel is used."""
    dst_ptr, perm = _build_csr(edge_dst.contiguous(), num_nodes)
    return _cfconv(x, filter_out, edge_weight, edge_src, dst_ptr, perm, num_nodes, cutoff_upper)


def fused_grad_filter_out(x, grad_output, edge_weight, edge_src, edge_dst, cutoff_upper, out_dtype=None):
    """grad_filter_out[e] = x[src_e] * grad_output[dst_e] * C(d_e) — reference cfconv_kernels.py:261-337."""
    _check_cuda(x, grad_output, edge_weight, edge_src, edge_dst)
    E, F = edge_src.numel(), x.shape[1]
    out_dtype = out_dtype or torch.float32
    out = torch.empty((E, F), dtype=out_dtype, device=x.device)
    if E == 0:
        return out
    L.call("fmd_cfconv_grad_filter", L.ptr(x.contiguous()), L.ptr(grad_output.contiguous()), L.ptr(edge_weight),
           L.ptr(edge_src), L.ptr(edge_dst), L.idx_bytes(edge_src), E, None, F, float(cutoff_upper), L.ptr(out),
           L.dt_code(out), None, 0, None, 0, _st())
    return out


def _grad_edge_weight(x, grad_output, filter_out, edge_weight, edge_src, edge_dst, cutoff_upper):
    """exact d/d(edge_weight): C'(d_e) * sum_f x[src_e,f] * filter_out[e,f] * grad_output[dst_e,f]
    (kept by the reference's PyTorch path models/schnet.py:710-715, dropped by its Triton paths)."""
    E, F = edge_src.numel(), x.shape[1]
    g = torch.zeros(E, dtype=torch.float32, device=x.device)
    if E == 0:
        return g
    L.call("fmd_cfconv_grad_filter", L.ptr(x.contiguous()), L.ptr(grad_output.contiguous()), L.ptr(edge_weight),
           L.ptr(edge_src), L.ptr(edge_dst), L.idx_bytes(edge_src), E, None, F, float(cutoff_upper), None, 0,
           L.ptr(filter_out), L.dt_code(filter_out), L.ptr(g), 0, _st())
    return g


#: When True (default) CFConv backward also returns the exact gradient w.r.t. the edge distances.
#: Set to False to reproduce the reference Triton paths, which return None there
#: (kernels/csr_kernels.py:912, kernels/cfconv_kernels.py:415).
EXACT_CUTOFF_GRADIENT = True


class FusedCSRCFConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, filter_out, edge_weight, edge_src, edge_dst, dst_ptr, csr_perm, num_nodes, cutoff_upper,
                src_ptr, src_perm):
        ctx.save_for_backward(x, filter_out, edge_weight, edge_src, edge_dst, dst_ptr, csr_perm, src_ptr, src_perm)
        ctx.num_nodes, ctx.cutoff_upper = num_nodes, cutoff_upper
        return fused_csr_cfconv(x, filter_out, edge_weight, edge_src, dst_ptr, csr_perm, num_nodes, cutoff_upper)

    @staticmethod
    def backward(ctx, grad_output):
        x, filter_out, edge_weight, edge_src, edge_dst, dst_ptr, csr_perm, src_ptr, src_perm = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_x = grad_f = grad_w = None
        if ctx.needs_input_grad[0]:
            if src_ptr is None or src_perm is None:
                src_ptr, src_perm = _build_csr(edge_src, ctx.num_nodes)
            grad_x = fused_src_csr_grad_x(grad_output, filter_out, edge_weight, edge_dst, src_ptr, src_perm,
                                          ctx.num_nodes, ctx.cutoff_upper)
        if ctx.needs_input_grad[1]:
            grad_f = fused_grad_filter_out(x, grad_output, edge_weight, edge_src, edge_dst, ctx.cutoff_upper,
                                           out_dtype=filter_out.dtype)
        if ctx.needs_input_grad[2] and EXACT_CUTOFF_GRADIENT:
            grad_w = _grad_edge_weight(x, grad_output, filter_out, edge_weight, edge_src, edge_dst, ctx.cutoff_upper)
        return grad_x, grad_f, grad_w, None, None, None, None, None, None, None, None


def fused_csr_cfconv_autograd(x, filter_out, edge_weight, edge_src, edge_dst, dst_ptr, csr_perm, num_nodes,
                              cutoff_upper, src_ptr=None, src_perm=None):
    """reference kernels/csr_kernels.py:915."""
    return FusedCSRCFConvFunction.apply(x, filter_out, edge_weight, edge_src, edge_dst, dst_ptr, csr_perm, num_nodes,
                                        cutoff_upper, src_ptr, src_perm)


def fused_cutoff_gather_multiply_scatter_autograd(x, filter_out, edge_weight, edge_src, edge_dst, num_nodes,
                                                  cutoff_upper, src_ptr=None, src_perm=None):
    """reference kernels/cfconv_kernels.py:418."""
    dst_ptr, perm = _build_csr(edge_dst.contiguous(), num_nodes)
    return FusedCSRCFConvFunction.apply(x, filter_out, edge_weight, edge_src.contiguous(), edge_dst.contiguous(),
                                        dst_ptr, perm, num_nodes, cutoff_upper, src_ptr, src_perm)


# ------------------------------------------------------------------------------------------ distance + RBF


def fused_distance_gaussian_rbf_cutoff(pos, edge_src, edge_dst, centers, gamma, cutoff_upper):
    """(distances [E], rbf [E,R]) — reference kernels/cfconv_kernels.py:1578-1656."""
    _check_cuda(pos, edge_src, edge_dst, centers)
    assert pos.is_contiguous() and edge_src.is_contiguous() and edge_dst.is_contiguous()
    E, R = edge_src.numel(), centers.numel()
    dist = torch.empty(E, dtype=torch.float32, device=pos.device)
    rbf = torch.empty((E, R), dtype=torch.float32, device=pos.device)
    if E == 0:
        return dist, rbf
    L.call("fmd_dist_rbf_cutoff_fwd", L.ptr(pos), L.ptr(edge_src), L.ptr(edge_dst), L.idx_bytes(edge_src), E, None,
           L.ptr(centers.contiguous().float()), R, float(gamma), float(cutoff_upper), L.ptr(dist), L.ptr(rbf), _st())
    return dist, rbf


class FusedDistanceGaussianRBFCutoffFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, edge_src, edge_dst, centers, gamma, cutoff_upper):
        dist, rbf = fused_distance_gaussian_rbf_cutoff(pos, edge_src, edge_dst, centers, gamma, cutoff_upper)
        ctx.save_for_backward(pos, edge_src, edge_dst, centers, dist)
        ctx.gamma, ctx.cutoff_upper = float(gamma), float(cutoff_upper)
        return dist, rbf

    @staticmethod
    def backward(ctx, grad_distances, grad_rbf):
        """reference kernels/cfconv_kernels.py:1679-1735 (which needs `math` injected to run)."""
        pos, edge_src, edge_dst, centers, dist = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        E = edge_src.numel()
        grad_pos = torch.zeros_like(pos)
        if E == 0:
            return grad_pos, None, None, None, None, None
        g_d = torch.empty(E, dtype=torch.float32, device=pos.device)
        gd_in = grad_distances.contiguous() if grad_distances is not None else None
        if grad_rbf is None:
            g_d = gd_in if gd_in is not None else torch.zeros_like(dist)
        else:
            L.call("fmd_rbf_bwd", L.ptr(dist), L.ptr(grad_rbf.contiguous()), L.ptr(gd_in), E, None,
                   L.ptr(centers.contiguous().float()), centers.numel(), ctx.gamma, ctx.cutoff_upper, L.ptr(g_d), 0,
                   _st())
        L.call("fmd_edge_grad_to_pos_atomic", L.ptr(pos), L.ptr(edge_src), L.ptr(edge_dst), L.idx_bytes(edge_src),
               L.ptr(dist), L.ptr(g_d), E, None, L.ptr(grad_pos), _st())
        return grad_pos, None, None, None, None, None


def fused_distance_gaussian_rbf_cutoff_autograd(pos, edge_src, edge_dst, centers, gamma, cutoff_upper):
    """reference kernels/cfconv_kernels.py:1738."""
    return FusedDistanceGaussianRBFCutoffFunction.apply(pos, edge_src, edge_dst, centers, gamma, cutoff_upper)


# ------------------------------------------------------------------------------------------ dense layers


def fused_tanh_linear(x, weight, bias=None):
    """Y = tanh(X) @ W[K,N] + b, fp32 — reference kernels/cfconv_kernels.py:1846-1898."""
    _check_cuda(x, weight, bias)
    return _linear(x.contiguous(), weight.contiguous(), bias, pro_act=L.ACT_TANH)


class FusedTanhLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        tanh_x = torch.empty_like(x)  # saved for backward, like the reference (:1915)
        torch.tanh(x, out=tanh_x)
        ctx.save_for_backward(tanh_x, weight)
        return _linear(tanh_x, weight.contiguous(), bias)

    @staticmethod
    def backward(ctx, grad_output):
        tanh_x, weight = ctx.saved_tensors
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("flashmd B200 kernels: weight gradients are not provided (inference/MD only)")
        grad_x = None
        if ctx.needs_input_grad[0]:
            # (grad_y @ W^T) * (1 - tanh(x)^2); W[K,N]^T as a [K'=N, N'=K] matrix
            grad_x = _linear(grad_output.contiguous(), weight.t().contiguous(), aux=tanh_x)
        return grad_x, None, None


def fused_tanh_linear_autograd(x, weight, bias=None):
    """reference kernels/cfconv_kernels.py:1944."""
    return FusedTanhLinearFunction.apply(x, weight, bias)


def fused_linear_tanh(x, weight, bias=None):
    """Y = tanh(X @ W + b), fp32 — reference kernels/cfconv_kernels.py:543-593 (dead code there)."""
    _check_cuda(x, weight, bias)
    return _linear(x.contiguous(), weight.contiguous(), bias, epi_act=L.ACT_TANH)


class FusedLinearTanhFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        y = fused_linear_tanh(x, weight, bias)
        ctx.save_for_backward(y, weight)
        return y

    @staticmethod
    def backward(ctx, grad_output):
        y, weight = ctx.saved_tensors
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("flashmd B200 kernels: weight gradients are not provided (inference/MD only)")
        gz = grad_output.contiguous() * (1.0 - y * y)
        return _linear(gz, weight.t().contiguous()), None, None


def fused_linear_tanh_autograd(x, weight, bias=None):
    return FusedLinearTanhFunction.apply(x, weight, bias)


def fused_linear_tanh_fp16(x, weight, bias=None):
    """Y16 = tanh(X @ W16 + b16); X fp32 (rounded to fp16 in-kernel) or fp16 — reference
    kernels/cfconv_kernels.py:723-777 (clamped exp-based tanh, :449-454)."""
    _check_cuda(x, weight, bias)
    assert weight.dtype == torch.float16
    return _linear(x.contiguous(), weight.contiguous(), bias, out_dtype=torch.float16,
                   x_round_f16=(x.dtype == torch.float32), epi_act=L.ACT_TANH_CLAMPED)


def linear_fp16(x, weight, out_dtype=torch.float32):
    """Y = X16 @ W16 (fp32 accumulate) — reference kernels/cfconv_kernels.py:835-905."""
    _check_cuda(x, weight)
    assert weight.dtype == torch.float16
    return _linear(x.contiguous(), weight.contiguous(), None, out_dtype=out_dtype,
                   x_round_f16=(x.dtype == torch.float32))


class FusedLinearTanhFP16Function(torch.autograd.Function):
    """reference kernels/cfconv_kernels.py:1305-1361."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = fused_linear_tanh_fp16(x, weight, bias)
        ctx.save_for_backward(weight, y)
        ctx.input_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, grad_output):
        weight, y = ctx.saved_tensors
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("flashmd B200 kernels: weight gradients are not provided (inference/MD only)")
        grad_x = None
        if ctx.needs_input_grad[0]:
            # reference :963-1035: gz = (g * (1-y^2)) cast to fp16, then gz @ W^T, fp32 out
            g = grad_output.contiguous()
            gz = (g.float() * (1.0 - y.float() ** 2)).half()
            grad_x = _linear(gz, weight.t().contiguous(), out_dtype=torch.float32)
            if ctx.input_dtype == torch.float16:
                grad_x = grad_x.half()
        return grad_x, None, None


class LinearFP16Function(torch.autograd.Function):
    """reference kernels/cfconv_kernels.py:1364-1434 (both output dtypes)."""

    @staticmethod
    def forward(ctx, x, weight, out_dtype):
        ctx.save_for_backward(weight)
        return linear_fp16(x, weight, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        (weight,) = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("flashmd B200 kernels: weight gradients are not provided (inference/MD only)")
        grad_x = None
        if ctx.needs_input_grad[0]:
            # reference :1165-1226: fp32 grad x fp16 W^T -> fp16
            grad_x = _linear(grad_output.contiguous().float(), weight.t().contiguous(), out_dtype=torch.float16,
                             x_round_f16=True)
        return grad_x, None, None


def fused_linear_tanh_fp16_autograd(x, weight, bias=None):
    """reference kernels/cfconv_kernels.py:1437."""
    return FusedLinearTanhFP16Function.apply(x, weight, bias)


def linear_fp16_autograd(x, weight, out_dtype=torch.float32):
    """reference kernels/cfconv_kernels.py:1442."""
    return LinearFP16Function.apply(x, weight, out_dtype)


class _KernelHandle:
    """Placeholder for the reference's `*_kernel` Triton JIT objects: names the CUDA kernel that
    does the work here (they are launched through the C ABI, not directly from Python)."""

    def __init__(self, cuda_name):
        self.cuda_name = cuda_name

    def __repr__(self):
        return f"<sm_100a kernel {self.cuda_name} in libfmd_b200.so>"


fused_cutoff_gather_multiply_scatter_kernel = _KernelHandle("cfconv_csr_kernel")
fused_linear_tanh_kernel = _KernelHandle("linear_kernel")
fused_linear_tanh_fp16_kernel = _KernelHandle("linear_kernel")
linear_fp16_kernel = _KernelHandle("linear_kernel")
fused_distance_gaussian_rbf_cutoff_kernel = _KernelHandle("dist_rbf_fwd_kernel")
fused_tanh_linear_kernel = _KernelHandle("linear_kernel")
