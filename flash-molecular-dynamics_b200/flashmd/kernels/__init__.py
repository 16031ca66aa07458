"""`flashmd.kernels` — the operator API of the reference (kernels/__init__.py:47-75), backed by
sm_100a CUDA kernels instead of Triton."""
from .cfconv_kernels import *  # noqa: F401,F403
from .cfconv_kernels import fused_grad_filter_out  # noqa: F401
from .csr_kernels import *  # noqa: F401,F403

__all__ = [
    "fused_cutoff_gather_multiply_scatter_kernel", "fused_cutoff_gather_multiply_scatter",
    "fused_cutoff_gather_multiply_scatter_autograd",
    "fused_linear_tanh_kernel", "fused_linear_tanh", "fused_linear_tanh_autograd",
    "fused_linear_tanh_fp16_kernel", "fused_linear_tanh_fp16", "fused_linear_tanh_fp16_autograd",
    "linear_fp16_kernel", "linear_fp16", "linear_fp16_autograd",
    "fused_distance_gaussian_rbf_cutoff_kernel", "fused_distance_gaussian_rbf_cutoff",
    "fused_distance_gaussian_rbf_cutoff_autograd",
    "fused_tanh_linear_kernel", "fused_tanh_linear", "fused_tanh_linear_autograd",
    "build_csr_index", "build_src_csr_index", "fused_csr_cfconv", "fused_csr_cfconv_autograd",
    "fused_src_csr_grad_x",
]
