"""Import-path parity with the reference's `flashmd.kernels.cfconv_kernels`."""
from .ops import (  # noqa: F401
    fused_cutoff_gather_multiply_scatter_kernel, fused_cutoff_gather_multiply_scatter,
    fused_cutoff_gather_multiply_scatter_autograd, fused_grad_filter_out,
    fused_linear_tanh_kernel, fused_linear_tanh, fused_linear_tanh_autograd,
    fused_linear_tanh_fp16_kernel, fused_linear_tanh_fp16, fused_linear_tanh_fp16_autograd,
    linear_fp16_kernel, linear_fp16, linear_fp16_autograd,
    fused_distance_gaussian_rbf_cutoff_kernel, fused_distance_gaussian_rbf_cutoff,
    fused_distance_gaussian_rbf_cutoff_autograd,
    fused_tanh_linear_kernel, fused_tanh_linear, fused_tanh_linear_autograd,
)
