"""Import-path parity with the reference's `flashmd.kernels.csr_kernels`."""
from .ops import (  # noqa: F401
    build_csr_index, build_src_csr_index, fused_csr_cfconv, fused_csr_cfconv_autograd, fused_src_csr_grad_x,
)
