""""W16A16" modules (reference models/gptq.py:30-443 — not GPTQ at all: the filter and output MLPs are cast
with .half() and evaluated by the FP16 operators).  The modules below keep that surface; on the step
path the fused engine consumes the same fp16 weights directly."""
import torch

from .mlp import MLP


def _linears(mlp: MLP):
    return [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]


class GPTQW16A16FilterNetwork(torch.nn.Module):
    """tanh(x W0^T + b0) W1^T with fp16 weights/bias/activations and fp32 accumulation."""

    def __init__(self, mlp: MLP):
        super().__init__()
        lin = _linears(mlp)
        assert len(lin) == 2 and lin[1].bias is None, "filter network must be Linear-Tanh-Linear(no bias)"
        self.register_buffer("w0", lin[0].weight.detach().t().contiguous().half())      # [K, N]
        self.register_buffer("b0", lin[0].bias.detach().half() if lin[0].bias is not None else None)
        self.register_buffer("w1", lin[1].weight.detach().t().contiguous().half())
        self.in_features, self.out_features = lin[0].in_features, lin[1].out_features

    def reset_parameters(self):
        pass

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("GPTQW16A16FilterNetwork is CUDA-only (as in the reference); use gptq=None on CPU")
        from .. import kernels as K
        t = K.fused_linear_tanh_fp16_autograd(x, self.w0, self.b0)
        return K.linear_fp16_autograd(t, self.w1, out_dtype=torch.float16)


class GPTQW16A16OutputNetwork(torch.nn.Module):
    """Output MLP: fused linear+tanh fp16 layers, last layer fp16 in -> fp32 out (no bias)."""

    def __init__(self, mlp: MLP):
        super().__init__()
        lin = _linears(mlp)
        self.n_layers = len(lin)
        for i, l in enumerate(lin):
            self.register_buffer(f"w{i}", l.weight.detach().t().contiguous().half())
            self.register_buffer(f"b{i}", l.bias.detach().half() if l.bias is not None else None)

    def reset_parameters(self):
        pass

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("GPTQW16A16OutputNetwork is CUDA-only (as in the reference); use gptq=None on CPU")
        from .. import kernels as K
        for i in range(self.n_layers - 1):
            x = K.fused_linear_tanh_fp16_autograd(x, getattr(self, f"w{i}"), getattr(self, f"b{i}"))
        i = self.n_layers - 1
        y = K.linear_fp16_autograd(x if x.dtype == torch.float16 else x.half(), getattr(self, f"w{i}"),
                                   out_dtype=torch.float32)
        b = getattr(self, f"b{i}")
        return y if b is None else y + b.float()


def _schnets(model):
    from .schnet import SchNet
    return [m for m in model.modules() if isinstance(m, SchNet)]


def apply_gptq_w16a16_to_model(model: torch.nn.Module, verbose: bool = False) -> torch.nn.Module:
    """Swap every SchNet's filter networks and output network for the W16A16 modules (in place)."""
    for net in _schnets(model):
        for block in net.interaction_blocks:
            if isinstance(block.conv.filter_network, MLP):
                block.conv.filter_network = GPTQW16A16FilterNetwork(block.conv.filter_network)
        if isinstance(net.output_network, MLP):
            net.output_network = GPTQW16A16OutputNetwork(net.output_network)
        net.gptq = "w16a16"
    return model


def validate_gptq_w16a16(model: torch.nn.Module) -> bool:
    nets = _schnets(model)
    return bool(nets) and all(isinstance(b.conv.filter_network, GPTQW16A16FilterNetwork)
                              for n in nets for b in n.interaction_blocks)
