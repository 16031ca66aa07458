""""W16A16" modules (reference models/gptq.py:30-443 — not GPTQ at all: the filter and output MLPs are cast
with .half() and evaluated by the FP16 operators).  The modules below keep that surface; on the step
path the fused engine consumes the same fp16 weights directly."""
import torch

from .mlp import MLP


def _linears(mlp: MLP):
    return [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]


class GPTQW16A16FilterNetwork(torch.nn.Module):
    """tanh(x W0^T + b0) W1^T with fp16 weights/bias/activations and fp32 accumulation."""

    def __init__(self, in_features, hidden_features: int = None, out_features: int = None):
        """Either the reference's signature (in_features, hidden_features, out_features: Xavier-initialised fp16 weights,
        zero bias; models/gptq.py:51-76) or an existing fp32 MLP to convert (what `from_mlp` and
        apply_gptq_w16a16_to_model do)."""
        super().__init__()
        mlp = in_features if isinstance(in_features, torch.nn.Module) else MLP(
            [int(in_features), int(hidden_features), int(out_features)], torch.nn.Tanh(), last_bias=False)
        lin = _linears(mlp)
        assert len(lin) == 2 and lin[1].bias is None, "filter network must be Linear-Tanh-Linear(no bias)"
        self.register_buffer("w0", lin[0].weight.detach().t().contiguous().half())      # [K, N]
        self.register_buffer("b0", lin[0].bias.detach().half() if lin[0].bias is not None else None)
        self.register_buffer("w1", lin[1].weight.detach().t().contiguous().half())
        self.in_features, self.hidden_features, self.out_features = lin[0].in_features, lin[0].out_features, lin[1].out_features

    @classmethod
    def from_mlp(cls, mlp: torch.nn.Module) -> "GPTQW16A16FilterNetwork":
        return cls(mlp)

    # the reference's attribute names ([in, out] fp16 tensors)
    weight0 = property(lambda self: self.w0)
    bias0 = property(lambda self: self.b0)
    weight1 = property(lambda self: self.w1)

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

    def __init__(self, in_features, hidden1_features: int = None, hidden2_features: int = None, out_features: int = 1):
        """Either the reference's signature (in, hidden1, hidden2, out; models/gptq.py:215-256) or an existing fp32 MLP."""
        super().__init__()
        mlp = in_features if isinstance(in_features, torch.nn.Module) else MLP(
            [int(in_features), int(hidden1_features), int(hidden2_features), int(out_features)], torch.nn.Tanh(),
            last_bias=False)
        lin = _linears(mlp)
        self.n_layers = len(lin)
        for i, l in enumerate(lin):
            self.register_buffer(f"w{i}", l.weight.detach().t().contiguous().half())
            self.register_buffer(f"b{i}", l.bias.detach().half() if l.bias is not None else None)

    @classmethod
    def from_mlp(cls, mlp: torch.nn.Module) -> "GPTQW16A16OutputNetwork":
        return cls(mlp)

    def __getattr__(self, name):
        # the reference's attribute names: weight<i> / bias<i>
        if name.startswith("weight") and name[6:].isdigit():
            return super().__getattr__("w" + name[6:])
        if name.startswith("bias") and name[4:].isdigit():
            return super().__getattr__("b" + name[4:])
        return super().__getattr__(name)

    def reset_parameters(self):
        pass

    def forward(self, x: torch.Tensor, data=None) -> torch.Tensor:
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
