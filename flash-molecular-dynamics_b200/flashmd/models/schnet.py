"""CGSchNet modules with the reference's constructor / attribute / toggle surface
(reference models/schnet.py:52-56 toggles, :104-369 SchNet, :440-548 InteractionBlock, :551-740 CFConv,
:743-841 StandardSchNet).  On CUDA tensors the forward dispatches to the sm_100a operators in
`flashmd.kernels` (same dispatch rules as the reference); on CPU (or with the toggles off) it is plain
PyTorch — the reference's `--disable_optim` semantics.  Simulations do not go through this module-level
autograd path at all: `flashmd.simulation` lowers the whole model to the fused engine (flashmd/engine.py)."""
import os
import warnings
from typing import List, Optional

import torch

from ..data._keys import ATOM_TYPE_KEY, ENERGY_KEY, POSITIONS_KEY
from ..geometry import compute_distances
from ..neighbor_list import atomic_data2neighbor_list, validate_neighborlist
from .cutoff import CosineCutoff
from .mlp import MLP, init_xavier_uniform

# read once at import, like the reference: `--disable_optim` must set them before importing this module
USE_TRITON_MESSAGE_PASSING = os.environ.get("MLCG_USE_TRITON_MESSAGE_PASSING", "1") == "1"
USE_FUSED_RBF = os.environ.get("MLCG_USE_FUSED_RBF", "1")          # "1" / "0" / anything else = auto
USE_FUSED_TANH_LINEAR = os.environ.get("MLCG_USE_FUSED_TANH_LINEAR", "1") == "1"
USE_CSR = os.environ.get("MLCG_USE_CSR", "1") == "1"
USE_SRC_CSR_GRAD_X = os.environ.get("MLCG_USE_SRC_CSR_GRAD_X", "1") == "1"
TRITON_AVAILABLE = True   # kept for API compatibility: the operators are CUDA kernels, always present on a GPU box


def _kernels():
    from .. import kernels
    return kernels


class CFConv(torch.nn.Module):
    """Continuous-filter convolution: lin1 -> filter(rbf) * x[src] * C(d) summed at dst -> lin2."""

    def __init__(self, filter_network: torch.nn.Module, cutoff: torch.nn.Module, in_channels: int = 128,
                 out_channels: int = 128, num_filters: int = 128, aggr: str = "add", use_triton: bool = True):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("only aggr='add' is supported")
        self.aggr = aggr
        self.lin1 = torch.nn.Linear(in_channels, num_filters, bias=False)
        self.lin2 = torch.nn.Linear(num_filters, out_channels)
        self.filter_network = filter_network
        self.cutoff = cutoff
        self.use_triton = use_triton and USE_TRITON_MESSAGE_PASSING
        self.reset_parameters()

    def reset_parameters(self):
        self.filter_network.reset_parameters()
        init_xavier_uniform(self.lin1)
        init_xavier_uniform(self.lin2)

    def forward(self, x, edge_index, edge_weight, edge_attr, csr_data: Optional[dict] = None):
        a = self.lin1(x)
        W = self.filter_network(edge_attr)
        src, dst = edge_index[0], edge_index[1]
        n = x.shape[0]
        rc = float(self.cutoff.cutoff_upper)
        fusable = isinstance(self.cutoff, CosineCutoff) and self.cutoff.cutoff_lower == 0 and x.is_cuda
        if fusable and csr_data is not None and "dst_ptr" in csr_data:
            m = _kernels().fused_csr_cfconv_autograd(a, W, edge_weight, src, dst, csr_data["dst_ptr"], csr_data["csr_perm"],
                                                     n, rc, csr_data.get("src_ptr"), csr_data.get("src_perm"))
        elif fusable and self.use_triton:
            m = _kernels().fused_cutoff_gather_multiply_scatter_autograd(a, W, edge_weight, src, dst, n, rc)
        else:
            msg = a[src] * W.to(a.dtype) * self.cutoff(edge_weight).unsqueeze(-1)
            m = torch.zeros_like(a).index_add(0, dst, msg)
        return self.lin2(m)


class InteractionBlock(torch.nn.Module):
    """CFConv -> activation -> Linear (the residual add happens in SchNet.forward)."""

    def __init__(self, cfconv_layer: torch.nn.Module, hidden_channels: int = 128,
                 activation: torch.nn.Module = torch.nn.Tanh()):
        super().__init__()
        self.conv = cfconv_layer
        self.activation = activation
        self.lin = torch.nn.Linear(hidden_channels, hidden_channels)
        self.use_fused_tanh_linear = USE_FUSED_TANH_LINEAR and isinstance(activation, torch.nn.Tanh)
        self.reset_parameters()

    def reset_parameters(self):
        init_xavier_uniform(self.lin)

    def forward(self, x, edge_index, edge_weight, edge_attr, csr_data: Optional[dict] = None):
        c = self.conv(x, edge_index, edge_weight, edge_attr, csr_data)
        if self.use_fused_tanh_linear and c.is_cuda and c.dtype == torch.float32:
            return _kernels().fused_tanh_linear_autograd(c, self.lin.weight.t().contiguous(), self.lin.bias)
        return self.lin(self.activation(c))


class SchNet(torch.nn.Module):
    name: str = "SchNet"

    def __init__(self, embedding_layer: torch.nn.Module, interaction_blocks: List[torch.nn.Module],
                 rbf_layer: torch.nn.Module, output_network: torch.nn.Module, self_interaction: bool = False,
                 max_num_neighbors: int = 1000):
        super().__init__()
        if self_interaction:
            raise NotImplementedError("`self_interaction` only exists for compatibility and must stay False")
        self.embedding_layer = embedding_layer
        self.rbf_layer = rbf_layer
        self.max_num_neighbors = max_num_neighbors
        self.self_interaction = self_interaction
        if isinstance(interaction_blocks, InteractionBlock):
            interaction_blocks = [interaction_blocks]
        if not isinstance(interaction_blocks, (list, tuple, torch.nn.Sequential)):
            raise RuntimeError("interaction_blocks must be a list of InteractionBlock")
        self.interaction_blocks = torch.nn.Sequential(*interaction_blocks)
        self.output_network = output_network
        self.reset_parameters()

    def reset_parameters(self):
        self.embedding_layer.reset_parameters()
        self.rbf_layer.reset_parameters()
        for block in self.interaction_blocks:
            block.conv.reset_parameters()
            block.reset_parameters()
        self.output_network.reset_parameters()

    # ---- helpers
    @property
    def cutoff_upper(self) -> float:
        return float(self.rbf_layer.cutoff.cutoff_upper)

    @staticmethod
    def neighbor_list(data, rcut: float, max_num_neighbors: int = 1000) -> dict:
        return {SchNet.name: atomic_data2neighbor_list(data, rcut, self_interaction=False,
                                                       max_num_neighbors=max_num_neighbors)}

    def is_nl_compatible(self, nl) -> bool:
        return validate_neighborlist(nl) and nl.get("rcut") == self.cutoff_upper and not nl.get("self_interaction")

    def _fused_rbf_ok(self, pos) -> bool:
        c = self.rbf_layer.cutoff
        return (USE_FUSED_RBF == "1" and pos.is_cuda and pos.dtype == torch.float32 and isinstance(c, CosineCutoff)
                and c.cutoff_lower == 0 and not getattr(self.rbf_layer, "trainable", False)
                and hasattr(self.rbf_layer, "offset"))

    def forward(self, data):
        pos = data[POSITIONS_KEY]
        x = self.embedding_layer(data[ATOM_TYPE_KEY])
        nl = data.neighbor_list.get(self.name) if isinstance(data.neighbor_list, dict) else None
        if not self.is_nl_compatible(nl):
            nl = self.neighbor_list(data, self.cutoff_upper, self.max_num_neighbors)[self.name]
        edge_index = nl["index_mapping"]
        src, dst = edge_index[0], edge_index[1]
        if self._fused_rbf_ok(pos):
            d, rbf = _kernels().fused_distance_gaussian_rbf_cutoff_autograd(
                pos, src, dst, self.rbf_layer.offset, float(self.rbf_layer.coeff), self.cutoff_upper)
        else:
            d = compute_distances(pos, edge_index)
            rbf = self.rbf_layer(d)
        csr = None
        if pos.is_cuda and USE_CSR:
            k = _kernels()
            dst_ptr, csr_perm = k.build_csr_index(dst, x.shape[0])
            csr = {"dst_ptr": dst_ptr, "csr_perm": csr_perm}
            if USE_SRC_CSR_GRAD_X:
                csr["src_ptr"], csr["src_perm"] = k.build_src_csr_index(src, x.shape[0])
        for block in self.interaction_blocks:
            x = x + block(x, edge_index, d, rbf, csr)
        e_atom = self.output_network(x).flatten().to(pos.dtype)
        n_mol = data.ptr.numel() - 1
        energy = torch.zeros(n_mol, dtype=e_atom.dtype, device=e_atom.device).index_add(0, data.batch, e_atom)
        data.out[self.name] = {ENERGY_KEY: energy}
        return data


class StandardSchNet(SchNet):
    """The usual architecture: Embedding, `num_interactions` x (filter MLP [R, F, F] no last bias, CFConv,
    InteractionBlock), output MLP [hidden, *widths, 1] without last bias."""

    def __init__(self, rbf_layer: torch.nn.Module, cutoff: torch.nn.Module, output_hidden_layer_widths: List[int],
                 hidden_channels: int = 128, embedding_size: int = 100, num_filters: int = 128,
                 num_interactions: int = 3, activation: torch.nn.Module = torch.nn.Tanh(),
                 max_num_neighbors: int = 1000, aggr: str = "add"):
        if num_interactions < 1:
            raise ValueError("At least one interaction block must be specified")
        for side in ("lower", "upper"):
            a, b = getattr(cutoff, f"cutoff_{side}"), getattr(rbf_layer.cutoff, f"cutoff_{side}")
            if a != b:
                warnings.warn(f"Cutoff function {side} cutoff, {a}, and radial basis function  {side} cutoff, {b}, "
                              "do not match.")
        embedding = torch.nn.Embedding(embedding_size, hidden_channels)
        blocks = []
        for _ in range(num_interactions):
            filt = MLP([rbf_layer.num_rbf, num_filters, num_filters], activation_func=activation, last_bias=False)
            conv = CFConv(filt, cutoff=cutoff, num_filters=num_filters, in_channels=hidden_channels,
                          out_channels=hidden_channels, aggr=aggr)
            blocks.append(InteractionBlock(conv, hidden_channels, activation))
        out_net = MLP([hidden_channels] + list(output_hidden_layer_widths) + [1], activation_func=activation,
                      last_bias=False)
        super().__init__(embedding, blocks, rbf_layer, out_net, max_num_neighbors=max_num_neighbors)
