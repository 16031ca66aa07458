"""Lower a reference-style model object (SumOut{name: GradientsOut(SchNet | prior)}) and a collated
batch to the fused engine (flashmd/engine.py): flat SchNet weights + condensed prior tables.  This is the
set-up-time counterpart of reference simulation/specialize_prior.py:76-207 (per-term parameter vectors)
and of simulation/base.py:338-369 (_attach_model: eval, W16A16 swap, freeze)."""
from typing import Dict, List, Optional

import torch

from .. import _lib as L
from ..engine import ForceField, PriorTerm, SchNetWeights
from ..models.gptq import GPTQW16A16FilterNetwork, GPTQW16A16OutputNetwork
from ..models.gradients import GradientsOut, SumOut
from ..models.mlp import MLP
from ..models.schnet import SchNet
from ..models.cutoff import CosineCutoff
from ..models.radial_basis.gaussian import GaussianBasis


class NotLowerable(RuntimeError):
    pass


def _unwrap(m):
    return m.model if isinstance(m, GradientsOut) else m


def schnet_flat_tensors(net: SchNet) -> Dict[str, torch.Tensor]:
    """nn.Linear-layout ([out, in]) fp32 tensors with the engine's key names."""
    if (not isinstance(net.rbf_layer, GaussianBasis) or not isinstance(net.rbf_layer.cutoff, CosineCutoff)
            or net.rbf_layer.cutoff.cutoff_lower != 0):
        raise NotLowerable("the fused step needs GaussianBasis(CosineCutoff(0, rc))")
    t = {"embedding": net.embedding_layer.weight}
    for l, blk in enumerate(net.interaction_blocks):
        cf = blk.conv
        if not isinstance(blk.activation, torch.nn.Tanh):
            raise NotLowerable("the fused step supports Tanh activations only")
        if float(cf.cutoff.cutoff_upper) != float(net.rbf_layer.cutoff.cutoff_upper) or cf.cutoff.cutoff_lower != 0:
            raise NotLowerable("CFConv cutoff must equal the radial-basis cutoff")
        t[f"b{l}.lin1_w"] = cf.lin1.weight
        fn = cf.filter_network
        if isinstance(fn, GPTQW16A16FilterNetwork):
            t[f"b{l}.f0_w"], t[f"b{l}.f1_w"] = fn.w0.t().float(), fn.w1.t().float()
            t[f"b{l}.f0_b"] = fn.b0.float() if fn.b0 is not None else torch.zeros(fn.w0.shape[1], device=fn.w0.device)
        elif isinstance(fn, MLP):
            lin = [m for m in fn.layers if isinstance(m, torch.nn.Linear)]
            if len(lin) != 2 or lin[1].bias is not None:
                raise NotLowerable("filter network must be Linear-Tanh-Linear(no bias)")
            t[f"b{l}.f0_w"], t[f"b{l}.f1_w"] = lin[0].weight, lin[1].weight
            t[f"b{l}.f0_b"] = lin[0].bias if lin[0].bias is not None else torch.zeros_like(lin[0].weight[:, 0])
        else:
            raise NotLowerable(f"unsupported filter network {type(fn).__name__}")
        t[f"b{l}.lin2_w"], t[f"b{l}.lin2_b"] = cf.lin2.weight, cf.lin2.bias
        t[f"b{l}.lin_w"], t[f"b{l}.lin_b"] = blk.lin.weight, blk.lin.bias
    on = net.output_network
    if isinstance(on, GPTQW16A16OutputNetwork):
        for i in range(on.n_layers):
            t[f"out{i}_w"] = getattr(on, f"w{i}").t().float()
            b = getattr(on, f"b{i}")
            if b is not None:
                t[f"out{i}_b"] = b.float()
    elif isinstance(on, MLP):
        lin = [m for m in on.layers if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lin):
            t[f"out{i}_w"] = m.weight
            if m.bias is not None:
                t[f"out{i}_b"] = m.bias
        if lin[-1].bias is not None or lin[-1].out_features != 1:
            raise NotLowerable("output network must end in Linear(., 1, bias=False)")
    else:
        raise NotLowerable(f"unsupported output network {type(on).__name__}")
    return {k: v.detach() for k, v in t.items()}


def prior_term(prior, data, device) -> PriorTerm:
    """Per-term flat parameter vectors of one prior on the collated batch (what condense_* produces)."""
    if getattr(prior, "kernel_kind", None) is None:
        raise NotLowerable(f"prior {type(prior).__name__} has no CUDA kernel")
    nl = data.neighbor_list[prior.name]
    mapping = nl["index_mapping"].to(device=device, dtype=torch.int32).contiguous()
    mb = nl["mapping_batch"].to(device=device, dtype=torch.int32).contiguous()
    p = prior.data2parameters(data)
    f = lambda x: x.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
    kind = prior.kernel_kind
    if kind in (L.PRIOR_BONDS, L.PRIOR_ANGLES, L.PRIOR_RAW_ANGLES, L.PRIOR_IMPROPERS, L.PRIOR_SHIFTED_IMPROPERS):
        return PriorTerm(kind, mapping, mb, f(p["k"]), f(p["x0"]))
    if kind == L.PRIOR_DIHEDRALS:
        return PriorTerm(kind, mapping, mb, f(p["k1s"]), f(p["k2s"]), f(p["v_0"].flatten()), int(p["k1s"].shape[1]))
    if kind in (L.PRIOR_POLY_BONDS, L.PRIOR_POLY_ANGLES):
        if kind == L.PRIOR_POLY_BONDS:
            from ..geometry import compute_distances
            probe = torch.tensor([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
            try:     # a Polynomial on a bond set must measure plain distances (users attach the feature function)
                ok = float(prior.compute_features(probe, torch.tensor([[0], [1]]))) == 5.0
            except Exception:  # noqa: BLE001
                ok = False
            if not ok:
                raise NotLowerable("Polynomial prior of order 2 whose feature is not the bond length")
        return PriorTerm(kind, mapping, mb, f(p["ks"]), None, f(p["v_0s"].flatten()))
    if kind == L.PRIOR_RESTRICTED_ANGLES:
        ks = torch.stack([p["a"], p["b"], p["c"], p["d"], p["k"]], 1)
        return PriorTerm(kind, mapping, mb, f(ks), None, f(p["v_0"].flatten()))
    return PriorTerm(kind, mapping, mb, f(p["sigma"]))


def lower(model: torch.nn.Module, data, precision: str, exact_cutoff_grad: bool = True,
          edge_capacity: Optional[int] = None) -> ForceField:
    """Build the fused ForceField for `model` on the CUDA batch `data`; raises NotLowerable otherwise."""
    if not isinstance(model, SumOut):
        raise NotLowerable("expected SumOut(ModuleDict{name: GradientsOut(model)})")
    dev = data.pos.device
    if dev.type != "cuda":
        raise NotLowerable("the fused step needs a CUDA device")
    weights, priors = None, []
    for name, sub in model.models.items():
        m = _unwrap(sub)
        if isinstance(m, SchNet):
            if weights is not None:
                raise NotLowerable("more than one SchNet")
            rc = float(m.rbf_layer.cutoff.cutoff_upper)
            # the basis parameters come from the module (a trained GaussianBasis carries its own offset / coeff)
            weights = SchNetWeights.from_flat(schnet_flat_tensors(m), rc, int(m.rbf_layer.num_rbf), dev,
                                              rbf_centers=m.rbf_layer.offset.detach(),
                                              rbf_gamma=float(m.rbf_layer.coeff.detach()))
            max_nn = int(m.max_num_neighbors)
            sizes = (data.ptr[1:] - data.ptr[:-1])
            if int(sizes.max()) - 1 > max_nn:
                raise NotLowerable(f"max_num_neighbors = {max_nn} truncates the neighbour list of a {int(sizes.max())}-bead "
                                   "molecule: the fused step needs the full symmetric list")
        else:
            priors.append(prior_term(m, data, dev))
    ptr = data.ptr.to(dev)
    kw = {}
    if weights is not None:
        kw["max_num_neighbors"] = max_nn
    return ForceField(weights, priors, data.atom_types.to(dev), ptr, precision=precision,
                      exact_cutoff_grad=exact_cutoff_grad, edge_capacity=edge_capacity, **kw)


SchNetWeights.from_module = staticmethod(
    lambda net, device=None: SchNetWeights.from_flat(schnet_flat_tensors(net), float(net.rbf_layer.cutoff.cutoff_upper),
                                                     int(net.rbf_layer.num_rbf),
                                                     device or net.embedding_layer.weight.device,
                                                     rbf_centers=net.rbf_layer.offset.detach(),
                                                     rbf_gamma=float(net.rbf_layer.coeff.detach())))
