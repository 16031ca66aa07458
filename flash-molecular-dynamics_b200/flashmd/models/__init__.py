from .cutoff import CosineCutoff, IdentityCutoff, ShiftedCosineCutoff  # noqa: F401
from .gradients import EnergyOut, GradientsOut, SumOut  # noqa: F401
from .mlp import MLP, TypesMLP  # noqa: F401
from .utils import desparsify_prior_module, sparsify_prior_module  # noqa: F401
from .radial_basis import GaussianBasis  # noqa: F401
from .schnet import CFConv, InteractionBlock, SchNet, StandardSchNet  # noqa: F401
from .gptq import (GPTQW16A16FilterNetwork, GPTQW16A16OutputNetwork, apply_gptq_w16a16_to_model,  # noqa: F401
                   validate_gptq_w16a16)


def load_and_adapt_old_checkpoint(f, **kwargs):
    """Reference models/pyg_forward_compatibility.py:17-243 re-wires pickled mlcg / PyG-era checkpoints (MessagePassing
    internals of old torch_geometric versions) onto the current classes.  That archaeology needs torch_geometric and is out
    of scope here: a checkpoint pickled from THIS package (or any module that unpickles cleanly) is
    loaded and returned; anything else raises with the reason."""
    import torch
    try:
        return torch.load(f, weights_only=False, **kwargs)
    except Exception as err:  # noqa: BLE001
        raise NotImplementedError(
            "load_and_adapt_old_checkpoint: this checkpoint does not unpickle against the drop-in classes; adapting "
            f"legacy mlcg/PyG pickles is not implemented (no torch_geometric here): {err!r}") from err


def _pyg_compat_unavailable(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            f"flashmd.models.{name} rewires torch_geometric MessagePassing internals of legacy checkpoints (reference "
            "models/pyg_forward_compatibility.py); the drop-in has no torch_geometric dependency and its CFConv is not a "
            "MessagePassing module, so there is nothing to refresh")
    f.__name__ = name
    return f


# names the reference exports from models/pyg_forward_compatibility.py (checkpoint archaeology)
fixed_pyg_inspector = _pyg_compat_unavailable("fixed_pyg_inspector")
get_refreshed_cfconv_layer = _pyg_compat_unavailable("get_refreshed_cfconv_layer")
refresh_module_ = _pyg_compat_unavailable("refresh_module_")
