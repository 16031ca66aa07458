from .cutoff import CosineCutoff, IdentityCutoff, ShiftedCosineCutoff  # noqa: F401
from .gradients import EnergyOut, GradientsOut, SumOut  # noqa: F401
from .mlp import MLP, TypesMLP  # noqa: F401
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
