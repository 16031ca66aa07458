from .cutoff import CosineCutoff, IdentityCutoff, ShiftedCosineCutoff  # noqa: F401
from .gradients import EnergyOut, GradientsOut, SumOut  # noqa: F401
from .mlp import MLP  # noqa: F401
from .radial_basis import GaussianBasis  # noqa: F401
from .schnet import CFConv, InteractionBlock, SchNet, StandardSchNet  # noqa: F401
from .gptq import (GPTQW16A16FilterNetwork, GPTQW16A16OutputNetwork, apply_gptq_w16a16_to_model,  # noqa: F401
                   validate_gptq_w16a16)
