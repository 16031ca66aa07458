from .gaussian import GaussianBasis, _RadialBasis  # noqa: F401
