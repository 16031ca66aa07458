"""flashmd — B200-native drop-in for the CGSchNet force-field + Langevin step of FlashMD.

Same import paths as the reference package (`flashmd.kernels`, `flashmd.models`, `flashmd.prior`,
`flashmd.simulation`, ...); the per-step work runs in hand-written sm_100a CUDA kernels behind the
C ABI in include/fmd_b200.h.
"""
__version__ = "0.1.0"
