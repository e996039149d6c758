"""hybridmc_b200 — B200-native ensemble engine for HybridMC's event-driven hybrid
Monte Carlo hot path (reference: margaritacolberg/hybridmc, src/main.cc:45-305).

The package holds only what the path needs: `csrc/` (the sm_100a CUDA kernels and
the C-ABI of include/hybridmc_b200.h) and the host-side mirror of the reference's
interface (`params`, `engine`).  There is no CPU fallback: importing works
anywhere, but creating an engine without the built CUDA library or without a
CUDA device raises.
"""
from .params import HmcParams, params_from_json, transition_jsons  # noqa: F401

__all__ = ["HmcParams", "params_from_json", "transition_jsons"]
