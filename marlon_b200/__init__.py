"""marlon_b200 -- B200-native batched environment step for MARLon / CyberBattleSim.

Hot path: hand-written sm_100a CUDA kernels behind the C ABI of ``include/cbx.h`` (``marlon_b200/libcbx.so``).
Host side: the reference's Python surface (CyberBattleEnv, AttackerEnvWrapper, DefenderEnvWrapper,
MaskedDiscreteAttackerWrapper, MultiAgentUniverse, an SB3-style VecEnv).  No CPU fallback.
"""
__version__ = "0.1.0"

from . import _abi, config, model, registry, scenario, scenarios  # noqa: F401
from .config import (AttackerGoal, DefenderConstraint, DefenderGoal, ScanAndReimageCompromisedMachines,  # noqa: F401
                     make_config)


def __getattr__(name):
    # heavier modules (torch) are imported on first use
    import importlib

    lazy = {
        "Batch": ("batch", "Batch"),
        "CyberBattleEnv": ("cyberbattle_env", "CyberBattleEnv"),
        "AttackerEnvWrapper": ("wrappers", "AttackerEnvWrapper"),
        "DefenderEnvWrapper": ("wrappers", "DefenderEnvWrapper"),
        "MaskedDiscreteAttackerWrapper": ("wrappers", "MaskedDiscreteAttackerWrapper"),
        "EnvironmentEventSource": ("wrappers", "EnvironmentEventSource"),
        "MultiAgentUniverse": ("universe", "MultiAgentUniverse"),
        "MultiAgentUniversalEnv": ("universe", "MultiAgentUniversalEnv"),
        "BatchedVecEnv": ("vec_env", "BatchedVecEnv"),
        "make": ("cyberbattle_env", "make"),
    }
    if name in lazy:
        mod, attr = lazy[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
