"""offlinerlkit_b200 -- B200-native engine for OfflineRL-Kit's offline actor-critic gradient step.

The package mirrors the reference's import surface for the hot path only
(``buffer``, ``nets``, ``modules``, ``policy``, ``dynamics``, ``policy_trainer``,
``utils``); every ``learn`` / ``sample`` / ``rollout`` / ``step`` runs hand-written
sm_100a CUDA kernels from ``csrc/`` through the C-ABI in ``include/orlk_b200.h``.
There is no CPU fallback: using a compute entry point without the built library
or without a CUDA device raises.
"""
__version__ = "0.1.0"


def attach(policy, buffer=None):
    """Put the CUDA engine behind objects built by the unmodified reference (see attach_mode.py)."""
    from offlinerlkit_b200.attach_mode import attach as _attach      # (absolute: this file is exec'd by the root import shim)
    return _attach(policy, buffer)
