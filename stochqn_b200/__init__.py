"""stochqn_b200 - B200-native (sm_100a) implementation of the stochQN optimizer step.

The product is the pair of CUDA shared libraries in ``stochqn_b200/lib`` that export the
reference's C ABI (``include/stochqn.h``); this package is the thin Python host side above
it (ctypes binding + free-mode classes mirroring the reference's ``stochqn/_optimizers.py``).
"""
from . import _abi  # noqa: F401

__all__ = ["_abi"]
__version__ = "0.1.0"
