"""pfpn_b200 -- B200-native (sm_100a) implementation of the PFPN policy-head hot path.

Host side mirrors the reference's Python interface for this path
(``networks/utils.py``, ``networks/actor_critic/{a2c,ppo,sac}.py``); compute is
hand-written CUDA behind the C ABI of ``include/pfpn_b200.h``.  Importing this
package loads ``libpfpn_b200.so`` and raises if it is missing -- there is no
CPU / PyTorch fallback.
"""
from . import _cabi  # noqa: F401  (fails loudly when the library is absent)
from .distribution import MixtureGaussianDistribution  # noqa: F401

__all__ = ["MixtureGaussianDistribution"]
