"""border_b200 -- B200-native training hot path of laboroai/border behind a C ABI.

Python here is only the host-side mirror of the reference's trait surface (for tests, bench.py
and the examples); the product is libborder_b200.so (include/border_b200.h).
"""
from .replay import GenericTransitionBatch, PerConfig, SimpleReplayBuffer, SimpleReplayBufferConfig  # noqa: F401
from .agents import (AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, Iqn, IqnConfig, MlpConfig,  # noqa: F401
                     OptimizerConfig, Sac, SacConfig, Softmax)
from .atari import AtariPreprocessor  # noqa: F401
