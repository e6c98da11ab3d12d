"""Tiny driver for ncu captures: N opt() steps of the bench workload on a small ring."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from border_b200 import *
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cap = 1 << 16
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
rb.fill_synthetic(cap, 6, 1234)
agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                            soft_update_interval=10000, tau=1.0, batch_size=256, train=True, device=0))
for _ in range(n):
    agent.opt(rb)
print(agent.opt_with_record(rb))
