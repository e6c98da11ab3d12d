"""Scratch: full per-kernel CUDA-event breakdown of one DQN opt step (bench workload, small ring)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from border_b200 import *
cap = 1 << 16
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
rb.fill_synthetic(cap, 6, 1234)
agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                            soft_update_interval=10000, tau=1.0, batch_size=256, train=True, device=0))
for _ in range(5):
    agent.opt(rb)
runs = [agent.opt_profiled(rb) for _ in range(8)][2:]
agg, order = {}, []
for run in runs:
    for k, v in run:
        if k not in agg:
            order.append(k)
        agg[k] = agg.get(k, 0.0) + v / len(runs)
tot = sum(agg.values())
print("total %.1f us over %d kernels" % (tot * 1e3, len(order)))
for k in order:
    print("%-50s %7.1f us %5.1f%%" % (k, agg[k] * 1e3, 100 * agg[k] / tot))
groups = {}
for k, v in agg.items():
    g = k.split(":")[-1]
    groups[g] = groups.get(g, 0) + v
print("--- by kernel")
for g, v in sorted(groups.items(), key=lambda kv: -kv[1]):
    print("%-30s %7.1f us %5.1f%%" % (g, v * 1e3, 100 * v / tot))
