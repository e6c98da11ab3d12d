"""Scratch: per-kernel CUDA-event breakdown of one IQN / SAC opt step (bench side workloads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from border_b200 import *
kind = sys.argv[1] if len(sys.argv) > 1 else "iqn"
if kind == "iqn":
    cap = 1 << 15
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42, per_config=PerConfig()))
    rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
    rb.fill_synthetic(cap, 4, 1234)
    agent = Iqn.build(IqnConfig(f_config=AtariCnnConfig(n_stack=4, out_dim=0, skip_linear=True), m_config=MlpConfig(3136, [512], 4),
                                opt_config=OptimizerConfig(lr=1e-4), feature_dim=3136, embed_dim=64, soft_update_interval=10000,
                                batch_size=256, tau=1.0, train=True, sample_percents_pred="Uniform64",
                                sample_percents_tgt="Uniform64", device=0))
else:
    cap = 1 << 18
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
    rb.allocate((17,), np.float32, (8,), np.float32)
    rb.fill_synthetic(cap, 0, 99)
    agent = Sac.build(SacConfig(pi_config=MlpConfig(17, [256, 256], 8), q_config=MlpConfig(25, [256, 256], 1), batch_size=512,
                                train=True, n_critics=int(sys.argv[2]) if len(sys.argv) > 2 else 1, device=0))
for _ in range(5):
    agent.opt(rb)
runs = [agent.opt_profiled(rb) for _ in range(6)][2:]
agg, order = {}, []
for run in runs:
    for k, v in run:
        if k not in agg:
            order.append(k)
        agg[k] = agg.get(k, 0.0) + v / len(runs)
tot = sum(agg.values())
print("%s: total %.1f us over %d marks" % (kind, tot * 1e3, len(order)))
for k in order:
    print("%-56s %8.1f us %5.1f%%" % (k, agg[k] * 1e3, 100 * agg[k] / tot))
