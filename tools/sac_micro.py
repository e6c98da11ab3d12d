"""Scratch: SAC Ant-like B=512 update step time (graph and eager)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from border_b200 import *
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=1 << 16, seed=42))
rb.allocate((17,), np.float32, (8,), np.float32)
rb.fill_synthetic(1 << 16, 0, 99)
rb.set_stream(ts.cuda_stream)
for nc in (1, 2):
    sac = Sac.build(SacConfig(pi_config=MlpConfig(17, [256, 256], 8), q_config=MlpConfig(25, [256, 256], 1), batch_size=512, train=True, n_critics=nc, device=0))
    sac.set_stream(ts.cuda_stream)
    for _ in range(30): sac.opt(rb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500): sac.opt(rb)
    e1.record(); torch.cuda.synchronize()
    print("sac critics=%d: %.1f us/step" % (nc, e0.elapsed_time(e1) / 500 * 1e3))
