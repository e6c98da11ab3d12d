"""Scratch: timings of the prioritized-replay calls and of a DQN step on a PER ring."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from border_b200 import *
cap = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 18)
B = 256
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42, per_config=PerConfig()))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
rb.fill_synthetic(cap, 6, 1234)
agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                            soft_update_interval=10000, tau=1.0, batch_size=B, train=True, device=0))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
rb.set_stream(ts.cuda_stream); agent.set_stream(ts.cuda_stream)
def timeit(fn, n, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("PER sample+gather: %.1f us" % timeit(lambda: rb.batch_device(B), 500))
ix = np.random.default_rng(0).integers(0, cap, B).astype(np.uint64)
td = np.random.default_rng(1).random(B).astype(np.float32)
print("update_priority (host ixs/td): %.1f us" % timeit(lambda: rb.update_priority(ix, td), 200))
print("DQN+PER opt step: %.1f us" % timeit(lambda: agent.opt(rb), 200, 10))
runs = [agent.opt_profiled(rb) for _ in range(6)][2:]
agg = {}
for run in runs:
    for k, v in run: agg[k] = agg.get(k, 0.0) + v / len(runs)
for k, v in agg.items():
    if "replay" in k or "adam" in k or "loss" in k: print("  %-40s %.1f us" % (k, v * 1e3))
tr = GenericTransitionBatch(np.zeros((1, 4, 84, 84), np.uint8), np.zeros((1, 1), np.int64), np.zeros((1, 4, 84, 84), np.uint8),
                            np.ones(1, np.float32), np.zeros(1, np.int8), np.zeros(1, np.int8))
t0 = time.perf_counter()
for _ in range(500): rb.push(tr)
torch.cuda.synchronize()
print("PER push(1 host transition): %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
