"""Scratch: where does an env step go?  bb_actor_step vs its pieces (host timing, synchronous calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from border_b200 import *
rng = np.random.default_rng(0)
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=4096, seed=1))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
ag = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)), batch_size=32,
                         train=True, explorer=EpsilonGreedy(), device=0))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
rb.set_stream(ts.cuda_stream); ag.set_stream(ts.cuda_stream)
obs = rng.integers(0, 256, (16, 4, 84, 84), dtype=np.uint8)
def t(fn, n=2000, w=50):
    for i in range(w): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("actor_step            %.1f us" % t(lambda i: ag.actor_step(rb, obs[i % 16], 0.5, 0, 0)))
o8 = np.ascontiguousarray(obs[:8]); r8 = np.ones(8, np.float32); z8 = np.zeros(8, np.int8)
ag.actor_reset()
us8 = t(lambda i: ag.actor_step_n(rb, o8, r8, z8, z8), n=500)
print("actor_step_n (8 envs)  %.1f us per call = %.0f env-steps/s" % (us8, 8e6 / us8))
ag.actor_reset()
print("sample (host explorer) %.1f us" % t(lambda i: ag.sample(obs[i % 16][None])))
tr = GenericTransitionBatch(obs[:1], np.zeros((1, 1), np.int64), obs[1:2], np.ones(1, np.float32), np.zeros(1, np.int8), np.zeros(1, np.int8))
print("host push             %.1f us" % t(lambda i: rb.push(tr)))
x = torch.empty(28240, dtype=torch.uint8, device="cuda"); h = torch.empty(28240, dtype=torch.uint8).pin_memory()
def h2d(i):
    x.copy_(h, non_blocking=True); torch.cuda.current_stream().synchronize()
print("pinned H2D 28 KB + sync %.1f us" % t(h2d))
ev = torch.cuda.Event()
def evs(i):
    ev.record(); ev.synchronize()
print("event record + sync    %.1f us" % t(evs))
