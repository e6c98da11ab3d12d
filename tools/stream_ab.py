"""Scratch: DQN opt step time vs which stream the agent runs on (legacy default / torch stream / library stream) and BB_GRAPH."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from border_b200 import *
mode = sys.argv[1]
cap = 65536
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
rb.fill_synthetic(cap, 6, 1234)
agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                            soft_update_interval=10000, tau=1.0, batch_size=256, train=True, device=0))
if mode == "legacy":
    s = torch.cuda.current_stream().cuda_stream
    rb.set_stream(s); agent.set_stream(s)
elif mode == "torch":
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    rb.set_stream(ts.cuda_stream); agent.set_stream(ts.cuda_stream)
elif mode == "torch_hi":
    ts = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(ts)
    rb.set_stream(ts.cuda_stream); agent.set_stream(ts.cuda_stream)
# "lib": the library's own non-blocking stream
for _ in range(20): agent.opt(rb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): agent.opt(rb)
t1 = time.perf_counter()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 300
print("mode %-8s BB_GRAPH=%s BB_SERIAL=%s: %.1f us/step total, %.1f us/step host issue" % (mode, os.environ.get("BB_GRAPH", "1"), os.environ.get("BB_SERIAL", "0"), dt * 1e6, (t1 - t0) / 300 * 1e6))
t0 = time.perf_counter()
for _ in range(2000): rb.batch_device(256)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("   gather: %.2f us total, %.2f us host issue" % ((time.perf_counter() - t0) / 2000 * 1e6, (t1 - t0) / 2000 * 1e6))
