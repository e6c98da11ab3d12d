"""Scratch timing of the two hot calls (not the driver's bench): sample+gather and the DQN opt step."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from border_b200 import *
from border_b200 import _lib as L

cap = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
B = 256
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
t0 = time.time(); rb.fill_synthetic(cap, 6, 1234); torch.cuda.synchronize(); print("fill s", time.time() - t0)
agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                            soft_update_interval=10000, tau=1.0, batch_size=B, train=True, device=0))
ts = torch.cuda.Stream(priority=int(os.environ.get("BB_PRIO", "0"))); torch.cuda.set_stream(ts)
s = ts.cuda_stream
rb.set_stream(s); agent.set_stream(s)
def timeit(fn, n, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: rb.batch_device(B), 2000)
bytes_alg = B * (2 * 28224 + 8 + 4 + 1 + 1)
print("sample+gather: %.2f us  read %.1f GB/s  read+write %.1f GB/s" % (ms * 1e3, bytes_alg / ms / 1e6, 2 * bytes_alg / ms / 1e6))
ms = timeit(lambda: agent.opt(rb), 200, 10)
print("dqn opt step: %.1f us -> %.1f grad-steps/s" % (ms * 1e3, 1e3 / ms))
rec = agent.opt_with_record(rb); print(rec)
obs = np.zeros((1, 4, 84, 84), np.uint8)
t0 = time.time()
for _ in range(500): agent.sample(obs)
print("policy sample: %.1f us" % ((time.time() - t0) / 500 * 1e6))
import ctypes as C
e = C.c_int32(); L.check(L.lib().bb_device_error(C.byref(e), 0)); print("device error flag:", e.value)
