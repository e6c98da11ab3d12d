import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from border_b200 import *
cap = 1 << 14
rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42, per_config=PerConfig()))
rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
rb.fill_synthetic(cap, 6, 1234)
ix = np.random.default_rng(0).integers(0, cap, 256).astype(np.uint64)
td = np.random.default_rng(1).random(256).astype(np.float32)
for _ in range(4): rb.update_priority(ix, td)
print(rb.state())
