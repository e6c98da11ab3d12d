"""torchrun worker (needs >= 2 GPUs): data-parallel DQN with the fused P2P all-reduce + Adam.

Checks (SURVEY.md 8e parity definition; there is no reference twin):
  1. identical data on every rank  -> parameters after 3 steps equal the single-GPU run bit for bit
     (the pairwise-tree mean of identical gradients is the gradient itself, exactly, for a power-of-two world);
  2. different data per rank       -> all ranks hold bit-identical parameters, and the applied
     gradient (first Adam moment / 0.1 after one step) is the mean of the ranks' own gradients.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from border_b200 import dist as bd
from border_b200.agents import Dqn, DqnConfig, DqnModelConfig, MlpConfig, OptimizerConfig
from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig


def make(rank, data_seed, sync):
    rng = np.random.default_rng(data_seed)
    n = 256
    obs = rng.standard_normal((n, 4)).astype(np.float32)
    tr = GenericTransitionBatch(obs, rng.integers(0, 2, (n, 1)).astype(np.int64), rng.standard_normal((n, 4)).astype(np.float32),
                                rng.standard_normal(n).astype(np.float32), np.zeros(n, np.int8), np.zeros(n, np.int8))
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=512, seed=7), device=rank)
    rb.push(tr)
    agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=MlpConfig(4, [64, 64], 2), opt_config=OptimizerConfig(lr=1e-3)),
                                soft_update_interval=100, batch_size=64, train=True, device=rank, init_seed=5))
    if sync:
        bd.connect_gradient_peers(agent, dist, torch)
    return rb, agent


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # 1. identical data everywhere vs a single-GPU (unsynchronised) run
    rb, agent = make(local, 123, True)
    rb1, solo = make(local, 123, False)
    for _ in range(3):
        agent.opt(rb)
        solo.opt(rb1)
    pa, ps = agent.named_parameters("qnet"), solo.named_parameters("qnet")
    pow2 = world & (world - 1) == 0  # pairwise-tree mean of identical gradients is exact only then
    for k in pa:
        if pow2:
            assert np.array_equal(pa[k], ps[k]), ("identical-data run diverged from single GPU", k)
        else:
            assert np.allclose(pa[k], ps[k], rtol=1e-5, atol=1e-7), ("identical-data run diverged from single GPU", k)
    # 2. different data per rank
    rb, agent = make(local, 1000 + rank, True)
    rb1, solo = make(local, 1000 + rank, False)
    agent.opt(rb)
    solo.opt(rb1)
    for k, v in agent.named_parameters("qnet").items():
        t = torch.from_numpy(v).cuda()
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        for o in ts:
            assert torch.equal(o, t), ("ranks hold different parameters", k)
        m_sync = agent.opt_state("qnet", k, v.shape)[0] / 0.1
        g_own = torch.from_numpy(solo.opt_state("qnet", k, v.shape)[0] / 0.1).cuda()
        dist.all_reduce(g_own)
        g_mean = (g_own / world).cpu().numpy()
        assert np.allclose(m_sync, g_mean, rtol=1e-5, atol=1e-8), ("applied gradient is not the mean", k)
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
