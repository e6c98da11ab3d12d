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
from border_b200.agents import AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, MlpConfig, OptimizerConfig
from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig


KIND = "mlp"


def make(rank, data_seed, sync):
    rng = np.random.default_rng(data_seed)
    n = 256
    if KIND == "cnn":  # the bench's network (AtariCnn on 4x84x84 u8 frames): exercises the early exchange of the FC region
        obs = rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)
        nxt = rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)
        qcfg, n_act, lr, B = AtariCnnConfig(4, 6), 6, 1e-4, 32
    else:
        obs = rng.standard_normal((n, 4)).astype(np.float32)
        nxt = rng.standard_normal((n, 4)).astype(np.float32)
        qcfg, n_act, lr, B = MlpConfig(4, [64, 64], 2), 2, 1e-3, 64
    tr = GenericTransitionBatch(obs, rng.integers(0, n_act, (n, 1)).astype(np.int64), nxt,
                                rng.standard_normal(n).astype(np.float32), np.zeros(n, np.int8), np.zeros(n, np.int8))
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=512, seed=7), device=rank)
    rb.push(tr)
    agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=qcfg, opt_config=OptimizerConfig(lr=lr)),
                                soft_update_interval=100, batch_size=B, train=True, device=rank, init_seed=5))
    if sync:
        bd.connect_gradient_peers(agent, dist, torch)
    return rb, agent


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for kind in ("mlp", "cnn"):
        global KIND
        KIND = kind
        check(rank, world, local)
        if rank == 0:
            print("MGPU_OK world=%d net=%s sync=%s" % (world, kind, os.environ.get("BB_GRAD_SYNC", "default")))
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d" % world)
    dist.destroy_process_group()


def check(rank, world, local):
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
    lr_tol = 1e-5 if KIND == "mlp" else 1e-4  # (the CNN's fp32 GEMMs reorder sums between the split-K choices of one run: none here, same kernels)
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
        assert np.allclose(m_sync, g_mean, rtol=lr_tol, atol=1e-8), ("applied gradient is not the mean", k)
    # 3. keep going (graph replays, epochs advance): ranks stay bit-identical over 20 more steps on different data
    for _ in range(20):
        agent.opt(rb)
    for k, v in agent.named_parameters("qnet").items():
        t = torch.from_numpy(v).cuda()
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        for o in ts:
            assert torch.equal(o, t), ("ranks diverged after 21 steps", k)


if __name__ == "__main__":
    main()
