"""GPU parity: the device SAC update vs the torch-CPU oracle of border-tch-agent/src/sac/base.rs
(noise z injected on both sides; the reference draws it from libtorch's CPU generator)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from border_b200.agents import MlpConfig, OptimizerConfig, Sac, SacConfig
from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
from oracle import agent_oracle as ao
from oracle import replay_oracle as ro

OBS, ACT = 17, 8


def _setup(n_critics, mode, critic_loss, B, units=(256, 256), lr=3e-4, seed=0):
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    cap = 600
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
    orc = ro.ReplayOracle(cap, 42, (OBS,), np.float32, (ACT,), np.float32)
    n = 580
    obs = rng.standard_normal((n, OBS)).astype(np.float32)
    tr = GenericTransitionBatch(obs, rng.uniform(-1, 1, (n, ACT)).astype(np.float32),
                                rng.standard_normal((n, OBS)).astype(np.float32), rng.standard_normal(n).astype(np.float32),
                                (rng.random(n) < 0.1).astype(np.int8), np.zeros(n, np.int8))
    dev.push(tr)
    orc.push(*tr.unpack()[:6])
    pi_p = ao.mlp2_params(OBS, list(units), ACT, gen)
    q_ps = [ao.mlp_params(OBS + ACT, list(units), 1, gen) for _ in range(n_critics)]
    cfg = SacConfig(pi_config=MlpConfig(OBS, list(units), ACT), pi_opt_config=OptimizerConfig(lr=lr),
                    q_config=MlpConfig(OBS + ACT, list(units), 1), q_opt_config=OptimizerConfig(lr=lr), gamma=0.99, tau=0.005,
                    ent_coef_mode=mode, batch_size=B, train=True, critic_loss=critic_loss, reward_scale=1.5,
                    n_critics=n_critics, device=0)
    agent = Sac.build(cfg)
    agent.set_parameters("pi", {k: v.numpy() for k, v in pi_p.items()})
    for i, q in enumerate(q_ps):
        agent.set_parameters("qnet_%d" % i, {k: v.numpy() for k, v in q.items()})
        agent.set_parameters("qnet_tgt_%d" % i, {k: v.numpy() for k, v in q.items()})
    oracle = ao.SacOracle(pi_p, q_ps, len(units), len(units) + 1, lr, lr, B, 0.99, 0.005, mode, reward_scale=1.5,
                          critic_loss=critic_loss)
    return rng, dev, orc, agent, oracle


def _tb(b):
    return dict(obs=torch.from_numpy(b["obs"]), act=torch.from_numpy(b["act"]), next_obs=torch.from_numpy(b["next_obs"]),
                reward=torch.from_numpy(b["reward"]), is_terminated=torch.from_numpy(b["is_terminated"]))


def _close_params(agent, model, ref, lr):
    got = agent.named_parameters(model)
    for k, v in ref.items():
        d = np.abs(got[k] - v.detach().numpy())
        assert d.max() <= 6.3 * lr, (model, k, d.max())
        assert (d > 0.02 * lr + 1e-7).mean() <= 5e-3, (model, k, (d > 0.02 * lr).mean(), d.max())


@pytest.mark.parametrize("n_critics,mode,critic_loss", [(1, ("Fix", 1.0), "Mse"), (2, ("Auto", -8.0, 3e-4), "Mse"),
                                                        (2, ("Fix", 0.2), "SmoothL1")])
def test_sac_update_parity(n_critics, mode, critic_loss):
    B, lr = 64, 3e-4
    rng, dev, orc, agent, oracle = _setup(n_critics, mode, critic_loss, B, lr=lr)
    for step in range(3):
        z1 = rng.standard_normal((B, ACT)).astype(np.float32)
        z2 = rng.standard_normal((B, ACT)).astype(np.float32)
        agent.inject_noise(0, z1)
        agent.inject_noise(1, z2)
        rec = agent.opt_with_record(dev)
        ref = oracle.opt_(_tb(orc.batch(B)), torch.from_numpy(z1), torch.from_numpy(z2))
        for k in ("loss_critic", "loss_actor", "ent_coef"):
            assert abs(rec[k] - ref[k]) <= 1e-4 * abs(ref[k]) + 1e-6, (step, k, rec[k], ref[k])
        _close_params(agent, "pi", oracle.pi, lr)
        for i in range(n_critics):
            _close_params(agent, "qnet_%d" % i, oracle.qnets[i], lr)
            got = agent.named_parameters("qnet_tgt_%d" % i)
            for k, v in oracle.qnets_tgt[i].items():
                assert np.abs(got[k] - v.numpy()).max() <= 6.3 * lr * 0.005 + 1e-7
        la = agent.named_parameters("ent_coef")["log_alpha"]
        assert abs(float(la[0]) - float(oracle.log_alpha.detach()[0])) < 1e-6


def test_sac_b512_ant_shapes_one_step():
    """BASELINE configs[2]: obs 17, act 8, MLP[256,256], batch 512."""
    B = 512
    rng, dev, orc, agent, oracle = _setup(1, ("Fix", 1.0), "Mse", B)
    z1 = rng.standard_normal((B, ACT)).astype(np.float32)
    z2 = rng.standard_normal((B, ACT)).astype(np.float32)
    agent.inject_noise(0, z1)
    agent.inject_noise(1, z2)
    rec = agent.opt_with_record(dev)
    ref = oracle.opt_(_tb(orc.batch(B)), torch.from_numpy(z1), torch.from_numpy(z2))
    for k in ("loss_critic", "loss_actor"):
        assert abs(rec[k] - ref[k]) <= 1e-4 * abs(ref[k]) + 1e-6, (k, rec[k], ref[k])


def test_sac_policy_sample_and_inkernel_noise():
    rng, dev, orc, agent, oracle = _setup(1, ("Fix", 1.0), "Mse", 32, units=(64, 64))
    obs = rng.standard_normal((4, OBS)).astype(np.float32)
    agent.eval()
    a = agent.sample(obs)
    mean, _ = ao.mlp2_forward(oracle.pi, torch.from_numpy(obs), 2)
    assert a.shape == (4, ACT) and np.allclose(a, np.tanh(mean.detach().numpy()), atol=1e-5)
    agent.train()
    a2 = agent.sample(obs)
    assert np.abs(a2).max() <= 1.0 and not np.allclose(a, a2)
    # without injection the update draws its own N(0,1) noise on the device and stays finite
    rec = agent.opt_with_record(dev)
    assert np.isfinite(rec["loss_critic"]) and np.isfinite(rec["loss_actor"])
    n_opts, blob = agent.model_info()  # SyncModel ships pi only (sac/base.rs:377-386)
    assert blob.size == sum(v.numel() for v in oracle.pi.values())


def test_sac_graph_replay_is_bit_identical_to_eager_launches(monkeypatch):
    """After three eager updates Sac captures the whole update (sample .. soft_update, optimizers included: their
    step-dependent scalars live in device memory) into a CUDA graph; 12 updates with in-kernel noise must leave the same
    parameters as the eager run, bit for bit."""
    outs = []
    stream = torch.cuda.Stream(device=0)
    for graph in ("1", "0"):
        monkeypatch.setenv("BB_GRAPH", graph)
        rng, dev, orc, agent, oracle = _setup(2, ("Auto", -8.0, 3e-4), "Mse", 64)
        dev.set_stream(stream.cuda_stream)
        agent.set_stream(stream.cuda_stream)
        recs = [agent.opt_with_record(dev) for _ in range(12)]
        outs.append((recs, {m: agent.named_parameters(m) for m in ("pi", "qnet_0", "qnet_1", "qnet_tgt_0", "ent_coef")}))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert a["loss_critic"] == b["loss_critic"] and a["loss_actor"] == b["loss_actor"] and a["ent_coef"] == b["ent_coef"]
    for m in outs[0][1]:
        for k in outs[0][1][m]:
            assert np.array_equal(outs[0][1][m][k], outs[1][1][m][k]), (m, k)
