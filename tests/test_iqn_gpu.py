"""GPU parity: the device IQN update vs the torch-CPU oracle of border-tch-agent/src/iqn/base.rs
(percent points tau / tau' injected on both sides)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L
from border_b200.agents import AtariCnnConfig, EpsilonGreedy, Iqn, IqnConfig, MlpConfig, OptimizerConfig
from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
from oracle import agent_oracle as ao
from oracle import replay_oracle as ro


def _setup(kind, B, n_act, lr, sample="Uniform8", per=None, cap=200, n=180):
    rng = np.random.default_rng(11)
    gen = torch.Generator().manual_seed(3)
    # feature extractor = AtariCnn.skip_linear, merge net = Mlp(3136 -> 512 -> A): the shapes of
    # BASELINE configs[3] (an Mlp extractor would collide with the merge net's `mlp.ln*` names in
    # the reference's single VarStore)
    assert kind == "cnn"
    obs_shape, obs_dtype, F, E = (4, 84, 84), np.uint8, 3136, 64
    f_params = ao.atari_cnn_params(4, 0, gen, skip_linear=True)
    psi_fn = lambda p, x: ao.atari_cnn_forward(p, x, skip_linear=True)
    f_cfg = AtariCnnConfig(n_stack=4, out_dim=0, skip_linear=True)
    m_units = [512]
    m_params = ao.mlp_params(F, m_units, n_act, gen)
    params = ao.iqn_params(f_params, F, E, m_params, gen)
    m_fn = lambda p, m: ao.mlp_forward(p, m, len(m_units) + 1)
    from border_b200.replay import PerConfig
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=9, per_config=PerConfig(**per) if per else None))
    orc = ro.ReplayOracle(cap, 9, obs_shape, obs_dtype, (1,), np.int64, per=per)
    obs = rng.integers(0, 256, (n,) + obs_shape, dtype=np.uint8)
    nxt = rng.integers(0, 256, (n,) + obs_shape, dtype=np.uint8)
    tr = GenericTransitionBatch(obs, rng.integers(0, n_act, (n, 1)).astype(np.int64), nxt,
                                rng.standard_normal(n).astype(np.float32), (rng.random(n) < 0.2).astype(np.int8),
                                np.zeros(n, np.int8))
    dev.push(tr)
    orc.push(*tr.unpack()[:6])
    cfg = IqnConfig(f_config=f_cfg, m_config=MlpConfig(F, m_units, n_act), opt_config=OptimizerConfig(lr=lr), feature_dim=F,
                    embed_dim=E, soft_update_interval=2, batch_size=B, discount_factor=0.99, tau=0.5, train=True,
                    sample_percents_pred=sample, sample_percents_tgt=sample, sample_percents_act="Uniform32",
                    explorer=EpsilonGreedy(eps_start=0.0, eps_final=0.0), device=0)
    agent = Iqn.build(cfg)
    agent.set_parameters("iqn", {k: v.numpy() for k, v in params.items()})
    agent.set_parameters("iqn_tgt", {k: v.numpy() for k, v in params.items()})
    oracle = ao.IqnOracle(params, psi_fn, m_fn, E, lr, B, 0.99, 0.5, 2)
    return rng, dev, orc, agent, oracle, params, psi_fn, m_fn, E


def _tb(b):
    return dict(obs=torch.from_numpy(b["obs"]), act=torch.from_numpy(b["act"]), next_obs=torch.from_numpy(b["next_obs"]),
                reward=torch.from_numpy(b["reward"]), is_terminated=torch.from_numpy(b["is_terminated"]))


def _close(agent, model, ref, lr):
    got = agent.named_parameters(model)
    for k, v in ref.items():
        d = np.abs(got[k] - v.detach().numpy())
        assert d.max() <= 4.2 * lr, (k, d.max())
        assert (d > 0.02 * lr + 1e-7).mean() <= 5e-3, (k, (d > 0.02 * lr).mean(), d.max())


@pytest.mark.parametrize("sample,N", [("Uniform8", 8), ("Uniform64", 64)])
def test_iqn_atari_update_parity(sample, N):
    B, lr = 16, 1e-4
    rng, dev, orc, agent, oracle, *_ = _setup("cnn", B, 4, lr, sample)
    for step in range(2):
        t1 = rng.random((B, N), dtype=np.float32)
        t2 = rng.random((B, N), dtype=np.float32)
        agent.inject_noise(0, t1)
        agent.inject_noise(1, t2)
        rec = agent.opt_with_record(dev)
        ref = oracle.opt_(_tb(orc.batch(B)), torch.from_numpy(t1), torch.from_numpy(t2))
        assert abs(rec["loss_critic"] - ref) <= 1e-4 * abs(ref) + 1e-7, (step, rec["loss_critic"], ref)
        _close(agent, "iqn", oracle.iqn, lr)
        _close(agent, "iqn_tgt", oracle.iqn_tgt, lr)


def test_iqn_baseline_shape_b256_n64_on_a_prioritized_ring():
    """BASELINE configs[3] as the bench runs it: B = 256, N = N' = 64, prioritized replay (sampling + IS weights; the
    reference's IQN never updates priorities): rows drawn identical to the oracle's, loss within 1e-4."""
    B, N, lr = 256, 64, 1e-4
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=500000, normalize="All")
    rng, dev, orc, agent, oracle, *_ = _setup("cnn", B, 4, lr, "Uniform64", per=per, cap=512, n=400)
    t1 = rng.random((B, N), dtype=np.float32)
    t2 = rng.random((B, N), dtype=np.float32)
    u = rng.random(B, dtype=np.float32)
    agent.inject_noise(0, t1)
    agent.inject_noise(1, t2)
    dev.inject_uniforms(u)
    rec = agent.opt_with_record(dev)
    b = orc.batch(B, u)
    assert np.array_equal(dev.last_indices(), np.asarray(b["ix_sample"], dtype=np.uint64))
    ref = oracle.opt_(_tb(b), torch.from_numpy(t1), torch.from_numpy(t2))
    assert abs(rec["loss_critic"] - ref) <= 1e-4 * abs(ref) + 1e-7, (rec["loss_critic"], ref)
    _close(agent, "iqn", oracle.iqn, lr)


def test_iqn_const_modes_policy_and_errors():
    B, lr = 8, 1e-4
    rng, dev, orc, agent, oracle, params, psi_fn, m_fn, E = _setup("cnn", B, 4, lr, "Const10")
    # Const10 needs no injection: tau is the fixed grid 0.05..0.95 on both sides
    tau = torch.tensor([0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75, 0.85, 0.95]).unsqueeze(0).repeat(B, 1)
    rec = agent.opt_with_record(dev)
    ref = oracle.opt_(_tb(orc.batch(B)), tau, tau)
    assert abs(rec["loss_critic"] - ref) <= 1e-4 * abs(ref) + 1e-7
    # greedy policy (eps = 0) with the deterministic Const10 grid for `sample_percents_act`:
    # argmax of the tau-averaged action values (average(), iqn/model/base.rs:394-418)
    cfg = agent.config
    cfg.sample_percents_act = "Const10"
    pol = Iqn.build(cfg)
    pol.sync_model(agent.model_info()[1])
    for k, v in pol.named_parameters("iqn").items():
        assert np.array_equal(v, agent.named_parameters("iqn")[k]), k
    checked = 0
    for _ in range(6):
        obs = rng.integers(0, 256, (1, 4, 84, 84), dtype=np.uint8)
        q = ao.iqn_forward(oracle.iqn, psi_fn, m_fn, torch.from_numpy(obs), tau[:1], E).mean(1).detach().numpy()[0]
        top2 = np.sort(q)[-2:]
        a = int(pol.sample(obs)[0, 0])
        if top2[1] - top2[0] > 1e-3 * (np.abs(q).max() + 1e-6):
            assert a == int(q.argmax()), (a, q)
            checked += 1
    assert checked >= 1
    pol.eval()
    assert pol.sample(obs).shape == (1, 1)
    assert agent.sample(obs).shape == (1, 1)  # Uniform32 draws on the device
    # Const32 for pred/tgt panics in the reference (33 points vs 32): same here, as an error
    cfg = agent.config
    cfg.sample_percents_pred = "Const32"
    bad = Iqn.build(cfg)
    with pytest.raises(L.BorderB200Error, match="Const32"):
        bad.opt(dev)


def test_iqn_graph_replay_is_bit_identical_to_eager_launches(monkeypatch):
    """Iqn captures sample .. backward into a CUDA graph after three eager updates (uniform ring, in-kernel tau draws read
    their counters from device memory); the parameters after 8 updates equal the eager run's bit for bit."""
    outs = []
    stream = torch.cuda.Stream(device=0)
    for graph in ("1", "0"):
        monkeypatch.setenv("BB_GRAPH", graph)
        rng, dev, orc, agent, oracle, params, psi_fn, m_fn, E = _setup("cnn", 16, 4, 1e-4)
        dev.set_stream(stream.cuda_stream)
        agent.set_stream(stream.cuda_stream)
        losses = [agent.opt_with_record(dev)["loss_critic"] for _ in range(8)]
        outs.append((losses, agent.named_parameters("iqn"), agent.named_parameters("iqn_tgt")))
    assert outs[0][0] == outs[1][0]
    for k in outs[0][1]:
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k
        assert np.array_equal(outs[0][2][k], outs[1][2][k]), k
