"""The committed DQN fixture (tests/golden/dqn_mlp_trace.npz, written by tests/golden/make_golden.py from the torch-CPU
oracle of border-tch-agent/src/dqn/base.rs): the oracle must keep reproducing it (CPU), and the device agent must match
it through the C ABI (GPU).  The reference ships no vectors for this path, so this pins the oracle against
itself across torch versions and the device against a file that travels to the GPU box."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dqn_mlp_trace.npz")


def test_oracle_reproduces_the_dqn_fixture():
    from tests.golden import make_golden as mg
    gold = np.load(GOLD)
    out = mg.run_dqn_mlp_trace()
    assert np.array_equal(out["ixs"], gold["ixs"])  # StdRng index stream: integer-exact
    assert np.allclose(out["losses"], gold["losses"], rtol=1e-6, atol=0)
    for k in gold.files:
        if k.startswith("qnet"):
            assert np.allclose(out[k], gold[k], rtol=0, atol=2e-6), k


@pytest.mark.gpu
def test_device_dqn_matches_the_fixture():
    from border_b200.agents import Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, MlpConfig, OptimizerConfig
    from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
    from tests.golden import make_golden as mg
    gold = np.load(GOLD)
    params, tr, cap = mg.dqn_mlp_setup()
    lr, B, steps = 1e-3, 32, len(gold["losses"])
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42))
    dev.push(GenericTransitionBatch(*tr))
    agent = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=MlpConfig(in_dim=4, units=[64, 64], out_dim=2),
                                                            opt_config=OptimizerConfig(lr=lr)),
                                soft_update_interval=2, n_updates_per_opt=1, batch_size=B, discount_factor=0.99, tau=0.5,
                                train=True, explorer=EpsilonGreedy(), double_dqn=True, device=0, critic_loss="SmoothL1"))
    agent.set_parameters("qnet", {k: v.numpy() for k, v in params.items()})
    agent.set_parameters("qnet_tgt", {k: v.numpy() for k, v in params.items()})
    for s in range(steps):
        rec = agent.opt_with_record(dev)
        assert abs(rec["loss"] - gold["losses"][s]) <= 1e-4 * abs(gold["losses"][s]), (s, rec["loss"], gold["losses"][s])
    assert dev.state()["rng_words"] == steps * B  # the same StdRng words were consumed
    for model in ("qnet", "qnet_tgt"):
        got = agent.named_parameters(model)
        for k, v in got.items():
            d = np.abs(v - gold[model + "." + k])
            assert d.max() <= 4.2 * lr, (model, k, d.max())
            assert (d > 2e-2 * lr + 1e-7).mean() <= 2e-3 * steps, (model, k, d.max())


# ---------------------------------------------------------------------------------------------------------------------
# SURVEY 8c G6 / G7 (+ G5 at tau = 0.005): IQN and SAC one-step fixtures with injected tau / z.  The fixtures hold the
# losses, the injected noise and a digest (sum, sum |x|, first 32 values) of every parameter tensor after the updates.
def _check_digest(prefix, got, gold, head_tol, sum_rtol):
    """`got`: name -> np array of a model; compares with the fixture's digest entries `prefix.name.{sum,head}`."""
    for k, v in got.items():
        a = np.asarray(v, np.float64).ravel()
        head, sums = gold["%s.%s.head" % (prefix, k)], gold["%s.%s.sum" % (prefix, k)]
        assert np.abs(a[:32] - head).max() <= head_tol, (prefix, k, np.abs(a[:32] - head).max())
        assert abs(np.abs(a).sum() - sums[1]) <= sum_rtol * sums[1] + 1e-6, (prefix, k)


def test_oracle_reproduces_the_sac_and_iqn_fixtures():
    from tests.golden import make_golden as mg
    for name, run, loss_keys in (("sac", mg.run_sac_trace, ("loss_critic", "loss_actor", "ent_coef")),
                                 ("iqn", mg.run_iqn_trace, ("loss",))):
        gold = np.load(os.path.join(os.path.dirname(GOLD), name + "_trace.npz"))
        out = run()
        for k in loss_keys:
            assert np.allclose(out[k], gold[k], rtol=1e-5, atol=0), (name, k, out[k], gold[k])
        for k in gold.files:
            if k.endswith(".head"):
                assert np.allclose(out[k], gold[k], rtol=0, atol=2e-6), (name, k)


@pytest.mark.gpu
def test_device_sac_matches_the_fixture():
    from border_b200.agents import MlpConfig, OptimizerConfig, Sac, SacConfig
    from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
    from tests.golden import make_golden as mg
    gold = np.load(os.path.join(os.path.dirname(GOLD), "sac_trace.npz"))
    _, tr, pi_p, q_ps = mg.sac_setup()
    lr, B = 3e-4, 64
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=600, seed=42))
    dev.push(GenericTransitionBatch(*tr))
    agent = Sac.build(SacConfig(pi_config=MlpConfig(17, [64, 64], 8), pi_opt_config=OptimizerConfig(lr=lr),
                                q_config=MlpConfig(25, [64, 64], 1), q_opt_config=OptimizerConfig(lr=lr), gamma=0.99, tau=0.005,
                                ent_coef_mode=("Auto", -8.0, 3e-4), batch_size=B, train=True, critic_loss="Mse", reward_scale=1.5,
                                n_critics=2, device=0))
    agent.set_parameters("pi", {k: v.numpy() for k, v in pi_p.items()})
    for i, q in enumerate(q_ps):
        agent.set_parameters("qnet_%d" % i, {k: v.numpy() for k, v in q.items()})
        agent.set_parameters("qnet_tgt_%d" % i, {k: v.numpy() for k, v in q.items()})
    for s in range(len(gold["loss_critic"])):
        agent.inject_noise(0, gold["z1"][s])
        agent.inject_noise(1, gold["z2"][s])
        rec = agent.opt_with_record(dev)
        for k in ("loss_critic", "loss_actor", "ent_coef"):
            assert abs(rec[k] - gold[k][s]) <= 1e-4 * abs(gold[k][s]) + 1e-6, (s, k, rec[k], gold[k][s])
    _check_digest("pi", agent.named_parameters("pi"), gold, 6.3 * lr, 1e-3)
    for i in range(2):
        _check_digest("qnet_%d" % i, agent.named_parameters("qnet_%d" % i), gold, 6.3 * lr, 1e-3)
        _check_digest("qnet_tgt_%d" % i, agent.named_parameters("qnet_tgt_%d" % i), gold, 6.3 * lr * 0.005 * 3 + 1e-7, 1e-4)  # tau = 0.005
    assert abs(float(agent.named_parameters("ent_coef")["log_alpha"][0]) - float(gold["log_alpha"][0])) < 1e-6


@pytest.mark.gpu
def test_device_iqn_matches_the_fixture():
    from border_b200.agents import AtariCnnConfig, EpsilonGreedy, Iqn, IqnConfig, MlpConfig, OptimizerConfig
    from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
    from tests.golden import make_golden as mg
    gold = np.load(os.path.join(os.path.dirname(GOLD), "iqn_trace.npz"))
    _, tr, params = mg.iqn_setup()
    lr, B = 1e-4, 16
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=200, seed=9))
    dev.push(GenericTransitionBatch(*tr))
    agent = Iqn.build(IqnConfig(f_config=AtariCnnConfig(n_stack=4, out_dim=0, skip_linear=True), m_config=MlpConfig(3136, [512], 4),
                                opt_config=OptimizerConfig(lr=lr), feature_dim=3136, embed_dim=64, soft_update_interval=2, batch_size=B,
                                discount_factor=0.99, tau=0.5, train=True, sample_percents_pred="Uniform8", sample_percents_tgt="Uniform8",
                                sample_percents_act="Uniform32", explorer=EpsilonGreedy(eps_start=0.0, eps_final=0.0), device=0))
    agent.set_parameters("iqn", {k: v.numpy() for k, v in params.items()})
    agent.set_parameters("iqn_tgt", {k: v.numpy() for k, v in params.items()})
    for s in range(len(gold["loss"])):
        agent.inject_noise(0, gold["t1"][s])
        agent.inject_noise(1, gold["t2"][s])
        rec = agent.opt_with_record(dev)
        assert abs(rec["loss_critic"] - gold["loss"][s]) <= 1e-4 * abs(gold["loss"][s]) + 1e-7, (s, rec["loss_critic"], gold["loss"][s])
    _check_digest("iqn", agent.named_parameters("iqn"), gold, 4.2 * lr, 1e-3)
    _check_digest("iqn_tgt", agent.named_parameters("iqn_tgt"), gold, 4.2 * lr, 1e-3)
