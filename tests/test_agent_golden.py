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
