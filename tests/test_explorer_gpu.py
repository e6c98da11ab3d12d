"""GPU: Policy::sample of the device Dqn (dqn/base.rs:211-241) against the explorer oracle (oracle/agent_oracle.py:
EpsilonGreedyOracle over the wyrand restatement): the epsilon schedule, the random / greedy decision stream, the random
actions and the eval branch's 1 % random actions are reproduced draw for draw from the same explorer seed.  The Softmax
explorer draws from libtorch's global generator in the reference (`multinomial`), which no other implementation can
reproduce: it is checked for determinism and against the softmax probabilities instead."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from border_b200.agents import Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, MlpConfig, OptimizerConfig, Softmax
from oracle import agent_oracle as ao


def _agent(explorer, seed, train=True):
    gen = torch.Generator().manual_seed(5)
    params = ao.mlp_params(6, [32, 32], 5, gen)
    cfg = DqnConfig(model_config=DqnModelConfig(q_config=MlpConfig(in_dim=6, units=[32, 32], out_dim=5), opt_config=OptimizerConfig(lr=1e-3)),
                    batch_size=8, train=train, explorer=explorer, device=0, explorer_seed=seed)
    a = Dqn.build(cfg)
    a.set_parameters("qnet", {k: v.numpy() for k, v in params.items()})
    return a, params


@pytest.mark.parametrize("n_procs", [1, 3])
def test_epsilon_greedy_stream_matches_the_oracle(n_procs):
    seed = 1234567
    agent, params = _agent(EpsilonGreedy(eps_start=1.0, eps_final=0.1, final_step=60), seed)
    orc = ao.EpsilonGreedyOracle(ao.FastRandPy(seed), 1.0, 0.1, 60)
    rng = np.random.default_rng(0)
    n_random = 0
    for t in range(150):
        obs = rng.standard_normal((n_procs, 6)).astype(np.float32)
        with torch.no_grad():
            q = ao.mlp_forward(params, torch.from_numpy(obs), 3)
        top2 = q.topk(2, -1).values
        assert float((top2[:, 0] - top2[:, 1]).min()) > 1e-4   # greedy choice not a near tie
        want = orc.action(q)
        got = agent.sample(obs).reshape(-1).tolist()
        assert got == want, (t, got, want)
        n_random += got != [int(x) for x in q.argmax(-1)]
    assert n_random > 10   # both branches were exercised


def test_eval_mode_takes_one_percent_random_actions_like_the_reference():
    seed = 99
    agent, params = _agent(EpsilonGreedy(), seed, train=False)
    orc = ao.EpsilonGreedyOracle(ao.FastRandPy(seed))
    rng = np.random.default_rng(1)
    diffs = 0
    for t in range(600):
        obs = rng.standard_normal((1, 6)).astype(np.float32)
        with torch.no_grad():
            q = ao.mlp_forward(params, torch.from_numpy(obs), 3)
        want = orc.eval_action(q)
        got = agent.sample(obs).reshape(-1).tolist()
        assert got == want, (t, got, want)
        diffs += got[0] != int(q.argmax(-1))
    assert diffs <= 30


def test_softmax_explorer_is_deterministic_and_follows_the_softmax_probabilities():
    a1, params = _agent(Softmax(), 7)
    a2, _ = _agent(Softmax(), 7)
    obs = np.random.default_rng(2).standard_normal((1, 6)).astype(np.float32)
    with torch.no_grad():
        p = torch.softmax(ao.mlp_forward(params, torch.from_numpy(obs), 3), -1).numpy()[0]
    n = 4000
    s1 = [int(a1.sample(obs)[0, 0]) for _ in range(n)]
    s2 = [int(a2.sample(obs)[0, 0]) for _ in range(n)]
    assert s1 == s2
    freq = np.bincount(s1, minlength=5) / n
    assert np.abs(freq - p).max() < 4.0 * np.sqrt(0.25 / n)
