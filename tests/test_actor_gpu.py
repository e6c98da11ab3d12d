"""GPU: the device-side actor path (bb_actor_step) against Policy::sample + host push.

Two agents with the same weights and explorer seed walk the same scripted episode stream: one through
bb_agent_sample + bb_replay_push (host transitions, the reference's Sampler::sample_and_push call for call), the other
through bb_actor_step (observation uploaded once, explorer in the kernel tail, push from device-resident copies).  Actions
must be identical step for step and the two rings must hold identical rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200.agents import AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, MlpConfig, OptimizerConfig
from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig


def _walk(kind, explorer, train, steps=60, ep_len=7):
    rng = np.random.default_rng(5)
    if kind == "iqn":
        from border_b200.agents import Iqn, IqnConfig
        shape, dtype, n_act = (4, 84, 84), np.uint8, 4
        frames = rng.integers(0, 256, (steps + 2,) + shape, dtype=np.uint8)
        resets = rng.integers(0, 256, (steps + 2,) + shape, dtype=np.uint8)
    elif kind == "cnn":
        qcfg, shape, dtype, n_act = AtariCnnConfig(4, 6), (4, 84, 84), np.uint8, 6
        frames = rng.integers(0, 256, (steps + 2,) + shape, dtype=np.uint8)
        resets = rng.integers(0, 256, (steps + 2,) + shape, dtype=np.uint8)
    else:
        qcfg, shape, dtype, n_act = MlpConfig(4, [64, 64], 3), (4,), np.float32, 3
        frames = rng.standard_normal((steps + 2,) + shape).astype(np.float32)
        resets = rng.standard_normal((steps + 2,) + shape).astype(np.float32)
    rewards = rng.standard_normal(steps + 2).astype(np.float32)

    def make():
        rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=64, seed=1))
        if kind == "iqn":
            ag = Iqn.build(IqnConfig(f_config=AtariCnnConfig(n_stack=4, out_dim=0, skip_linear=True),
                                     m_config=MlpConfig(3136, [512], n_act), opt_config=OptimizerConfig(lr=1e-4), feature_dim=3136,
                                     embed_dim=64, batch_size=8, train=train, sample_percents_act="Uniform32", explorer=explorer,
                                     device=0, init_seed=3, explorer_seed=77))
            return rb, ag
        ag = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=qcfg, opt_config=OptimizerConfig(lr=1e-3)), batch_size=8,
                                 train=train, explorer=explorer, device=0, init_seed=3, explorer_seed=77))
        return rb, ag

    # reference-shaped loop: act = sample(prev_obs); (next_obs, r, done) = env.step(act); push; prev_obs = reset or next_obs
    rb_a, ag_a = make()
    acts_a = []
    prev = frames[0]
    for t in range(steps):
        a = int(ag_a.sample(prev[None])[0, 0])
        acts_a.append(a)
        nxt, r = frames[t + 1], rewards[t]
        term = int((t + 1) % ep_len == 0)
        rb_a.push(GenericTransitionBatch(prev[None], np.array([[a]], np.int64), nxt[None], np.array([r], np.float32),
                                         np.array([term], np.int8), np.zeros(1, np.int8)))
        prev = resets[t] if term else nxt
    # device actor: the first call only uploads the initial observation and acts on it
    rb_b, ag_b = make()
    rb_b.allocate(shape, dtype, (1,), np.int64)
    acts_b = [ag_b.actor_step(rb_b, frames[0])]
    for t in range(steps):
        term = int((t + 1) % ep_len == 0)
        acts_b.append(ag_b.actor_step(rb_b, frames[t + 1], rewards[t], term, 0, reset_obs=resets[t] if term else None))
    assert acts_a == acts_b[:steps]
    assert len(rb_b) == steps and len(rb_a) == steps
    # same seed, same rows => the same batches
    for _ in range(3):
        ba, bb = rb_a.batch(32), rb_b.batch(32)
        for k in ("obs", "act", "next_obs", "reward", "is_terminated", "is_truncated", "ix_sample"):
            assert np.array_equal(getattr(ba, k), getattr(bb, k)), k
    return acts_a


def test_actor_step_eps_greedy_cnn_matches_sample_and_push():
    acts = _walk("cnn", EpsilonGreedy(eps_start=0.5, eps_final=0.1, final_step=40), True)
    assert len(set(acts)) > 1


def test_actor_step_softmax_and_eval_mlp():
    from border_b200.agents import Softmax
    _walk("mlp", Softmax(), True)
    _walk("mlp", EpsilonGreedy(), False)   # eval mode: 1 % random actions, argmax otherwise


def test_actor_step_iqn_matches_sample_and_push():
    """IQN: quantile mean + IqnExplorer::EpsilonGreedy (iqn/explorer.rs:78-97) through the same device path; the tau draws
    of the acting forward (Uniform32) advance identically on both paths."""
    acts = _walk("iqn", EpsilonGreedy(eps_start=0.4, eps_final=0.1, final_step=30), True, steps=24, ep_len=5)
    assert len(set(acts)) > 1
    _walk("iqn", EpsilonGreedy(), False, steps=10, ep_len=4)


@pytest.mark.parametrize("kind,explorer", [("cnn", EpsilonGreedy(eps_start=0.5, eps_final=0.2, final_step=20)), ("mlp", None)])
def test_actor_step_n_vectorised_envs_match_batched_sample_and_push(kind, explorer):
    """n_envs = 5 per call: the actions equal Policy::sample on the batch of 5 observations (one epsilon draw per call, one
    action draw per process, dqn/explorer.rs:68-90) and the ring holds the 5 transitions of every step in environment order;
    episodes end at different steps per environment (reset observations replace only those rows)."""
    from border_b200.agents import Softmax
    rng = np.random.default_rng(9)
    n, steps = 5, 14
    if kind == "cnn":
        qcfg, shape, dtype, n_act = AtariCnnConfig(4, 6), (4, 84, 84), np.uint8, 6
        frames = rng.integers(0, 256, (steps + 2, n) + shape, dtype=np.uint8)
        resets = rng.integers(0, 256, (steps + 2, n) + shape, dtype=np.uint8)
    else:
        qcfg, shape, dtype, n_act = MlpConfig(4, [64, 64], 3), (4,), np.float32, 3
        frames = rng.standard_normal((steps + 2, n) + shape).astype(np.float32)
        resets = rng.standard_normal((steps + 2, n) + shape).astype(np.float32)
        explorer = Softmax()
    rewards = rng.standard_normal((steps + 2, n)).astype(np.float32)
    done = (rng.random((steps + 2, n)) < 0.25).astype(np.int8)

    def make():
        rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=128, seed=1))
        rb.allocate(shape, dtype, (1,), np.int64)
        ag = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=qcfg, opt_config=OptimizerConfig(lr=1e-3)), batch_size=8,
                                 train=True, explorer=explorer, device=0, init_seed=3, explorer_seed=77))
        return rb, ag

    rb_a, ag_a = make()
    acts_a, prev = [], frames[0].copy()
    for t in range(steps):
        a = ag_a.sample(prev)[:, 0]
        acts_a.append(a.tolist())
        rb_a.push(GenericTransitionBatch(prev, a[:, None].astype(np.int64), frames[t + 1], rewards[t], done[t], np.zeros(n, np.int8)))
        prev = np.where(done[t].reshape((n,) + (1,) * len(shape)) != 0, resets[t], frames[t + 1])
    rb_b, ag_b = make()
    zeros = np.zeros(n, np.int8)
    acts_b = [ag_b.actor_step_n(rb_b, frames[0], np.zeros(n, np.float32), zeros, zeros).tolist()]
    for t in range(steps):
        any_done = bool(done[t].any())
        acts_b.append(ag_b.actor_step_n(rb_b, frames[t + 1], rewards[t], done[t], zeros, reset_obs=resets[t] if any_done else None,
                                        reset_mask=done[t] if any_done else None).tolist())
    assert acts_a == acts_b[:steps]
    assert len(rb_a) == len(rb_b) == steps * n
    for _ in range(2):
        ba, bb = rb_a.batch(48), rb_b.batch(48)
        for k in ("obs", "act", "next_obs", "reward", "is_terminated", "is_truncated", "ix_sample"):
            assert np.array_equal(getattr(ba, k), getattr(bb, k)), k


def test_actor_step_rejects_sac():
    from border_b200 import _lib as L
    from border_b200.agents import Sac, SacConfig
    sac = Sac.build(SacConfig(pi_config=MlpConfig(3, [16], 2), q_config=MlpConfig(5, [16], 1), batch_size=4, device=0))
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=8, seed=1))
    rb.allocate((3,), np.float32, (2,), np.float32)
    with pytest.raises(L.BorderB200Error, match="no device-side actor path"):
        sac.actor_step(rb, np.zeros(3, np.float32))
