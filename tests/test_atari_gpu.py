"""GPU: the device Atari observation pipeline (bb_atari_*) bit for bit against oracle/atari_oracle.py
(border-atari-env/src/env.rs:126-199, 263-300; the resize restates image 0.23.14 -- see the oracle's header: unpinned)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200 import AtariPreprocessor
from oracle import atari_oracle as ato


def _frames(rng, n, h=210, w=160):
    # smooth structure + noise + flat regions, so that rounding ties and saturation both occur
    base = rng.integers(0, 256, (n, h // 10 + 1, w // 10 + 1, 3)).repeat(10, 1).repeat(10, 2)[:, :h, :w]
    noise = rng.integers(-20, 21, (n, h, w, 3))
    out = np.clip(base + noise, 0, 255).astype(np.uint8)
    out[:, :20] = 255
    out[:, -15:] = 0
    return out


def test_reset_and_steps_match_the_oracle_bit_for_bit():
    rng = np.random.default_rng(0)
    fr = _frames(rng, 13)
    dev = AtariPreprocessor(train=True)
    orc = ato.FrameStack()
    dev.reset(fr[0])
    assert np.array_equal(dev.obs(), orc.reset(fr[0]))
    for t in range(6):
        a, b = fr[1 + 2 * t], fr[2 + 2 * t]
        r = float(rng.choice([-3.0, 0.0, 0.5, 7.0]))
        assert dev.step(a, b, r) == ato.clip_reward(r, True)
        got, want = dev.obs(), orc.step(a, b)
        assert np.array_equal(got, want), (t, int(np.abs(got.astype(int) - want.astype(int)).max()))
    dev.close()


def test_eval_mode_keeps_the_reward_and_other_frame_sizes_work():
    rng = np.random.default_rng(1)
    dev = AtariPreprocessor(width=200, height=250, train=False)
    f = _frames(rng, 2, 250, 200)
    dev.reset(f[0])
    assert dev.step(f[0], f[1], -2.5) == -2.5
    want = ato.warp_and_grayscale(np.maximum(f[0], f[1]))
    assert np.array_equal(dev.obs()[0], want)
    assert np.array_equal(dev.obs()[1], ato.warp_and_grayscale(f[0]))
    dev.close()


def test_frame_stack_feeds_the_device_actor_without_leaving_hbm():
    """AtariPreprocessor.obs_device() -> Dqn.actor_step_dev: same actions and ring rows as the host path fed with the
    oracle's observations."""
    from border_b200.agents import AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, OptimizerConfig
    from border_b200.replay import SimpleReplayBuffer, SimpleReplayBufferConfig
    rng = np.random.default_rng(2)
    fr = _frames(rng, 21)

    def make():
        rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=32, seed=1))
        rb.allocate((4, 84, 84), np.uint8, (1,), np.int64)
        ag = Dqn.build(DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                                 batch_size=8, train=True, explorer=EpsilonGreedy(eps_start=0.3, eps_final=0.3), device=0,
                                 init_seed=3, explorer_seed=5))
        return rb, ag

    import torch
    rb_h, ag_h = make()
    rb_d, ag_d = make()
    orc, dev = ato.FrameStack(), AtariPreprocessor(train=True)
    stream = torch.cuda.Stream(device=0)   # preprocessor, policy and ring share one stream: no cross-stream waits needed
    for o in (dev, ag_d, rb_d):
        o.set_stream(stream.cuda_stream)
    acts_h = [ag_h.actor_step(rb_h, orc.reset(fr[0]))]
    dev.reset(fr[0])
    acts_d = [ag_d.actor_step_dev(rb_d, dev.obs_device(stream.cuda_stream))]
    for t in range(10):
        a, b = fr[1 + 2 * t], fr[2 + 2 * t]
        r = float(t % 3 - 1)
        acts_h.append(ag_h.actor_step(rb_h, orc.step(a, b), ato.clip_reward(r, True), 0, 0))
        rc = dev.step(a, b, r)
        acts_d.append(ag_d.actor_step_dev(rb_d, dev.obs_device(stream.cuda_stream), rc, 0, 0))
    assert acts_h == acts_d
    bh, bd = rb_h.batch(16), rb_d.batch(16)
    for k in ("obs", "act", "next_obs", "reward", "is_terminated"):
        assert np.array_equal(getattr(bh, k), getattr(bd, k)), k
