"""CPU: the transitions the C++ mirror of `Sampler::sample_and_push` + `SimpleStepProcessor::process` emits, against a Python
restatement of border-core/src/trainer/sampler.rs:99-144, generic_replay_buffer/step_proc.rs:103-137 and
base/env.rs:138-160 (step_with_reset) over the same synthetic environment: obs / next_obs identity of every transition
incl. the prev_obs <- init_obs swap at episode ends (terminated and truncated), actions, rewards, flags."""
import numpy as np
import pytest

M64 = (1 << 64) - 1


def mix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class Env:
    """SyntheticEnv of border_b200/host/host_capi.cpp: obs = (episode, word(episode, t))."""

    def __init__(self, seed, episode_len, truncate_len):
        self.seed, self.el, self.tl = seed, episode_len, truncate_len
        self.episode = self.t = self.steps = 0

    def _obs(self):
        return (self.episode, mix64(mix64(self.seed ^ ((self.episode << 32) & M64) ^ self.t)))

    def reset(self):
        self.episode += 1
        self.t = 0
        return self._obs()

    def step_with_reset(self, act):  # env.rs:138-160
        self.t += 1
        self.steps += 1
        obs = self._obs()
        r = mix64(self.seed ^ ((self.steps * 0x100000001B3) & M64)) % 100
        reward = -1 if r < 5 else (0 if r < 95 else 1)
        term = 1 if (self.el and self.t % self.el == 0) else 0
        trunc = 1 if (not term and self.tl and self.t % self.tl == 0) else 0
        init_obs = self.reset() if (term or trunc) else None
        return dict(act=act, obs=obs, reward=reward, term=term, trunc=trunc, init_obs=init_obs)


def reference_trace(seed, episode_len, truncate_len, n):
    env = Env(seed, episode_len, truncate_len)
    out = []
    sampler_prev = None      # Sampler.prev_obs (sampler.rs:109-115)
    proc_prev = None         # SimpleStepProcessor.prev_obs
    for i in range(n):
        if sampler_prev is None:
            sampler_prev = env.reset()
            proc_prev = sampler_prev                      # step_proc.reset(init_obs)
        step = env.step_with_reset(i)                     # act = agent.sample(prev_obs) = i
        done = step["term"] or step["trunc"]
        sampler_prev = step["init_obs"] if done else step["obs"]   # sampler.rs:126-129
        # step_proc.rs:111-123: obs = prev_obs.replace(step.obs); on done prev_obs <- init_obs
        obs, proc_prev = proc_prev, step["obs"]
        if done:
            proc_prev = step["init_obs"]
        out.append((obs, step["obs"], i, step["reward"], step["term"], step["trunc"]))
        if done:
            proc_prev = sampler_prev                      # sampler.rs:138-141: step_proc.reset(prev_obs)
    return out


@pytest.mark.parametrize("episode_len,truncate_len", [(7, 0), (0, 5), (11, 4), (0, 0)])
def test_emitted_transitions_match_the_reference_loop(episode_len, truncate_len):
    from border_b200 import host_loops as hl
    n, seed = 60, 12345
    got = hl.sampler_trace(4 * 84 * 84, episode_len, truncate_len, seed, n)
    ref = reference_trace(seed, episode_len, truncate_len, n)
    ends = 0
    for i, (obs, nxt, act, r, term, trunc) in enumerate(ref):
        row = [int(x) & M64 for x in got[i][:4]]
        assert row == [obs[0], obs[1], nxt[0], nxt[1]], i
        assert [int(x) for x in got[i][4:]] == [act, r, term, trunc], i
        ends += term or trunc
    if episode_len or truncate_len:
        assert ends >= 4
        # the transition after an episode end starts from the NEW episode's initial observation
        k = next(i for i, t in enumerate(ref) if t[4] or t[5])
        assert ref[k + 1][0][0] == ref[k][1][0] + 1 and int(got[k + 1][0]) == ref[k + 1][0][0]
