"""GPU: border's Trainer / train_async loop semantics (C++ mirror over the C ABI) with a synthetic env."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L
from border_b200 import host_loops as H
from border_b200.agents import (AtariCnnConfig, DqnConfig, DqnModelConfig, EpsilonGreedy, MlpConfig, OptimizerConfig,
                                SacConfig)


def _replay_cfg(capacity, obs_kind, obs_elems, act_kind, act_elems, per=False):
    c = L.bb_replay_cfg()
    L.lib().bb_replay_cfg_default(C.byref(c))
    c.capacity, c.seed, c.per_config_some = capacity, 42, int(per)
    c.n_opts_final = 1000
    c.obs_kind, c.obs_elems, c.act_kind, c.act_elems, c.device = obs_kind, obs_elems, act_kind, act_elems, 0
    return c


def _dqn_mlp(batch):
    return DqnConfig(model_config=DqnModelConfig(q_config=MlpConfig(4, [64, 64], 2), opt_config=OptimizerConfig(lr=1e-3)),
                     soft_update_interval=10, batch_size=batch, train=True, explorer=EpsilonGreedy(final_step=1000),
                     device=0).to_c()


def test_trainer_cadence_matches_reference():
    """trainer.rs:197-228: no opt before warmup_period env steps, then one every opt_interval."""
    env = H.bbh_env_cfg(L.BB_F32, 4, 17, 0)
    tc = H.trainer_cfg(max_opts=50, opt_interval=3, warmup_period=100, record_agent_info_interval=10, env_seed=1)
    st = H.train("dqn", _dqn_mlp(32), _replay_cfg(10000, L.BB_F32, 4, L.BB_I64, 1), env, tc)
    assert st["opt_steps"] == 50 and st["agent_n_opts"] == 50
    # first opt at the first env step >= 100 divisible by 3 (= 102), then every 3 steps
    assert st["env_steps"] == 102 + 3 * 49
    assert st["buffer_len"] == st["env_steps"]          # one transition pushed per env step
    assert st["records"] == 5                            # (opt_steps + 1) % 10 == 0
    assert np.isfinite(st["last_loss"]) and st["last_loss"] > 0


def test_trainer_ring_wraps_and_saves(tmp_path):
    env = H.bbh_env_cfg(L.BB_F32, 4, 5, 0)
    tc = H.trainer_cfg(max_opts=40, opt_interval=1, warmup_period=64, save_interval=20, env_seed=2)
    st = H.train("dqn", _dqn_mlp(16), _replay_cfg(50, L.BB_F32, 4, L.BB_I64, 1, per=True), env, tc, str(tmp_path))
    assert st["buffer_len"] == 50 and st["env_steps"] == 64 + 39
    assert st["saves"] == 2
    assert os.path.exists(tmp_path / "20" / "qnet.pt.tch.b200") and os.path.exists(tmp_path / "40" / "qnet_tgt.pt.tch.b200")


def test_trainer_atari_dqn_and_sac():
    cfg = DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                    soft_update_interval=10, batch_size=32, train=True, explorer=EpsilonGreedy(), device=0).to_c()
    env = H.bbh_env_cfg(L.BB_U8, 4 * 84 * 84, 50, 0)
    st = H.train("dqn", cfg, _replay_cfg(512, L.BB_U8, 4 * 84 * 84, L.BB_I64, 1), env,
                 H.trainer_cfg(max_opts=20, warmup_period=32, record_agent_info_interval=20))
    assert st["opt_steps"] == 20 and st["env_steps"] == 51 and np.isfinite(st["last_loss"])
    sac = SacConfig(pi_config=MlpConfig(17, [64, 64], 8), q_config=MlpConfig(25, [64, 64], 1), batch_size=32, train=True,
                    n_critics=2, device=0).to_c()
    env = H.bbh_env_cfg(L.BB_F32, 17, 40, 0)
    st = H.train("sac", sac, _replay_cfg(1000, L.BB_F32, 17, L.BB_F32, 8), env,
                 H.trainer_cfg(max_opts=15, warmup_period=40, record_agent_info_interval=15))
    assert st["opt_steps"] == 15 and np.isfinite(st["last_loss"])


def test_train_async_actors_feed_the_learner():
    """util.rs:31-92: actors push through ReplayBufferProxy in bulks of n_buffer; the learner warms
    up on buffer.len() (async_trainer/base.rs:205), syncs the model every sync_interval."""
    env = H.bbh_env_cfg(L.BB_F32, 4, 23, 0)
    tc = H.trainer_cfg(max_opts=60, warmup_period=200, sync_interval=5, n_actors=3, n_buffer=25,
                       record_agent_info_interval=60)
    st = H.train_async("dqn", _dqn_mlp(32), _replay_cfg(5000, L.BB_F32, 4, L.BB_I64, 1), env, tc)
    assert st["opt_steps"] == 60
    assert st["samples_total"] >= 200 and st["samples_total"] % 25 == 0   # bulks of n_buffer
    assert st["buffer_len"] == min(5000, st["samples_total"])
    assert st["env_steps"] >= st["samples_total"]
    assert st["syncs"] >= 1 + 60 // 5
    assert st["samples_per_sec"] > 0 and st["opt_per_sec"] > 0 and np.isfinite(st["last_loss"])


def test_host_errors_surface():
    env = H.bbh_env_cfg(L.BB_F32, 5, 10, 0)  # obs_elems 5 != network in_dim 4
    with pytest.raises(L.BorderB200Error):
        H.train("dqn", _dqn_mlp(8), _replay_cfg(100, L.BB_F32, 5, L.BB_I64, 1), env, H.trainer_cfg(max_opts=2, warmup_period=8))


def test_train_async_learner_hook_phases():
    """bbh_train_async_ex: the hook sees the learner's handle after creation (phase 0: where a data-parallel job connects
    its gradient peers) and after the last update (phase 1), on the calling thread."""
    env = H.bbh_env_cfg(L.BB_F32, 4, 23, 0)
    tc = H.trainer_cfg(max_opts=20, warmup_period=64, sync_interval=5, n_actors=2, n_buffer=8, record_agent_info_interval=20)
    seen = []

    def hook(handle, phase):
        n = C.c_uint64()
        L.check(L.lib().bb_agent_n_opts(handle, C.byref(n)))
        seen.append((phase, n.value))

    st = H.train_async("dqn", _dqn_mlp(16), _replay_cfg(1000, L.BB_F32, 4, L.BB_I64, 1), env, tc, on_learner=hook)
    assert st["opt_steps"] == 20
    assert seen == [(0, 0), (2, 0), (1, 20)]


def test_train_offline_runs_max_opts_on_a_prefilled_dataset(tmp_path):
    """trainer.rs:330-384: no sampling, one optimisation step per loop trip (warmup 0, opt_interval 1), env_steps counts the
    trips, the dataset is untouched, records / saves keep the trainer's cadence."""
    from border_b200.replay import GenericTransitionBatch, SimpleReplayBuffer, SimpleReplayBufferConfig
    rng = np.random.default_rng(0)
    n = 400
    obs = rng.standard_normal((n, 4)).astype(np.float32)
    ds = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=512, seed=3))
    ds.push(GenericTransitionBatch(obs, rng.integers(0, 2, (n, 1)).astype(np.int64), obs + 0.1, rng.standard_normal(n).astype(np.float32),
                                   np.zeros(n, np.int8), np.zeros(n, np.int8)))
    tc = H.trainer_cfg(max_opts=30, opt_interval=7, warmup_period=1000, record_agent_info_interval=10, save_interval=15)
    st = H.train_offline("dqn", _dqn_mlp(32), ds, tc, str(tmp_path))
    assert st["opt_steps"] == 30 and st["agent_n_opts"] == 30 and st["env_steps"] == 30   # warmup / opt_interval are overridden
    assert st["buffer_len"] == n and len(ds) == n
    assert st["records"] == 3 and st["saves"] == 2 and np.isfinite(st["last_loss"]) and st["last_loss"] > 0
    assert ds.state()["rng_words"] == 30 * 32                                             # 30 batches were drawn from the dataset
