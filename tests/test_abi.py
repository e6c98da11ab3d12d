"""CPU: the C-ABI library loads and exports every symbol include/border_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from border_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "border_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    hdr = _header_symbols()
    assert len(hdr) >= 40
    for name in hdr:
        assert hasattr(lib, name), name
    assert sorted(hdr) == L.declared_symbols()


def test_abi_version_and_defaults_without_gpu():
    lib = L.lib()
    assert lib.bb_abi_version() == 1
    c = L.bb_replay_cfg()
    lib.bb_replay_cfg_default(C.byref(c))
    # SimpleReplayBufferConfig::default (config.rs:199-207) and PerConfig::default (:23-33)
    assert (c.capacity, c.seed, c.per_config_some) == (10000, 42, 0)
    assert abs(c.alpha - 0.6) < 1e-7 and abs(c.beta_0 - 0.4) < 1e-7 and c.beta_final == 1.0 and c.n_opts_final == 500000
    d = L.bb_dqn_cfg()
    lib.bb_dqn_cfg_default(C.byref(d))
    # DqnConfig::default (dqn/config.rs:82-102)
    assert (d.soft_update_interval, d.n_updates_per_opt, d.batch_size) == (1, 1, 1)
    assert d.discount_factor == 0.99 and d.tau == 0.005 and d.train == 0 and d.double_dqn == 0
    assert d.explorer == L.BB_EXPLORER_SOFTMAX and d.critic_loss == L.BB_LOSS_MSE
    s = L.bb_sac_cfg()
    lib.bb_sac_cfg_default(C.byref(s))
    # SacConfig::default (sac/config.rs:85-105)
    assert s.gamma == 0.99 and s.tau == 0.005 and s.epsilon == 1e-4 and (s.min_lstd, s.max_lstd) == (-20.0, 2.0)
    assert s.n_critics == 1 and s.reward_scale == 1.0 and s.ent_coef_mode == L.BB_ENTCOEF_FIX and s.ent_coef_fix == 1.0


def test_errors_do_not_cross_the_abi():
    lib = L.lib()
    n = C.c_int32(-1)
    assert lib.bb_device_count(C.byref(n)) == 0 and n.value >= 0
    if n.value == 0:  # no GPU here: creating a handle must fail with a message, not crash
        c = L.bb_replay_cfg()
        lib.bb_replay_cfg_default(C.byref(c))
        h = C.c_void_p()
        assert lib.bb_replay_create(C.byref(c), C.byref(h)) != 0
        assert len(lib.bb_last_error()) > 0
    assert lib.bb_replay_len(None, None) != 0
    assert b"null" in lib.bb_last_error()


def test_host_loop_library_loads():
    from border_b200 import host_loops as H
    lib = H.host_lib()
    for name in ("bbh_last_error", "bbh_trainer_cfg_default", "bbh_train", "bbh_train_async", "bbh_e2e_steps", "bbh_env_steps"):
        assert hasattr(lib, name)
    c = H.trainer_cfg()
    # TrainerConfig::default (trainer/config.rs:49-62), ActorManagerConfig::default n_buffer = 100
    assert (c.max_opts, c.opt_interval, c.warmup_period, c.n_buffer, c.sync_interval) == (0, 1, 0, 100, 1)
