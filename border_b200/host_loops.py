"""ctypes binding of include/border_host.h: border's Trainer / train_async loops restated in C++
(border_b200/host/border_host.hpp) over the C ABI, with synthetic environments."""
import ctypes as C
import os

from . import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libborder_host.so")
ALGO = {"dqn": 0, "iqn": 1, "sac": 2}


class bbh_env_cfg(C.Structure):
    _fields_ = [("obs_kind", C.c_int32), ("obs_elems", C.c_uint32), ("episode_len", C.c_uint64),
                ("truncate_len", C.c_uint64)]


class bbh_trainer_cfg(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("max_opts", "opt_interval", "eval_interval", "flush_record_interval",
                                           "record_compute_cost_interval", "record_agent_info_interval", "warmup_period",
                                           "save_interval", "sync_interval", "n_actors", "n_buffer", "env_seed")]


class bbh_train_stat(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("env_steps", "opt_steps", "records", "saves", "buffer_len", "agent_n_opts",
                                           "samples_total", "syncs")] + \
               [(n, C.c_double) for n in ("opt_seconds", "sample_seconds", "total_seconds", "samples_per_sec", "opt_per_sec")] + \
               [("last_loss", C.c_float)]


_hl = None


def host_lib():
    global _hl
    if _hl is None:
        L.lib()  # libborder_b200.so first (RTLD_GLOBAL)
        if not os.path.exists(HOST_LIB_PATH):
            raise L.BorderB200Error("%s is missing: build it with `make`" % HOST_LIB_PATH)
        l = C.CDLL(HOST_LIB_PATH)
        l.bbh_last_error.restype = C.c_char_p
        l.bbh_trainer_cfg_default.argtypes = [C.POINTER(bbh_trainer_cfg)]
        l.bbh_train.restype = C.c_int32
        l.bbh_train.argtypes = [C.c_int32, C.c_void_p, C.POINTER(L.bb_replay_cfg), C.POINTER(bbh_env_cfg),
                                C.POINTER(bbh_trainer_cfg), C.c_char_p, C.POINTER(bbh_train_stat)]
        l.bbh_train_async.restype = C.c_int32
        l.bbh_train_async.argtypes = [C.c_int32, C.c_void_p, C.POINTER(L.bb_replay_cfg), C.POINTER(bbh_env_cfg),
                                      C.POINTER(bbh_trainer_cfg), C.POINTER(bbh_train_stat)]
        l.bbh_e2e_steps.restype = C.c_int32
        l.bbh_e2e_steps.argtypes = [C.c_void_p] * 8 + [C.c_uint64] * 4 + [C.POINTER(C.c_float)]
        l.bbh_sampler_trace.restype = C.c_int32
        l.bbh_sampler_trace.argtypes = [C.POINTER(bbh_env_cfg), C.c_uint64, C.c_uint64, C.c_void_p]
        l.bbh_env_steps.restype = C.c_int32
        l.bbh_env_steps.argtypes = [C.c_void_p] * 7 + [C.c_uint64] * 3 + [C.POINTER(C.c_int64)]
        l.bbh_actor_steps.restype = C.c_int32
        l.bbh_actor_steps.argtypes = [C.c_void_p] * 7 + [C.c_uint64] * 3 + [C.POINTER(C.c_int64)]
        _hl = l
    return _hl


def _check(rc):
    if rc != 0:
        raise L.BorderB200Error(host_lib().bbh_last_error().decode("utf-8", "replace"))


def trainer_cfg(**kw):
    c = bbh_trainer_cfg()
    host_lib().bbh_trainer_cfg_default(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _stat(st):
    return {n: getattr(st, n) for n, _ in bbh_train_stat._fields_}


def train(algo, agent_cfg, replay_cfg, env_cfg, tcfg, save_dir=None):
    """Trainer::train (border-core/src/trainer.rs:267-327)."""
    st = bbh_train_stat()
    _check(host_lib().bbh_train(ALGO[algo], C.cast(C.pointer(agent_cfg), C.c_void_p), C.byref(replay_cfg), C.byref(env_cfg),
                                C.byref(tcfg), save_dir.encode() if save_dir else None, C.byref(st)))
    return _stat(st)


def train_offline(algo, agent_cfg, dataset, tcfg, save_dir=None):
    """Trainer::train_offline (border-core/src/trainer.rs:330-384) on a replay buffer that already holds the dataset."""
    st = bbh_train_stat()
    lib = host_lib()
    lib.bbh_train_offline.restype = C.c_int32
    lib.bbh_train_offline.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p]
    _check(lib.bbh_train_offline(ALGO[algo], C.cast(C.pointer(agent_cfg), C.c_void_p), dataset.handle, C.byref(tcfg),
                                 save_dir.encode() if save_dir else None, C.byref(st)))
    return _stat(st)


LEARNER_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_void_p)


def train_async(algo, agent_cfg, replay_cfg, env_cfg, tcfg, on_learner=None):
    """train_async (border-async-trainer/src/util.rs:31-92).  on_learner(handle, phase): called on the learner thread with
    the learner's bb_agent handle, phase 0 after it was created (connect gradient peers here), 1 after the last update."""
    st = bbh_train_stat()
    lib = host_lib()
    lib.bbh_train_async_ex.restype = C.c_int32
    lib.bbh_train_async_ex.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, LEARNER_HOOK, C.c_void_p,
                                       C.c_void_p]
    errors = []

    def _hook(handle, phase, _user):
        try:
            on_learner(C.c_void_p(handle), phase)
        except BaseException as e:  # an exception must not unwind through the C++ frames
            errors.append(e)

    cb = LEARNER_HOOK(_hook) if on_learner else LEARNER_HOOK(0)
    _check(lib.bbh_train_async_ex(ALGO[algo], C.cast(C.pointer(agent_cfg), C.c_void_p), C.byref(replay_cfg), C.byref(env_cfg),
                                  C.byref(tcfg), cb, None, C.byref(st)))
    if errors:
        raise errors[0]
    return _stat(st)


def e2e_steps(agent, buffer, obs, act, next_obs, reward, is_terminated, is_truncated, n_steps):
    """n_steps x [ExperienceBufferBase::push(one host transition); Agent::opt_with_record] over the C ABI, in C++
    (the Trainer inner loop, border-core/src/trainer.rs:206-228).  Arrays hold n_slots transitions, pushed round-robin."""
    import numpy as np
    n = len(reward)
    arrs = [np.ascontiguousarray(a) for a in (obs, act, next_obs, reward, is_terminated, is_truncated)]
    assert arrs[3].dtype == np.float32 and arrs[4].dtype == np.int8 and arrs[5].dtype == np.int8
    loss = C.c_float()
    _check(host_lib().bbh_e2e_steps(agent.handle, buffer.handle, *[a.ctypes.data for a in arrs],
                                    arrs[0].nbytes // n, arrs[1].nbytes // n, n, n_steps, C.byref(loss)))
    return loss.value


def env_steps(agent, buffer, obs, next_obs, reward, is_terminated, is_truncated, n_steps):
    """n_steps x Sampler::sample_and_push (trainer/sampler.rs:99-144) with a zero-cost environment, in C++ over the C ABI:
    Policy::sample on a host observation + push of the host transition.  Returns the last action."""
    import numpy as np
    n = len(reward)
    arrs = [np.ascontiguousarray(a) for a in (obs, next_obs, reward, is_terminated, is_truncated)]
    act = C.c_int64()
    _check(host_lib().bbh_env_steps(agent.handle, buffer.handle, *[a.ctypes.data for a in arrs], arrs[0].nbytes // n, n,
                                    n_steps, C.byref(act)))
    return act.value


def actor_steps(agent, buffer, obs, next_obs, reward, is_terminated, is_truncated, n_steps):
    """env_steps through bb_actor_step: the observation crosses PCIe once per step, the explorer runs in the tail of the
    policy forward and the transition is pushed from its device-resident copies.  Returns the last action."""
    import numpy as np
    n = len(reward)
    arrs = [np.ascontiguousarray(a) for a in (obs, next_obs, reward, is_terminated, is_truncated)]
    act = C.c_int64()
    _check(host_lib().bbh_actor_steps(agent.handle, buffer.handle, *[a.ctypes.data for a in arrs], arrs[0].nbytes // n, n,
                                      n_steps, C.byref(act)))
    return act.value


def sampler_trace(obs_elems, episode_len, truncate_len, seed, n_steps):
    """The transitions Sampler + SimpleStepProcessor (C++ mirror) emit over the synthetic u8 environment with a scripted
    policy: an int64 array [n_steps][8] (see include/border_host.h: bbh_sampler_trace)."""
    import numpy as np
    cfg = bbh_env_cfg(L.BB_U8, obs_elems, episode_len, truncate_len)
    out = np.zeros((n_steps, 8), np.int64)
    _check(host_lib().bbh_sampler_trace(C.byref(cfg), seed, n_steps, out.ctypes.data))
    return out
