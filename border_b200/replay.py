"""Host-side mirror of border-core's generic replay buffer over the C ABI.

Names, argument meaning and error behaviour follow the reference:
  SimpleReplayBufferConfig / PerConfig   border-core/src/generic_replay_buffer/config.rs:45-65,185-197
  GenericTransitionBatch                  border-core/src/generic_replay_buffer/batch.rs:89-117
  SimpleReplayBuffer                      border-core/src/generic_replay_buffer/base.rs:86-426
  ExperienceBufferBase / ReplayBufferBase border-core/src/base/replay_buffer.rs:38-127
Storage lives in HBM behind a bb_replay handle; like TensorBatch (border-tch-agent/src/tensor_batch.rs:85-110)
the row dtype/shape is taken from the first pushed item.
"""
import ctypes as C
from dataclasses import dataclass, field, asdict
from typing import Optional

import numpy as np

from . import _lib as L

_KIND = {np.dtype(np.uint8): L.BB_U8, np.dtype(np.float32): L.BB_F32, np.dtype(np.int64): L.BB_I64,
         np.dtype(np.int32): L.BB_I32}
_NORM = {"All": L.BB_NORM_ALL, "Batch": L.BB_NORM_BATCH}


@dataclass
class PerConfig:
    alpha: float = 0.6
    beta_0: float = 0.4
    beta_final: float = 1.0
    n_opts_final: int = 500_000
    normalize: str = "All"  # WeightNormalizer::{All, Batch}


@dataclass
class SimpleReplayBufferConfig:
    capacity: int = 10000
    seed: int = 42
    per_config: Optional[PerConfig] = None

    # chained builders, as in config.rs:209-221
    def with_capacity(self, v):
        self.capacity = v
        return self

    def with_seed(self, v):
        self.seed = v
        return self

    def with_per_config(self, v):
        self.per_config = v
        return self

    @staticmethod
    def load(path):
        import yaml
        d = yaml.safe_load(open(path))
        per = d.get("per_config")
        return SimpleReplayBufferConfig(d["capacity"], d["seed"], PerConfig(**per) if per else None)

    def save(self, path):
        import yaml
        yaml.safe_dump(asdict(self), open(path, "w"))


@dataclass
class GenericTransitionBatch:
    obs: np.ndarray
    act: np.ndarray
    next_obs: np.ndarray
    reward: np.ndarray
    is_terminated: np.ndarray
    is_truncated: np.ndarray
    ix_sample: Optional[np.ndarray] = None
    weight: Optional[np.ndarray] = None

    def len(self):
        return len(self.reward)

    def unpack(self):
        return (self.obs, self.act, self.next_obs, self.reward, self.is_terminated, self.is_truncated,
                self.ix_sample, self.weight)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class SimpleReplayBuffer:
    """ExperienceBufferBase + ReplayBufferBase over a device-resident ring."""

    def __init__(self, config: SimpleReplayBufferConfig, device=0, fastrand_seed=0x5EED5EED5EED):
        self.config = config
        self.device = device
        self.fastrand_seed = fastrand_seed
        self._h = None
        self._obs_shape = self._act_shape = None
        self._obs_dtype = self._act_dtype = None

    @classmethod
    def build(cls, config, **kw):  # ReplayBufferBase::build, base.rs:336-356
        return cls(config, **kw)

    # -- handle management -------------------------------------------------------------------
    def _ensure(self, obs, act):
        if self._h is not None:
            return
        self.allocate(obs.shape[1:], obs.dtype, act.shape[1:], act.dtype)

    def allocate(self, obs_shape, obs_dtype, act_shape, act_dtype):
        """Explicit allocation (TensorBatch does this lazily on the first push)."""
        lib = L.lib()
        cfg = L.bb_replay_cfg()
        lib.bb_replay_cfg_default(C.byref(cfg))
        cfg.capacity = self.config.capacity
        cfg.seed = self.config.seed
        per = self.config.per_config
        cfg.per_config_some = 1 if per is not None else 0
        if per is not None:
            cfg.alpha, cfg.beta_0, cfg.beta_final = per.alpha, per.beta_0, per.beta_final
            cfg.n_opts_final = per.n_opts_final
            cfg.normalize = _NORM[per.normalize]
        self._obs_shape, self._act_shape = tuple(obs_shape), tuple(act_shape)
        self._obs_dtype, self._act_dtype = np.dtype(obs_dtype), np.dtype(act_dtype)
        cfg.obs_kind = _KIND[self._obs_dtype]
        cfg.obs_elems = int(np.prod(self._obs_shape, dtype=np.int64))
        cfg.act_kind = _KIND[self._act_dtype]
        cfg.act_elems = int(np.prod(self._act_shape, dtype=np.int64))
        cfg.fastrand_seed = self.fastrand_seed
        cfg.device = self.device
        h = C.c_void_p()
        L.check(lib.bb_replay_create(C.byref(cfg), C.byref(h)))
        self._h = h
        return self

    @property
    def handle(self):
        if self._h is None:
            raise L.BorderB200Error("replay buffer has no storage yet (nothing was pushed)")
        return self._h

    def close(self):
        if self._h is not None:
            L.lib().bb_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- ExperienceBufferBase ----------------------------------------------------------------
    def push(self, tr: GenericTransitionBatch):
        obs = np.ascontiguousarray(tr.obs)
        act = np.ascontiguousarray(tr.act)
        next_obs = np.ascontiguousarray(tr.next_obs, dtype=obs.dtype)
        n = tr.len()
        if n == 0:
            return
        self._ensure(obs, act)
        reward = np.ascontiguousarray(tr.reward, dtype=np.float32)
        term = np.ascontiguousarray(tr.is_terminated, dtype=np.int8)
        trunc = np.ascontiguousarray(tr.is_truncated, dtype=np.int8)
        assert obs.dtype == self._obs_dtype and act.dtype == self._act_dtype
        L.check(L.lib().bb_replay_push(self.handle, _p(obs), _p(act), _p(next_obs), _p(reward), _p(term), _p(trunc),
                                       n, 0))

    def len(self):
        if self._h is None:
            return 0
        out = C.c_uint64()
        L.check(L.lib().bb_replay_len(self._h, C.byref(out)))
        return out.value

    __len__ = len

    # -- ReplayBufferBase --------------------------------------------------------------------
    def batch_device(self, size):
        """batch() that leaves the sampled transitions in HBM (what the agents consume)."""
        view = L.bb_batch_view()
        L.check(L.lib().bb_replay_sample(self.handle, size, C.byref(view)))
        return view

    def batch(self, size) -> GenericTransitionBatch:  # base.rs:376-402
        self.batch_device(size)
        per = self.config.per_config is not None
        obs = np.empty((size,) + self._obs_shape, self._obs_dtype)
        next_obs = np.empty_like(obs)
        act = np.empty((size,) + self._act_shape, self._act_dtype)
        reward = np.empty(size, np.float32)
        term = np.empty(size, np.int8)
        trunc = np.empty(size, np.int8)
        ix = np.empty(size, np.uint64)
        w = np.empty(size, np.float32) if per else None
        L.check(L.lib().bb_replay_batch_to_host(self.handle, _p(obs), _p(act), _p(next_obs), _p(reward), _p(term),
                                                _p(trunc), _p(ix), _p(w)))
        return GenericTransitionBatch(obs, act, next_obs, reward, term, trunc, ix, w)

    def last_indices(self):
        """ix_sample of the last sampled batch (also when an agent's opt() drew it on the device)."""
        n = C.c_uint64()
        L.check(L.lib().bb_replay_last_batch(self.handle, C.byref(n)))
        ix = np.empty(n.value, np.uint64)
        L.check(L.lib().bb_replay_batch_to_host(self.handle, None, None, None, None, None, None, _p(ix), None))
        return ix

    def update_priority(self, ixs, td_errs):  # base.rs:413-426
        if self.config.per_config is None:
            return
        if ixs is None:
            raise L.BorderB200Error("ixs should be Some(_) in update_priority().")
        if td_errs is None:
            raise L.BorderB200Error("td_errs should be Some(_) in update_priority().")
        ixs = np.ascontiguousarray(ixs, dtype=np.uint64)
        td = np.ascontiguousarray(td_errs, dtype=np.float32)
        L.check(L.lib().bb_replay_update_priority(self.handle, _p(ixs), _p(td), len(ixs), 0))

    # -- test hooks / benchmark set-up -------------------------------------------------------
    def inject_uniforms(self, u):
        u = np.ascontiguousarray(u, dtype=np.float32)
        L.check(L.lib().bb_replay_inject_uniforms(self.handle, _p(u), len(u)))

    def dump_sum_tree(self):
        tree = np.empty(2 * self.config.capacity - 1, np.float32)
        ns, no = C.c_uint64(), C.c_uint64()
        L.check(L.lib().bb_replay_dump_sum_tree(self.handle, _p(tree), C.byref(ns), C.byref(no)))
        return tree, ns.value, no.value

    def state(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        L.check(L.lib().bb_replay_state(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return dict(i=a.value, size=b.value, rng_words=c.value)

    def fill_synthetic(self, n_rows, n_actions=6, seed=1234):
        L.check(L.lib().bb_replay_fill_synthetic(self.handle, n_rows, n_actions, seed))

    def set_stream(self, cuda_stream_ptr):
        L.check(L.lib().bb_replay_set_stream(self.handle, C.c_void_p(cuda_stream_ptr)))
