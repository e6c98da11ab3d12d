"""ctypes wrapper around oracle/replay_oracle.c (the CPU restatement of border's replay path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product (border_b200/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libreplay_oracle.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    src = os.path.join(_HERE, "replay_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_SO)
        l.bo_sumtree_new.restype = C.c_void_p
        l.bo_sumtree_new.argtypes = [C.c_size_t, C.c_float, C.c_int]
        l.bo_sumtree_free.argtypes = [C.c_void_p]
        l.bo_sumtree_total.restype = C.c_float
        l.bo_sumtree_total.argtypes = [C.c_void_p]
        l.bo_sumtree_max.restype = C.c_float
        l.bo_sumtree_max.argtypes = [C.c_void_p]
        l.bo_sumtree_update.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        l.bo_sumtree_add.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        l.bo_sumtree_get.restype = C.c_size_t
        l.bo_sumtree_get.argtypes = [C.c_void_p, C.c_float]
        l.bo_sumtree_sample.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        l.bo_sumtree_tree.restype = C.POINTER(C.c_float)
        l.bo_sumtree_tree.argtypes = [C.c_void_p]
        l.bo_sumtree_n_samples.restype = C.c_size_t
        l.bo_sumtree_n_samples.argtypes = [C.c_void_p]
        l.bo_replay_build.restype = C.c_void_p
        l.bo_replay_build.argtypes = [C.c_size_t, C.c_uint64, C.c_size_t, C.c_size_t, C.c_int, C.c_float, C.c_float,
                                      C.c_float, C.c_size_t, C.c_int, C.c_uint64]
        l.bo_replay_free.argtypes = [C.c_void_p]
        l.bo_replay_len.restype = C.c_size_t
        l.bo_replay_len.argtypes = [C.c_void_p]
        l.bo_replay_head.restype = C.c_size_t
        l.bo_replay_head.argtypes = [C.c_void_p]
        l.bo_replay_push.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_size_t]
        l.bo_replay_sample_indices.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        l.bo_replay_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 6
        l.bo_replay_update_priority.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        l.bo_replay_sumtree.restype = C.c_void_p
        l.bo_replay_sumtree.argtypes = [C.c_void_p]
        l.bo_replay_beta.restype = C.c_float
        l.bo_replay_beta.argtypes = [C.c_void_p]
        l.bo_powf.restype = C.c_float
        l.bo_powf.argtypes = [C.c_float, C.c_float]
        l.bo_chacha_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.bo_stdrng_seed_from_u64.argtypes = [C.c_void_p, C.c_uint64]
        l.bo_stdrng_next_u32.restype = C.c_uint32
        l.bo_stdrng_next_u32.argtypes = [C.c_void_p]
        l.bo_fastrand_seed.argtypes = [C.c_void_p, C.c_uint64]
        l.bo_fastrand_u64.restype = C.c_uint64
        l.bo_fastrand_u64.argtypes = [C.c_void_p]
        l.bo_fastrand_f32.restype = C.c_float
        l.bo_fastrand_f32.argtypes = [C.c_void_p]
        l.bo_fastrand_f64.restype = C.c_double
        l.bo_fastrand_f64.argtypes = [C.c_void_p]
        l.bo_fastrand_u32_below.restype = C.c_uint32
        l.bo_fastrand_u32_below.argtypes = [C.c_void_p, C.c_uint32]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class SumTree:
    """SumTree of border-core/src/generic_replay_buffer/base/sum_tree.rs."""

    def __init__(self, capacity, alpha, normalize="Batch"):
        self.capacity = capacity
        self._h = lib().bo_sumtree_new(capacity, alpha, 0 if normalize == "All" else 1)

    def __del__(self):
        try:
            lib().bo_sumtree_free(self._h)
        except Exception:
            pass

    def add(self, ix, p):
        lib().bo_sumtree_add(self._h, ix, p)

    def update(self, ix, p):
        lib().bo_sumtree_update(self._h, ix, p)

    def get(self, s):
        return lib().bo_sumtree_get(self._h, s)

    def total(self):
        return lib().bo_sumtree_total(self._h)

    def max(self):
        return lib().bo_sumtree_max(self._h)

    def tree(self):
        return np.ctypeslib.as_array(lib().bo_sumtree_tree(self._h), shape=(2 * self.capacity - 1,)).copy()

    def sample(self, u, beta):
        u = np.ascontiguousarray(u, np.float32)
        ix = np.empty(len(u), np.int64)
        w = np.empty(len(u), np.float32)
        lib().bo_sumtree_sample(self._h, len(u), beta, _p(u), _p(ix), _p(w))
        return ix, w


class StdRng:
    """rand 0.8.5 StdRng::seed_from_u64 / next_u32."""

    def __init__(self, seed):
        self._s = (C.c_uint8 * 256)()
        lib().bo_stdrng_seed_from_u64(self._s, seed)

    def next_u32(self):
        return lib().bo_stdrng_next_u32(self._s)

    def key_words(self):
        """The 8 little-endian key words seed_from_u64 produced (first field of bo_stdrng)."""
        return [int.from_bytes(bytes(self._s[4 * i:4 * i + 4]), "little") for i in range(8)]


class FastRand:
    def __init__(self, seed):
        self._s = (C.c_uint64 * 1)()
        lib().bo_fastrand_seed(self._s, seed)

    def u64(self):
        return lib().bo_fastrand_u64(self._s)

    def f32(self):
        return lib().bo_fastrand_f32(self._s)

    def f64(self):
        return lib().bo_fastrand_f64(self._s)

    def u32_below(self, n):
        return lib().bo_fastrand_u32_below(self._s, n)


class ReplayOracle:
    """SimpleReplayBuffer of border-core/src/generic_replay_buffer/base.rs over byte rows."""

    def __init__(self, capacity, seed, obs_shape, obs_dtype, act_shape, act_dtype, per=None, fastrand_seed=0x5EED5EED5EED):
        self.capacity = capacity
        self.obs_shape, self.obs_dtype = tuple(obs_shape), np.dtype(obs_dtype)
        self.act_shape, self.act_dtype = tuple(act_shape), np.dtype(act_dtype)
        ob = int(np.prod(self.obs_shape, dtype=np.int64)) * self.obs_dtype.itemsize
        ab = int(np.prod(self.act_shape, dtype=np.int64)) * self.act_dtype.itemsize
        self.per = per
        if per is None:
            args = (0, 0.6, 0.4, 1.0, 1, 0)
        else:
            args = (1, per["alpha"], per["beta_0"], per["beta_final"], per["n_opts_final"],
                    0 if per["normalize"] == "All" else 1)
        self._h = lib().bo_replay_build(capacity, seed, ob, ab, *args, fastrand_seed)

    def __del__(self):
        try:
            lib().bo_replay_free(self._h)
        except Exception:
            pass

    def len(self):
        return lib().bo_replay_len(self._h)

    def head(self):
        return lib().bo_replay_head(self._h)

    def push(self, obs, act, next_obs, reward, term, trunc):
        obs = np.ascontiguousarray(obs, self.obs_dtype)
        next_obs = np.ascontiguousarray(next_obs, self.obs_dtype)
        act = np.ascontiguousarray(act, self.act_dtype)
        reward = np.ascontiguousarray(reward, np.float32)
        term = np.ascontiguousarray(term, np.int8)
        trunc = np.ascontiguousarray(trunc, np.int8)
        lib().bo_replay_push(self._h, _p(obs), _p(act), _p(next_obs), _p(reward), _p(term), _p(trunc), len(reward))

    def sample_indices(self, batch, u_inject=None):
        ix = np.empty(batch, np.uint64)
        w = np.empty(batch, np.float32) if self.per is not None else None
        u = np.ascontiguousarray(u_inject, np.float32) if u_inject is not None else None
        lib().bo_replay_sample_indices(self._h, batch, _p(u), _p(ix), _p(w))
        return ix, w

    def gather(self, ix):
        ix = np.ascontiguousarray(ix, np.uint64)
        B = len(ix)
        obs = np.empty((B,) + self.obs_shape, self.obs_dtype)
        next_obs = np.empty_like(obs)
        act = np.empty((B,) + self.act_shape, self.act_dtype)
        reward = np.empty(B, np.float32)
        term = np.empty(B, np.int8)
        trunc = np.empty(B, np.int8)
        lib().bo_replay_gather(self._h, _p(ix), B, _p(obs), _p(act), _p(next_obs), _p(reward), _p(term), _p(trunc))
        return dict(obs=obs, act=act, next_obs=next_obs, reward=reward, is_terminated=term, is_truncated=trunc)

    def batch(self, size, u_inject=None):
        ix, w = self.sample_indices(size, u_inject)
        b = self.gather(ix)
        b["ix_sample"], b["weight"] = ix, w
        return b

    def update_priority(self, ixs, td):
        ixs = np.ascontiguousarray(ixs, np.uint64)
        td = np.ascontiguousarray(td, np.float32)
        lib().bo_replay_update_priority(self._h, _p(ixs), _p(td), len(ixs))

    def sum_tree(self):
        st = lib().bo_replay_sumtree(self._h)
        n = 2 * self.capacity - 1
        return np.ctypeslib.as_array(lib().bo_sumtree_tree(st), shape=(n,)).copy(), lib().bo_sumtree_n_samples(st)

    def beta(self):
        return lib().bo_replay_beta(self._h)


def powf(x, y):
    return lib().bo_powf(float(np.float32(x)), float(np.float32(y)))
