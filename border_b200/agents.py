"""Host-side mirror of border-tch-agent's agents (Policy + Agent + Configurable + SyncModel) over
the C ABI.  Config field names are the reference's:
  DqnConfig  border-tch-agent/src/dqn/config.rs:26-48     MlpConfig      mlp/config.rs:7-12
  SacConfig  border-tch-agent/src/sac/config.rs:23-47     AtariCnnConfig cnn/config.rs:13-18
  IqnConfig  border-tch-agent/src/iqn/config.rs:21-41     OptimizerConfig opt.rs:13-28
Agent surface: border-core/src/base/agent.rs:24-136, policy.rs:49-63; SyncModel:
border-async-trainer/src/sync_model.rs:2-13.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _lib as L
from .replay import SimpleReplayBuffer, _p


@dataclass
class MlpConfig:
    in_dim: int = 4
    units: List[int] = field(default_factory=lambda: [64, 64])
    out_dim: int = 2
    activation_out: bool = False

    def to_c(self):
        n = L.bb_net_cfg()
        n.kind = L.BB_NET_MLP
        n.in_dim, n.n_units, n.out_dim = self.in_dim, len(self.units), self.out_dim
        for i, u in enumerate(self.units):
            n.units[i] = u
        n.activation_out = int(self.activation_out)
        return n


@dataclass
class AtariCnnConfig:
    n_stack: int = 4
    out_dim: int = 6
    skip_linear: bool = False

    def to_c(self):
        n = L.bb_net_cfg()
        n.kind = L.BB_NET_ATARI_CNN
        n.n_stack, n.out_dim, n.skip_linear = self.n_stack, self.out_dim, int(self.skip_linear)
        return n


@dataclass
class OptimizerConfig:
    """OptimizerConfig::{Adam{lr}, AdamW{lr,beta1,beta2,wd,eps,amsgrad}} (opt.rs:13-28)."""
    kind: str = "Adam"
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    wd: float = 0.0
    eps: float = 1e-8
    amsgrad: bool = False

    def to_c(self):
        o = L.bb_opt_cfg()
        o.kind = L.BB_OPT_ADAMW if self.kind == "AdamW" else L.BB_OPT_ADAM
        o.lr, o.beta1, o.beta2, o.wd, o.eps, o.amsgrad = self.lr, self.beta1, self.beta2, self.wd, self.eps, int(self.amsgrad)
        return o


@dataclass
class EpsilonGreedy:
    eps_start: float = 1.0
    eps_final: float = 0.02
    final_step: int = 100_000


@dataclass
class Softmax:
    pass


@dataclass
class DqnModelConfig:
    q_config: object = None
    opt_config: OptimizerConfig = field(default_factory=OptimizerConfig)


@dataclass
class DqnConfig:
    model_config: DqnModelConfig = field(default_factory=DqnModelConfig)
    soft_update_interval: int = 1
    n_updates_per_opt: int = 1
    batch_size: int = 1
    discount_factor: float = 0.99
    tau: float = 0.005
    train: bool = False
    explorer: object = field(default_factory=Softmax)
    clip_reward: Optional[float] = None
    double_dqn: bool = False
    clip_td_err: Optional[Tuple[float, float]] = None
    device: Optional[int] = None  # Device::Cuda(n)
    critic_loss: str = "Mse"
    record_verbose_level: int = 0
    init_seed: int = 0
    explorer_seed: int = 0x0123456789ABCDEF

    def to_c(self):
        lib = L.lib()
        c = L.bb_dqn_cfg()
        lib.bb_dqn_cfg_default(C.byref(c))
        if self.device is None:
            raise L.BorderB200Error("No device is given for DQN agent")  # dqn/base.rs:258
        c.q_config = self.model_config.q_config.to_c()
        c.opt_config = self.model_config.opt_config.to_c()
        c.soft_update_interval, c.n_updates_per_opt, c.batch_size = (self.soft_update_interval,
                                                                       self.n_updates_per_opt, self.batch_size)
        c.discount_factor, c.tau, c.train = self.discount_factor, self.tau, int(self.train)
        if isinstance(self.explorer, EpsilonGreedy):
            c.explorer = L.BB_EXPLORER_EPS_GREEDY
            c.eps_start, c.eps_final, c.final_step = self.explorer.eps_start, self.explorer.eps_final, self.explorer.final_step
        else:
            c.explorer = L.BB_EXPLORER_SOFTMAX
        c.clip_reward_some = int(self.clip_reward is not None)
        c.clip_reward = self.clip_reward or 0.0
        c.double_dqn = int(self.double_dqn)
        c.clip_td_err_some = int(self.clip_td_err is not None)
        if self.clip_td_err is not None:
            c.clip_td_err_min, c.clip_td_err_max = self.clip_td_err
        c.device = self.device
        c.critic_loss = L.BB_LOSS_SMOOTH_L1 if self.critic_loss == "SmoothL1" else L.BB_LOSS_MSE
        c.record_verbose_level = self.record_verbose_level
        c.init_seed, c.explorer_seed = self.init_seed, self.explorer_seed
        return c


@dataclass
class SacConfig:
    pi_config: MlpConfig = None       # actor_config.pi_config (Mlp2: out_dim = action dim)
    pi_opt_config: OptimizerConfig = field(default_factory=lambda: OptimizerConfig(lr=3e-4))
    q_config: MlpConfig = None        # critic_config.q_config (in_dim = obs+act, out_dim = 1)
    q_opt_config: OptimizerConfig = field(default_factory=lambda: OptimizerConfig(lr=3e-4))
    gamma: float = 0.99
    tau: float = 0.005
    ent_coef_mode: object = ("Fix", 1.0)  # ("Fix", alpha) | ("Auto", target_entropy, lr)
    epsilon: float = 1e-4
    min_lstd: float = -20.0
    max_lstd: float = 2.0
    n_updates_per_opt: int = 1
    batch_size: int = 1
    train: bool = False
    critic_loss: str = "Mse"
    reward_scale: float = 1.0
    n_critics: int = 1
    seed: Optional[int] = None
    device: Optional[int] = None
    init_seed: int = 0
    noise_seed: int = 0x5AC5AC5AC

    def to_c(self):
        lib = L.lib()
        c = L.bb_sac_cfg()
        lib.bb_sac_cfg_default(C.byref(c))
        if self.device is None:
            raise L.BorderB200Error("No device is given for SAC agent")
        c.pi_config, c.pi_opt_config = self.pi_config.to_c(), self.pi_opt_config.to_c()
        c.q_config, c.q_opt_config = self.q_config.to_c(), self.q_opt_config.to_c()
        c.gamma, c.tau = self.gamma, self.tau
        if self.ent_coef_mode[0] == "Auto":
            c.ent_coef_mode = L.BB_ENTCOEF_AUTO
            c.ent_coef_target, c.ent_coef_lr = self.ent_coef_mode[1], self.ent_coef_mode[2]
        else:
            c.ent_coef_mode = L.BB_ENTCOEF_FIX
            c.ent_coef_fix = self.ent_coef_mode[1]
        c.epsilon, c.min_lstd, c.max_lstd = self.epsilon, self.min_lstd, self.max_lstd
        c.n_updates_per_opt, c.batch_size, c.train = self.n_updates_per_opt, self.batch_size, int(self.train)
        c.critic_loss = L.BB_LOSS_SMOOTH_L1 if self.critic_loss == "SmoothL1" else L.BB_LOSS_MSE
        c.reward_scale, c.n_critics = self.reward_scale, self.n_critics
        c.seed_some, c.seed = int(self.seed is not None), self.seed or 0
        c.device, c.init_seed, c.noise_seed = self.device, self.init_seed, self.noise_seed
        return c


_IQN_SAMPLE = {"Const10": L.BB_IQN_CONST10, "Uniform8": L.BB_IQN_UNIFORM8, "Uniform10": L.BB_IQN_UNIFORM10,
               "Uniform32": L.BB_IQN_UNIFORM32, "Uniform64": L.BB_IQN_UNIFORM64, "Median": L.BB_IQN_MEDIAN,
               "Const32": L.BB_IQN_CONST32}


@dataclass
class IqnConfig:
    f_config: object = None           # model_config.f_config (feature extractor)
    m_config: MlpConfig = None        # model_config.m_config (merge net)
    opt_config: OptimizerConfig = field(default_factory=OptimizerConfig)
    feature_dim: int = 64
    embed_dim: int = 64
    soft_update_interval: int = 1
    n_updates_per_opt: int = 1
    batch_size: int = 1
    discount_factor: float = 0.99
    tau: float = 0.005
    train: bool = False
    sample_percents_pred: str = "Uniform64"
    sample_percents_tgt: str = "Uniform64"
    sample_percents_act: str = "Uniform32"
    explorer: EpsilonGreedy = field(default_factory=EpsilonGreedy)
    device: Optional[int] = None
    init_seed: int = 0
    explorer_seed: int = 0x0123456789ABCDEF
    tau_seed: int = 0x7A07A0

    def to_c(self):
        lib = L.lib()
        c = L.bb_iqn_cfg()
        lib.bb_iqn_cfg_default(C.byref(c))
        if self.device is None:
            raise L.BorderB200Error("No device is given for IQN agent")
        c.f_config, c.m_config, c.opt_config = self.f_config.to_c(), self.m_config.to_c(), self.opt_config.to_c()
        c.feature_dim, c.embed_dim = self.feature_dim, self.embed_dim
        c.soft_update_interval, c.n_updates_per_opt, c.batch_size = (self.soft_update_interval,
                                                                       self.n_updates_per_opt, self.batch_size)
        c.discount_factor, c.tau, c.train = self.discount_factor, self.tau, int(self.train)
        c.sample_percents_pred = _IQN_SAMPLE[self.sample_percents_pred]
        c.sample_percents_tgt = _IQN_SAMPLE[self.sample_percents_tgt]
        c.sample_percents_act = _IQN_SAMPLE[self.sample_percents_act]
        c.eps_start, c.eps_final, c.final_step = self.explorer.eps_start, self.explorer.eps_final, self.explorer.final_step
        c.device, c.init_seed, c.explorer_seed, c.tau_seed = self.device, self.init_seed, self.explorer_seed, self.tau_seed
        return c


class Agent:
    """Policy<E> + Agent<E, R> + SyncModel behind a bb_agent handle."""
    _act_dtype = np.int64
    _models = ()

    def __init__(self, handle, config):
        self._h = handle
        self.config = config

    def close(self):
        if self._h is not None:
            L.lib().bb_agent_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # Agent::train / eval / is_train
    def train(self):
        L.check(L.lib().bb_agent_set_train(self._h, 1))

    def eval(self):
        L.check(L.lib().bb_agent_set_train(self._h, 0))

    def is_train(self):
        out = C.c_int32()
        L.check(L.lib().bb_agent_is_train(self._h, C.byref(out)))
        return bool(out.value)

    # Policy::sample
    def sample(self, obs):
        obs = np.ascontiguousarray(obs)
        n = obs.shape[0]
        out = np.empty((n, self._act_dim()), self._act_dtype)
        L.check(L.lib().bb_agent_sample(self._h, _p(obs), n, _p(out)))
        return out

    def _act_dim(self):
        return 1

    def actor_step(self, buffer, obs, reward=0.0, is_terminated=0, is_truncated=0, reset_obs=None):
        """Sampler::sample_and_push (trainer/sampler.rs:99-144) for one env step with device-resident observations
        (include/border_b200.h: bb_actor_step): pushes (previous obs, previous action, obs, reward, flags) -- nothing on
        the first call -- and returns the action for `reset_obs if the episode ended else obs`."""
        obs = np.ascontiguousarray(obs)
        ro = None if reset_obs is None else np.ascontiguousarray(reset_obs)
        act = C.c_int64()
        L.check(L.lib().bb_actor_step(self._h, buffer.handle, _p(obs), None if ro is None else _p(ro), float(reward),
                                      int(is_terminated), int(is_truncated), C.byref(act)))
        return act.value

    def actor_step_dev(self, buffer, obs_dev, reward=0.0, is_terminated=0, is_truncated=0, reset_obs_dev=None):
        """actor_step with the observation already in HBM (device pointers, e.g. AtariPreprocessor.obs_device())."""
        act = C.c_int64()
        L.check(L.lib().bb_actor_step_dev(self._h, buffer.handle, C.c_void_p(obs_dev),
                                          None if reset_obs_dev is None else C.c_void_p(reset_obs_dev), float(reward),
                                          int(is_terminated), int(is_truncated), C.byref(act)))
        return act.value

    def actor_step_n(self, buffer, obs, reward, is_terminated, is_truncated, reset_obs=None, reset_mask=None):
        """actor_step for n environments at once (obs [n, ...], the other arguments [n]); returns the n actions."""
        obs = np.ascontiguousarray(obs)
        n = obs.shape[0]
        r = np.ascontiguousarray(reward, np.float32)
        t = np.ascontiguousarray(is_terminated, np.int8)
        tr = np.ascontiguousarray(is_truncated, np.int8)
        ro = None if reset_obs is None else np.ascontiguousarray(reset_obs, obs.dtype)
        rm = None if reset_mask is None else np.ascontiguousarray(reset_mask, np.int8)
        out = np.empty(n, np.int64)
        L.check(L.lib().bb_actor_step_n(self._h, buffer.handle, n, _p(obs), None if ro is None else _p(ro),
                                        None if rm is None else _p(rm), _p(r), _p(t), _p(tr), _p(out), 0))
        return out

    def actor_reset(self):
        L.check(L.lib().bb_actor_reset(self._h))

    # Agent::opt / opt_with_record
    def opt(self, buffer: SimpleReplayBuffer):
        L.check(L.lib().bb_agent_opt(self._h, buffer.handle, None))

    def opt_with_record(self, buffer: SimpleReplayBuffer):
        rec = L.bb_record()
        L.check(L.lib().bb_agent_opt(self._h, buffer.handle, C.byref(rec)))
        return self._record(rec)

    def _record(self, rec):
        return {"loss": rec.loss}

    def opt_profiled(self, buffer: SimpleReplayBuffer):
        """One opt() with per-kernel CUDA-event timing: list of (label, ms)."""
        buf = C.create_string_buffer(1 << 16)
        L.check(L.lib().bb_agent_opt_profiled(self._h, buffer.handle, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            k, v = line.rsplit(" ", 1)
            out.append((k, float(v)))
        return out

    def n_opts(self):
        out = C.c_uint64()
        L.check(L.lib().bb_agent_n_opts(self._h, C.byref(out)))
        return out.value

    # Agent::save_params / load_params
    def set_precision(self, fast):
        """fast=False (default): 3xTF32, fp32 parity; fast=True: single-pass TF32 contractions (see border_b200.h)."""
        L.check(L.lib().bb_agent_set_precision(self._h, 1 if fast else 0))

    def save_params(self, path):
        """Agent::save_params (dqn/base.rs:348-362, sac/base.rs:313-345): one tch `VarStore` archive `<model>.pt.tch` per
        model -- the file names and tensor names a reference-side `load_params` expects (checkpoint.py) -- and returns their
        paths like the trait does.  The Adam moments / step counts the reference does not save go to this library's
        side-car files `<model>.pt.tch.b200` in the same directory (true resume)."""
        from . import checkpoint
        os.makedirs(str(path), exist_ok=True)  # fs::create_dir_all
        L.check(L.lib().bb_agent_save_params(self._h, str(path).encode()))
        out = []
        for m in self._models:
            f = os.path.join(str(path), m + ".pt.tch")
            checkpoint.write_varstore(f, self.named_parameters(m))
            out.append(f)
        return out

    def load_params(self, path):
        """Agent::load_params: reads `<model>.pt.tch` archives written by the reference (tch VarStore::save) or by
        save_params.  When this library's side-car is present the optimizer state is restored from it first; otherwise
        the Adam moments and step of the live models are reset (a reference checkpoint carries none)."""
        from . import checkpoint
        side = all(os.path.exists(os.path.join(str(path), m + ".pt.tch.b200")) for m in self._models)
        if side:
            L.check(L.lib().bb_agent_load_params(self._h, str(path).encode()))
        for m in self._models:
            f = os.path.join(str(path), m + ".pt.tch")
            if os.path.exists(f):
                self.set_parameters(m, checkpoint.read_varstore(f))
            elif not side:
                raise FileNotFoundError(f)
        if not side:
            L.check(L.lib().bb_agent_reset_opt_state(self._h))

    # named tensors in the reference layout
    def named_parameters(self, model):
        lib = L.lib()
        nt = C.c_uint64()
        L.check(lib.bb_agent_param_count(self._h, model.encode(), C.byref(nt), None))
        out = {}
        for i in range(nt.value):
            name = C.create_string_buffer(128)
            shape = (C.c_int64 * 4)()
            nd = C.c_int32()
            L.check(lib.bb_agent_param_info(self._h, model.encode(), i, name, 128, shape, C.byref(nd)))
            shp = tuple(shape[k] for k in range(nd.value))
            a = np.empty(shp, np.float32)
            L.check(lib.bb_agent_get_param(self._h, model.encode(), name.value, _p(a), a.size))
            out[name.value.decode()] = a
        return out

    def set_parameters(self, model, tensors):
        for k, v in tensors.items():
            a = np.ascontiguousarray(v, dtype=np.float32)
            L.check(L.lib().bb_agent_set_param(self._h, model.encode(), k.encode(), _p(a), a.size))

    def opt_state(self, model, name, shape):
        m = np.empty(shape, np.float32)
        v = np.empty(shape, np.float32)
        step = C.c_uint64()
        L.check(L.lib().bb_agent_get_opt_state(self._h, model.encode(), name.encode(), _p(m), _p(v), m.size,
                                               C.byref(step)))
        return m, v, step.value

    # SyncModel
    def model_info(self):
        n = C.c_uint64()
        L.check(L.lib().bb_agent_model_info_size(self._h, C.byref(n)))
        blob = np.empty(n.value, np.float32)
        n_opts = C.c_uint64()
        L.check(L.lib().bb_agent_model_info(self._h, _p(blob), blob.size, C.byref(n_opts)))
        return n_opts.value, blob

    def sync_model(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        L.check(L.lib().bb_agent_sync_model(self._h, _p(blob), blob.size))

    def sync_model_from(self, other):
        L.check(L.lib().bb_agent_sync_model_from(self._h, other._h))

    def inject_noise(self, slot, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        L.check(L.lib().bb_agent_inject_noise(self._h, slot, _p(a), a.size))

    def set_stream(self, cuda_stream_ptr):
        L.check(L.lib().bb_agent_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))


class Dqn(Agent):
    _models = ("qnet", "qnet_tgt")

    @classmethod
    def build(cls, config: DqnConfig):  # Configurable::build, dqn/base.rs:255-287
        c = config.to_c()
        h = C.c_void_p()
        L.check(L.lib().bb_dqn_create(C.byref(c), C.byref(h)))
        return cls(h, config)

    def _record(self, rec):
        r = {"loss": rec.loss}
        if self.config.record_verbose_level >= 2:
            r.update(pred_mean=rec.pred_mean, tgt_mean=rec.tgt_mean, reward_mean=rec.reward_mean,
                     tgt_minus_pred_mean=rec.tgt_minus_pred_mean)
        return r


class Sac(Agent):
    _act_dtype = np.float32

    @classmethod
    def build(cls, config: SacConfig):
        c = config.to_c()
        h = C.c_void_p()
        L.check(L.lib().bb_sac_create(C.byref(c), C.byref(h)))
        a = cls(h, config)
        a._models = tuple(["pi", "ent_coef"] + ["qnet_%d" % i for i in range(config.n_critics)] +
                          ["qnet_tgt_%d" % i for i in range(config.n_critics)])
        return a

    def _act_dim(self):
        return self.config.pi_config.out_dim

    def _record(self, rec):
        return {"loss_critic": rec.loss_critic, "loss_actor": rec.loss_actor, "ent_coef": rec.ent_coef}


class Iqn(Agent):
    _models = ("iqn", "iqn_tgt")

    @classmethod
    def build(cls, config: IqnConfig):
        c = config.to_c()
        h = C.c_void_p()
        L.check(L.lib().bb_iqn_create(C.byref(c), C.byref(h)))
        return cls(h, config)

    def _record(self, rec):
        return {"loss_critic": rec.loss_critic}
