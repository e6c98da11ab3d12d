"""CPU restatement of border-tch-agent's agents with the same ATen ops in the same order.

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module, as the checker / CPU baseline.  The product never does.

The reference's arithmetic lives in libtorch, reached through tch 0.16 (not under
/root/reference); this file calls the same operators through Python torch (fp32, CPU):
  AtariCnn            border-tch-agent/src/cnn/base.rs:23-36
  Mlp / Mlp2          border-tch-agent/src/mlp/base.rs:13-41, mlp/mlp2.rs:23-50
  Dqn::update_critic  border-tch-agent/src/dqn/base.rs:60-160, opt_ :182-200
  Optimizer           border-tch-agent/src/opt.rs:74-83 -> torch::optim::Adam (C++ frontend form)
  track               border-tch-agent/src/util.rs:31-45
  Sac                 border-tch-agent/src/sac/base.rs:73-198, ent_coef.rs:27-75
  Iqn                 border-tch-agent/src/iqn/base.rs:63-170, iqn/model/base.rs:162-234,
                      util/quantile_loss.rs:7-12
Parity pinning: the reference holds no golden vectors for these paths, so
the numerics are "parity unpinned" against Rust-tch itself; they are pinned by construction to
the ATen operators tch binds.  The C++-frontend Adam is restated op by op below because Python's
torch.optim.Adam uses a different (lerp / fused) formulation.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------ models

def atari_cnn_params(n_stack, out_dim, gen, skip_linear=False):
    """Named tensors in tch VarStore naming/layout (conv OIHW, linear [out,in])."""
    def u(shape, bound):
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    p = OrderedDict()
    for name, (o, i, k) in (("c1", (32, n_stack, 8)), ("c2", (64, 32, 4)), ("c3", (64, 64, 3))):
        fan = i * k * k
        p[name + ".weight"] = u((o, i, k, k), math.sqrt(6.0 / fan))
        p[name + ".bias"] = u((o,), 0.05)  # non-zero so that bias paths are exercised
    if not skip_linear:
        for name, (o, i) in (("l1", (512, 3136)), ("l2", (out_dim, 512))):
            p[name + ".weight"] = u((o, i), math.sqrt(6.0 / i))
            p[name + ".bias"] = u((o,), 1.0 / math.sqrt(i))
    return p


def atari_cnn_forward(p, x, skip_linear=False):
    """cnn/base.rs:23-36.  x: [B, n_stack, 1, 84, 84] or [B, n_stack, 84, 84], any dtype."""
    if x.dim() == 5:
        x = x.squeeze(2)
    x = x.to(torch.float32) / 255
    x = F.relu(F.conv2d(x, p["c1.weight"], p["c1.bias"], stride=4))
    x = F.relu(F.conv2d(x, p["c2.weight"], p["c2.bias"], stride=2))
    x = F.relu(F.conv2d(x, p["c3.weight"], p["c3.bias"], stride=1)).flatten(1)
    if skip_linear:
        return x
    x = F.relu(F.linear(x, p["l1.weight"], p["l1.bias"]))
    return F.linear(x, p["l2.weight"], p["l2.bias"])


def mlp_params(in_dim, units, out_dim, gen, prefix="mlp.ln"):
    p = OrderedDict()
    dims = [in_dim] + list(units) + [out_dim]
    for i in range(len(dims) - 1):
        bound = 1.0 / math.sqrt(dims[i])
        p["%s%d.weight" % (prefix, i)] = (torch.rand((dims[i + 1], dims[i]), generator=gen) * 2 - 1) * math.sqrt(6.0 / dims[i])
        p["%s%d.bias" % (prefix, i)] = (torch.rand((dims[i + 1],), generator=gen) * 2 - 1) * bound
    return p


def mlp_forward(p, x, n_layers, activation_out=False, prefix="mlp.ln"):
    """mlp/base.rs:13-41: Linear+ReLU per hidden layer, final Linear (+ReLU iff activation_out)."""
    for i in range(n_layers):
        x = F.linear(x, p["%s%d.weight" % (prefix, i)], p["%s%d.bias" % (prefix, i)])
        if i < n_layers - 1 or activation_out:
            x = F.relu(x)
    return x


# ------------------------------------------------------------------------------------ optimizer

class CppAdam:
    """torch::optim::Adam / AdamW step (torch/csrc/api/src/optim/adam.cpp, adamw.cpp), which is
    what tch::nn::Adam / AdamW bind (opt.rs:32-57).  Adam::default(): betas (0.9, 0.999), eps 1e-8,
    wd 0."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, adamw=False):
        self.params = params  # OrderedDict name -> leaf tensor (requires_grad)
        self.lr, self.b1, self.b2, self.eps, self.wd, self.adamw = lr, beta1, beta2, eps, wd, adamw
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def zero_grad(self):
        for v in self.params.values():
            v.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        for k, p in self.params.items():
            if p.grad is None:
                continue
            g = p.grad
            if self.adamw:
                p.mul_(1 - self.lr * self.wd)
            elif self.wd != 0:
                g = g.add(p, alpha=self.wd)
            m, v = self.m[k], self.v[k]
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-(self.lr / bc1))

    def backward_step(self, loss):  # tch nn::Optimizer::backward_step: zero_grad; backward; step
        self.zero_grad()
        loss.backward()
        self.step()


@torch.no_grad()
def track(dest, src, tau):
    """util.rs:31-45: dest.copy_(tau * src + (1.0 - tau) * dest) per named variable."""
    for k in src:
        dest[k].copy_(tau * src[k] + (1.0 - tau) * dest[k])


def leafify(p):
    return OrderedDict((k, v.clone().detach().requires_grad_(True)) for k, v in p.items())


# ------------------------------------------------------------------------------------ DQN

class DqnOracle:
    """Dqn (dqn/base.rs) for Q in {AtariCnn, Mlp}."""

    def __init__(self, params, forward, lr, batch_size, discount_factor=0.99, tau=0.005, soft_update_interval=1,
                 n_updates_per_opt=1, double_dqn=False, clip_td_err=None, critic_loss="Mse", opt_kwargs=None):
        self.forward = forward
        self.qnet = leafify(params)
        self.qnet_tgt = OrderedDict((k, v.clone().detach()) for k, v in params.items())
        self.opt = CppAdam(self.qnet, lr, **(opt_kwargs or {}))
        self.batch_size, self.gamma, self.tau = batch_size, discount_factor, tau
        self.soft_update_interval, self.n_updates_per_opt = soft_update_interval, n_updates_per_opt
        self.soft_update_counter = 0
        self.double_dqn, self.clip_td_err, self.critic_loss = double_dqn, clip_td_err, critic_loss
        self.n_opts = 0

    def update_critic(self, batch):
        """batch: dict(obs, act [B,1] int64, next_obs, reward [B] f32, is_terminated [B] int8, weight or None).
        Returns (loss, td_errs or None) as in dqn/base.rs:60-160."""
        obs, act, next_obs = batch["obs"], batch["act"], batch["next_obs"]
        reward, is_terminated, weight = batch["reward"], batch["is_terminated"], batch.get("weight")
        pred = self.forward(self.qnet, obs).gather(-1, act).squeeze()
        with torch.no_grad():
            if self.double_dqn:
                x = self.forward(self.qnet, next_obs)
                y = x.argmax(-1, False).unsqueeze(-1)
                q = self.forward(self.qnet_tgt, next_obs).gather(-1, y).squeeze()
            else:
                x = self.forward(self.qnet_tgt, next_obs)
                y = x.argmax(-1, False).unsqueeze(-1)
                q = x.gather(-1, y).squeeze()
            tgt = reward + (1 - is_terminated) * self.gamma * q
        td = None
        if weight is not None:
            td_errs = (pred - tgt).abs()
            if self.clip_td_err is not None:
                td_errs = td_errs.clip(self.clip_td_err[0], self.clip_td_err[1])
            loss = weight * td_errs
            zeros = torch.zeros(len(weight))
            loss = F.smooth_l1_loss(loss, zeros, reduction="mean", beta=1.0) if self.critic_loss == "SmoothL1" \
                else F.mse_loss(loss, zeros, reduction="mean")
            self.opt.backward_step(loss)
            td = td_errs.detach().clone()
        else:
            loss = F.smooth_l1_loss(pred, tgt, reduction="mean", beta=1.0) if self.critic_loss == "SmoothL1" \
                else F.mse_loss(pred, tgt, reduction="mean")
            self.opt.backward_step(loss)
        self.last = dict(pred=pred.detach(), tgt=tgt)
        return float(loss.detach()), td

    def opt_(self, sample_fn, update_priority_fn=None):
        """dqn/base.rs:182-200. sample_fn() -> batch dict (with 'ix_sample'); returns last loss."""
        loss = None
        for _ in range(self.n_updates_per_opt):
            batch = sample_fn()
            loss, td = self.update_critic(batch)
            if td is not None and update_priority_fn is not None:
                update_priority_fn(batch["ix_sample"], td.numpy())
        self.soft_update_counter += 1
        if self.soft_update_counter == self.soft_update_interval:
            self.soft_update_counter = 0
            track(self.qnet_tgt, self.qnet, self.tau)
        self.n_opts += 1
        return loss


# ------------------------------------------------------------------------------------ SAC

def mlp2_params(in_dim, units, out_dim, gen):
    """Mlp2 (mlp/mlp2.rs:31-50): trunk mlp.al{i}, heads ml / sl."""
    p = OrderedDict()
    dims = [in_dim] + list(units)
    for i in range(len(units)):
        p["mlp.al%d.weight" % i] = (torch.rand((dims[i + 1], dims[i]), generator=gen) * 2 - 1) * math.sqrt(6.0 / dims[i])
        p["mlp.al%d.bias" % i] = (torch.rand((dims[i + 1],), generator=gen) * 2 - 1) / math.sqrt(dims[i])
    h = dims[-1]
    for name in ("ml", "sl"):
        p[name + ".weight"] = (torch.rand((out_dim, h), generator=gen) * 2 - 1) * math.sqrt(1.0 / h)
        p[name + ".bias"] = (torch.rand((out_dim,), generator=gen) * 2 - 1) / math.sqrt(h)
    return p


def mlp2_forward(p, x, n_hidden):
    """mlp2.rs:23-28: (head1(x), exp(head2(x)))."""
    for i in range(n_hidden):
        x = F.relu(F.linear(x, p["mlp.al%d.weight" % i], p["mlp.al%d.bias" % i]))
    return F.linear(x, p["ml.weight"], p["ml.bias"]), F.linear(x, p["sl.weight"], p["sl.bias"]).exp()


def normal_logp(x):
    """sac/base.rs:25-29"""
    tmp = torch.tensor(-0.5 * math.log(2.0 * 3.1415927410125732), dtype=torch.float32) - 0.5 * x.pow(2)
    return tmp.sum(-1)


class SacOracle:
    """Sac (sac/base.rs:73-198) with Actor = Mlp2, Critic = Mlp over cat[obs, act]."""

    def __init__(self, pi_params, q_params_list, n_pi_hidden, n_q_layers, lr_pi, lr_q, batch_size, gamma=0.99, tau=0.005,
                 ent_coef_mode=("Fix", 1.0), epsilon=1e-4, min_lstd=-20.0, max_lstd=2.0, reward_scale=1.0,
                 critic_loss="Mse"):
        self.pi = leafify(pi_params)
        self.qnets = [leafify(q) for q in q_params_list]
        self.qnets_tgt = [OrderedDict((k, v.clone().detach()) for k, v in q.items()) for q in q_params_list]
        self.opt_pi = CppAdam(self.pi, lr_pi)
        self.opt_q = [CppAdam(q, lr_q) for q in self.qnets]
        self.n_pi_hidden, self.n_q_layers = n_pi_hidden, n_q_layers
        self.batch_size, self.gamma, self.tau = batch_size, gamma, tau
        self.epsilon, self.min_lstd, self.max_lstd = epsilon, min_lstd, max_lstd
        self.reward_scale, self.critic_loss = reward_scale, critic_loss
        if ent_coef_mode[0] == "Fix":  # ent_coef.rs:31-36
            self.log_alpha = torch.tensor([math.log(ent_coef_mode[1])], dtype=torch.float32)
            self.target_entropy, self.opt_alpha = None, None
        else:  # Auto(target_entropy, lr), ent_coef.rs:37-45
            self.log_alpha = torch.zeros(1, dtype=torch.float32, requires_grad=True)
            self.target_entropy = ent_coef_mode[1]
            self.opt_alpha = CppAdam(OrderedDict(log_alpha=self.log_alpha), ent_coef_mode[2])
        self.n_opts = 0

    def alpha(self):
        return self.log_alpha.detach().exp()

    def action_logp(self, o, z):
        mean, lstd = mlp2_forward(self.pi, o, self.n_pi_hidden)
        std = lstd.clip(self.min_lstd, self.max_lstd).exp()
        a = (std * z + mean).tanh()
        log_p = normal_logp(z) - (torch.tensor(1.0) - a.pow(2.0) + torch.tensor(self.epsilon, dtype=torch.float32)).log().sum(-1)
        return a, log_p

    def qvals(self, qnets, obs, act):
        return [mlp_forward(q, torch.cat([obs, act], -1), self.n_q_layers).squeeze() for q in qnets]

    def update_actor(self, batch, z):
        a, log_p = self.action_logp(batch["obs"], z)
        if self.target_entropy is not None:  # ent_coef.update(&log_p.detach())
            loss_a = -(self.log_alpha * (log_p.detach() + torch.tensor(self.target_entropy, dtype=torch.float64)).detach()).mean(dtype=torch.float32)
            self.opt_alpha.backward_step(loss_a)
        qval = torch.vstack(self.qvals(self.qnets, batch["obs"], a)).min(0)[0]
        loss = (self.alpha() * log_p - qval).mean(dtype=torch.float32)
        self.opt_pi.backward_step(loss)
        return float(loss.detach())

    def update_critic(self, batch, z):
        reward, is_terminated = batch["reward"], batch["is_terminated"]
        preds = self.qvals(self.qnets, batch["obs"], batch["act"])
        with torch.no_grad():
            next_a, next_log_p = self.action_logp(batch["next_obs"], z)
            next_q = torch.vstack(self.qvals(self.qnets_tgt, batch["next_obs"], next_a)).min(0)[0]
            next_q = next_q - self.alpha() * next_log_p
        tgt = self.reward_scale * reward + (1.0 - is_terminated) * torch.tensor(self.gamma, dtype=torch.float64) * next_q
        tgt = tgt.to(torch.float32)
        if self.critic_loss == "Mse":
            losses = [F.mse_loss(p, tgt, reduction="mean") for p in preds]
        else:
            losses = [F.smooth_l1_loss(p, tgt, reduction="mean", beta=1.0) for p in preds]
        for opt, loss in zip(self.opt_q, losses):
            opt.backward_step(loss)
        return sum(float(l.detach()) for l in losses) / len(losses)

    def opt_(self, batch, z_actor, z_critic):
        """one update of sac/base.rs:175-198 (n_updates_per_opt = 1)"""
        la = self.update_actor(batch, z_actor)
        lc = self.update_critic(batch, z_critic)
        for qt, q in zip(self.qnets_tgt, self.qnets):
            track(qt, q, self.tau)
        self.n_opts += 1
        return dict(loss_critic=lc, loss_actor=la, ent_coef=float(self.alpha()[0]))


# ------------------------------------------------------------------------------------ IQN

def iqn_params(f_params, feature_dim, embed_dim, m_params, gen):
    """One VarStore: psi vars, iqn_cos_to_feature, merge-net vars (iqn/model/base.rs:64-86)."""
    p = OrderedDict(f_params)
    p["iqn_cos_to_feature.weight"] = (torch.rand((feature_dim, embed_dim), generator=gen) * 2 - 1) * math.sqrt(6.0 / embed_dim)
    p["iqn_cos_to_feature.bias"] = (torch.rand((feature_dim,), generator=gen) * 2 - 1) / math.sqrt(embed_dim)
    p.update(m_params)
    return p


def iqn_forward(p, psi_fn, m_fn, x, tau, embed_dim):
    """IqnModel::forward (iqn/model/base.rs:198-234) with cos_embed_nn (:162-191)."""
    psi = psi_fn(p, x)
    B, N = tau.shape
    i = torch.arange(1, embed_dim + 1, dtype=torch.float32).unsqueeze(0).unsqueeze(0)
    cos = torch.cos(tau.unsqueeze(-1) * (math.pi * i)).reshape(-1, embed_dim)
    phi = F.relu(F.linear(cos, p["iqn_cos_to_feature.weight"], p["iqn_cos_to_feature.bias"]))
    phi = phi.reshape(B, N, -1)
    m = psi.unsqueeze(1) * phi
    return m_fn(p, m)


def quantile_huber_loss(x, tau):
    """util/quantile_loss.rs:7-12"""
    lt_0 = x.lt(0.0).detach()
    loss = F.smooth_l1_loss(x, torch.zeros_like(x), reduction="none", beta=1.0)
    return (tau - torch.where(lt_0, 1.0, 0.0)).abs() * loss


class IqnOracle:
    """Iqn::update_critic / opt_ (iqn/base.rs:63-190)."""

    def __init__(self, params, psi_fn, m_fn, embed_dim, lr, batch_size, discount_factor=0.99, tau=0.005,
                 soft_update_interval=1):
        self.iqn = leafify(params)
        self.iqn_tgt = OrderedDict((k, v.clone().detach()) for k, v in params.items())
        self.opt = CppAdam(self.iqn, lr)
        self.psi_fn, self.m_fn, self.embed_dim = psi_fn, m_fn, embed_dim
        self.batch_size, self.gamma, self.tau = batch_size, discount_factor, tau
        self.soft_update_interval, self.soft_update_counter = soft_update_interval, 0

    def update_critic(self, batch, tau_pred, tau_tgt):
        obs, act, next_obs = batch["obs"], batch["act"], batch["next_obs"]
        reward = batch["reward"].unsqueeze(-1)
        is_terminated = batch["is_terminated"].unsqueeze(-1)
        n_pred, n_tgt = tau_pred.shape[1], tau_tgt.shape[1]
        z = iqn_forward(self.iqn, self.psi_fn, self.m_fn, obs, tau_pred, self.embed_dim)
        a = act.unsqueeze(1).repeat(1, n_pred, 1)
        pred = z.gather(-1, a).squeeze(-1).unsqueeze(1)
        with torch.no_grad():
            zt = iqn_forward(self.iqn_tgt, self.psi_fn, self.m_fn, next_obs, tau_tgt, self.embed_dim)
            y = zt.clone().mean(1)
            a = y.argmax(-1, False).unsqueeze(-1).unsqueeze(-1).repeat(1, n_tgt, 1)
            zt = zt.gather(2, a).squeeze(-1)
            tgt = (reward + (1 - is_terminated) * self.gamma * zt).unsqueeze(-1)
        diff = tgt - pred
        tau = tau_pred.unsqueeze(1).repeat(1, n_tgt, 1)
        loss = quantile_huber_loss(diff, tau).mean(dtype=torch.float32)
        self.opt.backward_step(loss)
        return float(loss.detach())

    def opt_(self, batch, tau_pred, tau_tgt):
        loss = self.update_critic(batch, tau_pred, tau_tgt)
        self.soft_update_counter += 1
        if self.soft_update_counter == self.soft_update_interval:
            self.soft_update_counter = 0
            track(self.iqn_tgt, self.iqn, self.tau)
        return loss


# ------------------------------------------------------------------------------------ explorers

class FastRandPy:
    """fastrand 1.x wyrand (the global RNG the reference's explorers draw from: dqn/explorer.rs:71,82, dqn/base.rs:231-233),
    restated in pure Python integers: s += 0xA0761D6478BD642F; t = s * (s ^ 0xE7037ED1A0B428DB) (128 bit); out = lo ^ hi.
    f64() = from_bits(0x3FF0.. | (u64 >> 12)) - 1; f32() likewise on the low 32 bits; u32(..n) / u64(..n) = Lemire's
    multiply-high with rejection.  Seeded explicitly here (the reference never seeds it: no run-to-run vector exists)."""
    M64 = (1 << 64) - 1

    def __init__(self, seed):
        self.s = seed & self.M64

    def u64(self):
        self.s = (self.s + 0xA0761D6478BD642F) & self.M64
        t = self.s * (self.s ^ 0xE7037ED1A0B428DB)
        return (t & self.M64) ^ (t >> 64)

    def u32(self):
        return self.u64() & 0xFFFFFFFF

    def f64(self):
        import struct
        return struct.unpack("<d", struct.pack("<Q", 0x3FF0000000000000 | (self.u64() >> 12)))[0] - 1.0

    def f32(self):
        import struct
        import numpy as np
        return float(np.float32(struct.unpack("<f", struct.pack("<I", 0x3F800000 | (self.u32() >> 9)))[0]) - np.float32(1.0))

    def _below(self, n, bits):
        mask = (1 << bits) - 1
        draw = self.u32 if bits == 32 else self.u64
        x = draw()
        m = x * n
        hi, lo = m >> bits, m & mask
        if lo < n:
            t = ((1 << bits) - n) % n
            while lo < t:
                x = draw()
                m = x * n
                hi, lo = m >> bits, m & mask
        return hi

    def u32_below(self, n):
        return self._below(n, 32)

    def u64_below(self, n):
        return self._below(n, 64)


class EpsilonGreedyOracle:
    """dqn/explorer.rs:33-90 (`EpsilonGreedy::action`) and the eval branch of `Dqn::sample` (dqn/base.rs:229-236)."""

    def __init__(self, fr, eps_start=1.0, eps_final=0.02, final_step=100_000):
        self.fr, self.n_opts = fr, 0
        self.eps_start, self.eps_final, self.final_step = eps_start, eps_final, final_step

    def action(self, q):
        """q: [n_procs, n_actions] tensor of action values."""
        d = (self.eps_start - self.eps_final) / float(self.final_step)
        eps = max(self.eps_start - d * float(self.n_opts), self.eps_final)
        r = self.fr.f64()
        is_random = r < eps
        self.n_opts += 1
        best = [int(x) for x in q.argmax(-1)]
        if is_random:
            return [self.fr.u32_below(q.shape[1]) for _ in range(q.shape[0])]
        return best

    def eval_action(self, q):
        out = []
        for i in range(q.shape[0]):
            if self.fr.f32() < 0.01:
                out.append(self.fr.u64_below(q.shape[1]))
            else:
                out.append(int(q[i].argmax(-1)))
        return out
