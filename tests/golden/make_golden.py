"""Generates the committed golden fixtures from the CPU oracle.

The reference ships no vectors for these paths (only test_sum_tree_odd), and the Rust
reference cannot be built in this image, so these are oracle-derived regression fixtures:
    python -m tests.golden.make_golden
"""
import json
import os

import numpy as np

from oracle import replay_oracle as ro

HERE = os.path.dirname(os.path.abspath(__file__))


def run_per_trace(normalize, capacity=1000, pushes=1500, rounds=40, batch=64):
    """pushes (with wrap) then rounds x (sample with injected uniforms -> update_priority)."""
    rng = np.random.default_rng(2024)
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize=normalize)
    r = ro.ReplayOracle(capacity, 42, (4,), np.float32, (1,), np.int64, per=per)
    out = {}
    done = 0
    while done < pushes:
        n = int(rng.integers(1, 8))
        obs = rng.standard_normal((n, 4)).astype(np.float32)
        r.push(obs, rng.integers(0, 3, (n, 1)), obs + 1, rng.standard_normal(n).astype(np.float32),
               np.zeros(n, np.int8), np.zeros(n, np.int8))
        done += n
    out["tree_after_push"], _ = r.sum_tree()
    ixs, ws, trees = [], [], []
    for k in range(rounds):
        u = rng.random(batch, dtype=np.float32)
        ix, w = r.sample_indices(batch, u)
        td = np.abs(rng.standard_normal(batch)).astype(np.float32) * (3.0 if k % 3 else 0.01)
        if k % 5 == 0:
            ix2 = ix.copy()
            ix2[1::2] = ix2[0::2][: len(ix2[1::2])]  # duplicates inside one update batch
            r.update_priority(ix2, td)
        else:
            r.update_priority(ix, td)
        ixs.append(ix); ws.append(w); trees.append(r.sum_tree()[0][:64].copy())
    out["ixs"] = np.stack(ixs); out["ws"] = np.stack(ws); out["tree_heads"] = np.stack(trees)
    out["tree_final"], _ = r.sum_tree()
    return out


def run_uniform_ixs(seed=42):
    """`ixs[k] = next_u32() % size` (base.rs:384-387) for seed 42 over rings of capacity 32, 10 k, 262,144 and
    2^20 -- 10^4 draws each, taken as batches WHILE the ring fills (size grows between batches, then the ring wraps).  Stored
    per capacity: the (size, batch) schedule, the first 8 indices of every batch and a SHA-256 over all draws."""
    import hashlib
    out = {"seed": seed, "cases": []}
    for cap in (32, 10000, 262144, 1 << 20):
        rng = ro.StdRng(seed)
        size, drawn, sched, heads, h = 0, 0, [], [], hashlib.sha256()
        k = 0
        while drawn < 10000:
            push = [1, 7, cap // 3 + 1, 5, cap][k % 5]           # grow (and wrap) the ring between batches
            size = min(cap, size + push)
            B = [1, 32, 256, 1000, 64][k % 5]
            B = min(B, 10000 - drawn)
            ix = np.array([rng.next_u32() % size for _ in range(B)], np.uint64)
            sched.append([push, size, B]); heads.append(ix[:8].tolist()); h.update(ix.tobytes())
            drawn += B
            k += 1
        out["cases"].append({"capacity": cap, "schedule": sched, "heads": heads, "sha256": h.hexdigest()})
    return out


def dqn_mlp_setup():
    """The fixed DQN-MLP case of the agent fixture: parameters, ring contents and the oracle agent (seeds only)."""
    import torch
    from oracle import agent_oracle as ao
    rng = np.random.default_rng(7)
    gen = torch.Generator().manual_seed(0)
    params = ao.mlp_params(4, [64, 64], 2, gen)
    n, cap = 280, 300
    obs = rng.standard_normal((n, 4)).astype(np.float32)
    nxt = rng.standard_normal((n, 4)).astype(np.float32)
    act = rng.integers(0, 2, (n, 1)).astype(np.int64)
    reward = rng.standard_normal(n).astype(np.float32)
    term = (rng.random(n) < 0.2).astype(np.int8)
    trunc = np.zeros(n, np.int8)
    return params, (obs, act, nxt, reward, term, trunc), cap


def run_dqn_mlp_trace(steps=5, B=32, lr=1e-3):
    """DQN (dqn/base.rs:60-200) on an Mlp[64,64] Q net: SmoothL1, double DQN, soft update every 2 opts with tau 0.5,
    uniform replay seed 42 -- losses, sampled indices and the parameters after `steps` updates."""
    import torch
    from oracle import agent_oracle as ao
    params, tr, cap = dqn_mlp_setup()
    orc = ro.ReplayOracle(cap, 42, (4,), np.float32, (1,), np.int64)
    orc.push(*tr)
    oracle = ao.DqnOracle(params, lambda p, x: ao.mlp_forward(p, x, 3), lr, B, 0.99, 0.5, 2, 1, True, None, "SmoothL1")
    losses, ixs = [], []

    def sample():
        b = orc.batch(B)
        ixs.append(np.asarray(b["ix_sample"], dtype=np.uint64))
        return dict(obs=torch.from_numpy(b["obs"]), act=torch.from_numpy(b["act"]), next_obs=torch.from_numpy(b["next_obs"]),
                    reward=torch.from_numpy(b["reward"]), is_terminated=torch.from_numpy(b["is_terminated"]),
                    ix_sample=b["ix_sample"])

    for _ in range(steps):
        losses.append(oracle.opt_(sample))
    out = {"losses": np.asarray(losses, np.float64), "ixs": np.stack(ixs)}
    for k, v in oracle.qnet.items():
        out["qnet." + k] = v.detach().numpy().copy()
    for k, v in oracle.qnet_tgt.items():
        out["qnet_tgt." + k] = v.detach().numpy().copy()
    return out


def _digest(prefix, params, out):
    """Compact, order-free summary of a parameter dict: sum, sum of |x| (float64) and the first 32 values of every tensor."""
    for k, v in params.items():
        a = v.detach().numpy().astype(np.float64).ravel()
        out["%s.%s.sum" % (prefix, k)] = np.array([a.sum(), np.abs(a).sum()])
        out["%s.%s.head" % (prefix, k)] = a[:32].astype(np.float32)


def sac_setup(n_critics=2, units=(64, 64), seed=0, obs_dim=17, act_dim=8):
    """SURVEY 8c G7 set-up (oracle side only): parameters, ring contents."""
    import torch
    from oracle import agent_oracle as ao
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    n = 580
    obs = rng.standard_normal((n, obs_dim)).astype(np.float32)
    tr = (obs, rng.uniform(-1, 1, (n, act_dim)).astype(np.float32), rng.standard_normal((n, obs_dim)).astype(np.float32),
          rng.standard_normal(n).astype(np.float32), (rng.random(n) < 0.1).astype(np.int8), np.zeros(n, np.int8))
    pi_p = ao.mlp2_params(obs_dim, list(units), act_dim, gen)
    q_ps = [ao.mlp_params(obs_dim + act_dim, list(units), 1, gen) for _ in range(n_critics)]
    return rng, tr, pi_p, q_ps


def run_sac_trace(steps=3, B=64, lr=3e-4):
    """G7 (+ G5 at tau = 0.005): SAC (sac/base.rs:107-198) with two critics, EntCoefMode::Auto(-8, 3e-4), reward_scale 1.5,
    injected z / z' (the double-exp std quirk is in the oracle's action_logp): losses, alpha, injected noise and parameter
    digests of pi, both critics, both targets and log_alpha after `steps` updates."""
    import torch
    from oracle import agent_oracle as ao
    rng, tr, pi_p, q_ps = sac_setup()
    cap = 600
    orc = ro.ReplayOracle(cap, 42, (17,), np.float32, (8,), np.float32)
    orc.push(*tr)
    mode = ("Auto", -8.0, 3e-4)
    oracle = ao.SacOracle(pi_p, q_ps, 2, 3, lr, lr, B, 0.99, 0.005, mode, reward_scale=1.5, critic_loss="Mse")
    out = {"z1": [], "z2": [], "loss_critic": [], "loss_actor": [], "ent_coef": []}
    for _ in range(steps):
        z1 = rng.standard_normal((B, 8)).astype(np.float32)
        z2 = rng.standard_normal((B, 8)).astype(np.float32)
        b = orc.batch(B)
        tb = dict(obs=torch.from_numpy(b["obs"]), act=torch.from_numpy(b["act"]), next_obs=torch.from_numpy(b["next_obs"]),
                  reward=torch.from_numpy(b["reward"]), is_terminated=torch.from_numpy(b["is_terminated"]))
        ref = oracle.opt_(tb, torch.from_numpy(z1), torch.from_numpy(z2))
        out["z1"].append(z1); out["z2"].append(z2)
        for k in ("loss_critic", "loss_actor", "ent_coef"):
            out[k].append(float(ref[k]))
    out = {k: np.asarray(v) for k, v in out.items()}
    _digest("pi", oracle.pi, out)
    for i in range(2):
        _digest("qnet_%d" % i, oracle.qnets[i], out)
        _digest("qnet_tgt_%d" % i, oracle.qnets_tgt[i], out)
    out["log_alpha"] = oracle.log_alpha.detach().numpy().copy()
    return out


def iqn_setup(n_act=4):
    """SURVEY 8c G6 set-up (oracle side only): AtariCnn.skip_linear features, merge Mlp(3136 -> 512 -> A)."""
    import torch
    from oracle import agent_oracle as ao
    rng = np.random.default_rng(11)
    gen = torch.Generator().manual_seed(3)
    f_params = ao.atari_cnn_params(4, 0, gen, skip_linear=True)
    m_params = ao.mlp_params(3136, [512], n_act, gen)
    params = ao.iqn_params(f_params, 3136, 64, m_params, gen)
    n = 180
    obs = rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)
    nxt = rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)
    tr = (obs, rng.integers(0, n_act, (n, 1)).astype(np.int64), nxt, rng.standard_normal(n).astype(np.float32),
          (rng.random(n) < 0.2).astype(np.int8), np.zeros(n, np.int8))
    return rng, tr, params


def run_iqn_trace(steps=2, B=16, N=8, lr=1e-4):
    """G6: IQN update (iqn/base.rs:63-170) with injected tau / tau' [B][N]: losses, the injected percent points and parameter
    digests of the online and target models after `steps` updates (soft update every 2 opts, tau 0.5)."""
    import torch
    from oracle import agent_oracle as ao
    rng, tr, params = iqn_setup()
    orc = ro.ReplayOracle(200, 9, (4, 84, 84), np.uint8, (1,), np.int64)
    orc.push(*tr)
    psi_fn = lambda p, x: ao.atari_cnn_forward(p, x, skip_linear=True)
    m_fn = lambda p, m: ao.mlp_forward(p, m, 2)
    oracle = ao.IqnOracle(params, psi_fn, m_fn, 64, lr, B, 0.99, 0.5, 2)
    out = {"t1": [], "t2": [], "loss": []}
    for _ in range(steps):
        t1 = rng.random((B, N), dtype=np.float32)
        t2 = rng.random((B, N), dtype=np.float32)
        b = orc.batch(B)
        tb = dict(obs=torch.from_numpy(b["obs"]), act=torch.from_numpy(b["act"]), next_obs=torch.from_numpy(b["next_obs"]),
                  reward=torch.from_numpy(b["reward"]), is_terminated=torch.from_numpy(b["is_terminated"]))
        out["loss"].append(float(oracle.opt_(tb, torch.from_numpy(t1), torch.from_numpy(t2))))
        out["t1"].append(t1); out["t2"].append(t2)
    out = {k: np.asarray(v) for k, v in out.items()}
    _digest("iqn", oracle.iqn, out)
    _digest("iqn_tgt", oracle.iqn_tgt, out)
    return out


def write_agent_traces():
    np.savez_compressed(os.path.join(HERE, "sac_trace.npz"), **run_sac_trace())
    np.savez_compressed(os.path.join(HERE, "iqn_trace.npz"), **run_iqn_trace())


def main():
    np.savez_compressed(os.path.join(HERE, "dqn_mlp_trace.npz"), **run_dqn_mlp_trace())
    r = ro.StdRng(42)
    json.dump({"seed": 42, "words": [r.next_u32() for _ in range(64)]},
              open(os.path.join(HERE, "stdrng_seed42.json"), "w"))
    for norm in ("All", "Batch"):
        np.savez_compressed(os.path.join(HERE, "per_trace_%s.npz" % norm.lower()), **run_per_trace(norm))


def write_uniform_ixs():
    json.dump(run_uniform_ixs(), open(os.path.join(HERE, "uniform_ixs_seed42.json"), "w"))


if __name__ == "__main__":
    write_uniform_ixs()
    write_agent_traces()
    main()
