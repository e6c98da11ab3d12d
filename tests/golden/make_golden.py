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
    main()
