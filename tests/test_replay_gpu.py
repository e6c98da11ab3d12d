"""GPU parity: device replay ring (through the C ABI) vs the CPU oracle -- bit-exact indices,
gathers, priorities, IS weights."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from border_b200 import _lib as L
from border_b200.replay import (GenericTransitionBatch, PerConfig, SimpleReplayBuffer, SimpleReplayBufferConfig)
from oracle import replay_oracle as ro

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tr(rng, n, obs_shape, obs_dtype, act_shape, act_dtype, n_act=6):
    if np.dtype(obs_dtype) == np.uint8:
        obs = rng.integers(0, 256, (n,) + obs_shape, dtype=np.uint8)
        nxt = rng.integers(0, 256, (n,) + obs_shape, dtype=np.uint8)
    else:
        obs = rng.standard_normal((n,) + obs_shape).astype(np.float32)
        nxt = rng.standard_normal((n,) + obs_shape).astype(np.float32)
    if np.dtype(act_dtype) == np.int64:
        act = rng.integers(0, n_act, (n,) + act_shape).astype(np.int64)
    else:
        act = rng.uniform(-1, 1, (n,) + act_shape).astype(np.float32)
    return GenericTransitionBatch(obs, act, nxt, rng.standard_normal(n).astype(np.float32),
                                  (rng.random(n) < 0.1).astype(np.int8), (rng.random(n) < 0.05).astype(np.int8))


def _pair(capacity, seed, obs_shape, obs_dtype, act_shape, act_dtype, per=None):
    cfg = SimpleReplayBufferConfig(capacity=capacity, seed=seed, per_config=PerConfig(**per) if per else None)
    dev = SimpleReplayBuffer.build(cfg)
    orc = ro.ReplayOracle(capacity, seed, obs_shape, obs_dtype, act_shape, act_dtype, per=per)
    return dev, orc


def _same_batch(b, o):
    assert np.array_equal(b.ix_sample, o["ix_sample"])
    for k in ("obs", "act", "next_obs", "reward", "is_terminated", "is_truncated"):
        assert np.array_equal(getattr(b, k), o[k]), k


def test_device_powf_is_glibc_powf():
    rng = np.random.default_rng(0)
    n = 200000
    x = np.concatenate([rng.uniform(1e-8, 20, n // 2), np.exp(rng.uniform(-80, 80, n // 2))]).astype(np.float32)
    y = np.concatenate([rng.choice([0.6, 1 / 0.6, 0.4, -0.4, -1.0, 1.0, 0.5, 0.7], n // 2),
                        rng.uniform(-2, 2, n // 2)]).astype(np.float32)
    out = np.empty(n, np.float32)
    L.check(L.lib().bb_test_powf(0, x.ctypes.data, y.ctypes.data, out.ctypes.data, n))
    ref = np.array([ro.powf(a, b) for a, b in zip(x[:20000], y[:20000])], np.float32)
    assert np.array_equal(out[:20000].view(np.uint32), ref.view(np.uint32))
    ref2 = np.array([ro.powf(a, b) for a, b in zip(x[-20000:], y[-20000:])], np.float32)
    assert np.array_equal(out[-20000:].view(np.uint32), ref2.view(np.uint32))


@pytest.mark.parametrize("geom", [((4, 84, 84), np.uint8, (1,), np.int64), ((17,), np.float32, (8,), np.float32),
                                  ((5,), np.uint8, (1,), np.int64)])
def test_uniform_sample_gather_bit_exact_while_filling(geom):
    """base.rs:376-402 with growing size, ring wrap, ragged pushes."""
    obs_shape, obs_dtype, act_shape, act_dtype = geom
    cap = 97 if obs_shape == (4, 84, 84) else 1000
    dev, orc = _pair(cap, 42, obs_shape, obs_dtype, act_shape, act_dtype)
    rng = np.random.default_rng(1)
    for step in range(60):
        n = int(rng.integers(1, 9))
        tr = _tr(rng, n, obs_shape, obs_dtype, act_shape, act_dtype)
        dev.push(tr)
        orc.push(*tr.unpack()[:6])
        assert dev.len() == orc.len()
        B = int(rng.integers(1, 40))
        _same_batch(dev.batch(B), orc.batch(B))
    assert dev.state()["i"] == orc.head()


def test_uniform_indices_10k_draws_seed42_golden():
    import json
    gold = json.load(open(os.path.join(GOLD, "stdrng_seed42.json")))["words"]
    dev, orc = _pair(1 << 20, 42, (4,), np.float32, (1,), np.int64)
    dev.allocate((4,), np.float32, (1,), np.int64)
    dev.fill_synthetic(1 << 20)
    b = dev.batch(64)
    assert b.ix_sample.tolist() == [w % (1 << 20) for w in gold[:64]]
    rng = ro.StdRng(42)
    for _ in range(64):
        rng.next_u32()
    for B in (1, 7, 256, 1000, 4096):
        got = dev.batch(B).ix_sample
        want = np.array([rng.next_u32() % (1 << 20) for _ in range(B)], np.uint64)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("case_ix", [0, 1, 2, 3])
def test_uniform_ixs_golden_g2_sizes_on_device(case_ix):
    """SURVEY 8c G2 on the device: 10^4 draws of seed 42 over capacities 32 / 10 k / 262,144 / 2^20 while the ring fills and
    wraps -- every batch's indices equal the committed fixture (heads) and the SHA-256 over all draws matches."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(GOLD, "uniform_ixs_seed42.json")))
    case = gold["cases"][case_ix]
    cap = case["capacity"]
    dev = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=gold["seed"]))
    dev.allocate((2,), np.float32, (1,), np.int64)
    h = hashlib.sha256()
    row = 0
    for (push, size, B), head in zip(case["schedule"], case["heads"]):
        obs = np.arange(row, row + push, dtype=np.float32)[:, None].repeat(2, 1)
        dev.push(GenericTransitionBatch(obs, np.zeros((push, 1), np.int64), obs, np.zeros(push, np.float32),
                                        np.zeros(push, np.int8), np.zeros(push, np.int8)))
        row += push
        assert dev.len() == size
        b = dev.batch(B)
        assert b.ix_sample[:8].tolist() == head
        # the gathered row is the ring row the index names: row r holds the value of the LAST push that landed on r
        pos = b.ix_sample.astype(np.int64)
        newest = row - 1 - ((row - 1 - pos) % cap)
        assert np.array_equal(b.obs[:, 0], newest.astype(np.float32))
        h.update(np.ascontiguousarray(b.ix_sample, np.uint64).tobytes())
    assert h.hexdigest() == case["sha256"]
    dev.close()


def test_empty_buffer_and_zero_push():
    dev, _ = _pair(10, 1, (3,), np.float32, (1,), np.int64)
    dev.allocate((3,), np.float32, (1,), np.int64)
    with pytest.raises(L.BorderB200Error):
        dev.batch(4)
    z = np.zeros((0, 3), np.float32)
    dev.push(GenericTransitionBatch(z, np.zeros((0, 1), np.int64), z, np.zeros(0, np.float32), np.zeros(0, np.int8),
                                    np.zeros(0, np.int8)))
    assert dev.len() == 0


@pytest.mark.parametrize("normalize", ["All", "Batch"])
@pytest.mark.parametrize("capacity", [1000, 1024, 37])
def test_per_trace_bit_exact(normalize, capacity):
    """push -> sample (injected uniforms) -> update_priority, comparing the whole sum tree."""
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize=normalize)
    dev, orc = _pair(capacity, 42, (4,), np.float32, (1,), np.int64, per=per)
    rng = np.random.default_rng(2024)
    done = 0
    while done < int(capacity * 1.5):
        n = int(rng.integers(1, 8))
        tr = _tr(rng, n, (4,), np.float32, (1,), np.int64)
        dev.push(tr)
        orc.push(*tr.unpack()[:6])
        done += n
    t_dev, ns, _ = dev.dump_sum_tree()
    t_orc, ns_o = orc.sum_tree()
    assert ns == ns_o and np.array_equal(t_dev.view(np.uint32), t_orc.view(np.uint32))
    B = 64
    for k in range(40):
        u = rng.random(B, dtype=np.float32)
        dev.inject_uniforms(u)
        b = dev.batch(B)
        o = orc.batch(B, u)
        _same_batch(b, o)
        assert np.array_equal(b.weight.view(np.uint32), o["weight"].view(np.uint32)), k
        td = np.abs(rng.standard_normal(B)).astype(np.float32) * (3.0 if k % 3 else 0.01)
        ix = b.ix_sample.copy()
        if k % 5 == 0:
            ix[1::2] = ix[0::2][: len(ix[1::2])]  # duplicate indices inside one batch
        dev.update_priority(ix, td)
        orc.update_priority(ix, td)
        t_dev, ns, no = dev.dump_sum_tree()
        t_orc, _ = orc.sum_tree()
        assert np.array_equal(t_dev.view(np.uint32), t_orc.view(np.uint32)), k
        assert no == k + 1


def test_per_golden_fixture():
    gold = np.load(os.path.join(GOLD, "per_trace_all.npz"))
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize="All")
    dev, _ = _pair(1000, 42, (4,), np.float32, (1,), np.int64, per=per)
    rng = np.random.default_rng(2024)
    done = 0
    while done < 1500:
        n = int(rng.integers(1, 8))
        obs = rng.standard_normal((n, 4)).astype(np.float32)
        dev.push(GenericTransitionBatch(obs, rng.integers(0, 3, (n, 1)), obs + 1, rng.standard_normal(n).astype(np.float32),
                                        np.zeros(n, np.int8), np.zeros(n, np.int8)))
        done += n
    assert np.array_equal(dev.dump_sum_tree()[0], gold["tree_after_push"])
    for k in range(40):
        u = rng.random(64, dtype=np.float32)
        dev.inject_uniforms(u)
        b = dev.batch(64)
        assert np.array_equal(b.ix_sample, gold["ixs"][k]) and np.array_equal(b.weight, gold["ws"][k])
        td = np.abs(rng.standard_normal(64)).astype(np.float32) * (3.0 if k % 3 else 0.01)
        ix = b.ix_sample.copy()
        if k % 5 == 0:
            ix[1::2] = ix[0::2][: len(ix[1::2])]
        dev.update_priority(ix, td)
    assert np.array_equal(dev.dump_sum_tree()[0], gold["tree_final"])


def test_per_fastrand_stream_matches_oracle_wyrand():
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize="All")
    dev, orc = _pair(256, 3, (4,), np.float32, (1,), np.int64, per=per)
    rng = np.random.default_rng(5)
    tr = _tr(rng, 200, (4,), np.float32, (1,), np.int64)
    dev.push(tr)
    orc.push(*tr.unpack()[:6])
    for _ in range(3):
        _same_batch(dev.batch(32), orc.batch(32))


def test_update_priority_without_per_is_noop_and_errors_match():
    dev, _ = _pair(16, 1, (3,), np.float32, (1,), np.int64)
    dev.allocate((3,), np.float32, (1,), np.int64)
    dev.update_priority(None, None)  # base.rs:414: silently ignored without PER
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize="All")
    dev2, _ = _pair(16, 1, (3,), np.float32, (1,), np.int64, per=per)
    dev2.allocate((3,), np.float32, (1,), np.int64)
    with pytest.raises(L.BorderB200Error, match="ixs should be Some"):
        dev2.update_priority(None, np.zeros(1, np.float32))


def test_full_size_ring_properties():
    """BASELINE size (1M x 84x84x4 u8 = 56 GB): size-independent checks -- gathered rows equal the
    ring rows their indices name (checked through a second gather of the same indices via PER-free
    replay determinism) and indices are in range."""
    cap = 1 << 20
    cfg = SimpleReplayBufferConfig(capacity=cap, seed=42)
    dev = SimpleReplayBuffer.build(cfg)
    dev.allocate((4, 84, 84), np.uint8, (1,), np.int64)
    dev.fill_synthetic(cap, n_actions=6, seed=1234)
    assert dev.len() == cap
    b1 = dev.batch(256)
    assert b1.ix_sample.max() < cap
    rng = ro.StdRng(42)
    assert b1.ix_sample.tolist() == [rng.next_u32() % cap for _ in range(256)]
    # a second buffer with the same fill must hand back identical rows for identical indices
    dev.close()
    dev2 = SimpleReplayBuffer.build(cfg)
    dev2.allocate((4, 84, 84), np.uint8, (1,), np.int64)
    dev2.fill_synthetic(cap, n_actions=6, seed=1234)
    b2 = dev2.batch(256)
    assert np.array_equal(b1.obs, b2.obs) and np.array_equal(b1.next_obs, b2.next_obs)
    assert np.array_equal(b1.act, b2.act) and set(np.unique(b1.act)) <= set(range(6))
    # next_obs of row r is obs of row r+1 unless terminated (SURVEY 8d synthetic contract)
    assert b1.obs.dtype == np.uint8 and 100 < b1.obs.mean() < 155


@pytest.mark.parametrize("capacity,n,distinct", [(1000, 1024, 7), (1024, 1024, 1), (37, 300, 37), (4099, 1500, 50),
                                                 (4096, 2500, 4096), (1000, 1, 1)])
def test_update_priority_large_batches_and_heavy_duplicates_bit_exact(capacity, n, distinct):
    """SumTree::update in batch order (sum_tree.rs:93-107) on the device is a stable per-depth partition + parallel
    run folds: whole-tree bit equality with the sequential oracle for batches up to several 1024-chunks, all-equal
    indices, non-2^k capacities and a pushed ring (max-priority drift included)."""
    per = dict(alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=30, normalize="All")
    dev, orc = _pair(capacity, 7, (4,), np.float32, (1,), np.int64, per=per)
    rng = np.random.default_rng(capacity * 13 + n)
    tr = _tr(rng, capacity, (4,), np.float32, (1,), np.int64)
    dev.push(tr)
    orc.push(*tr.unpack()[:6])
    pool = rng.choice(capacity, size=min(distinct, capacity), replace=False)
    for k in range(3):
        ix = rng.choice(pool, size=n).astype(np.uint64)
        td = (rng.random(n) * (10.0 ** rng.integers(-3, 3))).astype(np.float32)
        dev.update_priority(ix, td)
        orc.update_priority(ix, td)
        t_dev, _, no = dev.dump_sum_tree()
        t_orc, _ = orc.sum_tree()
        assert np.array_equal(t_dev.view(np.uint32), t_orc.view(np.uint32)), (k, capacity, n, distinct)
        assert no == k + 1
    # a push after the updates re-derives the max priority from the max tree (sum_tree.rs:73-77)
    tr2 = _tr(rng, 5, (4,), np.float32, (1,), np.int64)
    dev.push(tr2)
    orc.push(*tr2.unpack()[:6])
    assert np.array_equal(dev.dump_sum_tree()[0].view(np.uint32), orc.sum_tree()[0].view(np.uint32))
