"""CPU: tch `VarStore` archives (`<model>.pt.tch`) -- this library's host-side reader / writer against the real libtorch calls
tch makes (oracle/varstore_oracle.cpp: OutputArchive::write + save_to, jit::load + named_parameters), both directions.
GPU: Agent::save_params / load_params write the reference's file names and tensor names and round-trip."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "_build", "varstore_oracle")


def _oracle():
    if not os.path.exists(ORACLE):
        subprocess.check_call(["make", "-C", ROOT, "oracle/_build/varstore_oracle"])
    return ORACLE


def _pattern(shape, seed):
    n = int(np.prod(shape))
    return np.sin((np.arange(n, dtype=np.float32) * np.float32(0.001)) + np.float32(seed)).astype(np.float32).reshape(shape)


SPEC = [("c1.weight", (32, 4, 8, 8)), ("c1.bias", (32,)), ("mlp.ln0.weight", (3, 5)), ("iqn_cos_to_feature.weight", (7, 4)),
        ("log_alpha", (1,))]


def test_reads_an_archive_written_the_way_tch_writes_it(tmp_path):
    from border_b200.checkpoint import read_varstore
    f = str(tmp_path / "qnet.pt.tch")
    args = [_oracle(), "save", f]
    for name, shape in SPEC:
        args += [name, str(len(shape))] + [str(d) for d in shape]
    subprocess.check_call(args)
    got = read_varstore(f)
    assert list(got.keys()) == [n for n, _ in SPEC]
    for i, (name, shape) in enumerate(SPEC):
        assert got[name].shape == shape
        # (torch's sin and numpy's may differ in the last bit)
        assert np.allclose(got[name], _pattern(shape, i + 1), rtol=0, atol=1e-6)


def test_tch_load_reads_an_archive_written_here(tmp_path):
    from border_b200.checkpoint import write_varstore
    f = str(tmp_path / "pi.pt.tch")
    tensors = {name: _pattern(shape, 10 + i) for i, (name, shape) in enumerate(SPEC)}
    write_varstore(f, tensors)
    out = subprocess.check_output([_oracle(), "load", f], text=True)
    seen = {}
    for line in out.strip().splitlines():
        t = line.split()
        nd = int(t[1])
        seen[t[0]] = (tuple(int(x) for x in t[2:2 + nd]), float(t[2 + nd]), float(t[3 + nd]), float(t[4 + nd]))
    assert set(seen) == set(tensors)
    for name, a in tensors.items():
        shape, s, first, last = seen[name]
        assert shape == a.shape
        assert abs(s - float(a.astype(np.float64).sum())) <= 1e-4 * max(1.0, abs(s))
        assert abs(first - float(a.flat[0])) < 1e-6 and abs(last - float(a.flat[-1])) < 1e-6


@pytest.mark.gpu
def test_agent_checkpoints_use_the_reference_file_and_tensor_names(tmp_path):
    from border_b200.agents import AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, OptimizerConfig
    from border_b200.checkpoint import read_varstore
    cfg = DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, 6), opt_config=OptimizerConfig(lr=1e-4)),
                    batch_size=32, train=True, device=0)
    a = Dqn.build(cfg)
    d = str(tmp_path / "nested" / "100")   # Trainer saves to save_dir/<opt_steps>: missing parents are created
    paths = a.save_params(d)
    assert sorted(os.path.basename(p) for p in paths) == ["qnet.pt.tch", "qnet_tgt.pt.tch"]   # dqn/base.rs:352-355
    names = list(read_varstore(paths[0]).keys())
    assert names == ["c1.weight", "c1.bias", "c2.weight", "c2.bias", "c3.weight", "c3.bias", "l1.weight", "l1.bias",
                     "l2.weight", "l2.bias"]                                                       # cnn/base.rs var names
    out = subprocess.check_output([_oracle(), "load", paths[0]], text=True)                        # tch's own load path
    assert [l.split()[0] for l in out.strip().splitlines()] == names
    b = Dqn.build(cfg.replace(init_seed=123) if hasattr(cfg, "replace") else cfg)
    b.load_params(d)
    pa, pb = a.named_parameters("qnet"), b.named_parameters("qnet")
    assert all(np.array_equal(pa[k], pb[k]) for k in pa)
    # a reference checkpoint has no side-car: parameters load, Adam state is reset
    for f in os.listdir(d):
        if f.endswith(".b200"):
            os.remove(os.path.join(d, f))
    c = Dqn.build(cfg)
    c.load_params(d)
    pc = c.named_parameters("qnet_tgt")
    assert all(np.array_equal(a.named_parameters("qnet_tgt")[k], pc[k]) for k in pc)
