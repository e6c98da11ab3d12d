"""CPU: the optimizer restatement of the agent oracle (oracle/agent_oracle.py: CppAdam, written op by op after
torch/csrc/api/src/optim/adam.cpp) against the REAL `torch::optim::Adam` / `AdamW` that tch's nn::Adam / AdamW bind
(opt.rs:32-57), run through oracle/varstore_oracle.cpp: parameters and both moments after several steps, bit for bit."""
import os
import subprocess
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import agent_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "_build", "varstore_oracle")


def _real_adam(p0, grads, lr, b1, b2, eps, wd, adamw):
    if not os.path.exists(ORACLE):
        subprocess.check_call(["make", "-C", ROOT, "oracle/_build/varstore_oracle"])
    n, steps = p0.size, grads.shape[0]
    out = subprocess.run([ORACLE, "adam", str(n), str(steps), repr(lr), repr(b1), repr(b2), repr(eps), repr(wd), "1" if adamw else "0"],
                         input=p0.tobytes() + grads.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    a = np.frombuffer(out, np.float32)
    return a[:n], a[n:2 * n], a[2 * n:]


@pytest.mark.parametrize("cfg", [dict(lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, adamw=False),      # Adam::default(), opt.rs:35
                                 dict(lr=3e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.01, adamw=False),
                                 dict(lr=1e-3, b1=0.8, b2=0.99, eps=1e-6, wd=0.05, adamw=True)])
def test_cpp_adam_restatement_is_torch_optim_adam_bit_for_bit(cfg):
    rng = np.random.default_rng(3)
    n, steps = 4099, 7
    p0 = rng.standard_normal(n).astype(np.float32)
    # gradients over many magnitudes, incl. exact zeros and denormal-scale values (Adam's eps regime)
    grads = (rng.standard_normal((steps, n)) * 10.0 ** rng.integers(-12, 2, (steps, n))).astype(np.float32)
    grads[:, ::17] = 0.0
    p_ref, m_ref, v_ref = _real_adam(p0, grads, **cfg)
    params = OrderedDict(w=torch.from_numpy(p0.copy()).requires_grad_(True))
    opt = ao.CppAdam(params, cfg["lr"], cfg["b1"], cfg["b2"], cfg["eps"], cfg["wd"], cfg["adamw"])
    for s in range(steps):
        params["w"].grad = torch.from_numpy(grads[s].copy())
        opt.step()
    assert np.array_equal(params["w"].detach().numpy().view(np.uint32), p_ref.view(np.uint32))
    assert np.array_equal(opt.m["w"].numpy().view(np.uint32), m_ref.view(np.uint32))
    assert np.array_equal(opt.v["w"].numpy().view(np.uint32), v_ref.view(np.uint32))
