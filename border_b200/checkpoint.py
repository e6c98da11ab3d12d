"""tch `VarStore` archives (`<model>.pt.tch`) read and written on the host.

The reference saves every model with `VarStore::save` (border-tch-agent/src/dqn/base.rs:348-362 -> qnet.pt.tch,
qnet_tgt.pt.tch; sac/base.rs:313-345 -> pi.pt.tch, qnet_i.pt.tch, ent_coef.pt.tch; iqn/base.rs likewise), which tch 0.16
implements as `torch::serialize::OutputArchive::write(name, tensor)` + `save_to(path)`, and loads them with
`torch::jit::load(path).named_parameters()`.  The archive is a TorchScript module zip, so it is read and written here with
`torch.jit` -- checkpointing is host plumbing, not the hot path; the C ABI stays torch-free (parameters cross it as plain
float arrays through bb_agent_get_param / bb_agent_set_param).  tests/test_checkpoint.py checks both directions against
the real libtorch calls (oracle/varstore_oracle.cpp).

What the reference's archives lack for a true resume -- Adam moments and step counts -- stays in this library's side-car
(`<model>.pt.tch.b200`, written by bb_agent_save_params).
"""
import os
from collections import OrderedDict

import numpy as np
import torch


class _Holder(torch.nn.Module):
    """A node of the VarStore name tree ("mlp.ln0.weight" -> mlp -> ln0 -> weight)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


def write_varstore(path, named_tensors):
    """Writes {name: ndarray} so that tch's `VarStore::load` (torch::jit::load + named_parameters()) reads the same names
    and values.  Names use '.' as tch's `nn::Path` does."""
    root = _Holder()
    for name, arr in named_tensors.items():
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                setattr(mod, p, _Holder())
            mod = getattr(mod, p)
        t = torch.from_numpy(np.ascontiguousarray(arr)).clone()
        mod.register_parameter(parts[-1], torch.nn.Parameter(t, requires_grad=t.is_floating_point()))
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    torch.jit.script(root).save(path)


def read_varstore(path):
    """Reads a `.pt.tch` archive (written by tch or by write_varstore) into an OrderedDict {name: ndarray}."""
    m = torch.jit.load(path, map_location="cpu")
    out = OrderedDict()
    for name, p in m.named_parameters():
        out[name] = p.detach().cpu().numpy().copy()
    for name, b in m.named_buffers():
        out.setdefault(name, b.detach().cpu().numpy().copy())
    return out
