"""Multi-GPU plumbing (one process per GPU, torch.distributed for the rendezvous only).

The data path has no NCCL call: ranks exchange CUDA-IPC handles of their gradient buffers once,
after which the fused all-reduce + Adam kernel reads every rank's gradients through peer pointers
over NVLink.  There is no reference twin: the reference's only
parallelism is actor threads.
"""
import ctypes as C

from . import _lib as L

HANDLE_BYTES = 64  # sizeof(cudaIpcMemHandle_t)


def rank_seed(base_seed, rank):
    """Replay / env seed of a rank: seed = actor id in the reference (actor_manager/base.rs:153,169)."""
    return base_seed + rank


def gather_blobs(mine: bytes, dist, torch, device):
    """all_gather of one fixed-size byte blob per rank, returned in rank order."""
    world = dist.get_world_size()
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [bytes(o.cpu().tolist()) for o in out]


def connect_gradient_peers(agent, dist, torch):
    """Maps every rank's gradient buffer and barrier flags into this process (CUDA IPC)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    h = (C.c_uint8 * HANDLE_BYTES)()
    f = (C.c_uint8 * HANDLE_BYTES)()
    L.check(L.lib().bb_agent_ipc_export(agent.handle, h, f))
    blobs = gather_blobs(bytes(h) + bytes(f), dist, torch, "cuda")
    hs = b"".join(b[:HANDLE_BYTES] for b in blobs)
    fs = b"".join(b[HANDLE_BYTES:] for b in blobs)
    L.check(L.lib().bb_agent_ipc_connect(agent.handle, rank, world, hs, fs))
    import os
    mode = os.environ.get("BB_GRAD_SYNC") or "overlapped"
    if mode in ("overlapped", "late"):
        return ("sharded mean over NVLink (each rank reduces 1/%d of a gradient region through CUDA-IPC peer loads and stores it to "
                "every rank; rendezvous flags folded into the reduce kernel and Adam's prologue%s)"
                % (world, "; FC region exchanged under the convolution backward" if mode == "overlapped" else ""))
    if mode == "sharded" or (mode == "legacy" and world >= 4):
        return "legacy: barrier + sharded mean over NVLink + barrier + local Adam"
    return "legacy: barrier + fused P2P all-reduce + Adam over NVLink + barrier"


def max_over_ranks(value, dist, torch, device):
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
