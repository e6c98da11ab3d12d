"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (handle exchange order,
per-rank seeds, max-over-ranks timing reduction)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from border_b200 import dist as bd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank * 16 + (i % 16) for i in range(2 * bd.HANDLE_BYTES)])
    blobs = bd.gather_blobs(mine, dist, torch, "cpu")
    ok = all(blobs[r] == bytes([r * 16 + (i % 16) for i in range(2 * bd.HANDLE_BYTES)]) for r in range(world))
    mx = bd.max_over_ranks(10.0 + rank, dist, torch, "cpu")
    q.put((rank, ok, mx, bd.rank_seed(42, rank)))
    dist.destroy_process_group()


def test_gloo_world2_handle_exchange_and_timing_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]          # every rank sees all blobs in rank order
    assert [r[2] for r in res] == [11.0, 11.0]          # max over ranks
    assert [r[3] for r in res] == [42, 43]              # distinct replay seeds per rank
