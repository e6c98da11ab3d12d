"""CPU: a discrete-event model of the data-parallel gradient exchange (nn.cu: grad_exchange_kernel,
grad_exchange_ll_kernel, adam_kernel's prologue wait) under random interleavings of W ranks.

The device code cannot run here; what can be checked without a GPU is the PROTOCOL: the order of flag writes, waits, peer
reads and buffer overwrites that the kernels implement, with the same flag slots (`xflag(kind, rank)`), epochs and
stream orders.  Every rank is a coroutine that yields before each globally visible action; a seeded scheduler picks the
next runnable rank.  Shadow version numbers on every buffer assert the hazards the design argues away:

  R1  a rank reduces region 0 of step t only from gradients of step t (never a peer's step t-1 or t+1 values);
  R2  Adam of step t on rank r reads region 0 only after EVERY rank's slice of the mean for step t has landed;
  R3  a rank's backward of step t+1 overwrites its gradient buffer only after every peer has finished reading step t;
  R4  a flag-in-data slot is overwritten by the sender's step t+1 push only after the receiver consumed step t;
  R5  all ranks finish every step with the same (bit-identical) mean.
"""
import random

import pytest


def xflag(kind, r):  # nn.cuh
    return 16 + kind * 8 + r


class World:
    def __init__(self, W):
        self.W = W
        self.flags = [[0] * 64 for _ in range(W)]          # flags[owner][slot], written by peers
        # gradient buffer of rank r, region 0 (FC): per slice s the (step, kind) it holds; kind 'g' = own gradient, 'm' = mean
        self.fc = [[(0, "m")] * W for _ in range(W)]
        self.conv = [(0, "m")] * W                          # region 1 of rank r: what it holds
        self.ll = [[0] * W for _ in range(W)]               # ll[receiver][sender] = epoch of the data in the slot
        self.ll_consumed = [[0] * W for _ in range(W)]      # last epoch the receiver read from that slot
        self.readers = [[set() for _ in range(W)] for _ in range(W)]   # readers[owner][slice]: ranks currently reading it
        self.result = [dict() for _ in range(W)]            # result[r][step] = tuple describing the mean it applied


def rank_proc(w, r, steps, early):
    """One rank: per step = backward (writes gradients) -> exchange(s) -> Adam, in the device's stream order."""
    W = w.W
    ctr = [0, 0]        # local epoch counters (xchg_ctr[region]); LL epoch
    ll_epoch = 0
    for t in range(1, steps + 1):
        # ---- backward writes this rank's FC gradients (every slice of its own buffer)
        yield
        for s in range(W):
            assert not w.readers[r][s], ("R3: rank %d overwrites FC slice %d at step %d while %s still read it" % (r, s, t, w.readers[r][s]))
            w.fc[r][s] = (t, "g")
        # ---- grad_exchange_kernel, region 0 (on the communication stream when `early`, else after the backward pass)
        epoch = ctr[0] + 1
        for peer in range(W):                                # announce ready in every rank's flag array
            yield
            w.flags[peer][xflag(0, r)] = epoch
        while any(w.flags[r][xflag(0, p)] < epoch for p in range(W)):
            yield                                            # wait_flags(kind 0)
        # reduce this rank's slice r from every rank's buffer
        vals = []
        for p in range(W):
            yield
            w.readers[p][r].add(r)
            assert w.fc[p][r] == (t, "g"), ("R1: rank %d reads slice %d of rank %d holding %s at step %d" % (r, r, p, w.fc[p][r], t))
            vals.append((p, t))
        for p in range(W):
            w.readers[p][r].discard(r)
        mean = tuple(sorted(vals))
        for p in range(W):                                   # store the mean to every rank (peer stores), then ONE fence
            yield
            w.fc[p][r] = (t, "m", mean)
        for peer in range(W):                                # announce delivered
            yield
            w.flags[peer][xflag(1, r)] = epoch
        ctr[0] = epoch
        # ---- conv backward finishes: region 1 gradients written (local buffer only: nobody else reads it in the LL scheme)
        yield
        w.conv[r] = (t, "g")
        # ---- grad_exchange_ll_kernel: push to every peer's receive slot, then consume the peers' pushes
        ll_epoch += 1
        for peer in range(W):
            if peer == r:
                continue
            yield
            assert w.ll_consumed[peer][r] >= ll_epoch - 1, ("R4: rank %d overwrites its slot at rank %d (epoch %d) before epoch %d was consumed"
                                                            % (r, peer, ll_epoch, ll_epoch - 1))
            w.ll[peer][r] = ll_epoch
        got = [(r, t)]
        for peer in range(W):
            if peer == r:
                continue
            while w.ll[r][peer] < ll_epoch:
                yield
            assert w.ll[r][peer] == ll_epoch
            w.ll_consumed[r][peer] = ll_epoch
            got.append((peer, t))
        w.conv[r] = (t, "m", tuple(sorted(got)))
        # ---- adam_kernel: prologue waits for every rank's delivered flag of region 0, then reads the whole vector
        while any(w.flags[r][xflag(1, p)] < ctr[0] for p in range(W)):
            yield
        yield
        for s in range(W):
            assert w.fc[r][s][:2] == (t, "m"), ("R2: Adam of rank %d at step %d reads FC slice %d holding %s" % (r, t, s, w.fc[r][s]))
        assert w.conv[r][:2] == (t, "m")
        w.result[r][t] = (tuple(w.fc[r][s][2] for s in range(W)), w.conv[r][2])


def run(W, steps, seed, early=True):
    rng = random.Random(seed)
    w = World(W)
    procs = [rank_proc(w, r, steps, early) for r in range(W)]
    alive = list(range(W))
    guard = 0
    while alive:
        guard += 1
        assert guard < 2_000_000, "deadlock: no rank can finish"
        r = rng.choice(alive)
        # bias the scheduler now and then so that one rank races far ahead of the others (the skew the flags must absorb)
        if rng.random() < 0.3:
            r = alive[0]
        try:
            next(procs[r])
        except StopIteration:
            alive.remove(r)
    return w


@pytest.mark.parametrize("W", [2, 3, 4, 8])
def test_exchange_protocol_random_interleavings(W):
    for seed in range(25):
        w = run(W, steps=4, seed=1000 * W + seed)
        for t in range(1, 5):
            ref = w.result[0][t]
            for r in range(1, W):
                assert w.result[r][t] == ref, ("R5: ranks disagree on the mean of step %d" % t)
            # every slice of the mean is built from all W ranks' step-t gradients
            for part in ref[0]:
                assert part == tuple((p, t) for p in range(W))
            assert ref[1] == tuple((p, t) for p in range(W))


def test_a_protocol_without_the_ready_wait_is_caught():
    """The model has teeth: with the ready flags pre-set (the rendezvous never blocks) a fast rank reduces gradients its
    peers have not written yet (R1) or overwrites a slice a slow peer still reads (R3).  (Pre-setting the DELIVERED flags
    instead is harmless in this schedule: the flag-in-data exchange that follows is itself a rendezvous behind every peer's
    stores -- the model says Adam's prologue wait is implied whenever the conv region goes through grad_exchange_ll.)"""
    hit = False
    for seed in range(40):
        rng = random.Random(seed)
        w = World(4)
        for r in range(4):
            for p in range(4):
                w.flags[r][xflag(0, p)] = 10 ** 6
        procs = [rank_proc(w, r, 2, True) for r in range(4)]
        alive = list(range(4))
        try:
            while alive:
                r = rng.choice(alive)
                try:
                    next(procs[r])
                except StopIteration:
                    alive.remove(r)
        except AssertionError as e:
            hit = hit or "R1" in str(e) or "R3" in str(e)
    assert hit
