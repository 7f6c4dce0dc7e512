"""Host-side model of the one-shot all-reduce protocol of csrc/comm.cu (allreduce_oneshot_kernel): P threads play the
ranks and execute the kernel's steps -- stage the own values in the buffer of the call's parity, write the epoch into a
flag word of every peer, wait for every peer's epoch, read the P staged copies and add them in rank order -- with random
delays between the steps and NO barrier between calls.  What it checks is the protocol's claim that double buffering by
the parity of the epoch needs no closing barrier: a reader must never see a staging buffer that its owner has already
overwritten for a later call (every staged value carries the epoch it was written in), and every rank must form the
same sums.  (Memory ordering on the GPU is the kernel's business -- release stores / acquire loads at system scope; the
exact-sum check on real GPUs is tests/check_multi_gpu.py::check_allreduce.)"""
import random
import threading
import time

import pytest


class Region:
    """What one rank exposes to its peers: two staging buffers and two rows of flag words (one per parity)."""

    def __init__(self, nranks, n):
        self.stage = [[(0.0, 0)] * n for _ in range(2)]  # (value, epoch it was staged in)
        self.flags = [[0] * nranks for _ in range(2)]
        self.epoch = 0


def allreduce(rank, regions, values, rng, errors):
    me = regions[rank]
    P = len(regions)
    e = me.epoch + 1
    par = e & 1
    # 1. own values -> own staging buffer of this parity
    for i, v in enumerate(values):
        me.stage[par][i] = (v, e)
        if rng.random() < 0.05:
            time.sleep(0)
    # flag every peer
    for r in range(P):
        if r != rank:
            regions[r].flags[par][rank] = e
    # 2. wait for the peers' epochs
    deadline = time.time() + 20
    for r in range(P):
        if r == rank:
            continue
        while me.flags[par][r] < e:
            if time.time() > deadline:
                errors.append(f"rank {rank}: no flag from {r} in call {e}")
                return None
            time.sleep(0)
    if rng.random() < 0.3:
        time.sleep(rng.random() * 2e-3)  # a slow reader: the peers may already be one call ahead
    # 3. read every staged copy, add in rank order
    out = []
    for i in range(len(values)):
        s = 0.0
        for r in range(P):
            v, tag = regions[r].stage[par][i]
            if tag != e:
                errors.append(f"rank {rank} call {e}: staging of rank {r} holds epoch {tag}")
            s = v if r == 0 else s + v
        out.append(s)
    me.epoch = e
    return out


@pytest.mark.parametrize("P", [2, 3, 8])
def test_double_buffered_epochs_need_no_closing_barrier(P):
    n, calls = 5, 120
    regions = [Region(P, n) for _ in range(P)]
    results = [[] for _ in range(P)]
    errors = []

    def run(rank):
        rng = random.Random(100 + rank)
        for c in range(calls):
            vals = [float((rank + 1) * (c + 1) + i) for i in range(n)]
            out = allreduce(rank, regions, vals, rng, errors)
            if out is None:
                return
            results[rank].append(out)
            if rng.random() < 0.2:
                time.sleep(rng.random() * 1e-3)  # skew between the ranks

    threads = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors[:3]
    for c in range(calls):
        want = [float(sum((r + 1) * (c + 1) + i for r in range(P))) for i in range(n)]
        for r in range(P):
            assert results[r][c] == want, (r, c)


def test_a_single_buffer_would_be_overwritten():
    """The same protocol with ONE staging buffer and one flag row does hit the hazard the parity avoids -- the model is
    able to see it (otherwise the test above would prove nothing)."""
    P, n = 2, 4
    regions = [Region(P, n) for _ in range(P)]
    seen_stale = []

    def call(rank, e, slow):
        me = regions[rank]
        for i in range(n):
            me.stage[0][i] = (float(rank + e), e)
        regions[1 - rank].flags[0][rank] = e
        while me.flags[0][1 - rank] < e:
            time.sleep(0)
        if slow:
            time.sleep(5e-3)
        for i in range(n):
            _, tag = regions[1 - rank].stage[0][i]
            if tag != e:
                seen_stale.append((rank, e, tag))

    def run(rank):
        for e in range(1, 30):
            call(rank, e, slow=(rank == 0))

    ts = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert seen_stale, "the fast rank never got a call ahead of the slow reader"
