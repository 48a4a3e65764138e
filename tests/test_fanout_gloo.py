"""CPU, world_size 2, gloo: the N>1 path (shard -> independent chains -> all-gather of assignments -> label
offsets).  The per-rank engine here is the CPU oracle (test infrastructure); on GPUs bench.py runs the same
host logic with the CUDA engine and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, D, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pybgmm_b200 import fanout
        X, _ = cases.gen(N, D, 4, 1)
        lo, hi = fanout.shard_bounds(N, world, rank)
        m_0, k_0, v_0, S_0 = cases.prior_for(D, "full")
        orc = O.Oracle(X[lo:hi], m_0, k_0, v_0, S_0, K_max=32)
        rng = np.random.RandomState(100 + rank)
        orc.set_assignments(O.init_assignments(hi - lo, np.arange(hi - lo) % 3))
        for _ in range(3):
            orc.sweep(rng.random_sample(hi - lo), 1.0)
        z_local = torch.from_numpy(orc.assignments.copy())
        outs, ks = fanout.gather_assignments(z_local, orc.K)
        glob = fanout.offset_labels([t.numpy() for t in outs], ks)
        q.put((rank, orc.assignments.tolist(), orc.K, glob.tolist(), ks))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world, N, D = 2, 240, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, z0, K0, g0, ks0), (_, z1, K1, g1, ks1) = res
    assert g0 == g1 and ks0 == ks1 == [K0, K1]           # every rank ends with the same global labelling
    assert len(g0) == N
    assert g0[:N // 2] == z0                              # shard 0 keeps its labels
    assert g0[N // 2:] == [k + K0 for k in z1]            # shard 1 is offset by K_0
    assert max(g0) == K0 + K1 - 1
