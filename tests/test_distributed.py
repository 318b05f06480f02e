"""CPU tests of the N > 1 host logic with the gloo backend (world_size 2): row sharding and
the single mean-NLL all-reduce.  The per-rank log-densities come from the oracle here (no GPU
in this container); on the GPU box bench.py exercises the same helpers over NCCL."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import build_flow, load
from zuko_b200.dist import NllRing, all_reduce_gradients, mean_nll, shard_rows


def test_shard_rows_partition():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    from oracle import oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load("flow_nsf35_row")
        spec = O.flowspec_from_module(build_flow("nsf35_row", g))
        lo, hi = shard_rows(g["x"].shape[0], rank, world)
        lp = spec.log_prob(g["x"][lo:hi], g["c"])
        local = torch.tensor([lp.sum()], dtype=torch.float64)
        nll = mean_nll(local, hi - lo)
        out[rank] = float(nll)
    finally:
        dist.destroy_process_group()


def test_mean_nll_allreduce_gloo_world2():
    g = load("flow_nsf35_row")
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        got = dict(out)
    expect = -float(np.mean(g["log_prob64"]))
    assert got[0] == got[1]  # every rank holds the same reduced scalar
    assert abs(got[0] - expect) < 1e-9 * abs(expect)


def _ring_worker(rank: int, world: int, port: int, out):
    """Seven steps through a 3-slot ring (two banks, so banks are rewritten while the previous
    reduction of the other bank is in flight): every step's global mean NLL, in order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ring = NllRing("cpu", slots=3)
        for step in range(7):
            rows = 10 + rank + step  # ragged and changing counts
            ring.slot(rows).fill_(-(step + 1.0) * rows * (rank + 1))  # the engine writes sum log p here
        out[rank] = ring.means().tolist()
    finally:
        dist.destroy_process_group()


def test_nll_ring_gloo_world2():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ring_worker, args=(2, port, out), nprocs=2, join=True)
        got = dict(out)
    expect = []
    for step in range(7):
        rows = [10 + r + step for r in range(2)]
        sums = [-(step + 1.0) * rows[r] * (r + 1) for r in range(2)]
        expect.append(-sum(sums) / sum(rows))
    assert got[0] == got[1]
    assert np.allclose(got[0], expect, rtol=1e-12)


def _ring_reuse_worker(rank: int, world: int, port: int, out):
    """Constant count, banks reused many times: the count column must not accumulate earlier reductions."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ring = NllRing("cpu", slots=2)
        vals = []
        for rnd in range(4):
            for step in range(5):
                ring.slot(8).fill_(-8.0 * (rank + 1) * (step + 1))
            vals.append(ring.means().tolist())
        out[rank] = vals
    finally:
        dist.destroy_process_group()


def test_nll_ring_bank_reuse_gloo_world2():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ring_reuse_worker, args=(2, port, out), nprocs=2, join=True)
        got = dict(out)
    expect = [(8.0 * 1 * (s + 1) + 8.0 * 2 * (s + 1)) / 16.0 for s in range(5)]
    assert got[0] == got[1]
    for rnd in got[0]:
        assert np.allclose(rnd, expect, rtol=1e-12)


def test_nll_ring_without_process_group():
    ring = NllRing("cpu", slots=2)
    for step in range(5):
        ring.slot(4).fill_(-8.0 * (step + 1))
    assert ring.means().tolist() == [2.0, 4.0, 6.0, 8.0, 10.0]
    assert ring.means().numel() == 0


def test_mean_nll_without_process_group():
    v = mean_nll(torch.tensor([-30.0], dtype=torch.float64), 10)
    assert v.item() == 3.0


def _grad_worker(rank: int, world: int, port: int, out):
    """Each rank differentiates its LOCAL mean NLL of a ragged shard with the (fp64) gradient oracle; the
    reduced gradient must equal the gradient of the GLOBAL mean NLL."""
    from cases import oracle_named_grads
    from oracle import oracle as O
    from oracle import oracle_grad as OG

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load("flow_maf35_batch")
        flow = build_flow("maf35_batch", g)
        spec = O.flowspec_from_module(flow)
        n = 101  # ragged: 51 + 50 rows
        lo, hi = shard_rows(n, rank, world)
        w = np.full(hi - lo, -1.0 / (hi - lo))  # d(local mean NLL) / d log_prob
        _, _, lgs = OG.flow_backward(spec, g["x"][lo:hi], g["c"][lo:hi], g_log_prob=w)
        named = oracle_named_grads(flow, lgs)
        for name, p in flow.named_parameters():
            p.grad = torch.from_numpy(named[name].reshape(tuple(p.shape))).to(p.dtype)
        all_reduce_gradients(flow, weights=(hi - lo, n))
        out[rank] = {name: p.grad.double().numpy().copy() for name, p in flow.named_parameters()}
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    from cases import oracle_named_grads
    from oracle import oracle as O
    from oracle import oracle_grad as OG

    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
        got = dict(out)
    g = load("flow_maf35_batch")
    flow = build_flow("maf35_batch", g)
    spec = O.flowspec_from_module(flow)
    n = 101
    _, _, lgs = OG.flow_backward(spec, g["x"][:n], g["c"][:n], g_log_prob=np.full(n, -1.0 / n))
    ref = oracle_named_grads(flow, lgs)
    for name in ref:
        np.testing.assert_array_equal(got[0][name], got[1][name])  # every rank holds the same gradient
        scale = max(float(np.abs(ref[name]).max()), 1e-30)
        assert np.abs(got[0][name].reshape(-1) - ref[name]).max() <= 2e-6 * scale, name  # fp32 transport
