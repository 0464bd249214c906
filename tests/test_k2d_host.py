"""K2D host logic on CPU: element partitioning, the periodic ring exchange over torch.distributed (gloo,
world_size 2 and 3), and the reference's K2D invariants on the oracle (test/unit_tests/k2d_unit_test.jl:75-120)."""
import os
import socket

import numpy as np
import pytest

from kid_b200 import k2d
from kid_b200 import model as M
from oracle.oracle import OracleHandle


def test_partition_elements_covers_the_domain():
    for helem, world in [(32, 1), (32, 2), (32, 3), (7, 4), (1024, 8)]:
        spans = [k2d.partition_elements(helem, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == helem
        for (o0, c0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_subdomain_cases_tile_the_global_case():
    glob = k2d.k2d_case(np.float64, nx=6, nz=8, npoly=3)
    parts = [k2d.k2d_case(np.float64, rank=r, world=3, nx=6, nz=8, npoly=3) for r in range(3)]
    assert sum(p.cfg.nc for p in parts) == glob.cfg.nc
    assert [p.cfg.helem_offset for p in parts] == [0, 2, 4]
    assert np.array_equal(np.concatenate([p.Y0 for p in parts]), glob.Y0)
    assert all(p.cfg.helem_global == 6 for p in parts)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ring_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5
    send_left = torch.full((n,), 10.0 * rank + 1)   # my first element's left-edge column
    send_right = torch.full((n,), 10.0 * rank + 2)  # my last element's right-edge column
    recv_left, recv_right = torch.zeros(n), torch.zeros(n)
    for _ in range(3):  # repeated exchanges keep matching (3 stages per step)
        k2d.exchange_edges(dist, send_left, send_right, recv_left, recv_right, rank, world)
    q.put((rank, recv_left[0].item(), recv_right[0].item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(world):
        r, rl, rr = q.get(timeout=120)
        got[r] = (rl, rr)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        left, right = (r - 1) % world, (r + 1) % world
        assert got[r] == (10.0 * left + 2, 10.0 * right + 1)  # left neighbour's right edge, right neighbour's left edge


def test_single_rank_exchange_wraps_on_itself():
    import torch
    sl, sr = torch.tensor([1.0]), torch.tensor([2.0])
    rl, rr = torch.zeros(1), torch.zeros(1)
    k2d.exchange_edges(None, sl, sr, rl, rr, 0, 1)
    assert rl.item() == 2.0 and rr.item() == 1.0


@pytest.mark.parametrize("ms,ps", [("EquilibriumMoisture", "NoPrecipitation"), ("NonEquilibriumMoisture", "Precipitation2M")])
def test_k2d_zero_flow_advection_adds_nothing(ms, ps):
    """k2d_unit_test.jl:95-119: at t = 1.1·t1 the prescribed flow is zero, so the advection operators (vertical,
    horizontal weak divergence, DSS) leave the tendency untouched: the rhs is the pointwise sources only, identical in
    every column of the horizontally uniform initial state."""
    cs = k2d.k2d_case(np.float64, nx=4, nz=8, npoly=4, moisture_choice=ms, precipitation_choice=ps)
    o = M.make_handle(cs, OracleHandle)
    t = 1.1 * cs.meta["opts"]["t1"]
    dY = o.rhs(t)
    assert np.all(dY == dY[0:1])
    assert np.all(dY[:, 0] == 0)  # ρq_tot has no source


def test_k2d_duplicated_gll_nodes_stay_continuous():
    cs = k2d.k2d_case(np.float64, nx=5, nz=12, npoly=3, precipitation_choice="Precipitation2M", t1=300.0)
    o = M.make_handle(cs, OracleHandle)
    o.step(0.0, 120)
    Y = o.get_state()
    Nq = 4
    assert np.all(np.isfinite(Y))
    assert np.array_equal(Y[Nq - 1::Nq][:-1], Y[Nq::Nq]) and np.array_equal(Y[-1], Y[0])
    assert np.abs(Y[:, 0] - Y[0:1, 0]).max() > 0  # the flow really varies in x
