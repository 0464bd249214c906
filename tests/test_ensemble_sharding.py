"""Ensemble sharding across ranks (bench.py --gpus N: contiguous column ranges, no collective): the shards of the
hash-ordered ensemble are exactly the global ensemble, and a world_size-2 gloo job stepping its shards on the CPU oracle
reproduces the single-process run (the only cross-rank traffic is the max-reduction of the timing)."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

from kid_b200 import model as M
from kid_b200.lib import COL_PRESCRIBED_ND, COL_W1

KW = dict(n_elem=16, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")


def test_shards_tile_the_global_ensemble():
    full = M.k1d_ensemble_case(np.float32, nc=96, n_profiles=4, order="hash", **KW)
    parts = [M.k1d_ensemble_case(np.float32, nc=32, n_profiles=4, order="hash", column_offset=32 * r, **KW) for r in range(3)]
    for r, p in enumerate(parts):
        sl = slice(32 * r, 32 * r + 32)
        assert np.array_equal(p.Y0, full.Y0[sl]) and np.array_equal(p.ρ_dry, full.ρ_dry[sl])
        for k in (COL_W1, COL_PRESCRIBED_ND):
            assert np.array_equal(p.col_scalars[k], full.col_scalars[k][sl])
    # device-built initial condition: the per-member surface pressures shard the same way
    fd = M.k1d_ensemble_case(np.float32, nc=96, ic="device", order="hash", **KW)
    pd = M.k1d_ensemble_case(np.float32, nc=32, ic="device", order="hash", column_offset=64, **KW)
    assert np.array_equal(pd.device_ic["p0"], fd.device_ic["p0"][64:]) and len(np.unique(fd.device_ic["p0"])) == 96
    # slice_columns is the same operation on an existing case
    s = M.slice_columns(full, 32, 64)
    assert np.array_equal(s.Y0, parts[1].Y0) and s.cfg.nc == 32


def _rank(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleHandle
    import torch
    nc = 8
    case = M.k1d_ensemble_case(np.float64, nc=nc, n_profiles=2, order="hash", column_offset=rank * nc, **KW)
    h = M.make_handle(case, OracleHandle)
    dist.barrier()
    h.step(0.0, 40)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # stands for the per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    np.save(out + f".{rank}.npy", h.get_state())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_job_reproduces_the_single_process_ensemble(tmp_path):
    from oracle.oracle import OracleHandle
    out = str(tmp_path / "Y")
    mp.spawn(_rank, args=(2, 29641, out), nprocs=2, join=True)
    full = M.k1d_ensemble_case(np.float64, nc=16, n_profiles=2, order="hash", **KW)
    h = M.make_handle(full, OracleHandle)
    h.step(0.0, 40)
    Y = h.get_state()
    got = np.concatenate([np.load(out + f".{r}.npy") for r in range(2)])
    assert np.array_equal(got, Y)


def test_reordering_columns_does_not_change_any_member():
    """warp_coherent_order + reorder_columns + inverse_permutation: results come back in the caller's order, unchanged."""
    from oracle.oracle import OracleHandle
    case = M.k1d_ensemble_case(np.float64, nc=24, n_profiles=3, order="hash", **KW)
    perm = M.warp_coherent_order(case)
    assert sorted(perm.tolist()) == list(range(24)) and not np.array_equal(perm, np.arange(24))
    ordered = M.reorder_columns(case, perm)
    a, b = M.make_handle(case, OracleHandle), M.make_handle(ordered, OracleHandle)
    a.step(0.0, 60)
    b.step(0.0, 60)
    assert np.array_equal(b.get_state()[M.inverse_permutation(perm)], a.get_state())
    # neighbours of the ordered ensemble are similar in every distinguishing scalar
    w = np.asarray(ordered.col_scalars[COL_W1], dtype=np.float64)
    assert np.abs(np.diff(w)).mean() < 0.5 * np.abs(np.diff(np.asarray(case.col_scalars[COL_W1], dtype=np.float64))).mean()
    # device-built initial conditions reorder their per-member surface pressure too
    dcase = M.k1d_ensemble_case(np.float32, nc=64, ic="device", order="hash", **KW)
    dperm = M.warp_coherent_order(dcase)
    dord = M.reorder_columns(dcase, dperm)
    assert np.array_equal(dord.device_ic["p0"], dcase.device_ic["p0"][dperm])


def test_nested_bin_order_groups_members_by_profile_then_updraft():
    """The storage order of large ensembles (model.nested_bin_order): a permutation; neighbouring members are close in the
    primary key (surface pressure) AND, inside a primary bin, in the secondary key; small ensembles fall back to the Z-order
    curve.  warp_coherent_order picks it for ensembles that differ in the sounding, w1 and Nd."""
    rng = np.random.default_rng(5)
    n = 1 << 14
    p0, w1, nd = rng.uniform(0.98e5, 1.02e5, n), rng.uniform(1.0, 2.2, n), rng.uniform(7.0, 8.5, n)
    perm = M.nested_bin_order(p0, w1, nd)
    assert sorted(perm.tolist()) == list(range(n))
    d_p0 = np.abs(np.diff(p0[perm])).mean() / np.abs(np.diff(p0)).mean()
    d_w1 = np.abs(np.diff(w1[perm])).mean() / np.abs(np.diff(w1)).mean()
    assert d_p0 < 0.05 and d_w1 < 0.2, (d_p0, d_w1)     # random order: both ratios are 1
    small = M.nested_bin_order(p0[:100], w1[:100], nd[:100])
    assert sorted(small.tolist()) == list(range(100))
    case = M.k1d_ensemble_case(np.float32, nc=4096, ic="device", order="hash", **KW)
    perm = M.warp_coherent_order(case)
    ref = M.nested_bin_order(np.asarray(case.device_ic["p0"], dtype=np.float64), np.asarray(case.col_scalars[COL_W1], dtype=np.float64),
                             np.log10(np.asarray(case.col_scalars[COL_PRESCRIBED_ND], dtype=np.float64)))
    assert np.array_equal(perm, ref)
    sorted_case = M.k1d_ensemble_case(np.float32, nc=4096, ic="device", order="sorted", **KW)
    assert np.array_equal(sorted_case.meta["column_perm"], ref)   # the benchmark's "sorted" order is this order


def test_round_robin_block_shards_cover_the_ensemble_and_take_columns_slices_everything():
    """bench.py --gpus N: blocks of 4096 sorted columns are dealt round-robin (every rank gets the same mix of members);
    model.take_columns gives a rank its columns of a host-built or device-built case."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    total = 3 * bench.SHARD_BLOCK * 4 + 100
    for world in (1, 2, 4, 8):
        parts = [bench.shard_columns(total, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(total))
        assert max(p.size for p in parts) - min(p.size for p in parts) <= bench.SHARD_BLOCK
        assert all(np.all(np.diff(p) > 0) for p in parts)
    full = M.k1d_ensemble_case(np.float32, nc=96, n_profiles=4, order="sorted", **KW)
    idx = np.array([5, 6, 7, 40, 41, 90])
    sub = M.take_columns(full, idx)
    assert sub.cfg.nc == 6 and full.cfg.nc == 96
    assert np.array_equal(sub.Y0, full.Y0[idx]) and np.array_equal(sub.ρ_dry, full.ρ_dry[idx])
    assert np.array_equal(sub.col_scalars[COL_W1], full.col_scalars[COL_W1][idx])
    dev = M.k1d_ensemble_case(np.float32, nc=96, ic="device", order="sorted", **KW)
    sd = M.take_columns(dev, idx)
    assert np.array_equal(sd.device_ic["p0"], dev.device_ic["p0"][idx])
    assert np.array_equal(sd.meta["column_perm"], np.asarray(dev.meta["column_perm"])[idx])


def test_bench_parity_measures():
    """bench.py's parity helpers: field-maximum errors, columns over tolerance, column-integrated number."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    names = ("ρq_tot", "ρq_liq", "N_liq", "N_rai", "N_aer")
    rng = np.random.default_rng(0)
    b = rng.random((10, 5, 8)) + 0.5
    a = b.copy()
    a[3, 2, 4] += 0.5   # one column: activation moved by a level (N_liq up, N_aer down: the sum is conserved)
    a[3, 4, 4] -= 0.5
    m = bench.knife_edge_measures(a, b, names)
    assert m["columns_over_tolerance"] == {"ρq_tot": 0, "ρq_liq": 0, "N_liq": 1, "N_rai": 0, "N_aer": 1}
    assert m["column_integrated_number_rel_err"] < 1e-12
    e = bench.sample_errors(a, b, names)
    assert e["ρq_tot"] == 0.0 and 0.3 < e["N_liq"] < 0.4
