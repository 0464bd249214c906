"""Host <-> device layout path: host arrays go through the copy engine and a staging buffer, or (page-locked arrays, on request
or with a column map) are transposed in place over PCIe by the layout kernels; the column map (kid_set_column_map) lets the host arrays keep the caller's member order while the
device stores the warp-coherent order.  Everything here is data movement: bitwise."""
import numpy as np
import pytest

from kid_b200 import model as M
from kid_b200.lib import Handle, KidCudaError, pinned_empty
from kid_b200.solve import PipelinedEnsemble

pytestmark = pytest.mark.gpu


def case(nc, FT=np.float32, nz=128, **kw):
    return M.k1d_ensemble_case(FT, nc=nc, n_elem=nz, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M",
                               order="sorted", ic="device", **kw)


@pytest.mark.parametrize("direct", [0, 2])
@pytest.mark.parametrize("FT,nz,nc", [(np.float32, 128, 100), (np.float32, 48, 37), (np.float64, 128, 70)])
def test_pinned_and_pageable_host_arrays_give_the_same_state(FT, nz, nc, direct):
    cs = case(nc, FT, nz)
    cs.cfg.reserved[4] = direct  # 2: the layout kernels access the page-locked array in place (default: copy engine + staging)
    h = M.make_handle(cs, Handle)
    h.step(0.0, 5)
    Y_pageable = h.get_state()
    Y_pinned = pinned_empty(Y_pageable.shape, FT)
    h.get_state(Y_pinned)
    assert np.array_equal(Y_pageable, Y_pinned)
    rng = np.random.default_rng(1)
    Ynew = (Y_pageable * (1 + 1e-3 * rng.standard_normal(Y_pageable.shape))).astype(FT)
    Y_pinned[...] = Ynew
    h.set_state(Y_pinned)
    a = h.get_state()
    h.set_state(Ynew)
    assert np.array_equal(a, h.get_state())
    assert np.array_equal(a, Ynew)


def test_column_map_gathers_and_scatters_through_the_host_order():
    nc = 96
    cs = case(nc)
    h = M.make_handle(cs, Handle)
    h.step(0.0, 3)
    Y = h.get_state()
    perm = np.random.default_rng(2).permutation(nc).astype(np.int32)   # device column i <-> host row perm[i]
    host = pinned_empty(Y.shape, np.float32)
    h.set_column_map(perm)
    h.get_state(host)
    assert np.array_equal(host[perm], Y)
    host2 = pinned_empty(Y.shape, np.float32)
    host2[perm] = Y[::-1]
    h.set_state(host2)
    h.set_column_map(None)
    assert np.array_equal(h.get_state(), Y[::-1])
    # a permutation of one contiguous block of host rows travels through the copy engine: pageable arrays work too
    h.set_column_map(perm)
    pageable = np.empty_like(Y)
    h.get_state(pageable)
    assert np.array_equal(pageable[perm], Y[::-1])


def test_scattered_and_block_local_column_maps():
    """Scattered map: device column i <-> an arbitrary row of a LARGER page-locked array (gathered in place over PCIe; pageable
    memory is refused: no host-side gather fallback).  Block-local map: a permutation of rows [base, base + nc) of a larger
    array (copy engine + permutation on the device)."""
    nc = 64
    cs = case(nc)
    h = M.make_handle(cs, Handle)
    h.step(0.0, 3)
    Y = h.get_state()
    rng = np.random.default_rng(3)
    big = pinned_empty((3 * nc,) + Y.shape[1:], np.float32); big[...] = -1.0
    rows = rng.permutation(3 * nc)[:nc].astype(np.int32)               # scattered
    def get(dst):
        h.get_state_async(dst); h.synchronize()
    def put(src):
        h.set_state_async(src); h.synchronize()
    h.set_column_map(rows)
    get(big)
    assert np.array_equal(big[rows], Y)
    untouched = np.setdiff1d(np.arange(3 * nc), rows)
    assert np.all(big[untouched] == -1.0)
    with pytest.raises(KidCudaError):
        get(np.empty_like(big))
    big[...] = -1.0
    local = (nc + rng.permutation(nc)).astype(np.int32)                # block-local: rows [nc, 2 nc)
    h.set_column_map(local)
    get(big)
    assert np.array_equal(big[local], Y) and np.all(big[:nc] == -1.0) and np.all(big[2 * nc:] == -1.0)
    big[local] = Y[::-1]
    put(big)
    h.set_column_map(None)
    assert np.array_equal(h.get_state(), Y[::-1])


def test_pipelined_ensemble_in_member_order_equals_storage_order():
    nc = 256
    cs = case(nc)
    g = M.make_handle(cs, Handle)
    g.step(0.0, 30)
    Y0 = g.get_state()
    a = pinned_empty(Y0.shape, np.float32); a[...] = Y0
    PipelinedEnsemble(cs, nchunks=4).solve(a, 30.0, 10)
    members = np.asarray(cs.meta["column_perm"])
    host_row = np.argsort(np.argsort(members)).astype(np.int32)
    b = pinned_empty(Y0.shape, np.float32); b[host_row] = Y0
    PipelinedEnsemble(cs, nchunks=4, host_column_of=host_row).solve(b, 30.0, 10)
    assert np.array_equal(b[host_row], a)
    g.step(30.0, 10)
    assert np.array_equal(a, g.get_state())
    # the caller's arrays in member order, every chunk stored in its own warp-coherent order (block-local maps)
    by_member = np.argsort(members)
    c = pinned_empty(Y0.shape, np.float32); c[...] = Y0[by_member]
    PipelinedEnsemble.for_member_order(M.reorder_columns(cs, by_member), nchunks=4).solve(c, 30.0, 10)
    assert np.array_equal(c, a[by_member])
