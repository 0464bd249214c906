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
    h.set_column_map(perm)
    with pytest.raises(KidCudaError):   # a map needs page-locked host memory (no host-side gather fallback)
        h.get_state(np.empty_like(Y))


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
