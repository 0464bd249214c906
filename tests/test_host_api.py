"""Host-side mirror of the reference API (CPU only): option factories, state layouts, error behaviour.
Reference: test/unit_tests/common_unit_test.jl:3-158."""
import re
from pathlib import Path

import numpy as np
import pytest

from kid_b200 import equation_types as ET
from kid_b200 import lib
from kid_b200 import model as M
from kid_b200 import parameters as PR

ROOT = Path(__file__).resolve().parents[1]


def test_moisture_and_precipitation_factories():
    td = PR.override_toml_dict(FT=np.float64)
    assert isinstance(ET.get_moisture_type("EquilibriumMoisture", td), ET.EquilibriumMoisture)
    assert isinstance(ET.get_moisture_type("NonEquilibriumMoisture", td), ET.NonEquilibriumMoisture)
    with pytest.raises(ValueError):
        ET.get_moisture_type("_", td)
    assert isinstance(ET.get_precipitation_type("NoPrecipitation", td), ET.NoPrecipitation)
    assert isinstance(ET.get_precipitation_type("Precipitation0M", td), ET.Precipitation0M)
    for rf in ("CliMA_1M", "KK2000", "B1994", "TC1980", "LD2004", "VarTimeScaleAcnv"):
        p = ET.get_precipitation_type("Precipitation1M", td, rain_formation_choice=rf)
        assert isinstance(p, ET.Precipitation1M) and p.rain_formation == ET.RF[rf]
    for rf in ("SB2006", "SB2006NL"):
        p = ET.get_precipitation_type("Precipitation2M", td, rain_formation_choice=rf)
        assert isinstance(p, ET.Precipitation2M) and p.rain_formation == ET.RF[rf]
    with pytest.raises(ValueError):
        ET.get_precipitation_type("_", td, rain_formation_choice="_")
    with pytest.raises(ValueError):
        ET.get_precipitation_type("Precipitation1M", td, rain_formation_choice="_")
    with pytest.raises(ValueError):
        ET.get_precipitation_type("Precipitation2M", td, rain_formation_choice="KK2000")
    with pytest.raises(ValueError):
        ET.get_precipitation_type("Precipitation1M", td, sedimentation_choice="_")


@pytest.mark.parametrize("choice", ["CloudyMoisture", "MoistureP3"])
def test_out_of_scope_moisture_styles_raise_loudly(choice):
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        ET.get_moisture_type(choice, None)


@pytest.mark.parametrize("choice", ["CloudyPrecip", "PrecipitationP3"])
def test_out_of_scope_precip_styles_raise_loudly(choice):
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        ET.get_precipitation_type(choice, None)


def test_initialise_state_field_sets():
    """src/Common/ode_utils.jl:23-54"""
    td = PR.override_toml_dict(FT=np.float64)
    eq, ne = ET.get_moisture_type("EquilibriumMoisture", td), ET.get_moisture_type("NonEquilibriumMoisture", td)
    g = lambda s: ET.get_precipitation_type(s, td)  # noqa: E731
    assert ET.state_variable_names(eq, g("NoPrecipitation")) == ("ρq_tot",)
    assert ET.state_variable_names(eq, g("Precipitation0M")) == ("ρq_tot",)
    assert ET.state_variable_names(ne, g("Precipitation0M")) == ("ρq_tot", "ρq_liq", "ρq_ice")
    assert ET.state_variable_names(eq, g("Precipitation1M")) == ("ρq_tot", "ρq_rai", "ρq_sno")
    assert ET.state_variable_names(ne, g("Precipitation1M")) == ("ρq_tot", "ρq_liq", "ρq_ice", "ρq_rai", "ρq_sno")
    assert ET.state_variable_names(eq, g("Precipitation2M")) == ("ρq_tot", "ρq_rai", "ρq_sno", "N_liq", "N_rai", "N_aer")
    assert ET.state_variable_names(ne, g("Precipitation2M")) == (
        "ρq_tot", "ρq_liq", "ρq_ice", "ρq_rai", "ρq_sno", "N_liq", "N_rai", "N_aer")


def test_k1d_case_shapes_and_defaults():
    cs = M.k1d_case(np.float32, nc=4, moisture_choice="NonEquilibriumMoisture", precipitation_choice="Precipitation2M")
    assert cs.Y0.shape == (4, 8, 128) and cs.Y0.dtype == np.float32
    assert cs.ρ_dry.shape == (128,) and cs.cfg.nz == 128 and cs.cfg.dt == 1.0
    n = list(cs.var_names)
    assert np.all(cs.Y0[:, n.index("ρq_liq")] == 0) and np.all(cs.Y0[:, n.index("N_liq")] == 0)
    # IC N_aer = prescribed_Nd·ρ_dry/1.225 (initial_condition.jl:171)
    np.testing.assert_allclose(cs.Y0[0, n.index("N_aer")], 1e8 * cs.ρ_dry / 1.225, rtol=1e-6)
    with pytest.raises(ValueError):
        M.k1d_case(np.float64, z_max=4000.0)
    with pytest.raises(KeyError):
        M.k1d_case(np.float64, not_an_option=1)
    with pytest.raises(NotImplementedError):
        M.k1d_case(np.float64, velocity="Jouan2020")


def test_param_table_matches_header():
    names = PR.param_names()
    hdr = (ROOT / "include" / "kid_params.h").read_text()
    assert len(names) == len(set(names)) and len(names) > 100
    assert "KID_NPARAMS" in hdr
    tbl = PR.flatten_params(PR.override_toml_dict(FT=np.float32), "KK2000")
    assert tbl.shape == (len(names),) and np.all(np.isfinite(tbl))
    # values are FT-rounded: the float32 table round-trips through float32 exactly
    assert np.array_equal(tbl.astype(np.float32).astype(np.float64), tbl)


def test_ensemble_case_is_deterministic():
    a = M.k1d_ensemble_case(np.float32, nc=64, n_profiles=2, n_elem=16, precipitation_choice="Precipitation2M",
                            moisture_choice="NonEquilibriumMoisture")
    b = M.k1d_ensemble_case(np.float32, nc=32, column_offset=32, n_profiles=2, n_elem=16,
                            precipitation_choice="Precipitation2M", moisture_choice="NonEquilibriumMoisture")
    assert np.array_equal(a.Y0[32:], b.Y0) and np.array_equal(a.col_scalars[lib.COL_W1][32:], b.col_scalars[lib.COL_W1])
    assert a.ρ_dry.shape == (64, 16)
    w1 = a.col_scalars[lib.COL_W1]
    assert w1.min() >= 1.0 and w1.max() <= 2.2 and np.unique(w1).size > 32


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    """The drop-in boundary: libkid_cuda.so exports exactly what include/kid_cuda.h declares."""
    hdr = (ROOT / "include" / "kid_cuda.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kid_[a-z0-9_]+)\s*\(", hdr))
    assert {"kid_create", "kid_step", "kid_set_state", "kid_get_state", "kid_rhs", "kid_get_aux"} <= declared
    l = lib.load_library()
    for name in declared:
        assert hasattr(l, name), f"libkid_cuda.so does not export {name}"
    assert set(l._kid_symbols) == declared
    assert l.kid_nvars(1, 3) == 8 and l.kid_nvars(0, 0) == 1 and l.kid_nvars(5, 0) == -1
    assert b"sm_100a" in l.kid_version()


def test_product_fails_loudly_without_gpu():
    l = lib.load_library()
    if l.kid_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cs = M.k1d_case(np.float32, nc=2, n_elem=8)
    with pytest.raises(lib.KidCudaError, match="no CPU fallback"):
        lib.Handle(cs.cfg)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = ROOT / "kinematicdriver.jl_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "oracle" not in txt.lower() or f.name == "model.py" and "CPU oracle" in txt, f


def test_morton_order_keeps_neighbours_close_in_every_key():
    from kid_b200 import model as M
    rng = np.random.default_rng(3)
    a, b, c = rng.uniform(1, 2.2, 32768), rng.uniform(7, 8.5, 32768), rng.uniform(0.98e5, 1.02e5, 32768)
    perm = M.morton_order(a, b, c)
    assert sorted(perm.tolist()) == list(range(32768))
    def spread(x):  # mean range of a key inside a warp of 32 neighbouring members, relative to the key's full range
        g = x[perm].reshape(-1, 32)
        return float(np.mean(g.max(1) - g.min(1)) / (x.max() - x.min()))
    lex = np.lexsort((c, b, a))
    def spread_lex(x):
        g = x[lex].reshape(-1, 32)
        return float(np.mean(g.max(1) - g.min(1)) / (x.max() - x.min()))
    assert max(spread(a), spread(b), spread(c)) < 0.25      # close in all three keys
    assert max(spread_lex(b), spread_lex(c)) > 0.8          # a lexicographic sort only helps the first key


def test_chen2022_sedimentation_choice():
    for ps in ("Precipitation2M", "Precipitation1M"):
        p = ET.get_precipitation_type(ps, PR.override_toml_dict(), sedimentation_choice="Chen2022")
        assert p.sedimentation == 2 and p.sedimentation_name == "Chen2022"
    with pytest.raises(ValueError):
        ET.get_precipitation_type("Precipitation1M", PR.override_toml_dict(), sedimentation_choice="SB2006")
