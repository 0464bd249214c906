"""Parameter creation: the host-side mirror of the reference's ClimaParams plumbing.

Mirrors (reference paths):
  test/create_parameters.jl:8-88   override_toml_dict, create_*_parameters
  src/Common/parameters.jl:18-41   CommonParameters
  src/K1DModel/parameters.jl:17-73 KinematicDriverParameters
and flattens everything into the versioned table of include/kid_params.h.

The default VALUES of ClimaParams 1.0 are not in the reference checkout (un-vendored
dependency).  DEFAULTS restates them from the published ClimaParams TOML; the names
are ClimaParams names so a maintainer can diff them against the real file.  On the
Julia side the shim does not need this file at all: it reads the struct fields the
reference already built and passes the values through.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------- defaults
DEFAULTS: dict[str, float] = {
    # Thermodynamics
    "temperature_reference": 273.16,  # T_0
    "mean_sea_level_pressure": 101325.0,
    "pressure_triple_point": 611.657,
    "temperature_triple_point": 273.16,
    "temperature_water_freeze": 273.15,
    "temperature_homogenous_nucleation": 233.0,
    "pow_icenuc": 1.0,
    "universal_gas_constant": 8.3144598,
    "molar_mass_dry_air": 0.02897,
    "molar_mass_water": 0.01801528,
    "adiabatic_exponent_dry_air": 2.0 / 7.0,
    "gas_constant_dry_air": 8.3144598 / 0.02897,
    "gas_constant_vapor": 8.3144598 / 0.01801528,
    "isobaric_specific_heat_vapor": 1859.0,
    "isobaric_specific_heat_liquid": 4181.0,
    "isobaric_specific_heat_ice": 2100.0,
    "latent_heat_vaporization_at_reference": 2500800.0,
    "latent_heat_sublimation_at_reference": 2834400.0,
    "gravitational_acceleration": 9.81,
    # air properties
    "thermal_conductivity_of_air": 2.4e-2,
    "diffusivity_of_water_vapor": 2.26e-5,
    "kinematic_viscosity_of_air": 1.6e-5,
    # aerosol activation (ARG2000)
    "surface_tension_water": 0.072,
    "density_liquid_water": 1000.0,
    "density_ice_water": 916.7,
    "ARG2000_f_coeff_1": 0.5,
    "ARG2000_f_coeff_2": 2.5,
    "ARG2000_g_coeff_1": 1.0,
    "ARG2000_g_coeff_2": 0.25,
    "ARG2000_pow_1": 1.5,
    "ARG2000_pow_2": 0.75,
    # non-equilibrium cloud formation
    "condensation_evaporation_timescale": 10.0,
    "sublimation_deposition_timescale": 10.0,
    # 0-moment
    "precipitation_timescale": 1000.0,
    "specific_humidity_precipitation_threshold": 5e-6,
    "supersaturation_precipitation_threshold": 0.02,
    # 1-moment: rain
    "rain_autoconversion_timescale": 1000.0,
    "cloud_liquid_water_specific_humidity_autoconversion_threshold": 5e-4,
    "threshold_smooth_transition_steepness": 2.0,
    "rain_drop_length_scale": 1e-3,
    "rain_mass_size_relation_coefficient_me": 3.0,
    "rain_mass_size_relation_coefficient_delm": 0.0,
    "rain_mass_size_relation_coefficient_chim": 1.0,
    "rain_cross_section_size_relation_coefficient_ae": 2.0,
    "rain_cross_section_size_relation_coefficient_dela": 0.0,
    "rain_cross_section_size_relation_coefficient_chia": 1.0,
    "rain_drop_size_distribution_coefficient_n0": 16e6,
    "rain_ventillation_coefficient_a": 1.5,
    "rain_ventillation_coefficient_b": 0.53,
    "rain_terminal_velocity_size_relation_coefficient_chiv": 1.0,
    "rain_terminal_velocity_size_relation_coefficient_ve": 0.5,
    "rain_terminal_velocity_size_relation_coefficient_delv": 0.0,
    "rain_drop_drag_coefficient": 0.55,
    # 1-moment: snow
    "snow_autoconversion_timescale": 100.0,
    "cloud_ice_specific_humidity_autoconversion_threshold": 1e-6,
    "snow_flake_length_scale": 1e-3,
    "snow_mass_size_relation_coefficient_m0_prefactor": 1e-1,  # m0 = 0.1 r0^2
    "snow_mass_size_relation_coefficient_me": 2.0,
    "snow_mass_size_relation_coefficient_delm": 0.0,
    "snow_mass_size_relation_coefficient_chim": 1.0,
    "snow_cross_section_size_relation_coefficient_a0_prefactor": 0.3,  # a0 = 0.3 π r0^2
    "snow_cross_section_size_relation_coefficient_ae": 2.0,
    "snow_cross_section_size_relation_coefficient_dela": 0.0,
    "snow_cross_section_size_relation_coefficient_chia": 1.0,
    "snow_flake_size_distribution_coefficient_mu": 4.36e9,
    "snow_flake_size_distribution_coefficient_nu": 0.63,
    "snow_ventillation_coefficient_a": 0.65,
    "snow_ventillation_coefficient_b": 0.44,
    "snow_terminal_velocity_size_relation_coefficient_chiv": 1.0,
    "snow_terminal_velocity_size_relation_coefficient_ve": 0.25,
    "snow_terminal_velocity_size_relation_coefficient_delv": 0.0,
    # 1-moment: cloud ice
    "cloud_ice_crystals_length_scale": 1e-5,
    "cloud_ice_mass_size_relation_coefficient_me": 3.0,
    "cloud_ice_mass_size_relation_coefficient_delm": 0.0,
    "cloud_ice_mass_size_relation_coefficient_chim": 1.0,
    "cloud_ice_size_distribution_coefficient_n0": 2e7,
    "ice_snow_threshold_radius": 62.5e-6,
    # collision efficiencies
    "cloud_liquid_rain_collision_efficiency": 0.8,
    "cloud_liquid_snow_collision_efficiency": 0.1,
    "cloud_ice_rain_collision_efficiency": 1.0,
    "cloud_ice_snow_collision_efficiency": 0.1,
    "rain_snow_collision_efficiency": 1.0,
    # 2-moment-style rain formation usable with 1M (KK2000, B1994, TC1980, LD2004, VarTimescale)
    "KK2000_autoconversion_coeff_A": 7.42e13,
    "KK2000_autoconversion_coeff_a": 2.47,
    "KK2000_autoconversion_coeff_b": -1.79,
    "KK2000_autoconversion_coeff_c": -1.47,
    "KK2000_accretion_coeff_A": 67.0,
    "KK2000_accretion_coeff_a": 1.15,
    "KK2000_accretion_coeff_b": -1.3,
    "B1994_autoconversion_coeff_C": 6e28,
    "B1994_autoconversion_coeff_a": -1.7,
    "B1994_autoconversion_coeff_b": 4.7,
    "B1994_autoconversion_coeff_c": -3.3,
    "B1994_autoconversion_coeff_N_0": 2e8,
    "B1994_autoconversion_coeff_d_low": 3.9,
    "B1994_autoconversion_coeff_d_high": 9.9,
    "B1994_accretion_coeff_A": 6.0,
    "TC1980_autoconversion_coeff_a": 7.0 / 3.0,
    "TC1980_autoconversion_coeff_b": -1.0 / 3.0,
    "TC1980_autoconversion_coeff_D": 3268.0,
    "TC1980_autoconversion_coeff_r_0": 7e-6,
    "TC1980_autoconversion_coeff_me_liq": 3.0,
    "TC1980_accretion_coeff_A": 4.7,
    "LD2004_autoconversion_coeff_R_6C_0": 7.5,
    "LD2004_autoconversion_coeff_E_0": 1.08e10,
    "Variable_time_scale_autoconversion_coeff_tau": 1000.0,
    "Variable_time_scale_autoconversion_coeff_alpha": 1.0,
    # SB2006
    "SB2006_collection_kernel_coeff_kcc": 4.44e9,
    "SB2006_raindrops_min_mass": 2.6e-10,  # x*, also pdf_r.xr_min
    "SB2006_reference_air_density": 1.225,
    "SB2006_autoconversion_correcting_function_coeff_A": 400.0,
    "SB2006_autoconversion_correcting_function_coeff_a": 0.7,
    "SB2006_autoconversion_correcting_function_coeff_b": 3.0,
    "SB2006_collection_kernel_coeff_kcr": 5.25,
    "SB2006_accretion_correcting_function_coeff_tau0": 5e-5,
    "SB2006_accretion_correcting_function_coeff_c": 4.0,
    "SB2006_collection_kernel_coeff_krr": 7.12,
    "SB2006_collection_kernel_coeff_kapparr": 60.7,
    "SB2006_raindrops_self_collection_coeff_d": -5.0,
    "SB2006_raindrops_equlibrium_mean_diameter": 9e-4,
    "SB2006_raindrops_breakup_mean_diameter_threshold": 3.5e-4,
    "SB2006_raindrops_breakup_coeff_kbr": 1000.0,
    "SB2006_raindrops_breakup_coeff_kappabr": 2300.0,
    "SB2006_ventilation_factor_coeff_av": 0.78,
    "SB2006_ventilation_factor_coeff_bv": 0.308,
    "SB2006_raindrops_terminal_velocity_coeff_alpha": 159.0,
    "SB2006_raindrops_terminal_velocity_coeff_beta": 0.266,
    "SB2006_number_adjustment_timescale": 100.0,
    "SB2006_cloud_gamma_distribution_parameter": 2.0,  # ν_c
    "SB2006_cloud_droplets_min_mass": 4.2e-15,
    "SB2006_cloud_droplets_max_mass": 2.6e-10,
    "SB2006_raindrops_max_mass": 5e-6,
    "SB2006_raindrops_size_distribution_coeff_N0_min": 3.5e5,
    "SB2006_raindrops_size_distribution_coeff_N0_max": 2e10,
    "SB2006_raindrops_size_distribution_coeff_lambda_min": 1e3,
    "SB2006_raindrops_size_distribution_coeff_lambda_max": 1e4,
    "SB2006_raindrops_terminal_velocity_coeff_aR": 9.65,
    "SB2006_raindrops_terminal_velocity_coeff_bR": 10.3,
    "SB2006_raindrops_terminal_velocity_coeff_cR": 600.0,
    # CMP.Chen2022VelTypeRain: Chen et al. (2022) Table B1
    "Chen2022_table_B1_q_coeff": 0.115231,
    "Chen2022_table_B1_a1_coeff": 0.044612,
    "Chen2022_table_B1_a2_coeff": -0.263166,
    "Chen2022_table_B1_a3_coeff": 4.7178,
    "Chen2022_table_B1_a3_pow_coeff": -0.47335,
    "Chen2022_table_B1_b1_coeff": 2.2955,
    "Chen2022_table_B1_b2_coeff": 2.2955,
    "Chen2022_table_B1_b3_coeff": 1.1451,
    "Chen2022_table_B1_b_rho_coeff": 0.038465,
    "Chen2022_table_B1_c1_coeff": 0.0,
    "Chen2022_table_B1_c2_coeff": 0.184325,
    "Chen2022_table_B1_c3_coeff": 0.184325,
    # KinematicDriver / Common (no defaults in ClimaParams: set by override_toml_dict)
    "prescribed_flow_w1": 2.0,
    "prescribed_flow_t1": 600.0,
    "surface_pressure": 100700.0,
    "precipitation_sources_flag": 1,
    "precipitation_sinks_flag": 1,
    "qtot_flux_correction_flag": 0,
    "prescribed_Nd": 100e6,
    "open_system_activation": 0,
    "local_activation": 0,
    "r_dry": 0.04e-6,
    "std_dry": 1.4,
    "kappa": 1.12,
    "init_cond_z0": 0.0,
    "init_cond_z1": 740.0,
    "init_cond_z2": 3260.0,
    "init_cond_rv0": 0.015,
    "init_cond_rv1": 0.0138,
    "init_cond_rv2": 0.0024,
    "init_cond_theta0": 297.9,
    "init_cond_theta1": 297.9,
    "init_cond_theta2": 312.66,
    # restated ClimaCore FluxCorrectionC2C scaling
    "flux_correction_scale": 1.0,
}


def create_toml_dict(FT=np.float64, override_file: dict | None = None) -> dict:
    """CP.create_toml_dict(FT; override_file) — a name -> value dict tagged with FT."""
    d = copy.deepcopy(DEFAULTS)
    if override_file:
        for k, v in override_file.items():
            d[k] = v["value"] if isinstance(v, dict) else v
    d["__FT__"] = np.dtype(FT).type
    return d


def override_toml_dict(
    toml_dict: dict | None = None,
    *,
    FT=None,
    w1=2.0, t1=600.0, p0=100700.0, z_0=0.0, z_1=740.0, z_2=3260.0,
    rv_0=0.015, rv_1=0.0138, rv_2=0.0024, tht_0=297.9, tht_1=297.9, tht_2=312.66,
    precip_sources=1, precip_sinks=1, qtot_flux_correction=0, prescribed_Nd=100e6,
    open_system_activation=0, local_activation=0, r_dry=0.04e-6, std_dry=1.4, kappa=1.12,
) -> dict:
    """test/create_parameters.jl:8-66 — the PySDM-matching constants plus the KiD options."""
    if FT is None:
        FT = toml_dict["__FT__"] if toml_dict is not None else np.float64
    override_file = {
        "mean_sea_level_pressure": 100000.0,
        "gravitational_acceleration": 9.80665,
        "universal_gas_constant": 8.314462618,
        "adiabatic_exponent_dry_air": 0.2855747338575384,
        "isobaric_specific_heat_vapor": 1850.0,
        "gas_constant_vapor": 461.52998157091315,
        "gas_constant_dry_air": 287.0027047999343,
        "cloud_liquid_water_specific_humidity_autoconversion_threshold": 0.0001,
        "prescribed_flow_w1": w1,
        "prescribed_flow_t1": t1,
        "surface_pressure": p0,
        "precipitation_sources_flag": precip_sources,
        "precipitation_sinks_flag": precip_sinks,
        "qtot_flux_correction_flag": qtot_flux_correction,
        "prescribed_Nd": prescribed_Nd,
        "open_system_activation": open_system_activation,
        "local_activation": local_activation,
        "r_dry": r_dry,
        "std_dry": std_dry,
        "kappa": kappa,
        "init_cond_z0": z_0, "init_cond_z1": z_1, "init_cond_z2": z_2,
        "init_cond_rv0": rv_0, "init_cond_rv1": rv_1, "init_cond_rv2": rv_2,
        "init_cond_theta0": tht_0, "init_cond_theta1": tht_1, "init_cond_theta2": tht_2,
    }
    return create_toml_dict(FT, override_file)


# ------------------------------------------------------------ parameter structs
@dataclass(frozen=True)
class CommonParameters:
    """src/Common/parameters.jl:18-41"""
    precip_sources: bool
    precip_sinks: bool
    prescribed_Nd: float
    open_system_activation: bool
    local_activation: bool


@dataclass(frozen=True)
class KinematicDriverParameters:
    """src/K1DModel/parameters.jl:17-73"""
    w1: float
    t1: float
    p0: float
    z_0: float
    z_1: float
    z_2: float
    rv_0: float
    rv_1: float
    rv_2: float
    tht_0: float
    tht_1: float
    tht_2: float
    qtot_flux_correction: int
    r_dry: float
    std_dry: float
    κ: float


@dataclass(frozen=True)
class ThermodynamicsParameters:
    """Thermodynamics.Parameters.ThermodynamicsParameters (fields the hot path reads)."""
    T_0: float
    MSLP: float
    press_triple: float
    T_triple: float
    T_freeze: float
    T_icenuc: float
    pow_icenuc: float
    R_d: float
    R_v: float
    cp_d: float
    cp_v: float
    cp_l: float
    cp_i: float
    LH_v0: float
    LH_s0: float
    grav: float

    @property
    def Rv_over_Rd(self):
        return self.R_v / self.R_d


def _ft(td):
    return td["__FT__"]


def create_common_parameters(td: dict) -> CommonParameters:
    FT = _ft(td)
    return CommonParameters(
        bool(td["precipitation_sources_flag"]), bool(td["precipitation_sinks_flag"]), FT(td["prescribed_Nd"]),
        bool(td["open_system_activation"]), bool(td["local_activation"]),
    )


def create_kid_parameters(td: dict) -> KinematicDriverParameters:
    FT = _ft(td)
    g = lambda k: FT(td[k])  # noqa: E731
    return KinematicDriverParameters(
        g("prescribed_flow_w1"), g("prescribed_flow_t1"), g("surface_pressure"),
        g("init_cond_z0"), g("init_cond_z1"), g("init_cond_z2"),
        g("init_cond_rv0"), g("init_cond_rv1"), g("init_cond_rv2"),
        g("init_cond_theta0"), g("init_cond_theta1"), g("init_cond_theta2"),
        int(td["qtot_flux_correction_flag"]), g("r_dry"), g("std_dry"), g("kappa"),
    )


def create_thermodynamics_parameters(td: dict) -> ThermodynamicsParameters:
    FT = _ft(td)
    g = lambda k: FT(td[k])  # noqa: E731
    R_d = g("gas_constant_dry_air")
    return ThermodynamicsParameters(
        g("temperature_reference"), g("mean_sea_level_pressure"), g("pressure_triple_point"),
        g("temperature_triple_point"), g("temperature_water_freeze"), g("temperature_homogenous_nucleation"),
        g("pow_icenuc"), R_d, g("gas_constant_vapor"), FT(R_d / g("adiabatic_exponent_dry_air")),
        g("isobaric_specific_heat_vapor"), g("isobaric_specific_heat_liquid"), g("isobaric_specific_heat_ice"),
        g("latent_heat_vaporization_at_reference"), g("latent_heat_sublimation_at_reference"),
        g("gravitational_acceleration"),
    )


# -------------------------------------------------------------- flat param table
def param_names() -> list[str]:
    """Names of include/kid_params.h in enum order (parsed from the header so the two cannot drift)."""
    import re
    from pathlib import Path

    hdr = (Path(__file__).resolve().parents[2] / "include" / "kid_params.h").read_text()
    body = hdr.split("#define KID_PARAM_LIST(X)")[1].split("enum kid_param_id")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return re.findall(r"X\((\w+)\)", body)


_RF_IDS = {"CliMA_1M": 0, "KK2000": 1, "B1994": 2, "TC1980": 3, "LD2004": 4, "VarTimeScaleAcnv": 5,
           "SB2006": 6, "SB2006NL": 7}


def flatten_params(td: dict, rain_formation: str | int | None = None) -> np.ndarray:
    """Build the KID_NPARAMS table from a toml dict, doing derived-field arithmetic in FT
    the way the CloudMicrophysics/Thermodynamics constructors do."""
    FT = _ft(td)
    g = lambda k: FT(td[k])  # noqa: E731
    pi = FT(math.pi)
    tp = create_thermodynamics_parameters(td)
    rho_w = g("density_liquid_water")
    rho_i = g("density_ice_water")
    v: dict[str, float] = {}
    v.update(td_T_0=tp.T_0, td_MSLP=tp.MSLP, td_p_triple=tp.press_triple, td_T_triple=tp.T_triple,
             td_T_freeze=tp.T_freeze, td_T_icenuc=tp.T_icenuc, td_pow_icenuc=tp.pow_icenuc, td_R_d=tp.R_d,
             td_R_v=tp.R_v, td_cp_d=tp.cp_d, td_cp_v=tp.cp_v, td_cp_l=tp.cp_l, td_cp_i=tp.cp_i,
             td_LH_v0=tp.LH_v0, td_LH_s0=tp.LH_s0, td_grav=tp.grav)
    v.update(air_K_therm=g("thermal_conductivity_of_air"), air_D_vapor=g("diffusivity_of_water_vapor"),
             air_nu_air=g("kinematic_viscosity_of_air"))
    v.update(arg_sigma=g("surface_tension_water"), arg_M_w=g("molar_mass_water"), arg_R=g("universal_gas_constant"),
             arg_rho_w=rho_w, arg_rho_i=rho_i, arg_g=g("gravitational_acceleration"),
             arg_f1=g("ARG2000_f_coeff_1"), arg_f2=g("ARG2000_f_coeff_2"), arg_g1=g("ARG2000_g_coeff_1"),
             arg_g2=g("ARG2000_g_coeff_2"), arg_p1=g("ARG2000_pow_1"), arg_p2=g("ARG2000_pow_2"))
    v.update(kid_w1=g("prescribed_flow_w1"), kid_t1=g("prescribed_flow_t1"), kid_r_dry=g("r_dry"),
             kid_std_dry=g("std_dry"), kid_kappa=g("kappa"), com_prescribed_Nd=g("prescribed_Nd"))
    v.update(ne_tau_relax_liq=g("condensation_evaporation_timescale"),
             ne_tau_relax_ice=g("sublimation_deposition_timescale"))
    v.update(p0m_tau_precip=g("precipitation_timescale"), p0m_qc_0=g("specific_humidity_precipitation_threshold"),
             p0m_S_0=g("supersaturation_precipitation_threshold"))
    r0 = g("rain_drop_length_scale")
    v.update(rain_acnv_tau=g("rain_autoconversion_timescale"),
             rain_acnv_q_thr=g("cloud_liquid_water_specific_humidity_autoconversion_threshold"),
             rain_acnv_k=g("threshold_smooth_transition_steepness"),
             rain_r0=r0, rain_m0=FT(FT(4 / 3) * pi * rho_w * r0 ** FT(3)),
             rain_me=g("rain_mass_size_relation_coefficient_me"), rain_dm=g("rain_mass_size_relation_coefficient_delm"),
             rain_chim=g("rain_mass_size_relation_coefficient_chim"), rain_a0=FT(pi * r0 ** FT(2)),
             rain_ae=g("rain_cross_section_size_relation_coefficient_ae"),
             rain_da=g("rain_cross_section_size_relation_coefficient_dela"),
             rain_chia=g("rain_cross_section_size_relation_coefficient_chia"),
             rain_n0=g("rain_drop_size_distribution_coefficient_n0"),
             rain_a_vent=g("rain_ventillation_coefficient_a"), rain_b_vent=g("rain_ventillation_coefficient_b"),
             rain_chiv=g("rain_terminal_velocity_size_relation_coefficient_chiv"),
             rain_ve=g("rain_terminal_velocity_size_relation_coefficient_ve"),
             rain_dv=g("rain_terminal_velocity_size_relation_coefficient_delv"),
             rain_C_drag=g("rain_drop_drag_coefficient"), rain_rho_w=rho_w)
    s0 = g("snow_flake_length_scale")
    v.update(snow_acnv_tau=g("snow_autoconversion_timescale"),
             snow_acnv_q_thr=g("cloud_ice_specific_humidity_autoconversion_threshold"),
             snow_acnv_k=g("threshold_smooth_transition_steepness"),
             snow_r0=s0, snow_m0=FT(g("snow_mass_size_relation_coefficient_m0_prefactor") * s0 ** FT(2)),
             snow_me=g("snow_mass_size_relation_coefficient_me"), snow_dm=g("snow_mass_size_relation_coefficient_delm"),
             snow_chim=g("snow_mass_size_relation_coefficient_chim"),
             snow_a0=FT(g("snow_cross_section_size_relation_coefficient_a0_prefactor") * pi * s0 ** FT(2)),
             snow_ae=g("snow_cross_section_size_relation_coefficient_ae"),
             snow_da=g("snow_cross_section_size_relation_coefficient_dela"),
             snow_chia=g("snow_cross_section_size_relation_coefficient_chia"),
             snow_mu=g("snow_flake_size_distribution_coefficient_mu"),
             snow_nu=g("snow_flake_size_distribution_coefficient_nu"),
             snow_a_vent=g("snow_ventillation_coefficient_a"), snow_b_vent=g("snow_ventillation_coefficient_b"),
             snow_chiv=g("snow_terminal_velocity_size_relation_coefficient_chiv"),
             snow_v0=FT(FT(2) ** FT(9 / 4) * s0 ** FT(1 / 4)),
             snow_ve=g("snow_terminal_velocity_size_relation_coefficient_ve"),
             snow_dv=g("snow_terminal_velocity_size_relation_coefficient_delv"))
    i0 = g("cloud_ice_crystals_length_scale")
    v.update(ice_r0=i0, ice_m0=FT(FT(4 / 3) * pi * rho_i * i0 ** FT(3)),
             ice_me=g("cloud_ice_mass_size_relation_coefficient_me"),
             ice_dm=g("cloud_ice_mass_size_relation_coefficient_delm"),
             ice_chim=g("cloud_ice_mass_size_relation_coefficient_chim"),
             ice_n0=g("cloud_ice_size_distribution_coefficient_n0"), ice_r_ice_snow=g("ice_snow_threshold_radius"))
    v.update(ce_liq_rai=g("cloud_liquid_rain_collision_efficiency"),
             ce_liq_sno=g("cloud_liquid_snow_collision_efficiency"),
             ce_ice_rai=g("cloud_ice_rain_collision_efficiency"), ce_ice_sno=g("cloud_ice_snow_collision_efficiency"),
             ce_rai_sno=g("rain_snow_collision_efficiency"))
    rf = _RF_IDS.get(rain_formation, rain_formation) if rain_formation is not None else 0
    slots = [FT(0)] * 11
    if rf == 1:
        slots[0:4] = [g("KK2000_autoconversion_coeff_A"), g("KK2000_autoconversion_coeff_a"),
                      g("KK2000_autoconversion_coeff_b"), g("KK2000_autoconversion_coeff_c")]
        slots[8:11] = [g("KK2000_accretion_coeff_A"), g("KK2000_accretion_coeff_a"), g("KK2000_accretion_coeff_b")]
    elif rf == 2:
        slots[0:8] = [g("B1994_autoconversion_coeff_C"), g("B1994_autoconversion_coeff_a"),
                      g("B1994_autoconversion_coeff_b"), g("B1994_autoconversion_coeff_c"),
                      g("B1994_autoconversion_coeff_N_0"), g("B1994_autoconversion_coeff_d_low"),
                      g("B1994_autoconversion_coeff_d_high"), g("threshold_smooth_transition_steepness")]
        slots[8] = g("B1994_accretion_coeff_A")
    elif rf == 3:
        slots[0:7] = [g("TC1980_autoconversion_coeff_a"), g("TC1980_autoconversion_coeff_b"),
                      g("TC1980_autoconversion_coeff_D"), g("TC1980_autoconversion_coeff_r_0"),
                      g("TC1980_autoconversion_coeff_me_liq"), FT(FT(4 / 3) * pi * rho_w),
                      g("threshold_smooth_transition_steepness")]
        slots[8] = g("TC1980_accretion_coeff_A")
    elif rf == 4:
        slots[0:4] = [g("LD2004_autoconversion_coeff_R_6C_0"), g("LD2004_autoconversion_coeff_E_0"), rho_w,
                      g("threshold_smooth_transition_steepness")]
    elif rf == 5:
        slots[0:2] = [g("Variable_time_scale_autoconversion_coeff_tau"),
                      g("Variable_time_scale_autoconversion_coeff_alpha")]
    for i in range(8):
        v[f"rf_acnv{i}"] = slots[i]
    for i in range(3):
        v[f"rf_accr{i}"] = slots[8 + i]
    v.update(sb_kcc=g("SB2006_collection_kernel_coeff_kcc"), sb_x_star=g("SB2006_raindrops_min_mass"),
             sb_rho0=g("SB2006_reference_air_density"),
             sb_A=g("SB2006_autoconversion_correcting_function_coeff_A"),
             sb_a=g("SB2006_autoconversion_correcting_function_coeff_a"),
             sb_b=g("SB2006_autoconversion_correcting_function_coeff_b"),
             sb_kcr=g("SB2006_collection_kernel_coeff_kcr"),
             sb_tau0=g("SB2006_accretion_correcting_function_coeff_tau0"),
             sb_c=g("SB2006_accretion_correcting_function_coeff_c"),
             sb_krr=g("SB2006_collection_kernel_coeff_krr"), sb_kappa_rr=g("SB2006_collection_kernel_coeff_kapparr"),
             sb_d=g("SB2006_raindrops_self_collection_coeff_d"),
             sb_Deq=g("SB2006_raindrops_equlibrium_mean_diameter"),
             sb_Dr_th=g("SB2006_raindrops_breakup_mean_diameter_threshold"),
             sb_kbr=g("SB2006_raindrops_breakup_coeff_kbr"), sb_kappa_br=g("SB2006_raindrops_breakup_coeff_kappabr"),
             sb_av=g("SB2006_ventilation_factor_coeff_av"), sb_bv=g("SB2006_ventilation_factor_coeff_bv"),
             sb_alpha=g("SB2006_raindrops_terminal_velocity_coeff_alpha"),
             sb_beta=g("SB2006_raindrops_terminal_velocity_coeff_beta"),
             sb_numadj_tau=g("SB2006_number_adjustment_timescale"),
             sb_nu_c=g("SB2006_cloud_gamma_distribution_parameter"),
             sb_xc_min=g("SB2006_cloud_droplets_min_mass"), sb_xc_max=g("SB2006_cloud_droplets_max_mass"),
             sb_xr_min=g("SB2006_raindrops_min_mass"), sb_xr_max=g("SB2006_raindrops_max_mass"),
             sb_N0_min=g("SB2006_raindrops_size_distribution_coeff_N0_min"),
             sb_N0_max=g("SB2006_raindrops_size_distribution_coeff_N0_max"),
             sb_lam_min=g("SB2006_raindrops_size_distribution_coeff_lambda_min"),
             sb_lam_max=g("SB2006_raindrops_size_distribution_coeff_lambda_max"), sb_rho_w=rho_w,
             sb_aR=g("SB2006_raindrops_terminal_velocity_coeff_aR"),
             sb_bR=g("SB2006_raindrops_terminal_velocity_coeff_bR"),
             sb_cR=g("SB2006_raindrops_terminal_velocity_coeff_cR"),
             sb_vel_rho0=g("SB2006_reference_air_density"),
             ch_rho0=g("Chen2022_table_B1_q_coeff"), ch_a1=g("Chen2022_table_B1_a1_coeff"),
             ch_a2=g("Chen2022_table_B1_a2_coeff"), ch_a3=g("Chen2022_table_B1_a3_coeff"),
             ch_a3_pow=g("Chen2022_table_B1_a3_pow_coeff"), ch_b1=g("Chen2022_table_B1_b1_coeff"),
             ch_b2=g("Chen2022_table_B1_b2_coeff"), ch_b3=g("Chen2022_table_B1_b3_coeff"),
             ch_b_rho=g("Chen2022_table_B1_b_rho_coeff"), ch_c1=g("Chen2022_table_B1_c1_coeff"),
             ch_c2=g("Chen2022_table_B1_c2_coeff"), ch_c3=g("Chen2022_table_B1_c3_coeff"))
    v["num_fcc_scale"] = g("flux_correction_scale")
    names = param_names()
    missing = [n for n in names if n not in v]
    extra = [n for n in v if n not in names]
    if missing or extra:
        raise RuntimeError(f"parameter table out of sync with include/kid_params.h: missing={missing} extra={extra}")
    return np.array([float(v[n]) for n in names], dtype=np.float64)
