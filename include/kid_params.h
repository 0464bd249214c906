/*
 * kid_params.h — flat, versioned parameter table of libkid_cuda.so.
 *
 * The reference carries its physical parameters in Julia structs built from a
 * ClimaParams TOML dict (reference: test/create_parameters.jl:68-88,
 * src/Common/helper_functions.jl:4-93, src/Common/equation_types.jl:23-61,
 * src/Common/parameters.jl:18-41, src/K1DModel/parameters.jl:17-73).  Julia
 * struct layouts are not ABI-stable across CloudMicrophysics 0.29-0.32, so the
 * boundary takes ONE flat array of doubles, indexed by the enum below.  The
 * Julia shim (kinematicdriver.jl_b200/julia/KiDCuda.jl) fills it by reading
 * struct fields by name; the Python host mirror fills it from
 * kid_b200/parameters.py.  Every value is the struct field AS THE REFERENCE
 * HOLDS IT (already rounded to FLOAT_TYPE where FT=Float32), widened to double.
 *
 * Names: <owner>_<field>; owner = td (Thermodynamics), air (CMP.AirProperties),
 * arg (CMP.AerosolActivationParameters), kid (KinematicDriverParameters),
 * com (CommonParameters), ne (CloudLiquid/CloudIce relaxation), p0m
 * (CMP.Parameters0M), rain/snow/ice/liq/ce (1-moment CMP structs), rf
 * (1-moment rain-formation variants), sb (CMP.SB2006 + SB2006VelType),
 * num (numerical knobs that restate un-vendored operator constants).
 */
#ifndef KID_PARAMS_H
#define KID_PARAMS_H

#define KID_PARAMS_VERSION 2

#define KID_PARAM_LIST(X)                                                      \
    /* Thermodynamics.Parameters.ThermodynamicsParameters */                   \
    X(td_T_0) X(td_MSLP) X(td_p_triple) X(td_T_triple) X(td_T_freeze)          \
    X(td_T_icenuc) X(td_pow_icenuc) X(td_R_d) X(td_R_v) X(td_cp_d) X(td_cp_v)  \
    X(td_cp_l) X(td_cp_i) X(td_LH_v0) X(td_LH_s0) X(td_grav)                   \
    /* CMP.AirProperties */                                                    \
    X(air_K_therm) X(air_D_vapor) X(air_nu_air)                                \
    /* CMP.AerosolActivationParameters (ARG2000) */                            \
    X(arg_sigma) X(arg_M_w) X(arg_R) X(arg_rho_w) X(arg_rho_i) X(arg_g)        \
    X(arg_f1) X(arg_f2) X(arg_g1) X(arg_g2) X(arg_p1) X(arg_p2)                \
    /* K1D KinematicDriverParameters (device-relevant subset) */               \
    X(kid_w1) X(kid_t1) X(kid_r_dry) X(kid_std_dry) X(kid_kappa)               \
    /* CommonParameters */                                                     \
    X(com_prescribed_Nd)                                                       \
    /* NonEquilibriumMoisture: CloudLiquid.tau_relax, CloudIce.tau_relax */    \
    X(ne_tau_relax_liq) X(ne_tau_relax_ice)                                    \
    /* CMP.Parameters0M */                                                     \
    X(p0m_tau_precip) X(p0m_qc_0) X(p0m_S_0)                                   \
    /* CMP.Rain (+ Blk1MVelTypeRain) */                                        \
    X(rain_acnv_tau) X(rain_acnv_q_thr) X(rain_acnv_k)                         \
    X(rain_r0) X(rain_m0) X(rain_me) X(rain_dm) X(rain_chim)                   \
    X(rain_a0) X(rain_ae) X(rain_da) X(rain_chia) X(rain_n0)                   \
    X(rain_a_vent) X(rain_b_vent)                                              \
    X(rain_chiv) X(rain_ve) X(rain_dv) X(rain_C_drag) X(rain_rho_w)            \
    /* CMP.Snow (+ Blk1MVelTypeSnow) */                                        \
    X(snow_acnv_tau) X(snow_acnv_q_thr) X(snow_acnv_k)                         \
    X(snow_r0) X(snow_m0) X(snow_me) X(snow_dm) X(snow_chim)                   \
    X(snow_a0) X(snow_ae) X(snow_da) X(snow_chia) X(snow_mu) X(snow_nu)        \
    X(snow_a_vent) X(snow_b_vent)                                              \
    X(snow_chiv) X(snow_v0) X(snow_ve) X(snow_dv)                              \
    /* CMP.CloudIce */                                                         \
    X(ice_r0) X(ice_m0) X(ice_me) X(ice_dm) X(ice_chim) X(ice_n0)              \
    X(ice_r_ice_snow)                                                          \
    /* CMP.CollisionEff */                                                     \
    X(ce_liq_rai) X(ce_liq_sno) X(ce_ice_rai) X(ce_ice_sno) X(ce_rai_sno)      \
    /* 1-moment rain-formation variants (CMP.KK2000/B1994/TC1980/LD2004/       \
       VarTimescaleAcnv); meaning of the slots depends on                      \
       kid_config_t.rain_formation, see kid_b200/parameters.py */              \
    X(rf_acnv0) X(rf_acnv1) X(rf_acnv2) X(rf_acnv3) X(rf_acnv4) X(rf_acnv5)    \
    X(rf_acnv6) X(rf_acnv7) X(rf_accr0) X(rf_accr1) X(rf_accr2)                \
    /* CMP.SB2006: acnv, accr, self, brek, evap, numadj, pdf_c, pdf_r */       \
    X(sb_kcc) X(sb_x_star) X(sb_rho0) X(sb_A) X(sb_a) X(sb_b)                  \
    X(sb_kcr) X(sb_tau0) X(sb_c)                                               \
    X(sb_krr) X(sb_kappa_rr) X(sb_d)                                           \
    X(sb_Deq) X(sb_Dr_th) X(sb_kbr) X(sb_kappa_br)                             \
    X(sb_av) X(sb_bv) X(sb_alpha) X(sb_beta)                                   \
    X(sb_numadj_tau)                                                           \
    X(sb_nu_c) X(sb_xc_min) X(sb_xc_max)                                       \
    X(sb_xr_min) X(sb_xr_max) X(sb_N0_min) X(sb_N0_max)                        \
    X(sb_lam_min) X(sb_lam_max) X(sb_rho_w)                                    \
    /* CMP.SB2006VelType */                                                    \
    X(sb_aR) X(sb_bR) X(sb_cR) X(sb_vel_rho0)                                  \
    /* CMP.Chen2022VelTypeRain (Chen et al. 2022, Table B1): q = exp(rho0·ρ),  \
       a_i, a3·ρ^a3_pow, b_i - b_rho·ρ, c_i */                                 \
    X(ch_rho0) X(ch_a1) X(ch_a2) X(ch_a3) X(ch_a3_pow)                         \
    X(ch_b1) X(ch_b2) X(ch_b3) X(ch_b_rho) X(ch_c1) X(ch_c2) X(ch_c3)          \
    /* numerical knobs restating un-vendored ClimaCore operators */            \
    X(num_fcc_scale)

enum kid_param_id {
#define KID_X_ENUM(name) KID_P_##name,
    KID_PARAM_LIST(KID_X_ENUM)
#undef KID_X_ENUM
    KID_NPARAMS
};

#endif /* KID_PARAMS_H */
