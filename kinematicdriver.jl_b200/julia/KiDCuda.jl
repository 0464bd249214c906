"""
    KiDCuda — thin `ccall` shim that lets KinematicDriver.jl's drivers run the rhs + SSPRK33 loop on
    libkid_cuda.so (include/kid_cuda.h) instead of `ODE.solve(problem, ODE.SSPRK33(); ...)`.

UNTESTED IN THIS REPOSITORY'S CI: no Julia toolchain exists in the build image.  The Python host
mirror (kid_b200/) exercises exactly the same C ABI and is what the parity tests drive.

Usage inside e.g. test/experiments/KiD_driver/run_KiD_simulation.jl, replacing lines 159-166:

    sol = KiDCuda.solve(moisture, precip, Y, aux, (FT(opts["t_ini"]), FT(opts["t_end"]));
                        model = :K1D, dt = TS.dt, saveat = TS.dt_io, space = space, callback = affect_output!)
"""
module KiDCuda

import ClimaCore as CC
import CloudMicrophysics.Parameters as CMP
import KinematicDriver.Common as CO

const libkid = get(ENV, "KID_CUDA_LIB", "libkid_cuda.so")

# ---------------------------------------------------------------- C structs (include/kid_cuda.h)
Base.@kwdef mutable struct KidConfig
    abi_version::Int32 = 1
    params_version::Int32 = 2
    model::Int32 = 1
    moisture::Int32 = 0
    precip::Int32 = 0
    rain_formation::Int32 = 0
    sedimentation::Int32 = 0
    float_type::Int32 = 1
    nz::Int32 = 1
    nc::Int32 = 1
    helem::Int32 = 0
    npoly::Int32 = 0
    helem_global::Int32 = 0
    helem_offset::Int32 = 0
    precip_sources::Int32 = 1
    precip_sinks::Int32 = 1
    open_system_activation::Int32 = 0
    local_activation::Int32 = 0
    qtot_flux_correction::Int32 = 0
    device::Int32 = 0
    reserved::NTuple{8, Int32} = ntuple(_ -> Int32(0), 8)
    z_min::Float64 = 0.0
    z_max::Float64 = 1.0
    dt::Float64 = 1.0
    domain_width::Float64 = 0.0
    domain_height::Float64 = 0.0
end

struct KidError <: Exception
    msg::String
end
check(rc) = rc == 0 ? nothing : throw(KidError(unsafe_string(ccall((:kid_last_error, libkid), Cstring, ()))))

# ---------------------------------------------------------------- style -> enum
moisture_enum(::CO.EquilibriumMoisture) = Int32(0)
moisture_enum(::CO.NonEquilibriumMoisture) = Int32(1)
moisture_enum(ms) = error("$(typeof(ms).name.name) is not on the device hot path of libkid_cuda (no CPU fallback)")
precip_enum(::CO.NoPrecipitation) = Int32(0)
precip_enum(::CO.Precipitation0M) = Int32(1)
precip_enum(::CO.Precipitation1M) = Int32(2)
precip_enum(::CO.Precipitation2M) = Int32(3)
precip_enum(ps) = error("$(typeof(ps).name.name) is not on the device hot path of libkid_cuda (no CPU fallback)")

# rain-formation / sedimentation choice -> enum kid_rain_formation / kid_sedimentation (src/Common/helper_functions.jl:22-93)
rain_formation_enum(::Any) = Int32(0)
function rain_formation_enum(ps::CO.Precipitation1M)
    rf = ps.rain_formation
    rf isa CMP.KK2000 && return Int32(1)
    rf isa CMP.B1994 && return Int32(2)
    rf isa CMP.TC1980 && return Int32(3)
    rf isa CMP.LD2004 && return Int32(4)
    rf isa CMP.VarTimescaleAcnv && return Int32(5)
    return Int32(0)                                            # CMP.Rain: CliMA_1M
end
rain_formation_enum(ps::CO.Precipitation2M) = hasproperty(ps.rain_formation.pdf_r, :N0_min) ? Int32(6) : Int32(7)  # SB2006 / SB2006NL
sedimentation_enum(::Any) = Int32(0)
sedimentation_enum(ps::CO.Precipitation1M) =
    ps.sedimentation isa CMP.Chen2022VelType ? error("Chen2022 sedimentation with Precipitation1M is not on the device path") : Int32(0)
sedimentation_enum(ps::CO.Precipitation2M) = ps.sedimentation isa CMP.Chen2022VelTypeRain ? Int32(2) : Int32(1)

# The flat parameter table: every entry of include/kid_params.h, read from the structs the reference
# already built (thermo_params, air_params, activation_params, kid_params, common_params, the style
# structs).  `param_table` must list the names in the enum order of kid_params.h; the values are
# whatever the reference structs hold (already of type FT), widened to Float64.
function param_table(aux, moisture, precip; fcc_scale = 1.0)::Vector{Float64}
    tp, ap, act, cp = aux.thermo_params, aux.air_params, aux.activation_params, aux.common_params
    kp = hasproperty(aux, :kid_params) ? aux.kid_params : nothing
    TDP = CO.TD.Parameters
    FT = eltype(tp)
    z(n) = zeros(Float64, n)
    v = Float64[]
    # td_* (16), air_* (3), arg_* (12)
    append!(v, Float64[TDP.T_0(tp), TDP.MSLP(tp), TDP.press_triple(tp), TDP.T_triple(tp), TDP.T_freeze(tp),
                       TDP.T_icenuc(tp), TDP.pow_icenuc(tp), TDP.R_d(tp), TDP.R_v(tp), TDP.cp_d(tp), TDP.cp_v(tp),
                       TDP.cp_l(tp), TDP.cp_i(tp), TDP.LH_v0(tp), TDP.LH_s0(tp), TDP.grav(tp)])
    append!(v, Float64[ap.K_therm, ap.D_vapor, ap.ν_air])
    append!(v, Float64[act.σ, act.M_w, act.R, act.ρ_w, act.ρ_i, act.g, act.f1, act.f2, act.g1, act.g2, act.p1, act.p2])
    # kid_* (5), com_prescribed_Nd
    append!(v, kp === nothing ? z(5) : Float64[kp.w1, kp.t1, kp.r_dry, kp.std_dry, kp.κ])
    push!(v, Float64(cp.prescribed_Nd))
    # ne_tau_relax_liq / ice
    append!(v, moisture isa CO.NonEquilibriumMoisture ? Float64[moisture.liquid.τ_relax, moisture.ice.τ_relax] : z(2))
    # p0m_tau_precip, qc_0, S_0
    append!(v, precip isa CO.Precipitation0M ? Float64[precip.params.τ_precip, precip.params.qc_0, precip.params.S_0] : z(3))
    if precip isa CO.Precipitation1M
        r, sn, ic, ce, sed = precip.rain, precip.snow, precip.ice, precip.ce, precip.sedimentation
        append!(v, Float64[r.acnv1M.τ, r.acnv1M.q_threshold, r.acnv1M.k,
                           r.mass.r0, r.mass.m0, r.mass.me, r.mass.Δm, r.mass.χm,
                           r.area.a0, r.area.ae, r.area.Δa, r.area.χa, r.pdf.n0, r.vent.a, r.vent.b,
                           sed.rain.χv, sed.rain.ve, sed.rain.Δv, sed.rain.C_drag, sed.rain.ρw])
        append!(v, Float64[sn.acnv1M.τ, sn.acnv1M.q_threshold, sn.acnv1M.k,
                           sn.mass.r0, sn.mass.m0, sn.mass.me, sn.mass.Δm, sn.mass.χm,
                           sn.area.a0, sn.area.ae, sn.area.Δa, sn.area.χa, sn.pdf.μ, sn.pdf.ν, sn.vent.a, sn.vent.b,
                           sed.snow.χv, sed.snow.v0, sed.snow.ve, sed.snow.Δv])
        append!(v, Float64[ic.mass.r0, ic.mass.m0, ic.mass.me, ic.mass.Δm, ic.mass.χm, ic.pdf.n0, ic.r_ice_snow])
        append!(v, Float64[ce.e_lcl_rai, ce.e_lcl_sno, ce.e_icl_rai, ce.e_icl_sno, ce.e_rai_sno])
        rf = precip.rain_formation
        slots = z(11)
        if rf isa CMP.KK2000
            slots[1:4] = Float64[rf.acnv.A, rf.acnv.a, rf.acnv.b, rf.acnv.c]; slots[9:11] = Float64[rf.accr.A, rf.accr.a, rf.accr.b]
        elseif rf isa CMP.B1994
            slots[1:8] = Float64[rf.acnv.C, rf.acnv.a, rf.acnv.b, rf.acnv.c, rf.acnv.N_0, rf.acnv.d_low, rf.acnv.d_high, rf.acnv.k]
            slots[9] = rf.accr.A
        elseif rf isa CMP.TC1980
            slots[1:7] = Float64[rf.acnv.a, rf.acnv.b, rf.acnv.D, rf.acnv.r_0, rf.acnv.me_liq, rf.acnv.m0_liq_coeff, rf.acnv.k]
            slots[9] = rf.accr.A
        elseif rf isa CMP.LD2004
            slots[1:4] = Float64[rf.R_6C_0, rf.E_0, rf.ρ_w, rf.k]
        elseif rf isa CMP.VarTimescaleAcnv
            slots[1:2] = Float64[rf.τ, rf.α]
        end
        append!(v, slots)
    else
        append!(v, z(20 + 20 + 7 + 5 + 11))
    end
    if precip isa CO.Precipitation2M
        sb, vel = precip.rain_formation, precip.sedimentation
        limited = hasproperty(sb.pdf_r, :N0_min)
        append!(v, Float64[sb.acnv.kcc, sb.acnv.x_star, sb.acnv.ρ0, sb.acnv.A, sb.acnv.a, sb.acnv.b,
                           sb.accr.kcr, sb.accr.τ0, sb.accr.c, sb.self.krr, sb.self.κrr, sb.self.d,
                           sb.brek.Deq, sb.brek.Dr_th, sb.brek.kbr, sb.brek.κbr,
                           sb.evap.av, sb.evap.bv, sb.evap.α, sb.evap.β, sb.numadj.τ,
                           sb.pdf_c.νc, sb.pdf_c.xc_min, sb.pdf_c.xc_max, sb.pdf_r.xr_min, sb.pdf_r.xr_max,
                           limited ? sb.pdf_r.N0_min : 0, limited ? sb.pdf_r.N0_max : 0,
                           limited ? sb.pdf_r.λ_min : 0, limited ? sb.pdf_r.λ_max : 0, sb.pdf_r.ρw,
                           ])
        if vel isa CMP.Chen2022VelTypeRain   # sedimentation_choice = "Chen2022" (helper_functions.jl:69-70)
            append!(v, z(4))
            append!(v, Float64[vel.ρ0, vel.a[1], vel.a[2], vel.a[3], vel.a3_pow, vel.b[1], vel.b[2], vel.b[3], vel.b_ρ,
                               vel.c[1], vel.c[2], vel.c[3]])
        else
            append!(v, Float64[vel.aR, vel.bR, vel.cR, vel.ρ0])
            append!(v, z(12))
        end
    else
        append!(v, z(35 + 12))
    end
    push!(v, Float64(fcc_scale))  # num_fcc_scale
    return v
end

# ---------------------------------------------------------------- the ODE.solve replacement
"""
    solve(moisture, precip, Y, aux, tspan; model, dt, saveat, callback = nothing, columns = 1)

Uploads `parent(Y)` (host layout [column][variable][level] == ClimaCore's column parent array) and the
time-invariant ρ_dry/T of `aux.thermo_variables`, steps on the device between save times, downloads the
state into `Y` at every save time, calls `callback(Y, aux_getter, t)` there, and returns `(; t, u)`.
"""
function solve(moisture, precip, Y, aux, tspan; model::Symbol, dt, saveat, nz, z_min, z_max, callback = nothing, device = 0)
    FT = eltype(parent(Y))
    cfg = KidConfig(model = Int32(Dict(:Box => 0, :K1D => 1, :K1D_col_sed => 2, :K2D => 3)[model]),
                    moisture = moisture_enum(moisture), precip = precip_enum(precip),
                    rain_formation = rain_formation_enum(precip), sedimentation = sedimentation_enum(precip),
                    float_type = Int32(FT == Float32 ? 0 : 1), nz = Int32(nz), nc = Int32(1),
                    precip_sources = Int32(aux.common_params.precip_sources), precip_sinks = Int32(aux.common_params.precip_sinks),
                    open_system_activation = Int32(aux.common_params.open_system_activation),
                    local_activation = Int32(aux.common_params.local_activation),
                    qtot_flux_correction = Int32(hasproperty(aux, :kid_params) ? aux.kid_params.qtot_flux_correction : 0),
                    device = Int32(device), z_min = z_min, z_max = z_max, dt = dt)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:kid_create, libkid), Cint, (Ref{KidConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    try
        p = param_table(aux, moisture, precip)
        check(ccall((:kid_set_params, libkid), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Int32}), h[], p, length(p), 1, C_NULL))
        ρ_dry = Array(vec(parent(aux.thermo_variables.ρ_dry)))
        T = Array(vec(parent(aux.thermo_variables.T)))
        check(ccall((:kid_set_profiles, libkid), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32), h[], ρ_dry, T, 0))
        if hasproperty(aux, :q_surf)
            check(ccall((:kid_set_column_scalar, libkid), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h[], 1, FT[aux.q_surf]))
        end
        Yh = parent(Y)                       # (nz, nvars) column-major == [variable][level]
        check(ccall((:kid_set_state, libkid), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h[], Yh))
        ts, us = FT[], typeof(Y)[]
        t = FT(tspan[1])
        while t < tspan[2]
            n = Int(round(min(saveat, tspan[2] - t) / dt))
            check(ccall((:kid_step, libkid), Cint, (Ptr{Cvoid}, Float64, Int32), h[], Float64(t), n))
            t += n * dt
            check(ccall((:kid_get_state, libkid), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h[], Yh))
            push!(ts, t); push!(us, copy(Y))
            callback === nothing || callback(Y, (field, out) -> check(ccall((:kid_get_aux, libkid), Cint,
                (Ptr{Cvoid}, Float64, Int32, Ptr{Cvoid}), h[], Float64(t), field, out)), t)
        end
        return (; t = ts, u = us)
    finally
        ccall((:kid_destroy, libkid), Cvoid, (Ptr{Cvoid},), h[])
    end
end

# ------------------------------------------------------------------ batched ensembles (run_ensembles on one handle)
"""
    Ensemble: one device handle whose columns are independent simulations (ensemble member x case).

    ens = create_ensemble(moisture, precip, Ys, auxs; model = :K1D, dt, nz, z_min, z_max)
    step_to!(ens, t)                    # kid_step up to time t (a multiple of dt after the current time)
    destroy!(ens)

`Ys[i]`, `auxs[i]` are what the reference builds per simulation (`CO.initialise_state`, `K1D.initialise_aux`,
calibration/KiDUtils.jl:83-144); column i takes its state, ρ_dry/T profiles, w1, q_surf, prescribed_Nd and its whole
parameter set from them.  Replaces the `Threads.@threads` loop of `run_ensembles` (calibration/OptimizationUtils.jl:206-209).
"""
mutable struct Ensemble
    h::Ptr{Cvoid}
    t::Float64
    dt::Float64
    nc::Int
    nz::Int
end

function create_ensemble(moisture, precip, Ys::Vector, auxs::Vector; model::Symbol = :K1D, dt, nz, z_min, z_max, t_ini = 0.0, device = 0)
    FT = eltype(parent(Ys[1]))
    nc = length(Ys)
    a1 = auxs[1]
    cfg = KidConfig(model = Int32(Dict(:Box => 0, :K1D => 1, :K1D_col_sed => 2)[model]),
                    moisture = moisture_enum(moisture), precip = precip_enum(precip),
                    rain_formation = rain_formation_enum(precip), sedimentation = sedimentation_enum(precip),
                    float_type = Int32(FT == Float32 ? 0 : 1), nz = Int32(nz), nc = Int32(nc),
                    precip_sources = Int32(a1.common_params.precip_sources), precip_sinks = Int32(a1.common_params.precip_sinks),
                    open_system_activation = Int32(a1.common_params.open_system_activation),
                    local_activation = Int32(a1.common_params.local_activation),
                    qtot_flux_correction = Int32(hasproperty(a1, :kid_params) ? a1.kid_params.qtot_flux_correction : 0),
                    device = Int32(device), z_min = z_min, z_max = z_max, dt = dt)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:kid_create, libkid), Cint, (Ref{KidConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    tables = reduce(hcat, [param_table(a, moisture, precip) for a in auxs])          # one parameter set per column
    c2s = Int32.(0:(nc - 1))
    check(ccall((:kid_set_params, libkid), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Int32}),
                h[], tables, size(tables, 1), nc, c2s))
    ρ_dry = reduce(hcat, [Array(vec(parent(a.thermo_variables.ρ_dry))) for a in auxs])  # (nz, nc) column-major = [column][level]
    T = reduce(hcat, [Array(vec(parent(a.thermo_variables.T))) for a in auxs])
    check(ccall((:kid_set_profiles, libkid), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32), h[], ρ_dry, T, 1))
    if hasproperty(a1, :kid_params)
        check(ccall((:kid_set_column_scalar, libkid), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h[], 0, FT[a.kid_params.w1 for a in auxs]))
    end
    if hasproperty(a1, :q_surf)
        check(ccall((:kid_set_column_scalar, libkid), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h[], 1, FT[a.q_surf for a in auxs]))
    end
    check(ccall((:kid_set_column_scalar, libkid), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h[], 2,
                FT[a.common_params.prescribed_Nd for a in auxs]))
    Yall = reduce((a, b) -> cat(a, b; dims = 3), [reshape(Array(parent(Y)), nz, :, 1) for Y in Ys])  # (nz, nvars, nc) = [column][variable][level]
    check(ccall((:kid_set_state, libkid), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h[], Yall))
    return Ensemble(h[], Float64(t_ini), Float64(dt), nc, nz)
end

function step_to!(ens::Ensemble, t)
    n = Int(round((t - ens.t) / ens.dt))
    n >= 0 && abs(ens.t + n * ens.dt - t) < 1e-9 * max(1.0, abs(t)) ||
        error("save time $t is not a multiple of dt after t = $(ens.t): the device path saves at step boundaries only")
    n > 0 && check(ccall((:kid_step, libkid), Cint, (Ptr{Cvoid}, Float64, Int32), ens.h, ens.t, n))
    ens.t += n * ens.dt
    return ens
end

destroy!(ens::Ensemble) = ccall((:kid_destroy, libkid), Cvoid, (Ptr{Cvoid},), ens.h)

# ------------------------------------------------------------------ initial condition on the device
"""
    init_shipway_hill!(h, kid_params, p0::Vector{FT})

Replaces `ρ_ivp` + `initial_condition_1d` + `initialise_state` (src/Common/initial_condition.jl:50-231) for an ensemble
whose members differ in the surface pressure: one hydrostatic profile and initial state per column, built on the device.
"""
function init_shipway_hill!(h::Ptr{Cvoid}, kp, p0::Vector{FT}; dry = false) where {FT}
    snd = Float64[kp.z_0, kp.z_1, kp.z_2, kp.rv_0, kp.rv_1, kp.rv_2, kp.tht_0, kp.tht_1, kp.tht_2]
    check(ccall((:kid_init_shipway_hill, libkid), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Int32), h, snd, p0, dry ? 1 : 0))
end

# ------------------------------------------------------------------ calibration observables on the device
const OBS_NAMES = ["qt", "ql", "qr", "qv", "qlr", "rt", "rl", "rr", "rv", "rlr", "Nl", "Nr", "Na", "rho",
                   "rain averaged terminal velocity", "rainrate"]   # enum kid_obs_var

"""
    obs_configure(h, variables, filter) / obs_accumulate(h, time_cell) / obs_get(h, nc)

Device replacement of `ODEsolution2Gvector(ODEsol, aux, precip, variables[, filter])` (calibration/KiDUtils.jl:329-376):
call `obs_accumulate` at every `filter["saveat_t"]` time instead of saving the state; `obs_get` returns the G vectors,
one column per simulated column.
"""
function obs_configure(h::Ptr{Cvoid}, variables::Vector{String}, filter::Dict)
    ids = Int32[findfirst(==(v), OBS_NAMES) - 1 for v in variables]
    per = Int32.(filter["nz_per_filtered_cell"])
    n_out = Ref{Int64}(0)
    check(ccall((:kid_obs_configure, libkid), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Ref{Int64}),
                h, length(ids), ids, per, filter["nt_filtered"], filter["nt_per_filtered_cell"], n_out))
    return n_out[]
end
obs_accumulate(h::Ptr{Cvoid}, time_cell::Integer) =
    check(ccall((:kid_obs_accumulate, libkid), Cint, (Ptr{Cvoid}, Int32), h, time_cell))
function obs_get(h::Ptr{Cvoid}, n_out::Integer, nc::Integer)
    G = zeros(Float64, n_out, nc)
    check(ccall((:kid_obs_get, libkid), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, G))
    return G
end

end # module
