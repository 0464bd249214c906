# ORACLE support (test infrastructure) — closes the "parity unpinned" gap the moment a Julia toolchain is available.
#
# Runs the UNMODIFIED reference (KinematicDriver.jl + its pinned deps) and dumps, per case,
#   <case>.json   metadata: styles, FT, grid, dt, save times, variable order, every parameter value used
#   <case>.Y.bin  the prognostic state at every save time  [time][variable][level]  (FT, little endian)
#   <case>.dY.bin one rhs!(dY, Y0, aux, t=0) evaluation    [variable][level]
#   <case>.aux.bin aux fields after that call (order in the json)
# into tests/golden/julia/.  tests/test_oracle_vs_julia_fixtures.py then compares the C++ oracle against them
# (it is skipped while the directory is empty).  Usage (on a machine with Julia >= 1.10):
#   julia --project=/path/to/KinematicDriver.jl/test oracle/julia/dump_reference_states.jl /path/to/KinematicDriver.jl tests/golden/julia
import JSON
import OrdinaryDiffEq as ODE
import ClimaCore as CC
import ClimaParams as CP
import CloudMicrophysics.Parameters as CMP

const REF = ARGS[1]
const OUT = ARGS[2]
mkpath(OUT)
import KinematicDriver
import KinematicDriver.Common as CO
import KinematicDriver.K1DModel as K1D
include(joinpath(REF, "test", "create_parameters.jl"))

function dump_case(name, FT, moisture_choice, precipitation_choice; rain_formation_choice = nothing, n_elem = 64,
                   t_end = 900.0, dt = 1.0, saveat = 300.0)
    toml_dict = override_toml_dict(CP.create_toml_dict(FT); w1 = FT(2), t1 = FT(600), p0 = FT(100000), prescribed_Nd = FT(1e8),
                                   κ = FT(0.9))
    common_params = create_common_parameters(toml_dict)
    kid_params = create_kid_parameters(toml_dict)
    thermo_params = create_thermodynamics_parameters(toml_dict)
    air_params = CMP.AirProperties(toml_dict)
    activation_params = CMP.AerosolActivationParameters(toml_dict)
    moisture = CO.get_moisture_type(moisture_choice, toml_dict)
    precip = CO.get_precipitation_type(precipitation_choice, toml_dict; rain_formation_choice)
    TS = CO.TimeStepping(FT(dt), FT(saveat), FT(t_end))
    space, face_space = K1D.make_function_space(FT, FT(0), FT(2000), n_elem)
    coord = CC.Fields.coordinate_field(space)
    ρ_profile = CO.ρ_ivp(FT, kid_params, thermo_params, "ShipwayHill2012")
    init = CO.initial_condition_1d.(FT, common_params, kid_params, thermo_params, (ρ_profile,), coord.z, "ShipwayHill2012")
    aux = K1D.initialise_aux(FT, init, common_params, kid_params, thermo_params, air_params, activation_params, TS,
                             nothing, face_space, moisture, precip, "ShipwayHill2012")
    Y = CO.initialise_state(moisture, precip, init)
    rhs! = K1D.make_rhs_function(moisture, precip, "ShipwayHill2012")
    dY = similar(Y)
    rhs!(dY, Y, aux, FT(0))
    vars = collect(string.(propertynames(Y)))
    open(joinpath(OUT, name * ".dY.bin"), "w") do io
        for v in propertynames(Y); write(io, vec(parent(getproperty(dY, v)))); end
    end
    auxnames = String[]
    open(joinpath(OUT, name * ".aux.bin"), "w") do io
        for grp in (:thermo_variables, :microph_variables, :velocities, :precip_sources, :activation_sources, :cloud_sources)
            f = getproperty(aux, grp)
            f === nothing && continue
            for v in propertynames(f)
                push!(auxnames, string(grp, ".", v)); write(io, vec(parent(getproperty(f, v))))
            end
        end
    end
    problem = ODE.ODEProblem(rhs!, Y, (FT(0), FT(t_end)), aux)
    sol = ODE.solve(problem, ODE.SSPRK33(), dt = FT(dt), saveat = FT(saveat))
    open(joinpath(OUT, name * ".Y.bin"), "w") do io
        for u in sol.u, v in propertynames(u); write(io, vec(parent(getproperty(u, v)))); end
    end
    meta = Dict("name" => name, "FT" => string(FT), "moisture" => moisture_choice, "precipitation" => precipitation_choice,
                "rain_formation" => rain_formation_choice, "n_elem" => n_elem, "z_min" => 0.0, "z_max" => 2000.0, "dt" => dt,
                "t" => collect(Float64.(sol.t)), "variables" => vars, "aux" => auxnames,
                "rho_dry" => collect(Float64.(vec(parent(aux.thermo_variables.ρ_dry)))),
                "T" => collect(Float64.(vec(parent(aux.thermo_variables.T)))), "q_surf" => Float64(aux.q_surf),
                "toml" => Dict(k => v["value"] for (k, v) in toml_dict.data if haskey(v, "value") && v["value"] isa Number),
                "versions" => Dict(string(p.name) => string(p.version) for p in values(Pkg.dependencies()) if p.is_direct_dep))
    open(joinpath(OUT, name * ".json"), "w") do io; JSON.print(io, meta); end
    @info "dumped $name"
end

import Pkg
for FT in (Float64, Float32), (ms, ps, rf) in (("EquilibriumMoisture", "Precipitation0M", nothing),
        ("NonEquilibriumMoisture", "Precipitation1M", "CliMA_1M"), ("NonEquilibriumMoisture", "Precipitation1M", "KK2000"),
        ("EquilibriumMoisture", "Precipitation2M", "SB2006"), ("NonEquilibriumMoisture", "Precipitation2M", "SB2006"),
        ("NonEquilibriumMoisture", "Precipitation2M", "SB2006NL"))
    dump_case("k1d_$(ms[1:2])_$(ps[end-1:end])_$(something(rf, "none"))_$(FT)", FT, ms, ps; rain_formation_choice = rf)
end
