"""TEST INFRASTRUCTURE ONLY (see oracle/README): numpy restatement of the reference's calibration observables.

Follows (reference):
  get_variable_data_from_ODE   src/Common/helper_functions.jl:98-191
  ODEsolution2Gvector          calibration/KiDUtils.jl:329-340 (unfiltered), :342-376 (filtered)
The terminal velocity of "rainrate" / "rain averaged terminal velocity" is an upstream CloudMicrophysics call; the
caller passes it in (tests take it from the C++ oracle's aux field).  Parity unpinned against Julia, like the oracle.
"""
import numpy as np


def get_variable_data_from_ODE(u: dict, rho_dry: np.ndarray, var: str, vt: np.ndarray | None = None) -> np.ndarray:
    """u: {state variable name: [nz] array} of ONE column; returns [nz] in the dtype of the state."""
    rho = rho_dry + u["ρq_tot"]
    if var == "qt":
        return u["ρq_tot"] / rho
    if var == "ql":
        return u["ρq_liq"] / rho
    if var == "qr":
        return u["ρq_rai"] / rho
    if var == "qv":
        return (u["ρq_tot"] - u["ρq_liq"]) / rho
    if var == "qlr":
        return (u["ρq_liq"] + u["ρq_rai"]) / rho
    qtot = u["ρq_tot"] / rho
    if var == "rt":
        return qtot / (1 - qtot)
    if var == "rl":
        return u["ρq_liq"] / rho / (1 - qtot)
    if var == "rr":
        return u["ρq_rai"] / rho / (1 - qtot)
    if var == "rv":
        return (u["ρq_tot"] - u["ρq_liq"]) / rho / (1 - qtot)
    if var == "rlr":
        return (u["ρq_liq"] + u["ρq_rai"]) / rho / (1 - qtot)
    if var == "Nl":
        return u["N_liq"]
    if var == "Nr":
        return u["N_rai"]
    if var == "Na":
        return u["N_aer"]
    if var == "rho":
        return rho
    if var == "rain averaged terminal velocity":
        return vt
    if var == "rainrate":
        return (u["ρq_rai"] / rho) * rho * vt * 3600
    raise ValueError(f'Data name "{var}" not recognized!!')


def z_top(u: dict, rho_dry: np.ndarray, z_span: float, rain: dict) -> float:
    """"Z_top" (helper_functions.jl:164-185) of one column.  rain: n0, r0, m0, me, Δm, χm of CMP.Rain (Marshall-Palmer);
    CMD.radar_reflectivity_1M restated as 10 log10(720 n0 λ^-7 / 1e-18) [UPSTREAM, recollected]."""
    import math
    rho = rho_dry + u["ρq_tot"]
    qliq, qrai = u["ρq_liq"] / rho, u["ρq_rai"] / rho
    qlr = qliq + qrai
    nz = qlr.size
    zt = nz  # 1-based
    while qlr[zt - 1] < 1e-3 and zt > 1:
        zt -= 1
    dz = z_span / (nz + 1)  # the reference takes length(z) of the FACE field
    steps = int(round(500.0 / dz))
    ind = max(zt - steps + 1, 1)
    qr = float(np.mean(qrai[ind - 1:zt]))
    rh = float(np.mean(rho[ind - 1:zt]))
    me = rain["me"] + rain["Δm"]
    if qr > 0 and rh > 0:
        li = (rh * qr * rain["r0"] ** me / (rain["χm"] * rain["m0"] * rain["n0"] * math.gamma(me + 1.0))) ** (1.0 / (me + 1.0))
    else:
        li = 0.0
    if li <= 0.0:
        return -300.0
    Z = 10.0 * (math.log10(720.0 * rain["n0"]) + 7.0 * math.log10(li) + 18.0)
    if Z != Z:
        return 0.0
    return Z


def gvector_unfiltered(states: list, rho_dry, variables, vts=None) -> np.ndarray:
    """states: one {name: [nz]} dict per saved time (KiDUtils.jl:329-340)."""
    out = []
    for i, u in enumerate(states):
        for var in variables:
            out.append(np.asarray(get_variable_data_from_ODE(u, rho_dry, var, None if vts is None else vts[i]), dtype=np.float64))
    return np.concatenate(out)


def gvector_filtered(states: list, rho_dry, variables, filt: dict, vts=None) -> np.ndarray:
    """KiDUtils.jl:342-376: mean over nt_per_filtered_cell snapshots x nz_per_filtered_cell levels, in Float64."""
    nzf, per = np.asarray(filt["nz_filtered"]), np.asarray(filt["nz_per_filtered_cell"])
    nz = nzf * per
    ntf, ntp = filt["nt_filtered"], filt["nt_per_filtered_cell"]
    out = []
    for i in range(ntf):
        data = np.zeros((int(nz.sum()), ntp))
        for j in range(ntp):
            s = i * ntp + j
            for k, var in enumerate(variables):
                a = int(nz[:k].sum())
                data[a:a + nz[k], j] = get_variable_data_from_ODE(states[s], rho_dry, var, None if vts is None else vts[s])
        for j in range(len(variables)):
            for k in range(nzf[j]):
                a = int(nz[:j].sum()) + k * per[j]
                out.append(data[a:a + per[j], :].mean())
    return np.asarray(out)
