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
