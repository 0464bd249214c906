"""Shared parity metric of the test-suite.

For every prognostic variable v the error is
    err_v = max |a_v - b_v| / max( max|b_v| , FLOOR * group_scale )
where group_scale is the largest magnitude among the variables of the same kind (mass
densities ρq_* or number densities N_*).  The floor keeps fields that are pure round-off
noise in the reference formulation itself (e.g. ρq_sno in a warm run, N_aer in the box
model: `triangle_inequality_limiter(0, M) = M - sqrt(M*M)` is ±ulp(M), not 0) from
dominating a relative measure."""
import numpy as np

FLOOR = 1e-6


def group_of(name: str) -> str:
    return "N" if name.startswith("N_") else "q"


def errors(a, b, var_names):
    """a, b: [nc, nvars, nz] -> {var: err}"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = {}
    for i, v in enumerate(var_names):
        g = group_of(v)
        scale[g] = max(scale.get(g, 0.0), float(np.max(np.abs(b[:, i]))))
    out = {}
    for i, v in enumerate(var_names):
        den = max(float(np.max(np.abs(b[:, i]))), FLOOR * scale[group_of(v)], 1e-300)
        out[v] = float(np.max(np.abs(a[:, i] - b[:, i])) / den)
    return out


def field_error(a, b, floor_scale=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), floor_scale, 1e-300)
    return float(np.max(np.abs(a - b)) / den)


def rhs_errors(a, b, Y, var_names, dt, FT):
    """Tendency parity: err_v = max|a_v - b_v| / max(max|b_v|, 1000·eps(FT)·state_scale/dt).
    A tendency that is itself a difference of nearly equal numbers (e.g. condensation at equilibrium,
    (q_eq - q_liq)/τ ≈ 0) carries absolute noise ~eps·|q|/τ; it is measured against the change it
    produces in the state over one step, not against its own (vanishing) magnitude."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    eps = float(np.finfo(FT).eps)
    scale = {}
    for i, v in enumerate(var_names):
        g = group_of(v)
        scale[g] = max(scale.get(g, 0.0), float(np.max(np.abs(Y[:, i]))))
    out = {}
    for i, v in enumerate(var_names):
        den = max(float(np.max(np.abs(b[:, i]))), 1000.0 * eps * scale[group_of(v)] / dt, 1e-300)
        out[v] = float(np.max(np.abs(a[:, i] - b[:, i])) / den)
    return out


def pointwise_errors(a, b, var_names, floor=1e-3):
    """Pointwise relative error with an absolute floor: max over grid points of |a - b| / max(|b|, floor * max|b_v|).
    Unlike `errors` (normalised by the field maximum) a large relative error in a small-valued region shows, down to
    `floor` of the field maximum."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = {}
    for i, v in enumerate(var_names):
        m = float(np.max(np.abs(b[:, i])))
        if m == 0.0:
            out[v] = float(np.max(np.abs(a[:, i])))
            continue
        out[v] = float(np.max(np.abs(a[:, i] - b[:, i]) / np.maximum(np.abs(b[:, i]), floor * m)))
    return out


def column_integrated_number_error(a, b, var_names):
    """Relative error of Σ_levels (N_liq + N_rai + N_aer) per column (max over columns): insensitive to WHICH level a
    knife-edge decision (find_cloud_base!) activates on."""
    idx = [var_names.index(v) for v in ("N_liq", "N_rai", "N_aer") if v in var_names]
    a = np.asarray(a, dtype=np.float64)[:, idx].sum(axis=(1, 2))
    b = np.asarray(b, dtype=np.float64)[:, idx].sum(axis=(1, 2))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def cloud_base_levels(Y, var_names, threshold=1e6):
    """Lowest level with N_liq above `threshold` per column (-1: none)."""
    N = np.asarray(Y)[:, var_names.index("N_liq")]
    has = N > threshold
    return np.where(has.any(axis=1), has.argmax(axis=1), -1)


def on_backend(handle_type, fn, *args, **kwargs):
    """Run a product entry point with `handle_type` (the CPU oracle's handle class) in place of the CUDA library's handle:
    test plumbing only — the product API has no backend switch (kid_b200.model.Handle is patched for the call)."""
    from kid_b200 import model as M
    saved = M.Handle
    M.Handle = handle_type
    try:
        return fn(*args, **kwargs)
    finally:
        M.Handle = saved
