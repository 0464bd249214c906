// Initial condition on the device (sm_100a): the ShipwayHill2012 sounding, the hydrostatic density profile and the
// per-level initial state of the 1-D rainshaft, one thread per column, so that every ensemble member / calibration case
// can carry its own surface pressure p0 without a host ODE solve per member.
//
// Replaces (reference): init_profile (src/Common/initial_condition.jl:50-84), dρ_dz + ρ_ivp (:100-149, a Tsit5 solve whose
// dense output is sampled at the cell centres), initial_condition_1d (:155-231), the PySDM helpers
// (src/Common/pysdm_functions.jl:8-63) and initialise_state / the ρ_dry, T, q_surf fields of initialise_aux
// (src/Common/ode_utils.jl:23-54, src/K1DModel/ode_utils.jl:80-117).
// The IVP is integrated in double with classical RK4 from z_0 straight to the cell centres (`nsub` substeps per level,
// steps split at the sounding's kink z_1); at Δz/4 ≈ 4 m the result agrees with the adaptive solve to ~1e-11.
#pragma once
#include <cuda_runtime.h>

namespace kid {

template <class FT>
struct ICArgs {
    // sounding (kid_params z_0..2, rv_0..2, tht_0..2) and thermodynamic constants
    double z_0, z_1, z_2, rv_0, rv_1, rv_2, th_0, th_1, th_2;
    double R_d, R_v, cp_d, cp_v, grav, MSLP;
    double z_min, dz;
    int nz, nc, ncp, nv, nsub, dry;
    int i_tot, i_Na;      // state indices (i_Na < 0: no N_aer)
    const FT* p0;         // [ncp] surface pressure per column
    const FT* Nd;         // [ncp] prescribed_Nd per column
    FT* rho_dry;          // [nz][ncp]
    FT* T;                // [nz][ncp]
    FT* Y;                // [nv][nz][ncp] (zero-filled by the caller)
    FT* q_surf;           // [ncp]
};

template <class FT>
struct ICProfile {
    const ICArgs<FT>& A;
    __device__ explicit ICProfile(const ICArgs<FT>& a) : A(a) {}
    __device__ void sounding(double z, double& qv, double& th) const {
        const double rv0 = A.dry ? 0.0 : A.rv_0, rv1 = A.dry ? 0.0 : A.rv_1, rv2 = A.dry ? 0.0 : A.rv_2;
        const double rv = z < A.z_1 ? rv0 + (rv1 - rv0) / (A.z_1 - A.z_0) * (z - A.z_0) : rv1 + (rv2 - rv1) / (A.z_2 - A.z_1) * (z - A.z_1);
        qv = rv / (1.0 + rv);
        th = z < A.z_1 ? A.th_0 : A.th_1 + (A.th_2 - A.th_1) / (A.z_2 - A.z_1) * (z - A.z_1);
    }
    __device__ double theta_dry(double th, double qv) const {  // SDM_θ_dry
        const double r = qv / (1.0 - qv);
        return th * pow(1.0 + r * (A.R_v / A.R_d), A.R_d / A.cp_d);
    }
    __device__ double T_of(double th_dry, double rho_dry) const {  // SDM_T
        const double kap = A.R_d / A.cp_d;
        return th_dry * pow(rho_dry * th_dry / A.MSLP * A.R_d, kap / (1.0 - kap));
    }
    __device__ double rho_surface(double p0) const {  // ρ_0 of init_profile: SDM_ρ_dry at the surface, then SDM_ρ_of_ρ_dry
        double qv0, th0;
        sounding(A.z_0, qv0, th0);  // θ_0 = tht_0 (z_0 < z_1)
        const double kap = A.R_d / A.cp_d;
        double rho_dry0;
        if (qv0 > 0.0) {
            const double r = qv0 / (1.0 - qv0);
            rho_dry0 = p0 * (1.0 - 1.0 / (1.0 + 1.0 / (A.R_v / A.R_d) / r)) / (pow(p0 / A.MSLP, kap) * A.R_d * A.th_0);
            return rho_dry0 * (1.0 + r);
        }
        return p0 / (pow(p0 / A.MSLP, kap) * A.R_d * A.th_0);
    }
    __device__ double drho_dz(double z, double rho) const {  // initial_condition.jl:100-130
        double qv, th;
        sounding(z, qv, th);
        const double thd = theta_dry(th, qv);
        const double r = qv / (1.0 - qv);
        const double rho_dry = rho / (1.0 + r);
        const double T = T_of(thd, rho_dry);
        double R_m = A.R_d, cp_m = A.cp_d, p = rho_dry * A.R_d * T;
        if (qv > 0.0) {
            R_m = A.R_v / (1.0 / r + 1.0) + A.R_d / (1.0 + r);
            cp_m = A.cp_v / (1.0 / r + 1.0) + A.cp_d / (1.0 + r);
            p = rho_dry * (1.0 + r) * R_m * T;  // SDM_p
        }
        const double rho_sdm = p / R_m / T;
        return A.grav / T * rho_sdm * (R_m / cp_m - 1.0) / R_m;
    }
    __device__ double rk4(double z, double rho, double h) const {
        const double k1 = drho_dz(z, rho);
        const double k2 = drho_dz(z + 0.5 * h, rho + 0.5 * h * k1);
        const double k3 = drho_dz(z + 0.5 * h, rho + 0.5 * h * k2);
        const double k4 = drho_dz(z + h, rho + h * k3);
        return rho + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
    }
    // integrate from za to zb (zb > za), splitting at the kink of the sounding
    __device__ double advance(double za, double zb, double rho, int nsub) const {
        if (za < A.z_1 && A.z_1 < zb) {
            rho = advance(za, A.z_1, rho, nsub);
            return advance(A.z_1, zb, rho, nsub);
        }
        const double h = (zb - za) / nsub;
        for (int s = 0; s < nsub; ++s) rho = rk4(za + s * h, rho, h);
        return rho;
    }
};

template <class FT>
__global__ void init_shipway_hill_kernel(const ICArgs<FT> A) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.nc) return;
    ICProfile<FT> pr(A);
    const size_t plane = size_t(A.nz) * A.ncp;
    double z = A.z_0, rho = pr.rho_surface(double(A.p0[c]));
    // aux.q_surf = init_profile(kid_params, thermo_params, z = 0).qv (K1DModel/ode_utils.jl:96): the reference calls it
    // WITHOUT the dry flag, so q_surf is the moist sounding's surface humidity also for a dry initial state
    {
        const double rv_s = 0.0 < A.z_1 ? A.rv_0 + (A.rv_1 - A.rv_0) / (A.z_1 - A.z_0) * (0.0 - A.z_0)
                                        : A.rv_1 + (A.rv_2 - A.rv_1) / (A.z_2 - A.z_1) * (0.0 - A.z_1);
        A.q_surf[c] = FT(rv_s / (1.0 + rv_s));
    }
    const double Nd = double(A.Nd[c]);
    for (int k = 0; k < A.nz; ++k) {
        const double zc = A.z_min + (k + 0.5) * A.dz;
        if (zc > z) { rho = pr.advance(z, zc, rho, k == 0 ? A.nsub * 2 : A.nsub); z = zc; }
        double qv, th;
        pr.sounding(zc, qv, th);
        // initial_condition_1d: ρ = FT(ρ_profile(z)), everything after in FT
        const FT rho_ft = FT(rho), qv_ft = FT(qv);
        const FT rho_dry = FT(double(rho_ft) / (1.0 + double(qv_ft) / (1.0 - double(qv_ft))));
        const FT thd = FT(pr.theta_dry(double(FT(th)), double(qv_ft)));
        const FT T = FT(pr.T_of(double(thd), double(rho_dry)));
        const size_t pi = size_t(k) * A.ncp + c;
        A.rho_dry[pi] = rho_dry;
        A.T[pi] = T;
        A.Y[size_t(A.i_tot) * plane + pi] = qv_ft * rho_ft;
        if (A.i_Na >= 0) A.Y[size_t(A.i_Na) * plane + pi] = FT(Nd) * rho_dry / FT(1.225);
    }
}

// ------------------------------------------------------------------ initial_condition_0d (Box, col_sed)
// Regularised upper incomplete gamma function Q(a, x), a > 0, x >= 0 (series for x < a + 1, continued fraction otherwise),
// in double: SpecialFunctions.gamma_inc(a, x)[2] of src/Common/initial_condition.jl:257,275.
__device__ inline double gamma_inc_Q(double a, double x) {
    if (!(x > 0.0)) return 1.0;
    const double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 2000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return 1.0 - sum * exp(-x + a * log(x) - gln);
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -double(i) * (double(i) - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c;
        if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-17) break;
    }
    return exp(-x + a * log(x) - gln) * h;
}

template <class FT>
struct IC0dArgs {
    const FT *qt, *Nd, *k, *rho_dry;  // [ncp] per column: total water (cloud + rain), number, gamma shape, dry density
    FT* rho_dry_out;                  // [nz][ncp]
    FT* T_out;                        // [nz][ncp]
    FT* Y;                            // [nv][nz][ncp] (zero-filled by the caller)
    int nz, nc, ncp;
    int i_tot, i_liq, i_rai, i_Nl, i_Nr;  // state indices (-1: absent)
};
// initial_condition_0d (src/Common/initial_condition.jl:237-308): the box model's gamma-distribution state — liquid and
// rain split at a fixed 40 um radius — per column, copied to every level (run_KiD_col_sed_simulation.jl:80-103).
// The arithmetic follows the reference's FT roundings; the incomplete gamma function is evaluated in double.
template <class FT>
__global__ void init_gamma_0d_kernel(const IC0dArgs<FT> A) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.nc) return;
    const FT qt = A.qt[c], Nd = A.Nd[c], k = A.k[c], rho_dry = A.rho_dry[c];
    const FT Lt = rho_dry * qt / (FT(1) - qt);
    const FT rhow = FT(1000), radius_th = FT(40 * 1e-6);
    const FT thr = FT(4.0 / 3.0) * FT(M_PI) * (radius_th * radius_th * radius_th) * rhow / (Lt / Nd / k);
    // gamma_inc(thrshld, k + 1)[2]: a = thrshld, x = k + 1, as the reference writes it
    const FT mass_ratio = FT(gamma_inc_Q(double(thr), double(k) + 1.0));
    const FT rq_liq = mass_ratio * Lt, rq_rai = Lt - rq_liq;
    const FT num_ratio = FT(gamma_inc_Q(double(thr), double(k)));
    const FT N_liq = num_ratio * Nd, N_rai = Nd - N_liq;
    const size_t plane = size_t(A.nz) * A.ncp;
    for (int l = 0; l < A.nz; ++l) {
        const size_t pi = size_t(l) * A.ncp + c;
        A.rho_dry_out[pi] = rho_dry;
        A.T_out[pi] = FT(300);
        A.Y[size_t(A.i_tot) * plane + pi] = Lt;
        if (A.i_liq >= 0) A.Y[size_t(A.i_liq) * plane + pi] = rq_liq;
        if (A.i_rai >= 0) A.Y[size_t(A.i_rai) * plane + pi] = rq_rai;
        if (A.i_Nl >= 0) A.Y[size_t(A.i_Nl) * plane + pi] = N_liq;
        if (A.i_Nr >= 0) A.Y[size_t(A.i_Nr) * plane + pi] = N_rai;
    }
}

}  // namespace kid
