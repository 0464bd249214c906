// Packed Float32 pairs for sm_100a: two grid points per thread in one 64-bit register pair.
//
// Blackwell (sm_100) has packed single-precision instructions — FFMA2 / FADD2 / FMUL2 in SASS, reached from PTX
// fma.rn.f32x2 / add.f32x2 / mul.f32x2 — that produce two results per issue slot (measured on the box:
// scripts/microbench/f32x2_peak.cu, profiles/r02_f32x2_peak.jsonl: the same flop rate as FFMA at half the
// instructions, and the freed issue slots are usable by the ALU / MUFU / LSU streams).  k1d_pair_kernel is bound by
// instruction issue, so it evaluates TWO vertically adjacent levels per thread and keeps every add / multiply / fma
// of the pair in these instructions.  Operations without a packed form (min/max, compares, selects, MUFU
// transcendentals) are done per component.
#pragma once
#include <cuda_runtime.h>

namespace kid {

#define KID_F2_DEV __device__ __forceinline__

struct F2 { float x, y; };   // x: lower level of the pair, y: upper level
struct B2 { bool x, y; };

namespace f2detail {
KID_F2_DEV unsigned long long pk(F2 a) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
KID_F2_DEV F2 up(unsigned long long r) { F2 a; asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a; }
}  // namespace f2detail

KID_F2_DEV F2 f2(float a) { return F2{a, a}; }
KID_F2_DEV F2 f2(float a, float b) { return F2{a, b}; }

KID_F2_DEV F2 operator+(F2 a, F2 b) { unsigned long long r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2detail::pk(a)), "l"(f2detail::pk(b))); return f2detail::up(r); }
KID_F2_DEV F2 operator-(F2 a, F2 b) { unsigned long long r; asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2detail::pk(a)), "l"(f2detail::pk(b))); return f2detail::up(r); }
KID_F2_DEV F2 operator*(F2 a, F2 b) { unsigned long long r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2detail::pk(a)), "l"(f2detail::pk(b))); return f2detail::up(r); }
// a * b + c
KID_F2_DEV F2 fma2(F2 a, F2 b, F2 c) {
    unsigned long long r;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2detail::pk(a)), "l"(f2detail::pk(b)), "l"(f2detail::pk(c)));
    return f2detail::up(r);
}
KID_F2_DEV F2 operator-(F2 a) { return F2{-a.x, -a.y}; }  // folds into the operand negation of the consuming FFMA2 / FADD2
KID_F2_DEV F2 operator+(F2 a, float b) { return a + f2(b); }
KID_F2_DEV F2 operator+(float a, F2 b) { return f2(a) + b; }
KID_F2_DEV F2 operator-(F2 a, float b) { return a - f2(b); }
KID_F2_DEV F2 operator-(float a, F2 b) { return f2(a) - b; }
KID_F2_DEV F2 operator*(F2 a, float b) { return a * f2(b); }
KID_F2_DEV F2 operator*(float a, F2 b) { return f2(a) * b; }
KID_F2_DEV F2 fma2(F2 a, float b, F2 c) { return fma2(a, f2(b), c); }
KID_F2_DEV F2 fma2(F2 a, F2 b, float c) { return fma2(a, b, f2(c)); }
KID_F2_DEV F2 fma2(F2 a, float b, float c) { return fma2(a, f2(b), f2(c)); }
KID_F2_DEV F2& operator+=(F2& a, F2 b) { a = a + b; return a; }
KID_F2_DEV F2& operator-=(F2& a, F2 b) { a = a - b; return a; }

// ---- per-component operations
KID_F2_DEV F2 vmax(F2 a, F2 b) { return F2{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }
KID_F2_DEV F2 vmin(F2 a, F2 b) { return F2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
KID_F2_DEV F2 vmax(F2 a, float b) { return F2{fmaxf(a.x, b), fmaxf(a.y, b)}; }
KID_F2_DEV F2 vmin(F2 a, float b) { return F2{fminf(a.x, b), fminf(a.y, b)}; }
KID_F2_DEV F2 vabs(F2 a) { return F2{fabsf(a.x), fabsf(a.y)}; }
KID_F2_DEV float rcp1(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KID_F2_DEV float ex2_1(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KID_F2_DEV float lg2_1(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KID_F2_DEV float sqrt_approx1(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KID_F2_DEV float rsqrt_approx1(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KID_F2_DEV F2 vrcp(F2 a) { return F2{rcp1(a.x), rcp1(a.y)}; }
// a / b as the Float32 build of the scalar kernels evaluates it (-use_fast_math: rcp.approx + mul, 2 ulp)
KID_F2_DEV F2 operator/(F2 a, F2 b) { return a * vrcp(b); }
KID_F2_DEV F2 operator/(float a, F2 b) { return a * vrcp(b); }
KID_F2_DEV F2 operator/(F2 a, float b) { return a * rcp1(b); }
KID_F2_DEV F2 vex2(F2 a) { return F2{ex2_1(a.x), ex2_1(a.y)}; }
KID_F2_DEV F2 vlg2(F2 a) { return F2{lg2_1(a.x), lg2_1(a.y)}; }
// the MUFU formulations of Mth<float> (kid_physics.cuh)
KID_F2_DEV F2 vexp(F2 a) { return vex2(a * 1.4426950408889634f); }
KID_F2_DEV F2 vlog(F2 a) { return vlg2(a) * 0.6931471805599453f; }
KID_F2_DEV F2 vpow(F2 a, float y) { return vex2(vlg2(a) * y); }
KID_F2_DEV F2 vcbrt(F2 a) { return vex2(vlg2(a) * 0.333333343f); }
// kid_physics.cuh::pow_small_int for two values (the exponent is uniform)
KID_F2_DEV F2 vpow_small_int(F2 x, float y) {
    if (y == 3.f) return x * x * x;
    if (y == 4.f) { const F2 x2 = x * x; return x2 * x2; }
    if (y == -5.f) { const F2 x2 = x * x; return vrcp(x2 * x2 * x); }
    if (y == 2.f) return x * x;
    return vpow(x, y);
}
KID_F2_DEV F2 vsqrt_fast(F2 a) { return F2{sqrt_approx1(a.x), sqrt_approx1(a.y)}; }
// Mth<float>::sqrt_lim for two values: rsqrt + one Newton step in FMA form, sqrt_lim(0) = 0 exactly
KID_F2_DEV F2 vsqrt_lim(F2 x) {
    const F2 y = F2{rsqrt_approx1(fmaxf(x.x, 1.17549435e-38f)), rsqrt_approx1(fmaxf(x.y, 1.17549435e-38f))};
    const F2 g = x * y, h = y * 0.5f;
    const F2 r = fma2(-g, g, x);
    return fma2(r, h, g);
}

// ---- masks
KID_F2_DEV B2 operator<(F2 a, F2 b) { return B2{a.x < b.x, a.y < b.y}; }
KID_F2_DEV B2 operator>(F2 a, F2 b) { return B2{a.x > b.x, a.y > b.y}; }
KID_F2_DEV B2 operator<=(F2 a, F2 b) { return B2{a.x <= b.x, a.y <= b.y}; }
KID_F2_DEV B2 operator<(F2 a, float b) { return B2{a.x < b, a.y < b}; }
KID_F2_DEV B2 operator>(F2 a, float b) { return B2{a.x > b, a.y > b}; }
KID_F2_DEV B2 operator<=(F2 a, float b) { return B2{a.x <= b, a.y <= b}; }
KID_F2_DEV B2 operator>=(F2 a, float b) { return B2{a.x >= b, a.y >= b}; }
KID_F2_DEV B2 veq(F2 a, float b) { return B2{a.x == b, a.y == b}; }
KID_F2_DEV B2 operator!(B2 a) { return B2{!a.x, !a.y}; }
KID_F2_DEV B2 operator&&(B2 a, B2 b) { return B2{a.x && b.x, a.y && b.y}; }
KID_F2_DEV B2 operator||(B2 a, B2 b) { return B2{a.x || b.x, a.y || b.y}; }
KID_F2_DEV bool any(B2 a) { return a.x || a.y; }
KID_F2_DEV bool all(B2 a) { return a.x && a.y; }
KID_F2_DEV F2 sel(B2 m, F2 a, F2 b) { return F2{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
KID_F2_DEV F2 sel(B2 m, F2 a, float b) { return F2{m.x ? a.x : b, m.y ? a.y : b}; }
KID_F2_DEV F2 sel(B2 m, float a, F2 b) { return F2{m.x ? a : b.x, m.y ? a : b.y}; }

}  // namespace kid
