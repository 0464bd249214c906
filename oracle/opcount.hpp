// ORACLE — test/measurement infrastructure only.
// Operation-counting scalar for the CPU restatement ("instrument the C++ oracle with an op-counting scalar
// type once per style to get {add/mul, div, sqrt, exp/log/pow, erf/gamma} per rhs").  Cnt wraps a double; every
// arithmetic operator and math call increments a global counter.  Include BEFORE kid_physics.hpp; run single-threaded.
#pragma once
#include <cmath>
#include <limits>

namespace kidm {

struct OpCounters {
    double add = 0, mul = 0, div = 0, sqrt = 0, transc = 0, special = 0, cmp = 0;  // cmp: comparisons, min/max, abs
};
inline OpCounters& counters() { static OpCounters c; return c; }

struct Cnt {
    double v;
    Cnt() = default;
    Cnt(double x) : v(x) {}
    explicit operator double() const { return v; }
    explicit operator float() const { return float(v); }
    explicit operator int() const { return int(v); }
};
inline Cnt operator+(Cnt a, Cnt b) { counters().add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { counters().add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { counters().mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { counters().div++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt operator+(Cnt a) { return a; }
inline Cnt& operator+=(Cnt& a, Cnt b) { counters().add++; a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { counters().add++; a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { counters().mul++; a.v *= b.v; return a; }
inline Cnt& operator/=(Cnt& a, Cnt b) { counters().div++; a.v /= b.v; return a; }
#define KIDM_MIXED(op)                                                   \
    inline Cnt operator op(Cnt a, double b) { return a op Cnt(b); }      \
    inline Cnt operator op(double a, Cnt b) { return Cnt(a) op b; }      \
    inline Cnt operator op(Cnt a, int b) { return a op Cnt(double(b)); } \
    inline Cnt operator op(int a, Cnt b) { return Cnt(double(a)) op b; }
KIDM_MIXED(+) KIDM_MIXED(-) KIDM_MIXED(*) KIDM_MIXED(/)
#undef KIDM_MIXED
#define KIDM_CMP(op)                                                                     \
    inline bool operator op(Cnt a, Cnt b) { counters().cmp++; return a.v op b.v; }       \
    inline bool operator op(Cnt a, double b) { counters().cmp++; return a.v op b; }      \
    inline bool operator op(double a, Cnt b) { counters().cmp++; return a op b.v; }      \
    inline bool operator op(Cnt a, int b) { counters().cmp++; return a.v op double(b); } \
    inline bool operator op(int a, Cnt b) { counters().cmp++; return double(a) op b.v; }
KIDM_CMP(<) KIDM_CMP(>) KIDM_CMP(<=) KIDM_CMP(>=) KIDM_CMP(==) KIDM_CMP(!=)
#undef KIDM_CMP

inline Cnt pow(Cnt a, Cnt b) { counters().transc++; return Cnt(std::pow(a.v, b.v)); }
inline Cnt pow(Cnt a, double b) { return pow(a, Cnt(b)); }
inline Cnt pow(double a, Cnt b) { return pow(Cnt(a), b); }
inline Cnt pow(Cnt a, int b) { return pow(a, Cnt(double(b))); }
inline Cnt exp(Cnt a) { counters().transc++; return Cnt(std::exp(a.v)); }
inline Cnt log(Cnt a) { counters().transc++; return Cnt(std::log(a.v)); }
inline Cnt cbrt(Cnt a) { counters().transc++; return Cnt(std::cbrt(a.v)); }
inline Cnt sin(Cnt a) { counters().transc++; return Cnt(std::sin(a.v)); }
inline Cnt cos(Cnt a) { counters().transc++; return Cnt(std::cos(a.v)); }
inline Cnt sqrt(Cnt a) { counters().sqrt++; return Cnt(std::sqrt(a.v)); }
inline Cnt erf(Cnt a) { counters().special++; return Cnt(std::erf(a.v)); }
inline Cnt tgamma(Cnt a) { counters().special++; return Cnt(std::tgamma(a.v)); }
inline Cnt fabs(Cnt a) { counters().cmp++; return Cnt(std::fabs(a.v)); }
inline Cnt max(Cnt a, Cnt b) { counters().cmp++; return a.v < b.v ? b : a; }
inline Cnt min(Cnt a, Cnt b) { counters().cmp++; return b.v < a.v ? b : a; }
inline bool isnan(Cnt a) { return std::isnan(a.v); }

}  // namespace kidm

namespace std {
template <>
struct numeric_limits<kidm::Cnt> : numeric_limits<double> {
    static kidm::Cnt epsilon() { return kidm::Cnt(numeric_limits<double>::epsilon()); }
    static kidm::Cnt max() { return kidm::Cnt(numeric_limits<double>::max()); }
    static kidm::Cnt min() { return kidm::Cnt(numeric_limits<double>::min()); }
    static kidm::Cnt infinity() { return kidm::Cnt(numeric_limits<double>::infinity()); }
    static kidm::Cnt quiet_NaN() { return kidm::Cnt(numeric_limits<double>::quiet_NaN()); }
};
}  // namespace std
