// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path (see oracle/README.md).
//
// Minimal Jacobian-taping reverse-mode AD type, restating what the reference obtains from
// CoDiPack's `codi::RealReverse` + global tape (reference src/adjoint/DASolver/DASolver.H:789-794;
// record: DASolver.C:1411-1441; evaluate: DASolver.C:1396; clearAdjoints: DASolver.C:1404).
// CoDiPack itself is a third-party dependency that is absent from /root/reference and not pinned
// there; the published algorithm restated here is a Jacobian tape: every
// active statement stores its partial derivatives and the identifiers of its arguments, and
// `evaluate()` sweeps the statements backwards accumulating adjoints.  Non-smooth intrinsics
// (max/min/fabs) take the derivative of the active branch, sqrt has zero derivative at zero.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc
{

struct Tape
{
    std::vector<int32_t> ia, ib;
    std::vector<double> pa, pb;
    inline int push0()
    {
        ia.push_back(0); ib.push_back(0); pa.push_back(0.0); pb.push_back(0.0);
        return (int)ia.size();
    }
    inline int push1(int a, double p)
    {
        ia.push_back(a); ib.push_back(0); pa.push_back(p); pb.push_back(0.0);
        return (int)ia.size();
    }
    inline int push2(int a, double p, int b, double q)
    {
        ia.push_back(a); ib.push_back(b); pa.push_back(p); pb.push_back(q);
        return (int)ia.size();
    }
    void reset()
    {
        ia.clear(); ib.clear(); pa.clear(); pb.clear();
    }
    size_t size() const { return ia.size(); }
    // adj has size()+1 entries; entry 0 collects the (discarded) adjoint of passive arguments
    void evaluate(std::vector<double>& adj) const
    {
        for (long k = (long)ia.size() - 1; k >= 0; --k)
        {
            const double a = adj[k + 1];
            if (a != 0.0)
            {
                adj[ia[k]] += pa[k] * a;
                adj[ib[k]] += pb[k] * a;
            }
        }
    }
};

inline Tape& tape()
{
    static Tape t;
    return t;
}

struct AReal
{
    double v;
    int id;
    AReal() : v(0.0), id(0) {}
    AReal(double x) : v(x), id(0) {}
    AReal(double x, int i) : v(x), id(i) {}
    void registerInput() { id = tape().push0(); }
};

inline double val(const double& x) { return x; }
inline double val(const AReal& x) { return x.v; }

inline AReal un(const AReal& a, double v, double p) { return AReal(v, a.id ? tape().push1(a.id, p) : 0); }
inline AReal bin(const AReal& a, const AReal& b, double v, double p, double q)
{
    if (a.id || b.id) return AReal(v, tape().push2(a.id, p, b.id, q));
    return AReal(v, 0);
}

inline AReal operator+(const AReal& a, const AReal& b) { return bin(a, b, a.v + b.v, 1.0, 1.0); }
inline AReal operator-(const AReal& a, const AReal& b) { return bin(a, b, a.v - b.v, 1.0, -1.0); }
inline AReal operator*(const AReal& a, const AReal& b) { return bin(a, b, a.v * b.v, b.v, a.v); }
inline AReal operator/(const AReal& a, const AReal& b) { return bin(a, b, a.v / b.v, 1.0 / b.v, -a.v / (b.v * b.v)); }
inline AReal operator+(const AReal& a, double b) { return un(a, a.v + b, 1.0); }
inline AReal operator-(const AReal& a, double b) { return un(a, a.v - b, 1.0); }
inline AReal operator*(const AReal& a, double b) { return un(a, a.v * b, b); }
inline AReal operator/(const AReal& a, double b) { return un(a, a.v / b, 1.0 / b); }
inline AReal operator+(double a, const AReal& b) { return un(b, a + b.v, 1.0); }
inline AReal operator-(double a, const AReal& b) { return un(b, a - b.v, -1.0); }
inline AReal operator*(double a, const AReal& b) { return un(b, a * b.v, a); }
inline AReal operator/(double a, const AReal& b) { return un(b, a / b.v, -a / (b.v * b.v)); }
inline AReal operator-(const AReal& a) { return un(a, -a.v, -1.0); }
inline AReal& operator+=(AReal& a, const AReal& b) { a = a + b; return a; }
inline AReal& operator-=(AReal& a, const AReal& b) { a = a - b; return a; }
inline AReal& operator*=(AReal& a, const AReal& b) { a = a * b; return a; }
inline AReal& operator/=(AReal& a, const AReal& b) { a = a / b; return a; }
inline AReal& operator+=(AReal& a, double b) { a = a + b; return a; }
inline AReal& operator-=(AReal& a, double b) { a = a - b; return a; }
inline AReal& operator*=(AReal& a, double b) { a = a * b; return a; }
inline AReal& operator/=(AReal& a, double b) { a = a / b; return a; }

inline AReal sqrt(const AReal& a)
{
    const double r = std::sqrt(a.v);
    return un(a, r, r != 0.0 ? 0.5 / r : 0.0);
}
inline AReal fabs(const AReal& a) { return un(a, std::fabs(a.v), a.v < 0.0 ? -1.0 : 1.0); }
inline AReal pow(const AReal& a, double e)
{
    const double r = std::pow(a.v, e);
    return un(a, r, a.v != 0.0 ? e * r / a.v : 0.0);
}
inline AReal exp(const AReal& a)
{
    const double r = std::exp(a.v);
    return un(a, r, r);
}
inline AReal max(const AReal& a, const AReal& b) { return a.v > b.v ? a : b; }
inline AReal min(const AReal& a, const AReal& b) { return a.v < b.v ? a : b; }
inline AReal max(const AReal& a, double b) { return a.v > b ? a : AReal(b); }
inline AReal min(const AReal& a, double b) { return a.v < b ? a : AReal(b); }

inline double sqrt(double a) { return std::sqrt(a); }
inline double exp(double a) { return std::exp(a); }
inline double fabs(double a) { return std::fabs(a); }
inline double pow(double a, double e) { return std::pow(a, e); }
inline double max(double a, double b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }

// ---- forward-mode dual number (tangent of ONE direction): the independent check of the reverse tape.  psi^T (J v) from the
// dual instantiation of the residual must equal v^T (J^T psi) from the tape to rounding (1e-12), where central finite differences
// only give 1e-7 (tests/test_oracle.py).  Same conventions as the tape for the non-smooth intrinsics (active branch; sqrt'(0) = 0).
struct Dual
{
    double v, d;
    Dual() : v(0.0), d(0.0) {}
    Dual(double x) : v(x), d(0.0) {}
    Dual(double x, double t) : v(x), d(t) {}
};
inline double val(const Dual& x) { return x.v; }
inline Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator*(const Dual& a, const Dual& b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
inline Dual operator/(const Dual& a, const Dual& b) { return Dual(a.v / b.v, a.d / b.v - a.v * b.d / (b.v * b.v)); }
inline Dual operator+(const Dual& a, double b) { return Dual(a.v + b, a.d); }
inline Dual operator-(const Dual& a, double b) { return Dual(a.v - b, a.d); }
inline Dual operator*(const Dual& a, double b) { return Dual(a.v * b, a.d * b); }
inline Dual operator/(const Dual& a, double b) { return Dual(a.v / b, a.d / b); }
inline Dual operator+(double a, const Dual& b) { return Dual(a + b.v, b.d); }
inline Dual operator-(double a, const Dual& b) { return Dual(a - b.v, -b.d); }
inline Dual operator*(double a, const Dual& b) { return Dual(a * b.v, a * b.d); }
inline Dual operator/(double a, const Dual& b) { return Dual(a / b.v, -a * b.d / (b.v * b.v)); }
inline Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
inline Dual& operator+=(Dual& a, const Dual& b) { a = a + b; return a; }
inline Dual& operator-=(Dual& a, const Dual& b) { a = a - b; return a; }
inline Dual& operator*=(Dual& a, const Dual& b) { a = a * b; return a; }
inline Dual& operator/=(Dual& a, const Dual& b) { a = a / b; return a; }
inline Dual& operator+=(Dual& a, double b) { a = a + b; return a; }
inline Dual& operator-=(Dual& a, double b) { a = a - b; return a; }
inline Dual& operator*=(Dual& a, double b) { a = a * b; return a; }
inline Dual& operator/=(Dual& a, double b) { a = a / b; return a; }
inline Dual sqrt(const Dual& a)
{
    const double r = std::sqrt(a.v);
    return Dual(r, r != 0.0 ? 0.5 / r * a.d : 0.0);
}
inline Dual fabs(const Dual& a) { return Dual(std::fabs(a.v), a.v < 0.0 ? -a.d : a.d); }
inline Dual pow(const Dual& a, double e)
{
    const double r = std::pow(a.v, e);
    return Dual(r, a.v != 0.0 ? e * r / a.v * a.d : 0.0);
}
inline Dual exp(const Dual& a)
{
    const double r = std::exp(a.v);
    return Dual(r, r * a.d);
}
inline Dual max(const Dual& a, const Dual& b) { return a.v > b.v ? a : b; }
inline Dual min(const Dual& a, const Dual& b) { return a.v < b.v ? a : b; }
inline Dual max(const Dual& a, double b) { return a.v > b ? a : Dual(b); }
inline Dual min(const Dual& a, double b) { return a.v < b ? a : Dual(b); }

} // namespace orc
