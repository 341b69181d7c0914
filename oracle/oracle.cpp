// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see below).
//
// CPU restatement of the reference's adjoint hot path for DASimpleFoam (+ Spalart-Allmaras):
//   R(W)            reference src/adjoint/DAResidual/DAResidualSimpleFoam.C:106-237,
//                   src/adjoint/DAModel/DATurbulenceModel/DASpalartAllmaras.C:124-178,215-233,407-488,
//                   src/adjoint/DAModel/DATurbulenceModel/DATurbulenceModel.C:360-398
//   dR/dW^T psi     reference src/adjoint/DASolver/DASolver.C:1364-1441 (tape record + evaluate)
//   state layout    reference src/adjoint/DAIndex/DAIndex.C:188-257 ("state" ordering)
//   scalings        reference src/include/DAMacroFunctions.H:28-51, DASolver.C:2356-2455
//   force function  reference src/adjoint/DAFunction/DAFunctionForce.C:79-153
//
// The arithmetic of R(W) is not in the reference repository: it is OpenFOAM-v1812 operator algebra
// (fvm::div, fvm::laplacian, fvc::grad, fvMatrix::relax/A/H/flux/operator&, boundary-condition
// coefficients), a third-party dependency that is absent from /root/reference and not pinned there
// (SURVEY.md section 8c).  This file restates those operators one by one, in the same order and
// with the same data structures (diag/lower/upper/source/internalCoeffs/boundaryCoeffs), and
// differentiates them with the tape of tape.hpp exactly as the reference does with CoDiPack.
//
// PARITY UNPINNED: the reference's golden numbers for this path (tests/runUnitTests_DATurbModel.py:
// 96-130, tests/refs/*.txt) need case directories that are downloaded at test time and are not in
// /root/reference; neither OpenFOAM nor CoDiPack nor PETSc can be built here.  The oracle is
// therefore checked only for self-consistency (tape vs finite differences, dot-product identity).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may load
// this library.
#include "tape.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace orc
{

template <class T>
struct V3
{
    T x, y, z;
    V3() : x(0.0), y(0.0), z(0.0) {}
    V3(T a, T b, T c) : x(a), y(b), z(c) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> V3<T> operator+(const V3<T>& a, const V3<T>& b) { return V3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> V3<T> operator-(const V3<T>& a, const V3<T>& b) { return V3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> V3<T> operator*(const T& s, const V3<T>& a) { return V3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> V3<T> operator*(const V3<T>& a, const T& s) { return V3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> V3<T> cross(const V3<T>& a, const V3<T>& b)
{
    return V3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T> T mag(const V3<T>& a) { return sqrt(dot(a, a)); }

// boundary-condition kinds (shared numeric codes with include/dab200.h)
enum BCKind
{
    BC_FIXED_VALUE = 0,
    BC_ZERO_GRADIENT = 1,
    BC_INLET_OUTLET = 2,
    BC_OUTLET_INLET = 3,
    BC_SYMMETRY = 4,
    BC_CALCULATED = 5,
    BC_NUT_LOW_RE = 6,
    BC_NUT_SPALDING = 7
};
enum
{
    F_U = 0,
    F_P = 1,
    F_NUTILDA = 2,
    F_NUT = 3,
    N_FIELDS = 4
};
enum DivScheme
{
    DIV_UPWIND = 0,
    DIV_LINEAR_UPWIND = 1,
    DIV_LINEAR = 2,
    DIV_LINEAR_UPWIND_V = 3,
    DIV_LIMITED_LINEAR = 4 // limitedLinear k (NVD/TVD limited scheme, weights only)
};

struct Topo
{
    int nP, nF, nIF, nBF, nC, nPatch;
    std::vector<int> fOff, fLab, own, nei;
    std::vector<int> pStart, pSize, pGeom;
    std::vector<int> bPatch;           // patch of boundary face b (b = f - nIF)
    std::vector<int> cOff, cFace;      // cell -> faces CSR
};

struct Params
{
    double nu;
    double alphaU;    // momentum relaxation factor (fvSolution relaxationFactors.equations.U)
    int turb;         // 0: dummyTurbulenceModel (laminar, no nuTilda state), 1: SpalartAllmaras, 2: SpalartAllmarasFv3
    int divU, divNut; // DivScheme for div(phi,U), div(phi,nuTilda)
    double sU, sP, sNut, sPhi; // normalizeStates
    int phiNorm = 1;           // "phi" listed in normalizeStates (DASolver.C:2431-2452: the phi rows are scaled only then)
    int nrU, nrP, nrNut, nrPhi; // 1 = residual name listed in normalizeResiduals
    int constrainHbyA;
};

struct BCSpec
{
    std::vector<int> kind;     // [field][patch]
    std::vector<double> value; // [field][patch][3]
};

template <class T>
struct Geom
{
    std::vector<V3<T>> Sf, Cf, C, corr; // corr: nonOrthCorrectionVectors (internal faces; zero on boundary)
    std::vector<T> magSf, V, w, delta;  // w: linear weights; delta: nonOrthDeltaCoeffs (all faces)
};

// ---- geometry: OpenFOAM primitiveMesh / surfaceInterpolation definitions ---------------------
template <class T>
void computeGeometry(const Topo& t, const std::vector<V3<T>>& pts, Geom<T>& g)
{
    g.Sf.assign(t.nF, V3<T>());
    g.Cf.assign(t.nF, V3<T>());
    g.magSf.assign(t.nF, T(0.0));
    for (int f = 0; f < t.nF; f++)
    {
        const int n = t.fOff[f + 1] - t.fOff[f];
        const int* l = &t.fLab[t.fOff[f]];
        if (n == 3)
        {
            g.Cf[f] = T(1.0 / 3.0) * (pts[l[0]] + pts[l[1]] + pts[l[2]]);
            g.Sf[f] = T(0.5) * cross(pts[l[1]] - pts[l[0]], pts[l[2]] - pts[l[0]]);
        }
        else
        {
            V3<T> est;
            for (int i = 0; i < n; i++) est = est + pts[l[i]];
            est = T(1.0 / n) * est;
            V3<T> sumN, sumAc;
            T sumA(0.0);
            for (int i = 0; i < n; i++)
            {
                const V3<T>& p0 = pts[l[i]];
                const V3<T>& p1 = pts[l[(i + 1) % n]];
                V3<T> c = p0 + p1 + est;
                V3<T> nn = cross(p1 - p0, est - p0);
                T a = mag(nn);
                sumN = sumN + nn;
                sumA = sumA + a;
                sumAc = sumAc + a * c;
            }
            g.Cf[f] = (T(1.0 / 3.0) / sumA) * sumAc;
            g.Sf[f] = T(0.5) * sumN;
        }
        g.magSf[f] = mag(g.Sf[f]);
    }
    // cell centres and volumes
    g.C.assign(t.nC, V3<T>());
    g.V.assign(t.nC, T(0.0));
    std::vector<V3<T>> cEst(t.nC);
    std::vector<int> nCF(t.nC, 0);
    for (int f = 0; f < t.nF; f++)
    {
        cEst[t.own[f]] = cEst[t.own[f]] + g.Cf[f];
        nCF[t.own[f]]++;
        if (f < t.nIF)
        {
            cEst[t.nei[f]] = cEst[t.nei[f]] + g.Cf[f];
            nCF[t.nei[f]]++;
        }
    }
    for (int c = 0; c < t.nC; c++) cEst[c] = T(1.0 / nCF[c]) * cEst[c];
    for (int f = 0; f < t.nF; f++)
    {
        {
            const int c = t.own[f];
            T pyr3 = dot(g.Sf[f], g.Cf[f] - cEst[c]);
            V3<T> pc = T(0.75) * g.Cf[f] + T(0.25) * cEst[c];
            g.C[c] = g.C[c] + pyr3 * pc;
            g.V[c] = g.V[c] + pyr3;
        }
        if (f < t.nIF)
        {
            const int c = t.nei[f];
            T pyr3 = dot(g.Sf[f], cEst[c] - g.Cf[f]);
            V3<T> pc = T(0.75) * g.Cf[f] + T(0.25) * cEst[c];
            g.C[c] = g.C[c] + pyr3 * pc;
            g.V[c] = g.V[c] + pyr3;
        }
    }
    for (int c = 0; c < t.nC; c++)
    {
        g.C[c] = (T(1.0) / g.V[c]) * g.C[c];
        g.V[c] = g.V[c] * T(1.0 / 3.0);
    }
    // weights, nonOrthDeltaCoeffs, nonOrthCorrectionVectors
    g.w.assign(t.nF, T(1.0));
    g.delta.assign(t.nF, T(0.0));
    g.corr.assign(t.nF, V3<T>());
    for (int f = 0; f < t.nF; f++)
    {
        V3<T> nHat = (T(1.0) / g.magSf[f]) * g.Sf[f];
        if (f < t.nIF)
        {
            T dOwn = fabs(dot(g.Sf[f], g.Cf[f] - g.C[t.own[f]]));
            T dNei = fabs(dot(g.Sf[f], g.C[t.nei[f]] - g.Cf[f]));
            g.w[f] = dNei / (dOwn + dNei);
            V3<T> d = g.C[t.nei[f]] - g.C[t.own[f]];
            g.delta[f] = T(1.0) / max(dot(nHat, d), T(0.05) * mag(d));
            g.corr[f] = nHat - g.delta[f] * d;
        }
        else
        {
            // ESI fvPatch::delta(): patch-normal delta for non-coupled patches
            T dn = dot(nHat, g.Cf[f] - g.C[t.own[f]]);
            V3<T> d = dn * nHat;
            g.delta[f] = T(1.0) / max(dot(nHat, d), T(0.05) * mag(d));
        }
    }
}

// ---- boundary-condition coefficients ----------------------------------------------------------
// For every boundary face and field component: value, snGrad and the four fvMatrix coefficients
// (valueInternalCoeffs, valueBoundaryCoeffs, gradientInternalCoeffs, gradientBoundaryCoeffs).
template <class T>
struct BF
{
    int nc, nBF;
    std::vector<T> val, sng, vic, vbc, gic, gbc;
    void init(int nc_, int nBF_)
    {
        nc = nc_;
        nBF = nBF_;
        size_t n = (size_t)nc * nBF;
        val.assign(n, T(0.0)); sng.assign(n, T(0.0)); vic.assign(n, T(0.0));
        vbc.assign(n, T(0.0)); gic.assign(n, T(0.0)); gbc.assign(n, T(0.0));
    }
    size_t at(int k, int b) const { return (size_t)k * nBF + b; }
};

// mixed form: x_b = f*ref + (1-f)*x_P (refGrad = 0 for all kinds used here)
template <class T>
void mixedCoeffs(BF<T>& bf, int k, int b, double fr, const T& ref, const T& xP, const T& delta)
{
    size_t i = bf.at(k, b);
    bf.val[i] = fr * ref + (1.0 - fr) * xP;
    bf.sng[i] = fr * (ref - xP) * delta;
    bf.vic[i] = T(1.0 - fr);
    bf.vbc[i] = fr * ref;
    bf.gic[i] = -fr * delta;
    bf.gbc[i] = fr * ref * delta;
}

// bcv: the boundary reference values as scalars of type T ([field][patch][3]) so that they can be AD inputs
template <class T>
void evalBC(const Topo& t, const Geom<T>& g, const BCSpec& bc, const std::vector<T>& bcv, int field, int nc, const std::vector<T>& x /*nc*nC*/,
            const std::vector<T>& phi, BF<T>& bf)
{
    bf.init(nc, t.nBF);
    for (int b = 0; b < t.nBF; b++)
    {
        const int f = t.nIF + b, c = t.own[f], pa = t.bPatch[b];
        const int kind = bc.kind[field * t.nPatch + pa];
        const T* ref = &bcv[(field * t.nPatch + pa) * 3];
        const T& dl = g.delta[f];
        double fr = 0.0;
        switch (kind)
        {
        case BC_FIXED_VALUE:
        case BC_NUT_LOW_RE:
            fr = 1.0;
            break;
        case BC_ZERO_GRADIENT:
            fr = 0.0;
            break;
        case BC_INLET_OUTLET: // valueFraction = 1 - pos0(phi)
            fr = (val(phi[f]) >= 0.0) ? 0.0 : 1.0;
            break;
        case BC_OUTLET_INLET: // valueFraction = pos0(phi)
            fr = (val(phi[f]) >= 0.0) ? 1.0 : 0.0;
            break;
        default:
            break;
        }
        if (kind == BC_SYMMETRY && nc == 3)
        {
            // basicSymmetryFvPatchField<vector> + transformFvPatchField coefficients
            V3<T> n = (T(1.0) / g.magSf[f]) * g.Sf[f];
            V3<T> xP(x[0 * t.nC + c], x[1 * t.nC + c], x[2 * t.nC + c]);
            T xn = dot(n, xP);
            for (int k = 0; k < 3; k++)
            {
                size_t i = bf.at(k, b);
                T an = fabs(n[k]);
                bf.val[i] = xP[k] - xn * n[k];
                bf.sng[i] = -(xn * n[k]) * dl;
                bf.vic[i] = 1.0 - an;
                bf.vbc[i] = bf.val[i] - bf.vic[i] * xP[k];
                bf.gic[i] = -(dl * an);
                bf.gbc[i] = bf.sng[i] - bf.gic[i] * xP[k];
            }
        }
        else if (kind == BC_SYMMETRY || kind == BC_CALCULATED || kind == BC_NUT_SPALDING)
        {
            // scalar symmetry == zero gradient; `calculated` is overwritten by the caller (nut)
            for (int k = 0; k < nc; k++) mixedCoeffs(bf, k, b, 0.0, T(0.0), x[(size_t)k * t.nC + c], dl);
        }
        else
        {
            for (int k = 0; k < nc; k++) mixedCoeffs(bf, k, b, fr, ref[k], x[(size_t)k * t.nC + c], dl);
        }
    }
}

// ---- fvMatrix -----------------------------------------------------------------------------------
template <class T>
struct Mat
{
    int nc, nC, nIF, nBF;
    std::vector<T> diag, lower, upper;
    std::vector<T> src;    // [k][cell]
    std::vector<T> ic, bc; // internalCoeffs / boundaryCoeffs [k][bface]
    std::vector<T> ffc;    // faceFluxCorrection (all faces), nc==1 only
    void init(const Topo& t, int nc_)
    {
        nc = nc_; nC = t.nC; nIF = t.nIF; nBF = t.nBF;
        diag.assign(nC, T(0.0)); lower.assign(nIF, T(0.0)); upper.assign(nIF, T(0.0));
        src.assign((size_t)nc * nC, T(0.0));
        ic.assign((size_t)nc * nBF, T(0.0)); bc.assign((size_t)nc * nBF, T(0.0));
        ffc.assign(t.nF, T(0.0));
    }
};

// Gauss linear gradient of an nc-component cell field; out[(k*3+i)*nC + c] = d_i x_k
template <class T>
void fvcGrad(const Topo& t, const Geom<T>& g, int nc, const std::vector<T>& x, const BF<T>& bf, std::vector<T>& out)
{
    out.assign((size_t)nc * 3 * t.nC, T(0.0));
    for (int k = 0; k < nc; k++)
    {
        for (int f = 0; f < t.nF; f++)
        {
            T xf;
            const int o = t.own[f];
            if (f < t.nIF)
                xf = g.w[f] * x[(size_t)k * t.nC + o] + (1.0 - g.w[f]) * x[(size_t)k * t.nC + t.nei[f]];
            else
                xf = bf.val[bf.at(k, f - t.nIF)];
            for (int i = 0; i < 3; i++)
            {
                T s = g.Sf[f][i] * xf;
                out[((size_t)k * 3 + i) * t.nC + o] += s;
                if (f < t.nIF) out[((size_t)k * 3 + i) * t.nC + t.nei[f]] -= s;
            }
        }
        for (int i = 0; i < 3; i++)
            for (int c = 0; c < t.nC; c++) out[((size_t)k * 3 + i) * t.nC + c] /= g.V[c];
    }
}

// fvm::div(phi, x) with Gauss <scheme>; gaussConvectionScheme::fvmDiv (+ linearUpwind::correction)
template <class T>
void fvmDiv(Mat<T>& m, const Topo& t, const Geom<T>& g, const std::vector<T>& phi, int scheme, const BF<T>& bf,
            const std::vector<T>& gradX, bool bounded, const std::vector<T>* xc = nullptr, double limitedLinearK = 1.0)
{
    for (int f = 0; f < t.nIF; f++)
    {
        T w;
        if (scheme == DIV_LINEAR)
            w = g.w[f];
        else if (scheme == DIV_LIMITED_LINEAR)
        {
            // limitedScheme::weights with LimitedLinearLimiter (OpenFOAM limitedLinear.H, NVDTVD.H::r): scalar fields only
            const std::vector<T>& x = *xc;
            const int o = t.own[f], n = t.nei[f];
            const T gradf = x[n] - x[o];
            const int u = (val(phi[f]) > 0.0) ? o : n;
            V3<T> d = g.C[n] - g.C[o];
            const T gradcf = d.x * gradX[(size_t)0 * t.nC + u] + d.y * gradX[(size_t)1 * t.nC + u] + d.z * gradX[(size_t)2 * t.nC + u];
            T r;
            auto sgn = [](double v) { return v >= 0.0 ? 1.0 : -1.0; };
            if (std::fabs(val(gradcf)) >= 1000.0 * std::fabs(val(gradf)))
                r = T(2.0 * 1000.0 * sgn(val(gradcf)) * sgn(val(gradf)) - 1.0);
            else
                r = 2.0 * (gradcf / gradf) - 1.0;
            T lim = (2.0 / std::max(limitedLinearK, 1e-15)) * r;
            if (val(lim) > 1.0) lim = T(1.0);
            if (val(lim) < 0.0) lim = T(0.0);
            w = lim * g.w[f] + (1.0 - lim) * (val(phi[f]) >= 0.0 ? 1.0 : 0.0);
        }
        else
            w = T(val(phi[f]) >= 0.0 ? 1.0 : 0.0);
        T lo = -(w * phi[f]);
        T up = lo + phi[f];
        m.lower[f] += lo;
        m.upper[f] += up;
        m.diag[t.own[f]] -= lo;
        m.diag[t.nei[f]] -= up;
        if (scheme == DIV_LINEAR_UPWIND || scheme == DIV_LINEAR_UPWIND_V)
        {
            const int u = (val(phi[f]) > 0.0) ? t.own[f] : t.nei[f];
            V3<T> d = g.Cf[f] - g.C[u];
            std::vector<T> corr(m.nc);
            for (int k = 0; k < m.nc; k++)
                corr[k] = d.x * gradX[((size_t)k * 3 + 0) * t.nC + u] + d.y * gradX[((size_t)k * 3 + 1) * t.nC + u]
                    + d.z * gradX[((size_t)k * 3 + 2) * t.nC + u];
            if (scheme == DIV_LINEAR_UPWIND_V)
            {
                // linearUpwindV<vector>::correction (OpenFOAM-v1812): limit the correction along the linear increment
                const std::vector<T>& x = *xc;
                const int o = t.own[f], n = t.nei[f];
                T sfCorrs(0.0), maxCorrs(0.0);
                for (int k = 0; k < 3; k++)
                {
                    T mc = (val(phi[f]) > 0.0) ? (1.0 - g.w[f]) * (x[(size_t)k * t.nC + n] - x[(size_t)k * t.nC + o])
                                               : g.w[f] * (x[(size_t)k * t.nC + o] - x[(size_t)k * t.nC + n]);
                    sfCorrs += corr[k] * corr[k];
                    maxCorrs += corr[k] * mc;
                }
                if (val(sfCorrs) > 0.0)
                {
                    if (val(maxCorrs) < 0.0)
                        for (int k = 0; k < 3; k++) corr[k] = T(0.0);
                    else if (val(sfCorrs) > val(maxCorrs))
                    {
                        T ratio = maxCorrs / (sfCorrs + 1e-300);
                        for (int k = 0; k < 3; k++) corr[k] = corr[k] * ratio;
                    }
                }
            }
            for (int k = 0; k < m.nc; k++)
            {
                T fl = phi[f] * corr[k];
                m.src[(size_t)k * t.nC + t.own[f]] -= fl;
                m.src[(size_t)k * t.nC + t.nei[f]] += fl;
            }
        }
    }
    for (int b = 0; b < t.nBF; b++)
        for (int k = 0; k < m.nc; k++)
        {
            m.ic[(size_t)k * t.nBF + b] += phi[t.nIF + b] * bf.vic[bf.at(k, b)];
            m.bc[(size_t)k * t.nBF + b] -= phi[t.nIF + b] * bf.vbc[bf.at(k, b)];
        }
    if (bounded)
    {
        // boundedConvectionScheme: - fvm::Sp(fvc::surfaceIntegrate(phi), x)
        for (int f = 0; f < t.nF; f++)
        {
            m.diag[t.own[f]] -= phi[f];
            if (f < t.nIF) m.diag[t.nei[f]] += phi[f];
        }
    }
}

// sign * fvm::laplacian(gamma, x), Gauss linear corrected (gaussLaplacianScheme + correctedSnGrad)
template <class T>
void fvmLaplacian(Mat<T>& m, const Topo& t, const Geom<T>& g, double sign, const std::vector<T>& gammaC,
                  const std::vector<T>& gammaB, const BF<T>& bf, const std::vector<T>& gradX, bool storeFlux)
{
    for (int f = 0; f < t.nIF; f++)
    {
        const int o = t.own[f], n = t.nei[f];
        T gf = (g.w[f] * gammaC[o] + (1.0 - g.w[f]) * gammaC[n]) * g.magSf[f];
        T up = sign * (g.delta[f] * gf);
        m.upper[f] += up;
        m.lower[f] += up;
        m.diag[o] -= up;
        m.diag[n] -= up;
        for (int k = 0; k < m.nc; k++)
        {
            T cg(0.0);
            for (int i = 0; i < 3; i++)
                cg += g.corr[f][i]
                    * (g.w[f] * gradX[((size_t)k * 3 + i) * t.nC + o] + (1.0 - g.w[f]) * gradX[((size_t)k * 3 + i) * t.nC + n]);
            T cf = sign * (gf * cg);
            m.src[(size_t)k * t.nC + o] -= cf;
            m.src[(size_t)k * t.nC + n] += cf;
            if (storeFlux) m.ffc[f] += cf;
        }
    }
    for (int b = 0; b < t.nBF; b++)
        for (int k = 0; k < m.nc; k++)
        {
            T pg = gammaB[b] * g.magSf[t.nIF + b];
            m.ic[(size_t)k * t.nBF + b] += sign * (pg * bf.gic[bf.at(k, b)]);
            m.bc[(size_t)k * t.nBF + b] -= sign * (pg * bf.gbc[bf.at(k, b)]);
        }
}

// fvMatrix<Type>::relax(alpha), OpenFOAM-v1812
template <class T>
void relax(Mat<T>& m, const Topo& t, double alpha, const std::vector<T>& x)
{
    std::vector<T> D0(m.diag), sumOff(t.nC, T(0.0));
    for (int f = 0; f < t.nIF; f++)
    {
        sumOff[t.own[f]] += fabs(m.upper[f]);
        sumOff[t.nei[f]] += fabs(m.lower[f]);
    }
    for (int b = 0; b < t.nBF; b++)
    {
        const int c = t.own[t.nIF + b];
        T mx = fabs(m.ic[b]);
        for (int k = 1; k < m.nc; k++) mx = max(mx, fabs(m.ic[(size_t)k * t.nBF + b]));
        m.diag[c] += mx;
    }
    for (int c = 0; c < t.nC; c++)
    {
        m.diag[c] = max(fabs(m.diag[c]), sumOff[c]);
        m.diag[c] /= alpha;
    }
    for (int b = 0; b < t.nBF; b++)
    {
        const int c = t.own[t.nIF + b];
        T mn = m.ic[b];
        for (int k = 1; k < m.nc; k++) mn = min(mn, m.ic[(size_t)k * t.nBF + b]);
        m.diag[c] -= mn;
    }
    for (int k = 0; k < m.nc; k++)
        for (int c = 0; c < t.nC; c++) m.src[(size_t)k * t.nC + c] += (m.diag[c] - D0[c]) * x[(size_t)k * t.nC + c];
}

// (M & x): (A x - b)/V with boundary coefficients
template <class T>
void matResidual(const Mat<T>& m, const Topo& t, const Geom<T>& g, const std::vector<T>& x, std::vector<T>& out)
{
    out.assign((size_t)m.nc * t.nC, T(0.0));
    for (int k = 0; k < m.nc; k++)
    {
        T* o = &out[(size_t)k * t.nC];
        const T* xk = &x[(size_t)k * t.nC];
        for (int c = 0; c < t.nC; c++) o[c] = m.diag[c] * xk[c] - m.src[(size_t)k * t.nC + c];
        for (int f = 0; f < t.nIF; f++)
        {
            o[t.own[f]] += m.upper[f] * xk[t.nei[f]];
            o[t.nei[f]] += m.lower[f] * xk[t.own[f]];
        }
        for (int b = 0; b < t.nBF; b++)
        {
            const int c = t.own[t.nIF + b];
            o[c] += m.ic[(size_t)k * t.nBF + b] * xk[c] - m.bc[(size_t)k * t.nBF + b];
        }
        for (int c = 0; c < t.nC; c++) o[c] /= g.V[c];
    }
}

// A() and H() of fvMatrix<vector> (component-averaged boundary diagonal)
template <class T>
void matAH(const Mat<T>& m, const Topo& t, const Geom<T>& g, const std::vector<T>& x, std::vector<T>& A, std::vector<T>& H)
{
    A.assign(t.nC, T(0.0));
    H.assign((size_t)m.nc * t.nC, T(0.0));
    std::vector<T> av(t.nC, T(0.0));
    for (int b = 0; b < t.nBF; b++)
    {
        T s(0.0);
        for (int k = 0; k < m.nc; k++) s += m.ic[(size_t)k * t.nBF + b];
        av[t.own[t.nIF + b]] += s / double(m.nc);
    }
    for (int c = 0; c < t.nC; c++) A[c] = (m.diag[c] + av[c]) / g.V[c];
    for (int k = 0; k < m.nc; k++)
    {
        T* h = &H[(size_t)k * t.nC];
        const T* xk = &x[(size_t)k * t.nC];
        std::vector<T> bd(t.nC, T(0.0));
        for (int b = 0; b < t.nBF; b++) bd[t.own[t.nIF + b]] += m.ic[(size_t)k * t.nBF + b];
        for (int c = 0; c < t.nC; c++) h[c] = (av[c] - bd[c]) * xk[c] + m.src[(size_t)k * t.nC + c];
        for (int f = 0; f < t.nIF; f++)
        {
            h[t.own[f]] -= m.upper[f] * xk[t.nei[f]];
            h[t.nei[f]] -= m.lower[f] * xk[t.own[f]];
        }
        for (int b = 0; b < t.nBF; b++) h[t.own[t.nIF + b]] += m.bc[(size_t)k * t.nBF + b];
        for (int c = 0; c < t.nC; c++) h[c] /= g.V[c];
    }
}

// fvMatrix<scalar>::flux()
template <class T>
void matFlux(const Mat<T>& m, const Topo& t, const std::vector<T>& x, std::vector<T>& fl)
{
    fl.assign(t.nF, T(0.0));
    for (int f = 0; f < t.nIF; f++) fl[f] = m.upper[f] * x[t.nei[f]] - m.lower[f] * x[t.own[f]] + m.ffc[f];
    for (int b = 0; b < t.nBF; b++) fl[t.nIF + b] = m.ic[b] * x[t.own[t.nIF + b]] - m.bc[b];
}

// ---- SA closures (DASpalartAllmaras.C:41-178) ------------------------------------------------------
struct SAConst
{
    double sigmaNut = 0.66666, kappa = 0.41, Cb1 = 0.1355, Cb2 = 0.622, Cw2 = 0.3, Cw3 = 2.0, Cv1 = 7.1, Cs = 0.3;
    double Cw1() const { return Cb1 / (kappa * kappa) + (1.0 + Cb2) / sigmaNut; }
};

template <class T>
T fv1f(const T& chi)
{
    SAConst k;
    T chi3 = chi * chi * chi;
    return chi3 / (chi3 + k.Cv1 * k.Cv1 * k.Cv1);
}

// ---- the case ---------------------------------------------------------------------------------------
struct Case
{
    Topo t;
    BCSpec bc;
    Params par;
    std::vector<double> pts;   // 3*nP
    std::vector<double> yWall; // nC (frozen wall distance)
    Geom<double> gd;
    // recorded tape state
    std::vector<int> inId, outId;
    std::vector<double> adj;
    bool recorded = false;
    // fvSource: actuator disks (cylinderAnnulusSmooth), 15 numbers per disk: the 13 actuatorDiskPars, eps, rotLeft
    std::vector<double> disks;
    // DARhoSimpleFoam (compressible) extension
    struct Comp
    {
        int on = 0;
        int heIsE = 1;      // energy variable: 1 sensibleInternalEnergy (e), 0 sensibleEnthalpy (h)
        int sutherland = 0; // transport: 0 const (mu, Pr), 1 sutherland (As, Ts)
        int divE = 0, divEkp = 0, nrT = 1;
        int transonic = 0, divPhidP = 0, transonicPCOption = -1; // simple_.transonic(): fvm::div(phid, p) pressure equation
        double phidK = 1.0;                                      // k of 'limitedLinear k' for div(phid,p)
        int turbo = 0; // DATurboFoam: the sensibleEnthalpy energy equation carries the viscous-work and p(U - URel) terms
        double R = 287.0, Cp = 1005.0, mu = 1.8e-5, Pr = 0.7, Prt = 1.0, As = 1.4792e-6, Ts = 116.0, TRef = 298.15, sT = 1.0;
        std::vector<int> kindT;     // [nPatch]
        std::vector<double> valueT; // [nPatch]
    } comp;
    // MRF zone (reference src/adjoint/DAMisc/MRFDF/MRFZoneDF.C): one rotating cellZone
    struct Mrf
    {
        int on = 0;
        double omega[3] = {0, 0, 0}, origin[3] = {0, 0, 0}; // Omega = omega*axis [rad/s]
        std::vector<unsigned char> cell;     // [nC] in the zone
        std::vector<unsigned char> faceType; // [nF] MRFZoneDF::setMRFFaces: 1 internal-or-included, 2 excluded boundary face
    } mrf;
    int nDof() const { return ((par.turb ? 5 : 4) + (comp.on ? 1 : 0)) * t.nC + t.nF; }
};

// Omega x (x - origin)
template <class T>
V3<T> mrfVelocity(const Case& cs, const V3<T>& x)
{
    const double* w = cs.mrf.omega;
    const double* o = cs.mrf.origin;
    T r0 = x[0] - o[0], r1 = x[1] - o[1], r2 = x[2] - o[2];
    V3<T> v;
    v[0] = w[1] * r2 - w[2] * r1;
    v[1] = w[2] * r0 - w[0] * r2;
    v[2] = w[0] * r1 - w[1] * r0;
    return v;
}

// (Omega x (Cf - origin)) . Sf of one face (MRFZoneDF::makeRelativeRhoFlux, MRFZoneTemplatesDF.C:34-95)
template <class T>
T mrfFaceFlux(const Case& cs, const Geom<T>& g, int f)
{
    V3<T> v = mrfVelocity(cs, g.Cf[f]);
    return v[0] * g.Sf[f][0] + v[1] * g.Sf[f][1] + v[2] * g.Sf[f][2];
}

// MRFZoneDF::correctBoundaryVelocity (MRFZoneDF.C: included faces get U_b = Omega x (Cf - origin), forced like a fixed value)
template <class T>
void mrfWallVelocity(const Case& cs, const Geom<T>& g, const std::vector<T>& U, BF<T>& bU)
{
    const Topo& t = cs.t;
    for (int b = 0; b < t.nBF; b++)
    {
        const int f = t.nIF + b, c = t.own[f];
        if (cs.mrf.faceType[f] != 1) continue;
        V3<T> v = mrfVelocity(cs, g.Cf[f]);
        for (int k = 0; k < 3; k++) mixedCoeffs(bU, k, b, 1.0, v[k], U[(size_t)k * t.nC + c], g.delta[f]);
    }
}

template <class T>
struct Work
{
    // everything the force function needs after a residual evaluation
    BF<T> bU, bP, bNt, bNut;
    std::vector<T> gradU, nutC;
    std::vector<T> muEB; // compressible: rho_b*nuEff_b on the boundary faces
    std::vector<T> rhoB; // compressible: boundary density
    std::vector<T> TB;   // compressible: boundary temperature
    // capture of the relaxed nuTilda fvMatrix (calcPCMatWithFvMatrix, turbOnly): D(), upper(), lower()
    bool captureNut = false;
    double nutAlpha = 0.7;
    std::vector<T> nutD, nutUpper, nutLower;
};

// DAFvSourceActuatorDisk::calcFvSource, source = cylinderAnnulusSmooth (reference DAFvSourceActuatorDisk.C:205-407), adjustThrust 0;
// S[(j*nC + c)] = force per unit volume (depends on the cell centres, hence on the mesh points)
template <class T>
void actuatorSource(const Case& cs, const Geom<T>& g, std::vector<T>& S)
{
    const int nC = cs.t.nC;
    S.assign((size_t)3 * nC, T(0.0));
    const int nd = (int)cs.disks.size() / 15;
    for (int k = 0; k < nd; k++)
    {
        const double* a = &cs.disks[(size_t)15 * k];
        const double dm = std::sqrt(a[3] * a[3] + a[4] * a[4] + a[5] * a[5]);
        const double dn[3] = {a[3] / dm, a[4] / dm, a[5] / dm};
        const double rin = a[6], rout = a[7], scale = a[8], POD = a[9], expM = a[10], expN = a[11], eps = a[13];
        const bool rotLeft = a[14] != 0.0;
        const double epsR = eps / (rout - rin), rMin = epsR, rMax = 1.0 - epsR;
        const double fRMin = std::pow(rMin, expM) * std::pow(1.0 - rMin, expN), fRMax = std::pow(rMax, expM) * std::pow(1.0 - rMax, expN);
        for (int c = 0; c < nC; c++)
        {
            V3<T> v(g.C[c][0] - a[0], g.C[c][1] - a[1], g.C[c][2] - a[2]);
            V3<T> vA(v[0] * dn[0], v[1] * dn[1], v[2] * dn[2]);
            V3<T> vR = v - vA;
            V3<T> dnT;
            dnT.x = T(dn[0]); dnT.y = T(dn[1]); dnT.z = T(dn[2]);
            V3<T> vC = rotLeft ? cross(vR, dnT) : cross(dnT, vR);
            T rLen = mag(vR), cLen = mag(vC);
            T dA2 = dot(vA, vA);
            T rPrime = rLen / rout;
            const double rHub = rin / rout;
            T rStar = (rPrime - rHub) / (1.0 - rHub);
            T fR;
            if (val(rStar) < rMin) fR = fRMin * exp(-((rStar - rMin) * (rStar - rMin)) / epsR / epsR) * scale;
            else if (val(rStar) <= rMax) fR = pow(rStar, expM) * pow(1.0 - rStar, expN) * scale;
            else fR = fRMax * exp(-((rStar - rMax) * (rStar - rMax)) / epsR / epsR) * scale;
            T fAxial = fR * exp(-dA2 / eps / eps);
            T fCirc = fAxial * POD / 3.14159265358979323846 / (rPrime + 0.01 * eps / rout);
            for (int j = 0; j < 3; j++)
            {
                S[(size_t)j * nC + c] += fAxial * dn[j];
                if (val(cLen) > 0.0) S[(size_t)j * nC + c] += fCirc * vC[j] / cLen;
            }
        }
    }
}

template <class T>
void residualComp(const Case& cs, const Geom<T>& g, const std::vector<T>& W, int isPC, std::vector<T>& R, Work<T>* wk,
                  const std::vector<T>* bcvIn);

// R(W): DAResidualSimpleFoam::calcResiduals + DASpalartAllmaras::calcResiduals, preceded by
// DASolver::updateStateBoundaryConditions (BCs + correctNut).
template <class T>
void residual(const Case& cs, const Geom<T>& g, const std::vector<T>& W, int isPC, std::vector<T>& R, Work<T>* wk = nullptr,
              const std::vector<T>* bcvIn = nullptr)
{
    if (cs.comp.on)
    {
        residualComp<T>(cs, g, W, isPC, R, wk, bcvIn);
        return;
    }
    std::vector<T> bcvLocal;
    if (!bcvIn)
    {
        bcvLocal.resize(cs.bc.value.size());
        for (size_t i = 0; i < bcvLocal.size(); i++) bcvLocal[i] = T(cs.bc.value[i]);
    }
    const std::vector<T>& bcv = bcvIn ? *bcvIn : bcvLocal;
    const Topo& t = cs.t;
    const Params& par = cs.par;
    const int nC = t.nC, nF = t.nF, nIF = t.nIF, nBF = t.nBF;
    const bool turb = par.turb != 0;
    SAConst sa;
    // unpack the state vector (DAIndex state ordering)
    std::vector<T> U((size_t)3 * nC), p(nC), nt(nC, T(0.0)), phi(nF);
    for (int c = 0; c < nC; c++)
        for (int k = 0; k < 3; k++) U[(size_t)k * nC + c] = W[(size_t)3 * c + k];
    size_t off = (size_t)3 * nC;
    for (int c = 0; c < nC; c++) p[c] = W[off + c];
    off += nC;
    if (turb)
    {
        for (int c = 0; c < nC; c++) nt[c] = W[off + c];
        off += nC;
    }
    for (int f = 0; f < nF; f++) phi[f] = W[off + f];

    // --- boundary conditions and intermediate variables (DASolver::updateStateBoundaryConditions)
    BF<T> bU, bP, bNt, bNut;
    evalBC(t, g, cs.bc, bcv, F_U, 3, U, phi, bU);
    if (cs.mrf.on) mrfWallVelocity(cs, g, U, bU);
    evalBC(t, g, cs.bc, bcv, F_P, 1, p, phi, bP);
    std::vector<T> nut(nC, T(0.0));
    if (turb)
    {
        evalBC(t, g, cs.bc, bcv, F_NUTILDA, 1, nt, phi, bNt);
        // DASpalartAllmaras::correctNut: nut = nuTilda*fv1 (internal and boundary), then nut BCs
        for (int c = 0; c < nC; c++) nut[c] = nt[c] * fv1f(T(nt[c] / par.nu));
        evalBC(t, g, cs.bc, bcv, F_NUT, 1, nut, phi, bNut);
        for (int b = 0; b < nBF; b++)
        {
            const int kindN = cs.bc.kind[F_NUT * t.nPatch + t.bPatch[b]];
            if (kindN == BC_CALCULATED)
            {
                const T& nb = bNt.val[b];
                bNut.val[b] = nb * fv1f(T(nb / par.nu));
            }
            else if (kindN == BC_NUT_SPALDING)
            {
                // nutUSpaldingWallFunctionFvPatchScalarFieldDF::calcNut / calcUTau, differentiated through the
                // Newton iterations like the reference's CoDiPack build
                const int f = nIF + b, c = t.own[f];
                const double kappa = 0.41, E = 9.8, ROOTVSMALL = 1.0e-150;
                T d2(0.0);
                for (int k = 0; k < 3; k++)
                {
                    T dd = U[(size_t)k * nC + c] - bU.val[bU.at(k, b)];
                    d2 += dd * dd;
                }
                T magUp = sqrt(d2);
                T G = magUp * g.delta[f];
                T y = 1.0 / g.delta[f];
                T ut = sqrt(par.nu * G);
                if (val(ut) > ROOTVSMALL)
                {
                    for (int it = 0; it < 1000; it++)
                    {
                        T kUu = min(kappa * magUp / ut, T(50.0));
                        T fk = exp(kUu) - 1.0 - kUu * (1.0 + 0.5 * kUu);
                        T ff = -(ut * y) / par.nu + magUp / ut + (fk - kUu * kUu * kUu / 6.0) / E;
                        T df = y / par.nu + magUp / (ut * ut) + kUu * fk / ut / E;
                        T un = ut + ff / df;
                        const double err = std::fabs((val(ut) - val(un)) / val(ut));
                        ut = un;
                        if (!(val(ut) > ROOTVSMALL) || err < 1.0e-14) break;
                    }
                }
                ut = max(ut, T(0.0));
                bNut.val[b] = max(T(0.0), ut * ut / (G + ROOTVSMALL) - par.nu);
            }
        }
    }
    else
    {
        bNut.init(1, nBF);
    }
    std::vector<T> nuEff(nC), nuEffB(nBF);
    for (int c = 0; c < nC; c++) nuEff[c] = nut[c] + par.nu;
    for (int b = 0; b < nBF; b++) nuEffB[b] = bNut.val[b] + par.nu;

    // --- gradients (Gauss linear)
    std::vector<T> gradU, gradP, gradNt;
    fvcGrad(t, g, 3, U, bU, gradU); // gradU[(j*3+i)*nC+c] = d_i U_j
    fvcGrad(t, g, 1, p, bP, gradP);
    if (turb) fvcGrad(t, g, 1, nt, bNt, gradNt);
    auto GU = [&](int i, int j, int c) -> const T& { return gradU[((size_t)j * 3 + i) * nC + c]; };

    // --- UEqn = div(phi,U) - laplacian(nuEff,U) - div(nuEff*dev2(T(grad(U))))
    const int schemeU = isPC ? DIV_UPWIND : par.divU;
    Mat<T> UEqn;
    UEqn.init(t, 3);
    fvmDiv(UEqn, t, g, phi, schemeU, bU, gradU, true, &U);
    fvmLaplacian(UEqn, t, g, -1.0, nuEff, nuEffB, bU, gradU, false);
    {
        // - fvc::div(nuEff*dev2(T(grad(U)))), "Gauss linear": source += sum_f Sf & T_f
        // T = nuEff*( (gradU)^T - 2/3 tr(gradU) I ), (Sf & T)_j = nuEff*( S_i d_j U_i - 2/3 divU S_j )
        auto cellT = [&](int c, const V3<T>& S, V3<T>& out) {
            T tr = GU(0, 0, c) + GU(1, 1, c) + GU(2, 2, c);
            for (int j = 0; j < 3; j++)
            {
                T s = S[0] * GU(j, 0, c) + S[1] * GU(j, 1, c) + S[2] * GU(j, 2, c);
                out[j] = nuEff[c] * (s - (2.0 / 3.0) * tr * S[j]);
            }
        };
        for (int f = 0; f < nIF; f++)
        {
            const int o = t.own[f], n = t.nei[f];
            V3<T> a, b2;
            cellT(o, g.Sf[f], a);
            cellT(n, g.Sf[f], b2);
            for (int j = 0; j < 3; j++)
            {
                T fl = g.w[f] * a[j] + (1.0 - g.w[f]) * b2[j];
                UEqn.src[(size_t)j * nC + o] += fl;
                UEqn.src[(size_t)j * nC + n] -= fl;
            }
        }
        for (int b = 0; b < nBF; b++)
        {
            // boundary value of grad(U): cell value with the normal component replaced by snGrad
            const int f = nIF + b, c = t.own[f];
            V3<T> nh = (T(1.0) / g.magSf[f]) * g.Sf[f];
            T Gb[3][3]; // Gb[i][j] = d_i U_j
            for (int j = 0; j < 3; j++)
            {
                T nG = nh[0] * GU(0, j, c) + nh[1] * GU(1, j, c) + nh[2] * GU(2, j, c);
                for (int i = 0; i < 3; i++) Gb[i][j] = GU(i, j, c) + nh[i] * (bU.sng[bU.at(j, b)] - nG);
            }
            T tr = Gb[0][0] + Gb[1][1] + Gb[2][2];
            for (int j = 0; j < 3; j++)
            {
                T s = g.Sf[f][0] * Gb[j][0] + g.Sf[f][1] * Gb[j][1] + g.Sf[f][2] * Gb[j][2];
                UEqn.src[(size_t)j * nC + c] += nuEffB[b] * (s - (2.0 / 3.0) * tr * g.Sf[f][j]);
            }
        }
    }
    if (!cs.disks.empty())
    {
        // ... - fvSource (UEqnSimple.H): matrix source += fvSource*V
        std::vector<T> S;
        actuatorSource(cs, g, S);
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < nC; c++) UEqn.src[(size_t)k * nC + c] += g.V[c] * S[(size_t)k * nC + c];
    }
    if (cs.mrf.on)
    {
        // + MRF.DDt(U): the Coriolis acceleration Omega x U of the zone cells, explicit (fvMatrix + field: source -= V*field)
        const double* w = cs.mrf.omega;
        for (int c = 0; c < nC; c++)
        {
            if (!cs.mrf.cell[c]) continue;
            const T &u0 = U[c], &u1 = U[(size_t)nC + c], &u2 = U[(size_t)2 * nC + c];
            UEqn.src[c] -= g.V[c] * (w[1] * u2 - w[2] * u1);
            UEqn.src[(size_t)nC + c] -= g.V[c] * (w[2] * u0 - w[0] * u2);
            UEqn.src[(size_t)2 * nC + c] -= g.V[c] * (w[0] * u1 - w[1] * u0);
        }
    }
    relax(UEqn, t, par.alphaU, U);

    // --- URes = (UEqn & U) + grad(p)
    std::vector<T> URes;
    matResidual(UEqn, t, g, U, URes);
    for (int k = 0; k < 3; k++)
        for (int c = 0; c < nC; c++)
        {
            URes[(size_t)k * nC + c] += gradP[(size_t)k * nC + c];
            if (!par.nrU) URes[(size_t)k * nC + c] *= g.V[c];
        }

    // --- rAU, HbyA, phiHbyA
    std::vector<T> A, H;
    matAH(UEqn, t, g, U, A, H);
    std::vector<T> rAU(nC), HbyA((size_t)3 * nC);
    for (int c = 0; c < nC; c++)
    {
        rAU[c] = 1.0 / A[c];
        for (int k = 0; k < 3; k++) HbyA[(size_t)k * nC + c] = rAU[c] * H[(size_t)k * nC + c];
    }
    std::vector<T> phiHbyA(nF);
    for (int f = 0; f < nIF; f++)
    {
        T s(0.0);
        for (int k = 0; k < 3; k++)
            s += g.Sf[f][k] * (g.w[f] * HbyA[(size_t)k * nC + t.own[f]] + (1.0 - g.w[f]) * HbyA[(size_t)k * nC + t.nei[f]]);
        phiHbyA[f] = s;
    }
    for (int b = 0; b < nBF; b++)
    {
        // constrainHbyA: HbyA_b = U_b unless the U patch field is assignable (inletOutlet/zeroGradient),
        // where the extrapolated cell value rAU*H is used
        const int f = nIF + b, c = t.own[f];
        const int kind = cs.bc.kind[F_U * t.nPatch + t.bPatch[b]];
        const bool assignable = (kind == BC_INLET_OUTLET || kind == BC_OUTLET_INLET || kind == BC_ZERO_GRADIENT);
        T s(0.0);
        for (int k = 0; k < 3; k++)
        {
            const T& hb = (par.constrainHbyA && !assignable) ? bU.val[bU.at(k, b)] : HbyA[(size_t)k * nC + c];
            s += g.Sf[f][k] * hb;
        }
        phiHbyA[f] = s;
    }
    if (cs.mrf.on)
    {
        // MRF.makeRelative(phiHbyA)
        for (int f = 0; f < nF; f++)
        {
            const int ty = cs.mrf.faceType[f];
            if (ty == 0) continue;
            if (f >= nIF && ty == 1) phiHbyA[f] = T(0.0);
            else phiHbyA[f] -= mrfFaceFlux(cs, g, f);
        }
    }

    // --- pEqn: laplacian(rAU, p) == div(phiHbyA)
    Mat<T> pEqn;
    pEqn.init(t, 1);
    std::vector<T> rAUB(nBF);
    for (int b = 0; b < nBF; b++) rAUB[b] = rAU[t.own[nIF + b]];
    fvmLaplacian(pEqn, t, g, 1.0, rAU, rAUB, bP, gradP, true);
    for (int f = 0; f < nF; f++)
    {
        pEqn.src[t.own[f]] += phiHbyA[f];
        if (f < nIF) pEqn.src[t.nei[f]] -= phiHbyA[f];
    }
    std::vector<T> pRes;
    matResidual(pEqn, t, g, p, pRes);
    if (!par.nrP)
        for (int c = 0; c < nC; c++) pRes[c] *= g.V[c];

    // --- phiRes = phiHbyA - pEqn.flux() - phi
    std::vector<T> pFlux;
    matFlux(pEqn, t, p, pFlux);
    std::vector<T> phiRes(nF);
    for (int f = 0; f < nF; f++)
    {
        phiRes[f] = phiHbyA[f] - pFlux[f] - phi[f];
        if (par.nrPhi) phiRes[f] /= g.magSf[f];
    }

    // --- SA residual
    std::vector<T> ntRes;
    if (turb)
    {
        const int schemeN = isPC ? DIV_UPWIND : par.divNut;
        Mat<T> nEqn;
        nEqn.init(t, 1);
        fvmDiv(nEqn, t, g, phi, schemeN, bNt, gradNt, true);
        std::vector<T> Dn(nC), DnB(nBF);
        for (int c = 0; c < nC; c++) Dn[c] = (nt[c] + par.nu) / sa.sigmaNut;
        for (int b = 0; b < nBF; b++) DnB[b] = (bNt.val[b] + par.nu) / sa.sigmaNut;
        fvmLaplacian(nEqn, t, g, -1.0, Dn, DnB, bNt, gradNt, false);
        const double Cw1 = sa.Cw1();
        for (int c = 0; c < nC; c++)
        {
            T chi = nt[c] / par.nu;
            T fv1 = fv1f(chi);
            T fv2 = 1.0 - chi / (1.0 + chi * fv1);
            // Omega = sqrt(2)*mag(skew(grad(U)))
            T w01 = 0.5 * (GU(0, 1, c) - GU(1, 0, c)), w02 = 0.5 * (GU(0, 2, c) - GU(2, 0, c)),
              w12 = 0.5 * (GU(1, 2, c) - GU(2, 1, c));
            T Omega = std::sqrt(2.0) * sqrt(2.0 * (w01 * w01 + w02 * w02 + w12 * w12));
            const double ky2 = (sa.kappa * cs.yWall[c]) * (sa.kappa * cs.yWall[c]);
            T Stilda;
            if (par.turb == 2)
            {
                // DASpalartAllmarasFv3.C:158-175, 452-456: fv2 = (1 + chi/Cv2)^-3, fv3, no Cs clip
                const double Cv2 = 5.0;
                T tq = 1.0 + chi / Cv2;
                T t3 = tq * tq * tq;
                T f2 = 1.0 / t3;
                T cb = chi / Cv2;
                T f3 = (1.0 + chi * fv1) * (1.0 / Cv2) * (3.0 * tq + cb * cb) / t3;
                Stilda = f3 * Omega + f2 * nt[c] / ky2;
            }
            else
                Stilda = max(Omega + fv2 * nt[c] / ky2, sa.Cs * Omega);
            T r = min(nt[c] / (max(Stilda, T(1e-15)) * ky2), T(10.0));
            T r2 = r * r;
            T gg = r + sa.Cw2 * (r2 * r2 * r2 - r);
            T g2 = gg * gg;
            const double c6 = std::pow(sa.Cw3, 6.0);
            T fw = gg * pow((1.0 + c6) / (g2 * g2 * g2 + c6), 1.0 / 6.0);
            T mg2 = gradNt[(size_t)0 * nC + c] * gradNt[(size_t)0 * nC + c] + gradNt[(size_t)1 * nC + c] * gradNt[(size_t)1 * nC + c]
                + gradNt[(size_t)2 * nC + c] * gradNt[(size_t)2 * nC + c];
            // - Cb2/sigma*magSqr(grad(nuTilda)) == Cb1*Stilda*nuTilda - Sp(Cw1*fw*nuTilda/y^2, nuTilda)
            nEqn.src[c] += g.V[c] * (sa.Cb2 / sa.sigmaNut * mg2 + sa.Cb1 * Stilda * nt[c]);
            nEqn.diag[c] += g.V[c] * (Cw1 * fw * nt[c] / (cs.yWall[c] * cs.yWall[c]));
        }
        // relax() does not change (M & x) evaluated at x; omitted.
        if (wk && wk->captureNut)
        {
            // DASpalartAllmaras::getFvMatrixFields (DASpalartAllmaras.C:490-529): D(), upper(), lower() of the RELAXED matrix;
            // fvMatrix::D() = diag + (component-averaged) boundary internalCoeffs
            Mat<T> mr = nEqn;
            relax(mr, t, wk->nutAlpha, nt);
            wk->nutD.assign(nC, T(0.0));
            for (int c = 0; c < nC; c++) wk->nutD[c] = mr.diag[c];
            for (int b = 0; b < nBF; b++) wk->nutD[t.own[t.nIF + b]] += mr.ic[b];
            wk->nutUpper = mr.upper;
            wk->nutLower = mr.lower;
        }
        matResidual(nEqn, t, g, nt, ntRes);
        if (!par.nrNut)
            for (int c = 0; c < nC; c++) ntRes[c] *= g.V[c];
    }

    // --- pack residual vector (same layout as W)
    R.assign(cs.nDof(), T(0.0));
    for (int c = 0; c < nC; c++)
        for (int k = 0; k < 3; k++) R[(size_t)3 * c + k] = URes[(size_t)k * nC + c];
    off = (size_t)3 * nC;
    for (int c = 0; c < nC; c++) R[off + c] = pRes[c];
    off += nC;
    if (turb)
    {
        for (int c = 0; c < nC; c++) R[off + c] = ntRes[c];
        off += nC;
    }
    for (int f = 0; f < nF; f++) R[off + f] = phiRes[f];

    if (wk)
    {
        wk->bU = bU; wk->bP = bP; wk->bNt = bNt; wk->bNut = bNut;
        wk->gradU = gradU; wk->nutC = nut;
    }
}


// ---- DARhoSimpleFoam: compressible residual -----------------------------------------------------------------------
// R(W) of DAResidualRhoSimpleFoam::calcResiduals (reference src/adjoint/DAResidual/DAResidualRhoSimpleFoam.C:84-211) with
// DAResidual::updateThermoVars (DAResidual.C:179-293: hePsiThermo, pureMixture, perfectGas, hConst, const/sutherland
// transport), the compressible branches of DATurbulenceModel (rho(), nu() = mu/rho, divDevRhoReff, correctAlphat:
// DATurbulenceModel.C:195-212, 259-330, 378-398) and of DASpalartAllmaras::calcResiduals (DASpalartAllmaras.C:452-462).
// State ordering [U | p | T | nuTilda | phi] (DAStateInfoRhoSimpleFoam.C:40-46); phi is the mass flux.
// OpenFOAM semantics restated: heThermo::alphaEff = CpByCpv*(alpha + alphat) (gamma for e, 1 for h); the energy BCs
// fixedEnergy/gradientEnergy/mixedEnergy reduce, for constant Cp, to T's BC mapped through the linear he(T).
template <class T>
void residualComp(const Case& cs, const Geom<T>& g, const std::vector<T>& W, int isPC, std::vector<T>& R, Work<T>* wk,
                  const std::vector<T>* bcvIn)
{
    const Topo& t = cs.t;
    const Params& par = cs.par;
    const Case::Comp& cp = cs.comp;
    const int nC = t.nC, nF = t.nF, nIF = t.nIF, nBF = t.nBF;
    const bool turb = par.turb != 0;
    SAConst sa;
    std::vector<T> bcvLocal;
    if (!bcvIn)
    {
        bcvLocal.resize(cs.bc.value.size());
        for (size_t i = 0; i < bcvLocal.size(); i++) bcvLocal[i] = T(cs.bc.value[i]);
    }
    const std::vector<T>& bcv = bcvIn ? *bcvIn : bcvLocal;
    // unpack
    std::vector<T> U((size_t)3 * nC), p(nC), Tt(nC), nt(nC, T(0.0)), phi(nF);
    for (int c = 0; c < nC; c++)
        for (int k = 0; k < 3; k++) U[(size_t)k * nC + c] = W[(size_t)3 * c + k];
    size_t off = (size_t)3 * nC;
    for (int c = 0; c < nC; c++) p[c] = W[off + c];
    off += nC;
    for (int c = 0; c < nC; c++) Tt[c] = W[off + c];
    off += nC;
    if (turb)
    {
        for (int c = 0; c < nC; c++) nt[c] = W[off + c];
        off += nC;
    }
    for (int f = 0; f < nF; f++) phi[f] = W[off + f];

    // --- boundary conditions of the states
    BF<T> bU, bP, bT, bNt, bNut;
    evalBC(t, g, cs.bc, bcv, F_U, 3, U, phi, bU);
    if (cs.mrf.on) mrfWallVelocity(cs, g, U, bU);
    evalBC(t, g, cs.bc, bcv, F_P, 1, p, phi, bP);
    BCSpec bcT;
    bcT.kind.assign((size_t)N_FIELDS * t.nPatch, BC_ZERO_GRADIENT);
    std::vector<T> bcvT((size_t)N_FIELDS * t.nPatch * 3, T(0.0));
    for (int pa = 0; pa < t.nPatch; pa++)
    {
        bcT.kind[pa] = cp.kindT[pa]; // stored in slot 0 of a private table
        bcvT[(size_t)pa * 3] = T(cp.valueT[pa]);
    }
    evalBC(t, g, bcT, bcvT, 0, 1, Tt, phi, bT);
    // --- thermo (DAResidual::updateThermoVars): psi = 1/(R T), rho = psi p, he, mu, alpha; cells and boundary faces
    const double Rg = cp.R, Cp = cp.Cp, Cv = cp.Cp - cp.R;
    const double heA = cp.heIsE ? (Cp - Rg) : Cp, heB = -Cp * cp.TRef; // he = heA*T + heB
    const double CpByCpv = cp.heIsE ? Cp / Cv : 1.0;
    auto muOf = [&](const T& Tv) -> T {
        if (!cp.sutherland) return T(cp.mu);
        return cp.As * sqrt(Tv) / (1.0 + cp.Ts / Tv);
    };
    auto alphaOf = [&](const T& muv) -> T {
        if (!cp.sutherland) return muv / cp.Pr;
        return muv * Cv * (1.32 + 1.77 * Rg / Cv) / Cp;
    };
    std::vector<T> rho(nC), mu(nC), alpha(nC), he(nC), nu(nC), rhoB(nBF), muB(nBF), alphaB(nBF), nuB(nBF);
    for (int c = 0; c < nC; c++)
    {
        rho[c] = p[c] / (Rg * Tt[c]);
        mu[c] = muOf(Tt[c]);
        alpha[c] = alphaOf(mu[c]);
        he[c] = heA * Tt[c] + heB;
        nu[c] = mu[c] / rho[c];
    }
    for (int b = 0; b < nBF; b++)
    {
        rhoB[b] = bP.val[b] / (Rg * bT.val[b]);
        muB[b] = muOf(bT.val[b]);
        alphaB[b] = alphaOf(muB[b]);
        nuB[b] = muB[b] / rhoB[b];
    }
    // he boundary coefficients: T's BC kinds with the reference value mapped through he(T)
    BF<T> bHe;
    {
        std::vector<T> bcvH(bcvT);
        for (int pa = 0; pa < t.nPatch; pa++) bcvH[(size_t)pa * 3] = heA * bcvT[(size_t)pa * 3] + heB;
        evalBC(t, g, bcT, bcvH, 0, 1, he, phi, bHe);
    }
    // --- turbulence closures: nut = nuTilda*fv1(nuTilda/nu) (cells and boundaries), alphat = rho*nut/Prt
    std::vector<T> nut(nC, T(0.0));
    if (turb)
    {
        evalBC(t, g, cs.bc, bcv, F_NUTILDA, 1, nt, phi, bNt);
        for (int c = 0; c < nC; c++) nut[c] = nt[c] * fv1f(T(nt[c] / nu[c]));
        evalBC(t, g, cs.bc, bcv, F_NUT, 1, nut, phi, bNut);
        for (int b = 0; b < nBF; b++)
        {
            const int kindN = cs.bc.kind[F_NUT * t.nPatch + t.bPatch[b]];
            if (kindN == BC_CALCULATED)
            {
                const T& nb = bNt.val[b];
                bNut.val[b] = nb * fv1f(T(nb / nuB[b]));
            }
            else if (kindN == BC_NUT_SPALDING)
            {
                const int f = nIF + b, c = t.own[f];
                const double kappa = 0.41, E = 9.8, ROOTVSMALL = 1.0e-150;
                const T nuw = nuB[b];
                T d2(0.0);
                for (int k = 0; k < 3; k++)
                {
                    T dd = U[(size_t)k * nC + c] - bU.val[bU.at(k, b)];
                    d2 += dd * dd;
                }
                T magUp = sqrt(d2);
                T G = magUp * g.delta[f];
                T y = 1.0 / g.delta[f];
                T ut = sqrt(nuw * G);
                if (val(ut) > ROOTVSMALL)
                {
                    for (int it = 0; it < 1000; it++)
                    {
                        T kUu = min(kappa * magUp / ut, T(50.0));
                        T fk = exp(kUu) - 1.0 - kUu * (1.0 + 0.5 * kUu);
                        T ff = -(ut * y) / nuw + magUp / ut + (fk - kUu * kUu * kUu / 6.0) / E;
                        T df = y / nuw + magUp / (ut * ut) + kUu * fk / ut / E;
                        T un = ut + ff / df;
                        const double err = std::fabs((val(ut) - val(un)) / val(ut));
                        ut = un;
                        if (!(val(ut) > ROOTVSMALL) || err < 1.0e-14) break;
                    }
                }
                ut = max(ut, T(0.0));
                bNut.val[b] = max(T(0.0), ut * ut / (G + ROOTVSMALL) - nuw);
            }
        }
    }
    else
        bNut.init(1, nBF);
    // effective transport coefficients (cell fields interpolated linearly by the laplacians)
    std::vector<T> muE(nC), muEB(nBF), aE(nC), aEB(nBF);
    for (int c = 0; c < nC; c++)
    {
        muE[c] = rho[c] * (nu[c] + nut[c]);
        aE[c] = CpByCpv * (alpha[c] + rho[c] * nut[c] / cp.Prt);
    }
    for (int b = 0; b < nBF; b++)
    {
        muEB[b] = rhoB[b] * (nuB[b] + bNut.val[b]);
        aEB[b] = CpByCpv * (alphaB[b] + rhoB[b] * bNut.val[b] / cp.Prt);
    }

    // --- gradients
    std::vector<T> gradU, gradP, gradNt, gradHe;
    fvcGrad(t, g, 3, U, bU, gradU);
    fvcGrad(t, g, 1, p, bP, gradP);
    fvcGrad(t, g, 1, he, bHe, gradHe);
    if (turb) fvcGrad(t, g, 1, nt, bNt, gradNt);
    auto GU = [&](int i, int j, int c) -> const T& { return gradU[((size_t)j * 3 + i) * nC + c]; };

    // --- UEqn = div(phi,U) - laplacian(rho*nuEff,U) - div(rho*nuEff*dev2(T(grad(U))))
    const int schemeU = isPC ? DIV_UPWIND : par.divU;
    Mat<T> UEqn;
    UEqn.init(t, 3);
    fvmDiv(UEqn, t, g, phi, schemeU, bU, gradU, true, &U);
    fvmLaplacian(UEqn, t, g, -1.0, muE, muEB, bU, gradU, false);
    {
        auto cellT = [&](int c, const V3<T>& S, V3<T>& out) {
            T tr = GU(0, 0, c) + GU(1, 1, c) + GU(2, 2, c);
            for (int j = 0; j < 3; j++)
            {
                T s = S[0] * GU(j, 0, c) + S[1] * GU(j, 1, c) + S[2] * GU(j, 2, c);
                out[j] = muE[c] * (s - (2.0 / 3.0) * tr * S[j]);
            }
        };
        for (int f = 0; f < nIF; f++)
        {
            const int o = t.own[f], n = t.nei[f];
            V3<T> a, b2;
            cellT(o, g.Sf[f], a);
            cellT(n, g.Sf[f], b2);
            for (int j = 0; j < 3; j++)
            {
                T fl = g.w[f] * a[j] + (1.0 - g.w[f]) * b2[j];
                UEqn.src[(size_t)j * nC + o] += fl;
                UEqn.src[(size_t)j * nC + n] -= fl;
            }
        }
        for (int b = 0; b < nBF; b++)
        {
            const int f = nIF + b, c = t.own[f];
            V3<T> nh = (T(1.0) / g.magSf[f]) * g.Sf[f];
            T Gb[3][3];
            for (int j = 0; j < 3; j++)
            {
                T nG = nh[0] * GU(0, j, c) + nh[1] * GU(1, j, c) + nh[2] * GU(2, j, c);
                for (int i = 0; i < 3; i++) Gb[i][j] = GU(i, j, c) + nh[i] * (bU.sng[bU.at(j, b)] - nG);
            }
            T tr = Gb[0][0] + Gb[1][1] + Gb[2][2];
            for (int j = 0; j < 3; j++)
            {
                T s = g.Sf[f][0] * Gb[j][0] + g.Sf[f][1] * Gb[j][1] + g.Sf[f][2] * Gb[j][2];
                UEqn.src[(size_t)j * nC + c] += muEB[b] * (s - (2.0 / 3.0) * tr * g.Sf[f][j]);
            }
        }
    }
    std::vector<T> fvS;
    if (!cs.disks.empty())
    {
        actuatorSource(cs, g, fvS);
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < nC; c++) UEqn.src[(size_t)k * nC + c] += g.V[c] * fvS[(size_t)k * nC + c];
    }
    if (cs.mrf.on)
    {
        // + MRF.DDt(rho, U) = rho * (Omega x U) in the zone cells
        const double* w = cs.mrf.omega;
        for (int c = 0; c < nC; c++)
        {
            if (!cs.mrf.cell[c]) continue;
            const T &u0 = U[c], &u1 = U[(size_t)nC + c], &u2 = U[(size_t)2 * nC + c];
            UEqn.src[c] -= g.V[c] * rho[c] * (w[1] * u2 - w[2] * u1);
            UEqn.src[(size_t)nC + c] -= g.V[c] * rho[c] * (w[2] * u0 - w[0] * u2);
            UEqn.src[(size_t)2 * nC + c] -= g.V[c] * rho[c] * (w[0] * u1 - w[1] * u0);
        }
    }
    relax(UEqn, t, par.alphaU, U);
    std::vector<T> URes;
    matResidual(UEqn, t, g, U, URes);
    for (int k = 0; k < 3; k++)
        for (int c = 0; c < nC; c++)
        {
            URes[(size_t)k * nC + c] += gradP[(size_t)k * nC + c];
            if (!par.nrU) URes[(size_t)k * nC + c] *= g.V[c];
        }

    // --- EEqn = div(phi,he) + div(phi, Ekp|K) - laplacian(alphaEff, he)   (relax does not change EEqn & he)
    const int schemeE = isPC ? DIV_UPWIND : cp.divE;
    Mat<T> EEqn;
    EEqn.init(t, 1);
    fvmDiv(EEqn, t, g, phi, schemeE, bHe, gradHe, true);
    {
        std::vector<T> Ek(nC), EkB(nBF);
        for (int c = 0; c < nC; c++)
        {
            T k2 = U[c] * U[c] + U[(size_t)nC + c] * U[(size_t)nC + c] + U[(size_t)2 * nC + c] * U[(size_t)2 * nC + c];
            Ek[c] = 0.5 * k2;
            if (cp.heIsE) Ek[c] += p[c] / rho[c];
        }
        for (int b = 0; b < nBF; b++)
        {
            T k2(0.0);
            for (int k = 0; k < 3; k++) k2 += bU.val[bU.at(k, b)] * bU.val[bU.at(k, b)];
            EkB[b] = 0.5 * k2;
            if (cp.heIsE) EkB[b] += bP.val[b] / rhoB[b];
        }
        // fvc::div(phi, Ekp) with "bounded Gauss upwind|linear": sum_f phi_f Ekp_f - (sum_f phi_f) Ekp_P
        std::vector<T> dv(nC, T(0.0));
        for (int f = 0; f < nF; f++)
        {
            const int o = t.own[f];
            T ef;
            if (f < nIF)
            {
                const int n = t.nei[f];
                if (cp.divEkp == DIV_LINEAR) ef = g.w[f] * Ek[o] + (1.0 - g.w[f]) * Ek[n];
                else ef = (val(phi[f]) >= 0.0) ? Ek[o] : Ek[n];
            }
            else
                ef = EkB[f - nIF];
            T fl = phi[f] * ef;
            dv[o] += fl - phi[f] * Ek[o];
            if (f < nIF) dv[t.nei[f]] -= fl - phi[f] * Ek[t.nei[f]];
        }
        for (int c = 0; c < nC; c++) EEqn.src[c] -= dv[c];
    }
    if (cp.turbo && !cp.heIsE)
    {
        // DAResidualTurboFoam.C:117-121 (the ternary binds the two extra terms to the enthalpy branch):
        //   - fvc::div(Teff.T() & U) + fvc::div(p*(U - URel)),  Teff = -devRhoReff = muEff dev(twoSymm(grad U)),  Gauss linear
        auto workVec = [&](const T G[3][3], const T& mu, const T* u, V3<T>& out) {
            T tr = G[0][0] + G[1][1] + G[2][2];
            for (int j = 0; j < 3; j++)
            {
                T a(0.0);
                for (int i = 0; i < 3; i++)
                {
                    T te = G[i][j] + G[j][i];
                    if (i == j) te -= (2.0 / 3.0) * tr;
                    a += mu * te * u[i];
                }
                out[j] = a;
            }
        };
        std::vector<V3<T>> qC(nC), qB(nBF), vC(nC), vB(nBF);
        for (int c = 0; c < nC; c++)
        {
            T G[3][3], u[3];
            for (int i = 0; i < 3; i++)
            {
                u[i] = U[(size_t)i * nC + c];
                for (int j = 0; j < 3; j++) G[i][j] = GU(i, j, c);
            }
            workVec(G, muE[c], u, qC[c]);
            vC[c] = (cs.mrf.on && cs.mrf.cell[c]) ? mrfVelocity(cs, g.C[c]) : V3<T>();
        }
        for (int b = 0; b < nBF; b++)
        {
            const int f = nIF + b, c = t.own[f];
            V3<T> nh = (T(1.0) / g.magSf[f]) * g.Sf[f];
            T Gb[3][3], u[3];
            for (int j = 0; j < 3; j++)
            {
                T nG = nh[0] * GU(0, j, c) + nh[1] * GU(1, j, c) + nh[2] * GU(2, j, c);
                for (int i = 0; i < 3; i++) Gb[i][j] = GU(i, j, c) + nh[i] * (bU.sng[bU.at(j, b)] - nG);
                u[j] = bU.val[bU.at(j, b)];
            }
            workVec(Gb, muEB[b], u, qB[b]);
            // U - URel on the boundary (MRFZoneDF::makeRelative): Omega x r on the included and excluded faces of the zone
            vB[b] = (cs.mrf.on && cs.mrf.faceType[f] != 0) ? mrfVelocity(cs, g.Cf[f]) : V3<T>();
        }
        for (int f = 0; f < nF; f++)
        {
            const int o = t.own[f];
            T fl(0.0);
            if (f < nIF)
            {
                const int n = t.nei[f];
                for (int k = 0; k < 3; k++)
                    fl += g.Sf[f][k] * (g.w[f] * (qC[o][k] - p[o] * vC[o][k]) + (1.0 - g.w[f]) * (qC[n][k] - p[n] * vC[n][k]));
            }
            else
                for (int k = 0; k < 3; k++) fl += g.Sf[f][k] * (qB[f - nIF][k] - bP.val[f - nIF] * vB[f - nIF][k]);
            EEqn.src[o] += fl;
            if (f < nIF) EEqn.src[t.nei[f]] -= fl;
        }
    }
    fvmLaplacian(EEqn, t, g, -1.0, aE, aEB, bHe, gradHe, false);
    if (!fvS.empty())
        for (int c = 0; c < nC; c++) // - fvSourceEnergy = -(fvSource & U)
            EEqn.src[c] += g.V[c] * (fvS[c] * U[c] + fvS[(size_t)nC + c] * U[(size_t)nC + c] + fvS[(size_t)2 * nC + c] * U[(size_t)2 * nC + c]);
    std::vector<T> TRes;
    matResidual(EEqn, t, g, he, TRes);
    if (!cp.nrT)
        for (int c = 0; c < nC; c++) TRes[c] *= g.V[c];

    // --- rAU, HbyA, phiHbyA = interpolate(rho)*flux(HbyA)
    std::vector<T> A, H;
    matAH(UEqn, t, g, U, A, H);
    std::vector<T> rAU(nC), HbyA((size_t)3 * nC), rhorAU(nC), rhorAUB(nBF);
    for (int c = 0; c < nC; c++)
    {
        rAU[c] = 1.0 / A[c];
        rhorAU[c] = rho[c] * rAU[c];
        for (int k = 0; k < 3; k++) HbyA[(size_t)k * nC + c] = rAU[c] * H[(size_t)k * nC + c];
    }
    for (int b = 0; b < nBF; b++) rhorAUB[b] = rhoB[b] * rAU[t.own[nIF + b]];
    std::vector<T> phiHbyA(nF);
    for (int f = 0; f < nIF; f++)
    {
        T s(0.0);
        for (int k = 0; k < 3; k++)
            s += g.Sf[f][k] * (g.w[f] * HbyA[(size_t)k * nC + t.own[f]] + (1.0 - g.w[f]) * HbyA[(size_t)k * nC + t.nei[f]]);
        phiHbyA[f] = (g.w[f] * rho[t.own[f]] + (1.0 - g.w[f]) * rho[t.nei[f]]) * s;
    }
    for (int b = 0; b < nBF; b++)
    {
        const int f = nIF + b, c = t.own[f];
        const int kind = cs.bc.kind[F_U * t.nPatch + t.bPatch[b]];
        const bool assignable = (kind == BC_INLET_OUTLET || kind == BC_OUTLET_INLET || kind == BC_ZERO_GRADIENT);
        T s(0.0);
        for (int k = 0; k < 3; k++)
        {
            const T& hb = (par.constrainHbyA && !assignable) ? bU.val[bU.at(k, b)] : HbyA[(size_t)k * nC + c];
            s += g.Sf[f][k] * hb;
        }
        phiHbyA[f] = rhoB[b] * s;
    }
    if (cs.mrf.on)
    {
        // MRF.makeRelative(fvc::interpolate(rho), phiHbyA)
        for (int f = 0; f < nF; f++)
        {
            const int ty = cs.mrf.faceType[f];
            if (ty == 0) continue;
            if (f >= nIF)
            {
                if (ty == 1) phiHbyA[f] = T(0.0);
                else phiHbyA[f] -= rhoB[f - nIF] * mrfFaceFlux(cs, g, f);
            }
            else
                phiHbyA[f] -= (g.w[f] * rho[t.own[f]] + (1.0 - g.w[f]) * rho[t.nei[f]]) * mrfFaceFlux(cs, g, f);
        }
    }
    // --- pEqn = div(phiHbyA) - laplacian(rhorAUf, p)
    Mat<T> pEqn;
    pEqn.init(t, 1);
    fvmLaplacian(pEqn, t, g, -1.0, rhorAU, rhorAUB, bP, gradP, true);
    if (cp.transonic)
    {
        // simple_.transonic() (DAResidualTurboFoam.C:148-189; DAResidualRhoSimpleCFoam.C:148-200 reduces to the same rows because
        // interpolate(psi*p) == interpolate(rho) cancels its phiHbyA): phid = interpolate(psi)*(interpolate(HbyA) & Sf), made
        // relative; pEqn = fvm::div(phid, p) - fvm::laplacian(rho*rAU, p); phiRes = pEqn.flux() - phi.  pEqn.relax() leaves
        // (pEqn & p) and the flux unchanged.  phiHbyA (already rho_f * relative volume flux) / rho_f = the relative volume flux.
        std::vector<T> phid(nF);
        for (int f = 0; f < nF; f++)
        {
            T psif, rhof;
            if (f < nIF)
            {
                const int o = t.own[f], n = t.nei[f];
                psif = g.w[f] * (rho[o] / p[o]) + (1.0 - g.w[f]) * (rho[n] / p[n]);
                rhof = g.w[f] * rho[o] + (1.0 - g.w[f]) * rho[n];
            }
            else
            {
                psif = rhoB[f - nIF] / bP.val[f - nIF];
                rhof = rhoB[f - nIF];
            }
            phid[f] = psif * (phiHbyA[f] / rhof);
            phiHbyA[f] = T(0.0);
        }
        const bool dropped = isPC && cp.transonicPCOption == 1; // "for PC we do not include the div(phid, p) term"
        if (!dropped) fvmDiv(pEqn, t, g, phid, isPC ? (int)DIV_UPWIND : cp.divPhidP, bP, gradP, false, &p, cp.phidK);
    }
    for (int f = 0; f < nF; f++)
    {
        pEqn.src[t.own[f]] -= phiHbyA[f];
        if (f < nIF) pEqn.src[t.nei[f]] += phiHbyA[f];
    }
    std::vector<T> pRes;
    matResidual(pEqn, t, g, p, pRes);
    if (!par.nrP)
        for (int c = 0; c < nC; c++) pRes[c] *= g.V[c];
    // --- phiRes = phiHbyA + pEqn.flux() - phi
    std::vector<T> pFlux;
    matFlux(pEqn, t, p, pFlux);
    std::vector<T> phiRes(nF);
    for (int f = 0; f < nF; f++)
    {
        phiRes[f] = phiHbyA[f] + pFlux[f] - phi[f];
        if (cp.transonic && isPC && cp.transonicPCOption == 2) phiRes[f] = phi[f]; // DAResidualTurboFoam.C:173-178
        if (par.nrPhi) phiRes[f] /= g.magSf[f];
    }

    // --- SA residual (compressible form)
    std::vector<T> ntRes;
    if (turb)
    {
        const int schemeN = isPC ? DIV_UPWIND : par.divNut;
        Mat<T> nEqn;
        nEqn.init(t, 1);
        fvmDiv(nEqn, t, g, phi, schemeN, bNt, gradNt, true);
        std::vector<T> Dn(nC), DnB(nBF);
        for (int c = 0; c < nC; c++) Dn[c] = rho[c] * (nt[c] + nu[c]) / sa.sigmaNut;
        for (int b = 0; b < nBF; b++) DnB[b] = rhoB[b] * (bNt.val[b] + nuB[b]) / sa.sigmaNut;
        fvmLaplacian(nEqn, t, g, -1.0, Dn, DnB, bNt, gradNt, false);
        const double Cw1 = sa.Cw1();
        for (int c = 0; c < nC; c++)
        {
            T chi = nt[c] / nu[c];
            T fv1 = fv1f(chi);
            T fv2 = 1.0 - chi / (1.0 + chi * fv1);
            T w01 = 0.5 * (GU(0, 1, c) - GU(1, 0, c)), w02 = 0.5 * (GU(0, 2, c) - GU(2, 0, c)),
              w12 = 0.5 * (GU(1, 2, c) - GU(2, 1, c));
            T Omega = std::sqrt(2.0) * sqrt(2.0 * (w01 * w01 + w02 * w02 + w12 * w12));
            const double ky2 = (sa.kappa * cs.yWall[c]) * (sa.kappa * cs.yWall[c]);
            T Stilda;
            if (par.turb == 2)
            {
                const double Cv2 = 5.0;
                T tq = 1.0 + chi / Cv2;
                T t3 = tq * tq * tq;
                T f2 = 1.0 / t3;
                T cb = chi / Cv2;
                T f3 = (1.0 + chi * fv1) * (1.0 / Cv2) * (3.0 * tq + cb * cb) / t3;
                Stilda = f3 * Omega + f2 * nt[c] / ky2;
            }
            else
                Stilda = max(Omega + fv2 * nt[c] / ky2, sa.Cs * Omega);
            T r = min(nt[c] / (max(Stilda, T(1e-15)) * ky2), T(10.0));
            T r2 = r * r;
            T gg = r + sa.Cw2 * (r2 * r2 * r2 - r);
            T g2 = gg * gg;
            const double c6 = std::pow(sa.Cw3, 6.0);
            T fw = gg * pow((1.0 + c6) / (g2 * g2 * g2 + c6), 1.0 / 6.0);
            T mg2 = gradNt[(size_t)0 * nC + c] * gradNt[(size_t)0 * nC + c] + gradNt[(size_t)1 * nC + c] * gradNt[(size_t)1 * nC + c]
                + gradNt[(size_t)2 * nC + c] * gradNt[(size_t)2 * nC + c];
            nEqn.src[c] += g.V[c] * rho[c] * (sa.Cb2 / sa.sigmaNut * mg2 + sa.Cb1 * Stilda * nt[c]);
            nEqn.diag[c] += g.V[c] * rho[c] * (Cw1 * fw * nt[c] / (cs.yWall[c] * cs.yWall[c]));
        }
        matResidual(nEqn, t, g, nt, ntRes);
        if (!par.nrNut)
            for (int c = 0; c < nC; c++) ntRes[c] *= g.V[c];
    }

    // --- pack [URes | pRes | TRes | nuTildaRes | phiRes]
    R.assign(cs.nDof(), T(0.0));
    for (int c = 0; c < nC; c++)
        for (int k = 0; k < 3; k++) R[(size_t)3 * c + k] = URes[(size_t)k * nC + c];
    off = (size_t)3 * nC;
    for (int c = 0; c < nC; c++) R[off + c] = pRes[c];
    off += nC;
    for (int c = 0; c < nC; c++) R[off + c] = TRes[c];
    off += nC;
    if (turb)
    {
        for (int c = 0; c < nC; c++) R[off + c] = ntRes[c];
        off += nC;
    }
    for (int f = 0; f < nF; f++) R[off + f] = phiRes[f];
    if (wk)
    {
        wk->bU = bU; wk->bP = bP; wk->bNt = bNt; wk->bNut = bNut;
        wk->gradU = gradU; wk->nutC = nut;
        wk->muEB = muEB;
        wk->rhoB = rhoB;
        wk->TB = bT.val;
    }
}

// DAFunctionForce::calcFunction: sum over the faces of one patch of (Sf*p_b + Sf & devRhoReff_b) . dir
// mode 0: force . dir (DAFunctionForce); mode 1: ((Cf - center) x force) . dir (DAFunctionMoment.C:60-140)
template <class T>
T forceFunction(const Case& cs, const Geom<T>& g, const std::vector<T>& W, int patch, const double* dir, double scale, int mode = 0,
                 const double* center = nullptr)
{
    std::vector<T> R;
    Work<T> wk;
    residual(cs, g, W, 0, R, &wk);
    const Topo& t = cs.t;
    const int nC = t.nC;
    T F(0.0);
    if (mode >= 2)
    {
        // mode 2: DAFunctionTotalPressure (area-averaged p + 0.5 rho |U|^2); mode 3: DAFunctionMassFlowRate (rho U.Sf);
        // mode 4: the area-averaged isentropic total pressure p (1 + (gamma-1)/2 Ma^2)^(gamma/(gamma-1)) of one side of
        // DAFunctionTotalPressureRatio (DAFunctionTotalPressureRatio.C:50-140), gamma = dir[0], R = Cp - Cp/gamma
        T areaSum(0.0);
        for (int b = 0; b < t.nBF; b++)
            if (t.bPatch[b] == patch) areaSum += g.magSf[t.nIF + b];
        for (int b = 0; b < t.nBF; b++)
        {
            if (t.bPatch[b] != patch) continue;
            const int f = t.nIF + b;
            T rhob = cs.comp.on ? wk.rhoB[b] : T(1.0);
            T U2(0.0), SU(0.0);
            for (int k = 0; k < 3; k++)
            {
                U2 += wk.bU.val[wk.bU.at(k, b)] * wk.bU.val[wk.bU.at(k, b)];
                SU += g.Sf[f][k] * wk.bU.val[wk.bU.at(k, b)];
            }
            if (mode == 2) F += scale * (wk.bP.val[b] + 0.5 * rhob * U2) * g.magSf[f] / areaSum;
            else if (mode == 4)
            {
                const double gam = dir[0], Rg = cs.comp.Cp - cs.comp.Cp / gam;
                T Ma2 = U2 / (gam * Rg * wk.TB[b]);
                T pT = wk.bP.val[b] * pow(T(1.0 + 0.5 * (gam - 1.0) * Ma2), gam / (gam - 1.0));
                F += scale * pT * g.magSf[f] / areaSum;
            }
            else F += scale * rhob * SU;
        }
        return F;
    }
    for (int b = 0; b < t.nBF; b++)
    {
        if (t.bPatch[b] != patch) continue;
        const int f = t.nIF + b, c = t.own[f];
        V3<T> nh = (T(1.0) / g.magSf[f]) * g.Sf[f];
        T Gb[3][3];
        for (int j = 0; j < 3; j++)
        {
            T nG(0.0);
            for (int i = 0; i < 3; i++) nG += nh[i] * wk.gradU[((size_t)j * 3 + i) * nC + c];
            for (int i = 0; i < 3; i++) Gb[i][j] = wk.gradU[((size_t)j * 3 + i) * nC + c] + nh[i] * (wk.bU.sng[wk.bU.at(j, b)] - nG);
        }
        T tr = Gb[0][0] + Gb[1][1] + Gb[2][2];
        T nuEffB = cs.comp.on ? wk.muEB[b] : T(wk.bNut.val[b] + cs.par.nu); // compressible: devRhoReff = -rho*nuEff*dev(twoSymm(grad U))
        T fv(0.0);
        for (int j = 0; j < 3; j++)
        {
            // (Sf & devRhoReff)_j = -nuEff * S_i * dev(twoSymm(G))_ij
            T s(0.0);
            for (int i = 0; i < 3; i++) s += g.Sf[f][i] * (Gb[i][j] + Gb[j][i]);
            s -= (2.0 / 3.0) * tr * g.Sf[f][j];
            T fj = g.Sf[f][j] * wk.bP.val[b] - nuEffB * s;
            if (mode == 0)
                fv += fj * dir[j];
            else
            {
                // (r x F) . a = F . (a x r)
                const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                T r1 = g.Cf[f][j1] - center[j1], r2 = g.Cf[f][j2] - center[j2];
                fv += fj * (dir[j1] * r2 - dir[j2] * r1);
            }
        }
        F += scale * fv;
    }
    return F;
}

} // namespace orc

// =====================================================================================================
// C interface (ctypes), test infrastructure only
// =====================================================================================================
using namespace orc;

extern "C"
{

void* orc_create(int nP, const double* points, int nF, const int* fOff, const int* fLab, const int* owner, int nIF,
                 const int* neighbour, int nPatch, const int* pStart, const int* pSize, const int* pGeom,
                 const int* bcKind /*[4][nPatch]*/, const double* bcValue /*[4][nPatch][3]*/,
                 const double* dpar /*nu, alphaU, sU, sP, sNut, sPhi*/,
                 const int* ipar /*turb, divU, divNut, nrU, nrP, nrNut, nrPhi, constrainHbyA*/, const double* yWall)
{
    Case* cs = new Case();
    Topo& t = cs->t;
    t.nP = nP; t.nF = nF; t.nIF = nIF; t.nBF = nF - nIF; t.nPatch = nPatch;
    t.fOff.assign(fOff, fOff + nF + 1);
    t.fLab.assign(fLab, fLab + fOff[nF]);
    t.own.assign(owner, owner + nF);
    t.nei.assign(neighbour, neighbour + nIF);
    t.nC = *std::max_element(t.own.begin(), t.own.end()) + 1;
    t.pStart.assign(pStart, pStart + nPatch);
    t.pSize.assign(pSize, pSize + nPatch);
    t.pGeom.assign(pGeom, pGeom + nPatch);
    t.bPatch.assign(t.nBF, -1);
    for (int pa = 0; pa < nPatch; pa++)
        for (int i = 0; i < pSize[pa]; i++) t.bPatch[pStart[pa] - nIF + i] = pa;
    cs->bc.kind.assign(bcKind, bcKind + N_FIELDS * nPatch);
    cs->bc.value.assign(bcValue, bcValue + N_FIELDS * nPatch * 3);
    Params& q = cs->par;
    q.nu = dpar[0]; q.alphaU = dpar[1]; q.sU = dpar[2]; q.sP = dpar[3]; q.sNut = dpar[4]; q.sPhi = dpar[5];
    q.turb = ipar[0]; q.divU = ipar[1]; q.divNut = ipar[2];
    q.nrU = ipar[3]; q.nrP = ipar[4]; q.nrNut = ipar[5]; q.nrPhi = ipar[6]; q.constrainHbyA = ipar[7];
    q.phiNorm = ipar[8];
    cs->pts.assign(points, points + 3 * (size_t)nP);
    std::vector<V3<double>> P(nP);
    for (int i = 0; i < nP; i++) P[i] = V3<double>(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    computeGeometry(t, P, cs->gd);
    if (yWall)
        cs->yWall.assign(yWall, yWall + t.nC);
    else
    {
        // frozen wall distance: distance from the cell centre to the nearest wall-face centre
        cs->yWall.assign(t.nC, 1e30);
        for (int c = 0; c < t.nC; c++)
            for (int b = 0; b < t.nBF; b++)
                if (t.pGeom[t.bPatch[b]] == 1)
                {
                    V3<double> d = cs->gd.C[c] - cs->gd.Cf[t.nIF + b];
                    cs->yWall[c] = std::min(cs->yWall[c], mag(d));
                }
    }
    return cs;
}

// switch the case to DARhoSimpleFoam: dpar = R, Cp, mu, Pr, Prt, As, Ts, TRef, sT; ipar = heIsE, sutherland, divE, divEkp, nrT
void orc_set_compressible(void* h, const double* dpar, const int* ipar, const int* kindT, const double* valueT)
{
    Case* cs = (Case*)h;
    Case::Comp& c = cs->comp;
    c.on = 1;
    c.R = dpar[0]; c.Cp = dpar[1]; c.mu = dpar[2]; c.Pr = dpar[3]; c.Prt = dpar[4]; c.As = dpar[5]; c.Ts = dpar[6]; c.TRef = dpar[7]; c.sT = dpar[8];
    c.heIsE = ipar[0]; c.sutherland = ipar[1]; c.divE = ipar[2]; c.divEkp = ipar[3]; c.nrT = ipar[4];
    c.kindT.assign(kindT, kindT + cs->t.nPatch);
    c.valueT.assign(valueT, valueT + cs->t.nPatch);
    cs->recorded = false;
}

// transonic pressure equation: scheme of div(phid,p) (DivScheme; 4 = limitedLinear k), transonicPCOption (-1, 1, 2)
void orc_set_transonic(void* h, int on, int scheme, double k, int pcOption)
{
    Case* cs = (Case*)h;
    cs->comp.transonic = on;
    cs->comp.divPhidP = scheme;
    cs->comp.phidK = k;
    cs->comp.transonicPCOption = pcOption;
    cs->recorded = false;
}

void orc_set_turbo(void* h, int on)
{
    Case* cs = (Case*)h;
    cs->comp.turbo = on;
    cs->recorded = false;
}

// actuator disks: 15 numbers each (13 actuatorDiskPars, eps, rotLeft)
void orc_set_fvsource(void* h, int nDisk, const double* pars)
{
    Case* cs = (Case*)h;
    cs->disks.assign(pars, pars + (size_t)15 * nDisk);
    cs->recorded = false;
}

// MRF zone: Omega = omega*axis [rad/s], origin, cell mask [nC], mask of the nonRotatingPatches [nPatch]; the face types follow
// MRFZoneDF::setMRFFaces (reference MRFZoneDF.C)
void orc_set_mrf(void* h, const double* omega, const double* origin, const int* cellMask, const int* excludedPatch)
{
    Case* cs = (Case*)h;
    const Topo& t = cs->t;
    Case::Mrf& m = cs->mrf;
    m.on = 1;
    for (int k = 0; k < 3; k++)
    {
        m.omega[k] = omega[k];
        m.origin[k] = origin[k];
    }
    m.cell.assign(t.nC, 0);
    for (int c = 0; c < t.nC; c++) m.cell[c] = cellMask[c] ? 1 : 0;
    m.faceType.assign(t.nF, 0);
    for (int f = 0; f < t.nIF; f++)
        if (m.cell[t.own[f]] || m.cell[t.nei[f]]) m.faceType[f] = 1;
    for (int b = 0; b < t.nBF; b++)
    {
        const int f = t.nIF + b;
        if (m.cell[t.own[f]]) m.faceType[f] = excludedPatch[t.bPatch[b]] ? 2 : 1;
    }
    cs->recorded = false;
}

void orc_destroy(void* h) { delete (Case*)h; }
int orc_ndof(void* h) { return ((Case*)h)->nDof(); }
int orc_ncells(void* h) { return ((Case*)h)->t.nC; }

// geometry getters for cross-checking the product's host geometry: what = 0 V,1 magSf,2 w,3 delta,4 yWall,
// 5 C(3nC),6 Sf(3nF),7 Cf(3nF),8 corr(3nF)
void orc_get_geometry(void* h, int what, double* out)
{
    Case* cs = (Case*)h;
    const Geom<double>& g = cs->gd;
    const Topo& t = cs->t;
    switch (what)
    {
    case 0: for (int c = 0; c < t.nC; c++) out[c] = g.V[c]; break;
    case 1: for (int f = 0; f < t.nF; f++) out[f] = g.magSf[f]; break;
    case 2: for (int f = 0; f < t.nF; f++) out[f] = g.w[f]; break;
    case 3: for (int f = 0; f < t.nF; f++) out[f] = g.delta[f]; break;
    case 4: for (int c = 0; c < t.nC; c++) out[c] = cs->yWall[c]; break;
    case 5: for (int c = 0; c < t.nC; c++) for (int k = 0; k < 3; k++) out[3 * c + k] = g.C[c][k]; break;
    case 6: for (int f = 0; f < t.nF; f++) for (int k = 0; k < 3; k++) out[3 * f + k] = g.Sf[f][k]; break;
    case 7: for (int f = 0; f < t.nF; f++) for (int k = 0; k < 3; k++) out[3 * f + k] = g.Cf[f][k]; break;
    case 8: for (int f = 0; f < t.nF; f++) for (int k = 0; k < 3; k++) out[3 * f + k] = g.corr[f][k]; break;
    }
}

void orc_residual(void* h, const double* W, int isPC, double* R)
{
    Case* cs = (Case*)h;
    std::vector<double> w(W, W + cs->nDof()), r;
    residual<double>(*cs, cs->gd, w, isPC, r);
    std::copy(r.begin(), r.end(), R);
}

// DASpalartAllmaras::getFvMatrixFields with the `div(pc)` scheme (isPC = 1): D [nC], upper [nIF], lower [nIF] of the relaxed matrix
void orc_nut_fvmatrix(void* h, const double* W, double alpha, double* D, double* upper, double* lower)
{
    Case* cs = (Case*)h;
    std::vector<double> w(W, W + cs->nDof()), r;
    Work<double> wk;
    wk.captureNut = true;
    wk.nutAlpha = alpha;
    residual<double>(*cs, cs->gd, w, 1, r, &wk);
    std::copy(wk.nutD.begin(), wk.nutD.end(), D);
    std::copy(wk.nutUpper.begin(), wk.nutUpper.end(), upper);
    std::copy(wk.nutLower.begin(), wk.nutLower.end(), lower);
}

// J v by forward-mode dual numbers (one tangent direction): the exact counterpart of orc_jtvec, used by the tests to check the
// tape to rounding instead of to finite-difference accuracy
void orc_jvec(void* h, const double* W, const double* v, int isPC, double* out)
{
    Case* cs = (Case*)h;
    const int n = cs->nDof();
    std::vector<Dual> w(n), r;
    for (int i = 0; i < n; i++) w[i] = Dual(W[i], v[i]);
    Geom<Dual> g;
    std::vector<V3<Dual>> P(cs->t.nP);
    for (int i = 0; i < cs->t.nP; i++) P[i] = V3<Dual>(Dual(cs->pts[3 * i]), Dual(cs->pts[3 * i + 1]), Dual(cs->pts[3 * i + 2]));
    computeGeometry(cs->t, P, g);
    residual<Dual>(*cs, g, w, isPC, r);
    for (int i = 0; i < n; i++) out[i] = r[i].d;
}

// state scaling of a product vector: DASolver::normalizeGradientVec (DASolver.C:2356-2455)
static void scaleStates(const Case* cs, double* y)
{
    const Topo& t = cs->t;
    const Params& q = cs->par;
    size_t off = 0;
    for (int i = 0; i < 3 * t.nC; i++) y[off + i] *= q.sU;
    off += (size_t)3 * t.nC;
    for (int i = 0; i < t.nC; i++) y[off + i] *= q.sP;
    off += t.nC;
    if (cs->comp.on)
    {
        for (int i = 0; i < t.nC; i++) y[off + i] *= cs->comp.sT;
        off += t.nC;
    }
    if (q.turb)
    {
        for (int i = 0; i < t.nC; i++) y[off + i] *= q.sNut;
        off += t.nC;
    }
    if (q.phiNorm)
        for (int f = 0; f < t.nF; f++) y[off + f] *= q.sPhi * cs->gd.magSf[f];
}

// DASolver::initializeGlobalADTape4dRdWT: record R(W) once on the global tape
long orc_record(void* h, const double* W, int isPC)
{
    Case* cs = (Case*)h;
    Tape& tp = tape();
    tp.reset();
    const int n = cs->nDof();
    std::vector<AReal> w(n), r;
    cs->inId.resize(n);
    for (int i = 0; i < n; i++)
    {
        w[i] = AReal(W[i]);
        w[i].registerInput();
        cs->inId[i] = w[i].id;
    }
    Geom<AReal> g;
    std::vector<V3<AReal>> P(cs->t.nP);
    for (int i = 0; i < cs->t.nP; i++) P[i] = V3<AReal>(AReal(cs->pts[3 * i]), AReal(cs->pts[3 * i + 1]), AReal(cs->pts[3 * i + 2]));
    computeGeometry(cs->t, P, g);
    residual<AReal>(*cs, g, w, isPC, r);
    cs->outId.resize(n);
    for (int i = 0; i < n; i++) cs->outId[i] = r[i].id;
    cs->recorded = true;
    return (long)tp.size();
}

// DASolver::dRdWTMatVecMultFunction: seed residual adjoints, evaluate tape, read state adjoints, scale
void orc_jtvec(void* h, const double* psi, double* out, int normalize)
{
    Case* cs = (Case*)h;
    Tape& tp = tape();
    const int n = cs->nDof();
    cs->adj.assign(tp.size() + 1, 0.0);
    for (int i = 0; i < n; i++)
        if (cs->outId[i]) cs->adj[cs->outId[i]] += psi[i];
    tp.evaluate(cs->adj);
    for (int i = 0; i < n; i++) out[i] = cs->adj[cs->inId[i]];
    if (normalize) scaleStates(cs, out);
}

double orc_force(void* h, const double* W, int patch, const double* dir, double scale, int mode, const double* center)
{
    Case* cs = (Case*)h;
    std::vector<double> w(W, W + cs->nDof());
    return forceFunction<double>(*cs, cs->gd, w, patch, dir, scale, mode, center);
}

// calcJacTVecProduct(stateVar -> function): dF/dW * seed, scaled like the reference (DASolver.C:1819-1820)
void orc_dforce_dw(void* h, const double* W, int patch, const double* dir, double scale, double seed, double* out, int normalize, int mode,
                   const double* center)
{
    Case* cs = (Case*)h;
    Tape& tp = tape();
    tp.reset();
    const int n = cs->nDof();
    std::vector<AReal> w(n);
    std::vector<int> ids(n);
    for (int i = 0; i < n; i++)
    {
        w[i] = AReal(W[i]);
        w[i].registerInput();
        ids[i] = w[i].id;
    }
    Geom<AReal> g;
    std::vector<V3<AReal>> P(cs->t.nP);
    for (int i = 0; i < cs->t.nP; i++) P[i] = V3<AReal>(AReal(cs->pts[3 * i]), AReal(cs->pts[3 * i + 1]), AReal(cs->pts[3 * i + 2]));
    computeGeometry(cs->t, P, g);
    AReal F = forceFunction<AReal>(*cs, g, w, patch, dir, scale, mode, center);
    std::vector<double> adj(tp.size() + 1, 0.0);
    if (F.id) adj[F.id] = seed;
    tp.evaluate(adj);
    for (int i = 0; i < n; i++) out[i] = adj[ids[i]];
    if (normalize) scaleStates(cs, out);
    tp.reset();
    cs->recorded = false;
}

// calcJacTVecProduct(volCoord -> residual): [dR/dXv]^T psi (DAInputVolCoord.C:35-70)
void orc_jtvec_xv(void* h, const double* W, const double* psi, double* out /*3*nP*/)
{
    Case* cs = (Case*)h;
    Tape& tp = tape();
    tp.reset();
    const int n = cs->nDof();
    std::vector<V3<AReal>> P(cs->t.nP);
    std::vector<int> ids(3 * (size_t)cs->t.nP);
    for (int i = 0; i < cs->t.nP; i++)
        for (int k = 0; k < 3; k++)
        {
            AReal a(cs->pts[3 * i + k]);
            a.registerInput();
            ids[3 * (size_t)i + k] = a.id;
            P[i][k] = a;
        }
    Geom<AReal> g;
    computeGeometry(cs->t, P, g);
    std::vector<AReal> w(n), r;
    for (int i = 0; i < n; i++) w[i] = AReal(W[i]);
    residual<AReal>(*cs, g, w, 0, r);
    std::vector<double> adj(tp.size() + 1, 0.0);
    for (int i = 0; i < n; i++)
        if (r[i].id) adj[r[i].id] += psi[i];
    tp.evaluate(adj);
    for (size_t i = 0; i < ids.size(); i++) out[i] = adj[ids[i]];
    tp.reset();
    cs->recorded = false;
}

// calcJacTVecProduct(patchVelocity -> residual) building block: [dR/d(U reference value of one patch)]^T psi
// (the reference sets the patch values from (|U|, angle of attack) in DAInputPatchVelocity)
void orc_jtvec_bc(void* h, const double* W, const double* psi, int field, int patch, double* out3);
void orc_jtvec_bcU(void* h, const double* W, const double* psi, int patch, double* out3) { orc_jtvec_bc(h, W, psi, F_U, patch, out3); }

// [dR/d(boundary reference value of `field` on `patch`)]^T psi (DAInputPatchVar / DAInputPatchVelocity through the tape)
void orc_jtvec_bc(void* h, const double* W, const double* psi, int field, int patch, double* out3)
{
    Case* cs = (Case*)h;
    Tape& tp = tape();
    tp.reset();
    const int n = cs->nDof();
    std::vector<AReal> bcv(cs->bc.value.size());
    for (size_t i = 0; i < bcv.size(); i++) bcv[i] = AReal(cs->bc.value[i]);
    int ids[3];
    for (int k = 0; k < 3; k++)
    {
        AReal& a = bcv[(field * cs->t.nPatch + patch) * 3 + k];
        a.registerInput();
        ids[k] = a.id;
    }
    Geom<AReal> g;
    std::vector<V3<AReal>> P(cs->t.nP);
    for (int i = 0; i < cs->t.nP; i++) P[i] = V3<AReal>(AReal(cs->pts[3 * i]), AReal(cs->pts[3 * i + 1]), AReal(cs->pts[3 * i + 2]));
    computeGeometry(cs->t, P, g);
    std::vector<AReal> w(n), r;
    for (int i = 0; i < n; i++) w[i] = AReal(W[i]);
    residual<AReal>(*cs, g, w, 0, r, nullptr, &bcv);
    std::vector<double> adj(tp.size() + 1, 0.0);
    for (int i = 0; i < n; i++)
        if (r[i].id) adj[r[i].id] += psi[i];
    tp.evaluate(adj);
    for (int k = 0; k < 3; k++) out3[k] = adj[ids[k]];
    tp.reset();
    cs->recorded = false;
}

void orc_set_bc_value(void* h, int field, int patch, const double* v3)
{
    Case* cs = (Case*)h;
    for (int k = 0; k < 3; k++) cs->bc.value[(field * cs->t.nPatch + patch) * 3 + k] = v3[k];
    cs->recorded = false;
}

long orc_tape_size() { return (long)tape().size(); }

} // extern "C"
