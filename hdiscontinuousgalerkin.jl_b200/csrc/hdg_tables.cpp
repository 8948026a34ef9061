// Host-side construction of the reference-element tables (see hdg_tables.h).
#include "hdg_tables.h"

#include <cmath>
#include <cstddef>

namespace hdg {
namespace {

// Minimal forward-mode dual number in (r,s): value + both partials.  Plays the role Tensors.jl's
// `gradient(f, xi, :all)` plays in the reference (src/ScalarFunctionSpaces.jl:82).
struct D2 {
    double v, r, s;
};
inline D2 operator+(D2 a, D2 b) { return {a.v + b.v, a.r + b.r, a.s + b.s}; }
inline D2 operator-(D2 a, D2 b) { return {a.v - b.v, a.r - b.r, a.s - b.s}; }
inline D2 operator*(D2 a, D2 b) { return {a.v * b.v, a.r * b.v + a.v * b.r, a.s * b.v + a.v * b.s}; }
inline D2 operator*(double c, D2 a) { return {c * a.v, c * a.r, c * a.s}; }
inline D2 cst(double c) { return {c, 0.0, 0.0}; }

// Q_n(a,b) = b^n P_n(a/b), P_n Legendre: the collapsed-coordinate factor of the Dubiner basis
// written as a homogeneous polynomial, so no division by (1-s) is needed anywhere.
D2 homog_legendre(int n, D2 a, D2 b) {
    D2 q0 = cst(1.0);
    if (n == 0) return q0;
    D2 q1 = a;
    D2 b2 = b * b;
    for (int k = 1; k < n; ++k) {
        D2 q2 = (1.0 / (k + 1)) * ((2.0 * k + 1.0) * (a * q1) - double(k) * (b2 * q0));
        q0 = q1;
        q1 = q2;
    }
    return q1;
}

// Jacobi polynomial P_m^{(alpha,0)}(x), three-term recurrence.
D2 jacobi_a0(int m, double alpha, D2 x) {
    D2 p0 = cst(1.0);
    if (m == 0) return p0;
    D2 p1 = 0.5 * ((alpha + 2.0) * x + cst(alpha));
    for (int k = 2; k <= m; ++k) {
        double c = 2.0 * k + alpha;
        double a1 = 2.0 * k * (k + alpha) * (c - 2.0);
        double a2 = (c - 1.0) * alpha * alpha;
        double a3 = (c - 2.0) * (c - 1.0) * c;
        double a4 = 2.0 * (k - 1.0 + alpha) * (k - 1.0) * c;
        D2 p2 = (1.0 / a1) * ((cst(a2) + a3 * x) * p1 - a4 * p0);
        p0 = p1;
        p1 = p2;
    }
    return p1;
}

// (degree-in-xi n, degree-in-eta m) of basis function j (1-based): within total degree d the
// functions are ordered n = d..0 (src/basis.jl:211-218).
void dubiner_degrees(int j, int& n, int& m) {
    int d = 0;
    while ((d + 1) * (d + 2) / 2 < j) ++d;
    n = (d + 1) * (d + 2) / 2 - j;
    m = d - n;
}

// dense helpers, row-major
std::vector<double> matmul(const std::vector<double>& A, const std::vector<double>& B, int p, int q, int r) {
    std::vector<double> C(size_t(p) * r, 0.0);
    for (int i = 0; i < p; ++i)
        for (int k = 0; k < q; ++k) {
            double a = A[size_t(i) * q + k];
            for (int j = 0; j < r; ++j) C[size_t(i) * r + j] += a * B[size_t(k) * r + j];
        }
    return C;
}
std::vector<double> transpose(const std::vector<double>& A, int p, int q) {
    std::vector<double> At(size_t(p) * q);
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < q; ++j) At[size_t(j) * p + i] = A[size_t(i) * q + j];
    return At;
}
std::vector<double> inverse(const std::vector<double>& A, int n) {
    std::vector<long double> a(size_t(n) * 2 * n, 0.0L);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) a[size_t(i) * 2 * n + j] = A[size_t(i) * n + j];
        a[size_t(i) * 2 * n + n + i] = 1.0L;
    }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int i = c + 1; i < n; ++i)
            if (fabsl(a[size_t(i) * 2 * n + c]) > fabsl(a[size_t(piv) * 2 * n + c])) piv = i;
        if (a[size_t(piv) * 2 * n + c] == 0.0L) throw std::string("reference mass matrix is singular");
        if (piv != c)
            for (int j = 0; j < 2 * n; ++j) std::swap(a[size_t(c) * 2 * n + j], a[size_t(piv) * 2 * n + j]);
        long double d = 1.0L / a[size_t(c) * 2 * n + c];
        for (int j = 0; j < 2 * n; ++j) a[size_t(c) * 2 * n + j] *= d;
        for (int i = 0; i < n; ++i) {
            if (i == c) continue;
            long double f = a[size_t(i) * 2 * n + c];
            if (f == 0.0L) continue;
            for (int j = 0; j < 2 * n; ++j) a[size_t(i) * 2 * n + j] -= f * a[size_t(c) * 2 * n + j];
        }
    }
    std::vector<double> R(size_t(n) * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) R[size_t(i) * n + j] = double(a[size_t(i) * 2 * n + n + j]);
    return R;
}

// Strang-Fix rules, degree 1..6 (the data of src/StrangQuad.jl:1-61), stored as orbits:
// {barycentric pattern id, a, b, weight*2}.  Expanded in the reference's point order.
struct StrangPoint { double r, s, w; };
std::vector<StrangPoint> strang_points(int degree) {
    std::vector<StrangPoint> P;
    auto add = [&](double r, double s, double w) { P.push_back({r, s, 0.5 * w}); };
    switch (degree) {
        case 0:
        case 1:
            add(1.0 / 3.0, 1.0 / 3.0, 1.0);
            break;
        case 2: {
            const double a = 1.0 / 6.0, b = 2.0 / 3.0, w = 1.0 / 3.0;
            add(a, a, w); add(a, b, w); add(b, a, w);
            break;
        }
        case 3: {
            const double a = 0.659027622374092, b = 0.231933368553031, c = 0.109039009072877, w = 1.0 / 6.0;
            add(a, b, w); add(a, c, w); add(b, a, w); add(b, c, w); add(c, a, w); add(c, b, w);
            break;
        }
        case 4: {
            const double a = 0.816847572980459, b = 0.091576213509771, w1 = 0.109951743655322;
            const double c = 0.108103018168070, d = 0.445948490915965, w2 = 0.223381589678011;
            add(a, b, w1); add(b, a, w1); add(b, b, w1);
            add(c, d, w2); add(d, c, w2); add(d, d, w2);
            break;
        }
        case 5: {
            const double t = 0.33333333333333333, w0 = 0.22500000000000000;
            const double a = 0.79742698535308720, b = 0.10128650732345633, w1 = 0.12593918054482717;
            const double c = 0.05971587178976981, d = 0.47014206410511505, w2 = 0.13239415278850616;
            add(t, t, w0);
            add(a, b, w1); add(b, a, w1); add(b, b, w1);
            add(c, d, w2); add(d, c, w2); add(d, d, w2);
            break;
        }
        case 6: {
            const double a = 0.873821971016996, b = 0.063089014491502, w1 = 0.050844906370207;
            const double c = 0.501426509658179, d = 0.249286745170910, w2 = 0.116786275726379;
            const double e = 0.636502499121399, f = 0.310352451033785, g = 0.053145049844816,
                         w3 = 0.082851075618374;
            add(a, b, w1); add(b, a, w1); add(b, b, w1);
            add(c, d, w2); add(d, c, w2); add(d, d, w2);
            add(e, f, w3); add(e, g, w3); add(f, e, w3); add(f, g, w3); add(g, e, w3); add(g, f, w3);
            break;
        }
        default:
            throw std::string("Strang rule of order " + std::to_string(degree) + " not available");
    }
    return P;
}

double factorial(int k) {
    double f = 1.0;
    for (int i = 2; i <= k; ++i) f *= i;
    return f;
}

// Grundmann-Moeller rule of index s on the triangle (src/GrundmannMoellerQuad.jl:10-27):
// degree 2s+1, binomial(s+3,s) points, alternating-sign weights, normalised to sum 1/2.
void grundmann_moeller_2d(int s, std::vector<double>& pts, std::vector<double>& w) {
    const int dim = 2, d = 2 * s + 1;
    pts.clear();
    w.clear();
    for (int i = 0; i <= s; ++i) {
        double den = double(d + dim - 2 * i);
        double wi = ((i % 2) ? -1.0 : 1.0) * std::pow(2.0, -2 * s) * std::pow(den, d) /
                    (factorial(i) * factorial(d + dim - i));
        int tot = s - i;
        // compositions of tot into 3 parts, last part slowest (the reference's enumeration order)
        for (int c = 0; c <= tot; ++c)
            for (int b = 0; b <= tot - c; ++b) {
                pts.push_back((2.0 * b + 1.0) / den);
                pts.push_back((2.0 * c + 1.0) / den);
                w.push_back(wi);
            }
    }
    double sum = 0.0;
    for (double x : w) sum += 2.0 * x;
    for (double& x : w) x /= sum;
}

}  // namespace

void dubiner_eval(int j, double r, double s, double* val, double* dr, double* ds) {
    int n, m;
    dubiner_degrees(j, n, m);
    D2 a = {2.0 * r + s - 1.0, 2.0, 1.0};
    D2 b = {1.0 - s, 0.0, -1.0};
    D2 eta = {2.0 * s - 1.0, 0.0, 2.0};
    double scale = 2.0 * std::sqrt((2.0 * n + 1.0) * (m + n + 1.0) / 2.0);
    D2 phi = scale * (homog_legendre(n, a, b) * jacobi_a0(m, 2.0 * n + 1.0, eta));
    if (val) *val = phi.v;
    if (dr) *dr = phi.r;
    if (ds) *ds = phi.s;
}

double legendre01_eval(int k, double x) {
    // sqrt(2(k-1)+1) P_{k-1}(2x-1), src/basis.jl:351-354
    int deg = k - 1;
    double y = 2.0 * x - 1.0, p0 = 1.0, p1 = y;
    if (deg == 0) return 1.0;
    for (int i = 1; i < deg; ++i) {
        double p2 = ((2.0 * i + 1.0) * y * p1 - i * p0) / (i + 1.0);
        p0 = p1;
        p1 = p2;
    }
    return std::sqrt(2.0 * deg + 1.0) * p1;
}

void cell_rule(int degree, std::vector<double>& pts, std::vector<double>& w) {
    // DefaultQuad dispatch, src/quadrature.jl:17-26
    if (degree <= 6) {
        auto P = strang_points(degree);
        pts.clear();
        w.clear();
        for (auto& p : P) {
            pts.push_back(p.r);
            pts.push_back(p.s);
            w.push_back(p.w);
        }
    } else if (degree % 2 == 1) {
        grundmann_moeller_2d((degree - 1) / 2, pts, w);
    } else {
        throw std::string("Quadrature rule of order " + std::to_string(degree) + " not available");
    }
}

void gauss_legendre01(int npts, std::vector<double>& x, std::vector<double>& w) {
    // npts-point Gauss-Legendre mapped to (0,1), ascending (src/quadrature.jl:29-39).
    x.assign(npts, 0.0);
    w.assign(npts, 0.0);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < (npts + 1) / 2; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (npts + 0.5L)), dp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1.0L, p1 = z;
            for (int k = 1; k < npts; ++k) {
                long double p2 = ((2.0L * k + 1.0L) * z * p1 - k * p0) / (k + 1.0L);
                p0 = p1;
                p1 = p2;
            }
            dp = npts * (z * p1 - p0) / (z * z - 1.0L);
            long double dz = p1 / dp;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        {   // derivative at the converged node for the weight
            long double p0 = 1.0L, p1 = z;
            for (int k = 1; k < npts; ++k) {
                long double p2 = ((2.0L * k + 1.0L) * z * p1 - k * p0) / (k + 1.0L);
                p0 = p1;
                p1 = p2;
            }
            dp = npts * (z * p1 - p0) / (z * z - 1.0L);
        }
        long double wt = 2.0L / ((1.0L - z * z) * dp * dp);
        // z is the i-th largest root; ascending order puts -z first
        x[i] = double((1.0L - z) / 2.0L);
        x[npts - 1 - i] = double((1.0L + z) / 2.0L);
        w[i] = w[npts - 1 - i] = double(wt / 2.0L);
    }
    if (npts % 2 == 1) x[npts / 2] = 0.5;
}

RefTables build_ref_tables(int order, int quad_degree) {
    if (order < 1 || order > MAX_ORDER)
        throw std::string("order must be 1.." + std::to_string(MAX_ORDER) + " (Dubiner closed forms cover order <= 4)");
    RefTables R;
    R.order = order;
    R.quad_degree = quad_degree > 0 ? quad_degree : order + 1;  // src/ScalarFunctionSpaces.jl:24-25
    const int n = (order + 1) * (order + 2) / 2, nt = order + 1, t = 3 * nt;
    R.n = n; R.nt = nt; R.m = 3 * n; R.t = t;
    cell_rule(R.quad_degree, R.qpts, R.qw);
    gauss_legendre01(R.quad_degree, R.fpts, R.fw);
    const int nq = int(R.qw.size()), nfq = int(R.fw.size());
    R.nq = nq; R.nfq = nfq;

    R.N.assign(size_t(n) * nq, 0.0);
    R.dN.assign(size_t(n) * nq * 2, 0.0);
    R.Mgeo.assign(size_t(nq) * 3, 0.0);
    R.WN.assign(size_t(nq) * n, 0.0);
    for (int q = 0; q < nq; ++q) {
        double r = R.qpts[2 * q], s = R.qpts[2 * q + 1];
        for (int i = 0; i < n; ++i) {
            double v, dr, ds;
            dubiner_eval(i + 1, r, s, &v, &dr, &ds);
            R.N[i + size_t(n) * q] = v;
            R.dN[(i + size_t(n) * q) * 2 + 0] = dr;
            R.dN[(i + size_t(n) * q) * 2 + 1] = ds;
            R.WN[size_t(q) * n + i] = R.qw[q] * v;
        }
        R.Mgeo[3 * q + 0] = 1.0 - r - s;
        R.Mgeo[3 * q + 1] = r;
        R.Mgeo[3 * q + 2] = s;
    }
    // reference edges: edge 1 (1,0)->(0,1), edge 2 (0,1)->(0,0), edge 3 (0,0)->(1,0)  src/shapes.jl:19-23
    const double e1[3][2] = {{1, 0}, {0, 1}, {0, 0}}, e2[3][2] = {{0, 1}, {0, 0}, {1, 0}};
    R.E.assign(size_t(n) * nfq * 3, 0.0);
    for (int l = 0; l < 3; ++l)
        for (int p = 0; p < nfq; ++p) {
            double sp = R.fpts[p];
            double r = (1.0 - sp) * e1[l][0] + sp * e2[l][0];
            double s = (1.0 - sp) * e1[l][1] + sp * e2[l][1];
            for (int i = 0; i < n; ++i) {
                double v;
                dubiner_eval(i + 1, r, s, &v, nullptr, nullptr);
                R.E[i + size_t(n) * (p + size_t(nfq) * l)] = v;
            }
        }
    R.T.assign(size_t(nt) * nfq, 0.0);
    for (int p = 0; p < nfq; ++p)
        for (int j = 0; j < nt; ++j) R.T[j + size_t(nt) * p] = legendre01_eval(j + 1, R.fpts[p]);

    // ---- reference matrices -------------------------------------------------------------
    std::vector<double> Br(size_t(n) * n, 0.0), Bs(size_t(n) * n, 0.0);
    R.Mhat.assign(size_t(n) * n, 0.0);
    for (int q = 0; q < nq; ++q)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double Nj = R.N[j + size_t(n) * q];
                R.Mhat[size_t(i) * n + j] += R.qw[q] * R.N[i + size_t(n) * q] * Nj;
                Br[size_t(i) * n + j] += R.qw[q] * R.dN[(i + size_t(n) * q) * 2 + 0] * Nj;
                Bs[size_t(i) * n + j] += R.qw[q] * R.dN[(i + size_t(n) * q) * 2 + 1] * Nj;
            }
    R.Minv = inverse(R.Mhat, n);
    R.Tr = matmul(R.Minv, Br, n, n, n);
    R.Ts = matmul(R.Minv, Bs, n, n, n);
    auto BrT = transpose(Br, n, n), BsT = transpose(Bs, n, n);
    R.Prr = matmul(BrT, R.Tr, n, n, n);
    R.Pss = matmul(BsT, R.Ts, n, n, n);
    auto P1 = matmul(BrT, R.Ts, n, n, n), P2 = matmul(BsT, R.Tr, n, n, n);
    R.Prs.assign(size_t(n) * n, 0.0);
    for (size_t i = 0; i < R.Prs.size(); ++i) R.Prs[i] = P1[i] + P2[i];

    R.Chat.assign(size_t(3) * n * n, 0.0);
    R.Fhat.assign(size_t(n) * t, 0.0);
    for (int l = 0; l < 3; ++l)
        for (int p = 0; p < nfq; ++p)
            for (int i = 0; i < n; ++i) {
                double Ei = R.E[i + size_t(n) * (p + size_t(nfq) * l)];
                for (int j = 0; j < n; ++j)
                    R.Chat[(size_t(l) * n + i) * n + j] += R.fw[p] * Ei * R.E[j + size_t(n) * (p + size_t(nfq) * l)];
                for (int j = 0; j < nt; ++j)
                    R.Fhat[size_t(i) * t + l * nt + j] += R.fw[p] * Ei * R.T[j + size_t(nt) * p];
            }
    R.MF = matmul(R.Minv, R.Fhat, n, n, t);
    R.Qr = matmul(BrT, R.MF, n, n, t);
    R.Qs = matmul(BsT, R.MF, n, n, t);
    R.Hhat.assign(size_t(nt) * nt, 0.0);
    for (int p = 0; p < nfq; ++p)
        for (int i = 0; i < nt; ++i)
            for (int j = 0; j < nt; ++j)
                R.Hhat[size_t(i) * nt + j] += R.fw[p] * R.T[i + size_t(nt) * p] * R.T[j + size_t(nt) * p];
    return R;
}

}  // namespace hdg
