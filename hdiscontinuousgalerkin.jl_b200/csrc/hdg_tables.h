// Reference-element tables of the HDG path, built on the host once per context.
//
// Replaces the table builders of the reference: _scalar_fs (src/ScalarFunctionSpaces.jl:31-99),
// VectorFunctionSpace (src/VectorFunctionSpaces.jl:10-18), ScalarTraceFunctionSpace
// (src/TraceFunctionSpaces.jl:11-28), the quadrature rules (src/quadrature.jl:17-75,
// src/StrangQuad.jl, src/GrundmannMoellerQuad.jl) and the bases (src/basis.jl:65-232,351-354).
// On top of the raw tables it forms the "reference matrices": on an affine triangle every local
// block of examples/poisson2D_HDG.jl:88-153 is a geometry scalar times one of these
// (SURVEY.md Appendix A), which is what the device kernels consume.
#pragma once
#include <string>
#include <vector>

namespace hdg {

constexpr int MAX_ORDER = 4;

struct RefTables {
    int order = 0, quad_degree = 0;
    int n = 0, nt = 0, m = 0, t = 0, nq = 0, nfq = 0;
    // raw tables (layouts documented at hdg_get_table in include/hdg_b200.h)
    std::vector<double> qpts;   // nq*2
    std::vector<double> qw;     // nq
    std::vector<double> fpts;   // nfq   GL points on (0,1), ascending
    std::vector<double> fw;     // nfq
    std::vector<double> N;      // n*nq     N[i + n*q]
    std::vector<double> dN;     // n*nq*2   dN[(i + n*q)*2 + a]
    std::vector<double> E;      // n*nfq*3  E[i + n*(p + nfq*l)]  (un-reversed face points)
    std::vector<double> T;      // nt*nfq   T[j + nt*p]
    // reference matrices, row-major [i*cols + j]
    std::vector<double> Mhat;   // n*n   sum_q w N_i N_j
    std::vector<double> Minv;   // n*n
    std::vector<double> Tr, Ts; // n*n   Minv * Br, Minv * Bs with Br[i,j] = sum_q w dN_r[i] N[j]
    std::vector<double> Prr, Prs, Pss;  // n*n  Br'Tr, Br'Ts + Bs'Tr, Bs'Ts
    std::vector<double> Chat;   // 3*n*n  Chat[l][i][j] = sum_p w_p E_l[i,p] E_l[j,p]
    std::vector<double> Fhat;   // n*t    Fhat[i][l*nt+j] = sum_p w_p E_l[i,p] T[j,p]   (orientation true)
    std::vector<double> MF;     // n*t    Minv * Fhat
    std::vector<double> Qr, Qs; // n*t    Br' * MF, Bs' * MF
    std::vector<double> Hhat;   // nt*nt  sum_p w_p T_i T_j
    std::vector<double> WN;     // nq*n   WN[q][i] = w_q N[i,q]
    std::vector<double> Mgeo;   // nq*3   P1 geometry map at the cell points (1-r-s, r, s)
};

// Throws std::string on an unavailable rule (maps to HDG_ERR_UNSUPPORTED_RULE) or bad order.
RefTables build_ref_tables(int order, int quad_degree);

// individual pieces, exposed for the C-ABI table getter and for tests
void dubiner_eval(int j /*1-based*/, double r, double s, double* val, double* dr, double* ds);
double legendre01_eval(int k /*1-based*/, double x);
void cell_rule(int degree, std::vector<double>& pts, std::vector<double>& w);
void gauss_legendre01(int npts, std::vector<double>& x, std::vector<double>& w);

}  // namespace hdg
