// Vertex-space term of the multigrid preconditioner for meshes WITHOUT grid structure, hierarchy-free variant
// (tools/cheb_prototype.py):   z += P C_m(A_c) P' r,   A_c = P'AP in ELL form on the vertex graph,
// C_m = m steps of the Jacobi-scaled Chebyshev iteration on [lmax/alpha, lmax], lmax from the Gershgorin bound.
//
// Every kernel is a gather over one index (vertex or face) with no shared memory and no atomics except the max-reduction,
// and its body is a __host__ __device__ function of that index: tools/check_mg_general.py compiles this header for the
// HOST (-DHDG_HOST_EMU), runs the bodies in a serial loop and checks them against scipy on jittered and Delaunay meshes.
// Inputs are the arrays the library already holds: Kd/Ko/kcol (block-row trace matrix), isbc, facenode, and the
// vertex -> face adjacency vcnt/vface built by mg_adj_fill / mg_adj_sort / mg_adj_fix (with MGX_MAXVAL slots).
#pragma once
#include <cstdint>
#include <cmath>

#ifdef HDG_HOST_EMU
#define MGX_HD
#else
#define MGX_HD __host__ __device__ __forceinline__
#endif

namespace hdg {

constexpr int MGX_MAXVAL = 16;                              // faces per vertex on a general triangle mesh
constexpr double MGX_C1 = 0.28867513459481287;              // 1 / (2 sqrt 3)

// neighbour vertex of slot k = the other endpoint of the k-th incident face (bit 31 of vface: v is the face's hi vertex)
MGX_HD void mgx_neighbours_row(int64_t v, const int32_t* vcnt, const int32_t* vface, const int32_t* facenode, int32_t* nbr) {
    const int cnt = vcnt[v];
    for (int k = 0; k < MGX_MAXVAL; ++k) {
        int32_t u = -1;
        if (k < cnt) {
            const int64_t f = vface[v * MGX_MAXVAL + k] & 0x7fffffff;
            const int32_t a = facenode[2 * f], b = facenode[2 * f + 1];
            u = (a == int32_t(v)) ? b : a;
        }
        nbr[v * MGX_MAXVAL + k] = u;
    }
}

// row v of A_c = P'AP (A = -K on free rows): diag[v], val[v][k] for neighbour slot k.  Fixed vertices (vcnt < 0): zero row.
MGX_HD void mgx_operator_row(int64_t v, int NT, const double* Kd, const double* Ko, const int32_t* kcol, const uint8_t* isbc,
                             const int32_t* facenode, const int32_t* vcnt, const int32_t* vface, const int32_t* nbr,
                             double* diag, double* val) {
    const int NT2 = NT * NT;
    double acc[MGX_MAXVAL], d = 0.0;
    for (int k = 0; k < MGX_MAXVAL; ++k) acc[k] = 0.0;
    const int cnt = vcnt[v];
    for (int k = 0; k < cnt; ++k) {
        const int32_t e = vface[v * MGX_MAXVAL + k];
        const int64_t f = e & 0x7fffffff;
        const double pf0 = 0.5, pf1 = (e < 0) ? MGX_C1 : -MGX_C1;
        for (int s = -1; s < 4; ++s) {
            const int64_t g = s < 0 ? f : kcol[4 * f + s];
            if (g < 0 || isbc[g]) continue;
            const double* blk = s < 0 ? Kd + f * NT2 : Ko + (f * 4 + s) * NT2;     // column-major: blk[b*NT + a] = K[a][b]
            const double t0 = -(pf0 * blk[0] + pf1 * blk[1]);
            const double t1 = -(pf0 * blk[NT] + pf1 * blk[NT + 1]);
            const int32_t g1 = facenode[2 * g], g2 = facenode[2 * g + 1];
            const int32_t lo = g1 < g2 ? g1 : g2, hi = g1 < g2 ? g2 : g1;
            const double wlo = 0.5 * t0 - MGX_C1 * t1, whi = 0.5 * t0 + MGX_C1 * t1;
            for (int side = 0; side < 2; ++side) {
                const int32_t w = side ? hi : lo;
                const double a = side ? whi : wlo;
                if (vcnt[w] < 0) continue;                       // fixed vertex: no unknown
                if (w == int32_t(v)) { d += a; continue; }
                for (int q = 0; q < cnt; ++q)
                    if (nbr[v * MGX_MAXVAL + q] == w) { acc[q] += a; break; }
            }
        }
    }
    diag[v] = cnt < 0 ? 0.0 : d;
    for (int k = 0; k < MGX_MAXVAL; ++k) val[v * MGX_MAXVAL + k] = (cnt < 0 || k >= cnt) ? 0.0 : acc[k];
}

// inverse diagonal (0 at fixed vertices / non-positive diagonals) and the Gershgorin bound of row v of D^-1 A_c
MGX_HD double mgx_dinv_row(int64_t v, const int32_t* vcnt, const double* diag, const double* val, double* dinv) {
    const double d = diag[v];
    if (vcnt[v] < 0 || !(d > 0.0)) { dinv[v] = 0.0; return 0.0; }
    dinv[v] = 1.0 / d;
    double s = d;
    for (int k = 0; k < vcnt[v]; ++k) s += fabs(val[v * MGX_MAXVAL + k]);
    return s / d;
}

MGX_HD double mgx_apply_row(int64_t v, const int32_t* vcnt, const int32_t* nbr, const double* diag, const double* val, const double* x) {
    double s = diag[v] * x[v];
    const int cnt = vcnt[v];
    for (int k = 0; k < cnt; ++k) s = fma(val[v * MGX_MAXVAL + k], x[nbr[v * MGX_MAXVAL + k]], s);
    return s;
}

// P' r at vertex v (r: trace vector, NT entries per face)
MGX_HD double mgx_restrict_row(int64_t v, int NT, const int32_t* vcnt, const int32_t* vface, const double* r) {
    double s = 0.0;
    const int cnt = vcnt[v];
    for (int k = 0; k < cnt; ++k) {
        const int32_t e = vface[v * MGX_MAXVAL + k];
        const int64_t f = e & 0x7fffffff;
        s += 0.5 * r[f * NT] + ((e < 0) ? MGX_C1 : -MGX_C1) * r[f * NT + 1];
    }
    return s;
}

// z += P e on face f
MGX_HD void mgx_prolong_face(int64_t f, int NT, const int32_t* facenode, const uint8_t* isbc, const double* e, double* z) {
    if (isbc[f]) return;
    const int32_t v1 = facenode[2 * f], v2 = facenode[2 * f + 1];
    const double a = e[v1 < v2 ? v1 : v2], b = e[v1 < v2 ? v2 : v1];
    z[f * NT] += 0.5 * (a + b);
    z[f * NT + 1] += MGX_C1 * (b - a);
}

// Chebyshev iteration for A_c x = r, zero start, interval [lmin, lmax] of D^-1 A_c (Saad, Alg. 12.1 with Jacobi scaling):
//   theta = (lmax + lmin)/2, delta = (lmax - lmin)/2, sigma = theta/delta, rho_0 = 1/sigma, d_0 = Dinv r / theta,
//   step i:  x += d;  res -= A d;  rho' = 1/(2 sigma - rho);  d = rho' rho d + (2 rho'/delta) Dinv res;  rho = rho'
// One step = the two point-wise kernels below (res needs the neighbours of d, hence two kernels).
MGX_HD void mgx_cheb_residual_row(int64_t v, const int32_t* vcnt, const int32_t* nbr, const double* diag, const double* val,
                                  const double* d, double* x, double* res) {
    x[v] += d[v];
    res[v] -= mgx_apply_row(v, vcnt, nbr, diag, val, d);
}
MGX_HD void mgx_cheb_direction_row(int64_t v, const double* dinv, const double* res, double c_old, double c_new, double* d) {
    d[v] = c_old * d[v] + c_new * dinv[v] * res[v];
}

#ifndef HDG_HOST_EMU
// ---- device wrappers (one thread per index) ------------------------------------------------------------------------------
__global__ void mgx_neighbours(int64_t n, const int32_t* vcnt, const int32_t* vface, const int32_t* facenode, int32_t* nbr) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < n) mgx_neighbours_row(v, vcnt, vface, facenode, nbr);
}
__global__ void mgx_operator(int64_t n, int NT, const double* Kd, const double* Ko, const int32_t* kcol, const uint8_t* isbc,
                             const int32_t* facenode, const int32_t* vcnt, const int32_t* vface, const int32_t* nbr, double* diag, double* val) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < n) mgx_operator_row(v, NT, Kd, Ko, kcol, isbc, facenode, vcnt, vface, nbr, diag, val);
}
__global__ void mgx_dinv(int64_t n, const int32_t* vcnt, const double* diag, const double* val, double* dinv, unsigned long long* lmax_bits) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const double g = mgx_dinv_row(v, vcnt, diag, val, dinv);
    if (g > 0.0) atomicMax(lmax_bits, (unsigned long long)__double_as_longlong(g));      // positive doubles order like their bit patterns
}
__global__ void mgx_restrict(int64_t n, int NT, const int32_t* vcnt, const int32_t* vface, const double* r, double* rc) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < n) rc[v] = mgx_restrict_row(v, NT, vcnt, vface, r);
}
__global__ void mgx_prolong(int64_t nface, int NT, const int32_t* facenode, const uint8_t* isbc, const double* e, double* z) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f < nface) mgx_prolong_face(f, NT, facenode, isbc, e, z);
}
__global__ void mgx_cheb_residual(int64_t n, const int32_t* vcnt, const int32_t* nbr, const double* diag, const double* val,
                                  const double* d, double* x, double* res) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < n) mgx_cheb_residual_row(v, vcnt, nbr, diag, val, d, x, res);
}
__global__ void mgx_cheb_direction(int64_t n, const double* dinv, const double* res, double c_old, double c_new, double* d) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < n) mgx_cheb_direction_row(v, dinv, res, c_old, c_new, d);
}
#endif

}  // namespace hdg
