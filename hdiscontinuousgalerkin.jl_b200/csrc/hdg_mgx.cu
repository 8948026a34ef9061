// General-mesh vertex-space term of the multigrid preconditioner (hdg_set_preconditioner(ctx, 2) on meshes WITHOUT the grid
// structure of rectangle_mesh: parse_mesh_triangle input, permuted node ids, Delaunay meshes - the reference solves any mesh
// with K \ b, examples/poisson2D_HDG.jl:195).  Hierarchy-free:  z += P C_m(A_c) P' r  with A_c = P'AP in ELL form on the
// vertex graph and C_m = m steps of the Jacobi-scaled Chebyshev iteration (a fixed polynomial, so plain CG still applies).
// Kernel bodies: hdg_mg_general.cuh (also compiled for the host and checked against scipy by tools/check_mg_general.py).
// One GPU.
#include <cstring>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"
#include "hdg_mg_general.cuh"

namespace hdg {

__global__ void __launch_bounds__(RB) mgx_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ part) {
    double s = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * RB + threadIdx.x; i < n; i += int64_t(gridDim.x) * RB) s = fma(a[i], b[i], s);
    const double tot = block_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

constexpr int MGX_CHEB_STEPS = 16;          // tools/cheb_prototype.py: m = 16, alpha = 100
constexpr double MGX_ALPHA = 100.0;

struct MgGeneral {
    int64_t nnode = 0, nface = 0;
    int32_t *vcnt = nullptr, *vface = nullptr, *nbr = nullptr;
    double *val = nullptr, *diag = nullptr, *dinv = nullptr, *rc = nullptr, *x = nullptr, *res = nullptr, *d = nullptr, *fc = nullptr;
    unsigned long long* lmax_bits = nullptr;
    double lmax = 0.0;
    bool adjacency_ok = false;
};

__global__ void mgx_adj_fill(const int32_t* __restrict__ facenode, int64_t nface, int32_t* __restrict__ vcnt, int32_t* __restrict__ vface,
                             int32_t* __restrict__ flags) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    const int64_t v1 = facenode[2 * f], v2 = facenode[2 * f + 1];
    const int64_t lo = min(v1, v2), hi = max(v1, v2);
    int k = atomicAdd(&vcnt[lo], 1);
    if (k < MGX_MAXVAL) vface[lo * MGX_MAXVAL + k] = int32_t(f); else atomicExch(&flags[FLAG_MG], 1);
    k = atomicAdd(&vcnt[hi], 1);
    if (k < MGX_MAXVAL) vface[hi * MGX_MAXVAL + k] = int32_t(uint32_t(f) | 0x80000000u); else atomicExch(&flags[FLAG_MG], 1);
}
__global__ void mgx_adj_sort(int64_t nnode, int32_t* __restrict__ vcnt, int32_t* __restrict__ vface, const uint8_t* __restrict__ isbc) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    const int cnt = min(vcnt[v], MGX_MAXVAL);
    int32_t a[MGX_MAXVAL];
    bool fixed = false;
    for (int k = 0; k < cnt; ++k) {
        a[k] = vface[v * MGX_MAXVAL + k];
        fixed = fixed || isbc[a[k] & 0x7fffffff];
    }
    for (int i = 1; i < cnt; ++i) {
        const int32_t key = a[i];
        int j = i - 1;
        while (j >= 0 && (a[j] & 0x7fffffff) > (key & 0x7fffffff)) { a[j + 1] = a[j]; --j; }
        a[j + 1] = key;
    }
    for (int k = 0; k < cnt; ++k) vface[v * MGX_MAXVAL + k] = a[k];
    vcnt[v] = (fixed || cnt == 0) ? -1 : cnt;
}
// rc = P'r, res = rc, x = 0, d = Dinv res / theta
__global__ void mgx_cheb_start(int64_t n, int NT, const int32_t* __restrict__ vcnt, const int32_t* __restrict__ vface, const double* __restrict__ r,
                               const double* __restrict__ dinv, double inv_theta, double* __restrict__ rc, double* __restrict__ res,
                               double* __restrict__ x, double* __restrict__ d) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const double s = mgx_restrict_row(v, NT, vcnt, vface, r);
    rc[v] = s; res[v] = s; x[v] = 0.0;
    d[v] = dinv[v] * s * inv_theta;
}

void mgx_free(hdg_context* c) {
    MgGeneral* g = static_cast<MgGeneral*>(c->mg_general);
    if (!g) return;
    void* ptrs[] = {g->vcnt, g->vface, g->nbr, g->val, g->diag, g->dinv, g->rc, g->x, g->res, g->d, g->lmax_bits};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete g;
    c->mg_general = nullptr;
}

hdg_status mgx_setup(hdg_context* c) {
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "the general-mesh multigrid term runs on one GPU");
    const int NT = c->tab.nt;
    MgGeneral* g = static_cast<MgGeneral*>(c->mg_general);
    if (g && (g->nnode != c->nnode || g->nface != c->nface)) { mgx_free(c); g = nullptr; }
    const int64_t n = c->nnode;
    if (!g) {
        g = new MgGeneral();
        c->mg_general = g;
        g->nnode = n; g->nface = c->nface;
        HDG_CUDA(c, cudaMalloc(&g->vcnt, sizeof(int32_t) * n));
        HDG_CUDA(c, cudaMalloc(&g->vface, sizeof(int32_t) * n * MGX_MAXVAL));
        HDG_CUDA(c, cudaMalloc(&g->nbr, sizeof(int32_t) * n * MGX_MAXVAL));
        HDG_CUDA(c, cudaMalloc(&g->val, sizeof(double) * n * MGX_MAXVAL));
        HDG_CUDA(c, cudaMalloc(&g->diag, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->dinv, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->rc, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->x, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->res, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->d, sizeof(double) * n));
        HDG_CUDA(c, cudaMalloc(&g->lmax_bits, sizeof(unsigned long long)));
    }
    const unsigned nb = (unsigned)ceil_div(n, 256);
    if (!g->adjacency_ok) {
        HDG_CUDA(c, cudaMemsetAsync(g->vcnt, 0, sizeof(int32_t) * n, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_MG, 0, sizeof(int32_t), c->stream));
        mgx_adj_fill<<<(unsigned)ceil_div(c->nface, 256), 256, 0, c->stream>>>(c->d_facenode, c->nface, g->vcnt, g->vface, c->d_flags);
        mgx_adj_sort<<<nb, 256, 0, c->stream>>>(n, g->vcnt, g->vface, c->d_isbc);
        mgx_neighbours<<<nb, 256, 0, c->stream>>>(n, g->vcnt, g->vface, c->d_facenode, g->nbr);
        c->launches += 3;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->h_flags[FLAG_MG]) return set_err(c, HDG_ERR_INVALID, "a vertex has more than 16 faces");
        g->adjacency_ok = true;
    }
    // operator (every solve: the matrix may have changed), inverse diagonal, Gershgorin bound of D^-1 A_c
    mgx_operator<<<nb, 256, 0, c->stream>>>(n, NT, c->d_Kd, c->d_Ko, c->d_kcol, c->d_isbc, c->d_facenode, g->vcnt, g->vface, g->nbr, g->diag, g->val);
    HDG_CUDA(c, cudaMemsetAsync(g->lmax_bits, 0, sizeof(unsigned long long), c->stream));
    mgx_dinv<<<nb, 256, 0, c->stream>>>(n, g->vcnt, g->diag, g->val, g->dinv, g->lmax_bits);
    c->launches += 2;
    unsigned long long bits = 0;
    HDG_CUDA(c, cudaMemcpyAsync(&bits, g->lmax_bits, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    std::memcpy(&g->lmax, &bits, sizeof(double));
    if (!(g->lmax > 0.0)) g->lmax = 2.0;      // no free vertex at all: the term vanishes anyway
    return HDG_OK;
}

hdg_status mgx_apply(hdg_context* c, const double* r, double* z, double* part, int np) {
    MgGeneral* g = static_cast<MgGeneral*>(c->mg_general);
    cudaStream_t s = c->stream;
    const int NT = c->tab.nt;
    const int64_t n = g->nnode;
    const unsigned nb = (unsigned)ceil_div(n, 256);
    const double lmax = g->lmax, lmin = lmax / MGX_ALPHA;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    mgx_cheb_start<<<nb, 256, 0, s>>>(n, NT, g->vcnt, g->vface, r, g->dinv, 1.0 / theta, g->rc, g->res, g->x, g->d);
    for (int i = 0; i < MGX_CHEB_STEPS; ++i) {
        mgx_cheb_residual<<<nb, 256, 0, s>>>(n, g->vcnt, g->nbr, g->diag, g->val, g->d, g->x, g->res);
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        mgx_cheb_direction<<<nb, 256, 0, s>>>(n, g->dinv, g->res, rho_new * rho, 2.0 * rho_new / delta, g->d);
        rho = rho_new;
    }
    mgx_dot<<<np, RB, 0, s>>>(g->rc, g->x, n, part);
    mgx_prolong<<<(unsigned)ceil_div(c->nface_own, 256), 256, 0, s>>>(c->nface_own, NT, c->d_facenode, c->d_isbc, g->x, z);
    return HDG_OK;
}

bool mgx_active(const hdg_context* c) { return c->mg_general != nullptr; }
void mgx_invalidate(hdg_context* c) {      // same-size mesh change: keep the buffers, rebuild the adjacency
    if (c->mg_general) static_cast<MgGeneral*>(c->mg_general)->adjacency_ok = false;
}
int mgx_launches_per_apply() { return 3 + 2 * MGX_CHEB_STEPS; }

}  // namespace hdg
