// K5 + K6: element-wise recovery of (sigma, u) from the trace solution, and the squared L2 error.
//
// Replaces get_uσ! (examples/poisson2D_HDG.jl:197-212, with the hard-coded nt = 2 slice at
// :205-206 generalised to nt) and errornorm (src/DiscreteFunctions.jl:97-120).
// Both are streaming kernels, one thread per element: [K_e | b_e] is stored in tiles of 32 cells
// (entry-major inside a tile) so every load of a warp is one contiguous 256-byte segment, and the
// TrialFunction.m_values outputs are column-major ncell x nb (src/DiscreteFunctions.jl:38-54),
// i.e. unit stride across the cells of a warp.
#include <algorithm>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"

namespace hdg {

template <int K>
__global__ void __launch_bounds__(128) recover_kernel(const double* __restrict__ Ke, const double* __restrict__ x,
                                                     const int32_t* __restrict__ cellinfo, int64_t ncell,
                                                     double* __restrict__ sigma, double* __restrict__ u,
                                                     double* __restrict__ uhat_h) {
    constexpr int n = Ord<K>::n, nt = Ord<K>::nt, t = Ord<K>::t, m = Ord<K>::m, ke = Ord<K>::ke;
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int4* ci = reinterpret_cast<const int4*>(cellinfo + CI * c);
    int4 p0 = __ldg(ci), p1 = __ldg(ci + 1);
    const int64_t f[3] = {int64_t(uint32_t(p0.w) & 0x7fffffffu), int64_t(uint32_t(p1.x) & 0x7fffffffu),
                          int64_t(uint32_t(p1.y) & 0x7fffffffu)};
    double ue[t];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int j = 0; j < nt; ++j) {
            ue[l * nt + j] = x[f[l] * nt + j];
            uhat_h[c + ncell * (j + nt * l)] = ue[l * nt + j];   // m_values[cell, :, k]
        }
    const double* __restrict__ tile = Ke + ((c >> 5) * ke) * 32 + (c & 31);
#pragma unroll
    for (int i = 0; i < m; ++i) {
        double d = tile[int64_t(i * (t + 1) + t) * 32];   // b_e[i]
#pragma unroll
        for (int j = 0; j < t; ++j) d = fma(tile[int64_t(i * (t + 1) + j) * 32], ue[j], d);
        if (i < 2 * n) sigma[c + ncell * i] = d;
        else u[c + ncell * (i - 2 * n)] = d;
    }
}

template <int K> static hdg_status recover_t(hdg_context* c) {
    const int B = 128;
    recover_kernel<K><<<(unsigned)ceil_div(c->ncell_own, B), B, 0, c->stream>>>(c->d_Ke, c->d_x, c->d_cellinfo, c->ncell_own,
                                                                           c->d_sigma, c->d_u, c->d_uhat_h);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

hdg_status recover(hdg_context* c) {
    const int n = c->tab.n, nt = c->tab.nt;
    if (!c->d_sigma) {
        HDG_CUDA(c, cudaMalloc(&c->d_sigma, sizeof(double) * c->ncell_own * 2 * n));
        HDG_CUDA(c, cudaMalloc(&c->d_u, sizeof(double) * c->ncell_own * n));
        HDG_CUDA(c, cudaMalloc(&c->d_uhat_h, sizeof(double) * c->ncell_own * nt * 3));
    }
    timer_start(c, c->t_recover);
    hdg_status st = HDG_ERR_INVALID;
    switch (c->tab.order) {
        case 1: st = recover_t<1>(c); break;
        case 2: st = recover_t<2>(c); break;
        case 3: st = recover_t<3>(c); break;
        case 4: st = recover_t<4>(c); break;
    }
    timer_stop(c, c->t_recover);
    if (st) return st;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->recovered = true;
    return HDG_OK;
}

// ---- squared L2 error ----------------------------------------------------------------------------
__global__ void __launch_bounds__(RB) errornorm_kernel(const double* __restrict__ u, const int32_t* __restrict__ cellinfo,
                                                       const double* __restrict__ nodes, RawTablesDev R, int64_t ncell,
                                                       int exact_id, const double* __restrict__ uexq, double* __restrict__ part) {
    const double pi = 3.141592653589793;
    double acc = 0.0;
    for (int64_t c = int64_t(blockIdx.x) * RB + threadIdx.x; c < ncell; c += int64_t(gridDim.x) * RB) {
        const int32_t* ci = cellinfo + CI * c;
        double x[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[k][0] = nodes[2 * int64_t(ci[k])];
            x[k][1] = nodes[2 * int64_t(ci[k]) + 1];
        }
        double detJ = (x[1][0] - x[0][0]) * (x[2][1] - x[0][1]) - (x[2][0] - x[0][0]) * (x[1][1] - x[0][1]);
        double el = 0.0;
        for (int q = 0; q < R.nq; ++q) {
            double uq = 0.0;
            for (int i = 0; i < R.n; ++i) uq += u[c + ncell * i] * R.N[i + R.n * q];
            double xq = R.Mgeo[3 * q] * x[0][0] + R.Mgeo[3 * q + 1] * x[1][0] + R.Mgeo[3 * q + 2] * x[2][0];
            double yq = R.Mgeo[3 * q] * x[0][1] + R.Mgeo[3 * q + 1] * x[1][1] + R.Mgeo[3 * q + 2] * x[2][1];
            // u_ex: built in (poisson2D_HDG.jl:216) or the caller's values at the cell quadrature points
            double ex = exact_id == 1 ? sin(pi * xq) * sin(pi * yq) : (uexq ? uexq[c * R.nq + q] : 0.0);
            double d = uq - ex;
            el += d * d * (detJ * R.qw[q]);
        }
        acc += el;
    }
    double tot = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

hdg_status errornorm(hdg_context* c, int exact_id, const double* uex_host, double* err2) {
    int np = int(std::min<int64_t>(ceil_div(c->ncell_own, RB), 1024));
    double* d_uex = nullptr;
    if (exact_id == 0) {
        const size_t bytes = sizeof(double) * size_t(c->ncell_own) * c->tab.nq;
        HDG_CUDA(c, cudaMalloc(&d_uex, bytes));
        cudaError_t e = cudaMemcpyAsync(d_uex, uex_host, bytes, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { cudaFree(d_uex); return set_err(c, HDG_ERR_CUDA, cudaGetErrorString(e)); }
    }
    timer_start(c, c->t_err);
    errornorm_kernel<<<np, RB, 0, c->stream>>>(c->d_u, c->d_cellinfo, c->d_nodes, c->raw, c->ncell_own, exact_id, d_uex, c->d_partials);
    final_sum<<<1, RB, 0, c->stream>>>(c->d_partials, np, c->d_scal + 3);
    c->launches += 2;
    hdg_status st = comm_allreduce_sum(c, c->d_scal + 3, 1);
    if (st) { if (d_uex) { cudaStreamSynchronize(c->stream); cudaFree(d_uex); } return st; }
    timer_stop(c, c->t_err);
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
    cudaError_t es = cudaStreamSynchronize(c->stream);
    if (d_uex) cudaFree(d_uex);
    if (es != cudaSuccess) return set_err(c, HDG_ERR_CUDA, cudaGetErrorString(es));
    *err2 = c->h_scal[3];
    return HDG_OK;
}

// ---- nodal_avg(u_h) (src/DiscreteFunctions.jl:81-95): vertex values of the discontinuous u_h averaged per node --------
// value(u_h,node,cell) (:70-79) evaluates the cell's expansion at xi = Jinv (x_node - x_1), i.e. at the reference vertices.
__global__ void nodal_accumulate(const double* __restrict__ u, const int32_t* __restrict__ cellinfo, int64_t ncell, int n,
                                 const double* __restrict__ vtab /* n x 3 */, double* __restrict__ sum, int32_t* __restrict__ cnt) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n; ++i) {
        double ui = u[c + ncell * i];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = fma(ui, vtab[3 * i + k], v[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int32_t node = cellinfo[CI * c + k];
        atomicAdd(&sum[node], v[k]);
        atomicAdd(&cnt[node], 1);
    }
}
__global__ void nodal_divide(double* __restrict__ sum, const int32_t* __restrict__ cnt, int64_t nnode) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < nnode) sum[i] = sum[i] / double(cnt[i]);
}

hdg_status nodal_average(hdg_context* c, double* out) {
    const int n = c->tab.n;
    std::vector<double> vt(size_t(n) * 3);
    const double vx[3] = {0.0, 1.0, 0.0}, vy[3] = {0.0, 0.0, 1.0};   // reference vertices, src/shapes.jl:14-18
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) dubiner_eval(i + 1, vx[k], vy[k], &vt[3 * i + k], nullptr, nullptr);
    double *d_vt = nullptr, *d_sum = nullptr;
    int32_t* d_cnt = nullptr;
    HDG_CUDA(c, cudaMalloc(&d_vt, sizeof(double) * vt.size()));
    HDG_CUDA(c, cudaMalloc(&d_sum, sizeof(double) * c->nnode));
    HDG_CUDA(c, cudaMalloc(&d_cnt, sizeof(int32_t) * c->nnode));
    HDG_CUDA(c, cudaMemcpyAsync(d_vt, vt.data(), sizeof(double) * vt.size(), cudaMemcpyHostToDevice, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(d_sum, 0, sizeof(double) * c->nnode, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(d_cnt, 0, sizeof(int32_t) * c->nnode, c->stream));
    nodal_accumulate<<<(unsigned)ceil_div(c->ncell_own, 128), 128, 0, c->stream>>>(c->d_u, c->d_cellinfo, c->ncell_own, n, d_vt, d_sum, d_cnt);
    nodal_divide<<<(unsigned)ceil_div(c->nnode, 256), 256, 0, c->stream>>>(d_sum, d_cnt, c->nnode);
    c->launches += 2;
    HDG_CUDA(c, cudaMemcpyAsync(out, d_sum, sizeof(double) * c->nnode, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_vt); cudaFree(d_sum); cudaFree(d_cnt);
    return HDG_OK;
}

// ---- K_element[cell], b_element[cell] ----------------------------------------------------------------
__global__ void gather_local(const double* __restrict__ Ke, int64_t c, int m, int t, double* __restrict__ out) {
    // out: K_e column-major m x t, then b_e (m)
    int ke = m * (t + 1);
    const double* tile = Ke + ((c >> 5) * int64_t(ke)) * 32 + (c & 31);
    for (int e = threadIdx.x; e < ke; e += blockDim.x) {
        int i = e / (t + 1), j = e - i * (t + 1);
        double v = tile[int64_t(e) * 32];
        if (j < t) out[j * m + i] = v;
        else out[m * t + i] = v;
    }
}

hdg_status local_download(hdg_context* c, int64_t cell, double* Ke, double* be) {
    const int m = c->tab.m, t = c->tab.t;
    double* d = nullptr;
    HDG_CUDA(c, cudaMalloc(&d, sizeof(double) * m * (t + 1)));
    gather_local<<<1, 128, 0, c->stream>>>(c->d_Ke, cell, m, t, d);
    c->launches += 1;
    std::vector<double> h(size_t(m) * (t + 1));
    HDG_CUDA(c, cudaMemcpyAsync(h.data(), d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d);
    if (Ke) std::copy(h.begin(), h.begin() + size_t(m) * t, Ke);
    if (be) std::copy(h.begin() + size_t(m) * t, h.end(), be);
    return HDG_OK;
}

}  // namespace hdg
