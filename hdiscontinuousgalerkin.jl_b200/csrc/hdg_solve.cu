// K3b + K4: Dirichlet apply! and the Jacobi-PCG trace solve.
//
// Replaces apply!(K,b,dbc) (src/boundary.jl:121-175) and u_hat = K \ b
// (examples/poisson2D_HDG.jl:195; UMFPACK LU in the reference).  The condensed matrix is
// symmetric NEGATIVE semi-definite and apply! puts +meandiag on the Dirichlet rows, so PCG runs
// on D*K*u = D*b with D = -1 on free rows, +1 on Dirichlet rows (trap 3);
// Dirichlet rows/columns are decoupled after apply!, hence D*K is SPD and the solution is the
// solution of K u = b.
//
// Matrix storage (hdg_internal.h): per face row-block one diagonal block and four off-diagonal
// blocks (the two other faces of each adjacent cell), every block nt x nt column-major.  No row
// pointers, 16 B of column indices per face.
//
// PCG iteration = 3 kernels, no host synchronisation:
//   pcg_spmv    Ap = D K p            + partial sums of p.Ap
//   pcg_update  alpha = rz/pAp; x += alpha p; r -= alpha Ap; partial sums of r.Dinv r and r.r
//   pcg_dir     beta = rz'/rz; convergence test; p = Dinv r + beta p
// Dot products are reduced deterministically: every block writes one partial, and every block
// of the next kernel re-reduces the (<= 1184) partials in a fixed order.  Iterations are
// launched in CUDA-graph chunks; kernels turn into no-ops once the device-side flag says
// converged, and the host polls that flag once per chunk.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"

namespace hdg {

constexpr int MAX_PARTIALS = 2048;

// d_scal layout
enum Scal : int { S_BNORM2 = 0, S_RELRES = 1, S_MEANSUM = 2, S_ERR2 = 3, S_RZ = 4, NSCAL = 8 };
// d_partials layout: 5 arrays of MAX_PARTIALS
enum Part : int { P_PAP = 0, P_RZ0 = 1, P_RZ1 = 2, P_RR = 3, P_BB = 4, NPART = 5 };

// ---- apply! ----------------------------------------------------------------------------------
template <int NT>
__global__ void diag_abs_partial(const double* __restrict__ Kd, int64_t nface, double* __restrict__ part) {
    double s = 0.0;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < nface; f += int64_t(gridDim.x) * RB) {
#pragma unroll
        for (int a = 0; a < NT; ++a) s += fabs(Kd[f * NT * NT + a * NT + a]);
    }
    double tot = block_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

__device__ inline int64_t bc_index(const int32_t* __restrict__ bfaces, int64_t nb, int32_t f) {
    int64_t lo = 0, hi = nb - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (bfaces[mid] < f) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int NT>
__global__ void apply_bc_kernel(double* __restrict__ Kd, double* __restrict__ Ko, double* __restrict__ rhs,
                                const int32_t* __restrict__ kcol, const uint8_t* __restrict__ isbc,
                                const int32_t* __restrict__ bfaces, int64_t nb, const double* __restrict__ bcval,
                                double m, int64_t nface) {
    constexpr int NT2 = NT * NT;
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    if (isbc[f]) {
        // zero row, K[d,d] = m, f[d] = v*m   (src/boundary.jl:139-157)
        int64_t bi = bcval ? bc_index(bfaces, nb, int32_t(f)) : 0;
#pragma unroll
        for (int e = 0; e < NT2; ++e) Kd[f * NT2 + e] = (e % NT == e / NT) ? m : 0.0;
#pragma unroll
        for (int e = 0; e < 4 * NT2; ++e) Ko[f * 4 * NT2 + e] = 0.0;
#pragma unroll
        for (int a = 0; a < NT; ++a) rhs[f * NT + a] = bcval ? bcval[bi * NT + a] * m : 0.0;
        return;
    }
    for (int s = 0; s < 4; ++s) {
        int32_t g = kcol[4 * f + s];
        if (g < 0 || !isbc[g]) continue;
        double* blk = Ko + (f * 4 + s) * NT2;
        if (bcval) {   // rhs lift f -= v * K[:,d]  (:129-138)
            int64_t bi = bc_index(bfaces, nb, g);
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                double v = bcval[bi * NT + b];
                if (v != 0.0)
#pragma unroll
                    for (int a = 0; a < NT; ++a) rhs[f * NT + a] -= v * blk[b * NT + a];
            }
        }
#pragma unroll
        for (int e = 0; e < NT2; ++e) blk[e] = 0.0;   // zero_out_columns!
    }
}

template <int NT> static hdg_status apply_t(hdg_context* c) {
    int np = int(std::min<int64_t>(ceil_div(c->nface_own, RB), 1024));
    diag_abs_partial<NT><<<np, RB, 0, c->stream>>>(c->d_Kd, c->nface_own, c->d_partials);
    final_sum<<<1, RB, 0, c->stream>>>(c->d_partials, np, c->d_scal + S_MEANSUM);
    c->launches += 2;
    hdg_status st = comm_allreduce_sum(c, c->d_scal + S_MEANSUM, 1);   // mean over ALL dofs of the global system
    if (st) return st;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * NSCAL, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    const int64_t nface_global = comm_active(c) ? c->comm->nface_global : c->nface;
    c->meandiag = c->h_scal[S_MEANSUM] / double(nface_global * NT);   // meandiag, src/boundary.jl:169-175
    apply_bc_kernel<NT><<<(unsigned)ceil_div(c->nface, 128), 128, 0, c->stream>>>(
        c->d_Kd, c->d_Ko, c->d_rhs, c->d_kcol, c->d_isbc, c->d_bfaces, c->nbface, c->d_bcval, c->meandiag, c->nface);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

hdg_status apply_dirichlet(hdg_context* c, const double* values) {
    const int nt = c->tab.nt;
    if (c->d_bcval) { cudaFree(c->d_bcval); c->d_bcval = nullptr; }
    if (values && c->nbface) {
        HDG_CUDA(c, cudaMalloc(&c->d_bcval, sizeof(double) * c->nbface * nt));
        HDG_CUDA(c, cudaMemcpyAsync(c->d_bcval, values, sizeof(double) * c->nbface * nt, cudaMemcpyHostToDevice, c->stream));
    }
    switch (nt) {
        case 2: return apply_t<2>(c);
        case 3: return apply_t<3>(c);
        case 4: return apply_t<4>(c);
        case 5: return apply_t<5>(c);
    }
    return set_err(c, HDG_ERR_INVALID, "unsupported order");
}

// ---- PCG ---------------------------------------------------------------------------------------
struct PcgArgs;
__device__ __forceinline__ double get_sum(const PcgArgs& a, int which);
struct PcgArgs {
    const double* Kd;
    const double* Ko;
    const int32_t* kcol;
    const uint8_t* isbc;
    const double* rhs;
    double *x, *r, *p, *Ap, *dinv;
    double* part;      // NPART x MAX_PARTIALS
    double* scal;
    int32_t* flags;
    int64_t nface;     // OWNED faces: rows of this rank (vectors also carry the ghost faces behind them)
    const double* gscal;   // multi-GPU: all-reduced sums, one per partial array; nullptr on one GPU
    int np;            // number of partials == gridDim of the vector kernels
    double rtol;
    const int32_t* ghost_ridx;   // ghost face -> local face index on its owner (peer-memory SpMV)
    int64_t nbelow;              // ghost faces owned by rank-1 (they come first)
};

__device__ __forceinline__ double get_sum(const PcgArgs& a, int which) {
    if (a.gscal) return a.gscal[which];                                   // summed over ranks by NCCL
    return reduce_partials(a.part + which * MAX_PARTIALS, a.np);          // fixed-order, every block identically
}

// multi-GPU: local sums of all partial arrays -> gscal slots (then all-reduced in place)
__global__ void reduce_all(const double* __restrict__ part, int np, double* __restrict__ gscal) {
    double s = reduce_partials(part + blockIdx.x * MAX_PARTIALS, np);
    if (threadIdx.x == 0) gscal[blockIdx.x] = s;
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg_init(const PcgArgs a) {
    // r = D b ; dinv = 1/diag(D K) ; p = z = dinv r ; x = 0 ; partials of r.z and b.b
    constexpr int NT2 = NT * NT;
    const int64_t N = a.nface * NT;
    double rz = 0.0, bb = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB) {
        int64_t f = row / NT;
        int aa = int(row - f * NT);
        double sgn = a.isbc[f] ? 1.0 : -1.0;
        double d = sgn * a.Kd[f * NT2 + aa * NT + aa];
        double di = 1.0 / d;
        double r = sgn * a.rhs[row];
        double z = di * r;
        a.dinv[row] = di;
        a.r[row] = r;
        a.p[row] = z;
        a.x[row] = 0.0;
        rz += r * z;
        bb += r * r;
    }
    double t1 = block_sum(rz), t2 = block_sum(bb);
    if (threadIdx.x == 0) {
        a.part[P_RZ0 * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_BB * MAX_PARTIALS + blockIdx.x] = t2;
    }
}

__global__ void pcg_init_final(const PcgArgs a) {
    double bb = get_sum(a, P_BB);
    if (threadIdx.x == 0) {
        a.scal[S_BNORM2] = bb;
        a.scal[S_RELRES] = bb > 0.0 ? 1.0 : 0.0;
        a.flags[FLAG_DONE] = bb > 0.0 ? 0 : 1;   // b == 0 -> x = 0 is the solution
        a.flags[FLAG_ITERS] = 0;
    }
}

// widest aligned vector load of N doubles whose address is a multiple of 8*N bytes (256-bit LDG on sm_100a)
template <int N> __device__ __forceinline__ void load_vec(const double* __restrict__ src, double* v) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4)
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[i]), "=d"(v[i + 1]), "=d"(v[i + 2]), "=d"(v[i + 3]) : "l"(src + i));
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            double2 t = *reinterpret_cast<const double2*>(src + i);
            v[i] = t.x; v[i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = src[i];
    }
}
template <int N> __device__ __forceinline__ void store_vec(double* __restrict__ dst, const double* v) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4)
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "d"(v[i]), "d"(v[i + 1]), "d"(v[i + 2]), "d"(v[i + 3]) : "memory");
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = v[i];
    }
}

// One thread per face = per block row: the diagonal block and the 4 off-diagonal blocks of a face are
// 5*nt*nt contiguous-per-array doubles, fetched with the widest aligned vector loads (one 256-bit LDG per
// 2x2 block at k=1); every sector that reaches the SM is fully used.
template <int NT>
__global__ void __launch_bounds__(RB) pcg_spmv(const PcgArgs a) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    constexpr int NT2 = NT * NT;
    double pap = 0.0;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < a.nface; f += int64_t(gridDim.x) * RB) {
        double y[NT], pf[NT], blk[NT2];
        load_vec<NT>(a.p + f * NT, pf);
        load_vec<NT2>(a.Kd + f * NT2, blk);
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < NT; ++b) s = fma(blk[b * NT + r], pf[b], s);
            y[r] = s;
        }
        const int4 cols = *reinterpret_cast<const int4*>(a.kcol + 4 * f);
        const int cc[4] = {cols.x, cols.y, cols.z, cols.w};
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            if (cc[s4] < 0) continue;
            double pg[NT];
            load_vec<NT>(a.p + int64_t(cc[s4]) * NT, pg);
            load_vec<NT2>(a.Ko + (f * 4 + s4) * NT2, blk);
#pragma unroll
            for (int r = 0; r < NT; ++r)
#pragma unroll
                for (int b = 0; b < NT; ++b) y[r] = fma(blk[b * NT + r], pg[b], y[r]);
        }
        const bool bc = a.isbc[f];
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            y[r] = bc ? y[r] : -y[r];
            pap = fma(pf[r], y[r], pap);
        }
        store_vec<NT>(a.Ap + f * NT, y);
    }
    double tot = block_sum(pap);
    if (threadIdx.x == 0) a.part[P_PAP * MAX_PARTIALS + blockIdx.x] = tot;
}

// One thread per scalar row (used for nt = 5, where the blocks are not 16-byte aligned).
template <int NT>
__global__ void __launch_bounds__(RB) pcg_spmv_rows(const PcgArgs a) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    constexpr int NT2 = NT * NT;
    const int64_t N = a.nface * NT;
    double pap = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB) {
        int64_t f = row / NT;
        int aa = int(row - f * NT);
        double y = 0.0;
        const double* kd = a.Kd + f * NT2 + aa;
        const double* pf = a.p + f * NT;
#pragma unroll
        for (int b = 0; b < NT; ++b) y = fma(kd[b * NT], pf[b], y);
        const int4 cols = *reinterpret_cast<const int4*>(a.kcol + 4 * f);
        const int cc[4] = {cols.x, cols.y, cols.z, cols.w};
        const double* ko = a.Ko + f * 4 * NT2 + aa;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (cc[s] < 0) continue;
            const double* pg = a.p + int64_t(cc[s]) * NT;
#pragma unroll
            for (int b = 0; b < NT; ++b) y = fma(ko[s * NT2 + b * NT], pg[b], y);
        }
        y = a.isbc[f] ? y : -y;
        a.Ap[row] = y;
        pap = fma(pf[aa], y, pap);
    }
    double tot = block_sum(pap);
    if (threadIdx.x == 0) a.part[P_PAP * MAX_PARTIALS + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(RB) pcg_update(const PcgArgs a, int64_t N, int parity) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const double pap = get_sum(a, P_PAP);
    const double rz = get_sum(a, parity ? P_RZ1 : P_RZ0);
    const double alpha = rz / pap;
    double rz_new = 0.0, rr = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB) {
        double p = a.p[row], r = a.r[row];
        a.x[row] = fma(alpha, p, a.x[row]);
        r = fma(-alpha, a.Ap[row], r);
        a.r[row] = r;
        rz_new = fma(r * a.dinv[row], r, rz_new);
        rr = fma(r, r, rr);
    }
    double t1 = block_sum(rz_new), t2 = block_sum(rr);
    if (threadIdx.x == 0) {
        a.part[(parity ? P_RZ0 : P_RZ1) * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = t2;
    }
}

__global__ void __launch_bounds__(RB) pcg_dir(const PcgArgs a, int64_t N, int parity, int iter) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const double rz_old = get_sum(a, parity ? P_RZ1 : P_RZ0);
    const double rz_new = get_sum(a, parity ? P_RZ0 : P_RZ1);
    const double rr = get_sum(a, P_RR);
    const double bb = a.scal[S_BNORM2];
    const bool conv = rr <= a.rtol * a.rtol * bb;
    // the decision is recomputed identically by every block; a block that starts after block 0
    // has already raised FLAG_DONE returns at the top, which is the same outcome.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.scal[S_RELRES] = sqrt(rr / bb);
        a.flags[FLAG_ITERS] = iter;
    }
    if (conv) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.flags[FLAG_DONE] = 1;
        return;
    }
    const double beta = rz_new / rz_old;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB)
        a.p[row] = fma(beta, a.p[row], a.dinv[row] * a.r[row]);
}

template <int NT> static hdg_status pcg_t(hdg_context* c, double rtol, int maxit, hdg_solve_info* info) {
    const int64_t N = c->nface_own * NT;          // owned rows
    const int64_t Nloc = c->nface * NT;           // owned + ghost entries of the vectors
    const bool multi = comm_active(c);
    if (!c->d_x) HDG_CUDA(c, cudaMalloc(&c->d_x, sizeof(double) * Nloc));
    if (!c->d_r) {
        HDG_CUDA(c, cudaMalloc(&c->d_r, sizeof(double) * Nloc));
        HDG_CUDA(c, cudaMalloc(&c->d_p, sizeof(double) * Nloc));
        HDG_CUDA(c, cudaMalloc(&c->d_Ap, sizeof(double) * Nloc));
        HDG_CUDA(c, cudaMalloc(&c->d_dinv, sizeof(double) * Nloc));
    }
    if (multi) {
        HDG_CUDA(c, cudaMemsetAsync(c->d_p, 0, sizeof(double) * Nloc, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_x, 0, sizeof(double) * Nloc, c->stream));
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PcgArgs a{};
    a.Kd = c->d_Kd; a.Ko = c->d_Ko; a.kcol = c->d_kcol; a.isbc = c->d_isbc; a.rhs = c->d_rhs;
    a.x = c->d_x; a.r = c->d_r; a.p = c->d_p; a.Ap = c->d_Ap; a.dinv = c->d_dinv;
    a.part = c->d_partials; a.scal = c->d_scal; a.flags = c->d_flags; a.nface = c->nface_own;
    a.gscal = multi ? c->comm->d_gscal : nullptr;
    a.np = int(std::min<int64_t>(ceil_div(N, RB), std::min<int64_t>(int64_t(sms) * 8, MAX_PARTIALS)));
    a.rtol = rtol;
    const int G = a.np;

    timer_start(c, c->t_solve);
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
    hdg_status cst = HDG_OK;
    auto global_sums = [&]() {   // multi-GPU: partial arrays -> all-reduced scalars
        if (!multi) return;
        reduce_all<<<NPART, RB, 0, c->stream>>>(c->d_partials, G, c->comm->d_gscal);
        c->launches += 1;
        hdg_status s2 = comm_allreduce_sum(c, c->comm->d_gscal, NPART);
        if (s2) cst = s2;
    };
    pcg_init<NT><<<G, RB, 0, c->stream>>>(a);
    global_sums();
    pcg_init_final<<<1, RB, 0, c->stream>>>(a);
    c->launches += 2;

    // one CUDA graph = CHUNK iterations (even, so the rz double-buffer parity restarts at 0)
    const int CHUNK = 32;
    const bool use_graph = getenv("HDG_NO_GRAPH") == nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    auto enqueue_iter = [&](int it) {
        int parity = it & 1;
        if (multi) {   // ghost entries of p from the neighbouring strips
            hdg_status s2 = comm_halo_exchange(c, c->d_p, NT);
            if (s2) cst = s2;
        }
        if constexpr (NT == 5) pcg_spmv_rows<NT><<<G, RB, 0, c->stream>>>(a);
        else pcg_spmv<NT><<<G, RB, 0, c->stream>>>(a);
        global_sums();
        pcg_update<<<G, RB, 0, c->stream>>>(a, N, parity);
        global_sums();
        pcg_dir<<<G, RB, 0, c->stream>>>(a, N, parity, it + 1);
    };
    int it = 0;
    bool done = false;
    while (it < maxit && !done) {
        int chunk = std::min(CHUNK, maxit - it);
        if (chunk == CHUNK && use_graph) {
            // the iteration number baked into pcg_dir is relative; FLAG_ITERS is fixed up below
            if (!gexec) {
                HDG_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                for (int k = 0; k < CHUNK; ++k) enqueue_iter(k);
                HDG_CUDA(c, cudaStreamEndCapture(c->stream, &graph));
                HDG_CUDA(c, cudaGraphInstantiate(&gexec, graph, 0));
            }
            HDG_CUDA(c, cudaGraphLaunch(gexec, c->stream));
        } else {
            for (int k = 0; k < chunk; ++k) enqueue_iter(k);
        }
        c->launches += 3 * chunk;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        done = c->h_flags[FLAG_DONE] != 0;
        if (done) it += c->h_flags[FLAG_ITERS];
        else it += chunk;
    }
    if (multi) {   // recovery reads the trace on the ghost faces below the strip
        hdg_status s2 = comm_halo_exchange(c, c->d_x, NT);
        if (s2) cst = s2;
    }
    timer_stop(c, c->t_solve);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * NSCAL, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (info) {
        info->iterations = it;
        info->converged = done ? 1 : 0;
        info->relres = c->h_scal[S_RELRES];
        info->bnorm = std::sqrt(c->h_scal[S_BNORM2]);
        info->solve_ms = timer_ms(c->t_solve);
    }
    c->solved = true;
    if (cst) return cst;
    if (!done) return set_err(c, HDG_ERR_NOT_CONVERGED, "PCG did not converge in " + std::to_string(maxit) + " iterations");
    return HDG_OK;
}

// =================================================================================================
// PCG, main path (one GPU, and several GPUs over peer memory): 3 kernels per iteration and NOTHING else -
// no reduction kernels, no NCCL, no host synchronisation.
//   pcg3_spmv    Ap = D K p + p.Ap          p of faces owned by a neighbouring rank is loaded straight from that
//                                           rank's memory over NVLink (CUDA IPC mapping): the halo exchange is
//                                           part of the SpMV
//   pcg3_update  alpha = rz/pAp; x += alpha p; r -= alpha Ap; r.Dinv r, r.r
//   pcg3_dir     beta = rz'/rz; convergence test; p = Dinv r + beta p
// Reductions: every block writes its partial sums; the block that finishes LAST (ticket counter) adds the
// partials in index order (bitwise reproducible) and stores the result, tagged with the iteration number, into
// the mailbox of every rank (plain stores over NVLink + one system fence).  The consumer kernel polls its own
// mailbox until all ranks of that iteration have arrived and adds the contributions in rank order, so all
// ranks hold identical scalars.  A message is therefore both the all-reduce and the inter-GPU barrier that
// orders the peer reads of p against its next overwrite; pcg3_dir additionally posts a value-less "p ready"
// message that the next SpMV waits for.
// =================================================================================================
enum Msg : int { MSG_PAP = 0, MSG_RZRR = 1, MSG_PREADY = 2, NMSG = 3 };

struct Pcg3Sync {
    unsigned long long base_iter;   // absolute index of iteration 0 of the current chunk (message tags = index + 1)
    double val[NMSG][2];            // the reduced (all ranks) values of the latest message of each kind
    unsigned int ticket[NMSG];
};

struct Pcg3Args {
    PcgArgs a;
    Pcg3Sync* sync;
    double* my_mail;            // [NMSG][2][nranks][MAILW]
    double* const* peer_mail;   // the same region of every rank, as mapped here
    const double* peer_p[2];    // p of rank-1 / rank+1
    int rank, nranks;
};

__device__ __forceinline__ double* mail_slot(double* base, int msg, int buf, int nranks, int q) {
    return base + (size_t(msg * 2 + buf) * nranks + q) * MAILW;
}

// called by every thread of every block after the block's partial sums were stored to part[slot_v][blockIdx]
template <int NV>
__device__ __forceinline__ void last_block_send(const Pcg3Args& A, int msg, unsigned long long tag, const int (&slots)[NV == 0 ? 1 : NV]) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&A.sync->ticket[msg], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double vals[NV == 0 ? 1 : NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double s = 0.0;
        for (int i = threadIdx.x; i < A.a.np; i += RB) s += __ldcg(A.a.part + slots[v] * MAX_PARTIALS + i);
        vals[v] = block_sum(s);
    }
    if (A.nranks == 1) {
        if (threadIdx.x == 0)
#pragma unroll
            for (int v = 0; v < NV; ++v) A.sync->val[msg][v] = vals[v];
    } else if (threadIdx.x < A.nranks) {
        volatile double* dst = mail_slot(A.peer_mail[threadIdx.x], msg, int(tag & 1ull), A.nranks, A.rank);
#pragma unroll
        for (int v = 0; v < NV; ++v) dst[1 + v] = vals[v];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(dst) = tag;
    }
    if (threadIdx.x == 0) A.sync->ticket[msg] = 0;
}

// Several GPUs: a one-warp kernel between producer and consumer waits (stream-ordered) until every rank has
// posted message `msg` of this iteration and adds the contributions in rank order -> sync->val.  Keeping the
// polling out of the big kernels avoids thousands of pollers hammering the L2 line the remote store targets.
__global__ void pcg3_wait(const Pcg3Args A, int msg, int nv, int kiter, int tag_offset) {
    if (*reinterpret_cast<volatile int32_t*>(A.a.flags + FLAG_DONE)) return;
    const unsigned long long tag = A.sync->base_iter + kiter + 1 + tag_offset;
    const int buf = int(tag & 1ull);
    if (threadIdx.x < A.nranks) {
        volatile unsigned long long* flag =
            reinterpret_cast<volatile unsigned long long*>(mail_slot(A.my_mail, msg, buf, A.nranks, threadIdx.x));
        while (*flag != tag) { }
    }
    __syncwarp();
    if (threadIdx.x < nv) {
        double s = 0.0;
        for (int q = 0; q < A.nranks; ++q)
            s += reinterpret_cast<volatile double*>(mail_slot(A.my_mail, msg, buf, A.nranks, q))[1 + threadIdx.x];
        A.sync->val[msg][threadIdx.x] = s;
    }
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg3_spmv(const Pcg3Args A, int kiter) {
    const PcgArgs& a = A.a;
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    constexpr int NT2 = NT * NT;
    const unsigned long long tag = A.sync->base_iter + kiter + 1;
    double pap = 0.0;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < a.nface; f += int64_t(gridDim.x) * RB) {
        double y[NT], pf[NT], blk[NT2];
        load_vec<NT>(a.p + f * NT, pf);
        load_vec<NT2>(a.Kd + f * NT2, blk);
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < NT; ++b) s = fma(blk[b * NT + r], pf[b], s);
            y[r] = s;
        }
        const int4 cols = *reinterpret_cast<const int4*>(a.kcol + 4 * f);
        const int cc[4] = {cols.x, cols.y, cols.z, cols.w};
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int64_t g = cc[s4];
            if (g < 0) continue;
            const double* pg;
            if (g < a.nface) pg = a.p + g * NT;
            else {   // face owned by a neighbouring rank: its p comes over NVLink
                const int64_t gi = g - a.nface;
                pg = A.peer_p[gi < a.nbelow ? 0 : 1] + int64_t(a.ghost_ridx[gi]) * NT;
            }
            double pv[NT];
            load_vec<NT>(pg, pv);
            load_vec<NT2>(a.Ko + (f * 4 + s4) * NT2, blk);
#pragma unroll
            for (int r = 0; r < NT; ++r)
#pragma unroll
                for (int b = 0; b < NT; ++b) y[r] = fma(blk[b * NT + r], pv[b], y[r]);
        }
        const bool bc = a.isbc[f];
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            y[r] = bc ? y[r] : -y[r];
            pap = fma(pf[r], y[r], pap);
        }
        store_vec<NT>(a.Ap + f * NT, y);
    }
    double tot = block_sum(pap);
    if (threadIdx.x == 0) a.part[P_PAP * MAX_PARTIALS + blockIdx.x] = tot;
    const int slots[1] = {P_PAP};
    last_block_send<1>(A, MSG_PAP, tag, slots);
}

__global__ void __launch_bounds__(RB) pcg3_update(const Pcg3Args A, int64_t N, int kiter) {
    const PcgArgs& a = A.a;
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const unsigned long long tag = A.sync->base_iter + kiter + 1;
    const double rz = a.scal[S_RZ];
    const double alpha = rz / A.sync->val[MSG_PAP][0];
    double rz_new = 0.0, rr = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB) {
        double r = a.r[row];
        a.x[row] = fma(alpha, a.p[row], a.x[row]);
        r = fma(-alpha, a.Ap[row], r);
        a.r[row] = r;
        rz_new = fma(r * a.dinv[row], r, rz_new);
        rr = fma(r, r, rr);
    }
    double t1 = block_sum(rz_new), t2 = block_sum(rr);
    if (threadIdx.x == 0) {
        a.part[P_RZ0 * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = t2;
    }
    const int slots[2] = {P_RZ0, P_RR};
    last_block_send<2>(A, MSG_RZRR, tag, slots);
}

__global__ void __launch_bounds__(RB) pcg3_dir(const Pcg3Args A, int64_t N, int kiter, int z_in_ap) {
    const PcgArgs& a = A.a;
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const unsigned long long tag = A.sync->base_iter + kiter + 1;
    const double rz_new = A.sync->val[MSG_RZRR][0], rr = A.sync->val[MSG_RZRR][1], rz_old = a.scal[S_RZ], bb = a.scal[S_BNORM2];
    const bool conv = rr <= a.rtol * a.rtol * bb;
    // S_RZ is read by every block of this kernel before the LAST block overwrites it (see below)
    const double beta = rz_new / rz_old;
    if (!conv) {
        if (z_in_ap)   // block-Jacobi: z = M^-1 r was left in the Ap buffer by pcg3_update_blk
            for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB)
                a.p[row] = fma(beta, a.p[row], a.Ap[row]);
        else
            for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB)
                a.p[row] = fma(beta, a.p[row], a.dinv[row] * a.r[row]);
    }
    // last block: publish rz for the next iteration, the convergence state, and "p ready" to the neighbours
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&A.sync->ticket[MSG_PREADY], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) {
        a.scal[S_RZ] = rz_new;
        a.scal[S_RELRES] = sqrt(rr / bb);
        a.flags[FLAG_ITERS] = kiter + 1;
        if (conv) a.flags[FLAG_DONE] = 1;
        A.sync->ticket[MSG_PREADY] = 0;
    }
    if (!conv && A.nranks > 1 && threadIdx.x < A.nranks) {
        volatile double* dst = mail_slot(A.peer_mail[threadIdx.x], MSG_PREADY, int((tag + 1) & 1ull), A.nranks, A.rank);
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(dst) = tag + 1;
    }
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg3_init(const PcgArgs a) {
    // r = D b ; dinv = 1/diag(D K) ; p = z = dinv r ; x = 0 ; partials of r.z and b.b
    constexpr int NT2 = NT * NT;
    const int64_t N = a.nface * NT;
    double rz = 0.0, bb = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB) {
        int64_t f = row / NT;
        int aa = int(row - f * NT);
        double sgn = a.isbc[f] ? 1.0 : -1.0;
        double di = 1.0 / (sgn * a.Kd[f * NT2 + aa * NT + aa]);
        double r = sgn * a.rhs[row];
        double z = di * r;
        a.dinv[row] = di;
        a.r[row] = r;
        a.p[row] = z;
        a.x[row] = 0.0;
        rz += r * z;
        bb += r * r;
    }
    double t1 = block_sum(rz), t2 = block_sum(bb);
    if (threadIdx.x == 0) {
        a.part[P_RZ0 * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_BB * MAX_PARTIALS + blockIdx.x] = t2;
        a.part[P_PAP * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RZ1 * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = 0.0;
    }
}

// ---- block-Jacobi variant: M = blockdiag(D K_ff), the nt x nt face-diagonal blocks --------------
// The trace matrix is naturally blocked by face, so the preconditioner costs one nt x nt mat-vec per face; z is
// kept in the Ap buffer (dead after the update) for the direction update.  Thread per face, vector loads.
template <int NT> __device__ __forceinline__ void invert_block(double (&a)[NT * NT]) {
    // in-place Gauss-Jordan without pivoting on a symmetric positive definite block (column-major)
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        const double ip = 1.0 / a[k * NT + k];
        a[k * NT + k] = 1.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) a[j * NT + k] *= ip;            // row k
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            if (i == k) continue;
            const double f = a[k * NT + i];
            a[k * NT + i] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) a[j * NT + i] = fma(-f, a[j * NT + k], a[j * NT + i]);
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg3_init_blk(const PcgArgs a, double* __restrict__ binv) {
    constexpr int NT2 = NT * NT;
    double rz = 0.0, bb = 0.0;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < a.nface; f += int64_t(gridDim.x) * RB) {
        double B[NT2], r[NT], z[NT], zero[NT];
        load_vec<NT2>(a.Kd + f * NT2, B);
        load_vec<NT>(a.rhs + f * NT, r);
        const double sgn = a.isbc[f] ? 1.0 : -1.0;
#pragma unroll
        for (int e = 0; e < NT2; ++e) B[e] *= sgn;
#pragma unroll
        for (int e = 0; e < NT; ++e) { r[e] *= sgn; zero[e] = 0.0; }
        invert_block<NT>(B);
        store_vec<NT2>(binv + f * NT2, B);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) s = fma(B[j * NT + i], r[j], s);
            z[i] = s;
            rz = fma(r[i], s, rz);
            bb = fma(r[i], r[i], bb);
        }
        store_vec<NT>(a.r + f * NT, r);
        store_vec<NT>(a.p + f * NT, z);
        store_vec<NT>(a.x + f * NT, zero);
    }
    double t1 = block_sum(rz), t2 = block_sum(bb);
    if (threadIdx.x == 0) {
        a.part[P_RZ0 * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_BB * MAX_PARTIALS + blockIdx.x] = t2;
        a.part[P_PAP * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RZ1 * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = 0.0;
    }
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg3_update_blk(const Pcg3Args A, const double* __restrict__ binv, int kiter) {
    const PcgArgs& a = A.a;
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    constexpr int NT2 = NT * NT;
    const unsigned long long tag = A.sync->base_iter + kiter + 1;
    const double alpha = a.scal[S_RZ] / A.sync->val[MSG_PAP][0];
    double rz_new = 0.0, rr = 0.0;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < a.nface; f += int64_t(gridDim.x) * RB) {
        double r[NT], p[NT], q[NT], x[NT], B[NT2];
        load_vec<NT>(a.r + f * NT, r);
        load_vec<NT>(a.p + f * NT, p);
        load_vec<NT>(a.Ap + f * NT, q);
        load_vec<NT>(a.x + f * NT, x);
        load_vec<NT2>(binv + f * NT2, B);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            x[i] = fma(alpha, p[i], x[i]);
            r[i] = fma(-alpha, q[i], r[i]);
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) s = fma(B[j * NT + i], r[j], s);
            q[i] = s;                                   // z = M^-1 r
            rz_new = fma(r[i], s, rz_new);
            rr = fma(r[i], r[i], rr);
        }
        store_vec<NT>(a.x + f * NT, x);
        store_vec<NT>(a.r + f * NT, r);
        store_vec<NT>(a.Ap + f * NT, q);                // Ap is dead now: keep z there for pcg3_dir_blk
    }
    double t1 = block_sum(rz_new), t2 = block_sum(rr);
    if (threadIdx.x == 0) {
        a.part[P_RZ0 * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = t2;
    }
    const int slots[2] = {P_RZ0, P_RR};
    last_block_send<2>(A, MSG_RZRR, tag, slots);
}

// after the (all-reduced) sums of pcg3_init are available: scalars, flags, and the "p ready" tags of iteration 0
// (the all-reduce that precedes this kernel already is a barrier over all ranks)
__global__ void pcg3_init_final(const Pcg3Args A) {
    const PcgArgs& a = A.a;
    const double bb = get_sum(a, P_BB), rz = get_sum(a, P_RZ0);
    if (threadIdx.x == 0) {
        a.scal[S_BNORM2] = bb;
        a.scal[S_RZ] = rz;
        a.scal[S_RELRES] = bb > 0.0 ? 1.0 : 0.0;
        a.flags[FLAG_DONE] = bb > 0.0 ? 0 : 1;   // b == 0 -> x = 0 is the solution
        a.flags[FLAG_ITERS] = 0;
        for (int m = 0; m < NMSG; ++m) A.sync->ticket[m] = 0;
    }
    if (threadIdx.x < A.nranks) {
        const unsigned long long tag = A.sync->base_iter + 1;
        *reinterpret_cast<volatile unsigned long long*>(mail_slot(A.my_mail, MSG_PREADY, int(tag & 1ull), A.nranks, threadIdx.x)) = tag;
    }
}

// end of a chunk of `n` iterations: advance the absolute iteration index (skip ahead after convergence so that no
// message tag is ever reused by a later solve)
__global__ void pcg3_advance(const Pcg3Args A, int n) {
    if (threadIdx.x == 0) A.sync->base_iter += A.a.flags[FLAG_DONE] ? n + 8 : n;
}

// after convergence: ghost entries of x (the faces below the strip, read by the recovery) from the owners' memory
__global__ void pcg3_fetch_ghost_x(const Pcg3Args A, int64_t nghost, int nt, double* __restrict__ x) {
    int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= nghost * nt) return;
    int64_t gi = k / nt;
    int a = int(k - gi * nt);
    const double* src = A.peer_p[gi < A.a.nbelow ? 0 : 1];
    x[(A.a.nface + gi) * nt + a] = src[int64_t(A.a.ghost_ridx[gi]) * nt + a];
}

template <int NT> static hdg_status pcg3_t(hdg_context* c, double rtol, int maxit, hdg_solve_info* info) {
    const int64_t N = c->nface_own * NT, Nloc = c->nface * NT;
    const bool multi = comm_active(c);
    const int nranks = multi ? c->comm->nranks : 1, rank = multi ? c->comm->rank : 0;
    if (!c->d_x) HDG_CUDA(c, cudaMalloc(&c->d_x, sizeof(double) * Nloc));
    if (!c->d_Ap) HDG_CUDA(c, cudaMalloc(&c->d_Ap, sizeof(double) * Nloc));
    if (!c->d_vreg) {
        HDG_CUDA(c, cudaMalloc(&c->d_vreg, sizeof(double) * 3 * N));
        if (multi) {
            hdg_status st = comm_share_vectors(c, c->d_vreg, N);
            if (st) return st;
        }
    }
    if (!c->d_pcg_sync) {
        HDG_CUDA(c, cudaMalloc(&c->d_pcg_sync, sizeof(Pcg3Sync)));
        HDG_CUDA(c, cudaMemset(c->d_pcg_sync, 0, sizeof(Pcg3Sync)));
    }
    double* my_mail = nullptr;
    double* const* peer_mail = nullptr;
    if (multi) {
        my_mail = c->comm->d_mail + 2 * nranks * MAILW;             // behind the xgpu_allreduce mailbox
        peer_mail = c->comm->d_peer_mail + nranks;                  // second row: pointers to the PCG mailboxes
    } else {
        if (!c->d_pcg_mail) {
            HDG_CUDA(c, cudaMalloc(&c->d_pcg_mail, sizeof(double) * NMSG * 2 * MAILW + sizeof(double*)));
            HDG_CUDA(c, cudaMemset(c->d_pcg_mail, 0, sizeof(double) * NMSG * 2 * MAILW));
            double* self = c->d_pcg_mail;
            HDG_CUDA(c, cudaMemcpy(c->d_pcg_mail + NMSG * 2 * MAILW, &self, sizeof(double*), cudaMemcpyHostToDevice));
        }
        my_mail = c->d_pcg_mail;
        peer_mail = reinterpret_cast<double* const*>(c->d_pcg_mail + NMSG * 2 * MAILW);
    }
    if (multi) HDG_CUDA(c, cudaMemsetAsync(c->d_x, 0, sizeof(double) * Nloc, c->stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Pcg3Args A{};
    PcgArgs& a = A.a;
    a.Kd = c->d_Kd; a.Ko = c->d_Ko; a.kcol = c->d_kcol; a.isbc = c->d_isbc; a.rhs = c->d_rhs;
    a.x = c->d_x; a.r = c->d_vreg; a.dinv = c->d_vreg + N; a.p = c->d_vreg + 2 * N; a.Ap = c->d_Ap;
    a.part = c->d_partials; a.scal = c->d_scal; a.flags = c->d_flags; a.nface = c->nface_own;
    a.gscal = multi ? c->comm->d_gscal : nullptr;
    a.np = int(std::min<int64_t>(ceil_div(c->nface_own, RB), std::min<int64_t>(int64_t(sms) * 8, MAX_PARTIALS)));
    a.rtol = rtol;
    A.sync = static_cast<Pcg3Sync*>(c->d_pcg_sync);
    A.my_mail = my_mail; A.peer_mail = peer_mail; A.rank = rank; A.nranks = nranks;
    if (multi) {
        for (int w = 0; w < 2; ++w) {
            const double* base = static_cast<const double*>(c->comm->peer_vec[w]);
            A.peer_p[w] = base ? base + 2 * c->comm->peer_ndof[w] : nullptr;
        }
        a.ghost_ridx = c->comm->d_ghost_ridx;
        a.nbelow = c->comm->nbelow;
    }
    const int G = a.np;
    hdg_status cst = HDG_OK;
    timer_start(c, c->t_solve);
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
    const bool blockjac = c->precond == 1;
    if (blockjac && !c->d_binv) HDG_CUDA(c, cudaMalloc(&c->d_binv, sizeof(double) * c->nface_own * NT * NT));
    if (blockjac) pcg3_init_blk<NT><<<G, RB, 0, c->stream>>>(a, c->d_binv);
    else pcg3_init<NT><<<G, RB, 0, c->stream>>>(a);
    if (multi) {   // sums of b.b and r.z over all ranks; also the barrier before the first peer read of p
        hdg_status s2 = comm_p2p_allreduce(c, c->d_partials, G, NPART);
        if (s2) cst = s2;
    }
    pcg3_init_final<<<1, RB, 0, c->stream>>>(A);
    c->launches += 2;
    const int CHUNK = 32;
    const bool use_graph = getenv("HDG_NO_GRAPH") == nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const bool waits = multi && getenv("HDG_DBG_NOWAIT") == nullptr;   // debug: free-running ranks (wrong numerics, timing only)
    auto enqueue_chunk = [&](int n) {
        for (int k = 0; k < n; ++k) {
            if (waits) pcg3_wait<<<1, 32, 0, c->stream>>>(A, MSG_PREADY, 0, k, 0);   // neighbours' p complete
            pcg3_spmv<NT><<<G, RB, 0, c->stream>>>(A, k);
            if (waits) pcg3_wait<<<1, 32, 0, c->stream>>>(A, MSG_PAP, 1, k, 0);
            if (blockjac) pcg3_update_blk<NT><<<G, RB, 0, c->stream>>>(A, c->d_binv, k);
            else pcg3_update<<<G, RB, 0, c->stream>>>(A, N, k);
            if (waits) pcg3_wait<<<1, 32, 0, c->stream>>>(A, MSG_RZRR, 2, k, 0);
            pcg3_dir<<<G, RB, 0, c->stream>>>(A, N, k, blockjac ? 1 : 0);
        }
        pcg3_advance<<<1, 32, 0, c->stream>>>(A, n);
    };
    int it = 0;
    bool done = false;
    while (it < maxit && !done) {
        int chunk = std::min(CHUNK, maxit - it);
        if (chunk == CHUNK && use_graph) {
            if (!gexec) {
                HDG_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                enqueue_chunk(CHUNK);
                HDG_CUDA(c, cudaStreamEndCapture(c->stream, &graph));
                HDG_CUDA(c, cudaGraphInstantiate(&gexec, graph, 0));
            }
            HDG_CUDA(c, cudaGraphLaunch(gexec, c->stream));
        } else {
            enqueue_chunk(chunk);
        }
        c->launches += 3 * chunk + 1;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        done = c->h_flags[FLAG_DONE] != 0;
        it += done ? c->h_flags[FLAG_ITERS] : chunk;
    }
    if (multi) {
        // recovery reads the trace on the ghost faces below the strip: publish x through the shared p slot,
        // barrier, pull the ghost values over NVLink, barrier (p is rewritten by the next solve)
        const int64_t nghost = c->nface - c->nface_own;
        HDG_CUDA(c, cudaMemcpyAsync(a.p, c->d_x, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
        hdg_status s2 = comm_p2p_allreduce(c, c->d_partials, G, 0);
        if (s2) cst = s2;
        if (nghost > 0) {
            pcg3_fetch_ghost_x<<<(unsigned)ceil_div(nghost * NT, 256), 256, 0, c->stream>>>(A, nghost, NT, c->d_x);
            c->launches += 1;
        }
        s2 = comm_p2p_allreduce(c, c->d_partials, G, 0);
        if (s2) cst = s2;
    }
    timer_stop(c, c->t_solve);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double) * NSCAL, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (info) {
        info->iterations = it;
        info->converged = done ? 1 : 0;
        info->relres = c->h_scal[S_RELRES];
        info->bnorm = std::sqrt(c->h_scal[S_BNORM2]);
        info->solve_ms = timer_ms(c->t_solve);
    }
    c->solved = true;
    if (cst) return cst;
    if (!done) return set_err(c, HDG_ERR_NOT_CONVERGED, "PCG did not converge in " + std::to_string(maxit) + " iterations");
    return HDG_OK;
}

hdg_status pcg_solve(hdg_context* c, double rtol, int maxit, hdg_solve_info* info) {
    // mailbox PCG on one GPU and, over peer memory, on several; the NCCL variant is the
    // fallback when the GPUs cannot map each other's memory (or HDG_PCG_LEGACY is set)
    const bool fused = getenv("HDG_PCG_LEGACY") == nullptr && (!comm_active(c) || comm_p2p(c));
    if (fused) switch (c->tab.nt) {
        case 2: return pcg3_t<2>(c, rtol, maxit, info);
        case 3: return pcg3_t<3>(c, rtol, maxit, info);
        case 4: return pcg3_t<4>(c, rtol, maxit, info);
        case 5: return pcg3_t<5>(c, rtol, maxit, info);
    }
    switch (c->tab.nt) {
        case 2: return pcg_t<2>(c, rtol, maxit, info);
        case 3: return pcg_t<3>(c, rtol, maxit, info);
        case 4: return pcg_t<4>(c, rtol, maxit, info);
        case 5: return pcg_t<5>(c, rtol, maxit, info);
    }
    return set_err(c, HDG_ERR_INVALID, "unsupported order");
}

}  // namespace hdg
