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
//   pcg_spmv    Ap = D K p            + partial sums of p.Ap     (one thread per face row-block, 256-bit loads;
//                                                                 p of faces owned by a neighbouring rank is read
//                                                                 in place over NVLink)
//   pcg_update  alpha = rz/pAp; x += alpha p; r -= alpha Ap; partial sums of r.Dinv r and r.r
//   pcg_dir     beta = rz'/rz; convergence test; p = Dinv r + beta p
// Dot products are reduced deterministically: every block writes one partial, and every block
// of the next kernel re-reduces the (<= 1184) partials in a fixed order (measured faster than a
// last-block reduction: no fences in the big kernels, the reduction overlaps the launch ramp).
// Iterations are launched in CUDA-graph chunks; kernels turn into no-ops once the device-side flag
// says converged, and the host polls that flag once per chunk.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"

namespace hdg {

constexpr int MAX_PARTIALS = 2048;

// d_scal layout
enum Scal : int { S_BNORM2 = 0, S_RELRES = 1, S_MEANSUM = 2, S_ERR2 = 3, S_RZ = 4, NSCAL = 8 };
// d_partials layout: NPART arrays of MAX_PARTIALS
// (P_RZC*: the multigrid part of r.z, same parity scheme as P_RZ*)
enum Part : int { P_PAP = 0, P_RZ0 = 1, P_RZ1 = 2, P_RR = 3, P_BB = 4, P_RZC0 = 5, P_RZC1 = 6, NPART = 7 };

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
    const int32_t* ghost_owner;  // ... and the rank that owns it
    const double* peer_p[MAXR];  // p of the other ranks, mapped over NVLink (ghost_mode 0: ghost values live behind the owned part of p)
    const double* peer_r[MAXR];  // ghost_mode 2: the neighbours' r, Dinv (peer_p then points at their PREVIOUS direction)
    const double* peer_dinv[MAXR];
    double* pnext;               // where pcg_dir writes the next direction (p is double-buffered)
    int ghost_mode;              // 0 local ghost segment (NCCL halo), 1 peer p read in place, 2 peer p recomputed on the fly
    int parity;                  // iteration parity (which r.z partial array is current)
    int mg;                      // r.z = block-Jacobi part (P_RZ*) + multigrid part (P_RZC*)
};

__device__ __forceinline__ double get_sum(const PcgArgs& a, int which) {
    if (a.gscal) return a.gscal[which];                                   // summed over ranks by NCCL
    return reduce_partials(a.part + which * MAX_PARTIALS, a.np);          // fixed-order, every block identically
}

// r.z of the parity slot `which` (P_RZ0 / P_RZ1)
__device__ __forceinline__ double get_rz(const PcgArgs& a, int which) {
    double s = get_sum(a, which);
    if (a.mg) s += get_sum(a, which + (P_RZC0 - P_RZ0));
    return s;
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
        a.part[P_RZ1 * MAX_PARTIALS + blockIdx.x] = blockIdx.x == 0 ? 1.0 : 0.0;   // "r.z of iteration -1": any finite value (p_{-1} = 0)
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
            asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[i]), "=d"(v[i + 1]), "=d"(v[i + 2]), "=d"(v[i + 3]) : "l"(src + i));
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
__device__ __forceinline__ void spmv_face(const PcgArgs& a, int64_t f, double (&y)[NT], double (&pf)[NT]) {
    constexpr int NT2 = NT * NT;
    double blk[NT2];
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
        const double* src = a.p + g * NT;
        double pg[NT];
        if (g >= a.nface && a.ghost_mode) {   // face owned by a neighbouring rank: its p comes straight over NVLink
            const int64_t gi = g - a.nface;
            const int w = a.ghost_owner[gi];
            const int64_t ro = int64_t(a.ghost_ridx[gi]) * NT;
            if (a.ghost_mode == 2) {
                // the neighbour is still writing p_k (pcg_dir runs concurrently, no barrier in between): rebuild it from
                // what is complete - its r_k, Dinv and p_{k-1} (other buffer) - exactly as its pcg_dir does
                const double beta = get_sum(a, a.parity ? P_RZ1 : P_RZ0) / get_sum(a, a.parity ? P_RZ0 : P_RZ1);
                double rg[NT], dg[NT];
                load_vec<NT>(a.peer_p[w] + ro, pg);
                load_vec<NT>(a.peer_r[w] + ro, rg);
                load_vec<NT>(a.peer_dinv[w] + ro, dg);
#pragma unroll
                for (int b = 0; b < NT; ++b) pg[b] = fma(beta, pg[b], dg[b] * rg[b]);
            } else {
                load_vec<NT>(a.peer_p[w] + ro, pg);
            }
        } else {
            load_vec<NT>(src, pg);
        }
        load_vec<NT2>(a.Ko + (f * 4 + s4) * NT2, blk);
#pragma unroll
        for (int r = 0; r < NT; ++r)
#pragma unroll
            for (int b = 0; b < NT; ++b) y[r] = fma(blk[b * NT + r], pg[b], y[r]);
    }
    const bool bc = a.isbc[f];
#pragma unroll
    for (int r = 0; r < NT; ++r) y[r] = bc ? y[r] : -y[r];
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg_spmv(const PcgArgs a) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    double pap = 0.0;
    const int64_t stride = int64_t(gridDim.x) * RB;
    // two faces per trip: the loads of both are in flight before either result is stored
    for (int64_t f0 = int64_t(blockIdx.x) * RB + threadIdx.x; f0 < a.nface; f0 += 2 * stride) {
        const int64_t f1 = f0 + stride;
        double y0[NT], p0[NT], y1[NT], p1[NT];
        spmv_face<NT>(a, f0, y0, p0);
        if (f1 < a.nface) spmv_face<NT>(a, f1, y1, p1);
#pragma unroll
        for (int r = 0; r < NT; ++r) pap = fma(p0[r], y0[r], pap);
        store_vec<NT>(a.Ap + f0 * NT, y0);
        if (f1 < a.nface) {
#pragma unroll
            for (int r = 0; r < NT; ++r) pap = fma(p1[r], y1[r], pap);
            store_vec<NT>(a.Ap + f1 * NT, y1);
        }
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
            if (cc[s] >= a.nface && a.ghost_mode) {
                const int64_t gi = cc[s] - a.nface;
                pg = a.peer_p[a.ghost_owner[gi]] + int64_t(a.ghost_ridx[gi]) * NT;
            }
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
    // 4 rows per thread and trip, all loads issued before the first store (memory-level parallelism: these
    // streaming kernels are latency-, not instruction-bound); the summation order per thread stays fixed
    const double* __restrict__ pp = a.p;
    const double* __restrict__ qq = a.Ap;
    const double* __restrict__ dd = a.dinv;
    double* __restrict__ xx = a.x;
    double* __restrict__ rrp = a.r;
    constexpr int U = 4;
    const int64_t stride = int64_t(gridDim.x) * RB;
    for (int64_t row0 = int64_t(blockIdx.x) * RB + threadIdx.x; row0 < N; row0 += stride * U) {
        double p[U], r[U], q[U], x[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = row0 + u * stride;
            const bool ok = row < N;
            p[u] = ok ? pp[row] : 0.0; r[u] = ok ? rrp[row] : 0.0; q[u] = ok ? qq[row] : 0.0;
            x[u] = ok ? xx[row] : 0.0; d[u] = ok ? dd[row] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = row0 + u * stride;
            if (row < N) {
                xx[row] = fma(alpha, p[u], x[u]);
                const double rn = fma(-alpha, q[u], r[u]);
                rrp[row] = rn;
                rz_new = fma(rn * d[u], rn, rz_new);
                rr = fma(rn, rn, rr);
            }
        }
    }
    double t1 = block_sum(rz_new), t2 = block_sum(rr);
    if (threadIdx.x == 0) {
        a.part[(parity ? P_RZ0 : P_RZ1) * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = t2;
    }
}

__global__ void __launch_bounds__(RB) pcg_dir(const PcgArgs a, int64_t N, int parity, int iter, int z_in_ap) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const double rz_old = get_rz(a, parity ? P_RZ1 : P_RZ0);
    const double rz_new = get_rz(a, parity ? P_RZ0 : P_RZ1);
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
    if (z_in_ap)   // block-Jacobi: z = M^-1 r was left in the Ap buffer by pcg_update_blk
        for (int64_t row = int64_t(blockIdx.x) * RB + threadIdx.x; row < N; row += int64_t(gridDim.x) * RB)
            a.pnext[row] = fma(beta, a.p[row], a.Ap[row]);
    else {
        const double* __restrict__ pp = a.p;
        double* __restrict__ pn = a.pnext;
        const double* __restrict__ dd = a.dinv;
        const double* __restrict__ rp = a.r;
        constexpr int U = 4;
        const int64_t stride = int64_t(gridDim.x) * RB;
        for (int64_t row0 = int64_t(blockIdx.x) * RB + threadIdx.x; row0 < N; row0 += stride * U) {
            double p[U], r[U], d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t row = row0 + u * stride;
                const bool ok = row < N;
                p[u] = ok ? pp[row] : 0.0; r[u] = ok ? rp[row] : 0.0; d[u] = ok ? dd[row] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t row = row0 + u * stride;
                if (row < N) pn[row] = fma(beta, p[u], d[u] * r[u]);
            }
        }
    }
}

// Direction update of the multigrid-PCG with the last stage of the V-cycle folded in: z = (block-Jacobi part, in Ap) + P t, where t is
// the vertex-space result the V-cycle kernel left on level 0, then p = z + beta p.  Thread per owned face.  The V-cycle used to
// end with a grid barrier and a pass that read and rewrote z for every face (12 us + barrier of a 128 us V-cycle at k = 1, 1 M
// elements; 74 of 343 us at k = 3, 4 M); here the two vertex values are gathered while z and p stream through anyway.  Same
// operations on the same values as the separate pass (fma(0.5, a + b, z0), fma(C1, b - a, z1), then fma(beta, p, z)): same bits.
template <int NT>
__global__ void __launch_bounds__(RB) pcg_dir_mg(const PcgArgs a, int parity, int iter, const int32_t* __restrict__ facenode,
                                                const double* __restrict__ t0) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    const double rz_old = get_rz(a, parity ? P_RZ1 : P_RZ0);
    const double rz_new = get_rz(a, parity ? P_RZ0 : P_RZ1);
    const double rr = get_sum(a, P_RR);
    const double bb = a.scal[S_BNORM2];
    const bool conv = rr <= a.rtol * a.rtol * bb;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.scal[S_RELRES] = sqrt(rr / bb);
        a.flags[FLAG_ITERS] = iter;
    }
    if (conv) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.flags[FLAG_DONE] = 1;
        return;
    }
    const double beta = rz_new / rz_old;
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < a.nface; f += int64_t(gridDim.x) * RB) {
        const int2 vv = *reinterpret_cast<const int2*>(facenode + 2 * f);
        const int lo = min(vv.x, vv.y), hi = max(vv.x, vv.y);
        const double ta = t0[lo], tb = t0[hi];
        const bool bc = a.isbc[f] != 0;
        double z[NT], p[NT];
        load_vec<NT>(a.Ap + f * NT, z);
        load_vec<NT>(a.p + f * NT, p);
        if (!bc) {
            z[0] = fma(0.5, ta + tb, z[0]);
            z[1] = fma(MG_C1, tb - ta, z[1]);
        }
#pragma unroll
        for (int e = 0; e < NT; ++e) p[e] = fma(beta, p[e], z[e]);
        store_vec<NT>(a.pnext + f * NT, p);
    }
}

template <int NT> static hdg_status pcg_t(hdg_context* c, double rtol, int maxit, hdg_solve_info* info);

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
__global__ void __launch_bounds__(RB) pcg_init_blk(const PcgArgs a, double* __restrict__ binv) {
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
        a.part[P_RZ1 * MAX_PARTIALS + blockIdx.x] = blockIdx.x == 0 ? 1.0 : 0.0;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RZC0 * MAX_PARTIALS + blockIdx.x] = 0.0;
        a.part[P_RZC1 * MAX_PARTIALS + blockIdx.x] = 0.0;
    }
}

template <int NT>
__global__ void __launch_bounds__(RB) pcg_update_blk(const PcgArgs a, const double* __restrict__ binv, int parity) {
    if (*reinterpret_cast<volatile int32_t*>(a.flags + FLAG_DONE)) return;
    constexpr int NT2 = NT * NT;
    const double alpha = get_rz(a, parity ? P_RZ1 : P_RZ0) / get_sum(a, P_PAP);
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
        store_vec<NT>(a.Ap + f * NT, q);                // Ap is dead now: keep z there for pcg_dir
    }
    double t1 = block_sum(rz_new), t2 = block_sum(rr);
    if (threadIdx.x == 0) {
        a.part[(parity ? P_RZ0 : P_RZ1) * MAX_PARTIALS + blockIdx.x] = t1;
        a.part[P_RR * MAX_PARTIALS + blockIdx.x] = t2;
    }
}

// after convergence (several GPUs, peer memory): ghost entries of x - the faces below the strip, read by the
// recovery - straight from the owners' memory (x was published through the shared p array)
__global__ void pcg_fetch_ghost_x(const PcgArgs a, int64_t nghost, int nt, double* __restrict__ x) {
    int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= nghost * nt) return;
    int64_t gi = k / nt;
    int e = int(k - gi * nt);
    const double* src = a.peer_p[a.ghost_owner[gi]];
    x[(a.nface + gi) * nt + e] = src[int64_t(a.ghost_ridx[gi]) * nt + e];
}

// Host driver.  One GPU: 3 kernels per iteration.  Several GPUs over peer memory (default): the SpMV reads the
// neighbours' p in place, `xgpu_allreduce` (hdg_comm.cu) turns the per-block partials into sums over all ranks and
// doubles as the barrier that orders those peer reads against the next overwrite of p - no NCCL call inside the
// loop.  Without peer access the same kernels run with an NCCL halo exchange into ghost segments and ncclAllReduce.
template <int NT> static hdg_status pcg_t(hdg_context* c, double rtol, int maxit, hdg_solve_info* info) {
    const int64_t N = c->nface_own * NT;          // owned rows
    const int64_t Nloc = c->nface * NT;           // owned + ghost entries of the vectors
    const bool multi = comm_active(c);
    const bool p2p = multi && comm_p2p(c);
    if (multi && !p2p && c->comm->general_mesh)
        return set_err(c, HDG_ERR_NCCL, "partitioned hdg_set_mesh meshes need the peer-memory path (CUDA IPC between the GPUs)");
    const bool mg = c->precond == 2;            // block-Jacobi + P1-vertex multigrid (hdg_mg.cu)
    const bool blockjac = c->precond >= 1;
    if (!c->d_x) HDG_CUDA(c, cudaMalloc(&c->d_x, sizeof(double) * Nloc));
    if (!c->d_p) {
        // one region [p0 | p1 | r | Dinv], each Nloc long: the neighbouring ranks map it (CUDA IPC) and read all four
        HDG_CUDA(c, cudaMalloc(&c->d_p, sizeof(double) * 4 * Nloc));
        HDG_CUDA(c, cudaMalloc(&c->d_Ap, sizeof(double) * Nloc));
        c->d_r = c->d_p + 2 * Nloc;
        c->d_dinv = c->d_p + 3 * Nloc;
        if (p2p) {   // collective: every rank gets here in its first solve
            hdg_status st = comm_share_vectors(c, c->d_p, Nloc);
            if (st) return st;
        }
    }
    if (blockjac && !c->d_binv) HDG_CUDA(c, cudaMalloc(&c->d_binv, sizeof(double) * c->nface_own * NT * NT));
    HDG_CUDA(c, cudaMemsetAsync(c->d_p, 0, sizeof(double) * 2 * Nloc, c->stream));   // p_{-1} = 0, ghost segments = 0
    if (multi) HDG_CUDA(c, cudaMemsetAsync(c->d_x, 0, sizeof(double) * Nloc, c->stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PcgArgs a{};
    a.Kd = c->d_Kd; a.Ko = c->d_Ko; a.kcol = c->d_kcol; a.isbc = c->d_isbc; a.rhs = c->d_rhs;
    a.x = c->d_x; a.r = c->d_r; a.p = c->d_p; a.pnext = c->d_p + Nloc; a.Ap = c->d_Ap; a.dinv = c->d_dinv;
    a.part = c->d_partials; a.scal = c->d_scal; a.flags = c->d_flags; a.nface = c->nface_own;
    a.gscal = multi ? c->comm->d_gscal : nullptr;
    a.mg = mg ? 1 : 0;
    // ghost_mode 2 (no barrier between pcg_dir and the next SpMV) needs the point-Jacobi z = Dinv r; with block-Jacobi
    // or the row-wise nt = 5 SpMV the neighbours' finished p is read after a barrier (mode 1)
    const int ghost_mode = !p2p ? 0 : ((blockjac || NT == 5 || getenv("HDG_PCG_BARRIER")) ? 1 : 2);
    if (p2p) {
        a.ghost_ridx = c->comm->d_ghost_ridx;
        a.ghost_owner = c->comm->d_ghost_owner;
        a.ghost_mode = ghost_mode;
    }
    a.np = int(std::min<int64_t>(ceil_div(c->nface_own, RB), std::min<int64_t>(int64_t(sms) * 8, MAX_PARTIALS)));
    a.rtol = rtol;
    // argument sets of even / odd iterations: p_k lives in buffer k & 1
    PcgArgs arg[2] = {a, a};
    for (int par = 0; par < 2; ++par) {
        arg[par].parity = par;
        arg[par].p = c->d_p + par * Nloc;
        arg[par].pnext = c->d_p + (par ^ 1) * Nloc;
        if (p2p)
            for (int w = 0; w < MAXR; ++w) {
                const double* base = static_cast<const double*>(c->comm->peer_vec[w]);
                if (!base) continue;
                const int64_t Np = c->comm->peer_ndof[w];
                // mode 1: the neighbour's p_k; mode 2: its p_{k-1} (other buffer) + r + Dinv
                arg[par].peer_p[w] = base + (ghost_mode == 2 ? (par ^ 1) : par) * Np;
                arg[par].peer_r[w] = base + 2 * Np;
                arg[par].peer_dinv[w] = base + 3 * Np;
            }
    }
    a = arg[0];
    const int G = a.np;

    timer_start(c, c->t_solve);
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
    if (mg) {
        timer_start(c, c->t_mgsetup);
        hdg_status st = mg_setup(c);
        timer_stop(c, c->t_mgsetup);
        if (st) { timer_stop(c, c->t_solve); return st; }
    }
    const int32_t* mg_facenode = nullptr;
    const double* mg_t0 = nullptr;
    const bool mg_fused = mg && mg_fused_ptrs(c, &mg_facenode, &mg_t0);      // after mg_setup: the level arrays exist
    hdg_status cst = HDG_OK;
    auto note = [&](hdg_status s2) { if (s2) cst = s2; };
    auto global_sums = [&](unsigned mask) {   // several GPUs: selected partial arrays -> sums over all ranks (+ inter-GPU barrier)
        if (!multi) return;
        if (p2p) { note(comm_p2p_allreduce(c, c->d_partials, G, mask)); return; }
        if (mask == 0) return;
        reduce_all<<<NPART, RB, 0, c->stream>>>(c->d_partials, G, c->comm->d_gscal);
        c->launches += 1;
        note(comm_allreduce_sum(c, c->comm->d_gscal, NPART));
    };
    if (blockjac) pcg_init_blk<NT><<<G, RB, 0, c->stream>>>(a, c->d_binv);
    else pcg_init<NT><<<G, RB, 0, c->stream>>>(a);
    if (mg) note(mg_apply(c, a.r, a.p, c->d_partials + P_RZC0 * MAX_PARTIALS, G));      // p_0 = z_0 = M^-1 r_0
    global_sums(0x1fu | (mg ? 0x60u : 0u));
    pcg_init_final<<<1, RB, 0, c->stream>>>(a);
    c->launches += 2;

    // one CUDA graph = CHUNK iterations (even, so the rz double-buffer parity restarts at 0)
    // (the multigrid kernels do not test the converged flag: a short chunk bounds the work done after convergence)
    const int CHUNK = mg ? 4 : 32;
    bool use_graph = getenv("HDG_NO_GRAPH") == nullptr && !(multi && !p2p);      // the NCCL fallback path is launched directly
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    auto enqueue_iter = [&](int it) {
        const int parity = it & 1;
        const PcgArgs& ak = arg[parity];
        if (multi && !p2p) note(comm_halo_exchange(c, ak.p, NT));   // NCCL fallback: ghost entries of p
        if constexpr (NT == 5) pcg_spmv_rows<NT><<<G, RB, 0, c->stream>>>(ak);
        else pcg_spmv<NT><<<G, RB, 0, c->stream>>>(ak);
        global_sums(1u << P_PAP);      // p.Ap; every rank has finished reading the neighbours' vectors
        if (blockjac) pcg_update_blk<NT><<<G, RB, 0, c->stream>>>(ak, c->d_binv, parity);
        else pcg_update<<<G, RB, 0, c->stream>>>(ak, N, parity);
        if (mg) note(mg_apply(c, ak.r, ak.Ap, c->d_partials + (parity ? P_RZC0 : P_RZC1) * MAX_PARTIALS, G, mg_fused));   // z (in Ap) += P V(P'r); fused: P t is added by pcg_dir_mg
        global_sums((1u << (parity ? P_RZ0 : P_RZ1)) | (1u << P_RR) | (mg ? 1u << (parity ? P_RZC0 : P_RZC1) : 0u));   // r.z (+ its vertex-space part), r.r; r complete on every rank
        if (mg_fused) pcg_dir_mg<NT><<<G, RB, 0, c->stream>>>(ak, parity, it + 1, mg_facenode, mg_t0);
        else pcg_dir<<<G, RB, 0, c->stream>>>(ak, N, parity, it + 1, blockjac ? 1 : 0);
        if (ghost_mode == 1) global_sums(0);  // barrier: p complete on every rank before the next SpMV reads it
    };
    int it = 0;
    bool done = false;
    timer_start(c, c->t_loop);
    while (it < maxit && !done) {
        int chunk = std::min(CHUNK, maxit - it);
        if (chunk == CHUNK && use_graph) {
            // the iteration number baked into pcg_dir is relative; FLAG_ITERS is fixed up below
            if (!gexec) {
                HDG_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                for (int k = 0; k < CHUNK; ++k) enqueue_iter(k);
                cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
                if (ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
                if (ce != cudaSuccess || cst != HDG_OK) {      // e.g. a launch kind the driver cannot capture: run the iterations directly
                    cudaGetLastError();
                    if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
                    gexec = nullptr;
                    use_graph = false;
                    cst = HDG_OK;
                    c->err.clear();
                    continue;
                }
            }
            HDG_CUDA(c, cudaGraphLaunch(gexec, c->stream));
        } else {
            for (int k = 0; k < chunk; ++k) enqueue_iter(k);
        }
        c->launches += int64_t(3 + (mg ? mg_launches_per_apply(c) : 0)) * chunk;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        done = c->h_flags[FLAG_DONE] != 0;
        if (done) it += c->h_flags[FLAG_ITERS];
        else it += chunk;
    }
    timer_stop(c, c->t_loop);
    if (multi) {   // recovery reads the trace on the ghost faces below the strip
        if (p2p) {   // publish x through the shared p array, barrier, pull the ghost values over NVLink, barrier
            const int64_t nghost = c->nface - c->nface_own;
            global_sums(0);   // nobody reads p any more
            HDG_CUDA(c, cudaMemcpyAsync(c->d_p, c->d_x, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
            global_sums(0);
            PcgArgs af = arg[0];
            for (int w = 0; w < MAXR; ++w) af.peer_p[w] = static_cast<const double*>(c->comm->peer_vec[w]);   // buffer 0 of the neighbours
            if (nghost > 0) {
                pcg_fetch_ghost_x<<<(unsigned)ceil_div(nghost * NT, 256), 256, 0, c->stream>>>(af, nghost, NT, c->d_x);
                c->launches += 1;
            }
            global_sums(0);
        } else {
            note(comm_halo_exchange(c, c->d_x, NT));
        }
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
    switch (c->tab.nt) {
        case 2: return pcg_t<2>(c, rtol, maxit, info);
        case 3: return pcg_t<3>(c, rtol, maxit, info);
        case 4: return pcg_t<4>(c, rtol, maxit, info);
        case 5: return pcg_t<5>(c, rtol, maxit, info);
    }
    return set_err(c, HDG_ERR_INVALID, "unsupported order");
}

}  // namespace hdg
