// P1-vertex multigrid preconditioner for the trace PCG (hdg_set_preconditioner(ctx, 2)).
//
// The Jacobi iteration count of the condensed trace system grows like 1/h (4 472 iterations at 1 M elements, 27 643 at
// 16 M).  This preconditioner is the additive two-space method
//     M^-1 r = Binv r + P V(P' r)
// Binv = the inverted nt x nt face-diagonal blocks (block-Jacobi), P = the trace of a continuous P1 function (vertex
// values a, b on a face with vertices lo < hi give the Legendre coefficients (a+b)/2 and (b-a)/(2 sqrt 3), higher modes
// 0), V = one V(1,1) cycle of geometric multigrid for the Galerkin vertex operator A_c = P'AP.  A_c lives on the
// (nx+1) x (ny+1) vertex grid of rectangle_mesh (src/generate_mesh.jl:101-143: node id = iy (nx+1) + ix, cell diagonal
// from (i+1,j) to (i,j+1)) as a 7-point stencil; the coarser operators are Galerkin products with the P1 interpolation of
// that triangulation (stays 7-point), down to <= 64 points, which are solved with a dense inverse.  Vertices on
// Dirichlet faces carry no coarse unknown.  Everything is gather-formulated (no atomics): bitwise reproducible.
// Measured in the scipy prototype (tools/mg_prototype.py): 34-40 PCG iterations to 1e-12 for k = 1..4, independent of h.
// rectangle_mesh triangulations only (a general mesh needs an algebraic hierarchy for A_c - next).  On several GPUs (strips of
// hdg_set_rectangle_mesh) the vertex hierarchy is replicated: every rank builds the rows of A_c of the faces it owns and
// restricts the residual of its own faces, the partial level-0 stencils (once per solve) and vertex residuals (once per
// iteration) are summed with ncclAllReduce, and every rank runs the same V-cycle on the whole vertex grid (1 vertex per 2
// cells and 1 double each: small against the trace system).  The PCG then runs without CUDA-graph capture.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"

namespace hdg {

constexpr int MG_MAXVAL = 8;        // faces per vertex (6 on rectangle_mesh)
constexpr int MG_DENSE = 64;        // the coarsest grid has at most this many points
constexpr int MG_MAXLEV = 24;
#ifndef MG_FUSE_MAX
#define MG_FUSE_MAX 640            // levels with at most this many points run inside ONE single-block kernel (mg_fused_vcycle)
#endif
constexpr int MG_FUSE_LEVELS = 8;   // capacity of the argument struct (640 -> 160 -> 40: 3 fused levels by default)
constexpr double MG_OMEGA = 0.8;    // damped Jacobi on the vertex grids
constexpr double MG_C1 = 0.28867513459481287;   // 1 / (2 sqrt 3)

// stencil slots: centre, E, W, N, S, SE, NW  (the neighbours of a vertex in the triangulation)
__device__ __constant__ int MG_DX[7] = {0, 1, -1, 0, 0, 1, -1};
__device__ __constant__ int MG_DY[7] = {0, 0, 0, 1, -1, -1, 1};
__device__ __forceinline__ int mg_slot(int dx, int dy) {
    if (dy == 0) return dx == 0 ? 0 : (dx == 1 ? 1 : (dx == -1 ? 2 : -1));
    if (dx == 0) return dy == 1 ? 3 : (dy == -1 ? 4 : -1);
    if (dx == 1 && dy == -1) return 5;
    if (dx == -1 && dy == 1) return 6;
    return -1;
}

struct MgLevel {
    int px = 0, py = 0;
    int64_t n = 0;
    double* st = nullptr;     // 7 x n, slot-major
    double* dinv = nullptr;   // 1/diagonal, 0 at fixed points (identity rows)
    double *r = nullptr, *x = nullptr, *t = nullptr;
};

struct MgData {
    int nlev = 0;
    MgLevel lev[MG_MAXLEV];
    double* pool = nullptr;          // one allocation for all levels
    double* ainv = nullptr;          // dense inverse of the coarsest operator
    int32_t* vface = nullptr;        // nnode x MG_MAXVAL: incident faces ascending, bit 31 = the vertex is the face's hi vertex
    int32_t* vcnt = nullptr;         // nnode: number of incident faces, -1 = fixed (touches a Dirichlet face)
    int64_t nnode = 0, nface = 0;
    int nx = 0, ny = 0;
    int lf = 0;                      // first level of the fused tail (levels lf .. nlev-1 run in mg_fused_vcycle)
    bool adjacency_ok = false;
    // several GPUs (strips of rectangle_mesh): the vertex hierarchy is REPLICATED on every rank over the global vertex grid;
    // a rank contributes the rows of the faces it owns and the partial vertex vectors / stencils are all-reduced
    bool multi = false;
    int64_t node0 = 0;               // global id of local node 0
    int64_t nface_rows = 0;          // owned faces (rows of the trace system held by this rank)
};

// ---- vertex -> faces adjacency ------------------------------------------------------------------------------------
__global__ void mg_adj_fill(const int32_t* __restrict__ facenode, int64_t nface, int64_t node0, int32_t* __restrict__ vcnt,
                            int32_t* __restrict__ vface, int32_t* __restrict__ flags) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    const int64_t v1 = facenode[2 * f] + node0, v2 = facenode[2 * f + 1] + node0;
    const int64_t lo = min(v1, v2), hi = max(v1, v2);
    int k = atomicAdd(&vcnt[lo], 1);
    if (k < MG_MAXVAL) vface[lo * MG_MAXVAL + k] = int32_t(f); else atomicExch(&flags[FLAG_MG], 1);
    k = atomicAdd(&vcnt[hi], 1);
    if (k < MG_MAXVAL) vface[hi * MG_MAXVAL + k] = int32_t(uint32_t(f) | 0x80000000u); else atomicExch(&flags[FLAG_MG], 1);
}

// sorts the incident faces; fc[v] = 1 if the vertex touches a Dirichlet face, fc[nnode + v] = number of incident faces
// (doubles: summed over the ranks by an all-reduce when the mesh is distributed)
__global__ void mg_adj_sort(int64_t nnode, const int32_t* __restrict__ vcnt, int32_t* __restrict__ vface, const uint8_t* __restrict__ isbc,
                            double* __restrict__ fc) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    const int cnt = min(vcnt[v], MG_MAXVAL);
    int32_t a[MG_MAXVAL];
    bool fixed = false;
    for (int k = 0; k < cnt; ++k) {
        a[k] = vface[v * MG_MAXVAL + k];
        fixed = fixed || isbc[a[k] & 0x7fffffff];
    }
    for (int i = 1; i < cnt; ++i) {     // insertion sort by face id: the atomic append order is arbitrary
        const int32_t key = a[i];
        int j = i - 1;
        while (j >= 0 && (a[j] & 0x7fffffff) > (key & 0x7fffffff)) { a[j + 1] = a[j]; --j; }
        a[j + 1] = key;
    }
    for (int k = 0; k < cnt; ++k) vface[v * MG_MAXVAL + k] = a[k];
    fc[v] = fixed ? 1.0 : 0.0;
    fc[nnode + v] = double(cnt);
}
// a vertex carries no coarse unknown if it touches a Dirichlet face (on any rank) or has no face at all
__global__ void mg_adj_fix(int64_t nnode, int32_t* __restrict__ vcnt, const double* __restrict__ fc) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    if (fc[v] > 0.0 || fc[nnode + v] == 0.0) vcnt[v] = -1;
}

// ---- A_c = P'AP on the vertex grid, gathered per vertex ---------------------------------------------------------------
template <int NT>
__global__ void mg_vertex_operator(const double* __restrict__ Kd, const double* __restrict__ Ko, const int32_t* __restrict__ kcol,
                                   const uint8_t* __restrict__ isbc, const int32_t* __restrict__ facenode, int64_t node0,
                                   const int32_t* __restrict__ vcnt, const int32_t* __restrict__ vface, int px, int64_t nnode,
                                   double* __restrict__ st) {
    constexpr int NT2 = NT * NT;
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int cnt = vcnt[v];        // < 0: fixed vertex (identity row, written by mg_finalize_rows)
    const int vx = int(v % px), vy = int(v / px);
    for (int k = 0; k < cnt; ++k) {
        const int32_t e = vface[v * MG_MAXVAL + k];
        const int64_t f = e & 0x7fffffff;
        const double pf0 = 0.5, pf1 = (e < 0) ? MG_C1 : -MG_C1;       // coefficients of this vertex in face f
        for (int s = -1; s < 4; ++s) {
            const int64_t g = s < 0 ? f : kcol[4 * f + s];
            if (g < 0 || isbc[g]) continue;
            const double* blk = s < 0 ? Kd + f * NT2 : Ko + (f * 4 + s) * NT2;    // column-major: blk[b*NT + a] = K[a][b]
            // t_b = - sum_a pf_a K[a][b]   (A = -K on free rows), b = 0, 1
            const double t0 = -(pf0 * blk[0] + pf1 * blk[1]);
            const double t1 = -(pf0 * blk[NT] + pf1 * blk[NT + 1]);
            const int64_t g1 = facenode[2 * g] + node0, g2 = facenode[2 * g + 1] + node0;
            const int64_t lo = min(g1, g2), hi = max(g1, g2);
            if (vcnt[lo] >= 0) {
                const int sl = mg_slot(int(lo % px) - vx, int(lo / px) - vy);
                if (sl >= 0) acc[sl] += 0.5 * t0 - MG_C1 * t1;
            }
            if (vcnt[hi] >= 0) {
                const int sl = mg_slot(int(hi % px) - vx, int(hi / px) - vy);
                if (sl >= 0) acc[sl] += 0.5 * t0 + MG_C1 * t1;
            }
        }
    }
    for (int k = 0; k < 7; ++k) st[k * nnode + v] = acc[k];
}
// after the rows are complete (all-reduced over the ranks on a distributed mesh): inverse diagonal, identity rows at the
// fixed vertices and wherever the diagonal is not positive
__global__ void mg_finalize_rows(const int32_t* __restrict__ vcnt, int64_t nnode, double* __restrict__ st, double* __restrict__ dinv) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    const double d = st[v];
    if (vcnt[v] >= 0 && d > 0.0) { dinv[v] = 1.0 / d; return; }
    st[v] = 1.0;
    for (int k = 1; k < 7; ++k) st[k * nnode + v] = 0.0;
    dinv[v] = 0.0;
}

// ---- Galerkin coarse operator: A_c[I][J] = sum_p sum_q R[I,p] A[p,q] P[q,J], gathered per coarse point -----------------
__global__ void mg_rap(const double* __restrict__ stf, const double* __restrict__ dinvf, int px, int py, int cx, int cy,
                       double* __restrict__ stc, double* __restrict__ dinvc) {
    const int64_t nc = int64_t(cx) * cy, nf = int64_t(px) * py;
    int64_t I = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    const int Ix = int(I % cx), Iy = int(I / cx);
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool free_twin = dinvf[int64_t(2 * Iy) * px + 2 * Ix] != 0.0;
    if (free_twin) {
        for (int d = 0; d < 7; ++d) {
            const int fx = 2 * Ix + MG_DX[d], fy = 2 * Iy + MG_DY[d];
            if (fx < 0 || fy < 0 || fx >= px || fy >= py) continue;
            const int64_t p = int64_t(fy) * px + fx;
            if (dinvf[p] == 0.0) continue;
            const double wr = d == 0 ? 1.0 : 0.5;
            for (int e = 0; e < 7; ++e) {
                const int qx = fx + MG_DX[e], qy = fy + MG_DY[e];
                if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
                const int64_t q = int64_t(qy) * px + qx;
                if (dinvf[q] == 0.0) continue;
                const double aw = wr * stf[e * nf + p];
                if (aw == 0.0) continue;
                const int a2 = qx & 1, b2 = qy & 1, hx = qx >> 1, hy = qy >> 1;
                int jx[2], jy[2], cnt;
                double wp;
                if (!a2 && !b2) { jx[0] = hx; jy[0] = hy; cnt = 1; wp = 1.0; }
                else if (a2 && !b2) { jx[0] = hx; jy[0] = hy; jx[1] = hx + 1; jy[1] = hy; cnt = 2; wp = 0.5; }
                else if (!a2 && b2) { jx[0] = hx; jy[0] = hy; jx[1] = hx; jy[1] = hy + 1; cnt = 2; wp = 0.5; }
                else { jx[0] = hx + 1; jy[0] = hy; jx[1] = hx; jy[1] = hy + 1; cnt = 2; wp = 0.5; }
                for (int m = 0; m < cnt; ++m) {
                    if (jx[m] >= cx || jy[m] >= cy) continue;
                    if (dinvf[int64_t(2 * jy[m]) * px + 2 * jx[m]] == 0.0) continue;     // fixed coarse point
                    const int sl = mg_slot(jx[m] - Ix, jy[m] - Iy);
                    if (sl >= 0) acc[sl] += aw * wp;
                }
            }
        }
    }
    if (free_twin && acc[0] > 0.0) {
        for (int k = 0; k < 7; ++k) stc[k * nc + I] = acc[k];
        dinvc[I] = 1.0 / acc[0];
    } else {
        stc[I] = 1.0;
        for (int k = 1; k < 7; ++k) stc[k * nc + I] = 0.0;
        dinvc[I] = 0.0;
    }
}

// a coarse point whose fine twin is free but whose own diagonal vanished is fixed as well: drop the couplings to it
__global__ void mg_drop_fixed(double* __restrict__ st, const double* __restrict__ dinv, int px, int py) {
    const int64_t n = int64_t(px) * py;
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= n || dinv[p] == 0.0) return;
    const int x = int(p % px), y = int(p / px);
    for (int k = 1; k < 7; ++k) {
        const int qx = x + MG_DX[k], qy = y + MG_DY[k];
        if (qx < 0 || qy < 0 || qx >= px || qy >= py || dinv[int64_t(qy) * px + qx] == 0.0) st[k * n + p] = 0.0;
    }
}

// ---- V-cycle kernels --------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mg_apply_row(const double* __restrict__ st, const double* __restrict__ x, int64_t p, int px, int py, int64_t n) {
    const int ix = int(p % px), iy = int(p / px);
    double s = st[p] * x[p];
#pragma unroll
    for (int k = 1; k < 7; ++k) {
        const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
        if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
        s = fma(st[k * n + p], x[int64_t(qy) * px + qx], s);
    }
    return s;
}

// the same row product without __restrict__: inside mg_fused_vcycle the vectors are written and read within one launch, so the
// loads must not be routed through the non-coherent read-only path
__device__ __forceinline__ double mg_apply_row_rw(const double* st, const double* x, int64_t p, int px, int py, int64_t n) {
    const int ix = int(p % px), iy = int(p / px);
    double s = st[p] * x[p];
#pragma unroll
    for (int k = 1; k < 7; ++k) {
        const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
        if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
        s = fma(st[k * n + p], x[int64_t(qy) * px + qx], s);
    }
    return s;
}
__device__ __forceinline__ double mg_restrict_pt(const double* tf, int px, int py, const double* dinvc, int64_t I, int cx) {
    double s = 0.0;
    if (dinvc[I] != 0.0) {
        const int Ix = int(I % cx), Iy = int(I / cx);
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            const int fx = 2 * Ix + MG_DX[d], fy = 2 * Iy + MG_DY[d];
            if (fx < 0 || fy < 0 || fx >= px || fy >= py) continue;
            s += (d == 0 ? 1.0 : 0.5) * tf[int64_t(fy) * px + fx];
        }
    }
    return s;
}
__device__ __forceinline__ double mg_prolong_pt(const double* ec, int cx, int cy, int64_t p, int px) {
    const int ix = int(p % px), iy = int(p / px);
    const int a2 = ix & 1, b2 = iy & 1, hx = ix >> 1, hy = iy >> 1;
    auto get = [&](int jx, int jy) { return (jx < cx && jy < cy) ? ec[int64_t(jy) * cx + jx] : 0.0; };
    if (!a2 && !b2) return get(hx, hy);
    if (a2 && !b2) return 0.5 * (get(hx, hy) + get(hx + 1, hy));
    if (!a2 && b2) return 0.5 * (get(hx, hy) + get(hx, hy + 1));
    return 0.5 * (get(hx + 1, hy) + get(hx, hy + 1));
}

// x = omega Dinv r ; t = r - A x needs the neighbours of x, hence two kernels
__global__ void mg_smooth0(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ x, int64_t n) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < n) x[p] = MG_OMEGA * dinv[p] * r[p];
}
__global__ void mg_residual(const double* __restrict__ st, const double* __restrict__ r, const double* __restrict__ x, double* __restrict__ t,
                            int px, int py) {
    const int64_t n = int64_t(px) * py;
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < n) t[p] = r[p] - mg_apply_row(st, x, p, px, py, n);
}
// t = x + omega Dinv (r - A x)
__global__ void mg_smooth(const double* __restrict__ st, const double* __restrict__ dinv, const double* __restrict__ r,
                          const double* __restrict__ x, double* __restrict__ t, int px, int py) {
    const int64_t n = int64_t(px) * py;
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < n) t[p] = fma(MG_OMEGA * dinv[p], r[p] - mg_apply_row(st, x, p, px, py, n), x[p]);
}
// r_c = R t_f (transpose of the P1 interpolation), 0 at fixed coarse points
__global__ void mg_restrict(const double* __restrict__ tf, int px, int py, const double* __restrict__ dinvc, double* __restrict__ rc, int cx, int cy) {
    const int64_t nc = int64_t(cx) * cy;
    int64_t I = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (I < nc) rc[I] = mg_restrict_pt(tf, px, py, dinvc, I, cx);
}
// x_f += P e_c at the free fine points
__global__ void mg_prolong_add(const double* __restrict__ ec, int cx, int cy, const double* __restrict__ dinvf, double* __restrict__ x, int px, int py) {
    const int64_t n = int64_t(px) * py;
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= n || dinvf[p] == 0.0) return;
    x[p] += mg_prolong_pt(ec, cx, cy, p, px);
}

// coarsest grid: dense inverse by Gauss-Jordan without pivoting (SPD + identity rows), one block
__global__ void mg_dense_inverse(const double* __restrict__ st, int px, int py, double* __restrict__ ainv) {
    extern __shared__ double A[];      // n x n
    const int n = px * py;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[e] = 0.0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const int ix = p % px, iy = p / px;
        for (int k = 0; k < 7; ++k) {
            const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
            if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
            A[p * n + qy * px + qx] = st[k * n + p];
        }
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double piv = 1.0 / A[k * n + k];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) A[k * n + j] = (j == k) ? piv : A[k * n + j] * piv;
        __syncthreads();
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            if (i == k || j == k) continue;
            A[e] = fma(-A[i * n + k], A[k * n + j], A[e]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (i != k) A[i * n + k] = -A[i * n + k] * piv;
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) ainv[e] = A[e];
}
// ---- the coarse tail of the V-cycle in one kernel ----------------------------------------------------------------------
// The smallest levels are pure launch latency (five kernels for a microsecond of work each).  The levels with at most
// MG_FUSE_MAX points run inside one block: the same point-wise operations in the same order (results are bitwise those of the
// per-level kernels), separated by block barriers instead of kernel boundaries.  Measured (k=1, 1 M elements, same box):
// threshold 640 -> 0.297 ms per PCG iteration against 0.312 unfused; 2304 and 8192 are no faster than unfused (one SM cannot
// keep up with a few thousand points per stage), so the gain is small - inside a CUDA graph the tiny kernels cost ~1-2 us each.
struct MgFused {
    int nl;                                        // fused levels; lev 0 = finest fused, lev nl-1 = the dense one
    int px[MG_FUSE_LEVELS], py[MG_FUSE_LEVELS];
    double *st[MG_FUSE_LEVELS], *dinv[MG_FUSE_LEVELS], *r[MG_FUSE_LEVELS], *x[MG_FUSE_LEVELS], *t[MG_FUSE_LEVELS];
    const double* ainv;
};

__global__ void __launch_bounds__(1024) mg_fused_vcycle(const MgFused A) {
    const int T = blockDim.x, tid = threadIdx.x;
    for (int l = 0; l + 1 < A.nl; ++l) {
        const int px = A.px[l], py = A.py[l], cx = A.px[l + 1], cy = A.py[l + 1];
        const int64_t n = int64_t(px) * py, nc = int64_t(cx) * cy;
        for (int64_t p = tid; p < n; p += T) A.x[l][p] = MG_OMEGA * A.dinv[l][p] * A.r[l][p];
        __syncthreads();
        for (int64_t p = tid; p < n; p += T) A.t[l][p] = A.r[l][p] - mg_apply_row_rw(A.st[l], A.x[l], p, px, py, n);
        __syncthreads();
        for (int64_t I = tid; I < nc; I += T) A.r[l + 1][I] = mg_restrict_pt(A.t[l], px, py, A.dinv[l + 1], I, cx);
        __syncthreads();
    }
    {
        const int l = A.nl - 1, n = A.px[l] * A.py[l];
        for (int i = tid; i < n; i += T) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(A.ainv[i * n + j], A.r[l][j], s);
            A.t[l][i] = s;
        }
        __syncthreads();
    }
    for (int l = A.nl - 2; l >= 0; --l) {
        const int px = A.px[l], py = A.py[l], cx = A.px[l + 1], cy = A.py[l + 1];
        const int64_t n = int64_t(px) * py;
        for (int64_t p = tid; p < n; p += T)
            if (A.dinv[l][p] != 0.0) A.x[l][p] += mg_prolong_pt(A.t[l + 1], cx, cy, p, px);
        __syncthreads();
        for (int64_t p = tid; p < n; p += T)
            A.t[l][p] = fma(MG_OMEGA * A.dinv[l][p], A.r[l][p] - mg_apply_row_rw(A.st[l], A.x[l], p, px, py, n), A.x[l][p]);
        __syncthreads();
    }
}

// ---- transfers between the trace space and the vertex grid ----------------------------------------------------------------
template <int NT>
__global__ void mg_restrict_trace(const double* __restrict__ r, const int32_t* __restrict__ vcnt, const int32_t* __restrict__ vface, int64_t nnode,
                                  double* __restrict__ rc) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nnode) return;
    const int cnt = vcnt[v];
    double s = 0.0;
    for (int k = 0; k < cnt; ++k) {      // cnt < 0: fixed vertex
        const int32_t e = vface[v * MG_MAXVAL + k];
        const int64_t f = e & 0x7fffffff;
        s += 0.5 * r[f * NT] + ((e < 0) ? MG_C1 : -MG_C1) * r[f * NT + 1];
    }
    rc[v] = s;
}
// z += P e
template <int NT>
__global__ void __launch_bounds__(RB) mg_prolong_trace(const double* __restrict__ e, const int32_t* __restrict__ facenode, int64_t node0,
                                                       const uint8_t* __restrict__ isbc, int64_t nface, double* __restrict__ z) {
    for (int64_t f = int64_t(blockIdx.x) * RB + threadIdx.x; f < nface; f += int64_t(gridDim.x) * RB) {
        if (isbc[f]) continue;
        const int64_t v1 = facenode[2 * f] + node0, v2 = facenode[2 * f + 1] + node0;
        const double a = e[min(v1, v2)], b = e[max(v1, v2)];
        z[f * NT] += 0.5 * (a + b);
        z[f * NT + 1] += MG_C1 * (b - a);
    }
}
// scale = 0 on the ranks > 0 of a distributed mesh: the vertex vectors are replicated, their dot product counts once
__global__ void __launch_bounds__(RB) mg_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double scale, double* __restrict__ part) {
    double s = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * RB + threadIdx.x; i < n; i += int64_t(gridDim.x) * RB) s = fma(a[i], b[i], s);
    const double tot = block_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = scale * tot;
}

// ======================================================================================================================
// Meshes without grid structure (parse_mesh_triangle input, permuted node ids, Delaunay meshes): hierarchy-free vertex-space
// term  z += P C_m(A_c) P' r  with the kernels of hdg_mg_general.cuh (ELL vertex operator, Jacobi-scaled Chebyshev polynomial;
// bodies also checked on the CPU against scipy: tools/check_mg_general.py).  A fixed polynomial, so plain CG still applies.
// ======================================================================================================================
}  // namespace hdg
#include "hdg_mg_general.cuh"
namespace hdg {

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

static void mgx_free(hdg_context* c) {
    MgGeneral* g = static_cast<MgGeneral*>(c->mg_general);
    if (!g) return;
    void* ptrs[] = {g->vcnt, g->vface, g->nbr, g->val, g->diag, g->dinv, g->rc, g->x, g->res, g->d, g->lmax_bits};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete g;
    c->mg_general = nullptr;
}

static hdg_status mgx_setup(hdg_context* c) {
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

static hdg_status mgx_apply(hdg_context* c, const double* r, double* z, double* part, int np) {
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
    mg_dot<<<np, RB, 0, s>>>(g->rc, g->x, n, 1.0, part);
    mgx_prolong<<<(unsigned)ceil_div(c->nface_own, 256), 256, 0, s>>>(c->nface_own, NT, c->d_facenode, c->d_isbc, g->x, z);
    return HDG_OK;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static inline unsigned nblk(int64_t n, int b = 256) { return (unsigned)ceil_div(n, b); }

static void mgx_free(hdg_context* c);
void mg_free(hdg_context* c) {
    mgx_free(c);
    MgData* m = static_cast<MgData*>(c->mg);
    if (!m) return;
    if (m->pool) cudaFree(m->pool);
    if (m->ainv) cudaFree(m->ainv);
    if (m->vface) cudaFree(m->vface);
    if (m->vcnt) cudaFree(m->vcnt);
    delete m;
    c->mg = nullptr;
}

template <int NT> static hdg_status mg_setup_t(hdg_context* c) {
    const bool multi = comm_active(c);
    if (multi && c->comm->general_mesh)
        return set_err(c, HDG_ERR_INVALID, "on several GPUs the multigrid preconditioner needs hdg_set_rectangle_mesh (strip partition)");
    if (c->grid_px < 2 || c->grid_py < 2) return mgx_setup(c);        // no grid structure: hierarchy-free vertex term
    mgx_free(c);
    MgData* m = static_cast<MgData*>(c->mg);
    const int64_t nnode_g = c->grid_px * c->grid_py;       // the GLOBAL vertex grid (== the local one on a single GPU)
    if (m && (m->nnode != nnode_g || m->nface != c->nface || m->nx != c->grid_px - 1 || m->ny != c->grid_py - 1)) { mg_free(c); m = nullptr; }
    if (!m) {
        m = new MgData();
        c->mg = m;
        m->nnode = nnode_g; m->nface = c->nface; m->nx = int(c->grid_px) - 1; m->ny = int(c->grid_py) - 1;
        m->multi = multi;
        m->node0 = multi ? c->comm->j0 * c->grid_px : 0;    // local node ids are global ids minus j0 (nx+1), see Strip
        m->nface_rows = c->nface_own;
        // grid hierarchy
        int px = int(c->grid_px), py = int(c->grid_py);
        int64_t total = 0;
        while (true) {
            MgLevel& L = m->lev[m->nlev++];
            L.px = px; L.py = py; L.n = int64_t(px) * py;
            total += 11 * L.n;
            if (L.n <= MG_DENSE || std::min(px, py) < 3 || m->nlev == MG_MAXLEV) break;
            px = (px + 1) / 2; py = (py + 1) / 2;
        }
        const MgLevel& last = m->lev[m->nlev - 1];
        if (last.n > MG_DENSE) { mg_free(c); return set_err(c, HDG_ERR_INVALID, "mesh too anisotropic for the multigrid preconditioner"); }
        // fused tail: the levels with at most MG_FUSE_MAX points (always includes the dense one)
        m->lf = m->nlev - 1;
        int64_t fuse_max = MG_FUSE_MAX;
        if (const char* e = getenv("HDG_MG_FUSE_MAX")) fuse_max = atoll(e);      // tuning knob
        while (m->lf > 0 && m->lev[m->lf - 1].n <= fuse_max && m->nlev - (m->lf - 1) <= MG_FUSE_LEVELS) --m->lf;
        HDG_CUDA(c, cudaMalloc(&m->pool, sizeof(double) * total));
        double* q = m->pool;
        for (int l = 0; l < m->nlev; ++l) {
            MgLevel& L = m->lev[l];
            L.st = q; q += 7 * L.n;
            L.dinv = q; q += L.n;
            L.r = q; q += L.n;
            L.x = q; q += L.n;
            L.t = q; q += L.n;
        }
        HDG_CUDA(c, cudaMalloc(&m->ainv, sizeof(double) * last.n * last.n));
        HDG_CUDA(c, cudaMalloc(&m->vface, sizeof(int32_t) * nnode_g * MG_MAXVAL));
        HDG_CUDA(c, cudaMalloc(&m->vcnt, sizeof(int32_t) * nnode_g));
    }
    MgLevel& L0 = m->lev[0];
    if (!m->adjacency_ok) {
        // vertex -> OWNED faces (a face row is counted by exactly one rank); the "fixed" flags are global
        double* fc = L0.st;      // 2 n doubles of scratch (the stencil array is filled below)
        HDG_CUDA(c, cudaMemsetAsync(m->vcnt, 0, sizeof(int32_t) * nnode_g, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_MG, 0, sizeof(int32_t), c->stream));
        mg_adj_fill<<<nblk(m->nface_rows), 256, 0, c->stream>>>(c->d_facenode, m->nface_rows, m->node0, m->vcnt, m->vface, c->d_flags);
        mg_adj_sort<<<nblk(nnode_g), 256, 0, c->stream>>>(nnode_g, m->vcnt, m->vface, c->d_isbc, fc);
        if (multi) {
            hdg_status st = comm_allreduce_sum(c, fc, int(2 * nnode_g));
            if (st) return st;
        }
        mg_adj_fix<<<nblk(nnode_g), 256, 0, c->stream>>>(nnode_g, m->vcnt, fc);
        c->launches += 3;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->h_flags[FLAG_MG]) return set_err(c, HDG_ERR_INVALID, "a vertex has more than 8 faces");
        m->adjacency_ok = true;
    }
    // operators (every solve: the matrix may have changed)
    mg_vertex_operator<NT><<<nblk(L0.n), 256, 0, c->stream>>>(c->d_Kd, c->d_Ko, c->d_kcol, c->d_isbc, c->d_facenode, m->node0, m->vcnt,
                                                              m->vface, L0.px, L0.n, L0.st);
    if (multi) {     // rows of the faces of the other ranks
        hdg_status st = comm_allreduce_sum(c, L0.st, int(7 * L0.n));
        if (st) return st;
    }
    mg_finalize_rows<<<nblk(L0.n), 256, 0, c->stream>>>(m->vcnt, L0.n, L0.st, L0.dinv);
    c->launches += 2;
    for (int l = 0; l + 1 < m->nlev; ++l) {
        MgLevel &F = m->lev[l], &C = m->lev[l + 1];
        mg_rap<<<nblk(C.n), 256, 0, c->stream>>>(F.st, F.dinv, F.px, F.py, C.px, C.py, C.st, C.dinv);
        mg_drop_fixed<<<nblk(C.n), 256, 0, c->stream>>>(C.st, C.dinv, C.px, C.py);
        c->launches += 2;
    }
    const MgLevel& last = m->lev[m->nlev - 1];
    mg_dense_inverse<<<1, 256, sizeof(double) * last.n * last.n, c->stream>>>(last.st, last.px, last.py, m->ainv);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

hdg_status mg_setup(hdg_context* c) {
    switch (c->tab.nt) {
        case 2: return mg_setup_t<2>(c);
        case 3: return mg_setup_t<3>(c);
        case 4: return mg_setup_t<4>(c);
        case 5: return mg_setup_t<5>(c);
    }
    return set_err(c, HDG_ERR_INVALID, "unsupported order");
}

// z += P V(P' r);  part[0..np) = partial sums of (P' r) . V(P' r).  Enqueued on c->stream (capturable).
template <int NT> static hdg_status mg_apply_t(hdg_context* c, const double* r, double* z, double* part, int np) {
    MgData* m = static_cast<MgData*>(c->mg);
    cudaStream_t s = c->stream;
    MgLevel& L0 = m->lev[0];
    mg_restrict_trace<NT><<<nblk(L0.n), 256, 0, s>>>(r, m->vcnt, m->vface, L0.n, L0.r);
    if (m->multi) {     // P'r summed over the ranks; the V-cycle below is replicated
        hdg_status st = comm_allreduce_sum(c, L0.r, int(L0.n));
        if (st) return st;
    }
    const int nl = m->nlev, lf = m->lf;
    for (int l = 0; l < lf; ++l) {
        MgLevel &F = m->lev[l], &C = m->lev[l + 1];
        mg_smooth0<<<nblk(F.n), 256, 0, s>>>(F.dinv, F.r, F.x, F.n);
        mg_residual<<<nblk(F.n), 256, 0, s>>>(F.st, F.r, F.x, F.t, F.px, F.py);
        mg_restrict<<<nblk(C.n), 256, 0, s>>>(F.t, F.px, F.py, C.dinv, C.r, C.px, C.py);
    }
    {   // levels lf .. nl-1 in one block (mg_fused_vcycle); the solution of level lf ends up in lev[lf].t
        MgFused A{};
        A.nl = nl - lf;
        for (int l = lf; l < nl; ++l) {
            MgLevel& L = m->lev[l];
            const int k = l - lf;
            A.px[k] = L.px; A.py[k] = L.py; A.st[k] = L.st; A.dinv[k] = L.dinv; A.r[k] = L.r; A.x[k] = L.x; A.t[k] = L.t;
        }
        A.ainv = m->ainv;
        const int64_t nmax = m->lev[lf].n;
        const int threads = int(std::min<int64_t>(1024, std::max<int64_t>(64, (nmax + 31) / 32 * 32)));
        mg_fused_vcycle<<<1, threads, 0, s>>>(A);
    }
    for (int l = lf - 1; l >= 0; --l) {
        MgLevel &F = m->lev[l], &C = m->lev[l + 1];
        mg_prolong_add<<<nblk(F.n), 256, 0, s>>>(C.t, C.px, C.py, F.dinv, F.x, F.px, F.py);
        mg_smooth<<<nblk(F.n), 256, 0, s>>>(F.st, F.dinv, F.r, F.x, F.t, F.px, F.py);
    }
    mg_dot<<<np, RB, 0, s>>>(L0.r, L0.t, L0.n, (m->multi && c->comm->rank != 0) ? 0.0 : 1.0, part);
    mg_prolong_trace<NT><<<np, RB, 0, s>>>(L0.t, c->d_facenode, m->node0, c->d_isbc, c->nface_own, z);
    return HDG_OK;
}

hdg_status mg_apply(hdg_context* c, const double* r, double* z, double* part, int np) {
    if (c->mg_general) return mgx_apply(c, r, z, part, np);
    switch (c->tab.nt) {
        case 2: return mg_apply_t<2>(c, r, z, part, np);
        case 3: return mg_apply_t<3>(c, r, z, part, np);
        case 4: return mg_apply_t<4>(c, r, z, part, np);
        case 5: return mg_apply_t<5>(c, r, z, part, np);
    }
    return HDG_OK;
}

int mg_levels(const hdg_context* c) { return c->mg ? static_cast<const MgData*>(c->mg)->nlev : 0; }
// kernels one mg_apply enqueues: restrict_trace, fused tail, dot, prolong_trace + 5 per unfused level
int mg_launches_per_apply(const hdg_context* c) {
    if (c->mg_general) return 3 + 2 * MGX_CHEB_STEPS;
    return c->mg ? 4 + 5 * static_cast<const MgData*>(c->mg)->lf : 0;
}

}  // namespace hdg
