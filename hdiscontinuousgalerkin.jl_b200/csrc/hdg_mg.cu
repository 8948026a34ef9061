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
//
// ONE V-cycle = ONE persistent cooperative kernel (mg_vcycle_kernel): the stages are separated by grid barriers instead of
// kernel boundaries (the ~30 launches of 3-5 us each were 2/3 of a multigrid-PCG iteration), and the two sweeps of a level are
// fused into one stage each way - the smoothed iterate x = omega Dinv r (+ P e on the way up) is recomputed at the neighbours
// from r, Dinv (and the coarse correction) instead of being stored and re-read:
//     down, level l:  r_{l+1} = R (r_l - A_l (omega Dinv_l r_l))                     one stage, gathers over a 2-ring
//     up,   level l:  t_l = x + omega Dinv_l (r_l - A_l x),  x = omega Dinv_l r_l + P t_{l+1}     one stage
// Same operations in the same order as the unfused kernels, so the iterates are bitwise those of the launch-per-stage version.
//
// SEVERAL GPUs (strips of hdg_set_rectangle_mesh): the hierarchy is DISTRIBUTED.  A rank owns the vertex rows of its quad
// rows (level 0: rows [j0, j1), the last rank also row ny; level l+1: the rows Y with 2Y owned on level l).  Level arrays hold
// the owned rows plus TWO GHOST ROWS on each side - what a fused stage reads.  Ghost rows are PUSHED: the stage that produces a
// vector stores the first / last two owned rows also into the neighbouring ranks' arrays (plain stores over NVLink into CUDA IPC
// mappings), so every load of a stage is local and no stage waits on a remote load (a first version that read the neighbours'
// rows in place spent 30-40 us per stage in the few threads of the boundary rows).  The operators' ghost rows are copied once
// per solve.  P'r and the level-0 stencil rows of the vertex row shared by two strips are formed as partial sums by both ranks
// and added up by the owner.  The grid barrier between two stages then carries a barrier across the GPUs (xg_barrier_thread:
// mailbox flags over peer memory), executed by the last block to arrive.  Levels with <= MG_REP_MAX points (and all levels of a
// mesh too small for two rows per rank) are REPLICATED: the owners store their rows of the first such level into every rank's
// copy, and everything below runs redundantly on every GPU with local barriers only.  No NCCL call in the solve.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"
#include "hdg_xgpu.cuh"

namespace hdg {

// general-mesh term (hdg_mgx.cu)
hdg_status mgx_setup(hdg_context* c);
hdg_status mgx_apply(hdg_context* c, const double* r, double* z, double* part, int np);
void mgx_free(hdg_context* c);
void mgx_invalidate(hdg_context* c);
bool mgx_active(const hdg_context* c);
int mgx_launches_per_apply();

constexpr int MG_MAXVAL = 8;        // faces per vertex (6 on rectangle_mesh)
constexpr int MG_DENSE = 64;        // the coarsest grid has at most this many points
constexpr int MG_MAXLEV = 16;       // 2^16 vertices per direction
constexpr int MG_TAIL_MAX = 640;    // levels with at most this many points run inside ONE block (block barriers)
constexpr int MG_REP_MAX = 32768;   // several GPUs: levels with at most this many points are replicated on every rank
constexpr double MG_OMEGA = 0.8;    // damped Jacobi on the vertex grids
constexpr int MG_THREADS = 256;

// stencil slots: centre, E, W, N, S, SE, NW  (the neighbours of a vertex in the triangulation)
__device__ constexpr int MG_DX[7] = {0, 1, -1, 0, 0, 1, -1};
__device__ constexpr int MG_DY[7] = {0, 0, 0, 1, -1, -1, 1};
__device__ __forceinline__ int mg_slot(int dx, int dy) {
    if (dy == 0) return dx == 0 ? 0 : (dx == 1 ? 1 : (dx == -1 ? 2 : -1));
    if (dx == 0) return dy == 1 ? 3 : (dy == -1 ? 4 : -1);
    if (dx == 1 && dy == -1) return 5;
    if (dx == -1 && dy == 1) return 6;
    return -1;
}

constexpr int MG_GHOST = 2;         // ghost rows on each side of the owned rows of a distributed level

// One level as the kernels see it.  Rows [oy0, oy1) are owned by this rank; the arrays hold rows [rb, rb + n / px): on a
// distributed level the owned rows and MG_GHOST ghost rows on each side (clipped to the grid), on a replicated level all rows.
// b_* / a_*: the same vector on the rank below / above (distributed levels) - where the producer pushes its boundary rows.
struct LvDev {
    int px, py;
    int rb, oy0, oy1;
    int rep;
    int n;                      // stride between the arrays of the level (>= the stored points; levels have < 2^31 points)
    double *st, *dinv, *r, *t, *x;
    double *b_r, *b_t, *b_x;  int b_rb;
    double *a_r, *a_t, *a_x;  int a_rb;
};

enum MgArr : int { AR_R = 1, AR_T = 2, AR_X = 3 };

__device__ __forceinline__ int lv_idx(const LvDev& L, int x, int y) { return (y - L.rb) * L.px + x; }
// The dynamic vectors (r, t) are written by other SMs / GPUs inside the same launch, yet the stage loads are PLAIN (L1-cached)
// loads - the 7-point stencils of neighbouring threads re-read the same values, and served from L2 (ld.global.cg) those gathers
// were what bounded the big levels.  That is safe because every vector is written exactly once per launch (own rows by this
// GPU, ghost rows by the neighbours, all in the producing stage) and first read after the grid barrier that follows, so no SM
// can hold a stale line of it; L1 is invalidated between launches.  The one stage that reads a vector while the neighbours are
// still pushing its ghost rows (the shared-row addition at level 0) uses ld.global.cg.
__device__ __forceinline__ double lv_r(const LvDev& L, int x, int y) { return L.r[lv_idx(L, x, y)]; }
__device__ __forceinline__ double lv_t(const LvDev& L, int x, int y) { return L.t[lv_idx(L, x, y)]; }
// store into the own array and, for the first / last MG_GHOST owned rows of a distributed level, into the ghost rows of the neighbours
template <int W> __device__ __forceinline__ void lv_store(const LvDev& L, int x, int y, double v) {
    double* own = W == AR_R ? L.r : (W == AR_T ? L.t : L.x);
    own[lv_idx(L, x, y)] = v;
    double* b = W == AR_R ? L.b_r : (W == AR_T ? L.b_t : L.b_x);
    double* a = W == AR_R ? L.a_r : (W == AR_T ? L.a_t : L.a_x);
    if (b && y < L.oy0 + MG_GHOST) b[(y - L.b_rb) * L.px + x] = v;
    if (a && y >= L.oy1 - MG_GHOST) a[(y - L.a_rb) * L.px + x] = v;
}

// ---- the point-wise operations of the V-cycle ---------------------------------------------------------------------------
// Neighbours outside the grid are handled by MASKS, not branches: the address is clamped to the centre point and the value
// replaced by 0 (their stencil coefficients are 0 as well), which leaves every sum bit-identical to a skipping version and lets
// the loads of a whole stencil be in flight together - the persistent kernel runs 512 threads per SM and lives on memory-level
// parallelism inside a thread.

// x = omega Dinv r (first sweep from a zero start) at (qx, qy), 0 outside the grid; (cx, cy) is a point inside
__device__ __forceinline__ double mg_x0(const LvDev& L, int qx, int qy, int cx, int cy) {
    const bool in = (qx >= 0) & (qy >= 0) & (qx < L.px) & (qy < L.py);
    const int sx = in ? qx : cx, sy = in ? qy : cy;
    const double v = MG_OMEGA * L.dinv[lv_idx(L, sx, sy)] * lv_r(L, sx, sy);
    return in ? v : 0.0;
}

// P e at the fine point (x, y), e = t of the coarse level: the mean of two coarse values (twice the same one at a coarse twin)
__device__ __forceinline__ double mg_prolong_pt(const LvDev& C, int x, int y) {
    const int a2 = x & 1, b2 = y & 1, hx = x >> 1, hy = y >> 1;
    const int x1 = hx + (a2 & b2), y1 = hy, x2 = hx + (a2 & (b2 ^ 1)), y2 = hy + b2;
    const bool in1 = (x1 < C.px) & (y1 < C.py), in2 = (x2 < C.px) & (y2 < C.py);
    const double g1 = lv_t(C, in1 ? x1 : hx, in1 ? y1 : hy), g2 = lv_t(C, in2 ? x2 : hx, in2 ? y2 : hy);
    return 0.5 * ((in1 ? g1 : 0.0) + (in2 ? g2 : 0.0));
}
// x after the coarse correction at (qx, qy): omega Dinv r + P e at the free points; 0 outside the grid
__device__ __forceinline__ double mg_x1(const LvDev& F, const LvDev& C, int qx, int qy, int cx, int cy) {
    const bool in = (qx >= 0) & (qy >= 0) & (qx < F.px) & (qy < F.py);
    const int sx = in ? qx : cx, sy = in ? qy : cy;
    const double di = F.dinv[lv_idx(F, sx, sy)];
    const double pe = mg_prolong_pt(C, sx, sy);
    const double v = MG_OMEGA * di * lv_r(F, sx, sy) + (di != 0.0 ? pe : 0.0);
    return in ? v : 0.0;
}

// down: r_C(I) = sum_d w_d (r - A x0)(fine neighbour d of 2I), 0 at fixed coarse points.
// The 7 residuals need x0 on the 19 points of the 2-ring around 2I: loaded once into a 5 x 5 window.
__device__ __forceinline__ double mg_down_point(const LvDev& F, int Ix, int Iy) {
    const int cx = 2 * Ix, cy = 2 * Iy;
    double xw[5][5];
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx)
            if (dx + dy >= -2 && dx + dy <= 2) xw[dy + 2][dx + 2] = mg_x0(F, cx + dx, cy + dy, cx, cy);
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const int fx = cx + MG_DX[d], fy = cy + MG_DY[d];
        const bool in = (fx >= 0) & (fy >= 0) & (fx < F.px) & (fy < F.py);
        const int p = in ? lv_idx(F, fx, fy) : lv_idx(F, cx, cy);
        double a = F.st[p] * xw[MG_DY[d] + 2][MG_DX[d] + 2];
#pragma unroll
        for (int k = 1; k < 7; ++k) a = fma(F.st[k * int64_t(F.n) + p], xw[MG_DY[d] + MG_DY[k] + 2][MG_DX[d] + MG_DX[k] + 2], a);
        const double t = F.r[p] - a;
        s += in ? (d == 0 ? 1.0 : 0.5) * t : 0.0;
    }
    return s;
}
// up: t_F = x1 + omega Dinv (r - A x1)
__device__ __forceinline__ double mg_up_point(const LvDev& F, const LvDev& C, int x, int y) {
    double xq[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) xq[k] = mg_x1(F, C, x + MG_DX[k], y + MG_DY[k], x, y);
    const int p = lv_idx(F, x, y);
    double s = F.st[p] * xq[0];
#pragma unroll
    for (int k = 1; k < 7; ++k) s = fma(F.st[k * int64_t(F.n) + p], xq[k], s);
    return fma(MG_OMEGA * F.dinv[p], F.r[p] - s, xq[0]);
}

// ---- vertex <-> trace adjacency ---------------------------------------------------------------------------------------------
// vertex -> OWNED faces (a face row of K is counted by exactly one rank), local node ids (row - j0) px + x
__global__ void mg_adj_fill(const int32_t* __restrict__ facenode, int64_t nface, int32_t* __restrict__ vcnt,
                            int32_t* __restrict__ vface, int32_t* __restrict__ flags) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    const int64_t v1 = facenode[2 * f], v2 = facenode[2 * f + 1];
    const int64_t lo = min(v1, v2), hi = max(v1, v2);
    int k = atomicAdd(&vcnt[lo], 1);
    if (k < MG_MAXVAL) vface[lo * MG_MAXVAL + k] = int32_t(f); else atomicExch(&flags[FLAG_MG], 1);
    k = atomicAdd(&vcnt[hi], 1);
    if (k < MG_MAXVAL) vface[hi * MG_MAXVAL + k] = int32_t(uint32_t(f) | 0x80000000u); else atomicExch(&flags[FLAG_MG], 1);
}
// sorts the incident faces (the atomic append order is arbitrary); partial "touches a Dirichlet face" flag and face count of
// the local vertex rows into two level-0 scratch vectors - the row shared by two strips is summed by its owner afterwards
__global__ void mg_adj_sort(int64_t nv, int32_t* __restrict__ vcnt, int32_t* __restrict__ vface, const uint8_t* __restrict__ isbc,
                            double* __restrict__ flag, double* __restrict__ count) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int cnt = min(vcnt[v], MG_MAXVAL);
    int32_t a[MG_MAXVAL];
    bool fixed = false;
    for (int k = 0; k < cnt; ++k) {
        a[k] = vface[v * MG_MAXVAL + k];
        fixed = fixed || isbc[a[k] & 0x7fffffff];
    }
    for (int i = 1; i < cnt; ++i) {
        const int32_t key = a[i];
        int j = i - 1;
        while (j >= 0 && (a[j] & 0x7fffffff) > (key & 0x7fffffff)) { a[j + 1] = a[j]; --j; }
        a[j + 1] = key;
    }
    for (int k = 0; k < cnt; ++k) vface[v * MG_MAXVAL + k] = a[k];
    vcnt[v] = cnt;
    flag[v] = fixed ? 1.0 : 0.0;
    count[v] = double(cnt);
}
// the owner of the shared vertex row adds the partial sums the rank below formed in its spare row: narr arrays, one row each
__global__ void mg_add_shared_row(double* __restrict__ mine, int64_t stride, const double* __restrict__ below, int64_t bstride, int narr, int px) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= px) return;
    for (int k = 0; k < narr; ++k) mine[k * stride + i] += __ldcg(below + k * bstride + i);
}
// a vertex carries no coarse unknown (fx = 1) if it touches a Dirichlet face on any rank or has no face at all
__global__ void mg_fix_flags(int64_t cnt, const double* __restrict__ flag, const double* __restrict__ count, double* __restrict__ fx) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v < cnt) fx[v] = (flag[v] > 0.0 || count[v] == 0.0) ? 1.0 : 0.0;
}

// set-up, replicated levels: every rank copies the rows it does not own from their owners: narr arrays of stride n each
struct GatherArgs {
    const double* src[MAXR];      // the same array on every rank (replicated levels share one layout)
    int ry0[MAXR + 1];            // owned rows of rank q: [ry0[q], ry0[q+1])
    int nranks, rank;
};
__global__ void mg_gather_rows(const GatherArgs g, double* dst, int64_t n, int narr, int px) {
    const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, T = int64_t(gridDim.x) * blockDim.x;
    for (int q = 0; q < g.nranks; ++q) {
        if (q == g.rank) continue;
        const int64_t o = int64_t(g.ry0[q]) * px, cnt = int64_t(g.ry0[q + 1] - g.ry0[q]) * px;
        for (int k = 0; k < narr; ++k)
            for (int64_t i = tid; i < cnt; i += T) dst[k * n + o + i] = __ldcg(g.src[q] + k * n + o + i);
    }
}
// set-up, distributed levels: the ghost rows of narr arrays (stride n) from the owners' copies (stride bn / an, row base brb / arb)
__global__ void mg_pull_ghost_rows(double* dst, int64_t n, int rb, int oy0, int oy1, int rows, int narr, int px,
                                   const double* below, int64_t bn, int brb, const double* above, int64_t an, int arb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = (oy0 - rb) * px, na = (rb + rows - oy1) * px;      // ghost points below / above
    if (i < nb && below) {
        const int y = rb + i / px, x = i % px;
        for (int k = 0; k < narr; ++k) dst[k * n + (y - rb) * px + x] = __ldcg(below + k * bn + (y - brb) * px + x);
    } else if (i >= nb && i < nb + na && above) {
        const int y = oy1 + (i - nb) / px, x = (i - nb) % px;
        for (int k = 0; k < narr; ++k) dst[k * n + (y - rb) * px + x] = __ldcg(above + k * an + (y - arb) * px + x);
    }
}

// ---- A_c = P'AP on the vertex grid: partial rows from the owned faces of the local vertex rows ---------------------------
// fixed flags in the level-0 layout (ghost rows included)
struct FxView { const double* fx; int rb; };
__device__ __forceinline__ bool mg_fixed(const FxView& f, int px, int x, int y) { return f.fx[(y - f.rb) * px + x] != 0.0; }
template <int NT>
__global__ void mg_vertex_operator(const double* __restrict__ Kd, const double* __restrict__ Ko, const int32_t* __restrict__ kcol,
                                   const uint8_t* __restrict__ isbc, const int32_t* __restrict__ facenode, int node_row0,
                                   const int32_t* __restrict__ vcnt, const int32_t* __restrict__ vface, const FxView fxv, int px,
                                   int64_t nv, double* __restrict__ st, int64_t stride) {
    constexpr int NT2 = NT * NT;
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;      // local node id = (row - node_row0) px + x
    if (v >= nv) return;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int vx = int(v % px), vy = int(v / px) + node_row0;
    const int cnt = mg_fixed(fxv, px, vx, vy) ? 0 : vcnt[v];
    for (int k = 0; k < cnt; ++k) {
        const int32_t e = vface[v * MG_MAXVAL + k];
        const int64_t f = e & 0x7fffffff;
        const double pf0 = 0.5, pf1 = (e < 0) ? MG_C1 : -MG_C1;       // coefficients of this vertex in face f
        // the 5 block columns of row f (diagonal + 4 neighbours): all loads of all of them issued before the first use
        // (absent / Dirichlet columns read the diagonal block's addresses and are masked out)
        const int4 kc = *reinterpret_cast<const int4*>(kcol + 4 * f);
        const int cols[5] = {int(f), kc.x, kc.y, kc.z, kc.w};
        double t0[5], t1[5];
        int lox[5], loy[5], hix[5], hiy[5];
        bool ok[5], flo[5], fhi[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int64_t g = cols[s] < 0 ? f : cols[s];
            ok[s] = cols[s] >= 0 && !isbc[g];
            const double* blk = s == 0 ? Kd + f * NT2 : Ko + (f * 4 + (s - 1)) * NT2;    // column-major: blk[b*NT + a] = K[a][b]
            const double b00 = blk[0], b10 = blk[1], b01 = blk[NT], b11 = blk[NT + 1];
            // t_b = - sum_a pf_a K[a][b]   (A = -K on free rows), b = 0, 1
            t0[s] = -(pf0 * b00 + pf1 * b10);
            t1[s] = -(pf0 * b01 + pf1 * b11);
            const int2 gn = *reinterpret_cast<const int2*>(facenode + 2 * g);
            const int lo = min(gn.x, gn.y), hi = max(gn.x, gn.y);
            lox[s] = lo % px; loy[s] = lo / px + node_row0; hix[s] = hi % px; hiy[s] = hi / px + node_row0;
            flo[s] = mg_fixed(fxv, px, lox[s], loy[s]);
            fhi[s] = mg_fixed(fxv, px, hix[s], hiy[s]);
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (!ok[s]) continue;
            if (!flo[s]) {
                const int sl = mg_slot(lox[s] - vx, loy[s] - vy);
                if (sl >= 0) acc[sl] += 0.5 * t0[s] - MG_C1 * t1[s];
            }
            if (!fhi[s]) {
                const int sl = mg_slot(hix[s] - vx, hiy[s] - vy);
                if (sl >= 0) acc[sl] += 0.5 * t0[s] + MG_C1 * t1[s];
            }
        }
    }
    for (int k = 0; k < 7; ++k) st[k * stride + v] = acc[k];
}
// complete rows -> inverse diagonal; identity rows at the fixed vertices and wherever the diagonal is not positive
__global__ void mg_finalize_rows(const double* __restrict__ fx, int64_t cnt, double* __restrict__ st, int64_t stride, double* __restrict__ dinv) {
    int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= cnt) return;
    const double d = st[v];
    if (fx[v] == 0.0 && d > 0.0) { dinv[v] = 1.0 / d; return; }
    st[v] = 1.0;
    for (int k = 1; k < 7; ++k) st[k * stride + v] = 0.0;
    dinv[v] = 0.0;
}

// ---- Galerkin coarse operator: A_c[I][J] = sum_p sum_q R[I,p] A[p,q] P[q,J], gathered per coarse point; coarse rows [cy0, cy1).
// Like the stages of the V-cycle this is written for memory-level parallelism: the "is this point free" flags of the 2-ring
// around the fine twin come from ONE 5 x 5 window of Dinv (they also answer "is the coarse neighbour fixed": its twin lies in the
// window), and the 49 stencil coefficients are loaded unconditionally from clamped addresses, so nothing waits behind a branch
// (the first version took ~35 us per level whatever its size).  Same sums in the same order.
__global__ void mg_rap(const LvDev F, const LvDev C, int cy0, int cy1) {
    const int cnt = (cy1 - cy0) * C.px;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int Iy = cy0 + i / C.px, Ix = i - (Iy - cy0) * C.px;
    const int px = F.px, py = F.py, cx = C.px, cy = C.py;
    const int fx0 = 2 * Ix, fy0 = 2 * Iy;
    bool fr[5][5];      // free fine point (in the grid and Dinv != 0)
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            fr[dy + 2][dx + 2] = false;
            if (dx + dy >= -2 && dx + dy <= 2) {
                const int qx = fx0 + dx, qy = fy0 + dy;
                const bool in = (qx >= 0) & (qy >= 0) & (qx < px) & (qy < py);
                const double dv = F.dinv[in ? lv_idx(F, qx, qy) : lv_idx(F, fx0, fy0)];
                fr[dy + 2][dx + 2] = in & (dv != 0.0);
            }
        }
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool free_twin = fr[2][2];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const bool pfree = fr[MG_DY[d] + 2][MG_DX[d] + 2];
        const int p = pfree ? lv_idx(F, fx0 + MG_DX[d], fy0 + MG_DY[d]) : lv_idx(F, fx0, fy0);
        const double wr = d == 0 ? 1.0 : 0.5;
#pragma unroll
        for (int e = 0; e < 7; ++e) {
            const int ddx = MG_DX[d] + MG_DX[e], ddy = MG_DY[d] + MG_DY[e];      // q relative to the twin (compile time)
            const double aw = wr * F.st[e * int64_t(F.n) + p];
            if (!(free_twin && pfree && fr[ddy + 2][ddx + 2]) || aw == 0.0) continue;
            // P[q, J]: the coarse points q interpolates from, by the parity of q - the twin is even, so the parity of (ddx, ddy)
            const int a2 = ddx & 1, b2 = ddy & 1;
            const int hx = (ddx - a2) / 2, hy = (ddy - b2) / 2;                  // floor(dd / 2): coarse offset of q's lower-left twin
            int jx[2], jy[2], cn;
            double wp;
            if (!a2 && !b2) { jx[0] = hx; jy[0] = hy; cn = 1; wp = 1.0; }
            else if (a2 && !b2) { jx[0] = hx; jy[0] = hy; jx[1] = hx + 1; jy[1] = hy; cn = 2; wp = 0.5; }
            else if (!a2 && b2) { jx[0] = hx; jy[0] = hy; jx[1] = hx; jy[1] = hy + 1; cn = 2; wp = 0.5; }
            else { jx[0] = hx + 1; jy[0] = hy; jx[1] = hx; jy[1] = hy + 1; cn = 2; wp = 0.5; }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (m >= cn) continue;
                if (jx[m] < -1 || jx[m] > 1 || jy[m] < -1 || jy[m] > 1) continue;           // not a stencil neighbour of I
                if (Ix + jx[m] >= cx || Iy + jy[m] >= cy) continue;                          // (negative ones are not free: outside the grid)
                if (!fr[2 * jy[m] + 2][2 * jx[m] + 2]) continue;                             // fixed coarse point: its twin is not free
                const int sl = mg_slot(jx[m], jy[m]);
                if (sl >= 0) acc[sl] += aw * wp;
            }
        }
    }
    const int I = lv_idx(C, Ix, Iy);
    if (free_twin && acc[0] > 0.0) {
        for (int k = 0; k < 7; ++k) C.st[k * int64_t(C.n) + I] = acc[k];
        C.dinv[I] = 1.0 / acc[0];
    } else {
        C.st[I] = 1.0;
        for (int k = 1; k < 7; ++k) C.st[k * int64_t(C.n) + I] = 0.0;
        C.dinv[I] = 0.0;
    }
}
// a coarse point whose fine twin is free but whose own diagonal vanished is fixed as well: drop the couplings to it
__global__ void mg_drop_fixed(const LvDev C, int cy0, int cy1) {
    const int64_t cnt = int64_t(cy1 - cy0) * C.px;
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int x = int(i % C.px), y = cy0 + int(i / C.px);
    const int p = lv_idx(C, x, y);
    if (C.dinv[p] == 0.0) return;
    for (int k = 1; k < 7; ++k) {
        const int qx = x + MG_DX[k], qy = y + MG_DY[k];
        if (qx < 0 || qy < 0 || qx >= C.px || qy >= C.py || C.dinv[lv_idx(C, qx, qy)] == 0.0) C.st[k * int64_t(C.n) + p] = 0.0;
    }
}

// coarsest grid: dense inverse by Gauss-Jordan without pivoting (SPD + identity rows), one block
__global__ void mg_dense_inverse(const double* __restrict__ st, int64_t stride, int px, int py, double* __restrict__ ainv) {
    extern __shared__ double A[];      // n x n
    const int n = px * py;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[e] = 0.0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const int ix = p % px, iy = p / px;
        for (int k = 0; k < 7; ++k) {
            const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
            if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
            A[p * n + qy * px + qx] = st[k * stride + p];
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

// ---- the V-cycle kernel ---------------------------------------------------------------------------------------------------
struct VcArgs {
    int nlev;          // all levels; the last one is the dense one
    int lrep;          // several GPUs: levels [0, lrep) are distributed over the ranks, [lrep, nlev) replicated
    int lt;            // levels [lt, nlev) run inside block 0 (tail); lrep <= lt
    LvDev lev[MG_MAXLEV];
    const double* ainv;
    // trace side
    const double* r;   // trace residual (nt entries per owned face)
    double* z;         // z += P V(P'r)
    double* part;      // partial sums of (P'r).V(P'r), np entries
    int np;
    int64_t nface;     // owned faces
    const int32_t *vcnt, *vface, *facenode;
    const uint8_t* isbc;
    int node_row0;     // global vertex row of local node row 0 (j0 of the strip) == first owned row of level 0
    int prows;         // local vertex rows with (partial) P'r: the owned rows + the row shared with the rank above
    const double* fx;  // level-0 fixed flags of the owned rows
    double* above_x0;  // where the partial P'r of the row shared with the rank above goes: that row of its level-0 x vector
    // grid barrier + cross-GPU barrier
    unsigned* bar_count;
    volatile unsigned* bar_gen;
    XgComm xg;
    double* rep_r[MAXR];   // r of the first replicated level on every rank (its owners store their rows into all copies)
    unsigned long long* trace;   // measurement: globaltimer (ns) of block 0 at every barrier of the last V-cycle, [0] = count
    int fuse;          // 1: stop after the level-0 result t; the caller's direction update adds P t to z (pcg_dir_mg, hdg_solve.cu)
};

// all blocks of the grid; with `cross` the last block to arrive also runs the barrier across the GPUs before releasing
__device__ __forceinline__ unsigned long long mg_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mg_grid_barrier(const VcArgs& A, bool cross) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (A.trace && blockIdx.x == 0) { const unsigned long long k = A.trace[0] + 1; if (k < 62) { A.trace[k] = mg_now(); A.trace[0] = k; } }
        const unsigned g = *A.bar_gen;
        __threadfence();
        if (atomicAdd(A.bar_count, 1u) + 1u == gridDim.x) {
            *A.bar_count = 0u;
            if (cross && A.xg.nranks > 1) xg_barrier_thread(A.xg);
            __threadfence();
            *A.bar_gen = g + 1u;
        } else {
            while (*A.bar_gen == g) { }
            __threadfence();
        }
        if (A.trace && blockIdx.x == 0) { const unsigned long long k = A.trace[0] + 1; if (k < 62) { A.trace[k] = mg_now(); A.trace[0] = k; } }
    }
    __syncthreads();
}

// tail: the smallest levels inside one block, in SHARED memory (operators copied in at kernel start, while the other blocks
// form P'r), unfused sweeps with stored x and block barriers between them
struct TailLv { int px, py, n; double *st, *dinv, *r, *t, *x; };
__device__ __forceinline__ TailLv mg_tail_level(const VcArgs& A, double* sm, int l) {
    int o = 0;
    for (int k = A.lt; k < l; ++k) o += 11 * A.lev[k].px * A.lev[k].py;
    const int n = A.lev[l].px * A.lev[l].py;
    return TailLv{A.lev[l].px, A.lev[l].py, n, sm + o, sm + o + 7 * n, sm + o + 8 * n, sm + o + 9 * n, sm + o + 10 * n};
}
__device__ void mg_tail_load(const VcArgs& A, double* sm) {      // st (7 arrays) and dinv of every tail level
    for (int l = A.lt; l < A.nlev; ++l) {
        const TailLv L = mg_tail_level(A, sm, l);
        const double* __restrict__ src = A.lev[l].st;
        const int stride = A.lev[l].n, cnt = 8 * L.n;
        constexpr int U = 8;      // independent loads in flight per thread: one block fetches ~60 kB here
        for (int i0 = threadIdx.x; i0 < cnt; i0 += blockDim.x * U) {
            double v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < cnt ? src[(i / L.n) * int64_t(stride) + i % L.n] : 0.0; }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int i = i0 + u * blockDim.x; if (i < cnt) L.st[i] = v[u]; }
        }
    }
}
__device__ void mg_tail(const VcArgs& A, double* sm) {
    const int T = blockDim.x, tid = threadIdx.x;
    auto row = [](const TailLv& L, const double* x, int p) {
        const int iy = p / L.px, ix = p - iy * L.px;
        double s = L.st[p] * x[p];
#pragma unroll
        for (int k = 1; k < 7; ++k) {
            const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
            const bool in = (qx >= 0) & (qy >= 0) & (qx < L.px) & (qy < L.py);
            s = fma(L.st[k * L.n + p], in ? x[in ? qy * L.px + qx : p] : 0.0, s);
        }
        return s;
    };
    {
        const TailLv F = mg_tail_level(A, sm, A.lt);
        for (int p = tid; p < F.n; p += T) F.r[p] = __ldcg(A.lev[A.lt].r + p);
        __syncthreads();
    }
    for (int l = A.lt; l + 1 < A.nlev; ++l) {
        const TailLv F = mg_tail_level(A, sm, l), C = mg_tail_level(A, sm, l + 1);
        for (int p = tid; p < F.n; p += T) F.x[p] = MG_OMEGA * F.dinv[p] * F.r[p];
        __syncthreads();
        for (int p = tid; p < F.n; p += T) F.t[p] = F.r[p] - row(F, F.x, p);
        __syncthreads();
        for (int I = tid; I < C.n; I += T) {
            double s = 0.0;
            if (C.dinv[I] != 0.0) {
                const int Iy = I / C.px, Ix = I - Iy * C.px;
#pragma unroll
                for (int d = 0; d < 7; ++d) {
                    const int fx = 2 * Ix + MG_DX[d], fy = 2 * Iy + MG_DY[d];
                    if (fx < 0 || fy < 0 || fx >= F.px || fy >= F.py) continue;
                    s += (d == 0 ? 1.0 : 0.5) * F.t[fy * F.px + fx];
                }
            }
            C.r[I] = s;
        }
        __syncthreads();
    }
    {
        // dense coarsest solve t = Ainv r (n <= 64): one warp per row, lanes over the columns, fixed-order shuffle tree
        const TailLv L = mg_tail_level(A, sm, A.nlev - 1);
        const int lane = tid & 31, w = tid >> 5, nw = T >> 5;
        for (int i = w; i < L.n; i += nw) {
            double s = 0.0;
            for (int j = lane; j < L.n; j += 32) s = fma(A.ainv[i * L.n + j], L.r[j], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) L.t[i] = s;
        }
        __syncthreads();
    }
    for (int l = A.nlev - 2; l >= A.lt; --l) {
        const TailLv F = mg_tail_level(A, sm, l), C = mg_tail_level(A, sm, l + 1);
        for (int p = tid; p < F.n; p += T)
            if (F.dinv[p] != 0.0) {
                const int y = p / F.px, x = p - y * F.px;
                const int a2 = x & 1, b2 = y & 1, hx = x >> 1, hy = y >> 1;
                const int x1 = hx + (a2 & b2), y1 = hy, x2 = hx + (a2 & (b2 ^ 1)), y2 = hy + b2;
                const double g1 = (x1 < C.px && y1 < C.py) ? C.t[y1 * C.px + x1] : 0.0, g2 = (x2 < C.px && y2 < C.py) ? C.t[y2 * C.px + x2] : 0.0;
                F.x[p] += 0.5 * (g1 + g2);
            }
        __syncthreads();
        for (int p = tid; p < F.n; p += T) F.t[p] = fma(MG_OMEGA * F.dinv[p], F.r[p] - row(F, F.x, p), F.x[p]);
        __syncthreads();
    }
    {
        const TailLv F = mg_tail_level(A, sm, A.lt);
        for (int p = tid; p < F.n; p += T) A.lev[A.lt].t[p] = F.t[p];
    }
}

#ifndef MG_BLOCKS_PER_SM
#define MG_BLOCKS_PER_SM 2
#endif
template <int NT>
__global__ void __launch_bounds__(MG_THREADS, MG_BLOCKS_PER_SM) mg_vcycle_kernel(const VcArgs A) {
    const int T = int(gridDim.x * blockDim.x), tid = int(blockIdx.x * blockDim.x + threadIdx.x);
    const bool multi = A.xg.nranks > 1;
    const LvDev& L0 = A.lev[0];
    const int o0 = (L0.oy0 - L0.rb) * L0.px;            // first owned row == local node row 0 inside the level-0 arrays
    const int nown = (L0.oy1 - L0.oy0) * L0.px;
    extern __shared__ double mg_sm[];
    if (A.trace && tid == 0) { A.trace[0] = 1; A.trace[1] = mg_now(); }
    if (blockIdx.x == 0) mg_tail_load(A, mg_sm);
    // ---- P'r at the local vertex rows: sums over the OWNED faces, fixed vertices 0.  The row shared with the rank above holds
    // a partial sum: it goes to the owner (into its x vector, unused on a level that is not in the tail)
    {
        const int cnt = A.prows * L0.px;
        constexpr int UV = 2;      // vertices per trip: the adjacency rows of both are loaded first, then all 2 x 16 gathers are in flight together
        for (int v0 = tid; v0 < cnt; v0 += UV * T) {
            int c[UV];
            int32_t e[UV][MG_MAXVAL];
            double r0[UV][MG_MAXVAL], r1[UV][MG_MAXVAL];
#pragma unroll
            for (int u = 0; u < UV; ++u) {
                const int v = v0 + u * T < cnt ? v0 + u * T : v0;
                c[u] = A.vcnt[v];
                const int4 e0 = *reinterpret_cast<const int4*>(A.vface + int64_t(v) * MG_MAXVAL), e1 = *reinterpret_cast<const int4*>(A.vface + int64_t(v) * MG_MAXVAL + 4);
                e[u][0] = e0.x; e[u][1] = e0.y; e[u][2] = e0.z; e[u][3] = e0.w; e[u][4] = e1.x; e[u][5] = e1.y; e[u][6] = e1.z; e[u][7] = e1.w;
            }
#pragma unroll
            for (int u = 0; u < UV; ++u)
#pragma unroll
                for (int k = 0; k < MG_MAXVAL; ++k) {      // slots beyond the count read face 0
                    const int64_t f = k < c[u] ? (e[u][k] & 0x7fffffff) : 0;
                    if constexpr (NT % 2 == 0) {
                        const double2 q = *reinterpret_cast<const double2*>(A.r + f * NT);
                        r0[u][k] = q.x; r1[u][k] = q.y;
                    } else {
                        r0[u][k] = A.r[f * NT]; r1[u][k] = A.r[f * NT + 1];
                    }
                }
#pragma unroll
            for (int u = 0; u < UV; ++u) {
                const int v = v0 + u * T;
                if (v >= cnt) continue;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < MG_MAXVAL; ++k)
                    if (k < c[u]) s += 0.5 * r0[u][k] + ((e[u][k] < 0) ? MG_C1 : -MG_C1) * r1[u][k];
                if (v < nown) L0.r[o0 + v] = A.fx[v] != 0.0 ? 0.0 : s;
                else A.above_x0[v - nown] = s;
            }
        }
    }
    mg_grid_barrier(A, true);
    if (multi) {
        // the owner of a shared row adds the partial sums of the rank below; then the boundary rows of the completed vector go
        // to the neighbours' ghost rows (distributed level 0) or all owned rows to every rank (replicated level 0)
        const bool add = A.xg.rank > 0;      // row oy0 is shared with the rank below
        if (L0.rep) {
            for (int i = tid; i < nown; i += T) {
                double v = __ldcg(L0.r + o0 + i);
                if (add && i < L0.px) { v = A.fx[i] != 0.0 ? 0.0 : v + __ldcg(L0.x + o0 + i); L0.r[o0 + i] = v; }
                for (int q = 0; q < A.xg.nranks; ++q)
                    if (q != A.xg.rank) A.rep_r[q][o0 + i] = v;
            }
        } else {
            const int nb = min(MG_GHOST, L0.oy1 - L0.oy0) * L0.px;
            for (int i = tid; i < 2 * nb; i += T) {
                const int k = i < nb ? i : nown - 2 * nb + i;          // the first / the last MG_GHOST owned rows
                if (i >= nb && k < nb) continue;                        // fewer than 2 MG_GHOST owned rows: each row once
                const int y = L0.oy0 + k / L0.px, x = k % L0.px;
                double v = __ldcg(L0.r + o0 + k);
                if (add && k < L0.px) v = A.fx[k] != 0.0 ? 0.0 : v + __ldcg(L0.x + o0 + k);
                lv_store<AR_R>(L0, x, y, v);
            }
        }
        mg_grid_barrier(A, true);
    }
    // ---- down
    for (int l = 0; l < A.lt; ++l) {
        const LvDev &F = A.lev[l], &C = A.lev[l + 1];
        const bool own_rows = multi && l + 1 <= A.lrep;            // coarse level distributed, or the first replicated one
        const bool to_all = multi && l + 1 == A.lrep;
        const int cy0 = own_rows ? C.oy0 : 0, cy1 = own_rows ? C.oy1 : C.py;
        const int cnt = (cy1 - cy0) * C.px;
        for (int i = tid; i < cnt; i += T) {
            const int Iy = cy0 + i / C.px, Ix = i - (Iy - cy0) * C.px;
            const double s = C.dinv[lv_idx(C, Ix, Iy)] != 0.0 ? mg_down_point(F, Ix, Iy) : 0.0;
            if (to_all) {
                for (int q = 0; q < A.xg.nranks; ++q) A.rep_r[q][lv_idx(C, Ix, Iy)] = s;
            } else lv_store<AR_R>(C, Ix, Iy, s);
        }
        mg_grid_barrier(A, own_rows);
    }
    // ---- tail (block 0; the others wait at the barrier)
    if (blockIdx.x == 0) mg_tail(A, mg_sm);
    mg_grid_barrier(A, false);
    // ---- up.  The level-0 stage also accumulates (P'r).V(P'r) = sum r t over the owned rows: the thread that forms t(i) adds
    // r(i) t(i), in the order i = tid, tid + T, ... of the separate pass that used to follow the last barrier
    double dot = 0.0;
    const bool dot_in_up = A.lt > 0;
    for (int l = A.lt - 1; l >= 0; --l) {
        const LvDev &F = A.lev[l], &C = A.lev[l + 1];
        const bool dist = multi && l < A.lrep;
        const int fy0 = dist ? F.oy0 : 0, fy1 = dist ? F.oy1 : F.py;
        const int cnt = (fy1 - fy0) * F.px;
        const bool last = l == 0;
        for (int i = tid; i < cnt; i += T) {
            const int y = fy0 + i / F.px, x = i - (y - fy0) * F.px;
            const double tv = mg_up_point(F, C, x, y);
            lv_store<AR_T>(F, x, y, tv);
            if (last && y >= F.oy0 && y < F.oy1) dot = fma(F.r[lv_idx(F, x, y)], tv, dot);
        }
        // fused with the caller's direction update on one GPU: nothing in this launch reads t any more, the kernel boundary orders it
        if (!(last && A.fuse && !multi)) mg_grid_barrier(A, dist);
    }
    // ---- (P'r).V(P'r) over the owned rows, z += P t on the owned faces
    {
        double s = dot;
        if (!dot_in_up)
            for (int i = tid; i < nown; i += T) s = fma(__ldcg(L0.r + o0 + i), __ldcg(L0.t + o0 + i), s);
        const double tot = block_sum(s);
        if (threadIdx.x == 0) A.part[blockIdx.x] = tot;
        if (blockIdx.x == 0)
            for (int i = gridDim.x + threadIdx.x; i < A.np; i += blockDim.x) A.part[i] = 0.0;
        constexpr int U = 4;      // faces per trip: the vertex loads of all of them are issued before the first update of z
        for (int64_t f0 = tid; f0 < (A.fuse ? 0 : A.nface); f0 += int64_t(T) * U) {
            double a[U], b[U], z0[U], z1[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t f = f0 + int64_t(u) * T;
                ok[u] = f < A.nface;
                const int64_t fs = ok[u] ? f : 0;
                const int2 vv = *reinterpret_cast<const int2*>(A.facenode + 2 * fs);
                const int lo = min(vv.x, vv.y), hi = max(vv.x, vv.y);
                a[u] = __ldcg(L0.t + o0 + lo);          // local node ids index the level-0 arrays from the first owned row on
                b[u] = __ldcg(L0.t + o0 + hi);
                z0[u] = A.z[fs * NT]; z1[u] = A.z[fs * NT + 1];
                ok[u] = ok[u] && !A.isbc[fs];
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) {
                    const int64_t f = f0 + int64_t(u) * T;
                    A.z[f * NT] = fma(0.5, a[u] + b[u], z0[u]);
                    A.z[f * NT + 1] = fma(MG_C1, b[u] - a[u], z1[u]);
                }
        }
        if (A.trace && tid == 0) { const unsigned long long k = A.trace[0] + 1; if (k < 62) { A.trace[k] = mg_now(); A.trace[0] = k; } }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct MgLevelHost {
    int px = 0, py = 0;
    bool rep = false;
    int oy0[MAXR + 1] = {};          // owned rows of rank q: [oy0[q], oy0[q+1])
    int rb[MAXR] = {};               // first stored row on rank q
    int64_t off[MAXR] = {};          // offset of the level's arrays in rank q's pool (doubles)
    int64_t n[MAXR] = {};            // array stride on rank q: the stored points rounded up to a multiple of 16
};

struct MgData {
    int nlev = 0, lrep = 0, lt = 0;
    MgLevelHost lev[MG_MAXLEV];
    int R = 1, rank = 0;
    double* pool = nullptr;          // one allocation for all levels (+ the level-0 fixed flags behind them)
    int64_t pool_doubles = 0;
    int64_t fx_off[MAXR] = {};       // where the fixed flags start in rank q's pool
    void* peer_pool[MAXR] = {};      // the pools of all ranks, mapped (peer_pool[rank] = pool)
    double* ainv = nullptr;          // dense inverse of the coarsest operator
    int32_t* vface = nullptr;        // local nodes x MG_MAXVAL: incident owned faces ascending, bit 31 = the vertex is the face's hi vertex
    int32_t* vcnt = nullptr;
    unsigned* bar = nullptr;         // grid barrier words: count, generation
    unsigned long long* trace = nullptr;   // 64 words, only with HDG_MG_TRACE=1 (hdg_mg_trace)
    int64_t nnode_local = 0, nface = 0;
    int nx = 0, ny = 0;
    bool adjacency_ok = false;
    int grid_blocks = 0;
    size_t tail_smem = 0;            // shared memory of the V-cycle kernel: the tail levels
    VcArgs args{};                   // everything but r / z / part
};

static inline unsigned nblk(int64_t n, int b = 256) { return (unsigned)std::max<int64_t>(1, ceil_div(n, b)); }

// arrays of one level inside a pool: st (7 n) | dinv | r | t | x
static double* level_array(const MgData* m, int q, int l, int which) {
    return static_cast<double*>(m->peer_pool[q]) + m->lev[l].off[q] + which * m->lev[l].n[q];
}

static LvDev level_dev(const MgData* m, int l) {
    const MgLevelHost& H = m->lev[l];
    const int q = m->rank;
    LvDev L{};
    L.px = H.px; L.py = H.py; L.rep = H.rep ? 1 : 0;
    L.oy0 = H.oy0[q]; L.oy1 = H.oy0[q + 1];
    L.rb = H.rb[q];
    L.n = int(H.n[q]);
    L.st = level_array(m, q, l, 0); L.dinv = level_array(m, q, l, 7); L.r = level_array(m, q, l, 8);
    L.t = level_array(m, q, l, 9); L.x = level_array(m, q, l, 10);
    if (!H.rep && m->R > 1) {
        if (q > 0) { L.b_r = level_array(m, q - 1, l, 8); L.b_t = level_array(m, q - 1, l, 9); L.b_x = level_array(m, q - 1, l, 10); L.b_rb = H.rb[q - 1]; }
        if (q + 1 < m->R) { L.a_r = level_array(m, q + 1, l, 8); L.a_t = level_array(m, q + 1, l, 9); L.a_x = level_array(m, q + 1, l, 10); L.a_rb = H.rb[q + 1]; }
    }
    return L;
}

void mg_free(hdg_context* c) {
    mgx_free(c);
    MgData* m = static_cast<MgData*>(c->mg);
    if (!m) return;
    if (m->R > 1) comm_close_buffer(c, m->peer_pool);
    if (m->pool) cudaFree(m->pool);
    if (m->ainv) cudaFree(m->ainv);
    if (m->vface) cudaFree(m->vface);
    if (m->vcnt) cudaFree(m->vcnt);
    if (m->bar) cudaFree(m->bar);
    if (m->trace) cudaFree(m->trace);
    delete m;
    c->mg = nullptr;
}

// the mesh or the Dirichlet set changed: the buffers stay (if the sizes still fit at the next solve), the adjacency is rebuilt
void mg_invalidate(hdg_context* c) {
    mgx_invalidate(c);
    MgData* m = static_cast<MgData*>(c->mg);
    if (m) m->adjacency_ok = false;
}

static hdg_status xbarrier(hdg_context* c) {      // stream-ordered barrier across the GPUs (set-up only)
    if (!comm_active(c)) return HDG_OK;
    return comm_p2p_allreduce(c, c->d_partials, 1, 0u);
}

template <int NT> static hdg_status mg_setup_t(hdg_context* c) {
    const bool multi = comm_active(c);
    if (multi && c->comm->general_mesh)
        return set_err(c, HDG_ERR_INVALID, "on several GPUs the multigrid preconditioner needs hdg_set_rectangle_mesh (strip partition)");
    if (c->grid_px < 2 || c->grid_py < 2) return mgx_setup(c);        // no grid structure: hierarchy-free vertex term
    mgx_free(c);
    if (multi && !comm_p2p(c))
        return set_err(c, HDG_ERR_INVALID, "the distributed multigrid preconditioner needs peer access between the GPUs (CUDA IPC)");
    if (c->grid_px * c->grid_py >= (int64_t(1) << 31)) return set_err(c, HDG_ERR_INVALID, "vertex grid too large for 32-bit level indices");
    const int R = multi ? c->comm->nranks : 1, rank = multi ? c->comm->rank : 0;
    MgData* m = static_cast<MgData*>(c->mg);
    if (m && (m->nnode_local != c->nnode || m->nface != c->nface || m->nx != c->grid_px - 1 || m->ny != c->grid_py - 1 || m->R != R)) { mg_free(c); m = nullptr; }
    if (!m) {
        m = new MgData();
        c->mg = m;
        m->nnode_local = c->nnode; m->nface = c->nface; m->nx = int(c->grid_px) - 1; m->ny = int(c->grid_py) - 1;
        m->R = R; m->rank = rank;
        // grid hierarchy (global) and the rows every rank owns
        int px = int(c->grid_px), py = int(c->grid_py);
        while (true) {
            MgLevelHost& H = m->lev[m->nlev++];
            H.px = px; H.py = py;
            if (int64_t(px) * py <= MG_DENSE || std::min(px, py) < 3 || m->nlev == MG_MAXLEV) break;
            px = (px + 1) / 2; py = (py + 1) / 2;
        }
        const MgLevelHost& last = m->lev[m->nlev - 1];
        if (int64_t(last.px) * last.py > MG_DENSE) { mg_free(c); return set_err(c, HDG_ERR_INVALID, "mesh too anisotropic for the multigrid preconditioner"); }
        int64_t rep_max = MG_REP_MAX;
        if (const char* e = getenv("HDG_MG_REP_MAX")) rep_max = atoll(e);      // test / tuning knob
        const int64_t ny = m->ny;
        m->lrep = R > 1 ? m->nlev : 0;
        for (int l = 0; l < m->nlev; ++l) {
            MgLevelHost& H = m->lev[l];
            int minrows = H.py;
            for (int q = 0; q <= R; ++q) {
                if (l == 0) H.oy0[q] = q == R ? H.py : int(ny * q / R);         // strips of quad rows (mesh_rectangle); the last rank also owns row ny
                else H.oy0[q] = q == R ? H.py : (m->lev[l - 1].oy0[q] + 1) / 2;
                if (q > 0) minrows = std::min(minrows, H.oy0[q] - H.oy0[q - 1]);
            }
            // a distributed level needs MG_GHOST owned rows on every rank: ghost rows come from the immediate neighbours only
            if (R > 1 && m->lrep == m->nlev && (int64_t(H.px) * H.py <= rep_max || minrows < MG_GHOST)) m->lrep = l;
        }
        // tail: the levels with at most MG_TAIL_MAX points (always includes the dense one)
        m->lt = m->nlev - 1;
        int64_t tail_max = MG_TAIL_MAX;
        if (const char* e = getenv("HDG_MG_FUSE_MAX")) tail_max = std::min<int64_t>(MG_TAIL_MAX, atoll(e));      // the tail lives in shared memory
        while (m->lt > 0 && int64_t(m->lev[m->lt - 1].px) * m->lev[m->lt - 1].py <= tail_max) --m->lt;
        if (R > 1) m->lrep = std::min(m->lrep, m->lt);       // the tail runs on replicated levels
        // pool layout of every rank: replicated levels first (identical offsets everywhere), then the distributed ones
        for (int q = 0; q < R; ++q) {
            int64_t o = 0;
            for (int pass = 0; pass < 2; ++pass)
                for (int l = 0; l < m->nlev; ++l) {
                    MgLevelHost& H = m->lev[l];
                    H.rep = R == 1 || l >= m->lrep;
                    if ((pass == 0) != H.rep) continue;
                    H.rb[q] = H.rep ? 0 : std::max(0, H.oy0[q] - MG_GHOST);
                    const int re = H.rep ? H.py : std::min(H.py, H.oy0[q + 1] + MG_GHOST);
                    H.n[q] = (int64_t(re - H.rb[q]) * H.px + 15) / 16 * 16;      // array stride: every array starts on its own 128-byte line
                    H.off[q] = o;
                    o += 11 * H.n[q];
                }
            m->fx_off[q] = o;
            if (q == rank) m->pool_doubles = o + m->lev[0].n[q];
        }
        HDG_CUDA(c, cudaMalloc(&m->pool, sizeof(double) * m->pool_doubles));
        HDG_CUDA(c, cudaMemsetAsync(m->pool, 0, sizeof(double) * m->pool_doubles, c->stream));
        HDG_CUDA(c, cudaMalloc(&m->ainv, sizeof(double) * int64_t(last.px) * last.py * last.px * last.py));
        HDG_CUDA(c, cudaMalloc(&m->vface, sizeof(int32_t) * c->nnode * MG_MAXVAL));
        HDG_CUDA(c, cudaMalloc(&m->vcnt, sizeof(int32_t) * c->nnode));
        HDG_CUDA(c, cudaMalloc(&m->bar, sizeof(unsigned) * 2));
        HDG_CUDA(c, cudaMemsetAsync(m->bar, 0, sizeof(unsigned) * 2, c->stream));
        if (getenv("HDG_MG_TRACE")) {
            HDG_CUDA(c, cudaMalloc(&m->trace, sizeof(unsigned long long) * 64));
            HDG_CUDA(c, cudaMemsetAsync(m->trace, 0, sizeof(unsigned long long) * 64, c->stream));
        }
        m->peer_pool[rank] = m->pool;
        if (R > 1) {     // the neighbours' pools for the ghost rows, everybody's for the first replicated level (collective)
            HDG_CUDA(c, cudaStreamSynchronize(c->stream));
            hdg_status st = comm_share_buffer(c, m->pool, (1u << R) - 1u, m->peer_pool);
            if (st) return st;
        }
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        for (int l = m->lt; l < m->nlev; ++l) m->tail_smem += sizeof(double) * 11 * size_t(m->lev[l].px) * m->lev[l].py;
        if (m->tail_smem > 48 * 1024)
            HDG_CUDA(c, cudaFuncSetAttribute(mg_vcycle_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(m->tail_smem)));
        HDG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mg_vcycle_kernel<NT>, MG_THREADS, m->tail_smem));
        m->grid_blocks = sms * std::max(1, std::min(occ, MG_BLOCKS_PER_SM));
    }
    const int q = rank;
    const MgLevelHost& H0 = m->lev[0];
    const LvDev L0 = level_dev(m, 0);
    double* fx = m->pool + m->fx_off[q];                    // level-0 layout: rows [rb, ...)
    const int j0 = H0.oy0[q];                               // global vertex row of local node row 0
    const int own_rows = H0.oy0[q + 1] - H0.oy0[q];
    const int prows = own_rows + (q + 1 < R ? 1 : 0);       // + the row shared with the rank above (partial sums)
    const int64_t lo = int64_t(j0 - L0.rb) * L0.px;         // local node 0 inside the level-0 arrays
    const int64_t pcount = int64_t(prows) * L0.px, ocount = int64_t(own_rows) * L0.px;
    if (multi && j0 != int(c->comm->j0)) return set_err(c, HDG_ERR_INVALID, "internal: strip rows of the multigrid hierarchy and of the mesh differ");
    // row j0 in the arrays of the rank below: its copy of the shared row, holding its partial sums
    const int64_t bo = q > 0 ? int64_t(j0 - H0.rb[q - 1]) * H0.px : 0;
    hdg_status st = HDG_OK;
    // ghost rows of a distributed level (or all foreign rows of the first replicated one) of narr adjacent arrays starting at
    // array index `which`, after their owners have written them
    auto complete = [&](int l, int which, int narr) -> hdg_status {
        if (!multi) return HDG_OK;
        const MgLevelHost& H = m->lev[l];
        if (H.rep && l != m->lrep) return HDG_OK;            // computed redundantly by every rank
        hdg_status s2 = xbarrier(c);
        if (s2) return s2;
        double* dst = level_array(m, q, l, which);
        if (H.rep) {
            GatherArgs g{};
            g.nranks = R; g.rank = q;
            for (int k = 0; k <= R; ++k) g.ry0[k] = H.oy0[k];
            for (int k = 0; k < R; ++k) g.src[k] = level_array(m, k, l, which);
            mg_gather_rows<<<nblk(H.n[q]), 256, 0, c->stream>>>(g, dst, H.n[q], narr, H.px);
        } else {
            const int rows = (H.rep ? H.py : std::min(H.py, H.oy0[q + 1] + MG_GHOST)) - H.rb[q];
            const int ghosts = (H.oy0[q] - H.rb[q]) + (H.rb[q] + rows - H.oy0[q + 1]);
            mg_pull_ghost_rows<<<nblk(int64_t(ghosts) * H.px), 256, 0, c->stream>>>(
                dst, H.n[q], H.rb[q], H.oy0[q], H.oy0[q + 1], rows, narr, H.px,
                q > 0 ? level_array(m, q - 1, l, which) : nullptr, q > 0 ? H.n[q - 1] : 0, q > 0 ? H.rb[q - 1] : 0,
                q + 1 < R ? level_array(m, q + 1, l, which) : nullptr, q + 1 < R ? H.n[q + 1] : 0, q + 1 < R ? H.rb[q + 1] : 0);
        }
        c->launches += 1;
        return xbarrier(c);                                    // nobody overwrites a row another rank is still copying
    };
    if (!m->adjacency_ok) {
        HDG_CUDA(c, cudaMemsetAsync(m->vcnt, 0, sizeof(int32_t) * c->nnode, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_MG, 0, sizeof(int32_t), c->stream));
        mg_adj_fill<<<nblk(c->nface_own), 256, 0, c->stream>>>(c->d_facenode, c->nface_own, m->vcnt, m->vface, c->d_flags);
        // partial flags / counts of the local rows into the scratch vectors t / x of level 0 (adjacent arrays, stride n)
        mg_adj_sort<<<nblk(pcount), 256, 0, c->stream>>>(pcount, m->vcnt, m->vface, c->d_isbc, L0.t + lo, L0.x + lo);
        c->launches += 2;
        if (multi) {
            if ((st = xbarrier(c))) return st;
            if (q > 0) {
                mg_add_shared_row<<<nblk(H0.px), 256, 0, c->stream>>>(L0.t + lo, L0.n, level_array(m, q - 1, 0, 9) + bo, H0.n[q - 1], 2, H0.px);
                c->launches += 1;
            }
        }
        mg_fix_flags<<<nblk(ocount), 256, 0, c->stream>>>(ocount, L0.t + lo, L0.x + lo, fx + lo);
        c->launches += 1;
        if (multi) {      // the flags of the ghost rows (all rows on a replicated level 0): the array behind the levels, same layout
            if ((st = xbarrier(c))) return st;
            if (H0.rep) {
                GatherArgs g{};
                g.nranks = R; g.rank = q;
                for (int k = 0; k <= R; ++k) g.ry0[k] = H0.oy0[k];
                for (int k = 0; k < R; ++k) g.src[k] = static_cast<const double*>(m->peer_pool[k]) + m->fx_off[k];
                mg_gather_rows<<<nblk(H0.n[q]), 256, 0, c->stream>>>(g, fx, H0.n[q], 1, H0.px);
            } else {
                const int rows = std::min(H0.py, H0.oy0[q + 1] + MG_GHOST) - H0.rb[q];
                const int ghosts = (H0.oy0[q] - H0.rb[q]) + (H0.rb[q] + rows - H0.oy0[q + 1]);
                mg_pull_ghost_rows<<<nblk(int64_t(ghosts) * H0.px), 256, 0, c->stream>>>(
                    fx, H0.n[q], H0.rb[q], H0.oy0[q], H0.oy0[q + 1], rows, 1, H0.px,
                    q > 0 ? static_cast<const double*>(m->peer_pool[q - 1]) + m->fx_off[q - 1] : nullptr, 0, q > 0 ? H0.rb[q - 1] : 0,
                    q + 1 < R ? static_cast<const double*>(m->peer_pool[q + 1]) + m->fx_off[q + 1] : nullptr, 0, q + 1 < R ? H0.rb[q + 1] : 0);
            }
            c->launches += 1;
            if ((st = xbarrier(c))) return st;
        }
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->h_flags[FLAG_MG]) return set_err(c, HDG_ERR_INVALID, "a vertex has more than 8 faces");
        m->adjacency_ok = true;
    }
    // ---- operators (every solve: the matrix may have changed)
    const FxView fxv{fx, L0.rb};
    mg_vertex_operator<NT><<<nblk(pcount), 256, 0, c->stream>>>(c->d_Kd, c->d_Ko, c->d_kcol, c->d_isbc, c->d_facenode, j0, m->vcnt, m->vface,
                                                                fxv, L0.px, pcount, L0.st + lo, L0.n);
    c->launches += 1;
    if (multi) {
        if ((st = xbarrier(c))) return st;
        if (q > 0) {
            mg_add_shared_row<<<nblk(H0.px), 256, 0, c->stream>>>(L0.st + lo, L0.n, level_array(m, q - 1, 0, 0) + bo, H0.n[q - 1], 7, H0.px);
            c->launches += 1;
        }
    }
    mg_finalize_rows<<<nblk(ocount), 256, 0, c->stream>>>(fx + lo, ocount, L0.st + lo, L0.n, L0.dinv + lo);
    c->launches += 1;
    if ((st = complete(0, 0, 8))) return st;
    for (int l = 0; l + 1 < m->nlev; ++l) {
        const LvDev F = level_dev(m, l), C = level_dev(m, l + 1);
        const bool own = multi && l + 1 <= m->lrep;          // coarse rows computed by their owners
        const int cy0 = own ? C.oy0 : 0, cy1 = own ? C.oy1 : C.py;
        const int64_t cnt = int64_t(cy1 - cy0) * C.px;
        mg_rap<<<nblk(cnt), 256, 0, c->stream>>>(F, C, cy0, cy1);
        c->launches += 1;
        if (own && m->lev[l + 1].rep) {                      // first replicated level: collect all rows, then drop everywhere
            if ((st = complete(l + 1, 0, 8))) return st;
            mg_drop_fixed<<<nblk(C.n), 256, 0, c->stream>>>(C, 0, C.py);
        } else {
            if (own && (st = complete(l + 1, 7, 1))) return st;     // drop_fixed reads Dinv of the ghost rows
            mg_drop_fixed<<<nblk(cnt), 256, 0, c->stream>>>(C, cy0, cy1);
            if (own && (st = complete(l + 1, 0, 7))) return st;     // the stencils after the drop
        }
        c->launches += 1;
    }
    const LvDev last = level_dev(m, m->nlev - 1);
    mg_dense_inverse<<<1, 256, sizeof(double) * last.px * last.py * last.px * last.py, c->stream>>>(last.st, last.n, last.px, last.py, m->ainv);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    // ---- arguments of the V-cycle kernel
    VcArgs& A = m->args;
    A = VcArgs{};
    A.nlev = m->nlev; A.lrep = multi ? m->lrep : m->nlev; A.lt = m->lt;
    for (int l = 0; l < m->nlev; ++l) A.lev[l] = level_dev(m, l);
    A.ainv = m->ainv;
    A.nface = c->nface_own;
    A.vcnt = m->vcnt; A.vface = m->vface; A.facenode = c->d_facenode; A.isbc = c->d_isbc;
    A.node_row0 = j0; A.prows = prows;
    A.fx = fx + lo;
    if (q + 1 < R) A.above_x0 = level_array(m, q + 1, 0, 10) + int64_t(H0.oy0[q + 1] - H0.rb[q + 1]) * H0.px;      // its first owned row
    A.bar_count = m->bar; A.bar_gen = m->bar + 1;
    comm_xg(c, &A.xg);
    if (multi)
        for (int k = 0; k < R; ++k) A.rep_r[k] = level_array(m, k, m->lrep, 8);      // r of the first replicated level, everywhere
    A.trace = m->trace;
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

// z += P V(P' r);  part[0..np) = partial sums of (P' r) . V(P' r).  One cooperative launch on c->stream (capturable).
template <int NT> static hdg_status mg_apply_t(hdg_context* c, const double* r, double* z, double* part, int np, bool fuse) {
    MgData* m = static_cast<MgData*>(c->mg);
    VcArgs A = m->args;
    A.r = r; A.z = z; A.part = part; A.np = np; A.fuse = fuse ? 1 : 0;
    const int grid = std::max(1, std::min(m->grid_blocks, np));
    void* args[] = {&A};
    HDG_CUDA(c, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(mg_vcycle_kernel<NT>), dim3(grid), dim3(MG_THREADS), args, m->tail_smem, c->stream));
    return HDG_OK;
}

hdg_status mg_apply(hdg_context* c, const double* r, double* z, double* part, int np, bool fuse) {
    if (mgx_active(c)) return mgx_apply(c, r, z, part, np);
    switch (c->tab.nt) {
        case 2: return mg_apply_t<2>(c, r, z, part, np, fuse);
        case 3: return mg_apply_t<3>(c, r, z, part, np, fuse);
        case 4: return mg_apply_t<4>(c, r, z, part, np, fuse);
        case 5: return mg_apply_t<5>(c, r, z, part, np, fuse);
    }
    return HDG_OK;
}

bool mg_fused_ptrs(const hdg_context* c, const int32_t** facenode, const double** t0) {
    const MgData* m = static_cast<const MgData*>(c->mg);
    if (!m || mgx_active(c) || getenv("HDG_MG_NOFUSE")) return false;
    const LvDev& L0 = m->args.lev[0];
    *facenode = m->args.facenode;
    *t0 = L0.t + int64_t(L0.oy0 - L0.rb) * L0.px;      // local node ids index the level-0 arrays from the first owned row on
    return true;
}

// measurement (HDG_MG_TRACE=1): microseconds since the kernel start at every barrier entry / exit of the last V-cycle
int mg_trace(hdg_context* c, double* us, int cap) {
    MgData* m = static_cast<MgData*>(c->mg);
    if (!m || !m->trace) return 0;
    unsigned long long h[64];
    if (cudaMemcpy(h, m->trace, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    const int n = int(std::min<unsigned long long>(h[0], 62));
    for (int i = 0; i < n && i < cap; ++i) us[i] = double(h[1 + i] - h[1]) * 1e-3;
    return std::min(n, cap);
}

int mg_levels(const hdg_context* c) { return c->mg ? static_cast<const MgData*>(c->mg)->nlev : 0; }
int mg_launches_per_apply(const hdg_context* c) {
    if (mgx_active(c)) return mgx_launches_per_apply();
    return c->mg ? 1 : 0;
}

}  // namespace hdg
