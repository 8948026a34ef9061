// element_quad_kernel<K>: the fused blocks + condensation + scatter pass for orders k >= 2, FOUR threads per element.
//
// Same mathematics and the same outputs as element_schur_kernel<K> (hdg_element.cu; reference:
// examples/poisson2D_HDG.jl:77-184), different mapping.  At k >= 2 a thread per element needs 220-255 registers
// and 0.8-1.9 kB of shared memory, which leaves 8 warps per SM and long FP64 dependency chains (ncu: issue active
// 27-30 %, FP64 pipe 28-33 %).  The t+1 right-hand-side columns of an element are independent once
// S = C + B'A^-1 B is factored, so a block of 128 threads takes one 32-cell tile and splits every element 4 ways:
//   phase 0  (thread = cell tid%32, warp w = tid/32)  geometry; warp w forms rows w, w+4, ... of S and integrates the
//            load vector be over the quadrature points w, w+4, ...; all reference-matrix operands are warp-uniform
//            (constant bank); S and the partial load vectors go to shared memory;
//   phase 1  (thread = cell tid/4, lane q = tid%4)  S is factored L D L' by 4 adjacent lanes: lane q owns rows
//            q, q+4, ... and keeps them in registers, the pivot row is read from shared memory, pivots travel by
//            shuffle; L and D^-1 overwrite S in shared memory;
//   phase 2  (thread = cell tid%32, warp w = tid/32)  warp w solves the columns w, w+4, ... of [K_e | b_e] for the 32
//            cells of the tile, CB columns at a time so that one shared-memory load of an L entry feeds CB FMAs.  The
//            column index is warp-uniform, so reference-matrix operands come from the constant bank, and a store
//            instruction writes one 256-byte row segment of the tile.  (k = 3: the 12 face columns deal evenly; the
//            load-vector column is solved by every warp and its rows are split 4 ways, see SPLITB.)  sigma = A^-1(r1 + B u) is formed row by row,
//            stored, and folded into the 3 nt accumulators of the Ate column; products with structural zeros of
//            Tr, Ts, Fhat are skipped at compile time (hdg_sparsity.h, validated on the host against the tables);
//   phase 3  the face-diagonal blocks and rhs entries staged in shared memory are paired with the neighbour cell
//            of the same tile and stored (RED.ADD.F64 only when the neighbour is in another tile), warp w = face w.
#pragma once
#include <type_traits>

namespace hdg {

#ifndef QCB2
#define QCB2 1
#endif
#ifndef QCB3
#define QCB3 1
#endif
#ifndef QCB4
#define QCB4 1
#endif
#ifndef QMINB2
#define QMINB2 4
#endif
// Stage the off-diagonal trace blocks in shared memory and store them cooperatively in phase 3 (16-byte pieces of the
// two adjacent blocks a cell owns in a face row, consecutive threads -> consecutive pieces).  Pays when nt is not a
// multiple of 4: per-lane stores of nt doubles are then split into 8-byte STGs that each touch their own sector
// (k = 2, ncu: 263 M store sectors where 66 M would do, L1->L2 request path 70 % busy).  k = 3 writes whole 32-byte
// columns already.
#ifndef QSTAGE2
#define QSTAGE2 1
#endif
#ifndef QSTAGE4
#define QSTAGE4 0
#endif
// QINV: phase 1 inverts S in place (Gauss-Jordan without pivoting on the SPD matrix, rows dealt to the 4 lanes of a cell,
// pivot rows by shuffle) and phase 2 applies S^-1 as a symmetric matrix-vector product: n(n+1)/2 shared-memory loads and n^2
// INDEPENDENT FMAs per column instead of two triangular sweeps (n(n-1) loads, serial dependency chains of length 2n).
#ifndef QINV
#define QINV 0
#endif
#ifndef QMINB3
#define QMINB3 4
#endif
#ifndef QMINB4
#define QMINB4 3
#endif

// QFOLD (bit k-2 = order k): the Legendre-parity sign of a column on a reversed face is applied to its right-hand side (three
// geometry factors) instead of to every stored value.  Sign flips commute exactly with IEEE products and FMAs, so the output is
// bit-identical; the stores then read the solution registers directly instead of a temporary that the next product has to wait
// for (ncu: the WAR dependency on the store operand showed up as long-scoreboard stalls on the sign multiplications).
// Measured (4 M / 1 M cells): k=4 3.62 -> 3.49 ms; k=3 4.92 vs 4.94 ms and k=2 2.28 vs 2.32 ms (off there).
#ifndef QFOLD
#define QFOLD 4
#endif

// QPFk: tiles ahead whose cell records are prefetched into L2 by warp 0 of a block (0 = off).  The blocks of a launch start in index
// order, so the tile that some SM picks up next is about (resident blocks = 148 x 4) ahead of this one; the block that gets it then
// finds its first load in L2 (the cell records -> vertices -> geometry chain at block start is 8 % of the warp time, ncu).
// Measured (4 M / 1 M cells, 296 / 592 / 1184 tiles ahead): k=2 2.101 -> 2.013 / 2.013 / 2.049 ms; k=3 4.670 -> 4.685 / 4.745 / 4.629;
// k=4 3.285 -> 3.323 / 3.331 / 3.318  ->  on for k = 2 only.
#ifndef QPF2
#define QPF2 592
#endif
#ifndef QPF3
#define QPF3 0
#endif
#ifndef QPF4
#define QPF4 0
#endif

template <int K> struct QuadCfg {
    static constexpr int n = Ord<K>::n, nt = Ord<K>::nt, t = Ord<K>::t;
    static constexpr int G = 4;                       // threads per element
    static constexpr int cells = 32;                  // cells per block = one Ke tile
    static constexpr int threads = G * cells;
    static constexpr int CS = 33;                     // stride of one entry row in shared memory (odd: the 4-lanes-per-cell phase spreads over the banks)
    static constexpr int R = (n + G - 1) / G;         // rows of S per lane (phase 1)
    static constexpr int CC = (t + 1 + G - 1) / G;    // columns per warp (phase 2)
    static constexpr int CB = K == 2 ? QCB2 : (K == 3 ? QCB3 : QCB4);   // columns per batch
#ifndef QSPLITB
#define QSPLITB 0
#endif
    // t = 0 mod 4 (k = 3): the t face columns deal evenly to the 4 warps and the load-vector column t would make warp 0
    // carry 4 columns against 3; instead every warp solves it and takes the rows i = w mod 4 of sigma / K_e / bte
    static constexpr bool SPLITB = QSPLITB && (t % 4 == 0) && CB == 1;
    static constexpr int NB = SPLITB ? t / 4 + 1 : (CC + CB - 1) / CB;
    static constexpr int nL = n * (n - 1) / 2;
    // shared-memory record per cell, stored [entry][cell]
    static constexpr int o_L = 0;                     // strict lower triangle of S, overwritten by L
    static constexpr int o_dinv = o_L + nL;           // diagonal of S, overwritten by D^-1
    static constexpr int o_status = o_dinv + n;
    static constexpr int o_diag = o_status + 1;       // staging of the face-diagonal blocks (phases 2-3); partial be sums (phases 0-1)
    static constexpr int o_rhs = o_diag + 3 * nt * nt;
    // the load vector be shares the rhs staging entries: the warp that solves the load-vector column reads be into registers at
    // the start of that column and stages bte (3 nt >= n entries) at its end, lane = cell both times
    static constexpr int o_be = o_rhs;
    static_assert(3 * nt >= n && 3 * nt * nt >= 4 * n && !SPLITB, "be aliases the rhs staging entries, behind the four partial load vectors");
    static constexpr int o_scr = o_rhs + (SPLITB ? 4 : 1) * 3 * nt;      // phase 2: solutions of the 2nd ... CB-th column of a batch, per warp
    static constexpr bool STAGE_OFF = (K == 2 && QSTAGE2) || (K == 4 && QSTAGE4);
    static constexpr bool INV = ((QINV >> (K - 2)) & 1) != 0;          // bit 0: k = 2, bit 1: k = 3, bit 2: k = 4
    static constexpr int o_off = o_scr + 4 * (CB - 1) * n;             // staged off-diagonal blocks: [(lp*2 + s)*nt*nt + j*nt + ip]
    static constexpr int entries = o_off + (STAGE_OFF ? 6 * nt * nt : 0);
    static_assert(3 * nt * nt + 3 * nt >= 4 * n, "partial load vectors alias the staging area");
    static constexpr size_t smem_rec = sizeof(double) * entries * CS;
    static constexpr size_t smem = smem_rec + (STAGE_OFF ? sizeof(uint32_t) * 3 * cells : 0);   // + face words of the tile
    static constexpr int min_blocks = K == 2 ? QMINB2 : (K == 3 ? QMINB3 : QMINB4);
    static constexpr bool col_sweep = K <= 3;         // ordering of the triangular sweeps, see phase 2
    static constexpr int PF = K == 2 ? QPF2 : (K == 3 ? QPF3 : QPF4);
    static constexpr bool FOLD = ((QFOLD >> (K - 2)) & 1) != 0;
};

// phase 0, rows i = W, W+4, ... of S = C + B'A^-1 B = tau sum_l |wn_l| Chat_l + detJ (al Prr + be Prs + ga Pss); W is a
// template parameter so that every reference-matrix operand has a compile-time constant-bank address
template <int K, int W>
__device__ __forceinline__ void form_S_rows(double* __restrict__ sm, const double cf0, const double cf1, const double cf2,
                                            const double al, const double be, const double ga) {
    using Q = QuadCfg<K>;
    constexpr int n = Q::n, CS = Q::CS;
    const DevTables<K>& T = ctab<K>();
#pragma unroll
    for (int i = W; i < n; i += 4)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = cf0 * T.Chat[(0 * n + i) * n + j];
            s = fma(cf1, T.Chat[(1 * n + i) * n + j], s);
            s = fma(cf2, T.Chat[(2 * n + i) * n + j], s);
            s = fma(al, T.Prr[i * n + j], s);
            s = fma(be, T.Prs[i * n + j], s);
            s = fma(ga, T.Pss[i * n + j], s);
            sm[(i == j ? Q::o_dinv + j : Q::o_L + tri(i, j)) * CS] = s;
        }
}

template <int K>
__global__ void __launch_bounds__(QuadCfg<K>::threads, QuadCfg<K>::min_blocks) element_quad_kernel(const ElemArgs a) {
    using Q = QuadCfg<K>;
    using Sp = Sparsity<K>;
    constexpr int n = Q::n, nt = Q::nt, t = Q::t, ke = Ord<K>::ke, CS = Q::CS, R = Q::R, CB = Q::CB;
    constexpr int64_t nt2 = nt * nt;
    const DevTables<K>& T = ctab<K>();          // warp-uniform operands (constant bank)
    extern __shared__ double smem[];
    const bool dbg = a.dbg_At != nullptr;
    const int64_t tile0 = a.cell_begin + int64_t(blockIdx.x) * Q::cells;
    const int w = threadIdx.x >> 5, ci = threadIdx.x & 31;
    double* const sm = smem + ci;               // record of cell ci: sm[entry * CS]
    const int64_t c = tile0 + ci;
    bool active = c < a.cell_end;
    const int64_t cl = active ? c : a.cell_end - 1;      // padding cells compute on the last cell; nothing of theirs is stored
    const double tau = a.tau;

    // =========================== phase 0: lane = cell, warp w = rows / quadrature points w, w+4, ... ========
    if constexpr (Q::PF > 0) {
        const int64_t cn = c + int64_t(Q::PF) * Q::cells;
        if (w == 0 && cn < a.cell_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.cellinfo + CI * cn));
    }
    CellGeom g;
    load_geometry(a, cl, g);
    const double cf0 = tau * g.dJf[0], cf1 = tau * g.dJf[1], cf2 = tau * g.dJf[2];
    {
        const double al = g.detJ * (g.G00 * g.G00 + g.G01 * g.G01);
        const double be = g.detJ * (g.G00 * g.G10 + g.G01 * g.G11);
        const double ga = g.detJ * (g.G10 * g.G10 + g.G11 * g.G11);
        switch (w) {
            case 0: form_S_rows<K, 0>(sm, cf0, cf1, cf2, al, be, ga); break;
            case 1: form_S_rows<K, 1>(sm, cf0, cf1, cf2, al, be, ga); break;
            case 2: form_S_rows<K, 2>(sm, cf0, cf1, cf2, al, be, ga); break;
            default: form_S_rows<K, 3>(sm, cf0, cf1, cf2, al, be, ga); break;
        }
        // partial load vector: detJ sum_{q = w, w+4, ...} w_q f(x_q) N[i,q]   (poisson2D_HDG.jl:106-114)
        double bev[n];
#pragma unroll
        for (int i = 0; i < n; ++i) bev[i] = 0.0;
        for (int qq = w; qq < a.nq; qq += 4) {
            double fv;
            if (a.source_id == 0) fv = a.fq[cl * a.nq + qq];
            else {
                const double m0 = T.Mgeo[3 * qq], m1 = T.Mgeo[3 * qq + 1], m2 = T.Mgeo[3 * qq + 2];
                const double xq = m0 * g.x[0][0] + m1 * g.x[1][0] + m2 * g.x[2][0];
                const double yq = m0 * g.x[0][1] + m1 * g.x[1][1] + m2 * g.x[2][1];
                fv = source_value(a.source_id, xq, yq);
            }
#pragma unroll
            for (int i = 0; i < n; ++i) bev[i] = fma(T.WN[qq * n + i], fv, bev[i]);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) sm[(Q::o_diag + w * n + i) * CS] = bev[i] * g.detJ;
    }
    __syncthreads();

    // =========================== phase 1: 4 lanes per cell, L D L' of S ======================================
    {
        const int q = threadIdx.x & 3, pc = threadIdx.x >> 2;
        const unsigned lane = threadIdx.x & 31u;
        double* const sp = smem + pc;           // record of cell pc
        bool spd = true;
        if constexpr (Q::INV) {
        // lane q owns rows q, q+4, ... of the full symmetric matrix, in registers
        double A[R][n];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = q + 4 * r;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                A[r][j] = 0.0;
                if (i < n) A[r][j] = sp[(i == j ? Q::o_dinv + j : (i > j ? Q::o_L + i * (i - 1) / 2 + j : Q::o_L + j * (j - 1) / 2 + i)) * CS];
            }
        }
#pragma unroll
        for (int k = 0; k < n; ++k) {
            double prow[n];
#pragma unroll
            for (int j = 0; j < n; ++j) prow[j] = __shfl_sync(0xffffffffu, A[k >> 2][j], (lane & ~3u) | unsigned(k & 3));
            const double dk = prow[k];
            spd = spd && (dk > 0.0);
            const double dinv = 1.0 / dk;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (4 * r >= n) continue;
                const bool own = (r == (k >> 2)) && (q == (k & 3));       // this lane holds the pivot row in slot r
                const double f = A[r][k] * dinv;
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    if (j == k) continue;
                    A[r][j] = own ? prow[j] * dinv : fma(-f, prow[j], A[r][j]);
                }
                A[r][k] = own ? dinv : -f;
            }
        }
        // S^-1 (lower triangle + diagonal) overwrites S
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = q + 4 * r;
            if (i < n) {
#pragma unroll
                for (int j = 0; j < n; ++j)
                    if (j <= i) sp[(i == j ? Q::o_dinv + j : Q::o_L + i * (i - 1) / 2 + j) * CS] = A[r][j];
            }
        }
        __syncwarp();
        } else {
        // lane q owns rows q, q+4, ...; its rows of L stay in registers, the pivot row is read from shared memory
        double Lr[R][n], dd[n];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            double ljd[n];
#pragma unroll
            for (int k = 0; k < j; ++k) ljd[k] = sp[(Q::o_L + tri(j, k)) * CS] * dd[k];     // L[j][k] d_k
            double colr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                colr[r] = 1.0;
                if (4 * r + 3 < j || 4 * r >= n) continue;         // compile-time: no row of this slot is in [j, n)
                const int i = q + 4 * r;
                if (i >= j && i < n) {
                    double s = sp[(i == j ? Q::o_dinv + j : Q::o_L + i * (i - 1) / 2 + j) * CS];
#pragma unroll
                    for (int k = 0; k < j; ++k) s = fma(-Lr[r][k], ljd[k], s);
                    colr[r] = s;
                }
            }
            const double dj = __shfl_sync(0xffffffffu, colr[j >> 2], (lane & ~3u) | unsigned(j & 3));
            spd = spd && (dj > 0.0);
            dd[j] = dj;
            const double dinv = 1.0 / dj;
            if (q == (j & 3)) sp[(Q::o_dinv + j) * CS] = dinv;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                Lr[r][j] = 0.0;
                if (4 * r + 3 <= j || 4 * r >= n) continue;
                const int i = q + 4 * r;
                if (i > j && i < n) {
                    Lr[r][j] = colr[r] * dinv;
                    sp[(Q::o_L + i * (i - 1) / 2 + j) * CS] = Lr[r][j];
                }
            }
            __syncwarp();
        }
        }
        if (q == 0) sp[Q::o_status * CS] = spd ? 1.0 : -1.0;
    }
    // be = sum of the 4 partial vectors (fixed order), lane = cell again, warp w takes rows w, w+4, ...: with the 4-lanes-per-cell
    // mapping of the factorisation these loads were 4-way bank conflicts (ncu: 45 M of the kernel's 66 M excess wavefronts at k = 3)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = w + 4 * r;
        if (i < n) {
            const double* pp = sm + (Q::o_diag + i) * CS;
            sm[(Q::o_be + i) * CS] = ((pp[0] + pp[n * CS]) + pp[2 * n * CS]) + pp[3 * n * CS];
        }
    }
    __syncthreads();

    // =========================== phase 2: lane = cell, warp = column group =================================
    if (active && !g.ok) {
        if (w == 0) atomicCAS(&a.flags[FLAG_BAD_GEOM], 0, int32_t(c + 1));
        active = false;
    }
    if (active && sm[Q::o_status * CS] < 0.0) {
        if (w == 0) atomicCAS(&a.flags[FLAG_SINGULAR], 0, int32_t(c + 1));
        active = false;
    }
    const bool o0 = g.v[2] > g.v[1], o1 = g.v[0] > g.v[2], o2 = g.v[1] > g.v[0];   // face_orientation, src/mesh.jl:51-54
    const double idet = 1.0 / g.detJ;
    double* __restrict__ const Ke_tile = a.Ke + ((c >> 5) * ke) * 32 + (c & 31);
    const volatile double* const smv = sm;      // volatile: keeps the L loads next to their uses (hoisted they spill)

#pragma unroll 1
    for (int b = 0; b < Q::NB; ++b) {
        const bool bpart = Q::SPLITB && b == Q::NB - 1;          // the shared load-vector column: rows i = w mod 4 only
        const int col0 = bpart ? t : w + 4 * CB * b;             // columns col0, col0 + 4, ... (warp-uniform)
        if (col0 > t) break;
        double u[CB][n];
        // right-hand sides: [-E;F] columns reduced to the u-block  (B'A^-1 E_l = ca_l Qr_l + cb_l Qs_l), or be
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
            const int col = col0 + 4 * cb;
            if (col < t) {
                const int l = col / nt;
                const double wnx = l == 0 ? g.wn[0][0] : (l == 1 ? g.wn[1][0] : g.wn[2][0]);
                const double wny = l == 0 ? g.wn[0][1] : (l == 1 ? g.wn[1][1] : g.wn[2][1]);
                double cf = l == 0 ? cf0 : (l == 1 ? cf1 : cf2);
                double ca_l = g.G00 * wnx + g.G01 * wny, cb_l = g.G10 * wnx + g.G11 * wny;
                if constexpr (Q::FOLD) {         // Legendre parity of a reversed face, applied to the whole column through its right-hand side
                    const bool o_l = l == 0 ? o0 : (l == 1 ? o1 : o2);
                    if (!o_l && ((col - l * nt) & 1)) { cf = -cf; ca_l = -ca_l; cb_l = -cb_l; }
                }
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    double r = cf * T.RT[(col * n + i) * 4];
                    r = fma(ca_l, T.RT[(col * n + i) * 4 + 1], r);
                    u[cb][i] = fma(cb_l, T.RT[(col * n + i) * 4 + 2], r);
                }
            } else if (col == t) {
#pragma unroll
                for (int i = 0; i < n; ++i) u[cb][i] = sm[(Q::o_be + i) * CS];
            } else {
#pragma unroll
                for (int i = 0; i < n; ++i) u[cb][i] = 0.0;
            }
        }
        // S u = r  by L D L'; one load of every L entry serves the CB columns of the batch.  Two orderings of the same
        // operations: column-oriented (axpy: once u[k] is final the updates of u[k+1..] are independent FMAs) and
        // row-oriented (dot products).  Measured per order (4 M / 1 M elements): k=2 2.35 vs 2.43 ms, k=3 4.91 vs 4.93 ms,
        // k=4 4.00 vs 3.62 ms  ->  column-oriented for k <= 3.
        if constexpr (Q::INV) {
            double x[CB][n];
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                for (int i = 0; i < n; ++i) x[cb][i] = 0.0;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const double ajj = smv[(Q::o_dinv + j) * CS];
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) x[cb][j] = fma(ajj, u[cb][j], x[cb][j]);
#pragma unroll
                for (int i = 0; i < j; ++i) {
                    const double aji = smv[(Q::o_L + tri(j, i)) * CS];
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        x[cb][j] = fma(aji, u[cb][i], x[cb][j]);
                        x[cb][i] = fma(aji, u[cb][j], x[cb][i]);
                    }
                }
            }
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                for (int i = 0; i < n; ++i) u[cb][i] = x[cb][i];
        } else {
        if constexpr (Q::col_sweep) {
#pragma unroll
            for (int k = 0; k < n - 1; ++k) {
                double lk[n];
#pragma unroll
                for (int i = k + 1; i < n; ++i) lk[i] = smv[(Q::o_L + tri(i, k)) * CS];
#pragma unroll
                for (int i = k + 1; i < n; ++i)
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) u[cb][i] = fma(-lk[i], u[cb][k], u[cb][i]);
            }
        } else {
#pragma unroll
            for (int i = 1; i < n; ++i)
#pragma unroll
                for (int k = 0; k < i; ++k) {
                    const double lik = smv[(Q::o_L + tri(i, k)) * CS];
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) u[cb][i] = fma(-lik, u[cb][k], u[cb][i]);
                }
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const double di = smv[(Q::o_dinv + i) * CS];
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) u[cb][i] *= di;
        }
        if constexpr (Q::col_sweep) {
#pragma unroll
            for (int k = n - 1; k > 0; --k) {
                double lk[n];
#pragma unroll
                for (int i = 0; i < k; ++i) lk[i] = smv[(Q::o_L + tri(k, i)) * CS];
#pragma unroll
                for (int i = 0; i < k; ++i)
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) u[cb][i] = fma(-lk[i], u[cb][k], u[cb][i]);
            }
        } else {
#pragma unroll
            for (int i = n - 2; i >= 0; --i)
#pragma unroll
                for (int k = i + 1; k < n; ++k) {
                    const double lki = smv[(Q::o_L + tri(k, i)) * CS];
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) u[cb][i] = fma(-lki, u[cb][k], u[cb][i]);
                }
        }

        }
        // per column: sigma = A^-1 (r1 + B u) row by row; column of Ate = [E;F]'K_e - He accumulated on the fly.
        // The columns of a batch are processed one after the other by ONE copy of the code (the solutions of the 2nd,
        // 3rd ... column wait in a per-thread shared-memory scratch), so the register need is that of a single column.
        double uc[n];
#pragma unroll
        for (int i = 0; i < n; ++i) uc[i] = u[0][i];
        double* const scr = sm + (Q::o_scr + w * (CB - 1) * n) * CS;
#pragma unroll
        for (int cb = 1; cb < CB; ++cb)
#pragma unroll
            for (int i = 0; i < n; ++i) scr[((cb - 1) * n + i) * CS] = u[cb][i];
        // one column: generic lambda so that the row-split variant of the load-vector column (BP = true_type) is a second,
        // separately scheduled copy of the code instead of a branch inside every row of the common one
        auto column = [&](auto BP, const int col) {
            constexpr bool kBP = decltype(BP)::value;
                const bool isb = col == t;
                const int l = isb ? 0 : col / nt;
                const int j = isb ? 0 : col - l * nt;
                const bool o_l = l == 0 ? o0 : (l == 1 ? o1 : o2);
                const double scol = (isb || o_l || !(j & 1)) ? 1.0 : -1.0;      // Legendre parity of a reversed face
                constexpr bool FOLD = Q::FOLD;                                  // ... already in the solution (see the right-hand sides)
                const double dJf_l = l == 0 ? g.dJf[0] : (l == 1 ? g.dJf[1] : g.dJf[2]);
                const double wnx_l = l == 0 ? g.wn[0][0] : (l == 1 ? g.wn[1][0] : g.wn[2][0]);
                const double wny_l = l == 0 ? g.wn[0][1] : (l == 1 ? g.wn[1][1] : g.wn[2][1]);
                const double ex = isb ? 0.0 : (FOLD ? scol : 1.0) * (wnx_l * idet), ey = isb ? 0.0 : (FOLD ? scol : 1.0) * (wny_l * idet);
                const int mcol = isb ? 0 : col;
                double val[3][nt];
    #pragma unroll
                for (int lp = 0; lp < 3; ++lp)
    #pragma unroll
                    for (int ip = 0; ip < nt; ++ip) val[lp][ip] = 0.0;
                double* __restrict__ const Kp = Ke_tile + int64_t(col) * 32;
    #pragma unroll
                for (int i = 0; i < n; ++i) {
                    if (kBP && (i & 3) != w) continue;
                    double p = 0.0, s = 0.0;
    #pragma unroll
                    for (int k = 0; k < n; ++k) {
                        if ((Sp::tr(i) >> k) & 1u) p = fma(T.Tr[i * n + k], uc[k], p);
                        if ((Sp::ts(i) >> k) & 1u) s = fma(T.Ts[i * n + k], uc[k], s);
                    }
                    const double mf = T.RT[(mcol * n + i) * 4 + 3];
                    const double sx = fma(g.G00, p, fma(g.G10, s, -ex * mf));
                    const double sy = fma(g.G01, p, fma(g.G11, s, -ey * mf));
                    if (active && !dbg) {               // 256-byte row segments of the tile
                        Kp[int64_t(i * (t + 1)) * 32] = FOLD ? sx : scol * sx;
                        Kp[int64_t((n + i) * (t + 1)) * 32] = FOLD ? sy : scol * sy;
                        Kp[int64_t((2 * n + i) * (t + 1)) * 32] = FOLD ? uc[i] : scol * uc[i];
                    }
                    const double w0 = fma(g.wn[0][0], sx, fma(g.wn[0][1], sy, cf0 * uc[i]));
                    const double w1 = fma(g.wn[1][0], sx, fma(g.wn[1][1], sy, cf1 * uc[i]));
                    const double w2 = fma(g.wn[2][0], sx, fma(g.wn[2][1], sy, cf2 * uc[i]));
    #pragma unroll
                    for (int ip = 0; ip < nt; ++ip) {
                        if ((Sp::fh(i) >> (0 * nt + ip)) & 1u) val[0][ip] = fma(T.Fhat[i * t + 0 * nt + ip], w0, val[0][ip]);
                        if ((Sp::fh(i) >> (1 * nt + ip)) & 1u) val[1][ip] = fma(T.Fhat[i * t + 1 * nt + ip], w1, val[1][ip]);
                        if ((Sp::fh(i) >> (2 * nt + ip)) & 1u) val[2][ip] = fma(T.Fhat[i * t + 2 * nt + ip], w2, val[2][ip]);
                    }
                }
    #pragma unroll
                for (int lp = 0; lp < 3; ++lp) {
                    const bool o_lp = lp == 0 ? o0 : (lp == 1 ? o1 : o2);
                    double v[nt];
    #pragma unroll
                    for (int ip = 0; ip < nt; ++ip) {
                        const double srow = (o_lp || !(ip & 1)) ? 1.0 : -1.0;
                        v[ip] = val[lp][ip] * (FOLD ? srow : srow * scol);
                    }
                    if (isb) {                                   // bte = -[E;F]' b_e
    #pragma unroll
                        for (int ip = 0; ip < nt; ++ip) {
                            if (dbg && !Q::SPLITB) { if (active) a.dbg_bt[lp * nt + ip] = -v[ip]; }
                            else sm[(Q::o_rhs + (Q::SPLITB ? w * 3 * nt : 0) + lp * nt + ip) * CS] = -v[ip];
                        }
                    } else if (lp == l) {                        // face-diagonal block, minus He (poisson2D_HDG.jl:144-151)
    #pragma unroll
                        for (int ip = 0; ip < nt; ++ip) {
                            const double h = fma(-dJf_l, T.Hhat[ip * nt + j], v[ip]);
                            if (dbg) { if (active) a.dbg_At[col * t + lp * nt + ip] = h; }
                            else sm[(Q::o_diag + (lp * nt + j) * nt + ip) * CS] = h;
                        }
                    } else if (dbg) {
    #pragma unroll
                        for (int ip = 0; ip < nt; ++ip) if (active) a.dbg_At[col * t + lp * nt + ip] = v[ip];
                    } else if (Q::STAGE_OFF) {                   // off-diagonal block, staged: stored cooperatively in phase 3
                        const int s = (l - lp + 3) % 3 - 1;
    #pragma unroll
                        for (int ip = 0; ip < nt; ++ip) sm[(Q::o_off + ((lp * 2 + s) * nt + j) * nt + ip) * CS] = v[ip];
                    } else if (active) {                         // off-diagonal block: exactly one contributing cell
                        const int s = (l - lp + 3) % 3 - 1;
                        const int64_t f_lp = g.f[lp] & 0x7fffffffu;
                        const int slot = int(g.f[lp] >> 31) * 2 + s;
                        store_vec<nt>(a.Ko + (f_lp * 4 + slot) * nt2 + j * nt, v);
                    }
                }
        };
#pragma unroll 1
        for (int cb = 0; cb < CB; ++cb) {
            const int col = col0 + 4 * cb;
            if (col > t) break;
            if (cb > 0) {
#pragma unroll
                for (int i = 0; i < n; ++i) uc[i] = scr[((cb - 1) * n + i) * CS];
            }
            if (bpart) column(std::true_type{}, col);
            else column(std::false_type{}, col);
        }
    }
    if (dbg) {
        if constexpr (Q::SPLITB) {      // bte = sum of the 4 partial vectors
            __syncthreads();
            if (w == 0 && active)
                for (int e = 0; e < 3 * nt; ++e)
                    a.dbg_bt[e] = ((sm[(Q::o_rhs + e) * CS] + sm[(Q::o_rhs + 3 * nt + e) * CS]) + sm[(Q::o_rhs + 6 * nt + e) * CS]) + sm[(Q::o_rhs + 9 * nt + e) * CS];
        }
        return;
    }

    // =========================== phase 3: scatter of the staged blocks, warp w = local face w =================
    if constexpr (Q::STAGE_OFF) {
        // face words of the tile, so that any thread can address any cell's face rows (0xffffffff: nothing to store)
        uint32_t* const fword = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem) + Q::smem_rec);
        if (w < 3) fword[3 * ci + w] = active ? (w == 0 ? g.f[0] : (w == 1 ? g.f[1] : g.f[2])) : 0xffffffffu;
        __syncthreads();
        // A cell owns, in the row of each of its faces, the two adjacent blocks of its side: 2 nt^2 doubles, 16-byte aligned
        constexpr int PC = nt * nt;                        // 16-byte pieces per (cell, face) chunk
        for (int idx = threadIdx.x; idx < 3 * Q::cells * PC; idx += Q::threads) {
            const int chunk = idx / PC, piece = idx - chunk * PC;
            const int cell = chunk / 3, lp = chunk - 3 * cell;
            const uint32_t fw = fword[chunk];
            if (fw == 0xffffffffu) continue;
            const double* const src = smem + cell + (Q::o_off + lp * 2 * PC + 2 * piece) * CS;
            double* const dst = a.Ko + (int64_t(fw & 0x7fffffffu) * 4 + 2 * int64_t(fw >> 31)) * nt2 + 2 * piece;
            *reinterpret_cast<double2*>(dst) = make_double2(src[0], src[CS]);
        }
    } else {
        __syncthreads();
    }
    if (!active || w == 3) return;
    {
        const int l = w;
        const uint32_t fl = l == 0 ? g.f[0] : (l == 1 ? g.f[1] : g.f[2]);
        const int64_t f = fl & 0x7fffffffu;
        const uint32_t sec = fl >> 31;
        const uint32_t pb = (uint32_t(g.partner) >> (8 * l)) & 0xffu;
        double* const kd = a.Kd + f * nt2;
        double* const rh = a.rhs + f * nt;
        const double* const sd = sm + (Q::o_diag + l * nt2) * CS;
        const double* const sr = sm + (Q::o_rhs + l * nt) * CS;
        auto rsum = [&](const double* base, int e) -> double {          // bte entry: one value, or the 4 per-warp partial sums in fixed order
            if constexpr (Q::SPLITB) return ((base[e * CS] + base[(3 * nt + e) * CS]) + base[(6 * nt + e) * CS]) + base[(9 * nt + e) * CS];
            else return base[e * CS];
        };
        if (pb & 0x80u) {
            if (!sec) {
                const int pl = int(pb & 31u), plf = int((pb >> 5) & 3u);
                const double* const pd = smem + pl + (Q::o_diag + plf * nt2) * CS;     // partner cell's staged block
                const double* const pr = smem + pl + (Q::o_rhs + plf * nt) * CS;
                double v[nt2], r[nt];
#pragma unroll
                for (int e = 0; e < nt2; ++e) v[e] = sd[e * CS] + pd[e * CS];
#pragma unroll
                for (int e = 0; e < nt; ++e) r[e] = rsum(sr, e) + rsum(pr, e);
                store_vec<nt2>(kd, v);
                store_vec<nt>(rh, r);
            }
        } else if ((uint32_t(g.bflags) >> l) & 1u) {
            double v[nt2], r[nt];
#pragma unroll
            for (int e = 0; e < nt2; ++e) v[e] = sd[e * CS];
#pragma unroll
            for (int e = 0; e < nt; ++e) r[e] = rsum(sr, e);
            store_vec<nt2>(kd, v);
            store_vec<nt>(rh, r);
        } else {
#pragma unroll
            for (int e = 0; e < nt2; ++e) atomicAdd(kd + e, sd[e * CS]);
#pragma unroll
            for (int e = 0; e < nt; ++e) atomicAdd(rh + e, rsum(sr, e));
        }
    }
}

}  // namespace hdg
