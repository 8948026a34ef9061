// K1+K2+K3: fused local-block formation, static condensation and scatter, one pass per element.
//
// Replaces the per-element body of doassemble (examples/poisson2D_HDG.jl:77-184): reinit!
// (src/ScalarFunctionSpaces.jl:101-132), the quadrature loops (:88-153), the block solve
// Me \ [-E;F], Ate = [E;F]'K_e - He, b_e, bte (:155-174) and the scatter assemble!(...)
// (:176-183, src/assembler.jl:31-60).
//
// Two device paths, both FP64:
//  * element_schur_kernel<K>  (default when the cell rule integrates the mass matrix exactly):
//    one thread per element.  On an affine triangle every block is a geometry scalar times a
//    reference matrix, so Me is eliminated block-wise:
//      S = C + B' A^-1 B = tau*sum_l |wn_l| Chat_l + detJ (alpha Prr + beta Prs + gamma Pss)   (n x n, SPD)
//    factored LDL' per element; the t+1 right-hand sides ([-E;F] columns and [0;be]) are then
//    solved one column at a time, each column immediately contracted into its column of
//    Ate / bte and scattered.  Reference matrices live in __constant__ memory so the operand of
//    most DFMAs comes straight from the constant bank.
//  * element_lu_kernel (any quad_degree; literal quadrature + dense partial-pivot LU, one warp
//    per element in shared memory) - the reference's formulation, used when the mass matrix of
//    the chosen rule is not the identity (e.g. the reference default quad_degree = 3 for k = 2).
//
// Scatter: the trace matrix is stored block-wise (hdg_internal.h): every (row face, column
// face) block of Ate is nt*nt contiguous doubles.  Off-diagonal blocks have exactly one
// contributing cell -> plain stores.  Face-diagonal blocks and rhs entries have at most two
// contributions -> RED.ADD.F64 onto zeroed storage, which is bitwise deterministic because a
// two-term IEEE sum is commutative.
#include <utility>
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "hdg_internal.h"
#include "hdg_sparsity.h"

namespace hdg {

__constant__ DevTables<1> c_tab1;
__constant__ DevTables<2> c_tab2;
__constant__ DevTables<3> c_tab3;
__constant__ DevTables<4> c_tab4;

template <int K> __device__ __forceinline__ const DevTables<K>& ctab();
template <> __device__ __forceinline__ const DevTables<1>& ctab<1>() { return c_tab1; }
template <> __device__ __forceinline__ const DevTables<2>& ctab<2>() { return c_tab2; }
template <> __device__ __forceinline__ const DevTables<3>& ctab<3>() { return c_tab3; }
template <> __device__ __forceinline__ const DevTables<4>& ctab<4>() { return c_tab4; }

struct ElemArgs {
    const int32_t* cellinfo;
    const double* nodes;
    const double* fq;      // ncell x nq (source_id == 0)
    double* Ke;            // [K_e | b_e], KE_TILE32 layout
    double* Kd;
    double* Ko;
    double* rhs;
    int32_t* flags;
    int64_t cell_begin, cell_end;
    double tau;
    int nq;
    int source_id;
    double* dbg_At;        // when non-null: write Ate (t x t column-major) / bte here instead of scattering
    double* dbg_bt;
};

struct CellGeom {
    double detJ, G00, G01, G10, G11;   // G = J^-1
    double wn[3][2];                   // weighted normals, src/shapes.jl:80-87
    double dJf[3];
    double x[3][2];
    int32_t v[3];
    uint32_t f[3];                     // face id | sec<<31
    int32_t partner, bflags;           // in-warp neighbour per local face (8 bits each); boundary-face bits
    bool ok;
};

__device__ __forceinline__ void load_geometry(const ElemArgs& a, int64_t c, CellGeom& g) {
    const int4* ci = reinterpret_cast<const int4*>(a.cellinfo + CI * c);
    int4 p0 = __ldg(ci), p1 = __ldg(ci + 1);
    g.v[0] = p0.x; g.v[1] = p0.y; g.v[2] = p0.z;
    g.f[0] = uint32_t(p0.w); g.f[1] = uint32_t(p1.x); g.f[2] = uint32_t(p1.y);
    g.partner = p1.z; g.bflags = p1.w;
    const double2* nd = reinterpret_cast<const double2*>(a.nodes);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double2 xy = __ldg(nd + g.v[k]);
        g.x[k][0] = xy.x; g.x[k][1] = xy.y;
    }
    // J = [x2-x1 | x3-x1]  (reinit!, src/ScalarFunctionSpaces.jl:105-112)
    double J00 = g.x[1][0] - g.x[0][0], J01 = g.x[2][0] - g.x[0][0];
    double J10 = g.x[1][1] - g.x[0][1], J11 = g.x[2][1] - g.x[0][1];
    g.detJ = J00 * J11 - J01 * J10;
    g.ok = g.detJ > 0.0;
    double id = 1.0 / g.detJ;
    g.G00 = J11 * id; g.G01 = -J01 * id; g.G10 = -J10 * id; g.G11 = J00 * id;
    g.wn[0][0] = -(J10 - J11); g.wn[0][1] = J00 - J01;
    g.wn[1][0] = -J11;         g.wn[1][1] = J01;
    g.wn[2][0] = J10;          g.wn[2][1] = -J00;
#pragma unroll
    for (int l = 0; l < 3; ++l) g.dJf[l] = sqrt(g.wn[l][0] * g.wn[l][0] + g.wn[l][1] * g.wn[l][1]);
}

__device__ __forceinline__ double source_value(int source_id, double x, double y) {
    // f of examples/poisson2D_HDG.jl:55
    const double pi = 3.141592653589793;
#ifdef HDG_USE_SIN
    return 2.0 * (pi * pi) * sin(pi * x) * sin(pi * y);
#else
    // sinpi(x) = sin(pi x) without the rounding of pi*x and with a cheaper argument reduction (measured 1-4 % of the
    // element kernels at k >= 2); differs from sin(pi * x) by an ulp or two
    return 2.0 * (pi * pi) * sinpi(x) * sinpi(y);
#endif
}

// strict lower triangle index
__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i - 1) / 2 + j; }

// widest aligned vector store of N doubles (256-bit STG on sm_100a where the block is 32-byte aligned)
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

// Staging area in shared memory, entry-major ([entry][thread]) so that every access is conflict free.
template <int K> struct Stage {
    static constexpr int nt = Ord<K>::nt;
    __host__ __device__ static constexpr int diag(int l, int j, int ip) { return (l * nt + j) * nt + ip; }          // face-diagonal blocks
    __host__ __device__ static constexpr int rhs(int l, int ip) { return 3 * nt * nt + l * nt + ip; }              // bte
    __host__ __device__ static constexpr int off(int lp, int s, int j, int ip) { return 3 * nt * nt + 3 * nt + ((lp * 2 + s) * nt + j) * nt + ip; }
    static constexpr int n_base = 3 * nt * nt + 3 * nt;
    static constexpr int n_off = 6 * nt * nt;
};

#ifndef MINB1
#define MINB1 5
#endif
#ifndef MINB2
#define MINB2 2
#endif
template <int K> struct SchurCfg {
    static constexpr int threads = K >= 4 ? 64 : 128;
    static constexpr int min_blocks = K == 1 ? MINB1 : (K == 2 ? MINB2 : 1);
#ifndef CU1
#define CU1 7
#endif
#ifndef CU2
#define CU2 10
#endif
#ifndef CU3
#define CU3 1
#endif
    // unroll factor of the column loop: fully unrolled for k=1 (column-dependent selects and indexed constant
    // loads disappear: measured 0.170 -> 0.153 ms per 1M elements)
    static constexpr int col_unroll = K == 1 ? CU1 : (K == 2 ? CU2 : (K == 3 ? CU3 : 1));
    static constexpr bool stage_diag = K <= 3;      // face-diagonal blocks + rhs staged in shared memory (in-warp pairing)
    static constexpr bool l_smem = K >= 3;          // LDL' factor in shared memory (register pressure)
#ifndef K1_STAGE_OFF
#define K1_STAGE_OFF 1
#endif
    static constexpr bool stage_off = K == 1 && K1_STAGE_OFF == 1;  // off-diagonal blocks staged in shared memory -> one 256-bit store per block
    static constexpr bool hold_off = K == 1 && K1_STAGE_OFF == 2;   // ... or held in registers until the block is complete (needs the fully unrolled column loop)
    static constexpr int nL = Ord<K>::n * (Ord<K>::n - 1) / 2;
    static constexpr int smem_doubles = (l_smem ? nL : 0) + (stage_diag ? Stage<K>::n_base : 0) + (stage_off ? Stage<K>::n_off : 0);
};

template <int K>
__global__ void __launch_bounds__(SchurCfg<K>::threads, SchurCfg<K>::min_blocks) element_schur_kernel(const ElemArgs a) {
    constexpr int n = Ord<K>::n, nt = Ord<K>::nt, t = Ord<K>::t, ke = Ord<K>::ke;
    constexpr int nL = SchurCfg<K>::nL;
    constexpr bool L_SMEM = SchurCfg<K>::l_smem, STAGE_OFF = SchurCfg<K>::stage_off, STAGE_DIAG = SchurCfg<K>::stage_diag;
    constexpr bool HOLD_OFF = SchurCfg<K>::hold_off;
    static_assert(!HOLD_OFF || SchurCfg<K>::col_unroll == Ord<K>::t + 1, "hold_off needs the fully unrolled column loop");
    constexpr int B = SchurCfg<K>::threads;
    using St = Stage<K>;
    const DevTables<K>& T = ctab<K>();
    extern __shared__ double smem[];
    double* const smem_L = smem + threadIdx.x;                                   // [nL][B]
    double* const stg = smem + (L_SMEM ? nL * B : 0) + threadIdx.x;              // [entries][B]

    const int64_t c = a.cell_begin + int64_t(blockIdx.x) * B + threadIdx.x;
    bool active = c < a.cell_end;
    const bool dbg = a.dbg_At != nullptr;
    CellGeom g;
    if (active) {
        load_geometry(a, c, g);
        if (!g.ok) {
            atomicCAS(&a.flags[FLAG_BAD_GEOM], 0, int32_t(c + 1));
            active = false;
        }
    }
    const double tau = a.tau;
    // orientation bits, face_orientation src/mesh.jl:51-54: local edge l joins local nodes ((1,2),(2,0),(0,1))
    const bool o0 = g.v[2] > g.v[1], o1 = g.v[0] > g.v[2], o2 = g.v[1] > g.v[0];

    // ---- S = C + B'A^-1 B, LDL' factorisation ------------------------------------------------
    double Lr[L_SMEM ? 1 : (nL > 0 ? nL : 1)];
    double dinv[n];
    auto Lget = [&](int i, int j) -> double {
        if constexpr (L_SMEM) return smem_L[tri(i, j) * B];
        else return Lr[tri(i, j)];
    };
    auto Lset = [&](int i, int j, double v) {
        if constexpr (L_SMEM) smem_L[tri(i, j) * B] = v;
        else Lr[tri(i, j)] = v;
    };
    if (active) {
        double dd[n];
        const double al = g.detJ * (g.G00 * g.G00 + g.G01 * g.G01);
        const double be = g.detJ * (g.G00 * g.G10 + g.G01 * g.G11);
        const double ga = g.detJ * (g.G10 * g.G10 + g.G11 * g.G11);
        const double c0 = tau * g.dJf[0], c1 = tau * g.dJf[1], c2 = tau * g.dJf[2];
        bool spd = true;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            // column j of S below the diagonal, eliminated against columns < j
            double col[n];
#pragma unroll
            for (int i = j; i < n; ++i) {
                double s = c0 * T.Chat[(0 * n + i) * n + j];
                s = fma(c1, T.Chat[(1 * n + i) * n + j], s);
                s = fma(c2, T.Chat[(2 * n + i) * n + j], s);
                s = fma(al, T.Prr[i * n + j], s);
                s = fma(be, T.Prs[i * n + j], s);
                s = fma(ga, T.Pss[i * n + j], s);
                col[i] = s;
            }
#pragma unroll
            for (int k = 0; k < j; ++k) {
                const double ljk_dk = Lget(j, k) * dd[k];   // L[j][k] * d_k
#pragma unroll
                for (int i = j; i < n; ++i) col[i] = fma(-Lget(i, k), ljk_dk, col[i]);
            }
            spd = spd && (col[j] > 0.0);
            dd[j] = col[j];
            dinv[j] = 1.0 / col[j];
#pragma unroll
            for (int i = j + 1; i < n; ++i) Lset(i, j, col[i] * dinv[j]);
        }
        if (!spd) {
            atomicCAS(&a.flags[FLAG_SINGULAR], 0, int32_t(c + 1));
            active = false;
        }
    }

    if (active) {
        // ---- rhs vector be[i] = detJ sum_q w_q f(x_q) N[i,q]  (poisson2D_HDG.jl:106-114) ---------
        double bev[n];
#pragma unroll
        for (int i = 0; i < n; ++i) bev[i] = 0.0;
        for (int q = 0; q < a.nq; ++q) {
            double fv;
            if (a.source_id == 0) fv = a.fq[c * a.nq + q];
            else {
                double xq = T.Mgeo[3 * q] * g.x[0][0] + T.Mgeo[3 * q + 1] * g.x[1][0] + T.Mgeo[3 * q + 2] * g.x[2][0];
                double yq = T.Mgeo[3 * q] * g.x[0][1] + T.Mgeo[3 * q + 1] * g.x[1][1] + T.Mgeo[3 * q + 2] * g.x[2][1];
                fv = source_value(a.source_id, xq, yq);
            }
#pragma unroll
            for (int i = 0; i < n; ++i) bev[i] = fma(T.WN[q * n + i], fv, bev[i]);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) bev[i] *= g.detJ;

        // per-face coefficients of the E-columns:  B'A^-1 E_l = a_l Qr_l + b_l Qs_l
        double ca[3], cb[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            ca[l] = g.G00 * g.wn[l][0] + g.G01 * g.wn[l][1];
            cb[l] = g.G10 * g.wn[l][0] + g.G11 * g.wn[l][1];
        }
        const double idet = 1.0 / g.detJ;
        double* __restrict__ Ke_tile = a.Ke + ((c >> 5) * ke) * 32 + (c & 31);
        constexpr int64_t nt2 = nt * nt;

        double hold[HOLD_OFF ? 3 : 1][2][nt][nt];   // HOLD_OFF: columns of the off-diagonal blocks until a block is complete
        // ---- one column of [K_e | b_e] at a time ------------------------------------------------
        constexpr int col_unroll = SchurCfg<K>::col_unroll;
#pragma unroll col_unroll
        for (int col = 0; col <= t; ++col) {
            const bool isb = col == t;
            const int l = isb ? 0 : col / nt;
            const int j = isb ? 0 : col - l * nt;
            const double dJf_l = l == 0 ? g.dJf[0] : (l == 1 ? g.dJf[1] : g.dJf[2]);
            const double wnx = l == 0 ? g.wn[0][0] : (l == 1 ? g.wn[1][0] : g.wn[2][0]);
            const double wny = l == 0 ? g.wn[0][1] : (l == 1 ? g.wn[1][1] : g.wn[2][1]);
            const double ca_l = l == 0 ? ca[0] : (l == 1 ? ca[1] : ca[2]);
            const double cb_l = l == 0 ? cb[0] : (l == 1 ? cb[1] : cb[2]);
            const bool o_l = l == 0 ? o0 : (l == 1 ? o1 : o2);
            const double scol = (isb || o_l || !(j & 1)) ? 1.0 : -1.0;   // Legendre parity of a reversed face

            double u[n];
            if (isb) {
#pragma unroll
                for (int i = 0; i < n; ++i) u[i] = bev[i];
            } else {
                const double cf = tau * dJf_l;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    double r = cf * T.RT[(col * n + i) * 4];
                    r = fma(ca_l, T.RT[(col * n + i) * 4 + 1], r);
                    u[i] = fma(cb_l, T.RT[(col * n + i) * 4 + 2], r);
                }
            }
            // S u = r  by L D L'
#pragma unroll
            for (int i = 1; i < n; ++i)
#pragma unroll
                for (int k = 0; k < i; ++k) u[i] = fma(-Lget(i, k), u[k], u[i]);
#pragma unroll
            for (int i = 0; i < n; ++i) u[i] *= dinv[i];
#pragma unroll
            for (int i = n - 2; i >= 0; --i)
#pragma unroll
                for (int k = i + 1; k < n; ++k) u[i] = fma(-Lget(k, i), u[k], u[i]);

            // sigma = A^-1 (r1 + B u)
            double sx[n], sy[n];
            const double ex = isb ? 0.0 : wnx * idet, ey = isb ? 0.0 : wny * idet;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                double p = 0.0, q = 0.0;
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    p = fma(T.Tr[i * n + k], u[k], p);
                    q = fma(T.Ts[i * n + k], u[k], q);
                }
                double mf = isb ? 0.0 : T.RT[(col * n + i) * 4 + 3];
                sx[i] = fma(g.G00, p, fma(g.G10, q, -ex * mf));
                sy[i] = fma(g.G01, p, fma(g.G11, q, -ey * mf));
            }
            // store column of [K_e | b_e]   (rows: sigma_x, sigma_y, u); 256-byte segments per warp
            if (!dbg) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    Ke_tile[int64_t((i) * (t + 1) + col) * 32] = scol * sx[i];
                    Ke_tile[int64_t((n + i) * (t + 1) + col) * 32] = scol * sy[i];
                    Ke_tile[int64_t((2 * n + i) * (t + 1) + col) * 32] = scol * u[i];
                }
            }
            // column of Ate = [E;F]' K_e - He  /  bte = -[E;F]' b_e
#pragma unroll
            for (int lp = 0; lp < 3; ++lp) {
                const bool o_lp = lp == 0 ? o0 : (lp == 1 ? o1 : o2);
                const double cf = tau * g.dJf[lp];
                double w[n];
#pragma unroll
                for (int i = 0; i < n; ++i) w[i] = fma(g.wn[lp][0], sx[i], fma(g.wn[lp][1], sy[i], cf * u[i]));
                double val[nt];
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < n; ++k) s = fma(T.Fhat[k * t + lp * nt + ip], w[k], s);
                    const double srow = (o_lp || !(ip & 1)) ? 1.0 : -1.0;
                    val[ip] = s * (srow * scol);
                }
                if (isb) {
#pragma unroll
                    for (int ip = 0; ip < nt; ++ip) {
                        if (dbg) a.dbg_bt[lp * nt + ip] = -val[ip];
                        else if constexpr (STAGE_DIAG) stg[St::rhs(lp, ip) * B] = -val[ip];
                        else atomicAdd(&a.rhs[int64_t(g.f[lp] & 0x7fffffffu) * nt + ip], -val[ip]);
                    }
                } else {
                    if (lp == l) {
#pragma unroll
                        for (int ip = 0; ip < nt; ++ip) val[ip] = fma(-dJf_l, T.Hhat[ip * nt + j], val[ip]);
                    }
                    if (dbg) {
#pragma unroll
                        for (int ip = 0; ip < nt; ++ip) a.dbg_At[col * t + lp * nt + ip] = val[ip];
                    } else if (lp == l) {
                        if constexpr (STAGE_DIAG) {
#pragma unroll
                            for (int ip = 0; ip < nt; ++ip) stg[(St::diag(lp, 0, ip) + j * nt) * B] = val[ip];
                        } else {
                            double* dst = a.Kd + int64_t(g.f[lp] & 0x7fffffffu) * nt2 + j * nt;
#pragma unroll
                            for (int ip = 0; ip < nt; ++ip) atomicAdd(dst + ip, val[ip]);
                        }
                    } else {
                        const int s = (l - lp + 3) % 3 - 1;
                        if constexpr (HOLD_OFF) {
#pragma unroll
                            for (int ip = 0; ip < nt; ++ip) hold[lp][s][j][ip] = val[ip];
                            if (j == nt - 1) {   // block (row face lp, column face l) complete: one 256-bit store
                                double blk[nt * nt];
#pragma unroll
                                for (int jj = 0; jj < nt; ++jj)
#pragma unroll
                                    for (int ip = 0; ip < nt; ++ip) blk[jj * nt + ip] = hold[lp][s][jj][ip];
                                const int64_t f_lp = g.f[lp] & 0x7fffffffu;
                                store_vec<nt * nt>(a.Ko + (f_lp * 4 + int(g.f[lp] >> 31) * 2 + s) * nt2, blk);
                            }
                        } else if constexpr (STAGE_OFF) {
#pragma unroll
                            for (int ip = 0; ip < nt; ++ip) stg[(St::off(lp, 0, 0, ip) + (s * nt + j) * nt) * B] = val[ip];
                        } else {
                            const int64_t f_lp = g.f[lp] & 0x7fffffffu;
                            const int slot = int(g.f[lp] >> 31) * 2 + s;
                            store_vec<nt>(a.Ko + (f_lp * 4 + slot) * nt2 + j * nt, val);
                        }
                    }
                }
            }
        }
    }
    if (dbg || !STAGE_DIAG) return;
    // ---- scatter of the staged blocks ------------------------------------------------------------
    // Every trace-matrix entry has one or two contributing cells.  If the neighbour across a face
    // sits in the same warp its contribution is read from shared memory and the first cell stores
    // the two-term sum; boundary faces are stored directly; only faces whose neighbour lives in
    // another warp use RED.ADD.F64 onto the zeroed array (two-term sums are commutative, so the
    // result is bitwise independent of the order in every case).
    __syncwarp();
    if (!active) return;
    constexpr int64_t nt2 = nt * nt;
    const uint32_t pinfo = uint32_t(g.partner), binfo = uint32_t(g.bflags);
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const int64_t f = g.f[l] & 0x7fffffffu;
        const uint32_t sec = g.f[l] >> 31;
        const uint32_t pb = (pinfo >> (8 * l)) & 0xffu;
        double* const kd = a.Kd + f * nt2;
        double* const rh = a.rhs + f * nt;
        if (pb & 0x80u) {
            if (!sec) {
                const int pl = int(pb & 31u), plf = int((pb >> 5) & 3u);
                const double* ps = stg + (pl - int(threadIdx.x & 31));           // partner's column of the staging area
                double v[nt2], r[nt];
#pragma unroll
                for (int e = 0; e < nt2; ++e) v[e] = stg[(St::diag(l, 0, 0) + e) * B] + ps[(plf * nt2 + e) * B];
#pragma unroll
                for (int e = 0; e < nt; ++e) r[e] = stg[St::rhs(l, e) * B] + ps[(3 * nt2 + plf * nt + e) * B];
                store_vec<nt2>(kd, v);
                store_vec<nt>(rh, r);
            }
        } else if ((binfo >> l) & 1u) {
            double v[nt2], r[nt];
#pragma unroll
            for (int e = 0; e < nt2; ++e) v[e] = stg[(St::diag(l, 0, 0) + e) * B];
#pragma unroll
            for (int e = 0; e < nt; ++e) r[e] = stg[St::rhs(l, e) * B];
            store_vec<nt2>(kd, v);
            store_vec<nt>(rh, r);
        } else {
#pragma unroll
            for (int e = 0; e < nt2; ++e) atomicAdd(kd + e, stg[(St::diag(l, 0, 0) + e) * B]);
#pragma unroll
            for (int e = 0; e < nt; ++e) atomicAdd(rh + e, stg[St::rhs(l, e) * B]);
        }
        if constexpr (STAGE_OFF) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                double v[nt2];
#pragma unroll
                for (int e = 0; e < nt2; ++e) v[e] = stg[(St::off(l, s, 0, 0) + e) * B];
                store_vec<nt2>(a.Ko + (f * 4 + sec * 2 + s) * nt2, v);
            }
        }
    }
}

}  // namespace hdg
#include "hdg_element_quad.cuh"
namespace hdg {

// ---------------------------------------------------------------------------------------------
// General path: literal quadrature + dense LU with partial pivoting, one warp per element.
// Follows examples/poisson2D_HDG.jl:88-174 term by term.
// ---------------------------------------------------------------------------------------------
struct LuArgs {
    ElemArgs e;
    RawTablesDev raw;
    int n, nt;
};

__global__ void __launch_bounds__(128) element_lu_kernel(const LuArgs A) {
    const ElemArgs& a = A.e;
    const RawTablesDev& R = A.raw;
    const int n = A.n, nt = A.nt, nv = 2 * n, m = 3 * n, t = 3 * nt, nc = t + 1;   // rhs columns
    const int ld = (m + nc) | 1;   // row length of [Me | rhs], odd so that lane-strided rows hit distinct banks
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    // per warp: aug[m][ld], G[m][t] ( = [E;F] ), be[n], perm scratch
    const int per_warp = m * ld + m * t + n + 8;
    double* aug = sm + size_t(warp) * per_warp;
    double* Gm = aug + m * ld;
    double* bev = Gm + m * t;

    const int64_t c = a.cell_begin + int64_t(blockIdx.x) * wpb + warp;
    if (c >= a.cell_end) return;
    CellGeom g;
    load_geometry(a, c, g);
    if (!g.ok) {
        if (lane == 0) atomicCAS(&a.flags[FLAG_BAD_GEOM], 0, int32_t(c + 1));
        return;
    }
    const double tau = a.tau;
    const bool ori[3] = {g.v[2] > g.v[1], g.v[0] > g.v[2], g.v[1] > g.v[0]};
    for (int k = lane; k < m * ld + m * t + n; k += 32) aug[k] = 0.0;
    __syncwarp();

    // cell integrals: A (block diagonal mass) and B  (:88-104).  One (i,j) pair per lane-iteration.
    for (int idx = lane; idx < n * n; idx += 32) {
        int i = idx / n, j = idx - i * n;
        double mass = 0.0, bx = 0.0, by = 0.0;
        for (int q = 0; q < R.nq; ++q) {
            double dO = g.detJ * R.qw[q];
            double Ni = R.N[i + n * q], Nj = R.N[j + n * q];
            double dr = R.dN[(i + n * q) * 2], ds = R.dN[(i + n * q) * 2 + 1];
            double gx = dr * g.G00 + ds * g.G10, gy = dr * g.G01 + ds * g.G11;   // dNdxi . Jinv
            mass += (Nj * Ni) * dO;
            bx += (Nj * gx) * dO;
            by += (Nj * gy) * dO;
        }
        aug[i * ld + j] = mass;                     // A, x component
        aug[(n + i) * ld + n + j] = mass;           // A, y component
        aug[i * ld + nv + j] = -bx;                 // -B
        aug[(n + i) * ld + nv + j] = -by;
        aug[(nv + j) * ld + i] = bx;                // B'
        aug[(nv + j) * ld + n + i] = by;
    }
    // rhs be (:106-114)
    for (int i = lane; i < n; i += 32) {
        double s = 0.0;
        for (int q = 0; q < R.nq; ++q) {
            double fv;
            if (a.source_id == 0) fv = a.fq[c * R.nq + q];
            else {
                double xq = R.Mgeo[3 * q] * g.x[0][0] + R.Mgeo[3 * q + 1] * g.x[1][0] + R.Mgeo[3 * q + 2] * g.x[2][0];
                double yq = R.Mgeo[3 * q] * g.x[0][1] + R.Mgeo[3 * q + 1] * g.x[1][1] + R.Mgeo[3 * q + 2] * g.x[2][1];
                fv = source_value(a.source_id, xq, yq);
            }
            s += fv * R.N[i + n * q] * (g.detJ * R.qw[q]);
        }
        bev[i] = s;
    }
    // face integrals C (:120-126)
    for (int idx = lane; idx < n * n; idx += 32) {
        int i = idx / n, j = idx - i * n;
        double s = 0.0;
        for (int l = 0; l < 3; ++l)
            for (int p = 0; p < R.nfq; ++p) {
                double dS = g.dJf[l] * R.fw[p];
                s += tau * (R.E[j + n * (p + R.nfq * l)] * R.E[i + n * (p + R.nfq * l)]) * dS;
            }
        aug[(nv + i) * ld + nv + j] = s;
    }
    // F, E (:127-143): G = [E;F]
    for (int idx = lane; idx < n * t; idx += 32) {
        int i = idx / t, cj = idx - i * t, l = cj / nt, j = cj - l * nt;
        double f = 0.0;
        for (int p = 0; p < R.nfq; ++p) {
            int po = ori[l] ? p : R.nfq - 1 - p;
            double dS = g.dJf[l] * R.fw[p];
            f += (R.T[j + nt * p] * R.E[i + n * (po + R.nfq * l)]) * dS;
        }
        double nx = g.wn[l][0] / g.dJf[l], ny = g.wn[l][1] / g.dJf[l];
        Gm[i * t + cj] = f * nx;
        Gm[(n + i) * t + cj] = f * ny;
        Gm[(nv + i) * t + cj] = tau * f;
    }
    __syncwarp();
    // right-hand sides [-E;F | 0;be]
    for (int idx = lane; idx < m * nc; idx += 32) {
        int i = idx / nc, cj = idx - i * nc;
        double v;
        if (cj < t) v = i < nv ? -Gm[i * t + cj] : Gm[i * t + cj];
        else v = i < nv ? 0.0 : bev[i - nv];
        aug[i * ld + m + cj] = v;
    }
    __syncwarp();

    // LU with partial pivoting (LAPACK getrf semantics) on [Me | rhs], rows distributed over lanes
    bool singular = false;
    for (int k = 0; k < m; ++k) {
        // pivot search
        double best = -1.0;
        int bi = k;
        for (int i = k + lane; i < m; i += 32) {
            double v = fabs(aug[i * ld + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!(best > 0.0)) { singular = true; break; }
        if (bi != k)
            for (int j2 = lane; j2 < ld; j2 += 32) {
                double tmp = aug[k * ld + j2];
                aug[k * ld + j2] = aug[bi * ld + j2];
                aug[bi * ld + j2] = tmp;
            }
        __syncwarp();
        double ip = 1.0 / aug[k * ld + k];
        for (int i = k + 1 + lane; i < m; i += 32) {
            double lik = aug[i * ld + k] * ip;
            for (int j2 = k + 1; j2 < ld; ++j2) aug[i * ld + j2] = fma(-lik, aug[k * ld + j2], aug[i * ld + j2]);
        }
        __syncwarp();
    }
    if (singular) {
        if (lane == 0) atomicCAS(&a.flags[FLAG_SINGULAR], 0, int32_t(c + 1));
        return;
    }
    // back substitution, one rhs column per lane
    for (int cj = lane; cj < nc; cj += 32) {
        for (int i = m - 1; i >= 0; --i) {
            double s = aug[i * ld + m + cj];
            for (int k = i + 1; k < m; ++k) s = fma(-aug[i * ld + k], aug[k * ld + m + cj], s);
            aug[i * ld + m + cj] = s / aug[i * ld + i];
        }
    }
    __syncwarp();
    // store [K_e | b_e]
    if (a.dbg_At == nullptr) {
        double* Ke_tile = a.Ke + ((c >> 5) * int64_t(m * nc)) * 32 + (c & 31);
        for (int idx = lane; idx < m * nc; idx += 32) {
            int i = idx / nc, cj = idx - i * nc;
            Ke_tile[int64_t(idx) * 32] = aug[i * ld + m + cj];
        }
    }
    // Ate = G' K_e - He ; bte = -G' b_e ; scatter
    const int nt2 = nt * nt;
    for (int idx = lane; idx < t * nc; idx += 32) {
        int col = idx / t, row = idx - col * t;        // (row, col) of Ate, col == t -> bte
        double s = 0.0;
        for (int i = 0; i < m; ++i) s = fma(Gm[i * t + row], aug[i * ld + m + col], s);
        int lp = row / nt, ip = row - lp * nt;
        int64_t f_lp = g.f[lp] & 0x7fffffffu;
        if (col == t) {
            if (a.dbg_At) a.dbg_bt[row] = -s;
            else atomicAdd(&a.rhs[f_lp * nt + ip], -s);
            continue;
        }
        int l = col / nt, j = col - l * nt;
        if (l == lp) {   // He (:144-151)
            double h = 0.0;
            for (int p = 0; p < R.nfq; ++p) h += (R.T[j + nt * p] * R.T[ip + nt * p]) * (g.dJf[l] * R.fw[p]);
            s -= h;
        }
        if (a.dbg_At) a.dbg_At[col * t + row] = s;
        else if (l == lp) atomicAdd(&a.Kd[f_lp * nt2 + j * nt + ip], s);
        else {
            int slot = int(g.f[lp] >> 31) * 2 + ((l - lp + 3) % 3 - 1);
            a.Ko[(f_lp * 4 + slot) * nt2 + j * nt + ip] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// General path, order 1: one THREAD per element (a warp per element leaves most lanes idle at m = 9; measured 4.8x
// faster than the warp kernel.  At k = 2 the per-thread [Me | rhs] no longer fits registers and the warp kernel wins).
// Literal quadrature (examples/poisson2D_HDG.jl:88-153) into a per-thread [Me | rhs] (registers at k=1, L1-backed
// local memory at k=2), LU WITHOUT pivoting: Me = [A -B; B' C] has a positive definite symmetric part
// (x'Me x = s'As + u'Cu), so elimination in the natural order is stable, every lane walks the same index sequence
// (coalesced local-memory traffic, no divergence), and the result agrees with LAPACK's pivoted LU to rounding.
// A zero / non-finite pivot raises the singular flag (the reference's SingularException).
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) element_lu_thread_kernel(const LuArgs A) {
    constexpr int n = Ord<K>::n, nt = Ord<K>::nt, nv = 2 * n, m = Ord<K>::m, t = Ord<K>::t, nc = t + 1, ke = Ord<K>::ke;
    const ElemArgs& a = A.e;
    const RawTablesDev& R = A.raw;
    const int64_t c = a.cell_begin + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= a.cell_end) return;
    CellGeom g;
    load_geometry(a, c, g);
    if (!g.ok) {
        atomicCAS(&a.flags[FLAG_BAD_GEOM], 0, int32_t(c + 1));
        return;
    }
    const double tau = a.tau;
    const bool ori[3] = {g.v[2] > g.v[1], g.v[0] > g.v[2], g.v[1] > g.v[0]};
    double Me[m][m], Rh[m][nc], Fl[n][t];
#pragma unroll
    for (int i = 0; i < m; ++i) {
#pragma unroll
        for (int j = 0; j < m; ++j) Me[i][j] = 0.0;
#pragma unroll
        for (int j = 0; j < nc; ++j) Rh[i][j] = 0.0;
    }
    // cell integrals A, B (:88-104) and rhs be (:106-114)
    for (int q = 0; q < R.nq; ++q) {
        const double dO = g.detJ * R.qw[q];
        double fv;
        if (a.source_id == 0) fv = a.fq[c * R.nq + q];
        else {
            double xq = R.Mgeo[3 * q] * g.x[0][0] + R.Mgeo[3 * q + 1] * g.x[1][0] + R.Mgeo[3 * q + 2] * g.x[2][0];
            double yq = R.Mgeo[3 * q] * g.x[0][1] + R.Mgeo[3 * q + 1] * g.x[1][1] + R.Mgeo[3 * q + 2] * g.x[2][1];
            fv = source_value(a.source_id, xq, yq);
        }
        double Nq[n], gx[n], gy[n];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            Nq[i] = R.N[i + n * q];
            const double dr = R.dN[(i + n * q) * 2], ds = R.dN[(i + n * q) * 2 + 1];
            gx[i] = dr * g.G00 + ds * g.G10;   // dNdxi . Jinv
            gy[i] = dr * g.G01 + ds * g.G11;
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const double mass = (Nq[j] * Nq[i]) * dO;
                Me[i][j] += mass;
                Me[n + i][n + j] += mass;
                const double bx = (Nq[j] * gx[i]) * dO, by = (Nq[j] * gy[i]) * dO;
                Me[i][nv + j] -= bx;
                Me[n + i][nv + j] -= by;
                Me[nv + j][i] += bx;
                Me[nv + j][n + i] += by;
            }
            Rh[nv + i][t] += fv * Nq[i] * dO;
        }
    }
    // face integrals C, F (:116-143); E = F (x) normal
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < t; ++j) Fl[i][j] = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l)
        for (int p = 0; p < R.nfq; ++p) {
            const double dS = g.dJf[l] * R.fw[p];
            const int po = ori[l] ? p : R.nfq - 1 - p;
            double Ep[n], Eo[n], Tp[nt];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                Ep[i] = R.E[i + n * (p + R.nfq * l)];
                Eo[i] = R.E[i + n * (po + R.nfq * l)];
            }
#pragma unroll
            for (int j = 0; j < nt; ++j) Tp[j] = R.T[j + nt * p];
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j < n; ++j) Me[nv + i][nv + j] += tau * (Ep[j] * Ep[i]) * dS;
#pragma unroll
                for (int j = 0; j < nt; ++j) Fl[i][l * nt + j] += (Tp[j] * Eo[i]) * dS;
            }
        }
    double nrm[3][2];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        nrm[l][0] = g.wn[l][0] / g.dJf[l];
        nrm[l][1] = g.wn[l][1] / g.dJf[l];
    }
    // right-hand sides [-E; F | 0; be]
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int col = 0; col < t; ++col) {
            const int l = col / nt;
            Rh[i][col] = -Fl[i][col] * nrm[l][0];
            Rh[n + i][col] = -Fl[i][col] * nrm[l][1];
            Rh[nv + i][col] = tau * Fl[i][col];
        }
    // LU without pivoting + forward elimination of the right-hand sides
    // k = 1: everything unrolled (arrays live in registers); k = 2: outer loops rolled, [Me | rhs] in local memory
    constexpr int UO = K == 1 ? 64 : 1;
    bool singular = false;
#pragma unroll UO
    for (int k = 0; k < m; ++k) {
        const double piv = Me[k][k];
        singular = singular || !(fabs(piv) > 0.0) || !isfinite(piv);
        const double ip = 1.0 / piv;
#pragma unroll UO
        for (int i = k + 1; i < m; ++i) {
            const double lik = Me[i][k] * ip;
#pragma unroll UO
            for (int j = k + 1; j < m; ++j) Me[i][j] = fma(-lik, Me[k][j], Me[i][j]);
#pragma unroll
            for (int j = 0; j < nc; ++j) Rh[i][j] = fma(-lik, Rh[k][j], Rh[i][j]);
        }
    }
    if (singular) {
        atomicCAS(&a.flags[FLAG_SINGULAR], 0, int32_t(c + 1));
        return;
    }
#pragma unroll UO
    for (int i = m - 1; i >= 0; --i) {
        const double id = 1.0 / Me[i][i];
#pragma unroll
        for (int j = 0; j < nc; ++j) {
            double s = Rh[i][j];
#pragma unroll UO
            for (int k = i + 1; k < m; ++k) s = fma(-Me[i][k], Rh[k][j], s);
            Rh[i][j] = s * id;
        }
    }
    const bool dbg = a.dbg_At != nullptr;
    if (!dbg) {
        double* __restrict__ Ke_tile = a.Ke + ((c >> 5) * ke) * 32 + (c & 31);
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int j = 0; j < nc; ++j) Ke_tile[int64_t(i * nc + j) * 32] = Rh[i][j];
    }
    // Ate = [E;F]' K_e - He ; bte = -[E;F]' b_e ; scatter (plain stores for single-contribution blocks, RED otherwise)
    constexpr int64_t nt2 = nt * nt;
#pragma unroll
    for (int lp = 0; lp < 3; ++lp) {
        const int64_t f_lp = g.f[lp] & 0x7fffffffu;
        const int sec = int(g.f[lp] >> 31);
#pragma unroll UO
        for (int col = 0; col < nc; ++col) {
            double val[nt];
#pragma unroll
            for (int ip = 0; ip < nt; ++ip) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const double f = Fl[i][lp * nt + ip];
                    s = fma(f * nrm[lp][0], Rh[i][col], s);
                    s = fma(f * nrm[lp][1], Rh[n + i][col], s);
                    s = fma(tau * f, Rh[nv + i][col], s);
                }
                val[ip] = s;
            }
            if (col == t) {
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) {
                    if (dbg) a.dbg_bt[lp * nt + ip] = -val[ip];
                    else atomicAdd(&a.rhs[f_lp * nt + ip], -val[ip]);
                }
                continue;
            }
            const int l = col / nt, j = col - l * nt;
            if (l == lp) {   // He (:144-151)
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) {
                    double h = 0.0;
                    for (int p = 0; p < R.nfq; ++p) h += (R.T[j + nt * p] * R.T[ip + nt * p]) * (g.dJf[l] * R.fw[p]);
                    val[ip] -= h;
                }
            }
            if (dbg) {
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) a.dbg_At[col * t + lp * nt + ip] = val[ip];
            } else if (l == lp) {
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) atomicAdd(&a.Kd[f_lp * nt2 + j * nt + ip], val[ip]);
            } else {
                const int slot = sec * 2 + ((l - lp + 3) % 3 - 1);
                store_vec<nt>(a.Ko + (f_lp * 4 + slot) * nt2 + j * nt, val);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int K> static void fill_dev_tables(const RefTables& R, DevTables<K>& D) {
    auto cp = [](const std::vector<double>& src, double* dst, size_t cap) {
        for (size_t i = 0; i < cap; ++i) dst[i] = i < src.size() ? src[i] : 0.0;
    };
    constexpr int n = Ord<K>::n, t = Ord<K>::t, nt = Ord<K>::nt;
    cp(R.Tr, D.Tr, n * n); cp(R.Ts, D.Ts, n * n);
    cp(R.Prr, D.Prr, n * n); cp(R.Prs, D.Prs, n * n); cp(R.Pss, D.Pss, n * n);
    cp(R.Chat, D.Chat, 3 * n * n);
    cp(R.Fhat, D.Fhat, n * t);
    for (int col = 0; col < t; ++col)
        for (int i = 0; i < n; ++i) {
            double* q = D.RT + (col * n + i) * 4;
            q[0] = R.Fhat[i * t + col]; q[1] = R.Qr[i * t + col]; q[2] = R.Qs[i * t + col]; q[3] = R.MF[i * t + col];
        }
    cp(R.Hhat, D.Hhat, nt * nt);
    cp(R.WN, D.WN, MAX_NQ * n);
    cp(R.Mgeo, D.Mgeo, MAX_NQ * 3);
    cp(R.qw, D.qw, MAX_NQ);
    D.nq = R.nq;
    D.pad = 0;
}

// element_quad_kernel skips the products hdg_sparsity.h calls structurally zero: check the header against the tables
// this context actually built (any rule that integrates the mass matrix exactly gives the same reference matrices).
template <int K> static bool sparsity_matches(const DevTables<K>& D) {
    constexpr int n = Ord<K>::n, t = Ord<K>::t;
    auto amax = [](const double* p, int cnt) { double m = 0.0; for (int i = 0; i < cnt; ++i) m = std::max(m, std::fabs(p[i])); return m; };
    const double mr = amax(D.Tr, n * n), ms = amax(D.Ts, n * n), mf = amax(D.Fhat, n * t);
    const double tol = 1e-10;
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < n; ++k) {
            if (!((Sparsity<K>::tr(i) >> k) & 1u) && std::fabs(D.Tr[i * n + k]) > tol * mr) return false;
            if (!((Sparsity<K>::ts(i) >> k) & 1u) && std::fabs(D.Ts[i * n + k]) > tol * ms) return false;
        }
        for (int j = 0; j < t; ++j)
            if (!((Sparsity<K>::fh(i) >> j) & 1u) && std::fabs(D.Fhat[i * t + j]) > tol * mf) return false;
    }
    return true;
}

// The constant tables are per (device, order), shared by every context of that order on the device.  Whoever uploaded last
// owns them: a context re-uploads when another one (e.g. a different quad_degree) has used them since; hdg_destroy releases.
constexpr int MAX_DEVICES = 16;
static const hdg_context* g_table_owner[MAX_DEVICES][MAX_ORDER + 1] = {};
static const hdg_context*& table_owner(const hdg_context* c) { return g_table_owner[c->device & (MAX_DEVICES - 1)][c->tab.order]; }
void release_tables(const hdg_context* c) {
    if (c->tab.order >= 1 && c->tab.order <= MAX_ORDER && table_owner(c) == c) table_owner(c) = nullptr;
}

template <int K, typename Sym> static hdg_status upload_dev_tables(hdg_context* c, const Sym& symbol) {
    static DevTables<K> D;
    fill_dev_tables<K>(c->tab, D);
    // the tables are shared by every context of this order on this device: kernels of the previous owner may still be
    // running on its (non-blocking) stream, and from here on the tables are this context's
    HDG_CUDA(c, cudaDeviceSynchronize());
    HDG_CUDA(c, cudaMemcpyToSymbol(symbol, &D, sizeof(D)));
    table_owner(c) = c;
    c->quad_ok = sparsity_matches<K>(D);
    return HDG_OK;
}

hdg_status upload_tables(hdg_context* c) {
    const RefTables& R = c->tab;
    if (R.nq > MAX_NQ || R.nfq > MAX_NFQ) return set_err(c, HDG_ERR_UNSUPPORTED_RULE, "quadrature rule too large for the device tables");
    if (!c->use_lu) {
        hdg_status st = HDG_OK;
        switch (R.order) {
            case 1: st = upload_dev_tables<1>(c, c_tab1); break;
            case 2: st = upload_dev_tables<2>(c, c_tab2); break;
            case 3: st = upload_dev_tables<3>(c, c_tab3); break;
            case 4: st = upload_dev_tables<4>(c, c_tab4); break;
        }
        if (st) return st;
    }
    // raw tables in global memory (LU path, error norm)
    std::vector<double> buf;
    auto push = [&](const std::vector<double>& v) { size_t o = buf.size(); buf.insert(buf.end(), v.begin(), v.end()); return o; };
    size_t oN = push(R.N), odN = push(R.dN), oE = push(R.E), oT = push(R.T), oqw = push(R.qw), ofw = push(R.fw), oM = push(R.Mgeo);
    if (c->d_rawtab) cudaFree(c->d_rawtab);
    HDG_CUDA(c, cudaMalloc(&c->d_rawtab, sizeof(double) * buf.size()));
    HDG_CUDA(c, cudaMemcpy(c->d_rawtab, buf.data(), sizeof(double) * buf.size(), cudaMemcpyHostToDevice));
    c->raw.N = c->d_rawtab + oN; c->raw.dN = c->d_rawtab + odN; c->raw.E = c->d_rawtab + oE; c->raw.T = c->d_rawtab + oT;
    c->raw.qw = c->d_rawtab + oqw; c->raw.fw = c->d_rawtab + ofw; c->raw.Mgeo = c->d_rawtab + oM;
    c->raw.n = R.n; c->raw.nt = R.nt; c->raw.nq = R.nq; c->raw.nfq = R.nfq;
    return HDG_OK;
}


template <int K> static hdg_status launch_schur(hdg_context* c, const ElemArgs& a) {
    constexpr int B = SchurCfg<K>::threads;
    constexpr size_t smem = sizeof(double) * SchurCfg<K>::smem_doubles * B;
    auto kern = element_schur_kernel<K>;
    if (smem > 48 * 1024) HDG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t ncell = a.cell_end - a.cell_begin;
    kern<<<(unsigned)ceil_div(ncell, B), B, smem, c->stream>>>(a);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

// Orders >= 2: four lanes per element (hdg_element_quad.cuh).  HDG_ELEM_V1=1 selects the thread-per-element kernel
// (A/B measurements); it is also the fallback if hdg_sparsity.h does not match the tables of this context.
template <int K> static hdg_status launch_quad(hdg_context* c, const ElemArgs& a) {
    using Q = QuadCfg<K>;
    auto kern = element_quad_kernel<K>;
    if (Q::smem > 48 * 1024) HDG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Q::smem)));
    int64_t ncell = a.cell_end - a.cell_begin;
    kern<<<(unsigned)ceil_div(ncell, Q::cells), Q::threads, Q::smem, c->stream>>>(a);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

static bool use_quad_kernel(const hdg_context* c) {
    static const bool v1 = getenv("HDG_ELEM_V1") != nullptr;
    return c->tab.order >= 2 && c->quad_ok && !v1;
}

static hdg_status launch_elements(hdg_context* c, ElemArgs a) {
    if (c->use_lu && c->tab.order == 1 && getenv("HDG_LU_WARP") == nullptr) {   // k = 2 measured 2.4x slower than the warp kernel (local-memory bound)
        LuArgs A{a, c->raw, c->tab.n, c->tab.nt};
        int64_t ncell = a.cell_end - a.cell_begin;
        element_lu_thread_kernel<1><<<(unsigned)ceil_div(ncell, 128), 128, 0, c->stream>>>(A);
        c->launches += 1;
        HDG_CUDA(c, cudaGetLastError());
        return HDG_OK;
    }
    if (c->use_lu) {
        LuArgs A{a, c->raw, c->tab.n, c->tab.nt};
        const int m = c->tab.m, t = c->tab.t, ld = (m + t + 1) | 1;
        const int wpb = 4;
        size_t smem = sizeof(double) * size_t(wpb) * (m * ld + m * t + c->tab.n + 8);
        if (smem > 48 * 1024) HDG_CUDA(c, cudaFuncSetAttribute(element_lu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        int64_t ncell = a.cell_end - a.cell_begin;
        element_lu_kernel<<<(unsigned)ceil_div(ncell, wpb), wpb * 32, smem, c->stream>>>(A);
        c->launches += 1;
        HDG_CUDA(c, cudaGetLastError());
        return HDG_OK;
    }
    if (table_owner(c) != c) {
        hdg_status st = upload_tables(c);      // synchronises the device and takes ownership
        if (st) return st;
    }
    if (use_quad_kernel(c)) {
        switch (c->tab.order) {
            case 2: return launch_quad<2>(c, a);
            case 3: return launch_quad<3>(c, a);
            case 4: return launch_quad<4>(c, a);
        }
    }
    switch (c->tab.order) {
        case 1: return launch_schur<1>(c, a);
        case 2: return launch_schur<2>(c, a);
        case 3: return launch_schur<3>(c, a);
        case 4: return launch_schur<4>(c, a);
    }
    return set_err(c, HDG_ERR_INVALID, "unsupported order");
}

hdg_status launch_element_kernels(hdg_context* c) {
    const int nt = c->tab.nt;
    // RED.ADD targets start from zero: Kd and rhs (contiguous).  Zeroing only the faces that are accumulated with RED.ADD (cells in
    // two tiles, 1/3 of the faces of rectangle_mesh) through a face list was measured SLOWER than the plain memset of everything:
    // 0.1866 vs 0.1751 ms per step at k = 1 (scattered 16/32-byte writes, and the element kernel itself loses 5 us)
    HDG_CUDA(c, cudaMemsetAsync(c->d_Kd, 0, sizeof(double) * c->nface * (nt * nt + nt), c->stream));
    ElemArgs a{};
    a.cellinfo = c->d_cellinfo; a.nodes = c->d_nodes; a.fq = c->d_fq;
    a.Ke = c->d_Ke; a.Kd = c->d_Kd; a.Ko = c->d_Ko; a.rhs = c->d_rhs; a.flags = c->d_flags;
    a.cell_begin = 0; a.cell_end = c->ncell; a.tau = c->prm.tau; a.nq = c->tab.nq; a.source_id = c->prm.source_id;
    a.dbg_At = nullptr; a.dbg_bt = nullptr;
    timer_start(c, c->t_elem);
    hdg_status st = launch_elements(c, a);
    timer_stop(c, c->t_elem);
    return st;
}

hdg_status condensed_of_cell(hdg_context* c, int64_t cell, double* At, double* bt) {
    const int t = c->tab.t;
    double* d = nullptr;
    HDG_CUDA(c, cudaMalloc(&d, sizeof(double) * (t * t + t)));
    ElemArgs a{};
    a.cellinfo = c->d_cellinfo; a.nodes = c->d_nodes; a.fq = c->d_fq;
    a.Ke = c->d_Ke; a.Kd = c->d_Kd; a.Ko = c->d_Ko; a.rhs = c->d_rhs; a.flags = c->d_flags;
    a.cell_begin = cell; a.cell_end = cell + 1; a.tau = c->prm.tau; a.nq = c->tab.nq; a.source_id = c->prm.source_id;
    a.dbg_At = d; a.dbg_bt = d + t * t;
    hdg_status st = launch_elements(c, a);
    if (st) { cudaFree(d); return st; }
    std::vector<double> h(t * t + t);
    HDG_CUDA(c, cudaMemcpyAsync(h.data(), d, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d);
    for (int i = 0; i < t * t; ++i) At[i] = h[i];
    for (int i = 0; i < t; ++i) bt[i] = h[t * t + i];
    return HDG_OK;
}

}  // namespace hdg
