// element_quad_kernel<K>: the fused blocks + condensation + scatter pass for orders k >= 2, FOUR lanes per element.
//
// Same mathematics and the same outputs as element_schur_kernel<K> (hdg_element.cu; reference:
// examples/poisson2D_HDG.jl:77-184), different mapping.  At k >= 2 a thread per element needs 220-255 registers
// and 0.8-1.9 kB of shared memory, which leaves 8 warps per SM and long FP64 dependency chains (ncu: issue active
// 27-30 %).  Here the t+1 right-hand-side columns of an element - they are independent once S = C + B'A^-1 B is
// factored - are dealt round-robin to 4 adjacent lanes, so a lane carries one column at a time:
//   phase 1  S (n x n) is formed and factored L D L' cooperatively (lane q owns rows q, q+4, ...), L in shared
//            memory; the load vector be is integrated with the quadrature points split over the 4 lanes and
//            summed with two shuffles;
//   phase 2  every lane solves its columns: u = S^-1 r, sigma = A^-1 (r1 + B u) row by row; each row is stored
//            to [K_e | b_e] as soon as it exists and folded into the 3 nt accumulators of its Ate column.
//            Products with structural zeros of the reference matrices Tr, Ts, Fhat are skipped at compile time
//            (hdg_sparsity.h, validated on the host against the tables actually built);
//   phase 3  the face-diagonal blocks and rhs entries staged in shared memory are paired with the neighbour cell
//            of the same 32-cell tile and stored (RED.ADD.F64 only when the neighbour is in another tile).
// A block is 128 threads = one 32-cell tile of the [K_e | b_e] layout; a warp stores 4 columns x 8 cells per
// instruction = 4 full 64-byte segments.
#pragma once

namespace hdg {

template <int K> struct QuadCfg {
    static constexpr int n = Ord<K>::n, nt = Ord<K>::nt, t = Ord<K>::t;
    static constexpr int G = 4;                       // lanes per element
    static constexpr int cells = 32;                  // cells per block = one Ke tile
    static constexpr int threads = G * cells;
    static constexpr int R = (n + G - 1) / G;         // rows of S per lane
    static constexpr int CC = (t + 1 + G - 1) / G;    // columns per lane
    static constexpr int nL = n * (n - 1) / 2;
    // shared-memory record per cell, stored [entry][cell]
    static constexpr int o_L = 0;
    static constexpr int o_dinv = o_L + nL;
    static constexpr int o_be = o_dinv + n;
    static constexpr int o_ca = o_be + n;             // ca[3], cb[3], dJf[3]
    static constexpr int o_diag = o_ca + 9;
    static constexpr int o_rhs = o_diag + 3 * nt * nt;
    static constexpr int entries = o_rhs + 3 * nt;
    static constexpr size_t smem = sizeof(double) * entries * cells;
#ifndef QMINB2
#define QMINB2 4
#endif
#ifndef QMINB3
#define QMINB3 4
#endif
#ifndef QMINB4
#define QMINB4 3
#endif
    static constexpr int min_blocks = K == 2 ? QMINB2 : (K == 3 ? QMINB3 : QMINB4);
};

template <int K>
__global__ void __launch_bounds__(QuadCfg<K>::threads, QuadCfg<K>::min_blocks)
element_quad_kernel(const ElemArgs a, const DevTables<K>* __restrict__ gt) {
    using Q = QuadCfg<K>;
    using Sp = Sparsity<K>;
    constexpr int n = Q::n, nt = Q::nt, t = Q::t, ke = Ord<K>::ke, CS = Q::cells, R = Q::R;
    constexpr int64_t nt2 = nt * nt;
    const DevTables<K>& T = ctab<K>();          // uniform-index operands (constant bank)
    extern __shared__ double smem[];
    const int q = threadIdx.x & 3, ci = threadIdx.x >> 2;
    double* const sm = smem + ci;               // this cell's record: sm[entry * CS]
    const unsigned lane = threadIdx.x & 31u;

    int64_t c = a.cell_begin + int64_t(blockIdx.x) * CS + ci;
    bool active = c < a.cell_end;
    if (!active) c = a.cell_end - 1;            // all lanes run (shuffles); only the stores are masked
    const bool dbg = a.dbg_At != nullptr;
    CellGeom g;
    load_geometry(a, c, g);
    if (active && !g.ok) {
        if (q == 0) atomicCAS(&a.flags[FLAG_BAD_GEOM], 0, int32_t(c + 1));
        active = false;
    }
    const double tau = a.tau;
    const bool o0 = g.v[2] > g.v[1], o1 = g.v[0] > g.v[2], o2 = g.v[1] > g.v[0];   // face_orientation, src/mesh.jl:51-54
    const double cf0 = tau * g.dJf[0], cf1 = tau * g.dJf[1], cf2 = tau * g.dJf[2];

    // ---- phase 1a: load vector be, quadrature points dealt to the 4 lanes ---------------------------------
    {
        double bev[n];
#pragma unroll
        for (int i = 0; i < n; ++i) bev[i] = 0.0;
        for (int qq = q; qq < a.nq; qq += 4) {
            double fv;
            if (a.source_id == 0) fv = a.fq[c * a.nq + qq];
            else {
                const double m0 = __ldg(&gt->Mgeo[3 * qq]), m1 = __ldg(&gt->Mgeo[3 * qq + 1]), m2 = __ldg(&gt->Mgeo[3 * qq + 2]);
                const double xq = m0 * g.x[0][0] + m1 * g.x[1][0] + m2 * g.x[2][0];
                const double yq = m0 * g.x[0][1] + m1 * g.x[1][1] + m2 * g.x[2][1];
                fv = source_value(a.source_id, xq, yq);
            }
#pragma unroll
            for (int i = 0; i < n; ++i) bev[i] = fma(__ldg(&gt->WN[qq * n + i]), fv, bev[i]);
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double s = bev[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == (i & 3)) sm[(Q::o_be + i) * CS] = s * g.detJ;
        }
        if (q == 0) {
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                sm[(Q::o_ca + l) * CS] = g.G00 * g.wn[l][0] + g.G01 * g.wn[l][1];        // B'A^-1 E_l = ca_l Qr_l + cb_l Qs_l
                sm[(Q::o_ca + 3 + l) * CS] = g.G10 * g.wn[l][0] + g.G11 * g.wn[l][1];
                sm[(Q::o_ca + 6 + l) * CS] = g.dJf[l];
            }
        }
    }

    // ---- phase 1b: S = C + B'A^-1 B, L D L' with the rows dealt to the 4 lanes ------------------------------
    {
        const double al = g.detJ * (g.G00 * g.G00 + g.G01 * g.G01);
        const double be = g.detJ * (g.G00 * g.G10 + g.G01 * g.G11);
        const double ga = g.detJ * (g.G10 * g.G10 + g.G11 * g.G11);
        const double* __restrict__ gC = gt->Chat + q * n;      // row i = q + 4r  ->  offset (4r) n + j
        const double* __restrict__ gPrr = gt->Prr + q * n;
        const double* __restrict__ gPrs = gt->Prs + q * n;
        const double* __restrict__ gPss = gt->Pss + q * n;
        double dd[n];
        bool spd = true;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < n; ++j) {
            double ljd[n];
#pragma unroll
            for (int k = 0; k < j; ++k) ljd[k] = sm[(Q::o_L + tri(j, k)) * CS] * dd[k];     // L[j][k] d_k
            double colr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                colr[r] = 1.0;
                if (4 * r + 3 < j || 4 * r >= n) continue;         // compile-time: no row of this slot is in [j, n)
                const int i = q + 4 * r;
                if (i >= j && i < n) {
                    const int o = 4 * r * n + j;
                    double s = cf0 * __ldg(gC + o);
                    s = fma(cf1, __ldg(gC + n * n + o), s);
                    s = fma(cf2, __ldg(gC + 2 * n * n + o), s);
                    s = fma(al, __ldg(gPrr + o), s);
                    s = fma(be, __ldg(gPrs + o), s);
                    s = fma(ga, __ldg(gPss + o), s);
                    const double* Li = sm + (Q::o_L + i * (i - 1) / 2) * CS;
#pragma unroll
                    for (int k = 0; k < j; ++k) s = fma(-Li[k * CS], ljd[k], s);
                    colr[r] = s;
                }
            }
            const double dj = __shfl_sync(0xffffffffu, colr[j >> 2], (lane & ~3u) | unsigned(j & 3));
            spd = spd && (dj > 0.0);
            dd[j] = dj;
            const double dinv = 1.0 / dj;
            if (q == (j & 3)) sm[(Q::o_dinv + j) * CS] = dinv;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (4 * r + 3 <= j || 4 * r >= n) continue;
                const int i = q + 4 * r;
                if (i > j && i < n) sm[(Q::o_L + i * (i - 1) / 2 + j) * CS] = colr[r] * dinv;
            }
            __syncwarp();
        }
        if (active && !spd) {
            if (q == 0) atomicCAS(&a.flags[FLAG_SINGULAR], 0, int32_t(c + 1));
            active = false;
        }
    }

    // ---- phase 2: one column of [K_e | b_e] at a time, columns q, q+4, ... ----------------------------------
    const double idet = 1.0 / g.detJ;
    double* __restrict__ const Ke_tile = a.Ke + ((c >> 5) * ke) * 32 + (c & 31);
#pragma unroll 1
    for (int cc = 0; cc < Q::CC; ++cc) {
        const int col = q + 4 * cc;
        if (col > t) break;
        const bool isb = col == t;
        const int l = isb ? 0 : col / nt;
        const int j = isb ? 0 : col - l * nt;
        const bool o_l = l == 0 ? o0 : (l == 1 ? o1 : o2);
        const double scol = (isb || o_l || !(j & 1)) ? 1.0 : -1.0;      // Legendre parity of a reversed face
        const double dJf_l = sm[(Q::o_ca + 6 + l) * CS];

        double u[n];
        if (isb) {
#pragma unroll
            for (int i = 0; i < n; ++i) u[i] = sm[(Q::o_be + i) * CS];
        } else {
            const double cf = tau * dJf_l, ca_l = sm[(Q::o_ca + l) * CS], cb_l = sm[(Q::o_ca + 3 + l) * CS];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                double r = cf * __ldg(&gt->Fhat[i * t + col]);
                r = fma(ca_l, __ldg(&gt->Qr[i * t + col]), r);
                u[i] = fma(cb_l, __ldg(&gt->Qs[i * t + col]), r);
            }
        }
        // S u = r  by L D L'
        const volatile double* const smv = sm;
        // (the compiler barriers keep the n(n-1) shared-memory loads next to their uses: hoisted to the top they
        //  cost 2 registers each and spill)
#pragma unroll
        for (int i = 1; i < n; ++i) {
#pragma unroll
            for (int k = 0; k < i; ++k) u[i] = fma(-smv[(Q::o_L + tri(i, k)) * CS], u[k], u[i]);
            asm volatile("" ::: "memory");
        }
#pragma unroll
        for (int i = 0; i < n; ++i) u[i] *= sm[(Q::o_dinv + i) * CS];
        asm volatile("" ::: "memory");
#pragma unroll
        for (int i = n - 2; i >= 0; --i) {
#pragma unroll
            for (int k = i + 1; k < n; ++k) u[i] = fma(-smv[(Q::o_L + tri(k, i)) * CS], u[k], u[i]);
            asm volatile("" ::: "memory");
        }

        // sigma = A^-1 (r1 + B u) row by row; column of Ate = [E;F]'K_e - He accumulated on the fly
        const double wnx_l = l == 0 ? g.wn[0][0] : (l == 1 ? g.wn[1][0] : g.wn[2][0]);
        const double wny_l = l == 0 ? g.wn[0][1] : (l == 1 ? g.wn[1][1] : g.wn[2][1]);
        const double ex = isb ? 0.0 : wnx_l * idet, ey = isb ? 0.0 : wny_l * idet;
        double val[3][nt];
#pragma unroll
        for (int lp = 0; lp < 3; ++lp)
#pragma unroll
            for (int ip = 0; ip < nt; ++ip) val[lp][ip] = 0.0;
        double* __restrict__ const Kp = Ke_tile + int64_t(col) * 32;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double p = 0.0, s = 0.0;
#pragma unroll
            for (int k = 0; k < n; ++k) {
                if ((Sp::tr(i) >> k) & 1u) p = fma(T.Tr[i * n + k], u[k], p);
                if ((Sp::ts(i) >> k) & 1u) s = fma(T.Ts[i * n + k], u[k], s);
            }
            const double mf = isb ? 0.0 : __ldg(&gt->MF[i * t + col]);
            const double sx = fma(g.G00, p, fma(g.G10, s, -ex * mf));
            const double sy = fma(g.G01, p, fma(g.G11, s, -ey * mf));
            if (active && !dbg) {
                Kp[int64_t(i * (t + 1)) * 32] = scol * sx;
                Kp[int64_t((n + i) * (t + 1)) * 32] = scol * sy;
                Kp[int64_t((2 * n + i) * (t + 1)) * 32] = scol * u[i];
            }
            const double w0 = fma(g.wn[0][0], sx, fma(g.wn[0][1], sy, cf0 * u[i]));
            const double w1 = fma(g.wn[1][0], sx, fma(g.wn[1][1], sy, cf1 * u[i]));
            const double w2 = fma(g.wn[2][0], sx, fma(g.wn[2][1], sy, cf2 * u[i]));
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
                v[ip] = val[lp][ip] * (srow * scol);
            }
            if (isb) {                                   // bte = -[E;F]' b_e
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) {
                    if (dbg) { if (active) a.dbg_bt[lp * nt + ip] = -v[ip]; }
                    else sm[(Q::o_rhs + lp * nt + ip) * CS] = -v[ip];
                }
            } else if (lp == l) {                        // face-diagonal block, minus He (poisson2D_HDG.jl:144-151)
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) {
                    const double h = fma(-dJf_l, __ldg(&gt->Hhat[ip * nt + j]), v[ip]);
                    if (dbg) { if (active) a.dbg_At[col * t + lp * nt + ip] = h; }
                    else sm[(Q::o_diag + (lp * nt + j) * nt + ip) * CS] = h;
                }
            } else if (dbg) {
#pragma unroll
                for (int ip = 0; ip < nt; ++ip) if (active) a.dbg_At[col * t + lp * nt + ip] = v[ip];
            } else if (active) {                         // off-diagonal block: exactly one contributing cell
                const int s = (l - lp + 3) % 3 - 1;
                const int64_t f_lp = g.f[lp] & 0x7fffffffu;
                const int slot = int(g.f[lp] >> 31) * 2 + s;
                store_vec<nt>(a.Ko + (f_lp * 4 + slot) * nt2 + j * nt, v);
            }
        }
    }
    if (dbg) return;

    // ---- phase 3: scatter of the staged face-diagonal blocks and rhs entries (see element_schur_kernel) ------
    __syncthreads();
    if (!active || q == 3) return;
    {
        const int l = q;
        const uint32_t fl = l == 0 ? g.f[0] : (l == 1 ? g.f[1] : g.f[2]);
        const int64_t f = fl & 0x7fffffffu;
        const uint32_t sec = fl >> 31;
        const uint32_t pb = (uint32_t(g.partner) >> (8 * l)) & 0xffu;
        double* const kd = a.Kd + f * nt2;
        double* const rh = a.rhs + f * nt;
        const double* const sd = sm + (Q::o_diag + l * nt2) * CS;
        const double* const sr = sm + (Q::o_rhs + l * nt) * CS;
        if (pb & 0x80u) {
            if (!sec) {
                const int pl = int(pb & 31u), plf = int((pb >> 5) & 3u);
                const double* const pd = smem + pl + (Q::o_diag + plf * nt2) * CS;     // partner cell's staged block
                const double* const pr = smem + pl + (Q::o_rhs + plf * nt) * CS;
                double v[nt2], r[nt];
#pragma unroll
                for (int e = 0; e < nt2; ++e) v[e] = sd[e * CS] + pd[e * CS];
#pragma unroll
                for (int e = 0; e < nt; ++e) r[e] = sr[e * CS] + pr[e * CS];
                store_vec<nt2>(kd, v);
                store_vec<nt>(rh, r);
            }
        } else if ((uint32_t(g.bflags) >> l) & 1u) {
            double v[nt2], r[nt];
#pragma unroll
            for (int e = 0; e < nt2; ++e) v[e] = sd[e * CS];
#pragma unroll
            for (int e = 0; e < nt; ++e) r[e] = sr[e * CS];
            store_vec<nt2>(kd, v);
            store_vec<nt>(rh, r);
        } else {
#pragma unroll
            for (int e = 0; e < nt2; ++e) atomicAdd(kd + e, sd[e * CS]);
#pragma unroll
            for (int e = 0; e < nt; ++e) atomicAdd(rh + e, sr[e * CS]);
        }
    }
}

}  // namespace hdg
