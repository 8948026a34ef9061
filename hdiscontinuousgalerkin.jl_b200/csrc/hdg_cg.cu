// CG side of the exported API on the device: examples/poisson2D_CG.jl on the mesh of the context.
//
//   DofHandler([u_h], mesh)        _distribute_dofs, src/dofhandler.jl:80-152      -> cg_setup: first-encounter dof numbering
//   create_sparsity_pattern(dh)    src/dofhandler.jl:181-220                        -> cg_setup: CSC pattern, bit-identical
//   doassemble(Wh, K, dh)          examples/poisson2D_CG.jl:72-126, assemble! src/assembler.jl:62-137   -> cg_assemble
//   Dirichlet(u_h, dh, "boundary", [0.0]) + apply!     src/boundary.jl:48-96, 121-158                   -> cg_apply_dirichlet
//   u = K \ b                      examples/poisson2D_CG.jl:134                      -> cg_solve (Jacobi-PCG, K is SPD here)
//   reconstruct! + errornorm       src/dofhandler.jl:218-228, src/DiscreteFunctions.jl:97-120            -> cg_errornorm
//
// ContinuousLagrange{2,RefTetrahedron,order}, order 1 or 2, one scalar field, quad_degree = order + 1 (the default of
// ScalarFunctionSpace).  The reference numbers dofs with a sequential dictionary walk over the cells (vertices of a cell, then
// its faces); here an entity (vertex / face) is numbered by the exclusive scan over "first encounter" flags in the same
// (cell, slot) order - the technique of hdg_number_faces.  The pattern is gathered per column from a dof -> cells adjacency
// (sorted in registers), the element matrices are added with RED.ADD at positions found by binary search (the job of
// AssemblerSparsityPattern's sorted-dof walk).  One GPU.
#include <algorithm>
#include <cmath>
#include <vector>

#include "hdg_internal.h"
#include "hdg_reduce.cuh"

namespace hdg {

constexpr int CG_MAXDPC = 6;     // dofs per cell: 3 (P1), 6 (P2)
constexpr int CG_MAXNQ = 8;      // Strang rule of degree 3 has 6 points
constexpr int CG_ADJ = 16;       // cells per dof (vertex valence); more -> HDG_ERR_INVALID
constexpr int CG_MAXROW = 64;    // stored entries per column

struct CgTables {
    double N[CG_MAXDPC * CG_MAXNQ];        // N[i*nq + q]
    double dN[CG_MAXDPC * CG_MAXNQ * 2];   // dN[(i*nq + q)*2 + a]
    double M[3 * CG_MAXNQ];                // geometry map (1-r-s, r, s) at the points: M[g*nq + q]
    double qw[CG_MAXNQ];
    int ndpc, nq;
};

struct CgData {
    int order = 0, ndpc = 0;
    int64_t ncell = 0, nface = 0, nnode = 0, ndofs = 0, nnz = 0;
    CgTables tab{};
    int32_t* celldofs = nullptr;   // ncell x ndpc, 0-based
    int64_t* colptr = nullptr;     // ndofs + 1
    int32_t* rowval = nullptr;     // nnz
    double *nzval = nullptr, *rhs = nullptr, *u = nullptr;
    double *r = nullptr, *p = nullptr, *Ap = nullptr, *dinv = nullptr;
    uint8_t* isdir = nullptr;
    double meandiag = 0.0;
    bool assembled = false, applied = false, solved = false;
};

static void cg_free(CgData* g) {
    if (!g) return;
    void* ptrs[] = {g->celldofs, g->colptr, g->rowval, g->nzval, g->rhs, g->u, g->r, g->p, g->Ap, g->dinv, g->isdir};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete g;
}
void cg_release(hdg_context* c) { cg_free(static_cast<CgData*>(c->cg)); c->cg = nullptr; }

// ---- Lagrange tables (host): the nodal basis on the points of get_nodal_points (src/shapes.jl:46-57) - vertices, then the
// midpoints of the reference edges (1,0)-(0,1), (0,1)-(0,0), (0,0)-(1,0); closed forms of what src/basis.jl:264-293 obtains by
// inverting the Dubiner Vandermonde matrix
static void lagrange_eval(int order, double r, double s, double* v, double* dr, double* ds) {
    const double l0 = 1.0 - r - s, l1 = r, l2 = s;
    if (order == 1) {
        v[0] = l0; v[1] = l1; v[2] = l2;
        dr[0] = -1; dr[1] = 1; dr[2] = 0;
        ds[0] = -1; ds[1] = 0; ds[2] = 1;
        return;
    }
    v[0] = l0 * (2 * l0 - 1); v[1] = l1 * (2 * l1 - 1); v[2] = l2 * (2 * l2 - 1);
    v[3] = 4 * l1 * l2; v[4] = 4 * l2 * l0; v[5] = 4 * l0 * l1;
    dr[0] = -(4 * l0 - 1); dr[1] = 4 * l1 - 1; dr[2] = 0;
    ds[0] = -(4 * l0 - 1); ds[1] = 0; ds[2] = 4 * l2 - 1;
    dr[3] = 4 * l2; ds[3] = 4 * l1;
    dr[4] = -4 * l2; ds[4] = 4 * (l0 - l2);
    dr[5] = 4 * (l0 - l1); ds[5] = -4 * l1;
}

// ---- dof numbering ----------------------------------------------------------------------------------------------------
__global__ void cg_vertex_first(const int32_t* __restrict__ cellinfo, int64_t ncell, int32_t* __restrict__ vfirst) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicMin(&vfirst[cellinfo[CI * c + k]], int32_t(c));
}
// flag[(c, slot)] = 1 if cell c is the first to meet the entity of the slot (slots 0-2: its vertices, 3-5: its faces)
__global__ void cg_first_flags(const int32_t* __restrict__ cellinfo, const int32_t* __restrict__ facecell, const int32_t* __restrict__ vfirst,
                               int64_t ncell, int ndpc, int32_t* __restrict__ flag) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ncell * ndpc) return;
    const int64_t c = i / ndpc;
    const int s = int(i % ndpc);
    if (s < 3) flag[i] = vfirst[cellinfo[CI * c + s]] == int32_t(c);
    else flag[i] = facecell[2 * (uint32_t(cellinfo[CI * c + s]) & 0x7fffffffu)] == int32_t(c);
}
__global__ void cg_entity_dofs(const int32_t* __restrict__ cellinfo, const int32_t* __restrict__ flag, const int64_t* __restrict__ offs,
                               int64_t ncell, int ndpc, int32_t* __restrict__ vdof, int32_t* __restrict__ fdof) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ncell * ndpc || !flag[i]) return;
    const int64_t c = i / ndpc;
    const int s = int(i % ndpc);
    if (s < 3) vdof[cellinfo[CI * c + s]] = int32_t(offs[i]);
    else fdof[uint32_t(cellinfo[CI * c + s]) & 0x7fffffffu] = int32_t(offs[i]);
}
__global__ void cg_cell_dofs(const int32_t* __restrict__ cellinfo, const int32_t* __restrict__ vdof, const int32_t* __restrict__ fdof,
                             int64_t ncell, int ndpc, int32_t* __restrict__ celldofs) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ncell * ndpc) return;
    const int64_t c = i / ndpc;
    const int s = int(i % ndpc);
    celldofs[i] = s < 3 ? vdof[cellinfo[CI * c + s]] : fdof[uint32_t(cellinfo[CI * c + s]) & 0x7fffffffu];
}

// ---- sparsity pattern --------------------------------------------------------------------------------------------------
__global__ void cg_adj_fill(const int32_t* __restrict__ celldofs, int64_t ncell, int ndpc, int32_t* __restrict__ cnt, int32_t* __restrict__ adj,
                            int32_t* __restrict__ flags) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ncell * ndpc) return;
    const int32_t d = celldofs[i];
    const int k = atomicAdd(&cnt[d], 1);
    if (k < CG_ADJ) adj[int64_t(d) * CG_ADJ + k] = int32_t(i / ndpc); else atomicExch(&flags[FLAG_MG], 1);
}
// column j of sparse(I, J, V): the dofs of all cells that hold dof j (and j itself), ascending, duplicates merged
__device__ int cg_column(const int32_t* __restrict__ celldofs, int ndpc, const int32_t* __restrict__ cnt, const int32_t* __restrict__ adj,
                         int64_t j, int32_t* rows) {
    int m = 0;
    rows[m++] = int32_t(j);
    const int nc = min(cnt[j], CG_ADJ);
    for (int k = 0; k < nc; ++k) {
        const int64_t c = adj[j * CG_ADJ + k];
        for (int s = 0; s < ndpc; ++s) {
            const int32_t d = celldofs[c * ndpc + s];
            bool seen = false;
            for (int q = 0; q < m; ++q) seen = seen || rows[q] == d;
            if (!seen && m < CG_MAXROW) rows[m++] = d;
            else if (!seen) return -1;
        }
    }
    for (int a = 1; a < m; ++a) {      // insertion sort
        const int32_t key = rows[a];
        int b = a - 1;
        while (b >= 0 && rows[b] > key) { rows[b + 1] = rows[b]; --b; }
        rows[b + 1] = key;
    }
    return m;
}
__global__ void cg_col_count(const int32_t* __restrict__ celldofs, int ndpc, const int32_t* __restrict__ cnt, const int32_t* __restrict__ adj,
                             int64_t ndofs, int32_t* __restrict__ count, int32_t* __restrict__ flags) {
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= ndofs) return;
    int32_t rows[CG_MAXROW];
    const int m = cg_column(celldofs, ndpc, cnt, adj, j, rows);
    if (m < 0) { atomicExch(&flags[FLAG_MG], 2); count[j] = 0; return; }
    count[j] = m;
}
__global__ void cg_col_fill(const int32_t* __restrict__ celldofs, int ndpc, const int32_t* __restrict__ cnt, const int32_t* __restrict__ adj,
                            int64_t ndofs, const int64_t* __restrict__ colptr, int32_t* __restrict__ rowval) {
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= ndofs) return;
    int32_t rows[CG_MAXROW];
    const int m = cg_column(celldofs, ndpc, cnt, adj, j, rows);
    for (int q = 0; q < m; ++q) rowval[colptr[j] + q] = rows[q];
}

__device__ __forceinline__ int64_t cg_find(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int32_t col, int32_t row) {
    int64_t lo = colptr[col], hi = colptr[col + 1] - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < row) lo = mid + 1; else hi = mid; }
    return lo;
}

// ---- doassemble: one thread per cell --------------------------------------------------------------------------------------
__global__ void cg_assemble_kernel(const CgTables T, const int32_t* __restrict__ cellinfo, const double* __restrict__ nodes,
                                   const int32_t* __restrict__ celldofs, const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                   int64_t ncell, double* __restrict__ nzval, double* __restrict__ rhs, int32_t* __restrict__ flags) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int n = T.ndpc, nq = T.nq;
    double x[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) { const int64_t v = cellinfo[CI * c + g]; x[g][0] = nodes[2 * v]; x[g][1] = nodes[2 * v + 1]; }
    // reinit! (src/ScalarFunctionSpaces.jl:101-132): J = [x2 - x1; x3 - x1], dNdx = dNdxi . Jinv
    const double J00 = x[1][0] - x[0][0], J01 = x[1][1] - x[0][1], J10 = x[2][0] - x[0][0], J11 = x[2][1] - x[0][1];
    const double detJ = J00 * J11 - J01 * J10;
    if (!(detJ > 0.0)) { atomicCAS(&flags[FLAG_BAD_GEOM], 0, int32_t(c + 1)); return; }
    const double i00 = J11 / detJ, i01 = -J01 / detJ, i10 = -J10 / detJ, i11 = J00 / detJ;      // inverse of [J00 J01; J10 J11] (rows = reference directions)
    double Ke[CG_MAXDPC][CG_MAXDPC], fe[CG_MAXDPC];
    for (int i = 0; i < n; ++i) { fe[i] = 0.0; for (int j = 0; j < n; ++j) Ke[i][j] = 0.0; }
    for (int q = 0; q < nq; ++q) {
        const double dO = detJ * T.qw[q];
        const double xq = T.M[0 * nq + q] * x[0][0] + T.M[1 * nq + q] * x[1][0] + T.M[2 * nq + q] * x[2][0];
        const double yq = T.M[0 * nq + q] * x[0][1] + T.M[1 * nq + q] * x[1][1] + T.M[2 * nq + q] * x[2][1];
        const double fh = 2.0 * (M_PI * M_PI) * sin(M_PI * xq) * sin(M_PI * yq);       // examples/poisson2D_CG.jl:69
        double gx[CG_MAXDPC], gy[CG_MAXDPC];
        for (int i = 0; i < n; ++i) {
            const double dr = T.dN[(i * nq + q) * 2], ds = T.dN[(i * nq + q) * 2 + 1];
            gx[i] = dr * i00 + ds * i01;      // dNdx = dNdxi . Jinv with Jinv = inverse of [dx/dr dx/ds; dy/dr dy/ds] = transpose of (i..)
            gy[i] = dr * i10 + ds * i11;
        }
        for (int i = 0; i < n; ++i) {
            fe[i] += fh * T.N[i * nq + q] * dO;
            for (int j = 0; j < n; ++j) Ke[i][j] += (gx[i] * gx[j] + gy[i] * gy[j]) * dO;
        }
    }
    // assemble!(assembler, cell_dofs, fe, Ke): K[dof_i, dof_j] += Ke[i, j] at the stored position, f[dof_i] += fe[i]
    for (int j = 0; j < n; ++j) {
        const int32_t dj = celldofs[c * n + j];
        atomicAdd(&rhs[dj], fe[j]);
        for (int i = 0; i < n; ++i) atomicAdd(&nzval[cg_find(colptr, rowval, dj, celldofs[c * n + i])], Ke[i][j]);
    }
}

// ---- Dirichlet(u_h, dh, "boundary", [0.0]) and apply! ---------------------------------------------------------------------
__global__ void cg_mark_dirichlet(const int32_t* __restrict__ bfaces, int64_t nb, const int32_t* __restrict__ facecell,
                                  const int32_t* __restrict__ cellinfo, const int32_t* __restrict__ celldofs, int ndpc, uint8_t* __restrict__ isdir) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int32_t f = bfaces[i];
    const int64_t c = facecell[2 * f];
    int l = 0;
    for (int k = 0; k < 3; ++k)
        if (int32_t(uint32_t(cellinfo[CI * c + 3 + k]) & 0x7fffffffu) == f) l = k;
    const int e0[3] = {1, 2, 0}, e1[3] = {2, 0, 1};      // reference_edge_nodes ((2,3),(3,1),(1,2))
    isdir[celldofs[c * ndpc + e0[l]]] = 1;
    isdir[celldofs[c * ndpc + e1[l]]] = 1;
    if (ndpc == 6) isdir[celldofs[c * ndpc + 3 + l]] = 1;
}
__global__ void cg_diag_abs(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const double* __restrict__ nzval, int64_t ndofs,
                            double* __restrict__ part) {
    double s = 0.0;
    for (int64_t j = int64_t(blockIdx.x) * RB + threadIdx.x; j < ndofs; j += int64_t(gridDim.x) * RB)
        s += fabs(nzval[cg_find(colptr, rowval, int32_t(j), int32_t(j))]);
    const double tot = block_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}
// zero the prescribed rows and columns in place (pattern unchanged), K[d,d] = m, f[d] = 0   (src/boundary.jl:139-157, g = 0)
__global__ void cg_apply_kernel(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, double* __restrict__ nzval,
                                double* __restrict__ rhs, const uint8_t* __restrict__ isdir, double m, int64_t ndofs) {
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= ndofs) return;
    const bool dj = isdir[j];
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
        const int32_t i = rowval[p];
        if (dj || isdir[i]) nzval[p] = (dj && i == j) ? m : 0.0;
    }
    if (dj) rhs[j] = 0.0;
}

// ---- Jacobi-PCG on the CSC matrix (symmetric: the columns are the rows) --------------------------------------------------
__global__ void __launch_bounds__(RB) cg_pcg_init(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const double* __restrict__ nzval,
                                                  const double* __restrict__ b, int64_t n, double* __restrict__ x, double* __restrict__ r,
                                                  double* __restrict__ p, double* __restrict__ dinv, double* __restrict__ part) {
    double rz = 0.0, bb = 0.0;
    for (int64_t j = int64_t(blockIdx.x) * RB + threadIdx.x; j < n; j += int64_t(gridDim.x) * RB) {
        const double d = 1.0 / nzval[cg_find(colptr, rowval, int32_t(j), int32_t(j))];
        dinv[j] = d; x[j] = 0.0; r[j] = b[j]; p[j] = d * b[j];
        rz = fma(b[j], d * b[j], rz); bb = fma(b[j], b[j], bb);
    }
    const double t1 = block_sum(rz), t2 = block_sum(bb);
    if (threadIdx.x == 0) { part[blockIdx.x] = t1; part[2048 + blockIdx.x] = t2; }
}
__global__ void __launch_bounds__(RB) cg_pcg_spmv(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const double* __restrict__ nzval,
                                                  const double* __restrict__ p, int64_t n, double* __restrict__ Ap, double* __restrict__ part) {
    double pap = 0.0;
    for (int64_t j = int64_t(blockIdx.x) * RB + threadIdx.x; j < n; j += int64_t(gridDim.x) * RB) {
        double s = 0.0;
        for (int64_t q = colptr[j]; q < colptr[j + 1]; ++q) s = fma(nzval[q], p[rowval[q]], s);
        Ap[j] = s;
        pap = fma(p[j], s, pap);
    }
    const double tot = block_sum(pap);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}
// scal: [0] r.z, [1] p.Ap, [2] r.z new, [3] r.r, [4] b.b
__global__ void __launch_bounds__(RB) cg_pcg_update(const double* __restrict__ p, const double* __restrict__ Ap, const double* __restrict__ dinv, int64_t n,
                                                    const double* __restrict__ scal, double* __restrict__ x, double* __restrict__ r, double* __restrict__ part) {
    const double alpha = scal[0] / scal[1];
    double rz = 0.0, rr = 0.0;
    for (int64_t j = int64_t(blockIdx.x) * RB + threadIdx.x; j < n; j += int64_t(gridDim.x) * RB) {
        x[j] = fma(alpha, p[j], x[j]);
        const double rn = fma(-alpha, Ap[j], r[j]);
        r[j] = rn;
        rz = fma(rn * dinv[j], rn, rz); rr = fma(rn, rn, rr);
    }
    const double t1 = block_sum(rz), t2 = block_sum(rr);
    if (threadIdx.x == 0) { part[blockIdx.x] = t1; part[2048 + blockIdx.x] = t2; }
}
__global__ void __launch_bounds__(RB) cg_pcg_dir(const double* __restrict__ r, const double* __restrict__ dinv, int64_t n, double* __restrict__ scal,
                                                 double* __restrict__ p) {
    const double beta = scal[2] / scal[0];
    for (int64_t j = int64_t(blockIdx.x) * RB + threadIdx.x; j < n; j += int64_t(gridDim.x) * RB) p[j] = fma(beta, p[j], dinv[j] * r[j]);
}
__global__ void cg_sum2(const double* __restrict__ part, int np, double* __restrict__ out0, double* __restrict__ out1) {
    const double a = reduce_partials(part, np), b = out1 ? reduce_partials(part + 2048, np) : 0.0;
    if (threadIdx.x == 0) { *out0 = a; if (out1) *out1 = b; }
}
__global__ void cg_shift(double* scal) { scal[0] = scal[2]; }      // r.z of the next iteration

// ---- errornorm(u_h, u_ex) after reconstruct!: squared L2 error with the cell rule -----------------------------------------
__global__ void __launch_bounds__(RB) cg_errornorm_kernel(const CgTables T, const int32_t* __restrict__ cellinfo, const double* __restrict__ nodes,
                                                          const int32_t* __restrict__ celldofs, const double* __restrict__ u, int64_t ncell,
                                                          double* __restrict__ part) {
    double tot = 0.0;
    const int n = T.ndpc, nq = T.nq;
    for (int64_t c = int64_t(blockIdx.x) * RB + threadIdx.x; c < ncell; c += int64_t(gridDim.x) * RB) {
        double x[3][2];
#pragma unroll
        for (int g = 0; g < 3; ++g) { const int64_t v = cellinfo[CI * c + g]; x[g][0] = nodes[2 * v]; x[g][1] = nodes[2 * v + 1]; }
        const double detJ = (x[1][0] - x[0][0]) * (x[2][1] - x[0][1]) - (x[2][0] - x[0][0]) * (x[1][1] - x[0][1]);
        for (int q = 0; q < nq; ++q) {
            double uq = 0.0;
            for (int i = 0; i < n; ++i) uq += u[celldofs[c * n + i]] * T.N[i * nq + q];
            const double xq = T.M[0 * nq + q] * x[0][0] + T.M[1 * nq + q] * x[1][0] + T.M[2 * nq + q] * x[2][0];
            const double yq = T.M[0 * nq + q] * x[0][1] + T.M[1 * nq + q] * x[1][1] + T.M[2 * nq + q] * x[2][1];
            const double d = uq - sin(M_PI * xq) * sin(M_PI * yq);
            tot += d * d * (detJ * T.qw[q]);
        }
    }
    const double s = block_sum(tot);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// export helpers (1-based int64 like the Julia structs)
__global__ void cg_export_plus1_32(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = int64_t(in[i]) + 1;
}
__global__ void cg_export_plus1_64(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + 1;
}

static inline unsigned nb(int64_t n, int b = 256) { return (unsigned)std::max<int64_t>(1, ceil_div(n, b)); }

hdg_status cg_setup(hdg_context* c, int order, int64_t* ndofs_out) {
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "the CG side runs on one GPU");
    if (order < 1 || order > 2) return set_err(c, HDG_ERR_INVALID, "ContinuousLagrange order must be 1 or 2");
    cg_release(c);
    CgData* g = new CgData();
    c->cg = g;
    g->order = order; g->ndpc = order == 1 ? 3 : 6;
    g->ncell = c->ncell; g->nface = c->nface; g->nnode = c->nnode;
    const int ndpc = g->ndpc;
    // tables of ScalarFunctionSpace(mesh, ContinuousLagrange{2,RefTetrahedron,order}()) with the default rule quad_degree = order + 1
    {
        std::vector<double> pts, w;
        try { cell_rule(order + 1, pts, w); } catch (const std::string& e) { return set_err(c, HDG_ERR_UNSUPPORTED_RULE, e); }
        const int nq = int(w.size());
        if (nq > CG_MAXNQ) return set_err(c, HDG_ERR_UNSUPPORTED_RULE, "cell rule too large for the CG tables");
        CgTables& T = g->tab;
        T.ndpc = ndpc; T.nq = nq;
        for (int q = 0; q < nq; ++q) {
            double v[6], dr[6], ds[6];
            lagrange_eval(order, pts[2 * q], pts[2 * q + 1], v, dr, ds);
            for (int i = 0; i < ndpc; ++i) { T.N[i * nq + q] = v[i]; T.dN[(i * nq + q) * 2] = dr[i]; T.dN[(i * nq + q) * 2 + 1] = ds[i]; }
            T.M[0 * nq + q] = 1.0 - pts[2 * q] - pts[2 * q + 1]; T.M[1 * nq + q] = pts[2 * q]; T.M[2 * nq + q] = pts[2 * q + 1];
            T.qw[q] = w[q];
        }
    }
    const int64_t nslots = g->ncell * ndpc;
    int32_t *vfirst = nullptr, *flag = nullptr, *vdof = nullptr, *fdof = nullptr, *cnt = nullptr, *adj = nullptr, *count = nullptr;
    int64_t* offs = nullptr;
    auto fail = [&](hdg_status s) { cudaStreamSynchronize(c->stream); for (void* q : {(void*)vfirst, (void*)flag, (void*)vdof, (void*)fdof, (void*)cnt, (void*)adj, (void*)count, (void*)offs}) if (q) cudaFree(q); return s; };
    HDG_CUDA(c, cudaMalloc(&vfirst, sizeof(int32_t) * g->nnode));
    HDG_CUDA(c, cudaMalloc(&flag, sizeof(int32_t) * nslots));
    HDG_CUDA(c, cudaMalloc(&offs, sizeof(int64_t) * (nslots + 1)));
    HDG_CUDA(c, cudaMalloc(&vdof, sizeof(int32_t) * g->nnode));
    HDG_CUDA(c, cudaMalloc(&fdof, sizeof(int32_t) * g->nface));
    HDG_CUDA(c, cudaMalloc(&g->celldofs, sizeof(int32_t) * nslots));
    HDG_CUDA(c, cudaMemsetAsync(vfirst, 0x7f, sizeof(int32_t) * g->nnode, c->stream));
    cg_vertex_first<<<nb(g->ncell), 256, 0, c->stream>>>(c->d_cellinfo, g->ncell, vfirst);
    cg_first_flags<<<nb(nslots), 256, 0, c->stream>>>(c->d_cellinfo, c->d_facecell, vfirst, g->ncell, ndpc, flag);
    hdg_status st = exclusive_scan_i32(c, flag, nslots, offs, offs + nslots);
    if (st) return fail(st);
    HDG_CUDA(c, cudaMemcpy(&g->ndofs, offs + nslots, sizeof(int64_t), cudaMemcpyDeviceToHost));
    cg_entity_dofs<<<nb(nslots), 256, 0, c->stream>>>(c->d_cellinfo, flag, offs, g->ncell, ndpc, vdof, fdof);
    cg_cell_dofs<<<nb(nslots), 256, 0, c->stream>>>(c->d_cellinfo, vdof, fdof, g->ncell, ndpc, g->celldofs);
    c->launches += 4;
    // pattern
    HDG_CUDA(c, cudaMalloc(&cnt, sizeof(int32_t) * g->ndofs));
    HDG_CUDA(c, cudaMalloc(&adj, sizeof(int32_t) * g->ndofs * CG_ADJ));
    HDG_CUDA(c, cudaMalloc(&count, sizeof(int32_t) * g->ndofs));
    HDG_CUDA(c, cudaMalloc(&g->colptr, sizeof(int64_t) * (g->ndofs + 1)));
    HDG_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * g->ndofs, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_MG, 0, sizeof(int32_t), c->stream));
    cg_adj_fill<<<nb(nslots), 256, 0, c->stream>>>(g->celldofs, g->ncell, ndpc, cnt, adj, c->d_flags);
    cg_col_count<<<nb(g->ndofs, 128), 128, 0, c->stream>>>(g->celldofs, ndpc, cnt, adj, g->ndofs, count, c->d_flags);
    c->launches += 2;
    st = exclusive_scan_i32(c, count, g->ndofs, g->colptr, g->colptr + g->ndofs);
    if (st) return fail(st);
    HDG_CUDA(c, cudaMemcpy(&g->nnz, g->colptr + g->ndofs, sizeof(int64_t), cudaMemcpyDeviceToHost));
    HDG_CUDA(c, cudaMemcpy(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost));
    if (c->h_flags[FLAG_MG]) return fail(set_err(c, HDG_ERR_INVALID, c->h_flags[FLAG_MG] == 1 ? "a vertex belongs to more than 16 cells" : "a column of the CG matrix has more than 64 entries"));
    HDG_CUDA(c, cudaMalloc(&g->rowval, sizeof(int32_t) * g->nnz));
    cg_col_fill<<<nb(g->ndofs, 128), 128, 0, c->stream>>>(g->celldofs, ndpc, cnt, adj, g->ndofs, g->colptr, g->rowval);
    c->launches += 1;
    HDG_CUDA(c, cudaMalloc(&g->nzval, sizeof(double) * g->nnz));
    for (double** v : {&g->rhs, &g->u, &g->r, &g->p, &g->Ap, &g->dinv}) HDG_CUDA(c, cudaMalloc(v, sizeof(double) * g->ndofs));
    HDG_CUDA(c, cudaMalloc(&g->isdir, g->ndofs));
    fail(HDG_OK);
    if (ndofs_out) *ndofs_out = g->ndofs;
    return HDG_OK;
}

static CgData* cg_get(hdg_context* c) { return static_cast<CgData*>(c->cg); }

hdg_status cg_sizes(hdg_context* c, int64_t out[4]) {
    CgData* g = cg_get(c);
    if (!g) return set_err(c, HDG_ERR_INVALID, "hdg_cg_setup first");
    out[0] = g->ndofs; out[1] = g->nnz; out[2] = g->ndpc; out[3] = g->ncell;
    return HDG_OK;
}

hdg_status cg_download(hdg_context* c, int64_t* cell_dofs, int64_t* colptr, int64_t* rowval, double* nzval, double* rhs, double* u) {
    CgData* g = cg_get(c);
    if (!g) return set_err(c, HDG_ERR_INVALID, "hdg_cg_setup first");
    const int64_t nmax = std::max({g->ncell * g->ndpc, g->ndofs + 1, g->nnz});
    int64_t* tmp = nullptr;
    if (cell_dofs || colptr || rowval) HDG_CUDA(c, cudaMalloc(&tmp, sizeof(int64_t) * nmax));
    auto out64 = [&](int64_t* host, int64_t n) -> hdg_status {
        HDG_CUDA(c, cudaMemcpyAsync(host, tmp, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        return HDG_OK;
    };
    hdg_status st = HDG_OK;
    if (cell_dofs) { cg_export_plus1_32<<<nb(g->ncell * g->ndpc), 256, 0, c->stream>>>(g->celldofs, g->ncell * g->ndpc, tmp); st = out64(cell_dofs, g->ncell * g->ndpc); }
    if (!st && colptr) { cg_export_plus1_64<<<nb(g->ndofs + 1), 256, 0, c->stream>>>(g->colptr, g->ndofs + 1, tmp); st = out64(colptr, g->ndofs + 1); }
    if (!st && rowval) { cg_export_plus1_32<<<nb(g->nnz), 256, 0, c->stream>>>(g->rowval, g->nnz, tmp); st = out64(rowval, g->nnz); }
    if (tmp) cudaFree(tmp);
    if (st) return st;
    if ((nzval || rhs) && !g->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_cg_assemble first");
    if (nzval) HDG_CUDA(c, cudaMemcpyAsync(nzval, g->nzval, sizeof(double) * g->nnz, cudaMemcpyDeviceToHost, c->stream));
    if (rhs) HDG_CUDA(c, cudaMemcpyAsync(rhs, g->rhs, sizeof(double) * g->ndofs, cudaMemcpyDeviceToHost, c->stream));
    if (u) {
        if (!g->solved) return set_err(c, HDG_ERR_INVALID, "hdg_cg_solve first");
        HDG_CUDA(c, cudaMemcpyAsync(u, g->u, sizeof(double) * g->ndofs, cudaMemcpyDeviceToHost, c->stream));
    }
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status cg_assemble(hdg_context* c) {
    CgData* g = cg_get(c);
    if (!g) return set_err(c, HDG_ERR_INVALID, "hdg_cg_setup first");
    // start_assemble(K, b): fill!(K.nzval, 0), fill!(f, 0)   (src/assembler.jl:78-82)
    HDG_CUDA(c, cudaMemsetAsync(g->nzval, 0, sizeof(double) * g->nnz, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(g->rhs, 0, sizeof(double) * g->ndofs, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
    cg_assemble_kernel<<<nb(g->ncell, 128), 128, 0, c->stream>>>(g->tab, c->d_cellinfo, c->d_nodes, g->celldofs, g->colptr, g->rowval, g->ncell, g->nzval, g->rhs, c->d_flags);
    c->launches += 1;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->h_flags[FLAG_BAD_GEOM]) return set_err(c, HDG_ERR_BAD_GEOMETRY, "det(J) is not positive: det(J) <= 0 in cell " + std::to_string(c->h_flags[FLAG_BAD_GEOM]));
    g->assembled = true; g->applied = g->solved = false;
    return HDG_OK;
}

hdg_status cg_apply_dirichlet(hdg_context* c) {
    CgData* g = cg_get(c);
    if (!g || !g->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_cg_assemble first");
    if (g->applied) return set_err(c, HDG_ERR_INVALID, "the system was already modified: assemble again first");
    HDG_CUDA(c, cudaMemsetAsync(g->isdir, 0, g->ndofs, c->stream));
    if (c->nbface) cg_mark_dirichlet<<<nb(c->nbface), 256, 0, c->stream>>>(c->d_bfaces, c->nbface, c->d_facecell, c->d_cellinfo, g->celldofs, g->ndpc, g->isdir);
    const int np = int(std::min<int64_t>(ceil_div(g->ndofs, RB), 1024));
    cg_diag_abs<<<np, RB, 0, c->stream>>>(g->colptr, g->rowval, g->nzval, g->ndofs, c->d_partials);
    cg_sum2<<<1, RB, 0, c->stream>>>(c->d_partials, np, c->d_scal, nullptr);
    c->launches += 3;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    g->meandiag = c->h_scal[0] / double(g->ndofs);      // meandiag, src/boundary.jl:169-175
    cg_apply_kernel<<<nb(g->ndofs), 256, 0, c->stream>>>(g->colptr, g->rowval, g->nzval, g->rhs, g->isdir, g->meandiag, g->ndofs);
    c->launches += 1;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    g->applied = true;
    return HDG_OK;
}

hdg_status cg_solve(hdg_context* c, double rtol, int maxit, hdg_solve_info* info) {
    CgData* g = cg_get(c);
    if (!g || !g->applied) return set_err(c, HDG_ERR_INVALID, "hdg_cg_apply_dirichlet first");
    const int64_t n = g->ndofs;
    const int G = int(std::min<int64_t>(ceil_div(n, RB), 1184));
    double* scal = c->d_scal;      // [0] r.z [1] p.Ap [2] r.z new [3] r.r [4] b.b
    timer_start(c, c->t_solve);
    cg_pcg_init<<<G, RB, 0, c->stream>>>(g->colptr, g->rowval, g->nzval, g->rhs, n, g->u, g->r, g->p, g->dinv, c->d_partials);
    cg_sum2<<<1, RB, 0, c->stream>>>(c->d_partials, G, scal + 0, scal + 4);
    c->launches += 2;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    const double bb = c->h_scal[4];
    int it = 0;
    bool done = !(bb > 0.0);
    double rr = bb;
    while (!done && it < maxit) {
        const int chunk = std::min(16, maxit - it);
        for (int k = 0; k < chunk; ++k) {
            cg_pcg_spmv<<<G, RB, 0, c->stream>>>(g->colptr, g->rowval, g->nzval, g->p, n, g->Ap, c->d_partials);
            cg_sum2<<<1, RB, 0, c->stream>>>(c->d_partials, G, scal + 1, nullptr);
            cg_pcg_update<<<G, RB, 0, c->stream>>>(g->p, g->Ap, g->dinv, n, scal, g->u, g->r, c->d_partials);
            cg_sum2<<<1, RB, 0, c->stream>>>(c->d_partials, G, scal + 2, scal + 3);
            cg_pcg_dir<<<G, RB, 0, c->stream>>>(g->r, g->dinv, n, scal, g->p);
            cg_shift<<<1, 1, 0, c->stream>>>(scal);
        }
        c->launches += 6 * chunk;
        it += chunk;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        rr = c->h_scal[3];
        done = rr <= rtol * rtol * bb;      // checked every 16 iterations: the extra iterations only sharpen the solution
    }
    timer_stop(c, c->t_solve);
    if (info) {
        info->iterations = it; info->converged = done ? 1 : 0;
        info->relres = bb > 0.0 ? std::sqrt(rr / bb) : 0.0; info->bnorm = std::sqrt(bb); info->solve_ms = timer_ms(c->t_solve);
    }
    g->solved = true;
    if (!done) return set_err(c, HDG_ERR_NOT_CONVERGED, "CG did not converge in " + std::to_string(maxit) + " iterations");
    return HDG_OK;
}

hdg_status cg_errornorm(hdg_context* c, double* err2) {
    CgData* g = cg_get(c);
    if (!g || !g->solved) return set_err(c, HDG_ERR_INVALID, "hdg_cg_solve first");
    const int np = int(std::min<int64_t>(ceil_div(g->ncell, RB), 1024));
    cg_errornorm_kernel<<<np, RB, 0, c->stream>>>(g->tab, c->d_cellinfo, c->d_nodes, g->celldofs, g->u, g->ncell, c->d_partials);
    cg_sum2<<<1, RB, 0, c->stream>>>(c->d_partials, np, c->d_scal, nullptr);
    c->launches += 2;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    *err2 = c->h_scal[0];
    return HDG_OK;
}

double cg_meandiag(const hdg_context* c) { return c->cg ? static_cast<const CgData*>(c->cg)->meandiag : 0.0; }

}  // namespace hdg
