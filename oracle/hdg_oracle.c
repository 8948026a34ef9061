/*
 * C restatement of the reference's CPU hot path - TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file; the product (libhdg_b200.so) never does.  The reference itself
 * (Julia) cannot run in this image, so this loop-faithful port is what is timed as "the
 * reference's CPU path" (cpu_baseline.kind = "port").  It is validated against the numpy
 * oracle (oracle/hdg_oracle.py), which in turn is pinned on the reference's golden vectors
 * (tests/test_oracle_goldens.py).
 *
 * Follows examples/poisson2D_HDG.jl:58-186 (doassemble): reinit! per cell
 * (src/ScalarFunctionSpaces.jl:101-132), the quadrature loops :88-153 in the same loop order,
 * dense LU with partial pivoting for factorize(Array(Me)) :160 (LAPACK getrf semantics), the
 * solves and products :171-174, the COO append of assemble! (src/assembler.jl:31-40) and
 * sparse(I,J,V) (src/assembler.jl:47-49: column-major CSC, rows ascending, duplicates summed,
 * explicit zeros kept).  The reference is single-threaded; `nthreads > 1` parallelises the
 * element loop with OpenMP for the "all host cores" baseline.
 *
 * Tables are passed in by the caller (built by hdg_oracle.build_tables) with the numpy layouts:
 *   N[n][nq], dN[n][nq][2], E[n][nfq][3], T[nt][nfq], M[3][nq], qw[nq], fw[nfq]   (C order)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, nt, nq, nfq;
    const double *N, *dN, *E, *T, *M, *qw, *fw;
} tables_t;

static double source_f(double x, double y) {   /* examples/poisson2D_HDG.jl:55 */
    const double pi = 3.141592653589793;
    return 2 * (pi * pi) * sin(pi * x) * sin(pi * y);
}

/* LU with partial pivoting in place (row-major m x m), then solve nrhs columns of B (m x nrhs). */
static int lu_solve(double* A, int m, double* B, int nrhs, int* piv) {
    for (int k = 0; k < m; ++k) {
        int p = k;
        double best = fabs(A[k * m + k]);
        for (int i = k + 1; i < m; ++i)
            if (fabs(A[i * m + k]) > best) { best = fabs(A[i * m + k]); p = i; }
        piv[k] = p;
        if (best == 0.0) return k + 1;
        if (p != k) {
            for (int j = 0; j < m; ++j) { double t = A[k * m + j]; A[k * m + j] = A[p * m + j]; A[p * m + j] = t; }
            for (int j = 0; j < nrhs; ++j) { double t = B[k * nrhs + j]; B[k * nrhs + j] = B[p * nrhs + j]; B[p * nrhs + j] = t; }
        }
        double ip = 1.0 / A[k * m + k];
        for (int i = k + 1; i < m; ++i) {
            double l = A[i * m + k] * ip;
            A[i * m + k] = l;
            for (int j = k + 1; j < m; ++j) A[i * m + j] -= l * A[k * m + j];
            for (int j = 0; j < nrhs; ++j) B[i * nrhs + j] -= l * B[k * nrhs + j];
        }
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < nrhs; ++j) {
            double s = B[i * nrhs + j];
            for (int k = i + 1; k < m; ++k) s -= A[i * m + k] * B[k * nrhs + j];
            B[i * nrhs + j] = s / A[i * m + i];
        }
    return 0;
}

/* One cell: blocks, condensation.  Outputs Ke (m x t row-major), be_out (m), At (t x t row-major), bt (t). */
static int cell_work(const tables_t* tb, const double x[3][2], const int ori[3], double tau, const double* fq,
                     double* scratch, int* piv, double* Ke, double* be_out, double* At, double* bt) {
    const int n = tb->n, nt = tb->nt, nq = tb->nq, nfq = tb->nfq, nv = 2 * n, m = 3 * n, t = 3 * nt, nr = t + 1;
    double* Me = scratch;            /* m*m */
    double* R = Me + m * m;          /* m*(t+1): [-E;F | 0;be] -> solution */
    double* G = R + m * nr;          /* m*t : [E;F] */
    double* be = G + m * t;          /* n */
    memset(scratch, 0, sizeof(double) * (m * m + m * nr + m * t + n));
    /* reinit! */
    double J00 = x[1][0] - x[0][0], J01 = x[2][0] - x[0][0], J10 = x[1][1] - x[0][1], J11 = x[2][1] - x[0][1];
    double detJ = J00 * J11 - J01 * J10;
    if (!(detJ > 0.0)) return -1;
    double G00 = J11 / detJ, G01 = -J01 / detJ, G10 = -J10 / detJ, G11 = J00 / detJ;
    double wn[3][2] = {{-(J10 - J11), J00 - J01}, {-J11, J01}, {J10, -J00}}, dJf[3], nrm[3][2];
    for (int l = 0; l < 3; ++l) {
        dJf[l] = sqrt(wn[l][0] * wn[l][0] + wn[l][1] * wn[l][1]);
        nrm[l][0] = wn[l][0] / dJf[l];
        nrm[l][1] = wn[l][1] / dJf[l];
    }
    /* cell integrals :88-104 */
    for (int q = 0; q < nq; ++q) {
        double dO = detJ * tb->qw[q];
        for (int i = 0; i < nv; ++i) {
            int ci = i / n, ii = i % n;
            double dr = tb->dN[(ii * nq + q) * 2], ds = tb->dN[(ii * nq + q) * 2 + 1];
            double div = ci == 0 ? dr * G00 + ds * G10 : dr * G01 + ds * G11;
            double vi = tb->N[ii * nq + q];
            for (int j = 0; j < nv; ++j) {
                int cj = j / n, jj = j % n;
                if (ci == cj) Me[i * m + j] += (tb->N[jj * nq + q] * vi) * dO;          /* A */
            }
            for (int j = 0; j < n; ++j) {
                double b = (tb->N[j * nq + q] * div) * dO;
                Me[i * m + nv + j] -= b;                                                /* -B */
                Me[(nv + j) * m + i] += b;                                              /* B' */
            }
        }
    }
    /* rhs :106-114 */
    for (int q = 0; q < nq; ++q) {
        double dO = detJ * tb->qw[q];
        double fh;
        if (fq) fh = fq[q];
        else {
            double xq = tb->M[0 * nq + q] * x[0][0] + tb->M[1 * nq + q] * x[1][0] + tb->M[2 * nq + q] * x[2][0];
            double yq = tb->M[0 * nq + q] * x[0][1] + tb->M[1 * nq + q] * x[1][1] + tb->M[2 * nq + q] * x[2][1];
            fh = source_f(xq, yq);
        }
        for (int i = 0; i < n; ++i) be[i] += fh * tb->N[i * nq + q] * dO;
    }
    /* face integrals :116-153 */
    double* H = At;   /* accumulate He in At's storage, subtracted at the end */
    memset(H, 0, sizeof(double) * t * t);
    for (int l = 0; l < 3; ++l)
        for (int q = 0; q < nfq; ++q) {
            double dS = dJf[l] * tb->fw[q];
            int qo = ori[l] ? q : nfq - 1 - q;
            for (int i = 0; i < n; ++i) {
                double w = tb->E[(i * nfq + q) * 3 + l];
                for (int j = 0; j < n; ++j) Me[(nv + i) * m + nv + j] += tau * (tb->E[(j * nfq + q) * 3 + l] * w) * dS;   /* C */
                double wo = tb->E[(i * nfq + qo) * 3 + l];
                for (int j = 0; j < nt; ++j) G[(nv + i) * t + nt * l + j] += (tau * (tb->T[j * nfq + q] * wo)) * dS;        /* F */
            }
            for (int i = 0; i < nv; ++i) {
                int ci = i / n, ii = i % n;
                double vn = tb->E[(ii * nfq + qo) * 3 + l] * nrm[l][ci];
                for (int j = 0; j < nt; ++j) G[i * t + nt * l + j] += (tb->T[j * nfq + q] * vn) * dS;                      /* E */
            }
            for (int i = 0; i < nt; ++i)
                for (int j = 0; j < nt; ++j) H[(nt * l + i) * t + nt * l + j] += (tb->T[j * nfq + q] * tb->T[i * nfq + q]) * dS;
        }
    /* condensation :155-174 */
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < t; ++j) R[i * nr + j] = i < nv ? -G[i * t + j] : G[i * t + j];
        R[i * nr + t] = i < nv ? 0.0 : be[i - nv];
    }
    int info = lu_solve(Me, m, R, nr, piv);
    if (info) return info;
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < t; ++j) Ke[i * t + j] = R[i * nr + j];
        be_out[i] = R[i * nr + t];
    }
    for (int r = 0; r < t; ++r) {
        for (int c = 0; c < t; ++c) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += G[i * t + r] * R[i * nr + c];
            At[r * t + c] = s - H[r * t + c];
        }
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += G[i * t + r] * R[i * nr + t];
        bt[r] = -s;
    }
    return 0;
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/*
 * doassemble.  cells (ncell x 3) / cell_faces (ncell x 3) 1-based int64, nodes (nnode x 2).
 * Outputs: Ke_all (ncell*m*t), be_all (ncell*m) [may be NULL], rhs (ndof), and the CSC matrix:
 * colptr (ndof+1, 0-based), rowval/nzval with capacity t*t*ncell; *nnz_out = merged entry count.
 * Returns 0, -1 (bad geometry), >0 (singular pivot).
 */
int hdg_c_doassemble(int n, int nt, int nq, int nfq, const double* N, const double* dN, const double* E,
                     const double* T, const double* M, const double* qw, const double* fw, int64_t ncell,
                     int64_t nface, const int64_t* cells, const int64_t* cell_faces, const double* nodes, double tau,
                     const double* fq_all, int nthreads, double* Ke_all, double* be_all, double* rhs, int64_t* colptr,
                     int64_t* rowval, double* nzval, int64_t* nnz_out) {
    tables_t tb = {n, nt, nq, nfq, N, dN, E, T, M, qw, fw};
    const int m = 3 * n, t = 3 * nt, nr = t + 1;
    const int64_t ndof = nface * nt;
    int status = 0;
    /* COO triplets exactly as assemble! appends them: V = Ate column-major, I = gdof per column, J = gdof[j] */
    int64_t* I = (int64_t*)malloc(sizeof(int64_t) * (size_t)ncell * t * t);
    int64_t* Jc = (int64_t*)malloc(sizeof(int64_t) * (size_t)ncell * t * t);
    double* V = (double*)malloc(sizeof(double) * (size_t)ncell * t * t);
    double* bt_all = (double*)malloc(sizeof(double) * (size_t)ncell * t);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double* scratch = (double*)malloc(sizeof(double) * (m * m + m * nr + m * t + n));
        double* Ke = (double*)malloc(sizeof(double) * m * t);
        double* be = (double*)malloc(sizeof(double) * m);
        double* At = (double*)malloc(sizeof(double) * t * t);
        int* piv = (int*)malloc(sizeof(int) * m);
#pragma omp for schedule(static)
        for (int64_t c = 0; c < ncell; ++c) {
            double x[3][2];
            int ori[3];
            const int64_t* v = cells + 3 * c;
            for (int k = 0; k < 3; ++k) { x[k][0] = nodes[2 * (v[k] - 1)]; x[k][1] = nodes[2 * (v[k] - 1) + 1]; }
            ori[0] = v[2] > v[1]; ori[1] = v[0] > v[2]; ori[2] = v[1] > v[0];   /* src/mesh.jl:51-54 */
            int info = cell_work(&tb, x, ori, tau, fq_all ? fq_all + c * nq : NULL, scratch, piv, Ke, be, At, bt_all + c * t);
            if (info) {
#pragma omp critical
                status = info;
                continue;
            }
            if (Ke_all) memcpy(Ke_all + (size_t)c * m * t, Ke, sizeof(double) * m * t);
            if (be_all) memcpy(be_all + (size_t)c * m, be, sizeof(double) * m);
            int64_t gdof[32];
            for (int l = 0; l < 3; ++l)
                for (int j = 1; j <= nt; ++j) gdof[l * nt + j - 1] = cell_faces[3 * c + l] * nt - (nt - j);   /* :176-181 */
            size_t o = (size_t)c * t * t;
            for (int j = 0; j < t; ++j)
                for (int i = 0; i < t; ++i) {
                    I[o + j * t + i] = gdof[i];
                    Jc[o + j * t + i] = gdof[j];
                    V[o + j * t + i] = At[i * t + j];
                }
        }
        free(scratch); free(Ke); free(be); free(At); free(piv);
    }
    if (status) { free(I); free(Jc); free(V); free(bt_all); return status; }
    /* rhs[gdof] += bte, sequential like the reference (order of additions matters bitwise) */
    memset(rhs, 0, sizeof(double) * ndof);
    for (int64_t c = 0; c < ncell; ++c)
        for (int l = 0; l < 3; ++l)
            for (int j = 0; j < nt; ++j) rhs[(cell_faces[3 * c + l] - 1) * nt + j] += bt_all[c * t + l * nt + j];
    /* sparse(I,J,V): counting sort by column, sort rows inside a column, sum duplicates */
    const int64_t ncoo = ncell * t * t;
    int64_t* cnt = (int64_t*)calloc(ndof + 1, sizeof(int64_t));
    for (int64_t k = 0; k < ncoo; ++k) cnt[Jc[k]]++;          /* Jc 1-based -> cnt[j] = entries of column j-1 */
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (ndof + 1));
    start[0] = 0;
    for (int64_t j = 0; j < ndof; ++j) start[j + 1] = start[j] + cnt[j + 1];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ndof);
    memcpy(fill, start, sizeof(int64_t) * ndof);
    int64_t* key = (int64_t*)malloc(sizeof(int64_t) * ncoo * 2);   /* (row, coo index) pairs */
    /* With nthreads > 1 the bucket fill and the per-column sort/merge run over all threads (the reference's sparse() is
       serial; this is the "all host cores" arm).  The result does not depend on the fill order: ties are re-sorted
       by coo index below. */
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t k = 0; k < ncoo; ++k) {
        int64_t p;
#pragma omp atomic capture
        p = fill[Jc[k] - 1]++;
        key[2 * p] = I[k] - 1;
        key[2 * p + 1] = k;
    }
    /* pass 1: sort every column by row, count its distinct rows */
    int64_t* nuniq = cnt;   /* reuse */
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t j = 0; j < ndof; ++j) {
        int64_t a = start[j], b = start[j + 1], u = 0;
        qsort(key + 2 * a, (size_t)(b - a), 2 * sizeof(int64_t), cmp_i64);
        for (int64_t p = a; p < b; ++p) u += (p == a || key[2 * p] != key[2 * (p - 1)]);
        nuniq[j] = u;
    }
    colptr[0] = 0;
    for (int64_t j = 0; j < ndof; ++j) colptr[j + 1] = colptr[j] + nuniq[j];
    const int64_t nnz = colptr[ndof];
    /* pass 2: duplicates are summed in COO order (ascending coo index), as sparse() does */
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t j = 0; j < ndof; ++j) {
        int64_t a = start[j], b = start[j + 1], o = colptr[j];
        int64_t p = a;
        while (p < b) {
            int64_t row = key[2 * p], q = p;
            while (q + 1 < b && key[2 * (q + 1)] == row) ++q;
            double s2;
            if (q == p) s2 = V[key[2 * p + 1]];
            else {
                int64_t cntd = q - p + 1, idx[8];
                for (int64_t r = 0; r < cntd && r < 8; ++r) idx[r] = key[2 * (p + r) + 1];
                for (int64_t r = 1; r < cntd && r < 8; ++r) { int64_t v2 = idx[r]; int64_t u = r - 1; while (u >= 0 && idx[u] > v2) { idx[u + 1] = idx[u]; --u; } idx[u + 1] = v2; }
                s2 = 0.0;
                for (int64_t r = 0; r < cntd && r < 8; ++r) s2 += V[idx[r]];
            }
            rowval[o] = row;
            nzval[o] = s2;
            ++o;
            p = q + 1;
        }
    }
    *nnz_out = nnz;
    free(cnt); free(start); free(fill); free(key); free(I); free(Jc); free(V); free(bt_all);
    return 0;
}

/*
 * Jacobi-PCG on the sign-fixed system D K x = D b (D = -1 on free rows, +1 on Dirichlet rows),
 * K given as CSC == CSR (symmetric pattern; values symmetric to rounding - the transpose is used,
 * which for CG is immaterial).  The CPU stand-in for K \ b (examples/poisson2D_HDG.jl:195).
 * Returns the iteration count; *relres_out = ||r|| / ||b||.
 */
int hdg_c_pcg(int64_t ndof, const int64_t* colptr, const int64_t* rowval, const double* nzval, const double* b,
              const uint8_t* isbc, double rtol, int maxit, int nthreads, double* x, double* relres_out) {
    double* r = (double*)malloc(sizeof(double) * ndof);
    double* p = (double*)malloc(sizeof(double) * ndof);
    double* Ap = (double*)malloc(sizeof(double) * ndof);
    double* dinv = (double*)malloc(sizeof(double) * ndof);
    if (nthreads < 1) nthreads = 1;
    double rz = 0.0, bb = 0.0;
    for (int64_t i = 0; i < ndof; ++i) {
        double sg = isbc[i] ? 1.0 : -1.0, d = 0.0;
        for (int64_t k = colptr[i]; k < colptr[i + 1]; ++k)
            if (rowval[k] == i) d = nzval[k];
        dinv[i] = 1.0 / (sg * d);
        r[i] = sg * b[i];
        p[i] = dinv[i] * r[i];
        x[i] = 0.0;
        rz += r[i] * p[i];
        bb += r[i] * r[i];
    }
    int it = 0;
    double rr = bb;
    if (bb > 0.0)
        for (it = 1; it <= maxit; ++it) {
            double pap = 0.0;
#pragma omp parallel for num_threads(nthreads) reduction(+ : pap) schedule(static)
            for (int64_t i = 0; i < ndof; ++i) {
                double s = 0.0;
                for (int64_t k = colptr[i]; k < colptr[i + 1]; ++k) s += nzval[k] * p[rowval[k]];
                s = isbc[i] ? s : -s;
                Ap[i] = s;
                pap += p[i] * s;
            }
            double alpha = rz / pap, rzn = 0.0;
            rr = 0.0;
#pragma omp parallel for num_threads(nthreads) reduction(+ : rzn, rr) schedule(static)
            for (int64_t i = 0; i < ndof; ++i) {
                x[i] += alpha * p[i];
                r[i] -= alpha * Ap[i];
                rzn += r[i] * dinv[i] * r[i];
                rr += r[i] * r[i];
            }
            if (rr <= rtol * rtol * bb) break;
            double beta = rzn / rz;
            rz = rzn;
#pragma omp parallel for num_threads(nthreads) schedule(static)
            for (int64_t i = 0; i < ndof; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
        }
    if (relres_out) *relres_out = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    free(r); free(p); free(Ap); free(dinv);
    return it > maxit ? maxit : it;
}

/*
 * rectangle_mesh(TriangleCell,(nx,ny),LL,UR), src/generate_mesh.jl:101-143: nodes by _generate_2d_nodes! (:1-18, the same
 * floating-point expression), two triangles per quad made counter-clockwise by _check_node_data (:49-57), and the sequential
 * first-encounter face numbering of _build_cells (:20-46) - the reference's Dict{(min,max) => face} is an open-addressing hash
 * table here.  Outputs (1-based, the numpy oracle's layouts): cells ncell x 3, cell_faces ncell x 3, nodes nnode x 2,
 * faces nface x 4 row-major (v1 v2 cell1 cell2|0).  Returns the number of faces.  Sequential like the reference.
 */
__attribute__((optimize("fp-contract=off")))
static void ccw(const double* nodes, int64_t n1, int64_t n2, int64_t n3, int64_t out[3]) {
    const double ax = nodes[2 * (n2 - 1)] - nodes[2 * (n1 - 1)], ay = nodes[2 * (n2 - 1) + 1] - nodes[2 * (n1 - 1) + 1];
    const double bx = nodes[2 * (n3 - 1)] - nodes[2 * (n1 - 1)], by = nodes[2 * (n3 - 1) + 1] - nodes[2 * (n1 - 1) + 1];
    out[0] = n1;
    if (ax * by - ay * bx < 0) { out[1] = n3; out[2] = n2; } else { out[1] = n2; out[2] = n3; }
}

/* no FMA contraction: the coordinates must be the reference's separately rounded products and sums, bit for bit */
__attribute__((optimize("fp-contract=off")))
int64_t hdg_c_rectangle_mesh(int64_t nx, int64_t ny, double llx, double lly, double urx, double ury, int64_t* cells,
                             int64_t* cell_faces, double* nodes, int64_t* faces) {
    const int64_t nnx = nx + 1, nny = ny + 1;
    const double LRx = urx, LRy = lly, ULx = llx, ULy = ury;
    int64_t p = 0;
    for (int64_t i = 0; i < nny; ++i) {
        const double rb = (double)i / (double)(nny - 1);
        const double x0 = llx * (1 - rb) + rb * ULx, x1 = LRx * (1 - rb) + rb * urx;
        const double y0 = lly * (1 - rb) + rb * ULy, y1 = LRy * (1 - rb) + rb * ury;
        for (int64_t j = 0; j < nnx; ++j) {
            const double r = (double)j / (double)(nnx - 1);
            nodes[2 * p] = x0 * (1 - r) + r * x1;
            nodes[2 * p + 1] = y0 * (1 - r) + r * y1;
            ++p;
        }
    }
    const int64_t ncell = 2 * nx * ny;
    /* hash table: key = (min << 32 | max), value = face id */
    uint64_t cap = 16;
    while (cap < (uint64_t)(3 * ncell) * 2) cap <<= 1;
    uint64_t* keys = (uint64_t*)calloc(cap, sizeof(uint64_t));
    int64_t* vals = (int64_t*)malloc(cap * sizeof(int64_t));
    static const int EN[3][2] = {{1, 2}, {2, 0}, {0, 1}};   /* reference_edge_nodes ((2,3),(3,1),(1,2)), src/shapes.jl:14 */
    int64_t face_idx = 0, n_el = 0;
    for (int64_t j = 1; j <= ny; ++j)
        for (int64_t i = 1; i <= nx; ++i)
            for (int half = 0; half < 2; ++half) {
                int64_t el[3];
#define NA(I, J) ((I) + ((J) - 1) * nnx)
                if (half == 0) ccw(nodes, NA(i, j), NA(i + 1, j), NA(i, j + 1), el);
                else ccw(nodes, NA(i + 1, j), NA(i + 1, j + 1), NA(i, j + 1), el);
#undef NA
                ++n_el;
                for (int e = 0; e < 3; ++e) {
                    const int64_t v1 = el[EN[e][0]], v2 = el[EN[e][1]];
                    const uint64_t lo = (uint64_t)(v1 < v2 ? v1 : v2), hi = (uint64_t)(v1 < v2 ? v2 : v1);
                    const uint64_t key = (lo << 32) | hi;
                    uint64_t h = (key * 0x9E3779B97F4A7C15ull) & (cap - 1);
                    while (keys[h] != 0 && keys[h] != key) h = (h + 1) & (cap - 1);
                    if (keys[h] == key) {
                        const int64_t fid = vals[h];
                        cell_faces[3 * (n_el - 1) + e] = fid;
                        if (n_el != faces[4 * (fid - 1) + 3]) faces[4 * (fid - 1) + 3] = n_el;
                    } else {
                        ++face_idx;
                        keys[h] = key; vals[h] = face_idx;
                        faces[4 * (face_idx - 1)] = v1; faces[4 * (face_idx - 1) + 1] = v2;
                        faces[4 * (face_idx - 1) + 2] = n_el; faces[4 * (face_idx - 1) + 3] = 0;
                        cell_faces[3 * (n_el - 1) + e] = face_idx;
                    }
                }
                cells[3 * (n_el - 1)] = el[0]; cells[3 * (n_el - 1) + 1] = el[1]; cells[3 * (n_el - 1) + 2] = el[2];
            }
    free(keys); free(vals);
    return face_idx;
}

/*
 * get_u_sigma!(sigma_h, u_h, uhat_h, uhat, K_e, b_e, mesh), examples/poisson2D_HDG.jl:197-212 (the hard-coded nt = 2 slice of
 * :205-206 generalised to nt): per cell gather uhat_e in local-face order, dof = K_e uhat_e + b_e.  Outputs are the m_values
 * arrays as the numpy oracle returns them: sigma ncell x 2n, u ncell x n (row-major).
 */
void hdg_c_recover(int n, int nt, int64_t ncell, const int64_t* cell_faces, const double* Ke, const double* be,
                   const double* uhat, int nthreads, double* sigma, double* u) {
    const int m = 3 * n, t = 3 * nt;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t c = 0; c < ncell; ++c) {
        double ue[15];
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < nt; ++a) ue[k * nt + a] = uhat[nt * (cell_faces[3 * c + k] - 1) + a];
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int j = 0; j < t; ++j) s += Ke[(c * m + i) * t + j] * ue[j];
            s += be[c * m + i];
            if (i < 2 * n) sigma[c * 2 * n + i] = s; else u[c * n + (i - 2 * n)] = s;
        }
    }
}

/* errornorm(u_h, u_ex) with u_ex = sin(pi x) sin(pi y) (examples/poisson2D_HDG.jl:216), src/DiscreteFunctions.jl:97-120:
 * squared L2 error with the cell rule.  N[n][nq], M[3][nq], qw[nq]; u ncell x n row-major. */
double hdg_c_errornorm(int n, int nq, const double* N, const double* M, const double* qw, int64_t ncell, const int64_t* cells,
                       const double* nodes, const double* u, int nthreads) {
    const double pi = 3.141592653589793;
    double tot = 0.0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : tot)
    for (int64_t c = 0; c < ncell; ++c) {
        double x[3][2];
        for (int g = 0; g < 3; ++g) { x[g][0] = nodes[2 * (cells[3 * c + g] - 1)]; x[g][1] = nodes[2 * (cells[3 * c + g] - 1) + 1]; }
        const double detJ = (x[1][0] - x[0][0]) * (x[2][1] - x[0][1]) - (x[2][0] - x[0][0]) * (x[1][1] - x[0][1]);
        double el = 0.0;
        for (int q = 0; q < nq; ++q) {
            double uq = 0.0, xq = 0.0, yq = 0.0;
            for (int i = 0; i < n; ++i) uq += u[c * n + i] * N[i * nq + q];
            for (int g = 0; g < 3; ++g) { xq += M[g * nq + q] * x[g][0]; yq += M[g * nq + q] * x[g][1]; }
            const double d = uq - sin(pi * xq) * sin(pi * yq);
            el += d * d * (detJ * qw[q]);
        }
        tot += el;
    }
    return tot;
}

int hdg_c_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
