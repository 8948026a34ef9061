/*
 * libhdg_b200 - C ABI of the B200-native HDG Poisson hot path.
 *
 * Drop-in boundary for Paulms/HDiscontinuousGalerkin.jl (reference, Julia).  The reference has
 * no FFI of its own; the boundary is the set of script-level calls of
 * examples/poisson2D_HDG.jl (doassemble :58-186, apply! :194, K\b :195, get_uσ! :197-212,
 * errornorm :217).  Each entry point below names the reference code it replaces.  The Julia
 * side binds these with `ccall` (see INTEGRATION.md and julia/HDGB200.jl); the test-suite binds
 * them with Python ctypes.
 *
 * Conventions
 *  - plain C types only; all ids that mirror reference data are Int64 and 1-based, exactly as
 *    the Julia structs hold them (mesh.cells = NTuple{3,Int} nodes + NTuple{3,Int} faces,
 *    mesh.faces = column-major Matrix{Int} nface x 4, src/mesh.jl:17-20,43-49).
 *  - every function returns an hdg_status; hdg_last_error(ctx) gives the message.
 *  - the caller owns every host buffer; the library copies during the call and keeps no host
 *    pointer.  All device memory lives inside the opaque context.
 *  - calls are synchronous on return unless stated; one context per host thread, not re-entrant.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    HDG_ERR_CUDA.
 */
#ifndef HDG_B200_H
#define HDG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdg_context hdg_context;

typedef enum hdg_status {
    HDG_OK = 0,
    HDG_ERR_INVALID = 1,        /* bad argument / call order                                      */
    HDG_ERR_BAD_GEOMETRY = 2,   /* det(J) <= 0: ArgumentError of src/ScalarFunctionSpaces.jl:110   */
    HDG_ERR_UNSUPPORTED_RULE = 3, /* ArgumentError of src/quadrature.jl:24 / StrangQuad.jl:58      */
    HDG_ERR_SINGULAR_LOCAL = 4, /* LAPACK SingularException of factorize, poisson2D_HDG.jl:160     */
    HDG_ERR_CUDA = 5,           /* CUDA runtime failure, or no device                              */
    HDG_ERR_NCCL = 6,
    HDG_ERR_NOT_CONVERGED = 7,  /* PCG hit maxit (the solution is still returned)                  */
    HDG_ERR_NOT_BOUNDARY = 8    /* AssertionError of src/boundary.jl:22                            */
} hdg_status;

/* Script constants / keyword arguments of the reference driver gathered in one struct. */
typedef struct hdg_params {
    int32_t order;        /* k of Dubiner{2,RefTetrahedron,k} / Legendre{1,RefTetrahedron,k}; 1..4  */
    int32_t quad_degree;  /* keyword quad_degree of ScalarFunctionSpace (src/ScalarFunctionSpaces.jl:24-25);
                             <= 0 selects the reference default order+1                            */
    double  tau;          /* stabilisation, doassemble(...; tau=1.0) poisson2D_HDG.jl:58            */
    int32_t source_id;    /* 0: caller supplies f at quadrature points (hdg_set_source_values);
                             1: f = 2*pi^2*sin(pi x)sin(pi y), poisson2D_HDG.jl:55                  */
    int32_t device;       /* CUDA device ordinal; -1 = current device                               */
    int32_t local_solver; /* 0: block elimination on the reference-matrix form (default);
                             reserved: 1 = literal quadrature + dense partial-pivot LU             */
    int32_t reserved;
} hdg_params;

/* Sizes derived from params + mesh, for sizing caller buffers. */
typedef struct hdg_sizes {
    int64_t ncell, nnode, nface, nbface;
    int32_t n;        /* scalar dofs per cell  (k+1)(k+2)/2          */
    int32_t nt;       /* trace dofs per face   k+1                   */
    int32_t m;        /* local system size 3n                        */
    int32_t t;        /* local trace dofs 3nt                        */
    int32_t nq;       /* cell quadrature points                      */
    int32_t nfq;      /* face quadrature points                      */
    int64_t ndof;     /* nt*nface, size of the trace system          */
    int64_t nnz;      /* stored entries of sparse(I,J,V)             */
} hdg_sizes;

typedef struct hdg_solve_info {
    int32_t iterations;
    int32_t converged;
    double  relres;       /* ||r||_2 / ||b||_2 of the sign-fixed system at exit */
    double  bnorm;
    double  solve_ms;     /* device time of the PCG loop (CUDA events)          */
} hdg_solve_info;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Builds the reference tables for (order, quad_degree) - the job of ScalarFunctionSpace /
 * VectorFunctionSpace / ScalarTraceFunctionSpace constructors (src/ScalarFunctionSpaces.jl:31-99,
 * src/TraceFunctionSpaces.jl:11-28) - and creates the device context. */
hdg_status hdg_create(const hdg_params* params, hdg_context** ctx_out);
void       hdg_destroy(hdg_context* ctx);
const char* hdg_last_error(const hdg_context* ctx);   /* ctx may be NULL: last create error */
const char* hdg_version(void);

/* ---- mesh --------------------------------------------------------------------------------
 * hdg_set_mesh: hand over a PolygonalMesh{2,3,3} (src/mesh.jl:43-49) as the Julia structs lay
 * it out.  cells: ncell x 6 Int64 row-major (3 node ids, 3 face ids) == Vector{Cell{2,3,3}};
 * nodes: nnode x 2 doubles == Vector{Node{2,Float64}}; faces: COLUMN-major nface x 4 Int64
 * (v1 v2 cell1 cell2|0) == Matrix{Int}; bfaces: the Dirichlet face set (getfaceset(mesh,
 * "boundary"), any order, 1-based).  Replaces CellIterator/reinit! gathers (src/iterator.jl:48-57).
 * faces may be NULL: the face table is then rebuilt on the device from the cells (the first cell of a face is
 * the adjacent cell with the smaller id, exactly what the first-encounter numbering produces), which saves
 * the host-to-device copy of the largest array; nface must still be given.
 * After hdg_comm_init every rank passes the SAME whole mesh (faces required, first-encounter numbering) and keeps
 * a contiguous range of cells and of the faces they create, plus ghost cells / ghost columns (hdg_get_partition). */
hdg_status hdg_set_mesh(hdg_context* ctx,
                        const int64_t* cells, int64_t ncell,
                        const double* nodes, int64_t nnode,
                        const int64_t* faces, int64_t nface,
                        const int64_t* bfaces, int64_t nbface);

/* rectangle_mesh(TriangleCell,(nx,ny),LL,UR) (src/generate_mesh.jl:101-143) generated on the
 * device with the reference's node coordinates and first-encounter face numbering; the
 * "boundary" face set becomes the Dirichlet set.  After hdg_comm_init every rank passes the SAME global
 * (nx, ny, LL, UR) and keeps the strip of quad rows it owns (hdg_get_partition) plus a one-cell ghost layer. */
hdg_status hdg_set_rectangle_mesh(hdg_context* ctx, int64_t nx, int64_t ny,
                                  double llx, double lly, double urx, double ury);

/* Dirichlet(u_hat, mesh, faceset, g) for another face set than the one handed over with the mesh (src/boundary.jl:7-42 takes
 * any named set of boundary faces, e.g. "bottom"/"right"/"top"/"left" of rectangle_mesh, src/generate_mesh.jl:60-89): replaces
 * the Dirichlet face set (1-based ids, any order).  Boundary faces outside the set keep the natural condition.  A face with
 * two cells gives HDG_ERR_NOT_BOUNDARY (the @assert of src/boundary.jl:22) and leaves an empty set.  Call after the mesh is set
 * and before hdg_apply_dirichlet; single GPU. */
hdg_status hdg_set_dirichlet_faces(hdg_context* ctx, const int64_t* bfaces, int64_t nbface);

/* First-encounter face numbering of an arbitrary triangle list on the device - the job of _build_cells
 * (src/generate_mesh.jl:20-46) / parse_cells! (src/triangle_mesh.jl:48-108), which are sequential hash-table
 * inserts in the reference.  tri: ncell x 3 Int64 node ids (1-based, either orientation: clockwise cells get
 * vertices 2,3 swapped like _check_node_data, src/generate_mesh.jl:49-57).  Outputs in the Julia layouts:
 * cells_out ncell x 6 (nodes, faces), faces_out nface x 4 column-major (v1 v2 cell1 cell2|0), bit-identical to the
 * reference's numbering.  Call with cells_out = faces_out = NULL to obtain *nface_out first. */
hdg_status hdg_number_faces(hdg_context* ctx, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode,
                            int64_t* cells_out, int64_t* faces_out, int64_t faces_capacity, int64_t* nface_out);

/* Locality-preserving cell order for hdg_set_mesh on several GPUs: perm_out[i] (1-based) = the input cell that comes i-th along
 * the Morton curve through the cell centroids (computed and sorted on the device; cells with equal keys keep their input
 * order).  hdg_set_mesh partitions by contiguous cell-id ranges, so a mesh whose cells are renumbered in this order - cells
 * permuted, faces renumbered by hdg_number_faces, which is what the reference's parse_cells! (src/triangle_mesh.jl:48-108)
 * would produce for the permuted element list - splits into compact patches whatever the order of the generator's output. */
hdg_status hdg_order_cells(hdg_context* ctx, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode,
                           int64_t* perm_out);

/* Deterministic interior-node jitter (fraction of h) to defeat translation invariance in
 * benchmarks (SURVEY Appendix A integrity note).  Applies to the device mesh in place. */
hdg_status hdg_perturb_nodes(hdg_context* ctx, double fraction, uint64_t seed);

hdg_status hdg_get_sizes(const hdg_context* ctx, hdg_sizes* out);
/* Download the device mesh in the Julia layouts described at hdg_set_mesh (1-based). */
hdg_status hdg_get_mesh(hdg_context* ctx, int64_t* cells, double* nodes, int64_t* faces,
                        int64_t* bfaces_sorted);

/* Reference tables, for checking against the Julia-side tables.  name is one of
 * "qpoints"(nq*2) "qweights"(nq) "fpoints"(nfq) "fweights"(nfq) "N"(n*nq, N[i,q] at i+n*q)
 * "dNdxi"(n*nq*2) "E"(n*nfq*3, [i + n*(p + nfq*l)]) "T"(nt*nfq), or one of the derived reference
 * matrices the kernels consume (row-major, hdg_tables.h): "Tr" "Ts" "Prr" "Prs" "Pss" (n*n),
 * "Chat"(3*n*n) "Fhat" "MF" "Qr" "Qs" (n*t) "Hhat"(nt*nt).  Returns the number of
 * doubles written through *count (buf may be NULL to query). */
hdg_status hdg_get_table(const hdg_context* ctx, const char* name, double* buf, int64_t* count);
/* Same tables without a context (host-only table builder; needs no device). */
hdg_status hdg_ref_table(int32_t order, int32_t quad_degree, const char* name, double* buf, int64_t* count);

/* value(ip, j, xi) / gradient_value(ip, j, xi) of the bases the tables are built from (src/basis.jl:65-86 Dubiner on the
 * reference triangle, kind 0, xi = (r, s); :351-354 orthonormal Legendre on (0,1), kind 1, xi = (x)); j is 1-based.  grad may
 * be NULL.  Host-only (needs no device): lets the reference's basis tests run against the library's own evaluation. */
hdg_status hdg_basis_value(int32_t kind, int32_t j, const double* xi, double* value, double* grad);

/* ---- source -------------------------------------------------------------------------------
 * Pre-evaluated f(x_q): ncell x nq doubles, fq[c*nq+q] = f(spatial_coordinate(Wh,q,coords_c))
 * (function_value, src/DiscreteFunctions.jl:6-24).  Required when source_id == 0.
 * Several GPUs: the rows are this rank's LOCAL cells - the owned cells in id order (hdg_get_partition out[0..1]) followed by
 * the ghost cells it recomputes, in the order of hdg_get_ghost_cells (out[6] of them): (ncell_own + nghost) x nq doubles. */
hdg_status hdg_set_source_values(hdg_context* ctx, const double* fq);

/* ---- hot path -----------------------------------------------------------------------------
 * hdg_assemble == doassemble(Vh,Wh,Mh,tau), poisson2D_HDG.jl:58-186: local blocks (:88-153),
 * static condensation (:155-174), dof map and scatter (:176-185, src/assembler.jl:31-60). */
hdg_status hdg_assemble(hdg_context* ctx);
/* apply!(K,b,dbc) with dbc = Dirichlet(u_hat, mesh, "boundary", g), src/boundary.jl:121-158.
 * values: nbface*nt doubles ordered like Dirichlet.prescribed_dofs (ascending face, then dof),
 * or NULL for g == 0. */
hdg_status hdg_apply_dirichlet(hdg_context* ctx, const double* values);
/* u_hat = K \ b (poisson2D_HDG.jl:195) by Jacobi-PCG on the sign-fixed system. */
hdg_status hdg_solve(hdg_context* ctx, double rtol, int32_t maxit, hdg_solve_info* info);
/* Preconditioner of hdg_solve: 0 = Jacobi (default), 1 = block-Jacobi with the nt x nt face-diagonal blocks
 * (fewer iterations for k >= 2; identical to Jacobi for k = 1 where the blocks are diagonal), 2 = block-Jacobi plus a
 * geometric multigrid V-cycle on the P1 vertex space (35-45 iterations independent of h and k).  The mesh must be the
 * triangulation of rectangle_mesh: from hdg_set_rectangle_mesh (one GPU, or strips over several GPUs - the vertex hierarchy
 * is then replicated and the partial stencils / vertex residuals are all-reduced), or handed over as arrays through
 * hdg_set_mesh on one GPU, where the grid numbering is recognised in the face table; any other configuration makes hdg_solve
 * return HDG_ERR_INVALID. */
hdg_status hdg_set_preconditioner(hdg_context* ctx, int32_t id);
/* get_uσ!(σ_h,u_h,û_h,û,K_e,b_e,mesh), poisson2D_HDG.jl:197-212 (nt-general). */
hdg_status hdg_recover(hdg_context* ctx);
/* errornorm(u_h,u_ex) (squared L2, src/DiscreteFunctions.jl:97-120).  exact_id 1:
 * sin(pi x) sin(pi y) (poisson2D_HDG.jl:216). */
hdg_status hdg_errornorm(hdg_context* ctx, int32_t exact_id, double* err2);

/* errornorm(u_h, u_ex) for any u_ex: uex_q holds u_ex at the cell quadrature points, ncell x nq doubles laid out like the
 * source values of hdg_set_source_values (uex_q[c*nq+q] = u_ex(spatial_coordinate(Wh,q,coords_c))).  Several GPUs: every
 * rank passes the values of the cells it owns; the result is the global sum. */
hdg_status hdg_errornorm_values(hdg_context* ctx, const double* uex_q, double* err2);

/* Asynchronous variant of hdg_assemble for timing loops: enqueues on the context stream and
 * returns; pair with hdg_sync.  hdg_stream returns the cudaStream_t as an integer handle. */
hdg_status hdg_assemble_async(hdg_context* ctx);
hdg_status hdg_sync(hdg_context* ctx);
uint64_t   hdg_stream(const hdg_context* ctx);

/* ---- results ------------------------------------------------------------------------------
 * CSC pattern of sparse(I,J,V) (src/assembler.jl:47-49): colptr ndof+1, rowval nnz, Int64,
 * 1-based, rows ascending per column. */
hdg_status hdg_get_pattern(hdg_context* ctx, int64_t* colptr, int64_t* rowval);
/* nzval in that CSC order (current state of K: raw after hdg_assemble, modified after apply). */
hdg_status hdg_get_values(hdg_context* ctx, double* nzval);
hdg_status hdg_get_rhs(hdg_context* ctx, double* rhs);
hdg_status hdg_get_trace(hdg_context* ctx, double* uhat);
hdg_status hdg_set_trace(hdg_context* ctx, const double* uhat);
hdg_status hdg_get_meandiag(const hdg_context* ctx, double* m);
/* K_element[cell] (m x t, column-major like a Julia Matrix) and b_element[cell] (m);
 * cell is 1-based. */
hdg_status hdg_get_local(hdg_context* ctx, int64_t cell, double* Ke, double* be);
/* Condensed element matrix / vector Ate (t x t column-major) and bte (t) of one cell,
 * recomputed on the device (debug / parity). */
hdg_status hdg_get_condensed(hdg_context* ctx, int64_t cell, double* Ate, double* bte);
/* TrialFunction.m_values (src/DiscreteFunctions.jl:27-54), column-major:
 * sigma ncell x 2n, u ncell x n, uhat ncell x nt x 3.  Any pointer may be NULL. */
hdg_status hdg_get_mvalues(hdg_context* ctx, double* sigma, double* u, double* uhat_h);

/* nodal_avg(u_h) (src/DiscreteFunctions.jl:81-95): the discontinuous u_h evaluated at the vertices of every cell
 * (value(u_h,node,cell), :70-79) and averaged over the cells sharing a node; nnode doubles.  Plotting helper. */
hdg_status hdg_nodal_avg(hdg_context* ctx, double* out);

/* ---- CG side of the exported API -------------------------------------------
 * examples/poisson2D_CG.jl on the mesh of the context: ContinuousLagrange{2,RefTetrahedron,order} (order 1 or 2), one scalar
 * field, quad_degree = order + 1 (the default of ScalarFunctionSpace), f = 2 pi^2 sin(pi x) sin(pi y).  One GPU.
 * hdg_cg_setup == DofHandler([u_h], mesh) (_distribute_dofs, src/dofhandler.jl:80-152) + create_sparsity_pattern(dh)
 * (:181-220) on the device; *ndofs receives ndofs(dh). */
hdg_status hdg_cg_setup(hdg_context* ctx, int32_t order, int64_t* ndofs);
/* out = {ndofs, nnz of the pattern, ndofs_per_cell, ncell} */
hdg_status hdg_cg_get_sizes(hdg_context* ctx, int64_t out[4]);
/* dh.cell_dofs (ncell x ndofs_per_cell, 1-based, cell-major == the reference's flat vector) and the CSC pattern of
 * create_sparsity_pattern (colptr ndofs+1, rowval nnz; Int64, 1-based).  Any pointer may be NULL. */
hdg_status hdg_cg_get_dofhandler(hdg_context* ctx, int64_t* cell_dofs, int64_t* colptr, int64_t* rowval);
/* doassemble(Wh, K, dh) of examples/poisson2D_CG.jl:72-126: start_assemble, element matrices, assemble!(assembler, dofs, fe, Ke)
 * (src/assembler.jl:62-137). */
hdg_status hdg_cg_assemble(hdg_context* ctx);
/* dbc = Dirichlet(u_h, dh, "boundary", [0.0]) (src/boundary.jl:48-96) on the context's Dirichlet face set + apply!(K, b, dbc)
 * (:121-158). */
hdg_status hdg_cg_apply_dirichlet(hdg_context* ctx);
/* u = K \ b (examples/poisson2D_CG.jl:134) by Jacobi-PCG. */
hdg_status hdg_cg_solve(hdg_context* ctx, double rtol, int32_t maxit, hdg_solve_info* info);
/* K.nzval (pattern order), b, u (ndofs each).  Any pointer may be NULL. */
hdg_status hdg_cg_get_system(hdg_context* ctx, double* nzval, double* rhs, double* u);
/* reconstruct!(u_h, u, dh) + errornorm(u_h, u_ex) with u_ex = sin(pi x) sin(pi y) (squared L2, test/test_CGExample.jl:92-93). */
hdg_status hdg_cg_errornorm(hdg_context* ctx, double* err2);
hdg_status hdg_cg_get_meandiag(const hdg_context* ctx, double* m);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------
 * The caller (torch.distributed / MPI / Julia Distributed) creates a 128-byte ncclUniqueId on
 * rank 0 with hdg_comm_unique_id, broadcasts it, and every rank calls hdg_comm_init.  After
 * that hdg_set_rectangle_mesh partitions quad rows across ranks, hdg_solve exchanges halo
 * trace values and all-reduces the dot products over NCCL, hdg_errornorm all-reduces. */
hdg_status hdg_comm_unique_id(uint8_t id_out[128]);
hdg_status hdg_comm_init(hdg_context* ctx, int32_t rank, int32_t nranks, const uint8_t id[128]);
/* What this rank owns, global 0-based half-open ranges: out = {cell_begin, cell_end, face_begin, face_end,
 * ncell_global, nface_global, ghost cells held, ghost faces held}.  Cell and face ids are the reference's
 * global numbering minus one; trace dofs of face f are nt*f .. nt*f+nt-1. */
hdg_status hdg_get_partition(const hdg_context* ctx, int64_t out[8]);
/* Global 0-based ids of the ghost cells this rank holds (out[6] of hdg_get_partition), in local order: the cells of other
 * ranks that touch an owned face and are recomputed here instead of exchanged. */
hdg_status hdg_get_ghost_cells(const hdg_context* ctx, int64_t* ids);
/* Measurement helper: mean latency (microseconds) of one all-to-all mailbox exchange over peer memory. */
hdg_status hdg_comm_pingpong(hdg_context* ctx, int32_t iters, double* usec_per_exchange);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* Device time (ms, CUDA events on the context stream) of the kernels of the last call of the
 * named phase: "assemble" (memsets + element kernel), "element_kernel", "apply", "solve",
 * "recover", "errornorm"; inside "solve": "mg_setup" (operators of the vertex hierarchy) and "solve_loop" (the PCG iterations). */
hdg_status hdg_last_phase_ms(const hdg_context* ctx, const char* phase, double* ms);
/* FP64 FMA throughput of the context's device (TFLOP/s, best of 3 launches of a register-only DFMA kernel, CUDA events):
 * the denominator for the FP64 roofline of the k >= 2 element kernels - MEASURED_PEAKS.json holds no FP64 figure. */
hdg_status hdg_measure_fp64_peak(hdg_context* ctx, double* tflops);
/* Stage timing of the multigrid V-cycle kernel (library loaded with HDG_MG_TRACE=1 in the environment): microseconds since the
 * kernel start at the entry and exit of every grid barrier of the LAST V-cycle; returns the number of values written. */
int32_t    hdg_mg_trace(hdg_context* ctx, double* usec, int32_t capacity);
/* Number of kernel launches issued by this context since creation. */
int64_t    hdg_launch_count(const hdg_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HDG_B200_H */
