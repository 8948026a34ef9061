// Internal declarations shared by the translation units of libhdg_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hdg_b200.h"
#include "hdg_tables.h"

namespace hdg {

// ---- compile-time sizes per order ---------------------------------------------------------
template <int K> struct Ord {
    static constexpr int n = (K + 1) * (K + 2) / 2;   // scalar dofs per cell
    static constexpr int nt = K + 1;                  // trace dofs per face
    static constexpr int m = 3 * n;                   // local system
    static constexpr int t = 3 * nt;                  // local trace dofs
    static constexpr int ke = m * (t + 1);            // doubles of [K_e | b_e] per cell
};

constexpr int CI = 8;        // int32 words per cell record (32 bytes)
constexpr int MAX_NQ = 40;    // largest supported cell rule: Grundmann-Moeller s=4 has 35 points
constexpr int MAX_NFQ = 12;

// Reference matrices in __constant__ memory, one instance per order (uniform-index reads fold
// into the DFMA operand).  Layouts: row-major.
template <int K> struct DevTables {
    double Tr[Ord<K>::n * Ord<K>::n], Ts[Ord<K>::n * Ord<K>::n];
    double Prr[Ord<K>::n * Ord<K>::n], Prs[Ord<K>::n * Ord<K>::n], Pss[Ord<K>::n * Ord<K>::n];
    double Chat[3 * Ord<K>::n * Ord<K>::n];
    double Fhat[Ord<K>::n * Ord<K>::t];
    // The four matrices that are read with a RUN-TIME column index (the right-hand side of a column and its MF term), packed per
    // (column, row): RT[(col * n + i) * 4 + {0: Fhat, 1: Qr, 2: Qs, 3: MF}].  An indexed constant load goes through the small
    // indexed-constant cache; row-major n x t matrices made the n loads of a column touch n different lines of each matrix
    // (ncu, k = 3: hit rate 57 %, the GPC-level constant cache at 63 % of its peak) - packed, a column is n * 32 contiguous bytes.
    double RT[Ord<K>::n * Ord<K>::t * 4];
    double Hhat[Ord<K>::nt * Ord<K>::nt];
    double WN[MAX_NQ * Ord<K>::n];        // w_q N[i,q] at [q*n + i]
    double Mgeo[MAX_NQ * 3];
    double qw[MAX_NQ];
    int nq;
    int pad;
};

// Raw tables for the literal-quadrature + LU kernel (global memory, small, L1/L2 resident).
struct RawTablesDev {
    const double* N;    // n*nq
    const double* dN;   // n*nq*2
    const double* E;    // n*nfq*3
    const double* T;    // nt*nfq
    const double* qw;   // nq
    const double* fw;   // nfq
    const double* Mgeo; // nq*3
    int n, nt, nq, nfq;
};

// Storage layout of [K_e | b_e]
enum KeLayout : int {
    KE_TILE32 = 0,   // tiles of 32 cells, entry-major inside: ((c/32)*ke + e)*32 + c%32 ; e = i*(t+1)+j
    KE_CELL = 1      // per cell contiguous, row-major m x (t+1): c*ke + i*(t+1) + j
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    float last_ms = 0.f;
    bool pending = false;
};

// Multi-GPU state (hdg_comm.cu): one process per GPU, NCCL over NVLink.  The mesh is split into
// strips of quad rows (contiguous cell-id and face-id ranges); every rank holds its owned cells and
// faces plus a one-cell-deep ghost layer above and the ghost faces it references.
constexpr int MAXR = 8;   // GPUs of one box

struct Comm {
    void* nccl = nullptr;          // ncclComm_t
    int rank = 0, nranks = 1;
    // strip [j0, j1) of quad rows of the global nx x ny mesh
    int64_t j0 = 0, j1 = 0, ny_global = 0;
    int64_t cell_begin = 0, face_begin = 0;          // global 0-based ids of the first owned cell / face
    int64_t ncell_global = 0, nface_global = 0;
    int64_t nbelow = 0, nabove = 0;                    // ghost faces owned by rank-1 / rank+1
    // halo exchange of trace vectors: pack lists (local face ids) and buffers
    int32_t* d_send_dn_idx = nullptr;  int64_t n_send_dn = 0;   // faces sent to rank-1 (its ghost-above layer)
    int32_t* d_send_up_idx = nullptr;  int64_t n_send_up = 0;   // faces sent to rank+1 (its ghost-below layer)
    double *d_send_dn = nullptr, *d_send_up = nullptr;
    double* d_gscal = nullptr;       // all-reduced scalars
    // peer-memory path (CUDA IPC over NVLink): mailboxes for the flag-based all-reduce, neighbours' PCG vectors
    bool p2p = false;
    double* d_mail = nullptr;               // my mailbox [2][nranks][MAILW] (double-buffered by epoch parity)
    double** d_peer_mail = nullptr;         // device array [nranks]: the mailbox of every rank as mapped in this process
    unsigned long long* d_epoch = nullptr;  // all-reduce epoch counter
    void* peer_vec[MAXR] = {};              // mapped PCG vector region of the ranks that own ghost faces of mine
    int64_t peer_ndof[MAXR] = {};           // their local dof counts (stride of the four vectors inside the region)
    unsigned need_rank = 0;                 // bit q: some ghost face of mine is owned by rank q
    int32_t* d_ghost_ridx = nullptr;        // for every ghost face: its local face index on the owning rank
    int32_t* d_ghost_owner = nullptr;       // ... and the owning rank
    bool general_mesh = false;              // partition of an hdg_set_mesh mesh (no NCCL halo lists)
    std::vector<int64_t> ghost_cells;       // global 0-based ids of the ghost cells, in local order (hdg_get_ghost_cells)
    std::vector<void*> ipc_opened;
};
constexpr int MAILW = 8;   // doubles per mailbox slot: epoch word + up to 7 values

}  // namespace hdg

struct hdg_context {
    hdg_params prm{};
    hdg::RefTables tab;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    bool use_lu = false;          // literal quadrature + LU path (mass matrix not identity / requested)
    int ke_layout = hdg::KE_TILE32;

    // mesh (device)
    int64_t ncell = 0, nnode = 0, nface = 0, nbface = 0;   // local totals (owned + ghost)
    int64_t ncell_own = 0, nface_own = 0;                 // owned by this rank (== totals on one GPU)
    // ncell x CI int32 records: v0 v1 v2 (0-based node ids), f0 f1 f2 (0-based face ids, bit31 = this cell is
    // the face's second cell), partner word (per local face 8 bits: bit7 = neighbour cell lies in the same
    // 32-cell tile, bits0-4 its index in the tile, bits5-6 its local face index), boundary-face bits
    int32_t* d_cellinfo = nullptr;
    double* d_nodes = nullptr;       // nnode x 2
    int32_t* d_facecell = nullptr;   // nface x 2 : cell1, cell2 (0-based, -1 = none)
    int32_t* d_facenode = nullptr;   // nface x 2 : v1, v2 (0-based)
    int32_t* d_bfaces = nullptr;     // nbface sorted ascending (0-based)
    uint8_t* d_isbc = nullptr;       // nface
    int32_t* d_kcol = nullptr;       // nface x 4 neighbour faces of the block-ELL rows (-1 = none)
    bool have_mesh = false;
    int64_t cap_ncell = 0, cap_nnode = 0, cap_nface = 0, cap_nbface = 0;   // sizes the device buffers were allocated for
    int64_t *d_stage_cells = nullptr, *d_stage_faces = nullptr;           // int64 staging of hdg_set_mesh inputs
    // structured-mesh parameters (0 when the mesh came from hdg_set_mesh)
    int64_t nx = 0, ny = 0;
    // vertex grid (px x py, node id = iy px + ix) when the triangulation is that of rectangle_mesh - either generated
    // (hdg_set_rectangle_mesh) or recognised in the arrays of hdg_set_mesh; 0 = none.  Used by the multigrid preconditioner.
    int64_t grid_px = 0, grid_py = 0;

    // source
    double* d_fq = nullptr;          // ncell x nq when source_id == 0

    // raw tables on device (LU path + error norm)
    double* d_rawtab = nullptr;
    hdg::RawTablesDev raw{};
    bool quad_ok = false;            // hdg_sparsity.h matches the tables of this context

    // trace system, block-ELL: diagonal blocks + 4 off-diagonal blocks per face, blocks column-major nt x nt
    double* d_Kd = nullptr;          // nface * nt*nt
    double* d_Ko = nullptr;          // nface * 4 * nt*nt
    double* d_rhs = nullptr;         // ndof
    double* d_Ke = nullptr;          // ncell * ke  ([K_e | b_e])
    bool assembled = false, applied = false;
    double meandiag = 0.0;
    double* d_bcval = nullptr;       // nbface*nt prescribed values (NULL = 0)

    // solver vectors
    double* d_x = nullptr;           // trace solution u_hat
    double *d_r = nullptr, *d_p = nullptr, *d_Ap = nullptr, *d_dinv = nullptr;   // d_p is mapped by the neighbouring ranks (CUDA IPC)
    double* d_binv = nullptr;        // block-Jacobi: inverted face-diagonal blocks
    int precond = 0;                 // 0 Jacobi, 1 block-Jacobi, 2 block-Jacobi + P1-vertex multigrid (hdg_mg.cu)
    void* mg = nullptr;              // hdg::MgData
    void* mg_general = nullptr;      // MgGeneral: vertex term on meshes without grid structure (hdg_mgx.cu)
    void* cg = nullptr;              // CgData: CG side of the exported API (hdg_cg.cu)
    double* d_partials = nullptr;    // reduction partials
    double* d_scal = nullptr;        // device scalars
    int32_t* d_flags = nullptr;      // error / convergence flags
    int32_t* h_flags = nullptr;      // pinned mirror
    double* h_scal = nullptr;        // pinned mirror
    bool solved = false;

    // recovered values, column-major ncell x nb
    double *d_sigma = nullptr, *d_u = nullptr, *d_uhat_h = nullptr;
    bool recovered = false;

    hdg::Timer t_assemble, t_apply, t_solve, t_recover, t_err, t_elem, t_mgsetup, t_loop;
    hdg::Comm* comm = nullptr;
};

namespace hdg {

// error helpers -------------------------------------------------------------------------------
hdg_status set_err(hdg_context* c, hdg_status s, const std::string& msg);
#define HDG_CUDA(c, call)                                                                        \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return hdg::set_err((c), HDG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

void timer_start(hdg_context* c, Timer& t);
void timer_stop(hdg_context* c, Timer& t);
float timer_ms(Timer& t);

// flag words in d_flags
enum Flag : int { FLAG_BAD_GEOM = 0, FLAG_SINGULAR = 1, FLAG_DONE = 2, FLAG_ITERS = 3, FLAG_NOT_BOUNDARY = 4, FLAG_MG = 5, FLAG_GRID_PX = 6, FLAG_GRID_BAD = 7, FLAG_BAD_ID = 8, NFLAGS = 12 };

// ---- per-translation-unit entry points ------------------------------------------------------
hdg_status upload_tables(hdg_context* c);                       // hdg_element.cu
void release_tables(const hdg_context* c);                      // hdg_element.cu: gives up the __constant__ tables on destroy
hdg_status launch_element_kernels(hdg_context* c);              // hdg_element.cu
hdg_status condensed_of_cell(hdg_context* c, int64_t cell, double* At, double* bt);  // hdg_element.cu

hdg_status mesh_from_host(hdg_context* c, const int64_t* cells, int64_t ncell, const double* nodes,
                          int64_t nnode, const int64_t* faces, int64_t nface, const int64_t* bfaces,
                          int64_t nbface);                      // hdg_mesh.cu
hdg_status mesh_set_dirichlet(hdg_context* c, const int64_t* bfaces, int64_t nbface);   // hdg_mesh.cu
hdg_status mesh_rectangle(hdg_context* c, int64_t nx, int64_t ny, double llx, double lly, double urx,
                          double ury);                          // hdg_mesh.cu
hdg_status mesh_perturb(hdg_context* c, double fraction, uint64_t seed);
hdg_status mesh_download(hdg_context* c, int64_t* cells, double* nodes, int64_t* faces, int64_t* bf);
hdg_status pattern_download(hdg_context* c, int64_t* colptr, int64_t* rowval);
hdg_status values_download(hdg_context* c, double* nzval);
int64_t pattern_nnz(hdg_context* c);
hdg_status order_cells_host(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode, int64_t* perm);   // hdg_order.cu
hdg_status number_faces_host(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode,
                             int64_t* cells_out, int64_t* faces_out, int64_t faces_capacity, int64_t* nface_out);
hdg_status alloc_system(hdg_context* c);                        // hdg_mesh.cu (after mesh known)
void free_mesh(hdg_context* c);

hdg_status apply_dirichlet(hdg_context* c, const double* values);   // hdg_solve.cu
hdg_status pcg_solve(hdg_context* c, double rtol, int maxit, hdg_solve_info* info);

// hdg_mg.cu: P1-vertex multigrid term of the preconditioner (one GPU, rectangle_mesh)
hdg_status mg_setup(hdg_context* c);                                              // operators of the current trace matrix
hdg_status mg_apply(hdg_context* c, const double* r, double* z, double* part, int np, bool fuse = false);   // z += P V(P'r); part = partials of (P'r).V(P'r)
// fuse: the V-cycle leaves its result t on the vertices and the caller's direction update adds P t itself (pcg_dir_mg); the
// pointers it needs, false when the general-mesh term (hdg_mgx.cu) is active - that one always updates z itself
bool mg_fused_ptrs(const hdg_context* c, const int32_t** facenode, const double** t0);
constexpr double MG_C1 = 0.28867513459481287;   // 1 / (2 sqrt 3): trace mode 1 of a P1 function with vertex values (a, b) is C1 (b - a)
void mg_free(hdg_context* c);
void mg_invalidate(hdg_context* c);                             // keep the buffers, rebuild adjacency / flags at the next solve
int mg_levels(const hdg_context* c);
int mg_trace(hdg_context* c, double* us, int cap);                // HDG_MG_TRACE=1: barrier timestamps of the last V-cycle
int mg_launches_per_apply(const hdg_context* c);

hdg_status recover(hdg_context* c);                             // hdg_recover.cu
hdg_status errornorm(hdg_context* c, int exact_id, const double* uex_host, double* err2);
hdg_status local_download(hdg_context* c, int64_t cell, double* Ke, double* be);
hdg_status nodal_average(hdg_context* c, double* out);

// hdg_cg.cu: CG side (examples/poisson2D_CG.jl) on the mesh of the context
hdg_status cg_setup(hdg_context* c, int order, int64_t* ndofs_out);
hdg_status cg_sizes(hdg_context* c, int64_t out[4]);
hdg_status cg_download(hdg_context* c, int64_t* cell_dofs, int64_t* colptr, int64_t* rowval, double* nzval, double* rhs, double* u);
hdg_status cg_assemble(hdg_context* c);
hdg_status cg_apply_dirichlet(hdg_context* c);
hdg_status cg_solve(hdg_context* c, double rtol, int maxit, hdg_solve_info* info);
hdg_status cg_errornorm(hdg_context* c, double* err2);
double cg_meandiag(const hdg_context* c);
void cg_release(hdg_context* c);
hdg_status exclusive_scan_i32(hdg_context* c, const int32_t* in, int64_t n, int64_t* out, int64_t* d_total);   // hdg_mesh.cu

// hdg_comm.cu
bool comm_active(const hdg_context* c);
hdg_status comm_allreduce_sum(hdg_context* c, double* d_buf, int count);                 // in place, on c->stream
hdg_status comm_halo_exchange(hdg_context* c, double* d_vec, int nt);                    // fills the ghost segments of d_vec
hdg_status comm_setup_halo(hdg_context* c, const std::vector<int32_t>& send_dn, const std::vector<int32_t>& send_up);
void comm_free_halo(hdg_context* c);
void comm_destroy(hdg_context* c);
bool comm_p2p(const hdg_context* c);
hdg_status comm_share_vectors(hdg_context* c, void* region, int64_t ndof_own);              // maps the neighbours' regions
void comm_unshare_vectors(hdg_context* c);
// generic: all-gathers the IPC handle of `mine` (a cudaMalloc base pointer) and maps the regions of the ranks in `rank_mask`
// into peers[q] (peers[rank] = mine); collective.  comm_close_buffer unmaps them.
hdg_status comm_share_buffer(hdg_context* c, void* mine, unsigned rank_mask, void* peers[MAXR]);
void comm_close_buffer(hdg_context* c, void* peers[MAXR]);
struct XgComm;
void comm_xg(const hdg_context* c, XgComm* out);                 // mailbox handles for in-kernel barriers (hdg_xgpu.cuh)
hdg_status comm_p2p_allreduce(hdg_context* c, const double* d_partials, int np, unsigned slot_mask);   // selected partial arrays -> d_gscal, all ranks (mask 0: barrier only)
hdg_status comm_set_ghosts(hdg_context* c, const std::vector<int32_t>& ridx, const std::vector<int32_t>& owner);

}  // namespace hdg
