// C ABI of libhdg_b200 (include/hdg_b200.h): argument checking, call-order state machine, error
// reporting.  No compute lives here and there is no CPU fallback: every compute entry point ends
// in a CUDA kernel launch or fails with HDG_ERR_CUDA.
#include <cmath>
#include <cstring>

#include "hdg_internal.h"

namespace hdg {

static thread_local std::string g_create_error;

hdg_status set_err(hdg_context* c, hdg_status s, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return s;
}

void timer_start(hdg_context* c, Timer& t) {
    if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    cudaEventRecord(t.a, c->stream);
}
void timer_stop(hdg_context* c, Timer& t) {
    cudaEventRecord(t.b, c->stream);
    t.pending = true;
}
float timer_ms(Timer& t) {
    if (t.pending) {
        cudaEventSynchronize(t.b);
        cudaEventElapsedTime(&t.last_ms, t.a, t.b);
        t.pending = false;
    }
    return t.last_ms;
}

static hdg_status check_flags(hdg_context* c) {
    HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->h_flags[FLAG_BAD_GEOM])
        return set_err(c, HDG_ERR_BAD_GEOMETRY, "det(J) is not positive in cell " + std::to_string(c->h_flags[FLAG_BAD_GEOM]));
    if (c->h_flags[FLAG_SINGULAR])
        return set_err(c, HDG_ERR_SINGULAR_LOCAL, "singular local matrix in cell " + std::to_string(c->h_flags[FLAG_SINGULAR]));
    return HDG_OK;
}

}  // namespace hdg

using namespace hdg;

// ---- measurement helper: FP64 FMA peak of this device (the denominator SURVEY 8(d) asks for) -------------------
// Every thread runs 8 independent DFMA chains; 148 SMs x 8 blocks x 256 threads keep the FP64 pipe saturated.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b, int iters) {
    double x0 = threadIdx.x, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;   // never true in practice; keeps the chains alive
}


extern "C" {

const char* hdg_version(void) { return "hdg_b200 0.2 (sm_100a)"; }

const char* hdg_last_error(const hdg_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

hdg_status hdg_create(const hdg_params* prm, hdg_context** out) {
    if (!prm || !out) return set_err(nullptr, HDG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (prm->source_id < 0 || prm->source_id > 1) return set_err(nullptr, HDG_ERR_INVALID, "unknown source_id");
    if (prm->local_solver < 0 || prm->local_solver > 1) return set_err(nullptr, HDG_ERR_INVALID, "unknown local_solver");
    RefTables tab;
    try {
        tab = build_ref_tables(prm->order, prm->quad_degree);
    } catch (const std::string& e) {
        bool rule = e.find("not available") != std::string::npos;
        return set_err(nullptr, rule ? HDG_ERR_UNSUPPORTED_RULE : HDG_ERR_INVALID, e);
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_err(nullptr, HDG_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
    hdg_context* c = new hdg_context();
    c->prm = *prm;
    c->tab = tab;
    if (prm->device >= 0) {
        if (cudaSetDevice(prm->device) != cudaSuccess) {
            delete c;
            return set_err(nullptr, HDG_ERR_CUDA, "cudaSetDevice failed");
        }
    }
    cudaGetDevice(&c->device);
    auto fail = [&](const char* what) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
        hdg_destroy(c);
        return set_err(nullptr, HDG_ERR_CUDA, msg);
    };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate");
    if (cudaMalloc(&c->d_flags, sizeof(int32_t) * NFLAGS) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&c->d_scal, sizeof(double) * 8) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&c->d_partials, sizeof(double) * 8 * 2048) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMallocHost(&c->h_flags, sizeof(int32_t) * NFLAGS) != cudaSuccess) return fail("cudaMallocHost");
    if (cudaMallocHost(&c->h_scal, sizeof(double) * 8) != cudaSuccess) return fail("cudaMallocHost");
    cudaMemset(c->d_flags, 0, sizeof(int32_t) * NFLAGS);
    // Block elimination needs the mass matrix of the chosen cell rule to be invertible; it is
    // the identity exactly when the rule integrates degree 2k (Dubiner is orthonormal).  Rules that
    // under-integrate it (e.g. the reference default quad_degree = k+1 for k = 2) go through the
    // literal quadrature + LU path, as the reference does.
    double dev = 0.0;
    for (int i = 0; i < tab.n; ++i)
        for (int j = 0; j < tab.n; ++j) dev = std::fmax(dev, std::fabs(tab.Mhat[size_t(i) * tab.n + j] - (i == j ? 1.0 : 0.0)));
    c->use_lu = prm->local_solver == 1 || !(dev < 1e-9);
    hdg_status st = upload_tables(c);
    if (st) {
        std::string msg = c->err;
        hdg_destroy(c);
        return set_err(nullptr, st, msg);
    }
    *out = c;
    return HDG_OK;
}

void hdg_destroy(hdg_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_mesh(c);
    release_tables(c);
    comm_destroy(c);
    if (c->d_rawtab) cudaFree(c->d_rawtab);
    if (c->d_flags) cudaFree(c->d_flags);
    if (c->d_scal) cudaFree(c->d_scal);
    mg_free(c);
    if (c->d_partials) cudaFree(c->d_partials);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    for (Timer* t : {&c->t_assemble, &c->t_apply, &c->t_solve, &c->t_recover, &c->t_err, &c->t_elem, &c->t_mgsetup, &c->t_loop}) {
        if (t->a) cudaEventDestroy(t->a);
        if (t->b) cudaEventDestroy(t->b);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

hdg_status hdg_set_mesh(hdg_context* c, const int64_t* cells, int64_t ncell, const double* nodes, int64_t nnode,
                        const int64_t* faces, int64_t nface, const int64_t* bfaces, int64_t nbface) {
    if (!c) return HDG_ERR_INVALID;
    if (!cells || !nodes || (nbface > 0 && !bfaces)) return set_err(c, HDG_ERR_INVALID, "null mesh array");
    cudaSetDevice(c->device);
    return mesh_from_host(c, cells, ncell, nodes, nnode, faces, nface, bfaces, nbface);
}

hdg_status hdg_set_rectangle_mesh(hdg_context* c, int64_t nx, int64_t ny, double llx, double lly, double urx, double ury) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return mesh_rectangle(c, nx, ny, llx, lly, urx, ury);
}

hdg_status hdg_set_dirichlet_faces(hdg_context* c, const int64_t* bfaces, int64_t nbface) {
    if (!c || nbface < 0 || (nbface > 0 && !bfaces)) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return mesh_set_dirichlet(c, bfaces, nbface);
}

hdg_status hdg_number_faces(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode,
                            int64_t* cells_out, int64_t* faces_out, int64_t faces_capacity, int64_t* nface_out) {
    if (!c || !tri || !nodes || !nface_out) return HDG_ERR_INVALID;
    if (ncell < 1 || nnode < 3) return set_err(c, HDG_ERR_INVALID, "empty mesh");
    cudaSetDevice(c->device);
    return number_faces_host(c, tri, ncell, nodes, nnode, cells_out, faces_out, faces_capacity, nface_out);
}

hdg_status hdg_order_cells(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode, int64_t* perm_out) {
    if (!c || !tri || !nodes || !perm_out) return HDG_ERR_INVALID;
    if (ncell < 1 || nnode < 3) return set_err(c, HDG_ERR_INVALID, "empty mesh");
    cudaSetDevice(c->device);
    return order_cells_host(c, tri, ncell, nodes, nnode, perm_out);
}

hdg_status hdg_perturb_nodes(hdg_context* c, double fraction, uint64_t seed) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return mesh_perturb(c, fraction, seed);
}

hdg_status hdg_get_sizes(const hdg_context* c, hdg_sizes* s) {
    if (!c || !s) return HDG_ERR_INVALID;
    std::memset(s, 0, sizeof(*s));
    s->n = c->tab.n; s->nt = c->tab.nt; s->m = c->tab.m; s->t = c->tab.t; s->nq = c->tab.nq; s->nfq = c->tab.nfq;
    if (c->have_mesh) {
        // multi-GPU: the counts of what this rank owns (hdg_get_partition gives the global ranges)
        s->ncell = c->ncell_own; s->nnode = c->nnode; s->nface = c->nface_own; s->nbface = c->nbface;
        s->ndof = c->nface_own * c->tab.nt;
        s->nnz = comm_active(c) ? 0 : pattern_nnz(const_cast<hdg_context*>(c));
    }
    return HDG_OK;
}

hdg_status hdg_get_mesh(hdg_context* c, int64_t* cells, double* nodes, int64_t* faces, int64_t* bf) {
    if (!c) return HDG_ERR_INVALID;
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_get_mesh is single-GPU (a rank holds a strip in local numbering)");
    cudaSetDevice(c->device);
    return mesh_download(c, cells, nodes, faces, bf);
}

static const std::vector<double>* table_by_name(const RefTables& T, const std::string& s) {
    if (s == "qpoints") return &T.qpts;
    if (s == "qweights") return &T.qw;
    if (s == "fpoints") return &T.fpts;
    if (s == "fweights") return &T.fw;
    if (s == "N") return &T.N;
    if (s == "dNdxi") return &T.dN;
    if (s == "E") return &T.E;
    if (s == "T") return &T.T;
    if (s == "Mhat") return &T.Mhat;
    // reference matrices consumed by the device kernels (hdg_tables.h)
    if (s == "Tr") return &T.Tr;
    if (s == "Ts") return &T.Ts;
    if (s == "Prr") return &T.Prr;
    if (s == "Prs") return &T.Prs;
    if (s == "Pss") return &T.Pss;
    if (s == "Chat") return &T.Chat;
    if (s == "Fhat") return &T.Fhat;
    if (s == "MF") return &T.MF;
    if (s == "Qr") return &T.Qr;
    if (s == "Qs") return &T.Qs;
    if (s == "Hhat") return &T.Hhat;
    return nullptr;
}

hdg_status hdg_get_table(const hdg_context* c, const char* name, double* buf, int64_t* count) {
    if (!c || !name || !count) return HDG_ERR_INVALID;
    const std::vector<double>* v = table_by_name(c->tab, name);
    if (!v) return set_err(const_cast<hdg_context*>(c), HDG_ERR_INVALID, "unknown table name");
    *count = int64_t(v->size());
    if (buf) std::memcpy(buf, v->data(), sizeof(double) * v->size());
    return HDG_OK;
}

hdg_status hdg_ref_table(int32_t order, int32_t quad_degree, const char* name, double* buf, int64_t* count) {
    if (!name || !count) return HDG_ERR_INVALID;
    RefTables tab;
    try {
        tab = build_ref_tables(order, quad_degree);
    } catch (const std::string& e) {
        bool rule = e.find("not available") != std::string::npos;
        return set_err(nullptr, rule ? HDG_ERR_UNSUPPORTED_RULE : HDG_ERR_INVALID, e);
    }
    const std::vector<double>* v = table_by_name(tab, name);
    if (!v) return set_err(nullptr, HDG_ERR_INVALID, "unknown table name");
    *count = int64_t(v->size());
    if (buf) std::memcpy(buf, v->data(), sizeof(double) * v->size());
    return HDG_OK;
}

hdg_status hdg_basis_value(int32_t kind, int32_t j, const double* xi, double* value, double* grad) {
    if (!xi || !value) return HDG_ERR_INVALID;
    if (kind == 0) {            // Dubiner on the reference triangle, src/basis.jl:65-86
        if (j < 1 || j > 15 || !(xi[0] >= -1e-12 && xi[1] >= -1e-12)) return set_err(nullptr, HDG_ERR_INVALID, "Dubiner: 1 <= j <= 15 (order <= 4)");
        double v, dr, ds;
        dubiner_eval(j, xi[0], xi[1], &v, &dr, &ds);
        *value = v;
        if (grad) { grad[0] = dr; grad[1] = ds; }
        return HDG_OK;
    }
    if (kind == 1) {            // orthonormal Legendre on (0,1), src/basis.jl:351-354
        if (j < 1 || j > 5) return set_err(nullptr, HDG_ERR_INVALID, "Legendre: 1 <= j <= 5 (order <= 4)");
        *value = legendre01_eval(j, xi[0]);
        if (grad) {             // central difference of the polynomial (table building never needs the derivative)
            const double h = 1e-6;
            grad[0] = (legendre01_eval(j, xi[0] + h) - legendre01_eval(j, xi[0] - h)) / (2.0 * h);
        }
        return HDG_OK;
    }
    return set_err(nullptr, HDG_ERR_INVALID, "unknown basis kind");
}

hdg_status hdg_set_source_values(hdg_context* c, const double* fq) {
    if (!c || !fq) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "hdg_set_source_values before a mesh is set");
    cudaSetDevice(c->device);
    size_t bytes = sizeof(double) * c->ncell * c->tab.nq;
    if (!c->d_fq) HDG_CUDA(c, cudaMalloc(&c->d_fq, bytes));
    HDG_CUDA(c, cudaMemcpyAsync(c->d_fq, fq, bytes, cudaMemcpyHostToDevice, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status hdg_assemble_async(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "hdg_assemble before a mesh is set");
    if (c->prm.source_id == 0 && !c->d_fq) return set_err(c, HDG_ERR_INVALID, "source_id 0 needs hdg_set_source_values");
    cudaSetDevice(c->device);
    timer_start(c, c->t_assemble);
    hdg_status st = launch_element_kernels(c);
    timer_stop(c, c->t_assemble);
    if (st) return st;
    c->assembled = true;
    c->applied = c->solved = c->recovered = false;
    return HDG_OK;
}

hdg_status hdg_sync(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

uint64_t hdg_stream(const hdg_context* c) { return c ? uint64_t(reinterpret_cast<uintptr_t>(c->stream)) : 0; }

hdg_status hdg_assemble(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
    hdg_status st = hdg_assemble_async(c);
    if (st) return st;
    st = check_flags(c);
    if (st) c->assembled = false;
    return st;
}

hdg_status hdg_apply_dirichlet(hdg_context* c, const double* values) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_apply_dirichlet before hdg_assemble");
    if (c->applied) return set_err(c, HDG_ERR_INVALID, "hdg_apply_dirichlet called twice on the same assembly");
    cudaSetDevice(c->device);
    timer_start(c, c->t_apply);
    hdg_status st = apply_dirichlet(c, values);
    timer_stop(c, c->t_apply);
    if (st) return st;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->applied = true;
    return HDG_OK;
}

hdg_status hdg_solve(hdg_context* c, double rtol, int32_t maxit, hdg_solve_info* info) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_solve before hdg_assemble");
    if (!c->applied) return set_err(c, HDG_ERR_INVALID, "hdg_solve before hdg_apply_dirichlet (the raw condensed matrix is singular)");
    if (!(rtol > 0.0) || maxit < 1) return set_err(c, HDG_ERR_INVALID, "need rtol > 0 and maxit >= 1");
    cudaSetDevice(c->device);
    return pcg_solve(c, rtol, maxit, info);
}

hdg_status hdg_set_preconditioner(hdg_context* c, int32_t id) {
    if (!c) return HDG_ERR_INVALID;
    if (id < 0 || id > 2) return set_err(c, HDG_ERR_INVALID, "unknown preconditioner id");
    c->precond = id;
    return HDG_OK;
}

hdg_status hdg_recover(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->assembled || !c->d_x) return set_err(c, HDG_ERR_INVALID, "hdg_recover needs hdg_assemble and a trace solution (hdg_solve / hdg_set_trace)");
    cudaSetDevice(c->device);
    return recover(c);
}

hdg_status hdg_errornorm(hdg_context* c, int32_t exact_id, double* err2) {
    if (!c || !err2) return HDG_ERR_INVALID;
    if (!c->recovered) return set_err(c, HDG_ERR_INVALID, "hdg_errornorm before hdg_recover");
    if (exact_id != 1) return set_err(c, HDG_ERR_INVALID, "unknown exact_id");
    cudaSetDevice(c->device);
    return errornorm(c, exact_id, nullptr, err2);
}

hdg_status hdg_errornorm_values(hdg_context* c, const double* uex_q, double* err2) {
    if (!c || !uex_q || !err2) return HDG_ERR_INVALID;
    if (!c->recovered) return set_err(c, HDG_ERR_INVALID, "hdg_errornorm_values before hdg_recover");
    cudaSetDevice(c->device);
    return errornorm(c, 0, uex_q, err2);
}

hdg_status hdg_get_pattern(hdg_context* c, int64_t* colptr, int64_t* rowval) {
    if (!c) return HDG_ERR_INVALID;
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_get_pattern is single-GPU");
    cudaSetDevice(c->device);
    return pattern_download(c, colptr, rowval);
}

hdg_status hdg_get_values(hdg_context* c, double* nzval) {
    if (!c || !nzval) return HDG_ERR_INVALID;
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_get_values is single-GPU");
    cudaSetDevice(c->device);
    return values_download(c, nzval);
}

hdg_status hdg_get_rhs(hdg_context* c, double* rhs) {
    if (!c || !rhs) return HDG_ERR_INVALID;
    if (!c->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_get_rhs before hdg_assemble");
    HDG_CUDA(c, cudaMemcpyAsync(rhs, c->d_rhs, sizeof(double) * c->nface_own * c->tab.nt, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status hdg_get_trace(hdg_context* c, double* uhat) {
    if (!c || !uhat) return HDG_ERR_INVALID;
    if (!c->d_x) return set_err(c, HDG_ERR_INVALID, "no trace solution yet");
    HDG_CUDA(c, cudaMemcpyAsync(uhat, c->d_x, sizeof(double) * c->nface_own * c->tab.nt, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status hdg_set_trace(hdg_context* c, const double* uhat) {
    if (!c || !uhat) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "no mesh");
    cudaSetDevice(c->device);
    if (comm_active(c) && c->comm->general_mesh)      // ghost entries live on arbitrary ranks there: only hdg_solve fills them
        return set_err(c, HDG_ERR_INVALID, "hdg_set_trace is not available on a partitioned hdg_set_mesh mesh");
    if (!c->d_x) {
        HDG_CUDA(c, cudaMalloc(&c->d_x, sizeof(double) * c->nface * c->tab.nt));
        HDG_CUDA(c, cudaMemsetAsync(c->d_x, 0, sizeof(double) * c->nface * c->tab.nt, c->stream));
    }
    HDG_CUDA(c, cudaMemcpyAsync(c->d_x, uhat, sizeof(double) * c->nface_own * c->tab.nt, cudaMemcpyHostToDevice, c->stream));
    hdg_status st = comm_halo_exchange(c, c->d_x, c->tab.nt);
    if (st) return st;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status hdg_get_meandiag(const hdg_context* c, double* m) {
    if (!c || !m) return HDG_ERR_INVALID;
    if (!c->applied) return set_err(const_cast<hdg_context*>(c), HDG_ERR_INVALID, "meandiag is computed by hdg_apply_dirichlet");
    *m = c->meandiag;
    return HDG_OK;
}

hdg_status hdg_get_local(hdg_context* c, int64_t cell, double* Ke, double* be) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_get_local before hdg_assemble");
    if (cell < 1 || cell > c->ncell_own) return set_err(c, HDG_ERR_INVALID, "cell out of range");
    cudaSetDevice(c->device);
    return local_download(c, cell - 1, Ke, be);
}

hdg_status hdg_get_condensed(hdg_context* c, int64_t cell, double* Ate, double* bte) {
    if (!c || !Ate || !bte) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "no mesh");
    if (cell < 1 || cell > c->ncell) return set_err(c, HDG_ERR_INVALID, "cell out of range");
    if (c->prm.source_id == 0 && !c->d_fq) return set_err(c, HDG_ERR_INVALID, "source_id 0 needs hdg_set_source_values");
    cudaSetDevice(c->device);
    return condensed_of_cell(c, cell - 1, Ate, bte);
}

hdg_status hdg_get_mvalues(hdg_context* c, double* sigma, double* u, double* uhat_h) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->recovered) return set_err(c, HDG_ERR_INVALID, "hdg_get_mvalues before hdg_recover");
    const int n = c->tab.n, nt = c->tab.nt;
    if (sigma) HDG_CUDA(c, cudaMemcpyAsync(sigma, c->d_sigma, sizeof(double) * c->ncell_own * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    if (u) HDG_CUDA(c, cudaMemcpyAsync(u, c->d_u, sizeof(double) * c->ncell_own * n, cudaMemcpyDeviceToHost, c->stream));
    if (uhat_h) HDG_CUDA(c, cudaMemcpyAsync(uhat_h, c->d_uhat_h, sizeof(double) * c->ncell_own * nt * 3, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    return HDG_OK;
}

hdg_status hdg_nodal_avg(hdg_context* c, double* out) {
    if (!c || !out) return HDG_ERR_INVALID;
    if (!c->recovered) return set_err(c, HDG_ERR_INVALID, "hdg_nodal_avg before hdg_recover");
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_nodal_avg is single-GPU");
    cudaSetDevice(c->device);
    return nodal_average(c, out);
}

hdg_status hdg_last_phase_ms(const hdg_context* cc, const char* phase, double* ms) {
    if (!cc || !phase || !ms) return HDG_ERR_INVALID;
    hdg_context* c = const_cast<hdg_context*>(cc);
    std::string s(phase);
    Timer* t = nullptr;
    if (s == "assemble") t = &c->t_assemble;
    else if (s == "apply") t = &c->t_apply;
    else if (s == "solve") t = &c->t_solve;
    else if (s == "recover") t = &c->t_recover;
    else if (s == "errornorm") t = &c->t_err;
    else if (s == "element_kernel") t = &c->t_elem;
    else if (s == "mg_setup") t = &c->t_mgsetup;       // operators of the vertex hierarchy (inside "solve")
    else if (s == "solve_loop") t = &c->t_loop;        // the PCG iterations alone (inside "solve")
    else return set_err(c, HDG_ERR_INVALID, "unknown phase");
    if (!t->a) { *ms = 0.0; return HDG_OK; }
    *ms = double(timer_ms(*t));
    return HDG_OK;
}

hdg_status hdg_measure_fp64_peak(hdg_context* c, double* tflops) {
    if (!c || !tflops) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    int sms = 0;
    HDG_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    double* d = nullptr;
    HDG_CUDA(c, cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    HDG_CUDA(c, cudaEventCreate(&e0));
    HDG_CUDA(c, cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {   // first launch warms up
        cudaEventRecord(e0, c->stream);
        dfma_peak_kernel<<<blocks, threads, 0, c->stream>>>(d, 0.999999, 1e-9, iters);
        cudaEventRecord(e1, c->stream);
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { cudaFree(d); return set_err(c, HDG_ERR_CUDA, cudaGetErrorString(err)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * double(iters) * double(blocks) * double(threads) / (double(ms) * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        c->launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return HDG_OK;
}

// ---- CG side of the exported API (examples/poisson2D_CG.jl), hdg_cg.cu ----------------------------------------------------
hdg_status hdg_cg_setup(hdg_context* c, int32_t order, int64_t* ndofs) {
    if (!c) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "hdg_cg_setup before a mesh is set");
    cudaSetDevice(c->device);
    return cg_setup(c, order, ndofs);
}
hdg_status hdg_cg_get_sizes(hdg_context* c, int64_t out[4]) {
    if (!c || !out) return HDG_ERR_INVALID;
    return cg_sizes(c, out);
}
hdg_status hdg_cg_get_dofhandler(hdg_context* c, int64_t* cell_dofs, int64_t* colptr, int64_t* rowval) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_download(c, cell_dofs, colptr, rowval, nullptr, nullptr, nullptr);
}
hdg_status hdg_cg_assemble(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_assemble(c);
}
hdg_status hdg_cg_apply_dirichlet(hdg_context* c) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_apply_dirichlet(c);
}
hdg_status hdg_cg_solve(hdg_context* c, double rtol, int32_t maxit, hdg_solve_info* info) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_solve(c, rtol, maxit, info);
}
hdg_status hdg_cg_get_system(hdg_context* c, double* nzval, double* rhs, double* u) {
    if (!c) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_download(c, nullptr, nullptr, nullptr, nzval, rhs, u);
}
hdg_status hdg_cg_errornorm(hdg_context* c, double* err2) {
    if (!c || !err2) return HDG_ERR_INVALID;
    cudaSetDevice(c->device);
    return cg_errornorm(c, err2);
}
hdg_status hdg_cg_get_meandiag(const hdg_context* c, double* m) {
    if (!c || !m) return HDG_ERR_INVALID;
    *m = cg_meandiag(c);
    return HDG_OK;
}

int32_t hdg_mg_trace(hdg_context* c, double* usec, int32_t capacity) {
    if (!c || !usec) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    return mg_trace(c, usec, capacity);
}

int64_t hdg_launch_count(const hdg_context* c) { return c ? c->launches : 0; }

}  // extern "C"
