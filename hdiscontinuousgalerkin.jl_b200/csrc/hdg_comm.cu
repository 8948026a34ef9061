// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process; this is the sharding of the same hot path:
// assembly / condensation / recovery are embarrassingly parallel over strips of cells (a one-cell
// ghost layer is recomputed instead of exchanged), the PCG needs one halo exchange of interface
// trace values per SpMV and all-reduced dot products.  NCCL is loaded with dlopen so that the
// single-GPU library has no NCCL dependency; the caller distributes the ncclUniqueId
// (torch.distributed / MPI / Julia Distributed are all fine - plain bytes through the C ABI).
#include <dlfcn.h>

#include <cstring>

#include "hdg_internal.h"

namespace hdg {

// minimal NCCL surface (nccl.h 2.x ABI)
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
typedef int nccl_result_t;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void* handle = nullptr;
    nccl_result_t (*GetUniqueId)(nccl_uid_t*) = nullptr;
    nccl_result_t (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(nccl_result_t) = nullptr;
};
static NcclApi g_nccl;

static bool load_nccl(std::string& why) {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        why = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_nccl.handle = h;
    return true;
}

#define HDG_NCCL(c, call)                                                                              \
    do {                                                                                               \
        nccl_result_t r_ = (call);                                                                     \
        if (r_ != 0)                                                                                   \
            return set_err((c), HDG_ERR_NCCL, std::string(#call) + ": " +                               \
                                                  (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
    } while (0)

bool comm_active(const hdg_context* c) { return c->comm && c->comm->nranks > 1; }

hdg_status comm_allreduce_sum(hdg_context* c, double* d_buf, int count) {
    if (!comm_active(c)) return HDG_OK;
    HDG_NCCL(c, g_nccl.AllReduce(d_buf, d_buf, size_t(count), NCCL_FLOAT64, NCCL_SUM, c->comm->nccl, c->stream));
    return HDG_OK;
}

__global__ void pack_faces(const double* __restrict__ v, const int32_t* __restrict__ idx, int64_t n, int nt,
                           double* __restrict__ out) {
    int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n * nt) return;
    int64_t f = k / nt;
    int a = int(k - f * nt);
    out[k] = v[int64_t(idx[f]) * nt + a];
}

// Vector layout on every rank: [owned faces | ghost-below faces | ghost-above faces] x nt.
hdg_status comm_halo_exchange(hdg_context* c, double* d_vec, int nt) {
    if (!comm_active(c)) return HDG_OK;
    Comm* m = c->comm;
    if (m->n_send_dn) {
        pack_faces<<<(unsigned)ceil_div(m->n_send_dn * nt, 256), 256, 0, c->stream>>>(d_vec, m->d_send_dn_idx, m->n_send_dn, nt, m->d_send_dn);
        c->launches += 1;
    }
    if (m->n_send_up) {
        pack_faces<<<(unsigned)ceil_div(m->n_send_up * nt, 256), 256, 0, c->stream>>>(d_vec, m->d_send_up_idx, m->n_send_up, nt, m->d_send_up);
        c->launches += 1;
    }
    double* ghost_below = d_vec + c->nface_own * nt;
    double* ghost_above = ghost_below + m->nbelow * nt;
    HDG_NCCL(c, g_nccl.GroupStart());
    if (m->n_send_dn) HDG_NCCL(c, g_nccl.Send(m->d_send_dn, size_t(m->n_send_dn * nt), NCCL_FLOAT64, m->rank - 1, m->nccl, c->stream));
    if (m->n_send_up) HDG_NCCL(c, g_nccl.Send(m->d_send_up, size_t(m->n_send_up * nt), NCCL_FLOAT64, m->rank + 1, m->nccl, c->stream));
    if (m->nbelow) HDG_NCCL(c, g_nccl.Recv(ghost_below, size_t(m->nbelow * nt), NCCL_FLOAT64, m->rank - 1, m->nccl, c->stream));
    if (m->nabove) HDG_NCCL(c, g_nccl.Recv(ghost_above, size_t(m->nabove * nt), NCCL_FLOAT64, m->rank + 1, m->nccl, c->stream));
    HDG_NCCL(c, g_nccl.GroupEnd());
    return HDG_OK;
}

void comm_free_halo(hdg_context* c) {
    if (!c->comm) return;
    Comm* m = c->comm;
    auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    F(m->d_send_dn_idx); F(m->d_send_up_idx); F(m->d_send_dn); F(m->d_send_up);
    m->n_send_dn = m->n_send_up = 0;
}

hdg_status comm_setup_halo(hdg_context* c, const std::vector<int32_t>& send_dn, const std::vector<int32_t>& send_up) {
    Comm* m = c->comm;
    comm_free_halo(c);
    const int nt = c->tab.nt;
    m->n_send_dn = int64_t(send_dn.size());
    m->n_send_up = int64_t(send_up.size());
    if (m->n_send_dn) {
        HDG_CUDA(c, cudaMalloc(&m->d_send_dn_idx, sizeof(int32_t) * m->n_send_dn));
        HDG_CUDA(c, cudaMalloc(&m->d_send_dn, sizeof(double) * m->n_send_dn * nt));
        HDG_CUDA(c, cudaMemcpy(m->d_send_dn_idx, send_dn.data(), sizeof(int32_t) * m->n_send_dn, cudaMemcpyHostToDevice));
    }
    if (m->n_send_up) {
        HDG_CUDA(c, cudaMalloc(&m->d_send_up_idx, sizeof(int32_t) * m->n_send_up));
        HDG_CUDA(c, cudaMalloc(&m->d_send_up, sizeof(double) * m->n_send_up * nt));
        HDG_CUDA(c, cudaMemcpy(m->d_send_up_idx, send_up.data(), sizeof(int32_t) * m->n_send_up, cudaMemcpyHostToDevice));
    }
    return HDG_OK;
}

void comm_destroy(hdg_context* c) {
    if (!c->comm) return;
    comm_free_halo(c);
    if (c->comm->d_gscal) cudaFree(c->comm->d_gscal);
    if (c->comm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->nccl);
    delete c->comm;
    c->comm = nullptr;
}

}  // namespace hdg

using namespace hdg;

extern "C" {

hdg_status hdg_comm_unique_id(uint8_t id_out[128]) {
    if (!id_out) return HDG_ERR_INVALID;
    std::string why;
    if (!load_nccl(why)) return set_err(nullptr, HDG_ERR_NCCL, why);
    nccl_uid_t id;
    nccl_result_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return set_err(nullptr, HDG_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(id_out, id.internal, 128);
    return HDG_OK;
}

hdg_status hdg_comm_init(hdg_context* c, int32_t rank, int32_t nranks, const uint8_t id[128]) {
    if (!c || !id) return HDG_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return set_err(c, HDG_ERR_INVALID, "bad rank / nranks");
    if (c->have_mesh) return set_err(c, HDG_ERR_INVALID, "hdg_comm_init must precede the mesh");
    if (c->comm) return set_err(c, HDG_ERR_INVALID, "communicator already initialised");
    std::string why;
    if (!load_nccl(why)) return set_err(c, HDG_ERR_NCCL, why);
    cudaSetDevice(c->device);
    Comm* m = new Comm();
    m->rank = rank;
    m->nranks = nranks;
    nccl_uid_t uid;
    std::memcpy(uid.internal, id, 128);
    nccl_result_t r = g_nccl.CommInitRank(&m->nccl, nranks, uid, rank);
    if (r != 0) {
        delete m;
        return set_err(c, HDG_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    }
    if (cudaMalloc(&m->d_gscal, sizeof(double) * 8) != cudaSuccess) {
        delete m;
        return set_err(c, HDG_ERR_CUDA, "cudaMalloc failed");
    }
    c->comm = m;
    return HDG_OK;
}

hdg_status hdg_get_partition(const hdg_context* c, int64_t out[8]) {
    if (!c || !out) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(const_cast<hdg_context*>(c), HDG_ERR_INVALID, "no mesh");
    if (c->comm && c->comm->nranks > 1) {
        const Comm* m = c->comm;
        out[0] = m->cell_begin; out[1] = m->cell_begin + c->ncell_own;
        out[2] = m->face_begin; out[3] = m->face_begin + c->nface_own;
        out[4] = m->ncell_global; out[5] = m->nface_global;
        out[6] = c->ncell - c->ncell_own; out[7] = c->nface - c->nface_own;
    } else {
        out[0] = 0; out[1] = c->ncell; out[2] = 0; out[3] = c->nface; out[4] = c->ncell; out[5] = c->nface; out[6] = 0; out[7] = 0;
    }
    return HDG_OK;
}

}
