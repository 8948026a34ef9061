// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process; this is the sharding of the same hot path:
// assembly / condensation / recovery are embarrassingly parallel over strips of cells (a one-cell
// ghost layer is recomputed instead of exchanged), the PCG needs one halo exchange of interface
// trace values per SpMV and all-reduced dot products.  NCCL is loaded with dlopen so that the
// single-GPU library has no NCCL dependency; the caller distributes the ncclUniqueId
// (torch.distributed / MPI / Julia Distributed are all fine - plain bytes through the C ABI).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include "hdg_internal.h"
#include "hdg_xgpu.cuh"

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
    nccl_result_t (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
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
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
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

// ---------------------------------------------------------------------------------------------
// Peer-memory path.  Every rank exposes (a) a small mailbox and (b) its PCG vector region through CUDA
// IPC; the handles travel through one ncclAllGather.  After that the PCG needs no NCCL call: the SpMV
// kernel loads the neighbours' interface values straight over NVLink, and the dot products are
// combined by `xgpu_allreduce`, a one-block kernel that stores this rank's partial sums into every
// rank's mailbox and spins (system-scope loads) until all ranks of the current epoch have arrived -
// an all-reduce and a device-wide barrier across GPUs in ~one NVLink round trip.
// ---------------------------------------------------------------------------------------------
struct IpcRecord {
    cudaIpcMemHandle_t handle;   // 64 bytes
    int64_t ndof;
    int64_t pad;
};

static hdg_status allgather_records(hdg_context* c, const IpcRecord& mine, std::vector<IpcRecord>& all) {
    Comm* m = c->comm;
    if (!g_nccl.AllGather) return set_err(c, HDG_ERR_NCCL, "ncclAllGather missing");
    IpcRecord* d = nullptr;
    HDG_CUDA(c, cudaMalloc(&d, sizeof(IpcRecord) * (m->nranks + 1)));
    HDG_CUDA(c, cudaMemcpyAsync(d + m->nranks, &mine, sizeof(IpcRecord), cudaMemcpyHostToDevice, c->stream));
    HDG_NCCL(c, g_nccl.AllGather(d + m->nranks, d, sizeof(IpcRecord), 0 /* ncclInt8 */, m->nccl, c->stream));
    all.resize(m->nranks);
    HDG_CUDA(c, cudaMemcpyAsync(all.data(), d, sizeof(IpcRecord) * m->nranks, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return HDG_OK;
}

bool comm_p2p(const hdg_context* c) { return comm_active(c) && c->comm->p2p; }

static hdg_status setup_mailboxes(hdg_context* c) {
    Comm* m = c->comm;
    const size_t bytes = sizeof(double) * 2 * m->nranks * MAILW;
    HDG_CUDA(c, cudaMalloc(&m->d_mail, bytes));
    HDG_CUDA(c, cudaMemset(m->d_mail, 0, bytes));
    HDG_CUDA(c, cudaMalloc(&m->d_epoch, sizeof(unsigned long long)));
    HDG_CUDA(c, cudaMemset(m->d_epoch, 0, sizeof(unsigned long long)));
    IpcRecord mine{};
    HDG_CUDA(c, cudaIpcGetMemHandle(&mine.handle, m->d_mail));
    std::vector<IpcRecord> all;
    hdg_status st = allgather_records(c, mine, all);
    if (st) return st;
    std::vector<double*> ptrs(m->nranks, nullptr);
    for (int q = 0; q < m->nranks; ++q) {
        if (q == m->rank) { ptrs[q] = m->d_mail; continue; }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(c, HDG_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        m->ipc_opened.push_back(p);
        ptrs[q] = static_cast<double*>(p);
    }
    HDG_CUDA(c, cudaMalloc(&m->d_peer_mail, sizeof(double*) * m->nranks));
    HDG_CUDA(c, cudaMemcpy(m->d_peer_mail, ptrs.data(), sizeof(double*) * m->nranks, cudaMemcpyHostToDevice));
    return HDG_OK;
}

void comm_unshare_vectors(hdg_context* c) {
    if (!c->comm) return;
    for (int w = 0; w < MAXR; ++w)
        if (c->comm->peer_vec[w]) { cudaIpcCloseMemHandle(c->comm->peer_vec[w]); c->comm->peer_vec[w] = nullptr; }
}

hdg_status comm_share_vectors(hdg_context* c, void* region, int64_t ndof_own) {
    Comm* m = c->comm;
    comm_unshare_vectors(c);
    IpcRecord mine{};
    HDG_CUDA(c, cudaIpcGetMemHandle(&mine.handle, region));
    mine.ndof = ndof_own;
    std::vector<IpcRecord> all;
    hdg_status st = allgather_records(c, mine, all);
    if (st) return st;
    for (int q = 0; q < m->nranks && q < MAXR; ++q) {
        if (q == m->rank) continue;      // every rank's region: a later mesh of the same sizes may have other neighbours
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(c, HDG_ERR_CUDA, std::string("cudaIpcOpenMemHandle(vectors): ") + cudaGetErrorString(e));
        }
        m->peer_vec[q] = p;
        m->peer_ndof[q] = all[q].ndof;
    }
    return HDG_OK;
}

hdg_status comm_share_buffer(hdg_context* c, void* mine, unsigned rank_mask, void* peers[MAXR]) {
    Comm* m = c->comm;
    for (int q = 0; q < MAXR; ++q) peers[q] = nullptr;
    IpcRecord rec{};
    HDG_CUDA(c, cudaIpcGetMemHandle(&rec.handle, mine));
    std::vector<IpcRecord> all;
    hdg_status st = allgather_records(c, rec, all);
    if (st) return st;
    for (int q = 0; q < m->nranks && q < MAXR; ++q) {
        if (q == m->rank) { peers[q] = mine; continue; }
        if (!((rank_mask >> q) & 1u)) continue;
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(c, HDG_ERR_CUDA, std::string("cudaIpcOpenMemHandle(buffer): ") + cudaGetErrorString(e));
        }
        peers[q] = p;
    }
    return HDG_OK;
}

void comm_close_buffer(hdg_context* c, void* peers[MAXR]) {
    const int me = c->comm ? c->comm->rank : -1;
    for (int q = 0; q < MAXR; ++q) {
        if (peers[q] && q != me) cudaIpcCloseMemHandle(peers[q]);
        peers[q] = nullptr;
    }
}

void comm_xg(const hdg_context* c, XgComm* out) {
    *out = XgComm{};
    if (!comm_p2p(c)) return;
    const Comm* m = c->comm;
    out->peer_mail = m->d_peer_mail; out->my_mail = m->d_mail; out->epoch = m->d_epoch; out->rank = m->rank; out->nranks = m->nranks;
}

hdg_status comm_set_ghosts(hdg_context* c, const std::vector<int32_t>& ridx, const std::vector<int32_t>& owner) {
    Comm* m = c->comm;
    if (m->d_ghost_ridx) { cudaFree(m->d_ghost_ridx); m->d_ghost_ridx = nullptr; }
    if (m->d_ghost_owner) { cudaFree(m->d_ghost_owner); m->d_ghost_owner = nullptr; }
    m->need_rank = 0;
    for (int32_t q : owner) {
        if (q < 0 || q >= MAXR) return set_err(c, HDG_ERR_INVALID, "ghost face owned by a rank outside the box");
        m->need_rank |= 1u << q;
    }
    if (ridx.empty()) return HDG_OK;
    HDG_CUDA(c, cudaMalloc(&m->d_ghost_ridx, sizeof(int32_t) * ridx.size()));
    HDG_CUDA(c, cudaMalloc(&m->d_ghost_owner, sizeof(int32_t) * owner.size()));
    HDG_CUDA(c, cudaMemcpy(m->d_ghost_ridx, ridx.data(), sizeof(int32_t) * ridx.size(), cudaMemcpyHostToDevice));
    HDG_CUDA(c, cudaMemcpy(m->d_ghost_owner, owner.data(), sizeof(int32_t) * owner.size(), cudaMemcpyHostToDevice));
    return HDG_OK;
}

static_assert(XG_MAILW == MAILW, "mailbox slot width");
constexpr int XG_THREADS = 256;
constexpr int XG_MAXPART = 2048;   // == MAX_PARTIALS of hdg_solve.cu

__global__ void __launch_bounds__(XG_THREADS) xgpu_allreduce(const double* __restrict__ part, int np, unsigned slot_mask,
                                                             double* __restrict__ gscal, double* const* __restrict__ peer_mail,
                                                             double* my_mail, unsigned long long* epoch_ctr, int rank, int nranks) {
    __shared__ double loc[MAILW];
    __shared__ unsigned long long e_sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // warp w adds the partials of slot w in a fixed order (lane-strided, then a shuffle tree)
    if (wid < MAILW - 1 && ((slot_mask >> wid) & 1u)) {
        double s = 0.0;
        for (int i = lane; i < np; i += 32) s += part[wid * XG_MAXPART + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) loc[wid] = s;
    }
    if (threadIdx.x == 0) {
        e_sh = *epoch_ctr + 1;
        *epoch_ctr = e_sh;
    }
    __syncthreads();
    const unsigned long long e = e_sh;
    const int buf = int(e & 1ull);
    if (threadIdx.x < nranks) {
        // my contribution -> slot [buf][rank] of rank `threadIdx.x` (values first, then the epoch word)
        volatile double* dst = peer_mail[threadIdx.x] + size_t(buf * nranks + rank) * MAILW;
        for (int w = 0; w < MAILW - 1; ++w)
            if ((slot_mask >> w) & 1u) dst[1 + w] = loc[w];
        xg_st_release_sys(reinterpret_cast<unsigned long long*>(const_cast<double*>(dst)), e);      // values first, then the epoch word
        // wait for rank `threadIdx.x`'s contribution to arrive in my mailbox
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(my_mail + size_t(buf * nranks + threadIdx.x) * MAILW);
        while (xg_ld_acquire_sys(flag) != e) { }
    }
    __syncthreads();
    if (threadIdx.x < MAILW - 1 && ((slot_mask >> threadIdx.x) & 1u)) {
        double s = 0.0;
        for (int q = 0; q < nranks; ++q)   // rank order: identical result on every rank
            s += reinterpret_cast<volatile double*>(my_mail)[size_t(buf * nranks + q) * MAILW + 1 + threadIdx.x];
        gscal[threadIdx.x] = s;
    }
}

hdg_status comm_p2p_allreduce(hdg_context* c, const double* d_partials, int np, unsigned slot_mask) {
    Comm* m = c->comm;
    xgpu_allreduce<<<1, XG_THREADS, 0, c->stream>>>(d_partials, np, slot_mask, m->d_gscal, m->d_peer_mail, m->d_mail, m->d_epoch,
                                                   m->rank, m->nranks);
    c->launches += 1;
    HDG_CUDA(c, cudaGetLastError());
    return HDG_OK;
}

void comm_destroy(hdg_context* c) {
    if (!c->comm) return;
    comm_free_halo(c);
    comm_unshare_vectors(c);
    for (void* p : c->comm->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->comm->d_mail) cudaFree(c->comm->d_mail);
    if (c->comm->d_peer_mail) cudaFree(c->comm->d_peer_mail);
    if (c->comm->d_epoch) cudaFree(c->comm->d_epoch);
    if (c->comm->d_ghost_ridx) cudaFree(c->comm->d_ghost_ridx);
    if (c->comm->d_ghost_owner) cudaFree(c->comm->d_ghost_owner);
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
    // peer-memory mailboxes; on failure (no P2P / IPC) the NCCL path stays in use
    if (nranks > 1 && getenv("HDG_NO_P2P") == nullptr && nranks <= MAXR) {
        hdg_status st = setup_mailboxes(c);
        m->p2p = st == HDG_OK;
        if (st != HDG_OK) c->err = "peer-memory path disabled: " + c->err;
    }
    return HDG_OK;
}

// Debug / measurement: round-trip latency of the mailbox exchange between all ranks (iters exchanges in ONE kernel).
__global__ void pingpong_kernel(double* const* peer_mail, double* my_mail, int rank, int nranks, int iters, unsigned long long tag0) {
    for (int it = 0; it < iters; ++it) {
        unsigned long long tag = tag0 + it + 1;
        int buf = int(tag & 1ull);
        if (threadIdx.x < nranks) {
            volatile double* dst = peer_mail[threadIdx.x] + size_t(buf * nranks + rank) * MAILW;
            dst[1] = double(it);
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(dst) = tag;
            volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(my_mail + size_t(buf * nranks + threadIdx.x) * MAILW);
            while (*flag != tag) { }
        }
        __syncthreads();
    }
}

hdg_status hdg_comm_pingpong(hdg_context* c, int32_t iters, double* usec_per_exchange) {
    if (!c || !usec_per_exchange) return HDG_ERR_INVALID;
    if (!comm_p2p(c)) return set_err(c, HDG_ERR_INVALID, "peer-memory path not active");
    Comm* m = c->comm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    // tags far above anything xgpu_allreduce uses in this mailbox
    static unsigned long long tag0 = 1ull << 40;
    pingpong_kernel<<<1, 32, 0, c->stream>>>(m->d_peer_mail, m->d_mail, m->rank, m->nranks, 8, tag0);
    tag0 += 1024;
    cudaEventRecord(a, c->stream);
    pingpong_kernel<<<1, 32, 0, c->stream>>>(m->d_peer_mail, m->d_mail, m->rank, m->nranks, iters, tag0);
    cudaEventRecord(b, c->stream);
    tag0 += (unsigned long long)iters + 1024;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    *usec_per_exchange = 1e3 * ms / iters;
    return HDG_OK;
}

hdg_status hdg_get_ghost_cells(const hdg_context* c, int64_t* ids) {
    if (!c || !ids) return HDG_ERR_INVALID;
    if (!c->have_mesh) return set_err(const_cast<hdg_context*>(c), HDG_ERR_INVALID, "no mesh");
    if (c->comm && c->comm->nranks > 1)
        for (size_t i = 0; i < c->comm->ghost_cells.size(); ++i) ids[i] = c->comm->ghost_cells[i];
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
