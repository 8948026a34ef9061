// Cross-GPU synchronisation over peer memory, usable from inside any kernel (hdg_comm.cu: xgpu_allreduce; hdg_mg.cu: the
// stage barriers of the persistent V-cycle kernel).  Every rank owns a mailbox [2][nranks][MAILW] doubles that all ranks have
// mapped (CUDA IPC over NVLink); slot [e & 1][q] receives rank q's contribution of exchange number e (values first, then the
// epoch word).  All ranks must perform the same sequence of exchanges; the epoch counter lives in device memory and is
// shared by every user of the mailbox on this rank.
#pragma once
#include <stdint.h>

namespace hdg {

constexpr int XG_MAILW = 8;   // == MAILW (hdg_internal.h): doubles per mailbox slot, word 0 = epoch

struct XgComm {
    double* const* peer_mail = nullptr;   // device array [nranks]: every rank's mailbox as mapped here
    double* my_mail = nullptr;
    unsigned long long* epoch = nullptr;
    int rank = 0, nranks = 1;
};

// Barrier across the GPUs, executed by ONE thread (the caller has already synchronised its own grid): everything this GPU
// wrote before the call is visible to every GPU after it returns there.  ~3 us on NVSwitch (tools/pingpong.py).
__device__ __forceinline__ void xg_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long xg_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ONE release fence before the flag stores and ONE acquire fence after the polls (relaxed system-scope accesses in between)
// instead of two sequentially-consistent system fences: the data this GPU wrote is already in its L2 (every block fenced at GPU
// scope before arriving), which is where the peers' loads and the peers' pushed rows are served
__device__ __forceinline__ void xg_barrier_thread(const XgComm& x) {
    const unsigned long long e = *x.epoch + 1;
    *x.epoch = e;
    const int buf = int(e & 1ull);
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    for (int q = 0; q < x.nranks; ++q)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(x.peer_mail[q] + size_t(buf * x.nranks + x.rank) * XG_MAILW), "l"(e) : "memory");
    for (int q = 0; q < x.nranks; ++q) {
        const double* flag = x.my_mail + size_t(buf * x.nranks + q) * XG_MAILW;
        unsigned long long v;
        do { asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); } while (v != e);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

}  // namespace hdg
