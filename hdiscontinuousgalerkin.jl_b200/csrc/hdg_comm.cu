// Multi-GPU plumbing (one process per GPU).  Placeholder until the NCCL halo-PCG lands.
#include "hdg_internal.h"

using namespace hdg;

extern "C" {

hdg_status hdg_comm_unique_id(uint8_t id_out[128]) {
    (void)id_out;
    return set_err(nullptr, HDG_ERR_NCCL, "multi-GPU support not built");
}

hdg_status hdg_comm_init(hdg_context* c, int32_t rank, int32_t nranks, const uint8_t id[128]) {
    (void)rank; (void)nranks; (void)id;
    return set_err(c, HDG_ERR_NCCL, "multi-GPU support not built");
}

}
