// Locality-preserving cell order for hdg_set_mesh on several GPUs ("a real partitioner").
//
// hdg_set_mesh partitions a mesh into contiguous ranges of cell ids (and of the face ids those cells create, which the
// reference's first-encounter numbering makes contiguous as well - src/generate_mesh.jl:20-46, src/triangle_mesh.jl:66-101),
// so a partition is exactly as local as the numbering of the input.  A mesh generator's output (Triangle, Delaunay of scattered
// points, a refined mesh) has no such order: cells shuffled in blocks of 4 096 leave 1 M ghost cells per rank at 4 M cells.
// hdg_order_cells computes, on the device, the permutation that sorts the cells along the Morton (Z-order) curve through their
// centroids; cutting that order into R contiguous pieces gives R compact patches (the classical space-filling-curve
// partition).  The host mirror (api.py renumber_mesh) then rebuilds cells / faces in the new order with hdg_number_faces, hands
// the renumbered mesh to hdg_set_mesh and maps results back to the caller's numbering.
//
// The one-off sort of (key, cell) pairs uses cub::DeviceRadixSort from the CUDA toolkit - set-up code, not on the measured
// path; it is stable, so cells with equal keys keep their input order and the permutation is deterministic.
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>
#include <vector>

#include "hdg_internal.h"

namespace hdg {

constexpr int ORDER_BLOCKS = 256;

// per-block bounding box of the node coordinates: out[4 * block + {0: min x, 1: min y, 2: max x, 3: max y}]
__global__ void order_bbox(const double* __restrict__ nodes, int64_t nnode, double* __restrict__ out) {
    double lox = DBL_MAX, loy = DBL_MAX, hix = -DBL_MAX, hiy = -DBL_MAX;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nnode; i += int64_t(gridDim.x) * blockDim.x) {
        const double2 p = reinterpret_cast<const double2*>(nodes)[i];
        lox = fmin(lox, p.x); hix = fmax(hix, p.x);
        loy = fmin(loy, p.y); hiy = fmax(hiy, p.y);
    }
    __shared__ double s[4][256];
    s[0][threadIdx.x] = lox; s[1][threadIdx.x] = loy; s[2][threadIdx.x] = hix; s[3][threadIdx.x] = hiy;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
        if (int(threadIdx.x) < d) {
            s[0][threadIdx.x] = fmin(s[0][threadIdx.x], s[0][threadIdx.x + d]);
            s[1][threadIdx.x] = fmin(s[1][threadIdx.x], s[1][threadIdx.x + d]);
            s[2][threadIdx.x] = fmax(s[2][threadIdx.x], s[2][threadIdx.x + d]);
            s[3][threadIdx.x] = fmax(s[3][threadIdx.x], s[3][threadIdx.x + d]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) out[4 * blockIdx.x + threadIdx.x] = s[threadIdx.x][0];
}

// spread the 32 bits of v over the even bits of a 64-bit word
__device__ __forceinline__ uint64_t spread_bits(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// Morton key of the centroid of every cell on a 2^31 x 2^31 lattice over the bounding box; ids out of range flag the mesh
__global__ void order_keys(const int64_t* __restrict__ tri, int64_t ncell, const double* __restrict__ nodes, int64_t nnode,
                           double x0, double y0, double sx, double sy, uint64_t* __restrict__ key, int32_t* __restrict__ idx,
                           int32_t* __restrict__ bad) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double cx = 0.0, cy = 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t v = tri[3 * c + k];
        if (v < 1 || v > nnode) { ok = false; continue; }
        const double2 p = reinterpret_cast<const double2*>(nodes)[v - 1];
        cx = __dadd_rn(cx, p.x); cy = __dadd_rn(cy, p.y);
    }
    if (!ok) atomicCAS(bad, 0, 1);
    // every product / difference rounded on its own (no FMA contraction): the keys are reproducible on the host
    const double fx = fmin(fmax(__dmul_rn(__dsub_rn(__dmul_rn(cx, 1.0 / 3.0), x0), sx), 0.0), 2147483647.0);
    const double fy = fmin(fmax(__dmul_rn(__dsub_rn(__dmul_rn(cy, 1.0 / 3.0), y0), sy), 0.0), 2147483647.0);
    key[c] = spread_bits(uint32_t(fx)) | (spread_bits(uint32_t(fy)) << 1);
    idx[c] = int32_t(c);
}

__global__ void order_emit(const int32_t* __restrict__ idx, int64_t ncell, int64_t* __restrict__ perm) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < ncell) perm[i] = int64_t(idx[i]) + 1;
}

hdg_status order_cells_host(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode, int64_t* perm) {
    if (ncell >= (int64_t(1) << 31)) return set_err(c, HDG_ERR_INVALID, "mesh too large for 32-bit device ids");
    int64_t *d_tri = nullptr, *d_perm = nullptr;
    double *d_nodes = nullptr, *d_box = nullptr;
    uint64_t *d_key = nullptr, *d_key2 = nullptr;
    int32_t *d_idx = nullptr, *d_idx2 = nullptr, *d_bad = nullptr;
    void* d_tmp = nullptr;
    hdg_status st = HDG_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && st == HDG_OK) st = set_err(c, HDG_ERR_CUDA, cudaGetErrorString(e)); return e != cudaSuccess; };
    do {
        if (fail(cudaMalloc(&d_tri, sizeof(int64_t) * 3 * ncell))) break;
        if (fail(cudaMalloc(&d_nodes, sizeof(double) * 2 * nnode))) break;
        if (fail(cudaMalloc(&d_box, sizeof(double) * 4 * ORDER_BLOCKS))) break;
        if (fail(cudaMalloc(&d_key, sizeof(uint64_t) * ncell)) || fail(cudaMalloc(&d_key2, sizeof(uint64_t) * ncell))) break;
        if (fail(cudaMalloc(&d_idx, sizeof(int32_t) * ncell)) || fail(cudaMalloc(&d_idx2, sizeof(int32_t) * ncell))) break;
        if (fail(cudaMalloc(&d_bad, sizeof(int32_t))) || fail(cudaMalloc(&d_perm, sizeof(int64_t) * ncell))) break;
        if (fail(cudaMemcpyAsync(d_tri, tri, sizeof(int64_t) * 3 * ncell, cudaMemcpyHostToDevice, c->stream))) break;
        if (fail(cudaMemcpyAsync(d_nodes, nodes, sizeof(double) * 2 * nnode, cudaMemcpyHostToDevice, c->stream))) break;
        if (fail(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), c->stream))) break;
        order_bbox<<<ORDER_BLOCKS, 256, 0, c->stream>>>(d_nodes, nnode, d_box);
        std::vector<double> box(4 * ORDER_BLOCKS);
        if (fail(cudaMemcpyAsync(box.data(), d_box, sizeof(double) * box.size(), cudaMemcpyDeviceToHost, c->stream))) break;
        if (fail(cudaStreamSynchronize(c->stream))) break;
        double lox = DBL_MAX, loy = DBL_MAX, hix = -DBL_MAX, hiy = -DBL_MAX;
        for (int b = 0; b < ORDER_BLOCKS; ++b) {
            lox = std::min(lox, box[4 * b]); loy = std::min(loy, box[4 * b + 1]);
            hix = std::max(hix, box[4 * b + 2]); hiy = std::max(hiy, box[4 * b + 3]);
        }
        if (!(hix >= lox) || !(hiy >= loy)) { st = set_err(c, HDG_ERR_INVALID, "node coordinates are not finite"); break; }
        // one lattice step for both directions: the curve then follows the geometry, not the aspect ratio of the box
        const double ext = std::max(std::max(hix - lox, hiy - loy), DBL_MIN);
        const double scale = 2147483647.0 / ext;
        const int B = 256;
        order_keys<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(d_tri, ncell, d_nodes, nnode, lox, loy, scale, scale, d_key, d_idx, d_bad);
        size_t tmp_bytes = 0;
        if (fail(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_idx, d_idx2, int(ncell), 0, 64, c->stream))) break;
        if (fail(cudaMalloc(&d_tmp, tmp_bytes))) break;
        if (fail(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_idx, d_idx2, int(ncell), 0, 64, c->stream))) break;
        order_emit<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(d_idx2, ncell, d_perm);
        c->launches += 3;
        int32_t bad = 0;
        if (fail(cudaMemcpyAsync(&bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream))) break;
        if (fail(cudaMemcpyAsync(perm, d_perm, sizeof(int64_t) * ncell, cudaMemcpyDeviceToHost, c->stream))) break;
        if (fail(cudaStreamSynchronize(c->stream))) break;
        if (fail(cudaGetLastError())) break;
        if (bad) st = set_err(c, HDG_ERR_INVALID, "mesh arrays: node id of a cell out of range (ids are 1-based: nodes 1..nnode)");
    } while (false);
    cudaFree(d_tri); cudaFree(d_nodes); cudaFree(d_box); cudaFree(d_key); cudaFree(d_key2); cudaFree(d_idx); cudaFree(d_idx2);
    cudaFree(d_bad); cudaFree(d_perm); cudaFree(d_tmp);
    return st;
}

}  // namespace hdg
