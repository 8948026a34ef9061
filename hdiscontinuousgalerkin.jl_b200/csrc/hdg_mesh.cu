// K0: mesh ingestion / generation, adjacency and sparsity pattern - integer work on the device.
//
// Replaces: rectangle_mesh + _build_cells + _generate_2d_nodes! + boundary sets
// (src/generate_mesh.jl:1-143), the CellIterator gathers (src/iterator.jl:48-57), the dof map
// gdof = face*nt-(nt-j) (examples/poisson2D_HDG.jl:176-181) and the pattern of sparse(I,J,V)
// (src/assembler.jl:47-49).
#include <algorithm>

#include "hdg_internal.h"

namespace hdg {

// ------------------------------------------------------------------------------------------
// small device scan (exclusive, int32 -> int64), three phases, deterministic
// ------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;   // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ inline int64_t block_exclusive_scan(int64_t v, int64_t* total) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int64_t s = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    int64_t base = wid > 0 ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[SCAN_BLOCK / 32 - 1];
    int64_t r = base + x - v;
    __syncthreads();
    return r;
}

__global__ void scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
    int64_t base = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int64_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void scan_tile_offsets(int64_t* tile_sums, int64_t ntiles, int64_t* grand_total) {
    // single block; serial over chunks of SCAN_BLOCK
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t c = 0; c < ntiles; c += SCAN_BLOCK) {
        int64_t i = c + threadIdx.x;
        int64_t v = i < ntiles ? tile_sums[i] : 0;
        int64_t tot;
        int64_t ex = block_exclusive_scan(v, &tot);
        int64_t cr = carry;
        if (i < ntiles) tile_sums[i] = cr + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = cr + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_offs,
                           int64_t* __restrict__ out) {
    int64_t base = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    int64_t ex = block_exclusive_scan(s, nullptr) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

// out[i] = sum_{j<i} in[j]; *d_total = sum of all. d_total is a device pointer.
static hdg_status exclusive_scan(hdg_context* c, const int32_t* in, int64_t n, int64_t* out, int64_t* d_total) {
    int64_t ntiles = ceil_div(n, SCAN_TILE);
    int64_t* tile_sums = nullptr;
    HDG_CUDA(c, cudaMalloc(&tile_sums, sizeof(int64_t) * std::max<int64_t>(ntiles, 1)));
    scan_tile_sums<<<(unsigned)ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, tile_sums);
    scan_tile_offsets<<<1, SCAN_BLOCK, 0, c->stream>>>(tile_sums, ntiles, d_total);
    scan_apply<<<(unsigned)ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, tile_sums, out);
    c->launches += 3;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    HDG_CUDA(c, cudaFree(tile_sums));
    return HDG_OK;
}

hdg_status exclusive_scan_i32(hdg_context* c, const int32_t* in, int64_t n, int64_t* out, int64_t* d_total) { return exclusive_scan(c, in, n, out, d_total); }

// ------------------------------------------------------------------------------------------
// mesh from host arrays (Julia layouts, 1-based int64)
// ------------------------------------------------------------------------------------------
// Range check of the 1-based ids handed over by hdg_set_mesh (the reference would throw a BoundsError at the first use):
// node ids in [1, nnode], face ids in [1, nface], cell ids of the face table in [0, ncell].  flag = 1 + first offender kind.
__global__ void check_mesh_ids(const int64_t* __restrict__ cells, int64_t ncell, int64_t nnode, const int64_t* __restrict__ faces,
                               int64_t nface, int32_t* __restrict__ flag) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < ncell) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int64_t v = cells[6 * i + k], f = cells[6 * i + 3 + k];
            if (v < 1 || v > nnode) atomicCAS(flag, 0, 1);
            if (f < 1 || f > nface) atomicCAS(flag, 0, 2);
        }
    }
    if (faces && i < nface) {
        const int64_t v1 = faces[i], v2 = faces[i + nface], c1 = faces[i + 2 * nface], c2 = faces[i + 3 * nface];
        if (v1 < 1 || v1 > nnode || v2 < 1 || v2 > nnode) atomicCAS(flag, 0, 3);
        if (c1 < 1 || c1 > ncell || c2 < 0 || c2 > ncell) atomicCAS(flag, 0, 4);
    }
}

__global__ void convert_faces(const int64_t* __restrict__ faces, int64_t nface, int32_t* __restrict__ facecell,
                              int32_t* __restrict__ facenode) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    facenode[2 * f + 0] = int32_t(faces[f] - 1);
    facenode[2 * f + 1] = int32_t(faces[f + nface] - 1);
    facecell[2 * f + 0] = int32_t(faces[f + 2 * nface] - 1);
    facecell[2 * f + 1] = int32_t(faces[f + 3 * nface] - 1);   // 0 -> -1 (no second cell)
}

// faces == NULL at the boundary: rebuild the face table from the cells alone.  A face's first cell is the one that
// encountered it first in the reference's sequential numbering (src/generate_mesh.jl:20-46) = the adjacent cell
// with the smaller id; (v1,v2) is that cell's local direction.
__global__ void derive_facecell(const int64_t* __restrict__ cells, int64_t ncell, int32_t* __restrict__ facecell) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int64_t f = cells[6 * c + 3 + k] - 1;
        atomicMin(&facecell[2 * f + 0], int32_t(c));
        atomicMax(&facecell[2 * f + 1], int32_t(c));
    }
}
__global__ void init_facecell(int32_t* __restrict__ facecell, int64_t nface) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f < nface) *reinterpret_cast<int2*>(facecell + 2 * f) = make_int2(0x7fffffff, -1);
}
__global__ void derive_facenode(const int64_t* __restrict__ cells, int64_t ncell, int32_t* __restrict__ facecell,
                                int32_t* __restrict__ facenode) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int k1[3] = {1, 2, 0}, k2[3] = {2, 0, 1};   // reference_edge_nodes ((2,3),(3,1),(1,2)), src/mesh.jl:26
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int64_t f = cells[6 * c + 3 + k] - 1;
        if (facecell[2 * f] == int32_t(c)) {
            facenode[2 * f + 0] = int32_t(cells[6 * c + k1[k]] - 1);
            facenode[2 * f + 1] = int32_t(cells[6 * c + k2[k]] - 1);
        }
    }
}
__global__ void finish_facecell(int32_t* __restrict__ facecell, int64_t nface) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    if (facecell[2 * f + 1] == facecell[2 * f]) facecell[2 * f + 1] = -1;   // one adjacent cell: boundary face
}

__global__ void convert_cells(const int64_t* __restrict__ cells, int64_t ncell, const int32_t* __restrict__ facecell,
                              int32_t* __restrict__ cellinfo) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) cellinfo[CI * c + k] = int32_t(cells[6 * c + k] - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int32_t f = int32_t(cells[6 * c + 3 + k] - 1);
        uint32_t sec = facecell[2 * f + 1] == int32_t(c) ? 0x80000000u : 0u;
        cellinfo[CI * c + 3 + k] = int32_t(uint32_t(f) | sec);
    }
}

__global__ void build_kcol(const int32_t* __restrict__ cellinfo, int64_t ncell, int32_t* __restrict__ kcol) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    uint32_t fr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) fr[k] = uint32_t(cellinfo[CI * c + 3 + k]);
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        int64_t f = fr[l] & 0x7fffffffu;
        int sec = fr[l] >> 31;
        kcol[4 * f + 2 * sec + 0] = int32_t(fr[(l + 1) % 3] & 0x7fffffffu);
        kcol[4 * f + 2 * sec + 1] = int32_t(fr[(l + 2) % 3] & 0x7fffffffu);
    }
}

// In-warp pairing hints for the element kernel: for every local face, is the neighbour cell in the same
// 32-cell tile (= warp of the element kernel), at which lane, and which of its local faces is the shared one.
__global__ void build_partner(int32_t* __restrict__ cellinfo, int64_t ncell, const int32_t* __restrict__ facecell) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    uint32_t pw = 0, bw = 0;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        int32_t f = int32_t(uint32_t(cellinfo[CI * c + 3 + l]) & 0x7fffffffu);
        int32_t c1 = facecell[2 * f], c2 = facecell[2 * f + 1];
        if (c2 < 0) { bw |= 1u << l; continue; }
        int64_t o = c1 == int32_t(c) ? c2 : c1;
        if ((o >> 5) != (c >> 5)) continue;
        int lo = 0;
        for (int k = 0; k < 3; ++k)
            if (int32_t(uint32_t(cellinfo[CI * o + 3 + k]) & 0x7fffffffu) == f) lo = k;
        pw |= (0x80u | uint32_t(o & 31) | (uint32_t(lo) << 5)) << (8 * l);
    }
    cellinfo[CI * c + 6] = int32_t(pw);
    cellinfo[CI * c + 7] = int32_t(bw);
}

__global__ void mark_bfaces(const int32_t* __restrict__ bfaces, int64_t nb, const int32_t* __restrict__ facecell,
                            uint8_t* __restrict__ isbc, int32_t* __restrict__ flags) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    int32_t f = bfaces[i];
    if (facecell[2 * f + 1] >= 0) atomicExch(&flags[FLAG_NOT_BOUNDARY], f + 1);  // src/boundary.jl:22
    isbc[f] = 1;
}

// ------------------------------------------------------------------------------------------
// structured mesh on the device: closed forms of the first-encounter numbering
// ------------------------------------------------------------------------------------------
__global__ void rect_nodes(double* __restrict__ nodes, int64_t nnx, int64_t nny_global, int64_t row0, int64_t nrows,
                           double llx, double lly, double urx, double ury) {
    int64_t id = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (id >= nnx * nrows) return;
    int64_t i = row0 + id / nnx, j = id % nnx;   // global row i, column j (src/generate_mesh.jl:1-18)
    // every product / sum rounded separately, as the Julia expression does (no FMA contraction)
    double rb = double(i) / double(nny_global - 1);
    double omr = __dsub_rn(1.0, rb);
    // LL=(llx,lly) UL=(llx,ury) LR=(urx,lly) UR=(urx,ury)
    double x0 = __dadd_rn(__dmul_rn(llx, omr), __dmul_rn(rb, llx));
    double x1 = __dadd_rn(__dmul_rn(urx, omr), __dmul_rn(rb, urx));
    double y0 = __dadd_rn(__dmul_rn(lly, omr), __dmul_rn(rb, ury));
    double y1 = __dadd_rn(__dmul_rn(lly, omr), __dmul_rn(rb, ury));
    double r = double(j) / double(nnx - 1);
    double om = __dsub_rn(1.0, r);
    nodes[2 * id + 0] = __dadd_rn(__dmul_rn(x0, om), __dmul_rn(r, x1));
    nodes[2 * id + 1] = __dadd_rn(__dmul_rn(y0, om), __dmul_rn(r, y1));
}

// 0-based face ids of quad (i,j), i in [0,nx), j in [0,ny)
struct QuadFaces { int64_t diag, left, bottom, top, right; };
__host__ __device__ inline int64_t quad_base(int64_t i, int64_t j, int64_t nx) {
    // number of faces created before quad (i,j) (0-based i,j)
    if (j == 0) return 4 * i + (i > 0 ? 1 : 0);
    return 4 * nx + 1 + (j - 1) * (3 * nx + 1) + 3 * i + (i > 0 ? 1 : 0);
}
__host__ __device__ inline QuadFaces quad_faces(int64_t i, int64_t j, int64_t nx) {
    QuadFaces q;
    int64_t b = quad_base(i, j, nx);
    int i0 = i == 0, j0 = j == 0;
    q.diag = b;
    q.top = b + 1 + i0 + j0;
    q.right = b + 2 + i0 + j0;
    if (i0) q.left = b + 1;
    else {
        int64_t bl = quad_base(i - 1, j, nx);
        q.left = bl + 2 + (i - 1 == 0) + j0;
    }
    if (j0) q.bottom = b + 1 + i0;
    else {
        int64_t bb = quad_base(i, j - 1, nx);
        q.bottom = bb + 1 + i0 + (j - 1 == 0);
    }
    return q;
}

// Strip [j0,j1) of quad rows of the global nx x ny mesh in LOCAL numbering:
//   cells : owned cells (global id - 2*nx*j0), then one ghost cell per column (the lower-left triangle of row j1)
//   faces : owned faces (global id - F0: everything the strip's quads create), then the ghost-below faces
//           (tops of row j0-1, owned by rank-1), then the ghost-above faces (left and diagonal of the ghost cells)
//   nodes : global id - j0*(nx+1)
// facecell: -1 = no cell (domain boundary), -2 = a cell of another rank.
struct Strip {
    int64_t nx, ny, j0, j1, F0, nown, nbelow, ncell_own, node0;
};

__global__ void rect_cells(Strip S, int32_t* __restrict__ cellinfo, int32_t* __restrict__ facecell,
                           int32_t* __restrict__ facenode) {
    const int64_t nx = S.nx, nnx = nx + 1;
    int64_t ql = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nrows = S.j1 - S.j0 + (S.j1 < S.ny ? 1 : 0);
    if (ql >= nx * nrows) return;
    const int64_t i = ql % nx, j = S.j0 + ql / nx;
    auto na = [&](int64_t a, int64_t b) { return int32_t(a + b * nnx - S.node0); };
    const uint32_t SEC = 0x80000000u;
    auto setf = [&](int64_t f, int32_t v1, int32_t v2, int64_t ca, int64_t cb) {
        facenode[2 * f] = v1; facenode[2 * f + 1] = v2;
        facecell[2 * f] = int32_t(ca); facecell[2 * f + 1] = int32_t(cb);
    };
    const int64_t ghost_above0 = S.nown + S.nbelow;
    if (j == S.j1) {
        // ghost cell: lower-left triangle of quad (i, j1); its bottom face is the owned interface face
        const int64_t cg = S.ncell_own + i;
        const int64_t f_left = ghost_above0 + 2 * i, f_diag = f_left + 1;
        const int64_t f_bottom = quad_faces(i, j - 1, nx).top - S.F0;
        cellinfo[CI * cg + 0] = na(i, j);
        cellinfo[CI * cg + 1] = na(i + 1, j);
        cellinfo[CI * cg + 2] = na(i, j + 1);
        cellinfo[CI * cg + 3] = int32_t(f_diag);
        cellinfo[CI * cg + 4] = int32_t(uint32_t(f_left) | (i > 0 ? SEC : 0u));
        cellinfo[CI * cg + 5] = int32_t(uint32_t(f_bottom) | SEC);
        setf(f_diag, na(i + 1, j), na(i, j + 1), cg, -2);
        if (i == 0) setf(f_left, na(i, j + 1), na(i, j), cg, -1);
        else setf(f_left, na(i, j), na(i, j + 1), -2, cg);   // right face of quad (i-1, j1): created by a remote cell
        return;
    }
    QuadFaces F = quad_faces(i, j, nx);
    const int64_t q = (j - S.j0) * nx + i;
    const int64_t c0 = 2 * q, c1 = 2 * q + 1;
    const bool bottom_ghost = j == S.j0 && S.j0 > 0;
    const int64_t f_diag = F.diag - S.F0, f_left = F.left - S.F0, f_top = F.top - S.F0, f_right = F.right - S.F0;
    const int64_t f_bottom = bottom_ghost ? S.nown + i : F.bottom - S.F0;
    // lower-left triangle: nodes (na(i,j), na(i+1,j), na(i,j+1)), faces (diag, left, bottom)
    cellinfo[CI * c0 + 0] = na(i, j);
    cellinfo[CI * c0 + 1] = na(i + 1, j);
    cellinfo[CI * c0 + 2] = na(i, j + 1);
    cellinfo[CI * c0 + 3] = int32_t(f_diag);
    cellinfo[CI * c0 + 4] = int32_t(uint32_t(f_left) | (i > 0 ? SEC : 0u));
    cellinfo[CI * c0 + 5] = int32_t(uint32_t(f_bottom) | (j > 0 ? SEC : 0u));
    // upper-right triangle: nodes (na(i+1,j), na(i+1,j+1), na(i,j+1)), faces (top, diag, right)
    cellinfo[CI * c1 + 0] = na(i + 1, j);
    cellinfo[CI * c1 + 1] = na(i + 1, j + 1);
    cellinfo[CI * c1 + 2] = na(i, j + 1);
    cellinfo[CI * c1 + 3] = int32_t(f_top);
    cellinfo[CI * c1 + 4] = int32_t(uint32_t(f_diag) | SEC);
    cellinfo[CI * c1 + 5] = int32_t(f_right);
    // faces created by this quad, (v1,v2) in the creating cell's local direction
    setf(f_diag, na(i + 1, j), na(i, j + 1), c0, c1);
    if (i == 0) setf(f_left, na(i, j + 1), na(i, j), c0, -1);
    if (j == 0) setf(f_bottom, na(i, j), na(i + 1, j), c0, -1);
    if (bottom_ghost) setf(f_bottom, na(i + 1, j), na(i, j), -2, c0);   // top face of quad (i, j0-1), created remotely
    int64_t above = -1;
    if (j + 1 < S.ny) above = j + 1 < S.j1 ? 2 * (q + nx) : S.ncell_own + i;
    setf(f_top, na(i + 1, j + 1), na(i, j + 1), c1, above);
    setf(f_right, na(i + 1, j), na(i + 1, j + 1), c1, i + 1 < nx ? 2 * (q + 1) : -1);
}

__global__ void flag_boundary(const int32_t* __restrict__ facecell, int64_t nface, int32_t* __restrict__ flag) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f < nface) flag[f] = facecell[2 * f + 1] == -1 ? 1 : 0;
}
__global__ void compact_boundary(const int32_t* __restrict__ flag, const int64_t* __restrict__ offs, int64_t nface,
                                 int32_t* __restrict__ bfaces, uint8_t* __restrict__ isbc) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    isbc[f] = uint8_t(flag[f]);
    if (flag[f]) bfaces[offs[f]] = int32_t(f);
}

__global__ void perturb_nodes_k(double* __restrict__ nodes, int64_t nnx, int64_t nny, double hx, double hy,
                                double frac, uint64_t seed) {
    int64_t id = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (id >= nnx * nny) return;
    int64_t i = id / nnx, j = id % nnx;
    if (i == 0 || j == 0 || i == nny - 1 || j == nnx - 1) return;   // boundary nodes stay
    uint64_t z = uint64_t(id) * 0x9E3779B97F4A7C15ull + seed;        // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    double a = double(z & 0xffffffffu) / 4294967296.0 - 0.5;
    double b = double(z >> 32) / 4294967296.0 - 0.5;
    nodes[2 * id + 0] += 2.0 * frac * hx * a;
    nodes[2 * id + 1] += 2.0 * frac * hy * b;
}

// ------------------------------------------------------------------------------------------
// First-encounter face numbering of an arbitrary triangle list on the device
// (src/generate_mesh.jl:20-46 / src/triangle_mesh.jl:66-101 are a sequential hash-table insert).
// Parallel equivalent: every (cell, local face) is an "encounter" at position p = 3*cell + local face.
//   1. an open-addressing hash table maps the edge key (min node, max node) to the SMALLEST position that
//      encounters it (atomicCAS on the key, atomicMin on the position);
//   2. a position is a first encounter iff the table holds it; an exclusive scan of those flags ranks the first
//      encounters in position order = the reference's face ids;
//   3. every position looks its face up; the first encounter writes (v1, v2, cell, 0), the other one writes cell2.
// ------------------------------------------------------------------------------------------
struct EdgeTable {
    unsigned long long* keys;   // 0 = empty (node ids are >= 1)
    int* minpos;
    unsigned long long mask;    // capacity - 1 (power of two)
};
__device__ __forceinline__ unsigned long long edge_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ void tri_nodes_ccw(const int64_t* __restrict__ tri, const double* __restrict__ nodes, int64_t c, int64_t v[3]) {
    v[0] = tri[3 * c]; v[1] = tri[3 * c + 1]; v[2] = tri[3 * c + 2];
    // _check_node_data, src/generate_mesh.jl:49-57: swap vertices 2 and 3 of a clockwise triangle
    double ax = nodes[2 * (v[1] - 1)] - nodes[2 * (v[0] - 1)], ay = nodes[2 * (v[1] - 1) + 1] - nodes[2 * (v[0] - 1) + 1];
    double bx = nodes[2 * (v[2] - 1)] - nodes[2 * (v[0] - 1)], by = nodes[2 * (v[2] - 1) + 1] - nodes[2 * (v[0] - 1) + 1];
    if (__dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx)) < 0) { int64_t t = v[1]; v[1] = v[2]; v[2] = t; }
}
__device__ __forceinline__ unsigned long long edge_key(const int64_t v[3], int l, int64_t* v1, int64_t* v2) {
    const int k1[3] = {1, 2, 0}, k2[3] = {2, 0, 1};
    int64_t a = v[k1[l]], b = v[k2[l]];
    if (v1) { *v1 = a; *v2 = b; }
    int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    return (static_cast<unsigned long long>(lo) << 32) | static_cast<unsigned long long>(hi);
}
__global__ void edge_insert(const int64_t* __restrict__ tri, const double* __restrict__ nodes, int64_t ncell, EdgeTable T) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= 3 * ncell) return;
    int64_t c = p / 3, v[3];
    int l = int(p - 3 * c);
    tri_nodes_ccw(tri, nodes, c, v);
    unsigned long long key = edge_key(v, l, nullptr, nullptr);
    unsigned long long h = edge_hash(key) & T.mask;
    while (true) {
        unsigned long long prev = atomicCAS(&T.keys[h], 0ull, key);
        if (prev == 0ull || prev == key) { atomicMin(&T.minpos[h], int(p)); return; }
        h = (h + 1) & T.mask;
    }
}
__device__ __forceinline__ unsigned long long edge_find(const EdgeTable& T, unsigned long long key) {
    unsigned long long h = edge_hash(key) & T.mask;
    while (T.keys[h] != key) h = (h + 1) & T.mask;
    return h;
}
__global__ void edge_flag_first(const int64_t* __restrict__ tri, const double* __restrict__ nodes, int64_t ncell, EdgeTable T,
                                int32_t* __restrict__ flag) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= 3 * ncell) return;
    int64_t c = p / 3, v[3];
    tri_nodes_ccw(tri, nodes, c, v);
    unsigned long long h = edge_find(T, edge_key(v, int(p - 3 * c), nullptr, nullptr));
    flag[p] = T.minpos[h] == int(p) ? 1 : 0;
}
__global__ void edge_number(const int64_t* __restrict__ tri, const double* __restrict__ nodes, int64_t ncell, EdgeTable T,
                            const int64_t* __restrict__ rank, int64_t nface, int64_t* __restrict__ cells_out,
                            int64_t* __restrict__ faces_out, int32_t* __restrict__ err) {
    int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= 3 * ncell) return;
    int64_t c = p / 3, v[3], v1, v2;
    int l = int(p - 3 * c);
    tri_nodes_ccw(tri, nodes, c, v);
    unsigned long long h = edge_find(T, edge_key(v, l, &v1, &v2));
    int first = T.minpos[h];
    int64_t f = rank[first];
    if (l == 0) { cells_out[6 * c] = v[0]; cells_out[6 * c + 1] = v[1]; cells_out[6 * c + 2] = v[2]; }
    cells_out[6 * c + 3 + l] = f + 1;
    if (first == int(p)) {            // column-major nface x 4: v1 v2 cell1 (cell2 stays 0 unless a second encounter writes it)
        faces_out[f] = v1; faces_out[f + nface] = v2; faces_out[f + 2 * nface] = c + 1;
    } else {
        unsigned long long old = atomicExch(reinterpret_cast<unsigned long long*>(&faces_out[f + 3 * nface]), static_cast<unsigned long long>(c + 1));
        if (old != 0ull) atomicExch(err, int32_t(f + 1));   // a third cell on one edge: non-manifold
    }
}

// tri: ncell x 3 Int64 (1-based node ids, any orientation); outputs in the Julia layouts (cells ncell x 6, faces nface x 4
// column-major).  *nface_out is the number of distinct edges.  Buffers are device pointers owned by the caller.
static hdg_status number_faces_device(hdg_context* c, const int64_t* d_tri, const double* d_nodes, int64_t ncell,
                                      int64_t* d_cells_out, int64_t* d_faces_out, int64_t* nface_out) {
    if (3 * ncell >= (int64_t(1) << 31)) return set_err(c, HDG_ERR_INVALID, "mesh too large");
    unsigned long long cap = 1;
    while (cap < static_cast<unsigned long long>(6 * ncell)) cap <<= 1;   // load factor <= 1/2 even if every edge were distinct
    EdgeTable T{};
    T.mask = cap - 1;
    int32_t* flag = nullptr;
    int64_t *rank = nullptr, *d_tot = nullptr;
    HDG_CUDA(c, cudaMalloc(&T.keys, sizeof(unsigned long long) * cap));
    HDG_CUDA(c, cudaMalloc(&T.minpos, sizeof(int) * cap));
    HDG_CUDA(c, cudaMalloc(&flag, sizeof(int32_t) * 3 * ncell));
    HDG_CUDA(c, cudaMalloc(&rank, sizeof(int64_t) * 3 * ncell));
    HDG_CUDA(c, cudaMalloc(&d_tot, sizeof(int64_t)));
    HDG_CUDA(c, cudaMemsetAsync(T.keys, 0, sizeof(unsigned long long) * cap, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(T.minpos, 0x7f, sizeof(int) * cap, c->stream));
    const int B = 256;
    const unsigned G = (unsigned)ceil_div(3 * ncell, B);
    edge_insert<<<G, B, 0, c->stream>>>(d_tri, d_nodes, ncell, T);
    edge_flag_first<<<G, B, 0, c->stream>>>(d_tri, d_nodes, ncell, T, flag);
    c->launches += 2;
    hdg_status st = exclusive_scan(c, flag, 3 * ncell, rank, d_tot);
    if (st) return st;
    int64_t nface = 0;
    HDG_CUDA(c, cudaMemcpy(&nface, d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost));
    *nface_out = nface;
    if (d_faces_out) {
        HDG_CUDA(c, cudaMemsetAsync(d_faces_out, 0, sizeof(int64_t) * 4 * nface, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
        edge_number<<<G, B, 0, c->stream>>>(d_tri, d_nodes, ncell, T, rank, nface, d_cells_out, d_faces_out, c->d_flags + FLAG_NOT_BOUNDARY);
        c->launches += 1;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    cudaFree(T.keys); cudaFree(T.minpos); cudaFree(flag); cudaFree(rank); cudaFree(d_tot);
    if (d_faces_out && c->h_flags[FLAG_NOT_BOUNDARY])
        return set_err(c, HDG_ERR_INVALID, "non-manifold mesh: an edge is shared by more than two cells (face " +
                                               std::to_string(c->h_flags[FLAG_NOT_BOUNDARY]) + ")");
    return HDG_OK;
}

// C-ABI helper behind hdg_number_faces: host in / host out
hdg_status number_faces_host(hdg_context* c, const int64_t* tri, int64_t ncell, const double* nodes, int64_t nnode,
                             int64_t* cells_out, int64_t* faces_out, int64_t faces_capacity, int64_t* nface_out) {
    int64_t *d_tri = nullptr, *d_cells = nullptr, *d_faces = nullptr;
    double* d_nodes = nullptr;
    HDG_CUDA(c, cudaMalloc(&d_tri, sizeof(int64_t) * 3 * ncell));
    HDG_CUDA(c, cudaMalloc(&d_nodes, sizeof(double) * 2 * nnode));
    HDG_CUDA(c, cudaMalloc(&d_cells, sizeof(int64_t) * 6 * ncell));
    HDG_CUDA(c, cudaMalloc(&d_faces, sizeof(int64_t) * 4 * 3 * ncell));
    HDG_CUDA(c, cudaMemcpyAsync(d_tri, tri, sizeof(int64_t) * 3 * ncell, cudaMemcpyHostToDevice, c->stream));
    HDG_CUDA(c, cudaMemcpyAsync(d_nodes, nodes, sizeof(double) * 2 * nnode, cudaMemcpyHostToDevice, c->stream));
    int64_t nface = 0;
    hdg_status st = number_faces_device(c, d_tri, d_nodes, ncell, nullptr, nullptr, &nface);   // count first
    if (st == HDG_OK) {
        *nface_out = nface;
        if (faces_out && cells_out) {
            if (faces_capacity < nface) st = set_err(c, HDG_ERR_INVALID, "faces buffer too small");
            else {
                st = number_faces_device(c, d_tri, d_nodes, ncell, d_cells, d_faces, &nface);
                if (st == HDG_OK) {
                    HDG_CUDA(c, cudaMemcpyAsync(cells_out, d_cells, sizeof(int64_t) * 6 * ncell, cudaMemcpyDeviceToHost, c->stream));
                    // device faces are column-major with leading dimension nface; so is the caller's (nface x 4)
                    HDG_CUDA(c, cudaMemcpyAsync(faces_out, d_faces, sizeof(int64_t) * 4 * nface, cudaMemcpyDeviceToHost, c->stream));
                    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
                }
            }
        }
    }
    cudaFree(d_tri); cudaFree(d_nodes); cudaFree(d_cells); cudaFree(d_faces);
    return st;
}

// ------------------------------------------------------------------------------------------
void free_mesh(hdg_context* c) {
    mg_free(c);
    cg_release(c);
    auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    F(c->d_cellinfo); F(c->d_nodes); F(c->d_facecell); F(c->d_facenode); F(c->d_bfaces); F(c->d_isbc);
    F(c->d_kcol); F(c->d_fq); F(c->d_Kd); F(c->d_Ko); F(c->d_Ke); F(c->d_bcval);
    c->d_rhs = nullptr;   // lives behind d_Kd (one allocation, one memset per assembly)
    if (c->d_p) comm_unshare_vectors(c);   // close the neighbours' mappings before the vectors go away
    F(c->d_x); F(c->d_p); F(c->d_Ap);
    c->d_r = c->d_dinv = nullptr;          // r and Dinv live inside the d_p region
    F(c->d_binv);
    F(c->d_sigma); F(c->d_u); F(c->d_uhat_h); F(c->d_stage_cells); F(c->d_stage_faces);
    c->cap_ncell = c->cap_nnode = c->cap_nface = c->cap_nbface = 0;
    c->have_mesh = c->assembled = c->applied = c->solved = c->recovered = false;
}

// Device buffers are kept across hdg_set_mesh / hdg_set_rectangle_mesh calls with unchanged sizes
// (cudaMalloc / cudaFree of ~1 GB costs far more than re-uploading the mesh).
static bool same_capacity(const hdg_context* c, int64_t ncell, int64_t nnode, int64_t nface, int64_t nbface) {
    return c->d_cellinfo && c->d_Ke && c->cap_ncell == ncell && c->cap_nnode == nnode && c->cap_nface == nface &&
           c->cap_nbface >= nbface;
}

static hdg_status alloc_mesh(hdg_context* c) {
    if (c->ncell <= 0 || c->nface <= 0 || c->nnode <= 0) return set_err(c, HDG_ERR_INVALID, "empty mesh");
    if (c->nface >= (int64_t(1) << 31) || c->ncell >= (int64_t(1) << 31) || c->nnode >= (int64_t(1) << 31))
        return set_err(c, HDG_ERR_INVALID, "mesh too large for 32-bit device ids");
    bool same = same_capacity(c, c->ncell, c->nnode, c->nface, c->nbface);
    if (comm_active(c)) {
        // several GPUs: the decision is COLLECTIVE - a rank whose sizes did change frees the vector region its neighbours have
        // mapped and re-shares it in its next solve, so either every rank keeps its buffers or none does
        double flag = same ? 0.0 : 1.0;
        HDG_CUDA(c, cudaMemcpyAsync(c->comm->d_gscal, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        hdg_status st = comm_allreduce_sum(c, c->comm->d_gscal, 1);
        if (st) return st;
        HDG_CUDA(c, cudaMemcpyAsync(&flag, c->comm->d_gscal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        same = flag == 0.0;
    }
    if (same) {
        c->have_mesh = c->assembled = c->applied = c->solved = c->recovered = false;
        HDG_CUDA(c, cudaMemsetAsync(c->d_isbc, 0, c->nface, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_kcol, 0xFF, sizeof(int32_t) * 4 * c->nface, c->stream));
        return HDG_OK;
    }
    {
        int64_t a = c->ncell, b = c->nnode, d = c->nface, e = c->nbface, x = c->nx, y = c->ny, ao = c->ncell_own, fo = c->nface_own;
        free_mesh(c);
        c->ncell = a; c->nnode = b; c->nface = d; c->nbface = e; c->nx = x; c->ny = y; c->ncell_own = ao; c->nface_own = fo;
    }
    HDG_CUDA(c, cudaMalloc(&c->d_bfaces, sizeof(int32_t) * std::max<int64_t>(c->nbface, 1)));
    c->cap_ncell = c->ncell; c->cap_nnode = c->nnode; c->cap_nface = c->nface; c->cap_nbface = std::max<int64_t>(c->nbface, 1);
    HDG_CUDA(c, cudaMalloc(&c->d_cellinfo, sizeof(int32_t) * CI * c->ncell));
    HDG_CUDA(c, cudaMalloc(&c->d_nodes, sizeof(double) * 2 * c->nnode));
    HDG_CUDA(c, cudaMalloc(&c->d_facecell, sizeof(int32_t) * 2 * c->nface));
    HDG_CUDA(c, cudaMalloc(&c->d_facenode, sizeof(int32_t) * 2 * c->nface));
    HDG_CUDA(c, cudaMalloc(&c->d_isbc, c->nface));
    HDG_CUDA(c, cudaMalloc(&c->d_kcol, sizeof(int32_t) * 4 * c->nface));
    HDG_CUDA(c, cudaMemsetAsync(c->d_isbc, 0, c->nface, c->stream));
    HDG_CUDA(c, cudaMemsetAsync(c->d_kcol, 0xFF, sizeof(int32_t) * 4 * c->nface, c->stream));
    return HDG_OK;
}

hdg_status alloc_system(hdg_context* c) {
    const int nt = c->tab.nt, ke = c->tab.m * (c->tab.t + 1);
    if (c->d_Ke) return HDG_OK;   // buffers kept from a previous mesh of the same size (unused Ko slots are never read)
    int64_t ncell_pad = ceil_div(c->ncell, 32) * 32;
    // face-diagonal blocks and rhs share one allocation: both are zeroed before every assembly (RED targets)
    HDG_CUDA(c, cudaMalloc(&c->d_Kd, sizeof(double) * c->nface * (nt * nt + nt)));
    c->d_rhs = c->d_Kd + c->nface * nt * nt;
    HDG_CUDA(c, cudaMalloc(&c->d_Ko, sizeof(double) * c->nface * 4 * nt * nt));
    HDG_CUDA(c, cudaMalloc(&c->d_Ke, sizeof(double) * ncell_pad * ke));
    HDG_CUDA(c, cudaMemsetAsync(c->d_Ko, 0, sizeof(double) * c->nface * 4 * nt * nt, c->stream));
    return HDG_OK;
}

// ------------------------------------------------------------------------------------------
// hdg_set_mesh on several GPUs: every rank is handed the whole mesh (the Julia process of a GPU holds it) but UPLOADS ONLY ITS
// PART: a contiguous range of cell ids, the rows of the face table those cells create - with the reference's first-encounter
// numbering (src/generate_mesh.jl:20-46) they are a contiguous id range too, = the trace rows the rank owns - and the node
// coordinates.  Everything else happens on the device: the cells of other ranks that touch an owned face (recomputed as ghost
// cells, like the ghost layer of the strips) are found from the owned face rows, the faces local cells reference beyond the
// owned range become ghost columns whose p values the SpMV reads from the owner's memory, local numbering / adjacency / tile
// pairing / Dirichlet flags are built by kernels.  The host only sorts the two small interface lists (ghost cells, ghost faces)
// and gathers the ghost cells' rows; there is no pass over the mesh on the host (the partition boundaries come from R binary
// searches in the face table's first-cell column).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* __restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
// owned face rows: ids in range, first cells inside the owned cell range and ascending; ghost-cell candidates = second cells
// outside the owned range (appended in any order, duplicates possible)
__global__ void part_check_faces(const int64_t* __restrict__ fv1, const int64_t* __restrict__ fv2, const int64_t* __restrict__ fc1,
                                 const int64_t* __restrict__ fc2, int64_t nown, int64_t cb, int64_t ce, int64_t ncell, int64_t nnode,
                                 int64_t* __restrict__ cand, unsigned long long* __restrict__ ncand, int32_t* __restrict__ flag) {
    const int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nown) return;
    const int64_t c1 = fc1[f] - 1, c2 = fc2[f] - 1;
    if (fv1[f] < 1 || fv1[f] > nnode || fv2[f] < 1 || fv2[f] > nnode) atomicCAS(flag, 0, 3);
    if (c1 < cb || c1 >= ce || c2 < -1 || c2 >= ncell) atomicCAS(flag, 0, 4);
    if (f > 0 && fc1[f] < fc1[f - 1]) atomicCAS(flag, 0, 5);
    if (c2 >= 0 && (c2 < cb || c2 >= ce)) cand[atomicAdd(ncand, 1ull)] = c2;
}
// faces of the local cells (owned + ghost) beyond the owned face range (any order, duplicates possible); id range check
__global__ void part_ghost_face_candidates(const int64_t* __restrict__ cells, int64_t ncloc, int64_t fb, int64_t fe, int64_t nface,
                                           int64_t nnode, int64_t* __restrict__ cand, unsigned long long* __restrict__ ncand,
                                           int32_t* __restrict__ flag) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncloc) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t v = cells[6 * c + k], f = cells[6 * c + 3 + k] - 1;
        if (v < 1 || v > nnode) atomicCAS(flag, 0, 1);
        if (f < 0 || f >= nface) { atomicCAS(flag, 0, 2); continue; }
        if (f < fb || f >= fe) cand[atomicAdd(ncand, 1ull)] = f;
    }
}
// device records of the local cells: global node ids, LOCAL face ids (owned: f - fb, ghost: nown + position in the sorted ghost
// list), bit 31 = this cell is the face's second cell (owned face: the face row says so; ghost face of an owned cell: its first
// cell lives on another rank, so always)
__global__ void part_convert_cells(const int64_t* __restrict__ cells, int64_t ncloc, int64_t ncown, int64_t cb, const int64_t* __restrict__ gcells,
                                   int64_t fb, int64_t fe, const int64_t* __restrict__ fc2, const int64_t* __restrict__ gfaces, int64_t ngf,
                                   int32_t* __restrict__ cellinfo) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncloc) return;
    const int64_t cg = c < ncown ? cb + c : gcells[c - ncown];
#pragma unroll
    for (int k = 0; k < 3; ++k) cellinfo[CI * c + k] = int32_t(cells[6 * c + k] - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t f = cells[6 * c + 3 + k] - 1;
        int64_t fl;
        uint32_t sec;
        if (f >= fb && f < fe) { fl = f - fb; sec = (fc2[fl] - 1 == cg) ? 0x80000000u : 0u; }
        else { fl = (fe - fb) + lower_bound_i64(gfaces, ngf, f); sec = c < ncown ? 0x80000000u : 0u; }
        cellinfo[CI * c + 3 + k] = int32_t(uint32_t(fl) | sec);
    }
}
// face -> (first, second) LOCAL cell of the owned faces from their rows (-1: no second cell, -2: a cell this rank does not hold)
__global__ void part_owned_facecell(const int64_t* __restrict__ fv1, const int64_t* __restrict__ fv2, const int64_t* __restrict__ fc1,
                                    const int64_t* __restrict__ fc2, int64_t nown, int64_t cb, int64_t ce, int64_t ncown,
                                    const int64_t* __restrict__ gcells, int64_t ngc, int32_t* __restrict__ facecell, int32_t* __restrict__ facenode) {
    const int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nown) return;
    facenode[2 * f] = int32_t(fv1[f] - 1);
    facenode[2 * f + 1] = int32_t(fv2[f] - 1);
    facecell[2 * f] = int32_t(fc1[f] - 1 - cb);
    const int64_t c2 = fc2[f] - 1;
    int32_t l2 = -1;
    if (c2 >= 0) {
        if (c2 >= cb && c2 < ce) l2 = int32_t(c2 - cb);
        else {
            const int64_t p = lower_bound_i64(gcells, ngc, c2);
            l2 = (p < ngc && gcells[p] == c2) ? int32_t(ncown + p) : -2;
        }
    }
    facecell[2 * f + 1] = l2;
}
// ghost faces: the local cells that hold them (min / max local id; a single holder leaves the second entry at -1) and the end
// nodes from one holder's edge - enough for the tile pairing hints, these rows are never used as rows of K
__global__ void part_ghost_facecell(const int32_t* __restrict__ cellinfo, int64_t ncloc, int64_t nown, int32_t* __restrict__ facecell,
                                    int32_t* __restrict__ facenode) {
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncloc) return;
    const int k1[3] = {1, 2, 0}, k2[3] = {2, 0, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t f = uint32_t(cellinfo[CI * c + 3 + k]) & 0x7fffffffu;
        if (f < nown) continue;
        atomicMin(&facecell[2 * f], int32_t(c));
        atomicMax(&facecell[2 * f + 1], int32_t(c));
        facenode[2 * f] = cellinfo[CI * c + k1[k]];        // any holder's edge (both holders write the same two nodes, possibly swapped)
        facenode[2 * f + 1] = cellinfo[CI * c + k2[k]];
    }
}
// Dirichlet set (global 1-based ids, any order) -> flags of the local faces; an owned face with two cells raises the assertion
__global__ void part_mark_dirichlet(const int64_t* __restrict__ bfaces, int64_t nb, int64_t nface, int64_t fb, int64_t fe,
                                    const int64_t* __restrict__ fc2, const int64_t* __restrict__ gfaces, int64_t ngf, int32_t* __restrict__ dflag,
                                    int32_t* __restrict__ flags) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t f = bfaces[i] - 1;
    if (f < 0 || f >= nface) { atomicCAS(&flags[FLAG_BAD_ID], 0, 6); return; }
    if (f >= fb && f < fe) {
        if (fc2[f - fb] != 0) atomicExch(&flags[FLAG_NOT_BOUNDARY], int32_t(f + 1));      // src/boundary.jl:22
        dflag[f - fb] = 1;
    } else {
        const int64_t p = lower_bound_i64(gfaces, ngf, f);
        if (p < ngf && gfaces[p] == f) dflag[(fe - fb) + p] = 1;
    }
}

static hdg_status mesh_from_host_partitioned(hdg_context* c, const int64_t* cells, int64_t ncell, const double* nodes,
                                             int64_t nnode, const int64_t* faces, int64_t nface, const int64_t* bfaces,
                                             int64_t nbface) {
    Comm* m = c->comm;
    const int R = m->nranks, r = m->rank;
    if (!faces) return set_err(c, HDG_ERR_INVALID, "hdg_set_mesh on several GPUs needs the faces array");
    if (ncell < R) return set_err(c, HDG_ERR_INVALID, "need at least one cell per rank");
    const int64_t* fcol[4] = {faces, faces + nface, faces + 2 * nface, faces + 3 * nface};      // column-major nface x 4, 1-based
    // partition: equal cell-id ranges; the faces whose first cell lies in the range (the first-cell column is ascending with
    // first-encounter numbering - checked on the device for the rows a rank owns, which together are all rows)
    std::vector<int64_t> c0(R + 1), F0(R + 1);
    for (int q = 0; q <= R; ++q) c0[q] = ncell * q / R;
    for (int q = 0; q <= R; ++q) {
        int64_t lo = 0, hi = nface;
        while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (fcol[2][mid] - 1 < c0[q]) lo = mid + 1; else hi = mid; }
        F0[q] = lo;
    }
    const int64_t cb = c0[r], ce = c0[r + 1], fb = F0[r], fe = F0[r + 1];
    const int64_t ncown = ce - cb, nown = fe - fb;
    if (nown <= 0) return set_err(c, HDG_ERR_INVALID, "hdg_set_mesh on several GPUs needs first-encounter face numbering (faces ordered by their first cell)");
    const int B = 256;
    hdg_status st = HDG_OK;
    // ---- upload the owned face rows; ghost cells
    int64_t *d_f = nullptr, *d_cand = nullptr;
    unsigned long long* d_n = nullptr;
    HDG_CUDA(c, cudaMalloc(&d_f, sizeof(int64_t) * 4 * nown));
    HDG_CUDA(c, cudaMalloc(&d_cand, sizeof(int64_t) * std::max<int64_t>(nown, 1)));
    HDG_CUDA(c, cudaMalloc(&d_n, sizeof(unsigned long long)));
    auto cleanup = [&](hdg_status s2) { cudaStreamSynchronize(c->stream); cudaFree(d_f); cudaFree(d_cand); cudaFree(d_n); return s2; };
    for (int k = 0; k < 4; ++k)
        if (cudaMemcpyAsync(d_f + k * nown, fcol[k] + fb, sizeof(int64_t) * nown, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            return cleanup(set_err(c, HDG_ERR_CUDA, "cudaMemcpyAsync(faces)"));
    cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), c->stream);
    cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream);
    part_check_faces<<<(unsigned)ceil_div(nown, B), B, 0, c->stream>>>(d_f, d_f + nown, d_f + 2 * nown, d_f + 3 * nown, nown, cb, ce, ncell, nnode,
                                                                        d_cand, d_n, c->d_flags + FLAG_BAD_ID);
    c->launches += 1;
    unsigned long long ncand = 0;
    cudaMemcpyAsync(&ncand, d_n, sizeof(ncand), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return cleanup(set_err(c, HDG_ERR_CUDA, "partition: face check failed"));
    if (c->h_flags[FLAG_BAD_ID] == 5 || c->h_flags[FLAG_BAD_ID] == 4)
        return cleanup(set_err(c, HDG_ERR_INVALID, "hdg_set_mesh on several GPUs needs first-encounter face numbering (faces ordered by their first cell) and cell ids in range"));
    if (c->h_flags[FLAG_BAD_ID]) return cleanup(set_err(c, HDG_ERR_INVALID, "mesh arrays: node id of a face out of range (ids are 1-based)"));
    std::vector<int64_t> gcells(ncand);
    if (ncand) HDG_CUDA(c, cudaMemcpy(gcells.data(), d_cand, sizeof(int64_t) * ncand, cudaMemcpyDeviceToHost));
    std::sort(gcells.begin(), gcells.end());
    gcells.erase(std::unique(gcells.begin(), gcells.end()), gcells.end());
    const int64_t ngc = int64_t(gcells.size()), ncloc = ncown + ngc;
    // ---- local cell rows: the owned slice straight from the caller's array, the ghost rows gathered
    int64_t *d_cells = nullptr, *d_gcells = nullptr, *d_fcand = nullptr, *d_gfaces = nullptr, *d_bf = nullptr;
    int32_t* d_dflag = nullptr;
    int64_t* d_offs = nullptr;
    auto cleanup2 = [&](hdg_status s2) {
        cudaStreamSynchronize(c->stream);
        cudaFree(d_cells); cudaFree(d_gcells); cudaFree(d_fcand); cudaFree(d_gfaces); cudaFree(d_bf); cudaFree(d_dflag); cudaFree(d_offs);
        return cleanup(s2);
    };
    if (cudaMalloc(&d_cells, sizeof(int64_t) * 6 * ncloc) != cudaSuccess || cudaMalloc(&d_gcells, sizeof(int64_t) * std::max<int64_t>(ngc, 1)) != cudaSuccess ||
        cudaMalloc(&d_fcand, sizeof(int64_t) * 3 * ncloc) != cudaSuccess)
        return cleanup2(set_err(c, HDG_ERR_CUDA, "cudaMalloc(partition)"));
    cudaMemcpyAsync(d_cells, cells + 6 * cb, sizeof(int64_t) * 6 * ncown, cudaMemcpyHostToDevice, c->stream);
    std::vector<int64_t> grows(size_t(ngc) * 6);
    for (int64_t g = 0; g < ngc; ++g) std::copy(cells + 6 * gcells[g], cells + 6 * gcells[g] + 6, grows.begin() + 6 * g);
    if (ngc) {
        cudaMemcpyAsync(d_cells + 6 * ncown, grows.data(), sizeof(int64_t) * 6 * ngc, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_gcells, gcells.data(), sizeof(int64_t) * ngc, cudaMemcpyHostToDevice, c->stream);
    }
    cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), c->stream);
    part_ghost_face_candidates<<<(unsigned)ceil_div(ncloc, B), B, 0, c->stream>>>(d_cells, ncloc, fb, fe, nface, nnode, d_fcand, d_n, c->d_flags + FLAG_BAD_ID);
    c->launches += 1;
    cudaMemcpyAsync(&ncand, d_n, sizeof(ncand), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return cleanup2(set_err(c, HDG_ERR_CUDA, "partition: cell pass failed"));
    if (c->h_flags[FLAG_BAD_ID]) return cleanup2(set_err(c, HDG_ERR_INVALID, "mesh arrays: node / face id of a cell out of range (ids are 1-based: nodes 1..nnode, faces 1..nface)"));
    std::vector<int64_t> gfaces(ncand);
    if (ncand) HDG_CUDA(c, cudaMemcpy(gfaces.data(), d_fcand, sizeof(int64_t) * ncand, cudaMemcpyDeviceToHost));
    std::sort(gfaces.begin(), gfaces.end());
    gfaces.erase(std::unique(gfaces.begin(), gfaces.end()), gfaces.end());
    const int64_t ngf = int64_t(gfaces.size()), nfloc = nown + ngf;
    // ghost face -> (owner rank, local index there)
    std::vector<int32_t> ridx, owner;
    for (int64_t f : gfaces) {
        const int q = int(std::upper_bound(F0.begin(), F0.end(), f) - F0.begin()) - 1;
        owner.push_back(q);
        ridx.push_back(int32_t(f - F0[q]));
    }
    // ---- device mesh
    c->ncell = ncloc; c->ncell_own = ncown; c->nnode = nnode; c->nface = nfloc; c->nface_own = nown;
    c->nbface = std::min<int64_t>(nbface, nfloc); c->nx = c->ny = 0; c->grid_px = c->grid_py = 0;
    st = alloc_mesh(c);
    if (st) return cleanup2(st);
    if (cudaMalloc(&d_gfaces, sizeof(int64_t) * std::max<int64_t>(ngf, 1)) != cudaSuccess || cudaMalloc(&d_bf, sizeof(int64_t) * std::max<int64_t>(nbface, 1)) != cudaSuccess ||
        cudaMalloc(&d_dflag, sizeof(int32_t) * nfloc) != cudaSuccess || cudaMalloc(&d_offs, sizeof(int64_t) * (nfloc + 1)) != cudaSuccess)
        return cleanup2(set_err(c, HDG_ERR_CUDA, "cudaMalloc(partition)"));
    if (ngf) cudaMemcpyAsync(d_gfaces, gfaces.data(), sizeof(int64_t) * ngf, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->d_nodes, nodes, sizeof(double) * 2 * nnode, cudaMemcpyHostToDevice, c->stream);
    part_convert_cells<<<(unsigned)ceil_div(ncloc, B), B, 0, c->stream>>>(d_cells, ncloc, ncown, cb, d_gcells, fb, fe, d_f + 3 * nown, d_gfaces, ngf, c->d_cellinfo);
    part_owned_facecell<<<(unsigned)ceil_div(nown, B), B, 0, c->stream>>>(d_f, d_f + nown, d_f + 2 * nown, d_f + 3 * nown, nown, cb, ce, ncown, d_gcells, ngc,
                                                                           c->d_facecell, c->d_facenode);
    if (ngf) {
        init_facecell<<<(unsigned)ceil_div(ngf, B), B, 0, c->stream>>>(c->d_facecell + 2 * nown, ngf);
        part_ghost_facecell<<<(unsigned)ceil_div(ncloc, B), B, 0, c->stream>>>(c->d_cellinfo, ncloc, nown, c->d_facecell, c->d_facenode);
        finish_facecell<<<(unsigned)ceil_div(ngf, B), B, 0, c->stream>>>(c->d_facecell + 2 * nown, ngf);
    }
    build_kcol<<<(unsigned)ceil_div(ncloc, B), B, 0, c->stream>>>(c->d_cellinfo, ncloc, c->d_kcol);
    build_partner<<<(unsigned)ceil_div(ncloc, B), B, 0, c->stream>>>(c->d_cellinfo, ncloc, c->d_facecell);
    c->launches += 7;
    // Dirichlet set: flags of the local faces, compacted ascending (the dof order of Dirichlet, src/boundary.jl:19-21)
    cudaMemsetAsync(d_dflag, 0, sizeof(int32_t) * nfloc, c->stream);
    int64_t nbloc = 0;
    if (nbface) {
        cudaMemcpyAsync(d_bf, bfaces, sizeof(int64_t) * nbface, cudaMemcpyHostToDevice, c->stream);
        part_mark_dirichlet<<<(unsigned)ceil_div(nbface, B), B, 0, c->stream>>>(d_bf, nbface, nface, fb, fe, d_f + 3 * nown, d_gfaces, ngf, d_dflag, c->d_flags);
        c->launches += 1;
    }
    int64_t* d_tot = d_offs + nfloc;
    st = exclusive_scan(c, d_dflag, nfloc, d_offs, d_tot);
    if (st) return cleanup2(st);
    cudaMemcpyAsync(&nbloc, d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return cleanup2(set_err(c, HDG_ERR_CUDA, "partition: Dirichlet pass failed"));
    if (c->h_flags[FLAG_BAD_ID]) return cleanup2(set_err(c, HDG_ERR_INVALID, "boundary face id out of range"));
    if (c->h_flags[FLAG_NOT_BOUNDARY]) return cleanup2(set_err(c, HDG_ERR_NOT_BOUNDARY, "Face " + std::to_string(c->h_flags[FLAG_NOT_BOUNDARY]) + " is not in boundary"));
    compact_boundary<<<(unsigned)ceil_div(nfloc, B), B, 0, c->stream>>>(d_dflag, d_offs, nfloc, c->d_bfaces, c->d_isbc);
    c->launches += 1;
    c->nbface = nbloc;
    cleanup2(HDG_OK);
    m->j0 = m->j1 = m->ny_global = 0;
    m->cell_begin = cb; m->face_begin = fb; m->ncell_global = ncell; m->nface_global = nface;
    m->nbelow = m->nabove = 0;
    m->general_mesh = true;
    m->ghost_cells = gcells;
    comm_free_halo(c);
    st = comm_set_ghosts(c, ridx, owner);
    if (st) return st;
    st = alloc_system(c);
    if (st) return st;
    c->have_mesh = true;
    return HDG_OK;
}

// ---- is this the triangulation of rectangle_mesh?  (hdg_set_mesh arrays; enables the vertex-grid multigrid) ---------------
// rectangle_mesh numbers the nodes iy (nx+1) + ix and joins (ix,iy) to (ix+1,iy), (ix,iy+1) and (ix+1,iy) to (ix,iy+1)
// (src/generate_mesh.jl:1-18,101-143).  Node 0 is the lower-left corner: its only neighbours are node 1 and node px.
__global__ void grid_find_px(const int32_t* __restrict__ facenode, int64_t nface, int32_t* __restrict__ flags) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    const int32_t v1 = facenode[2 * f], v2 = facenode[2 * f + 1];
    if (min(v1, v2) == 0) atomicMax(&flags[FLAG_GRID_PX], max(v1, v2));
}
__global__ void grid_check(const int32_t* __restrict__ facenode, int64_t nface, int64_t nnode, int32_t* __restrict__ flags) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    const int32_t px = flags[FLAG_GRID_PX];
    if (px < 2 || nnode % px != 0) { if (f == 0) flags[FLAG_GRID_BAD] = 1; return; }
    const int32_t v1 = facenode[2 * f], v2 = facenode[2 * f + 1];
    const int32_t lo = min(v1, v2), hi = max(v1, v2);
    const int dx = hi % px - lo % px, dy = hi / px - lo / px;
    const bool ok = (dx == 1 && dy == 0) || (dx == 0 && dy == 1) || (dx == -1 && dy == 1);
    if (!ok) flags[FLAG_GRID_BAD] = 1;
}

hdg_status mesh_from_host(hdg_context* c, const int64_t* cells, int64_t ncell, const double* nodes, int64_t nnode,
                          const int64_t* faces, int64_t nface, const int64_t* bfaces, int64_t nbface) {
    mg_invalidate(c);   // vertex adjacency of the previous mesh (buffers are kept while the sizes fit)
    cg_release(c);      // the DofHandler of the previous mesh

    if (comm_active(c)) return mesh_from_host_partitioned(c, cells, ncell, nodes, nnode, faces, nface, bfaces, nbface);
    c->ncell = c->ncell_own = ncell; c->nnode = nnode; c->nface = c->nface_own = nface; c->nbface = nbface; c->nx = c->ny = 0;
    c->grid_px = c->grid_py = 0;
    hdg_status st = alloc_mesh(c);
    if (st) return st;
    if (!c->d_stage_cells) HDG_CUDA(c, cudaMalloc(&c->d_stage_cells, sizeof(int64_t) * 6 * ncell));
    if (faces && !c->d_stage_faces) HDG_CUDA(c, cudaMalloc(&c->d_stage_faces, sizeof(int64_t) * 4 * nface));
    int64_t *d_cells = c->d_stage_cells, *d_faces = c->d_stage_faces;
    HDG_CUDA(c, cudaMemcpyAsync(d_cells, cells, sizeof(int64_t) * 6 * ncell, cudaMemcpyHostToDevice, c->stream));
    if (faces) HDG_CUDA(c, cudaMemcpyAsync(d_faces, faces, sizeof(int64_t) * 4 * nface, cudaMemcpyHostToDevice, c->stream));
    HDG_CUDA(c, cudaMemcpyAsync(c->d_nodes, nodes, sizeof(double) * 2 * nnode, cudaMemcpyHostToDevice, c->stream));
    const int B = 256;
    {   // ids are used as device indices from here on: reject a malformed (e.g. 0-based) mesh first
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_BAD_ID, 0, sizeof(int32_t), c->stream));
        check_mesh_ids<<<(unsigned)ceil_div(std::max(ncell, nface), B), B, 0, c->stream>>>(d_cells, ncell, nnode, faces ? d_faces : nullptr, nface,
                                                                                         c->d_flags + FLAG_BAD_ID);
        c->launches += 1;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags + FLAG_BAD_ID, c->d_flags + FLAG_BAD_ID, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
        static const char* what[] = {"", "node id of a cell", "face id of a cell", "node id of a face", "cell id of a face"};
        const int bad = c->h_flags[FLAG_BAD_ID];
        if (bad) return set_err(c, HDG_ERR_INVALID, std::string("mesh arrays: ") + what[bad & 7] + " out of range (ids are 1-based: nodes 1..nnode, faces 1..nface, cells 1..ncell)");
    }
    if (faces) {
        convert_faces<<<(unsigned)ceil_div(nface, B), B, 0, c->stream>>>(d_faces, nface, c->d_facecell, c->d_facenode);
    } else {
        init_facecell<<<(unsigned)ceil_div(nface, B), B, 0, c->stream>>>(c->d_facecell, nface);
        derive_facecell<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(d_cells, ncell, c->d_facecell);
        derive_facenode<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(d_cells, ncell, c->d_facecell, c->d_facenode);
        finish_facecell<<<(unsigned)ceil_div(nface, B), B, 0, c->stream>>>(c->d_facecell, nface);
        c->launches += 4;
    }
    convert_cells<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(d_cells, ncell, c->d_facecell, c->d_cellinfo);
    build_kcol<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(c->d_cellinfo, ncell, c->d_kcol);
    build_partner<<<(unsigned)ceil_div(ncell, B), B, 0, c->stream>>>(c->d_cellinfo, ncell, c->d_facecell);
    c->launches += 4;
    // Dirichlet face set: sort ascending on the host (the reference iterates faces 1..nface and
    // tests membership, src/boundary.jl:19-21, so the dof order is ascending face id)
    std::vector<int32_t> bf(nbface);
    for (int64_t i = 0; i < nbface; ++i) {
        if (bfaces[i] < 1 || bfaces[i] > nface) return set_err(c, HDG_ERR_INVALID, "boundary face id out of range");
        bf[i] = int32_t(bfaces[i] - 1);
    }
    std::sort(bf.begin(), bf.end());
    bf.erase(std::unique(bf.begin(), bf.end()), bf.end());
    c->nbface = int64_t(bf.size());
    if (c->nbface) {
        HDG_CUDA(c, cudaMemcpyAsync(c->d_bfaces, bf.data(), sizeof(int32_t) * c->nbface, cudaMemcpyHostToDevice, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t) * NFLAGS, c->stream));
        mark_bfaces<<<(unsigned)ceil_div(c->nbface, B), B, 0, c->stream>>>(c->d_bfaces, c->nbface, c->d_facecell, c->d_isbc, c->d_flags);
        c->launches += 1;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int32_t) * NFLAGS, cudaMemcpyDeviceToHost, c->stream));
    }
    // rectangle_mesh triangulation passed as arrays?  (two words of the flag array; after the Dirichlet check's reset)
    HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_GRID_PX, 0, sizeof(int32_t) * 2, c->stream));
    grid_find_px<<<(unsigned)ceil_div(nface, B), B, 0, c->stream>>>(c->d_facenode, nface, c->d_flags);
    grid_check<<<(unsigned)ceil_div(nface, B), B, 0, c->stream>>>(c->d_facenode, nface, nnode, c->d_flags);
    c->launches += 2;
    HDG_CUDA(c, cudaMemcpyAsync(c->h_flags + FLAG_GRID_PX, c->d_flags + FLAG_GRID_PX, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        const int64_t px = c->h_flags[FLAG_GRID_PX], py = px >= 2 ? nnode / px : 0;
        if (!c->h_flags[FLAG_GRID_BAD] && px >= 2 && py >= 2 && px * py == nnode &&
            nface == 3 * (px - 1) * (py - 1) + (px - 1) + (py - 1) && ncell == 2 * (px - 1) * (py - 1)) {
            c->grid_px = px; c->grid_py = py;
        }
    }
    if (c->nbface && c->h_flags[FLAG_NOT_BOUNDARY])
        return set_err(c, HDG_ERR_NOT_BOUNDARY, "Face " + std::to_string(c->h_flags[FLAG_NOT_BOUNDARY]) + " is not in boundary");
    st = alloc_system(c);
    if (st) return st;
    c->have_mesh = true;
    return HDG_OK;
}

// Dirichlet(u_hat, mesh, faceset, g) on another face set than the one given with the mesh (src/boundary.jl:7-42: any named set
// of boundary faces - "bottom", "left", ... of rectangle_mesh, src/generate_mesh.jl:60-89).  Faces of the boundary that are not
// in the set keep their natural (flux) condition.
hdg_status mesh_set_dirichlet(hdg_context* c, const int64_t* bfaces, int64_t nbface) {
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "hdg_set_dirichlet_faces before a mesh is set");
    if (comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_set_dirichlet_faces is single-GPU");
    if (c->applied) return set_err(c, HDG_ERR_INVALID, "the system was already modified by hdg_apply_dirichlet: call hdg_assemble again first");
    std::vector<int32_t> bf(nbface);
    for (int64_t i = 0; i < nbface; ++i) {
        if (bfaces[i] < 1 || bfaces[i] > c->nface) return set_err(c, HDG_ERR_INVALID, "boundary face id out of range");
        bf[i] = int32_t(bfaces[i] - 1);
    }
    std::sort(bf.begin(), bf.end());
    bf.erase(std::unique(bf.begin(), bf.end()), bf.end());
    const int64_t nb = int64_t(bf.size());
    if (nb > c->cap_nbface) {
        cudaFree(c->d_bfaces);
        c->d_bfaces = nullptr;
        c->cap_nbface = 0;
        HDG_CUDA(c, cudaMalloc(&c->d_bfaces, sizeof(int32_t) * nb));
        c->cap_nbface = nb;
    }
    mg_invalidate(c);   // the vertex hierarchy fixes the vertices of Dirichlet faces
    const int B = 256;
    HDG_CUDA(c, cudaMemsetAsync(c->d_isbc, 0, c->nface, c->stream));
    c->nbface = 0;
    c->solved = false;
    if (nb) {
        HDG_CUDA(c, cudaMemcpyAsync(c->d_bfaces, bf.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, c->stream));
        HDG_CUDA(c, cudaMemsetAsync(c->d_flags + FLAG_NOT_BOUNDARY, 0, sizeof(int32_t), c->stream));
        mark_bfaces<<<(unsigned)ceil_div(nb, B), B, 0, c->stream>>>(c->d_bfaces, nb, c->d_facecell, c->d_isbc, c->d_flags);
        c->launches += 1;
        HDG_CUDA(c, cudaMemcpyAsync(c->h_flags + FLAG_NOT_BOUNDARY, c->d_flags + FLAG_NOT_BOUNDARY, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (nb && c->h_flags[FLAG_NOT_BOUNDARY]) {
        const int bad = c->h_flags[FLAG_NOT_BOUNDARY];
        cudaMemsetAsync(c->d_isbc, 0, c->nface, c->stream);     // leave an empty set behind, not a half-checked one
        cudaStreamSynchronize(c->stream);
        return set_err(c, HDG_ERR_NOT_BOUNDARY, "Face " + std::to_string(bad) + " is not in boundary");
    }
    c->nbface = nb;
    return HDG_OK;
}

hdg_status mesh_rectangle(hdg_context* c, int64_t nx, int64_t ny, double llx, double lly, double urx, double ury) {
    mg_invalidate(c);   // vertex adjacency of the previous mesh (buffers are kept while the sizes fit)
    cg_release(c);      // the DofHandler of the previous mesh

    if (nx < 1 || ny < 1 || !(urx > llx) || !(ury > lly)) return set_err(c, HDG_ERR_INVALID, "rectangle_mesh: need nx,ny >= 1 and UR > LL");
    // strip of quad rows owned by this rank (whole mesh on one GPU)
    int64_t j0 = 0, j1 = ny;
    const bool multi = comm_active(c);
    if (multi) {
        if (ny < c->comm->nranks) return set_err(c, HDG_ERR_INVALID, "need at least one quad row per rank");
        j0 = ny * c->comm->rank / c->comm->nranks;
        j1 = ny * (c->comm->rank + 1) / c->comm->nranks;
    }
    const int64_t nface_global = 3 * nx * ny + nx + ny;   // src/generate_mesh.jl:121
    Strip S{};
    S.nx = nx; S.ny = ny; S.j0 = j0; S.j1 = j1;
    S.F0 = quad_base(0, j0, nx);
    const int64_t F1 = j1 < ny ? quad_base(0, j1, nx) : nface_global;
    S.nown = F1 - S.F0;
    S.nbelow = j0 > 0 ? nx : 0;
    const int64_t nabove = j1 < ny ? 2 * nx : 0;
    S.ncell_own = 2 * nx * (j1 - j0);
    S.node0 = j0 * (nx + 1);
    const int64_t node_rows = (j1 - j0 + 1) + (j1 < ny ? 1 : 0);
    c->nx = nx; c->ny = ny;
    c->grid_px = nx + 1; c->grid_py = ny + 1;      // the global vertex grid, also on a strip
    c->ncell_own = S.ncell_own;
    c->ncell = S.ncell_own + (j1 < ny ? nx : 0);
    c->nnode = node_rows * (nx + 1);
    c->nface_own = S.nown;
    c->nface = S.nown + S.nbelow + nabove;
    // Dirichlet faces among the local faces: left/right of every owned row, global bottom/top, left face of the ghost row
    c->nbface = 2 * (j1 - j0) + (j0 == 0 ? nx : 0) + (j1 == ny ? nx : 0) + (j1 < ny ? 1 : 0);
    hdg_status st = alloc_mesh(c);
    if (st) return st;
    const int B = 256;
    rect_nodes<<<(unsigned)ceil_div(c->nnode, B), B, 0, c->stream>>>(c->d_nodes, nx + 1, ny + 1, j0, node_rows, llx, lly, urx, ury);
    const int64_t nquad_threads = nx * (j1 - j0 + (j1 < ny ? 1 : 0));
    rect_cells<<<(unsigned)ceil_div(nquad_threads, B), B, 0, c->stream>>>(S, c->d_cellinfo, c->d_facecell, c->d_facenode);
    build_kcol<<<(unsigned)ceil_div(c->ncell, B), B, 0, c->stream>>>(c->d_cellinfo, c->ncell, c->d_kcol);
    build_partner<<<(unsigned)ceil_div(c->ncell, B), B, 0, c->stream>>>(c->d_cellinfo, c->ncell, c->d_facecell);
    c->launches += 4;
    // boundary = faces with one cell (src/generate_mesh.jl:60-89), ascending
    int32_t* flag = nullptr;
    int64_t *offs = nullptr, *d_tot = nullptr;
    HDG_CUDA(c, cudaMalloc(&flag, sizeof(int32_t) * c->nface));
    HDG_CUDA(c, cudaMalloc(&offs, sizeof(int64_t) * c->nface));
    HDG_CUDA(c, cudaMalloc(&d_tot, sizeof(int64_t)));
    flag_boundary<<<(unsigned)ceil_div(c->nface, B), B, 0, c->stream>>>(c->d_facecell, c->nface, flag);
    c->launches += 1;
    st = exclusive_scan(c, flag, c->nface, offs, d_tot);
    if (st) return st;
    int64_t tot = 0;
    HDG_CUDA(c, cudaMemcpy(&tot, d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (tot != c->nbface) return set_err(c, HDG_ERR_INVALID, "internal: boundary face count mismatch");
    compact_boundary<<<(unsigned)ceil_div(c->nface, B), B, 0, c->stream>>>(flag, offs, c->nface, c->d_bfaces, c->d_isbc);
    c->launches += 1;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(flag); cudaFree(offs); cudaFree(d_tot);
    if (multi) {
        Comm* m = c->comm;
        m->j0 = j0; m->j1 = j1; m->ny_global = ny;
        m->cell_begin = 2 * nx * j0; m->face_begin = S.F0;
        m->ncell_global = 2 * nx * ny; m->nface_global = nface_global;
        m->nbelow = S.nbelow; m->nabove = nabove;
        // halo lists: what the neighbours hold as ghosts of my owned faces
        std::vector<int32_t> send_dn, send_up;
        if (j0 > 0)      // rank-1's ghost-above layer: [left(i,j0), diag(i,j0)] for every column
            for (int64_t i = 0; i < nx; ++i) {
                QuadFaces F = quad_faces(i, j0, nx);
                send_dn.push_back(int32_t(F.left - S.F0));
                send_dn.push_back(int32_t(F.diag - S.F0));
            }
        if (j1 < ny)     // rank+1's ghost-below layer: top(i, j1-1)
            for (int64_t i = 0; i < nx; ++i) send_up.push_back(int32_t(quad_faces(i, j1 - 1, nx).top - S.F0));
        st = comm_setup_halo(c, send_dn, send_up);
        if (st) return st;
        // ghost face -> local face index on the rank that owns it (peer-memory SpMV)
        std::vector<int32_t> ridx, owner;
        if (j0 > 0) {
            const int64_t jp = ny * (m->rank - 1) / m->nranks, F0p = quad_base(0, jp, nx);
            for (int64_t i = 0; i < nx; ++i) {
                ridx.push_back(int32_t(quad_faces(i, j0 - 1, nx).top - F0p));
                owner.push_back(m->rank - 1);
            }
        }
        if (j1 < ny) {
            const int64_t F0n = quad_base(0, j1, nx);
            for (int64_t i = 0; i < nx; ++i) {
                QuadFaces F = quad_faces(i, j1, nx);
                ridx.push_back(int32_t(F.left - F0n));
                ridx.push_back(int32_t(F.diag - F0n));
                owner.push_back(m->rank + 1);
                owner.push_back(m->rank + 1);
            }
        }
        m->general_mesh = false;
        m->ghost_cells.clear();      // the lower-left triangles of the quad row above the strip
        if (j1 < ny) for (int64_t i = 0; i < nx; ++i) m->ghost_cells.push_back(2 * (j1 * nx + i));
        st = comm_set_ghosts(c, ridx, owner);
        if (st) return st;
    }
    st = alloc_system(c);
    if (st) return st;
    c->have_mesh = true;
    return HDG_OK;
}

hdg_status mesh_perturb(hdg_context* c, double fraction, uint64_t seed) {
    if (!c->have_mesh || c->nx == 0 || comm_active(c)) return set_err(c, HDG_ERR_INVALID, "hdg_perturb_nodes needs a single-GPU rectangle mesh");
    if (!(fraction >= 0.0 && fraction < 0.5)) return set_err(c, HDG_ERR_INVALID, "perturbation fraction must be in [0,0.5)");
    double h_nodes[4];
    // mesh extents from the first and last node
    HDG_CUDA(c, cudaMemcpy(h_nodes, c->d_nodes, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    HDG_CUDA(c, cudaMemcpy(h_nodes + 2, c->d_nodes + 2 * (c->nnode - 1), 2 * sizeof(double), cudaMemcpyDeviceToHost));
    double hx = (h_nodes[2] - h_nodes[0]) / double(c->nx), hy = (h_nodes[3] - h_nodes[1]) / double(c->ny);
    perturb_nodes_k<<<(unsigned)ceil_div(c->nnode, 256), 256, 0, c->stream>>>(c->d_nodes, c->nx + 1, c->ny + 1, hx, hy, fraction, seed);
    c->launches += 1;
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->assembled = c->applied = c->solved = c->recovered = false;
    return HDG_OK;
}

// ------------------------------------------------------------------------------------------
// downloads in the Julia layouts
// ------------------------------------------------------------------------------------------
__global__ void export_cells(const int32_t* __restrict__ cellinfo, int64_t ncell, int64_t* __restrict__ cells) {
    int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) cells[6 * c + k] = int64_t(cellinfo[CI * c + k]) + 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) cells[6 * c + 3 + k] = int64_t(uint32_t(cellinfo[CI * c + 3 + k]) & 0x7fffffffu) + 1;
}
__global__ void export_faces(const int32_t* __restrict__ facecell, const int32_t* __restrict__ facenode, int64_t nface,
                             int64_t* __restrict__ faces) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    faces[f] = int64_t(facenode[2 * f]) + 1;
    faces[f + nface] = int64_t(facenode[2 * f + 1]) + 1;
    faces[f + 2 * nface] = int64_t(facecell[2 * f]) + 1;
    faces[f + 3 * nface] = int64_t(facecell[2 * f + 1]) + 1;
}
__global__ void export_i32_plus1(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = int64_t(in[i]) + 1;
}

hdg_status mesh_download(hdg_context* c, int64_t* cells, double* nodes, int64_t* faces, int64_t* bf) {
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "no mesh");
    const int B = 256;
    int64_t nmax = std::max({6 * c->ncell, 4 * c->nface, c->nbface});
    int64_t* tmp = nullptr;
    HDG_CUDA(c, cudaMalloc(&tmp, sizeof(int64_t) * nmax));
    if (cells) {
        export_cells<<<(unsigned)ceil_div(c->ncell, B), B, 0, c->stream>>>(c->d_cellinfo, c->ncell, tmp);
        HDG_CUDA(c, cudaMemcpyAsync(cells, tmp, sizeof(int64_t) * 6 * c->ncell, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (faces) {
        export_faces<<<(unsigned)ceil_div(c->nface, B), B, 0, c->stream>>>(c->d_facecell, c->d_facenode, c->nface, tmp);
        HDG_CUDA(c, cudaMemcpyAsync(faces, tmp, sizeof(int64_t) * 4 * c->nface, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (bf && c->nbface) {
        export_i32_plus1<<<(unsigned)ceil_div(c->nbface, B), B, 0, c->stream>>>(c->d_bfaces, c->nbface, tmp);
        HDG_CUDA(c, cudaMemcpyAsync(bf, tmp, sizeof(int64_t) * c->nbface, cudaMemcpyDeviceToHost, c->stream));
        HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    c->launches += 3;
    if (nodes) HDG_CUDA(c, cudaMemcpy(nodes, c->d_nodes, sizeof(double) * 2 * c->nnode, cudaMemcpyDeviceToHost));
    HDG_CUDA(c, cudaFree(tmp));
    return HDG_OK;
}

// ---- CSC pattern of sparse(I,J,V) ------------------------------------------------------------
// Column (f,b) holds rows (f',a) for every face f' of the cells adjacent to f, ascending.
__device__ inline int sorted_neighbours(const int32_t* __restrict__ kcol, int64_t f, int32_t nb[5]) {
    int cnt = 0;
    nb[cnt++] = int32_t(f);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int32_t g = kcol[4 * f + s];
        if (g < 0) continue;
        bool dup = false;
        for (int k = 0; k < cnt; ++k) dup |= nb[k] == g;
        if (!dup) nb[cnt++] = g;
    }
    for (int a = 1; a < cnt; ++a) {   // insertion sort, <= 5 entries
        int32_t v = nb[a];
        int b = a - 1;
        while (b >= 0 && nb[b] > v) { nb[b + 1] = nb[b]; --b; }
        nb[b + 1] = v;
    }
    return cnt;
}

__global__ void count_neighbours(const int32_t* __restrict__ kcol, int64_t nface, int32_t* __restrict__ cnt) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    int32_t nb[5];
    cnt[f] = sorted_neighbours(kcol, f, nb);
}

__global__ void fill_pattern(const int32_t* __restrict__ kcol, const int64_t* __restrict__ offs, int64_t nface, int nt,
                             int64_t* __restrict__ colptr, int64_t* __restrict__ rowval) {
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    int32_t nb[5];
    int cnt = sorted_neighbours(kcol, f, nb);
    int64_t base = int64_t(nt) * nt * offs[f];
    for (int b = 0; b < nt; ++b) {
        int64_t p = base + int64_t(b) * cnt * nt;
        colptr[f * nt + b] = p + 1;
        if (rowval)
            for (int k = 0; k < cnt; ++k)
                for (int a = 0; a < nt; ++a) rowval[p++] = int64_t(nb[k]) * nt + a + 1;
    }
    if (f == nface - 1) colptr[nface * nt] = base + int64_t(nt) * cnt * nt + 1;
}

__global__ void fill_values(const int32_t* __restrict__ kcol, const int64_t* __restrict__ offs, int64_t nface, int nt,
                            const double* __restrict__ Kd, const double* __restrict__ Ko, double* __restrict__ nzval) {
    // one thread per face column block; value of entry (row (f',a), col (f,b)) = block(f',f)(a,b)
    int64_t f = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (f >= nface) return;
    int32_t nb[5];
    int cnt = sorted_neighbours(kcol, f, nb);
    int64_t base = int64_t(nt) * nt * offs[f];
    const int nt2 = nt * nt;
    for (int b = 0; b < nt; ++b) {
        int64_t p = base + int64_t(b) * cnt * nt;
        for (int k = 0; k < cnt; ++k) {
            int64_t fr = nb[k];   // row face
            for (int a = 0; a < nt; ++a) {
                double v = 0.0;
                if (fr == f) v = Kd[f * nt2 + b * nt + a];
                else
                    for (int s = 0; s < 4; ++s)
                        if (kcol[4 * fr + s] == int32_t(f)) v += Ko[(fr * 4 + s) * nt2 + b * nt + a];
                nzval[p++] = v;
            }
        }
    }
}

struct PatternTmp {
    int32_t* cnt = nullptr;
    int64_t* offs = nullptr;
    int64_t* d_tot = nullptr;
    int64_t total = 0;
};
static hdg_status pattern_offsets(hdg_context* c, PatternTmp& P) {
    HDG_CUDA(c, cudaMalloc(&P.cnt, sizeof(int32_t) * c->nface));
    HDG_CUDA(c, cudaMalloc(&P.offs, sizeof(int64_t) * c->nface));
    HDG_CUDA(c, cudaMalloc(&P.d_tot, sizeof(int64_t)));
    count_neighbours<<<(unsigned)ceil_div(c->nface, 256), 256, 0, c->stream>>>(c->d_kcol, c->nface, P.cnt);
    c->launches += 1;
    hdg_status st = exclusive_scan(c, P.cnt, c->nface, P.offs, P.d_tot);
    if (st) return st;
    HDG_CUDA(c, cudaMemcpy(&P.total, P.d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost));
    return HDG_OK;
}
static void pattern_free(PatternTmp& P) {
    cudaFree(P.cnt); cudaFree(P.offs); cudaFree(P.d_tot);
}

int64_t pattern_nnz(hdg_context* c) {
    if (!c->have_mesh) return 0;
    PatternTmp P;
    if (pattern_offsets(c, P) != HDG_OK) { pattern_free(P); return -1; }
    pattern_free(P);
    return P.total * c->tab.nt * c->tab.nt;
}

hdg_status pattern_download(hdg_context* c, int64_t* colptr, int64_t* rowval) {
    if (!c->have_mesh) return set_err(c, HDG_ERR_INVALID, "no mesh");
    PatternTmp P;
    hdg_status st = pattern_offsets(c, P);
    if (st) { pattern_free(P); return st; }
    const int nt = c->tab.nt;
    int64_t ndof = c->nface * nt, nnz = P.total * nt * nt;
    int64_t *d_cp = nullptr, *d_rv = nullptr;
    HDG_CUDA(c, cudaMalloc(&d_cp, sizeof(int64_t) * (ndof + 1)));
    if (rowval) HDG_CUDA(c, cudaMalloc(&d_rv, sizeof(int64_t) * nnz));
    fill_pattern<<<(unsigned)ceil_div(c->nface, 256), 256, 0, c->stream>>>(c->d_kcol, P.offs, c->nface, nt, d_cp, d_rv);
    c->launches += 1;
    if (colptr) HDG_CUDA(c, cudaMemcpyAsync(colptr, d_cp, sizeof(int64_t) * (ndof + 1), cudaMemcpyDeviceToHost, c->stream));
    if (rowval) HDG_CUDA(c, cudaMemcpyAsync(rowval, d_rv, sizeof(int64_t) * nnz, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_cp);
    if (d_rv) cudaFree(d_rv);
    pattern_free(P);
    return HDG_OK;
}

hdg_status values_download(hdg_context* c, double* nzval) {
    if (!c->assembled) return set_err(c, HDG_ERR_INVALID, "hdg_get_values before hdg_assemble");
    PatternTmp P;
    hdg_status st = pattern_offsets(c, P);
    if (st) { pattern_free(P); return st; }
    const int nt = c->tab.nt;
    int64_t nnz = P.total * nt * nt;
    double* d_v = nullptr;
    HDG_CUDA(c, cudaMalloc(&d_v, sizeof(double) * nnz));
    fill_values<<<(unsigned)ceil_div(c->nface, 256), 256, 0, c->stream>>>(c->d_kcol, P.offs, c->nface, nt, c->d_Kd, c->d_Ko, d_v);
    c->launches += 1;
    HDG_CUDA(c, cudaMemcpyAsync(nzval, d_v, sizeof(double) * nnz, cudaMemcpyDeviceToHost, c->stream));
    HDG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_v);
    pattern_free(P);
    return HDG_OK;
}

}  // namespace hdg
