// EXPERIMENTAL - NOT compiled into libhdg_b200.so and NOT run on a GPU yet (round-2 candidate).
//
// The V-cycle of the P1-vertex multigrid preconditioner (hdg_mg.cu) as ONE persistent cooperative kernel: one CTA per SM
// (or a few), grid-wide barriers between the stages instead of kernel boundaries.  Today a V-cycle is ~30 launches of
// 3-5 us each (profiles/r1_launches_k1.csv) that move ~190 MB in total; with ~36 grid barriers of 1-2 us the same work
// should fit in 60-80 us.  The point-wise operations and their order are those of mg_smooth0 / mg_residual / mg_restrict /
// mg_prolong_add / mg_smooth, so the result is bitwise the same.
//
// To try it: include this header at the end of hdg_mg.cu (it uses MG_OMEGA, mg_apply_row_rw, mg_restrict_pt, mg_prolong_pt),
// fill MgAllLevels in mg_apply_t from m->lev[0 .. nlev-1], and replace the per-level launches and mg_fused_vcycle by
//     void* args[] = {&A};
//     cudaLaunchCooperativeKernel((void*)mg_vcycle_persistent, grid, 256, args, 0, stream);
// with grid = (SM count) x cudaOccupancyMaxActiveBlocksPerMultiprocessor(mg_vcycle_persistent, 256, 0) (cooperative launches
// can be captured into the CUDA graph of the PCG chunk).  Syntax-checked with
//     nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -c tools/check_experimental.cu
#pragma once
#include <cooperative_groups.h>

namespace hdg {

constexpr int MG_ALL_LEVELS = 24;

struct MgAllLevels {
    int nl;                                        // all levels; lev nl-1 is the dense one
    int px[MG_ALL_LEVELS], py[MG_ALL_LEVELS];
    double *st[MG_ALL_LEVELS], *dinv[MG_ALL_LEVELS], *r[MG_ALL_LEVELS], *x[MG_ALL_LEVELS], *t[MG_ALL_LEVELS];
    const double* ainv;
};

__global__ void __launch_bounds__(256) mg_vcycle_persistent(const MgAllLevels A) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int64_t T = int64_t(gridDim.x) * blockDim.x, tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int l = 0; l + 1 < A.nl; ++l) {
        const int px = A.px[l], py = A.py[l], cx = A.px[l + 1], cy = A.py[l + 1];
        const int64_t n = int64_t(px) * py, nc = int64_t(cx) * cy;
        for (int64_t p = tid; p < n; p += T) A.x[l][p] = MG_OMEGA * A.dinv[l][p] * A.r[l][p];
        grid.sync();
        for (int64_t p = tid; p < n; p += T) A.t[l][p] = A.r[l][p] - mg_apply_row_rw(A.st[l], A.x[l], p, px, py, n);
        grid.sync();
        for (int64_t I = tid; I < nc; I += T) A.r[l + 1][I] = mg_restrict_pt(A.t[l], px, py, A.dinv[l + 1], I, cx);
        grid.sync();
    }
    {
        const int l = A.nl - 1, n = A.px[l] * A.py[l];
        for (int64_t i = tid; i < n; i += T) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(A.ainv[i * n + j], A.r[l][j], s);
            A.t[l][i] = s;
        }
        grid.sync();
    }
    for (int l = A.nl - 2; l >= 0; --l) {
        const int px = A.px[l], py = A.py[l], cx = A.px[l + 1], cy = A.py[l + 1];
        const int64_t n = int64_t(px) * py;
        for (int64_t p = tid; p < n; p += T)
            if (A.dinv[l][p] != 0.0) A.x[l][p] += mg_prolong_pt(A.t[l + 1], cx, cy, p, px);
        grid.sync();
        for (int64_t p = tid; p < n; p += T)
            A.t[l][p] = fma(MG_OMEGA * A.dinv[l][p], A.r[l][p] - mg_apply_row_rw(A.st[l], A.x[l], p, px, py, n), A.x[l][p]);
        if (l > 0) grid.sync();
    }
}

}  // namespace hdg
