// Syntax check of the round-2 candidates under csrc/experimental/ (they are not part of libhdg_b200.so):
//   nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -std=c++17 -I hdiscontinuousgalerkin.jl_b200/csrc -c tools/check_experimental.cu -o /dev/null
#include <cstdint>
namespace hdg {
constexpr double MG_OMEGA = 0.8;
__device__ __constant__ int MG_DX[7] = {0, 1, -1, 0, 0, 1, -1};
__device__ __constant__ int MG_DY[7] = {0, 0, 0, 1, -1, -1, 1};
__device__ __forceinline__ double mg_apply_row_rw(const double* st, const double* x, int64_t p, int px, int py, int64_t n) {
    const int ix = int(p % px), iy = int(p / px);
    double s = st[p] * x[p];
    for (int k = 1; k < 7; ++k) {
        const int qx = ix + MG_DX[k], qy = iy + MG_DY[k];
        if (qx < 0 || qy < 0 || qx >= px || qy >= py) continue;
        s = fma(st[k * n + p], x[int64_t(qy) * px + qx], s);
    }
    return s;
}
__device__ __forceinline__ double mg_restrict_pt(const double* tf, int px, int py, const double* dinvc, int64_t I, int cx) {
    double s = 0.0;
    if (dinvc[I] != 0.0) {
        const int Ix = int(I % cx), Iy = int(I / cx);
        for (int d = 0; d < 7; ++d) {
            const int fx = 2 * Ix + MG_DX[d], fy = 2 * Iy + MG_DY[d];
            if (fx < 0 || fy < 0 || fx >= px || fy >= py) continue;
            s += (d == 0 ? 1.0 : 0.5) * tf[int64_t(fy) * px + fx];
        }
    }
    return s;
}
__device__ __forceinline__ double mg_prolong_pt(const double* ec, int cx, int cy, int64_t p, int px) {
    const int ix = int(p % px), iy = int(p / px);
    const int a2 = ix & 1, b2 = iy & 1, hx = ix >> 1, hy = iy >> 1;
    auto get = [&](int jx, int jy) { return (jx < cx && jy < cy) ? ec[int64_t(jy) * cx + jx] : 0.0; };
    if (!a2 && !b2) return get(hx, hy);
    if (a2 && !b2) return 0.5 * (get(hx, hy) + get(hx + 1, hy));
    if (!a2 && b2) return 0.5 * (get(hx, hy) + get(hx, hy + 1));
    return 0.5 * (get(hx + 1, hy) + get(hx, hy + 1));
}
}  // namespace hdg
#include "experimental/mg_persistent.cuh"
#include "experimental/mg_general.cuh"
