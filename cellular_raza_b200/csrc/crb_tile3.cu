// crb_tile3.cu — translation unit of k_force_tile3 (see crb_launch.cuh)
#define CRB_TILE_TU
#include "crb_force_tile3.cuh"
#include "crb_launch.cuh"

namespace crb {
namespace {

template <class K>
cudaError_t go(K k, const KParams& P, const TileLaunch& T, u32 grid, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, threads, smem, st>>>(P, T);
    return cudaGetLastError();
}

template <int D, int INTER>
cudaError_t pick(bool force_only, const KParams& P, const TileLaunch& T, u32 grid, int threads, size_t smem,
                 cudaStream_t st) {
    if (threads == 256)
        return force_only ? go(k_force_tile3<D, INTER, true, 256>, P, T, grid, threads, smem, st)
                          : go(k_force_tile3<D, INTER, false, 256>, P, T, grid, threads, smem, st);
    return force_only ? go(k_force_tile3<D, INTER, true, 128>, P, T, grid, threads, smem, st)
                      : go(k_force_tile3<D, INTER, false, 128>, P, T, grid, threads, smem, st);
}

template <int D>
cudaError_t by_inter(int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid, int threads,
                     size_t smem, cudaStream_t st) {
    switch (inter) {
        case CRB_INT_CELL_SORTING: return pick<D, CRB_INT_CELL_SORTING>(force_only, P, T, grid, threads, smem, st);
        case CRB_INT_CELL_SORTING_BENCH: return pick<D, CRB_INT_CELL_SORTING_BENCH>(force_only, P, T, grid, threads, smem, st);
        case CRB_INT_MORSE: return pick<D, CRB_INT_MORSE>(force_only, P, T, grid, threads, smem, st);
        case CRB_INT_BOUND_LJ: return pick<D, CRB_INT_BOUND_LJ>(force_only, P, T, grid, threads, smem, st);
        case CRB_INT_MIE: return pick<D, CRB_INT_MIE>(force_only, P, T, grid, threads, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_k_force_tile3(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid, int threads,
                                 size_t smem, cudaStream_t st) {
    return D == 3 ? by_inter<3>(inter, force_only, P, T, grid, threads, smem, st)
                  : by_inter<2>(inter, force_only, P, T, grid, threads, smem, st);
}

}  // namespace crb
