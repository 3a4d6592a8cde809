// crb_tile2.cu — translation unit of k_force_tile2 (see crb_launch.cuh)
#define CRB_TILE_TU
#include "crb_force_tile2.cuh"
#include "crb_launch.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace crb {
namespace {

// cudaFuncSetAttribute is issued only when a kernel needs more dynamic shared memory than it was configured for
// (per device): the call may wait for work in flight, which must not happen between the queued phases of
// crb_slab_step (a device-side wait for a neighbouring slab may be pending in the stream)
template <class K>
cudaError_t go(K k, const KParams& P, const TileLaunch& T, u32 grid, int threads, size_t smem, cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> configured;
    int dev = 0;
    cudaGetDevice(&dev);
    if (grid == kPreloadOnly) {  // force the lazy loading of this kernel now (crb_slab_step: see crb_launch.cuh)
        cudaFuncAttributes at;
        return cudaFuncGetAttributes(&at, k);
    }
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& have = configured[{reinterpret_cast<const void*>(k), dev}];
        if (have < smem) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            have = smem;
        }
    }
    // grid (segments x chunks, tile rows, tile planes): no integer division by run-time values in the kernel's
    // prologue except the segment / chunk split, which is one multiply-high by inv_nseg
    TileLaunch U = T;
    U.inv_nseg = (u32)((0x100000000ull + (u64)T.nseg - 1) / (u64)T.nseg);
    const u32 gx = (u32)T.nseg * (u32)T.nchunks, gy = (u32)(P.g.ng[1] + 1) / 2;
    const u32 gz = grid / (gx * gy);
    if (gx >= 65536u || gy >= 65536u || gz >= 65536u) return cudaErrorInvalidConfiguration;
    k<<<dim3(gx, gy, gz), threads, smem, st>>>(P, U);
    return cudaGetLastError();
}

template <int D, int INTER>
cudaError_t pick(bool force_only, const KParams& P, const TileLaunch& T, u32 grid, int threads, size_t smem,
                 cudaStream_t st) {
    if (threads == 256)
        return force_only ? go(k_force_tile2<D, INTER, true, 256>, P, T, grid, threads, smem, st)
                          : go(k_force_tile2<D, INTER, false, 256>, P, T, grid, threads, smem, st);
    return force_only ? go(k_force_tile2<D, INTER, true, 128>, P, T, grid, threads, smem, st)
                      : go(k_force_tile2<D, INTER, false, 128>, P, T, grid, threads, smem, st);
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

cudaError_t launch_k_force_tile2(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid, int threads,
                                 size_t smem, cudaStream_t st) {
    return D == 3 ? by_inter<3>(inter, force_only, P, T, grid, threads, smem, st)
                  : by_inter<2>(inter, force_only, P, T, grid, threads, smem, st);
}

}  // namespace crb
