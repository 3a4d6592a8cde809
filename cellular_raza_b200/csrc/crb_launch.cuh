// crb_launch.cuh — launchers of the tile force kernels, one translation unit per kernel family so that the library
// compiles in parallel (crb_tile.cu, crb_tile2.cu, crb_tile3.cu).  Each returns the launch status
// (cudaFuncSetAttribute for the dynamic shared memory + the launch itself).
#pragma once
#include "crb_force_tile.cuh"

namespace crb {

cudaError_t launch_k_force_tile(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                int threads, size_t smem, cudaStream_t st);
cudaError_t launch_k_force_tile2(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                 int threads, size_t smem, cudaStream_t st);
cudaError_t launch_k_force_tile3(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                 int threads, size_t smem, cudaStream_t st);

}  // namespace crb
