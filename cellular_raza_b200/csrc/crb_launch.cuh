// crb_launch.cuh — launchers of the tile force kernels, one translation unit per kernel family so that the library
// compiles in parallel (crb_tile.cu, crb_tile2.cu, crb_tile3.cu).  Each returns the launch status
// (cudaFuncSetAttribute for the dynamic shared memory + the launch itself).
#pragma once
#include "crb_force_tile.cuh"

namespace crb {

// grid == kPreloadOnly: nothing is launched, the kernel is only loaded (CUDA loads kernels lazily at their first
// launch, and that load waits for the device to go idle: a first launch queued behind a k_slab_wait that spins for a
// neighbouring slab on the same device would never get there).  crb_slab_step preloads what it launches.
constexpr u32 kPreloadOnly = 0xffffffffu;

cudaError_t launch_k_force_tile(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                int threads, size_t smem, cudaStream_t st);
cudaError_t launch_k_force_tile2(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                 int threads, size_t smem, cudaStream_t st);
cudaError_t launch_k_force_tile3(int D, int inter, bool force_only, const KParams& P, const TileLaunch& T, u32 grid,
                                 int threads, size_t smem, cudaStream_t st);

}  // namespace crb
