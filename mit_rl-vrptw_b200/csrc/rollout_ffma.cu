// Translation unit of the FP32-pipe cross-check kernel (gemm_path 1).
#include "launchers.h"
#include "rollout_ffma.cuh"

namespace vrpb200 {

int smem_ffma(int rb, int ib, int n_nodes, int fd) { return make_layout(rb, ib, n_nodes, fd).total; }

template <int RB, bool TW>
static cudaError_t go(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rollout_ffma_kernel<RB, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rollout_ffma_kernel<RB, TW><<<grid, RB * 8, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_ffma(const RolloutArgs& a, int rb, bool tw, int grid, int smem, cudaStream_t st) {
    switch (rb) {
        case 16: return tw ? go<16, true>(a, grid, smem, st) : go<16, false>(a, grid, smem, st);
        case 32: return tw ? go<32, true>(a, grid, smem, st) : go<32, false>(a, grid, smem, st);
        case 48: return tw ? go<48, true>(a, grid, smem, st) : go<48, false>(a, grid, smem, st);
        case 64: return tw ? go<64, true>(a, grid, smem, st) : go<64, false>(a, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vrpb200
