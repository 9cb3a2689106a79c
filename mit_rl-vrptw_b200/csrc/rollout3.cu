// Translation unit of rollout3 (gemm_path 3): transposed tcgen05 GEMMs, lock-step phases, fused beam search.
#include "launchers.h"
#include "rollout3_kernel.cuh"

namespace vrpb200 {

int smem_rollout3(int rb, int ib, int n_nodes, int fd, int W, int max_smem, int* stages) {
    if (rb < W) return 1 << 30;                       // every beam of an instance lives in one CTA
    const int np = (n_nodes + 3) / 4 * 4;
    if (W > 1 && 1024 + rb * (np * 4 + np + kRowState * 4) > 8 * kC1TileBytes) return 1 << 30;   // beam gather scratch
    *stages = make_layout3(rb, ib, n_nodes, fd, 3).total <= max_smem ? 3 : 2;
    return make_layout3(rb, ib, n_nodes, fd, *stages).total;
}

template <int RB, bool TW, int STAGES>
static cudaError_t go(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rollout3_kernel<RB, TW, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rollout3_kernel<RB, TW, STAGES><<<grid, kV3Threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int RB>
static cudaError_t go_rb(const RolloutArgs& a, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    if (stages == 3) return tw ? go<RB, true, 3>(a, grid, smem, st) : go<RB, false, 3>(a, grid, smem, st);
    return tw ? go<RB, true, 2>(a, grid, smem, st) : go<RB, false, 2>(a, grid, smem, st);
}

cudaError_t launch_rollout3(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    switch (rb) {
        case 16: return go_rb<16>(a, tw, stages, grid, smem, st);
        case 32: return go_rb<32>(a, tw, stages, grid, smem, st);
        case 48: return go_rb<48>(a, tw, stages, grid, smem, st);
        case 64: return go_rb<64>(a, tw, stages, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vrpb200
