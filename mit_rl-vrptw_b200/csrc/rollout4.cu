// Translation unit of rollout4, split variant (gemm_path 6).
#include "launchers.h"
#include "rollout4_kernel.cuh"

namespace vrpb200 {

int smem_rollout4(int rb, int n_nodes, int max_smem, int* stages) {
    if (rb > 48) return 1 << 30;                      // TMEM: q^T and the score buffers need their own columns
    // ring depth: 4 stages hold the whole of W_q^T (FP16 hi/lo, 64 KB) - the query GEMM then never waits for weights
    *stages = make_layout4(rb, n_nodes, 4, 17).total <= max_smem ? 4 : make_layout4(rb, n_nodes, 3, 17).total <= max_smem ? 3 : 2;
    return make_layout4(rb, n_nodes, *stages, 17).total;
}

template <int RB, bool TW, int STAGES, bool SAMP>
static cudaError_t go2(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rollout4_kernel<RB, TW, STAGES, 16, true, SAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rollout4_kernel<RB, TW, STAGES, 16, true, SAMP><<<grid, kV4Threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int RB, bool TW, int STAGES>
static cudaError_t go(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    return a.mode == 1 ? go2<RB, TW, STAGES, true>(a, grid, smem, st) : go2<RB, TW, STAGES, false>(a, grid, smem, st);
}

template <int RB>
static cudaError_t go_rb(const RolloutArgs& a, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    if (stages == 4) return tw ? go<RB, true, 4>(a, grid, smem, st) : go<RB, false, 4>(a, grid, smem, st);
    if (stages == 3) return tw ? go<RB, true, 3>(a, grid, smem, st) : go<RB, false, 3>(a, grid, smem, st);
    return tw ? go<RB, true, 2>(a, grid, smem, st) : go<RB, false, 2>(a, grid, smem, st);
}

cudaError_t launch_rollout4(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    switch (rb) {
        case 16: return go_rb<16>(a, tw, stages, grid, smem, st);
        case 32: return go_rb<32>(a, tw, stages, grid, smem, st);
        case 48: return go_rb<48>(a, tw, stages, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vrpb200
