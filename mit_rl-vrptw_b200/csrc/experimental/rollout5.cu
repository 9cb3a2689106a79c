// Translation unit of rollout5 (gemm_path 7): the warp-specialised pipeline (score warps | aux warps | service warps).
// EXPERIMENT, not part of the build since the FP16-split GEMMs made rollout4 as fast or faster at every block size (DESIGN.md
// section 4.3, profiles/r2_rows_per_cta_sweep.txt).  Bit-identical to rollout4 on the whole GPU suite when it was retired.
// To build it again: move both files back to csrc/, declare smem_rollout5 / launch_rollout5 in launchers.h, add family 7 to
// launch_rollout (vrpb200_api.cu) and "tc5": 7 to engine.GEMM_PATHS.
#include "../launchers.h"
#include "rollout5_kernel.cuh"

namespace vrpb200 {

int smem_rollout5(int rb, int n_nodes, int max_smem, int* stages) {
    if (rb > 48) return 1 << 30;                      // TMEM: gates^T, q^T, the cell state and the score buffers share 512 columns
    *stages = make_layout5(rb, n_nodes, 4).total <= max_smem ? 4 : make_layout5(rb, n_nodes, 3).total <= max_smem ? 3 : 2;   // 4 stages = all of W_q^T
    return make_layout5(rb, n_nodes, *stages).total;
}

template <int RB, bool TW, int STAGES>
static cudaError_t go(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rollout5_kernel<RB, TW, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rollout5_kernel<RB, TW, STAGES><<<grid, kV5Threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int RB>
static cudaError_t go_rb(const RolloutArgs& a, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    if (stages == 4) return tw ? go<RB, true, 4>(a, grid, smem, st) : go<RB, false, 4>(a, grid, smem, st);
    if (stages == 3) return tw ? go<RB, true, 3>(a, grid, smem, st) : go<RB, false, 3>(a, grid, smem, st);
    return tw ? go<RB, true, 2>(a, grid, smem, st) : go<RB, false, 2>(a, grid, smem, st);
}

cudaError_t launch_rollout5(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    switch (rb) {
        case 16: return go_rb<16>(a, tw, stages, grid, smem, st);
        case 32: return go_rb<32>(a, tw, stages, grid, smem, st);
        case 48: return go_rb<48>(a, tw, stages, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vrpb200
