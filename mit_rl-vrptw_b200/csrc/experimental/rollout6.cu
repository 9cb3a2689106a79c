// Translation unit of rollout6: rollout4 with the rows of a CTA split into two phase-shifted teams.
// EXPERIMENT, not part of the build (csrc/experimental/ is not compiled into libvrpb200.so): bit-identical to rollout4 on
// the whole GPU suite (45 tests) but 3.5 % SLOWER on the headline workload (13.64 ms vs 13.17 ms per 14 208 UPS VRPTW-100
// instances; 13.38 ms without the score token) - profiles/r2_phase_clocks_rollout6_two_teams.txt and DESIGN.md section 4.4
// say why.  To try it again: move both files back to csrc/, declare smem_rollout6 / launch_rollout6 in launchers.h and add
// a family to launch_rollout (vrpb200_api.cu).
#include "../launchers.h"
#include "rollout6_kernel.cuh"

namespace vrpb200 {

int smem_rollout6(int rb, int n_nodes, int max_smem, int* stages) {
    if (rb != 32 && rb != 48) return 1 << 30;         // two teams of 16 or 24 rows
    *stages = make_layout6(rb, n_nodes, 3).total <= max_smem ? 3 : 2;
    return make_layout6(rb, n_nodes, *stages).total;
}

template <int RB, bool TW, int STAGES>
static cudaError_t go(const RolloutArgs& a, int grid, int smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rollout6_kernel<RB, TW, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rollout6_kernel<RB, TW, STAGES><<<grid, kV6Threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int RB>
static cudaError_t go_rb(const RolloutArgs& a, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    if (stages == 3) return tw ? go<RB, true, 3>(a, grid, smem, st) : go<RB, false, 3>(a, grid, smem, st);
    return tw ? go<RB, true, 2>(a, grid, smem, st) : go<RB, false, 2>(a, grid, smem, st);
}

cudaError_t launch_rollout6(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st) {
    switch (rb) {
        case 32: return go_rb<32>(a, tw, stages, grid, smem, st);
        case 48: return go_rb<48>(a, tw, stages, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vrpb200
