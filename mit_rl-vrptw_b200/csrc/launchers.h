// Internal: one launch function per fused-kernel family; each is defined in its own translation unit (rollout_*.cu) so
// that the families compile in parallel and the host logic (vrpb200_api.cu) instantiates no kernel template itself.
#pragma once
#include "common.cuh"

namespace vrpb200 {

// shared-memory bytes a family needs for `rb` rows per CTA (1 << 30: not available); stages is set for the ring depth
int smem_ffma(int rb, int ib, int n_nodes, int fd);
int smem_rollout3(int rb, int ib, int n_nodes, int fd, int W, int max_smem, int* stages);
int smem_rollout4(int rb, int n_nodes, int max_smem, int* stages);

cudaError_t launch_ffma(const RolloutArgs& a, int rb, bool tw, int grid, int smem, cudaStream_t st);
cudaError_t launch_rollout3(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st);
cudaError_t launch_rollout4(const RolloutArgs& a, int rb, bool tw, int stages, int grid, int smem, cudaStream_t st);

}  // namespace vrpb200
