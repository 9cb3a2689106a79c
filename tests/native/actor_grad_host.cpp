// TEST INFRASTRUCTURE: builds mit_rl-vrptw_b200/csrc/actor_grad_row.h as plain C++ (a phase = a loop over the 128 threads), so
// that the algebra of the CUDA back-propagation kernel can be checked against torch autograd on a machine without a GPU
// (tests/test_actor_grad_cpu.py).  Nothing under mit_rl-vrptw_b200/ loads this; the product runs the CUDA build only.
#include <vector>
#include <cstring>
#include "../../mit_rl-vrptw_b200/csrc/actor_grad_row.h"

using namespace vrpb200;

extern "C" int actor_grad_host(const float* const* wp /*21 pointers in ActorGradWeights order*/, int N, int E, int D, int T, int tw,
                               int use_tanh, int mask_pointer, float C, float forget_bias, long long R, const float* input,
                               const int* idx /*[T,R]*/, const float* weight, const float* demand /*[T,R,N]*/,
                               const float* mask, const float* load /*[T,R]*/, const float* time,
                               float* dgates /*[R,T,4H]*/, float* dq /*[R,T,H]*/, float* Dsum /*[R,N,H]*/, float* demb /*[R,N,E]*/,
                               float* sums /*[R,5,H]*/, float* hs /*[R,T+1,H]*/, float* xin /*[R,T,E]*/, float* emb /*[R,N,E]*/,
                               float* const* small /*15 pointers in SmallGradOut order*/) {
    ActorGradWeights w;
    static_assert(sizeof(ActorGradWeights) == 21 * sizeof(void*), "layout");
    std::memcpy(&w, wp, sizeof w);
    ActorGradDims d{N, E, D, T, tw, use_tanh, mask_pointer, C, forget_bias};
    const int H = kAgH;
    std::vector<float> cs((size_t)(T + 1) * H), act((size_t)T * 4 * H), dx((size_t)T * E), eref((size_t)N * H), erefT((size_t)H * 128);
    for (long long r = 0; r < R; ++r) {
        ActorGradRow row;
        row.input = input + r * N * D;
        row.idx = idx + r; row.idx_stride = R; row.weight = weight[r];
        row.demand = demand + r * N; row.mask = mask + r * N; row.node_stride = R * N;
        row.load = load + r; row.time = time ? time + r : nullptr; row.scalar_stride = R;
        row.hs = hs + r * (T + 1) * H; row.cs = cs.data(); row.act = act.data(); row.xin = xin + r * T * E; row.dx = dx.data();
        row.emb = emb + r * N * E; row.eref = eref.data(); row.erefT = erefT.data();
        row.dgates = dgates + r * T * 4 * H; row.dq = dq + r * T * H; row.Dsum = Dsum + r * N * H; row.demb = demb + r * N * E;
        row.sums = sums + r * 5 * H;
        actor_grad_row(w, d, row);
    }
    SmallGradOut o;
    static_assert(sizeof(SmallGradOut) == 15 * sizeof(void*), "layout");
    std::memcpy(&o, small, sizeof o);
    attention_small_grads(w, sums, R, tw, o);
    return 0;
}
