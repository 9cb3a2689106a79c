// TEST INFRASTRUCTURE: builds mit_rl-vrptw_b200/csrc/actor_grad_row.h as plain C++ (a phase = a loop over the 128 threads), so
// that the algebra of the CUDA back-propagation kernel can be checked against torch autograd on a machine without a GPU
// (tests/test_actor_grad_cpu.py).  Nothing under mit_rl-vrptw_b200/ loads this; the product runs the CUDA build only.
#include <vector>
#include <cstring>
#include "../../mit_rl-vrptw_b200/csrc/actor_grad_row.h"

using namespace vrpb200;

extern "C" int actor_grad_host(const float* const* wp /*23 pointers in ActorGradWeights order*/, int N, int E, int D, int T, int tw,
                               int use_tanh, int mask_pointer, float C, float forget_bias, long long R, const float* input,
                               const int* idx /*[T,R]*/, const float* weight, const float* demand /*[T,R,N]*/,
                               const float* mask, const float* load /*[T,R]*/, const float* time,
                               float* dgates /*[R,T,4H]*/, float* dq /*[R,T,H]*/, float* Dsum /*[R,N,H]*/, float* demb /*[R,N,E]*/,
                               float* sums /*[R,5,H]*/, float* hs /*[R,T+1,H]*/, float* xin /*[R,T,E]*/, float* emb /*[R,N,E]*/,
                               float* const* small /*15 pointers in SmallGradOut order*/, const float* xscale /*[T,R,E] or null*/) {
    ActorGradWeights w;
    static_assert(sizeof(ActorGradWeights) == 23 * sizeof(void*), "layout");
    std::memcpy(&w, wp, sizeof w);
    ActorGradDims d{N, E, D, T, tw, use_tanh, mask_pointer, C, forget_bias};
    const int H = kAgH;
    std::vector<float> cs((size_t)(T + 1) * H), act((size_t)T * 4 * H), dx((size_t)T * E), eref((size_t)N * H), erefT((size_t)H * 128);
    for (long long r = 0; r < R; ++r) {
        ActorGradRow row;
        row.input = input + r * N * D;
        row.idx = idx + r; row.idx_stride = R; row.weight = weight[r];
        row.xscale = xscale ? xscale + r * E : nullptr; row.xscale_stride = R * E;
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

#include "../../mit_rl-vrptw_b200/csrc/critic_grad_row.h"

extern "C" int critic_grad_host(const float* const* wp /*110 pointers in CriticGradWeights order*/, int N, int E, int D, int P, int tw,
                                long long R, const float* input, const float* Rw /*[R]*/, float* hyin /*[R,P+1,H]*/, float* De /*[R,P,N,H]*/,
                                float* sums /*[R,P,4,H]*/, float* dq /*[R,P,H]*/, float* l1 /*[R,H]*/, float* dl1 /*[R,H]*/, float* vdv /*[R,2]*/,
                                float* emb /*[R,N,E]*/, float* const* small /*[P][10] pointers in CriticSmallGradOut order*/) {
    CriticGradWeights w;
    static_assert(sizeof(CriticGradWeights) == 110 * sizeof(void*), "layout");
    std::memcpy(&w, wp, sizeof w);
    CriticGradDims d{N, E, D, P, tw, 2.0f / (float)R};
    const int H = kAgH;
    std::vector<float> ee((size_t)P * N * H), eeT((size_t)P * H * 128), prob((size_t)P * 128);
    for (long long r = 0; r < R; ++r) {
        CriticGradRow row;
        row.input = input + r * N * D; row.R = Rw[r]; row.emb = emb + r * N * E; row.ee = ee.data(); row.eeT = eeT.data(); row.prob = prob.data();
        row.hyin = hyin + r * (P + 1) * H; row.De = De + r * (long long)P * N * H; row.sums = sums + r * P * 4 * H; row.dq = dq + r * P * H;
        row.l1 = l1 + r * H; row.dl1 = dl1 + r * H; row.vdv = vdv + r * 2;
        critic_grad_row(w, d, row);
    }
    static_assert(sizeof(CriticSmallGradOut) == 10 * sizeof(void*), "layout");
    for (int b = 0; b < P; ++b) {
        CriticSmallGradOut o;
        std::memcpy(&o, small + b * 10, sizeof o);
        critic_small_grads(w.blk[b], sums, R, P, b, tw, o);
    }
    return 0;
}
