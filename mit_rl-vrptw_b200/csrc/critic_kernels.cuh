// Critic value and REINFORCE losses: v of RLAgent.build_model('stochastic') (model/attention_agent.py:243-270,
// critics VRPTW/vrptw_attention.py:87-169 and VRP/vrp_attention.py:79-140) and the two losses of build_train_step
// (attention_agent.py:286-294).  The rollout kernel already emits the per-step log-probabilities; the gradients live in train_kernels.cuh.
//
// The critic is n_process_blocks attention passes over the INITIAL instance (demand, time windows, actor embedding), none of
// them masked: hy_0 = 0; (e_i, u_i) = AttentionCritic_i(hy_i); hy_{i+1} = softmax(u_i) . e_i; v = L2(relu(L1(hy_last))).
// Every Conv1D chain is linear in a scalar and the embedding is linear, so per block
//   arg[n,h] = q_h + feat[n] . A[:,h] + demand[n] gd[h]           (the time-window terms (b + e) g_tw fold into A's rows 2, 3:
//                                                                  emb_begin_tw / proj_end_tw serve BOTH windows, :147-150)
//   hy = pf . Au + ce,  pf = sum_n prob[n] feat[n]   =>   q_{i+1} = pf . Mq + cq,   L1 pre-activation = pf . M1 + c1
// i.e. per instance 3 x N x H tanh and a handful of rank-(D-1) products: one warp per instance, lane = 4 hidden units.
#pragma once
#include "common.cuh"

namespace vrpb200 {

constexpr int kCriticBlockFloats = 4 * kH + kH + 4 * kH + kH + kH;   // A (2log2e-scaled) | gd (scaled) | Mq | cq (block 0: its constant) | v
constexpr int kCriticTailFloats = 4 * kH + kH + kH + 4;              // M1 | c1 | w2 | b2

template <bool TW>
__global__ void __launch_bounds__(256) critic_kernel(const float* __restrict__ input, long long B, int N, int D, int nblocks,
                                                     const float* __restrict__ cw, float* __restrict__ v_out) {
    constexpr int FD = TW ? 4 : 2;
    extern __shared__ __align__(16) float csm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int NP = align_up(N, 8);
    float* feat = csm + (size_t)warp * (NP * 5 + NP);   // per warp: [NP][4] features + [NP] demand + [NP] scores
    float* dem = feat + NP * 4;
    float* u = dem + NP;
    const long long b = (long long)blockIdx.x * nw + warp;
    if (b >= B) return;
    const float* in = input + b * N * D;
    for (int n = lane; n < NP; n += 32) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        float d = 0.0f;
        if (n < N) {
            f.x = in[n * D]; f.y = in[n * D + 1];
            if (TW) { f.z = in[n * D + 2]; f.w = in[n * D + 3]; }
            d = in[n * D + D - 1];
        }
        reinterpret_cast<float4*>(feat)[n] = f;
        dem[n] = d;
    }
    __syncwarp();
    float pf[4] = {0.f, 0.f, 0.f, 0.f};
    for (int blk = 0; blk < nblocks; ++blk) {
        const float* w = cw + (size_t)blk * kCriticBlockFloats;
        float cA[FD][4], cgd[4], cv[4], q[4];
#pragma unroll
        for (int k = 0; k < FD; ++k) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(w + k * kH) + lane);
            cA[k][0] = t4.x; cA[k][1] = t4.y; cA[k][2] = t4.z; cA[k][3] = t4.w;
        }
        {
            float4 t4 = __ldg(reinterpret_cast<const float4*>(w + 4 * kH) + lane);
            cgd[0] = t4.x; cgd[1] = t4.y; cgd[2] = t4.z; cgd[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(w + 10 * kH) + lane);
            cv[0] = t4.x; cv[1] = t4.y; cv[2] = t4.z; cv[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(w + 9 * kH) + lane);
            q[0] = t4.x; q[1] = t4.y; q[2] = t4.z; q[3] = t4.w;
        }
        if (blk > 0) {
#pragma unroll
            for (int k = 0; k < FD; ++k) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(w + 5 * kH + k * kH) + lane);
                q[0] = fmaf(pf[k], m4.x, q[0]); q[1] = fmaf(pf[k], m4.y, q[1]);
                q[2] = fmaf(pf[k], m4.z, q[2]); q[3] = fmaf(pf[k], m4.w, q[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] *= kTwoLog2e;
        for (int c0 = 0; c0 < NP; c0 += 8) {
            float part[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = c0 + e;
                const float4 f4 = reinterpret_cast<const float4*>(feat)[n];
                const float f[4] = {f4.x, f4.y, f4.z, f4.w};
                const float dn = dem[n];
                float p = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = q[j];
#pragma unroll
                    for (int k = 0; k < FD; ++k) x = fmaf(f[k], cA[k][j], x);
                    x = fmaf(dn, cgd[j], x);
                    p = fmaf(cv[j], fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(x)), 1.0f), p);
                }
                part[e] = p;
            }
            const float tot = reduce8(part, lane);
            if ((lane & 3) == 0) u[c0 + (lane >> 2)] = tot;
        }
        __syncwarp();
        // prob = tf.nn.softmax(logit) over all N nodes (no mask); pf = sum_n prob[n] feat[n]
        float mx = -INFINITY;
        for (int n = lane; n < N; n += 32) mx = fmaxf(mx, u[n]);
        mx = warp_max(mx);
        float se = 0.0f, acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = lane; n < N; n += 32) {
            const float ev = expf(u[n] - mx);
            const float4 f4 = reinterpret_cast<const float4*>(feat)[n];
            se += ev;
            acc[0] = fmaf(ev, f4.x, acc[0]); acc[1] = fmaf(ev, f4.y, acc[1]);
            acc[2] = fmaf(ev, f4.z, acc[2]); acc[3] = fmaf(ev, f4.w, acc[3]);
        }
        se = warp_sum(se);
        const float inv = 1.0f / se;
#pragma unroll
        for (int k = 0; k < 4; ++k) pf[k] = warp_sum(acc[k]) * inv;
        __syncwarp();
    }
    // v = L2(relu(L1(hy)))  (attention_agent.py:266-268), L1 folded to rank D-1
    const float* t = cw + (size_t)nblocks * kCriticBlockFloats;
    float4 h1 = __ldg(reinterpret_cast<const float4*>(t + 4 * kH) + lane);
#pragma unroll
    for (int k = 0; k < FD; ++k) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(t + k * kH) + lane);
        h1.x = fmaf(pf[k], m4.x, h1.x); h1.y = fmaf(pf[k], m4.y, h1.y);
        h1.z = fmaf(pf[k], m4.z, h1.z); h1.w = fmaf(pf[k], m4.w, h1.w);
    }
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(t + 5 * kH) + lane);
    float s = fmaxf(h1.x, 0.f) * w2.x + fmaxf(h1.y, 0.f) * w2.y + fmaxf(h1.z, 0.f) * w2.z + fmaxf(h1.w, 0.f) * w2.w;
    s = warp_sum(s);
    if (lane == 0) v_out[b] = s + __ldg(t + 6 * kH);
}

// build_train_step (attention_agent.py:286-294): actor_loss = mean((R - stop_gradient(v)) * add_n(logprobs)),
// critic_loss = mean_squared_error(R, v).  One block, fixed summation order (deterministic): thread i owns rows i, i + 1024, ...
__global__ void __launch_bounds__(1024) reinforce_loss_kernel(const float* __restrict__ R, const float* __restrict__ v,
                                                              const float* __restrict__ logprob, long long T, long long rows,
                                                              float* __restrict__ out /* actor, critic */) {
    __shared__ float sa[32], sc[32];
    float a = 0.0f, c = 0.0f;
    for (long long r = threadIdx.x; r < rows; r += 1024) {
        float slp = 0.0f;
        for (long long t = 0; t < T; ++t) slp = __fadd_rn(slp, logprob[t * rows + r]);   // tf.add_n in list order
        const float adv = __fsub_rn(R[r], v[r]);
        a += adv * slp;
        c += adv * adv;
    }
    a = warp_sum(a); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = warp_sum(sa[threadIdx.x]); c = warp_sum(sc[threadIdx.x]);
        if (threadIdx.x == 0) { out[0] = a / (float)rows; out[1] = c / (float)rows; }
    }
}

}  // namespace vrpb200
