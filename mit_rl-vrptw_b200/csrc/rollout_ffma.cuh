// rollout_ffma: the fused rollout (RLAgent.build_model, model/attention_agent.py:84-240) with EVERYTHING on the FP32
// pipe - the independent cross-check of the tensor-core kernels.  One block of instances per CTA, every decode step on-chip.
//
// One CTA owns RB rows (RB/W instances x W beams) for all T steps:
//   phase A  LSTM gates      [RB x 136] x [136 x 512]   (shared/decode_step.py:211-214; x-part folded to
//                                                         rank D-1 through the linear embedding)
//   phase B  query           [RB x 128] x [128 x 128]   (VRPTW/vrptw_attention.py:69)
//   phase C  per row, one warp: pointer scores u[n] = sum_h v_h tanh(.) over the un-masked nodes
//            (vrptw_attention.py:77, SURVEY App. C collapse), mask (decode_step.py:231-232),
//            log_softmax/exp (decode_step.py:125-126), selection (attention_agent.py:146-167),
//            Env.step (VRPTW/vrptw_utils.py:254-358), tour-length accumulation (vrptw_utils.py:397-414).
//            Glimpses (shared/decode_step.py:218-227) run inside the same per-row pass: scores of the glimpse module,
//            softmax, read-out hy = sum_n p[n] e[n,:] collapsed to pf = sum_n p[n] feat[n] (rank D-1), and
//            the next module's query pf . Mq + cq - no second GEMM, no extra launch.  UPS_old physics: a.physics.
// Node features, demand, masks and the LSTM state stay in shared memory / registers for the whole
// rollout; HBM sees the instance once on the way in and T int32 indices + one cost on the way out.
// The weight matrices of phases A and B (387 KB fp32) do not fit next to the instances, so they are
// streamed from L2 every step through a 2-stage shared-memory ring with cp.async.bulk (TMA bulk copy,
// SASS UBLKCP) completing on mbarriers.
#pragma once
#include "common.cuh"

namespace vrpb200 {

struct SmemLayout {
    int ring, hT, feat, dem, mask, rs, scr, bar, total;   // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int RB, int IB, int N, int FD) {
    SmemLayout L;
    const int NP = align_up(N, 4);
    const int nwarps = RB / 4;
    int off = 0;
    L.ring = off; off += kStages * kSlabBytes;            // also holds q[RB][128] during phase C
    L.hT = off;   off += kKG * RB * 4;                    // h (+ features of the previous node), k-major [136][RB]
    L.feat = off; off += align_up(IB * N * FD * 4, 16);
    L.dem = off;  off += RB * NP * 4;
    L.mask = off; off += align_up(RB * NP, 16);
    L.rs = off;   off += RB * kRowState * 4;
    L.scr = off;  off += nwarps * 640;                    // per warp: float u[128] + uint8 list[128]
    L.bar = off;  off += 64;                              // full[2], empty[2], work counter
    L.total = off;
    return L;
}

// =================================================================================================
// Phases A/B as register-tiled FFMA (RB*8 threads), phase C one warp per row over a compacted node list.
template <int RB, bool TW>
__global__ void __launch_bounds__(RB * 8, 1) rollout_ffma_kernel(const RolloutArgs a) {
    constexpr int NT = RB * 8;
    constexpr int NW = NT / 32;
    constexpr int RPW = RB / NW;          // rows per warp in phase C
    constexpr int FD = TW ? 4 : 2;
    static_assert(RB % NW == 0, "rows must divide over the warps");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), IB = a.IB, W = a.W, T = a.T, n_cust = a.n_cust;
    const SmemLayout L = make_layout(RB, IB, N, FD);
    float* ring = reinterpret_cast<float*>(smem + L.ring);
    float* qq = ring;   // phase C view of the ring memory
    float* hT = reinterpret_cast<float*>(smem + L.hT);          // FFMA: h, k-major [136][RB]
    float* feat = reinterpret_cast<float*>(smem + L.feat);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    unsigned char* scr = smem + L.scr;
    const uint32_t bar0 = smem_u32(smem + L.bar);               // full[0], full[1]
    const uint32_t bar_empty = bar0 + 16;
    int* next_row = reinterpret_cast<int*>(smem + L.bar + 48);   // phase C work counter
    const uint32_t ring_u32 = smem_u32(ring);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp >> 1, uh = warp & 1;                    // FFMA ownership
    const long long b0 = (long long)blockIdx.x * IB;
    const int IBv = (int)min((long long)IB, a.B - b0);
    const long long R = a.B * W;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        fence_mbar_init();
    }
    for (int i = tid; i < kKG * RB; i += NT) hT[i] = 0.0f;
    // instance -> shared memory (once): features (x, y[, b_tw, e_tw]) per instance, demand per row
    for (int i = tid; i < IBv * N; i += NT) {
        const int lb = i / N, n = i - lb * N;
        const float* src = a.input + ((b0 + lb) * N + n) * a.D;
        if (TW) {
            reinterpret_cast<float4*>(feat)[i] = make_float4(src[0], src[1], src[2], src[3]);
        } else {
            reinterpret_cast<float2*>(feat)[i] = make_float2(src[0], src[1]);
        }
        const float d = src[a.D - 1];
        for (int w = 0; w < W; ++w) dem[(w * IB + lb) * NP + n] = d;   // Env.reset tiles demand, vrptw_utils.py:213
    }
    __syncthreads();

    float* ulog = reinterpret_cast<float*>(scr + warp * 640);
    unsigned char* list = scr + warp * 640 + 512;

    // decoder input of row lr for the next step = features of node `n` (rank D-1 form of emb[row, idx])
    auto write_next_input = [&](int lr, const float* f) {
        if (lane < FD) hT[(kH + lane) * RB + lr] = f[lane];
    };

    // Env.reset (vrptw_utils.py:199-252) + rollout prologue (attention_agent.py:126-134), one warp per row
    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        for (int n = lane; n < N; n += 32)
            maskb[lr * NP + n] = (n == n_cust) ? 1 : (dem[lr * NP + n] == 0.0f ? 1 : 0);
        const float* fdep = feat + (size_t)(lb * N + n_cust) * FD;
        if (lane == 0) {
            float* rs = rowst + lr * kRowState;
            rs[0] = a.capacity;  // load
            rs[1] = 0.0f;        // time
            rs[2] = fdep[0];     // previous_x
            rs[3] = fdep[1];     // previous_y
            rs[4] = 0.0f;        // tour length so far
            rs[5] = 0.0f;        // x of first action
            rs[6] = 0.0f;        // y of first action
            reinterpret_cast<int*>(rs)[8] = 0;   // done
        }
        write_next_input(lr, fdep);   // decoder_input = embedding of the depot
    }
    __syncthreads();

    // ownership: (rows 8rg..8rg+7) x (units u0, u0+1); the LSTM cell state c lives in registers for all T steps
    const int u0 = uh * 64 + 2 * lane;
    const float2 bq2 = __ldg(reinterpret_cast<const float2*>(a.q_bq + u0));
    constexpr int NCST = 16;
    float cst[NCST];
#pragma unroll
    for (int r = 0; r < NCST; ++r) cst[r] = 0.0f;

#ifdef VRPB200_PHASE_CLOCKS
    long long pc_[6] = {0, 0, 0, 0, 0, 0}, pc_last_ = clock64();
#endif
    uint32_t slabcnt = 0;     // ring uses so far (producer and consumer count identically)
    unsigned n_eval = 0, n_rowsteps = 0;   // work counters of this warp (roofline accounting)
    int t = 0;
    for (; t < T; ++t) {
        if (tid == 0) *next_row = 0;   // ordered before phase C by the barriers of phases A/B
            // ================= phase A: gates = [h, feat(prev node)] . Wg ===================================
            float acc[8][8];
    #pragma unroll
            for (int r = 0; r < 8; ++r)
    #pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
            if (tid == 0) {
                fence_proxy_async();   // q[] was written into the ring memory by generic stores
                for (int s = 0; s < kStages; ++s) {
                    const uint32_t st = (slabcnt + s) % kStages;
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.Wg + (size_t)s * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
            }
            for (int s = 0; s < kGateSlabs; ++s) {
                const uint32_t st = slabcnt % kStages;
                mbar_wait(bar0 + 8 * st, (slabcnt / kStages) & 1);
                const float* sl = ring + st * kSlabFloats;
    #pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int k = s * 8 + kk;
                    const float4 a0 = *reinterpret_cast<const float4*>(hT + k * RB + 8 * rg);
                    const float4 a1 = *reinterpret_cast<const float4*>(hT + k * RB + 8 * rg + 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(sl + ((kk * 2 + uh) * 2 + 0) * 128 + lane * 4);
                    const float4 w1 = *reinterpret_cast<const float4*>(sl + ((kk * 2 + uh) * 2 + 1) * 128 + lane * 4);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    #pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        acc[r][0] = fmaf(av[r], w0.x, acc[r][0]); acc[r][1] = fmaf(av[r], w0.y, acc[r][1]);
                        acc[r][2] = fmaf(av[r], w0.z, acc[r][2]); acc[r][3] = fmaf(av[r], w0.w, acc[r][3]);
                        acc[r][4] = fmaf(av[r], w1.x, acc[r][4]); acc[r][5] = fmaf(av[r], w1.y, acc[r][5]);
                        acc[r][6] = fmaf(av[r], w1.z, acc[r][6]); acc[r][7] = fmaf(av[r], w1.w, acc[r][7]);
                    }
                }
                __syncthreads();   // every warp is done with this stage (and, after the last slab, with hT)
                if (tid == 0 && s + kStages < kGateSlabs) {
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.Wg + (size_t)(s + kStages) * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
                ++slabcnt;
            }
            // prefetch the first q slabs while the gate epilogue runs
            if (tid == 0) {
                for (int s = 0; s < kStages; ++s) {
                    const uint32_t st = (slabcnt + s) % kStages;
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.q_Wq + (size_t)s * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
            }
            // BasicLSTMCell (TF-1 rnn_cell_impl): c' = c*sig(f+fb) + sig(i)*tanh(j); h' = tanh(c')*sig(o)
            {
                const float4 bg0 = __ldg(reinterpret_cast<const float4*>(a.bg) + u0);
                const float4 bg1 = __ldg(reinterpret_cast<const float4*>(a.bg) + u0 + 1);
                float hv[2][8];
    #pragma unroll
                for (int r = 0; r < 8; ++r) {
    #pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const float4 bb = p ? bg1 : bg0;
                        const float gi = acc[r][p * 4 + 0] + bb.x, gj = acc[r][p * 4 + 1] + bb.y;
                        const float gf = acc[r][p * 4 + 2] + bb.z, go = acc[r][p * 4 + 3] + bb.w;
                        const float cn = __fadd_rn(__fmul_rn(cst[r * 2 + p], sigmoid_acc(gf + a.forget_bias)),
                                                   __fmul_rn(sigmoid_acc(gi), tanh_acc(gj)));
                        cst[r * 2 + p] = cn;
                        hv[p][r] = __fmul_rn(tanh_acc(cn), sigmoid_acc(go));
                    }
                }
    #pragma unroll
                for (int p = 0; p < 2; ++p) {
                    float* dst = hT + (u0 + p) * RB + 8 * rg;
                    *reinterpret_cast<float4*>(dst) = make_float4(hv[p][0], hv[p][1], hv[p][2], hv[p][3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[p][4], hv[p][5], hv[p][6], hv[p][7]);
                }
            }
            __syncthreads();   // h' complete

            // ================= phase B: q = h' . W_q =========================================================
            float qa[8][2];
    #pragma unroll
            for (int r = 0; r < 8; ++r) qa[r][0] = qa[r][1] = 0.0f;
            for (int s = 0; s < kQSlabs; ++s) {
                const uint32_t st = slabcnt % kStages;
                mbar_wait(bar0 + 8 * st, (slabcnt / kStages) & 1);
                const float* sl = ring + st * kSlabFloats;
    #pragma unroll 8
                for (int kk = 0; kk < 32; ++kk) {
                    const int k = s * 32 + kk;
                    const float4 a0 = *reinterpret_cast<const float4*>(hT + k * RB + 8 * rg);
                    const float4 a1 = *reinterpret_cast<const float4*>(hT + k * RB + 8 * rg + 4);
                    const float2 w = *reinterpret_cast<const float2*>(sl + kk * 128 + u0);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    #pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        qa[r][0] = fmaf(av[r], w.x, qa[r][0]);
                        qa[r][1] = fmaf(av[r], w.y, qa[r][1]);
                    }
                }
                __syncthreads();
                if (tid == 0 && s + kStages < kQSlabs) {
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.q_Wq + (size_t)(s + kStages) * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
                ++slabcnt;
            }
            // ring is idle from here to the next phase A: park q (+ every constant bias) in it
    #pragma unroll
            for (int r = 0; r < 8; ++r)
                *reinterpret_cast<float2*>(qq + (8 * rg + r) * kH + u0) = make_float2(qa[r][0] + bq2.x, qa[r][1] + bq2.y);
        __syncthreads();
        VRP_CLK(3);

        // ================= phase C: one warp per row ====================================================
        // attention constants of this lane's 4 hidden units h = 4*lane .. 4*lane+3, per module (glimpses, then the pointer)
        float cA[FD][4], cgd[4], cv[4], cgl[4], cgt[4];
        auto load_att = [&](const AttWeights& w) {
#pragma unroll
            for (int k = 0; k < FD; ++k) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(w.A + k * kH) + lane);
                cA[k][0] = t4.x; cA[k][1] = t4.y; cA[k][2] = t4.z; cA[k][3] = t4.w;
            }
            float4 t4 = __ldg(reinterpret_cast<const float4*>(w.gd) + lane);
            cgd[0] = t4.x; cgd[1] = t4.y; cgd[2] = t4.z; cgd[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(w.v) + lane);
            cv[0] = t4.x; cv[1] = t4.y; cv[2] = t4.z; cv[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(w.gl) + lane);
            cgl[0] = t4.x; cgl[1] = t4.y; cgl[2] = t4.z; cgl[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(w.gt) + lane);
            cgt[0] = t4.x; cgt[1] = t4.y; cgt[2] = t4.z; cgt[3] = t4.w;
        };
        const int G = a.n_glimpses;
        if (G == 0) load_att(a.att);
        int warp_done = 1;
        for (;;) {   // rows are handed out dynamically: their cost varies with the number of un-masked nodes
            int lr = 0;
            if (lane == 0) lr = atomicAdd(next_row, 1);
            lr = __shfl_sync(kFull, lr, 0);
            if (lr >= RB) break;
            const int wb = lr / IB, lb = lr - wb * IB;
            if (wb >= W || lb >= IBv) continue;
            float* rs = rowst + lr * kRowState;
            const long long row = (long long)wb * a.B + b0 + lb;
            if (reinterpret_cast<int*>(rs)[8]) {   // finished: the policy emits the depot with prob 1 forever
                if (lane == 0) {
                    a.idx_out[(long long)t * R + row] = n_cust;
                    if (a.logprob_out) a.logprob_out[(long long)t * R + row] = 0.0f;
                }
                continue;
            }
            const float load = rs[0], time = TW ? rs[1] : 0.0f;
            const unsigned char* mk = maskb + lr * NP;
            const float* frow = feat + (size_t)lb * N * FD;
            float* drow = dem + lr * NP;

            // --- node lists: lane owns nodes lane + 32 j
            int mcnt[4];
            unsigned actbits[4];
            int nact = 0, nun = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                mcnt[j] = (n < N) ? mk[n] : 255;
                nun += __popc(__ballot_sync(kFull, mcnt[j] == 0));
            }
            const bool all_nodes = a.full || nun == 0 || !a.mask_pointer || (G > 0 && !a.mask_glimpses);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                const bool act = (n < N) && (all_nodes || mcnt[j] == 0);
                actbits[j] = __ballot_sync(kFull, act);
                if (act) list[nact + __popc(actbits[j] & ((1u << lane) - 1))] = (unsigned char)n;
                nact += __popc(actbits[j]);
            }
            const int npad = align_up(nact, 8);
            n_eval += nact;
            ++n_rowsteps;
            if (lane < npad - nact) list[nact + lane] = 0;   // padding entries: node 0, result ignored
            __syncwarp();

            // --- u[n] = sum_h v_h tanh(q_h + e[n,h] + d[n,h] + ld[n,h] + t_h), all pre-scaled by 2 log2 e; once per glimpse
            //     module (query = LSTM output for the first one) and once for the pointer
            float qv[4];
            {
                const float4 q4 = *reinterpret_cast<const float4*>(qq + lr * kH + 4 * lane);
                qv[0] = q4.x; qv[1] = q4.y; qv[2] = q4.z; qv[3] = q4.w;
            }
#pragma unroll 1
            for (int mod = 0; mod <= G; ++mod) {
                if (G > 0) load_att(mod < G ? a.gatt[mod] : a.att);
                float base[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = fmaf(load, cgl[j], qv[j]);
                    if (TW) s = fmaf(time, cgt[j], s);
                    base[j] = s * kTwoLog2e;
                }
                for (int c0 = 0; c0 < npad; c0 += 8) {
                    const uint2 ids = *reinterpret_cast<const uint2*>(list + c0);
                    float part[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int n = ((e < 4 ? ids.x : ids.y) >> (8 * (e & 3))) & 0xff;
                        float f[FD];
                        if (TW) {
                            const float4 f4 = *reinterpret_cast<const float4*>(frow + n * 4);
                            f[0] = f4.x; f[1] = f4.y; f[FD - 2] = f4.z; f[FD - 1] = f4.w;
                        } else {
                            const float2 f2 = *reinterpret_cast<const float2*>(frow + n * 2);
                            f[0] = f2.x; f[1] = f2.y;
                        }
                        const float dn = drow[n];
                        float p = 0.0f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = base[j];
#pragma unroll
                            for (int k = 0; k < FD; ++k) x = fmaf(f[k], cA[k][j], x);
                            x = fmaf(dn, cgd[j], x);
                            const float rr = rcp_approx(1.0f + ex2_approx(x));
                            p = fmaf(cv[j], fmaf(-2.0f, rr, 1.0f), p);
                        }
                        part[e] = p;
                    }
                    const float tot = reduce8(part, lane);
                    const int pos = c0 + (lane >> 2);
                    if ((lane & 3) == 0 && pos < nact) ulog[list[pos]] = tot;
                }
                __syncwarp();
                if (mod == G) break;
                // --- glimpse (shared/decode_step.py:218-227): prob = softmax(logit - 1e5 mask); hy = prob . e  ->  the next
                //     module's query q' = hy . W_q' + b' = pf . Mq + cq with pf = sum_n prob[n] feat[n]
                float gl_[4], gmx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    float u = -INFINITY;
                    if ((actbits[j] >> lane) & 1) {
                        u = ulog[n];
                        if (a.mask_glimpses) u = __fsub_rn(u, __fmul_rn(kBig, (float)mcnt[j]));
                    }
                    gl_[j] = u;
                    gmx = fmaxf(gmx, u);
                }
                gmx = warp_max(gmx);
                float gse = 0.0f, pf[FD];
#pragma unroll
                for (int k = 0; k < FD; ++k) pf[k] = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    const float ev = (gl_[j] == -INFINITY) ? 0.0f : expf(gl_[j] - gmx);
                    gse += ev;
                    if (ev != 0.0f) {
#pragma unroll
                        for (int k = 0; k < FD; ++k) pf[k] = fmaf(ev, frow[n * FD + k], pf[k]);
                    }
                }
                gse = warp_sum(gse);
                const float ginv = 1.0f / gse;
#pragma unroll
                for (int k = 0; k < FD; ++k) pf[k] = warp_sum(pf[k]) * ginv;
                const AttWeights& gw = a.gatt[mod];
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(gw.cq) + lane);
                qv[0] = c4.x; qv[1] = c4.y; qv[2] = c4.z; qv[3] = c4.w;
#pragma unroll
                for (int k = 0; k < FD; ++k) {
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(gw.Mq + k * kH) + lane);
                    qv[0] = fmaf(pf[k], m4.x, qv[0]); qv[1] = fmaf(pf[k], m4.y, qv[1]);
                    qv[2] = fmaf(pf[k], m4.z, qv[2]); qv[3] = fmaf(pf[k], m4.w, qv[3]);
                }
                __syncwarp();
            }

            // --- mask, log_softmax, exp (shared/decode_step.py:231-232, 125-126)
            float lg[4], pr[4];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                const bool act = (actbits[j] >> lane) & 1;
                float u = -INFINITY;
                if (act) {
                    u = ulog[n];
                    if (a.use_tanh) u = a.tanh_C * tanhf(u);
                    if (a.mask_pointer) u = __fsub_rn(u, __fmul_rn(kBig, (float)mcnt[j]));
                }
                lg[j] = u;
                mx = fmaxf(mx, u);
            }
            mx = warp_max(mx);
            float se = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) se += (lg[j] == -INFINITY) ? 0.0f : expf(lg[j] - mx);
            se = warp_sum(se);
            const float lse = logf(se);
            float bestp = -1.0f;
            int besti = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                pr[j] = (lg[j] == -INFINITY) ? 0.0f : expf((lg[j] - mx) - lse);
                if (n < N && pr[j] > bestp) { bestp = pr[j]; besti = n; }
            }
            if (a.tr_logits || a.tr_probs || a.tr_masks) {
                const long long o = ((long long)t * R + row) * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    if (n < N) {
                        if (a.tr_logits) a.tr_logits[o + n] = lg[j];
                        if (a.tr_probs) a.tr_probs[o + n] = pr[j];
                        if (a.tr_masks) a.tr_masks[o + n] = (float)mcnt[j];
                    }
                }
            }

            // --- selection
            int idx;
            float psel;
            if (a.mode == 0) {   // greedy: argmax(prob), first maximal index (attention_agent.py:146-147)
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const float ov = __shfl_xor_sync(kFull, bestp, o);
                    const int oi = __shfl_xor_sync(kFull, besti, o);
                    if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                }
                idx = besti;
                psel = bestp;
            } else {             // inverse-CDF sampling with one retry (attention_agent.py:151-167)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    if (n < N) ulog[n] = pr[j];
                }
                __syncwarp();
                int sel = -1;
                if (lane == 0) {
                    const long long o = (long long)t * R + row;
                    int d0_, d1_;
                    draw_range(a, t, d0_, d1_);   // per-row retry, or the reference's whole-batch redraw (attention_agent.py:165-167)
                    for (int draw = d0_; draw < d1_ && sel < 0; ++draw) {
                        const float* ub = draw ? a.uniforms_retry : a.uniforms;
                        const float u = ub ? ub[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                        float cum = 0.0f;   // tf.cumsum: sequential fp32; skipped nodes have prob exactly 0
                        for (int k = 0; k < nact; ++k) {
                            const int n = list[k];
                            cum = __fadd_rn(cum, ulog[n]);
                            if (cum > u) { sel = n; break; }
                        }
                    }
                    if (sel < 0 && a.fail_steps && d0_ == 0) a.fail_steps[t] = 1;   // some row of the batch needs the redraw: the host re-runs with this step redrawn
                    if (sel < 0) sel = 0;   // argmin of an all-10000 row
                }
                idx = __shfl_sync(kFull, sel, 0);
                __syncwarp();
                psel = ulog[idx];
                __syncwarp();
            }

            // --- Env.step (VRPTW/vrptw_utils.py:294-348), tour length, next decoder input
            const float* fsel = frow + idx * FD;
            const float sx = fsel[0], sy = fsel[1];
            const float px = rs[2], py = rs[3];
            const float d_idx = drow[idx];
            __syncwarp();
            const float d_sat = fminf(d_idx, load);
            float nload = __fsub_rn(load, d_sat);
            const float ts = move_len(a.physics, px, py, sx, sy);   // what reward_func adds up; its travel time below
            float ntime = 0.0f;
            if (TW) {
                ntime = fmaxf(__fadd_rn(time, travel_time(a.physics, ts)), fsel[2]);
                if (a.physics) ntime = __fadd_rn(ntime, __fmul_rn(kGeoService, d_sat));   // UPS_old/vrptw_ups_utils.py:332-333
            }
            const float dep = (idx == n_cust) ? 1.0f : 0.0f;
            nload = __fadd_rn(__fmul_rn(nload, 1.0f - dep), __fmul_rn(dep, a.capacity));
            if (TW) ntime = __fmul_rn(ntime, 1.0f - dep);
            if (lane == 0) drow[idx] = __fsub_rn(d_idx, d_sat);
            __syncwarp();
            float dn[4], sumdem = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                dn[j] = (n < N) ? drow[n] : 0.0f;
                sumdem += dn[j];
            }
            sumdem = warp_sum(sumdem);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = lane + 32 * j;
                if (n < N) {
                    float nx = 0.f, ny = 0.f, ne = 0.f;
                    if (TW) {
                        const float4 f4 = *reinterpret_cast<const float4*>(frow + n * 4);
                        nx = f4.x; ny = f4.y; ne = f4.w;
                    }
                    maskb[lr * NP + n] =
                        (unsigned char)node_mask<TW>(n, n_cust, dn[j], nload, ntime, sx, sy, nx, ny, ne, dep, sumdem, a.physics);
                }
            }
            const int done_now = (!a.full && a.mask_pointer && idx == n_cust && sumdem == 0.0f) ? 1 : 0;
            if (lane == 0) {
                rs[0] = nload;
                rs[1] = ntime;
                rs[2] = sx;
                rs[3] = sy;
                if (t == 0) { rs[5] = sx; rs[6] = sy; }
                else rs[4] = __fadd_rn(rs[4], ts);   // |a_t - a_{t-1}| (vrptw_utils.py:403-408)
                reinterpret_cast<int*>(rs)[8] = done_now;
                a.idx_out[(long long)t * R + row] = idx;
                if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
            }
            write_next_input(lr, fsel);   // decoder_input = emb[row, idx] (attention_agent.py:211-212)
            warp_done &= done_now;
            __syncwarp();
        }
        VRP_CLK(4);
        const int all_done = __syncthreads_and(warp_done);
        VRP_CLK(5);
        if (all_done) { ++t; break; }
    }
    const int t_end = t;   // steps executed by this block

    // epilogue: finished rows keep choosing the depot (zero-length moves); close the tour
    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        const long long row = (long long)wb * a.B + b0 + lb;
        for (int tt = t_end + lane; tt < T; tt += 32) {
            a.idx_out[(long long)tt * R + row] = n_cust;
            if (a.logprob_out) a.logprob_out[(long long)tt * R + row] = 0.0f;
        }
        const float* rs = rowst + lr * kRowState;
        if (lane == 0) a.cost_out[row] = __fadd_rn(rs[4], move_len(a.physics, rs[2], rs[3], rs[5], rs[6]));   // + |a_0 - a_{T-1}|
        for (int n = lane; n < N; n += 32) {
            if (a.fin_demand) a.fin_demand[row * N + n] = dem[lr * NP + n];
            if (a.fin_mask) a.fin_mask[row * N + n] = (float)maskb[lr * NP + n];
        }
        if (lane == 0) {
            if (a.fin_load) a.fin_load[row] = rs[0];
            if (a.fin_time) a.fin_time[row] = rs[1];
        }
    }
    if (tid == 0 && a.steps_out) a.steps_out[blockIdx.x] = t_end;
    if (a.stats && lane == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)n_eval);
        atomicAdd(a.stats + 1, (unsigned long long)n_rowsteps);
        if (tid == 0) {
            atomicAdd(a.stats + 2, (unsigned long long)t_end * RB);
            atomicAdd(a.stats + 3, 1ull);
        }
#ifdef VRPB200_PHASE_CLOCKS
        if (tid == 64)
            for (int i = 0; i < 6; ++i) atomicAdd(a.stats + 4 + i, (unsigned long long)pc_[i]);
#endif
    }
}

}  // namespace vrpb200