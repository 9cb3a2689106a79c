// rollout5_kernel: the fused rollout (RLAgent.build_model of model/attention_agent.py:84-240, greedy and stochastic) as a
// warp-specialised pipeline in which the MUFU-bound score phase does not wait for the latency-bound rest of a decode step.
//
//   warps  0..15  score warps   per decode step: (1) Env mask -> item lists of the step (vrptw_utils.py:328-348; one warp per
//                               row, 4 nodes per lane, hidden behind the query GEMM), (2) query epilogue, (3) pointer scores
//                               u[n] = sum_h v_h tanh(.) (VRPTW/vrptw_attention.py:77) over 128-item tiles (item = un-masked
//                               (row, node)): E = F.Bc on tcgen05 (3xTF32), tanh at 1.25 MUFU per element (four reciprocals
//                               share one rcp, packed FP32x2 algebra); two groups of 8 warps, the two warps of a TMEM
//                               sub-partition share a tile's items and split the hidden units; (4) BETWEEN tiles, the LSTM
//                               epilogue of the next step (gates^T from TMEM + x part -> c, h'; shared/decode_step.py:211-214)
//                               for every 4-row chunk whose rows have moved
//   warps 16..19  aux warps     the per-row tail, STREAMED behind the score tiles: as soon as the tiles holding a row's items
//                               are done, one warp picks the row's node (arg-max with exact soft-max on near-ties, or sampling;
//                               attention_agent.py:146-167) and runs Env.step (vrptw_utils.py:254-325); four rows = one chunk
//   warp  20      weight feeder (cp.async.bulk ring from L2: W_q^T, then W_h^T for the next step's gates)
//   warp  21      LSTM / query MMA issuer (transposed GEMMs, M = 128 units, N = rows, TMEM accumulators)
//   warps 22, 23  score-MMA issuers of the two groups (a rotating ring of three 32-column TMEM buffers per group)
//
// At the end of the score phase only the tail of the last rows and one LSTM chunk are exposed.  Registers are re-distributed
// with setmaxnreg inside the CTA's own pool of 768 x 80 (score warps 96, aux warps 72, service warps 24: an increase can only
// be served from what the other warps of the CTA have released).  The LSTM cell state lives in TMEM (lane = unit, one column
// per row).  Long waits of many warps are hardware barriers (bar.sync), never polling mbarrier waits: 16 polling warps take
// the issue slots of the warps they wait for.
// One synchronisation point per decode step: T0 = "h'(t) is in the X tiles and every row has moved"; both GEMMs of the step
// are issued at T0, the gates GEMM (whose result is needed one step later) in the background of the score phase.
#pragma once
#include "../common.cuh"

namespace vrpb200 {

constexpr int kV5Threads = 768;   // 16 score warps + 4 aux warps + 4 service warps
constexpr int kRS5 = 18;          // floats of row state

struct Smem5 {
    int ring, X, qq, F, logits, items, bc, vsm, psum, dem, mask, rs, ibits, wx, meta, bar, total, stages;
};

__host__ __device__ inline Smem5 make_layout5(int RB, int N, int stages) {
    Smem5 L;
    const int NP = align_up(N, 4);
    int off = 0;
    L.stages = stages;
    L.ring = off;   off += stages * kSlabBytes;
    L.X = off;      off += 2 * RB * kH * 2;                    // h' as the K-major B operand, FP16 hi | lo
    L.qq = off;     off += RB * kH * 4;                        // base'[row][unit]
    L.F = off;      off += 2 * 2 * 2 * kC1TileBytes;           // 2 groups x 2 F tiles x (hi | lo)
    L.logits = off; off += RB * NP * 4;
    L.items = off;  off += align_up(RB * NP, 16);              // node ids of the items of every row
    L.bc = off;     off += 2 * 4096;
    L.vsm = off;    off += 512;                                // -2 v
    L.psum = off;   off += 2 * 2 * 128 * 4;                    // partial score sums, [group][tile parity][item]
    L.dem = off;    off += RB * NP * 4;
    L.mask = off;   off += align_up(RB * NP, 16);              // mask counts (exact tail only)
    L.rs = off;     off += align_up(RB * kRS5 * 4, 16);
    L.ibits = off;  off += RB * 16;                            // per row: 128-bit set of the nodes that are items
    L.wx = off;     off += 20 * kH * 4;                        // x part of the LSTM gates: W_x^T [4 features][4 gates][128] + bias [4][128]
    L.meta = off;   off += align_up((RB + 68 + RB + 8) * 4, 16);   // nact[RB], off[<=65 (+pad)], flags[RB], stop, t_end
    L.bar = off;    off += 320 + 64 * 8;
    L.total = off;
    return L;
}

// non-blocking test of an mbarrier phase
VRP_DEVINL bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// row state: 0 load, 1 time, 2-3 position, 4 tour length so far, 5-6 first node visited, 7 depot x, 8 done (int),
// 9-12 features of the node chosen last (x part of the LSTM input), 13 customers with demand left, 14 depot y,
// 15 "stands at the depot", 16-17 depot b_tw, e_tw
template <int RB, bool TW, int STAGES>
__global__ void __launch_bounds__(kV5Threads, 1) rollout5_kernel(const __grid_constant__ RolloutArgs a) {
    constexpr int RPT = RB / 4, FD = TW ? 4 : 2, NG = 2, NCH = RB / 4;
    constexpr int kGCol = 192;                                   // gates^T: tile m at kGCol + RB m
    constexpr int kQCol = 192 + 4 * RB;                          // q^T
    constexpr int kCCol = 192 + 5 * RB;                          // LSTM cell state c[unit][row]
    static_assert(RB % 16 == 0 && RB <= 48, "rows");
    static_assert(192 + 6 * RB <= kTcCols, "TMEM columns");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), T = a.T, n_cust = a.n_cust;
    const Smem5 L = make_layout5(RB, N, STAGES);
    unsigned char* xt = smem + L.X;
    float* qq = reinterpret_cast<float*>(smem + L.qq);
    unsigned char* ftile = smem + L.F;
    float* logits = reinterpret_cast<float*>(smem + L.logits);
    unsigned char* items = smem + L.items;
    float* vsm = reinterpret_cast<float*>(smem + L.vsm);
    float* psum = reinterpret_cast<float*>(smem + L.psum);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    unsigned* ibits = reinterpret_cast<unsigned*>(smem + L.ibits);
    const float* wxs = reinterpret_cast<const float*>(smem + L.wx);   // [(feature 4 + gate)][unit], then bias [gate][unit] at 16 * kH
    int* nact_s = reinterpret_cast<int*>(smem + L.meta);
    int* off_s = nact_s + RB;            // [RB + 1] exclusive prefix of nact (this step)
    int* flag_s = off_s + 68;            // [RB] every node of the row is an item (the mask COUNTS decide the choice)
    volatile int* stop_s = flag_s + RB;
    volatile int* tend_s = stop_s + 1;
    const uint32_t bar0 = smem_u32(smem + L.bar);
    const uint32_t bar_empty = bar0 + 32, bar_q = bar0 + 64, bar_g = bar0 + 72, bar_x = bar0 + 80, bar_items = bar0 + 88,
                   tmem_slot = bar0 + 120,
                   bar_grp = bar0 + 128,   // group g at + 64 g: full[3] (MMA commit), free[3] (4 warps copied the buffer), F ready[2]
                   bar_psum = bar0 + 256, bar_pack = bar0 + 272, bar_tile = bar0 + 320;
    const uint32_t ring_u32 = smem_u32(smem + L.ring), bc_u32 = smem_u32(smem + L.bc);
    const uint32_t x_hi = smem_u32(xt), x_lo = x_hi + RB * kH * 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sp = warp & 3;
    const long long b0 = (long long)blockIdx.x * RB;
    const int IBv = (int)min((long long)RB, a.B - b0);
    const long long R = a.B;
    const int D = a.D;
    // node features re-packed by repack_features_kernel: (x, y, b_tw, e_tw) of a node in ONE 16-byte load
    const float4* gfeat4 = a.feat4 + b0 * (long long)N;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_q, 1); mbar_init(bar_g, 1); mbar_init(bar_x, 1); mbar_init(bar_items, 1);
        for (int g = 0; g < NG; ++g) {
            for (int b = 0; b < 3; ++b) { mbar_init(bar_grp + 64 * g + 8 * b, 1); mbar_init(bar_grp + 64 * g + 24 + 8 * b, 4); }
            mbar_init(bar_grp + 64 * g + 48, 4); mbar_init(bar_grp + 64 * g + 56, 4);
            mbar_init(bar_psum + 8 * g, 4);    // the partial sums of a tile are in place
            mbar_init(bar_pack + 8 * g, 4);    // ... and have been consumed (the producing half runs at most one tile ahead)
        }
        for (int i = 0; i < 64; ++i) mbar_init(bar_tile + 8 * i, 4);    // 4 arrivals per step: the 4 warps that finish the tile
        *stop_s = 0;
        *tend_s = 0;
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc_alloc(tmem_slot, kTcCols);
    for (int i = tid; i < 128 * 8; i += kV5Threads) {
        const int h = i >> 3, k = i & 7;
        float val = 0.0f;
        if (k < FD) val = a.att.A[k * kH + h];
        else if (k == 4) val = a.att.gd[h];
        const float hi = tf32_rna(val), lo = tf32_rna(val - hi);
        const int o = (h >> 3) * 256 + (k >> 2) * 128 + (h & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(smem + L.bc + o) = hi;
        *reinterpret_cast<float*>(smem + L.bc + 4096 + o) = lo;
    }
    if (tid < 128) vsm[tid] = a.vm2[tid];
    for (int i = tid; i < 20 * kH; i += kV5Threads)
        reinterpret_cast<float*>(smem + L.wx)[i] = i < 16 * kH ? a.WxT[i] : a.bgT[i - 16 * kH];
    for (int i = tid; i < IBv * N; i += kV5Threads) {
        const int lb = i / N, n = i - lb * N;
        dem[lb * NP + n] = a.input[((b0 + lb) * N + n) * D + D - 1];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.bar + 120);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * sp) << 16);
    constexpr uint32_t idescL = tc_idesc_f16(128, RB), idescQ = tc_idesc(128, 32);
    const bool tracing = a.tr_logits || a.tr_probs || a.tr_masks;
    // the fast per-row path serves plain greedy decoding when nobody reads probabilities, log-probabilities or masks
    const bool fastk = a.mode == 0 && !a.full && a.mask_pointer && !a.use_tanh && a.logprob_out == nullptr && a.fin_mask == nullptr && !tracing;
    const unsigned ltm = (1u << lane) - 1;

    // 128-bit node set -> the row's item list (ascending node ids), the set itself, count and "all nodes" flag
    auto publish_items = [&](int lr, unsigned (&bits)[4], bool on) {
        int nit = __popc(bits[0]) + __popc(bits[1]) + __popc(bits[2]) + __popc(bits[3]);
        int fl = 0;
        if (on && (a.full || nit == 0 || !a.mask_pointer)) {   // every node is evaluated (the mask counts decide)
            fl = 1; nit = N;
#pragma unroll
            for (int k = 0; k < 4; ++k) bits[k] = (N - 32 * k >= 32) ? kFull : (N - 32 * k > 0 ? ((1u << (N - 32 * k)) - 1) : 0u);
        }
        if (!on) { nit = 0; bits[0] = bits[1] = bits[2] = bits[3] = 0u; }
        unsigned char* it = items + lr * NP;
        int base = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((bits[k] >> lane) & 1) it[base + __popc(bits[k] & ltm)] = (unsigned char)(lane + 32 * k);
            base += __popc(bits[k]);
        }
        if (lane < 4) ibits[lr * 4 + lane] = lane == 0 ? bits[0] : lane == 1 ? bits[1] : lane == 2 ? bits[2] : bits[3];
        if (lane == 0) { nact_s[lr] = nit; flag_s[lr] = fl; }
        return fl;
    };

    // ---- Env.reset (vrptw_utils.py:199-252) + prologue (attention_agent.py:126-134): one warp per row, lane owns nodes lane + 32 k
    if (warp < 16) {
        for (int lr = warp; lr < RB; lr += 16) {
            float* rs = rowst + lr * kRS5;
            unsigned bits[4] = {0u, 0u, 0u, 0u};
            if (lr >= IBv) {   // no instance behind this row slot: finished from the start, no items
                if (lane == 0) { reinterpret_cast<int*>(rs)[8] = 1; rs[9] = rs[10] = rs[11] = rs[12] = 0.0f; rs[0] = rs[1] = 0.0f; }
                publish_items(lr, bits, false);
                continue;
            }
            int nzc = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = lane + 32 * k;
                const bool has = n < n_cust && dem[lr * NP + n] != 0.0f;
                if (n < N) maskb[lr * NP + n] = (n == n_cust) ? 1 : (has ? 0 : 1);
                bits[k] = __ballot_sync(kFull, has);
                nzc += __popc(bits[k]);
            }
            publish_items(lr, bits, true);
            const float4 fdep = __ldcg(gfeat4 + (size_t)lr * N + n_cust);
            if (lane == 0) {
                rs[0] = a.capacity; rs[1] = 0.0f; rs[2] = fdep.x; rs[3] = fdep.y; rs[4] = 0.0f; rs[5] = 0.0f; rs[6] = 0.0f;
                rs[7] = fdep.x; rs[14] = fdep.y; rs[16] = TW ? fdep.z : 0.0f; rs[17] = TW ? fdep.w : 0.0f;
                reinterpret_cast<int*>(rs)[8] = 0;
                rs[9] = fdep.x; rs[10] = fdep.y; rs[11] = TW ? fdep.z : 0.0f; rs[12] = TW ? fdep.w : 0.0f;
                rs[13] = (float)nzc; rs[15] = 0.0f;
            }
        }
    }
    __syncthreads();

    if (warp >= 20) {
        // ============ service warps ================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 20 && lane == 0) {
            // ---- weight feeder: the ring is fed from L2 with cp.async.bulk; q(t+1)'s first stages are fetched as soon as
            //      ring slots free up, so that the query GEMM - the one the critical path waits for - finds operands in place
            uint32_t c = 0;
            int pre = 0;
            auto issue = [&](int u) {
                const uint32_t st = c % STAGES;
                mbar_wait(bar_empty + 8 * st, ((c / STAGES) & 1) ^ 1);
                mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                const float* src = u < kHUsesQ ? a.WqTh + (size_t)u * kSlabFloats : a.WgTh + (size_t)(u - kHUsesQ) * kSlabFloats;
                bulk_g2s(ring_u32 + st * kSlabBytes, src, kSlabBytes, bar0 + 8 * st);
                ++c;
            };
            for (int step = 0;; ++step) {
                mbar_wait(bar_x, step & 1);
                if (*stop_s) {   // loads already in flight must land before the CTA may exit
                    for (uint32_t j = c - pre; j < c; ++j) mbar_wait(bar0 + 8 * (j % STAGES), (j / STAGES) & 1);
                    break;
                }
                for (int u = pre; u < kHUsesQ + kHUsesG; ++u) issue(u);
                pre = STAGES < kHUsesQ ? STAGES : kHUsesQ;
                for (int u = 0; u < pre; ++u) issue(u);
            }
        } else if (warp == 21) {
            // ---- LSTM / query MMA issuer (the whole warp runs the loop, one elected lane issues: see elect_one in common.cuh)
            uint32_t c = 0;
            int step = 0;
            const uint64_t wdesc0 = tc_desc(ring_u32, 128, 256), xdesc_hi0 = VRP_XDESC_H(x_hi, 0), xdesc_lo0 = VRP_XDESC_H(x_lo, 0);
            for (;; ++step) {
                mbar_wait(bar_x, step & 1);
                tc_fence_after();
                if (*stop_s) break;
#pragma unroll 1
                for (int u = 0; u < kHUsesQ; ++u, ++c) {          // q^T[unit, row] = W_q^T . h^T  -> columns kQCol..
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
                    tc_fence_after();
                    const uint64_t w0 = wdesc0 + (uint64_t)(st * (kSlabBytes >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const int kc = 2 * u + kk;
                            const uint64_t w_hi = w0 + (uint64_t)(kk * (8192 >> 4)), w_lo = w_hi + (4096 >> 4);
                            const uint64_t dx_hi = xdesc_hi0 + (uint64_t)(kc * 16), dx_lo = xdesc_lo0 + (uint64_t)(kc * 16);
                            tc_mma_f16(tmem_base + kQCol, w_lo, dx_hi, idescL, kc > 0);
                            tc_mma_f16(tmem_base + kQCol, w_hi, dx_lo, idescL, 1);
                            tc_mma_f16(tmem_base + kQCol, w_hi, dx_hi, idescL, 1);
                        }
                        tc_commit(bar_empty + 8 * st);
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(bar_q);
                __syncwarp();
#pragma unroll 1
                for (int u = 0; u < kHUsesG; ++u, ++c) {          // gates^T of the NEXT step, h part
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
#ifdef VRP_THROTTLE_GATES
                    if (u > 0) {   // throttle: one ring stage in the tensor-pipe queue, the score MMAs queue behind it
                        const uint32_t cp = c - 1;
                        mbar_wait(bar_empty + 8 * (cp % STAGES), (cp / STAGES) & 1);
                    }
#endif
                    tc_fence_after();
                    const int kc = u >> 1, mp = u & 1;
                    const uint64_t dx_hi = xdesc_hi0 + (uint64_t)(kc * 16), dx_lo = xdesc_lo0 + (uint64_t)(kc * 16);
                    const uint64_t w0 = wdesc0 + (uint64_t)(st * (kSlabBytes >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi) {
                            const uint64_t w_hi = w0 + (uint64_t)(mi * (8192 >> 4)), w_lo = w_hi + (4096 >> 4);
                            const uint32_t d = tmem_base + kGCol + RB * (2 * mp + mi);
                            tc_mma_f16(d, w_lo, dx_hi, idescL, kc > 0);
                            tc_mma_f16(d, w_hi, dx_lo, idescL, 1);
                            tc_mma_f16(d, w_hi, dx_hi, idescL, 1);
                        }
                        tc_commit(bar_empty + 8 * st);
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(bar_g);
                __syncwarp();
            }
            if (step > 0) mbar_wait(bar_g, (step - 1) & 1);       // the last gates GEMM must have drained before TMEM is freed
        } else if (warp >= 22) {
            // ---- score-MMA issuer of group g: one tile per "F ready"; its quarters go to the group's three TMEM buffers in
            //      rotation, each as soon as the four score warps that read the buffer's previous content have copied it out
            const int g = warp - 22;
            const uint32_t bfull = bar_grp + 64 * g, bfree = bfull + 24, bfrdy = bfull + 48;
            const uint32_t fbase = smem_u32(smem + L.F) + g * 4 * kC1TileBytes;
            uint32_t buf = 0, use = 0;
            for (uint32_t tc = 0;; ++tc) {
                const uint32_t fb = tc & 1;
                mbar_wait(bfrdy + 8 * fb, (tc >> 1) & 1);
                if (*stop_s) break;
                const uint64_t dfh = tc_desc(fbase + fb * 2 * kC1TileBytes, 128, 256), dfl = tc_desc(fbase + fb * 2 * kC1TileBytes + kC1TileBytes, 128, 256);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    if (use > 0) mbar_wait(bfree + 8 * buf, (use - 1) & 1);
                    tc_fence_after();
                    const uint64_t dbh = tc_desc(bc_u32 + q * 1024, 128, 256), dbl = dbh + (4096 >> 4);
                    const uint32_t d = tmem_base + 96 * g + 32 * buf;
                    if (elect_one()) {
                        tc_mma_tf32(d, dfl, dbh, idescQ, 0);
                        tc_mma_tf32(d, dfh, dbl, idescQ, 1);
                        tc_mma_tf32(d, dfh, dbh, idescQ, 1);
                        tc_commit(bfull + 8 * buf);
                    }
                    __syncwarp();
                    if (++buf == 3) { buf = 0; ++use; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 16) {
        // ============ aux warps: the per-row tail (choice + Env.step), streamed behind the score tiles ====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int ax = warp - 16;                                // rows ax, ax + 4, ...: one row of every 4-row chunk
        unsigned long long n_rowsteps = 0;
#ifdef VRPB200_PHASE_CLOCKS
        long long pc_[8], pc_last_ = clock64();
        for (int i = 0; i < 8; ++i) pc_[i] = 0;
#endif
        int t = 0;
#pragma unroll 1
        for (;; ++t) {
            mbar_wait(bar_x, t & 1);                             // T0 of step t
            if (*stop_s) break;
            mbar_wait(bar_items, t & 1);                         // item lists, item sets and row offsets of step t are published
            VRP_CLK(0);
#pragma unroll 1
            for (int j = 0; j < NCH; ++j) {
                const int lr = 4 * j + ax;
                const bool rvalid = lr < IBv;
                float* rs = rowst + lr * kRS5;
                const long long row = b0 + lr;
                const bool was_done = reinterpret_cast<int*>(rs)[8] != 0;
                if (rvalid && was_done && lane == 0) {
                    a.idx_out[(long long)t * R + row] = n_cust;
                    if (a.logprob_out) a.logprob_out[(long long)t * R + row] = 0.0f;
                }
                if (rvalid && !was_done) {
                    if (lane == 0) ++n_rowsteps;
                    const float load = rs[0], time = TW ? rs[1] : 0.0f;
                    const float4* frow = gfeat4 + (size_t)lr * N;
                    float* drow = dem + lr * NP;
                    float* ulog = logits + lr * NP;
                    const unsigned char* mk = maskb + lr * NP;
                    const unsigned char* it = items + lr * NP;
                    const int fl = flag_s[lr], na = nact_s[lr];
                    const bool fast = fastk && fl == 0;
                    // (0) the records of this lane's items (the chosen node is one of them): issued before the wait for the
                    //     logits, used after the choice
                    float dk[4];
                    float4 fk[4];
                    bool isit[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int n = lane + 32 * k;
                        isit[k] = (ibits[lr * 4 + k] >> lane) & 1;
                        dk[k] = (n < N) ? drow[n] : 0.0f;
                        fk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (isit[k] && n < n_cust) fk[k] = ldcg4_volatile(frow + n);
                    }
                    if (na > 0) {   // wait for every tile that holds items of this row
                        const int i0 = off_s[lr], tl = (i0 + na - 1) >> 7;
                        for (int tile = i0 >> 7; tile <= tl; ++tile) mbar_wait(bar_tile + 8 * tile, t & 1);
                    }
                    VRP_CLK(1);
                    int idx;
                    float psel = 1.0f;
                    if (fast) {
                        // ---- fast choice (plain greedy decoding): arg-max of the logits over the items (first index on exact
                        //      ties, as argmax(prob) of the reference, attention_agent.py:146-147); only a runner-up within
                        //      1e-5 BELOW the maximum could round to the same probability: then the probabilities are evaluated
                        float u[4];
                        float mx = -INFINITY, second = -INFINITY;
                        int mxi = 0x7fffffff;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            u[k] = isit[k] ? ulog[lane + 32 * k] : -INFINITY;
                            if (u[k] > mx) { second = mx; mx = u[k]; mxi = lane + 32 * k; }
                            else second = fmaxf(second, u[k]);
                        }
#pragma unroll
                        for (int o = 16; o; o >>= 1) {
                            const float ov = __shfl_xor_sync(kFull, mx, o), os = __shfl_xor_sync(kFull, second, o);
                            const int oi = __shfl_xor_sync(kFull, mxi, o);
                            second = fmaxf(second, os);
                            if (ov != mx) second = fmaxf(second, fminf(mx, ov));
                            if (ov > mx || (ov == mx && oi < mxi)) { mx = ov; mxi = oi; }
                        }
                        if (second > mx - 1e-5f) {
                            // near-tie: exact probabilities, summed in the fixed order of the exact path (lane = node % 32 adds its
                            // nodes in sequence, then the xor butterfly); masked nodes have probability exactly 0
#ifdef VRPB200_PHASE_CLOCKS
                            pc_[7] += 1;
#endif
                            float se = 0.0f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) se = __fadd_rn(se, (u[k] == -INFINITY) ? 0.0f : expf(u[k] - mx));
#pragma unroll
                            for (int o = 16; o; o >>= 1) se = __fadd_rn(se, __shfl_xor_sync(kFull, se, o));
                            const float lse = logf(se);
                            float bestp = -1.0f;
                            int besti = 0x7fffffff;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float p = (u[k] == -INFINITY) ? 0.0f : expf((u[k] - mx) - lse);
                                if (lane + 32 * k < N && p > bestp) { bestp = p; besti = lane + 32 * k; }
                            }
#pragma unroll
                            for (int o = 16; o; o >>= 1) {
                                const float ov = __shfl_xor_sync(kFull, bestp, o);
                                const int oi = __shfl_xor_sync(kFull, besti, o);
                                if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                            }
                            mxi = besti;
                        }
                        idx = mxi;
                    } else {
                        // ---- exact choice: masked logits (decode_step.py:231-232), log-softmax / probabilities (:125-126), greedy
                        //      arg-max of the probabilities or inverse-CDF sampling (attention_agent.py:146-167), traces.  The mask
                        //      counts of the row were written by the score warp that built its item list.
                        const long long tro = ((long long)t * R + row) * N;
                        float u[4];
                        float mx = -INFINITY, second = -INFINITY;
                        int mxi = 0x7fffffff;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int n = lane + 32 * k;
                            u[k] = -INFINITY;
                            if (n < N) {
                                const int mm = mk[n];
                                if (fl || mm == 0) {
                                    float v = ulog[n];
                                    if (a.use_tanh) v = a.tanh_C * tanhf(v);
                                    if (a.mask_pointer) v = __fsub_rn(v, __fmul_rn(kBig, (float)mm));
                                    u[k] = v;
                                }
                                if (a.tr_logits) a.tr_logits[tro + n] = u[k];
                                if (a.tr_masks) a.tr_masks[tro + n] = (float)mm;
                            }
                            if (u[k] > mx) { second = mx; mx = u[k]; mxi = n; }
                            else second = fmaxf(second, u[k]);
                        }
#pragma unroll
                        for (int o = 16; o; o >>= 1) {
                            const float ov = __shfl_xor_sync(kFull, mx, o), os = __shfl_xor_sync(kFull, second, o);
                            const int oi = __shfl_xor_sync(kFull, mxi, o);
                            second = fmaxf(fmaxf(second, os), fminf(mx, ov));
                            if (ov > mx || (ov == mx && oi < mxi)) { mx = ov; mxi = oi; }
                        }
                        // exp is monotone up to a few ulp: with the runner-up more than 1e-5 below the maximum the arg-max of the
                        // logits IS argmax(prob); the soft-max is then only evaluated if somebody asked for its values
                        const bool need_softmax = a.mode != 0 || tracing || a.logprob_out != nullptr || !(second < mx - 1e-5f);
                        idx = mxi;
                        if (need_softmax) {
                            float se = 0.0f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) se = __fadd_rn(se, (u[k] == -INFINITY) ? 0.0f : expf(u[k] - mx));
#pragma unroll
                            for (int o = 16; o; o >>= 1) se = __fadd_rn(se, __shfl_xor_sync(kFull, se, o));
                            const float lse = logf(se);
                            float p[4];
                            float bestp = -1.0f;
                            int besti = 0x7fffffff;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int n = lane + 32 * k;
                                p[k] = (u[k] == -INFINITY) ? 0.0f : expf((u[k] - mx) - lse);
                                if (n < N) {
                                    if (p[k] > bestp) { bestp = p[k]; besti = n; }
                                    if (a.tr_probs) a.tr_probs[tro + n] = p[k];
                                }
                            }
                            if (a.mode == 0) {   // greedy: argmax(prob), first maximal index
#pragma unroll
                                for (int o = 16; o; o >>= 1) {
                                    const float ov = __shfl_xor_sync(kFull, bestp, o);
                                    const int oi = __shfl_xor_sync(kFull, besti, o);
                                    if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                                }
                                idx = besti;
                                psel = bestp;
                            } else {             // inverse-CDF sampling with one (per-row) retry (attention_agent.py:148-167)
                                __syncwarp();
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (lane + 32 * k < N) ulog[lane + 32 * k] = p[k];
                                __syncwarp();
                                int sel = -1;
                                if (lane == 0) {
                                    const long long o = (long long)t * R + row;
                                    int d0_, d1_;
                                    draw_range(a, t, d0_, d1_);   // per-row retry, or the reference's whole-batch redraw (attention_agent.py:165-167)
                                    for (int draw = d0_; draw < d1_ && sel < 0; ++draw) {
                                        const float* ubuf = draw ? a.uniforms_retry : a.uniforms;
                                        const float uu = ubuf ? ubuf[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                                        // tf.cumsum: sequential fp32; masked nodes add exactly 0, so when the row has an item
                                        // list the scan runs over the items only (ascending node ids) - no early exit, the
                                        // first node whose running sum exceeds u is kept by a select
                                        float cum = 0.0f;
                                        int first = -1;
                                        if (fl) {
                                            for (int n = 0; n < N; ++n) {
                                                cum = __fadd_rn(cum, ulog[n]);
                                                first = (first < 0 && cum > uu) ? n : first;
                                            }
                                        } else {
                                            for (int k = 0; k < na; ++k) {
                                                const int n = it[k];
                                                cum = __fadd_rn(cum, ulog[n]);
                                                first = (first < 0 && cum > uu) ? n : first;
                                            }
                                        }
                                        sel = first;
                                    }
                                    if (sel < 0 && a.fail_steps && d0_ == 0) a.fail_steps[t] = 1;   // some row of the batch needs the redraw: the host re-runs with this step redrawn
                                    if (sel < 0) sel = 0;
                                }
                                idx = __shfl_sync(kFull, sel, 0);
                                psel = ulog[idx];
                                __syncwarp();
                            }
                        }
                    }
                    VRP_CLK(2);
                    // ---- Env.step (vrptw_utils.py:254-325): the record of the chosen node comes from the lane that owns it
                    //      (a chosen customer is an item, whose record was fetched above; the depot's is part of the row state)
                    float sx, sy, sb, se_tw, d_idx;
                    {
                        const int ko = idx >> 5, lo = idx & 31;
                        const float4 fo = ko == 0 ? fk[0] : ko == 1 ? fk[1] : ko == 2 ? fk[2] : fk[3];
                        const float dko = ko == 0 ? dk[0] : ko == 1 ? dk[1] : ko == 2 ? dk[2] : dk[3];
                        const bool iko = ko == 0 ? isit[0] : ko == 1 ? isit[1] : ko == 2 ? isit[2] : isit[3];
                        sx = __shfl_sync(kFull, fo.x, lo); sy = __shfl_sync(kFull, fo.y, lo);
                        sb = __shfl_sync(kFull, fo.z, lo); se_tw = __shfl_sync(kFull, fo.w, lo);
                        d_idx = __shfl_sync(kFull, dko, lo);
                        const int have = __shfl_sync(kFull, (int)iko, lo);
                        if (idx == n_cust) { sx = rs[7]; sy = rs[14]; sb = rs[16]; se_tw = rs[17]; }
                        else if (!have) {   // (a node that is not an item can only be chosen when every node is masked)
                            const float4 fv = __ldcg(frow + idx);
                            sx = fv.x; sy = fv.y; sb = fv.z; se_tw = fv.w;
                        }
                        if (!TW) { sb = 0.0f; se_tw = 0.0f; }
                    }
                    const float px = rs[2], py = rs[3];
                    float nz = rs[13];                       // customers with demand left: sum(demand) > 0  <=>  nz > 0
                    const float d_sat = fminf(d_idx, load);
                    const float d_new = __fsub_rn(d_idx, d_sat);
                    float nload = __fsub_rn(load, d_sat);
                    const float ts = dist_rn(px, py, sx, sy);
                    float ntime = 0.0f;
                    if (TW) ntime = fmaxf(__fadd_rn(time, ts), sb);
                    const float dep = (idx == n_cust) ? 1.0f : 0.0f;
                    nload = __fadd_rn(__fmul_rn(nload, 1.0f - dep), __fmul_rn(dep, a.capacity));
                    if (TW) ntime = __fmul_rn(ntime, 1.0f - dep);
                    if (d_idx != 0.0f && d_new == 0.0f) nz -= 1.0f;
                    const int done_now = (!a.full && a.mask_pointer && idx == n_cust && nz == 0.0f) ? 1 : 0;
                    __syncwarp();
                    if (lane == 0) {
                        drow[idx] = d_new;
                        rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                        if (t == 0) { rs[5] = sx; rs[6] = sy; }
                        else rs[4] = __fadd_rn(rs[4], ts);
                        reinterpret_cast<int*>(rs)[8] = done_now;
                        rs[9] = sx; rs[10] = sy; rs[11] = sb; rs[12] = se_tw; rs[13] = nz; rs[15] = dep;
                        a.idx_out[(long long)t * R + row] = idx;
                        if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
                    }
                    VRP_CLK(3);
                }
                VRP_CLK(4);
            }
            named_bar_arrive(7, 640);                            // every row of this warp has moved (the score warps are parked there)
        }
        if (a.stats) {
            if (lane == 0 && n_rowsteps) atomicAdd(a.stats + 1, n_rowsteps);
#ifdef VRPB200_PHASE_CLOCKS
            if (tid == 16 * 32)
                for (int i = 0; i < 8; ++i) atomicAdd(a.stats + 20 + i, (unsigned long long)pc_[i]);
#endif
        }
    } else {
        // =============================== 16 score warps =================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const int ub = warp >> 2;
        const int g = warp >> 3;                                         // score group
        const int half = ub & 1;                                         // hidden units 64 half .. 64 half + 63 (quarters half, half + 2)
        const int m = 32 * sp + lane;                                    // item of the tile
        const int unit = 32 * sp + lane;                                 // hidden unit in the LSTM / query epilogues
        unsigned char* fgrp = ftile + g * 4 * kC1TileBytes;
        const uint32_t bfull = bar_grp + 64 * g, bfree = bfull + 24, bfrdy = bfull + 48, bpsum = bar_psum + 8 * g, bpack = bar_pack + 8 * g;
        float* ps = psum + g * 256;
        uint32_t c1_nq = 0;                                              // tiles of this group so far (all steps)
        unsigned long long n_eval = 0;
#ifdef VRPB200_PHASE_CLOCKS
        long long pc_[16], pc_last_ = clock64();
        for (int i = 0; i < 16; ++i) pc_[i] = 0;
#endif
        // LSTM epilogue of the rows 4j .. 4j+3 (this warp: units 32 sp .. 32 sp + 31; chunk j belongs to team j % 4 = ub):
        // gates = gates^T(TMEM, h part) + x part + bias; c, h'; h' -> X (hi | lo)
        auto lstm_chunk = [&](int j, bool first) {
            const int r0 = 4 * j;
            float gt[4][4], cc[4];
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                if (!first) tc_ld4(tmem_lane + kGCol + RB * mm + r0, gt[mm]);
                else { gt[mm][0] = gt[mm][1] = gt[mm][2] = gt[mm][3] = 0.0f; }
            }
            if (!first) tc_ld4(tmem_lane + kCCol + r0, cc);
            else { cc[0] = cc[1] = cc[2] = cc[3] = 0.0f; }
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* rs = rowst + (r0 + i) * kRS5;
                const float xf[4] = {rs[9], rs[10], rs[11], rs[12]};
                float gg[4];
#pragma unroll
                for (int mm = 0; mm < 4; ++mm) {
                    float s = wxs[(16 + mm) * kH + unit];
#pragma unroll
                    for (int k = 0; k < FD; ++k) s = fmaf(xf[k], wxs[(k * 4 + mm) * kH + unit], s);
                    gg[mm] = gt[mm][i] + s;
                }
                const float cn = __fadd_rn(__fmul_rn(cc[i], sigmoid_fast(gg[2] + a.forget_bias)),
                                           __fmul_rn(sigmoid_fast(gg[0]), tanh_fast(gg[1])));
                cc[i] = cn;
                const float hn = __fmul_rn(tanh_fast(cn), sigmoid_fast(gg[3]));
                hi[i] = __half2float(__float2half_rn(hn));      // FP16 two-term split (common.cuh): hi, then lo = fp16(h - hi) at the store
                lo[i] = hn - hi[i];
            }
            tc_st4(tmem_lane + kCCol + r0, cc);
            transpose4(hi, lane);
            transpose4(lo, lane);
            const int rr = r0 + (lane & 3);
            // [row/8][k/8][row%8][k%8] halves: this lane holds units 4 (unit/4) .. +3 of row rr = 8 bytes
            const int off = (rr >> 3) * 2048 + (unit >> 3) * 128 + (rr & 7) * 16 + ((unit >> 2) & 1) * 8;
            const __half2 h01 = __floats2half2_rn(hi[0], hi[1]), h23 = __floats2half2_rn(hi[2], hi[3]);
            const __half2 l01 = __floats2half2_rn(lo[0], lo[1]), l23 = __floats2half2_rn(lo[2], lo[3]);
            *reinterpret_cast<uint2*>(xt + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(xt + RB * kH * 2 + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        };
        // Env mask of a row after its last move (vrptw_utils.py:328-348) -> 128-bit item set, item list, count; one warp per
        // row, lane owns nodes lane + 32 k.  Plain greedy decoding keeps no mask counts (un-masked / masked only) unless a row
        // has no un-masked node left; then (and in every other mode) the counts are written for the aux warps.
        // One code path, kept small: this runs once per step and must not evict the score loop from the instruction cache.
        auto build_masks = [&](int lr, const float4 (&fk)[4], bool exact) {
            const float* rs = rowst + lr * kRS5;
            const bool on = lr < IBv && reinterpret_cast<const int*>(rs)[8] == 0;
            const float nload = rs[0], ntime = TW ? rs[1] : 0.0f, sx = rs[2], sy = rs[3], nz = rs[13], dep = rs[15];
            const float* drow = dem + lr * NP;
            unsigned char* mk = maskb + lr * NP;
            unsigned bits[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = lane + 32 * k;
                int mm = 1;
                if (n < N && on) {
                    if (n == n_cust) mm = (nz > 0.0f && dep != 0.0f) ? 1 : 0;
                    else {
                        mm = (drow[n] == 0.0f) + (nload == 0.0f);
                        if (TW && (exact || mm == 0)) mm += (__fadd_rn(ntime, dist_rn(sx, sy, fk[k].x, fk[k].y)) > fk[k].w);
                    }
                    if (exact) mk[n] = (unsigned char)mm;
                }
                bits[k] = __ballot_sync(kFull, mm == 0);
            }
            return publish_items(lr, bits, on);
        };
        float4 fnext[4];                                                 // mask records of the warp's next row (software pipeline)
        // the records a row's time-window test can need: its customers with demand left (a superset: demand only decreases)
        auto mask_records = [&](int lr, float4 (&fk)[4], bool all) {
            const bool on = lr < IBv && reinterpret_cast<const int*>(rowst + lr * kRS5)[8] == 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = lane + 32 * k;
                fk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (TW && on && lr < RB && n < n_cust && (all || dem[lr * NP + n] != 0.0f)) fk[k] = ldcg4_volatile(gfeat4 + (size_t)lr * N + n);
            }
        };
        int t = 0;
#pragma unroll 1
        for (;; ++t) {
            // ================= T0': every row has made its move of step t - 1 (the aux warps arrive on this barrier) =========
            if (t > 0) named_bar(7, 640);
            {
                const int d0 = (lane < IBv) ? reinterpret_cast<const int*>(rowst + lane * kRS5)[8] : 1;
                const int d1 = (lane + 32 < IBv) ? reinterpret_cast<const int*>(rowst + (lane + 32) * kRS5)[8] : 1;
                const bool stop = t > 0 && (__all_sync(kFull, d0 && d1) || t == T);   // (the same in every warp)
                if (stop) {
                    if (tid == 0) {
                        *tend_s = t; *stop_s = 1;
                        __threadfence_block();
                        mbar_arrive(bar_x);                              // feeder, MMA issuer and aux warps leave
                    }
                    break;
                }
            }
            VRP_CLK(0);
            if (t > 0) mask_records(warp, fnext, !fastk);                // in flight during the LSTM epilogue
            // ================= LSTM epilogue -> h'(t) (step 0: h_0 = c_0 = 0, decoder input = the depot, attention_agent.py:126-134)
            if (t > 0) {
                mbar_wait(bar_g, (t - 1) & 1);                           // gates^T(t) = W_h^T h'(t-1)^T: issued a whole step ago
                tc_fence_after();
            }
#pragma unroll 1
            for (int j = ub; j < NCH; j += 4) lstm_chunk(j, t == 0);
            fence_proxy_async();
            tc_fence_before();
            named_bar(5, 512);
            if (tid == 0) mbar_arrive(bar_x);                            // T0: h'(t) is in place - feeder, MMA issuer, aux warps
            VRP_CLK(1);
            // ================= Env mask -> item lists of this step (the query GEMM is running) ============================
            if (t > 0) {
                bool exact = !fastk;
#pragma unroll 1
                for (int lr = warp; lr < RB;) {
                    float4 fk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) fk[k] = fnext[k];
                    if (!exact || !fastk) mask_records(lr + 16, fnext, !fastk);   // the next row's records, in flight during this row
                    const int fl = build_masks(lr, fk, exact);
                    // no un-masked node: the mask COUNTS decide the choice - the same row once more, exactly (rare)
                    if (fastk && fl && !exact) { exact = true; mask_records(lr, fnext, true); }
                    else {
                        if (fastk && exact) mask_records(lr + 16, fnext, false);
                        exact = !fastk; lr += 16;
                    }
                }
            }
            VRP_CLK(9);
            named_bar(5, 512);                                           // item lists final
            // ================= row offsets of this step's item lists (every warp, redundantly; warp 0 publishes them) ======
            int m_tot;
            {
                const int n0 = (lane < RB) ? nact_s[lane] : 0, n1 = (lane + 32 < RB) ? nact_s[lane + 32] : 0;
                int s0 = n0, s1 = n1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t0 = __shfl_up_sync(kFull, s0, o), t1 = __shfl_up_sync(kFull, s1, o);
                    if (lane >= o) { s0 += t0; s1 += t1; }
                }
                const int tot0 = __shfl_sync(kFull, s0, 31);
                m_tot = tot0 + __shfl_sync(kFull, s1, 31);
                if (warp == 0) {
                    if (lane < RB) off_s[lane] = s0 - n0;
                    if (lane + 32 < RB) off_s[lane + 32] = tot0 + s1 - n1;
                    if (lane == 0) off_s[RB] = m_tot;
                }
                if (tid < 64 && tid >= ((m_tot + 127) >> 7)) mbar_arrive_n(bar_tile + 8 * tid, 4);   // tiles that do not exist in this step
            }
            named_bar(5, 512);                                           // off_s is in place
            VRP_CLK(15);
            if (tid == 0) mbar_arrive(bar_items);                        // the aux warps may look at item lists and offsets
            const int ntiles = (m_tot + 127) >> 7;
            // item `it` of the step -> (row, node): binary search in the row offsets, then the row's node list
            auto load_item = [&](int tile, int& r, int& n, bool& valid, float (&f)[5], bool feats) {
                const int it = tile * 128 + m;
                valid = tile < ntiles && it < m_tot;
                r = 0; n = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) f[k] = 0.0f;
                if (valid) {
                    int lo = 0, hi = RB;
#pragma unroll
                    for (int s = 0; s < 6; ++s) {
                        const int mid = (lo + hi) >> 1;
                        if (off_s[mid] <= it) lo = mid; else hi = mid;
                    }
                    r = lo;
                    n = items[r * NP + (it - off_s[r])];
                    if (feats) {
                        const float4 fv = ldcg4_volatile(gfeat4 + (size_t)r * N + n);
                        f[0] = fv.x; f[1] = fv.y;
                        if (TW) { f[2] = fv.z; f[3] = fv.w; }
                        f[4] = dem[r * NP + n];
                    }
                }
            };
            auto store_F = [&](unsigned char* fhi, const float (&f)[5]) {
                float hi[5], lo[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) { hi[k] = tf32_rna(f[k]); lo[k] = f[k] - hi[k]; }
                const int off = (m >> 3) * 256 + (m & 7) * 16;
                *reinterpret_cast<float4*>(fhi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(fhi + off + 128) = make_float4(hi[4], 0.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4*>(fhi + kC1TileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(fhi + kC1TileBytes + off + 128) = make_float4(lo[4], 0.0f, 0.0f, 0.0f);
                fence_proxy_async();
            };
            // 32 hidden units of one item: acc += sum_h w_h / (1 + 2^(E_h + base'_h)), w = -2 v; ONE reciprocal per 4 elements:
            //   (w0 t1 + w1 t0) t2 t3 + (w2 t3 + w3 t2) t0 t1  over  t0 t1 t2 t3,   t = 1 + 2^min(arg, 30)
            // (<= 2^30 + 1: the product cannot overflow, and tanh is 1.0f exactly from arg = 26 on).  8 elements at a time as four
            // pairs; the two quads that share one reciprocal each are the .x and the .y lanes of the pairs (packed FP32x2 algebra)
            auto quarter_arith = [&](const uint32_t (&cur)[32], const float* bq, const float* vq, float& acc0, float& acc1) {
                f32x2 acc = pk2(acc0, acc1);
                const f32x2 one = pk2(1.0f, 1.0f);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const float4 ba = *reinterpret_cast<const float4*>(bq + 8 * g8), bb = *reinterpret_cast<const float4*>(bq + 8 * g8 + 4);
                    const float4 va = *reinterpret_cast<const float4*>(vq + 8 * g8), vb = *reinterpret_cast<const float4*>(vq + 8 * g8 + 4);
                    f32x2 P[4];
                    P[0] = add2(pk2(__uint_as_float(cur[8 * g8 + 0]), __uint_as_float(cur[8 * g8 + 1])), pk2(ba.x, ba.y));
                    P[1] = add2(pk2(__uint_as_float(cur[8 * g8 + 2]), __uint_as_float(cur[8 * g8 + 3])), pk2(ba.z, ba.w));
                    P[2] = add2(pk2(__uint_as_float(cur[8 * g8 + 4]), __uint_as_float(cur[8 * g8 + 5])), pk2(bb.x, bb.y));
                    P[3] = add2(pk2(__uint_as_float(cur[8 * g8 + 6]), __uint_as_float(cur[8 * g8 + 7])), pk2(bb.z, bb.w));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float x, y;
                        upk2(P[i], x, y);
                        P[i] = add2(pk2(ex2_approx(fminf(x, 30.0f)), ex2_approx(fminf(y, 30.0f))), one);
                    }
                    const f32x2 V0 = pk2(va.x, va.y), V1 = pk2(va.z, va.w), V2 = pk2(vb.x, vb.y), V3 = pk2(vb.z, vb.w);
                    const f32x2 p01 = mul2(P[0], P[1]), p23 = mul2(P[2], P[3]);
                    const f32x2 n01 = fma2(V0, P[1], mul2(V1, P[0])), n23 = fma2(V2, P[3], mul2(V3, P[2]));
                    const f32x2 num = fma2(n01, p23, mul2(n23, p01));
                    float dx, dy;
                    upk2(mul2(p01, p23), dx, dy);
                    acc = fma2(num, pk2(rcp_approx(dx), rcp_approx(dy)), acc);
                }
                upk2(acc, acc0, acc1);
            };
            // ================= C1 set-up: the F tiles of the group's first two tiles, staged while the query GEMM runs =====
            int r_cur = 0, n_cur = 0, r_nx = 0, n_nx = 0;
            bool v_cur = false, v_nx = false;
            const bool scoring = g < ntiles;
            if (scoring) {
                float fc[5], f2[5];
                load_item(g, r_cur, n_cur, v_cur, fc, half == 0);
                load_item(g + NG, r_nx, n_nx, v_nx, f2, half == 0);
                if (half == 0) {
                    store_F(fgrp + (c1_nq & 1) * 2 * kC1TileBytes, fc);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfrdy + 8 * (c1_nq & 1));
                    if (g + NG < ntiles) {
                        store_F(fgrp + ((c1_nq + 1) & 1) * 2 * kC1TileBytes, f2);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bfrdy + 8 * ((c1_nq + 1) & 1));
                    }
                }
            }
            VRP_CLK(11);
            // ================= query epilogue: base'[row][unit] = 2log2e (q + b + load g_ld + time g_t) ===================
            const float bqv = __ldg(a.att.bq + unit), glv = __ldg(a.att.gl + unit), gtv = __ldg(a.att.gt + unit);
            if (warp == 0) mbar_wait(bar_q, t & 1);                      // one warp polls, the others are parked
            named_bar(5, 512);
            tc_fence_after();
            VRP_CLK(2);
#pragma unroll 1
            for (int c4 = 0; c4 < RPT / 4; ++c4) {
                const int r0 = ub * RPT + 4 * c4;
                float q4[4];
                tc_ld4(tmem_lane + kQCol + r0, q4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* rs = rowst + (r0 + i) * kRS5;
                    float s = fmaf(rs[0], glv, q4[i] + bqv);
                    if (TW) s = fmaf(rs[1], gtv, s);
                    qq[(r0 + i) * kH + unit] = s * kTwoLog2e;
                }
            }
            tc_fence_before();
            named_bar(5, 512);
            tc_fence_after();
            VRP_CLK(3);
            // ================= C1: scores (the aux warps stream the per-row tails behind the tiles) ========================
            // c1_nq counts the tiles of this group (all steps): tile k uses F tile k & 1; its quarter q is the (4k + q)-th use of
            // the group's TMEM buffers, which rotate over three.  Warps of half 0 gather the items, write F and take quarters 0, 2;
            // warps of half 1 take quarters 1, 3 of the same items and hand their partial sum over through shared memory.
            if (scoring) {
                int r_n2 = 0, n_n2 = 0;
                bool v_n2 = false;
                float fn[5];
#pragma unroll 1
                for (int tile = g; tile < ntiles; tile += NG) {
                const float* brow = qq + r_cur * kH;
                const uint32_t par = c1_nq & 1;
                float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                for (int qi = 0; qi < 2; ++qi) {
                    const int q = 2 * qi + half;
                    const uint32_t cq = 4 * c1_nq + q, buf = cq % 3, bpar = (cq / 3) & 1;
                    VRP_CLK(4);
                    mbar_wait(bfull + 8 * buf, bpar);
                    tc_fence_after();
                    VRP_CLK(5);
                    uint32_t cur[32];
                    tc_ld32_issue(tmem_lane + 96 * g + 32 * buf, cur);
                    tc_ld_wait32(cur);
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(bfree + 8 * buf);     // the issuer may send another quarter here
                    __syncwarp();
                    VRP_CLK(6);
                    // the items two tiles ahead: row look-up and feature loads are scheduled into the arithmetic below
                    if (qi == 0) load_item(tile + 2 * NG, r_n2, n_n2, v_n2, fn, half == 0);
                    quarter_arith(cur, brow + 32 * q, vsm + 32 * q, acc0, acc1);
                    VRP_CLK(7);
                }
                if (half == 1) {
                    if (c1_nq > 0) mbar_wait(bpack, (c1_nq - 1) & 1);   // the other half has taken the previous tile's sums
                    ps[par * 128 + m] = acc0 + acc1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bpsum);
                } else {
                    mbar_wait(bpsum, par);      // (also: quarters 1 and 3 of this tile have been read, its F tile is free)
                    const float psv = ps[par * 128 + m];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bpack);
                    // u = sum_h v_h tanh(.) = sum_h v_h - sum_h 2 v_h / (1 + 2^arg)
                    if (v_cur) logits[r_cur * NP + n_cur] = (a.sumv + (acc0 + acc1)) + psv;
                    n_eval += v_cur ? 1 : 0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tile + 8 * tile);   // the rows of this tile may be finished (release: logits visible)
                    if (tile + 2 * NG < ntiles) {
                        store_F(fgrp + par * 2 * kC1TileBytes, fn);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bfrdy + 8 * par);
                    }
                }
                VRP_CLK(8);
                ++c1_nq;
                r_cur = r_nx; n_cur = n_nx; v_cur = v_nx;
                r_nx = r_n2; n_nx = n_n2; v_nx = v_n2;
                }
            }
            VRP_CLK(10);
        }
        // wake the score-MMA issuers (they wait for the F tile of the group's next tile)
        if (half == 0 && lane == 0) mbar_arrive(bfrdy + 8 * (c1_nq & 1));
        if (a.stats) {
            if (n_eval) atomicAdd(a.stats + 0, n_eval);
#ifdef VRPB200_PHASE_CLOCKS
            if (tid == 64)
                for (int i = 0; i < 16; ++i) atomicAdd(a.stats + 4 + i, (unsigned long long)pc_[i]);
#endif
        }
    }

    // ============ every warp: outputs =================================================================================
    tc_fence_before();
    __syncthreads();
    const int t_end = *tend_s;
    if (warp < 16) {
        for (int lr = warp; lr < IBv; lr += 16) {
            const long long row = b0 + lr;
            for (int tt = t_end + lane; tt < T; tt += 32) {
                a.idx_out[(long long)tt * R + row] = n_cust;
                if (a.logprob_out) a.logprob_out[(long long)tt * R + row] = 0.0f;
            }
            const float* rs = rowst + lr * kRS5;
            if (lane == 0) a.cost_out[row] = __fadd_rn(rs[4], dist_rn(rs[2], rs[3], rs[5], rs[6]));
            for (int n = lane; n < N; n += 32) {
                if (a.fin_demand) a.fin_demand[row * N + n] = dem[lr * NP + n];
                if (a.fin_mask) {   // Env.mask after the last move (vrptw_utils.py:328-348)
                    float nx = 0.f, ny = 0.f, ne = 0.f;
                    if (TW && n < n_cust) { const float4 fv = __ldcg(gfeat4 + (size_t)lr * N + n); nx = fv.x; ny = fv.y; ne = fv.w; }
                    a.fin_mask[row * N + n] = (float)node_mask<TW>(n, n_cust, dem[lr * NP + n], rs[0], TW ? rs[1] : 0.0f, rs[2], rs[3], nx, ny, ne, rs[15], rs[13]);
                }
            }
            if (lane == 0) {
                if (a.fin_load) a.fin_load[row] = rs[0];
                if (a.fin_time) a.fin_time[row] = rs[1];
            }
        }
        if (tid == 0 && a.steps_out) a.steps_out[blockIdx.x] = t_end;
        if (a.stats && tid == 0) {
            atomicAdd(a.stats + 2, (unsigned long long)t_end * RB);
            atomicAdd(a.stats + 3, 1ull);
        }
    }
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem_base, kTcCols);
}

}  // namespace vrpb200
