// rollout4_kernel: the fused rollout (RLAgent.build_model of model/attention_agent.py:84-240, greedy and stochastic) with
// the per-step work organised around the two resources that bind it, the MUFU pipe and the issue slots.
//
//  * 16 compute warps + 4 single-thread service warps (weight feeder, LSTM/query MMA issuer, two score-MMA issuers); the
//    service threads wait on mbarriers with a suspend hint (asleep in hardware, no issue slots);
//  * pointer scores (C1): 128-item tiles (item = un-masked (row, node)); E = F.Bc on tcgen05 in four N = 32 quarters; per
//    group 4 TMEM quarter buffers and 2 F tiles, so the MMAs of the NEXT tile are in flight while this one is computed.
//    Template variants: NSW 16 + SPLIT (default: two groups of 8 warps, the two warps of a TMEM sub-partition share a
//    tile's items and split the hidden units), NSW 16 (four groups, 2 buffers each, lock-step), NSW 8 (two groups score,
//    8 row warps stream the per-row tail behind them);
//  * tanh at 1.25 MUFU per element: four reciprocals share one rcp, evaluated as packed FP32x2 arithmetic;
//  * per-row item lists (node ids) and lists of the customers with demand left; a tile finds the row of an item with a
//    binary search in the row offsets; node features are gathered as 16-byte records (repack_features_kernel);
//  * per-row tail: a warp starts as soon as the tiles of its rows are done (one mbarrier per tile); plain greedy decoding
//    takes the fast tail (arg-max of the logits, exact soft-max only on near-ties, feasibility only for customers with
//    demand left), everything else the exact tail (log-softmax, probabilities, sampling, mask counts, traces);
//  * q^T has its own TMEM columns, so the first tiles of a step are staged while the query GEMM is still running.
//
// Beam search stays on rollout3 (rows of one instance compete for the beam slots, no per-row streaming).
#pragma once
#include "common.cuh"

namespace vrpb200 {


struct Smem4 {
    int ring, X, qq, F, logits, items, alive, bc, vsm, psum, dem, mask, rs, meta, bar, rowmap, total, stages;
};

__host__ __device__ inline Smem4 make_layout4(int RB, int N, int stages, int nsw) {
    Smem4 L;
    const int NP = align_up(N, 4);
    int off = 0;
    L.stages = stages;
    L.ring = off;   off += stages * kSlabBytes;
    L.X = off;      off += 2 * RB * kH * 2;                    // h' as the K-major B operand, FP16 hi | lo
    L.qq = off;     off += RB * kH * 4;
    L.F = off;      off += 4 * 2 * kC1TileBytes;               // nsw 16: 4 groups x 1 F tile; nsw 8: 2 groups x 2 F tiles
    (void)nsw;
    L.logits = off; off += RB * NP * 4;
    L.items = off;  off += align_up(RB * NP, 16);             // node ids of the items of every row
    L.alive = off;  off += align_up(RB * NP, 16);             // node ids of the customers with demand left, per row
    L.bc = off;     off += 2 * 4096;
    L.vsm = off;    off += 512;                               // -2 v
    L.psum = off;   off += 2 * 2 * 128 * 4;                   // split variant: partial score sums, [group][tile parity][item]
    L.dem = off;    off += RB * NP * 4;
    L.mask = off;   off += align_up(RB * NP, 16);
    L.rs = off;     off += RB * kRowState * 4;
    L.meta = off;   off += align_up((RB + 68 + RB + RB + 20) * 4, 16);   // nact[RB], off[<=65 (+pad)], flags[RB], nalive[RB], done[16], stop
    L.bar = off;    off += 320 + 64 * 8;                      // ... + one mbarrier per tile of the step (row warps wait on them)
    L.rowmap = off; off += align_up(RB * NP, 16);             // row of every item slot of the step (one byte; replaces a 6-step binary search per item)
    L.total = off;
    return L;
}


// NSW = number of score warps:
//   16 -> every warp scores (4 groups, 2 TMEM quarter buffers each, the MMAs are issued by a thread of the group), then
//         every warp runs the per-row tail (lock-step);
//    8 -> warps 0..7 score (2 groups; 4 TMEM quarter buffers and 2 F tiles each, so the MMAs of the NEXT tile are in
//         flight during the arithmetic of this one; they are issued by two dedicated warps, 18 and 19, and a score warp
//         never waits for its peers), warps 8..15 are row warps that stream the per-row tail behind the scores
//   16 + SPLIT -> every warp scores, in 2 groups of 8 warps with the deep buffering of the 8-warp variant: the two warps of a
//         TMEM sub-partition share the 32 items of a tile and take hidden units 0..63 / 64..127 (partial sums meet in
//         shared memory and are added in a fixed order); then every warp runs the per-row tail
// SAMP: the kernel launched for sampling (its fast per-row tail evaluates the soft-max and the CDF; the greedy kernel carries
// none of that code in its step loop - measured: 1.5 % of the greedy step)
template <int RB, bool TW, int STAGES, int NSW, bool SPLIT, bool SAMP>
__global__ void __launch_bounds__(kV4Threads, 1) rollout4_kernel(const __grid_constant__ RolloutArgs a) {
    constexpr int NT = 512, RPW = RB / 16, RPT = RB / 4, FD = TW ? 4 : 2;
    constexpr bool DEEP = NSW == 8 || SPLIT;                     // 4 TMEM buffers + 2 F tiles per group, MMAs issued by warps 18/19
    constexpr int NG = DEEP ? 2 : 4;                             // score groups
    constexpr int kQCol = 256 + 4 * RB;                          // q^T: its own columns
    constexpr int kGCol = 256;                                   // gates^T: tile m at kGCol + RB m
    static_assert(NSW == 8 || NSW == 16, "score warps");
    static_assert(!SPLIT || NSW == 16, "split");
    constexpr int LPR = 8;                                       // lanes per row in the per-row tail
    constexpr int NBT = NSW == 16 ? 1 : (RB + 31) / 32;          // row warps: batches of (up to) 32 rows
    static_assert(RB % 16 == 0 && RB <= 48, "rows");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), T = a.T, n_cust = a.n_cust;
    const Smem4 L = make_layout4(RB, N, STAGES, NSW);
    unsigned char* xt = smem + L.X;
    float* qq = reinterpret_cast<float*>(smem + L.qq);
    unsigned char* ftile = smem + L.F;
    float* logits = reinterpret_cast<float*>(smem + L.logits);
    unsigned char* items = smem + L.items;
    unsigned char* alive = smem + L.alive;
    float* vsm = reinterpret_cast<float*>(smem + L.vsm);
    float* psum = reinterpret_cast<float*>(smem + L.psum);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    unsigned char* rowmap = smem + L.rowmap;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    int* nact_s = reinterpret_cast<int*>(smem + L.meta);
    int* off_s = nact_s + RB;            // [RB + 1] exclusive prefix of nact (this step)
    int* flag_s = off_s + 68;            // [RB] every node of the row is an item
    int* nalive_s = flag_s + RB;         // [RB] customers with demand left
    int* done_s = nalive_s + RB;         // [16] per warp: every row of the warp is finished
    volatile int* stop_s = done_s + 16;
    const uint32_t bar0 = smem_u32(smem + L.bar);
    const uint32_t bar_empty = bar0 + 32, bar_q = bar0 + 64, bar_g = bar0 + 72, bar_x = bar0 + 80, tmem_slot = bar0 + 120,
                   bar_tile = bar0 + 320, bar_grp = bar0 + 128;   // group g at +40 g: full[0], full[1] (MMA commit), free[0], free[1] (4 warps copied the buffer), F ready (4 warps)
    const uint32_t ring_u32 = smem_u32(smem + L.ring), bc_u32 = smem_u32(smem + L.bc);
    const uint32_t x_hi = smem_u32(xt), x_lo = x_hi + RB * kH * 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sp = warp & 3, ub = warp >> 2;
    const long long b0 = (long long)blockIdx.x * RB;
    const int IBv = (int)min((long long)RB, a.B - b0);
    const long long R = a.B;
    const int D = a.D;
    // node features re-packed by repack_features_kernel: (x, y, b_tw, e_tw) of a node in ONE 16-byte load (the [B, N, D]
    // input has 12- or 20-byte records: three or four scattered 4-byte loads per node otherwise)
    const float4* gfeat4 = a.feat4 + b0 * (long long)N;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_q, 1); mbar_init(bar_g, 1); mbar_init(bar_x, 1);
        if (!DEEP) {
            for (int g = 0; g < 4; ++g) {   // per group: full[2] (MMA commit), free[2] (4 warps copied the buffer), F ready (4 warps)
                mbar_init(bar_grp + 40 * g, 1); mbar_init(bar_grp + 40 * g + 8, 1);
                mbar_init(bar_grp + 40 * g + 16, 4); mbar_init(bar_grp + 40 * g + 24, 4); mbar_init(bar_grp + 40 * g + 32, 4);
            }
        } else {
            for (int g = 0; g < 2; ++g) {   // per group: full[4], free[4], F ready[2]
                for (int q = 0; q < 4; ++q) { mbar_init(bar_grp + 80 * g + 8 * q, 1); mbar_init(bar_grp + 80 * g + 32 + 8 * q, 4); }
                mbar_init(bar_grp + 80 * g + 64, 4); mbar_init(bar_grp + 80 * g + 72, 4);
                mbar_init(bar_grp + 160 + 8 * g, 4);   // split variant: the partial sums of a tile are in place
                mbar_init(bar_grp + 176 + 8 * g, 4);   // split variant: ... and have been consumed (back-pressure: the producing
                                                       // half may run at most one tile ahead, or the parity wait would alias)
            }
        }
        for (int i = 0; i < 64; ++i) mbar_init(bar_tile + 8 * i, 4);   // 4 arrivals per step: the 4 warps that scored the tile
        *stop_s = 0;
        for (int w = 0; w < 16; ++w) done_s[w] = 1;
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc_alloc(tmem_slot, kTcCols);
    for (int i = tid; i < 128 * 8; i += kV4Threads) {
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
    for (int i = tid; i < RB * kH; i += kV4Threads) reinterpret_cast<float*>(xt)[i] = 0.0f;   // h_0 = 0
    for (int i = tid; i < IBv * N; i += kV4Threads) {
        const int lb = i / N, n = i - lb * N;
        dem[lb * NP + n] = a.input[((b0 + lb) * N + n) * D + D - 1];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.bar + 120);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * sp) << 16);
    constexpr uint32_t idescL = tc_idesc_f16(128, RB), idescQ = tc_idesc(128, 32);

    if (warp >= 16) {
        // ============ streaming warps: warp 16 feeds the weight ring from L2, warp 17 issues every LSTM / query MMA.
        // (Two warps: as two lanes of one warp a blocking mbarrier.try_wait of one role stalls the other.)
        if (warp == 16 && lane == 0) {
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
                if (*stop_s) {
                    for (uint32_t j = c - pre; j < c; ++j) mbar_wait(bar0 + 8 * (j % STAGES), (j / STAGES) & 1);
                    break;
                }
                for (int u = pre; u < kHUsesQ + kHUsesG; ++u) issue(u);
                pre = STAGES < kHUsesQ ? STAGES : kHUsesQ;
                for (int u = 0; u < pre; ++u) issue(u);
            }
        } else if (warp == 17) {   // the whole warp runs the loop (uniform control flow), one elected lane issues
            uint32_t c = 0;
            int step = 0;
            const uint64_t wdesc0 = tc_desc(ring_u32, 128, 256), xdesc_hi0 = VRP_XDESC_H(x_hi, 0), xdesc_lo0 = VRP_XDESC_H(x_lo, 0);
#ifdef VRPB200_PHASE_CLOCKS
            long long ic_[6] = {0, 0, 0, 0, 0, 0}, ic_last_ = clock64();   // wait h', q: wait weights, q: issue, gates: wait weights, gates: throttle, gates: issue
#define VRP_ICLK(i) do { const long long c_ = clock64(); ic_[i] += c_ - ic_last_; ic_last_ = c_; } while (0)
#else
#define VRP_ICLK(i) do { } while (0)
#endif
            for (;; ++step) {
                mbar_wait(bar_x, step & 1);
                tc_fence_after();
                VRP_ICLK(0);
                if (*stop_s) break;
#pragma unroll 1
                for (int u = 0; u < kHUsesQ; ++u, ++c) {          // q^T[unit, row] = W_q^T . h^T  -> columns kQCol..
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
                    tc_fence_after();
                    VRP_ICLK(1);
                    // descriptors by addition (the address field counts 16-byte units and never carries out of its 14 bits):
                    // the single issuing thread is latency-bound on the uniform datapath, every instruction per MMA counts
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
                    VRP_ICLK(2);
                }
                if (elect_one()) tc_commit(bar_q);
                __syncwarp();
#pragma unroll 1
                for (int u = 0; u < kHUsesG; ++u, ++c) {          // gates^T of the NEXT step, h part
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
                    VRP_ICLK(3);
#ifndef VRP_NO_THROTTLE_GATES
                    if (u > 0) {   // throttle: this GEMM has a whole step of slack, the score MMAs of C1 queue behind it
                        const uint32_t cp = c - 1;
                        mbar_wait(bar_empty + 8 * (cp % STAGES), (cp / STAGES) & 1);
                    }
#endif
                    tc_fence_after();
                    VRP_ICLK(4);
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
                    VRP_ICLK(5);
                }
                if (elect_one()) tc_commit(bar_g);
                __syncwarp();
            }
            if (step > 0) mbar_wait(bar_g, (step - 1) & 1);
#ifdef VRPB200_PHASE_CLOCKS
            if (a.stats && lane == 0) for (int i = 0; i < 6; ++i) atomicAdd(a.stats + 40 + i, (unsigned long long)ic_[i]);   // the clock variant's stats buffer has 64 slots
#endif
        }
        else if (DEEP && warp >= 18) {   // (whole warp, one elected lane issues: see elect_one)
            // ---- score-MMA issuer of group g: one tile per "F ready"; quarter q goes to TMEM buffer q once the four score
            //      warps have copied the previous tile's quarter q out of it
            const int g = warp - 18;
            const uint32_t bfull = bar_grp + 80 * g, bfree = bfull + 32, bfrdy = bfull + 64;
            const uint32_t fbase = smem_u32(smem + L.F) + g * 4 * kC1TileBytes;
            const uint64_t bdesc0 = tc_desc(bc_u32, 128, 256);
            for (uint32_t tc = 0;; ++tc) {
                const uint32_t fb = tc & 1;
                mbar_wait(bfrdy + 8 * fb, (tc >> 1) & 1);
                if (*stop_s) break;
                const uint64_t dfh = tc_desc(fbase + fb * 2 * kC1TileBytes, 128, 256), dfl = tc_desc(fbase + fb * 2 * kC1TileBytes + kC1TileBytes, 128, 256);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    if (tc > 0) mbar_wait(bfree + 8 * q, (tc - 1) & 1);
                    tc_fence_after();
                    const uint64_t dbh = bdesc0 + (uint64_t)(q * (1024 >> 4)), dbl = dbh + (4096 >> 4);
                    const uint32_t d = tmem_base + 128 * g + 32 * q;
                    if (elect_one()) {
                        tc_mma_tf32(d, dfl, dbh, idescQ, 0);
                        tc_mma_tf32(d, dfh, dbl, idescQ, 1);
                        tc_mma_tf32(d, dfh, dbh, idescQ, 1);
                        tc_commit(bfull + 8 * q);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }

    // =============================== 16 compute warps ===========================================================
    // Code size matters here: the kernel has three warp roles running different code at the same time, and straight-line
    // (fully unrolled) phases of a few thousand instructions each thrash the instruction caches.  Every per-node loop
    // below is a rolled loop with a uniform trip count NJ (guards inside), the LSTM epilogue rotates its cell-state
    // registers instead of unrolling over row chunks, and the C1 quarter loop is rolled.
    const int NJ = (N + LPR - 1) / LPR;   // nodes per lane of a row sub-warp: n = sl + 8 j
    // 8-lane sub-warp of a row: the item list (+ count, + "all nodes" flag) of the next step from the row's mask.
    // All four sub-warps of a warp call this together (ballots are warp-wide).
    auto build_items = [&](int lrs, bool rvalid, bool finished, int sub, int sl) -> int {
        const unsigned char* mk = maskb + lrs * NP;
        unsigned char* it = items + lrs * NP;
        const bool on = rvalid && !finished;
        int nun = 0;
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
            const int n = sl + LPR * j;
            const bool act = on && n < N && mk[n] == 0;
            const unsigned bits = (__ballot_sync(kFull, act) >> (8 * sub)) & 0xffu;
            if (act) it[nun + __popc(bits & ((1u << sl) - 1))] = (unsigned char)n;
            nun += __popc(bits);
        }
        int fl = 0, na = nun;
        if (on && (a.full || nun == 0 || !a.mask_pointer)) {   // every node is evaluated
            fl = 1; na = N;
            for (int n = sl; n < N; n += LPR) it[n] = (unsigned char)n;
        }
        if (sl == 0 && rvalid) { nact_s[lrs] = na; flag_s[lrs] = fl; }
        return nun;
    };

    // list of the customers with demand left, from the row's demands (8-lane sub-warp; all four sub-warps together)
    auto build_alive = [&](int lrs, bool rvalid, int sub, int sl) {
        const float* drow = dem + lrs * NP;
        unsigned char* al = alive + lrs * NP;
        int cnt = 0;
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
            const int n = sl + LPR * j;
            const bool keep = rvalid && n < N && n != n_cust && drow[n] != 0.0f;
            const unsigned bits = (__ballot_sync(kFull, keep) >> (8 * sub)) & 0xffu;
            if (keep) al[cnt + __popc(bits & ((1u << sl) - 1))] = (unsigned char)n;
            cnt += __popc(bits);
        }
        if (sl == 0 && rvalid) nalive_s[lrs] = cnt;
    };
    // exact mask of every node of a row from its state: Env.reset's mask before the first step (vrptw_utils.py:233-240),
    // Env.step's afterwards (:328-348).  Used when the fast per-row tail (which keeps no mask bytes) hands a row over.
    auto recompute_masks = [&](int lrs, bool rvalid, int sl, int step) {
        if (!rvalid) return;
        const float* rs = rowst + lrs * kRowState;
        const float* drow = dem + lrs * NP;
        const float4* frow = gfeat4 + (size_t)lrs * N;
        unsigned char* mk = maskb + lrs * NP;
        for (int n = sl; n < N; n += LPR) {
            int mm;
            if (step == 0) mm = (n == n_cust) ? 1 : (drow[n] == 0.0f ? 1 : 0);
            else {
                float nx = 0.f, ny = 0.f, ne = 0.f;
                if (TW) { const float4 fv = __ldcg(frow + n); nx = fv.x; ny = fv.y; ne = fv.w; }
                mm = node_mask<TW>(n, n_cust, drow[n], rs[0], TW ? rs[1] : 0.0f, rs[2], rs[3], nx, ny, ne, rs[15], rs[13]);
            }
            mk[n] = (unsigned char)mm;
        }
    };

    for (int i = 0; i < RPW; ++i) {   // Env.reset (vrptw_utils.py:199-252) + prologue
        const int lr = warp * RPW + i;
        if (lane == 0) { nact_s[lr] = 0; flag_s[lr] = 0; }
        if (lr >= IBv) continue;
        int nzc = 0;
        for (int n = lane; n < N; n += 32) {
            const bool z = dem[lr * NP + n] == 0.0f;
            maskb[lr * NP + n] = (n == n_cust) ? 1 : (z ? 1 : 0);
            nzc += z ? 0 : 1;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) nzc += __shfl_xor_sync(kFull, nzc, o);
        const float4 fdep = __ldcg(gfeat4 + (size_t)lr * N + n_cust);
        if (lane == 0) {
            float* rs = rowst + lr * kRowState;
            rs[13] = (float)nzc; rs[15] = 0.0f;
            rs[0] = a.capacity; rs[1] = 0.0f; rs[2] = fdep.x; rs[3] = fdep.y; rs[4] = 0.0f; rs[5] = 0.0f; rs[6] = 0.0f;
            reinterpret_cast<int*>(rs)[8] = 0;
            rs[9] = fdep.x; rs[10] = fdep.y; rs[11] = TW ? fdep.z : 0.0f; rs[12] = TW ? fdep.w : 0.0f;
        }
    }
    named_bar(5, NT);
    if (warp < RB / 4) {
        const int sub = lane >> 3, sl = lane & 7, lr = 4 * warp + sub;
        build_items(lr < IBv ? lr : 0, lr < IBv, false, sub, sl);
        build_alive(lr < IBv ? lr : 0, lr < IBv, sub, sl);
    }
    // the fast per-row tail serves plain greedy decoding when nobody reads probabilities, log-probabilities or masks
    // (sampling takes it too: the soft-max is then always evaluated - in the exact tail's summation order, so both tails draw
    //  the same node from the same uniform - and log-probabilities may be requested)
    const bool fastk = !a.full && a.mask_pointer && !a.use_tanh && a.fin_mask == nullptr && a.tr_logits == nullptr && a.tr_probs == nullptr &&
                       a.tr_masks == nullptr && (SAMP ? a.mode == 1 : (a.mode == 0 && a.logprob_out == nullptr));
    const int unit = 32 * sp + lane;
    float cst[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) cst[r] = 0.0f;
    named_bar(5, NT);

#ifdef VRPB200_PHASE_CLOCKS
    long long pc_[28], pc_last_ = clock64();
    for (int i = 0; i < 28; ++i) pc_[i] = 0;
#endif
    // C1: the two TMEM buffers of a group are used strictly in turn, so one counter of consumed quarters (every thread)
    // and one of issued quarters (issuer thread) give buffer index and mbarrier parity
    uint32_t c1_nq = 0, c1_ni = 0, c1_nf = 0;
    // x part of the LSTM gates (rank D-1 through the linear embedding) and biases of this thread's unit.  They are (re)loaded
    // right before the end-of-step barrier, so that the L2 latency is spent waiting there and not at the top of the epilogue
    // (L1 keeps next to nothing beside 226 KB of shared memory)
    float wx[4][4], bgt[4];
    auto load_lstm_consts = [&]() {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            bgt[m] = ldcg_volatile(a.bgT + m * kH + unit);
#pragma unroll
            for (int k = 0; k < 4; ++k) wx[k][m] = ldcg_volatile(a.WxT + (k * 4 + m) * kH + unit);
        }
    };
    load_lstm_consts();
    unsigned long long n_eval = 0, n_rowsteps = 0;
    int t = 0;
#pragma unroll 1
    for (; t < T; ++t) {
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
            if (NSW == 16) {   // item slot -> row: the warp that owns rows 4w..4w+3 in the per-row tail fills their slots (8 lanes per row)
                const int rr = 4 * warp + (lane >> 3);
                const int o0 = __shfl_sync(kFull, s0 - n0, rr & 31), o1 = __shfl_sync(kFull, tot0 + s1 - n1, rr & 31);
                if (warp < RB / 4) {
                    const int ofs = rr < 32 ? o0 : o1, cnt = nact_s[rr];
                    for (int k = lane & 7; k < cnt; k += 8) rowmap[ofs + k] = (unsigned char)rr;
                }
            }
        }
        // ================= LSTM epilogue: gates = gates^T(TMEM, h part) + x part + bias ; c, h' ; h' -> X ==========
        {
            VRP_CLK(19);      // (row offsets of the step)
            if (t > 0) {
                mbar_wait(bar_g, (t - 1) & 1);
                tc_fence_after();
            }
            VRP_CLK(0);
#pragma unroll 1
            for (int c4 = 0; c4 < RPT / 4; ++c4) {
                const int r0 = ub * RPT + 4 * c4;
                float g[4][4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (t > 0) tc_ld4(tmem_lane + kGCol + RB * m + r0, g[m]);
                    else { g[m][0] = g[m][1] = g[m][2] = g[m][3] = 0.0f; }
                }
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* rs = rowst + (r0 + i) * kRowState;
                    const float xf[4] = {rs[9], rs[10], rs[11], rs[12]};
                    float gg[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        float s = bgt[m];
#pragma unroll
                        for (int k = 0; k < FD; ++k) s = fmaf(xf[k], wx[k][m], s);
                        gg[m] = g[m][i] + s;
                    }
                    const float cn = __fadd_rn(__fmul_rn(cst[i], sigmoid_fast(gg[2] + a.forget_bias)),
                                               __fmul_rn(sigmoid_fast(gg[0]), tanh_fast(gg[1])));
                    cst[i] = cn;
                    const float hn = __fmul_rn(tanh_fast(cn), sigmoid_fast(gg[3]));
                    hi[i] = __half2float(__float2half_rn(hn));      // FP16 two-term split (common.cuh): hi, then lo = fp16(h - hi) at the store
                    lo[i] = hn - hi[i];
                }
                if (RPT > 4) {   // rotate the cell-state registers: the chunk just updated goes to the back
                    const float t0 = cst[0], t1 = cst[1], t2 = cst[2], t3 = cst[3];
#pragma unroll
                    for (int r = 0; r + 4 < RPT; ++r) cst[r] = cst[r + 4];
                    cst[RPT - 4] = t0; cst[RPT - 3] = t1; cst[RPT - 2] = t2; cst[RPT - 1] = t3;
                }
                transpose4(hi, lane);
                transpose4(lo, lane);
                {
                    const int rr = r0 + (lane & 3);
                    // [row/8][k/8][row%8][k%8] halves: this lane holds units 4 (unit/4) .. +3 of row rr = 8 bytes
                    const int off = (rr >> 3) * 2048 + (unit >> 3) * 128 + (rr & 7) * 16 + ((unit >> 2) & 1) * 8;
                    const __half2 h01 = __floats2half2_rn(hi[0], hi[1]), h23 = __floats2half2_rn(hi[2], hi[3]);
                    const __half2 l01 = __floats2half2_rn(lo[0], lo[1]), l23 = __floats2half2_rn(lo[2], lo[3]);
                    *reinterpret_cast<uint2*>(xt + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                    *reinterpret_cast<uint2*>(xt + RB * kH * 2 + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                }
            }
            fence_proxy_async();
            tc_fence_before();
            named_bar(5, NT);
            if (tid == 0) mbar_arrive(bar_x);      // h' is in place: the streaming warp may issue q(t) and gates(t+1)
        }
        VRP_CLK(1);

        // ================= C1 set-up (score warps): first tile(s) of every group, staged while the query GEMM runs ======
        const int ntiles = (m_tot + 127) >> 7;
        const int g = SPLIT ? (warp >> 3) : ub;                          // score group
        const int half = SPLIT ? (ub & 1) : 0;                           // split variant: hidden units 64 half .. 64 half + 63
        const int m = 32 * sp + lane;
        // NSW 16: group g owns F tile g and TMEM columns 64 g (2 buffers x 32); NSW 8: F tiles 2g, 2g+1, columns 128 g (4 x 32)
        unsigned char* fgrp = ftile + (g < NG ? g : 0) * (DEEP ? 4 : 2) * kC1TileBytes;
        const uint32_t f_hi = smem_u32(fgrp), f_lo = f_hi + kC1TileBytes;
        const uint32_t d_g = tmem_base + 64 * g;
        const uint32_t bfull = bar_grp + (DEEP ? 80 : 40) * (g < NG ? g : 0), bfree = bfull + (DEEP ? 32 : 16),
                       bfrdy = bfull + (DEEP ? 64 : 32), bpsum = bar_grp + 160 + 8 * (g < NG ? g : 0), bpack = bpsum + 16;
        const bool issuer = (!DEEP && sp == 0 && lane == 0);
        // item `it` of the step -> (row, node): binary search in the row offsets, then the row's node list
        auto load_item = [&](int tile, int& r, int& n, bool& valid, float (&f)[5], bool feats = true) {
            const int it = tile * 128 + m;
            valid = tile < ntiles && it < m_tot;
            r = 0; n = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) f[k] = 0.0f;
            if (valid) {
                if (NSW == 16) r = rowmap[it];
                else {
                    int lo = 0, hi = RB;
#pragma unroll
                    for (int s = 0; s < 6; ++s) {
                        const int mid = (lo + hi) >> 1;
                        if (off_s[mid] <= it) lo = mid; else hi = mid;
                    }
                    r = lo;
                }
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
        auto issue_quarter = [&](int q) {   // NSW 16, issuer thread only: E[:, 32q .. 32q+31] -> the next TMEM buffer in turn
            const uint32_t b = c1_ni & 1, k = c1_ni >> 1;
            ++c1_ni;
            if (k > 0) mbar_wait(bfree + 8 * b, (k - 1) & 1);   // the 4 warps have copied its previous content
            tc_fence_after();
            const uint64_t dfh = tc_desc(f_hi, 128, 256), dfl = tc_desc(f_lo, 128, 256);
            const uint64_t dbh = tc_desc(bc_u32 + q * 1024, 128, 256), dbl = tc_desc(bc_u32 + 4096 + q * 1024, 128, 256);
            tc_mma_tf32(d_g + 32 * b, dfl, dbh, idescQ, 0);
            tc_mma_tf32(d_g + 32 * b, dfh, dbl, idescQ, 1);
            tc_mma_tf32(d_g + 32 * b, dfh, dbh, idescQ, 1);
            tc_commit(bfull + 8 * b);
        };
        // 32 hidden units of one item: acc += sum_h w_h / (1 + 2^(E_h + base'_h)), w = -2 v, 16 independent elements at a
        // time, ONE reciprocal per 4 elements:  (w0 t1 + w1 t0) t2 t3 + (w2 t3 + w3 t2) t0 t1  over  t0 t1 t2 t3
        // (t = 1 + 2^min(arg, 30) <= 2^30 + 1: the product cannot overflow, and tanh is 1.0f exactly from arg = 26 on)
        // 8 elements at a time as four pairs P0..P3 = (e0,e1) .. (e6,e7); the two quads that share one reciprocal each are
        // the .x and the .y lanes of the pairs, so the whole shared-reciprocal algebra is packed FP32x2 arithmetic; the last
        // pair of every group of eight takes its 2^x from the FMA pipe instead of the MUFU pipe (6 ex2 + 2 rcp per 8 elements)
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
        int r_cur = 0, n_cur = 0, r_nx = 0, n_nx = 0;
        bool v_cur = false, v_nx = false;
        const bool scoring = warp < NSW && g < ntiles;
        if (scoring) {
            // split variant: the F tile of the group's tile number c (counted over all steps) lives in F buffer c & 1 and is
            // gathered and staged by the warps of half c & 1 - the two warps of a pair take turns, so neither carries the
            // gathers and the F stores of every tile
            const int st0 = SPLIT ? (int)(c1_nq & 1) : 0;        // the half that stages the group's first tile of this step
            float fc[5];
            load_item(g, r_cur, n_cur, v_cur, fc, half == st0);
            if (!DEEP) {
                store_F(fgrp, fc);
                __syncwarp();
                if (lane == 0) mbar_arrive(bfrdy);
                if (issuer) {
                    mbar_wait(bfrdy, c1_nf & 1); ++c1_nf;      // the four warps of the group have written their part of F
                    issue_quarter(0); issue_quarter(1);
                }
                __syncwarp();
            } else {   // the F tiles of the group's first two tiles of the step; warp 18 + g issues the MMAs
                float f2[5];
                load_item(g + NG, r_nx, n_nx, v_nx, f2, SPLIT ? half != st0 : half == 0);
                if (half == st0) {
                    store_F(fgrp + (c1_nq & 1) * 2 * kC1TileBytes, fc);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfrdy + 8 * (c1_nq & 1));
                }
                if ((SPLIT ? half != st0 : half == 0) && g + NG < ntiles) {
                    store_F(fgrp + ((c1_nq + 1) & 1) * 2 * kC1TileBytes, f2);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfrdy + 8 * ((c1_nq + 1) & 1));
                }
            }
        }
        VRP_CLK(2);
        // ================= query epilogue: base'[row][unit] = 2log2e (q + b + load g_ld + time g_t) ===================
        {
            const float bqv = __ldg(a.att.bq + unit), glv = __ldg(a.att.gl + unit), gtv = __ldg(a.att.gt + unit);
            mbar_wait(bar_q, t & 1);
            tc_fence_after();
            VRP_CLK(3);
#pragma unroll 1
            for (int c4 = 0; c4 < RPT / 4; ++c4) {
                const int r0 = ub * RPT + 4 * c4;
                float q4[4];
                tc_ld4(tmem_lane + kQCol + r0, q4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* rs = rowst + (r0 + i) * kRowState;
                    float s = fmaf(rs[0], glv, q4[i] + bqv);
                    if (TW) s = fmaf(rs[1], gtv, s);
                    qq[(r0 + i) * kH + unit] = s * kTwoLog2e;
                }
            }
        }
        tc_fence_before();
        named_bar(5, NT);
        tc_fence_after();
        VRP_CLK(4);

        if (warp < NSW) {
            // ================= C1: scores =================================================================
            if (scoring && !DEEP) {
                float fn[5];
#pragma unroll 1
                for (int tile = g; tile < ntiles; tile += NG) {
                    const bool has_next = tile + NG < ntiles;
                    load_item(tile + NG, r_nx, n_nx, v_nx, fn);          // in flight during the arithmetic below
                    const float* brow = qq + r_cur * kH;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t b = c1_nq & 1, par = (c1_nq >> 1) & 1;
                        ++c1_nq;
                        VRP_CLK(11);
                        mbar_wait(bfull + 8 * b, par);
                        tc_fence_after();
                        VRP_CLK(7);
                        if (q == 3 && has_next) {   // every MMA of this tile has completed: the F tile may be replaced
                            store_F(fgrp, fn);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bfrdy);
                            if (issuer) { mbar_wait(bfrdy, c1_nf & 1); ++c1_nf; issue_quarter(0); }
                            __syncwarp();
                        }
                        VRP_CLK(8);
                        uint32_t cur[32];
                        tc_ld32_issue(tmem_lane + 64 * g + 32 * b, cur);
                        tc_ld_wait32(cur);
                        tc_fence_before();
                        if (lane == 0) mbar_arrive(bfree + 8 * b);
                        if (issuer) {
                            if (q < 2) issue_quarter(q + 2);
                            else if (q == 3 && has_next) issue_quarter(1);
                        }
                        __syncwarp();
                        VRP_CLK(9);
                        quarter_arith(cur, brow + 32 * q, vsm + 32 * q, acc0, acc1);
                        VRP_CLK(10);
                    }
                    // u = sum_h v_h tanh(.) = sum_h v_h - sum_h 2 v_h / (1 + 2^arg)
                    if (v_cur) logits[r_cur * NP + n_cur] = a.sumv + (acc0 + acc1);
                    n_eval += v_cur ? 1 : 0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tile + 8 * tile);   // the rows of this tile may be finished (release: logits visible)
                    r_cur = r_nx; n_cur = n_nx; v_cur = v_nx;
                }
            }
            if (scoring && SPLIT) {
                // c1_nq counts the tiles of this group (all steps): tile k uses F tile k & 1, its quarter q TMEM buffer q.
                // Warps 8g..8g+3 (half 0) gather the items, write F and take quarters 0, 2; warps 8g+4..8g+7 take quarters 1, 3
                // of the same items and hand their partial sum over through shared memory.
                int r_n2 = 0, n_n2 = 0;
                bool v_n2 = false;
                float fn[5];
                float* ps = psum + g * 256;
#pragma unroll 1
                for (int tile = g; tile < ntiles; tile += NG) {
                    const float* brow = qq + r_cur * kH;
                    const uint32_t par = c1_nq & 1;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                    for (int qi = 0; qi < 2; ++qi) {
                        const int q = 2 * qi + half;      // half 0: quarters 0, 2; half 1: quarters 1, 3 (the MMAs land in order 0..3)
                        VRP_CLK(11);
                        mbar_wait(bfull + 8 * q, par);
                        tc_fence_after();
                        VRP_CLK(7);
                        uint32_t cur[32];
                        tc_ld32_issue(tmem_base + ((uint32_t)(32 * sp) << 16) + 128 * g + 32 * q, cur);
                        tc_ld_wait32(cur);
                        tc_fence_before();
                        if (lane == 0) mbar_arrive(bfree + 8 * q);      // warp 18 + g may send the next tile's quarter q here
                        __syncwarp();
                        VRP_CLK(9);
                        // the items two tiles ahead: row look-up and feature loads are scheduled into the arithmetic below
                        if (qi == 0) load_item(tile + 2 * NG, r_n2, n_n2, v_n2, fn, half == (int)par);      // (features: the half that will stage that tile)
                        quarter_arith(cur, brow + 32 * q, vsm + 32 * q, acc0, acc1);
                        VRP_CLK(10);
                    }
                    if (half == 1) {
                        if (par == 1 && tile + 2 * NG < ntiles) {   // this half has seen quarter 3 complete: every MMA of the tile has read F
                            store_F(fgrp + 2 * kC1TileBytes, fn);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bfrdy + 8);
                        }
                        if (c1_nq > 0) mbar_wait(bpack, (c1_nq - 1) & 1);   // the other half has taken the previous tile's sums
                        ps[par * 128 + m] = acc0 + acc1;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bpsum);
                    } else {
                        mbar_wait(bpsum, par);      // (also: quarters 1 and 3 of this tile have been read, its F tile is free)
                        const float psv = ps[par * 128 + m];
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bpack);
                        if (v_cur) logits[r_cur * NP + n_cur] = (a.sumv + (acc0 + acc1)) + psv;
                        n_eval += v_cur ? 1 : 0;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tile + 8 * tile);   // the rows of this tile may be finished
                        if (par == 0 && tile + 2 * NG < ntiles) {
                            store_F(fgrp, fn);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bfrdy);
                        }
                    }
                    VRP_CLK(8);
                    ++c1_nq;
                    r_cur = r_nx; n_cur = n_nx; v_cur = v_nx;
                    r_nx = r_n2; n_nx = n_n2; v_nx = v_n2;
                }
            }
            if (scoring && NSW == 8) {
                // c1_nq counts the tiles of this group (all steps): tile k uses F tile k & 1, its quarter q TMEM buffer q
                int r_n2 = 0, n_n2 = 0;
                bool v_n2 = false;
                float fn[5];
#pragma unroll 1
                for (int tile = g; tile < ntiles; tile += NG) {
                    load_item(tile + 2 * NG, r_n2, n_n2, v_n2, fn);      // two tiles ahead: in flight during the arithmetic below
                    const float* brow = qq + r_cur * kH;
                    const uint32_t par = c1_nq & 1;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        VRP_CLK(11);
                        mbar_wait(bfull + 8 * q, par);
                        tc_fence_after();
                        VRP_CLK(7);
                        uint32_t cur[32];
                        tc_ld32_issue(tmem_base + ((uint32_t)(32 * sp) << 16) + 128 * g + 32 * q, cur);
                        tc_ld_wait32(cur);
                        tc_fence_before();
                        if (lane == 0) mbar_arrive(bfree + 8 * q);      // warp 18 + g may send the next tile's quarter q here
                        __syncwarp();
                        VRP_CLK(9);
                        quarter_arith(cur, brow + 32 * q, vsm + 32 * q, acc0, acc1);
                        VRP_CLK(10);
                    }
                    if (v_cur) logits[r_cur * NP + n_cur] = a.sumv + (acc0 + acc1);
                    n_eval += v_cur ? 1 : 0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tile + 8 * tile);   // tell the row warps (release: the logits are visible)
                    if (tile + 2 * NG < ntiles) {   // all four MMAs of this tile are complete: its F tile takes the tile two ahead
                        store_F(fgrp + par * 2 * kC1TileBytes, fn);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bfrdy + 8 * par);
                    }
                    VRP_CLK(8);
                    ++c1_nq;
                    r_cur = r_nx; n_cur = n_nx; v_cur = v_nx;
                    r_nx = r_n2; n_nx = n_n2; v_nx = v_n2;
                }
            }
            VRP_CLK(5);
        }
        int my_done = 1;   // every row of this warp is finished (warps without rows: vacuously)
        if (NSW == 16 ? warp < RB / 4 : warp >= 8) {
            // ================= per-row tail: log-softmax, choice, Env.step, mask, next item list ==========================
            // 8 lanes per row, lane sl owns nodes sl, sl + 8, ...; masked logits and probabilities live in the row's logits
            // slots (shared memory), so no per-lane arrays are needed and every node loop stays rolled.
            // NSW == 8: warps 8..15 stream batches of up to 32 rows behind the score warps; NSW == 16: warp w owns rows 4w..4w+3
            const int sub = lane >> 3, sl = lane & 7;
            int warp_done = 1;
#pragma unroll 1
            for (int bt = 0; bt < NBT; ++bt) {
                const int nrow = NSW == 16 ? RB : min(32, RB - 32 * bt);   // rows of this batch, spread evenly over the row warps
                const int rpw = NSW == 16 ? 4 : nrow / 8;                  // rows per warp (4 or 2)
                const int r0 = NSW == 16 ? 4 * warp : 32 * bt + rpw * (warp - 8);
                const int lr = r0 + sub;
                {   // wait for every tile that holds items of rows r0 .. r0+rpw-1
                    const int i0 = off_s[r0], i1 = off_s[r0 + rpw];
                    if (i1 > i0) {
                        const int tl = (i1 - 1) >> 7;
                        for (int tile = i0 >> 7; tile <= tl; ++tile) mbar_wait(bar_tile + 8 * tile, t & 1);
                    }
                    __syncwarp();
                }
                VRP_CLK(12);
                const bool rvalid = sub < rpw && lr < IBv;
                const int lrs = rvalid ? lr : 0;
                float* rs = rowst + lrs * kRowState;
                const long long row = b0 + lrs;
                const bool was_done = rvalid && reinterpret_cast<int*>(rs)[8] != 0;
                const bool live = rvalid && !was_done;
                if (was_done && sl == 0) {
                    a.idx_out[(long long)t * R + row] = n_cust;
                    if (a.logprob_out) a.logprob_out[(long long)t * R + row] = 0.0f;
                }
                if (live && sl == 0) ++n_rowsteps;
                const float load = rs[0], time = TW ? rs[1] : 0.0f;
                unsigned char* mk = maskb + lrs * NP;
                const float4* frow = gfeat4 + (size_t)lrs * N;
                float* drow = dem + lrs * NP;
                float* ulog = logits + lrs * NP;
                const int fl = flag_s[lrs];
                const bool tracing = a.tr_logits || a.tr_probs || a.tr_masks;
                const long long tro = ((long long)t * R + row) * N;
                int done_now = 0;
                bool handled = false;
                if (fastk) {
                    // ============ fast tail (plain greedy decoding): the choice is the arg-max of the logits over the row's
                    // items; feasibility is only evaluated for the customers that still have demand, whose list (and the item
                    // list of the next step) is rebuilt by compaction; no mask bytes, no softmax.  A row with a near-tie of its
                    // two best logits, or without any un-masked node, goes through the exact tail below.
                    const int nal = live ? nalive_s[lrs] : 0;
                    unsigned char* al = alive + lrs * NP;
                    unsigned char* it = items + lrs * NP;
                    // (0) features of this lane's first 8 customers: issued now, used after the choice
                    int pn[8];
                    float pxx[8], pyy[8], pee[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = sl + LPR * j;
                        pn[j] = k < nal ? (int)al[k] : -1;
                        pxx[j] = pyy[j] = pee[j] = 0.0f;
                        if (TW && pn[j] >= 0) {
                            const float4 fv = ldcg4_volatile(frow + pn[j]);
                            pxx[j] = fv.x; pyy[j] = fv.y; pee[j] = fv.w;
                        }
                    }
                    VRP_CLK(14);
                    // (1) arg-max (first index on ties) over the items, and the largest logit strictly below the maximum.
                    //     Exactly equal logits have exactly equal probabilities (first index wins, as in the reference); only a
                    //     runner-up within 1e-5 BELOW the maximum could round to the same probability and needs the exact tail.
                    float mx = -INFINITY, second = -INFINITY;
                    int mxi = 0x7fffffff;
                    const int na = live ? nact_s[lrs] : 0;
#pragma unroll 2
                    for (int k = sl; k < na; k += LPR) {
                        const int n = it[k];
                        const float u = ulog[n];
                        if (u > mx) { second = mx; mx = u; mxi = n; }
                        else if (u == mx) mxi = min(mxi, n);
                        else second = fmaxf(second, u);
                    }
#pragma unroll
                    for (int o = LPR / 2; o; o >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, mx, o), os = __shfl_xor_sync(kFull, second, o);
                        const int oi = __shfl_xor_sync(kFull, mxi, o);
                        second = fmaxf(second, os);
                        if (ov != mx) second = fmaxf(second, fminf(mx, ov));
                        if (ov > mx || (ov == mx && oi < mxi)) { mx = ov; mxi = oi; }
                    }
                    handled = !__any_sync(kFull, live && fl != 0);
                    constexpr bool sampling = SAMP;
                    float psel = 1.0f;
                    if (handled && (sampling || __any_sync(kFull, live && second > mx - 1e-5f))) {
                        // near-tie: the reference takes argmax(prob) (attention_agent.py:146-147) and two close logits may round
                        // to one probability, so evaluate the probabilities exactly as the exact tail does (masked nodes have
                        // probability 0 and add 0 to the denominator; the mask bytes are kept as un-masked / masked flags)
#ifdef VRPB200_PHASE_CLOCKS
                        pc_[25] += 1;
#endif
                        float P0 = 0.0f, P1 = 0.0f, P2 = 0.0f, P3 = 0.0f;
#pragma unroll 1
                        for (int k = 0; k < 4; ++k) {
                            const int nb = sl + 32 * k;
                            float ev[4];
#pragma unroll
                            for (int qv = 0; qv < 4; ++qv) {   // branch-free: loads and expf for every slot, then a select (the
                                const int n = nb + 8 * qv;       // branchy form re-derived the shared-memory window per element)
                                const int nn = n < N ? n : N - 1;
                                const float e = expf(ulog[nn] - mx);
                                ev[qv] = (n < N && live && mk[nn] == 0) ? e : 0.0f;
                            }
                            P0 = __fadd_rn(P0, ev[0]); P1 = __fadd_rn(P1, ev[1]); P2 = __fadd_rn(P2, ev[2]); P3 = __fadd_rn(P3, ev[3]);
                        }
                        float se = __fadd_rn(__fadd_rn(P0, P2), __fadd_rn(P1, P3));
#pragma unroll
                        for (int o = LPR / 2; o; o >>= 1) se = __fadd_rn(se, __shfl_xor_sync(kFull, se, o));
                        const float lse = logf(se);
                        float bestp = -1.0f;
                        int besti = 0x7fffffff;
#pragma unroll 1
                        for (int j = 0; j < NJ; ++j) {
                            const int n = sl + LPR * j;
                            const int nn = n < N ? n : N - 1;
                            const float e = expf((ulog[nn] - mx) - lse);
                            const float p = (n < N && live && mk[nn] == 0) ? e : 0.0f;
                            if (n < N && p > bestp) { bestp = p; besti = n; }
                            if (sampling && live && n < N) ulog[n] = p;      // probabilities for the sequential CDF below
                        }
                        if (!sampling) {
#pragma unroll
                            for (int o = LPR / 2; o; o >>= 1) {
                                const float ov = __shfl_xor_sync(kFull, bestp, o);
                                const int oi = __shfl_xor_sync(kFull, besti, o);
                                if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                            }
                            mxi = besti;
                        } else {   // inverse-CDF sampling (attention_agent.py:148-167), as in the exact tail: sequential fp32 sums over the items
                            __syncwarp();
                            int sel = -1;
                            if (sl == 0 && live) {
                                const long long o = (long long)t * R + row;
                                int d0_, d1_;
                                draw_range(a, t, d0_, d1_);
                                for (int draw = d0_; draw < d1_ && sel < 0; ++draw) {
                                    const float* ubuf = draw ? a.uniforms_retry : a.uniforms;
                                    const float u = ubuf ? ubuf[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                                    float cum = 0.0f;
                                    for (int k = 0; k < na; ++k) {
                                        const int n = it[k];
                                        cum = __fadd_rn(cum, ulog[n]);
                                        if (cum > u) { sel = n; break; }
                                    }
                                }
                                if (sel < 0 && a.fail_steps && d0_ == 0) a.fail_steps[t] = 1;
                            }
                            if (sel < 0) sel = 0;
                            mxi = __shfl_sync(kFull, sel, sub * LPR);
                            __syncwarp();
                            psel = ulog[mxi < N ? mxi : 0];
                            __syncwarp();
                        }
                    }
                    VRP_CLK(15);
                    if (handled) {
                        const int idx = live ? mxi : 0;
                        // (2) Env.step (vrptw_utils.py:254-358): demand, load, time, position
                        const float4 fsel = __ldcg(frow + idx);
                        const float sx = fsel.x, sy = fsel.y;
                        const float sb = TW ? fsel.z : 0.0f, se_tw = TW ? fsel.w : 0.0f;
                        const float px = rs[2], py = rs[3];
                        const float d_idx = drow[idx];
                        float nz = rs[13];
                        __syncwarp();
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
                        if (live && sl == 0) drow[idx] = d_new;
                        done_now = (idx == n_cust && nz == 0.0f) ? 1 : 0;
                        if (live && sl == 0) {
                            rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                            if (t == 0) { rs[5] = sx; rs[6] = sy; }
                            else rs[4] = __fadd_rn(rs[4], ts);
                            reinterpret_cast<int*>(rs)[8] = done_now;
                            rs[9] = sx; rs[10] = sy; rs[11] = sb; rs[12] = se_tw; rs[13] = nz; rs[15] = dep;
                            a.idx_out[(long long)t * R + row] = idx;
                            if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
                        }
                        VRP_CLK(16);
                        // (3) customers with demand left -> new list; those that are feasible (vehicle not empty, time window
                        //     reachable: vrptw_utils.py:328-348) -> items of the next step; the depot unless the vehicle stands
                        //     there with demand left
                        const bool on = live && !done_now;
                        const bool cust_ok = nload != 0.0f;
                        int nal2 = 0, nit = 0;
                        const int jmax = __reduce_max_sync(kFull, (unsigned)((nal + LPR - 1) / LPR));
#pragma unroll 1
                        for (int jb = 0; jb < jmax; jb += 8) {
                            if (jb > 0) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int k = sl + LPR * (jb + j);
                                    pn[j] = k < nal ? (int)al[k] : -1;
                                    pxx[j] = pyy[j] = pee[j] = 0.0f;
                                    if (TW && pn[j] >= 0) {
                                        const float4 fv = ldcg4_volatile(frow + pn[j]);
                                        pxx[j] = fv.x; pyy[j] = fv.y; pee[j] = fv.w;
                                    }
                                }
                            }
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int n = pn[j];
                                const float dn = n < 0 ? 0.0f : (n == idx ? d_new : drow[n]);
                                const bool keep = on && n >= 0 && dn != 0.0f;
                                bool feas = keep && cust_ok;
                                if (TW && feas) feas = !(__fadd_rn(ntime, dist_rn(sx, sy, pxx[j], pyy[j])) > pee[j]);
                                const unsigned bk = (__ballot_sync(kFull, keep) >> (8 * sub)) & 0xffu;
                                const unsigned bf = (__ballot_sync(kFull, feas) >> (8 * sub)) & 0xffu;
                                const unsigned below = (1u << sl) - 1;
                                if (keep) al[nal2 + __popc(bk & below)] = (unsigned char)n;
                                if (feas) it[nit + __popc(bf & below)] = (unsigned char)n;
                                if (on && n >= 0) mk[n] = feas ? 0 : 1;      // un-masked / masked flag (the exact count is not needed)
                                nal2 += __popc(bk);
                                nit += __popc(bf);
                            }
                        }
                        const bool depot_ok = on && !(nz > 0.0f && dep != 0.0f);
                        if (depot_ok && sl == 0) it[nit] = (unsigned char)n_cust;
                        if (on && sl == 0) mk[n_cust] = depot_ok ? 0 : 1;
                        nit += depot_ok ? 1 : 0;
                        int flv = 0;
                        if (on && nit == 0) {   // nothing un-masked: every node is evaluated, the mask counts decide (exact tail)
                            flv = 1; nit = N;
                            for (int n = sl; n < N; n += LPR) it[n] = (unsigned char)n;
                        }
                        if (sl == 0 && rvalid) { nact_s[lrs] = on ? nit : 0; flag_s[lrs] = flv; nalive_s[lrs] = nal2; }
                        VRP_CLK(17);
                    }
                }
                if (!handled) {
#ifdef VRPB200_PHASE_CLOCKS
                    if (fastk) { pc_[26] += 1; pc_[27] += __popc(__ballot_sync(kFull, live && sl == 0 && (fl != 0))); }   // exact-tail visits of this warp; rows without un-masked node
#endif
                    if (fastk) { recompute_masks(lrs, rvalid, sl, t); __syncwarp(); }
                    // ---- masked logits (decode_step.py:231-232) back into the row's slots; maximum, its first index, runner-up
                    float mx = -INFINITY, second = -INFINITY;
                    int mxi = 0x7fffffff;
    #pragma unroll 2
                    for (int j = 0; j < NJ; ++j) {
                        const int n = sl + LPR * j;
                        float u = -INFINITY;
                        if (n < N && live) {
                            const int mm = mk[n];
                            if (fl || mm == 0) {
                                u = ulog[n];
                                if (a.use_tanh) u = a.tanh_C * tanhf(u);
                                if (a.mask_pointer) u = __fsub_rn(u, __fmul_rn(kBig, (float)mm));
                            }
                            ulog[n] = u;
                            if (tracing) {
                                if (a.tr_logits) a.tr_logits[tro + n] = u;
                                if (a.tr_masks) a.tr_masks[tro + n] = (float)mm;
                            }
                        }
                        if (u > mx) { second = mx; mx = u; mxi = n; }      // (nodes come in increasing order: first maximum wins)
                        else second = fmaxf(second, u);
                    }
    #pragma unroll
                    for (int o = LPR / 2; o; o >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, mx, o), os = __shfl_xor_sync(kFull, second, o);
                        const int oi = __shfl_xor_sync(kFull, mxi, o);
                        second = fmaxf(fmaxf(second, os), fminf(mx, ov));
                        if (ov > mx || (ov == mx && oi < mxi)) { mx = ov; mxi = oi; }
                    }
                    int idx;
                    float psel = 1.0f;
                    // Greedy decoding takes argmax(prob) (attention_agent.py:146-147).  exp is monotone up to a few ulp, so when the
                    // runner-up logit is more than 1e-5 below the maximum its probability is strictly smaller and the arg-max of
                    // the logits IS the reference's choice; the softmax is then only evaluated if somebody asked for its values.
                    const bool need_softmax = a.mode != 0 || tracing || a.logprob_out != nullptr || !(second < mx - 1e-5f);
                    if (__any_sync(kFull, need_softmax && live)) {
                        // ---- softmax denominator in the fixed order of softmax_denominator<LPR> (rollout3): node n belongs to
                        //      virtual lane n % 32 (= sl + 8 qv), which adds its nodes n, n + 32, ... in sequence
                        float P0 = 0.0f, P1 = 0.0f, P2 = 0.0f, P3 = 0.0f;
    #pragma unroll 1
                        for (int k = 0; k < 4; ++k) {
                            const int nb = sl + 32 * k;
                            float ev[4];
    #pragma unroll
                            for (int qv = 0; qv < 4; ++qv) {
                                const int n = nb + 8 * qv;
                                const float u = (n < N && live) ? ulog[n] : -INFINITY;
                                ev[qv] = (u == -INFINITY) ? 0.0f : expf(u - mx);
                            }
                            P0 = __fadd_rn(P0, ev[0]); P1 = __fadd_rn(P1, ev[1]); P2 = __fadd_rn(P2, ev[2]); P3 = __fadd_rn(P3, ev[3]);
                        }
                        float se = __fadd_rn(__fadd_rn(P0, P2), __fadd_rn(P1, P3));
    #pragma unroll
                        for (int o = LPR / 2; o; o >>= 1) se = __fadd_rn(se, __shfl_xor_sync(kFull, se, o));
                        const float lse = logf(se);
                        // ---- probabilities (decode_step.py:125-126) and the greedy candidate
                        float bestp = -1.0f;
                        int besti = 0x7fffffff;
    #pragma unroll 1
                        for (int j = 0; j < NJ; ++j) {
                            const int n = sl + LPR * j;
                            if (n < N) {
                                const float u = live ? ulog[n] : -INFINITY;
                                const float p = (u == -INFINITY) ? 0.0f : expf((u - mx) - lse);
                                if (p > bestp) { bestp = p; besti = n; }
                                if (live) {
                                    if (a.mode != 0) ulog[n] = p;   // probabilities for the sequential CDF below
                                    if (a.tr_probs) a.tr_probs[tro + n] = p;
                                }
                            }
                        }
                        if (a.mode == 0) {   // greedy: argmax(prob), first maximal index
    #pragma unroll
                            for (int o = LPR / 2; o; o >>= 1) {
                                const float ov = __shfl_xor_sync(kFull, bestp, o);
                                const int oi = __shfl_xor_sync(kFull, besti, o);
                                if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                            }
                            idx = besti;
                            psel = bestp;
                        } else {             // inverse-CDF sampling with one (per-row) retry (attention_agent.py:148-167)
                            __syncwarp();
                            int sel = -1;
                            if (sl == 0 && live) {
                                const long long o = (long long)t * R + row;
                                int d0_, d1_;
                                draw_range(a, t, d0_, d1_);   // per-row retry, or the reference's whole-batch redraw (attention_agent.py:165-167)
                                for (int draw = d0_; draw < d1_ && sel < 0; ++draw) {
                                    const float* ubuf = draw ? a.uniforms_retry : a.uniforms;
                                    const float u = ubuf ? ubuf[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                                    // tf.cumsum: sequential fp32 in node order.  Masked nodes have probability exactly 0 and cannot be
                                    // the first node whose running sum exceeds u, so the scan runs over the row's item list (ascending
                                    // node ids; every node when the row has no un-masked one) - same sums, a third of the serial adds
                                    const unsigned char* itl = items + lrs * NP;
                                    const int nal = nact_s[lrs];
                                    float cum = 0.0f;
                                    for (int k = 0; k < nal; ++k) {
                                        const int n = itl[k];
                                        cum = __fadd_rn(cum, ulog[n]);
                                        if (cum > u) { sel = n; break; }
                                    }
                                }
                                if (sel < 0 && a.fail_steps && d0_ == 0) a.fail_steps[t] = 1;   // some row of the batch needs the redraw: the host re-runs with this step redrawn
                            }
                            if (sel < 0) sel = 0;
                            idx = __shfl_sync(kFull, sel, sub * LPR);
                            __syncwarp();
                            psel = ulog[idx < N ? idx : 0];
                            __syncwarp();
                        }
                    } else {
                        idx = mxi;
                    }
                    if (idx < 0 || idx >= N) idx = 0;   // rows that are not live carry garbage
                    // ---- Env.step (vrptw_utils.py:254-358): demand, load, time, position
                    const float4 fsel = __ldcg(frow + idx);
                    const float sx = fsel.x, sy = fsel.y;
                    const float sb = TW ? fsel.z : 0.0f, se_tw = TW ? fsel.w : 0.0f;
                    const float px = rs[2], py = rs[3];
                    const float d_idx = drow[idx];
                    float nz = rs[13];                       // customers with demand left: sum(demand) > 0  <=>  nz > 0
                    __syncwarp();
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
                    if (live && sl == 0) drow[idx] = d_new;
                    __syncwarp();
                    done_now = (!a.full && a.mask_pointer && idx == n_cust && nz == 0.0f) ? 1 : 0;
                    if (live && sl == 0) {
                        rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                        if (t == 0) { rs[5] = sx; rs[6] = sy; }
                        else rs[4] = __fadd_rn(rs[4], ts);
                        reinterpret_cast<int*>(rs)[8] = done_now;
                        rs[9] = sx; rs[10] = sy; rs[11] = sb; rs[12] = se_tw; rs[13] = nz; rs[15] = dep;
                        a.idx_out[(long long)t * R + row] = idx;
                        if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
                    }
                    // ---- new mask (vrptw_utils.py:328-348); node features come from L2, eight nodes' loads in flight at a time
                    // (when nobody reads the mask VALUES, a customer without demand or an empty vehicle is masked whatever its time
                    //  window says and its features are not fetched - unless no node at all stays un-masked: then the choice
                    //  depends on the mask counts and the pass is repeated exactly)
    #pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass) {
                        const bool exact = a.full || a.fin_mask != nullptr || pass == 1;
                        if (live) {
    #pragma unroll 1
                            for (int jb = 0; jb < NJ; jb += 8) {
                                float nx[8], ny[8], ne[8];
    #pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int n = sl + LPR * (jb + j);
                                    const int nn = n < N ? n : N - 1;
                                    const float4* fp = frow + nn;
                                    nx[j] = ny[j] = ne[j] = 0.0f;
                                    if (TW && nn != n_cust && (exact || (nload != 0.0f && drow[nn] != 0.0f))) {
                                        const float4 fv = ldcg4_volatile(fp);
                                        nx[j] = fv.x; ny[j] = fv.y; ne[j] = fv.w;
                                    }
                                }
    #pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int n = sl + LPR * (jb + j);
                                    const int nn = n < N ? n : N - 1;
                                    const int mm = node_mask<TW>(nn, n_cust, drow[nn], nload, ntime, sx, sy, nx[j], ny[j], ne[j], dep, nz);
                                    if (n < N) mk[n] = (unsigned char)mm;
                                }
                            }
                        }
                        __syncwarp();
                        const int nun = build_items(lrs, rvalid, was_done || done_now, sub, sl);
                        if (exact || !__any_sync(kFull, live && !done_now && nun == 0)) break;
                    }
                    if (fastk) { __syncwarp(); build_alive(lrs, rvalid, sub, sl); }
                }
                warp_done &= __all_sync(kFull, !live || done_now);
                VRP_CLK(13);
            }
            my_done = warp_done;
            VRP_CLK(5);
        }
        load_lstm_consts();
        const int all_done = named_bar_and(5, NT, my_done);   // end of the step: one barrier that also tells whether every row is done
        VRP_CLK(6);
        if (all_done) { ++t; break; }
    }
    if (tid == 0) {
        *stop_s = 1;
        __threadfence_block();
        mbar_arrive(bar_x);
    }
    if (DEEP) {   // wake the score-MMA issuers (they wait for the F tile of the group's next tile)
        named_bar(5, NT);
        if (SPLIT) { if (warp < 16 && ((warp >> 2) & 1) == 0 && lane == 0) mbar_arrive(bar_grp + 80 * (warp >> 3) + 64 + 8 * (c1_nq & 1)); }
        else if (warp < 8 && lane == 0) mbar_arrive(bar_grp + 80 * (warp >> 2) + 64 + 8 * (c1_nq & 1));
    }
    const int t_end = t;

    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i;
        if (lr >= IBv) continue;
        const long long row = b0 + lr;
        for (int tt = t_end + lane; tt < T; tt += 32) {
            a.idx_out[(long long)tt * R + row] = n_cust;
            if (a.logprob_out) a.logprob_out[(long long)tt * R + row] = 0.0f;
        }
        const float* rs = rowst + lr * kRowState;
        if (lane == 0) a.cost_out[row] = __fadd_rn(rs[4], dist_rn(rs[2], rs[3], rs[5], rs[6]));
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
    if (a.stats) {
        if (n_eval) atomicAdd(a.stats + 0, n_eval);
        if (n_rowsteps) atomicAdd(a.stats + 1, n_rowsteps);
        if (tid == 0) {
            atomicAdd(a.stats + 2, (unsigned long long)t_end * RB);
            atomicAdd(a.stats + 3, 1ull);
        }
#ifdef VRPB200_PHASE_CLOCKS
        if (tid == 64)
            for (int i = 0; i < 25; ++i) atomicAdd(a.stats + 4 + i, (unsigned long long)pc_[i]);
        if (lane == 0 && warp < 16) { atomicAdd(a.stats + 29, (unsigned long long)pc_[25]); atomicAdd(a.stats + 30, (unsigned long long)pc_[26]); atomicAdd(a.stats + 31, (unsigned long long)pc_[27]); }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem_base, kTcCols);
}

}  // namespace vrpb200
