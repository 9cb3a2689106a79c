// rollout3_kernel: rollout2 with the weight stream and the LSTM/query GEMMs taken off the critical path.
//
//  * the GEMMs are issued TRANSPOSED: gates^T[gate, row] = W^T[gate, k] . h^T[k, row] with M = 128 gate columns of one
//    gate type (i | j | f | o -> four M-tiles), N = RB rows, K = 8 per instruction (TF32, 3x split).  tcgen05 runs
//    M = 128 at full rate (M = 64 runs at half), the accumulator of unit u sits in TMEM lane u, so the LSTM epilogue
//    thread (= one hidden unit, RB/4 rows) finds i, j, f, o of its unit in its own lane: no shuffles, all 32 lanes busy;
//  * h lives in shared memory as the K-major B operand X[row, k] (hi | lo); the K = 4 input part (features of the
//    previously chosen node) and the biases are added with 16 FMAs per cell in the epilogue, so the GEMM that needs
//    the decision of step t is gone: gates(t+1) = h(t) . W_h can be computed while step t is still choosing;
//  * a 17th warp streams the weights (cp.async.bulk -> mbarrier ring) and issues every LSTM/query MMA:
//    q(t) first (the critical path waits for it), then the 512 KB of gates(t+1) in the background of phases C0/C1/C2;
//  * TMEM: columns 0..255 pointer-score tiles (q^T re-uses 0..63 before them), 256..511 gates^T.
#pragma once
#include "common.cuh"

namespace vrpb200 {

constexpr int kV3GateCol = 256;

struct Smem3 {
    int ring, X, qq, F, logits, items, bc, vsm, feat, dem, mask, rs, meta, bar, total, stages;
};

__host__ __device__ inline Smem3 make_layout3(int RB, int IB, int N, int FD, int stages) {
    Smem3 L;
    const int NP = align_up(N, 4);
    int off = 0;
    L.stages = stages;
    L.ring = off;   off += stages * kSlabBytes;
    L.X = off;      off += 2 * RB * kH * 4;               // X_hi | X_lo: h as the [row][k] K-major B operand, [row/8][k/4][row%8][k%4]
    L.qq = off;     off += RB * kH * 4;
    L.F = off;      off += 8 * kC1TileBytes;
    L.logits = off; off += RB * NP * 4;
    L.items = off;  off += align_up(RB * NP * 2, 16);
    L.bc = off;     off += 2 * 4096;
    L.vsm = off;    off += 512;
    L.feat = off;   // node features stay in global memory (L2-resident, read-only): the space buys more rows per CTA
    (void)IB; (void)FD;
    L.dem = off;    off += RB * NP * 4;
    L.mask = off;   off += align_up(RB * NP, 16);
    L.rs = off;     off += RB * kRowState * 4;
    L.meta = off;   off += align_up((6 * RB + 32) * 4, 16);   // nact, flags, beam: sel_idx, sel_par, sel_lp, prevlog; done[16], stop
    L.bar = off;    off += 192;
    L.total = off;
    return L;
}


template <int RB, bool TW, int STAGES>
__global__ void __launch_bounds__(kV3Threads, 1) rollout3_kernel(const RolloutArgs a) {
    constexpr int NT = 512, RPW = RB / 16, RPT = RB / 4, FD = TW ? 4 : 2;
    constexpr int LPR = RB <= 16 ? 32 : (RB <= 32 ? 16 : 8), NPL = 128 / LPR;   // C2: lanes per row, nodes per lane
    
    static_assert(RB % 16 == 0 && RB <= 64, "rows");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), IB = a.IB, W = a.W, T = a.T, n_cust = a.n_cust;
    const Smem3 L = make_layout3(RB, IB, N, FD, STAGES);
    unsigned char* xt = smem + L.X;                           // X_hi, then X_lo at + RB*512
    float* qq = reinterpret_cast<float*>(smem + L.qq);
    unsigned char* ftile = smem + L.F;
    float* logits = reinterpret_cast<float*>(smem + L.logits);
    unsigned short* items = reinterpret_cast<unsigned short*>(smem + L.items);
    float* vsm = reinterpret_cast<float*>(smem + L.vsm);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    int* nact_s = reinterpret_cast<int*>(smem + L.meta);
    int* flag_s = nact_s + RB;
    int* sel_idx_s = flag_s + RB;                             // beam search: choice of every new beam row
    int* sel_par_s = sel_idx_s + RB;
    float* sel_lp_s = reinterpret_cast<float*>(sel_par_s + RB);
    float* prevlog_s = sel_lp_s + RB;                         // log-probability of the beam so far
    int* done_s = reinterpret_cast<int*>(prevlog_s + RB);     // [16] per-warp "all my rows are finished"
    volatile int* stop_s = done_s + 16;
    const uint32_t bar0 = smem_u32(smem + L.bar);            // full[0..3] at +0, empty[0..3] at +32
    const uint32_t bar_empty = bar0 + 32, bar_q = bar0 + 64, bar_g = bar0 + 72, bar_x = bar0 + 80, bar_grp = bar0 + 88,
                   tmem_slot = bar0 + 120;
    const uint32_t ring_u32 = smem_u32(smem + L.ring), bc_u32 = smem_u32(smem + L.bc);
    const uint32_t x_hi = smem_u32(xt), x_lo = x_hi + RB * kH * 4;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sp = warp & 3, ub = warp >> 2;                  // compute warps: TMEM sub-partition; row block / C1 group
    const long long b0 = (long long)blockIdx.x * IB;
    const int IBv = (int)min((long long)IB, a.B - b0);
    const long long R = a.B * W;
    const int D = a.D;
    const float* gfeat = a.input + b0 * (long long)N * D;      // [instance of this block][node][x, y, (b_tw, e_tw,) demand]

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_q, 1); mbar_init(bar_g, 1); mbar_init(bar_x, 1);
        for (int g = 0; g < 4; ++g) { mbar_init(bar_grp + 8 * g, 1); mbar_init(bar_grp + 8 * g + 40, 1); }
        *stop_s = 0;
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc_alloc(tmem_slot, kTcCols);
    for (int i = tid; i < 128 * 8; i += kV3Threads) {
        const int h = i >> 3, k = i & 7;
        float val = 0.0f;
        if (k < FD) val = a.att.A[k * kH + h];
        else if (k == 4) val = a.att.gd[h];
        const float hi = tf32_rna(val), lo = tf32_rna(val - hi);
        const int o = (h >> 3) * 256 + (k >> 2) * 128 + (h & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(smem + L.bc + o) = hi;
        *reinterpret_cast<float*>(smem + L.bc + 4096 + o) = lo;
    }
    if (tid < 128) vsm[tid] = a.att.v[tid];
    for (int i = tid; i < 2 * RB * kH; i += kV3Threads) reinterpret_cast<float*>(xt)[i] = 0.0f;   // h_0 = 0
    for (int i = tid; i < IBv * N; i += kV3Threads) {
        const int lb = i / N, n = i - lb * N;
        const float* src = a.input + ((b0 + lb) * N + n) * a.D;
        const float d = src[a.D - 1];
        for (int w = 0; w < W; ++w) dem[(w * IB + lb) * NP + n] = d;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.bar + 120);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * sp) << 16);
    constexpr uint32_t idescL = tc_idesc(128, RB), idescS = tc_idesc(128, 64);

    if (warp == 16) {
        // ============ streaming warp: lane 0 feeds the ring from L2, lane 1 issues every LSTM / query MMA ============
        if (lane == 0) {
            // the weights are constants: the first STAGES stages of q(t+1) are fetched as soon as ring slots free up,
            // so that the query GEMM - the one the critical path waits for - finds its operands already in place
            uint32_t c = 0;
            int pre = 0;
            auto issue = [&](int u) {
                const uint32_t st = c % STAGES;
                mbar_wait(bar_empty + 8 * st, ((c / STAGES) & 1) ^ 1);
                mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                const float* src = u < kV3UsesQ ? a.WqTc + (size_t)u * kSlabFloats : a.WgT3 + (size_t)(u - kV3UsesQ) * kSlabFloats;
                bulk_g2s(ring_u32 + st * kSlabBytes, src, kSlabBytes, bar0 + 8 * st);
                ++c;
            };
            for (int step = 0;; ++step) {
                mbar_wait(bar_x, step & 1);
                if (*stop_s) {   // loads already in flight must land before the CTA may exit
                    for (uint32_t j = c - pre; j < c; ++j) mbar_wait(bar0 + 8 * (j % STAGES), (j / STAGES) & 1);
                    break;
                }
                for (int u = pre; u < kV3UsesQ + kV3UsesG; ++u) issue(u);
                pre = STAGES < kV3UsesQ ? STAGES : kV3UsesQ;
                for (int u = 0; u < pre; ++u) issue(u);
            }
        } else if (lane == 1) {
            uint32_t c = 0;
            int step = 0;
            for (;; ++step) {
                mbar_wait(bar_x, step & 1);
                tc_fence_after();
                if (*stop_s) break;
                for (int u = 0; u < kV3UsesQ; ++u, ++c) {          // q^T[unit, row] = W_q^T . h^T  -> columns 0..RB-1
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int kc = 2 * u + kk;
                        const uint64_t w_hi = tc_desc(ring_u32 + st * kSlabBytes + kk * 8192, 128, 256);
                        const uint64_t w_lo = tc_desc(ring_u32 + st * kSlabBytes + kk * 8192 + 4096, 128, 256);
                        const uint64_t dx_hi = VRP_XDESC(x_hi, kc), dx_lo = VRP_XDESC(x_lo, kc);
                        tc_mma_tf32(tmem_base, w_lo, dx_hi, idescL, kc > 0);
                        tc_mma_tf32(tmem_base, w_hi, dx_lo, idescL, 1);
                        tc_mma_tf32(tmem_base, w_hi, dx_hi, idescL, 1);
                    }
                    tc_commit(bar_empty + 8 * st);
                }
                tc_commit(bar_q);
                for (int u = 0; u < kV3UsesG; ++u, ++c) {          // gates^T of the NEXT step, h part
                    const uint32_t st = c % STAGES;
                    mbar_wait(bar0 + 8 * st, (c / STAGES) & 1);
                    tc_fence_after();
                    const int kc = u >> 1, mp = u & 1;
                    const uint64_t dx_hi = VRP_XDESC(x_hi, kc), dx_lo = VRP_XDESC(x_lo, kc);
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        const uint64_t w_hi = tc_desc(ring_u32 + st * kSlabBytes + mi * 8192, 128, 256);
                        const uint64_t w_lo = tc_desc(ring_u32 + st * kSlabBytes + mi * 8192 + 4096, 128, 256);
                        const uint32_t d = tmem_base + kV3GateCol + 64 * (2 * mp + mi);
                        tc_mma_tf32(d, w_lo, dx_hi, idescL, kc > 0);
                        tc_mma_tf32(d, w_hi, dx_lo, idescL, 1);
                        tc_mma_tf32(d, w_hi, dx_hi, idescL, 1);
                    }
                    tc_commit(bar_empty + 8 * st);
                }
                tc_commit(bar_g);
            }
            if (step > 0) mbar_wait(bar_g, (step - 1) & 1);       // drain the last (unused) gates GEMM before TMEM is freed
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }

    // =============================== 16 compute warps ===========================================================
    for (int i = 0; i < RPW; ++i) {   // Env.reset + prologue
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        for (int n = lane; n < N; n += 32) maskb[lr * NP + n] = (n == n_cust) ? 1 : (dem[lr * NP + n] == 0.0f ? 1 : 0);
        const float* fdep = gfeat + (size_t)(lb * N + n_cust) * D;
        if (lane == 0) {
            float* rs = rowst + lr * kRowState;
            rs[0] = a.capacity; rs[1] = 0.0f; rs[2] = __ldcg(fdep + 0); rs[3] = __ldcg(fdep + 1); rs[4] = 0.0f; rs[5] = 0.0f; rs[6] = 0.0f;
            reinterpret_cast<int*>(rs)[8] = 0;
            reinterpret_cast<int*>(rs)[14] = 0;
            rs[9] = __ldcg(fdep + 0); rs[10] = __ldcg(fdep + 1); rs[11] = TW ? __ldcg(fdep + 2) : 0.0f; rs[12] = TW ? __ldcg(fdep + 3) : 0.0f;
        }
    }
    const int unit = 32 * sp + lane;                           // this thread's hidden unit in the LSTM / query epilogues
    float cst[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) cst[r] = 0.0f;
    named_bar(5, NT);

    // Env mask of the rows that moved in the previous step (vrptw_utils.py:328-348).  It is not needed before the item
    // list of the next step, so it runs after the LSTM epilogue has released the query GEMM: its L2 feature reads and
    // square roots hide behind the wait for q instead of sitting on the critical path of the choice.
    auto update_masks = [&]() {
        const int sub = lane / LPR, sl = lane % LPR;
        const int lr = warp * (32 / LPR) + sub;
        const int wb = lr / IB, lb = lr - wb * IB;
        const bool rvalid = lr < RB && wb < W && lb < IBv;
        const int lrs = rvalid ? lr : 0, lbs = rvalid ? lb : 0;
        float* rs = rowst + lrs * kRowState;
        const bool dirty = rvalid && reinterpret_cast<const int*>(rs)[14] != 0;
        const float nload = rs[0], ntime = rs[1], sx = rs[2], sy = rs[3], sumdem = rs[13], dep = rs[15];
        const float* frow = gfeat + (size_t)lbs * N * D;
        const float* drow = dem + lrs * NP;
        unsigned char* mk = maskb + lrs * NP;
        __syncwarp();
        if (dirty) {
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int n = sl + LPR * j;
                const int nn = n < N ? n : N - 1;
                float nx = 0.f, ny = 0.f, ne = 0.f;
                if (TW) { nx = __ldcg(frow + nn * D); ny = __ldcg(frow + nn * D + 1); ne = __ldcg(frow + nn * D + 3); }
                const int m = node_mask<TW>(nn, n_cust, drow[nn], nload, ntime, sx, sy, nx, ny, ne, dep, sumdem);
                if (n < N) mk[n] = (unsigned char)m;
            }
            if (sl == 0) reinterpret_cast<int*>(rs)[14] = 0;
        }
        __syncwarp();
    };

#ifdef VRPB200_PHASE_CLOCKS
    long long pc_[6] = {0, 0, 0, 0, 0, 0}, pc_last_ = clock64();
#endif
    uint32_t gpar = 0;
    unsigned long long n_eval = 0, n_rowsteps = 0;
    int t = 0;
    for (; t < T; ++t) {
        // ================= LSTM epilogue: gates = gates^T(TMEM, h part) + x part + bias ; c, h' ; h' -> X ==========
        {
            float wx[4][4], bgt[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                bgt[m] = __ldg(a.bgT + m * kH + unit);
#pragma unroll
                for (int k = 0; k < 4; ++k) wx[k][m] = __ldg(a.WxT + (k * 4 + m) * kH + unit);
            }
            if (t > 0) {
                mbar_wait(bar_g, (t - 1) & 1);
                tc_fence_after();
            }
            VRP_CLK(0);
#pragma unroll
            for (int c4 = 0; c4 < RPT / 4; ++c4) {
                const int r0 = ub * RPT + 4 * c4;
                float g[4][4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if (t > 0) tc_ld4(tmem_lane + kV3GateCol + 64 * m + r0, g[m]);
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
                    const float cn = __fadd_rn(__fmul_rn(cst[c4 * 4 + i], sigmoid_fast(gg[2] + a.forget_bias)),
                                               __fmul_rn(sigmoid_fast(gg[0]), tanh_fast(gg[1])));
                    cst[c4 * 4 + i] = cn;
                    const float hn = __fmul_rn(tanh_fast(cn), sigmoid_fast(gg[3]));
                    hi[i] = tf32_rna(hn);
                    lo[i] = hn - hi[i];
                }
                // 4 lanes (units 4q..4q+3) x 4 rows -> transpose so that each lane owns one row and 4 consecutive k
                transpose4(hi, lane);
                transpose4(lo, lane);
                {
                    const int rr = r0 + (lane & 3);
                    const int off = (rr >> 3) * 4096 + (unit >> 2) * 128 + (rr & 7) * 16;
                    *reinterpret_cast<float4*>(xt + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(xt + RB * kH * 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            named_bar(5, NT);
            if (tid == 0) mbar_arrive(bar_x);      // h' is in place: the streaming warp may issue q(t) and gates(t+1)
        }
        update_masks();   // masks of the previous step's moves, hidden behind the query GEMM
        named_bar(5, NT); // (the item list below reads masks written by other warps)
        VRP_CLK(1);
        // ================= C0: item list ============================================================================
        for (int i = 0; i < RPW; ++i) {
            const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
            int na = 0, fl = 0;
            if (wb < W && lb < IBv && !reinterpret_cast<const int*>(rowst + lr * kRowState)[8]) {
                const unsigned char* mk = maskb + lr * NP;
                int nun = 0, nval = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    const int m = (n < N) ? mk[n] : 255;
                    nun += __popc(__ballot_sync(kFull, m == 0));
                    nval += __popc(__ballot_sync(kFull, n < N));
                }
                fl = (a.full || nun == 0 || !a.mask_pointer) ? 1 : 0;
                na = fl ? nval : nun;
            }
            if (lane == 0) { nact_s[lr] = na; flag_s[lr] = fl; }
        }
        tc_fence_before();
        named_bar(5, 512);
        tc_fence_after();
        int m_tot;
        {
            // exclusive prefix of nact over rows (every warp redundantly), then this warp's rows write their items
            const int n0 = (lane < RB) ? nact_s[lane] : 0, n1 = (lane + 32 < RB) ? nact_s[lane + 32] : 0;
            int s0 = n0, s1 = n1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t0 = __shfl_up_sync(kFull, s0, o), t1 = __shfl_up_sync(kFull, s1, o);
                if (lane >= o) { s0 += t0; s1 += t1; }
            }
            const int tot0 = __shfl_sync(kFull, s0, 31);
            m_tot = tot0 + __shfl_sync(kFull, s1, 31);
            const int ex0 = s0 - n0, ex1 = tot0 + s1 - n1;
            for (int i = 0; i < RPW; ++i) {
                const int lr = warp * RPW + i;
                const int na = nact_s[lr];
                if (na == 0) continue;
                const int off = (lr < 32) ? __shfl_sync(kFull, ex0, lr) : __shfl_sync(kFull, ex1, lr - 32);
                const int fl = flag_s[lr];
                const unsigned char* mk = maskb + lr * NP;
                int pos = off;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = lane + 32 * j;
                    const bool act = (n < N) && (fl || mk[n] == 0);
                    const unsigned bal = __ballot_sync(kFull, act);
                    if (act) items[pos + __popc(bal & ((1u << lane) - 1))] = (unsigned short)((lr << 7) | n);
                    pos += __popc(bal);
                }
            }
        }
        // ================= query epilogue: base'[row][unit] = 2log2e (q + b + load g_ld + time g_t) ===================
        {
            const float bqv = __ldg(a.att.bq + unit), glv = __ldg(a.att.gl + unit), gtv = __ldg(a.att.gt + unit);
            mbar_wait(bar_q, t & 1);
            tc_fence_after();
            VRP_CLK(2);
#pragma unroll
            for (int c4 = 0; c4 < RPT / 4; ++c4) {
                const int r0 = ub * RPT + 4 * c4;
                float q4[4];
                tc_ld4(tmem_lane + r0, q4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* rs = rowst + (r0 + i) * kRowState;
                    float s = fmaf(rs[0], glv, q4[i] + bqv);
                    if (TW) s = fmaf(rs[1], gtv, s);
                    qq[(r0 + i) * kH + unit] = s * kTwoLog2e;
                }
            }
        }
        named_bar(5, 512);
        VRP_CLK(3);

        // ================= C1: scores, one 128-item tile per group of 4 warps at a time ===========================
        {
            const int g = ub;
            unsigned char* fhi = ftile + g * 2 * kC1TileBytes;
            const uint32_t f_hi = smem_u32(fhi), f_lo = f_hi + kC1TileBytes;
            const uint32_t d_g = tmem_base + 64 * g;
            const int ntiles = (m_tot + 127) >> 7;
            for (int tile = g; tile < ntiles; tile += 4) {
                const int m = 32 * sp + lane, it = tile * 128 + m;
                const bool valid = it < m_tot;
                int r = 0, n = 0;
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const int code = items[it];
                    r = code >> 7; n = code & 127;
                    const int lb = r % IB;
                    const float* fr = gfeat + ((size_t)lb * N + n) * D;
                    f[0] = __ldcg(fr); f[1] = __ldcg(fr + 1);
                    if (TW) { f[2] = __ldcg(fr + 2); f[3] = __ldcg(fr + 3); }
                    f[4] = dem[r * NP + n];
                }
                {
                    float hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { hi[k] = tf32_rna(f[k]); lo[k] = f[k] - hi[k]; }
                    const int off = (m >> 3) * 256 + (m & 7) * 16;
                    *reinterpret_cast<float4*>(fhi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(fhi + off + 128) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<float4*>(fhi + kC1TileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(fhi + kC1TileBytes + off + 128) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                }
                fence_proxy_async();
                tc_fence_before();
                named_bar(1 + g, 128);
                const float* brow = qq + r * kH;
                float acc = 0.0f;
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    if (sp == 0 && lane == 0) {
                        tc_fence_after();
                        const uint64_t dfh = tc_desc(f_hi, 128, 256), dfl = tc_desc(f_lo, 128, 256);
                        const uint64_t dbh = tc_desc(bc_u32 + hh * 2048, 128, 256), dbl = tc_desc(bc_u32 + 4096 + hh * 2048, 128, 256);
                        tc_mma_tf32(d_g, dfl, dbh, idescS, 0);
                        tc_mma_tf32(d_g, dfh, dbl, idescS, 1);
                        tc_mma_tf32(d_g, dfh, dbh, idescS, 1);
                        tc_commit(bar_grp + 8 * g);
                    }
                    __syncwarp();
                    mbar_wait(bar_grp + 8 * g, gpar);
                    gpar ^= 1;
                    tc_fence_after();
                    // 64 columns in 4 chunks of 16, double-buffered: the TMEM read of chunk c+1 is in flight while the
                    // MUFU pipe works on chunk c
                    {
                        uint32_t va[16], vb[16];
                        tc_ld16_issue(tmem_lane + 64 * g, va);
                        tc_ld_wait16(va);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t(&cur)[16] = (c & 1) ? vb : va;
                            uint32_t(&nxt)[16] = (c & 1) ? va : vb;
                            if (c < 3) tc_ld16_issue(tmem_lane + 64 * g + 16 * (c + 1), nxt);
                            const int h0 = 64 * hh + 16 * c;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(brow + h0 + 4 * k4);
                                const float4 v4 = *reinterpret_cast<const float4*>(vsm + h0 + 4 * k4);
                                const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float rr = rcp_approx(1.0f + ex2_approx(__uint_as_float(cur[4 * k4 + j]) + bb[j]));
                                    acc = fmaf(vv[j], fmaf(-2.0f, rr, 1.0f), acc);
                                }
                            }
                            if (c < 3) tc_ld_wait16(nxt);
                        }
                    }
                    tc_fence_before();
                    named_bar(1 + g, 128);   // every warp of the group has drained D_g (and, after hh = 1, the F tile)
                }
                if (valid) logits[r * NP + n] = acc;
                n_eval += valid ? 1 : 0;
            }
        }
        tc_fence_before();
        named_bar(5, 512);
        tc_fence_after();
        VRP_CLK(4);

        // ================= C2: per row - mask, log-softmax, selection, Env.step =====================================
        // This part is a long dependent chain per row (reductions, exp/log, sqrt): LPR lanes per row (8, 16 or 32 - every compute warp busy),
        // every row of the block in flight at once; lane sl owns nodes sl, sl+LPR, ...
        int warp_done;
        if (a.mode == 2) {
            // ---------- beam search (model/attention_agent.py:169-205): rows of one instance compete ----------
            const int sub = lane / LPR, sl = lane % LPR;
            const int lr = warp * (32 / LPR) + sub;
            const int wb = lr / IB, lb = lr - wb * IB;
            const bool rvalid = lr < RB && wb < W && lb < IBv;
            const int lrs = rvalid ? lr : 0, lbs = rvalid ? lb : 0;
            float* rs = rowst + lrs * kRowState;
            const long long row = (long long)wb * a.B + b0 + lb;
            const bool was_done = rvalid && reinterpret_cast<int*>(rs)[8] != 0;
            const bool live = rvalid && !was_done;
            if (live && sl == 0) ++n_rowsteps;
            float* ulog = logits + lrs * NP;
            const float* frow = gfeat + (size_t)lbs * N * D;
            {   // (a) every row: log(prob[n]) for all nodes, in place of the logits (finished rows: depot with prob 1)
                unsigned char* mk = maskb + lrs * NP;
                const int fl = flag_s[lrs];
                float lg[NPL];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int n = sl + LPR * j;
                    float u = -INFINITY;
                    if (n < N && live) {
                        const int m = mk[n];
                        if (fl || m == 0) {
                            u = ulog[n];
                            if (a.use_tanh) u = a.tanh_C * tanhf(u);
                            if (a.mask_pointer) u = __fsub_rn(u, __fmul_rn(kBig, (float)m));
                        }
                    }
                    if (n == n_cust && was_done) u = 0.0f;
                    lg[j] = u;
                    mx = fmaxf(mx, u);
                }
#pragma unroll
                for (int o = LPR / 2; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
                const float se = softmax_denominator<LPR>(lg, mx);
                const float lse = logf(se);
                const bool tracing = a.tr_logits || a.tr_probs || a.tr_masks;
                const long long tro = ((long long)t * R + row) * N;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int n = sl + LPR * j;
                    const float p = (lg[j] == -INFINITY) ? 0.0f : expf((lg[j] - mx) - lse);
                    if (n < N && rvalid) {
                        ulog[n] = logf(p);      // tf.log(prob): -inf for masked nodes
                        if (tracing && live) {
                            if (a.tr_logits) a.tr_logits[tro + n] = lg[j];
                            if (a.tr_probs) a.tr_probs[tro + n] = p;
                            if (a.tr_masks) a.tr_masks[tro + n] = (float)mk[n];
                        }
                    }
                }
            }
            named_bar(5, NT);
            // (b) per instance (one warp): W times the best (log-prob so far + log prob) over W x N candidates;
            //     ties -> lower flat index (tf.nn.top_k); chosen entries are struck out with NaN
            for (int ib = warp; ib < IBv; ib += 16) {
                const int wc = (t == 0) ? 1 : W;       // step 0: every beam is a copy of beam 0 (attention_agent.py:171)
                const int ncand = wc * N;
                constexpr int CPL = 20;                 // candidates per lane that fit registers (W x N <= 640: e.g. 10 beams x 51 nodes)
                if (ncand <= 32 * CPL) {
                    // candidates in registers (flat index c = lane + 32 j), one local (value, index) maximum per lane; per beam
                    // slot one warp arg-max, then only the winning lane strikes its candidate and rescans its registers
                    float cv[CPL];
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        const int c = lane + 32 * j;
                        float v = __int_as_float(0x7fc00000);          // NaN: never selected
                        if (c < ncand) {
                            const int w = c / N, n = c - w * N;
                            v = logits[(w * IB + ib) * NP + n];
                            if (t > 0) v = __fadd_rn(v, prevlog_s[w * IB + ib]);
                        }
                        cv[j] = v;
                    }
                    auto local_best = [&](float& bv, int& bc) {
                        bv = -INFINITY; bc = 0x7fffffff;
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            const int c = lane + 32 * j;
                            if (cv[j] > bv || (cv[j] == bv && c < bc)) { bv = cv[j]; bc = c; }   // (increasing c: first maximum wins)
                        }
                    };
                    float lv; int lc;
                    local_best(lv, lc);
                    for (int k = 0; k < W; ++k) {
                        float bv = lv; int bc = lc;
#pragma unroll
                        for (int o = 16; o; o >>= 1) {
                            const float ov = __shfl_xor_sync(kFull, bv, o);
                            const int oc = __shfl_xor_sync(kFull, bc, o);
                            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
                        }
                        if (bc == 0x7fffffff) bc = 0;      // cannot happen (candidates are never all NaN), keeps indices in range
                        const int w = bc / N, n = bc - w * N;
                        if ((bc & 31) == lane) {           // the owner of the winner
                            const int jj = bc >> 5;
#pragma unroll
                            for (int j = 0; j < CPL; ++j) if (j == jj) cv[j] = __int_as_float(0x7fc00000);
                            local_best(lv, lc);
                        }
                        if (lane == 0) {
                            const int nr = k * IB + ib;
                            sel_idx_s[nr] = n;
                            sel_par_s[nr] = w;
                            sel_lp_s[nr] = logits[(w * IB + ib) * NP + n];
                            reinterpret_cast<float*>(ftile)[nr] = bv; // new log-probability of beam nr, staged
                        }
                    }
                    __syncwarp();
                } else
                for (int k = 0; k < W; ++k) {
                    float bv = -INFINITY;
                    int bc = 0x7fffffff;
                    for (int c = lane; c < ncand; c += 32) {
                        const int w = c / N, n = c - w * N;
                        float v = logits[(w * IB + ib) * NP + n];
                        if (t > 0) v = __fadd_rn(v, prevlog_s[w * IB + ib]);
                        if (v > bv || (v == bv && c < bc)) { bv = v; bc = c; }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, bv, o);
                        const int oc = __shfl_xor_sync(kFull, bc, o);
                        if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
                    }
                    if (bc == 0x7fffffff) bc = 0;      // cannot happen (candidates are never all NaN), keeps indices in range
                    const int w = bc / N, n = bc - w * N;
                    if (lane == 0) {
                        const int nr = k * IB + ib;
                        float* cell = logits + (w * IB + ib) * NP + n;
                        sel_idx_s[nr] = n;
                        sel_par_s[nr] = w;
                        sel_lp_s[nr] = *cell;
                        *reinterpret_cast<volatile float*>(cell) = __int_as_float(0x7fc00000);
                        reinterpret_cast<float*>(ftile)[nr] = bv; // new log-probability of beam nr, staged
                    }
                    __syncwarp();
                }
            }
            named_bar(5, NT);
            // (c) gather the parent's environment state into scratch (the F tiles are idle now)
            const int SS = NP * 4 + align_up(NP, 4) + kRowState * 4;
            unsigned char* scratch = ftile + 1024 + lrs * SS;
            float* sdem = reinterpret_cast<float*>(scratch);
            unsigned char* smk = scratch + NP * 4;
            float* srs = reinterpret_cast<float*>(scratch + NP * 4 + align_up(NP, 4));
            const int idx = rvalid ? sel_idx_s[lrs] : 0;
            const int par = rvalid ? sel_par_s[lrs] : 0;
            if (rvalid) {   // (rows without an instance alias row 0 for addressing only and must not write)
                const int pr = par * IB + lbs;
                for (int n = sl; n < N; n += LPR) {
                    sdem[n] = dem[pr * NP + n];
                    smk[n] = maskb[pr * NP + n];
                }
                if (sl < kRowState) srs[sl] = rowst[pr * kRowState + sl];
                if (LPR < kRowState && sl + LPR < kRowState) srs[sl + LPR] = rowst[pr * kRowState + sl + LPR];
            }
            named_bar(5, NT);
            // (d) Env.step on the gathered state (vrptw_utils.py:262-348); LSTM state stays with the slot
            {
                const bool p_done = reinterpret_cast<const int*>(srs)[8] != 0;
                const float load = srs[0], time = TW ? srs[1] : 0.0f;
                const float* fsel = frow + idx * D;
                const float sx = __ldcg(fsel + 0), sy = __ldcg(fsel + 1);
                const float px = srs[2], py = srs[3];
                const float d_idx = sdem[idx];
                const float d_sat = fminf(d_idx, load);
                float nload = __fsub_rn(load, d_sat);
                const float ts = dist_rn(px, py, sx, sy);
                float ntime = 0.0f;
                if (TW) ntime = fmaxf(__fadd_rn(time, ts), __ldcg(fsel + 2));
                const float dep = (idx == n_cust) ? 1.0f : 0.0f;
                nload = __fadd_rn(__fmul_rn(nload, 1.0f - dep), __fmul_rn(dep, a.capacity));
                if (TW) ntime = __fmul_rn(ntime, 1.0f - dep);
                float sumdem = 0.0f;
                float dn[NPL];
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int n = sl + LPR * j;
                    dn[j] = (n < N) ? sdem[n] : 0.0f;
                    if (n == idx) dn[j] = __fsub_rn(dn[j], d_sat);
                    sumdem += dn[j];
                }
#pragma unroll
                for (int o = LPR / 2; o; o >>= 1) sumdem += __shfl_xor_sync(kFull, sumdem, o);
                if (rvalid) {
#pragma unroll
                    for (int j = 0; j < NPL; ++j) {
                        const int n = sl + LPR * j;
                        if (n < N) dem[lr * NP + n] = dn[j];
                    }
                }
                const int done_now = (!a.full && a.mask_pointer && idx == n_cust && sumdem == 0.0f) ? 1 : 0;
                if (rvalid && sl == 0) {
                    rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                    if (t == 0) { rs[4] = 0.0f; rs[5] = sx; rs[6] = sy; }
                    else { rs[4] = __fadd_rn(srs[4], ts); rs[5] = srs[5]; rs[6] = srs[6]; }
                    reinterpret_cast<int*>(rs)[8] = done_now;
                    rs[13] = sumdem; reinterpret_cast<int*>(rs)[14] = 1; rs[15] = dep;   // mask update is deferred
                    rs[9] = sx; rs[10] = sy; rs[11] = TW ? __ldcg(fsel + 2) : 0.0f; rs[12] = TW ? __ldcg(fsel + 3) : 0.0f;
                    prevlog_s[lr] = reinterpret_cast<float*>(ftile)[lr];
                    a.idx_out[(long long)t * R + row] = idx;
                    if (a.parent_out) a.parent_out[(long long)t * R + row] = par;
                    if (a.logprob_out) a.logprob_out[(long long)t * R + row] = sel_lp_s[lr];
                }
                (void)p_done;
                warp_done = __all_sync(kFull, !rvalid || done_now);
            }
        } else {
            const int sub = lane / LPR, sl = lane % LPR;
            const int lr = warp * (32 / LPR) + sub;
            const int wb = lr / IB, lb = lr - wb * IB;
            const bool rvalid = lr < RB && wb < W && lb < IBv;
            float* rs = rowst + (rvalid ? lr : 0) * kRowState;
            const long long row = (long long)wb * a.B + b0 + lb;
            const bool was_done = rvalid && reinterpret_cast<int*>(rs)[8] != 0;
            const bool live = rvalid && !was_done;
            if (was_done && sl == 0) {
                a.idx_out[(long long)t * R + row] = n_cust;
                if (a.logprob_out) a.logprob_out[(long long)t * R + row] = 0.0f;
            }
            if (live && sl == 0) ++n_rowsteps;
            const int lrs = rvalid ? lr : 0, lbs = rvalid ? lb : 0;
            const float load = rs[0], time = TW ? rs[1] : 0.0f;
            unsigned char* mk = maskb + lrs * NP;
            const float* frow = gfeat + (size_t)lbs * N * D;
            float* drow = dem + lrs * NP;
            float* ulog = logits + lrs * NP;
            const int fl = flag_s[lrs];
            float lg[NPL];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int n = sl + LPR * j;
                float u = -INFINITY;
                if (n < N) {
                    const int m = mk[n];
                    if (fl || m == 0) {
                        u = ulog[n];
                        if (a.use_tanh) u = a.tanh_C * tanhf(u);
                        if (a.mask_pointer) u = __fsub_rn(u, __fmul_rn(kBig, (float)m));
                    }
                }
                lg[j] = u;
                mx = fmaxf(mx, u);
            }
#pragma unroll
            for (int o = LPR / 2; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
            const float se = softmax_denominator<LPR>(lg, mx);
            const float lse = logf(se);
            float bestp = -1.0f;
            int besti = 0x7fffffff;
            const bool tracing = a.tr_logits || a.tr_probs || a.tr_masks;
            const long long tro = ((long long)t * R + row) * N;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int n = sl + LPR * j;
                const float p = (lg[j] == -INFINITY) ? 0.0f : expf((lg[j] - mx) - lse);
                if (n < N && p > bestp) { bestp = p; besti = n; }
                if (n < N && live) {
                    if (a.mode != 0) ulog[n] = p;   // probabilities for the sequential CDF below
                    if (tracing) {
                        if (a.tr_logits) a.tr_logits[tro + n] = lg[j];
                        if (a.tr_probs) a.tr_probs[tro + n] = p;
                        if (a.tr_masks) a.tr_masks[tro + n] = (float)mk[n];
                    }
                }
            }
            int idx;
            float psel;
            if (a.mode == 0) {   // greedy: argmax(prob), first maximal index
#pragma unroll
                for (int o = LPR / 2; o; o >>= 1) {
                    const float ov = __shfl_xor_sync(kFull, bestp, o);
                    const int oi = __shfl_xor_sync(kFull, besti, o);
                    if (ov > bestp || (ov == bestp && oi < besti)) { bestp = ov; besti = oi; }
                }
                idx = besti;
                psel = bestp;
            } else {             // inverse-CDF sampling with one (per-row) retry
                __syncwarp();
                int sel = -1;
                if (sl == 0 && live) {
                    const long long o = (long long)t * R + row;
                    int d0_, d1_;
                    draw_range(a, t, d0_, d1_);   // per-row retry, or the reference's whole-batch redraw (attention_agent.py:165-167)
                    for (int draw = d0_; draw < d1_ && sel < 0; ++draw) {
                        const float* ubuf = draw ? a.uniforms_retry : a.uniforms;
                        const float u = ubuf ? ubuf[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                        float cum = 0.0f;   // tf.cumsum: sequential fp32 (masked nodes add exactly 0)
                        for (int n = 0; n < N; ++n) {
                            cum = __fadd_rn(cum, ulog[n]);
                            if (cum > u) { sel = n; break; }
                        }
                    }
                    if (sel < 0 && a.fail_steps && d0_ == 0) a.fail_steps[t] = 1;   // some row of the batch needs the redraw: the host re-runs with this step redrawn
                }
                if (sel < 0) sel = 0;
                idx = __shfl_sync(kFull, sel, sub * LPR);
                __syncwarp();
                psel = ulog[idx];
                __syncwarp();
            }
            if (idx < 0 || idx >= N) idx = 0;   // rows that are not live carry garbage
            const float* fsel = frow + idx * D;
            const float sx = __ldcg(fsel + 0), sy = __ldcg(fsel + 1);
            const float px = rs[2], py = rs[3];
            const float d_idx = drow[idx];
            __syncwarp();
            const float d_sat = fminf(d_idx, load);
            float nload = __fsub_rn(load, d_sat);
            const float ts = dist_rn(px, py, sx, sy);
            float ntime = 0.0f;
            if (TW) ntime = fmaxf(__fadd_rn(time, ts), __ldcg(fsel + 2));
            const float dep = (idx == n_cust) ? 1.0f : 0.0f;
            nload = __fadd_rn(__fmul_rn(nload, 1.0f - dep), __fmul_rn(dep, a.capacity));
            if (TW) ntime = __fmul_rn(ntime, 1.0f - dep);
            if (live && sl == 0) drow[idx] = __fsub_rn(d_idx, d_sat);
            __syncwarp();
            float sumdem = 0.0f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int n = sl + LPR * j;
                lg[j] = (n < N) ? drow[n] : 0.0f;     // lg[] now holds the updated demands
                sumdem += lg[j];
            }
#pragma unroll
            for (int o = LPR / 2; o; o >>= 1) sumdem += __shfl_xor_sync(kFull, sumdem, o);
            const int done_now = (!a.full && a.mask_pointer && idx == n_cust && sumdem == 0.0f) ? 1 : 0;
            if (live && sl == 0) {
                rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                if (t == 0) { rs[5] = sx; rs[6] = sy; }
                else rs[4] = __fadd_rn(rs[4], ts);
                reinterpret_cast<int*>(rs)[8] = done_now;
                rs[13] = sumdem; reinterpret_cast<int*>(rs)[14] = 1; rs[15] = dep;   // mask update is deferred
                rs[9] = sx; rs[10] = sy; rs[11] = TW ? __ldcg(fsel + 2) : 0.0f; rs[12] = TW ? __ldcg(fsel + 3) : 0.0f;
                a.idx_out[(long long)t * R + row] = idx;
                if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
            }
            warp_done = __all_sync(kFull, !live || done_now);
        }
        VRP_CLK(5);
        if (lane == 0) done_s[warp] = warp_done;
        named_bar(5, NT);
        int all_done = 1;
#pragma unroll
        for (int w = 0; w < 16; ++w) all_done &= done_s[w];
        named_bar(5, NT);
        if (all_done) { ++t; break; }
    }
    if (tid == 0) {
        *stop_s = 1;
        __threadfence_block();
        mbar_arrive(bar_x);
    }
    update_masks();            // the masks of the last moves (final-state outputs)
    named_bar(5, NT);
    const int t_end = t;

    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        const long long row = (long long)wb * a.B + b0 + lb;
        for (int tt = t_end + lane; tt < T; tt += 32) {
            a.idx_out[(long long)tt * R + row] = n_cust;
            if (a.logprob_out) a.logprob_out[(long long)tt * R + row] = 0.0f;
            if (a.parent_out) a.parent_out[(long long)tt * R + row] = wb;   // finished beams keep their order
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
        atomicAdd(a.stats + 0, n_eval);
        if (n_rowsteps) atomicAdd(a.stats + 1, n_rowsteps);   // counted by the first lane of every row
        if (tid == 0) {
            atomicAdd(a.stats + 2, (unsigned long long)t_end * RB);
            atomicAdd(a.stats + 3, 1ull);
        }
#ifdef VRPB200_PHASE_CLOCKS
        if (tid == 64)
            for (int i = 0; i < 6; ++i) atomicAdd(a.stats + 4 + i, (unsigned long long)pc_[i]);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem_base, kTcCols);
}

}  // namespace vrpb200
