// rollout2_kernel: the fused rollout with EVERY dense product on the tensor cores.
//
// Same contract as rollout_kernel (rollout_kernel.cuh) - one CTA owns RB rows for all T decode steps - but the
// pointer scores are re-organised around tcgen05:
//
//   S0  restore     A tiles (hi/lo TF32) of [h | features of the previous node] from the TMEM save area + row state
//   A   gates       [64 x 136] x [136 x 512]   tcgen05.mma kind::tf32, 3xTF32 split, D in TMEM, LSTM epilogue from TMEM
//   B   query       [64 x 128] x [128 x 128]   same; q (+ load/time terms, pre-scaled) parked in the idle weight ring
//   C0  items       every un-masked (row, node) pair of the block becomes one "item"; items are packed densely
//   C1  scores      per 128-item tile:  E[item, h] = F[item, 0:8] x Bc[0:8, h]   (F = x, y, b_tw, e_tw, demand | 0;
//                   Bc = 2log2e * [W_emb.W_ref ; g_d - g_ld]), M=128, N=64, K=8, 3xTF32, D in TMEM; the epilogue
//                   thread owns ONE item: u = sum_h v_h * tanh(E[item,h] + base[row,h]) straight out of TMEM - no
//                   cross-lane reduction, 4 FP32 + 2 MUFU per (item, h): the loop runs at the MUFU pipe rate
//   C2  select      per row: mask, log-softmax, greedy / sampled choice, Env.step, tour length (as rollout_kernel)
//
// The 70 KB shared-memory region of the LSTM A tiles is time-shared with the phase-C scratch (item tiles, logits,
// item list); h survives phase C in TMEM (columns 384..447), the LSTM cell state c in registers.
#pragma once
#include "rollout_kernel.cuh"

namespace vrpb200 {

constexpr int kC1TileBytes = 128 * 8 * 4;   // one F tile (hi or lo): 128 items x K=8, canonical K-major
constexpr int kHSaveCol = 384;              // TMEM columns 384..447: h between steps (16 values per thread)

struct Smem2 {
    int ring, U, bc, vsm, feat, dem, mask, rs, meta, bar, total;   // byte offsets
    int u_logits, u_items;                                          // inside U during phase C
};

__host__ __device__ inline Smem2 make_layout2(int RB, int IB, int N, int FD) {
    Smem2 L;
    const int NP = align_up(N, 4);
    int off = 0;
    L.ring = off; off += kStages * kSlabBytes;
    L.U = off;    off += 2 * kTcATileBytes;
    L.bc = off;   off += 2 * 4096;                  // Bc hi | lo, [N=128][K=8]
    L.vsm = off;  off += 512;
    L.feat = off; off += align_up(IB * N * FD * 4, 16);
    L.dem = off;  off += RB * NP * 4;
    L.mask = off; off += align_up(RB * NP, 16);
    L.rs = off;   off += RB * kRowState * 4;
    L.meta = off; off += align_up((2 * RB + 8) * 4, 16);   // nact[RB], flags[RB], counters
    L.bar = off;  off += 128;
    L.total = off;
    L.u_logits = 8 * kC1TileBytes;                  // 4 groups x (hi + lo)
    L.u_items = L.u_logits + RB * NP * 4;
    return L;
}

__host__ __device__ inline bool layout2_fits_U(int RB, int N) {
    const int NP = align_up(N, 4);
    return 8 * kC1TileBytes + RB * NP * 4 + RB * NP * 2 <= 2 * kTcATileBytes;
}

VRP_DEVINL void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

VRP_DEVINL void tc_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
VRP_DEVINL void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int RB, bool TW>
__global__ void __launch_bounds__(512, 1) rollout2_kernel(const RolloutArgs a) {
    constexpr int NT = 512, NW = 16, RPW = RB / NW, FD = TW ? 4 : 2;
    static_assert(RB % NW == 0 && RB <= 64, "rows");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), IB = a.IB, W = a.W, T = a.T, n_cust = a.n_cust;
    const Smem2 L = make_layout2(RB, IB, N, FD);
    float* ring = reinterpret_cast<float*>(smem + L.ring);
    float* qq = ring;                                        // phase C: base'[row][h] = 2log2e (q + load g_ld + time g_t)
    unsigned char* atile = smem + L.U;                       // phases A/B: A_hi | A_lo
    unsigned char* ftile = smem + L.U;                       // phase C: F tiles, 4 groups x (hi | lo)
    float* logits = reinterpret_cast<float*>(smem + L.U + L.u_logits);
    unsigned short* items = reinterpret_cast<unsigned short*>(smem + L.U + L.u_items);
    float* vsm = reinterpret_cast<float*>(smem + L.vsm);
    float* feat = reinterpret_cast<float*>(smem + L.feat);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    int* nact_s = reinterpret_cast<int*>(smem + L.meta);     // [RB] items of each row this step
    int* flag_s = nact_s + RB;                               // [RB] bit0: every node evaluated
    int* next_row = flag_s + RB;
    const uint32_t bar0 = smem_u32(smem + L.bar);            // full[0], full[1]
    const uint32_t bar_empty = bar0 + 16, bar_dready = bar0 + 32, bar_grp = bar0 + 40 /* 4 x 8 */, tmem_slot = bar0 + 72;
    const uint32_t ring_u32 = smem_u32(ring), bc_u32 = smem_u32(smem + L.bc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sp = warp & 3, ub = warp >> 2;                 // TMEM sub-partition; unit block (phases A/B) = group (C1)
    const long long b0 = (long long)blockIdx.x * IB;
    const int IBv = (int)min((long long)IB, a.B - b0);
    const long long R = a.B * W;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_dready, 1);
        for (int g = 0; g < 4; ++g) mbar_init(bar_grp + 8 * g, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc_alloc(tmem_slot, kTcCols);
    // constants of the scores GEMM: Bc[h][k] = 2log2e * (A[k][h] (k<4) | (g_d - g_ld)[h] (k=4) | 0), hi/lo TF32 split
    for (int i = tid; i < 128 * 8; i += NT) {
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
    for (int i = tid; i < IBv * N; i += NT) {
        const int lb = i / N, n = i - lb * N;
        const float* src = a.input + ((b0 + lb) * N + n) * a.D;
        if (TW) reinterpret_cast<float4*>(feat)[i] = make_float4(src[0], src[1], src[2], src[3]);
        else reinterpret_cast<float2*>(feat)[i] = make_float2(src[0], src[1]);
        const float d = src[a.D - 1];
        for (int w = 0; w < W; ++w) dem[(w * IB + lb) * NP + n] = d;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.bar + 72);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * sp) << 16);

    // Env.reset + prologue, one warp per row (static)
    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        for (int n = lane; n < N; n += 32) maskb[lr * NP + n] = (n == n_cust) ? 1 : (dem[lr * NP + n] == 0.0f ? 1 : 0);
        const float* fdep = feat + (size_t)(lb * N + n_cust) * FD;
        if (lane == 0) {
            float* rs = rowst + lr * kRowState;
            rs[0] = a.capacity; rs[1] = 0.0f; rs[2] = fdep[0]; rs[3] = fdep[1]; rs[4] = 0.0f; rs[5] = 0.0f; rs[6] = 0.0f;
            reinterpret_cast<int*>(rs)[8] = 0;
            rs[9] = fdep[0]; rs[10] = fdep[1]; rs[11] = TW ? fdep[2] : 0.0f; rs[12] = TW ? fdep[FD - 1] : 0.0f;   // next LSTM input
        }
    }
    {   // h_0 = 0 in the TMEM save area
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.0f;
        tc_st16(tmem_lane + kHSaveCol + ub * 16, z);
    }
    float cst[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) cst[r] = 0.0f;
    const int tc_m = 16 * sp + (lane & 15);
    const bool tc_row_ok = tc_m < RB;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

#ifdef VRPB200_PHASE_CLOCKS
    long long pc_[6] = {0, 0, 0, 0, 0, 0}, pc_last_ = clock64();
#endif
    uint32_t slabcnt = 0, dcnt = 0, gpar = 0;
    unsigned long long n_eval = 0, n_rowsteps = 0;
    constexpr uint32_t idescA = tc_idesc(kTcRows, 256), idescB = tc_idesc(kTcRows, 128), idescS = tc_idesc(128, 64);
    const uint32_t a_hi = smem_u32(atile), a_lo = a_hi + kTcATileBytes;
    int t = 0;
    for (; t < T; ++t) {
        // ================= S0: A tiles <- h (TMEM) and the features of the previously chosen node =================
        {
            float hv[16];
            tc_ld16(tmem_lane + kHSaveCol + ub * 16, hv);
            if (tc_row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { hi[j] = tf32_rna(hv[g * 4 + j]); lo[j] = hv[g * 4 + j] - hi[j]; }
                    const int off = tc_a_off(tc_m, ub * 8 + g * 2 + (lane >> 4));
                    *reinterpret_cast<float4*>(atile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(atile + kTcATileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (tid < kTcRows) {   // k = 128..135: features of the previous node (hi/lo), zero pad
                float f[4] = {0.f, 0.f, 0.f, 0.f};
                if (tid < RB) {
                    const float* rs = rowst + tid * kRowState;
                    f[0] = rs[9]; f[1] = rs[10];
                    if (TW) { f[2] = rs[11]; f[3] = rs[12]; }
                }
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { hi[j] = tf32_rna(f[j]); lo[j] = f[j] - hi[j]; }
                const int off = tc_a_off(tid, kH / 4);
                *reinterpret_cast<float4*>(atile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(atile + off + 128) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(atile + kTcATileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(atile + kTcATileBytes + off + 128) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tid == 0) *next_row = 0;
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
        // ================= phase A: gates on tcgen05 ================================================================
        if (warp == 0 && lane == 0) {
            uint32_t c = slabcnt;
            for (int u = 0; u < kTcUsesA; ++u, ++c) {
                const uint32_t st = c & 1;
                mbar_wait(bar_empty + 8 * st, ((c >> 1) & 1) ^ 1);
                mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                bulk_g2s(ring_u32 + st * kSlabBytes, a.WgTc + (size_t)u * kSlabFloats, kSlabBytes, bar0 + 8 * st);
            }
        } else if (warp == 1 && lane == 0) {
            uint32_t c = slabcnt;
            for (int u = 0; u < kTcUsesA; ++u, ++c) {
                const uint32_t st = c & 1;
                mbar_wait(bar0 + 8 * st, (c >> 1) & 1);
                tc_fence_after();
                const int kc = u >> 1, half = u & 1;
                const uint64_t da_hi = tc_desc(a_hi + kc * 256, 128, kTcASbo), da_lo = tc_desc(a_lo + kc * 256, 128, kTcASbo);
                const uint64_t db_hi = tc_desc(ring_u32 + st * kSlabBytes, 128, 256);
                const uint64_t db_lo = tc_desc(ring_u32 + st * kSlabBytes + 8192, 128, 256);
                const uint32_t d = tmem_base + half * 256;
                tc_mma_tf32(d, da_lo, db_hi, idescA, kc > 0);
                tc_mma_tf32(d, da_hi, db_lo, idescA, 1);
                tc_mma_tf32(d, da_hi, db_hi, idescA, 1);
                tc_commit(bar_empty + 8 * st);
            }
            tc_commit(bar_dready);
        }
        slabcnt += kTcUsesA;
        __syncwarp();
        mbar_wait(bar_dready, dcnt & 1);
        ++dcnt;
        tc_fence_after();
        VRP_CLK(0);
        float hsave[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // BasicLSTMCell epilogue out of TMEM (see rollout_kernel.cuh)
            float v[32];
            tc_ld32(tmem_lane + ub * 128 + g * 32, v);
            float w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float up = __shfl_sync(kFull, v[16 + k], lane & 15);
                w[k] = (lane < 16) ? v[k] : up;
            }
            const int unit0 = ub * 32 + g * 8 + (lane >> 4) * 4;
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bg) + unit0 + j);
                const float gi = w[4 * j + 0] + bb.x, gj = w[4 * j + 1] + bb.y, gf = w[4 * j + 2] + bb.z, go = w[4 * j + 3] + bb.w;
                const float cn = __fadd_rn(__fmul_rn(cst[g * 4 + j], sigmoid_fast(gf + a.forget_bias)),
                                           __fmul_rn(sigmoid_fast(gi), tanh_fast(gj)));
                cst[g * 4 + j] = cn;
                const float hn = __fmul_rn(tanh_fast(cn), sigmoid_fast(go));
                hsave[g * 4 + j] = hn;
                hi[j] = tf32_rna(hn);
                lo[j] = hn - hi[j];
            }
            if (tc_row_ok) {
                const int off = tc_a_off(tc_m, unit0 >> 2);
                *reinterpret_cast<float4*>(atile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(atile + kTcATileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();   // h' visible to the tensor core; gate accumulators drained
        tc_fence_after();
        tc_st16(tmem_lane + kHSaveCol + ub * 16, hsave);   // h' survives phase C in TMEM
        VRP_CLK(1);

        // ================= phase B: q on tcgen05 ====================================================================
        if (warp == 0 && lane == 0) {
            uint32_t c = slabcnt;
            for (int u = 0; u < kTcUsesB; ++u, ++c) {
                const uint32_t st = c & 1;
                mbar_wait(bar_empty + 8 * st, ((c >> 1) & 1) ^ 1);
                mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                bulk_g2s(ring_u32 + st * kSlabBytes, a.WqTc + (size_t)u * kSlabFloats, kSlabBytes, bar0 + 8 * st);
            }
        } else if (warp == 1 && lane == 0) {
            uint32_t c = slabcnt;
            for (int u = 0; u < kTcUsesB; ++u, ++c) {
                const uint32_t st = c & 1;
                mbar_wait(bar0 + 8 * st, (c >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int kc = 2 * u + kk;
                    const uint64_t da_hi = tc_desc(a_hi + kc * 256, 128, kTcASbo), da_lo = tc_desc(a_lo + kc * 256, 128, kTcASbo);
                    const uint64_t db_hi = tc_desc(ring_u32 + st * kSlabBytes + kk * 8192, 128, 256);
                    const uint64_t db_lo = tc_desc(ring_u32 + st * kSlabBytes + kk * 8192 + 4096, 128, 256);
                    tc_mma_tf32(tmem_base, da_lo, db_hi, idescB, kc > 0);
                    tc_mma_tf32(tmem_base, da_hi, db_lo, idescB, 1);
                    tc_mma_tf32(tmem_base, da_hi, db_hi, idescB, 1);
                }
                tc_commit(bar_empty + 8 * st);
            }
            tc_commit(bar_dready);
        }
        slabcnt += kTcUsesB;
        __syncwarp();
        mbar_wait(bar_dready, dcnt & 1);
        ++dcnt;
        tc_fence_after();
        VRP_CLK(2);
        {   // q + constant biases + load/time terms, pre-scaled, into the idle ring:  base'[row][h]
            float v[32];
            tc_ld32(tmem_lane + ub * 32, v);
            if (tc_row_ok && lane < 16) {
                const float* rs = rowst + tc_m * kRowState;
                const float load = rs[0], time = TW ? rs[1] : 0.0f;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.att.bq) + ub * 8 + q4);
                    const float4 gl = __ldg(reinterpret_cast<const float4*>(a.att.gl) + ub * 8 + q4);
                    const float4 gt = __ldg(reinterpret_cast<const float4*>(a.att.gt) + ub * 8 + q4);
                    float o[4];
                    const float qv[4] = {v[4 * q4] + bb.x, v[4 * q4 + 1] + bb.y, v[4 * q4 + 2] + bb.z, v[4 * q4 + 3] + bb.w};
                    const float glv[4] = {gl.x, gl.y, gl.z, gl.w}, gtv[4] = {gt.x, gt.y, gt.z, gt.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float s = fmaf(load, glv[j], qv[j]);
                        if (TW) s = fmaf(time, gtv[j], s);
                        o[j] = s * kTwoLog2e;
                    }
                    *reinterpret_cast<float4*>(qq + tc_m * kH + ub * 32 + q4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        // ================= C0: item list ============================================================================
        // (the A tiles are dead from here on: U becomes the phase-C scratch; ordered by the barrier below)
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
        __syncthreads();
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
        __syncthreads();
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
                    const float* fr = feat + ((size_t)lb * N + n) * FD;
                    if (TW) { const float4 f4 = *reinterpret_cast<const float4*>(fr); f[0] = f4.x; f[1] = f4.y; f[2] = f4.z; f[3] = f4.w; }
                    else { const float2 f2 = *reinterpret_cast<const float2*>(fr); f[0] = f2.x; f[1] = f2.y; }
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
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        float v[32];
                        tc_ld32(tmem_lane + 64 * g + 32 * c, v);
                        const int h0 = 64 * hh + 32 * c;
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(brow + h0 + 4 * k4);
                            const float4 v4 = *reinterpret_cast<const float4*>(vsm + h0 + 4 * k4);
                            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float rr = rcp_approx(1.0f + ex2_approx(v[4 * k4 + j] + bb[j]));
                                acc = fmaf(vv[j], fmaf(-2.0f, rr, 1.0f), acc);
                            }
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
        __syncthreads();
        tc_fence_after();
        VRP_CLK(4);

        // ================= C2: per row - mask, log-softmax, selection, Env.step =====================================
        // This part is a long dependent chain per row (reductions, exp/log, sqrt): 8 lanes per row, 4 rows per warp,
        // every row of the block in flight at once; lane sl owns nodes sl, sl+8, ... (16 per lane).
        int warp_done;
        {
            const int sub = lane >> 3, sl = lane & 7;
            const int lr = warp * 4 + sub;
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
            const float* frow = feat + (size_t)lbs * N * FD;
            float* drow = dem + lrs * NP;
            float* ulog = logits + lrs * NP;
            const int fl = flag_s[lrs];
            float lg[16];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = sl + 8 * j;
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
            for (int o = 4; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
            float se = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; ++j) se += (lg[j] == -INFINITY) ? 0.0f : expf(lg[j] - mx);
#pragma unroll
            for (int o = 4; o; o >>= 1) se += __shfl_xor_sync(kFull, se, o);
            const float lse = logf(se);
            float bestp = -1.0f;
            int besti = 0x7fffffff;
            const bool tracing = a.tr_logits || a.tr_probs || a.tr_masks;
            const long long tro = ((long long)t * R + row) * N;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = sl + 8 * j;
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
                for (int o = 4; o; o >>= 1) {
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
                    for (int draw = 0; draw < 2 && sel < 0; ++draw) {
                        const float* ubuf = draw ? a.uniforms_retry : a.uniforms;
                        const float u = ubuf ? ubuf[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                        float cum = 0.0f;   // tf.cumsum: sequential fp32 (masked nodes add exactly 0)
                        for (int n = 0; n < N; ++n) {
                            cum = __fadd_rn(cum, ulog[n]);
                            if (cum > u) { sel = n; break; }
                        }
                    }
                }
                if (sel < 0) sel = 0;
                idx = __shfl_sync(kFull, sel, sub * 8);
                __syncwarp();
                psel = ulog[idx];
                __syncwarp();
            }
            if (idx < 0 || idx >= N) idx = 0;   // rows that are not live carry garbage
            const float* fsel = frow + idx * FD;
            const float sx = fsel[0], sy = fsel[1];
            const float px = rs[2], py = rs[3];
            const float d_idx = drow[idx];
            __syncwarp();
            const float d_sat = fminf(d_idx, load);
            float nload = __fsub_rn(load, d_sat);
            const float ts = dist_rn(px, py, sx, sy);
            float ntime = 0.0f;
            if (TW) ntime = fmaxf(__fadd_rn(time, ts), fsel[2]);
            const float dep = (idx == n_cust) ? 1.0f : 0.0f;
            nload = __fadd_rn(__fmul_rn(nload, 1.0f - dep), __fmul_rn(dep, a.capacity));
            if (TW) ntime = __fmul_rn(ntime, 1.0f - dep);
            if (live && sl == 0) drow[idx] = __fsub_rn(d_idx, d_sat);
            __syncwarp();
            float sumdem = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = sl + 8 * j;
                lg[j] = (n < N) ? drow[n] : 0.0f;     // lg[] now holds the updated demands
                sumdem += lg[j];
            }
#pragma unroll
            for (int o = 4; o; o >>= 1) sumdem += __shfl_xor_sync(kFull, sumdem, o);
            if (live) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = sl + 8 * j;
                    if (n < N) {
                        float nx = 0.f, ny = 0.f, ne = 0.f;
                        if (TW) {
                            const float4 f4 = *reinterpret_cast<const float4*>(frow + n * 4);
                            nx = f4.x; ny = f4.y; ne = f4.w;
                        }
                        mk[n] = (unsigned char)node_mask<TW>(n, n_cust, lg[j], nload, ntime, sx, sy, nx, ny, ne, dep, sumdem);
                    }
                }
            }
            const int done_now = (!a.full && idx == n_cust && sumdem == 0.0f) ? 1 : 0;
            if (live && sl == 0) {
                rs[0] = nload; rs[1] = ntime; rs[2] = sx; rs[3] = sy;
                if (t == 0) { rs[5] = sx; rs[6] = sy; }
                else rs[4] = __fadd_rn(rs[4], ts);
                reinterpret_cast<int*>(rs)[8] = done_now;
                rs[9] = sx; rs[10] = sy; rs[11] = TW ? fsel[2] : 0.0f; rs[12] = TW ? fsel[FD - 1] : 0.0f;
                a.idx_out[(long long)t * R + row] = idx;
                if (a.logprob_out) a.logprob_out[(long long)t * R + row] = logf(psel);
            }
            warp_done = __all_sync(kFull, !live || done_now);
        }
        VRP_CLK(5);
        if (__syncthreads_and(warp_done)) { ++t; break; }
    }
    const int t_end = t;

    for (int i = 0; i < RPW; ++i) {
        const int lr = warp * RPW + i, wb = lr / IB, lb = lr - wb * IB;
        if (wb >= W || lb >= IBv) continue;
        const long long row = (long long)wb * a.B + b0 + lb;
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
