// Shared device-side vocabulary of libvrpb200: constants, the kernel argument block, PTX helpers (mbarrier, bulk copy,
// tcgen05 / TMEM, packed FP32x2), the separately-rounded Env arithmetic and the fast activations.  Every fused kernel
// (rollout_ffma.cuh, rollout3_kernel.cuh, rollout5_kernel.cuh) and the stand-alone kernels include this file.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrpb200 {

#define VRP_DEVINL __device__ __forceinline__
#ifdef VRPB200_PHASE_CLOCKS   // profiling build only (tools/phase_clocks.py): per-phase SM cycles into stats[4..9]
#define VRP_CLK(i) do { const long long c_ = clock64(); pc_[i] += c_ - pc_last_; pc_last_ = c_; } while (0)
#else
#define VRP_CLK(i) do { } while (0)
#endif

constexpr int kH = 128;            // hidden_dim of this build
constexpr int kKG = 136;           // FP32-pipe kernel, phase-A depth: 128 (h) + 4 (features of the previous node) + 4 (zero pad)
constexpr int kSlabFloats = 4096;  // 16 KB per ring stage
constexpr int kSlabBytes = kSlabFloats * 4;
constexpr int kStages = 2;
constexpr int kGateSlabs = kKG / 8;   // 8 k-rows x 512 columns per slab
constexpr int kQSlabs = 4;            // 32 k-rows x 128 columns per slab
constexpr int kRowState = 16;         // floats per row in shared memory
constexpr int kTcCols = 512;          // TMEM columns allocated per CTA (all of them)
constexpr float kBig = 100000.0f;     // shared/decode_step.py:36
constexpr float kTwoLog2e = 2.8853900817779268f;  // tanh(x) = 1 - 2 / (1 + 2^(x * 2 log2 e))
constexpr unsigned kFull = 0xffffffffu;
constexpr int kV3Threads = 544;      // rollout3: 16 compute warps + one streaming warp
constexpr int kV4Threads = 640;      // rollout4: 16 compute warps + weight feeder, LSTM/query MMA issuer, 2 score-MMA issuers

// Per attention module (glimpses first, pointer last), all device pointers into the weight blob.
struct AttWeights {
    const float* Wq;   // [4 slabs][32][128]   proj_q kernel, k-major
    const float* bq;   // [128]  b_q + c_e + c_d + c_ld + c_t
    const float* A;    // [4][128]  2log2e * (W_emb . W_ref)            feature -> e
    const float* gd;   // [128]     2log2e * (g_d - g_ld)               per unit of demand[n]
    const float* gl;   // [128]     g_ld                                per unit of load
    const float* gt;   // [128]     g_t                                 per unit of time (0 for VRP)
    const float* v;    // [128]
    const float* Au;   // [4][128]  W_emb . W_ref (unscaled, glimpse read-out)
    const float* ce;   // [128]     b_emb . W_ref + b_ref
    // glimpse modules only: the query of the NEXT module as a function of the glimpse read-out.  hy = sum_n p[n] e[n,:] =
    // pf . Au + ce with pf = sum_n p[n] feat[n], so  hy . W_q' + b'  =  pf . Mq + cq
    const float* Mq;   // [4][128]  Au . W_q(next module)
    const float* cq;   // [128]     ce . W_q(next module) + bq(next module)   (bq = every constant of that module)
};

struct RolloutArgs {
    const float* input;      // [B, N, D]
    const float* Wg;         // [17 slabs][8][2][2][32][4]  permuted LSTM kernel (+ folded embedding rows)
    const float* WqTc;       // [8 uses][2 k-chunks][hi 4 KB | lo 4 KB], tiles [128 units][K=8] of W_q^T (TF32 hi/lo split)
    const float* WgT3;       // rollout3: [32 uses][2 gate types][hi 4 KB | lo 4 KB], tiles [128 units][K=8] of W_h^T
    const float* WqTh;       // rollout4/5: [4 uses][2 k-chunks][hi 4 KB | lo 4 KB], tiles [128 units][K=16] of W_q^T as FP16 hi + FP16 lo
    const float* WgTh;       // rollout4/5: [16 uses][2 gate types][hi 4 KB | lo 4 KB], tiles [128 units][K=16] of W_h^T, FP16 hi/lo
    const float* WxT;        // rollout3: [4 features][4 gate types][128 units]  W_emb . W_lstm[:E]
    const float* bgT;        // rollout3: [4 gate types][128 units]  b_lstm + b_emb . W_lstm[:E]
    const float* bg;         // [128][4]  gate bias per unit: i, j, f, o (b_lstm + b_emb . W_x)
    AttWeights att;          // pointer attention
    const float* uniforms;   // [T, R] or null
    const float* uniforms_retry;
    int32_t* idx_out;        // [T, R]
    float* cost_out;         // [R]
    float* logprob_out;      // [T, R] or null
    int32_t* parent_out;     // [T, R] or null (beam search)
    float* tr_logits;        // traces, any may be null
    float* tr_probs;
    float* tr_masks;
    float* fin_demand;
    float* fin_load;
    float* fin_time;
    float* fin_mask;
    int32_t* steps_out;      // [grid] or null
    unsigned long long* stats; // [4] or null: node evaluations, live row-steps, block-steps*RB, blocks
    unsigned long long seed;
    long long B;             // instances
    long long row_base;      // global index of instance 0 of this call (chunked host entry point), RNG only
    int W;                   // beams per instance (1 unless beam search)
    int IB;                  // instances per CTA
    int N, n_cust, D, T;
    int mode;                // 0 greedy, 1 stochastic, 2 beam search (rollout3 only)
    int full;                // evaluate every node every step, never exit early (trace mode)
    int use_tanh, mask_pointer;
    float capacity, tanh_C, forget_bias;
    const float4* feat4;     // rollout4: [B, N] (x, y, b_tw, e_tw) re-packed copy of the node features (repack_features_kernel)
    float sumv;              // rollout4: sum_h v_h of the pointer attention
    float vm2[128];          // rollout4: -2 v_h (read as constant-bank operands)
    // FP32-pipe kernel only (the tensor-core kernels are not launched with these set):
    const AttWeights* gatt;  // [n_glimpses] glimpse modules (device array), shared/decode_step.py:218-227
    const float* q_Wq;       // W_q / bias of the module that takes the LSTM output as its query (glimpse 0, else the pointer)
    const float* q_bq;
    int n_glimpses, mask_glimpses;
    int physics;             // 0: VRP / VRPTW / UPS Env (Euclidean travel time); 1: UPS_old/vrptw_ups_utils.py:321-361
    // whole-batch sampling retry of the reference (attention_agent.py:163-167), opt-in: retry_steps[t] != 0 = every row draws
    // step t from uniforms_retry; fail_steps[t] is set when a row's FIRST draw of a step that was not redrawn finds no node
    const int* retry_steps;  // [T] or null
    int* fail_steps;         // [T] or null (null: per-row retry, the default)
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

// ---- PTX helpers -------------------------------------------------------------------------------
VRP_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
VRP_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
VRP_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef VRP_SUSPEND_NS
#define VRP_SUSPEND_NS 0x989680u
#endif
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase completes instead of
// spinning through the loop (a spinning waiter takes issue slots from the warps that do the arithmetic)
VRP_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity), "r"(VRP_SUSPEND_NS) : "memory");
}
VRP_DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
VRP_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
VRP_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VRP_DEVINL float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
VRP_DEVINL float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- tcgen05 / TMEM (sm_100a) --------------------------------------------------------------------
VRP_DEVINL void tc_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VRP_DEVINL void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
VRP_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
VRP_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
VRP_DEVINL void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::tf32, issued by one thread
VRP_DEVINL void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
// one lane of a fully converged warp (elect.sync): tcgen05.mma / tcgen05.commit issued under this predicate from warp-uniform
// control flow compile to a single predicated UTCHMMA; under `if (lane == 0)` the compiler wraps every one of them into an
// ELECT ... BRA.U.ANY loop over the active threads (6 serial instructions per MMA on the issuing thread)
VRP_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
VRP_DEVINL void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
VRP_DEVINL uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// instruction descriptor for kind::f16: F32 accumulate, FP16 x FP16, both K-major, M x N
__host__ __device__ constexpr uint32_t tc_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
VRP_DEVINL float tf32_rna(float x) {   // round to TF32 (10 explicit mantissa bits), ties away
    uint32_t y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
    return __uint_as_float(y);
}

// LSTM activations (5 per unit per row-step: accuracy first, they feed every later step)
VRP_DEVINL float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }
VRP_DEVINL float tanh_acc(float x) {
    // 1 - 2/(1+e^{2x}); exact limits at +-inf, abs error ~1e-7
    return 1.0f - 2.0f * __fdividef(1.0f, 1.0f + expf(2.0f * x));
}

// MUFU forms: sigmoid(x) = 1/(1 + 2^(-x log2 e)), tanh(x) = 1 - 2/(1 + 2^(2x log2 e)); abs error ~1e-7
VRP_DEVINL float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
VRP_DEVINL float tanh_fast(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(kTwoLog2e * x)), 1.0f); }

// ---- Env arithmetic: every op separately rounded, exactly as the TF graph executes it ------------
// (VRPTW/vrptw_utils.py:304-309 and :340-344: sqrt(square(dx) + square(dy)))
VRP_DEVINL float dist_rn(float ax, float ay, float bx, float by) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// UPS_old/vrptw_ups_utils.py:321-322, 359-360, 421-423: x = latitude, y = longitude (radians);
// d = 6371 * sqrt(square((by - ay) * cos(0.5 * (bx + ax))) + square(bx - ax)) km (the same value whichever point is "a":
// the differences are squared and the sum is commutative); travel time = (100/13 * 1.2) * d clicks, 1.5 clicks per package.
constexpr float kGeoRadius = 6371.0f;
constexpr float kGeoClicksPerKm = (float)((100.0 / 13.0) * 1.2);
constexpr float kGeoService = 1.5f;
VRP_DEVINL float geo_km_rn(float ax, float ay, float bx, float by) {
    const float interm = __fmul_rn(__fsub_rn(by, ay), cosf(__fmul_rn(0.5f, __fadd_rn(bx, ax))));
    const float dx = __fsub_rn(bx, ax);
    return __fmul_rn(kGeoRadius, __fsqrt_rn(__fadd_rn(__fmul_rn(interm, interm), __fmul_rn(dx, dx))));
}
// length of a move (what reward_func adds up) and its travel time, by Env physics
VRP_DEVINL float move_len(int physics, float ax, float ay, float bx, float by) {
    return physics ? geo_km_rn(ax, ay, bx, by) : dist_rn(ax, ay, bx, by);
}
VRP_DEVINL float travel_time(int physics, float len) { return physics ? __fmul_rn(kGeoClicksPerKm, len) : len; }

VRP_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
VRP_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Sum 8 per-lane partials across the warp with 9 shuffles: on return lane L holds the total of a[(L>>2)&7].
VRP_DEVINL float reduce8(const float (&a)[8], int lane) {
    float x[4], y[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b4 ? a[i + 4] : a[i], send = b4 ? a[i] : a[i + 4];
        x[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b3 ? x[i + 2] : x[i], send = b3 ? x[i] : x[i + 2];
        y[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    const float keep = b2 ? y[1] : y[0], send = b2 ? y[0] : y[1];
    float z = keep + __shfl_xor_sync(kFull, send, 4);
    z += __shfl_xor_sync(kFull, z, 2);
    z += __shfl_xor_sync(kFull, z, 1);
    return z;
}

// Counter-based uniform in [0,1) for (seed, step, row, draw): splitmix64 finaliser, top 24 bits.
VRP_DEVINL float counter_uniform(unsigned long long seed, unsigned t, unsigned long long row, unsigned draw) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (row * 2ull + draw + 1ull) + 0xD1B54A32D192ED03ull * (t + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
}

// which of the two uniform draws a row may use at step t (see RolloutArgs::retry_steps)
VRP_DEVINL void draw_range(const RolloutArgs& a, int t, int& d0, int& d1) {
    if (a.fail_steps) { d0 = a.retry_steps[t] ? 1 : 0; d1 = d0 + 1; }
    else { d0 = 0; d1 = 2; }
}

// Env mask of one node after a step (VRPTW/vrptw_utils.py:328-348, VRP/vrp_utils.py:228-247), as a count 0..3.
template <bool TW>
VRP_DEVINL int node_mask(int n, int n_cust, float dem_n, float load, float time, float px, float py, float nx,
                         float ny, float e_tw, float dep, float sumdem, int physics = 0) {
    if (n == n_cust) return (sumdem > 0.0f && dep != 0.0f) ? 1 : 0;
    int m = (dem_n == 0.0f) + (load == 0.0f);
    if (TW) m += (__fadd_rn(time, travel_time(physics, move_len(physics, px, py, nx, ny))) > e_tw);
    return m;
}


// ---- tiles of the pointer-score GEMM and of the transposed LSTM / query GEMMs -------------------------------------
constexpr int kC1TileBytes = 128 * 8 * 4;   // one F tile (hi or lo): 128 items x K=8, canonical K-major
constexpr int kV3UsesQ = 8;                 // W_q^T: 2 k-chunks per 16 KB ring stage
#ifdef VRP_EXP_NOGATES                      // timing experiment only: drop the gates GEMM (results are wrong)
constexpr int kV3UsesG = 0;
#else
constexpr int kV3UsesG = 32;                // W_h^T: (k-chunk, pair of gate types) per ring stage
#endif
// rollout4/5: the LSTM / query GEMMs on kind::f16 with a two-term FP16 split of both operands, x = hi + lo, hi = fp16(x), lo =
// fp16(x - hi): |x| <= 1 (h') or <= 0.2 (Glorot weights), so lo lies in FP16's subnormal range where the spacing is 2^-24 -
// the residual x - hi - lo is <= 2^-25 ABSOLUTE, and hi.hi + lo.hi + hi.lo is as accurate as an FP32 matmul (measured on
// Glorot weights: rms error 7.7e-8 against 7.4e-8 for numpy's fp32 matmul and 2.8e-8 for 3xTF32).  K = 16 per MMA instead of 8:
// half the MMAs (the issue rate of the single issuing thread, ~65 clk per MMA, is what the query GEMM waits for), half the
// weight bytes streamed from L2 per step, half the shared memory of the h' operand tiles.
constexpr int kHUsesQ = 4;                  // W_q^T: 2 k-chunks (K = 16 each) per 16 KB ring stage
constexpr int kHUsesG = 16;                 // W_h^T: (k-chunk, pair of gate types) per ring stage
// h' as the [row][k] K-major B operand in FP16, [row/8][k/8][row%8][k%8]: descriptor of k-chunk kc (K = 16)
#define VRP_XDESC_H(base, kc) tc_desc((base) + (kc) * 256, 128, 2048)
// h as the [row][k] K-major B operand of the transposed GEMMs, [row/8][k/4][row%8][k%4]: descriptor of k-chunk kc
#define VRP_XDESC(base, kc) tc_desc((base) + (kc) * 256, 128, 4096)

VRP_DEVINL void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
VRP_DEVINL void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
VRP_DEVINL void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
VRP_DEVINL void tc_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
VRP_DEVINL void tc_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// TMEM load split into issue + wait so that the load can overlap arithmetic; the wait carries the destination
// registers as in/out operands, which keeps the compiler from consuming them early
VRP_DEVINL void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
VRP_DEVINL void tc_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// 4x4 transpose across the 4 lanes of a quad: in: a[i] = M[lane&3][i]; out: a[i] = M[i][lane&3]
VRP_DEVINL void transpose4(float (&a)[4], int lane) {
    const bool b0 = lane & 1, b1 = lane & 2;
    {
        const float s0 = b0 ? a[0] : a[1], s1 = b0 ? a[2] : a[3];
        const float r0 = __shfl_xor_sync(kFull, s0, 1), r1 = __shfl_xor_sync(kFull, s1, 1);
        if (b0) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
    }
    {
        const float s0 = b1 ? a[0] : a[2], s1 = b1 ? a[1] : a[3];
        const float r0 = __shfl_xor_sync(kFull, s0, 2), r1 = __shfl_xor_sync(kFull, s1, 2);
        if (b1) { a[0] = r0; a[1] = r1; } else { a[2] = r0; a[3] = r1; }
    }
}

// sum_n exp(logit[n] - max) of one row held by LPR lanes (lane sl owns nodes sl + LPR j), added in ONE fixed order
// whatever LPR is - the order a full warp would use (node n -> lane n % 32, slots n / 32 in sequence, then the xor
// butterfly 16, 8, 4, 2, 1) - so that the choice of rows per CTA can never move a near-tie.
template <int LPR>
VRP_DEVINL float softmax_denominator(const float (&lg)[128 / LPR], float mx) {
    constexpr int NQ = 32 / LPR;
    float P[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float u = lg[q + NQ * k];
            acc = __fadd_rn(acc, (u == -INFINITY) ? 0.0f : expf(u - mx));
        }
        P[q] = acc;
    }
    float se;
    if (NQ == 4) se = __fadd_rn(__fadd_rn(P[0], P[2 % NQ]), __fadd_rn(P[1 % NQ], P[3 % NQ]));
    else if (NQ == 2) se = __fadd_rn(P[0], P[1 % NQ]);
    else se = P[0];
#pragma unroll
    for (int o = LPR / 2; o; o >>= 1) se = __fadd_rn(se, __shfl_xor_sync(kFull, se, o));
    return se;
}

VRP_DEVINL void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
VRP_DEVINL void tc_ld_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
VRP_DEVINL void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
VRP_DEVINL float ldcg_volatile(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// ---- packed FP32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 - two results per issue slot) -------------------------
typedef unsigned long long f32x2;
VRP_DEVINL f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
VRP_DEVINL void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
VRP_DEVINL f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
VRP_DEVINL f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
VRP_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// named barrier that also AND-reduces a predicate over its participants (one barrier instead of flags + two barriers)
VRP_DEVINL int named_bar_and(int id, int nthreads, int pred) {
    int r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %3, 0;\n\tbar.red.and.pred p, %1, %2, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"(id), "r"(nthreads), "r"(pred) : "memory");
    return r;
}
VRP_DEVINL float4 ldcg4_volatile(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
VRP_DEVINL int lds_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

}  // namespace vrpb200
