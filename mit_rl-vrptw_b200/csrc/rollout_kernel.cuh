// The fused rollout kernel: RLAgent.build_model(decode_type) of the reference
// (model/attention_agent.py:84-240) for one block of instances per CTA, every decode step on-chip.
//
// One CTA owns RB rows (RB/W instances x W beams) for all T steps:
//   phase A  LSTM gates      [RB x 136] x [136 x 512]   (shared/decode_step.py:211-214; x-part folded to
//                                                         rank D-1 through the linear embedding)
//   phase B  query           [RB x 128] x [128 x 128]   (VRPTW/vrptw_attention.py:69)
//   phase C  per row, one warp: pointer scores u[n] = sum_h v_h tanh(.) over the un-masked nodes
//            (vrptw_attention.py:77, SURVEY App. C collapse), mask (decode_step.py:231-232),
//            log_softmax/exp (decode_step.py:125-126), selection (attention_agent.py:146-167),
//            Env.step (VRPTW/vrptw_utils.py:254-358), tour-length accumulation (vrptw_utils.py:397-414).
// Node features, demand, masks and the LSTM state stay in shared memory / registers for the whole
// rollout; HBM sees the instance once on the way in and T int32 indices + one cost on the way out.
// The weight matrices of phases A and B (387 KB fp32) do not fit next to the instances, so they are
// streamed from L2 every step through a 2-stage shared-memory ring with cp.async.bulk (TMA bulk copy,
// SASS UBLKCP) completing on mbarriers.
#pragma once
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
constexpr int kKG = 136;           // phase-A depth: 128 (h) + 4 (features of the previous node) + 4 (zero pad)
constexpr int kSlabFloats = 4096;  // 16 KB per ring stage
constexpr int kSlabBytes = kSlabFloats * 4;
constexpr int kStages = 2;
constexpr int kGateSlabs = kKG / 8;   // 8 k-rows x 512 columns per slab
constexpr int kQSlabs = 4;            // 32 k-rows x 128 columns per slab
constexpr int kRowState = 16;         // floats per row in shared memory
// tensor-core (tcgen05) variant of phases A/B: M=64 A tiles (hi + lo TF32 split), 16 KB B stages
constexpr int kTcRows = 64;                       // UMMA M
constexpr int kTcATileBytes = kTcRows * kKG * 4;  // one A tile (hi or lo): [64 rows][136 k] canonical K-major, no swizzle
constexpr int kTcASbo = (kKG / 4) * 128;          // byte stride between 8-row groups of an A tile
constexpr int kTcUsesA = 2 * (kKG / 8);           // (k-chunk, N-half) stage uses of phase A
constexpr int kTcUsesB = 8;                       // pairs of k-chunks of phase B
constexpr int kTcCols = 512;                      // TMEM columns (all of them: gates occupy 512 fp32 columns)
constexpr float kBig = 100000.0f;     // shared/decode_step.py:36
constexpr float kTwoLog2e = 2.8853900817779268f;  // tanh(x) = 1 - 2 / (1 + 2^(x * 2 log2 e))
constexpr unsigned kFull = 0xffffffffu;

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
};

struct RolloutArgs {
    const float* input;      // [B, N, D]
    const float* Wg;         // [17 slabs][8][2][2][32][4]  permuted LSTM kernel (+ folded embedding rows)
    const float* WgTc;       // tcgen05 path: [34 uses][hi 8 KB | lo 8 KB], tiles [N=256][K=8] canonical K-major
    const float* WqTc;       // tcgen05 path: [8 uses][2 k-chunks][hi 4 KB | lo 4 KB], tiles [N=128][K=8]
    const float* WgT3;       // rollout3: [32 uses][2 gate types][hi 4 KB | lo 4 KB], tiles [128 units][K=8] of W_h^T
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
};

struct SmemLayout {
    int ring, hT, feat, dem, mask, rs, scr, bar, total;   // byte offsets
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int RB, int IB, int N, int FD, bool tc) {
    SmemLayout L;
    const int NP = align_up(N, 4);
    const int nwarps = tc ? 16 : RB / 4;
    int off = 0;
    L.ring = off; off += kStages * kSlabBytes;            // also holds q[RB][128] during phase C
    L.hT = off;   off += tc ? 2 * kTcATileBytes : kKG * RB * 4;   // tc: A_hi tile, A_lo tile
    L.feat = off; off += align_up(IB * N * FD * 4, 16);
    L.dem = off;  off += RB * NP * 4;
    L.mask = off; off += align_up(RB * NP, 16);
    L.rs = off;   off += RB * kRowState * 4;
    L.scr = off;  off += nwarps * 640;                    // per warp: float u[128] + uint8 list[128]
    L.bar = off;  off += 64;                              // full[2], empty[2], dready, tmem base slot
    L.total = off;
    return L;
}

// ---- PTX helpers -------------------------------------------------------------------------------
VRP_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
VRP_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
VRP_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase completes instead of
// spinning through the loop (a spinning waiter takes issue slots from the warps that do the arithmetic)
VRP_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
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
// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
VRP_DEVINL uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
VRP_DEVINL void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
VRP_DEVINL float tf32_rna(float x) {   // round to TF32 (10 explicit mantissa bits), ties away
    uint32_t y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
    return __uint_as_float(y);
}
// byte offset of element (row m, k) in a canonical K-major no-swizzle A tile: [m/8][k/4][m%8][k%4]
VRP_DEVINL int tc_a_off(int m, int kq) { return (m >> 3) * kTcASbo + kq * 128 + (m & 7) * 16; }

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

// Env mask of one node after a step (VRPTW/vrptw_utils.py:328-348, VRP/vrp_utils.py:228-247), as a count 0..3.
template <bool TW>
VRP_DEVINL int node_mask(int n, int n_cust, float dem_n, float load, float time, float px, float py, float nx,
                         float ny, float e_tw, float dep, float sumdem) {
    if (n == n_cust) return (sumdem > 0.0f && dep != 0.0f) ? 1 : 0;
    int m = (dem_n == 0.0f) + (load == 0.0f);
    if (TW) m += (__fadd_rn(time, dist_rn(px, py, nx, ny)) > e_tw);
    return m;
}

// =================================================================================================
// TC = false: phases A/B on the FP32 pipe (FFMA, RB*8 threads).  TC = true: phases A/B on the tensor cores
// (tcgen05.mma kind::tf32, 3xTF32 error-compensated split, accumulators in TMEM; 512 threads).
template <int RB, bool TW, bool TC>
__global__ void __launch_bounds__(TC ? 512 : RB * 8, 1) rollout_kernel(const RolloutArgs a) {
    constexpr int NT = TC ? 512 : RB * 8;
    constexpr int NW = NT / 32;
    constexpr int RPW = RB / NW;          // rows per warp in phase C
    constexpr int FD = TW ? 4 : 2;
    static_assert(RB % NW == 0, "rows must divide over the warps");
    extern __shared__ __align__(128) unsigned char smem[];
    const int N = a.N, NP = align_up(N, 4), IB = a.IB, W = a.W, T = a.T, n_cust = a.n_cust;
    const SmemLayout L = make_layout(RB, IB, N, FD, TC);
    float* ring = reinterpret_cast<float*>(smem + L.ring);
    float* qq = ring;   // phase C view of the ring memory
    float* hT = reinterpret_cast<float*>(smem + L.hT);          // FFMA: h, k-major [136][RB]
    unsigned char* atile = smem + L.hT;                         // TC: A_hi tile then A_lo tile
    float* feat = reinterpret_cast<float*>(smem + L.feat);
    float* dem = reinterpret_cast<float*>(smem + L.dem);
    unsigned char* maskb = smem + L.mask;
    float* rowst = reinterpret_cast<float*>(smem + L.rs);
    unsigned char* scr = smem + L.scr;
    const uint32_t bar0 = smem_u32(smem + L.bar);               // full[0], full[1]
    const uint32_t bar_empty = bar0 + 16, bar_dready = bar0 + 32, tmem_slot = bar0 + 40;
    int* next_row = reinterpret_cast<int*>(smem + L.bar + 48);   // phase C work counter
    const uint32_t ring_u32 = smem_u32(ring);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp >> 1, uh = warp & 1;                    // FFMA ownership
    const int sp = warp & 3, ub = warp >> 2;                    // TC ownership: TMEM sub-partition, block of 32 units
    const long long b0 = (long long)blockIdx.x * IB;
    const int IBv = (int)min((long long)IB, a.B - b0);
    const long long R = a.B * W;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_dready, 1);
        fence_mbar_init();
    }
    if (TC) {
        __syncwarp();
        if (warp == 0) tc_alloc(tmem_slot, kTcCols);
        for (int i = tid; i < 2 * kTcATileBytes / 4; i += NT) reinterpret_cast<float*>(atile)[i] = 0.0f;
    } else {
        for (int i = tid; i < kKG * RB; i += NT) hT[i] = 0.0f;
    }
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
    if (TC) tc_fence_before();
    __syncthreads();
    if (TC) tc_fence_after();
    uint32_t tmem_base = 0;
    if (TC) tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L.bar + 40);

    float* ulog = reinterpret_cast<float*>(scr + warp * 640);
    unsigned char* list = scr + warp * 640 + 512;

    // decoder input of row lr for the next step = features of node `n` (rank D-1 form of emb[row, idx])
    auto write_next_input = [&](int lr, const float* f) {
        if (TC) {
            if (lane == 0) {
                float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < FD; ++k) { hi[k] = tf32_rna(f[k]); lo[k] = f[k] - hi[k]; }
                const int off = tc_a_off(lr, kH / 4);
                *reinterpret_cast<float4*>(atile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(atile + kTcATileBytes + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        } else {
            if (lane < FD) hT[(kH + lane) * RB + lr] = f[lane];
        }
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
    if (TC) fence_proxy_async();      // A tiles were written with generic stores; the tensor core reads them
    __syncthreads();

    // FFMA ownership: (rows 8rg..8rg+7) x (units u0, u0+1).  TC ownership: MMA row m = 16 sp + lane (lane < 16),
    // units 32 ub .. 32 ub + 31.  The LSTM cell state c lives in registers for all T steps either way.
    const int u0 = uh * 64 + 2 * lane;
    const float2 bq2 = TC ? make_float2(0.f, 0.f) : __ldg(reinterpret_cast<const float2*>(a.att.bq + u0));
    constexpr int NCST = 16;
    float cst[NCST];
#pragma unroll
    for (int r = 0; r < NCST; ++r) cst[r] = 0.0f;
    const int tc_m = 16 * sp + (lane & 15);                     // TC: this thread's MMA row; lane >> 4 picks the unit half
    const bool tc_row_ok = TC && tc_m < RB;

#ifdef VRPB200_PHASE_CLOCKS
    long long pc_[6] = {0, 0, 0, 0, 0, 0}, pc_last_ = clock64();
#endif
    uint32_t slabcnt = 0;     // ring uses so far (producer and consumer count identically)
    uint32_t dcnt = 0;        // completions of the "accumulator ready" barrier
    unsigned n_eval = 0, n_rowsteps = 0;   // work counters of this warp (roofline accounting)
    int t = 0;
    for (; t < T; ++t) {
        if (tid == 0) *next_row = 0;   // ordered before phase C by the barriers of phases A/B
        if constexpr (TC) {
            // ================= phase A on tcgen05: gates[64 x 512] = [h, feat] (hi+lo) . Wg (hi+lo) ==============
            constexpr uint32_t idescA = tc_idesc(kTcRows, 256), idescB = tc_idesc(kTcRows, 128);
            const uint32_t a_hi = smem_u32(atile), a_lo = a_hi + kTcATileBytes;
            if (warp == 0 && lane == 0) {          // TMA producer: 34 x 16 KB stages from L2
                fence_proxy_async();               // q[] of the previous step was written into the ring by generic stores
                uint32_t c = slabcnt;
                for (int u = 0; u < kTcUsesA; ++u, ++c) {
                    const uint32_t st = c & 1;
                    mbar_wait(bar_empty + 8 * st, ((c >> 1) & 1) ^ 1);
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.WgTc + (size_t)u * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
            } else if (warp == 1 && lane == 0) {   // MMA issuer
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
                    tc_mma_tf32(d, da_lo, db_hi, idescA, kc > 0);      // small terms first
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
            // BasicLSTMCell epilogue straight out of TMEM.  M=64 accumulators sit in lanes 0..15 of
            // each TMEM sub-partition: lane l < 16 receives 8 units x 4 gates of row 16 sp + l and hands units 4..7 to
            // lane l + 16, so that all 32 lanes run the cell update (4 units each).
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float v[32];
                tc_ld32(tmem_base + ((uint32_t)(32 * sp) << 16) + ub * 128 + g * 32, v);
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
            __syncthreads();   // h' complete and visible to the tensor core; TMEM drained
            VRP_CLK(1);

            // ================= phase B on tcgen05: q[64 x 128] = h' . W_q ===========================================
            if (warp == 0 && lane == 0) {
                uint32_t c = slabcnt;
                for (int u = 0; u < kTcUsesB; ++u, ++c) {
                    const uint32_t st = c & 1;
                    mbar_wait(bar_empty + 8 * st, ((c >> 1) & 1) ^ 1);
                    mbar_expect_tx(bar0 + 8 * st, kSlabBytes);
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.WqTc + (size_t)u * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
            } else if (warp == 1 && lane == 0) {
                tc_fence_after();
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
            {   // ring is idle from here to the next phase A: park q (+ every constant bias) in it
                float v[32];
                tc_ld32(tmem_base + ((uint32_t)(32 * sp) << 16) + ub * 32, v);
                if (tc_row_ok && lane < 16) {
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.att.bq) + ub * 8 + q4);
                        *reinterpret_cast<float4*>(qq + tc_m * kH + ub * 32 + q4 * 4) =
                            make_float4(v[4 * q4] + bb.x, v[4 * q4 + 1] + bb.y, v[4 * q4 + 2] + bb.z, v[4 * q4 + 3] + bb.w);
                    }
                }
            }
            tc_fence_before();
        } else {
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
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.att.Wq + (size_t)s * kSlabFloats, kSlabBytes, bar0 + 8 * st);
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
                    bulk_g2s(ring_u32 + st * kSlabBytes, a.att.Wq + (size_t)(s + kStages) * kSlabFloats, kSlabBytes, bar0 + 8 * st);
                }
                ++slabcnt;
            }
            // ring is idle from here to the next phase A: park q (+ every constant bias) in it
    #pragma unroll
            for (int r = 0; r < 8; ++r)
                *reinterpret_cast<float2*>(qq + (8 * rg + r) * kH + u0) = make_float2(qa[r][0] + bq2.x, qa[r][1] + bq2.y);
        }
        __syncthreads();
        VRP_CLK(3);

        // ================= phase C: one warp per row ====================================================
        // pointer-attention constants of this lane's 4 hidden units h = 4*lane .. 4*lane+3
        float cA[FD][4], cgd[4], cv[4], cgl[4], cgt[4];
#pragma unroll
        for (int k = 0; k < FD; ++k) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.att.A + k * kH) + lane);
            cA[k][0] = t4.x; cA[k][1] = t4.y; cA[k][2] = t4.z; cA[k][3] = t4.w;
        }
        {
            float4 t4 = __ldg(reinterpret_cast<const float4*>(a.att.gd) + lane);
            cgd[0] = t4.x; cgd[1] = t4.y; cgd[2] = t4.z; cgd[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(a.att.v) + lane);
            cv[0] = t4.x; cv[1] = t4.y; cv[2] = t4.z; cv[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(a.att.gl) + lane);
            cgl[0] = t4.x; cgl[1] = t4.y; cgl[2] = t4.z; cgl[3] = t4.w;
            t4 = __ldg(reinterpret_cast<const float4*>(a.att.gt) + lane);
            cgt[0] = t4.x; cgt[1] = t4.y; cgt[2] = t4.z; cgt[3] = t4.w;
        }
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
            const bool all_nodes = a.full || nun == 0 || !a.mask_pointer;
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

            // --- u[n] = sum_h v_h tanh(q_h + e[n,h] + d[n,h] + ld[n,h] + t_h), all pre-scaled by 2 log2 e
            float base[4];
            {
                const float4 q4 = *reinterpret_cast<const float4*>(qq + lr * kH + 4 * lane);
                const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = fmaf(load, cgl[j], qv[j]);
                    if (TW) s = fmaf(time, cgt[j], s);
                    base[j] = s * kTwoLog2e;
                }
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
                    for (int draw = 0; draw < 2 && sel < 0; ++draw) {
                        const float* ub = draw ? a.uniforms_retry : a.uniforms;
                        const float u = ub ? ub[o] : counter_uniform(a.seed, t, a.row_base + row, draw);
                        float cum = 0.0f;   // tf.cumsum: sequential fp32; skipped nodes have prob exactly 0
                        for (int k = 0; k < nact; ++k) {
                            const int n = list[k];
                            cum = __fadd_rn(cum, ulog[n]);
                            if (cum > u) { sel = n; break; }
                        }
                    }
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
            const float ts = dist_rn(px, py, sx, sy);
            float ntime = 0.0f;
            if (TW) ntime = fmaxf(__fadd_rn(time, ts), fsel[2]);
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
                        (unsigned char)node_mask<TW>(n, n_cust, dn[j], nload, ntime, sx, sy, nx, ny, ne, dep, sumdem);
                }
            }
            const int done_now = (!a.full && idx == n_cust && sumdem == 0.0f) ? 1 : 0;
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
        if (TC) fence_proxy_async();
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
        if (lane == 0) a.cost_out[row] = __fadd_rn(rs[4], dist_rn(rs[2], rs[3], rs[5], rs[6]));   // + |a_0 - a_{T-1}|
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
    if (TC) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tc_dealloc(tmem_base, kTcCols);
    }
}

}  // namespace vrpb200
