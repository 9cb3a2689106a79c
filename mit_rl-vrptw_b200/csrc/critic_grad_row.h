// Back-propagation of the critic loss of build_train_step for ONE instance:
//   critic_loss = tf.losses.mean_squared_error(R, v)  ->  d loss / d v[r] = 2 (v[r] - R[r]) / B
// (model/attention_agent.py:290, 298-299: gradients w.r.t. every variable under "Critic").  The graph differentiated is the
// critic of build_model('stochastic') (attention_agent.py:243-270; VRPTW/vrptw_attention.py:87-169, VRP/vrp_attention.py:79-140):
// hy_0 = 0; P process blocks (e_i, u_i) = AttentionCritic_i(hy_i, emb, env), hy_{i+1} = softmax(u_i) . e_i (no mask);
// v = L2(relu(L1(hy_P))).  The actor's embedding is a constant here (it is not a Critic variable); in VRPTW blocks
// emb_begin_tw / proj_end_tw serve BOTH windows (vrptw_attention.py:147-150), so their gradient sums both uses.
//
// Same construction as actor_grad_row.h: phases of "every thread i: body; barrier", reductions as serial loops in a fixed
// order, so the source also compiles as plain C++ for the CPU check of the algebra (tests/native/).
#pragma once
#include "actor_grad_row.h"

namespace vrpb200 {

constexpr int kCriticMaxBlocks = 8;

struct CriticBlockWeights {   // raw variables of Critic/Process/P<i>, TF layout
    const float *v, *emb_d_k, *emb_d_b, *proj_d_k, *proj_d_b, *proj_q_k, *proj_q_b, *proj_e_k, *proj_e_b;
    const float *emb_tw_k, *emb_tw_b, *proj_tw_k, *proj_tw_b;   // emb_begin_tw, proj_end_tw (VRPTW only)
};
struct CriticGradWeights {
    const float *emb_k, *emb_b;               // Actor/Embedding/conv1d (a constant of this graph)
    CriticBlockWeights blk[kCriticMaxBlocks];
    const float *l1_k, *l1_b, *l2_k, *l2_b;   // Critic/Linear/L1 [H, H], [H]; L2 [H, 1], [1]
};
struct CriticGradDims {
    int N, E, D, P, tw;
    float scale;                              // 2 / B
};
struct CriticGradRow {
    const float* input;   // [N, D]
    float R;              // the tour length the value is regressed on
    float *emb;           // scratch [N][E]
    float *ee, *eeT;      // scratch [P][N][H], [P][H][128]: e_i
    float *prob;          // scratch [P][128]
    // results of the row
    float *hyin;          // [P + 1][H]: the query input of every block, then the final read-out
    float *De;            // [P][N][H]  d loss / d e_i
    float *sums;          // [P][4][H]  S1 = sum darg, SD = sum demand darg, STW = sum (b_tw + e_tw) darg, Sv = d loss / d v_i
    float *dq;            // [P][H]
    float *l1, *dl1;      // [H]
    float *vdv;           // [2]: v, d loss / d v
};

AG_FN void critic_block_constants(const CriticBlockWeights& b, int tw, int i, const float* s_hy, float* s_base, float* s_gd, float* s_gt,
                                  float* s_v) {
    constexpr int H = kAgH;
    float g = 0.0f, c = b.proj_d_b[i];
    for (int k = 0; k < H; ++k) { g += b.emb_d_k[k] * b.proj_d_k[k * H + i]; c += b.emb_d_b[k] * b.proj_d_k[k * H + i]; }
    float gt = 0.0f, ct = 0.0f;
    if (tw) {
        ct = b.proj_tw_b[i];
        for (int k = 0; k < H; ++k) { gt += b.emb_tw_k[k] * b.proj_tw_k[k * H + i]; ct += b.emb_tw_b[k] * b.proj_tw_k[k * H + i]; }
    }
    float q = b.proj_q_b[i];
    for (int k = 0; k < H; ++k) q += s_hy[k] * b.proj_q_k[k * H + i];
    s_base[i] = q + c + 2.0f * ct;
    s_gd[i] = g;
    s_gt[i] = gt;
    s_v[i] = b.v[i];
}

AG_FN void critic_grad_row(const CriticGradWeights& w, const CriticGradDims& d, const CriticGradRow& r) {
    constexpr int H = kAgH;
    const int N = d.N, E = d.E, D = d.D, P = d.P, D1 = d.D - 1;
    AG_SHARED float s_hy[H], s_base[H], s_gd[H], s_gt[H], s_v[H], s_u[H], s_p[H], s_dp[H], s_du[H], s_dhy[H], s_dq[H], s_l1[H], s_dl1[H];
    AG_SHARED float s_red[2];

    AG_PAR(i)
        for (int e = i; e < N * E; e += H) {
            const int n = e / E, k = e - n * E;
            float s = w.emb_b[k];
            for (int c = 0; c < D1; ++c) s += r.input[n * D + c] * w.emb_k[c * E + k];
            r.emb[e] = s;
        }
        s_hy[i] = 0.0f;
        r.hyin[i] = 0.0f;
    AG_END

    // ---- forward ---------------------------------------------------------------------------------------------------------------
    for (int b = 0; b < P; ++b) {
        const CriticBlockWeights& bw = w.blk[b];
        float* ee = r.ee + (long long)b * N * H;
        float* eeT = r.eeT + (long long)b * H * 128;
        AG_PAR(i)
            critic_block_constants(bw, d.tw, i, s_hy, s_base, s_gd, s_gt, s_v);
            for (int n = 0; n < N; ++n) {
                float s = bw.proj_e_b[i];
                for (int k = 0; k < E; ++k) s += r.emb[n * E + k] * bw.proj_e_k[k * H + i];
                ee[n * H + i] = s;
                eeT[i * 128 + n] = s;
            }
        AG_END
        AG_PAR(n)
            if (n < N) {
                const float dn = r.input[n * D + D - 1], twn = d.tw ? r.input[n * D + 2] + r.input[n * D + 3] : 0.0f;
                float u = 0.0f;
                for (int h = 0; h < H; ++h) u += s_v[h] * tanhf(s_base[h] + eeT[h * 128 + n] + dn * s_gd[h] + twn * s_gt[h]);
                s_u[n] = u;
            }
        AG_END
        AG_PAR(i)
            if (i == 0) {
                float m = s_u[0];
                for (int n = 1; n < N; ++n) m = fmaxf(m, s_u[n]);
                float s = 0.0f;
                for (int n = 0; n < N; ++n) s += expf(s_u[n] - m);
                s_red[0] = m; s_red[1] = s;
            }
        AG_END
        AG_PAR(n)
            if (n < N) { const float p = expf(s_u[n] - s_red[0]) / s_red[1]; s_p[n] = p; r.prob[b * 128 + n] = p; }
        AG_END
        AG_PAR(i)
            float s = 0.0f;
            for (int n = 0; n < N; ++n) s += s_p[n] * ee[n * H + i];
            r.hyin[(b + 1) * H + i] = s;
        AG_END
        AG_PAR(i)
            s_hy[i] = r.hyin[(b + 1) * H + i];
        AG_END
    }
    // ---- v = L2(relu(L1(hy))) and its back-propagation into hy ------------------------------------------------------------------
    AG_PAR(i)
        float s = w.l1_b[i];
        for (int k = 0; k < H; ++k) s += s_hy[k] * w.l1_k[k * H + i];
        s = fmaxf(s, 0.0f);
        s_l1[i] = s;
        r.l1[i] = s;
    AG_END
    AG_PAR(i)
        if (i == 0) {
            float v = w.l2_b[0];
            for (int h = 0; h < H; ++h) v += s_l1[h] * w.l2_k[h];
            const float dv = d.scale * (v - r.R);
            r.vdv[0] = v; r.vdv[1] = dv;
            s_red[0] = dv;
        }
    AG_END
    AG_PAR(i)
        const float g = s_l1[i] > 0.0f ? s_red[0] * w.l2_k[i] : 0.0f;
        s_dl1[i] = g;
        r.dl1[i] = g;
    AG_END
    AG_PAR(i)
        float s = 0.0f;
        for (int h = 0; h < H; ++h) s += w.l1_k[i * H + h] * s_dl1[h];
        s_dhy[i] = s;
    AG_END

    // ---- backward through the process blocks ---------------------------------------------------------------------------------------
    for (int b = P - 1; b >= 0; --b) {
        const CriticBlockWeights& bw = w.blk[b];
        const float* ee = r.ee + (long long)b * N * H;
        const float* eeT = r.eeT + (long long)b * H * 128;
        AG_PAR(i)
            s_hy[i] = r.hyin[b * H + i];
            if (i < N) s_p[i] = r.prob[b * 128 + i];
        AG_END
        AG_PAR(i)
            critic_block_constants(bw, d.tw, i, s_hy, s_base, s_gd, s_gt, s_v);
        AG_END
        AG_PAR(n)   // hy' = sum_n p[n] e[n, :]:  d loss / d p[n]
            if (n < N) {
                float s = 0.0f;
                for (int h = 0; h < H; ++h) s += s_dhy[h] * eeT[h * 128 + n];
                s_dp[n] = s;
            }
        AG_END
        AG_PAR(i)
            if (i == 0) {
                float s = 0.0f;
                for (int n = 0; n < N; ++n) s += s_p[n] * s_dp[n];
                s_red[0] = s;
            }
        AG_END
        AG_PAR(n)   // through the soft-max
            if (n < N) s_du[n] = s_p[n] * (s_dp[n] - s_red[0]);
        AG_END
        AG_PAR(i)
            float S1 = 0.0f, SD = 0.0f, STW = 0.0f, Sv = 0.0f;
            float* De = r.De + (long long)b * N * H;
            for (int n = 0; n < N; ++n) {
                const float dn = r.input[n * D + D - 1], twn = d.tw ? r.input[n * D + 2] + r.input[n * D + 3] : 0.0f;
                const float th = tanhf(s_base[i] + ee[n * H + i] + dn * s_gd[i] + twn * s_gt[i]);
                const float du = s_du[n];
                Sv += du * th;
                const float da = du * s_v[i] * (1.0f - th * th);
                S1 += da; SD += dn * da; STW += twn * da;
                De[n * H + i] = s_p[n] * s_dhy[i] + da;
            }
            float* sm = r.sums + (long long)b * 4 * H;
            sm[i] = S1; sm[H + i] = SD; sm[2 * H + i] = STW; sm[3 * H + i] = Sv;
            r.dq[b * H + i] = S1;
            s_dq[i] = S1;
        AG_END
        AG_PAR(i)   // into the previous block's read-out (block 0's query input is the constant 0)
            float s = 0.0f;
            for (int h = 0; h < H; ++h) s += bw.proj_q_k[i * H + h] * s_dq[h];
            s_dhy[i] = s;
        AG_END
    }
}

// Gradients of one block's small variables from the per-row sums [R][P][4][H], one thread block.  The time-window chain is used
// twice per node (begin and end window): its constant part enters with 2 S1.
struct CriticSmallGradOut {
    float *v, *proj_q_b, *emb_d_k, *emb_d_b, *proj_d_k, *proj_d_b, *emb_tw_k, *emb_tw_b, *proj_tw_k, *proj_tw_b;   // any may be null
};

AG_FN void critic_small_grads(const CriticBlockWeights& bw, const float* sums, long long R, int P, int b, int tw, const CriticSmallGradOut& o) {
    constexpr int H = kAgH;
    AG_SHARED float tot[4 * H];
    AG_PAR(i)
        for (int a = 0; a < 4; ++a) {
            float s = 0.0f;
            for (long long r = 0; r < R; ++r) s += sums[((r * P + b) * 4 + a) * H + i];
            tot[a * H + i] = s;
        }
    AG_END
    AG_PAR(i)
        if (o.v) o.v[i] = tot[3 * H + i];
        if (o.proj_q_b) o.proj_q_b[i] = tot[i];
        if (o.proj_d_b) o.proj_d_b[i] = tot[i];
        float gk = 0.0f, gb = 0.0f;
        for (int j = 0; j < H; ++j) { gk += bw.proj_d_k[i * H + j] * tot[H + j]; gb += bw.proj_d_k[i * H + j] * tot[j]; }
        if (o.emb_d_k) o.emb_d_k[i] = gk;
        if (o.emb_d_b) o.emb_d_b[i] = gb;
        if (o.proj_d_k)
            for (int k = 0; k < H; ++k) o.proj_d_k[k * H + i] = bw.emb_d_k[k] * tot[H + i] + bw.emb_d_b[k] * tot[i];
        if (tw) {
            if (o.proj_tw_b) o.proj_tw_b[i] = 2.0f * tot[i];
            gk = 0.0f; gb = 0.0f;
            for (int j = 0; j < H; ++j) { gk += bw.proj_tw_k[i * H + j] * tot[2 * H + j]; gb += bw.proj_tw_k[i * H + j] * tot[j]; }
            if (o.emb_tw_k) o.emb_tw_k[i] = gk;
            if (o.emb_tw_b) o.emb_tw_b[i] = 2.0f * gb;
            if (o.proj_tw_k)
                for (int k = 0; k < H; ++k) o.proj_tw_k[k * H + i] = bw.emb_tw_k[k] * tot[2 * H + i] + 2.0f * bw.emb_tw_b[k] * tot[i];
        }
    AG_END
}

}  // namespace vrpb200
