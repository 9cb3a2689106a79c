// Back-propagation of the REINFORCE actor loss through the teacher-forced rollout of ONE row:
//   L = sum_r weight[r] * sum_t log prob_t[r, idx[t, r]],   weight[r] = (R[r] - v[r]) / B
// (model/attention_agent.py:286-300: actor_loss = reduce_mean((R - stop_gradient(v)) * add_n(logprobs)), gradients w.r.t.
// every variable under "Actor").  The graph differentiated is RNNDecodeStep.step without glimpses (shared/decode_step.py:
// 97-128, 182-234): BasicLSTMCell -> Attention*Actor (VRPTW/vrptw_attention.py:30-84, VRP/vrp_attention.py:27-76) -> mask ->
// log_softmax, with the chosen node's embedding as the next LSTM input (attention_agent.py:208-214).  Env state (load, time,
// demand, mask) is a constant of the graph (the reference's Env ops carry no gradient to the variables).
//
// One block of 128 threads per row: thread = hidden unit, or node, depending on the phase.  The source is written as a
// sequence of phases, each "for every thread i: body; barrier", with everything that lives across phases in block-shared
// arrays and every reduction a serial loop in a fixed order.  That makes the results independent of the schedule, and lets
// the SAME source compile as plain C++ (a phase becomes a loop over i) - tests/ builds it with g++ to check the algebra
// against torch autograd on a machine without a GPU.  The product (libvrpb200.so) contains the CUDA build only.
//
// Per row the block writes, for the device-wide reductions that follow (train_kernels.cuh):
//   dgates [T][4H]  d loss / d LSTM gate pre-activations      dq [T][H]   d loss / d (proj_q output)
//   Dsum [N][H]     sum_t d loss / d arg[t, n, :]             demb [N][E] d loss / d embedding of node n (both uses)
//   sums [5][H]     S1 = sum darg, SD = sum demand darg, SLD = sum (load - demand) darg, ST = sum time darg, Sv = d loss / d v
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AG_FN __device__ __forceinline__
#define AG_SHARED __shared__
#define AG_PAR(i) { const int i = threadIdx.x;
#define AG_END } __syncthreads();
#else
#define AG_FN static inline
#define AG_SHARED
#define AG_PAR(i) for (int i = 0; i < 128; ++i) {
#define AG_END }
#endif

namespace vrpb200 {

constexpr int kAgH = 128;          // hidden_dim of this build = threads per block
constexpr float kAgBig = 100000.0f;  // BIGNUMBER, shared/decode_step.py:36

struct ActorGradWeights {   // raw variables, TF layout (device pointers in the CUDA build)
    const float *emb_k, *emb_b;     // Actor/Embedding/conv1d: [D1, E], [E]
    const float *lstm_k, *lstm_b;   // basic_lstm_cell: [E + H, 4H] (gate order i, j, f, o), [4H]
    const float *v;                 // Actor/Decoder/Attention/v [H]
    const float *emb_d_k, *emb_d_b, *proj_d_k, *proj_d_b;       // [H], [H], [H, H], [H]
    const float *emb_ld_k, *emb_ld_b, *proj_ld_k, *proj_ld_b;
    const float *emb_t_k, *emb_t_b, *proj_t_k, *proj_t_b;       // VRPTW only
    const float *proj_q_k, *proj_q_b;                           // [H, H], [H]
    const float *proj_ref_k, *proj_ref_b;                       // [E, H], [H]
    const float *lstm_kT, *proj_q_kT;   // transposed copies [4H, E + H], [H(out), H(in)]: the backward products read along a thread's
                                        // own input index, which is the contiguous one here (coalesced over the threads)
};

struct ActorGradDims {
    int N, E, D, T, tw, use_tanh, mask_pointer;
    float C, forget_bias;
};

struct ActorGradRow {
    const float* input;        // [N, D] the row's instance
    const int* idx;            // chosen node of step t at idx[t * idx_stride]
    long long idx_stride;
    float weight;
    const float* xscale;       // LSTM-input dropout (shared/decode_step.py:177-179): the step's input is emb[node] * xscale[t * xscale_stride + k],
    long long xscale_stride;   // xscale = keep-mask / keep-probability; null = no dropout
    // Env state SEEN by step t (before the step's decision): demand / mask at [t * node_stride + n], load / time at [t * scalar_stride]
    const float *demand, *mask, *load, *time;
    long long node_stride, scalar_stride;
    // scratch of the row
    float *hs, *cs;            // [T + 1][H]
    float *act;                // [T][4H]: sigmoid(i), tanh(j), sigmoid(f + forget_bias), sigmoid(o)
    float *xin;                // [T][E] LSTM inputs
    float *dx;                 // [T][E] d loss / d LSTM inputs
    float *emb;                // [N][E]
    float *eref, *erefT;       // [N][H], [H][128]
    // results of the row
    float *dgates, *dq, *Dsum, *demb, *sums;
};

AG_FN float ag_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
AG_FN int ag_node(int idx, int N) { return idx < 0 ? 0 : (idx >= N ? N - 1 : idx); }   // a caller's bad index must not read out of bounds

AG_FN void actor_grad_row(const ActorGradWeights& w, const ActorGradDims& d, const ActorGradRow& r) {
    constexpr int H = kAgH;
    const int N = d.N, E = d.E, D = d.D, T = d.T, D1 = d.D - 1;
    AG_SHARED float s_gd[H], s_cd[H], s_gld[H], s_cld[H], s_gt[H], s_ct[H], s_v[H];
    AG_SHARED float s_base[H], s_u[H], s_lg[H], s_du[H], s_dq[H], s_dh[H], s_dc[H], s_x[H], s_hprev[H], s_dg[4 * H];
    AG_SHARED float s_acc[5 * H], s_red[2];

    // ---- embedding of the row's nodes; collapsed scalar -> H chains: chain(s)[h] = s * g[h] + c[h] --------------------------
    AG_PAR(i)
        for (int e = i; e < N * E; e += H) {
            const int n = e / E, k = e - n * E;
            float s = w.emb_b[k];
            for (int c = 0; c < D1; ++c) s += r.input[n * D + c] * w.emb_k[c * E + k];
            r.emb[e] = s;
        }
        float g = 0.0f, c = w.proj_d_b[i];
        for (int k = 0; k < H; ++k) { g += w.emb_d_k[k] * w.proj_d_k[k * H + i]; c += w.emb_d_b[k] * w.proj_d_k[k * H + i]; }
        s_gd[i] = g; s_cd[i] = c;
        g = 0.0f; c = w.proj_ld_b[i];
        for (int k = 0; k < H; ++k) { g += w.emb_ld_k[k] * w.proj_ld_k[k * H + i]; c += w.emb_ld_b[k] * w.proj_ld_k[k * H + i]; }
        s_gld[i] = g; s_cld[i] = c;
        g = 0.0f; c = 0.0f;
        if (d.tw) {
            c = w.proj_t_b[i];
            for (int k = 0; k < H; ++k) { g += w.emb_t_k[k] * w.proj_t_k[k * H + i]; c += w.emb_t_b[k] * w.proj_t_k[k * H + i]; }
        }
        s_gt[i] = g; s_ct[i] = c;
        s_v[i] = w.v[i];
        s_dh[i] = 0.0f; s_dc[i] = 0.0f;
        for (int a = 0; a < 5; ++a) s_acc[a * H + i] = 0.0f;
        r.hs[i] = 0.0f; r.cs[i] = 0.0f;
    AG_END
    AG_PAR(i)   // e[n, :] = emb[n, :] . proj_ref + bias, both layouts
        for (int n = 0; n < N; ++n) {
            float s = w.proj_ref_b[i];
            for (int k = 0; k < E; ++k) s += r.emb[n * E + k] * w.proj_ref_k[k * H + i];
            r.eref[n * H + i] = s;
            r.erefT[i * 128 + n] = s;
            r.Dsum[n * H + i] = 0.0f;
        }
    AG_END

    // ---- forward: the LSTM along the given tour (the attention does not feed back into the state) ---------------------------
    for (int t = 0; t < T; ++t) {
        const int node = t == 0 ? N - 1 : ag_node(r.idx[(long long)(t - 1) * r.idx_stride], N);   // depot first (attention_agent.py:133-137)
        AG_PAR(i)
            if (i < E) {
                float x = r.emb[node * E + i];
                if (r.xscale) x *= r.xscale[(long long)t * r.xscale_stride + i];
                s_x[i] = x; r.xin[t * E + i] = x;
            }
            s_hprev[i] = r.hs[t * H + i];
        AG_END
        AG_PAR(i)
            float g4[4] = {w.lstm_b[i], w.lstm_b[H + i], w.lstm_b[2 * H + i], w.lstm_b[3 * H + i]};   // the four gates side by side:
            for (int k = 0; k < E; ++k) {                                                               // four independent loads per k
                const float* kr = w.lstm_k + (long long)k * 4 * H + i;
                const float x = s_x[k];
                g4[0] += x * kr[0]; g4[1] += x * kr[H]; g4[2] += x * kr[2 * H]; g4[3] += x * kr[3 * H];
            }
            for (int k = 0; k < H; ++k) {
                const float* kr = w.lstm_k + (long long)(E + k) * 4 * H + i;
                const float x = s_hprev[k];
                g4[0] += x * kr[0]; g4[1] += x * kr[H]; g4[2] += x * kr[2 * H]; g4[3] += x * kr[3 * H];
            }
            const float si = ag_sigmoid(g4[0]), tj = tanhf(g4[1]), sf = ag_sigmoid(g4[2] + d.forget_bias), so = ag_sigmoid(g4[3]);
            const float cn = r.cs[t * H + i] * sf + si * tj;
            r.cs[(t + 1) * H + i] = cn;
            r.hs[(t + 1) * H + i] = tanhf(cn) * so;
            r.act[t * 4 * H + i] = si; r.act[t * 4 * H + H + i] = tj; r.act[t * 4 * H + 2 * H + i] = sf; r.act[t * 4 * H + 3 * H + i] = so;
        AG_END
    }

    // ---- backward sweep ------------------------------------------------------------------------------------------------------
    for (int t = T - 1; t >= 0; --t) {
        const float* dem = r.demand + (long long)t * r.node_stride;
        const float* msk = r.mask + (long long)t * r.node_stride;
        const float load = r.load[(long long)t * r.scalar_stride];
        const float time = d.tw ? r.time[(long long)t * r.scalar_stride] : 0.0f;
        const int chosen = ag_node(r.idx[(long long)t * r.idx_stride], N);
        AG_PAR(i)
            s_hprev[i] = r.hs[(t + 1) * H + i];        // the query of step t is the LSTM output of step t
        AG_END
        AG_PAR(i)
            float q = w.proj_q_b[i];
            for (int k = 0; k < H; ++k) q += s_hprev[k] * w.proj_q_k[k * H + i];
            s_base[i] = q + s_cd[i] + s_cld[i] + (d.tw ? time * s_gt[i] + s_ct[i] : 0.0f);
        AG_END
        AG_PAR(n)   // logits: thread = node
            if (n < N) {
                const float dn = dem[n], l = load - dn;
                float u = 0.0f;
                for (int h = 0; h < H; ++h) u += s_v[h] * tanhf(s_base[h] + r.erefT[h * 128 + n] + dn * s_gd[h] + l * s_gld[h]);
                s_u[n] = u;
                float lg = d.use_tanh ? d.C * tanhf(u) : u;
                if (d.mask_pointer) lg -= kAgBig * msk[n];
                s_lg[n] = lg;
            }
        AG_END
        AG_PAR(i)   // log_softmax statistics in one fixed order
            if (i == 0) {
                float m = s_lg[0];
                for (int n = 1; n < N; ++n) m = fmaxf(m, s_lg[n]);
                float s = 0.0f;
                for (int n = 0; n < N; ++n) s += expf(s_lg[n] - m);
                s_red[0] = m; s_red[1] = logf(s);
            }
        AG_END
        AG_PAR(n)   // d loss / d u[n] = weight (1[n = chosen] - prob[n]) (x C (1 - tanh^2 u) with use_tanh)
            if (n < N) {
                const float p = expf((s_lg[n] - s_red[0]) - s_red[1]);
                float du = r.weight * ((n == chosen ? 1.0f : 0.0f) - p);
                if (d.use_tanh) { const float tu = tanhf(s_u[n]); du *= d.C * (1.0f - tu * tu); }
                s_du[n] = du;
            }
        AG_END
        AG_PAR(i)   // thread = hidden unit: d arg[n, i] and the sums every variable's gradient is made of
            float S1 = 0.0f, SD = 0.0f, SLD = 0.0f, Sv = 0.0f;
            for (int n = 0; n < N; ++n) {
                const float du = s_du[n];
                if (du != 0.0f) {   // masked nodes have probability exactly 0
                    const float dn = dem[n], l = load - dn;
                    const float th = tanhf(s_base[i] + r.eref[n * H + i] + dn * s_gd[i] + l * s_gld[i]);
                    Sv += du * th;
                    const float da = du * s_v[i] * (1.0f - th * th);
                    S1 += da; SD += dn * da; SLD += l * da;
                    r.Dsum[n * H + i] += da;
                }
            }
            r.dq[t * H + i] = S1;
            s_dq[i] = S1;
            s_acc[i] += S1; s_acc[H + i] += SD; s_acc[2 * H + i] += SLD; s_acc[3 * H + i] += time * S1; s_acc[4 * H + i] += Sv;
        AG_END
        AG_PAR(i)   // through proj_q into the LSTM output, then the cell of step t
            float dh = s_dh[i];
            for (int h = 0; h < H; ++h) dh += w.proj_q_kT[h * H + i] * s_dq[h];
            const float si = r.act[t * 4 * H + i], tj = r.act[t * 4 * H + H + i], sf = r.act[t * 4 * H + 2 * H + i], so = r.act[t * 4 * H + 3 * H + i];
            const float tc = tanhf(r.cs[(t + 1) * H + i]);
            const float dc = s_dc[i] + dh * so * (1.0f - tc * tc);
            const float gi = dc * tj * si * (1.0f - si), gj = dc * si * (1.0f - tj * tj);
            const float gf = dc * r.cs[t * H + i] * sf * (1.0f - sf), go = dh * tc * so * (1.0f - so);
            s_dc[i] = dc * sf;
            s_dg[i] = gi; s_dg[H + i] = gj; s_dg[2 * H + i] = gf; s_dg[3 * H + i] = go;
            r.dgates[t * 4 * H + i] = gi; r.dgates[t * 4 * H + H + i] = gj; r.dgates[t * 4 * H + 2 * H + i] = gf; r.dgates[t * 4 * H + 3 * H + i] = go;
        AG_END
        AG_PAR(i)   // into the previous LSTM output and the step's input
            float dh = 0.0f;
            const float* kr = w.lstm_kT + E + i;
            for (int j = 0; j < 4 * H; ++j) dh += kr[(long long)j * (E + H)] * s_dg[j];
            s_dh[i] = dh;
            if (i < E) {
                float dxv = 0.0f;
                const float* kx = w.lstm_kT + i;
                for (int j = 0; j < 4 * H; ++j) dxv += kx[(long long)j * (E + H)] * s_dg[j];
                r.dx[t * E + i] = dxv;
            }
        AG_END
    }

    // ---- d loss / d embedding: through proj_ref (every step) and through the LSTM inputs (the chosen nodes) ---------------------
    AG_PAR(i)
        for (int e = i; e < N * E; e += H) {
            const int n = e / E, k = e - n * E;
            float s = 0.0f;
            for (int h = 0; h < H; ++h) s += r.Dsum[n * H + h] * w.proj_ref_k[k * H + h];
            r.demb[e] = s;
        }
        for (int a = 0; a < 5; ++a) r.sums[a * H + i] = s_acc[a * H + i];
    AG_END
    AG_PAR(i)
        if (i < E)
            for (int t = 0; t < T; ++t) {
                const int node = t == 0 ? N - 1 : ag_node(r.idx[(long long)(t - 1) * r.idx_stride], N);
                r.demb[node * E + i] += r.xscale ? r.dx[t * E + i] * r.xscale[(long long)t * r.xscale_stride + i] : r.dx[t * E + i];
            }
    AG_END
}

// Gradients of the pointer module's small variables from the per-row sums [R][5][H] (S1, SD, SLD, ST, Sv), one block.  A chain
// y[n, :] = (s[n] k1 + b1) . K2 + b2 of a scalar s has  dK2[i, j] = k1[i] SS[j] + b1[i] S1[j],  db2 = S1,
// dk1[i] = sum_j K2[i, j] SS[j],  db1[i] = sum_j K2[i, j] S1[j]   with SS = sum s darg, S1 = sum darg.
struct SmallGradOut {
    float *v, *proj_q_b, *proj_ref_b;
    float *emb_k[3], *emb_b[3], *proj_k[3], *proj_b[3];   // d, ld, time (any may be null)
};

AG_FN void attention_small_grads(const ActorGradWeights& w, const float* sums, long long R, int tw, const SmallGradOut& o) {
    constexpr int H = kAgH;
    AG_SHARED float tot[5 * H];
    AG_PAR(i)
        for (int a = 0; a < 5; ++a) {
            float s = 0.0f;
            for (long long r = 0; r < R; ++r) s += sums[(r * 5 + a) * H + i];
            tot[a * H + i] = s;
        }
    AG_END
    AG_PAR(i)
        if (o.v) o.v[i] = tot[4 * H + i];
        if (o.proj_q_b) o.proj_q_b[i] = tot[i];
        if (o.proj_ref_b) o.proj_ref_b[i] = tot[i];
        const float* k1[3] = {w.emb_d_k, w.emb_ld_k, w.emb_t_k};
        const float* b1[3] = {w.emb_d_b, w.emb_ld_b, w.emb_t_b};
        const float* K2[3] = {w.proj_d_k, w.proj_ld_k, w.proj_t_k};
        for (int c = 0; c < (tw ? 3 : 2); ++c) {
            const float* SS = tot + (c + 1) * H;
            if (o.proj_b[c]) o.proj_b[c][i] = tot[i];
            float gk = 0.0f, gb = 0.0f;
            for (int j = 0; j < H; ++j) {
                gk += K2[c][i * H + j] * SS[j];
                gb += K2[c][i * H + j] * tot[j];
            }
            if (o.emb_k[c]) o.emb_k[c][i] = gk;
            if (o.emb_b[c]) o.emb_b[c][i] = gb;
            if (o.proj_k[c])
                for (int k = 0; k < H; ++k) o.proj_k[c][k * H + i] = k1[c][k] * SS[i] + b1[c][k] * tot[i];
        }
    AG_END
}

}  // namespace vrpb200
