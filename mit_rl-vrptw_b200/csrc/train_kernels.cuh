// The backward / update side of RLAgent.build_train_step (model/attention_agent.py:272-322):
//   actor_grad_rows_kernel   one block per row: actor_grad_row.h (teacher-forced LSTM forward + back-propagation through time)
//   outer_reduce_kernel      dW[k, j] = sum over rows and steps of A[., k] B[., j]   (every kernel-matrix gradient)
//   col_reduce_kernel        db[j]    = sum over rows and steps of B[., j]           (bias gradients)
//   attention_small_grads_kernel   v, emb_* / proj_* (rank-1 structure of the scalar -> H chains), proj_q / proj_ref biases
//   critic_grad_rows_kernel / critic_small_grads_kernel   the same for critic_loss = mean_squared_error(R, v) (critic_grad_row.h)
//   clip_adam_kernel         tf.clip_by_norm per variable (:303-307) + tf.train.AdamOptimizer.apply_gradients (:310-311)
// Every sum runs in one fixed order (thread per output element, serial loop): gradients are bit-reproducible.
#pragma once
#include "actor_grad_row.h"
#include "critic_grad_row.h"
#include "common.cuh"

namespace vrpb200 {

struct ActorGradBatch {   // per-row strides (in floats) of the work arrays; row r uses base + r * stride
    const float* input;       // [R, N, D]
    const int* idx;           // [T, R]
    const float* weight;      // [R]
    const float* xscale;      // [T, R, E] LSTM-input dropout scale, or null
    const float *demand, *mask, *load, *time;   // [T + 1][R][N], [T + 1][R]
    float *hs, *cs, *act, *xin, *dx, *emb, *eref, *erefT, *dgates, *dq, *Dsum, *demb, *sums;
    long long R;
};

__global__ void __launch_bounds__(128) actor_grad_rows_kernel(ActorGradWeights w, ActorGradDims d, ActorGradBatch b) {
    const long long r = blockIdx.x;
    const int H = kAgH, N = d.N, E = d.E, T = d.T;
    ActorGradRow row;
    row.input = b.input + r * N * d.D;
    row.idx = b.idx + r;
    row.idx_stride = b.R;
    row.weight = b.weight[r];
    row.xscale = b.xscale ? b.xscale + r * E : nullptr;
    row.xscale_stride = b.R * E;
    row.demand = b.demand + r * N;
    row.mask = b.mask + r * N;
    row.node_stride = b.R * N;
    row.load = b.load + r;
    row.time = b.time ? b.time + r : nullptr;
    row.scalar_stride = b.R;
    row.hs = b.hs + r * (T + 1) * H;
    row.cs = b.cs + r * (T + 1) * H;
    row.act = b.act + r * T * 4 * H;
    row.xin = b.xin + r * T * E;
    row.dx = b.dx + r * T * E;
    row.emb = b.emb + r * N * E;
    row.eref = b.eref + r * N * H;
    row.erefT = b.erefT + r * H * 128;
    row.dgates = b.dgates + r * T * 4 * H;
    row.dq = b.dq + r * T * H;
    row.Dsum = b.Dsum + r * N * H;
    row.demb = b.demb + r * N * E;
    row.sums = b.sums + r * 5 * H;
    actor_grad_row(w, d, row);
}

__global__ void idx_to_i64_kernel(const int* __restrict__ in, long long* __restrict__ out, long long n, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ag_node(in[i], N);
}

// part[y][k * J + j] = sum over the rows r of split y, then over t < Tn, of A[r * sAr + t * sAt + k] * B[r * sBr + t * sBt + j];
// blockIdx.y = split (rows_per_split consecutive rows).  sum_splits_kernel adds the splits in order: two fixed-order stages.
__global__ void outer_reduce_kernel(const float* __restrict__ A, long long sAr, long long sAt, const float* __restrict__ Bm,
                                    long long sBr, long long sBt, long long R, int rows_per_split, int Tn, int K, int J,
                                    float* __restrict__ part) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)K * J) return;
    const int k = (int)(i / J), j = (int)(i - (long long)k * J);
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
    float s = 0.0f;
    for (long long r = r0; r < r1; ++r) {
        const float* a = A + r * sAr + k;
        const float* b = Bm + r * sBr + j;
        float sr = 0.0f;
        for (int t = 0; t < Tn; ++t) sr = fmaf(a[t * sAt], b[t * sBt], sr);
        s += sr;
    }
    part[(long long)blockIdx.y * K * J + i] = s;
}

__global__ void col_reduce_kernel(const float* __restrict__ Bm, long long sBr, long long sBt, long long R, int rows_per_split, int Tn,
                                  int J, float* __restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
    float s = 0.0f;
    for (long long r = r0; r < r1; ++r) {
        const float* b = Bm + r * sBr + j;
        float sr = 0.0f;
        for (int t = 0; t < Tn; ++t) sr += b[t * sBt];
        s += sr;
    }
    part[(long long)blockIdx.y * J + j] = s;
}

__global__ void sum_splits_kernel(const float* __restrict__ part, int S, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int y = 0; y < S; ++y) s += part[(long long)y * n + i];
    out[i] = s;
}

__global__ void __launch_bounds__(128) attention_small_grads_kernel(ActorGradWeights w, const float* __restrict__ sums, long long R,
                                                                    int tw, SmallGradOut o) {
    attention_small_grads(w, sums, R, tw, o);
}

struct CriticGradBatch {
    const float* input;     // [R, N, D]
    const float* Rw;        // [R]
    float *emb, *ee, *eeT, *prob, *hyin, *De, *sums, *dq, *l1, *dl1, *vdv;
};

__global__ void __launch_bounds__(128) critic_grad_rows_kernel(CriticGradWeights w, CriticGradDims d, CriticGradBatch b) {
    const long long r = blockIdx.x;
    const int H = kAgH, N = d.N, E = d.E, P = d.P;
    CriticGradRow row;
    row.input = b.input + r * N * d.D;
    row.R = b.Rw[r];
    row.emb = b.emb + r * N * E;
    row.ee = b.ee + r * P * N * H;
    row.eeT = b.eeT + r * P * H * 128;
    row.prob = b.prob + r * P * 128;
    row.hyin = b.hyin + r * (P + 1) * H;
    row.De = b.De + r * P * N * H;
    row.sums = b.sums + r * P * 4 * H;
    row.dq = b.dq + r * P * H;
    row.l1 = b.l1 + r * H;
    row.dl1 = b.dl1 + r * H;
    row.vdv = b.vdv + r * 2;
    critic_grad_row(w, d, row);
}

__global__ void __launch_bounds__(128) critic_small_grads_kernel(CriticBlockWeights bw, const float* __restrict__ sums, long long R, int P,
                                                                 int b, int tw, CriticSmallGradOut o) {
    critic_small_grads(bw, sums, R, P, b, tw, o);
}

// tf.clip_by_norm(grad, max_norm) per variable, then Adam (TF-1: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m, v moments;
// var -= lr_t m / (sqrt(v) + eps)).  One block per variable; the norm is summed in a fixed order.
struct AdamVar {
    float* var;
    const float* grad;
    float *m, *v;
    long long n;
};
__global__ void __launch_bounds__(256) clip_adam_kernel(const AdamVar* __restrict__ vars, float max_norm, float lr_t, float beta1,
                                                        float beta2, float eps) {
    __shared__ float part[256];
    const AdamVar a = vars[blockIdx.x];
    float s = 0.0f;
    for (long long i = threadIdx.x; i < a.n; i += 256) s = fmaf(a.grad[i], a.grad[i], s);
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    const float norm = sqrtf(part[0]);
    const float scale = norm > max_norm ? max_norm / norm : 1.0f;    // clip_by_norm: t * clip / max(norm, clip)
    for (long long i = threadIdx.x; i < a.n; i += 256) {
        const float g = a.grad[i] * scale;
        const float m = beta1 * a.m[i] + (1.0f - beta1) * g;
        const float v = beta2 * a.v[i] + (1.0f - beta2) * g * g;
        a.m[i] = m;
        a.v[i] = v;
        a.var[i] -= lr_t * m / (sqrtf(v) + eps);
    }
}

}  // namespace vrpb200
