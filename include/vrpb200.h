/*
 * vrpb200.h - C ABI of libvrpb200.so: the B200 (sm_100a) implementation of the batched
 * attention-policy rollout of jpoulletXaccount/MIT_rl-vrptw.
 *
 * The reference has no FFI: its seam is the Python protocol (Env, Attention*Actor, RNNDecodeStep,
 * RLAgent.build_model) selected in main.py:17-53 and consumed by model/attention_agent.py:13-66.
 * Each entry point below names the reference code it replaces (paths relative to the reference
 * repository root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side; `mit_rl-vrptw_b200/` is that binding plus the reference-named host classes.
 *
 * Conventions
 *  - every function returns 0 on success, a negative VRPB200_E_* code otherwise; the message of
 *    the last error of a handle is vrpb200_last_error(handle); nothing throws across the boundary;
 *  - plain pointers and sizes only; `d_` arguments are DEVICE pointers of the handle's device,
 *    `h_` arguments are HOST pointers; the caller owns every buffer, the library allocates only
 *    its weight blob (set_weights) and, for the *_host call, its own pinned/device staging;
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls are
 *    asynchronous on it unless stated otherwise;
 *  - a handle is not thread-safe; handles on different devices are independent;
 *  - rows: R = B * W, beam-major (row = w*B + b) exactly as model/attention_agent.py:173-174 and
 *    VRPTW/vrptw_utils.py:225 lay them out; node N-1 is the depot;
 *  - there is no CPU fallback: without a CUDA device vrpb200_create fails with VRPB200_E_CUDA.
 */
#ifndef VRPB200_H_
#define VRPB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRPB200_OK 0
#define VRPB200_E_ARG (-1)      /* bad argument / unsupported configuration */
#define VRPB200_E_CUDA (-2)     /* CUDA runtime error (message in last_error) */
#define VRPB200_E_WEIGHTS (-3)  /* weights missing / wrong shape */
#define VRPB200_E_STATE (-4)    /* call order (e.g. rollout before set_weights) */

#define VRPB200_TASK_VRP 0      /* input_dim 3: x, y, demand            (VRP/vrp_utils.py) */
#define VRPB200_TASK_VRPTW 1    /* input_dim 5: x, y, b_tw, e_tw, demand (VRPTW/vrptw_utils.py, UPS twin) */

#define VRPB200_PHYSICS_EUCLID 0  /* travel time = Euclidean distance (VRP, VRPTW and UPS Env) */
#define VRPB200_PHYSICS_UPS_OLD 1 /* UPS_old/vrptw_ups_utils.py:321-361, 416-424: x, y = latitude, longitude in radians;
                                     equirectangular km x (100/13 * 1.2) clicks per km, 1.5 clicks of service per package;
                                     the tour length (reward_func) is in km.  VRPTW only. */

#define VRPB200_GREEDY 0        /* model/attention_agent.py:146-147 */
#define VRPB200_STOCHASTIC 1    /* model/attention_agent.py:148-167 */
#define VRPB200_BEAM 2          /* model/attention_agent.py:169-205 */

/* args[...] keys read by Env.__init__ (VRPTW/vrptw_utils.py:172-181), RLAgent.__init__
 * (model/attention_agent.py:47-64) and task_specific_params.py:22-115. */
typedef struct vrpb200_config {
    int32_t task;            /* VRPB200_TASK_* */
    int32_t n_nodes;         /* N  (<= 128) */
    int32_t n_cust;          /* N - 1 */
    int32_t input_dim;       /* 3 or 5 */
    int32_t embedding_dim;   /* E  (<= 128) */
    int32_t hidden_dim;      /* H  (128 in this build) */
    int32_t decode_len;      /* T */
    int32_t n_glimpses;      /* shared/decode_step.py:41-46 */
    int32_t use_tanh;        /* C*tanh(u) on the pointer logits, shared/decode_step.py:49-53 */
    int32_t mask_glimpses;   /* shared/decode_step.py:222-223 */
    int32_t mask_pointer;    /* shared/decode_step.py:231-232 */
    int32_t rows_per_cta;    /* 0 = choose; else 16/32/48/64 (tuning knob, no effect on results) */
    int32_t gemm_path;       /* fused kernel family (every family gives the same results up to fp32 rounding):
                                0 = default: rollout4 (greedy, stochastic), rollout3 for beam search, the FP32-pipe kernel
                                    with glimpses or UPS_old physics,
                                1 = everything on the FP32 pipe (independent cross-check of the tensor-core kernels),
                                3 = rollout3 (transposed tcgen05 GEMMs, lock-step phases, fused beam search),
                                6 = rollout4 (per-row item lists, fast greedy tail, shared-reciprocal tanh, deep MMA
                                    buffering, FP16-split LSTM / query GEMMs, 640 threads).
                                6 also runs beam search on rollout3. */
    float capacity;
    float tanh_exploration;  /* C */
    float forget_bias;       /* BasicLSTMCell forget_bias, shared/decode_step.py:175 */
    int32_t physics;         /* VRPB200_PHYSICS_*: Env.step / mask / reward_func arithmetic (Env, env_step, reward, rollout) */
} vrpb200_config;

typedef struct vrpb200_handle vrpb200_handle;

/* Optional per-step dumps of a rollout (test instrumentation; any pointer may be NULL).
 * With any of logits/probs/masks set, every node is evaluated at every step (no early exit). */
typedef struct vrpb200_trace {
    float* d_logits;        /* [T, R, N] pointer logits after masking (shared/decode_step.py:230-232) */
    float* d_probs;         /* [T, R, N] exp(log_softmax)            (shared/decode_step.py:125-126) */
    float* d_masks;         /* [T, R, N] Env.mask seen by step t, float counts 0..3 */
    float* d_final_demand;  /* [R, N] */
    float* d_final_load;    /* [R] */
    float* d_final_time;    /* [R]  (VRPTW only) */
    float* d_final_mask;    /* [R, N] */
    int32_t* d_steps;       /* [gridDim] decode steps each instance block actually executed */
    uint64_t* d_stats;      /* [4] += node evaluations, live row-steps, block-steps x rows-per-CTA, blocks (work accounting) */
} vrpb200_trace;

/* ---- life cycle ------------------------------------------------------------------------- */
int vrpb200_create(const vrpb200_config* cfg, int device, vrpb200_handle** out);
int vrpb200_destroy(vrpb200_handle* h);
const char* vrpb200_last_error(const vrpb200_handle* h); /* h may be NULL: last create error */
int vrpb200_version(void);

/* Weights by TF-1 variable name (what tf.train.Saver stores, model/attention_agent.py:79-80):
 *   Actor/Embedding/conv1d/{kernel,bias}                                   shared/embeddings.py:37-38
 *   Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/{kernel,bias}  shared/decode_step.py:175-180
 *   Actor/Decoder/Attention/{v, emb_d, emb_ld, emb_time, proj_d, proj_ld, proj_t, proj_q, proj_ref}/...
 *   Actor/Glimpse<i>/...   (same set)                                      VRPTW/vrptw_attention.py:10-24
 * h_data[i] is a host fp32 array of n_elem[i] elements in TF layout.  Synchronous. */
int vrpb200_set_weights(vrpb200_handle* h, int n, const char* const* names,
                        const float* const* h_data, const int64_t* n_elem);

/* ---- the hot path: RLAgent.build_model(decode_type) + evaluate_batch --------------------- *
 * (model/attention_agent.py:84-240, 475-518).  d_input [B, N, input_dim] fp32.
 * d_uniforms / d_uniforms_retry [T, R] in [0,1): the two tf.random_uniform draws of my_multinomial
 * (attention_agent.py:163-167), STOCHASTIC only (NULL: counter-based in-kernel stream from `seed`).
 * Outputs (any may be NULL except d_cost): d_idx [T, R] int32 chosen node; d_cost [R] tour length
 * (reward_func, depot=None branch, VRPTW/vrptw_utils.py:397-414), beam rows back-tracked as
 * attention_agent.py:222-231; d_logprob [T, R] log(prob[row, idx]) (attention_agent.py:214);
 * d_beam_parent [T, R] int32 (BEAM only). */
int vrpb200_rollout(vrpb200_handle* h, const float* d_input, int64_t B, int mode, int beam_width,
                    const float* d_uniforms, const float* d_uniforms_retry, uint64_t seed,
                    int32_t* d_idx, float* d_cost, float* d_logprob, int32_t* d_beam_parent,
                    const vrpb200_trace* trace, void* stream);

/* Sampling retry policy of vrpb200_rollout(STOCHASTIC).  whole_batch = 0 (default): a row whose uniform lies beyond its
 * CDF redraws alone (multi-GPU friendly; probability ~1e-7 per row-step).  whole_batch = 1: the reference's my_multinomial
 * (model/attention_agent.py:148-167) - if ANY row of the batch fails its first draw, EVERY row of the batch decides that
 * step with d_uniforms_retry instead; rows failing again take node 0.  The call then synchronises the stream (it re-runs the
 * rollout once per redrawn step; vrpb200_last_redraws tells how many).  One batch = one vrpb200_rollout call on one GPU. */
int vrpb200_set_sampling_retry(vrpb200_handle* h, int whole_batch);
int vrpb200_last_redraws(vrpb200_handle* h);

/* The same call with HOST buffers (what evaluate_batch's sess.run(feed_dict) is to a caller): the batch is
 * cut into chunks that are copied H2D, decoded and copied back D2H on a few streams so that transfers
 * overlap decoding.  Pass page-locked (pinned) buffers for full PCIe speed; pageable memory works, slower.
 * h_stats (may be NULL): uint64[4] work counters as vrpb200_trace.d_stats.  Synchronous. */
int vrpb200_rollout_host(vrpb200_handle* h, const float* h_input, int64_t B, int mode, int beam_width,
                         uint64_t seed, int32_t* h_idx, float* h_cost, uint64_t* h_stats);

/* ---- Env (VRPTW/vrptw_utils.py:161-358, VRP/vrp_utils.py:133-255) ------------------------- *
 * State arrays are the Env attributes of the same name, fp32, R rows:
 *   demand [R,N], load [R], time [R] (VRPTW), mask [R,N], prev_xy [R,2] (previous_x, previous_y).
 * env_reset: Env.reset(beam_width) (vrptw_utils.py:199-252).
 * env_step:  Env.step(idx, beam_parent) (vrptw_utils.py:254-358); d_idx [R] int64, d_beam_parent
 *            [R] int64 or NULL; d_sat [R] out (may be NULL).  With beam_parent the state is gathered
 *            from the *_in arrays into the *_out arrays (they must not alias); without it in == out
 *            is allowed. */
int vrpb200_env_reset(vrpb200_handle* h, const float* d_input, int64_t B, int beam_width,
                      float* d_demand, float* d_load, float* d_time, float* d_mask, float* d_prev_xy,
                      void* stream);
int vrpb200_env_step(vrpb200_handle* h, const float* d_input, int64_t B, int beam_width,
                     const int64_t* d_idx, const int64_t* d_beam_parent,
                     const float* d_demand_in, const float* d_load_in, const float* d_time_in,
                     const float* d_prev_xy_in,
                     float* d_demand, float* d_load, float* d_time, float* d_mask, float* d_prev_xy,
                     float* d_sat, void* stream);

/* ---- standalone model pieces (drop-in surface; as-written arithmetic, not the fused path) -- *
 * embed:       LinearEmbedding.__call__ (shared/embeddings.py:40-46): d_input_pnt [M, D-1] -> [M, E].
 * attention:   Attention*Actor.__call__(query, ref, env) (VRPTW/vrptw_attention.py:30-84,
 *              VRP/vrp_attention.py:27-76): att = -1 pointer ("Actor/Decoder/Attention"), i>=0 glimpse i;
 *              d_query [R,H], d_ref [R,N,E], env demand/load/time -> d_e [R,N,H] (may be NULL), d_logits [R,N].
 * decode_step: RNNDecodeStep.step (shared/decode_step.py:97-128,182-234): d_decoder_inp [R,E],
 *              d_context [R,N,E], env demand/load/time/mask, d_c/d_h [R,H] in-out ->
 *              d_logit, d_prob, d_logprob [R,N]. */
int vrpb200_embed(vrpb200_handle* h, const float* d_input_pnt, int64_t M, float* d_emb, void* stream);
int vrpb200_attention(vrpb200_handle* h, int att, int64_t R, const float* d_query, const float* d_ref,
                      const float* d_demand, const float* d_load, const float* d_time,
                      float* d_e, float* d_logits, void* stream);
int vrpb200_decode_step(vrpb200_handle* h, int64_t R, const float* d_decoder_inp, const float* d_context,
                        const float* d_demand, const float* d_load, const float* d_time, const float* d_mask,
                        float* d_c, float* d_h, float* d_logit, float* d_prob, float* d_logprob, void* stream);

/* reward_func(sample_solution) depot=None branch (VRPTW/vrptw_utils.py:397-414): d_actions [T, R, A]
 * (first two columns x, y) -> d_cost [R]. */
int vrpb200_reward(vrpb200_handle* h, const float* d_actions, int64_t T, int64_t R, int A, float* d_cost,
                   void* stream);

/* reward_func(sample_solution, decode_len, n_nodes, depot) with a depot - the min_trucks objective
 * (VRPTW/vrptw_utils.py:384-395, 410-412; VRP/vrp_utils.py:275-285, 302-304): d_actions [T, R, A] (A = 2: x, y; A = 4:
 * x, y, b_tw, e_tw), d_depot [R, A] -> d_reward [R] = 70/n_nodes x depot returns + 30 x route length / sum of the
 * distances to the depot.  As in the reference the depot test compares column 0 only and the normalising distance
 * runs over all A columns. */
int vrpb200_reward_min_trucks(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R, int A,
                              float n_nodes, float* d_reward, void* stream);

/* reward_func with a depot, UPS VRPTW form (UPS/vrptw_ups_utils.py:387-419): 100 x returns to the depot (a stay counts
 * once, column 0 is compared) + closed route length over x, y.  d_actions [T, R, A], d_depot [R, A] -> d_reward [R]. */
int vrpb200_reward_min_trucks_ups(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R, int A,
                                  float* d_reward, void* stream);

/* reward_func with a depot, UPS_old form (UPS_old/vrptw_ups_utils.py:400-428): 70/n_nodes x (steps at the depot - every one
 * counts, column 0 is compared) + 30 x closed route length in equirectangular km over x, y (latitude, longitude in radians)
 * / sum of the Euclidean distances to the depot over all A columns.  d_actions [T, R, A], d_depot [R, A] -> d_reward [R]. */
int vrpb200_reward_min_trucks_ups_old(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R,
                                      int A, float n_nodes, float* d_reward, void* stream);

/* ---- REINFORCE pieces around the rollout: value, losses, gradients, update ------------------------------ *
 * Critic value v of RLAgent.build_model('stochastic') (model/attention_agent.py:243-270; critics
 * VRPTW/vrptw_attention.py:87-169, VRP/vrp_attention.py:79-140): n_process_blocks un-masked attention passes over the
 * initial instance + Linear/L1 (relu) + Linear/L2.  Needs the Critic/Process/P<i>/{v, emb_d, proj_d, proj_q, proj_e
 * [, emb_begin_tw, proj_end_tw]} and Critic/Linear/{L1, L2} variables in set_weights (VRPB200_E_STATE otherwise).
 * d_input [B, N, input_dim] -> d_v [B]. */
int vrpb200_critic(vrpb200_handle* h, const float* d_input, int64_t B, float* d_v, void* stream);

/* build_train_step's losses (model/attention_agent.py:286-294): d_losses[0] = mean((R - v) * sum_t logprob[t]) (actor,
 * v as a constant), d_losses[1] = mean((R - v)^2) (critic).  d_R, d_v [R], d_logprob [T, R] (vrpb200_rollout's). */
int vrpb200_reinforce_loss(vrpb200_handle* h, const float* d_R, const float* d_v, const float* d_logprob, int64_t T, int64_t R,
                           float* d_losses, void* stream);

/* compute_gradients(actor_loss, Actor variables) of build_train_step (model/attention_agent.py:289-297): back-propagation of
 *   L = sum_r d_weight[r] * sum_t log prob_t[r, d_idx[t, r]]        (d_weight[r] = (R[r] - v[r]) / B gives the reference's actor_loss)
 * through the teacher-forced rollout - Env state along the tour d_idx (constants of the graph), BasicLSTMCell through time,
 * the pointer attention, mask, log_softmax, and the embedding in both of its uses (shared/decode_step.py:97-128, 182-234;
 * VRPTW/vrptw_attention.py:30-84; shared/embeddings.py:40-46).  d_input [B, N, input_dim]; d_idx [T, B] int32 (vrpb200_rollout's
 * d_idx, greedy or stochastic, beam_width 1).  names / d_grads / n_elem: the TF variable names as in vrpb200_set_weights, one
 * DEVICE buffer per variable (same element count as the variable) that receives its gradient; variables left out are skipped.
 * d_input_scale [T, B, embedding_dim] or NULL: the LSTM-input dropout of DropoutWrapper(input_keep_prob = 1 - dropout)
 * (shared/decode_step.py:177-179) as keep-mask / keep-probability - step t's input is the chosen node's embedding times
 * d_input_scale[t, b, :]; it must be the scale the tour was decoded with (NULL: no dropout).  n_glimpses must be 0.  Indices outside
 * [0, N) are clamped.  Work memory (~13 * T * hidden_dim floats per row: 1.4 MB per row at VRPTW-100, sized for training batches,
 * VRPB200_E_CUDA if it cannot be allocated) is owned by the handle.  Deterministic: every sum runs in one fixed order. */
int vrpb200_actor_grad(vrpb200_handle* h, const float* d_input, int64_t B, const int32_t* d_idx, const float* d_weight,
                       const float* d_input_scale, int n,
                       const char* const* names, float* const* d_grads, const int64_t* n_elem, void* stream);

/* compute_gradients(critic_loss, Critic variables) of build_train_step (model/attention_agent.py:290, 298-299):
 * critic_loss = tf.losses.mean_squared_error(R, v), back-propagated through Linear/L2, relu, Linear/L1 and the n_process_blocks
 * attention blocks of the critic (VRPTW/vrptw_attention.py:87-169, VRP/vrp_attention.py:79-140; emb_begin_tw / proj_end_tw receive
 * the gradient of both windows).  The actor's embedding is a constant of this graph.  d_input [B, N, input_dim], d_R [B] (the
 * rollout's tour lengths); names / d_grads / n_elem as in vrpb200_actor_grad, over the Critic/* variables; d_v [B] (may be NULL)
 * receives the value v the gradient was taken at.  Needs the Critic/* variables in set_weights (VRPB200_E_STATE otherwise). */
int vrpb200_critic_grad(vrpb200_handle* h, const float* d_input, int64_t B, const float* d_R, int n, const char* const* names,
                        float* const* d_grads, const int64_t* n_elem, float* d_v, void* stream);

/* clip_by_norm + AdamOptimizer.apply_gradients of build_train_step (model/attention_agent.py:303-311) for n variables at once:
 * per variable g = grad * max_norm / max(||grad||, max_norm); m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
 * var -= lr sqrt(1 - b2^step) / (1 - b1^step) * m / (sqrt(v) + eps)  (TF-1 Adam; step = 1 for the first update).
 * Every pointer is a DEVICE buffer of n_elem[i] floats; d_m / d_v are the optimizer slots, zero before the first step. */
int vrpb200_clip_adam(vrpb200_handle* h, int n, float* const* d_vars, const float* const* d_grads, float* const* d_m,
                      float* const* d_v, const int64_t* n_elem, float max_norm, float lr, int64_t step, float beta1, float beta2,
                      float eps, void* stream);

/* Measured throughput of one SM pipe on the handle's device, for the roofline denominators that
 * MEASURED_PEAKS.json does not hold: kind 0 = MUFU (ex2/rcp, ops/s), 1 = FP32 FMA (FMA/s, 2 flop each).
 * Runs a register-only micro-benchmark on every SM for ~`iters` x 4096 ops per thread.  Synchronous. */
int vrpb200_pipe_peak(vrpb200_handle* h, int kind, int iters, double* ops_per_second);

/* Launch statistics of the last rollout on this handle (for bench.py): grid size, block size,
 * dynamic shared memory bytes, rows per CTA, number of kernels launched. */
int vrpb200_last_launch(const vrpb200_handle* h, int32_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* VRPB200_H_ */
