// C ABI of libvrpb200.so (include/vrpb200.h): handle, weight pre-processing, launchers.
// No torch types, no CPU fallback: every compute entry point launches sm_100a kernels.
#include "../../include/vrpb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cstdlib>
#include <cuda_fp16.h>
#include "launchers.h"
#include "standalone_kernels.cuh"
#include "critic_kernels.cuh"
#include "train_kernels.cuh"

using namespace vrpb200;

namespace {
thread_local std::string g_create_error;
}

struct vrpb200_handle {
    vrpb200_config cfg;
    int device = 0;
    bool tw = false;
    int D1 = 0;   // input_dim - 1
    std::string err;
    bool have_weights = false;
    float* d_blob = nullptr;          // processed + raw weights, one allocation
    size_t blob_floats = 0;
    // offsets (in floats) into the blob
    const float* Wg = nullptr;
    const float* bg = nullptr;
    const float *WgT3 = nullptr, *WxT = nullptr, *bgT = nullptr;
    const float* WqTc[8] = {nullptr};
    const float *WgTh = nullptr, *WqTh[8] = {nullptr};   // FP16 hi/lo tiles of W_h^T / W_q^T (rollout4/5)
    AttWeights att[8];
    AttWeights* d_gatt = nullptr;   // device copy of att[0..n_glimpses-1]
    float* d_critic = nullptr;      // critic constants (critic_kernels.cuh), null until set_weights saw Critic/* variables
    int critic_blocks = 0;
    float* d_loss = nullptr;        // [2] scratch of reinforce_loss
    float* d_critic_raw = nullptr;  // raw Critic/* variables for vrpb200_critic_grad
    CriticGradWeights critic_raw;   // device pointers into d_critic_raw (+ the actor's embedding)
    void* d_adam_vars = nullptr;    // variable table of vrpb200_clip_adam
    int adam_vars_cap = 0;
    char* d_train_ws = nullptr;     // work arrays of vrpb200_actor_grad / vrpb200_clip_adam (grow-only)
    size_t train_ws_bytes = 0;
    int whole_batch_retry = 0;      // vrpb200_set_sampling_retry
    int* d_retry = nullptr;         // [2][T]: retry_steps | fail_steps
    int last_redraws = 0;
    RawModel raw;
    const float *raw_emb_k = nullptr, *raw_emb_b = nullptr;
    const float *raw_lstm_kT = nullptr, *raw_pq_kT = nullptr;   // [4H, E + H], pointer module's proj_q [H(out), H(in)]
    // staging of the *_host call: kChunks in-flight chunks, one stream each
    static constexpr int kChunks = 3;
    float* dev_in[kChunks] = {nullptr, nullptr, nullptr};
    int32_t* dev_idx[kChunks] = {nullptr, nullptr, nullptr};
    float* dev_cost[kChunks] = {nullptr, nullptr, nullptr};
    size_t dev_in_bytes[kChunks] = {0, 0, 0}, dev_idx_bytes[kChunks] = {0, 0, 0}, dev_cost_bytes[kChunks] = {0, 0, 0};
    cudaStream_t streams[kChunks] = {nullptr, nullptr, nullptr};
    unsigned long long* dev_stats = nullptr;
    // rollout4: re-packed node features, one workspace per staging slot (+ one for vrpb200_rollout)
    float4* ws_feat4[kChunks + 1] = {nullptr, nullptr, nullptr, nullptr};
    size_t ws_feat4_bytes[kChunks + 1] = {0, 0, 0, 0};
    int ws_slot = kChunks;            // which workspace the next rollout_impl call uses
    int32_t last_launch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float vm2[128];                   // -2 v of the pointer attention and sum v (rollout4 kernel parameters)
    float sumv = 0.0f;
    int max_smem_optin = 0;
    int num_sms = 0;
};

namespace {

int fail(vrpb200_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(h, VRPB200_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct HostTensor { const float* p; int64_t n; };

float tf32_round(float x) {   // round-to-nearest-even onto TF32's 10 explicit mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

const char* kLstm = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell";

std::string att_prefix(const vrpb200_config& c, int a) {   // a in [0, n_glimpses]: glimpses then pointer
    if (a < c.n_glimpses) return "Actor/Glimpse" + std::to_string(a);
    return "Actor/Decoder/Attention";
}

template <class T>
int ensure(vrpb200_handle* h, T** p, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    CUDA_TRY(h, cudaMalloc((void**)p, need));
    *have = need;
    return 0;
}

}  // namespace

extern "C" {

int vrpb200_version(void) { return 1; }

const char* vrpb200_last_error(const vrpb200_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vrpb200_create(const vrpb200_config* cfg, int device, vrpb200_handle** out) {
    if (!cfg || !out) return fail(nullptr, VRPB200_E_ARG, "null argument");
    *out = nullptr;
    const vrpb200_config& c = *cfg;
    if (c.task != VRPB200_TASK_VRP && c.task != VRPB200_TASK_VRPTW) return fail(nullptr, VRPB200_E_ARG, "task must be 0 (vrp) or 1 (vrptw)");
    const int want_dim = c.task == VRPB200_TASK_VRPTW ? 5 : 3;
    if (c.input_dim != want_dim) return fail(nullptr, VRPB200_E_ARG, "input_dim %d does not match task (want %d)", c.input_dim, want_dim);
    if (c.n_nodes < 2 || c.n_nodes > 128) return fail(nullptr, VRPB200_E_ARG, "n_nodes %d outside [2,128]", c.n_nodes);
    if (c.n_cust != c.n_nodes - 1) return fail(nullptr, VRPB200_E_ARG, "n_cust must be n_nodes-1 (depot is the last node)");
    if (c.hidden_dim != kH) return fail(nullptr, VRPB200_E_ARG, "hidden_dim %d not supported by this build (128)", c.hidden_dim);
    if (c.embedding_dim < 1 || c.embedding_dim > 128) return fail(nullptr, VRPB200_E_ARG, "embedding_dim %d outside [1,128]", c.embedding_dim);
    if (c.decode_len < 1) return fail(nullptr, VRPB200_E_ARG, "decode_len must be >= 1");
    if (c.n_glimpses < 0 || c.n_glimpses > 7) return fail(nullptr, VRPB200_E_ARG, "n_glimpses outside [0,7]");
    if (c.gemm_path != 0 && c.gemm_path != 1 && c.gemm_path != 3 && c.gemm_path != 6)
        return fail(nullptr, VRPB200_E_ARG, "gemm_path must be 0 (default), 1 (FP32 pipe), 3 (rollout3) or 6 (rollout4)");
    if (c.physics != VRPB200_PHYSICS_EUCLID && c.physics != VRPB200_PHYSICS_UPS_OLD) return fail(nullptr, VRPB200_E_ARG, "physics must be 0 (Euclidean) or 1 (UPS_old)");
    if (c.physics == VRPB200_PHYSICS_UPS_OLD && c.task != VRPB200_TASK_VRPTW) return fail(nullptr, VRPB200_E_ARG, "UPS_old physics is a VRPTW environment");
    if (c.rows_per_cta != 0 && c.rows_per_cta != 16 && c.rows_per_cta != 32 && c.rows_per_cta != 48 && c.rows_per_cta != 64)
        return fail(nullptr, VRPB200_E_ARG, "rows_per_cta must be 0, 16, 32, 48 or 64");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VRPB200_E_CUDA, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, VRPB200_E_ARG, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, VRPB200_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, VRPB200_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    vrpb200_handle* h = new vrpb200_handle();
    h->cfg = c;
    h->device = device;
    h->tw = c.task == VRPB200_TASK_VRPTW;
    h->D1 = c.input_dim - 1;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return VRPB200_OK;
}

int vrpb200_destroy(vrpb200_handle* h) {
    if (!h) return VRPB200_OK;
    cudaSetDevice(h->device);
    if (h->d_blob) cudaFree(h->d_blob);
    if (h->d_gatt) cudaFree(h->d_gatt);
    if (h->d_critic) cudaFree(h->d_critic);
    if (h->d_loss) cudaFree(h->d_loss);
    if (h->d_train_ws) cudaFree(h->d_train_ws);
    if (h->d_critic_raw) cudaFree(h->d_critic_raw);
    if (h->d_adam_vars) cudaFree(h->d_adam_vars);
    if (h->d_retry) cudaFree(h->d_retry);
    for (int i = 0; i < vrpb200_handle::kChunks; ++i) {
        if (h->dev_in[i]) cudaFree(h->dev_in[i]);
        if (h->dev_idx[i]) cudaFree(h->dev_idx[i]);
        if (h->dev_cost[i]) cudaFree(h->dev_cost[i]);
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    if (h->dev_stats) cudaFree(h->dev_stats);
    for (int i = 0; i <= vrpb200_handle::kChunks; ++i)
        if (h->ws_feat4[i]) cudaFree(h->ws_feat4[i]);
    delete h;
    return VRPB200_OK;
}

int vrpb200_set_weights(vrpb200_handle* h, int n, const char* const* names, const float* const* h_data,
                        const int64_t* n_elem) {
    if (!h) return VRPB200_E_ARG;
    const vrpb200_config& c = h->cfg;
    const int E = c.embedding_dim, D1 = h->D1, H = kH;
    std::map<std::string, HostTensor> m;
    for (int i = 0; i < n; ++i) m[names[i]] = HostTensor{h_data[i], n_elem[i]};
    std::string missing;
    auto get = [&](const std::string& name, int64_t want) -> const float* {
        auto it = m.find(name);
        if (it == m.end()) { missing += " " + name; return nullptr; }
        if (it->second.n != want) { missing += " " + name + "(size " + std::to_string(it->second.n) + " != " + std::to_string(want) + ")"; return nullptr; }
        return it->second.p;
    };
    const float* emb_k = get("Actor/Embedding/conv1d/kernel", (int64_t)D1 * E);
    const float* emb_b = get("Actor/Embedding/conv1d/bias", E);
    const float* lstm_k = get(std::string(kLstm) + "/kernel", (int64_t)(E + H) * 4 * H);
    const float* lstm_b = get(std::string(kLstm) + "/bias", 4 * H);
    const int natt = c.n_glimpses + 1;
    struct RawHost { const float *v, *ed_k, *ed_b, *eld_k, *eld_b, *et_k, *et_b, *pd_k, *pd_b, *pld_k, *pld_b, *pt_k, *pt_b, *pq_k, *pq_b, *pr_k, *pr_b; };
    std::vector<RawHost> rh(natt);
    for (int a = 0; a < natt; ++a) {
        const std::string p = att_prefix(c, a);
        RawHost& r = rh[a];
        r.v = get(p + "/v", H);
        r.ed_k = get(p + "/emb_d/kernel", H);   r.ed_b = get(p + "/emb_d/bias", H);
        r.eld_k = get(p + "/emb_ld/kernel", H); r.eld_b = get(p + "/emb_ld/bias", H);
        r.pd_k = get(p + "/proj_d/kernel", H * H);   r.pd_b = get(p + "/proj_d/bias", H);
        r.pld_k = get(p + "/proj_ld/kernel", H * H); r.pld_b = get(p + "/proj_ld/bias", H);
        r.pq_k = get(p + "/proj_q/kernel", H * H);   r.pq_b = get(p + "/proj_q/bias", H);
        r.pr_k = get(p + "/proj_ref/kernel", (int64_t)E * H); r.pr_b = get(p + "/proj_ref/bias", H);
        if (h->tw) {
            r.et_k = get(p + "/emb_time/kernel", H); r.et_b = get(p + "/emb_time/bias", H);
            r.pt_k = get(p + "/proj_t/kernel", H * H); r.pt_b = get(p + "/proj_t/bias", H);
        } else {
            r.et_k = r.et_b = r.pt_k = r.pt_b = nullptr;
        }
    }
    if (!missing.empty()) return fail(h, VRPB200_E_WEIGHTS, "missing or mis-sized variables:%s", missing.c_str());

    // ---- build the blob on the host --------------------------------------------------------------
    std::vector<float> blob;
    auto reserve = [&](size_t nfl) { size_t o = blob.size(); blob.resize(o + ((nfl + 31) / 32) * 32, 0.0f); return o; };   // 128 B aligned pieces
    auto put = [&](const float* src, size_t nfl) { size_t o = reserve(nfl); memcpy(&blob[o], src, nfl * 4); return o; };

    // phase A: gates kernel, rows 0..127 = W_lstm[E + k], rows 128..128+D1-1 = W_emb . W_lstm[:E], rest zero;
    // layout [slab s][kk 8][uh 2][part 2][lane 32][gate 4], unit = uh*64 + 2*lane + part, k = 8 s + kk.
    const size_t oWg = reserve((size_t)kGateSlabs * kSlabFloats);
    const size_t obg = reserve(4 * H);
    const size_t oWgT3 = reserve((size_t)kV3UsesG * kSlabFloats), oWxT = reserve(16 * H), obgT = reserve(4 * H);
    const size_t oWgTh = reserve((size_t)kHUsesG * kSlabFloats);
    // FP16 two-term split into a [128 units][K = 16] K-major tile (8 x 16-byte core matrices, as the TF32 tiles): hi at hi_off, lo 4 KB on
    auto put_f16 = [&](size_t slab_float_off, int tile, int unit, int kl, float val) {
        uint16_t* base = reinterpret_cast<uint16_t*>(&blob[slab_float_off]) + (size_t)tile * 4096;
        const __half hh = __float2half_rn(val);
        const __half ll = __float2half_rn(val - __half2float(hh));
        const size_t idx = (size_t)(unit / 8) * 128 + (kl / 8) * 64 + (unit % 8) * 8 + (kl % 8);
        base[idx] = __half_as_ushort(hh);
        base[2048 + idx] = __half_as_ushort(ll);
    };
    {
        std::vector<double> Wx((size_t)D1 * 4 * H, 0.0), bx(4 * H, 0.0);
        for (int col = 0; col < 4 * H; ++col) {
            for (int cc = 0; cc < D1; ++cc) {
                double s = 0;
                for (int e2 = 0; e2 < E; ++e2) s += (double)emb_k[cc * E + e2] * lstm_k[(size_t)e2 * 4 * H + col];
                Wx[(size_t)cc * 4 * H + col] = s;
            }
            double s = lstm_b[col];
            for (int e2 = 0; e2 < E; ++e2) s += (double)emb_b[e2] * lstm_k[(size_t)e2 * 4 * H + col];
            bx[col] = s;
        }
        for (int k = 0; k < kKG; ++k) {
            const int s = k / 8, kk = k % 8;
            for (int unit = 0; unit < H; ++unit) {
                const int uh = unit / 64, lane = (unit % 64) / 2, part = unit % 2;
                for (int g = 0; g < 4; ++g) {
                    const int col = g * H + unit;
                    float val = 0.0f;
                    if (k < H) val = lstm_k[(size_t)(E + k) * 4 * H + col];
                    else if (k - H < D1) val = (float)Wx[(size_t)(k - H) * 4 * H + col];
                    blob[oWg + (size_t)s * kSlabFloats + ((size_t)((kk * 2 + uh) * 2 + part) * 32 + lane) * 4 + g] = val;
                }
            }
        }
        for (int unit = 0; unit < H; ++unit)
            for (int g = 0; g < 4; ++g) blob[obg + unit * 4 + g] = (float)bx[g * H + unit];
        // rollout3: transposed tiles of the h part, per (k-chunk, pair of gate types): [128 units][K=8] K-major; plus the
        // fp32 input part W_x^T [feature][gate type][unit] and bias [gate type][unit]
        for (int u = 0; u < kV3UsesG; ++u) {
            const int kc = u >> 1, mp = u & 1;
            for (int mi = 0; mi < 2; ++mi)
                for (int unit = 0; unit < H; ++unit)
                    for (int kl = 0; kl < 8; ++kl) {
                        const int m = 2 * mp + mi, k = kc * 8 + kl;
                        const float val = lstm_k[(size_t)(E + k) * 4 * H + m * H + unit];
                        const float hi = tf32_round(val), lo = tf32_round(val - hi);
                        const size_t o = (size_t)(unit / 8) * 64 + (kl / 4) * 32 + (unit % 8) * 4 + (kl % 4);
                        blob[oWgT3 + (size_t)u * kSlabFloats + mi * 2048 + o] = hi;
                        blob[oWgT3 + (size_t)u * kSlabFloats + mi * 2048 + 1024 + o] = lo;
                    }
        }
        for (int u = 0; u < kHUsesG; ++u) {   // rollout4/5: the same tiles with K = 16 per k-chunk, FP16 hi/lo
            const int kc = u >> 1, mp = u & 1;
            for (int mi = 0; mi < 2; ++mi)
                for (int unit = 0; unit < H; ++unit)
                    for (int kl = 0; kl < 16; ++kl)
                        put_f16(oWgTh + (size_t)u * kSlabFloats, mi, unit, kl, lstm_k[(size_t)(E + kc * 16 + kl) * 4 * H + (2 * mp + mi) * H + unit]);
        }
        for (int m = 0; m < 4; ++m)
            for (int unit = 0; unit < H; ++unit) {
                blob[obgT + m * H + unit] = (float)bx[m * H + unit];
                for (int cc = 0; cc < 4; ++cc)
                    blob[oWxT + (cc * 4 + m) * H + unit] = cc < D1 ? (float)Wx[(size_t)cc * 4 * H + m * H + unit] : 0.0f;
            }
    }
    struct AttOff { size_t Wq, bq, A, gd, gl, gt, v, Au, ce, WqTc, Mq, cq, WqTh; };
    std::vector<AttOff> ao(natt);
    std::vector<std::vector<double>> Au_d(natt, std::vector<double>(4 * H, 0.0)), ce_d(natt, std::vector<double>(H, 0.0)), bq_d(natt, std::vector<double>(H, 0.0));
    for (int a = 0; a < natt; ++a) {
        const RawHost& r = rh[a];
        AttOff& o = ao[a];
        o.Wq = put(r.pq_k, (size_t)H * H);   // [k][unit] row-major == 4 slabs of 32 k-rows
        o.bq = reserve(H); o.A = reserve(4 * H); o.gd = reserve(H); o.gl = reserve(H); o.gt = reserve(H);
        o.v = put(r.v, H); o.Au = reserve(4 * H); o.ce = reserve(H); o.Mq = reserve(4 * H); o.cq = reserve(H);
        o.WqTc = reserve((size_t)kV3UsesQ * kSlabFloats);
        o.WqTh = reserve((size_t)kHUsesQ * kSlabFloats);
        for (int u = 0; u < kHUsesQ; ++u)
            for (int kk = 0; kk < 2; ++kk)
                for (int n = 0; n < H; ++n)
                    for (int kl = 0; kl < 16; ++kl)
                        put_f16(o.WqTh + (size_t)u * kSlabFloats, kk, n, kl, r.pq_k[(size_t)((2 * u + kk) * 16 + kl) * H + n]);
        for (int u = 0; u < kV3UsesQ; ++u)
            for (int kk = 0; kk < 2; ++kk)
                for (int n = 0; n < H; ++n)
                    for (int kl = 0; kl < 8; ++kl) {
                        const int k = (2 * u + kk) * 8 + kl;
                        const float val = r.pq_k[(size_t)k * H + n];
                        const float hi = tf32_round(val), lo = tf32_round(val - hi);
                        const size_t oo = (size_t)(n / 8) * 64 + (kl / 4) * 32 + (n % 8) * 4 + (kl % 4);
                        blob[o.WqTc + (size_t)u * kSlabFloats + kk * 2048 + oo] = hi;
                        blob[o.WqTc + (size_t)u * kSlabFloats + kk * 2048 + 1024 + oo] = lo;
                    }
        for (int hh = 0; hh < H; ++hh) {
            double gdv = 0, cd = r.pd_b[hh], gld = 0, cld = r.pld_b[hh], gtv = 0, ct = 0;
            for (int k = 0; k < H; ++k) {
                gdv += (double)r.ed_k[k] * r.pd_k[k * H + hh];  cd += (double)r.ed_b[k] * r.pd_k[k * H + hh];
                gld += (double)r.eld_k[k] * r.pld_k[k * H + hh]; cld += (double)r.eld_b[k] * r.pld_k[k * H + hh];
            }
            if (h->tw) {
                ct = r.pt_b[hh];
                for (int k = 0; k < H; ++k) { gtv += (double)r.et_k[k] * r.pt_k[k * H + hh]; ct += (double)r.et_b[k] * r.pt_k[k * H + hh]; }
            }
            double ce = r.pr_b[hh];
            for (int e2 = 0; e2 < E; ++e2) ce += (double)emb_b[e2] * r.pr_k[e2 * H + hh];
            for (int cc = 0; cc < 4; ++cc) {
                double s = 0;
                if (cc < D1) for (int e2 = 0; e2 < E; ++e2) s += (double)emb_k[cc * E + e2] * r.pr_k[e2 * H + hh];
                blob[o.Au + cc * H + hh] = (float)s;
                blob[o.A + cc * H + hh] = (float)(s * (double)kTwoLog2e);
                Au_d[a][cc * H + hh] = s;
            }
            blob[o.ce + hh] = (float)ce;
            blob[o.bq + hh] = (float)((double)r.pq_b[hh] + ce + cd + cld + ct);
            ce_d[a][hh] = ce;
            bq_d[a][hh] = (double)r.pq_b[hh] + ce + cd + cld + ct;
            blob[o.gd + hh] = (float)((gdv - gld) * (double)kTwoLog2e);
            blob[o.gl + hh] = (float)gld;
            blob[o.gt + hh] = (float)gtv;
        }
    }
    // glimpse read-out folded into the next module's query: hy = pf . Au + ce, q' = hy . W_q' + b'
    for (int a = 0; a + 1 < natt; ++a) {
        const float* Wn = rh[a + 1].pq_k;   // [k][unit]
        for (int hh = 0; hh < H; ++hh) {
            double c0 = bq_d[a + 1][hh];
            for (int k = 0; k < H; ++k) c0 += ce_d[a][k] * (double)Wn[(size_t)k * H + hh];
            blob[ao[a].cq + hh] = (float)c0;
            for (int cc = 0; cc < 4; ++cc) {
                double m0 = 0;
                for (int k = 0; k < H; ++k) m0 += Au_d[a][cc * H + k] * (double)Wn[(size_t)k * H + hh];
                blob[ao[a].Mq + cc * H + hh] = (float)m0;
            }
        }
    }
    // raw copies for the stand-alone (as-written) kernels
    const size_t o_emb_k = put(emb_k, (size_t)D1 * E), o_emb_b = put(emb_b, E);
    const size_t o_lstm_k = put(lstm_k, (size_t)(E + H) * 4 * H), o_lstm_b = put(lstm_b, 4 * H);
    std::vector<float> lstm_kT((size_t)4 * H * (E + H)), pq_kT((size_t)H * H);   // transposed copies for the back-propagation kernel
    for (int k = 0; k < E + H; ++k)
        for (int j = 0; j < 4 * H; ++j) lstm_kT[(size_t)j * (E + H) + k] = lstm_k[(size_t)k * 4 * H + j];
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < H; ++j) pq_kT[(size_t)j * H + k] = rh[natt - 1].pq_k[(size_t)k * H + j];
    const size_t o_lstm_kT = put(lstm_kT.data(), lstm_kT.size()), o_pq_kT = put(pq_kT.data(), pq_kT.size());
    struct RawOff { size_t v, ed_k, ed_b, eld_k, eld_b, et_k, et_b, pd_k, pd_b, pld_k, pld_b, pt_k, pt_b, pq_k, pq_b, pr_k, pr_b; };
    std::vector<RawOff> ro(natt);
    for (int a = 0; a < natt; ++a) {
        const RawHost& r = rh[a];
        RawOff& o = ro[a];
        o.v = put(r.v, H);
        o.ed_k = put(r.ed_k, H); o.ed_b = put(r.ed_b, H); o.eld_k = put(r.eld_k, H); o.eld_b = put(r.eld_b, H);
        o.pd_k = put(r.pd_k, H * H); o.pd_b = put(r.pd_b, H); o.pld_k = put(r.pld_k, H * H); o.pld_b = put(r.pld_b, H);
        o.pq_k = put(r.pq_k, H * H); o.pq_b = put(r.pq_b, H); o.pr_k = put(r.pr_k, (size_t)E * H); o.pr_b = put(r.pr_b, H);
        if (h->tw) { o.et_k = put(r.et_k, H); o.et_b = put(r.et_b, H); o.pt_k = put(r.pt_k, H * H); o.pt_b = put(r.pt_b, H); }
        else { o.et_k = o.et_b = o.pt_k = o.pt_b = 0; }
    }

    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->d_blob && h->blob_floats < blob.size()) { cudaFree(h->d_blob); h->d_blob = nullptr; }
    if (!h->d_blob) CUDA_TRY(h, cudaMalloc((void**)&h->d_blob, blob.size() * 4));
    h->blob_floats = blob.size();
    CUDA_TRY(h, cudaMemcpy(h->d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    const float* d = h->d_blob;
    h->Wg = d + oWg;
    h->bg = d + obg;
    h->WgT3 = d + oWgT3; h->WxT = d + oWxT; h->bgT = d + obgT; h->WgTh = d + oWgTh;
    for (int a = 0; a < natt; ++a) {
        h->att[a] = AttWeights{d + ao[a].Wq, d + ao[a].bq, d + ao[a].A, d + ao[a].gd, d + ao[a].gl, d + ao[a].gt,
                               d + ao[a].v, d + ao[a].Au, d + ao[a].ce, d + ao[a].Mq, d + ao[a].cq};
        h->WqTc[a] = d + ao[a].WqTc;
        h->WqTh[a] = d + ao[a].WqTh;
        const RawOff& o = ro[a];
        RawAtt& ra = h->raw.att[a];
        ra.v = d + o.v;
        ra.emb_d_k = d + o.ed_k; ra.emb_d_b = d + o.ed_b; ra.emb_ld_k = d + o.eld_k; ra.emb_ld_b = d + o.eld_b;
        ra.emb_t_k = d + o.et_k; ra.emb_t_b = d + o.et_b;
        ra.proj_d_k = d + o.pd_k; ra.proj_d_b = d + o.pd_b; ra.proj_ld_k = d + o.pld_k; ra.proj_ld_b = d + o.pld_b;
        ra.proj_t_k = d + o.pt_k; ra.proj_t_b = d + o.pt_b;
        ra.proj_q_k = d + o.pq_k; ra.proj_q_b = d + o.pq_b; ra.proj_ref_k = d + o.pr_k; ra.proj_ref_b = d + o.pr_b;
    }
    {
        const RawHost& r = rh[natt - 1];
        double sv = 0.0;
        for (int hh = 0; hh < H; ++hh) { h->vm2[hh] = -2.0f * r.v[hh]; sv += (double)r.v[hh]; }
        h->sumv = (float)sv;
    }
    if (c.n_glimpses > 0) {   // device array of the glimpse modules for the fused FP32-pipe kernel
        if (!h->d_gatt) CUDA_TRY(h, cudaMalloc((void**)&h->d_gatt, sizeof(AttWeights) * 8));
        CUDA_TRY(h, cudaMemcpy(h->d_gatt, h->att, sizeof(AttWeights) * c.n_glimpses, cudaMemcpyHostToDevice));
    }

    // ---- optional: the critic (Critic/Process/P<i>/..., Critic/Linear/L1|L2; attention_agent.py:243-270) ------------------
    h->critic_blocks = 0;
    if (m.count("Critic/Process/P0/v")) {
        int nb = 0;
        while (nb < 8 && m.count("Critic/Process/P" + std::to_string(nb) + "/v")) ++nb;
        std::string miss;
        auto getc = [&](const std::string& name, int64_t want) -> const float* {
            auto it = m.find(name);
            if (it == m.end() || it->second.n != want) { miss += " " + name; return nullptr; }
            return it->second.p;
        };
        std::vector<float> cb((size_t)nb * kCriticBlockFloats + kCriticTailFloats, 0.0f);
        std::vector<double> Au_prev(4 * H, 0.0), ce_prev(H, 0.0);
        std::vector<float> craw;                       // raw copies for the back-propagation kernel
        std::vector<size_t> craw_off;                  // 13 offsets per block, then 4 for the dense layers
        auto putc = [&](const float* src, size_t cnt) { craw_off.push_back(craw.size()); if (src) craw.insert(craw.end(), src, src + cnt); };
        for (int i = 0; i < nb; ++i) {
            const std::string p = "Critic/Process/P" + std::to_string(i);
            const float* v = getc(p + "/v", H);
            const float *ed_k = getc(p + "/emb_d/kernel", H), *ed_b = getc(p + "/emb_d/bias", H);
            const float *pd_k = getc(p + "/proj_d/kernel", H * H), *pd_b = getc(p + "/proj_d/bias", H);
            const float *pq_k = getc(p + "/proj_q/kernel", H * H), *pq_b = getc(p + "/proj_q/bias", H);
            const float *pe_k = getc(p + "/proj_e/kernel", (int64_t)E * H), *pe_b = getc(p + "/proj_e/bias", H);
            const float *et_k = nullptr, *et_b = nullptr, *pt_k = nullptr, *pt_b = nullptr;
            if (h->tw) {   // emb_begin_tw embeds and proj_end_tw projects BOTH windows (vrptw_attention.py:147-150)
                et_k = getc(p + "/emb_begin_tw/kernel", H); et_b = getc(p + "/emb_begin_tw/bias", H);
                pt_k = getc(p + "/proj_end_tw/kernel", H * H); pt_b = getc(p + "/proj_end_tw/bias", H);
            }
            if (!miss.empty()) break;
            putc(v, H); putc(ed_k, H); putc(ed_b, H); putc(pd_k, (size_t)H * H); putc(pd_b, H); putc(pq_k, (size_t)H * H); putc(pq_b, H);
            putc(pe_k, (size_t)E * H); putc(pe_b, H); putc(et_k, H); putc(et_b, H); putc(pt_k, (size_t)H * H); putc(pt_b, H);
            float* blk = &cb[(size_t)i * kCriticBlockFloats];
            std::vector<double> Au(4 * H, 0.0), ce(H, 0.0), cst(H, 0.0);
            for (int hh = 0; hh < H; ++hh) {
                double gd = 0, cd = pd_b[hh], gt = 0, ct = 0;
                for (int k = 0; k < H; ++k) { gd += (double)ed_k[k] * pd_k[k * H + hh]; cd += (double)ed_b[k] * pd_k[k * H + hh]; }
                if (h->tw) {
                    ct = pt_b[hh];
                    for (int k = 0; k < H; ++k) { gt += (double)et_k[k] * pt_k[k * H + hh]; ct += (double)et_b[k] * pt_k[k * H + hh]; }
                }
                double c0 = pe_b[hh];
                for (int e2 = 0; e2 < E; ++e2) c0 += (double)emb_b[e2] * pe_k[e2 * H + hh];
                ce[hh] = c0;
                for (int cc = 0; cc < 4; ++cc) {
                    double s = 0;
                    if (cc < D1) for (int e2 = 0; e2 < E; ++e2) s += (double)emb_k[cc * E + e2] * pe_k[e2 * H + hh];
                    Au[cc * H + hh] = s;
                    blk[cc * H + hh] = (float)((s + ((h->tw && cc >= 2) ? gt : 0.0)) * (double)kTwoLog2e);
                }
                blk[4 * H + hh] = (float)(gd * (double)kTwoLog2e);
                cst[hh] = (double)pq_b[hh] + c0 + cd + (h->tw ? 2.0 * ct : 0.0);
                blk[10 * H + hh] = v[hh];
            }
            for (int hh = 0; hh < H; ++hh) {   // query of this block from the previous block's read-out (block 0: hy = 0)
                double c0 = cst[hh];
                if (i > 0) {
                    for (int k = 0; k < H; ++k) c0 += ce_prev[k] * (double)pq_k[(size_t)k * H + hh];
                    for (int cc = 0; cc < 4; ++cc) {
                        double m0 = 0;
                        for (int k = 0; k < H; ++k) m0 += Au_prev[cc * H + k] * (double)pq_k[(size_t)k * H + hh];
                        blk[5 * H + cc * H + hh] = (float)m0;
                    }
                }
                blk[9 * H + hh] = (float)c0;
            }
            Au_prev = Au; ce_prev = ce;
        }
        const float *l1_k = getc("Critic/Linear/L1/kernel", H * H), *l1_b = getc("Critic/Linear/L1/bias", H);
        const float *l2_k = getc("Critic/Linear/L2/kernel", H), *l2_b = getc("Critic/Linear/L2/bias", 1);
        if (!miss.empty()) return fail(h, VRPB200_E_WEIGHTS, "critic: missing or mis-sized variables:%s", miss.c_str());
        float* tail = &cb[(size_t)nb * kCriticBlockFloats];
        for (int hh = 0; hh < H; ++hh) {
            double c0 = l1_b[hh];
            for (int k = 0; k < H; ++k) c0 += ce_prev[k] * (double)l1_k[(size_t)k * H + hh];
            for (int cc = 0; cc < 4; ++cc) {
                double m0 = 0;
                for (int k = 0; k < H; ++k) m0 += Au_prev[cc * H + k] * (double)l1_k[(size_t)k * H + hh];
                tail[cc * H + hh] = (float)m0;
            }
            tail[4 * H + hh] = (float)c0;
            tail[5 * H + hh] = l2_k[hh];
        }
        tail[6 * H] = l2_b[0];
        if (h->d_critic) { cudaFree(h->d_critic); h->d_critic = nullptr; }
        CUDA_TRY(h, cudaMalloc((void**)&h->d_critic, cb.size() * 4));
        CUDA_TRY(h, cudaMemcpy(h->d_critic, cb.data(), cb.size() * 4, cudaMemcpyHostToDevice));
        h->critic_blocks = nb;
        putc(l1_k, (size_t)H * H); putc(l1_b, H); putc(l2_k, H); putc(l2_b, 1);
        if (h->d_critic_raw) { cudaFree(h->d_critic_raw); h->d_critic_raw = nullptr; }
        CUDA_TRY(h, cudaMalloc((void**)&h->d_critic_raw, craw.size() * 4));
        CUDA_TRY(h, cudaMemcpy(h->d_critic_raw, craw.data(), craw.size() * 4, cudaMemcpyHostToDevice));
        CriticGradWeights& cw = h->critic_raw;
        memset(&cw, 0, sizeof cw);
        cw.emb_k = d + o_emb_k; cw.emb_b = d + o_emb_b;
        const float* cr = h->d_critic_raw;
        for (int i = 0; i < nb; ++i) {
            const size_t* o = &craw_off[(size_t)i * 13];
            cw.blk[i] = CriticBlockWeights{cr + o[0], cr + o[1], cr + o[2], cr + o[3], cr + o[4], cr + o[5], cr + o[6], cr + o[7], cr + o[8],
                                           cr + o[9], cr + o[10], cr + o[11], cr + o[12]};
        }
        const size_t* o = &craw_off[(size_t)nb * 13];
        cw.l1_k = cr + o[0]; cw.l1_b = cr + o[1]; cw.l2_k = cr + o[2]; cw.l2_b = cr + o[3];
    }
    h->raw.lstm_k = d + o_lstm_k;
    h->raw.lstm_b = d + o_lstm_b;
    h->raw.n_glimpses = c.n_glimpses;
    h->raw_emb_k = d + o_emb_k;
    h->raw_emb_b = d + o_emb_b;
    h->raw_lstm_kT = d + o_lstm_kT;
    h->raw_pq_kT = d + o_pq_kT;
    h->have_weights = true;
    return VRPB200_OK;
}

}  // extern "C"

namespace {

// gemm_path -> kernel family.  0 picks the default (rollout4 split; beam search on rollout3).
enum Family { kFamFfma = 1, kFamV3 = 3, kFamV4 = 6 };

int launch_rollout(vrpb200_handle* h, RolloutArgs& a, cudaStream_t st) {
    const vrpb200_config& c = h->cfg;
    const int FD = h->tw ? 4 : 2;
    // gemm_path 0: rollout4 for greedy and stochastic decoding, rollout3 for beam search
    const bool automatic = c.gemm_path == 0;
    int fam = automatic ? kFamV4 : c.gemm_path;
    if (c.n_glimpses != 0 || c.physics != 0) fam = kFamFfma;   // the one fused kernel with glimpses / UPS_old physics
    if (fam == kFamV4 && a.mode == 2) fam = kFamV3;
    int stages = 2;
    auto smem_of = [&](int rb) {
        if (fam == kFamV4) return smem_rollout4(rb, c.n_nodes, h->max_smem_optin, &stages);
        if (fam == kFamV3) return smem_rollout3(rb, rb / a.W, c.n_nodes, FD, a.W, h->max_smem_optin, &stages);
        return smem_ffma(rb, rb, c.n_nodes, FD);
    };
    // rows per CTA.  A decode step of a block is a latency chain (LSTM epilogue -> query GEMM -> scores -> per-row tail) whose
    // length hardly depends on the number of rows: measured on B200 (profiles/r2_rows_per_cta_sweep.txt) the time of one wave
    // of CTAs is proportional to (85 + rows_per_cta) for VRPTW-20, -50 and -100 alike.  So the block size that minimises
    // waves x (85 + rows) wins: big blocks for big batches, but e.g. 32 rows for 8 192 or 4 096 instances (2 waves / 1 wave
    // instead of 4 / 2 waves of 16-row blocks: 1.4 - 1.7 x faster).  Beam search keeps the rule "largest block that still
    // fills the device twice, else the smallest that holds every beam of an instance" (rollout3, not re-measured).
    const int cand[4] = {64, 48, 32, 16};
    int RB = 0;
    if (c.rows_per_cta) {
        RB = c.rows_per_cta;
        if (fam == kFamV4 && RB > 48) RB = 48;   // rollout4 owns TMEM columns for q^T and the score buffers: at most 48 rows (a tuning knob only)
    } else if (a.mode != 2) {
        double best = 0.0;
        for (int i = 0; i < 4; ++i) {
            const int rb = cand[i];
            if (smem_of(rb) > h->max_smem_optin) continue;
            const long long ctas = (a.B + rb - 1) / rb;
            const double cost = (double)((ctas + h->num_sms - 1) / h->num_sms) * (85.0 + rb);
            if (!RB || cost < best) { RB = rb; best = cost; }
        }
    } else {
        int smallest_fit = 0;
        for (int i = 0; i < 4 && !RB; ++i) {
            const int rb = cand[i];
            if (smem_of(rb) > h->max_smem_optin) continue;
            smallest_fit = rb;
            if ((a.B * a.W + rb - 1) / rb >= 2LL * h->num_sms) RB = rb;
        }
        if (!RB) RB = smallest_fit;
    }
    if (!RB) return fail(h, VRPB200_E_ARG, "no rows_per_cta fits shared memory (n_nodes %d, beam_width %d)", c.n_nodes, a.W);
    a.IB = RB / a.W;
    const int total = smem_of(RB);
    if (total > h->max_smem_optin)
        return fail(h, VRPB200_E_ARG, "rows_per_cta %d needs %d B shared memory (> %d)", RB, total, h->max_smem_optin);
    const long long grid = (a.B + a.IB - 1) / a.IB;
    if (grid > 0x7fffffffLL) return fail(h, VRPB200_E_ARG, "batch too large");
    h->last_launch[0] = (int)grid;
    h->last_launch[1] = fam == kFamV4 ? kV4Threads : fam == kFamV3 ? kV3Threads : RB * 8;
    h->last_launch[2] = total; h->last_launch[3] = RB; h->last_launch[4] = 1; h->last_launch[5] = fam;
    cudaError_t e;
    if (fam == kFamV4) e = launch_rollout4(a, RB, h->tw, stages, (int)grid, total, st);
    else if (fam == kFamV3) e = launch_rollout3(a, RB, h->tw, stages, (int)grid, total, st);
    else e = launch_ffma(a, RB, h->tw, (int)grid, total, st);
    if (e != cudaSuccess) return fail(h, VRPB200_E_CUDA, "rollout launch (family %d, rows_per_cta %d): %s", fam, RB, cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" {

static int rollout_impl(vrpb200_handle* h, const float* d_input, int64_t B, int mode, int beam_width,
                        const float* d_uniforms, const float* d_uniforms_retry, uint64_t seed, int32_t* d_idx,
                        float* d_cost, float* d_logprob, int32_t* d_beam_parent, const vrpb200_trace* trace,
                        void* stream, long long row_base) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights) return fail(h, VRPB200_E_STATE, "rollout before set_weights");
    if (!d_input || !d_cost || !d_idx) return fail(h, VRPB200_E_ARG, "d_input, d_idx and d_cost are required");
    if (B < 1) return fail(h, VRPB200_E_ARG, "B must be >= 1");
    if (mode != VRPB200_GREEDY && mode != VRPB200_STOCHASTIC && mode != VRPB200_BEAM) return fail(h, VRPB200_E_ARG, "bad mode %d", mode);
    if (mode == VRPB200_BEAM && h->cfg.gemm_path == 1) return fail(h, VRPB200_E_ARG, "beam search is not implemented by the FP32-pipe cross-check kernel");
    if (mode == VRPB200_BEAM && (beam_width < 1 || beam_width > 64)) return fail(h, VRPB200_E_ARG, "beam_width outside [1,64]");
    if (mode == VRPB200_BEAM && beam_width > h->cfg.n_nodes)   // step 0 has n_nodes candidates: tf.nn.top_k(k > n) is an error in the reference too
        return fail(h, VRPB200_E_ARG, "beam_width %d exceeds n_nodes %d", beam_width, h->cfg.n_nodes);
    if ((h->cfg.n_glimpses != 0 || h->cfg.physics != 0) && mode == VRPB200_BEAM)
        return fail(h, VRPB200_E_ARG, "beam search with glimpses or UPS_old physics is not fused (drive decode_step / env_step)");
    if (mode == VRPB200_STOCHASTIC && ((d_uniforms == nullptr) != (d_uniforms_retry == nullptr)))
        return fail(h, VRPB200_E_ARG, "pass both uniform buffers or neither");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const vrpb200_config& c = h->cfg;
    RolloutArgs a;
    memset(&a, 0, sizeof a);
    a.input = d_input; a.Wg = h->Wg; a.bg = h->bg; a.att = h->att[c.n_glimpses];
    a.WqTc = h->WqTc[c.n_glimpses];
    a.WgT3 = h->WgT3; a.WxT = h->WxT; a.bgT = h->bgT;
    a.WqTh = h->WqTh[c.n_glimpses]; a.WgTh = h->WgTh;
    a.uniforms = d_uniforms; a.uniforms_retry = d_uniforms_retry;
    a.idx_out = d_idx; a.cost_out = d_cost; a.logprob_out = d_logprob;
    a.parent_out = mode == VRPB200_BEAM ? d_beam_parent : nullptr;
    if (trace) {
        a.tr_logits = trace->d_logits; a.tr_probs = trace->d_probs; a.tr_masks = trace->d_masks;
        a.fin_demand = trace->d_final_demand; a.fin_load = trace->d_final_load; a.fin_time = trace->d_final_time;
        a.fin_mask = trace->d_final_mask; a.steps_out = trace->d_steps;
        a.stats = (unsigned long long*)trace->d_stats;
    }
    a.full = (a.tr_logits || a.tr_probs || a.tr_masks) ? 1 : 0;
    a.seed = seed; a.B = B; a.row_base = row_base; a.W = mode == VRPB200_BEAM ? beam_width : 1; a.N = c.n_nodes; a.n_cust = c.n_cust; a.D = c.input_dim; a.T = c.decode_len;
    a.mode = mode; a.use_tanh = c.use_tanh; a.mask_pointer = c.mask_pointer;
    a.capacity = c.capacity; a.tanh_C = c.tanh_exploration; a.forget_bias = c.forget_bias;
    a.sumv = h->sumv;
    memcpy(a.vm2, h->vm2, sizeof a.vm2);
    a.gatt = h->d_gatt; a.n_glimpses = c.n_glimpses; a.mask_glimpses = c.mask_glimpses; a.physics = c.physics;
    a.q_Wq = h->att[0].Wq; a.q_bq = h->att[0].bq;   // module 0 = first glimpse, or the pointer when there is none
    const bool fp32_only = c.n_glimpses != 0 || c.physics != 0;   // glimpses and UPS_old physics live in the FP32-pipe kernel
    if ((c.gemm_path == 0 || c.gemm_path >= 6) && mode != VRPB200_BEAM && !fp32_only) {   // rollout4 gathers node features as aligned 16-byte records
        const int slot = h->ws_slot;
        const long long nn = (long long)B * c.n_nodes;
        int rc = ensure(h, &h->ws_feat4[slot], &h->ws_feat4_bytes[slot], (size_t)nn * sizeof(float4));
        if (rc) return rc;
        repack_features_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_input, h->ws_feat4[slot], nn, c.input_dim);
        CUDA_TRY(h, cudaGetLastError());
        a.feat4 = h->ws_feat4[slot];
    }
    if (mode == VRPB200_STOCHASTIC && h->whole_batch_retry) {
        // the reference's my_multinomial (attention_agent.py:148-167): if ANY row of the batch finds no node with its first
        // draw, EVERY row redraws that step once.  The kernel flags such steps; the rollout is repeated with the earliest
        // flagged step redrawn until no new step is flagged (probability ~1e-7 per row-step: one launch almost always).
        const int T = c.decode_len;
        if (!h->d_retry) CUDA_TRY(h, cudaMalloc((void**)&h->d_retry, sizeof(int) * 2 * 65536));
        if (T > 65536) return fail(h, VRPB200_E_ARG, "decode_len too large for whole-batch retry");
        std::vector<int> retry(T, 0), failed(T, 0);
        h->last_redraws = 0;
        a.retry_steps = h->d_retry; a.fail_steps = h->d_retry + T;
        for (;;) {
            CUDA_TRY(h, cudaMemcpyAsync(h->d_retry, retry.data(), sizeof(int) * T, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            CUDA_TRY(h, cudaMemsetAsync(h->d_retry + T, 0, sizeof(int) * T, (cudaStream_t)stream));
            int rc = launch_rollout(h, a, (cudaStream_t)stream);
            if (rc) return rc;
            CUDA_TRY(h, cudaMemcpyAsync(failed.data(), h->d_retry + T, sizeof(int) * T, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
            CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream));
            int first = -1;
            for (int t = 0; t < T && first < 0; ++t) if (failed[t] && !retry[t]) first = t;
            if (first < 0) return 0;
            retry[first] = 1;           // later flags came from a trajectory that this redraw changes: decide them again
            ++h->last_redraws;
        }
    }
    return launch_rollout(h, a, (cudaStream_t)stream);
}

int vrpb200_rollout(vrpb200_handle* h, const float* d_input, int64_t B, int mode, int beam_width,
                    const float* d_uniforms, const float* d_uniforms_retry, uint64_t seed, int32_t* d_idx,
                    float* d_cost, float* d_logprob, int32_t* d_beam_parent, const vrpb200_trace* trace,
                    void* stream) {
    return rollout_impl(h, d_input, B, mode, beam_width, d_uniforms, d_uniforms_retry, seed, d_idx, d_cost, d_logprob,
                        d_beam_parent, trace, stream, 0);
}

int vrpb200_rollout_host(vrpb200_handle* h, const float* h_input, int64_t B, int mode, int beam_width,
                         uint64_t seed, int32_t* h_idx, float* h_cost, uint64_t* h_stats) {
    if (!h) return VRPB200_E_ARG;
    if (!h_input || !h_cost) return fail(h, VRPB200_E_ARG, "h_input and h_cost are required");
    if (B < 1) return fail(h, VRPB200_E_ARG, "B must be >= 1");
    if (mode == VRPB200_BEAM) return fail(h, VRPB200_E_ARG, "rollout_host: beam search goes through vrpb200_rollout");
    if (mode == VRPB200_STOCHASTIC && h->whole_batch_retry) return fail(h, VRPB200_E_ARG, "rollout_host decodes in chunks: whole-batch retry needs vrpb200_rollout");
    const vrpb200_config& c = h->cfg;
    CUDA_TRY(h, cudaSetDevice(h->device));
    constexpr int NC = vrpb200_handle::kChunks;
    for (int i = 0; i < NC; ++i)
        if (!h->streams[i]) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    if (!h->dev_stats) CUDA_TRY(h, cudaMalloc((void**)&h->dev_stats, 4 * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemsetAsync(h->dev_stats, 0, 4 * sizeof(unsigned long long), h->streams[0]));
    CUDA_TRY(h, cudaStreamSynchronize(h->streams[0]));
    // chunk = 4 waves of 64-row blocks: large enough to amortise launch + tail, small enough that the first
    // H2D and the last D2H (the only un-overlapped transfers) are short
    const int64_t chunk = std::min<int64_t>(B, (int64_t)h->num_sms * 64 * 4);
    const int T = c.decode_len;
    const size_t inst_b = (size_t)c.n_nodes * c.input_dim * 4;
    int rc;
    vrpb200_trace tr;
    memset(&tr, 0, sizeof tr);
    tr.d_stats = (uint64_t*)h->dev_stats;
    int launches = 0;
    for (int64_t off = 0, k = 0; off < B; off += chunk, ++k) {
        const int s = (int)(k % NC);
        const int64_t nb = std::min<int64_t>(chunk, B - off);
        if ((rc = ensure(h, &h->dev_in[s], &h->dev_in_bytes[s], (size_t)chunk * inst_b))) return rc;
        if ((rc = ensure(h, &h->dev_idx[s], &h->dev_idx_bytes[s], (size_t)chunk * T * 4))) return rc;
        if ((rc = ensure(h, &h->dev_cost[s], &h->dev_cost_bytes[s], (size_t)chunk * 4))) return rc;
        cudaStream_t st = h->streams[s];   // stream order protects the staging buffers of slot s
        CUDA_TRY(h, cudaMemcpyAsync(h->dev_in[s], (const char*)h_input + (size_t)off * inst_b, (size_t)nb * inst_b,
                                    cudaMemcpyHostToDevice, st));
        h->ws_slot = s;
        rc = rollout_impl(h, h->dev_in[s], nb, mode, beam_width, nullptr, nullptr, seed, h->dev_idx[s], h->dev_cost[s],
                          nullptr, nullptr, &tr, st, off);
        h->ws_slot = vrpb200_handle::kChunks;
        if (rc) return rc;
        ++launches;
        CUDA_TRY(h, cudaMemcpyAsync(h_cost + off, h->dev_cost[s], (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        if (h_idx)   // [T, nb] device -> columns [off, off+nb) of the host's [T, B]
            CUDA_TRY(h, cudaMemcpy2DAsync(h_idx + off, (size_t)B * 4, h->dev_idx[s], (size_t)nb * 4, (size_t)nb * 4, T,
                                          cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < NC; ++i) CUDA_TRY(h, cudaStreamSynchronize(h->streams[i]));
    if (h_stats) CUDA_TRY(h, cudaMemcpy(h_stats, h->dev_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    h->last_launch[4] = launches;
    return VRPB200_OK;
}

int vrpb200_pipe_peak(vrpb200_handle* h, int kind, int iters, double* ops_per_second) {
    if (!h || !ops_per_second || (kind != 0 && kind != 1) || iters < 1) return fail(h, VRPB200_E_ARG, "pipe_peak: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    float* sink = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&sink, 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(h, cudaEventCreate(&e0));
    CUDA_TRY(h, cudaEventCreate(&e1));
    const int grid = h->num_sms * 4, block = 512;
    pipe_peak_kernel<<<grid, block>>>(kind, 8, sink);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(h, cudaEventRecord(e0));
        pipe_peak_kernel<<<grid, block>>>(kind, iters, sink);
        CUDA_TRY(h, cudaEventRecord(e1));
        CUDA_TRY(h, cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    const double per_thread = (double)iters * 32.0 * (kind == 0 ? 16.0 : 16.0);
    *ops_per_second = per_thread * grid * block / (best * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return VRPB200_OK;
}

int vrpb200_env_reset(vrpb200_handle* h, const float* d_input, int64_t B, int beam_width, float* d_demand,
                      float* d_load, float* d_time, float* d_mask, float* d_prev_xy, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_input || !d_demand || !d_load || !d_mask || B < 1 || beam_width < 1) return fail(h, VRPB200_E_ARG, "env_reset: bad argument");
    if (h->tw && (!d_time || !d_prev_xy)) return fail(h, VRPB200_E_ARG, "env_reset: VRPTW needs d_time and d_prev_xy");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long R = B * beam_width;
    const int wpb = 8;
    env_reset_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        d_input, B, beam_width, h->cfg.n_nodes, h->cfg.input_dim, h->cfg.capacity, d_demand, d_load, d_time, d_mask, d_prev_xy);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_env_step(vrpb200_handle* h, const float* d_input, int64_t B, int beam_width, const int64_t* d_idx,
                     const int64_t* d_beam_parent, const float* d_demand_in, const float* d_load_in,
                     const float* d_time_in, const float* d_prev_xy_in, float* d_demand, float* d_load, float* d_time,
                     float* d_mask, float* d_prev_xy, float* d_sat, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_input || !d_idx || !d_demand_in || !d_load_in || !d_demand || !d_load || !d_mask || B < 1 || beam_width < 1)
        return fail(h, VRPB200_E_ARG, "env_step: bad argument");
    if (h->tw && (!d_time_in || !d_prev_xy_in || !d_time || !d_prev_xy)) return fail(h, VRPB200_E_ARG, "env_step: VRPTW needs time and prev_xy");
    if (d_beam_parent && (d_demand_in == d_demand || d_load_in == d_load))
        return fail(h, VRPB200_E_ARG, "env_step: with beam_parent the in and out state must not alias");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long R = B * beam_width;
    const int wpb = 8;
    const unsigned grid = (unsigned)((R + wpb - 1) / wpb);
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    if (h->tw)
        env_step_kernel<true><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(
            d_input, B, beam_width, h->cfg.n_nodes, h->cfg.input_dim, h->cfg.capacity, (const long long*)d_idx,
            (const long long*)d_beam_parent, d_demand_in, d_load_in, d_time_in, d_prev_xy_in, d_demand, d_load, d_time,
            d_mask, d_prev_xy, d_sat, h->cfg.physics);
    else
        env_step_kernel<false><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(
            d_input, B, beam_width, h->cfg.n_nodes, h->cfg.input_dim, h->cfg.capacity, (const long long*)d_idx,
            (const long long*)d_beam_parent, d_demand_in, d_load_in, d_time_in, d_prev_xy_in, d_demand, d_load, d_time,
            d_mask, d_prev_xy, d_sat, 0);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_embed(vrpb200_handle* h, const float* d_input_pnt, int64_t M, float* d_emb, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights) return fail(h, VRPB200_E_STATE, "embed before set_weights");
    if (!d_input_pnt || !d_emb || M < 1) return fail(h, VRPB200_E_ARG, "embed: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const long long tot = M * h->cfg.embedding_dim;
    embed_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_input_pnt, M, h->D1, h->cfg.embedding_dim,
                                                                                 h->raw_emb_k, h->raw_emb_b, d_emb);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_attention(vrpb200_handle* h, int att, int64_t R, const float* d_query, const float* d_ref,
                      const float* d_demand, const float* d_load, const float* d_time, float* d_e, float* d_logits,
                      void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights) return fail(h, VRPB200_E_STATE, "attention before set_weights");
    if (!d_query || !d_ref || !d_demand || !d_load || !d_logits || R < 1) return fail(h, VRPB200_E_ARG, "attention: bad argument");
    if (h->tw && !d_time) return fail(h, VRPB200_E_ARG, "attention: VRPTW needs d_time");
    if (att < -1 || att >= h->cfg.n_glimpses) return fail(h, VRPB200_E_ARG, "attention: module %d out of range", att);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool pointer = att < 0;
    const RawAtt& w = h->raw.att[pointer ? h->cfg.n_glimpses : att];
    const int use_tanh = pointer ? h->cfg.use_tanh : 0;   // glimpses are built with use_tanh=False, decode_step.py:41-46
    if (h->tw)
        attention_kernel<true><<<(unsigned)R, kH, 0, (cudaStream_t)stream>>>(w, h->cfg.n_nodes, h->cfg.embedding_dim, d_query, d_ref,
                                                                             d_demand, d_load, d_time, use_tanh,
                                                                             h->cfg.tanh_exploration, d_e, d_logits);
    else
        attention_kernel<false><<<(unsigned)R, kH, 0, (cudaStream_t)stream>>>(w, h->cfg.n_nodes, h->cfg.embedding_dim, d_query, d_ref,
                                                                              d_demand, d_load, d_time, use_tanh,
                                                                              h->cfg.tanh_exploration, d_e, d_logits);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_decode_step(vrpb200_handle* h, int64_t R, const float* d_decoder_inp, const float* d_context,
                        const float* d_demand, const float* d_load, const float* d_time, const float* d_mask,
                        float* d_c, float* d_h, float* d_logit, float* d_prob, float* d_logprob, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights) return fail(h, VRPB200_E_STATE, "decode_step before set_weights");
    if (!d_decoder_inp || !d_context || !d_demand || !d_load || !d_mask || !d_c || !d_h || !d_logit || !d_prob || !d_logprob || R < 1)
        return fail(h, VRPB200_E_ARG, "decode_step: bad argument");
    if (h->tw && !d_time) return fail(h, VRPB200_E_ARG, "decode_step: VRPTW needs d_time");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const vrpb200_config& c = h->cfg;
    if (h->tw)
        decode_step_kernel<true><<<(unsigned)R, kH, 0, (cudaStream_t)stream>>>(
            h->raw, c.n_nodes, c.embedding_dim, d_decoder_inp, d_context, d_demand, d_load, d_time, d_mask, d_c, d_h,
            c.use_tanh, c.tanh_exploration, c.mask_glimpses, c.mask_pointer, c.forget_bias, d_logit, d_prob, d_logprob);
    else
        decode_step_kernel<false><<<(unsigned)R, kH, 0, (cudaStream_t)stream>>>(
            h->raw, c.n_nodes, c.embedding_dim, d_decoder_inp, d_context, d_demand, d_load, d_time, d_mask, d_c, d_h,
            c.use_tanh, c.tanh_exploration, c.mask_glimpses, c.mask_pointer, c.forget_bias, d_logit, d_prob, d_logprob);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_set_sampling_retry(vrpb200_handle* h, int whole_batch) {
    if (!h) return VRPB200_E_ARG;
    h->whole_batch_retry = whole_batch ? 1 : 0;
    return VRPB200_OK;
}

int vrpb200_last_redraws(vrpb200_handle* h) { return h ? h->last_redraws : VRPB200_E_ARG; }

int vrpb200_critic(vrpb200_handle* h, const float* d_input, int64_t B, float* d_v, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights || h->critic_blocks == 0) return fail(h, VRPB200_E_STATE, "critic: set_weights has not seen the Critic/* variables");
    if (!d_input || !d_v || B < 1) return fail(h, VRPB200_E_ARG, "critic: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int wpb = 8, NP = align_up(h->cfg.n_nodes, 8);
    const size_t smem = (size_t)wpb * 6 * NP * sizeof(float);
    const unsigned grid = (unsigned)((B + wpb - 1) / wpb);
    if (h->tw) critic_kernel<true><<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(d_input, B, h->cfg.n_nodes, h->cfg.input_dim, h->critic_blocks, h->d_critic, d_v);
    else critic_kernel<false><<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(d_input, B, h->cfg.n_nodes, h->cfg.input_dim, h->critic_blocks, h->d_critic, d_v);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_reinforce_loss(vrpb200_handle* h, const float* d_R, const float* d_v, const float* d_logprob, int64_t T, int64_t R,
                           float* d_losses, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_R || !d_v || !d_logprob || !d_losses || T < 1 || R < 1) return fail(h, VRPB200_E_ARG, "reinforce_loss: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    reinforce_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_R, d_v, d_logprob, T, R, d_losses);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_reward(vrpb200_handle* h, const float* d_actions, int64_t T, int64_t R, int A, float* d_cost, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_actions || !d_cost || T < 1 || R < 1 || A < 2) return fail(h, VRPB200_E_ARG, "reward: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    reward_kernel<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_actions, T, R, A, d_cost, h->cfg.physics);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_reward_min_trucks(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R, int A,
                              float n_nodes, float* d_reward, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_actions || !d_depot || !d_reward || T < 1 || R < 1 || A < 2 || A > 4 || !(n_nodes > 0.0f))
        return fail(h, VRPB200_E_ARG, "reward_min_trucks: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int XY = A > 2 ? 2 : A;   // VRPTW slices the actions to x, y for the route length; VRP actions are x, y already
    reward_min_trucks_kernel<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_actions, d_depot, T, R, A, XY,
                                                                                           (float)(1.0 / (double)n_nodes), 0, d_reward);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_reward_min_trucks_ups(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R, int A,
                                  float* d_reward, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_actions || !d_depot || !d_reward || T < 1 || R < 1 || A < 2 || A > 4) return fail(h, VRPB200_E_ARG, "reward_min_trucks_ups: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    reward_min_trucks_kernel<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_actions, d_depot, T, R, A, 2, 0.0f, 1, d_reward);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_reward_min_trucks_ups_old(vrpb200_handle* h, const float* d_actions, const float* d_depot, int64_t T, int64_t R, int A,
                                      float n_nodes, float* d_reward, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!d_actions || !d_depot || !d_reward || T < 1 || R < 1 || A < 2 || A > 4 || !(n_nodes > 0.0f))
        return fail(h, VRPB200_E_ARG, "reward_min_trucks_ups_old: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    reward_min_trucks_kernel<<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_actions, d_depot, T, R, A, 2,
                                                                                           (float)(1.0 / (double)n_nodes), 2, d_reward);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

// ---- the backward / update side of build_train_step ---------------------------------------------------------
namespace {
int train_workspace(vrpb200_handle* h, size_t bytes) {
    if (h->train_ws_bytes >= bytes) return VRPB200_OK;
    if (h->d_train_ws) { cudaFree(h->d_train_ws); h->d_train_ws = nullptr; h->train_ws_bytes = 0; }
    CUDA_TRY(h, cudaMalloc((void**)&h->d_train_ws, bytes));
    h->train_ws_bytes = bytes;
    return VRPB200_OK;
}
}  // namespace

int vrpb200_actor_grad(vrpb200_handle* h, const float* d_input, int64_t B, const int32_t* d_idx, const float* d_weight,
                       const float* d_input_scale, int n,
                       const char* const* names, float* const* d_grads, const int64_t* n_elem, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights) return fail(h, VRPB200_E_STATE, "actor_grad before set_weights");
    if (!d_input || !d_idx || !d_weight || B < 1 || n < 0 || (n > 0 && (!names || !d_grads || !n_elem)))
        return fail(h, VRPB200_E_ARG, "actor_grad: bad argument");
    const vrpb200_config& c = h->cfg;
    if (c.n_glimpses != 0) return fail(h, VRPB200_E_ARG, "actor_grad: glimpses are not differentiated (n_glimpses must be 0)");
    const int N = c.n_nodes, E = c.embedding_dim, D = c.input_dim, T = c.decode_len, H = kH, D1 = h->D1;
    const long long R = B;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(h, cudaSetDevice(h->device));

    // expected variables: name -> (slot, elements)
    const std::string att = "Actor/Decoder/Attention/", lstm = "Actor/Decoder/LSTM/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/";
    enum { G_EMB_K, G_EMB_B, G_LSTM_K, G_LSTM_B, G_V, G_PQ_K, G_PQ_B, G_PR_K, G_PR_B, G_EK0, G_EB0 = G_EK0 + 3, G_PK0 = G_EB0 + 3,
           G_PB0 = G_PK0 + 3, G_COUNT = G_PB0 + 3 };
    std::map<std::string, std::pair<int, int64_t>> known = {
        {"Actor/Embedding/conv1d/kernel", {G_EMB_K, (int64_t)D1 * E}}, {"Actor/Embedding/conv1d/bias", {G_EMB_B, E}},
        {lstm + "kernel", {G_LSTM_K, (int64_t)(E + H) * 4 * H}}, {lstm + "bias", {G_LSTM_B, 4 * H}},
        {att + "v", {G_V, H}}, {att + "proj_q/kernel", {G_PQ_K, (int64_t)H * H}}, {att + "proj_q/bias", {G_PQ_B, H}},
        {att + "proj_ref/kernel", {G_PR_K, (int64_t)E * H}}, {att + "proj_ref/bias", {G_PR_B, H}}};
    const char* eb[3] = {"emb_d", "emb_ld", "emb_time"};
    const char* pb[3] = {"proj_d", "proj_ld", "proj_t"};
    for (int k = 0; k < (h->tw ? 3 : 2); ++k) {
        known[att + eb[k] + "/kernel"] = {G_EK0 + k, H};
        known[att + eb[k] + "/bias"] = {G_EB0 + k, H};
        known[att + pb[k] + "/kernel"] = {G_PK0 + k, (int64_t)H * H};
        known[att + pb[k] + "/bias"] = {G_PB0 + k, H};
    }
    float* out[G_COUNT] = {nullptr};
    for (int i = 0; i < n; ++i) {
        auto it = known.find(names[i] ? names[i] : "");
        if (it == known.end()) return fail(h, VRPB200_E_ARG, "actor_grad: %s is not a variable of the differentiated graph", names[i] ? names[i] : "(null)");
        if (n_elem[i] != it->second.second || !d_grads[i])
            return fail(h, VRPB200_E_ARG, "actor_grad: %s has %lld elements, expected %lld", names[i], (long long)n_elem[i], (long long)it->second.second);
        out[it->second.first] = d_grads[i];
    }

    // work arrays
    size_t off = 0;
    auto take = [&](size_t floats) { const size_t o = off; off += (floats * 4 + 255) / 256 * 256; return o; };
    const size_t o_idx64 = take((size_t)T * R * 2);
    const size_t o_dem = take((size_t)(T + 1) * R * N), o_msk = take((size_t)(T + 1) * R * N), o_load = take((size_t)(T + 1) * R);
    const size_t o_time = take((size_t)(T + 1) * R), o_prev = take((size_t)(T + 1) * R * 2);
    const size_t o_hs = take((size_t)R * (T + 1) * H), o_cs = take((size_t)R * (T + 1) * H), o_act = take((size_t)R * T * 4 * H);
    const size_t o_xin = take((size_t)R * T * E), o_dx = take((size_t)R * T * E), o_emb = take((size_t)R * N * E);
    const size_t o_eref = take((size_t)R * N * H), o_erefT = take((size_t)R * H * 128), o_dg = take((size_t)R * T * 4 * H);
    const size_t o_dq = take((size_t)R * T * H), o_Ds = take((size_t)R * N * H), o_demb = take((size_t)R * N * E), o_sums = take((size_t)R * 5 * H);
    const int S = (int)std::min<long long>(R, 32);                  // splits of the device-wide reductions (fixed, so the order is)
    const int rps = (int)((R + S - 1) / S);
    const size_t o_part = take((size_t)S * H * 4 * H);
    if (int rc = train_workspace(h, off)) return rc;
    char* ws = h->d_train_ws;
    auto F = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
    long long* idx64 = reinterpret_cast<long long*>(ws + o_idx64);

    // Env state seen by every step: Env.reset, then Env.step along the given tour (the tested stand-alone kernels)
    idx_to_i64_kernel<<<(unsigned)(((long long)T * R + 255) / 256), 256, 0, st>>>(d_idx, idx64, (long long)T * R, N);
    const int wpb = 8;
    const unsigned egrid = (unsigned)((R + wpb - 1) / wpb);
    env_reset_kernel<<<egrid, wpb * 32, 0, st>>>(d_input, R, 1, N, D, c.capacity, F(o_dem), F(o_load), h->tw ? F(o_time) : nullptr, F(o_msk),
                                                 h->tw ? F(o_prev) : nullptr);
    for (int t = 0; t < T; ++t) {
        const size_t a = (size_t)t * R, b = (size_t)(t + 1) * R;
        if (h->tw)
            env_step_kernel<true><<<egrid, wpb * 32, 0, st>>>(d_input, R, 1, N, D, c.capacity, idx64 + a, nullptr, F(o_dem) + a * N,
                                                              F(o_load) + a, F(o_time) + a, F(o_prev) + a * 2, F(o_dem) + b * N, F(o_load) + b,
                                                              F(o_time) + b, F(o_msk) + b * N, F(o_prev) + b * 2, nullptr, c.physics);
        else
            env_step_kernel<false><<<egrid, wpb * 32, 0, st>>>(d_input, R, 1, N, D, c.capacity, idx64 + a, nullptr, F(o_dem) + a * N,
                                                               F(o_load) + a, nullptr, nullptr, F(o_dem) + b * N, F(o_load) + b, nullptr,
                                                               F(o_msk) + b * N, nullptr, nullptr, 0);
    }
    CUDA_TRY(h, cudaGetLastError());

    const RawAtt& ra = h->raw.att[0];   // n_glimpses == 0: module 0 is the pointer
    ActorGradWeights w{h->raw_emb_k, h->raw_emb_b, h->raw.lstm_k, h->raw.lstm_b, ra.v,
                       ra.emb_d_k, ra.emb_d_b, ra.proj_d_k, ra.proj_d_b, ra.emb_ld_k, ra.emb_ld_b, ra.proj_ld_k, ra.proj_ld_b,
                       ra.emb_t_k, ra.emb_t_b, ra.proj_t_k, ra.proj_t_b, ra.proj_q_k, ra.proj_q_b, ra.proj_ref_k, ra.proj_ref_b, h->raw_lstm_kT, h->raw_pq_kT};
    ActorGradDims dm{N, E, D, T, h->tw ? 1 : 0, c.use_tanh, c.mask_pointer, c.tanh_exploration, c.forget_bias};
    ActorGradBatch bt{d_input, d_idx, d_weight, d_input_scale, F(o_dem), F(o_msk), F(o_load), h->tw ? F(o_time) : nullptr,
                      F(o_hs), F(o_cs), F(o_act), F(o_xin), F(o_dx), F(o_emb), F(o_eref), F(o_erefT), F(o_dg), F(o_dq), F(o_Ds), F(o_demb),
                      F(o_sums), R};
    actor_grad_rows_kernel<<<(unsigned)R, 128, 0, st>>>(w, dm, bt);
    CUDA_TRY(h, cudaGetLastError());

    // device-wide reductions, one fixed order
    auto outer = [&](const float* A, long long sAr, long long sAt, const float* Bm, long long sBr, long long sBt, int Tn, int K, int J, float* o) {
        if (!o) return;
        const long long nel = (long long)K * J;
        outer_reduce_kernel<<<dim3((unsigned)((nel + 127) / 128), (unsigned)S), 128, 0, st>>>(A, sAr, sAt, Bm, sBr, sBt, R, rps, Tn, K, J, F(o_part));
        sum_splits_kernel<<<(unsigned)((nel + 127) / 128), 128, 0, st>>>(F(o_part), S, nel, o);
    };
    auto cols = [&](const float* Bm, long long sBr, long long sBt, int Tn, int J, float* o) {
        if (!o) return;
        col_reduce_kernel<<<dim3((unsigned)((J + 127) / 128), (unsigned)S), 128, 0, st>>>(Bm, sBr, sBt, R, rps, Tn, J, F(o_part));
        sum_splits_kernel<<<(unsigned)((J + 127) / 128), 128, 0, st>>>(F(o_part), S, J, o);
    };
    if (out[G_LSTM_K]) {   // rows 0..E-1: the inputs x_t; rows E..E+H-1: the previous outputs h_{t-1}
        outer(F(o_xin), (long long)T * E, E, F(o_dg), (long long)T * 4 * H, 4 * H, T, E, 4 * H, out[G_LSTM_K]);
        outer(F(o_hs), (long long)(T + 1) * H, H, F(o_dg), (long long)T * 4 * H, 4 * H, T, H, 4 * H, out[G_LSTM_K] + (size_t)E * 4 * H);
    }
    cols(F(o_dg), (long long)T * 4 * H, 4 * H, T, 4 * H, out[G_LSTM_B]);
    outer(F(o_hs) + H, (long long)(T + 1) * H, H, F(o_dq), (long long)T * H, H, T, H, H, out[G_PQ_K]);
    outer(F(o_emb), (long long)N * E, E, F(o_Ds), (long long)N * H, H, N, E, H, out[G_PR_K]);
    outer(d_input, (long long)N * D, D, F(o_demb), (long long)N * E, E, N, D1, E, out[G_EMB_K]);
    cols(F(o_demb), (long long)N * E, E, N, E, out[G_EMB_B]);
    SmallGradOut so{out[G_V], out[G_PQ_B], out[G_PR_B],
                    {out[G_EK0], out[G_EK0 + 1], out[G_EK0 + 2]}, {out[G_EB0], out[G_EB0 + 1], out[G_EB0 + 2]},
                    {out[G_PK0], out[G_PK0 + 1], out[G_PK0 + 2]}, {out[G_PB0], out[G_PB0 + 1], out[G_PB0 + 2]}};
    attention_small_grads_kernel<<<1, 128, 0, st>>>(w, F(o_sums), R, h->tw ? 1 : 0, so);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_critic_grad(vrpb200_handle* h, const float* d_input, int64_t B, const float* d_R, int n, const char* const* names,
                        float* const* d_grads, const int64_t* n_elem, float* d_v, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (!h->have_weights || h->critic_blocks == 0) return fail(h, VRPB200_E_STATE, "critic_grad: set_weights has not seen the Critic/* variables");
    if (!d_input || !d_R || B < 1 || n < 0 || (n > 0 && (!names || !d_grads || !n_elem))) return fail(h, VRPB200_E_ARG, "critic_grad: bad argument");
    const vrpb200_config& c = h->cfg;
    const int N = c.n_nodes, E = c.embedding_dim, D = c.input_dim, H = kH, P = h->critic_blocks;
    const long long R = B;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(h, cudaSetDevice(h->device));
    enum { C_V, C_EDK, C_EDB, C_PDK, C_PDB, C_PQK, C_PQB, C_PEK, C_PEB, C_ETK, C_ETB, C_PTK, C_PTB, C_PER_BLOCK };
    const char* bn[C_PER_BLOCK] = {"v", "emb_d/kernel", "emb_d/bias", "proj_d/kernel", "proj_d/bias", "proj_q/kernel", "proj_q/bias",
                                   "proj_e/kernel", "proj_e/bias", "emb_begin_tw/kernel", "emb_begin_tw/bias", "proj_end_tw/kernel",
                                   "proj_end_tw/bias"};
    const int64_t bsz[C_PER_BLOCK] = {H, H, H, (int64_t)H * H, H, (int64_t)H * H, H, (int64_t)E * H, H, H, H, (int64_t)H * H, H};
    std::map<std::string, std::pair<int, int64_t>> known;
    for (int b = 0; b < P; ++b)
        for (int k = 0; k < (h->tw ? C_PER_BLOCK : C_ETK); ++k)
            known["Critic/Process/P" + std::to_string(b) + "/" + bn[k]] = {b * C_PER_BLOCK + k, bsz[k]};
    const int L0 = P * C_PER_BLOCK;
    known["Critic/Linear/L1/kernel"] = {L0, (int64_t)H * H}; known["Critic/Linear/L1/bias"] = {L0 + 1, H};
    known["Critic/Linear/L2/kernel"] = {L0 + 2, H}; known["Critic/Linear/L2/bias"] = {L0 + 3, 1};
    std::vector<float*> out((size_t)L0 + 4, nullptr);
    for (int i = 0; i < n; ++i) {
        auto it = known.find(names[i] ? names[i] : "");
        if (it == known.end()) return fail(h, VRPB200_E_ARG, "critic_grad: %s is not a variable of the critic", names[i] ? names[i] : "(null)");
        if (n_elem[i] != it->second.second || !d_grads[i])
            return fail(h, VRPB200_E_ARG, "critic_grad: %s has %lld elements, expected %lld", names[i], (long long)n_elem[i], (long long)it->second.second);
        out[it->second.first] = d_grads[i];
    }
    size_t off = 0;
    auto take = [&](size_t floats) { const size_t o = off; off += (floats * 4 + 255) / 256 * 256; return o; };
    const size_t o_emb = take((size_t)R * N * E), o_ee = take((size_t)R * P * N * H), o_eeT = take((size_t)R * P * H * 128), o_prob = take((size_t)R * P * 128);
    const size_t o_hy = take((size_t)R * (P + 1) * H), o_De = take((size_t)R * P * N * H), o_sums = take((size_t)R * P * 4 * H), o_dq = take((size_t)R * P * H);
    const size_t o_l1 = take((size_t)R * H), o_dl1 = take((size_t)R * H), o_vdv = take((size_t)R * 2);
    const int S = (int)std::min<long long>(R, 32);
    const int rps = (int)((R + S - 1) / S);
    const size_t o_part = take((size_t)S * H * H);
    if (int rc = train_workspace(h, off)) return rc;
    char* ws = h->d_train_ws;
    auto F = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
    CriticGradDims dm{N, E, D, P, h->tw ? 1 : 0, 2.0f / (float)R};
    CriticGradBatch bt{d_input, d_R, F(o_emb), F(o_ee), F(o_eeT), F(o_prob), F(o_hy), F(o_De), F(o_sums), F(o_dq), F(o_l1), F(o_dl1), F(o_vdv)};
    critic_grad_rows_kernel<<<(unsigned)R, 128, 0, st>>>(h->critic_raw, dm, bt);
    CUDA_TRY(h, cudaGetLastError());
    auto outer = [&](const float* A, long long sAr, long long sAt, const float* Bm, long long sBr, long long sBt, int Tn, int K, int J, float* o) {
        if (!o) return;
        const long long nel = (long long)K * J;
        outer_reduce_kernel<<<dim3((unsigned)((nel + 127) / 128), (unsigned)S), 128, 0, st>>>(A, sAr, sAt, Bm, sBr, sBt, R, rps, Tn, K, J, F(o_part));
        sum_splits_kernel<<<(unsigned)((nel + 127) / 128), 128, 0, st>>>(F(o_part), S, nel, o);
    };
    auto cols = [&](const float* Bm, long long sBr, long long sBt, int Tn, int J, float* o) {
        if (!o) return;
        col_reduce_kernel<<<dim3((unsigned)((J + 127) / 128), (unsigned)S), 128, 0, st>>>(Bm, sBr, sBt, R, rps, Tn, J, F(o_part));
        sum_splits_kernel<<<(unsigned)((J + 127) / 128), 128, 0, st>>>(F(o_part), S, J, o);
    };
    for (int b = 0; b < P; ++b) {
        float** ob = &out[(size_t)b * C_PER_BLOCK];
        const float* De = F(o_De) + (size_t)b * N * H;
        outer(F(o_emb), (long long)N * E, E, De, (long long)P * N * H, H, N, E, H, ob[C_PEK]);
        cols(De, (long long)P * N * H, H, N, H, ob[C_PEB]);
        outer(F(o_hy) + (size_t)b * H, (long long)(P + 1) * H, 0, F(o_dq) + (size_t)b * H, (long long)P * H, 0, 1, H, H, ob[C_PQK]);
        CriticSmallGradOut so{ob[C_V], ob[C_PQB], ob[C_EDK], ob[C_EDB], ob[C_PDK], ob[C_PDB], ob[C_ETK], ob[C_ETB], ob[C_PTK], ob[C_PTB]};
        critic_small_grads_kernel<<<1, 128, 0, st>>>(h->critic_raw.blk[b], F(o_sums), R, P, b, h->tw ? 1 : 0, so);
    }
    outer(F(o_hy) + (size_t)P * H, (long long)(P + 1) * H, 0, F(o_dl1), H, 0, 1, H, H, out[L0]);
    cols(F(o_dl1), H, 0, 1, H, out[L0 + 1]);
    outer(F(o_l1), H, 0, F(o_vdv) + 1, 2, 0, 1, H, 1, out[L0 + 2]);
    cols(F(o_vdv) + 1, 2, 0, 1, 1, out[L0 + 3]);
    if (d_v) CUDA_TRY(h, cudaMemcpy2DAsync(d_v, 4, F(o_vdv), 8, 4, (size_t)R, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_clip_adam(vrpb200_handle* h, int n, float* const* d_vars, const float* const* d_grads, float* const* d_m, float* const* d_v,
                      const int64_t* n_elem, float max_norm, float lr, int64_t step, float beta1, float beta2, float eps, void* stream) {
    if (!h) return VRPB200_E_ARG;
    if (n < 1 || !d_vars || !d_grads || !d_m || !d_v || !n_elem || step < 1 || !(max_norm > 0.0f))
        return fail(h, VRPB200_E_ARG, "clip_adam: bad argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<AdamVar> hv(n);
    for (int i = 0; i < n; ++i) {
        if (!d_vars[i] || !d_grads[i] || !d_m[i] || !d_v[i] || n_elem[i] < 1) return fail(h, VRPB200_E_ARG, "clip_adam: variable %d: bad argument", i);
        hv[i] = AdamVar{d_vars[i], d_grads[i], d_m[i], d_v[i], (long long)n_elem[i]};
    }
    if (h->adam_vars_cap < n) {   // the variable table lives in a small buffer of the handle
        if (h->d_adam_vars) { CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream)); cudaFree(h->d_adam_vars); h->d_adam_vars = nullptr; }
        CUDA_TRY(h, cudaMalloc(&h->d_adam_vars, sizeof(AdamVar) * (size_t)n));
        h->adam_vars_cap = n;
    }
    AdamVar* dv = static_cast<AdamVar*>(h->d_adam_vars);
    // pageable source: the runtime stages it before the call returns, so hv may go out of scope; stream order protects dv
    CUDA_TRY(h, cudaMemcpyAsync(dv, hv.data(), sizeof(AdamVar) * n, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) / (1.0 - std::pow((double)beta1, (double)step));
    clip_adam_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(dv, max_norm, (float)lr_t, beta1, beta2, eps);
    CUDA_TRY(h, cudaGetLastError());
    return VRPB200_OK;
}

int vrpb200_last_launch(const vrpb200_handle* h, int32_t out[8]) {
    if (!h || !out) return VRPB200_E_ARG;
    memcpy(out, h->last_launch, sizeof h->last_launch);
    return VRPB200_OK;
}

}  // extern "C"
