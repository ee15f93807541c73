// C ABI of libertdiff_b200.so -- see include/ertdiff_b200.h for the contract.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "model.cuh"
#include "denoiser.cuh"

namespace ertdiff {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
std::atomic<int64_t> g_launches{0};

static int check_model(const ertdiff_model* m, bool need_weights = true) {
    if (!m) return fail(ERTDIFF_ERR_ARG, "model handle is NULL");
    if (need_weights && !m->loaded) return fail(ERTDIFF_ERR_STATE, "model weights not loaded");
    return 0;
}

// ---- chain ------------------------------------------------------------------------------
static int run_chain(ertdiff_model* m, const ertdiff_chain_args* a, const float* d_cond_bias,
                     cudaStream_t st) {
    ERT_REQUIRE(a->B > 0 && a->n_cond > 0, "sample_chain: B and n_cond must be positive");
    ERT_REQUIRE(a->T > 0 && a->num_steps > 0 && a->num_steps <= a->T,
                "sample_chain: need 0 < num_steps <= T");
    ERT_REQUIRE(a->d_betas && a->d_alphas && a->d_alpha_bar, "sample_chain: schedule is NULL");
    ERT_REQUIRE(d_cond_bias, "sample_chain: cond_bias is NULL");
    ERT_REQUIRE(a->d_x_out, "sample_chain: x_out is NULL");
    ERT_REQUIRE(!(a->d_noise && !a->d_x_T), "sample_chain: injected noise needs d_x_T as well (row 0 of the draws)");
    const bool split = a->precision == ERTDIFF_PREC_BF16X3;
    const bool use_umma = a->precision == ERTDIFF_PREC_BF16 || split;
    if (a->precision != ERTDIFF_PREC_FP32 && !use_umma) return fail(ERTDIFF_ERR_ARG, "sample_chain: bad precision");
    if (split && !(chain_umma_split_supported(m->H, m->P) && m->w1_pk))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "sample_chain: the split-precision tensor-core chain is built for hidden_dim 128, param_dim <= 29");
    if (use_umma && !(chain_umma_supported(m->H, m->P) && m->w1_pk))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "sample_chain: the bf16 tensor-core chain is built for hidden_dim 128 or 256, param_dim <= 29");
    const int S = a->num_steps, H = m->H, P = m->P;
    const int64_t nstride = a->noise_member_stride_B > 0 ? a->noise_member_stride_B : a->B;
    ERT_REQUIRE(nstride >= a->B, "sample_chain: noise_member_stride_B < B");

    // c_t rows depend only on the weights: computed once per load_state_dict, extended on demand
    if (m->time_table_n < (size_t)S * H) {
        if (int rc = grow(m->time_table, m->time_table_n, (size_t)S * H)) return rc;
        m->time_rows_valid = 0;
    }
    if (int rc = grow(m->coef_table, m->coef_table_n, (size_t)S * 4)) return rc;
    if (m->time_rows_valid < S) {
        const int t0 = m->time_rows_valid;
        k_time_table<<<S - t0, H, 0, st>>>(m->freq, m->wtT, m->raw[7], m->w0tT, H, t0, m->time_table);
        ERT_LAUNCH_CHECK("k_time_table");
        m->time_rows_valid = S;
    }
    k_step_coefficients<<<(S + 127) / 128, 128, 0, st>>>(a->d_betas, a->d_alphas, a->d_alpha_bar, S,
                                                        (double)a->temperature, m->coef_table);
    ERT_LAUNCH_CHECK("k_step_coefficients");

    ChainParams p{};
    p.B = a->B; p.n_cond = a->n_cond; p.S = S; p.t_hi = S - 1; p.t_count = S;
    p.w0xT = m->w0xT; p.w2p = m->w2p; p.b2p = m->b2p; p.table = m->time_table;
    p.coef = m->coef_table; p.cond_bias = d_cond_bias;
    p.x_in = a->d_x_T; p.x_in_stride = P;
    p.noise = a->d_noise; p.noise_B = nstride;
    p.keys = make_philox_keys(a->seed); p.offset = a->offset; p.member_offset = a->member_offset;
    p.x_out = a->d_x_out; p.eps_trace = a->d_eps_trace; p.P = P;
    int mpb = 1, upt = 1;
    chain_fp32_tiling(a->B, H, &mpb, &upt);

    auto launch_any = [&](const ChainParams& q, cudaStream_t s2) -> int {
        if (!use_umma) {
            if (m->floor_mode) return launch_chain_fp32_floor(H, q, mpb, upt, s2);
            return launch_chain_fp32(H, q, mpb, upt, s2);
        }
        UmmaChainExtra ex{reinterpret_cast<const uint4*>(m->w1_pk), reinterpret_cast<const uint4*>(m->w2_pk), m->umma_status,
                          m->umma_timing_on ? m->umma_timing : nullptr, UC_M};
        return launch_chain_umma(H, q, ex, split, s2);
    };

    if (a->loop_mode == ERTDIFF_LOOP_PERSISTENT) {
        if (m->profile) {
            if (!m->ev_chain[0]) { ERT_CUDA(cudaEventCreate(&m->ev_chain[0])); ERT_CUDA(cudaEventCreate(&m->ev_chain[1])); }
            ERT_CUDA(cudaEventRecord(m->ev_chain[0], st));
        }
        const int rc = launch_any(p, st);
        if (m->profile && rc == 0) { ERT_CUDA(cudaEventRecord(m->ev_chain[1], st)); m->ev_valid = true; }
        return rc;
    }

    // ---- one kernel per timestep: plain stream launches, or the same sequence as a graph ----
    if (int rc = grow(m->xbuf[0], m->xbuf_n[0], (size_t)a->B * P)) return rc;
    if (int rc = grow(m->xbuf[1], m->xbuf_n[1], (size_t)a->B * P)) return rc;
    auto enqueue_steps = [&](cudaStream_t s) -> int {
        const float* xin = a->d_x_T;
        for (int it = 0; it < S; ++it) {
            ChainParams q = p;
            q.t_hi = S - 1 - it; q.t_count = 1;
            q.x_in = xin; q.x_in_stride = P;
            q.x_out = (it == S - 1) ? a->d_x_out : m->xbuf[it & 1];
            if (int rc = launch_any(q, s)) return rc;
            xin = q.x_out;
        }
        return 0;
    };
    if (a->loop_mode == ERTDIFF_LOOP_STREAM) return enqueue_steps(st);
    if (a->loop_mode != ERTDIFF_LOOP_GRAPH) return fail(ERTDIFF_ERR_ARG, "sample_chain: bad loop_mode");

    ertdiff_model::GraphKey key;
    key.B = a->B; key.n_cond = a->n_cond; key.steps = S; key.noise = a->d_noise; key.xT = a->d_x_T;
    key.xout = a->d_x_out; key.cb = d_cond_bias; key.seed = a->seed; key.offset = a->offset;
    key.moff = a->member_offset; key.nstride = nstride; key.trace = a->d_eps_trace;
    key.precision = a->precision; key.mpb = use_umma ? chain_umma_mpc(a->B) : mpb * 16 + upt;
    key.variant = chain_variant_id();
    key.time_table = m->time_table; key.coef_table = m->coef_table; key.xbuf0 = m->xbuf[0]; key.xbuf1 = m->xbuf[1];
    if (!(m->graph_exec && m->graph_key == key)) {
        // Anything a node bakes in changed (a pointer, the RNG stream, a reallocated scratch buffer, the
        // kernel variant): capture the step sequence again.  When the topology is unchanged -- the usual case:
        // same ensemble shape, new seed/offset or output address -- the instantiated graph is updated in place
        // (cudaGraphExecUpdate: microseconds per node); only a structural change pays for a new instantiation.
        cudaStream_t cap;
        ERT_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        ERT_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        const int64_t before = g_launches.load();
        int rc = enqueue_steps(cap);
        g_launches.store(before);                       // captured, not launched
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(cap, &graph);
        cudaStreamDestroy(cap);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(ERTDIFF_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        bool updated = false;
        if (m->graph_exec) {
            cudaGraphExecUpdateResultInfo info{};
            updated = cudaGraphExecUpdate(m->graph_exec, graph, &info) == cudaSuccess &&
                      info.result == cudaGraphExecUpdateSuccess;
            if (!updated) { (void)cudaGetLastError(); cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
        }
        if (!updated) {
            e = cudaGraphInstantiate(&m->graph_exec, graph, 0);
            if (e != cudaSuccess) {
                m->graph_exec = nullptr; cudaGraphDestroy(graph);
                return fail(ERTDIFF_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
            }
            ++m->graph_instantiations;
        } else {
            ++m->graph_updates;
        }
        cudaGraphDestroy(graph);
        m->graph_key = key;
    }
    ERT_CUDA(cudaGraphLaunch(m->graph_exec, st));
    g_launches.fetch_add(S, std::memory_order_relaxed);
    return 0;
}

}  // namespace ertdiff

using namespace ertdiff;

#pragma GCC visibility push(default)
extern "C" {

int ertdiff_abi_version(void) { return ERTDIFF_ABI_VERSION; }
const char* ertdiff_last_error(void) { return last_error().c_str(); }
int64_t ertdiff_launch_count(void) { return g_launches.load(); }
void ertdiff_launch_count_reset(void) { g_launches.store(0); }

int ertdiff_model_create(ertdiff_model** out, int device, int param_dim, int hidden_dim) {
    ERT_REQUIRE(out, "model_create: out is NULL");
    *out = nullptr;
    ERT_REQUIRE(param_dim >= 1 && param_dim <= kPPad, "model_create: param_dim must be in [1,32]");
    if (!(hidden_dim == 32 || hidden_dim == 64 || hidden_dim == 128 || hidden_dim == 256 ||
          hidden_dim == 512))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "model_create: hidden_dim must be one of 32,64,128,256,512");
    int ndev = 0;
    ERT_CUDA(cudaGetDeviceCount(&ndev));
    ERT_REQUIRE(device >= 0 && device < ndev, "model_create: no such CUDA device");
    DeviceGuard g(device);
    if (!g.ok) return fail(ERTDIFF_ERR_CUDA, "model_create: cudaSetDevice failed");
    auto* m = new ertdiff_model();
    m->device = device; m->P = param_dim; m->H = hidden_dim;
    raw_shapes(m->P, m->H, m->raw_n);
    const int H = m->H;
    auto alloc = [&](float*& p, size_t n) -> bool { return cudaMalloc(&p, n * sizeof(float)) == cudaSuccess; };
    bool ok = true;
    for (int i = 0; i < 12; ++i) ok = ok && alloc(m->raw[i], m->raw_n[i]);
    ok = ok && alloc(m->conv1_w, 32 * kInChannels * 3) && alloc(m->conv2_w, 64 * 96) &&
         alloc(m->w6T, (size_t)64 * H) && alloc(m->wtT, (size_t)H * H) &&
         alloc(m->w0xT, (size_t)kPPad * H) && alloc(m->w0tT, (size_t)H * H) &&
         alloc(m->w0cT, (size_t)H * H) && alloc(m->w2p, (size_t)kPPad * H) &&
         alloc(m->b2p, kPPad) && alloc(m->freq, H / 2);
    ok = ok && cudaMalloc(&m->enc_w1_pk, encoder_umma_w1_bytes()) == cudaSuccess &&
         cudaMalloc(&m->enc_w2_pk, encoder_umma_w2_bytes()) == cudaSuccess;
    if (ok && !m->umma_status) ok = cudaMalloc(&m->umma_status, sizeof(int)) == cudaSuccess &&
                                    cudaMalloc(&m->umma_timing, 16 * sizeof(long long)) == cudaSuccess;
    if (ok && chain_umma_supported(H, m->P)) {
        ok = cudaMalloc(&m->w1_pk, (size_t)2 * H * UC_K1 * 2) == cudaSuccess && cudaMalloc(&m->w2_pk, (size_t)2 * UC_N2 * H * 2) == cudaSuccess;   // operand + residual tiles
    }
    if (!ok) {
        ertdiff_model_destroy(m);
        return fail(ERTDIFF_ERR_CUDA, "model_create: cudaMalloc failed");
    }
    *out = m;
    return 0;
}

int ertdiff_model_profile(ertdiff_model* m, int enable) {
    if (int rc = check_model(m, false)) return rc;
    m->profile = enable != 0;
    m->ev_valid = false;
    return 0;
}

int ertdiff_model_last_chain_ms(ertdiff_model* m, float* h_ms) {
    if (int rc = check_model(m, false)) return rc;
    ERT_REQUIRE(h_ms, "last_chain_ms: h_ms is NULL");
    if (!m->ev_valid) return fail(ERTDIFF_ERR_STATE, "last_chain_ms: no profiled persistent chain launch yet");
    ERT_CUDA(cudaEventSynchronize(m->ev_chain[1]));
    ERT_CUDA(cudaEventElapsedTime(h_ms, m->ev_chain[0], m->ev_chain[1]));
    return 0;
}

int ertdiff_model_umma_status(ertdiff_model* m, int* h_status) {
    if (int rc = check_model(m, false)) return rc;
    ERT_REQUIRE(h_status, "umma_status: h_status is NULL");
    *h_status = 0;
    if (!m->umma_status) return 0;
    DeviceGuard g(m->device);
    ERT_CUDA(cudaMemcpy(h_status, m->umma_status, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int ertdiff_debug_umma_timing(ertdiff_model* m, int enable, int64_t* h_out16) {
    if (int rc = check_model(m, false)) return rc;
    DeviceGuard g(m->device);
    if (h_out16) ERT_CUDA(cudaMemcpy(h_out16, m->umma_timing, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    m->umma_timing_on = enable != 0;
    if (enable) ERT_CUDA(cudaMemset(m->umma_timing, 0, 16 * sizeof(long long)));
    return 0;
}

int ertdiff_model_destroy(ertdiff_model* m) {
    if (!m) return 0;
    DeviceGuard g(m->device);
    if (m->ev_chain[0]) { cudaEventDestroy(m->ev_chain[0]); cudaEventDestroy(m->ev_chain[1]); }
    for (int i = 0; i < 12; ++i) cudaFree(m->raw[i]);
    float* ptrs[] = {m->conv1_w, m->conv2_w, m->w6T, m->wtT, m->w0xT, m->w0tT, m->w0cT, m->w2p,
                     m->b2p, m->freq, m->enc_partial, m->cond_bias, m->cond_emb, m->time_table,
                     m->coef_table, m->xbuf[0], m->xbuf[1]};
    for (float* p : ptrs) cudaFree(p);
    cudaFree(m->w1_pk); cudaFree(m->w2_pk); cudaFree(m->umma_status); cudaFree(m->umma_timing);
    cudaFree(m->enc_w1_pk); cudaFree(m->enc_w2_pk);
    if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
    delete m;
    return 0;
}

int ertdiff_model_load(ertdiff_model* m, const float* const* tensors12, int on_device,
                       const float* h_freq, void* stream) {
    if (int rc = check_model(m, false)) return rc;
    ERT_REQUIRE(tensors12 && h_freq, "model_load: tensors12 / h_freq is NULL");
    DeviceGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < 12; ++i) {
        ERT_REQUIRE(tensors12[i], "model_load: a tensor pointer is NULL");
        ERT_CUDA(cudaMemcpyAsync(m->raw[i], tensors12[i], m->raw_n[i] * sizeof(float),
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
    ERT_CUDA(cudaMemcpyAsync(m->freq, h_freq, (m->H / 2) * sizeof(float), cudaMemcpyHostToDevice, st));
    const int64_t nmax = (int64_t)m->H * m->H > 64 * 96 ? (int64_t)m->H * m->H : 64 * 96;
    k_pack_weights<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(
        m->raw[0], m->raw[2], m->raw[4], m->raw[6], m->raw[8], m->raw[10], m->raw[11], m->P, m->H,
        m->conv1_w, m->conv2_w, m->w6T, m->wtT, m->w0xT, m->w0tT, m->w0cT, m->w2p, m->b2p);
    ERT_LAUNCH_CHECK("k_pack_weights");
    if (int rc = pack_encoder_umma_weights(m, st)) return rc;
    ERT_CUDA(cudaMemsetAsync(m->umma_status, 0, sizeof(int), st));
    if (m->w1_pk) {
        if (int rc = pack_chain_umma_weights(m->H, m->w0xT, m->w2p, m->P, m->w1_pk, m->w2_pk, st)) return rc;
        ERT_CUDA(cudaMemsetAsync(m->umma_status, 0, sizeof(int), st));
    }
    ERT_CUDA(cudaStreamSynchronize(st));
    m->loaded = true;
    m->time_rows_valid = 0;
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    return 0;
}

int ertdiff_model_export(ertdiff_model* m, float* const* tensors12, int on_device, void* stream) {
    if (int rc = check_model(m)) return rc;
    ERT_REQUIRE(tensors12, "model_export: tensors12 is NULL");
    DeviceGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < 12; ++i) {
        ERT_REQUIRE(tensors12[i], "model_export: a tensor pointer is NULL");
        ERT_CUDA(cudaMemcpyAsync(tensors12[i], m->raw[i], m->raw_n[i] * sizeof(float),
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    ERT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int ertdiff_encode_condition(ertdiff_model* m, const float* d_condition, int64_t n_cond,
                             int64_t L, int64_t cond_member_stride, float* d_cond_emb,
                             float* d_cond_bias, void* stream) {
    if (int rc = check_model(m)) return rc;
    DeviceGuard g(m->device);
    return run_encoder(m, d_condition, n_cond, L, cond_member_stride, d_cond_emb, d_cond_bias,
                       (cudaStream_t)stream);
}

int ertdiff_encode_condition_prec(ertdiff_model* m, const float* d_condition, int64_t n_cond,
                                  int64_t L, int64_t cond_member_stride, float* d_cond_emb,
                                  float* d_cond_bias, int32_t precision, void* stream) {
    if (int rc = check_model(m)) return rc;
    DeviceGuard g(m->device);
    if (precision == ERTDIFF_PREC_BF16)
        return run_encoder_umma(m, d_condition, n_cond, L, cond_member_stride, d_cond_emb, d_cond_bias,
                                (cudaStream_t)stream);
    // (ERTDIFF_PREC_BF16X3 keeps the fp32 encoder: the mode exists to return fp32-class fields, and the encoder runs
    // once per condition, not once per step)
    if (precision != ERTDIFF_PREC_FP32 && precision != ERTDIFF_PREC_BF16X3) return fail(ERTDIFF_ERR_ARG, "encode_condition: bad precision");
    return run_encoder(m, d_condition, n_cond, L, cond_member_stride, d_cond_emb, d_cond_bias,
                       (cudaStream_t)stream);
}

int ertdiff_forward(ertdiff_model* m, const float* d_x, const int64_t* d_t,
                    const float* d_condition, int64_t B, int64_t L, int64_t cond_member_stride,
                    float* d_out, void* stream) {
    if (int rc = check_model(m)) return rc;
    ERT_REQUIRE(d_x && d_t && d_condition && d_out && B > 0, "forward: NULL pointer or B <= 0");
    DeviceGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_cond = (cond_member_stride == 0) ? 1 : B;
    if (int rc = grow(m->cond_bias, m->cond_bias_n, (size_t)n_cond * m->H)) return rc;
    // the encoder launches with blockIdx.y = condition: at most 65535 per launch
    for (int64_t c0 = 0; c0 < n_cond; c0 += 32768) {
        const int64_t nc = (n_cond - c0) < 32768 ? (n_cond - c0) : 32768;
        if (int rc = run_encoder(m, d_condition + c0 * cond_member_stride, nc, L,
                                 cond_member_stride, nullptr, m->cond_bias + c0 * m->H, st))
            return rc;
    }
    k_forward_rows<<<(unsigned)B, m->H, 0, st>>>(d_x, d_t, m->cond_bias, n_cond, m->freq, m->wtT,
                                                 m->raw[7], m->w0tT, m->w0xT, m->w2p, m->b2p,
                                                 m->P, m->H, d_out);
    ERT_LAUNCH_CHECK("k_forward_rows");
    return 0;
}

int ertdiff_sample_chain(ertdiff_model* m, const ertdiff_chain_args* args, void* stream) {
    if (int rc = check_model(m)) return rc;
    ERT_REQUIRE(args, "sample_chain: args is NULL");
    DeviceGuard g(m->device);
    return run_chain(m, args, args->d_cond_bias, (cudaStream_t)stream);
}

int ertdiff_sample_model(ertdiff_model* m, const float* d_condition, int64_t L,
                         int64_t cond_member_stride, const ertdiff_chain_args* args,
                         void* stream) {
    if (int rc = check_model(m)) return rc;
    ERT_REQUIRE(args, "sample_model: args is NULL");
    DeviceGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_cond = args->n_cond;
    ERT_REQUIRE(n_cond > 0, "sample_model: n_cond must be positive");
    if (int rc = grow(m->cond_bias, m->cond_bias_n, (size_t)n_cond * m->H)) return rc;
    if (args->precision == ERTDIFF_PREC_BF16) {      // tensor-core encoder (batches internally)
        if (int rc = run_encoder_umma(m, d_condition, n_cond, L, cond_member_stride, nullptr, m->cond_bias, st))
            return rc;
    } else {
        for (int64_t c0 = 0; c0 < n_cond; c0 += 32768) {
            const int64_t nc = (n_cond - c0) < 32768 ? (n_cond - c0) : 32768;
            if (int rc = run_encoder(m, d_condition + c0 * cond_member_stride, nc, L,
                                     cond_member_stride, nullptr, m->cond_bias + c0 * m->H, st))
                return rc;
        }
    }
    return run_chain(m, args, m->cond_bias, st);
}

int ertdiff_step_coefficients(const float* d_betas, const float* d_alphas,
                              const float* d_alpha_bar, int32_t num_steps, double temperature,
                              float* d_table, void* stream) {
    ERT_REQUIRE(d_betas && d_alphas && d_alpha_bar && d_table && num_steps > 0,
                "step_coefficients: bad arguments");
    k_step_coefficients<<<(num_steps + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_betas, d_alphas, d_alpha_bar, num_steps, temperature, d_table);
    ERT_LAUNCH_CHECK("k_step_coefficients");
    return 0;
}

int ertdiff_philox_normal(uint64_t seed, uint64_t offset, int64_t member_offset, int64_t B,
                          int32_t P, int32_t draws, float* d_out, void* stream) {
    ERT_REQUIRE(d_out && B > 0 && P > 0 && P <= kPPad && draws > 0, "philox_normal: bad arguments");
    const int64_t n = (int64_t)draws * B * 8;
    k_philox_fill<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        make_philox_keys(seed), offset, member_offset, B, P, draws, d_out);
    ERT_LAUNCH_CHECK("k_philox_fill");
    return 0;
}

int ertdiff_posterior_update(const float* d_x, const float* d_eps, const float* d_z, float coef,
                             float c1, float sigma, int64_t n, float* d_out, void* stream) {
    ERT_REQUIRE(d_x && d_eps && d_out && n >= 0, "posterior_update: bad arguments");
    if (n == 0) return 0;
    const int64_t threads = (n + 3) / 4;
    k_posterior_update<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_x, d_eps, d_z, coef, c1, sigma, n, d_out);
    ERT_LAUNCH_CHECK("k_posterior_update");
    return 0;
}

int ertdiff_debug_umma_gemm(const float* d_A, const float* d_B, int32_t N, int32_t K, float* d_D,
                            void* stream) {
    ERT_REQUIRE(d_A && d_B && d_D, "debug_umma_gemm: NULL pointer");
    return umma_selftest(d_A, d_B, N, K, d_D, (cudaStream_t)stream);
}

int ertdiff_debug_chain_floor(ertdiff_model* m, int enable) {
    if (int rc = check_model(m, false)) return rc;
    m->floor_mode = enable != 0;
    return 0;
}

int ertdiff_debug_graph_stats(ertdiff_model* m, int64_t* h_out2) {
    if (int rc = check_model(m, false)) return rc;
    ERT_REQUIRE(h_out2, "debug_graph_stats: h_out2 is NULL");
    h_out2[0] = m->graph_instantiations; h_out2[1] = m->graph_updates;
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
