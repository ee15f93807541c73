// C ABI of libertdiff_b200.so -- see include/ertdiff_b200.h for the contract.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "model.cuh"
#include "encoder.cuh"
#include "denoiser.cuh"
#include "stats.cuh"
#include "misfit.cuh"
#include "umma.cuh"
#include "chain_umma.cuh"
#include "encoder_umma.cuh"

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

// ---- per-device scratch for the statistics entry points (they take no model handle) --------
// Grown on demand, kept for the life of the process: no allocation on the hot path.  Calls on
// one device are expected to come from one stream at a time.
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
};
static Workspace g_ws[64];
static std::mutex g_ws_mutex;

static int workspace(size_t bytes, void** out) {
    int dev = 0;
    ERT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(ERTDIFF_ERR_ARG, "workspace: device index out of range");
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Workspace& w = g_ws[dev];
    if (w.bytes < bytes) {
        if (w.ptr) { ERT_CUDA(cudaDeviceSynchronize()); cudaFree(w.ptr); w.ptr = nullptr; w.bytes = 0; }
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        ERT_CUDA(cudaMalloc(&w.ptr, want));
        w.bytes = want;
    }
    *out = w.ptr;
    return 0;
}

// cudaFuncSetAttribute is per device: every "set the dynamic shared memory limit once" site keeps one
// flag per device (a process may drive several GPUs through several handles)
struct PerDeviceOnce {
    bool done[64] = {};
    bool* slot() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        return &done[dev];
    }
};

static int colstats_attr() {
    static PerDeviceOnce once;
    bool& done = *once.slot();
    if (!done) {
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        done = true;
    }
    return 0;
}

// ---- encoder ----------------------------------------------------------------------------
static int run_encoder(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L,
                       int64_t member_stride, float* d_cond_emb, float* d_cond_bias,
                       cudaStream_t st) {
    ERT_REQUIRE(d_cond && n_cond > 0 && L > 0, "encode_condition: bad condition/n_cond/L");
    ERT_REQUIRE(n_cond <= 65535, "encode_condition: n_cond > 65535 per call; split the batch");
    const int64_t L1 = conv_out_len(L), L2 = conv_out_len(L1);
    // few conditions: 32 positions per CTA so that one condition still spreads over ~37 SMs
    const bool small = n_cond * ((L2 + 127) / 128) < 2 * kNumSMs;
    const int tp = small ? 32 : 128;
    const int n_chunks = (int)((L2 + tp - 1) / tp);
    if (int rc = grow(m->enc_partial, m->enc_partial_n, (size_t)n_cond * n_chunks * kConv2Out)) return rc;
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_encoder_conv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(EncSmem<4>)));
        ERT_CUDA(cudaFuncSetAttribute(k_encoder_conv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(EncSmem<1>)));
        attr_set = true;
    }
    dim3 grid(n_chunks, (unsigned)n_cond);
    if (small)
        k_encoder_conv<1><<<grid, ENC_THREADS, sizeof(EncSmem<1>), st>>>(
            d_cond, member_stride, L, L1, L2, m->conv1_w, m->raw[1], m->conv2_w, m->raw[3],
            m->enc_partial, n_chunks);
    else
        k_encoder_conv<4><<<grid, ENC_THREADS, sizeof(EncSmem<4>), st>>>(
            d_cond, member_stride, L, L1, L2, m->conv1_w, m->raw[1], m->conv2_w, m->raw[3],
            m->enc_partial, n_chunks);
    ERT_LAUNCH_CHECK("k_encoder_conv");
    k_encoder_finish<<<(unsigned)n_cond, m->H, 0, st>>>(m->enc_partial, n_chunks, L2, m->w6T,
                                                       m->raw[5], m->w0cT, m->raw[9], m->H,
                                                       d_cond_emb, d_cond_bias);
    ERT_LAUNCH_CHECK("k_encoder_finish");
    return 0;
}

// ---- tensor-core encoder (precision = bf16) ---------------------------------------------
static int run_encoder_umma(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L,
                            int64_t member_stride, float* d_cond_emb, float* d_cond_bias, cudaStream_t st) {
    ERT_REQUIRE(d_cond && n_cond > 0 && L > 0, "encode_condition: bad condition/n_cond/L");
    ERT_REQUIRE(m->enc_w1_pk, "encode_condition: tensor-core encoder weights missing");
    const int64_t L1 = conv_out_len(L), L2 = conv_out_len(L1);
    ERT_REQUIRE(4 * (L2 + 128) + 16 < (int64_t)1 << 30, "encode_condition: L too large");
    // tiles per CTA: as few as it takes to give every CTA slot (two per SM) work, up to 10 (a whole
    // condition of the reference grid) -- longer chunks amortise the per-CTA set-up (weights, TMEM,
    // barriers): 4096 conditions 472 -> 400 us
    const int64_t tiles = (L2 + 127) / 128;
    int64_t want_chunks = (2 * kNumSMs) / n_cond;           // chunks per condition that fill the CTA slots
    want_chunks = want_chunks < 1 ? 1 : (want_chunks > tiles ? tiles : want_chunks);
    int tpc = (int)((tiles + want_chunks - 1) / want_chunks);
    if (const char* e = std::getenv("ERTDIFF_ENC_TPC")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) tpc = v; }
    const int n_chunks = (int)((tiles + tpc - 1) / tpc);
    if (int rc = grow(m->enc_partial, m->enc_partial_n, (size_t)n_cond * n_chunks * kConv2Out)) return rc;
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_encoder_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncUmmaSmem)));
        attr_set = true;
    }
    const int64_t batch = 32768;               // blockIdx.y
    for (int64_t c0 = 0; c0 < n_cond; c0 += batch) {
        const int64_t nc = (n_cond - c0) < batch ? (n_cond - c0) : batch;
        const float* first = d_cond + c0 * member_stride;
        const uintptr_t addr = (uintptr_t)first, base = addr & ~(uintptr_t)15;     // bulk copies need 16-byte aligned sources
        const int64_t elem0 = (int64_t)((addr - base) / 4);
        EncUmmaParams p{};
        p.base = (const float*)base; p.elem0 = elem0; p.member_stride = member_stride;
        p.total = (elem0 + (nc - 1) * member_stride + kInChannels * L + 3) & ~(int64_t)3; p.L = (int)L; p.L1 = (int)L1; p.L2 = (int)L2;
        p.n_chunks = n_chunks; p.tpc = tpc;
        p.w1_pk = reinterpret_cast<const uint4*>(m->enc_w1_pk); p.w2_pk = reinterpret_cast<const uint4*>(m->enc_w2_pk);
        p.b1 = m->raw[1]; p.b2 = m->raw[3]; p.conv1_w = m->conv1_w;
        p.partial = m->enc_partial + (size_t)c0 * n_chunks * kConv2Out; p.status = m->umma_status;
        p.timing = m->umma_timing_on ? m->umma_timing : nullptr;
        k_encoder_umma<<<dim3(n_chunks, (unsigned)nc), EU_THREADS, sizeof(EncUmmaSmem), st>>>(p);
        ERT_LAUNCH_CHECK("k_encoder_umma");
    }
    k_encoder_finish<<<(unsigned)n_cond, m->H, 0, st>>>(m->enc_partial, n_chunks, L2, m->w6T, m->raw[5], m->w0cT,
                                                       m->raw[9], m->H, d_cond_emb, d_cond_bias);
    ERT_LAUNCH_CHECK("k_encoder_finish");
    return 0;
}

// ---- chain ------------------------------------------------------------------------------
template <int H, int MPB>
static void launch_chain_hm(const ChainParams& p, unsigned grid, cudaStream_t st) {
    const bool replay = p.noise != nullptr, trace = p.eps_trace != nullptr;
    if (replay) {
        if (trace) k_chain<H, MPB, true, true><<<grid, H, 0, st>>>(p);
        else k_chain<H, MPB, true, false><<<grid, H, 0, st>>>(p);
    } else {
        if (trace) k_chain<H, MPB, false, true><<<grid, H, 0, st>>>(p);
        else k_chain<H, MPB, false, false><<<grid, H, 0, st>>>(p);
    }
}

template <int H>
static int launch_chain_h(const ChainParams& p, int mpb, cudaStream_t st) {
    if (H >= 512 && mpb > 4) mpb = 4;            // static shared memory budget
    const unsigned grid = (unsigned)((p.B + mpb - 1) / mpb);
    if (mpb == 1) launch_chain_hm<H, 1>(p, grid, st);
    else if (mpb == 2) launch_chain_hm<H, 2>(p, grid, st);
    else if (mpb == 4 || H >= 512) launch_chain_hm<H, 4>(p, grid, st);
    else launch_chain_hm<H, (H >= 512 ? 4 : 8)>(p, grid, st);
    ERT_LAUNCH_CHECK("k_chain");
    return 0;
}

// rows of the tensor-core chain's 128-row tile that carry members (ERTDIFF_UMMA_MPC overrides: 32, 64, 128)
static int pick_umma_mpc(int64_t B) {
    if (const char* e = std::getenv("ERTDIFF_UMMA_MPC")) {
        const int v = std::atoi(e);
        if (v == 32 || v == 64 || v == 128) return v;
    }
    // measured (T=1000, one CTA per SM): 1.08 / 1.21 / 1.38 us per step at 32 / 64 / 128 rows; two
    // part-filled CTAs per SM are slower than one fuller one (2 x 32 rows: 1.53 us)
    for (int mpc = 32; mpc < UC_M; mpc *= 2)
        if ((B + mpc - 1) / mpc <= (int64_t)kNumSMs) return mpc;
    return UC_M;
}

static int launch_chain(int H, const ChainParams& p, int mpb, cudaStream_t st) {
    switch (H) {
        case 32: return launch_chain_h<32>(p, mpb, st);
        case 64: return launch_chain_h<64>(p, mpb, st);
        case 128: return launch_chain_h<128>(p, mpb, st);
        case 256: return launch_chain_h<256>(p, mpb, st);
        case 512: return launch_chain_h<512>(p, mpb, st);
    }
    return fail(ERTDIFF_ERR_UNSUPPORTED, "hidden_dim must be one of 32,64,128,256,512");
}

// members per CTA: spread small ensembles over all SMs (the chain is latency-bound), pack
// large ones so that weight registers are amortised over more members
static int pick_mpb(int64_t B, int H) {
    if (const char* e = std::getenv("ERTDIFF_CHAIN_MPB")) {      // tuning override
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) return v;
    }
    const int64_t ctas_per_sm = (H <= 128) ? 4 : (H <= 256 ? 2 : 1);
    const int64_t slots = kNumSMs * ctas_per_sm;
    // the one-member variant trades registers for latency (see k_chain): 3 CTAs per SM at H <= 128
    if (B <= kNumSMs * ((H <= 128) ? 3 : ctas_per_sm)) return 1;
    if (B <= 2 * slots) return 2;
    if (B <= 4 * slots) return 4;
    return 8;
}

static int run_chain(ertdiff_model* m, const ertdiff_chain_args* a, const float* d_cond_bias,
                     cudaStream_t st) {
    ERT_REQUIRE(a->B > 0 && a->n_cond > 0, "sample_chain: B and n_cond must be positive");
    ERT_REQUIRE(a->T > 0 && a->num_steps > 0 && a->num_steps <= a->T,
                "sample_chain: need 0 < num_steps <= T");
    ERT_REQUIRE(a->d_betas && a->d_alphas && a->d_alpha_bar, "sample_chain: schedule is NULL");
    ERT_REQUIRE(d_cond_bias, "sample_chain: cond_bias is NULL");
    ERT_REQUIRE(a->d_x_out, "sample_chain: x_out is NULL");
    ERT_REQUIRE(!(a->d_noise && !a->d_x_T), "sample_chain: injected noise needs d_x_T as well (row 0 of the draws)");
    const bool use_umma = a->precision == ERTDIFF_PREC_BF16;
    if (a->precision != ERTDIFF_PREC_FP32 && !use_umma) return fail(ERTDIFF_ERR_ARG, "sample_chain: bad precision");
    if (use_umma && !(m->H == UC_H && m->P <= 29 && m->w1_pk))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "sample_chain: the bf16 tensor-core chain is built for hidden_dim = 128, param_dim <= 29");
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
    const int mpb = pick_mpb(a->B, H);

    auto launch_any = [&](const ChainParams& q, cudaStream_t s2) -> int {
        if (!use_umma) return launch_chain(H, q, mpb, s2);
        UmmaChainExtra ex{reinterpret_cast<const uint4*>(m->w1_pk), reinterpret_cast<const uint4*>(m->w2_pk), m->umma_status,
                          m->umma_timing_on ? m->umma_timing : nullptr, UC_M};
        // members per CTA: the fewest rows per tile that still fit one wave of CTAs, so that a mid-size
        // ensemble runs on all SMs (the step is latency-bound: a part-filled tile steps faster)
        ex.mpc = pick_umma_mpc(q.B);
        const unsigned grid = (unsigned)((q.B + ex.mpc - 1) / ex.mpc);
        // more tiles than SMs: the build that keeps two CTAs resident per SM
        const bool two = grid > (unsigned)kNumSMs && !std::getenv("ERTDIFF_UMMA_ONE_CTA");
        const size_t smem = two ? sizeof(UmmaChainSmem<uc_nslot(2)>) : sizeof(UmmaChainSmem<uc_nslot(1)>);
        const int variant = (q.noise != nullptr ? 4 : 0) | (q.eps_trace != nullptr ? 2 : 0) | (q.n_cond == 1 ? 1 : 0);
        using Kern = void (*)(const ChainParams, const UmmaChainExtra);
        static const Kern kerns[8] = {
            k_chain_umma<false, false, false, 1>, k_chain_umma<false, false, true, 1>,
            k_chain_umma<false, true, false, 1>,  k_chain_umma<false, true, true, 1>,
            k_chain_umma<true, false, false, 1>,  k_chain_umma<true, false, true, 1>,
            k_chain_umma<true, true, false, 1>,   k_chain_umma<true, true, true, 1>};
        static const Kern kerns2[8] = {             // two CTAs per SM
            k_chain_umma<false, false, false, 2>, k_chain_umma<false, false, true, 2>,
            k_chain_umma<false, true, false, 2>,  k_chain_umma<false, true, true, 2>,
            k_chain_umma<true, false, false, 2>,  k_chain_umma<true, false, true, 2>,
            k_chain_umma<true, true, false, 2>,   k_chain_umma<true, true, true, 2>};
        static PerDeviceOnce once;
        bool& attr_set = *once.slot();
        if (!attr_set) {
            for (Kern k : kerns)
                ERT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UmmaChainSmem<uc_nslot(1)>)));
            for (Kern k : kerns2)
                ERT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UmmaChainSmem<uc_nslot(2)>)));
            attr_set = true;
        }
        if (two) kerns2[variant]<<<grid, UC_THREADS, smem, s2>>>(q, ex);
        else kerns[variant]<<<grid, UC_THREADS, smem, s2>>>(q, ex);
        ERT_LAUNCH_CHECK("k_chain_umma");
        return 0;
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
    if (!(m->graph_exec && m->graph_key == key)) {
        if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
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
        e = cudaGraphInstantiate(&m->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { m->graph_exec = nullptr; return fail(ERTDIFF_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e)); }
        m->graph_key = key;
    }
    ERT_CUDA(cudaGraphLaunch(m->graph_exec, st));
    g_launches.fetch_add(S, std::memory_order_relaxed);
    return 0;
}

template <int N, int K>
static int run_umma_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
    int* d_status = nullptr;
    if (int rc = workspace(256, (void**)&d_status)) return rc;
    ERT_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    const size_t smem = umma::tile_bytes(128, K) + umma::tile_bytes(N, K);
    ERT_CUDA(cudaFuncSetAttribute(umma::k_umma_selftest<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma::k_umma_selftest<N, K><<<1, 128, smem, st>>>(A, B, D, d_status);
    ERT_LAUNCH_CHECK("k_umma_selftest");
    int h = 0;
    ERT_CUDA(cudaMemcpyAsync(&h, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    ERT_CUDA(cudaStreamSynchronize(st));
    if (h) return fail(ERTDIFF_ERR_CUDA, "umma selftest: mbarrier wait timed out (MMA never completed)");
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
    ok = ok && cudaMalloc(&m->enc_w1_pk, kConv1Out * EU_K1 * 2) == cudaSuccess &&
         cudaMalloc(&m->enc_w2_pk, kConv2Out * EU_K2 * 2) == cudaSuccess;
    if (ok && !m->umma_status) ok = cudaMalloc(&m->umma_status, sizeof(int)) == cudaSuccess &&
                                    cudaMalloc(&m->umma_timing, 16 * sizeof(long long)) == cudaSuccess;
    if (ok && H == UC_H) {
        ok = cudaMalloc(&m->w1_pk, UC_H * UC_K1 * 2) == cudaSuccess && cudaMalloc(&m->w2_pk, UC_N2 * UC_H * 2) == cudaSuccess;
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
    k_pack_encoder_umma<<<(kConv2Out * EU_K2 + 255) / 256, 256, 0, st>>>(m->raw[0], m->raw[2], m->enc_w1_pk, m->enc_w2_pk);
    ERT_LAUNCH_CHECK("k_pack_encoder_umma");
    ERT_CUDA(cudaMemsetAsync(m->umma_status, 0, sizeof(int), st));
    if (m->w1_pk) {
        k_pack_umma_weights<<<(UC_H * UC_K1 + 255) / 256, 256, 0, st>>>(m->w0xT, m->w2p, m->P, m->w1_pk, m->w2_pk);
        ERT_LAUNCH_CHECK("k_pack_umma_weights");
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
    if (precision != ERTDIFF_PREC_FP32) return fail(ERTDIFF_ERR_ARG, "encode_condition: bad precision");
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

int ertdiff_ensemble_moments(const void* d_a, int dtype, int64_t N, int64_t Q, void* d_mean,
                             void* d_std, void* d_var, void* stream) {
    ERT_REQUIRE(d_a && N > 0 && Q > 0, "ensemble_moments: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((Q + 127) / 128);
    if (Q <= SQ_MAXQ && N >= 64 && (dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64)) {
        // few columns: one CTA streams the rows through shared memory (see k_colstats_smallq)
        const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
        const size_t smem = (size_t)N * esz < (size_t)SQ_SMEM_BYTES ? (size_t)N * esz : (size_t)SQ_SMEM_BYTES;
        if (int rc = colstats_attr()) return rc;
        if (dtype == ERTDIFF_F32)
            k_colstats_smallq<float, 0><<<(unsigned)Q, 256, smem, st>>>((const float*)d_a, N, Q, (float*)d_mean, (float*)d_std, (float*)d_var, 0.0, nullptr);
        else
            k_colstats_smallq<double, 0><<<(unsigned)Q, 256, smem, st>>>((const double*)d_a, N, Q, (double*)d_mean, (double*)d_std, (double*)d_var, 0.0, nullptr);
        ERT_LAUNCH_CHECK("k_colstats_smallq");
        return 0;
    }
    if (dtype == ERTDIFF_F32)
        k_moments<float><<<grid, 128, 0, st>>>((const float*)d_a, N, Q, (float*)d_mean,
                                               (float*)d_std, (float*)d_var);
    else if (dtype == ERTDIFF_F64)
        k_moments<double><<<grid, 128, 0, st>>>((const double*)d_a, N, Q, (double*)d_mean,
                                                (double*)d_std, (double*)d_var);
    else
        return fail(ERTDIFF_ERR_ARG, "ensemble_moments: bad dtype");
    ERT_LAUNCH_CHECK("k_moments");
    return 0;
}

int ertdiff_ensemble_percentiles(const void* d_a, int dtype, int64_t N, int64_t Q,
                                 const double* h_q, int32_t nq, int index_dtype, void* d_out,
                                 void* stream) {
    ERT_REQUIRE(d_a && d_out && h_q && N > 0 && Q > 0 && nq > 0, "ensemble_percentiles: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_percentiles: bad dtype");
    ERT_REQUIRE(!(dtype == ERTDIFF_F64 && index_dtype == ERTDIFF_F32),
                "ensemble_percentiles: float64 data always uses float64 index arithmetic");
    cudaStream_t st = (cudaStream_t)stream;
    // numpy's index arithmetic (function_base._quantile, method 'linear'), in index_dtype
    std::vector<PctlQuery> qs(nq);
    for (int k = 0; k < nq; ++k) {
        ERT_REQUIRE(h_q[k] >= 0.0 && h_q[k] <= 100.0, "ensemble_percentiles: q outside [0,100]");
        PctlQuery& q = qs[k];
        if (index_dtype == ERTDIFF_F32) {
            const float quant = (float)h_q[k] / 100.0f;
            const float vi = (float)(N - 1) * quant;
            const float lo = floorf(vi);
            q.gamma_f = vi - lo; q.gamma_d = 0.0;
            q.lo = (int32_t)lo; q.hi = q.lo + 1;
            if (vi >= (float)(N - 1)) q.lo = q.hi = (int32_t)(N - 1);
        } else {
            const double quant = h_q[k] / 100.0;
            const double vi = (double)(N - 1) * quant;
            const double lo = floor(vi);
            q.gamma_d = vi - lo; q.gamma_f = 0.f;
            q.lo = (int32_t)lo; q.hi = q.lo + 1;
            if (vi >= (double)(N - 1)) q.lo = q.hi = (int32_t)(N - 1);
        }
    }
    int NP = 1;
    while (NP < N) NP <<= 1;
    const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
    const size_t budget = 200 * 1024;
    ERT_REQUIRE((size_t)NP * esz + 64 <= budget, "ensemble_percentiles: N too large for one CTA's shared memory");
    int CT = 32;
    while (CT > 1 && ((size_t)CT * NP * esz + CT * sizeof(int) > budget / 2 || (int64_t)CT > Q)) CT >>= 1;
    // keep enough CTAs in flight
    while (CT > 1 && (Q + CT - 1) / CT < 2 * kNumSMs) CT >>= 1;
    const size_t smem = (size_t)CT * NP * esz + CT * sizeof(int);
    const int threads = (int64_t)CT * NP / 2 >= 1024 ? 1024 : ((int64_t)CT * NP / 2 >= 256 ? 256 : 64);
    const unsigned grid = (unsigned)((Q + CT - 1) / CT);
    cudaError_t e = cudaSuccess;
    // queries travel as a by-value kernel argument, kMaxPctlQueries per launch
    for (int k0 = 0; k0 < nq && e == cudaSuccess; k0 += kMaxPctlQueries) {
        PctlQueryPack pack{};
        pack.n = (nq - k0) < kMaxPctlQueries ? (nq - k0) : kMaxPctlQueries;
        for (int k = 0; k < pack.n; ++k) pack.q[k] = qs[k0 + k];
#define ERT_PCT_LAUNCH(T, G, O)                                                              \
    do {                                                                                     \
        e = cudaFuncSetAttribute(k_percentiles<T, G, O>,                                     \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);  \
        if (e == cudaSuccess)                                                                \
            k_percentiles<T, G, O><<<grid, threads, smem, st>>>((const T*)d_a, N, Q, NP, CT, \
                                                               pack, (O*)d_out + (size_t)k0 * Q); \
    } while (0)
        if (dtype == ERTDIFF_F32 && index_dtype == ERTDIFF_F32) ERT_PCT_LAUNCH(float, float, float);
        else if (dtype == ERTDIFF_F32) ERT_PCT_LAUNCH(float, double, double);
        else ERT_PCT_LAUNCH(double, double, double);
#undef ERT_PCT_LAUNCH
        if (e != cudaSuccess) return fail(ERTDIFF_ERR_CUDA, std::string("percentiles attr: ") + cudaGetErrorString(e));
        ERT_LAUNCH_CHECK("k_percentiles");
    }
    return 0;
}

int ertdiff_interval_coverage(const double* d_low, const double* d_upp, const double* d_truth,
                              int32_t n_intervals, int64_t Q, int32_t P, int32_t* d_counts, void* stream) {
    ERT_REQUIRE(d_low && d_upp && d_truth && d_counts, "interval_coverage: NULL pointer");
    ERT_REQUIRE(n_intervals > 0 && Q > 0 && P > 0 && P <= kPPad && Q % P == 0,
                "interval_coverage: need n_intervals > 0, 0 < P <= 32 and Q a multiple of P");
    k_interval_coverage<<<(unsigned)n_intervals, 256, 0, (cudaStream_t)stream>>>(d_low, d_upp, d_truth, Q, P, d_counts);
    ERT_LAUNCH_CHECK("k_interval_coverage");
    return 0;
}

int ertdiff_minmax(const void* d_a, int dtype, int64_t n, double* d_out2, void* stream) {
    ERT_REQUIRE(d_a && d_out2 && n > 0, "minmax: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (blocks < 1) blocks = 1;
    double* part = nullptr;
    if (int rc = workspace((size_t)blocks * 3 * sizeof(double), (void**)&part)) return rc;
    if (dtype == ERTDIFF_F32) k_minmax_partial<float><<<blocks, 256, 0, st>>>((const float*)d_a, n, part);
    else if (dtype == ERTDIFF_F64) k_minmax_partial<double><<<blocks, 256, 0, st>>>((const double*)d_a, n, part);
    else return fail(ERTDIFF_ERR_ARG, "minmax: bad dtype");
    ERT_LAUNCH_CHECK("k_minmax_partial");
    k_minmax_final<<<1, 32, 0, st>>>(part, blocks, d_out2);
    ERT_LAUNCH_CHECK("k_minmax_final");
    return 0;
}

// persistent, self-cleaning per-column tickets (last-CTA-done pattern of the KDE kernels)
static int kde_tickets(unsigned int** out) {
    static unsigned int* tickets[64] = {};
    int dev = 0;
    ERT_CUDA(cudaGetDevice(&dev));
    ERT_REQUIRE(dev >= 0 && dev < 64, "kde: device index out of range");
    if (!tickets[dev]) {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        if (!tickets[dev]) {
            ERT_CUDA(cudaMalloc(&tickets[dev], 4096 * sizeof(unsigned int)));
            ERT_CUDA(cudaMemset(tickets[dev], 0, 4096 * sizeof(unsigned int)));
        }
    }
    *out = tickets[dev];
    return 0;
}

int ertdiff_ensemble_kde_mode(const void* d_a, int dtype, int64_t N, int64_t Q,
                              const double* d_lohi, int32_t n_grid, double* d_mode,
                              int64_t* d_index, void* stream) {
    ERT_REQUIRE(d_a && d_lohi && N > 1 && Q > 0 && n_grid > 1, "ensemble_kde_mode: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_kde_mode: bad dtype");
    ERT_REQUIRE((size_t)N * 8 <= 200 * 1024, "ensemble_kde_mode: N too large for shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    const int G = n_grid;
    // scipy: factor = neff**(-1/(d+4)) with d = 1, neff = N; covariance = data_cov * factor**2
    const double factor = std::pow((double)N, -1.0 / 5.0);
    const double f2 = factor * factor;
    // columns per batch: the fp32 scan of a batch lives in the workspace (<= 128 MiB)
    int64_t qb = (int64_t)((128u << 20) / ((size_t)G * sizeof(float)));
    if (qb < 1) qb = 1;
    if (qb > Q) qb = Q;
    if (qb > 65535 * 16) qb = 65535 * 16;
    void* ws = nullptr;
    const size_t cols_bytes = ((size_t)Q * sizeof(KdeColumn) + 255) & ~(size_t)255;
    // the float64 selection of a column is shared by several CTAs when there are few columns and many
    // members (one CTA would re-evaluate ~100 candidates x N members alone)
    const int sel_parts = (qb <= 4096 && N >= 1024) ? (qb * 8 <= 4 * kNumSMs ? 8 : (qb * 2 <= 4 * kNumSMs ? 2 : 1)) : 1;
    const size_t part_bytes = (size_t)qb * sel_parts * 2 * sizeof(double);
    if (int rc = workspace(cols_bytes + part_bytes + (size_t)qb * G * sizeof(float), &ws)) return rc;
    unsigned int* tk = nullptr;
    if (int rc = kde_tickets(&tk)) return rc;
    KdeColumn* cols = (KdeColumn*)ws;
    double* partials = (double*)((char*)ws + cols_bytes);
    float* s32 = (float*)((char*)ws + cols_bytes + part_bytes);
    const bool f32in = dtype == ERTDIFF_F32;
    if (Q <= SQ_MAXQ && N >= 64) {
        static_assert(sizeof(KdeColumn) == 2 * sizeof(double), "KdeColumn is written as two doubles");
        const size_t esz = f32in ? 4 : 8;
        const size_t smem = (size_t)N * esz < (size_t)SQ_SMEM_BYTES ? (size_t)N * esz : (size_t)SQ_SMEM_BYTES;
        if (int rc = colstats_attr()) return rc;
        if (f32in) k_colstats_smallq<float, 1><<<(unsigned)Q, 256, smem, st>>>((const float*)d_a, N, Q, nullptr, nullptr, nullptr, f2, (double*)cols);
        else k_colstats_smallq<double, 1><<<(unsigned)Q, 256, smem, st>>>((const double*)d_a, N, Q, nullptr, nullptr, nullptr, f2, (double*)cols);
        ERT_LAUNCH_CHECK("k_colstats_smallq");
    } else {
        const unsigned grid = (unsigned)((Q * 32 + 255) / 256);
        if (f32in) k_kde_prepare<float><<<grid, 256, 0, st>>>((const float*)d_a, N, Q, f2, cols);
        else k_kde_prepare<double><<<grid, 256, 0, st>>>((const double*)d_a, N, Q, f2, cols);
        ERT_LAUNCH_CHECK("k_kde_prepare");
    }
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    for (int64_t c0 = 0; c0 < Q; c0 += qb) {
        const int64_t nc = (Q - c0) < qb ? (Q - c0) : qb;
        // CTAs per column: one 256-thread pass over the grid each while nc * n_gchunks stays within ~4
        // CTAs per SM; with many columns (maps) one CTA walks the whole grid of its column
        int n_gchunks = (G + 255) / 256;
        while (n_gchunks > 1 && nc * n_gchunks > 4 * kNumSMs) n_gchunks = (n_gchunks + 1) / 2;
        const int gchunk = (G + n_gchunks - 1) / n_gchunks;
        const int threads = 256;
        const dim3 grid((unsigned)nc, (unsigned)n_gchunks);
        if (f32in) {
            k_kde_scan32<float><<<grid, threads, (size_t)N * 4, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, gchunk, cols, s32);
            ERT_LAUNCH_CHECK("k_kde_scan32");
            k_kde_select64<float><<<dim3((unsigned)nc, (unsigned)sel_parts), 256, (size_t)N * 8, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk);
        } else {
            k_kde_scan32<double><<<grid, threads, (size_t)N * 4, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, gchunk, cols, s32);
            ERT_LAUNCH_CHECK("k_kde_scan32");
            k_kde_select64<double><<<dim3((unsigned)nc, (unsigned)sel_parts), 256, (size_t)N * 8, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk);
        }
        ERT_LAUNCH_CHECK("k_kde_select64");
    }
    return 0;
}

int ertdiff_ensemble_kde_mode_auto(const void* d_a, int dtype, int64_t N, int64_t Q, int32_t n_grid,
                                   double* d_lohi, double* d_mode, int64_t* d_index, void* stream) {
    ERT_REQUIRE(d_a && d_lohi && N > 1 && Q > 0 && n_grid > 1, "ensemble_kde_mode_auto: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_kde_mode_auto: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    // (from ~1000 members on, the staged kernels win: their float64 selection is shared by 8 CTAs per column)
    const bool small = N * Q <= 65536 && N < 1024 && Q <= 4096;
    if (!small) {
        if (int rc = ertdiff_minmax(d_a, dtype, N * Q, d_lohi, stream)) return rc;
        return ertdiff_ensemble_kde_mode(d_a, dtype, N, Q, d_lohi, n_grid, d_mode, d_index, stream);
    }
    // one fused launch (k_kde_small); tickets: a persistent, self-cleaning counter per column
    unsigned int* tk = nullptr;
    if (int rc = kde_tickets(&tk)) return rc;
    const int G = n_grid;
    void* ws = nullptr;
    if (int rc = workspace((size_t)Q * G * sizeof(float), &ws)) return rc;
    // CTAs per column: one 256-thread pass over the grid each (every thread scans one point) when that
    // still fits in one wave of co-resident CTAs (a second, nearly empty wave would double the
    // duration); otherwise fewer, longer chunks
    int n_gchunks = (G + 255) / 256;
    if (const char* e = std::getenv("ERTDIFF_KDE_GCHUNKS")) n_gchunks = std::atoi(e) > 0 ? std::atoi(e) : 1;
    else
        while (n_gchunks > 1 && Q * n_gchunks > 4 * kNumSMs) n_gchunks = (n_gchunks + 1) / 2;
    const int gchunk = (G + n_gchunks - 1) / n_gchunks;
    const double factor = std::pow((double)N, -1.0 / 5.0);
    const dim3 grid((unsigned)Q, (unsigned)n_gchunks);
    const size_t smem = (size_t)N * 12;
    if (dtype == ERTDIFF_F32)
        k_kde_small<float><<<grid, 256, smem, st>>>((const float*)d_a, N, Q, 1, d_lohi, G, gchunk, factor * factor,
                                                    (float*)ws, tk, d_mode, d_index);
    else
        k_kde_small<double><<<grid, 256, smem, st>>>((const double*)d_a, N, Q, 1, d_lohi, G, gchunk, factor * factor,
                                                     (float*)ws, tk, d_mode, d_index);
    ERT_LAUNCH_CHECK("k_kde_small");
    return 0;
}

int ertdiff_debug_umma_gemm(const float* d_A, const float* d_B, int32_t N, int32_t K, float* d_D,
                            void* stream) {
    ERT_REQUIRE(d_A && d_B && d_D, "debug_umma_gemm: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 128 && K == 32) return run_umma_selftest<128, 32>(d_A, d_B, d_D, st);
    if (N == 32 && K == 128) return run_umma_selftest<32, 128>(d_A, d_B, d_D, st);
    if (N == 128 && K == 128) return run_umma_selftest<128, 128>(d_A, d_B, d_D, st);
    if (N == 64 && K == 96) return run_umma_selftest<64, 96>(d_A, d_B, d_D, st);
    return fail(ERTDIFF_ERR_UNSUPPORTED, "debug_umma_gemm: (N,K) must be (128,32), (32,128), (128,128) or (64,96)");
}

int ertdiff_untransform_bounds(const float* d_u, int64_t B, int32_t P, float a, float b,
                               const double* d_scaler_min, const double* d_scaler_scale,
                               const double* d_lim_lo, const double* d_lim_hi, float* d_phys,
                               uint8_t* d_valid, int32_t* d_first_bad, void* stream) {
    ERT_REQUIRE(d_u && B > 0 && P > 0 && P <= 32, "untransform_bounds: bad arguments");
    ERT_REQUIRE((d_scaler_min == nullptr) == (d_scaler_scale == nullptr), "untransform_bounds: give both scaler arrays or neither");
    ERT_REQUIRE((d_lim_lo == nullptr) == (d_lim_hi == nullptr), "untransform_bounds: give both limit arrays or neither");
    const int64_t threads = B * 32;
    k_untransform_bounds<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_u, B, P, a, b, d_scaler_min, d_scaler_scale, d_lim_lo, d_lim_hi, d_phys, d_valid, d_first_bad);
    ERT_LAUNCH_CHECK("k_untransform_bounds");
    return 0;
}

// ---- per-member misfit metrics (ECD.py:764-785, 927-930) ---------------------------------------
// numpy's pairwise-summation tree for n elements, flattened once per (device, n) and kept on the device
extern "C++" {
struct PairwiseNodes {
    std::vector<int2> leaves, nodes;      // nodes: children as (is_node ? -1 - k : leaf index) until fixed up
    std::vector<int> height;
    int build(int64_t start, int64_t n) {
        if (n <= 128) {
            leaves.push_back(make_int2((int)start, (int)n));
            return (int)leaves.size() - 1;
        }
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        const int l = build(start, n2), r = build(start + n2, n - n2);
        const int hl = l < 0 ? height[-1 - l] : 0, hr = r < 0 ? height[-1 - r] : 0;
        nodes.push_back(make_int2(l, r));
        height.push_back((hl > hr ? hl : hr) + 1);
        return -(int)nodes.size();
    }
};

static int pairwise_plan(int64_t n, PairwisePlan* out) {
    static std::map<std::pair<int, int64_t>, PairwisePlan> cache;
    int dev = 0;
    ERT_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto it = cache.find({dev, n});
    if (it != cache.end()) { *out = it->second; return 0; }
    PairwiseNodes t;
    t.build(0, n);
    const int nl = (int)t.leaves.size(), nn = (int)t.nodes.size();
    int n_levels = 0;
    for (int h : t.height) n_levels = h > n_levels ? h : n_levels;
    // order the internal nodes by height (children always sit on a lower level), remap the child indices
    std::vector<int> order, pos(nn), level_off(n_levels + 1, 0);
    for (int h = 1; h <= n_levels; ++h) {
        level_off[h - 1] = (int)order.size();
        for (int k = 0; k < nn; ++k)
            if (t.height[k] == h) { pos[k] = (int)order.size(); order.push_back(k); }
    }
    level_off[n_levels] = nn;
    std::vector<int2> nodes(nn);
    for (int i = 0; i < nn; ++i) {
        const int2 c = t.nodes[order[i]];
        nodes[i] = make_int2(c.x < 0 ? nl + pos[-1 - c.x] : c.x, c.y < 0 ? nl + pos[-1 - c.y] : c.y);
    }
    int2 *d_leaves = nullptr, *d_nodes = nullptr;
    int* d_off = nullptr;
    ERT_CUDA(cudaMalloc(&d_leaves, sizeof(int2) * nl));
    ERT_CUDA(cudaMalloc(&d_nodes, sizeof(int2) * (nn > 0 ? nn : 1)));
    ERT_CUDA(cudaMalloc(&d_off, sizeof(int) * (n_levels + 1)));
    ERT_CUDA(cudaMemcpy(d_leaves, t.leaves.data(), sizeof(int2) * nl, cudaMemcpyHostToDevice));
    if (nn) ERT_CUDA(cudaMemcpy(d_nodes, nodes.data(), sizeof(int2) * nn, cudaMemcpyHostToDevice));
    ERT_CUDA(cudaMemcpy(d_off, level_off.data(), sizeof(int) * (n_levels + 1), cudaMemcpyHostToDevice));
    PairwisePlan pl{d_leaves, d_nodes, d_off, nl, nn, n_levels};
    cache[{dev, n}] = pl;
    *out = pl;
    return 0;
}

template <typename T>
static int launch_misfit(const void* d_sims, const void* d_obs, int64_t N, int64_t L, int64_t C, double A, double B,
                         void* d_wsse, void* d_wsse_total, void* d_mse, cudaStream_t st) {
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_misfit_wsse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_misfit_mse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (d_wsse) {
        PairwisePlan pl;
        if (int rc = pairwise_plan(L, &pl)) return rc;
        // rows per chunk: whole leaves (<= 128 rows each), sized so that several CTAs stay resident per SM
        int rcap = (int)(L < 512 ? (L < 128 ? 128 : L) : 512);
        while (rcap > 128 && (size_t)C * (rcap | 1) * sizeof(T) > 64 * 1024) rcap = rcap / 2 < 128 ? 128 : rcap / 2;
        const size_t smem = ((size_t)C * (pl.n_leaves + pl.n_nodes) + C + (size_t)C * (rcap | 1)) * sizeof(T);
        ERT_REQUIRE(smem <= 200 * 1024, "misfit_metrics: L x C too large for the on-chip summation tree");
        k_misfit_wsse<T><<<(unsigned)N, kMisfitThreads, smem, st>>>((const T*)d_sims, (const T*)d_obs, (int)L, (int)C, (T)A, (T)B,
                                                                    rcap, pl, (T*)d_wsse, (T*)d_wsse_total);
        ERT_LAUNCH_CHECK("k_misfit_wsse");
    }
    if (d_mse) {
        PairwisePlan pl;
        if (int rc = pairwise_plan(L * C, &pl)) return rc;
        const size_t smem = (size_t)(pl.n_leaves + pl.n_nodes) * sizeof(T);
        ERT_REQUIRE(smem <= 200 * 1024, "misfit_metrics: map too large for the on-chip summation tree");
        k_misfit_mse<T><<<(unsigned)N, kMisfitThreads, smem, st>>>((const T*)d_sims, (const T*)d_obs, L * C, pl, (T*)d_mse);
        ERT_LAUNCH_CHECK("k_misfit_mse");
    }
    return 0;
}
}  // extern "C++"

int ertdiff_misfit_metrics(const void* d_sims, const void* d_obs, int dtype, int64_t N, int64_t L, int64_t C,
                           double A, double B, void* d_wsse, void* d_wsse_total, void* d_mse, void* stream) {
    ERT_REQUIRE(d_sims && d_obs && N > 0 && L > 0 && C > 0, "misfit_metrics: bad arguments");
    ERT_REQUIRE(C <= 128 && L * C < (int64_t(1) << 31), "misfit_metrics: need C <= 128 and L*C < 2^31");
    ERT_REQUIRE((d_wsse == nullptr) == (d_wsse_total == nullptr), "misfit_metrics: give both WSSE outputs or neither");
    ERT_REQUIRE(d_wsse || d_mse, "misfit_metrics: no output requested");
    if (dtype == ERTDIFF_F32) return launch_misfit<float>(d_sims, d_obs, N, L, C, A, B, d_wsse, d_wsse_total, d_mse, (cudaStream_t)stream);
    if (dtype == ERTDIFF_F64) return launch_misfit<double>(d_sims, d_obs, N, L, C, A, B, d_wsse, d_wsse_total, d_mse, (cudaStream_t)stream);
    return fail(ERTDIFF_ERR_ARG, "misfit_metrics: bad dtype");
}

extern "C++" {
template <typename T>
static int launch_sort_rows(const void* d_in, int64_t rows, int64_t stride, int n, int npad, double* d_out, cudaStream_t st) {
    k_sort_rows_f64<T><<<(unsigned)rows, 1024, 0, st>>>((const T*)d_in, stride, n, npad, d_out);
    ERT_LAUNCH_CHECK("k_sort_rows_f64");
    return 0;
}
}  // extern "C++"

int ertdiff_wasserstein_distance(const void* d_u, const void* d_v, int dtype, int64_t N, int64_t n, int64_t m,
                                 double* d_out, void* stream) {
    ERT_REQUIRE(d_u && d_v && d_out && N > 0 && n > 0 && m > 0, "wasserstein_distance: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "wasserstein_distance: bad dtype");
    ERT_REQUIRE(n <= (1 << 24) && m <= (1 << 24), "wasserstein_distance: at most 2^24 values per sample");
    cudaStream_t st = (cudaStream_t)stream;
    int upad = 1, vpad = 1;
    while (upad < n) upad <<= 1;
    while (vpad < m) vpad <<= 1;
    // scratch: sorted u rows, sorted v, the merged values and the u-counts of every pair
    const size_t b_us = (size_t)N * upad * 8, b_vs = (size_t)vpad * 8, b_mg = (size_t)N * (n + m) * 8, b_cu = (size_t)N * (n + m) * 4;
    char* ws = nullptr;
    if (int rc = workspace(b_us + b_vs + b_mg + b_cu, (void**)&ws)) return rc;
    double* us = (double*)ws;
    double* vs = (double*)(ws + b_us);
    double* mg = (double*)(ws + b_us + b_vs);
    int* cu = (int*)(ws + b_us + b_vs + b_mg);
    int rc = dtype == ERTDIFF_F32 ? launch_sort_rows<float>(d_u, N, n, (int)n, upad, us, st) : launch_sort_rows<double>(d_u, N, n, (int)n, upad, us, st);
    if (rc) return rc;
    rc = dtype == ERTDIFF_F32 ? launch_sort_rows<float>(d_v, 1, m, (int)m, vpad, vs, st) : launch_sort_rows<double>(d_v, 1, m, (int)m, vpad, vs, st);
    if (rc) return rc;
    k_wasserstein<<<(unsigned)N, 1024, 0, st>>>(us, (int)n, upad, vs, (int)m, mg, cu, d_out);
    ERT_LAUNCH_CHECK("k_wasserstein");
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
