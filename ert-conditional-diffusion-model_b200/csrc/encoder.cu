// Condition encoder launches: fp32 CUDA-core kernels (encoder.cuh) and the tcgen05/TMA kernel
// (encoder_umma.cuh).
#include <cstdlib>

#include "model.cuh"
#include "encoder.cuh"
#include "encoder_umma.cuh"

namespace ertdiff {

size_t encoder_umma_w1_bytes() { return (size_t)kConv1Out * EU_K1 * 2; }
size_t encoder_umma_w2_bytes() { return (size_t)kConv2Out * EU_K2 * 2; }

int pack_encoder_umma_weights(ertdiff_model* m, cudaStream_t st) {
    k_pack_encoder_umma<<<(kConv2Out * EU_K2 + 255) / 256, 256, 0, st>>>(m->raw[0], m->raw[2], m->enc_w1_pk, m->enc_w2_pk);
    ERT_LAUNCH_CHECK("k_pack_encoder_umma");
    return 0;
}

// ---- encoder ----------------------------------------------------------------------------
int run_encoder(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L,
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
int run_encoder_umma(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L,
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

}  // namespace ertdiff
