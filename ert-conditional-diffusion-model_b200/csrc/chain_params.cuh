// Host-visible parameter blocks of the chain kernels and the launch interface between the
// translation units of libertdiff_b200.so (capi.cu orchestrates; chain_fp32.cu, chain_umma.cu,
// encoder.cu and stats.cu hold the kernels, so that they compile in parallel).
#pragma once
#include "common.cuh"

struct ertdiff_model;

namespace ertdiff {

// The ten round keys of a Philox stream depend only on the seed: the host expands them once and
// they travel as kernel parameters, so every round reads its key straight from the constant bank.
struct PhiloxKeys {
    uint32_t k[20];
};
inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys ks;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        ks.k[2 * r] = k0; ks.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return ks;
}

struct ChainParams {
    int64_t B;             // members handled by this launch
    int64_t n_cond;
    int S;                 // chain length (num_steps); draw index of step t is S - t
    int t_hi;              // first timestep of this launch (S-1 for the whole chain)
    int t_count;           // steps in this launch
    const float* w0xT;     // (32,H)
    const float* w2p;      // (32,H)
    const float* b2p;      // (32)
    const float* table;    // (S,H)  c_t
    const float* coef;     // (S,4)
    const float* cond_bias;// (n_cond,H)
    const float* x_in;     // (B rows, P) or nullptr -> Philox draw 0
    int64_t x_in_stride;   // elements between rows of x_in
    const float* noise;    // (S-1, noise_B, P) or nullptr -> Philox
    int64_t noise_B;
    PhiloxKeys keys;       // round keys of the device RNG stream (make_philox_keys(seed))
    uint64_t offset;
    int64_t member_offset;
    float* x_out;          // (B,P)
    float* eps_trace;      // (S,B,P) indexed by t, or nullptr
    int P;
};

struct UmmaChainExtra {
    const uint4* w1_pk;     // bf16 B operand of GEMM1 (W0x augmented), H x 32, followed by the tile of its bf16 residuals
    const uint4* w2_pk;     // bf16 B operand of GEMM2 (W2 padded), 32 x H, followed by the tile of its bf16 residuals
    int* status;            // [0] = 1 when an mbarrier wait timed out
    long long* timing;      // optional (16 int64): phase cycle sums of CTA 0, see ertdiff_debug_umma_timing
    int mpc;                // members per CTA: 32, 64 or 128 rows of the 128-row tile are in use, so that a
                            // mid-size ensemble spreads over all SMs; the unused rows' warps only keep the barriers' counts
};

constexpr int UC_M = 128;       // members per CTA tile of the tensor-core chain
constexpr int UC_K1 = 32;       // padded param_dim + 3 augmentation columns
constexpr int UC_N2 = 32;       // padded param_dim

// ---- chain_fp32.cu -----------------------------------------------------------------------------
// members per CTA / hidden units per thread the launcher uses for (B, H) (env overrides included)
void chain_fp32_tiling(int64_t B, int H, int* mpb, int* upt);
// id of the kernel-variant selection in force (tuning env vars): part of the graph-mode cache key
int chain_variant_id();
int launch_chain_fp32(int H, const ChainParams& p, int mpb, int upt, cudaStream_t st);
// the same launch with the arithmetic removed (barriers, staging, shuffles and the dependent
// shared-memory hand-offs kept): the measured latency floor of the kernel's structure
int launch_chain_fp32_floor(int H, const ChainParams& p, int mpb, int upt, cudaStream_t st);

// ---- chain_umma.cu -----------------------------------------------------------------------------
bool chain_umma_supported(int H, int P);
int chain_umma_mpc(int64_t B);
bool chain_umma_split_supported(int H, int P);      // the split-precision build (ERTDIFF_PREC_BF16X3)
int launch_chain_umma(int H, const ChainParams& q, UmmaChainExtra ex, bool split, cudaStream_t st);
// (each packed buffer holds two tiles: the bf16 operand and its bf16 residual)
int pack_chain_umma_weights(int H, const float* w0xT, const float* w2p, int P, unsigned short* w1_pk,
                            unsigned short* w2_pk, cudaStream_t st);
int umma_selftest(const float* A, const float* B, int N, int K, float* D, cudaStream_t st);

// ---- encoder.cu --------------------------------------------------------------------------------
int run_encoder(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L, int64_t member_stride,
                float* d_cond_emb, float* d_cond_bias, cudaStream_t st);
int run_encoder_umma(ertdiff_model* m, const float* d_cond, int64_t n_cond, int64_t L, int64_t member_stride,
                     float* d_cond_emb, float* d_cond_bias, cudaStream_t st);
int pack_encoder_umma_weights(ertdiff_model* m, cudaStream_t st);
size_t encoder_umma_w1_bytes();
size_t encoder_umma_w2_bytes();

}  // namespace ertdiff
