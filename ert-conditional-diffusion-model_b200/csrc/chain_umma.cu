// Tensor-core reverse chain (k_chain_umma, chain_umma.cuh): variant selection and launch; the
// tcgen05 self-test GEMM.
#include <cstdlib>

#include "chain_umma.cuh"

namespace ertdiff {

bool chain_umma_supported(int H, int P) { return (H == 128 || H == 256) && P <= UC_AUG; }

// rows of the tensor-core chain's 128-row tile that carry members (ERTDIFF_UMMA_MPC overrides: 32, 64, 128)
int chain_umma_mpc(int64_t B) {
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

int pack_chain_umma_weights(int H, const float* w0xT, const float* w2p, int P, unsigned short* w1_pk,
                            unsigned short* w2_pk, cudaStream_t st) {
    k_pack_umma_weights<<<(H * UC_K1 + 255) / 256, 256, 0, st>>>(w0xT, w2p, P, H, w1_pk, w2_pk);
    ERT_LAUNCH_CHECK("k_pack_umma_weights");
    return 0;
}

using Kern = void (*)(const ChainParams, const UmmaChainExtra);

template <int H, int CTAS, bool SPLIT = false>
static Kern pick_kernel(int variant) {
    static const Kern kerns[8] = {
        k_chain_umma<H, false, false, false, CTAS, SPLIT>, k_chain_umma<H, false, false, true, CTAS, SPLIT>,
        k_chain_umma<H, false, true, false, CTAS, SPLIT>,  k_chain_umma<H, false, true, true, CTAS, SPLIT>,
        k_chain_umma<H, true, false, false, CTAS, SPLIT>,  k_chain_umma<H, true, false, true, CTAS, SPLIT>,
        k_chain_umma<H, true, true, false, CTAS, SPLIT>,   k_chain_umma<H, true, true, true, CTAS, SPLIT>};
    return kerns[variant];
}

bool chain_umma_split_supported(int H, int P) { return H == 128 && P <= UC_AUG; }

int launch_chain_umma(int H, const ChainParams& q, UmmaChainExtra ex, bool split, cudaStream_t st) {
    if (split && !chain_umma_split_supported(H, q.P))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "sample_chain: the split-precision tensor-core chain is built for hidden_dim 128, param_dim <= 29");
    if (!chain_umma_supported(H, q.P))
        return fail(ERTDIFF_ERR_UNSUPPORTED, "sample_chain: the bf16 tensor-core chain is built for hidden_dim 128 or 256, param_dim <= 29");
    // members per CTA: the fewest rows per tile that still fit one wave of CTAs, so that a mid-size
    // ensemble runs on all SMs (the step is latency-bound: a part-filled tile steps faster)
    ex.mpc = chain_umma_mpc(q.B);
    const unsigned grid = (unsigned)((q.B + ex.mpc - 1) / ex.mpc);
    // more tiles than SMs: the build that keeps two CTAs resident per SM (H = 128 only: at H = 256 one CTA's
    // operands take 172 KB of shared memory)
    // (the split-precision build holds two tiles per operand: 177 KB, one CTA per SM)
    const bool two = H == 128 && !split && grid > (unsigned)kNumSMs && !std::getenv("ERTDIFF_UMMA_ONE_CTA");
    const int variant = (q.noise != nullptr ? 4 : 0) | (q.eps_trace != nullptr ? 2 : 0) | (q.n_cond == 1 ? 1 : 0);
    Kern k;
    size_t smem;
    if (split) { k = pick_kernel<128, 1, true>(variant); smem = sizeof(UmmaChainSmem<uc_nslot(1), 128, 2>); }
    else if (H == 256) { k = pick_kernel<256, 1>(variant); smem = sizeof(UmmaChainSmem<uc_nslot(1), 256>); }
    else if (two) { k = pick_kernel<128, 2>(variant); smem = sizeof(UmmaChainSmem<uc_nslot(2), 128>); }
    else { k = pick_kernel<128, 1>(variant); smem = sizeof(UmmaChainSmem<uc_nslot(1), 128>); }
    static PerDeviceOnce once[4][8];
    bool& attr_set = *once[split ? 3 : (H == 256 ? 2 : (two ? 1 : 0))][variant].slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    k<<<grid, UC_THREADS, smem, st>>>(q, ex);
    ERT_LAUNCH_CHECK("k_chain_umma");
    return 0;
}

template <int N, int K>
static int run_umma_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
    int* d_status = nullptr;                       // (a debug entry point: its own small allocation per call)
    ERT_CUDA(cudaMalloc(&d_status, sizeof(int)));
    ERT_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    const size_t smem = umma::tile_bytes(128, K) + umma::tile_bytes(N, K);
    ERT_CUDA(cudaFuncSetAttribute(umma::k_umma_selftest<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma::k_umma_selftest<N, K><<<1, 128, smem, st>>>(A, B, D, d_status);
    ERT_LAUNCH_CHECK("k_umma_selftest");
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d_status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_status);
    if (e != cudaSuccess) return fail(ERTDIFF_ERR_CUDA, std::string("umma selftest: ") + cudaGetErrorString(e));
    if (h) return fail(ERTDIFF_ERR_CUDA, "umma selftest: mbarrier wait timed out (MMA never completed)");
    return 0;
}

int umma_selftest(const float* A, const float* B, int N, int K, float* D, cudaStream_t st) {
    if (N == 128 && K == 32) return run_umma_selftest<128, 32>(A, B, D, st);
    if (N == 32 && K == 128) return run_umma_selftest<32, 128>(A, B, D, st);
    if (N == 128 && K == 128) return run_umma_selftest<128, 128>(A, B, D, st);
    if (N == 64 && K == 96) return run_umma_selftest<64, 96>(A, B, D, st);
    if (N == 256 && K == 32) return run_umma_selftest<256, 32>(A, B, D, st);
    if (N == 32 && K == 256) return run_umma_selftest<32, 256>(A, B, D, st);
    return fail(ERTDIFF_ERR_UNSUPPORTED, "debug_umma_gemm: (N,K) must be (128,32), (32,128), (128,128), (64,96), (256,32) or (32,256)");
}

}  // namespace ertdiff
