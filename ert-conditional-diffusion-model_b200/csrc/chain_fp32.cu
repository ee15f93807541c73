// fp32 CUDA-core reverse chain (k_chain, denoiser.cuh): kernel-variant selection and launch.
#include <cstdlib>

#include "denoiser.cuh"

namespace ertdiff {

static int env_int(const char* name) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : 0;
}

// Tiling of the fp32 chain, from the sweep in profiles/r02_chain_fp32_variants.md (one B200, T = 1000):
//   hidden units per thread: 2 wherever that build exists (hidden_dim 128 / 256).  A member then occupies half
//     the warps (two at hidden 128), layer 2 needs one shuffle level less and a step's barriers synchronise two
//     warps instead of four: 256 members 0.356 -> 0.268 us per step, 1024 members 1.23 -> 0.70.
//   members per CTA: one while the ensemble is at most ~1.5 waves of co-resident CTAs (a CTA runs its members one
//     after the other inside a step, so a second member costs almost a second step: 0.26 -> 0.46 us), two beyond;
//     the two-units build holds 4 (hidden 128) or 2 (hidden 256) CTAs per SM in registers.
void chain_fp32_tiling(int64_t B, int H, int* mpb_out, int* upt_out) {
    const bool two_ok = (H == 128 || H == 256);
    int upt = two_ok ? 2 : 1;
    if (const int v = env_int("ERTDIFF_CHAIN_UPT"))
        if (v == 1 || (v == 2 && two_ok)) upt = v;
    int mpb;
    const int v = env_int("ERTDIFF_CHAIN_MPB");
    if (v == 1 || v == 2 || v == 4 || v == 8) {
        mpb = v;
    } else if (upt == 2) {
        // (measured: one member per CTA wins up to ~1.5 waves -- 700 members: 0.63 vs 0.68 us -- two members per CTA
        // beyond -- 1184: 0.70 vs 0.76, 8192: 4.74 vs 5.14; four never do)
        const int64_t slots = (int64_t)kNumSMs * (H == 128 ? 4 : 2);
        mpb = 2 * B <= 3 * slots ? 1 : 2;
    } else {
        const int64_t ctas_per_sm = (H <= 128) ? 4 : (H <= 256 ? 2 : 1);
        const int64_t slots = kNumSMs * ctas_per_sm;
        // the one-member variant trades registers for latency (see k_chain): 3 CTAs per SM at H <= 128
        if (B <= kNumSMs * ((H <= 128) ? 3 : ctas_per_sm)) mpb = 1;
        else if (B <= 2 * slots) mpb = 2;
        else if (B <= 4 * slots) mpb = 4;
        else mpb = 8;
    }
    if (H >= 512 && mpb > 4) mpb = 4;            // static shared memory budget
    if (upt == 2 && mpb > 4) mpb = 4;
    *mpb_out = mpb; *upt_out = upt;
}

int chain_variant_id() {
    return env_int("ERTDIFF_CHAIN_UPT") * 1000 + env_int("ERTDIFF_CHAIN_MPB") * 100 + env_int("ERTDIFF_UMMA_MPC") +
           (std::getenv("ERTDIFF_UMMA_ONE_CTA") ? 50000 : 0);
}

template <int H, int MPB, int UPT, bool FLOOR>
static void launch_hmu(const ChainParams& p, unsigned grid, cudaStream_t st) {
    const bool replay = p.noise != nullptr, trace = p.eps_trace != nullptr;
    constexpr int NT = H / UPT;
    if (FLOOR) {       // the floor build exists for the production data path only (device RNG, no trace)
        k_chain<H, MPB, UPT, false, false, FLOOR><<<grid, NT, 0, st>>>(p);
        return;
    }
    if (replay) {
        if (trace) k_chain<H, MPB, UPT, true, true, false><<<grid, NT, 0, st>>>(p);
        else k_chain<H, MPB, UPT, true, false, false><<<grid, NT, 0, st>>>(p);
    } else {
        if (trace) k_chain<H, MPB, UPT, false, true, false><<<grid, NT, 0, st>>>(p);
        else k_chain<H, MPB, UPT, false, false, false><<<grid, NT, 0, st>>>(p);
    }
}

template <int H, bool FLOOR>
static int launch_h(const ChainParams& p, int mpb, int upt, cudaStream_t st) {
    const unsigned grid = (unsigned)((p.B + mpb - 1) / mpb);
    constexpr bool kTwo = (H == 128 || H == 256);        // builds with two hidden units per thread
    if (upt == 2 && kTwo) {
        constexpr int H2 = kTwo ? H : 128;
        if (mpb == 1) launch_hmu<H2, 1, 2, FLOOR>(p, grid, st);
        else if (mpb == 2) launch_hmu<H2, 2, 2, FLOOR>(p, grid, st);
        else if (mpb == 4 && !FLOOR) launch_hmu<H2, 4, 2, false>(p, grid, st);
        else return fail(ERTDIFF_ERR_UNSUPPORTED, "k_chain: no such members-per-CTA build for two units per thread");
    } else if (upt != 1) {
        return fail(ERTDIFF_ERR_UNSUPPORTED, "k_chain: two units per thread need hidden_dim 128 or 256");
    } else if (FLOOR) {
        if (mpb == 1) launch_hmu<H, 1, 1, FLOOR>(p, grid, st);
        else if (mpb == 2) launch_hmu<H, 2, 1, FLOOR>(p, grid, st);
        else return fail(ERTDIFF_ERR_UNSUPPORTED, "k_chain floor: built for 1 or 2 members per CTA");
    } else {
        if (mpb == 1) launch_hmu<H, 1, 1, false>(p, grid, st);
        else if (mpb == 2) launch_hmu<H, 2, 1, false>(p, grid, st);
        else if (mpb == 4 || H >= 512) launch_hmu<H, 4, 1, false>(p, grid, st);
        else launch_hmu<H, (H >= 512 ? 4 : 8), 1, false>(p, grid, st);
    }
    ERT_LAUNCH_CHECK("k_chain");
    return 0;
}

int launch_chain_fp32(int H, const ChainParams& p, int mpb, int upt, cudaStream_t st) {
    switch (H) {
        case 32: return launch_h<32, false>(p, mpb, upt, st);
        case 64: return launch_h<64, false>(p, mpb, upt, st);
        case 128: return launch_h<128, false>(p, mpb, upt, st);
        case 256: return launch_h<256, false>(p, mpb, upt, st);
        case 512: return launch_h<512, false>(p, mpb, upt, st);
    }
    return fail(ERTDIFF_ERR_UNSUPPORTED, "hidden_dim must be one of 32,64,128,256,512");
}

int launch_chain_fp32_floor(int H, const ChainParams& p, int mpb, int upt, cudaStream_t st) {
    if (p.noise || p.eps_trace) return fail(ERTDIFF_ERR_UNSUPPORTED, "k_chain floor: device RNG, no trace");
    switch (H) {
        case 128: return launch_h<128, true>(p, mpb, upt, st);
        case 256: return launch_h<256, true>(p, mpb, upt, st);
    }
    return fail(ERTDIFF_ERR_UNSUPPORTED, "k_chain floor: built for hidden_dim 128 and 256");
}

}  // namespace ertdiff
