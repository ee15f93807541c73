// Tensor-core version of the reverse chain for large ensembles (precision = bf16 operands,
// fp32 accumulation; BASELINE configs 3/4): the two projections of the hoisted denoiser step run
// as tcgen05.mma with accumulators in TMEM, everything between them stays on chip.
//
// One CTA = a tile of up to 128 members for all steps, 13 warps in three roles, no __syncthreads in the
// step loop:
//   * 8 epilogue warps.  TMEM lane r <-> member r of the tile; a warp may only touch the lane
//     quarter (warp % 4), so warp w handles members 32*(w%4)..+31 and half = w/4 of the columns:
//     TWO threads share one member, each owning 64 hidden columns (epilogue 1) and 16 of the 32
//     padded parameters (epilogue 2, x).  With a part-filled tile (32 or 64 members per CTA, so that a
//     mid-size ensemble reaches every SM) the warps of the unused quarters retire at once.
//   * 4 noise warps.  They run ahead of the chain and fill a shared-memory ring (4 deep; 2 in the
//     two-CTAs-per-SM build) with the N(0,1) draws of the coming steps (Philox4x32-10 + Box-Muller, or
//     the caller's replayed noise), one thread per member and item; with a part-filled tile the warps
//     split into groups working on different items.  The generator is bound by the MUFU and
//     integer-multiply pipes, the epilogues by conversion / packed-fp32 / shared-memory work: separate
//     warps let the two overlap instead of alternating.
//   * 1 MMA-issue warp (one elected thread); it also prefetches the coming rows of the c_t table into L2.
//
//   GEMM1  D[128 members x 128 hidden] = Xaug[128 x 32] * W1aug[128 x 32]^T           (2 MMAs, K=16)
//          Xaug row     = [x_0..x_28, 1, 1, 1]   (bf16, rewritten by its two threads every step)
//          W1aug row j  = [W0x[j][0..28], v_hi[j], v_mid[j], v_lo[j]]
//          v = c_t (per-step vector), or c_t + c_b when all members share one condition; split
//          into three bf16 terms (exact to 2^-24) and folded into the contraction through the
//          three spare K columns -- the "embedding add" happens inside the MMA
//   epi 1  h = ReLU(D [+ c_b[member], fp32, parked in TMEM columns 256..383, only with distinct
//          conditions]) -> bf16 (cvt.rn.relu.bf16x2) into the K-major A operand of GEMM2
//   GEMM2  E[128 members x 32] = Hbf16[128 x 128] * W2pad[32 x 128]^T                  (8 MMAs, K=16)
//          issued per column half as soon as that half's 64 K-columns of H are written
//   epi 2  eps = E + b2; posterior update (packed fp32) of the thread's 16 parameters (x stays in fp32
//          registers for the whole chain) with the ring's noise, new Xaug chunks
// Hand-offs.  The two that a tcgen05.commit signals are mbarriers; all others are hardware named
// barriers (bar.arrive by the producing role, bar.sync by the consuming one), because a warp parked
// on a named barrier costs no issue slots, while every mbarrier.try_wait sleeper was woken by every
// arrival in the SM and polled again (ncu: a fifth of all issued instructions with mbarriers only):
//   NB_X      Xaug / W1aug of the next step written    epilogue (arrive) -> MMA warp (sync)
//   bar_d     (tcgen05.commit)   D complete            MMA warp -> epilogue (mbarrier wait)
//   NB_H+h    H columns of half h written              epilogue (arrive) -> MMA warp (sync)
//   bar_e     (tcgen05.commit)   E complete            MMA warp -> epilogue (mbarrier wait)
//   NB_FULL+s / NB_EMPTY+s       noise ring slot s     noise warps <-> epilogue
// A named barrier has no time-out, so a role never leaves its loop early: should an MMA never complete,
// the epilogue warps stop waiting on the mbarriers but keep every named-barrier operation (the tile's
// output is then poisoned and the status word set).
// TMEM: D columns 0..127, E 128..159, c_b 256..383 (distinct conditions only).
//
// Split precision (template parameter SPLIT; precision = ERTDIFF_PREC_BF16X3; H = 128, one CTA per SM).  Every
// fp32 operand is carried as TWO bf16 terms, v = v_hi + v_lo with v_hi = bf16(v), v_lo = bf16(v - v_hi) (the
// residual is exact in fp32, so v_hi + v_lo carries 16-17 mantissa bits), and each projection becomes three
// accumulating products  A_hi B_hi + A_lo B_hi + A_hi B_lo  (the lo x lo term, 2^-18 of a product, is dropped):
// the same kernel with a second copy of every operand tile in shared memory and the split arithmetic (cvt,
// shift/mask, subtract, cvt per pair) in both epilogues.  The tensor cores then return the fp32 kernel's
// projections to ~5e-6 of their scale instead of bf16's 3e-3 -- DESIGN.md section 5 has the measured figures.
// What the step costs is shared-memory traffic (every MMA re-reads its A tile: ~44 cycles per K=16 slice of H)
// and epilogue issue slots, not tensor-pipe time, so:
//   * GEMM2 reads H_hi once: its B operand is [W2_hi ; W2_lo] as ONE 64-row tile (the residual tile sits right
//     behind the operand tile, which is exactly rows 32..63 of the canonical layout), giving E1 = H_hi W_hi and
//     E2 = H_hi W_lo in adjacent TMEM columns; H_lo W_hi accumulates into E1; epilogue 2 adds E1 + E2.
//     16 MMAs instead of 24.  (GEMM1 keeps three products into one accumulator: it is off the critical path.)
//   * H is handed to the MMA warp in four pieces (two per column part, named barriers 12/13 for the second
//     halves), so that most of GEMM2 runs under the rest of epilogue 1.
//   * h >= 0: hi = cvt.rz.relu (truncation), so that the residual of a positive h is never negative and
//     cvt.rn.relu of (h - hi) is both the residual's rounding and the ReLU of a negative h -- no max().
// Measured (18,944 members, T = 1000): 1.78 us per step against 1.24 (bf16) and 11.4 (fp32 kernel); epilogue 1 --
// twice the conversions and shared-memory stores -- is what the extra time is.
#pragma once
#include "denoiser.cuh"
#include "umma.cuh"

// -DUC_TIMING=1 builds the phase-timing instrumentation read by ertdiff_debug_umma_timing (it costs
// ~16 registers, so production builds leave it out and that entry point reports zeros)
#ifndef UC_TIMING
#define UC_TIMING 0
#endif
#ifndef UC_PREFETCH_TABLE
#define UC_PREFETCH_TABLE 1
#endif
#ifndef UC_RNG_ONEPASS
#define UC_RNG_ONEPASS 1
#endif
#if UC_TIMING
#define UC_T(...) __VA_ARGS__
#else
#define UC_T(...)
#endif

namespace ertdiff {

// (UC_M = 128 members per CTA tile, UC_K1 = 32 = padded param_dim + 3 augmentation columns, UC_N2 = 32 = padded
// param_dim: chain_params.cuh.)  The kernel is a template over hidden_dim H = 128 (the reference's default,
// ECD.py:287) and 256 (the reference's one expressible widening, ECD.py:123; BASELINE config 5).
#ifndef UC_TPM_N
#define UC_TPM_N 2
#endif
constexpr int UC_TPM = UC_TPM_N;           // epilogue threads per member: 2 or 4
constexpr int UC_EPI_WARPS = 4 * UC_TPM;
#ifndef UC_RNG_WARPS_N
#define UC_RNG_WARPS_N 4
#endif
constexpr int UC_RNG_WARPS = UC_RNG_WARPS_N;   // 4 or 8
constexpr int UC_THREADS = (UC_EPI_WARPS + UC_RNG_WARPS + 1) * 32;   // + the MMA-issue warp
constexpr int UC_AUG = 29;      // first augmentation column (param_dim <= 29)
// Two builds of the kernel (template parameter CTAS): one CTA per SM (up to 128 registers per thread, a
// 4-deep noise ring, all 64 accumulator columns of a thread in flight per tcgen05.wait::ld) and, for
// ensembles of more than 148 tiles, two co-resident CTAs per SM (72 registers,
// 2-deep ring, 32 columns per wait): each CTA is then ~17 % slower, but their MMA / mbarrier waits
// overlap -- 37,888 members: 2.76 -> 2.29 ms.
__host__ __device__ constexpr int uc_nslot(int ctas) { return ctas == 2 ? 2 : 4; }        // depth of the noise ring (steps)
__host__ __device__ constexpr int uc_epi_chunk(int ctas) { return ctas == 2 ? 32 : 64; }

// NT = bf16 terms per operand: 1, or 2 for the split-precision build ([hi tile | lo tile], each a complete
// canonical K-major tile)
template <int NSLOT, int H, int NT = 1>
struct UmmaChainSmem {
    unsigned char x[NT * UC_M * UC_K1 * 2];      // A of GEMM1
    unsigned char h[NT * UC_M * H * 2];          // A of GEMM2
    unsigned char w1[NT * H * UC_K1 * 2];        // B of GEMM1 (W0x augmented | its bf16 residual, augmentation columns 0)
    unsigned char w2[NT * UC_N2 * H * 2];        // B of GEMM2 (W2 padded | its bf16 residual)
    float zring[NSLOT][UC_M][kPPad];     // noise ring; 16-byte chunk c of member m sits at chunk c ^ (m & 7)
    unsigned long long bar_d, bar_e;
    alignas(16) float b2[kPPad];
    uint32_t tmem_slot;
    int timeout;
};

// pack the bf16 B operands once per load_state_dict: byte layout = umma::elem_offset; each buffer holds the tile
// of bf16(w) followed by the tile of the residuals bf16(w - bf16(w)) (read by the split-precision build only)
static __global__ void k_pack_umma_weights(const float* __restrict__ w0xT /*(32,H)*/,
                                           const float* __restrict__ w2p /*(32,H)*/, int P, int H,
                                           unsigned short* __restrict__ w1_pk, unsigned short* __restrict__ w2_pk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H * UC_K1) {
        const int j = i / UC_K1, k = i % UC_K1;
        const float v = (k < P) ? w0xT[k * H + j] : 0.f;      // columns P..31 start as zero
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        const __nv_bfloat16 r = __float2bfloat16_rn(v - __bfloat162float(b));
        w1_pk[umma::elem_offset(j, k, UC_K1) / 2] = *reinterpret_cast<const unsigned short*>(&b);
        w1_pk[H * UC_K1 + umma::elem_offset(j, k, UC_K1) / 2] = *reinterpret_cast<const unsigned short*>(&r);
    }
    if (i < UC_N2 * H) {
        const int p = i / H, k = i % H;
        const float v = w2p[p * H + k];
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        const __nv_bfloat16 r = __float2bfloat16_rn(v - __bfloat162float(b));
        w2_pk[umma::elem_offset(p, k, H) / 2] = *reinterpret_cast<const unsigned short*>(&b);
        w2_pk[UC_N2 * H + umma::elem_offset(p, k, H) / 2] = *reinterpret_cast<const unsigned short*>(&r);
    }
}

// named-barrier ids of the chain kernel (0 is __syncthreads)
constexpr uint32_t UC_NB_X = 1, UC_NB_H = 2, UC_NB_FULL = 4, UC_NB_EMPTY = 8, UC_NB_H2 = 12;   // H2: second halves (split build)
static_assert(UC_NB_H + UC_TPM <= UC_NB_FULL && UC_NB_EMPTY + 4 <= UC_NB_H2 && UC_NB_H2 + UC_TPM <= 16, "named barrier ids");
__device__ __forceinline__ void nb_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// (a, b) as two packed bf16 pairs: hi = bf16(a), bf16(b); lo = bf16 of the (exact) fp32 residuals
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = umma::pack_bf16(a, b);
    lo = umma::pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// the same of (max(a,0), max(b,0)): hi truncated toward zero (cvt.rz.relu), so the residual of a positive value is
// >= 0 and the .relu of the second conversion only ever clears the residual (= a itself) of a negative one -- no
// max().  (Measured alternative: hi as a byte permute of the fp32 words' upper halves after max(), which moves one
// conversion per pair from the XU pipe to the ALU pipe, is SLOWER -- epilogue 1: 1490 -> 1920 cycles per step --
// because the noise warps' Philox rounds already load the ALU pipe.)
__device__ __forceinline__ void split_bf16_pair_relu(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    lo = umma::pack_bf16_relu(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

template <int H, bool REPLAY, bool TRACE, bool SHARED, int CTAS, bool SPLIT = false>
__global__ void __launch_bounds__(UC_THREADS, CTAS) k_chain_umma(const ChainParams a, const UmmaChainExtra ex) {
    static_assert(CTAS == 1 || CTAS == 2, "one or two CTAs per SM");
    static_assert(H == 128 || (H == 256 && CTAS == 1), "hidden_dim 128, or 256 with one CTA per SM");
    static_assert(!SPLIT || (H == 128 && CTAS == 1), "split precision: hidden_dim 128, one CTA per SM");
    constexpr int UC_H = H;
    constexpr int NT = SPLIT ? 2 : 1;
    // byte offsets of the lo tiles behind the hi tiles
    constexpr uint32_t X_LO = UC_M * UC_K1 * 2, H_LO = UC_M * H * 2, W1_LO = H * UC_K1 * 2, W2_LO = UC_N2 * H * 2;
    constexpr int UC_NSLOT = uc_nslot(CTAS);
    constexpr int UC_EPI_CHUNK = uc_epi_chunk(CTAS);
    using namespace umma;
    extern __shared__ __align__(128) unsigned char uc_smem_raw[];
    UmmaChainSmem<UC_NSLOT, H, NT>& s = *reinterpret_cast<UmmaChainSmem<UC_NSLOT, H, NT>*>(uc_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sX_ = smem_u32(s.x), sH_ = smem_u32(s.h), sW1_ = smem_u32(s.w1), sW2_ = smem_u32(s.w2);
    const uint32_t sZ_ = smem_u32(&s.zring[0][0][0]);
    const uint32_t bar_d_ = smem_u32(&s.bar_d), bar_e_ = smem_u32(&s.bar_e);
    uint32_t sX = sX_, sH = sH_, sW1 = sW1_, sW2 = sW2_, sZ = sZ_, bar_d = bar_d_, bar_e = bar_e_;
    // opaque to the optimiser: otherwise every use re-derives the shared window base
    // (S2R SR_CgaCtaId + LEA, a long-scoreboard read) inside the step loop
    asm volatile("" : "+r"(sX), "+r"(sH), "+r"(sW1), "+r"(sW2), "+r"(sZ), "+r"(bar_d), "+r"(bar_e));
    // TMEM columns, H = 128: D 0..127; E 128..159; c_b (distinct conditions) 256..383.  Two CTAs per SM have 256
    // columns each: with distinct conditions E then aliases D's first 32 columns (they belong to column part 0,
    // and the part-0 GEMM2 issue waits for the part-0 threads' H_0 arrival, i.e. until they have consumed them)
    // and c_b moves to 128..255.  H = 256: D 0..255; E 256..287 with a shared condition; with distinct
    // conditions c_b takes 256..511 and E aliases D's first 32 columns in the same way.
    constexpr bool E_ALIAS = !SHARED && (CTAS == 2 || H == 256);
    constexpr uint32_t TMEM_COLS = (H == 256) ? 512 : ((SHARED || CTAS == 2) ? 256 : 512);
    constexpr uint32_t E_COL = E_ALIAS ? 0 : (uint32_t)H;
    constexpr uint32_t CB_COL = (H == 256) ? 256 : ((CTAS == 2) ? 128 : 256);
    constexpr uint32_t SLOT_BYTES = UC_M * kPPad * 4;

    // ---- one-time setup ------------------------------------------------------------------------
    {
        uint4* d1 = reinterpret_cast<uint4*>(s.w1);
        uint4* d2 = reinterpret_cast<uint4*>(s.w2);
        for (int i = tid; i < NT * UC_H * UC_K1 * 2 / 16; i += UC_THREADS) d1[i] = ex.w1_pk[i];
        for (int i = tid; i < NT * UC_N2 * UC_H * 2 / 16; i += UC_THREADS) d2[i] = ex.w2_pk[i];
        if (tid < kPPad) s.b2[tid] = a.b2p[tid];
        if (tid == 0) s.timeout = 0;
    }
    if (warp == 0) tmem_alloc(smem_u32(&s.tmem_slot), TMEM_COLS);
    if (tid == 0) {
        mbar_init(bar_d, 1);
        mbar_init(bar_e, 1);
        fence_mbar_init();
    }
    fence_proxy_async();          // the weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_slot;
    const int n_steps = a.t_count;
#if UC_PREFETCH_TABLE
    // the c_t rows of this launch (512 B per step, each read once per CTA) are pulled into L2 up front, the
    // CTAs sharing the lines between them: a table left in DRAM by whatever ran before (bench.py flushes L2
    // between steps) otherwise costs every step part of a DRAM round trip
    {
        const char* base = reinterpret_cast<const char*>(a.table + (int64_t)(a.t_hi - n_steps + 1) * UC_H);
        const int64_t n_lines = (int64_t)n_steps * UC_H * 4 / 128;
        for (int64_t line = (int64_t)blockIdx.x * UC_THREADS + tid; line < n_lines; line += (int64_t)gridDim.x * UC_THREADS)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + line * 128));
    }
#endif
    const int d_first = a.S - a.t_hi;
    const int mpc = ex.mpc;                               // rows of the tile that carry members: 32, 64 or 128
    const int64_t m0 = (int64_t)blockIdx.x * mpc;
    const int P = a.P;
    // ring items, in consumption order: [x_T when the launch starts a chain] then one per step with
    // t > 0 (steps run t = t_hi, t_hi-1, ...; the step with t == 0 adds no noise, ECD.py:115)
    const int first_item_step = a.x_in ? 0 : 1;          // ring item of step `it` = it + first_item_step
    const int n_noisy = n_steps < a.t_hi ? n_steps : a.t_hi;
    const int n_items = first_item_step + n_noisy;
    // participants of the named barriers: the epilogue threads whose rows carry members, + the other role
    const uint32_t n_epi = (uint32_t)(UC_TPM * mpc);
    const uint32_t cnt_x = n_epi + 32u, cnt_h = (uint32_t)mpc + 32u, cnt_ring = n_epi + (uint32_t)mpc;   // + the noise warps of one item
    bool ok = true;

    if (warp == UC_EPI_WARPS + UC_RNG_WARPS) {
        // ===== MMA-issue warp: the whole warp walks the loop (converged), one elected lane issues ==========
        constexpr uint32_t IDESC1 = idesc_bf16_f32(UC_M, UC_H);
        constexpr uint32_t IDESC2 = idesc_bf16_f32(UC_M, UC_N2);
        constexpr uint32_t IDESC2W = idesc_bf16_f32(UC_M, 2 * UC_N2);     // split build: B = [W2_hi ; W2_lo], 64 rows
        // all operand descriptors are loop-invariant: build them once, so that a step costs this
        // (single, latency-bound) thread little more than the ten MMA issues themselves
        uint64_t dA1[UC_K1 / 16], dB1[UC_K1 / 16], dA2[UC_H / 16], dB2[UC_H / 16];
#pragma unroll
        for (int k = 0; k < UC_K1 / 16; ++k) {
            dA1[k] = smem_desc(sX + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1));
            dB1[k] = smem_desc(sW1 + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1));
        }
#pragma unroll
        for (int k = 0; k < UC_H / 16; ++k) {
            dA2[k] = smem_desc(sH + 2 * k * kLBO, kLBO, sbo_bytes(UC_H));
            dB2[k] = smem_desc(sW2 + 2 * k * kLBO, kLBO, sbo_bytes(UC_H));
        }
        const uint32_t tmemE = tmem + E_COL;
#if UC_TIMING
        const bool timed = ex.timing != nullptr && blockIdx.x == 0 && lane == 0;
        long long tm[4] = {0, 0, 0, 0}, c0 = 0, c1 = 0;
#endif
        // this warp has slack: it pulls the per-step rows of the c_t table (512 B each, read by every CTA one
        // step before use) and of the step scalars into L2 several steps ahead, so that a cold table costs the
        // epilogue threads an L2 hit, not a DRAM round trip inside the step
        constexpr int PF_AHEAD = 8;
        auto prefetch_rows = [&](int t) {
            if (t >= 0 && lane < UC_H / 32 + 1) {
                const void* ptr = lane < UC_H / 32 ? (const void*)(a.table + (int64_t)t * UC_H + lane * 32)
                                           : (const void*)(reinterpret_cast<const float4*>(a.coef) + t);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        };
        for (int d = 1; d < PF_AHEAD; ++d) prefetch_rows(a.t_hi - d);
#pragma unroll 1
        for (int it = 0; it < n_steps; ++it) {
            UC_T(if (timed) c0 = clock64();)
            prefetch_rows(a.t_hi - it - PF_AHEAD);
            nb_sync(UC_NB_X, cnt_x);
            UC_T(if (timed) { c1 = clock64(); tm[0] += c1 - c0; })
            tc_fence_after();
            if (elect_one()) {
                mma_bf16_first(tmem, dA1[0], dB1[0], IDESC1);
#pragma unroll
                for (int k = 1; k < UC_K1 / 16; ++k) mma_bf16_acc(tmem, dA1[k], dB1[k], IDESC1);
                if (SPLIT) {       // + x_lo W_hi + x_hi W_lo (the start-address field of a descriptor counts 16-byte units)
#pragma unroll
                    for (int k = 0; k < UC_K1 / 16; ++k) {
                        mma_bf16_acc(tmem, dA1[k] + (X_LO >> 4), dB1[k], IDESC1);
                        mma_bf16_acc(tmem, dA1[k], dB1[k] + (W1_LO >> 4), IDESC1);
                    }
                }
                mma_commit(bar_d);
            }
            __syncwarp();
            UC_T(if (timed) { c0 = clock64(); tm[1] += c0 - c1; })
            // each part's K columns as soon as they are written (split build: each half of each part)
            constexpr int NSUB = SPLIT ? 2 : 1, KPP = UC_H / 16 / UC_TPM, KPS = KPP / NSUB;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                for (int part = 0; part < UC_TPM; ++part) {
                    nb_sync((sub == 0 ? UC_NB_H : UC_NB_H2) + part, cnt_h);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < KPS; ++kk) {
                            const int k = KPP * part + KPS * sub + kk;
                            if (SPLIT) {      // [E1 | E2] (+)= H_hi [W_hi ; W_lo]^T,  E1 += H_lo W_hi^T
                                if (k == 0) mma_bf16_first(tmemE, dA2[0], dB2[0], IDESC2W);
                                else mma_bf16_acc(tmemE, dA2[k], dB2[k], IDESC2W);
                                mma_bf16_acc(tmemE, dA2[k] + (H_LO >> 4), dB2[k], IDESC2);
                            } else {
                                if (k == 0) mma_bf16_first(tmemE, dA2[0], dB2[0], IDESC2);
                                else mma_bf16_acc(tmemE, dA2[k], dB2[k], IDESC2);
                            }
                        }
                        if (sub == NSUB - 1 && part == UC_TPM - 1) mma_commit(bar_e);
                    }
                    __syncwarp();
                }
            }
            UC_T(if (timed) { c1 = clock64(); tm[2] += c1 - c0; })
        }
        UC_T(if (timed) { ex.timing[8] = tm[0]; ex.timing[9] = tm[1]; ex.timing[10] = tm[2]; })
    } else if (warp < UC_RNG_WARPS) {
        // ===== noise warps: a thread produces all 32 draws of one member for one ring item.  A tile with
        // fewer than 128 members needs fewer than 4 warps per item, so the warps split into groups that work on
        // different items at the same time: an item's latency (two dependent Philox + Box-Muller passes,
        // ~2000 cycles for a lone warp per scheduler) would otherwise bound the step of a part-filled tile ======
        const int G = mpc >> 5;                               // warps per item: 1, 2 or 4
        // a slot must belong to one group (two groups parked on the same EMPTY barrier would mix their counts):
        // at most UC_NSLOT groups; further warps stay idle
        const int n_groups = (UC_RNG_WARPS / G < UC_NSLOT) ? UC_RNG_WARPS / G : UC_NSLOT;
        const int grp = warp / G;
        const int m = (warp % G) * 32 + lane;
        const int64_t mg = (m0 + m) < a.B ? (m0 + m) : (a.B - 1);
        const int64_t gmember = a.member_offset + mg;
        const uint32_t zrow = sZ + (uint32_t)m * (kPPad * 4);
#if UC_TIMING
        const bool timed = ex.timing != nullptr && blockIdx.x == 0 && tid == 0;
        long long tg[2] = {0, 0}, g0 = 0, g1 = 0;
#endif
#pragma unroll 1
        for (int item = grp < n_groups ? grp : n_items; item < n_items; item += n_groups) {
            const int slot = item % UC_NSLOT;
            UC_T(if (timed) g0 = clock64();)
            if (item >= UC_NSLOT) nb_sync(UC_NB_EMPTY + slot, cnt_ring);     // the slot's previous item was consumed
            UC_T(if (timed) { g1 = clock64(); tg[0] += g1 - g0; })
            const uint32_t draw = (item < first_item_step) ? 0u : (uint32_t)(d_first + item - first_item_step);
            if (!REPLAY && CTAS == 1 && UC_RNG_ONEPASS) {
                // one pass: eight Philox chains in flight (the one-CTA build has the registers for it)
                float z[32];
                philox_normal32(a.keys, a.offset, gmember, draw, z);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 v = make_float4(z[4 * c], z[4 * c + 1], z[4 * c + 2], z[4 * c + 3]);
                    if (4 * c + 3 >= P) { v.x = (4 * c < P) ? v.x : 0.f; v.y = (4 * c + 1 < P) ? v.y : 0.f;
                                          v.z = (4 * c + 2 < P) ? v.z : 0.f; v.w = 0.f; }
                    sts128(zrow + (uint32_t)slot * SLOT_BYTES + (uint32_t)((c ^ (m & 7)) * 16), v);
                }
            } else {
#pragma unroll 1
            for (int u = 0; u < 4; u += 2) {                  // 16 draws per pass: four Philox chains in flight
                float z[16];
                if (REPLAY) {
                    const float* zr = a.noise + ((int64_t)(draw - 1) * a.noise_B + mg) * P + 8 * u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = (8 * u + i < P) ? zr[i] : 0.f;
                } else {
                    philox_normal8(a.keys, a.offset, gmember, draw, 2 * u, &z[0]);
                    philox_normal8(a.keys, a.offset, gmember, draw, 2 * u + 2, &z[8]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = (8 * u + i < P) ? z[i] : 0.f;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    sts128(zrow + (uint32_t)slot * SLOT_BYTES + (uint32_t)(((2 * u + c) ^ (m & 7)) * 16),
                           make_float4(z[4 * c], z[4 * c + 1], z[4 * c + 2], z[4 * c + 3]));
            }
            }
            nb_arrive(UC_NB_FULL + slot, cnt_ring);
            UC_T(if (timed) tg[1] += clock64() - g1;)
        }
        UC_T(if (timed) { ex.timing[11] = tg[0]; ex.timing[12] = tg[1]; })
    } else if ((warp & 3) * 32 < mpc) {
        // ===== epilogue warps (those whose TMEM lane quarter carries members) ==========================
        constexpr int CW = UC_H / UC_TPM;        // hidden columns per thread
        constexpr int PW = kPPad / UC_TPM;       // parameters per thread
        const int et = tid - UC_RNG_WARPS * 32;  // 0 .. 128*UC_TPM-1
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access (warp % 4)
        const int part = et >> 7;                // hidden columns CW*part..+CW-1, parameters PW*part..+PW-1
        const int row = quarter * 32 + lane;     // member of the tile = TMEM lane
        const bool mvalid = (m0 + row) < a.B;
        const int64_t mg = mvalid ? (m0 + row) : (a.B - 1);
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        const uint32_t tD = tlane + CW * part;           // this thread's hidden columns of D
        const uint32_t tE = tlane + E_COL + PW * part;   // this thread's parameter columns of E
        const uint32_t tCB = tlane + CB_COL + CW * part; // c_b of this member (distinct conditions)
        // H columns written so far -> the MMA warp (split build: the first half of this thread's columns early)
        auto h_arrive = [&](uint32_t base) {
            fence_proxy_async();
            tc_fence_before();
            nb_arrive(base + (uint32_t)part, cnt_h);
        };
        const uint32_t zrow = sZ + (uint32_t)row * (kPPad * 4);
        const uint32_t zsw = (uint32_t)(row & 7);

        // The working epilogue threads share the augmentation columns of W1aug: v[j] = c_t[t][j] (+ c_b[j] of
        // the shared condition) as its 3-term bf16 split (v_hi | v_mid v_lo) against the three constant-one
        // columns of Xaug.  Working thread wt = part * mpc + row owns rows j = wt and wt + 2 mpc (< 128): one
        // row per part-0 thread of a full tile, two rows per thread of a quarter-filled one.
        const int wt = part * mpc + row, aug_stride = UC_TPM * mpc;
        const bool aug_owner = wt < UC_H;
        constexpr int NAUG = UC_H / 64;     // W1aug rows per working thread, at most (quarter-filled tile)
        float cb0[NAUG] = {};
        if (aug_owner && SHARED) {
#pragma unroll
            for (int i = 0; i < NAUG; ++i)
                if (wt + i * aug_stride < UC_H) cb0[i] = a.cond_bias[wt + i * aug_stride];
        }
        const uint32_t w1aug = sW1 + elem_offset(wt & (UC_H - 1), UC_AUG, UC_K1);
        const uint32_t w1aug_pitch = (uint32_t)(aug_stride / 8) * sbo_bytes(UC_K1);      // aug_stride rows further down
        const float* ctcol = a.table + wt;
        auto load_ct = [&](int t, float (&ct)[NAUG]) {
#pragma unroll
            for (int i = 0; i < NAUG; ++i)
                if (wt + i * aug_stride < UC_H) ct[i] = __ldg(ctcol + (int64_t)t * UC_H + i * aug_stride);
        };
        auto refresh_w1aug = [&](const float (&ct)[NAUG]) {
#pragma unroll
            for (int i = 0; i < NAUG; ++i) {
                if (wt + i * aug_stride < UC_H) {
                    const float v = ct[i] + cb0[i];
                    const float v_hi = bf16_round(v);
                    const float r1 = v - v_hi;                    // exact
                    const float v_mid = bf16_round(r1);
                    const float v_lo = r1 - v_mid;                // exact; rounded to bf16 by the pack
                    const __nv_bfloat16 hb = __float2bfloat16_rn(v_hi);
                    const uint32_t dst = w1aug + (uint32_t)i * w1aug_pitch;
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(*reinterpret_cast<const unsigned short*>(&hb)) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 2u), "r"(pack_bf16(v_mid, v_lo)) : "memory");
                }
            }
        };

        if (!SHARED) {   // park c_b of this member's hidden columns in TMEM for the whole chain
            const float4* cbrow = reinterpret_cast<const float4*>(a.cond_bias + (mg % a.n_cond) * UC_H + CW * part);
#pragma unroll 1
            for (int c = 0; c < CW / 32; ++c) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 f = cbrow[8 * c + i];
                    v[4 * i] = __float_as_uint(f.x); v[4 * i + 1] = __float_as_uint(f.y);
                    v[4 * i + 2] = __float_as_uint(f.z); v[4 * i + 3] = __float_as_uint(f.w);
                }
                tmem_st32(tCB + 32 * c, v);
            }
            tmem_st_wait();
        }
        uint32_t b2a = smem_u32(&s.b2[PW * part]);
        asm volatile("" : "+r"(b2a));

        // one ring item: the thread's PW draws, 4 at a time
        auto ring_wait = [&](int item) -> uint32_t {
            const int slot = item % UC_NSLOT;
            nb_sync(UC_NB_FULL + slot, cnt_ring);
            return zrow + (uint32_t)slot * SLOT_BYTES;
        };
        auto ring_release = [&](int item) {       // only when the noise warps will wait for this slot again
            if (item + UC_NSLOT < n_items) nb_arrive(UC_NB_EMPTY + item % UC_NSLOT, cnt_ring);
        };

        // ---- x_T -------------------------------------------------------------------------------
        float2 x[PW / 2];
        if (a.x_in) {
            const float* xr = a.x_in + mg * a.x_in_stride + PW * part;
#pragma unroll
            for (int i = 0; i < PW / 2; ++i)
                x[i] = make_float2((PW * part + 2 * i < P) ? xr[2 * i] : 0.f, (PW * part + 2 * i + 1 < P) ? xr[2 * i + 1] : 0.f);
        } else {
            const uint32_t zb = ring_wait(0);
#pragma unroll
            for (int c = 0; c < PW / 4; ++c) {
                const float4 z4 = lds128(zb + (((uint32_t)(PW / 4 * part + c) ^ zsw) << 4));
                x[2 * c] = make_float2(z4.x, z4.y);
                x[2 * c + 1] = make_float2(z4.z, z4.w);
            }
            ring_release(0);
        }

        const uint32_t xchunk = sX + (uint32_t)(row / 8) * sbo_bytes(UC_K1) + (uint32_t)(row % 8) * 16u + (uint32_t)(PW / 8 * part) * kLBO;
        const uint32_t hrow = sH + (uint32_t)(row / 8) * sbo_bytes(UC_H) + (uint32_t)(row % 8) * 16u + (uint32_t)(CW / 8 * part) * kLBO;

        // operands of the next GEMM1: this thread's chunk(s) of Xaug (W1aug was refreshed earlier in the step)
        auto publish_gemm1_operands = [&]() {
#pragma unroll
            for (int c = 0; c < PW / 8; ++c) {
                const bool last = (PW / 8 * part + c) == 3;   // parameters 24..28 + three constant-one columns (bf16 1.0 = 0x3F80)
                if (SPLIT) {      // the residual tile carries zeros under the constant-one columns
                    uint32_t hi[4], lo[4];
                    split_bf16_pair(x[4 * c].x, x[4 * c].y, hi[0], lo[0]);
                    split_bf16_pair(x[4 * c + 1].x, x[4 * c + 1].y, hi[1], lo[1]);
                    split_bf16_pair(x[4 * c + 2].x, last ? 1.0f : x[4 * c + 2].y, hi[2], lo[2]);
                    split_bf16_pair(last ? 1.0f : x[4 * c + 3].x, last ? 1.0f : x[4 * c + 3].y, hi[3], lo[3]);
                    sts_u4(xchunk + (uint32_t)c * kLBO, hi[0], hi[1], hi[2], hi[3]);
                    sts_u4(xchunk + X_LO + (uint32_t)c * kLBO, lo[0], lo[1], lo[2], lo[3]);
                    continue;
                }
                const uint32_t w2 = last ? pack_bf16(x[4 * c + 2].x, 1.0f) : pack_bf16(x[4 * c + 2].x, x[4 * c + 2].y);
                const uint32_t w3 = last ? 0x3F803F80u : pack_bf16(x[4 * c + 3].x, x[4 * c + 3].y);
                sts_u4(xchunk + (uint32_t)c * kLBO, pack_bf16(x[4 * c].x, x[4 * c].y), pack_bf16(x[4 * c + 1].x, x[4 * c + 1].y), w2, w3);
            }
            fence_proxy_async();
            tc_fence_before();          // this thread's TMEM reads of the step are complete
            nb_arrive(UC_NB_X, cnt_x);
        };
        // c_t runs two steps ahead in registers: the row for step it+2 is requested at the top of step it and
        // written into W1aug in the middle of step it+1, so that even a DRAM-cold table row (each is read once)
        // has more than a full step to arrive
        float ct_pending[NAUG] = {};
        {
            float ct[NAUG];
            if (aug_owner) { load_ct(a.t_hi, ct); refresh_w1aug(ct); }
            if (aug_owner && n_steps > 1) load_ct(a.t_hi - 1, ct_pending);
            publish_gemm1_operands();
        }

#if UC_TIMING
        const bool timed = ex.timing != nullptr && blockIdx.x == 0 && et == 0;
        long long tw[6] = {0, 0, 0, 0, 0, 0}, k0 = 0, k1 = 0, tf[2] = {0, 0};
#endif
#pragma unroll 1
        for (int it = 0; it < n_steps; ++it) {
            UC_T(if (timed) k0 = clock64();)
            const int t = a.t_hi - it;
            const uint32_t ph = (uint32_t)it & 1u;
            // prefetches: the step scalars and the c_t elements of the step after next
            const float4 cf = __ldg(reinterpret_cast<const float4*>(a.coef) + t);
            float ct_far[NAUG] = {};
            const bool more = it + 1 < n_steps;
            if (aug_owner && it + 2 < n_steps) load_ct(t - 2, ct_far);
            // this step's noise: the ring runs ahead, so this rendezvous is normally already complete
            uint32_t zb = 0;
            if (t > 0) zb = ring_wait(it + first_item_step);
            UC_T(if (timed) { k1 = clock64(); tw[0] += k1 - k0; })
            if (ok) ok = mbar_wait(bar_d, ph);
            UC_T(if (timed) { k0 = clock64(); tw[1] += k0 - k1; })
            tc_fence_after();
            // ---- epilogue 1: h = ReLU(D [+ c_b]) -> bf16 A operand of GEMM2 ----------------------
            if (SHARED) {
#pragma unroll
                for (int half2 = 0; half2 < CW / UC_EPI_CHUNK; ++half2) {     // UC_EPI_CHUNK columns in flight per wait
                    uint32_t dv[UC_EPI_CHUNK];
#pragma unroll
                    for (int hh = 0; hh < UC_EPI_CHUNK / 16; ++hh)
                        tmem_ld16(tD + UC_EPI_CHUNK * half2 + 16 * hh, *reinterpret_cast<uint32_t(*)[16]>(&dv[16 * hh]));
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < UC_EPI_CHUNK / 8; ++q) {
                        if (SPLIT) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                split_bf16_pair_relu(__uint_as_float(dv[8 * q + 2 * i]), __uint_as_float(dv[8 * q + 2 * i + 1]), hi[i], lo[i]);
                            const uint32_t dst = hrow + (uint32_t)(half2 * (UC_EPI_CHUNK / 8) + q) * kLBO;
                            sts_u4(dst, hi[0], hi[1], hi[2], hi[3]);
                            sts_u4(dst + H_LO, lo[0], lo[1], lo[2], lo[3]);
                            if (q == UC_EPI_CHUNK / 16 - 1) h_arrive(UC_NB_H);      // first half of this thread's columns
                            continue;
                        }
                        sts_u4(hrow + (uint32_t)(half2 * (UC_EPI_CHUNK / 8) + q) * kLBO,
                               pack_bf16_relu(__uint_as_float(dv[8 * q]), __uint_as_float(dv[8 * q + 1])),
                               pack_bf16_relu(__uint_as_float(dv[8 * q + 2]), __uint_as_float(dv[8 * q + 3])),
                               pack_bf16_relu(__uint_as_float(dv[8 * q + 4]), __uint_as_float(dv[8 * q + 5])),
                               pack_bf16_relu(__uint_as_float(dv[8 * q + 6]), __uint_as_float(dv[8 * q + 7])));
                    }
                }
            } else {
#pragma unroll
                for (int hh = 0; hh < CW / 16; ++hh) {       // 16 columns at a time: accumulator + c_b
                    uint32_t dv[16], cv[16];
                    tmem_ld16(tD + 16 * hh, dv);
                    tmem_ld16(tCB + 16 * hh, cv);
                    tmem_ld_wait();
                    uint32_t pk[8], pl[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 hsum = fadd2(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])),
                                                  make_float2(__uint_as_float(cv[2 * i]), __uint_as_float(cv[2 * i + 1])));
                        if (SPLIT) split_bf16_pair_relu(hsum.x, hsum.y, pk[i], pl[i]);
                        else pk[i] = pack_bf16_relu(hsum.x, hsum.y);
                    }
                    sts_u4(hrow + (uint32_t)(2 * hh) * kLBO, pk[0], pk[1], pk[2], pk[3]);
                    sts_u4(hrow + (uint32_t)(2 * hh + 1) * kLBO, pk[4], pk[5], pk[6], pk[7]);
                    if (SPLIT) {
                        sts_u4(hrow + H_LO + (uint32_t)(2 * hh) * kLBO, pl[0], pl[1], pl[2], pl[3]);
                        sts_u4(hrow + H_LO + (uint32_t)(2 * hh + 1) * kLBO, pl[4], pl[5], pl[6], pl[7]);
                        if (hh == CW / 32 - 1) h_arrive(UC_NB_H);
                    }
                }
            }
            UC_T(long long f0 = 0; if (timed) f0 = clock64();)
            fence_proxy_async();
            UC_T(if (timed) { const long long f1 = clock64(); tf[0] += f1 - f0; f0 = f1; })
            tc_fence_before();
            nb_arrive((SPLIT ? UC_NB_H2 : UC_NB_H) + (uint32_t)part, cnt_h);
            UC_T(if (timed) tf[1] += clock64() - f0;)
            // GEMM1 of this step has long read W1aug: write the next step's v while GEMM2 runs (made visible
            // to the tensor core by the proxy fence of the publish below, before the next NB_X arrival)
            if (aug_owner && more) refresh_w1aug(ct_pending);
#pragma unroll
            for (int i = 0; i < NAUG; ++i) ct_pending[i] = ct_far[i];
            UC_T(if (timed) { k1 = clock64(); tw[2] += k1 - k0; })
            UC_T(if (timed) { k0 = clock64(); tw[3] += k0 - k1; })
            if (ok) ok = mbar_wait(bar_e, ph);
            UC_T(if (timed) { k1 = clock64(); tw[4] += k1 - k0; })
            tc_fence_after();
            // ---- epilogue 2: eps -> posterior update of this thread's parameters ----------------
            {
                uint32_t ev[PW], ev2[SPLIT ? PW : 1];
                if (PW == 16) tmem_ld16(tE, *reinterpret_cast<uint32_t(*)[16]>(&ev[0]));
                else tmem_ld8(tE, *reinterpret_cast<uint32_t(*)[8]>(&ev[0]));
                if (SPLIT) {      // E2 = H_hi W_lo^T, 32 columns further
                    if (PW == 16) tmem_ld16(tE + UC_N2, *reinterpret_cast<uint32_t(*)[16]>(&ev2[0]));
                    else tmem_ld8(tE + UC_N2, *reinterpret_cast<uint32_t(*)[8]>(&ev2[0]));
                }
                tmem_ld_wait();
                if (SPLIT) {
#pragma unroll
                    for (int i = 0; i < PW / 2; ++i) {
                        const float2 e12 = fadd2(make_float2(__uint_as_float(ev[2 * i]), __uint_as_float(ev[2 * i + 1])),
                                                 make_float2(__uint_as_float(ev2[2 * i]), __uint_as_float(ev2[2 * i + 1])));
                        ev[2 * i] = __float_as_uint(e12.x); ev[2 * i + 1] = __float_as_uint(e12.y);
                    }
                }
                // ECD.py:111-118: u = coef*eps ; v = x - u ; x' = c1*v ; [ w = sigma*z ; x' = x' + w ]
                // in packed fp32 (x - u formed as x + (-coef)*eps; the assembler is free to contract
                // these, the bf16 path's contract is its tolerance, not bit-exactness of the update)
                const float2 ncoef = make_float2(-cf.x, -cf.x), c1 = make_float2(cf.y, cf.y), sg = make_float2(cf.z, cf.z);
#pragma unroll
                for (int c = 0; c < PW / 4; ++c) {
                    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t > 0) z4 = lds128(zb + (((uint32_t)(PW / 4 * part + c) ^ zsw) << 4));
                    const float4 b4 = lds128(b2a + 16u * c);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = 2 * c + u;
                        const float2 e = fadd2(make_float2(__uint_as_float(ev[2 * i]), __uint_as_float(ev[2 * i + 1])),
                                               u == 0 ? make_float2(b4.x, b4.y) : make_float2(b4.z, b4.w));
                        float2 xn = fmul2(c1, fadd2(x[i], fmul2(ncoef, e)));
                        if (t > 0) xn = fadd2(xn, fmul2(sg, u == 0 ? make_float2(z4.x, z4.y) : make_float2(z4.z, z4.w)));
                        x[i] = xn;
                        if (TRACE && mvalid) {
                            float* dst = a.eps_trace + ((int64_t)t * a.B + m0 + row) * P + PW * part + 2 * i;
                            if (PW * part + 2 * i < P) dst[0] = e.x;
                            if (PW * part + 2 * i + 1 < P) dst[1] = e.y;
                        }
                    }
                }
            }
            if (t > 0) ring_release(it + first_item_step);
            if (more) publish_gemm1_operands();
            UC_T(if (timed) tw[5] += clock64() - k1;)
        }
#if UC_TIMING
        if (timed) {
#pragma unroll
            for (int i = 0; i < 6; ++i) ex.timing[i] = tw[i];
            ex.timing[6] = tf[0]; ex.timing[7] = tf[1];
            ex.timing[15] = n_steps;
        }
#endif
        // parameters >= P of the padded tile carry finite garbage that the zero weight columns
        // ignore; they are never stored
        if (mvalid) {
            float* dst = a.x_out + (m0 + row) * P + PW * part;
#pragma unroll
            for (int i = 0; i < PW / 2; ++i) {
                if (PW * part + 2 * i < P) dst[2 * i] = x[i].x;
                if (PW * part + 2 * i + 1 < P) dst[2 * i + 1] = x[i].y;
            }
        }
    }
    // (epilogue warps whose rows carry no member have nothing to do)
    if (!ok) s.timeout = 1;
    tc_fence_before();
    __syncthreads();
    if (s.timeout != 0) {                   // an MMA never completed: poison the tile's output
        for (int i = tid; i < mpc * a.P; i += UC_THREADS)
            if (m0 + i / a.P < a.B) a.x_out[m0 * a.P + i] = __int_as_float(0x7fc00000);
        if (tid == 0) ex.status[0] = 1;
    }
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace ertdiff
