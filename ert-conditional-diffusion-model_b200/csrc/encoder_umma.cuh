// Condition encoder (ECD.py:133-139) on the tensor cores: both Conv1d layers as implicit GEMMs
// (tcgen05.mma, bf16 operands, fp32 accumulators in TMEM), input fed by TMA, ReLU / bias / global
// average pool fused into the TMEM epilogues.  Used by precision = bf16 (BASELINE configs 3/4);
// the fp32 CUDA-core kernel of encoder.cuh serves the fp32 contract.
//
// One CTA = one condition x one chunk of `tpc` tiles of 128 conv2 positions (5 when there are few
// conditions, so that the grid still fills the machine; a whole condition otherwise, which halves the
// per-CTA set-up cost per tile).  Per tile
// (p0 = first conv2 position):
//   TMA     14 channel rows of the fp32 condition, window l = 4(p0-1) .. 4(p0+128)-1, into a
//           double-buffered staging area: one bulk copy (cp.async.bulk, mbarrier complete_tx) per
//           row, issued a whole tile ahead.  The rows of a (14, 4693) condition are only 4-byte
//           aligned, which rules out tensor maps with a row stride; each copy therefore starts at
//           the 16-byte boundary below its window and the readers skip the 0..3 leading elements
//   convert staging -> bf16 "phase" blocks PH[r][j][ci] = in[ci][4(p0-1+j)+r]  (r = l mod 4):
//           the stride-2 taps of both convolutions become unit-stride row windows
//   GEMM1   conv1 (ECD.py:134) for the even / odd output positions the tile needs:
//             E[i] = h1[2(p0+i)]   = W(0) PH3[i] + W(1) PH0[i+1] + W(2) PH1[i+1]
//             O[i] = h1[2(p0+i)+1] = W(0) PH1[i+1] + W(1) PH2[i+1] + W(2) PH3[i+1]
//           one MMA (M=128, N=32, K=16 channels) per tap; a tap's A operand is a 128-row window
//           of a phase block, selected purely by the descriptor's start address (see layout)
//   epi 1   +bias, ReLU, zero beyond L1 (conv2's padding), -> bf16 blocks HE[i] = E[i],
//           HO[i+1] = O[i]; HO[0] = h1[2p0-1] is carried over from the previous tile
//   GEMM2   conv2 (ECD.py:136): out[p0+i] = V(0) HO[i] + V(1) HE[i] + V(2) HO[i+1]
//           six MMAs (M=128, N=64, K=16)
//   epi 2   +bias, ReLU, rows beyond L2 dropped, added to 64 per-thread running sums
// After the last tile the 128 threads' sums are added in a fixed order and written as the
// chunk's pooling partial; k_encoder_finish (encoder.cuh) completes mean -> Linear -> c_b.
//
// Operand layout ("chunk-major", K-major without swizzle): a block holds one 16-byte K-chunk
// (8 bf16 channels) for 129 consecutive rows, rows 16 bytes apart.  In descriptor terms the
// stride between 8-row core matrices is SBO = 128 bytes and the stride between K-chunks is
// LBO = one block, so shifting the row window by one row is start address + 16 bytes.
//
// Algorithmic work per condition: 20.7 MFLOP (conv1 6.3 M + conv2 14.4 M), 262,808 B read once.
#pragma once
#include "denoiser.cuh"
#include "common.cuh"
#include "umma.cuh"

// -DEU_TIMING=1: thread 0 of CTA (0,0) accumulates per-phase cycle counts into EncUmmaParams::timing
#ifndef EU_TIMING
#define EU_TIMING 0
#endif
#if EU_TIMING
#define EU_T(...) __VA_ARGS__
#else
#define EU_T(...)
#endif

namespace ertdiff {

constexpr int EU_WORKERS = 256;                 // conversion + epilogue threads (8 warps)
constexpr int EU_THREADS = EU_WORKERS + 64;     // + the TMA producer warp + the MMA-issue warp
constexpr int EU_ROWS = 129;                    // rows per block: the tile's 128 + one neighbour
constexpr int EU_BLK = EU_ROWS * 16;            // bytes per block
constexpr int EU_COPY = 4 * EU_ROWS + 4;        // fp32 elements per bulk copy: the window + up to 3 leading elements
constexpr int EU_STG_ROW = EU_COPY + 8;         // floats per staged channel row (16-byte multiple)
constexpr int EU_K1 = 48;                       // conv1 K: 3 taps x 16 (14 channels + 2 zero)
constexpr int EU_K2 = 96;                       // conv2 K: 3 taps x 32 channels

struct EncUmmaSmem {
    float stage[2][kInChannels][EU_STG_ROW];    // TMA destinations (2 x 29,568 B); reused for the final reduction
    unsigned char ph[4][2][EU_BLK];             // input phases: [l mod 4][channel chunk]
    unsigned char he[4][EU_BLK];                // conv1 output, even positions: [channel chunk]
    unsigned char ho[4][EU_BLK];                // conv1 output, odd positions (row 0 = carry)
    unsigned char w1[kConv1Out * EU_K1 * 2];    // B of GEMM1
    unsigned char w2[kConv2Out * EU_K2 * 2];    // B of GEMM2
    alignas(16) float b1[kConv1Out];
    alignas(16) float b2[kConv2Out];
    unsigned long long desc[24];                // MMA operand descriptors (thread 0 builds and uses them)
    unsigned long long bar_tma[2], bar_free[2], bar_mma, bar_mma2;
    uint32_t tmem_slot;
    int timeout;
};

// bf16 B operands, tap-major K (k = tap*16 + ci for conv1, tap*32 + ci for conv2), umma::elem_offset layout
__global__ void k_pack_encoder_umma(const float* __restrict__ c1w /*(32,14,3)*/, const float* __restrict__ c2w /*(64,32,3)*/,
                                    unsigned short* __restrict__ w1_pk, unsigned short* __restrict__ w2_pk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kConv1Out * EU_K1) {
        const int co = i / EU_K1, k = i % EU_K1, tap = k / 16, ci = k % 16;
        const float v = ci < kInChannels ? c1w[(co * kInChannels + ci) * 3 + tap] : 0.f;
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        w1_pk[umma::elem_offset(co, k, EU_K1) / 2] = *reinterpret_cast<const unsigned short*>(&b);
    }
    if (i < kConv2Out * EU_K2) {
        const int co = i / EU_K2, k = i % EU_K2, tap = k / 32, ci = k % 32;
        const __nv_bfloat16 b = __float2bfloat16_rn(c2w[(co * kConv1Out + ci) * 3 + tap]);
        w2_pk[umma::elem_offset(co, k, EU_K2) / 2] = *reinterpret_cast<const unsigned short*>(&b);
    }
}

// TMA bulk copy global -> shared (16-byte aligned source, destination and size), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}

struct EncUmmaParams {
    const float* base;        // 16-byte aligned address at or below the first condition
    int64_t elem0;            // element index (from base) of condition 0, channel 0, l = 0
    int64_t total;            // elements from base to the end of the last condition, rounded up to 4
    int64_t member_stride;    // elements between conditions
    int L, L1, L2;
    int n_chunks;
    int tpc;                  // tiles per CTA
    const uint4* w1_pk;
    const uint4* w2_pk;
    const float* b1;
    const float* b2;
    const float* conv1_w;     // fp32 [(ci*3+k)][32]: the carried row of a chunk's first tile
    float* partial;           // (n_cond, n_chunks, 64)
    long long* timing;        // EU_TIMING only: [0] wait TMA [1] convert [2] sync+issue1 [3] wait MMA1 [4] epi1 [5] sync+issue2 [6] wait MMA2 [7] epi2 [15] tiles
    int* status;
};

__device__ __forceinline__ void worker_barrier() {      // the 8 worker warps only (named barrier 1)
    asm volatile("bar.sync 1, %0;" ::"n"(EU_WORKERS) : "memory");
}

__global__ void __launch_bounds__(EU_THREADS, 2)
k_encoder_umma(const EncUmmaParams a) {
    using namespace umma;
    extern __shared__ __align__(1024) unsigned char eu_smem_raw[];
    EncUmmaSmem& s = *reinterpret_cast<EncUmmaSmem*>(eu_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x;
    const int64_t cond = blockIdx.y;
    uint32_t sStage = smem_u32(&s.stage[0][0][0]);
    uint32_t sPH = smem_u32(&s.ph[0][0][0]), sHE = smem_u32(&s.he[0][0]), sHO = smem_u32(&s.ho[0][0]);
    uint32_t sW1 = smem_u32(s.w1), sW2 = smem_u32(s.w2), sB1 = smem_u32(s.b1), sB2 = smem_u32(s.b2);
    uint32_t bar_tma = smem_u32(&s.bar_tma[0]), bar_free = smem_u32(&s.bar_free[0]), bar_mma = smem_u32(&s.bar_mma);
    uint32_t bar_mma2 = smem_u32(&s.bar_mma2);
    asm volatile("" : "+r"(sStage), "+r"(sPH), "+r"(sHE), "+r"(sHO), "+r"(sW1), "+r"(sW2), "+r"(sB1), "+r"(sB2),
                      "+r"(bar_tma), "+r"(bar_free), "+r"(bar_mma), "+r"(bar_mma2));
    constexpr uint32_t STAGE_BYTES = kInChannels * EU_STG_ROW * 4;

    // ---- one-time setup ------------------------------------------------------------------------
    {
        uint4* d1 = reinterpret_cast<uint4*>(s.w1);
        uint4* d2 = reinterpret_cast<uint4*>(s.w2);
        for (int i = tid; i < kConv1Out * EU_K1 * 2 / 16; i += EU_THREADS) d1[i] = a.w1_pk[i];
        for (int i = tid; i < kConv2Out * EU_K2 * 2 / 16; i += EU_THREADS) d2[i] = a.w2_pk[i];
        if (tid < kConv1Out) s.b1[tid] = a.b1[tid];
        if (tid < kConv2Out) s.b2[tid] = a.b2[tid];
    }
    if (warp == 0) tmem_alloc(smem_u32(&s.tmem_slot), 128);
    if (tid == 0) {
        s.timeout = 0;
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tma + 8u * i, 1);
            mbar_init(bar_free + 8u * i, EU_WORKERS / 32);
        }
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma2, 1);
        fence_mbar_init();
        // MMA operand descriptors: all loop-invariant, built once by the issuing thread and parked in
        // shared memory (21 64-bit values would otherwise occupy registers in every thread)
        const int ev_ph[3] = {3, 0, 1}, ev_sh[3] = {0, 1, 1}, od_ph[3] = {1, 2, 3};
        for (int t = 0; t < 3; ++t) {
            s.desc[t] = smem_desc(sW1 + 2 * t * kLBO, kLBO, sbo_bytes(EU_K1));
            s.desc[3 + t] = smem_desc(sPH + (uint32_t)(ev_ph[t] * 2 * EU_BLK + ev_sh[t] * 16), EU_BLK, 128);
            s.desc[6 + t] = smem_desc(sPH + (uint32_t)(od_ph[t] * 2 * EU_BLK + 16), EU_BLK, 128);
        }
        for (int ks = 0; ks < 6; ++ks) {
            const int tap = ks >> 1, cp = ks & 1;
            s.desc[9 + ks] = smem_desc((tap == 1 ? sHE : sHO) + (uint32_t)(2 * cp * EU_BLK + (tap == 2 ? 16 : 0)), EU_BLK, 128);
            s.desc[15 + ks] = smem_desc(sW2 + 2 * ks * kLBO, kLBO, sbo_bytes(EU_K2));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_slot;

    const int p_begin = chunk * a.tpc * 128;
    int n_tiles = (a.L2 - p_begin + 127) / 128;
    if (n_tiles > a.tpc) n_tiles = a.tpc;
    const int64_t cond_elem = a.elem0 + cond * a.member_stride;
    // A bulk copy needs a 16-byte aligned source, but a channel row of an odd-length condition starts
    // anywhere: each row is fetched from the aligned element below its window, and readers skip the
    // 0..3 leading elements (lbase is a multiple of 4).  Copies are clipped to the caller's tensor.
    const int e0 = (int)(cond_elem & 3), Lm = a.L & 3;

    if (warp == EU_WORKERS / 32) {
        // ===== TMA producer warp: lane = channel row; runs up to two tiles ahead of the workers =========
        bool ok = true;
        for (int tile = 0; tile < n_tiles && ok; ++tile) {
            const int buf = tile & 1;
            // staging buffer `buf` is free once the workers have converted tile - 2
            if (tile >= 2) ok = mbar_wait(bar_free + 8u * buf, (uint32_t)((tile >> 1) + 1) & 1u);
            const int64_t lbase = 4 * (int64_t)(p_begin + tile * 128 - 1);
            const uint32_t bar = bar_tma + 8u * buf;
            int64_t first = 0;
            uint32_t bytes = 0, dst = 0;
            if (lane < kInChannels) {
                first = (cond_elem + (int64_t)lane * a.L + lbase) & ~(int64_t)3;
                int64_t last = first + EU_COPY;
                dst = sStage + (uint32_t)buf * STAGE_BYTES + 4u * (uint32_t)(lane * EU_STG_ROW);
                if (first < 0) { dst += (uint32_t)(-first) * 4u; first = 0; }
                if (last > a.total) last = a.total;
                bytes = last > first ? (uint32_t)(last - first) * 4u : 0u;
            }
            uint32_t sum = bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) mbar_expect_tx(bar, sum);
            __syncwarp();
            if (bytes) tma_bulk_load(dst, a.base + first, bytes, bar);
        }
        if (!ok) s.timeout = 1;
    } else if (warp == EU_WORKERS / 32 + 1) {
        // ===== MMA-issue warp: paced by the workers' arrive-only named barriers ===========================
        // barrier 2 ("phase blocks of the next tile written") -> GEMM1, barrier 3 ("HE/HO written") -> GEMM2
        constexpr uint32_t IDESC1 = idesc_bf16_f32(128, kConv1Out);
        constexpr uint32_t IDESC2 = idesc_bf16_f32(128, kConv2Out);
        uint64_t d[21];                          // the loop-invariant operand descriptors, in registers
#pragma unroll
        for (int i = 0; i < 21; ++i) d[i] = s.desc[i];
        auto gemm1 = [&]() {                     // even positions -> TMEM columns 0..31, odd -> 32..63
            asm volatile("bar.sync 2, %0;" ::"n"(EU_WORKERS + 32) : "memory");
            tc_fence_after();
            if (elect_one()) {
                mma_bf16_first(tmem, d[3], d[0], IDESC1);
                mma_bf16_acc(tmem, d[4], d[1], IDESC1);
                mma_bf16_acc(tmem, d[5], d[2], IDESC1);
                mma_bf16_first(tmem + 32, d[6], d[0], IDESC1);
                mma_bf16_acc(tmem + 32, d[7], d[1], IDESC1);
                mma_bf16_acc(tmem + 32, d[8], d[2], IDESC1);
                mma_commit(bar_mma);
            }
            __syncwarp();
        };
        if (n_tiles > 0) gemm1();
        for (int tile = 0; tile < n_tiles; ++tile) {
            asm volatile("bar.sync 3, %0;" ::"n"(EU_WORKERS + 32) : "memory");
            tc_fence_after();
            if (elect_one()) {
                mma_bf16_first(tmem + 64, d[9], d[15], IDESC2);
#pragma unroll
                for (int ks = 1; ks < 6; ++ks) mma_bf16_acc(tmem + 64, d[9 + ks], d[15 + ks], IDESC2);
                mma_commit(bar_mma2);
            }
            __syncwarp();
            if (tile + 1 < n_tiles) gemm1();
        }
    } else {
        // ===== worker warps: conversion and TMEM epilogues, software-pipelined across tiles ==============
        //   epilogue 1 (t) | convert (t+1) while GEMM2 (t) runs | epilogue 2 (t) while GEMM1 (t+1) runs
        // Workers never wait for each other inside the loop: they wait on the MMA / TMA mbarriers and only
        // ARRIVE on the named barriers that pace the MMA warp.  After a timeout every wait is skipped so that
        // all arrivals still happen (nobody is left blocked); the chunk's result is poisoned.
        // every warp may touch the TMEM lane quarter (warp % 4): warps 0-3 and 4-7 split the columns
        const int row = (warp & 3) * 32 + lane;      // tile row = TMEM lane
        const int wh = warp >> 2;                    // 0: even conv1 phase / conv2 channels 0..31, 1: odd / 32..63
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // conversion role: channel chunk wh (8 channels) for l_local = (tid & 127) + 128 n
        uint32_t src_row[8];                         // staging offset of channel 8 wh + u at l_local = 0 (shift included)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ci = 8 * wh + u;
            const int cic = ci < kInChannels ? ci : 0;
            src_row[u] = 4u * (uint32_t)(cic * EU_STG_ROW + ((e0 + cic * Lm) & 3) + (tid & 127));
        }
        EU_T(const bool timed = a.timing && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0; long long tt[8] = {0,0,0,0,0,0,0,0}, q0 = 0, q1 = 0;)
        volatile int* timeout_flag = &s.timeout;
        auto wait_bar = [&](uint32_t bar, uint32_t parity) {
            if (*timeout_flag) return;
            if (!mbar_wait(bar, parity)) *timeout_flag = 1;
        };

        // staging (fp32, [ci][l]) -> bf16 phase blocks of one tile, then "phase blocks written" (barrier 2).
        // lanes run along l: the eight 4-byte reads are conflict-free and the 16-byte stores of a quarter
        // warp fall into distinct banks.  Only the first / last tiles of a row need masks.
        auto convert_tile = [&](int tile) {
            const int p0 = p_begin + tile * 128;
            const int lbase = 4 * (p0 - 1);
            const int buf = tile & 1;
            EU_T(long long w0 = 0; if (timed) w0 = clock64();)
            wait_bar(bar_tma + 8u * buf, (uint32_t)(tile >> 1) & 1u);
            EU_T(if (timed) tt[0] += clock64() - w0;)
            const uint32_t stg = sStage + (uint32_t)buf * STAGE_BYTES;
            const bool edge = lbase < 0 || lbase + 4 * EU_ROWS > a.L;
            auto convert = [&](int ll) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = lds32(stg + src_row[u] + 4u * (uint32_t)(ll - (tid & 127)));
                if (wh == 1) { v[6] = 0.f; v[7] = 0.f; }                  // channels 14, 15 are padding
                if (edge) {
                    const int l = lbase + ll;
                    if (l < 0 || l >= a.L) {                              // conv1's zero padding
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = 0.f;
                    }
                }
                sts_u4(sPH + (uint32_t)(((ll & 3) * 2 + wh) * EU_BLK + (ll >> 2) * 16),
                       pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            };
#pragma unroll
            for (int n = 0; n < 4; ++n) convert((tid & 127) + 128 * n);      // 4 x 128 = 512 of the 516 positions
            if ((tid & 127) < 4 * EU_ROWS - 512) convert((tid & 127) + 512);
            // HO row 0 = h1[2 p0 - 1] for a chunk's first tile (later tiles: carried, see epilogue 2)
            if (tile == 0 && warp == 7) {            // lane = output channel; fp32 from the staged window
                const int q = 2 * p0 - 1;
                float h = 0.f;
                if (q >= 0 && q < a.L1) {
                    h = s.b1[lane];
                    for (int ci = 0; ci < kInChannels; ++ci)
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int ll = 1 + k;                  // l = 2q - 1 + k = 4 p0 - 3 + k
                            const int l = lbase + ll;
                            const float x = (l >= 0 && l < a.L) ? s.stage[buf][ci][ll + ((e0 + ci * Lm) & 3)] : 0.f;
                            h = fmaf(a.conv1_w[(ci * 3 + k) * kConv1Out + lane], x, h);
                        }
                    h = fmaxf(h, 0.f);
                }
                const __nv_bfloat16 hb = __float2bfloat16_rn(h);
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(sHO + (uint32_t)((lane >> 3) * EU_BLK + (lane & 7) * 2)),
                             "h"(*reinterpret_cast<const unsigned short*>(&hb)) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8u * buf);        // this warp is done with the staging buffer
            fence_proxy_async();
            tc_fence_before();                     // (this thread's TMEM reads of earlier tiles are complete)
            asm volatile("bar.arrive 2, %0;" ::"n"(EU_WORKERS + 32) : "memory");
        };

        float2 acc[16];                              // pooled running sums: conv2 channels 32 wh .. 32 wh + 31 of this row
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = make_float2(0.f, 0.f);
        uint4 carry[4];                              // thread (row 127, odd phase): O[127] = next tile's HO row 0
#pragma unroll
        for (int c = 0; c < 4; ++c) carry[c] = make_uint4(0u, 0u, 0u, 0u);
        if (n_tiles > 0) convert_tile(0);
#pragma unroll 1
        for (int tile = 0; tile < n_tiles; ++tile) {
            EU_T(if (timed) q0 = clock64();)
            const int p0 = p_begin + tile * 128;
            const uint32_t par = (uint32_t)tile & 1u;
            wait_bar(bar_mma, par);                 // GEMM1 (tile) complete
            EU_T(if (timed) { q1 = clock64(); tt[3] += q1 - q0; q0 = q1; })
            tc_fence_after();
            // ---- epilogue 1: bias + ReLU -> bf16 HE (warps 0-3) / HO (warps 4-7) --------------------
            {
                uint32_t dv[32];
                tmem_ld32(tlane + 32 * wh, dv);
                tmem_ld_wait();
                const int q = 2 * (p0 + row) + wh;
                const bool valid = q < a.L1;                          // conv2's zero padding beyond L1
                const uint32_t dst = (wh == 0 ? sHE : sHO) + (uint32_t)((row + wh) * 16);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 ba = lds128(sB1 + 32u * c), bb = lds128(sB1 + 32u * c + 16u);
                    const float2 h0 = fadd2(make_float2(__uint_as_float(dv[8 * c]), __uint_as_float(dv[8 * c + 1])), make_float2(ba.x, ba.y));
                    const float2 h1 = fadd2(make_float2(__uint_as_float(dv[8 * c + 2]), __uint_as_float(dv[8 * c + 3])), make_float2(ba.z, ba.w));
                    const float2 h2 = fadd2(make_float2(__uint_as_float(dv[8 * c + 4]), __uint_as_float(dv[8 * c + 5])), make_float2(bb.x, bb.y));
                    const float2 h3 = fadd2(make_float2(__uint_as_float(dv[8 * c + 6]), __uint_as_float(dv[8 * c + 7])), make_float2(bb.z, bb.w));
                    uint4 pk = make_uint4(pack_bf16_relu(h0.x, h0.y), pack_bf16_relu(h1.x, h1.y),
                                          pack_bf16_relu(h2.x, h2.y), pack_bf16_relu(h3.x, h3.y));
                    if (!valid) pk = make_uint4(0u, 0u, 0u, 0u);
                    sts_u4(dst + (uint32_t)(c * EU_BLK), pk.x, pk.y, pk.z, pk.w);
                    if (tid == EU_WORKERS - 1) carry[c] = pk;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            asm volatile("bar.arrive 3, %0;" ::"n"(EU_WORKERS + 32) : "memory");     // HE / HO written -> GEMM2 (tile)
            EU_T(if (timed) { q1 = clock64(); tt[4] += q1 - q0; q0 = q1; })
            // ---- next tile's phase blocks while GEMM2 runs (GEMM1 of this tile has completed: PH is free) ----
            if (tile + 1 < n_tiles) convert_tile(tile + 1);
            EU_T(if (timed) { q1 = clock64(); tt[1] += q1 - q0; q0 = q1; })
            wait_bar(bar_mma2, par);                // GEMM2 (tile) complete
            EU_T(if (timed) { q1 = clock64(); tt[6] += q1 - q0; q0 = q1; })
            tc_fence_after();
            // ---- epilogue 2: bias + ReLU + pooled running sums (GEMM1 of the next tile is in flight) ----
            {
                uint32_t dv[32];
                tmem_ld32(tlane + 64 + 32 * wh, dv);
                tmem_ld_wait();
                if ((p0 + row) < a.L2) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 b4 = lds128(sB2 + 128u * wh + 16u * c);
                        const float2 v0 = fadd2(make_float2(__uint_as_float(dv[4 * c]), __uint_as_float(dv[4 * c + 1])), make_float2(b4.x, b4.y));
                        const float2 v1 = fadd2(make_float2(__uint_as_float(dv[4 * c + 2]), __uint_as_float(dv[4 * c + 3])), make_float2(b4.z, b4.w));
                        acc[2 * c] = fadd2(acc[2 * c], make_float2(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f)));
                        acc[2 * c + 1] = fadd2(acc[2 * c + 1], make_float2(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f)));
                    }
                }
                if (tid == EU_WORKERS - 1) {       // next tile's HO row 0 (GEMM2 of this tile has completed)
#pragma unroll
                    for (int c = 0; c < 4; ++c) sts_u4(sHO + (uint32_t)(c * EU_BLK), carry[c].x, carry[c].y, carry[c].z, carry[c].w);
                }
            }
            tc_fence_before();
            EU_T(if (timed) { q1 = clock64(); tt[7] += q1 - q0; })
        }
        EU_T(if (timed) { for (int i = 0; i < 8; ++i) a.timing[i] = tt[i]; a.timing[15] = n_tiles; })
        // ---- deterministic reduction of the 128 rows' sums (staging buffer reused, rotated columns) --
        worker_barrier();
        float* red = &s.stage[0][0][0];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            red[row * kConv2Out + ((32 * wh + 2 * c + row) & (kConv2Out - 1))] = acc[c].x;
            red[row * kConv2Out + ((32 * wh + 2 * c + 1 + row) & (kConv2Out - 1))] = acc[c].y;
        }
        worker_barrier();
        if (tid < kConv2Out) {
            float sum = 0.f;
            for (int i = 0; i < 128; ++i) sum += red[i * kConv2Out + ((tid + i) & (kConv2Out - 1))];
            if (s.timeout) sum = __int_as_float(0x7fc00000);
            a.partial[(cond * a.n_chunks + chunk) * kConv2Out + tid] = sum;
        }
        if (tid == 0 && s.timeout) a.status[0] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace ertdiff
