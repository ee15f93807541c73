// Tensor-core version of the reverse chain for large ensembles (precision = bf16 operands,
// fp32 accumulation; BASELINE configs 3/4): the two projections of the hoisted denoiser step run
// as tcgen05.mma with accumulators in TMEM, everything between them stays on chip.
//
// One CTA = a tile of 128 members for all steps: 16 worker warps + 1 MMA-issue warp, no
// __syncthreads in the step loop (mbarriers only, so warps drift apart and fill each other's
// waits).  TMEM lane r <-> member r of the tile; a warp may only touch the lane quarter
// (warp % 4), so worker warp w handles members 32*(w%4)..+31 and column group g = w/4: FOUR
// threads share one member, each owning a quarter of the hidden columns (epilogue 1) and 8 of the
// 32 padded parameters (RNG, epilogue 2, x).
//
//   GEMM1  D[128 members x 128 hidden] = Xaug[128 x 32] * W1aug[128 x 32]^T           (2 MMAs, K=16)
//          Xaug row     = [x_0..x_28, 1, 1, 1]   (bf16, rewritten by its four threads every step)
//          W1aug row j  = [W0x[j][0..28], v_hi[j], v_mid[j], v_lo[j]]
//          v = c_t (per-step vector), or c_t + c_b when all members share one condition; split
//          into three bf16 terms (exact to 2^-24) and folded into the contraction through the
//          three spare K columns -- the "embedding add" happens inside the MMA
//   epi 1  h = ReLU(D [+ c_b[member], fp32 registers, only with distinct conditions])
//          -> bf16 (cvt.rn.relu.bf16x2) into the K-major A operand of GEMM2
//   GEMM2  E[128 members x 32] = Hbf16[128 x 128] * W2pad[32 x 128]^T                  (8 MMAs, K=16)
//          issued per column group as soon as that group's 32 K-columns of H are written
//   epi 2  eps = E + b2; bit-exact posterior update of the thread's 8 parameters (x stays in fp32
//          registers for the whole chain) with Philox / replayed noise, new Xaug chunk
// Barriers (all mbarriers, one completion per step each):
//   bar_x  (16 warp arrivals)  Xaug / W1aug of the next step written     workers -> MMA warp
//   bar_d  (tcgen05.commit)    D complete                                MMA warp -> workers
//   bar_h[g] (4 warp arrivals) H columns of group g written              workers -> MMA warp
//   bar_e  (tcgen05.commit)    E complete                                MMA warp -> workers
// The Philox + Box-Muller work of a step (two interleaved Philox calls, four Box-Muller pairs per
// thread) is issued ahead of the wait for D, i.e. while GEMM1 is in flight.
// TMEM: 256 columns (D: 0..127, E: 128..159).
// Algorithmic work: 14,848 FLOP per member-step, as in the fp32 kernel (the K/N padding to
// 32/32 is not counted).
#pragma once
#include "denoiser.cuh"
#include "umma.cuh"

// -DUC_TIMING=1 builds the phase-timing instrumentation read by ertdiff_debug_umma_timing (it costs
// ~16 registers, so production builds leave it out and that entry point reports zeros)
#ifndef UC_TIMING
#define UC_TIMING 0
#endif
#if UC_TIMING
#define UC_T(...) __VA_ARGS__
#else
#define UC_T(...)
#endif

namespace ertdiff {

constexpr int UC_M = 128;       // members per CTA
constexpr int UC_H = 128;       // hidden_dim this kernel is built for
constexpr int UC_K1 = 32;       // padded param_dim + 3 augmentation columns
constexpr int UC_N2 = 32;       // padded param_dim
constexpr int UC_WORKERS = 512; // 4 threads per member
constexpr int UC_THREADS = UC_WORKERS + 32;   // + the MMA-issue warp
constexpr int UC_AUG = 29;      // first augmentation column (param_dim <= 29)

struct UmmaChainSmem {
    unsigned char x[UC_M * UC_K1 * 2];      // A of GEMM1
    unsigned char h[UC_M * UC_H * 2];       // A of GEMM2
    unsigned char w1[UC_H * UC_K1 * 2];     // B of GEMM1 (W0x augmented)
    unsigned char w2[UC_N2 * UC_H * 2];     // B of GEMM2 (W2 padded)
    unsigned long long bar_x, bar_d, bar_e, bar_h[4];
    uint32_t tmem_slot;
    int timeout;
};

// pack the bf16 B operands once per load_state_dict: byte layout = umma::elem_offset
__global__ void k_pack_umma_weights(const float* __restrict__ w0xT /*(32,H)*/,
                                    const float* __restrict__ w2p /*(32,H)*/, int P,
                                    unsigned short* __restrict__ w1_pk, unsigned short* __restrict__ w2_pk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < UC_H * UC_K1) {
        const int j = i / UC_K1, k = i % UC_K1;
        const float v = (k < P) ? w0xT[k * UC_H + j] : 0.f;      // columns P..31 start as zero
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        w1_pk[umma::elem_offset(j, k, UC_K1) / 2] = *reinterpret_cast<const unsigned short*>(&b);
    }
    if (i < UC_N2 * UC_H) {
        const int p = i / UC_H, k = i % UC_H;
        const __nv_bfloat16 b = __float2bfloat16_rn(w2p[p * UC_H + k]);
        w2_pk[umma::elem_offset(p, k, UC_H) / 2] = *reinterpret_cast<const unsigned short*>(&b);
    }
}

struct UmmaChainExtra {
    const uint4* w1_pk;     // 8 KB
    const uint4* w2_pk;     // 8 KB
    int* status;            // [0] = 1 when an mbarrier wait timed out
    long long* timing;      // optional (16 int64): phase cycle sums of CTA 0, see ertdiff_debug_umma_timing
};

// packed fp32x2 add / mul (sm_100): two IEEE-rn operations per instruction
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
// {lo, hi} -> bf16x2 with ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

template <bool REPLAY, bool TRACE, bool SHARED>
__global__ void __launch_bounds__(UC_THREADS, 1) k_chain_umma(const ChainParams a, const UmmaChainExtra ex) {
    using namespace umma;
    extern __shared__ __align__(128) unsigned char uc_smem_raw[];
    UmmaChainSmem& s = *reinterpret_cast<UmmaChainSmem*>(uc_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sX = smem_u32(s.x), sH = smem_u32(s.h), sW1 = smem_u32(s.w1), sW2 = smem_u32(s.w2);
    const uint32_t bar_x = smem_u32(&s.bar_x), bar_d = smem_u32(&s.bar_d), bar_e = smem_u32(&s.bar_e);
    const uint32_t bar_h0 = smem_u32(&s.bar_h[0]);

    // ---- one-time setup ------------------------------------------------------------------------
    {
        uint4* d1 = reinterpret_cast<uint4*>(s.w1);
        uint4* d2 = reinterpret_cast<uint4*>(s.w2);
        for (int i = tid; i < UC_H * UC_K1 * 2 / 16; i += UC_THREADS) d1[i] = ex.w1_pk[i];
        for (int i = tid; i < UC_N2 * UC_H * 2 / 16; i += UC_THREADS) d2[i] = ex.w2_pk[i];
        if (tid == 0) s.timeout = 0;
    }
    if (warp == 0) tmem_alloc(smem_u32(&s.tmem_slot), 256);
    if (tid == 0) {
        mbar_init(bar_x, UC_WORKERS / 32);
        mbar_init(bar_d, 1);
        mbar_init(bar_e, 1);
        for (int g = 0; g < 4; ++g) mbar_init(bar_h0 + 8u * g, 4);
        fence_mbar_init();
    }
    fence_proxy_async();          // the weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_slot;
    const int n_steps = a.t_count;

    if (warp == UC_WORKERS / 32) {
        // ===== MMA-issue warp: one elected thread, everything it does is asynchronous ============
        if (lane == 0) {
            constexpr uint32_t IDESC1 = idesc_bf16_f32(UC_M, UC_H);
            constexpr uint32_t IDESC2 = idesc_bf16_f32(UC_M, UC_N2);
            bool ok = true;
#if UC_TIMING
            const bool timed = ex.timing != nullptr && blockIdx.x == 0;
            long long tm[4] = {0, 0, 0, 0}, c0 = 0, c1 = 0;
#endif
            for (int it = 0; it < n_steps && ok; ++it) {
                const uint32_t ph = (uint32_t)it & 1u;
                UC_T(if (timed) c0 = clock64();)
                ok = mbar_wait(bar_x, ph);
                UC_T(if (timed) { c1 = clock64(); tm[0] += c1 - c0; })
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < UC_K1 / 16; ++k)
                    mma_bf16(tmem, smem_desc(sX + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1)),
                             smem_desc(sW1 + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1)), IDESC1, k > 0);
                mma_commit(bar_d);
                UC_T(if (timed) { c0 = clock64(); tm[1] += c0 - c1; })
#pragma unroll 1
                for (int g = 0; g < 4 && ok; ++g) {
                    ok = mbar_wait(bar_h0 + 8u * g, ph);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k = 2 * g + kk;
                        mma_bf16(tmem + 128, smem_desc(sH + 2 * k * kLBO, kLBO, sbo_bytes(UC_H)),
                                 smem_desc(sW2 + 2 * k * kLBO, kLBO, sbo_bytes(UC_H)), IDESC2, k > 0);
                    }
                }
                mma_commit(bar_e);
                UC_T(if (timed) { c1 = clock64(); tm[2] += c1 - c0; })
            }
            if (!ok) s.timeout = 1;
            UC_T(if (timed) { ex.timing[8] = tm[0]; ex.timing[9] = tm[1]; ex.timing[10] = tm[2]; })
        }
    } else {
        // ===== worker warps ===========================================================================
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int g = warp >> 2;                 // column group: hidden 32g..32g+31, parameters 8g..8g+7
        const int row = quarter * 32 + lane;     // member of the tile = TMEM lane
        const int P = a.P;
        const int64_t m0 = (int64_t)blockIdx.x * UC_M;
        const bool mvalid = (m0 + row) < a.B;
        const int64_t mg = mvalid ? (m0 + row) : (a.B - 1);
        const int64_t gmember = a.member_offset + mg;
        constexpr bool shared_cond = SHARED;   // all members use one condition (n_cond == 1)
        const int d_first = a.S - a.t_hi;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        const uint32_t tD = tlane + 32 * g;              // this thread's 32 hidden columns of D
        const uint32_t tE = tlane + 128 + 8 * g;         // this thread's 8 parameter columns of E
        const uint32_t bar_h = bar_h0 + 8u * g;

        // distinct conditions: c_b of this member for the thread's hidden columns, fp32 registers
        float2 cb[16];
        if (!shared_cond) {
            const float4* cbrow = reinterpret_cast<const float4*>(a.cond_bias + (mg % a.n_cond) * UC_H + 32 * g);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 f = cbrow[i];
                cb[2 * i] = make_float2(f.x, f.y);
                cb[2 * i + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) cb[i] = make_float2(0.f, 0.f);
        }
        // threads 0..127 own row j = tid of W1aug: v = c_t[t][j] (+ c_b[j] of the shared condition)
        const float cb0 = (tid < UC_H && shared_cond) ? a.cond_bias[tid] : 0.f;
        float2 b2r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) b2r[i] = make_float2(a.b2p[8 * g + 2 * i], a.b2p[8 * g + 2 * i + 1]);

        // ---- x_T -------------------------------------------------------------------------------
        float2 x[4];
        float z[8];
        if (a.x_in) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = (8 * g + i < P) ? a.x_in[mg * a.x_in_stride + 8 * g + i] : 0.f;
        } else {
            philox_normal8(a.keys, a.offset, gmember, 0u, 2 * g, z);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = (8 * g + i < P) ? z[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = make_float2(z[2 * i], z[2 * i + 1]);

        const uint32_t xchunk = sX + (uint32_t)(row / 8) * sbo_bytes(UC_K1) + (uint32_t)(row % 8) * 16u + (uint32_t)g * kLBO;
        const uint32_t hrow = sH + (uint32_t)(row / 8) * sbo_bytes(UC_H) + (uint32_t)(row % 8) * 16u + (uint32_t)(4 * g) * kLBO;
        const uint32_t w1aug = sW1 + elem_offset(tid & (UC_H - 1), UC_AUG, UC_K1);   // (v_hi | v_mid v_lo) of row j = tid

        // operands of the next GEMM1: this thread's chunk of Xaug; threads 0..127 also refresh the
        // augmentation columns of W1 with the 3-term bf16 split of v; then one arrival per warp
        auto publish_gemm1_operands = [&](float ct) {
            if (g == 3)   // parameters 24..28 + the three constant-one columns (bf16 1.0 = 0x3F80)
                sts_u4(xchunk, pack_bf16(x[0].x, x[0].y), pack_bf16(x[1].x, x[1].y), pack_bf16(x[2].x, 1.0f), 0x3F803F80u);
            else
                sts_u4(xchunk, pack_bf16(x[0].x, x[0].y), pack_bf16(x[1].x, x[1].y), pack_bf16(x[2].x, x[2].y), pack_bf16(x[3].x, x[3].y));
            if (tid < UC_H) {
                const float v = ct + cb0;
                const float v_hi = bf16_round(v);
                const float r1 = v - v_hi;                    // exact
                const float v_mid = bf16_round(r1);
                const float v_lo = r1 - v_mid;                // exact; rounded to bf16 by the pack
                const __nv_bfloat16 hb = __float2bfloat16_rn(v_hi);
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(w1aug), "h"(*reinterpret_cast<const unsigned short*>(&hb)) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(w1aug + 2u), "r"(pack_bf16(v_mid, v_lo)) : "memory");
            }
            fence_proxy_async();
            tc_fence_before();          // this thread's TMEM reads of the step are complete
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_x);
        };
        const float* ctcol = a.table + (tid & (UC_H - 1));
        publish_gemm1_operands(tid < UC_H ? __ldg(ctcol + (int64_t)a.t_hi * UC_H) : 0.f);

        bool ok = true;
#if UC_TIMING
        const bool timed = ex.timing != nullptr && blockIdx.x == 0 && tid == 0;
        long long tw[6] = {0, 0, 0, 0, 0, 0}, k0 = 0, k1 = 0;
#endif
#pragma unroll 1
        for (int it = 0; it < n_steps && ok; ++it) {
            UC_T(if (timed) k0 = clock64();)
            const int t = a.t_hi - it;
            const int d = d_first + it;
            const uint32_t ph = (uint32_t)it & 1u;
            // prefetches: the step scalars and (threads 0..127) the next step's c_t element
            const float4 cf = __ldg(reinterpret_cast<const float4*>(a.coef) + t);
            float ct_next = 0.f;
            if (tid < UC_H && it + 1 < n_steps) ct_next = __ldg(ctcol + (int64_t)(t - 1) * UC_H);
            // ---- this step's noise (GEMM1 is in flight) ----------------------------------------------
            if (t > 0) {
                if (REPLAY) {
                    const float* zr = a.noise + ((int64_t)(d - 1) * a.noise_B + mg) * P + 8 * g;
#pragma unroll
                    for (int i = 0; i < 8; ++i) z[i] = (8 * g + i < P) ? zr[i] : 0.f;
                } else {
                    philox_normal8(a.keys, a.offset, gmember, (uint32_t)d, 2 * g, z);
                }
            }
            UC_T(if (timed) { k1 = clock64(); tw[0] += k1 - k0; })
            ok = mbar_wait(bar_d, ph);
            UC_T(if (timed) { k0 = clock64(); tw[1] += k0 - k1; })
            tc_fence_after();
            // ---- epilogue 1: h = ReLU(D [+ c_b]) -> bf16 A operand of GEMM2 ----------------------
            if (shared_cond) {
                uint32_t dv[32];
                tmem_ld32(tD, dv);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    pk[i] = pack_bf16_relu(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1]));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    sts_u4(hrow + (uint32_t)q * kLBO, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            } else {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {       // two halves of 16 columns: c_b occupies 32 registers
                    uint32_t dv[16];
                    tmem_ld16(tD + 16 * hh, dv);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 hsum = fadd2(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), cb[8 * hh + i]);
                        pk[i] = pack_bf16_relu(hsum.x, hsum.y);
                    }
                    sts_u4(hrow + (uint32_t)(2 * hh) * kLBO, pk[0], pk[1], pk[2], pk[3]);
                    sts_u4(hrow + (uint32_t)(2 * hh + 1) * kLBO, pk[4], pk[5], pk[6], pk[7]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
            UC_T(if (timed) { k1 = clock64(); tw[2] += k1 - k0; })
            UC_T(if (timed) { k0 = clock64(); tw[3] += k0 - k1; })
            ok = ok && mbar_wait(bar_e, ph);
            UC_T(if (timed) { k1 = clock64(); tw[4] += k1 - k0; })
            tc_fence_after();
            // ---- epilogue 2: eps -> posterior update of this thread's parameters ----------------
            {
                uint32_t ev[8];
                tmem_ld8(tE, ev);
                tmem_ld_wait();
                // ECD.py:111-118 with separately rounded operations (packed f32x2 ops are IEEE rn):
                //   u = coef*eps ; v = x - u ; x' = c1*v ; [ w = sigma*z ; x' = x' + w ]
                // (x - u is formed as x + (-coef)*eps: negation commutes with rounding)
                const float2 ncoef = make_float2(-cf.x, -cf.x), c1 = make_float2(cf.y, cf.y), sg = make_float2(cf.z, cf.z);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 e = fadd2(make_float2(__uint_as_float(ev[2 * i]), __uint_as_float(ev[2 * i + 1])), b2r[i]);
                    float2 xn = fmul2(c1, fadd2(x[i], fmul2(ncoef, e)));
                    if (t > 0) xn = fadd2(xn, fmul2(sg, make_float2(z[2 * i], z[2 * i + 1])));
                    x[i] = xn;
                    if (TRACE && mvalid) {
                        float* dst = a.eps_trace + ((int64_t)t * a.B + m0 + row) * P + 8 * g + 2 * i;
                        if (8 * g + 2 * i < P) dst[0] = e.x;
                        if (8 * g + 2 * i + 1 < P) dst[1] = e.y;
                    }
                }
            }
            if (it + 1 < n_steps) publish_gemm1_operands(ct_next);
            UC_T(if (timed) tw[5] += clock64() - k1;)
        }
        if (!ok) s.timeout = 1;
#if UC_TIMING
        if (timed) {
#pragma unroll
            for (int i = 0; i < 6; ++i) ex.timing[i] = tw[i];
            ex.timing[15] = n_steps;
        }
#endif
        // parameters >= P of the padded tile carry finite garbage that the zero weight columns
        // ignore; they are never stored
        if (mvalid) {
            float* dst = a.x_out + (m0 + row) * P + 8 * g;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (8 * g + 2 * i < P) dst[2 * i] = x[i].x;
                if (8 * g + 2 * i + 1 < P) dst[2 * i + 1] = x[i].y;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (s.timeout != 0) {                   // an MMA never completed: poison the tile's output
        const int64_t m0 = (int64_t)blockIdx.x * UC_M;
        for (int i = tid; i < UC_M * a.P; i += UC_THREADS)
            if (m0 + i / a.P < a.B) a.x_out[m0 * a.P + i] = __int_as_float(0x7fc00000);
        if (tid == 0) ex.status[0] = 1;
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace ertdiff
