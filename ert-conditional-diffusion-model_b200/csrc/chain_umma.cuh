// Tensor-core version of the reverse chain for large ensembles (precision = bf16 operands,
// fp32 accumulation; BASELINE configs 3/4): the two projections of the hoisted denoiser step run
// as tcgen05.mma with accumulators in TMEM, everything between them stays on chip.
//
// One CTA = a tile of 128 members for all steps; thread i <-> member i <-> TMEM lane i.
//   GEMM1  D[128 members x 128 hidden] = Xaug[128 x 32] * W0xaug[128 x 32]^T          (2 MMAs, K=16)
//          Xaug row  = [x_0..x_28, 0, 1, 1]            (bf16, rewritten by its thread every step)
//          W0xaug row j = [W0x[j][0..28], 0, ct_hi[j], ct_lo[j]]   -- the per-step vector c_t is
//          folded into the contraction through two spare K columns (bf16 hi + lo parts), i.e. the
//          "embedding add" happens inside the MMA
//   epi 1  h = ReLU(D + c_b[member])   c_b (fp32) lives in TMEM columns 128..255 for the whole
//          chain; h is written as bf16 into the K-major A operand of GEMM2
//   GEMM2  E[128 members x 32] = Hbf16[128 x 128] * W2pad[32 x 128]^T                  (8 MMAs, K=16)
//   epi 2  eps = E + b2; the thread applies the bit-exact posterior update to the 29 parameters of
//          its member (x stays in fp32 registers for the whole chain) with Philox / replayed noise
// TMEM: 256 columns (D: 0..127, reused as E: 0..31; c_b: 128..255)  ->  two CTAs per SM.
// Algorithmic work: 14,848 FLOP per member-step, as in the fp32 kernel (the K/N padding to
// 32/32 is not counted).
#pragma once
#include "denoiser.cuh"
#include "umma.cuh"

namespace ertdiff {

constexpr int UC_M = 128;       // members per CTA
constexpr int UC_H = 128;       // hidden_dim this kernel is built for
constexpr int UC_K1 = 32;       // padded param_dim + 2 augmentation columns
constexpr int UC_N2 = 32;       // padded param_dim

struct UmmaChainSmem {
    unsigned char x[UC_M * UC_K1 * 2];      // A of GEMM1
    unsigned char h[UC_M * UC_H * 2];       // A of GEMM2
    unsigned char w1[UC_H * UC_K1 * 2];     // B of GEMM1 (W0x augmented)
    unsigned char w2[UC_N2 * UC_H * 2];     // B of GEMM2 (W2 padded)
    float ctbuf[2][CHAIN_NB][UC_H + 4];     // staged c_t rows + step scalars
    float b2[kPPad];
    unsigned long long mbar;
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
};

template <bool REPLAY, bool TRACE>
__global__ void __launch_bounds__(UC_M, 2) k_chain_umma(const ChainParams a, const UmmaChainExtra ex) {
    using namespace umma;
    extern __shared__ __align__(128) unsigned char uc_smem_raw[];
    UmmaChainSmem& s = *reinterpret_cast<UmmaChainSmem*>(uc_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int P = a.P;
    const int64_t m0 = (int64_t)blockIdx.x * UC_M;
    const bool mvalid = (m0 + tid) < a.B;
    const int64_t mg = mvalid ? (m0 + tid) : (a.B - 1);
    const int64_t gmember = a.member_offset + mg;

    const uint32_t sX = smem_u32(s.x), sH = smem_u32(s.h), sW1 = smem_u32(s.w1), sW2 = smem_u32(s.w2);
    const uint32_t bar = smem_u32(&s.mbar);
    const uint32_t ct_a = smem_u32(&s.ctbuf[0][0][0]);

    // ---- one-time setup ------------------------------------------------------------------------
    {
        uint4* d1 = reinterpret_cast<uint4*>(s.w1);
        uint4* d2 = reinterpret_cast<uint4*>(s.w2);
        for (int i = tid; i < UC_H * UC_K1 * 2 / 16; i += UC_M) d1[i] = ex.w1_pk[i];
        for (int i = tid; i < UC_N2 * UC_H * 2 / 16; i += UC_M) d2[i] = ex.w2_pk[i];
        if (tid < kPPad) s.b2[tid] = a.b2p[tid];
        if (tid == 0) s.timeout = 0;
    }
    if (warp == 0) tmem_alloc(smem_u32(&s.tmem_slot), 256);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }

    const int nblocks = (a.t_count + CHAIN_NB - 1) / CHAIN_NB;
    const int d_first = a.S - a.t_hi;
    auto stage_block = [&](int b) {
        if (b < nblocks) {
            const int buf = b & 1, it0 = b * CHAIN_NB;
            const int r = tid / (UC_H / 4), c4 = tid % (UC_H / 4);
            if (it0 + r < a.t_count)
                cp_async16(ct_a + 4u * ((buf * CHAIN_NB + r) * (UC_H + 4) + 4 * c4),
                           a.table + (int64_t)(a.t_hi - it0 - r) * UC_H + 4 * c4);
            if (tid < CHAIN_NB && it0 + tid < a.t_count)
                cp_async16(ct_a + 4u * ((buf * CHAIN_NB + tid) * (UC_H + 4) + UC_H),
                           a.coef + 4 * (int64_t)(a.t_hi - it0 - tid));
        }
        cp_async_commit();
    };
    stage_block(0);

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's lane quarter

    // c_b of this member -> TMEM columns 128..255 (fp32, read back every step)
    {
        const float* cbrow = a.cond_bias + (mg % a.n_cond) * UC_H;
#pragma unroll 1
        for (int c = 0; c < UC_H / 32; ++c) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 f = *reinterpret_cast<const float4*>(cbrow + 32 * c + i);
                v[i] = __float_as_uint(f.x); v[i + 1] = __float_as_uint(f.y);
                v[i + 2] = __float_as_uint(f.z); v[i + 3] = __float_as_uint(f.w);
            }
            tmem_st32(tlane + 128 + 32 * c, v);
        }
        tmem_st_wait();
    }

    // ---- x_T -----------------------------------------------------------------------------------
    float x[kPPad];          // x[29..31] unused
    float z[kPPad];
    auto draw_normals = [&](uint32_t d) {
#pragma unroll
        for (int q = 0; q < 8; ++q) philox_normal4(a.seed, a.offset, gmember, d, q, &z[4 * q]);
    };
    if (a.x_in) {
#pragma unroll
        for (int p = 0; p < kPPad; ++p) x[p] = (p < P) ? a.x_in[mg * a.x_in_stride + p] : 0.f;
    } else {
        draw_normals(0);
#pragma unroll
        for (int p = 0; p < kPPad; ++p) x[p] = (p < P) ? z[p] : 0.f;
    }
    cp_async_wait<0>();
    __syncthreads();

    constexpr uint32_t IDESC1 = idesc_bf16_f32(UC_M, UC_H);
    constexpr uint32_t IDESC2 = idesc_bf16_f32(UC_M, UC_N2);
    const uint32_t xrow = sX + (uint32_t)(tid / 8) * sbo_bytes(UC_K1) + (uint32_t)(tid % 8) * 16u;
    const uint32_t hrow = sH + (uint32_t)(tid / 8) * sbo_bytes(UC_H) + (uint32_t)(tid % 8) * 16u;
    const uint32_t w1ct = sW1 + elem_offset(tid, 30, UC_K1);        // (ct_hi, ct_lo) slot of row j = tid
    uint32_t phase = 0;
    bool dead = false;

    for (int b = 0; b < nblocks && !dead; ++b) {
        const int buf = b & 1;
        stage_block(b + 1);
#pragma unroll 1
        for (int r = 0; r < CHAIN_NB; ++r) {
            const int it = b * CHAIN_NB + r;
            if (it >= a.t_count) break;
            const int t = a.t_hi - it;
            const int d = d_first + it;
            const uint32_t row_a = ct_a + 4u * ((buf * CHAIN_NB + r) * (UC_H + 4));
            // ---- operands of GEMM1 -------------------------------------------------------------
            {
                const float ct = lds32(row_a + 4u * tid);
                const float ct_hi = bf16_round(ct);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(w1ct), "r"(pack_bf16(ct_hi, ct - ct_hi)) : "memory");
                sts_u4(xrow + 0 * kLBO, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                sts_u4(xrow + 1 * kLBO, pack_bf16(x[8], x[9]), pack_bf16(x[10], x[11]), pack_bf16(x[12], x[13]), pack_bf16(x[14], x[15]));
                sts_u4(xrow + 2 * kLBO, pack_bf16(x[16], x[17]), pack_bf16(x[18], x[19]), pack_bf16(x[20], x[21]), pack_bf16(x[22], x[23]));
                sts_u4(xrow + 3 * kLBO, pack_bf16(x[24], x[25]), pack_bf16(x[26], x[27]), pack_bf16(x[28], 0.f), pack_bf16(1.f, 1.f));
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < UC_K1 / 16; ++k)
                    mma_bf16(tmem, smem_desc(sX + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1)),
                             smem_desc(sW1 + 2 * k * kLBO, kLBO, sbo_bytes(UC_K1)), IDESC1, k > 0);
                mma_commit(bar);
            }
            // ---- this step's noise, while the tensor core works --------------------------------
            if (t > 0) {
                if (REPLAY) {
                    const float* zr = a.noise + ((int64_t)(d - 1) * a.noise_B + mg) * P;
#pragma unroll
                    for (int p = 0; p < kPPad; ++p) z[p] = (p < P) ? zr[p] : 0.f;
                } else {
                    draw_normals((uint32_t)d);
                }
            }
            if (!dead && !mbar_wait(bar, phase)) { dead = true; s.timeout = 1; }
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue 1: h = ReLU(D + c_b) -> bf16 A operand of GEMM2 ----------------------
#pragma unroll 1
            for (int c = 0; c < UC_H / 32; ++c) {
                uint32_t dv[32], cv[32];
                tmem_ld32(tlane + 32 * c, dv);
                tmem_ld32(tlane + 128 + 32 * c, cv);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float h0 = fmaxf(__uint_as_float(dv[2 * i]) + __uint_as_float(cv[2 * i]), 0.f);
                    const float h1 = fmaxf(__uint_as_float(dv[2 * i + 1]) + __uint_as_float(cv[2 * i + 1]), 0.f);
                    pk[i] = pack_bf16(h0, h1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    sts_u4(hrow + (uint32_t)(4 * c + q) * kLBO, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < UC_H / 16; ++k)
                    mma_bf16(tmem, smem_desc(sH + 2 * k * kLBO, kLBO, sbo_bytes(UC_H)),
                             smem_desc(sW2 + 2 * k * kLBO, kLBO, sbo_bytes(UC_H)), IDESC2, k > 0);
                mma_commit(bar);
            }
            const float cf_coef = lds32(row_a + 4u * UC_H), cf_c1 = lds32(row_a + 4u * UC_H + 4u),
                        cf_sigma = lds32(row_a + 4u * UC_H + 8u);
            if (!dead && !mbar_wait(bar, phase)) { dead = true; s.timeout = 1; }
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue 2: eps -> posterior update of this member's parameters ---------------
            {
                uint32_t ev[32];
                tmem_ld32(tlane, ev);
                tmem_ld_wait();
#pragma unroll
                for (int p = 0; p < kPPad; ++p) {
                    if (p < P) {
                        const float e = __uint_as_float(ev[p]) + s.b2[p];
                        x[p] = posterior_update_rn(x[p], e, z[p], cf_coef, cf_c1, cf_sigma, t > 0);
                        if (TRACE && mvalid && p < P) a.eps_trace[((int64_t)t * a.B + m0 + tid) * P + p] = e;
                    }
                }
            }
            if (r == CHAIN_NB - 1) cp_async_wait<0>();
            tc_fence_before();
            __syncthreads();            // TMEM reads done before the next GEMM1 overwrites D; staged rows visible
            if (dead) break;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    const bool failed = s.timeout != 0;     // an MMA never completed: poison the tile's output
    if (mvalid) {
#pragma unroll
        for (int p = 0; p < kPPad; ++p)
            if (p < P) a.x_out[(m0 + tid) * P + p] = failed ? __int_as_float(0x7fc00000) : x[p];
    }
    if (tid == 0 && failed) ex.status[0] = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace ertdiff
