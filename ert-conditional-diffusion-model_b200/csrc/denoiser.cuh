// Hoisted denoiser step + DDPM posterior update + the reverse loop (ECD.py:102-119, 155-164).
//
// The reference evaluates, per step and per member,
//     h   = ReLU(W0 @ [x | t_emb | c_emb] + b0),   eps = W2 @ h + b2          (ECD.py:159-163)
// W0 @ [x|t|c] = W0x @ x + W0t @ t_emb + W0c @ c_emb.  In a chain all members share t, and a
// member's condition never changes, so
//     c_t[t]  = W0t @ ReLU(Wt @ emb(t) + bt)            one H-vector per step  (k_time_table)
//     c_b[m]  = W0c @ c_emb[m] + b0                     one H-vector per member (encoder.cuh)
//     h       = ReLU(W0x @ x + c_t[t] + c_b[m])         14,848 FLOP per member per step
// k_chain keeps x, the weights and both tables' current rows on chip for all steps of a tile of
// members: one launch runs the whole chain (no grid-wide dependency exists between members).
#pragma once
#include "common.cuh"
#include "chain_params.cuh"

namespace ertdiff {

// ------------------------------------------------------------------------------------------
// Posterior update, ECD.py:111-118.  Every operation is a separately rounded fp32 op (the
// _rn intrinsics are never contracted into FMAs), in the reference's order:
//   u = coef*eps ; v = x - u ; x' = c1*v ; [ w = sigma*z ; x' = x' + w ]
__device__ __forceinline__ float posterior_update_rn(float x, float eps, float z, float coef,
                                                     float c1, float sigma, bool add_noise) {
    const float u = __fmul_rn(coef, eps);
    const float v = __fsub_rn(x, u);
    float xn = __fmul_rn(c1, v);
    if (add_noise) xn = __fadd_rn(xn, __fmul_rn(sigma, z));
    return xn;
}

static __global__ void k_posterior_update(const float* __restrict__ x, const float* __restrict__ eps,
                                   const float* __restrict__ z, float coef, float c1,
                                   float sigma, int64_t n, float* __restrict__ out) {
    // 4 elements per thread, 128-bit accesses when the pointers allow it
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    const bool vec = (i4 + 4 <= n) &&
                     (((uintptr_t)x | (uintptr_t)eps | (uintptr_t)z | (uintptr_t)out) & 15) == 0;
    if (vec) {
        const float4 xv = *reinterpret_cast<const float4*>(x + i4);
        const float4 ev = *reinterpret_cast<const float4*>(eps + i4);
        float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (z) zv = *reinterpret_cast<const float4*>(z + i4);
        float4 o;
        o.x = posterior_update_rn(xv.x, ev.x, zv.x, coef, c1, sigma, z != nullptr);
        o.y = posterior_update_rn(xv.y, ev.y, zv.y, coef, c1, sigma, z != nullptr);
        o.z = posterior_update_rn(xv.z, ev.z, zv.z, coef, c1, sigma, z != nullptr);
        o.w = posterior_update_rn(xv.w, ev.w, zv.w, coef, c1, sigma, z != nullptr);
        *reinterpret_cast<float4*>(out + i4) = o;
    } else {
        for (int64_t i = i4; i < n && i < i4 + 4; ++i)
            out[i] = posterior_update_rn(x[i], eps[i], z ? z[i] : 0.f, coef, c1, sigma,
                                         z != nullptr);
    }
}

// ------------------------------------------------------------------------------------------
// Per-step scalars with the reference's rounding: (1-alpha) and (1-alpha_bar)
// are fp32 subtractions; math.sqrt works in double; the quotient and the two other scalars are
// rounded to fp32 once.  double sqrt and division are correctly rounded on the device, so the
// table is bit-identical to what the reference's python arithmetic produces.
__device__ __forceinline__ void step_coefficients(const float* betas, const float* alphas,
                                                  const float* alpha_bar, int t,
                                                  double temperature, float* out4) {
    const float oma = __fsub_rn(1.0f, alphas[t]);
    const float omab = __fsub_rn(1.0f, alpha_bar[t]);
    const double denom = __dadd_rn(sqrt((double)omab), 1e-8);
    out4[0] = __fdiv_rn(oma, (float)denom);
    out4[1] = (float)(1.0 / sqrt((double)alphas[t]));
    out4[2] = (float)__dmul_rn(sqrt((double)betas[t]), temperature);
    out4[3] = 0.f;
}

static __global__ void k_step_coefficients(const float* betas, const float* alphas,
                                    const float* alpha_bar, int steps, double temperature,
                                    float* table4) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < steps) step_coefficients(betas, alphas, alpha_bar, t, temperature, table4 + 4 * t);
}

// ------------------------------------------------------------------------------------------
// Timestep embedding (ECD.py:80-88) -> time_embed Linear+ReLU (ECD.py:144-147,160) ->
// t-block of mlp.0 : table[row] = W0t @ ReLU(Wt @ [sin(t f) | cos(t f)] + bt).
// grid = rows, block = H.  Row r is timestep t0 + r.
__device__ __forceinline__ void time_embed_row(float tf, const float* __restrict__ freq,
                                               const float* __restrict__ wtT,
                                               const float* __restrict__ bt, int H, int tid,
                                               float* emb /*smem H*/, float* te /*smem H*/) {
    const int half = H >> 1;
    if (tid < half) {
        const float arg = __fmul_rn(tf, freq[tid]);
        emb[tid] = sinf(arg);
        emb[half + tid] = cosf(arg);
    }
    __syncthreads();
    float a0 = bt[tid], a1 = 0.f;
    for (int k = 0; k < H; k += 2) {
        a0 = fmaf(wtT[(int64_t)k * H + tid], emb[k], a0);
        a1 = fmaf(wtT[(int64_t)(k + 1) * H + tid], emb[k + 1], a1);
    }
    te[tid] = fmaxf(a0 + a1, 0.f);
    __syncthreads();
}

static __global__ void k_time_table(const float* __restrict__ freq, const float* __restrict__ wtT,
                             const float* __restrict__ bt, const float* __restrict__ w0tT,
                             int H, int t0, float* __restrict__ table) {
    __shared__ float emb[512];
    __shared__ float te[512];
    const int tid = threadIdx.x;
    const int t = t0 + blockIdx.x;
    time_embed_row((float)t, freq, wtT, bt, H, tid, emb, te);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < H; k += 4) {
        a0 = fmaf(w0tT[(int64_t)(k + 0) * H + tid], te[k + 0], a0);
        a1 = fmaf(w0tT[(int64_t)(k + 1) * H + tid], te[k + 1], a1);
        a2 = fmaf(w0tT[(int64_t)(k + 2) * H + tid], te[k + 2], a2);
        a3 = fmaf(w0tT[(int64_t)(k + 3) * H + tid], te[k + 3], a3);
    }
    table[(int64_t)t * H + tid] = (a0 + a1) + (a2 + a3);
}

// ------------------------------------------------------------------------------------------
// Full forward for rows with their own t (ECD.py:155-164 as written; training passes random
// t per row, ECD.py:312-315).  cond_bias already holds W0c @ c_emb + b0 for the row.
// grid = B, block = H.
static __global__ void k_forward_rows(const float* __restrict__ x, const int64_t* __restrict__ t,
                               const float* __restrict__ cond_bias, int64_t n_cond,
                               const float* __restrict__ freq, const float* __restrict__ wtT,
                               const float* __restrict__ bt, const float* __restrict__ w0tT,
                               const float* __restrict__ w0xT, const float* __restrict__ w2p,
                               const float* __restrict__ b2p, int P, int H,
                               float* __restrict__ out) {
    __shared__ float emb[512];
    __shared__ float te[512];
    __shared__ float xs[kPPad];
    __shared__ float hs[512];
    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x;
    if (tid < kPPad) xs[tid] = (tid < P) ? x[row * P + tid] : 0.f;
    time_embed_row((float)t[row], freq, wtT, bt, H, tid, emb, te);
    float a0 = cond_bias[(row % n_cond) * H + tid], a1 = 0.f;
    for (int k = 0; k < H; k += 2) {
        a0 = fmaf(w0tT[(int64_t)k * H + tid], te[k], a0);
        a1 = fmaf(w0tT[(int64_t)(k + 1) * H + tid], te[k + 1], a1);
    }
#pragma unroll
    for (int k = 0; k < kPPad; ++k) a0 = fmaf(w0xT[k * H + tid], xs[k], a0);
    hs[tid] = fmaxf(a0 + a1, 0.f);
    __syncthreads();
    // output p handled by warp (p % nwarps): lanes stride over j, then a shuffle tree
    const int lane = tid & 31, warp = tid >> 5, nwarps = H >> 5;
    for (int p = warp; p < P; p += nwarps) {
        float acc = 0.f;
        for (int j = lane; j < H; j += 32) acc = fmaf(w2p[(int64_t)p * H + j], hs[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[row * P + p] = acc + b2p[p];
    }
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG + Box-Muller: device-side replacement for torch.randn in
// ECD.py:107,116 (a CPU mt19937 stream cannot be reproduced on the device; parity runs inject
// noise instead).  Stream layout: one Philox call yields the normals of ONE draw (draw 0 = x_T,
// draw k = k-th in-loop draw) for FOUR consecutive parameters of one member:
//     counter = (member_lo, offset_hi ^ member_hi, draw * 8 + pquad, offset_lo), key = seed
// so every kernel -- whatever its thread-to-work mapping -- and any sharding of the members
// draws identical numbers for (member, draw, parameter).
__device__ __forceinline__ void mul_wide_u32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    // one IMAD.WIDE.U32 for both halves of the product
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& ks) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mul_wide_u32(M0, c.x, hi0, lo0);
        mul_wide_u32(M1, c.z, hi1, lo1);
        c = make_uint4(hi1 ^ c.y ^ ks.k[2 * r], lo1, hi0 ^ c.w ^ ks.k[2 * r + 1], lo0);
    }
    return c;
}

// Box-Muller on the special-function unit: lg2 / sqrt / sin / cos are single MUFU instructions
// (abs. error ~1e-6 on a N(0,1) variate, far below what any statistic of the chain resolves).
// The generator runs once per member, step and parameter, so its instruction count -- and above all
// its count of XU-pipe operations (MUFU and int->float conversions share that 16-lane pipe) -- is what
// the chain kernels pay for.  The radius keeps all 32 random bits (one I2F; tails to 6.7 sigma, as
// torch's CUDA generator); the angle takes 23 bits straight into the mantissa of a float in [1, 2)
// with one logic op instead of a conversion: 5 XU operations per pair of normals instead of 6.
// Every kernel calls this one function, hence all of them (and k_philox_fill, which the parity tests
// replay) produce bit-identical draws.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0,1]
    // angle = 2*pi*(m - 1.5) in [-pi, pi), m = 1.mantissa(b): the MUFU sin/cos need no further range reduction
    const float m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    const float ang = fmaf(m, 6.2831853071795865f, -9.4247779607693797f);
    float lg, r, sn, cs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));            // u1 >= 2^-33: never denormal
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
    n0 = r * cs;
    n1 = r * sn;
}

// normals of draw `draw` for parameters 4*pquad .. 4*pquad+3 of `member`
__device__ __forceinline__ void philox_normal4(const PhiloxKeys& ks, uint64_t offset, int64_t member,
                                               uint32_t draw, uint32_t pquad, float out[4]) {
    const uint4 c = make_uint4((uint32_t)member, (uint32_t)(offset >> 32) ^ (uint32_t)(member >> 32),
                               (draw << 3) | pquad, (uint32_t)offset);
    const uint4 r = philox4x32_10(c, ks);
    box_muller(r.x, r.y, out[0], out[1]);
    box_muller(r.z, r.w, out[2], out[3]);
}
// two neighbouring quads at once: four independent multiply chains keep the pipes busy
__device__ __forceinline__ void philox_normal8(const PhiloxKeys& ks, uint64_t offset, int64_t member,
                                               uint32_t draw, uint32_t pquad0, float out[8]) {
    const uint32_t cy = (uint32_t)(offset >> 32) ^ (uint32_t)(member >> 32);
    const uint4 ra = philox4x32_10(make_uint4((uint32_t)member, cy, (draw << 3) | pquad0, (uint32_t)offset), ks);
    const uint4 rb = philox4x32_10(make_uint4((uint32_t)member, cy, (draw << 3) | (pquad0 + 1), (uint32_t)offset), ks);
    box_muller(ra.x, ra.y, out[0], out[1]);
    box_muller(ra.z, ra.w, out[2], out[3]);
    box_muller(rb.x, rb.y, out[4], out[5]);
    box_muller(rb.z, rb.w, out[6], out[7]);
}

// all 32 draws of one (member, draw) at once: eight independent Philox chains advance round by round, so that the
// multiply / xor latency of one chain is covered by the seven others (the noise warps of the tensor-core chain are
// bound by this latency, not by issue slots)
__device__ __forceinline__ void philox_normal32(const PhiloxKeys& ks, uint64_t offset, int64_t member, uint32_t draw,
                                                float out[32]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t cy = (uint32_t)(offset >> 32) ^ (uint32_t)(member >> 32);
    uint4 c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = make_uint4((uint32_t)member, cy, (draw << 3) | (uint32_t)q, (uint32_t)offset);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t hi0, lo0, hi1, lo1;
            mul_wide_u32(M0, c[q].x, hi0, lo0);
            mul_wide_u32(M1, c[q].z, hi1, lo1);
            c[q] = make_uint4(hi1 ^ c[q].y ^ ks.k[2 * r], lo1, hi0 ^ c[q].w ^ ks.k[2 * r + 1], lo0);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        box_muller(c[q].x, c[q].y, out[4 * q], out[4 * q + 1]);
        box_muller(c[q].z, c[q].w, out[4 * q + 2], out[4 * q + 3]);
    }
}

// standalone generator with the chain's stream layout (tests compare the chain in device-RNG
// mode against the oracle fed with these very draws): out[d][m][p], d = draw index.
static __global__ void k_philox_fill(const PhiloxKeys keys, uint64_t offset, int64_t member_offset, int64_t B,
                              int P, int draws, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (d, m, pquad)
    if (idx >= (int64_t)draws * B * 8) return;
    const uint32_t q = (uint32_t)(idx & 7);
    const int64_t m = (idx >> 3) % B;
    const uint32_t d = (uint32_t)((idx >> 3) / B);
    float z[4];
    philox_normal4(keys, offset, member_offset + m, d, q, z);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int p = 4 * q + u;
        if (p < P) out[((int64_t)d * B + m) * P + p] = z[u];
    }
}

// ---- small PTX helpers for the chain kernel ----------------------------------------------------
__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n"
                 ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// packed fp32x2 FMA (sm_100: one instruction, two IEEE fp32 FMAs) -- halves the issue slots of the
// two matrix-vector products; each lane of the pair is an ordinary fma.rn.f32
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;\n"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

constexpr int CHAIN_NB = 4;     // steps per staging block (double-buffered, cp.async one block ahead)

// minimum resident CTAs the register allocation must allow.  The one-member variant is the latency-critical
// small-ensemble kernel: with the default cap ptxas interleaves each shared-memory load with its consumer and
// exposes 8 load latencies per layer; a looser cap lets it issue all loads of a layer first (0.34 -> 0.28 us
// per step for a lone CTA).  Two hidden units per thread hold twice the weights: no cap beyond the block size.
__host__ __device__ constexpr int chain_min_blocks(int H, int MPB, int UPT) {
    return UPT > 1 ? 1 : ((H <= 128 && MPB <= 2) ? 2 : 0);
}

// One CTA = MPB members for all steps.  blockDim = NT = H / UPT (UPT hidden units per thread).
//   layer 1: thread j owns hidden units j, j + NT, ... (W0x rows in registers, x broadcast from smem)
//   layer 2: thread (p = tid / PARTS, part = tid % PARTS), PARTS = NT / 32, owns a 32*UPT-wide slice of
//            W2 row p; the PARTS partial sums are combined with xor-shuffles
//   update : every lane of a p-group computes it; the part-0 lane ("owner") writes x back.
// UPT = 2 halves the warps a member occupies (two warps at H = 128): a small ensemble that puts two members on
// most SMs then runs one warp per scheduler, and layer 2 needs one shuffle level less.
// Everything a step reads from global memory -- the c_t row, the three step scalars and (replay
// mode) the injected noise row -- is staged one block of CHAIN_NB steps ahead into shared memory
// with cp.async, so the dependent chain of a step never waits on L2/HBM latency.
// Device RNG: every RP = min(NT,128)/8 steps the first 8*RP threads run Philox once per member
// (thread -> (draw within the block, parameter quad)) and park the normals of the next RP draws
// in shared memory, where the owner lanes pick them up.
// FLOOR: the same kernel with the two matrix-vector products removed (every load, store, barrier, shuffle,
// the staging, the RNG and the posterior update stay): the latency floor of this structure, measured.
template <int H, int MPB, int UPT, bool REPLAY, bool TRACE, bool FLOOR>
__global__ void __launch_bounds__(H / UPT, chain_min_blocks(H, MPB, UPT)) k_chain(const ChainParams a) {
    constexpr int NT = H / UPT;
    constexpr int PARTS = NT / 32;
    constexpr int SLICE = 32 * UPT;         // width of this thread's slice of a W2 row
    constexpr int RP = (NT >= 128 ? 128 : NT) / 8;   // draws per RNG refill
    constexpr int HS_STRIDE = (H / 32) * 36;   // each 32-wide slice padded to 36: no bank conflicts
    constexpr int CT_STRIDE = H + 4;        // c_t row + [coef, c1, sigma, 0]
    constexpr int ZB = REPLAY ? 1 : 0;
    static_assert(NT >= 32 && NT % 32 == 0, "at least one warp, whole warps");
    static_assert(UPT == 1 || UPT == 2, "one or two hidden units per thread");
    __shared__ __align__(16) float xs[MPB][kPPad];
    __shared__ __align__(16) float hs[MPB][HS_STRIDE];
    __shared__ __align__(16) float ctbuf[2][CHAIN_NB][CT_STRIDE];
    __shared__ __align__(16) float zbuf[ZB ? 2 : 1][ZB ? CHAIN_NB : 1][ZB ? MPB : 1][kPPad];
    __shared__ __align__(16) float znorm[ZB ? 1 : MPB][ZB ? 1 : RP][kPPad];

    const int tid = threadIdx.x;
    const int p = tid / PARTS, part = tid % PARTS;
    const int64_t m0 = (int64_t)blockIdx.x * MPB;
    const int P = a.P;
    const bool owner = (part == 0) && (p < P);

    float2 w0x[UPT][kPPad / 2], w2r[SLICE / 2];
#pragma unroll
    for (int u = 0; u < UPT; ++u)
#pragma unroll
        for (int k = 0; k < kPPad / 2; ++k)
            w0x[u][k] = make_float2(a.w0xT[(2 * k) * H + tid + u * NT], a.w0xT[(2 * k + 1) * H + tid + u * NT]);
#pragma unroll
    for (int i = 0; i < SLICE / 2; ++i)
        w2r[i] = make_float2(a.w2p[p * H + part * SLICE + 2 * i], a.w2p[p * H + part * SLICE + 2 * i + 1]);
    const float b2 = a.b2p[p];

    // 32-bit shared-window addresses, computed once
    uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(&xs[0][0]);
    uint32_t hs_a = (uint32_t)__cvta_generic_to_shared(&hs[0][0]);
    uint32_t ct_a = (uint32_t)__cvta_generic_to_shared(&ctbuf[0][0][0]);
    uint32_t zb_a = (uint32_t)__cvta_generic_to_shared(&zbuf[0][0][0][0]);
    uint32_t zn_a = (uint32_t)__cvta_generic_to_shared(&znorm[0][0][0]);
    // make the bases opaque: otherwise the compiler re-derives them (S2UR SR_CgaCtaId + ULEA,
    // a long-latency special-register read) inside every step
    asm volatile("" : "+r"(xs_a), "+r"(hs_a), "+r"(ct_a), "+r"(zb_a), "+r"(zn_a));
    const uint32_t hs_w = hs_a + 4u * ((tid >> 5) * 36 + (tid & 31));   // this thread's first h slot (member 0)
    const uint32_t hs_r = hs_a + 4u * (part * UPT * 36);                // this thread's W2 slice of h
    const uint32_t xs_w = xs_a + 4u * p;

    float cb[MPB][UPT], x[MPB];
    int64_t mg[MPB];      // clamped local member index
    bool mvalid[MPB];
#pragma unroll
    for (int m = 0; m < MPB; ++m) {
        mvalid[m] = (m0 + m) < a.B;
        mg[m] = mvalid[m] ? (m0 + m) : (a.B - 1);
#pragma unroll
        for (int u = 0; u < UPT; ++u) cb[m][u] = a.cond_bias[(mg[m] % a.n_cond) * H + tid + u * NT];
    }

    // ---- block staging ---------------------------------------------------------------------
    const int nblocks = (a.t_count + CHAIN_NB - 1) / CHAIN_NB;
    const int d_first = a.S - a.t_hi;        // draw index used by the first step of this launch
    auto stage_block = [&](int b) {          // iterations [NB*b, NB*b+NB): t = t_hi - it
        if (b < nblocks) {
            const int buf = b & 1;
            const int it0 = b * CHAIN_NB;
            // c_t rows: H/4 16-byte chunks per row, NB rows -> UPT chunks per thread
#pragma unroll
            for (int c = tid; c < CHAIN_NB * (H / 4); c += NT) {
                const int r = c / (H / 4), c4 = c % (H / 4);
                if (it0 + r < a.t_count)
                    cp_async16(ct_a + 4u * ((buf * CHAIN_NB + r) * CT_STRIDE + 4 * c4),
                               a.table + (int64_t)(a.t_hi - it0 - r) * H + 4 * c4);
            }
            if (tid < CHAIN_NB && it0 + tid < a.t_count)
                cp_async16(ct_a + 4u * ((buf * CHAIN_NB + tid) * CT_STRIDE + H),
                           a.coef + 4 * (int64_t)(a.t_hi - it0 - tid));
            if (REPLAY && owner) {
#pragma unroll
                for (int r = 0; r < CHAIN_NB; ++r) {
                    const int it = it0 + r;
                    if (it < a.t_count && a.t_hi - it > 0) {
                        const int64_t row = (int64_t)(d_first + it - 1) * a.noise_B;
#pragma unroll
                        for (int m = 0; m < MPB; ++m)
                            cp_async4(zb_a + 4u * (((buf * CHAIN_NB + r) * MPB + m) * kPPad + p),
                                      a.noise + (row + mg[m]) * P + p);
                    }
                }
            }
        }
        cp_async_commit();
    };
    stage_block(0);

    // ---- device RNG ----------------------------------------------------------------------------
    // draws [blk*RP, blk*RP + RP): thread (dd = tid / 8, q = tid % 8) generates draw blk*RP + dd
    // for parameters 4q..4q+3 of every member of the CTA -> znorm[m][dd][4q..4q+3]
    auto refill_rng = [&](int d) {
        const uint32_t d0 = ((uint32_t)d / RP) * RP;
        if (tid < 8 * RP) {
            const uint32_t dd = tid >> 3, q = tid & 7;
#pragma unroll
            for (int m = 0; m < MPB; ++m) {
                float z4[4];
                philox_normal4(a.keys, a.offset, a.member_offset + mg[m], d0 + dd, q, z4);
                sts128(zn_a + 4u * ((m * RP + dd) * kPPad + 4 * q), make_float4(z4[0], z4[1], z4[2], z4[3]));
            }
        }
    };
    auto rng_draw = [&](int d, int m) -> float {
        return lds32(zn_a + 4u * ((m * RP + (d % RP)) * kPPad + (p & 31)));
    };
    // (a launch without x_in starts a chain: d_first == 1, and draw 0 = x_T sits in the same block)
    if (!REPLAY) {
        refill_rng(a.x_in ? d_first : 0);
        __syncthreads();
    }

    // ---- x_T -----------------------------------------------------------------------------------
#pragma unroll
    for (int m = 0; m < MPB; ++m) {
        if (a.x_in) {
            x[m] = (p < P) ? a.x_in[mg[m] * a.x_in_stride + p] : 0.f;
        } else if (!REPLAY) {
            const float z0 = rng_draw(0, m);
            x[m] = (p < P) ? z0 : 0.f;
        } else {
            x[m] = 0.f;       // replay launches always carry x_in
        }
        if (part == 0) sts32(xs_w + 4u * (m * kPPad), x[m]);
    }
    cp_async_wait<0>();
    __syncthreads();

    for (int b = 0; b < nblocks; ++b) {
        const int buf = b & 1;
        stage_block(b + 1);           // into the buffer block b-1 used (all its readers are past the barrier)
#pragma unroll
        for (int r = 0; r < CHAIN_NB; ++r) {
            const int it = b * CHAIN_NB + r;
            if (it >= a.t_count) break;
            const int t = a.t_hi - it;
            const int d = d_first + it;          // draw index of this step's noise
            const uint32_t row_a = ct_a + 4u * ((buf * CHAIN_NB + r) * CT_STRIDE);
            float ct[UPT];
#pragma unroll
            for (int u = 0; u < UPT; ++u) ct[u] = lds32(row_a + 4u * (tid + u * NT));
            // next block of normals: written here, published by the barrier after layer 1, first
            // read after it (the previous block's last reader finished before the last barrier)
            if (!REPLAY && t > 0 && (d % RP) == 0 && it > 0) refill_rng(d);
            // ---- layer 1 -----------------------------------------------------------------
#pragma unroll
            for (int m = 0; m < MPB; ++m) {
                float4 xv[kPPad / 4];
#pragma unroll
                for (int k4 = 0; k4 < kPPad / 4; ++k4) xv[k4] = lds128(xs_a + 16u * (m * (kPPad / 4) + k4));
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    float hval;
                    if (FLOOR) {
                        hval = (cb[m][u] + ct[u]) + xv[kPPad / 4 - 1].w;
                    } else {
                        float2 A01 = make_float2(cb[m][u] + ct[u], 0.f), A23 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int k4 = 0; k4 < kPPad / 4; ++k4) {
                            A01 = ffma2(w0x[u][2 * k4], make_float2(xv[k4].x, xv[k4].y), A01);
                            A23 = ffma2(w0x[u][2 * k4 + 1], make_float2(xv[k4].z, xv[k4].w), A23);
                        }
                        hval = (A01.x + A01.y) + (A23.x + A23.y);
                    }
                    sts32(hs_w + 4u * (m * HS_STRIDE + u * PARTS * 36), fmaxf(hval, 0.f));
                }
            }
            __syncthreads();          // hs complete (and, at block starts, staged rows are visible)
            // ---- layer 2 + posterior update --------------------------------------------------
            // (loads are issued in the order their consumers need them: h first, scalars last)
            float cf_coef, cf_c1, cf_sigma;
#pragma unroll
            for (int m = 0; m < MPB; ++m) {
                float4 hv[SLICE / 4];
#pragma unroll
                for (int u = 0; u < UPT; ++u)
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4)
                        hv[8 * u + i4] = lds128(hs_r + 4u * (m * HS_STRIDE + u * 36 + 4 * i4));
                float z = 0.f;
                if (t > 0) {
                    if (REPLAY) { if (owner) z = lds32(zb_a + 4u * (((buf * CHAIN_NB + r) * MPB + m) * kPPad + p)); }
                    else z = rng_draw(d, m);
                }
                if (m == 0) {
                    cf_coef = lds32(row_a + 4u * H);
                    cf_c1 = lds32(row_a + 4u * H + 4u);
                    cf_sigma = lds32(row_a + 4u * H + 8u);
                }
                float e;
                if (FLOOR) {
                    e = hv[SLICE / 4 - 1].w;
                } else {
                    float2 E01 = make_float2(0.f, 0.f), E23 = make_float2(0.f, 0.f);
                    if (UPT == 1) {
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            E01 = ffma2(w2r[2 * i4], make_float2(hv[i4].x, hv[i4].y), E01);
                            E23 = ffma2(w2r[2 * i4 + 1], make_float2(hv[i4].z, hv[i4].w), E23);
                        }
                        e = (E01.x + E01.y) + (E23.x + E23.y);
                    } else {
                        // two 32-wide sub-slices: four independent accumulator chains of 8
                        float2 F01 = make_float2(0.f, 0.f), F23 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            E01 = ffma2(w2r[2 * i4], make_float2(hv[i4].x, hv[i4].y), E01);
                            E23 = ffma2(w2r[2 * i4 + 1], make_float2(hv[i4].z, hv[i4].w), E23);
                            F01 = ffma2(w2r[16 + 2 * i4], make_float2(hv[8 + i4].x, hv[8 + i4].y), F01);
                            F23 = ffma2(w2r[16 + 2 * i4 + 1], make_float2(hv[8 + i4].z, hv[8 + i4].w), F23);
                        }
                        e = ((E01.x + E01.y) + (E23.x + E23.y)) + ((F01.x + F01.y) + (F23.x + F23.y));
                    }
                }
#pragma unroll
                for (int o = 1; o < PARTS; o <<= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
                e += b2;
                x[m] = posterior_update_rn(x[m], e, z, cf_coef, cf_c1, cf_sigma, t > 0);
                if (owner) {
                    sts32(xs_w + 4u * (m * kPPad), x[m]);
                    if (TRACE && mvalid[m]) a.eps_trace[((int64_t)t * a.B + m0 + m) * P + p] = e;
                }
            }
            if (r == CHAIN_NB - 1) cp_async_wait<0>();   // next block's rows landed (this thread's)
            __syncthreads();
        }
    }
    cp_async_wait<0>();
    if (owner) {
#pragma unroll
        for (int m = 0; m < MPB; ++m)
            if (mvalid[m]) a.x_out[(m0 + m) * P + p] = x[m];
    }
}

}  // namespace ertdiff
