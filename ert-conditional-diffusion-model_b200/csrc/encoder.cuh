// Condition encoder (ECD.py:133-142), fp32 CUDA-core version.
//
//   cond (14, L) -> conv1(k3,s2,p1)+ReLU (32, L1) -> conv2(k3,s2,p1)+ReLU (64, L2)
//        -> mean over L2 (64) -> Linear(64,H)+ReLU = cond_emb (H)
//        -> cond_bias = mlp.0.weight[:, P+H:] @ cond_emb + mlp.0.bias  (H)
//
// k_encoder_conv fuses both convolutions and the pooling partial sums: one CTA owns ENC_TP
// conv2 positions of one condition; the two intermediates (ECD.py's (B,32,2347) and
// (B,64,1174) tensors) live only in shared memory.  The stride-2 taps are turned into
// unit-stride shared-memory reads by storing the input in 4 phases (l mod 4) and the conv1
// output in 2 phases (even/odd position):
//   h1[2m]   = W0*in[4m-1] + W1*in[4m]   + W2*in[4m+1]   = W0*ph3[m-1] + W1*ph0[m] + W2*ph1[m]
//   h1[2m+1] = W0*in[4m+1] + W1*in[4m+2] + W2*in[4m+3]   = W0*ph1[m]   + W1*ph2[m] + W2*ph3[m]
//   h2[p]    = V0*h1[2p-1] + V1*h1[2p]   + V2*h1[2p+1]   = V0*h1o[p-1] + V1*h1e[p] + V2*h1o[p]
// Pooling is deterministic: per-chunk partial sums go to scratch and k_encoder_finish adds
// them in chunk order.
//
// Algorithmic work per condition: 20,751,232 FLOP (conv1 6.3 M + conv2 14.4 M + linear);
// HBM traffic: 14*L*4 bytes read once (262,808 B at L=4693) -- the 4-element halo between
// neighbouring chunks is served by L2.
#pragma once
#include "common.cuh"

namespace ertdiff {

constexpr int ENC_THREADS = 256;     // 8 warps

// NJ = conv2 positions per lane; a CTA owns ENC_TP = 32*NJ positions.  NJ = 4 for throughput
// (many conditions), NJ = 1 to spread a single condition over ~37 CTAs instead of 10.
template <int NJ>
struct EncSmem {
    static constexpr int TP = 32 * NJ;
    static constexpr int STRIDE = TP + 4;        // >= TP + 1, keeps rows 16-byte aligned
    float in_ph[4][kInChannels][STRIDE];         // in_ph[r][ci][i] = in[ci][4*(p0-1+i) + r]
    float h1e[kConv1Out][TP];                    // h1e[c][i] = h1[c][2*(p0+i)]
    float h1o[kConv1Out][STRIDE];                // h1o[c][i] = h1[c][2*(p0-1+i)+1]
    float w1[kInChannels * 3][kConv1Out];        // [(ci*3+k)][co]
    float w2[kConv1Out * 3][kConv2Out];          // [(ci*3+k)][co]
    float b1[kConv1Out];
    float b2[kConv2Out];
};

template <int NJ>
__global__ void __launch_bounds__(ENC_THREADS, 2)
k_encoder_conv(const float* __restrict__ cond, int64_t member_stride, int64_t L, int64_t L1,
               int64_t L2, const float* __restrict__ conv1_w, const float* __restrict__ b1g,
               const float* __restrict__ conv2_w, const float* __restrict__ b2g,
               float* __restrict__ partial, int n_chunks) {
    extern __shared__ __align__(16) unsigned char enc_smem_raw[];
    constexpr int ENC_TP = 32 * NJ;
    EncSmem<NJ>& s = *reinterpret_cast<EncSmem<NJ>*>(enc_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x;
    const int64_t member = blockIdx.y;
    const int64_t p0 = (int64_t)chunk * ENC_TP;
    const float* __restrict__ in = cond + member * member_stride;

    // ---- stage weights and the input window -------------------------------------------
    {
        const float4* src1 = reinterpret_cast<const float4*>(conv1_w);
        float4* dst1 = reinterpret_cast<float4*>(&s.w1[0][0]);
        for (int i = tid; i < kInChannels * 3 * kConv1Out / 4; i += ENC_THREADS) dst1[i] = src1[i];
        const float4* src2 = reinterpret_cast<const float4*>(conv2_w);
        float4* dst2 = reinterpret_cast<float4*>(&s.w2[0][0]);
        for (int i = tid; i < kConv1Out * 3 * kConv2Out / 4; i += ENC_THREADS) dst2[i] = src2[i];
        if (tid < kConv1Out) s.b1[tid] = b1g[tid];
        if (tid < kConv2Out) s.b2[tid] = b2g[tid];
    }
    {
        const int64_t lbase = 4 * (p0 - 1);
        constexpr int NW = 4 * (ENC_TP + 1);
#pragma unroll 2
        for (int ci = 0; ci < kInChannels; ++ci) {
            const float* __restrict__ row = in + (int64_t)ci * L;
            for (int lr = tid; lr < NW; lr += ENC_THREADS) {
                const int64_t l = lbase + lr;
                const float v = (l >= 0 && l < L) ? __ldg(row + l) : 0.f;   // conv1 zero padding
                s.in_ph[lr & 3][ci][lr >> 2] = v;
            }
        }
    }
    __syncthreads();

    // ---- conv1 + ReLU into the two output phases ----------------------------------------
    {
        const int co0 = warp * 4;   // 8 warps x 4 = 32 output channels
        // even phase: h1e[i], i = lane + 32 j ; m = p0 + i ; q = 2m
        {
            float acc[NJ][4];
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[j][c] = s.b1[co0 + c];
#pragma unroll 2
            for (int ci = 0; ci < kInChannels; ++ci) {
                const float4 wa = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 0][co0]);
                const float4 wb = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 1][co0]);
                const float4 wc = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 2][co0]);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int i = lane + 32 * j;
                    const float a = s.in_ph[3][ci][i];       // in[4m-1]
                    const float b = s.in_ph[0][ci][i + 1];   // in[4m]
                    const float c = s.in_ph[1][ci][i + 1];   // in[4m+1]
                    acc[j][0] = fmaf(wa.x, a, acc[j][0]); acc[j][1] = fmaf(wa.y, a, acc[j][1]);
                    acc[j][2] = fmaf(wa.z, a, acc[j][2]); acc[j][3] = fmaf(wa.w, a, acc[j][3]);
                    acc[j][0] = fmaf(wb.x, b, acc[j][0]); acc[j][1] = fmaf(wb.y, b, acc[j][1]);
                    acc[j][2] = fmaf(wb.z, b, acc[j][2]); acc[j][3] = fmaf(wb.w, b, acc[j][3]);
                    acc[j][0] = fmaf(wc.x, c, acc[j][0]); acc[j][1] = fmaf(wc.y, c, acc[j][1]);
                    acc[j][2] = fmaf(wc.z, c, acc[j][2]); acc[j][3] = fmaf(wc.w, c, acc[j][3]);
                }
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int i = lane + 32 * j;
                const int64_t q = 2 * (p0 + i);
                const bool ok = q < L1;                       // conv2 zero padding beyond L1
#pragma unroll
                for (int c = 0; c < 4; ++c) s.h1e[co0 + c][i] = ok ? fmaxf(acc[j][c], 0.f) : 0.f;
            }
        }
        // odd phase: h1o[i], i in [0, ENC_TP) ; m = p0 - 1 + i ; q = 2m + 1
        {
            float acc[NJ][4];
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[j][c] = s.b1[co0 + c];
#pragma unroll 2
            for (int ci = 0; ci < kInChannels; ++ci) {
                const float4 wa = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 0][co0]);
                const float4 wb = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 1][co0]);
                const float4 wc = *reinterpret_cast<const float4*>(&s.w1[ci * 3 + 2][co0]);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int i = lane + 32 * j;
                    const float a = s.in_ph[1][ci][i];   // in[4m+1]
                    const float b = s.in_ph[2][ci][i];   // in[4m+2]
                    const float c = s.in_ph[3][ci][i];   // in[4m+3]
                    acc[j][0] = fmaf(wa.x, a, acc[j][0]); acc[j][1] = fmaf(wa.y, a, acc[j][1]);
                    acc[j][2] = fmaf(wa.z, a, acc[j][2]); acc[j][3] = fmaf(wa.w, a, acc[j][3]);
                    acc[j][0] = fmaf(wb.x, b, acc[j][0]); acc[j][1] = fmaf(wb.y, b, acc[j][1]);
                    acc[j][2] = fmaf(wb.z, b, acc[j][2]); acc[j][3] = fmaf(wb.w, b, acc[j][3]);
                    acc[j][0] = fmaf(wc.x, c, acc[j][0]); acc[j][1] = fmaf(wc.y, c, acc[j][1]);
                    acc[j][2] = fmaf(wc.z, c, acc[j][2]); acc[j][3] = fmaf(wc.w, c, acc[j][3]);
                }
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int i = lane + 32 * j;
                const int64_t q = 2 * (p0 - 1 + i) + 1;
                const bool ok = q >= 0 && q < L1;
#pragma unroll
                for (int c = 0; c < 4; ++c) s.h1o[co0 + c][i] = ok ? fmaxf(acc[j][c], 0.f) : 0.f;
            }
        }
        // the one extra odd entry i = ENC_TP (needed by the chunk's last conv2 position)
        if (tid < kConv1Out) {
            const int i = ENC_TP;
            float a0 = s.b1[tid];
#pragma unroll
            for (int ci = 0; ci < kInChannels; ++ci) {
                a0 = fmaf(s.w1[ci * 3 + 0][tid], s.in_ph[1][ci][i], a0);
                a0 = fmaf(s.w1[ci * 3 + 1][tid], s.in_ph[2][ci][i], a0);
                a0 = fmaf(s.w1[ci * 3 + 2][tid], s.in_ph[3][ci][i], a0);
            }
            const int64_t q = 2 * (p0 - 1 + i) + 1;
            s.h1o[tid][i] = (q < L1) ? fmaxf(a0, 0.f) : 0.f;
        }
    }
    __syncthreads();

    // ---- conv2 + ReLU + pooled partial sum --------------------------------------------
    {
        const int co0 = warp * 8;   // 8 warps x 8 = 64 output channels
        float acc[NJ][8];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = s.b2[co0 + c];
#pragma unroll 2
        for (int ci = 0; ci < kConv1Out; ++ci) {
            float hv[3][NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int i = lane + 32 * j;
                hv[0][j] = s.h1o[ci][i];       // h1[2p-1]
                hv[1][j] = s.h1e[ci][i];       // h1[2p]
                hv[2][j] = s.h1o[ci][i + 1];   // h1[2p+1]
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 wlo = *reinterpret_cast<const float4*>(&s.w2[ci * 3 + k][co0]);
                const float4 whi = *reinterpret_cast<const float4*>(&s.w2[ci * 3 + k][co0 + 4]);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float h = hv[k][j];
                    acc[j][0] = fmaf(wlo.x, h, acc[j][0]); acc[j][1] = fmaf(wlo.y, h, acc[j][1]);
                    acc[j][2] = fmaf(wlo.z, h, acc[j][2]); acc[j][3] = fmaf(wlo.w, h, acc[j][3]);
                    acc[j][4] = fmaf(whi.x, h, acc[j][4]); acc[j][5] = fmaf(whi.y, h, acc[j][5]);
                    acc[j][6] = fmaf(whi.z, h, acc[j][6]); acc[j][7] = fmaf(whi.w, h, acc[j][7]);
                }
            }
        }
        float part[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) part[c] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const bool ok = (p0 + lane + 32 * j) < L2;
#pragma unroll
            for (int c = 0; c < 8; ++c) part[c] += ok ? fmaxf(acc[j][c], 0.f) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part[c] += __shfl_xor_sync(0xffffffffu, part[c], o);
        }
        if (lane == 0) {
            float* dst = partial + ((int64_t)member * n_chunks + chunk) * kConv2Out + co0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = part[c];
        }
    }
}

// pooled mean -> Linear(64,H)+ReLU -> cond_emb ; cond_bias = W0c @ cond_emb + b0.
// grid = n_cond, block = H.
__global__ void k_encoder_finish(const float* __restrict__ partial, int n_chunks, int64_t L2,
                                 const float* __restrict__ w6T, const float* __restrict__ b6,
                                 const float* __restrict__ w0cT, const float* __restrict__ b0,
                                 int H, float* __restrict__ cond_emb,
                                 float* __restrict__ cond_bias) {
    __shared__ float pooled[kConv2Out];
    __shared__ float cemb[512];
    const int tid = threadIdx.x;
    const int64_t member = blockIdx.x;
    for (int co = tid; co < kConv2Out; co += blockDim.x) {
        const float* src = partial + member * n_chunks * kConv2Out + co;
        float sum = 0.f;
        int c = 0;
        for (; c + 8 <= n_chunks; c += 8) {        // loads in flight together, adds in chunk order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = src[(int64_t)(c + u) * kConv2Out];
#pragma unroll
            for (int u = 0; u < 8; ++u) sum += v[u];
        }
        for (; c < n_chunks; ++c) sum += src[(int64_t)c * kConv2Out];
        pooled[co] = sum / (float)L2;
    }
    __syncthreads();
    // both matrix-vector products read their weight column with every load in flight at once (the
    // weights are cold after an L2 flush: a rolled loop would pay one DRAM latency per iteration)
    float w6c[kConv2Out];
#pragma unroll
    for (int k = 0; k < kConv2Out; ++k) w6c[k] = w6T[k * H + tid];
    float a0 = b6[tid], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < kConv2Out; k += 4) {
        a0 = fmaf(w6c[k + 0], pooled[k + 0], a0);
        a1 = fmaf(w6c[k + 1], pooled[k + 1], a1);
        a2 = fmaf(w6c[k + 2], pooled[k + 2], a2);
        a3 = fmaf(w6c[k + 3], pooled[k + 3], a3);
    }
    const float a = fmaxf((a0 + a1) + (a2 + a3), 0.f);
    cemb[tid] = a;
    if (cond_emb) cond_emb[member * H + tid] = a;
    __syncthreads();
    if (cond_bias) {
        float acc0 = b0[tid], acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int k0 = 0; k0 < H; k0 += 32) {          // H is a multiple of 32
            float wc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) wc[k] = w0cT[(int64_t)(k0 + k) * H + tid];
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                acc0 = fmaf(wc[k + 0], cemb[k0 + k + 0], acc0);
                acc1 = fmaf(wc[k + 1], cemb[k0 + k + 1], acc1);
                acc2 = fmaf(wc[k + 2], cemb[k0 + k + 2], acc2);
                acc3 = fmaf(wc[k + 3], cemb[k0 + k + 3], acc3);
            }
        }
        cond_bias[member * H + tid] = (acc0 + acc1) + (acc2 + acc3);
    }
}

}  // namespace ertdiff
