// tcgen05 / TMEM building blocks (sm_100a): shared-memory matrix descriptors, instruction
// descriptor, MMA issue, commit -> mbarrier, TMEM alloc / ld / st.  Raw inline PTX: CUTLASS is
// not used.
//
// Operand layout used throughout: K-major, no swizzle ("interleave").  In units of 16 bytes the
// canonical layout is ((8, n), 2) : ((1, SBO), LBO): a core matrix is 8 rows x 16 bytes stored
// as 128 contiguous bytes; the next 8 rows are SBO bytes further, the next 16-byte K chunk is
// LBO bytes further.  For an operand tile of R rows x K bf16 (K a multiple of 8) we store
//     byte(r, k) = (r / 8) * SBO + (r % 8) * 16 + (k / 8) * LBO + (k % 8) * 2
// with LBO = 128 and SBO = (K / 8) * 128, so one row's K chunks are 128 bytes apart and a thread
// that owns row r writes each chunk with one 16-byte store (8 consecutive rows fill 128
// contiguous bytes: conflict-free).  One tcgen05.mma of kind::f16 consumes K = 16 (two chunks);
// advancing along K adds 2 * LBO to the descriptor's start address.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace ertdiff {
namespace umma {

constexpr uint32_t kLBO = 128;

__host__ __device__ constexpr uint32_t sbo_bytes(int K) { return (uint32_t)(K / 8) * 128u; }
__host__ __device__ constexpr uint32_t tile_bytes(int rows, int K) { return (uint32_t)rows * K * 2u; }
__host__ __device__ inline uint32_t elem_offset(int r, int k, int K) {
    return (uint32_t)(r / 8) * sbo_bytes(K) + (uint32_t)(r % 8) * 16u + (uint32_t)(k / 8) * kLBO +
           (uint32_t)(k % 8) * 2u;
}

// 64-bit shared-memory matrix descriptor (sm_100 format, version 1, no swizzle)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// 32-bit instruction descriptor: kind::f16, A/B = bf16 (K-major), D = f32, M x N
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4)                 // c_format  = F32
         | (1u << 7)                 // a_format  = BF16
         | (1u << 10)                // b_format  = BF16
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier -------------------------------------------------------------------------------
#ifndef MBAR_SUSPEND_HINT_NS
#define MBAR_SUSPEND_HINT_NS 20000u
#endif
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival (release at CTA scope): the caller's earlier shared-memory writes / TMEM reads are
// ordered before the waiter's acquire
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// bounded spin: a kernel bug must not hang the GPU -- after 2^20 polls (each may sleep inside
// try_wait; a healthy MMA completes within a few polls) the wait gives up and the caller flags it
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
        // the suspend-time hint lets the hardware park the warp until the phase completes (or the hint
        // expires) instead of returning early: far fewer polls competing for issue slots
#if MBAR_SUSPEND_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
        if (ok) return true;
    }
    return false;
}

// ---- proxies / tcgen05 fences ---------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() {       // generic-proxy smem writes -> async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------
// one full warp calls alloc/dealloc; ncols is a power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_result_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// one lane of a converged warp (elect.sync): unlike `lane == 0` this keeps the branch warp-uniform for the
// compiler, so loop-invariant MMA descriptors can stay in uniform registers instead of being moved there
// (R2UR) before every tcgen05.mma
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// the same with the accumulate flag fixed at compile time (no per-issue predicate set-up)
__device__ __forceinline__ void mma_bf16_first(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_bf16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}

// warp w (w = warp index % 4) reads TMEM lanes 32w..32w+31; thread l gets lane 32w+l,
// 32 consecutive 32-bit columns starting at the column encoded in taddr
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// same mapping, 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// same mapping, 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);     // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// 16-byte shared store of 8 packed bf16
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

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

// ---------------------------------------------------------------------------------------------
// Self-test GEMM: D (128 x N) = A (128 x K) * B (N x K)^T with bf16-rounded operands and fp32
// accumulation, one CTA of 128 threads.  Exercises every convention above (operand layout,
// descriptors, instruction descriptor, commit/mbarrier, TMEM lane/column mapping); the tests
// compare it with a bf16 matmul.  status[0] = 1 if the mbarrier wait timed out.
template <int N, int K>
__global__ void __launch_bounds__(128) k_umma_selftest(const float* __restrict__ A,
                                                       const float* __restrict__ B,
                                                       float* __restrict__ D, int* status) {
    extern __shared__ __align__(128) unsigned char umma_smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sA = smem_u32(umma_smem);
    const uint32_t sB = sA + tile_bytes(128, K);
    constexpr uint32_t NCOLS = N < 32 ? 32 : N;

    for (int c = 0; c < K / 8; ++c) {           // thread = row of A
        const float* src = A + (size_t)tid * K + 8 * c;
        sts_u4(sA + elem_offset(tid, 8 * c, K), pack_bf16(src[0], src[1]), pack_bf16(src[2], src[3]),
               pack_bf16(src[4], src[5]), pack_bf16(src[6], src[7]));
    }
    for (int r = tid; r < N; r += 128)
        for (int c = 0; c < K / 8; ++c) {
            const float* src = B + (size_t)r * K + 8 * c;
            sts_u4(sB + elem_offset(r, 8 * c, K), pack_bf16(src[0], src[1]), pack_bf16(src[2], src[3]),
                   pack_bf16(src[4], src[5]), pack_bf16(src[6], src[7]));
        }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), NCOLS);
    if (tid == 0) { mbar_init(smem_u32(&mbar), 1); fence_mbar_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        constexpr uint32_t idesc = idesc_bf16_f32(128, N);
#pragma unroll
        for (int k = 0; k < K / 16; ++k)
            mma_bf16(tmem, smem_desc(sA + 2 * k * kLBO, kLBO, sbo_bytes(K)),
                     smem_desc(sB + 2 * k * kLBO, kLBO, sbo_bytes(K)), idesc, k > 0);
        mma_commit(smem_u32(&mbar));
    }
    const bool ok = mbar_wait(smem_u32(&mbar), 0);
    tc_fence_after();
    if (!ok && tid == 0) status[0] = 1;
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, NCOLS);
}

}  // namespace umma
}  // namespace ertdiff
