// Ensemble statistics over members (axis 0) of a row-major (N, Q) array:
// np.mean/std/var (ECD.py:867-869), np.percentile linear (ECD.py:870-872, 612, 1126-1127,
// 1199-1200) and the Gaussian-KDE mode (ECD.py:747-762).
//
// Bit-exactness rules: numpy adds rows one after another into the output
// for an axis-0 reduction, so each column is a left-to-right sum in the array's dtype; `_lerp`
// uses separate multiplies and adds.  All arithmetic that must match numpy goes through the
// _rn intrinsics so that nvcc never contracts it into FMAs.
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace ertdiff {

template <typename T> struct RN;
template <> struct RN<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
};
template <> struct RN<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    static __device__ __forceinline__ double nan() { return CUDART_NAN; }
};

// ------------------------------------------------------------------------------------------
// Moments: one thread per column, members visited in order (the order numpy uses).  Adjacent
// threads read adjacent columns, so every row access is coalesced; loads are issued 8 rows
// ahead of the dependent add chain.  HBM-bound: N*Q*sizeof(T) bytes, read twice when std/var
// are requested (the second pass normally hits L2).
template <typename T>
__global__ void __launch_bounds__(128)
k_moments(const T* __restrict__ a, int64_t N, int64_t Q, T* __restrict__ mean,
          T* __restrict__ stdv, T* __restrict__ var) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Q) return;
    const T* __restrict__ col = a + j;
    T acc = col[0];
    int64_t i = 1;
    for (; i + 8 <= N; i += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[(i + u) * Q];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = RN<T>::add(acc, v[u]);
    }
    for (; i < N; ++i) acc = RN<T>::add(acc, col[i * Q]);
    const T m = RN<T>::div(acc, (T)N);
    if (mean) mean[j] = m;
    if (!stdv && !var) return;
    T d0 = RN<T>::sub(col[0], m);
    T s2 = RN<T>::mul(d0, d0);
    i = 1;
    for (; i + 8 <= N; i += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = col[(i + u) * Q];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const T d = RN<T>::sub(v[u], m);
            s2 = RN<T>::add(s2, RN<T>::mul(d, d));
        }
    }
    for (; i < N; ++i) {
        const T d = RN<T>::sub(col[i * Q], m);
        s2 = RN<T>::add(s2, RN<T>::mul(d, d));
    }
    const T vv = RN<T>::div(s2, (T)N);
    if (var) var[j] = vv;
    if (stdv) stdv[j] = RN<T>::sqrt(vv);
}

// ------------------------------------------------------------------------------------------
// Few columns, many members (the chain's own output: Q = 29 parameters, N up to 10^5 members): one
// thread per column leaves the machine to 29 threads waiting on strided loads.  Here one CTA owns
// one column: all its threads gather the column into shared memory (every load in flight at once),
// then a single thread runs numpy's left-to-right chain out of shared memory with 128-bit reads, so
// the dependent fp add -- 4 cycles per member -- is all that is left.  The second pass (variance)
// re-reads shared memory, not HBM.  MODE 0: numpy mean / var / std (bit-exact order).  MODE 1: the
// KDE column constants (double sums, any order: all threads).
constexpr int SQ_MAXQ = 256;
constexpr int SQ_SMEM_BYTES = 192 * 1024;     // column chunk held in shared memory

struct KdeColumn;
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_colstats_smallq(const T* __restrict__ a, int64_t N, int64_t Q, T* __restrict__ mean, T* __restrict__ stdv,
                  T* __restrict__ var, double scott_factor_sq, double* __restrict__ kde_cols /* (Q,2) */) {
    extern __shared__ __align__(16) unsigned char sq_smem_raw[];
    T* xs = reinterpret_cast<T*>(sq_smem_raw);
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int64_t j = blockIdx.x;
    constexpr int64_t CH = SQ_SMEM_BYTES / (int64_t)sizeof(T);
    const bool second = MODE == 1 || stdv != nullptr || var != nullptr;
    const bool resident = N <= CH;               // the whole column stays in shared memory for pass 2
    T acc = (T)0, m = (T)0;
    double dacc = 0.0, dmean = 0.0;
    for (int pass = 0; pass < (second ? 2 : 1); ++pass) {
        for (int64_t c0 = 0; c0 < N; c0 += CH) {
            const int n = (int)(N - c0 < CH ? N - c0 : CH);
            if (pass == 0 || !resident) {
                __syncthreads();
                for (int i = tid; i < n; i += 256) xs[i] = a[(c0 + i) * Q + j];
                __syncthreads();
            }
            if (MODE == 0) {
                if (tid == 0) {
                    int i = 0;
                    if (c0 == 0) {
                        if (pass == 0) acc = xs[0];
                        else { const T d0 = RN<T>::sub(xs[0], m); acc = RN<T>::mul(d0, d0); }
                        i = 1;
                    }
                    for (; i < n && (i & 3); ++i) {           // up to a 16-byte boundary (fp32) / 32-byte (fp64)
                        if (pass == 0) acc = RN<T>::add(acc, xs[i]);
                        else { const T d = RN<T>::sub(xs[i], m); acc = RN<T>::add(acc, RN<T>::mul(d, d)); }
                    }
#pragma unroll 4
                    for (; i + 4 <= n; i += 4) {
                        T v[4];
                        if (sizeof(T) == 4) *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(xs + i);
                        else { *reinterpret_cast<double2*>(v) = *reinterpret_cast<const double2*>(xs + i);
                               *reinterpret_cast<double2*>(v + 2) = *reinterpret_cast<const double2*>(xs + i + 2); }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (pass == 0) acc = RN<T>::add(acc, v[u]);
                            else { const T d = RN<T>::sub(v[u], m); acc = RN<T>::add(acc, RN<T>::mul(d, d)); }
                        }
                    }
                    for (; i < n; ++i) {
                        if (pass == 0) acc = RN<T>::add(acc, xs[i]);
                        else { const T d = RN<T>::sub(xs[i], m); acc = RN<T>::add(acc, RN<T>::mul(d, d)); }
                    }
                }
            } else {
                for (int i = tid; i < n; i += 256) {
                    const double d = (double)xs[i] - dmean;      // dmean == 0 in pass 0
                    dacc += pass == 0 ? d : d * d;
                }
            }
        }
        if (MODE == 0) {
            if (tid == 0) {
                if (pass == 0) {
                    m = RN<T>::div(acc, (T)N);
                    if (mean) mean[j] = m;
                } else {
                    const T vv = RN<T>::div(acc, (T)N);
                    if (var) var[j] = vv;
                    if (stdv) stdv[j] = RN<T>::sqrt(vv);
                }
            }
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = dacc;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += red[w];
            if (pass == 0) { dmean = tot / (double)N; dacc = 0.0; }
            else if (tid == 0) {
                const double h2 = (tot / (double)(N - 1)) * scott_factor_sq;
                kde_cols[2 * j] = dmean;
                kde_cols[2 * j + 1] = -0.5 / h2;        // -inf when the column is constant
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Coverage of K nested probability intervals (ECD.py:1121-1132, 1195-1206): for interval k and
// column c (= condition m, parameter c % P), indicator = (low[k][c] < truth[c]) & (truth[c] <= upp[k][c]).
// counts[k][0] = number of ones over all columns, counts[k][1 + j] = over the columns of parameter j.
// grid = K, block = 256.  Integer counts: the means the reference takes are exact quotients of them.
__global__ void __launch_bounds__(256)
k_interval_coverage(const double* __restrict__ low, const double* __restrict__ upp,
                    const double* __restrict__ truth, int64_t Q, int P, int32_t* __restrict__ counts) {
    __shared__ int cnt[kPPad + 1];
    const int tid = threadIdx.x;
    const int64_t k = blockIdx.x;
    if (tid <= kPPad) cnt[tid] = 0;
    __syncthreads();
    const double* lo = low + k * Q;
    const double* up = upp + k * Q;
    for (int64_t c = tid; c < Q; c += 256) {
        const double t = truth[c];
        if (lo[c] < t && t <= up[c]) {
            atomicAdd(&cnt[0], 1);
            atomicAdd(&cnt[1 + (int)(c % P)], 1);
        }
    }
    __syncthreads();
    if (tid <= P) counts[k * (P + 1) + tid] = cnt[tid];
}

// ------------------------------------------------------------------------------------------
// Global min / max (ECD.py:749-750).  NaN propagates like np.min/np.max.
template <typename T>
__global__ void k_minmax_partial(const T* __restrict__ a, int64_t n, double* __restrict__ part) {
    __shared__ double smin[32], smax[32];
    __shared__ int snan[32];
    double lo = CUDART_INF, hi = -CUDART_INF;
    int has_nan = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)a[i];
        has_nan |= (v != v);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = lo; smax[warp] = hi; snan[warp] = has_nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; ++w) {
            lo = fmin(lo, smin[w]); hi = fmax(hi, smax[w]); has_nan |= snan[w];
        }
        part[3 * blockIdx.x + 0] = lo;
        part[3 * blockIdx.x + 1] = hi;
        part[3 * blockIdx.x + 2] = has_nan ? 1.0 : 0.0;
    }
}
__global__ void k_minmax_final(const double* __restrict__ part, int nblocks,
                               double* __restrict__ out2) {
    // one warp: lanes stride over the per-block partials, then a shuffle tree (min / max / any-NaN commute)
    const int lane = threadIdx.x;
    double lo = CUDART_INF, hi = -CUDART_INF, nn = 0.0;
    for (int b = lane; b < nblocks; b += 32) {
        lo = fmin(lo, part[3 * b]); hi = fmax(hi, part[3 * b + 1]); nn += part[3 * b + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) {
        out2[0] = nn > 0 ? CUDART_NAN : lo;
        out2[1] = nn > 0 ? CUDART_NAN : hi;
    }
}

// ------------------------------------------------------------------------------------------
// Percentiles.  A CTA owns CT adjacent columns: rows are read coalesced (CT*sizeof(T) bytes
// per row), transposed into shared memory, each column is sorted there with a bitonic
// network (padding = +inf), and every requested percentile is interpolated with numpy's
// `_lerp`.  Index arithmetic (lo, hi, gamma) is prepared on the host in the dtype numpy uses.
struct PctlQuery {
    int32_t lo, hi;
    float gamma_f;    // index dtype f32
    double gamma_d;   // index dtype f64
};

constexpr int kMaxPctlQueries = 96;     // by-value kernel argument: 96 * 24 B
struct PctlQueryPack {
    int32_t n;
    PctlQuery q[kMaxPctlQueries];
};

template <typename T, typename G, typename O>
__device__ __forceinline__ O lerp_numpy(T A, T Bv, G gamma) {
    const T d = RN<T>::sub(Bv, A);            // subtract(b, a) in the array's dtype
    const O dO = (O)d;
    const O g = (O)gamma;
    O r = RN<O>::add((O)A, RN<O>::mul(dO, g));
    if (gamma >= (G)0.5) r = RN<O>::sub((O)Bv, RN<O>::mul(dO, RN<O>::sub((O)1, g)));
    return r;
}

template <typename T, typename G, typename O>
__global__ void k_percentiles(const T* __restrict__ a, int64_t N, int64_t Q, int NP /*pow2>=N*/,
                              int CT, const __grid_constant__ PctlQueryPack qs,
                              O* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char pct_smem_raw[];
    T* sm = reinterpret_cast<T*>(pct_smem_raw);               // [CT][NP]
    int* nanflag = reinterpret_cast<int*>(sm + (size_t)CT * NP);  // [CT]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    for (int c = tid; c < CT; c += nthr) nanflag[c] = 0;
    __syncthreads();
    // transposed load, +inf padding
    const int64_t total = (int64_t)NP * CT;
    for (int64_t idx = tid; idx < total; idx += nthr) {
        const int c = (int)(idx % CT);
        const int64_t i = idx / CT;
        T v = RN<T>::inf();
        if (i < N && c0 + c < Q) {
            v = a[i * Q + c0 + c];
            if (v != v) { nanflag[c] = 1; v = RN<T>::inf(); }
        }
        sm[(size_t)c * NP + i] = v;
    }
    __syncthreads();
    // bitonic sort of each column (ascending)
    const int64_t pairs = total >> 1;
    const int half_np = NP >> 1;
    for (int k = 2; k <= NP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int64_t tpair = tid; tpair < pairs; tpair += nthr) {
                const int c = (int)(tpair / half_np);
                const int r = (int)(tpair % half_np);
                const int i = ((r & ~(j - 1)) << 1) | (r & (j - 1));   // bit log2(j) cleared
                const int ip = i | j;
                T* colp = sm + (size_t)c * NP;
                const T x = colp[i], y = colp[ip];
                const bool asc = (i & k) == 0;
                if ((x > y) == asc) { colp[i] = y; colp[ip] = x; }
            }
            __syncthreads();
        }
    }
    const int nq = qs.n;
    for (int idx = tid; idx < nq * CT; idx += nthr) {
        const int c = idx % CT, k = idx / CT;
        if (c0 + c >= Q) continue;
        const T* colp = sm + (size_t)c * NP;
        const PctlQuery qq = qs.q[k];
        O r;
        if (nanflag[c]) {
            r = RN<O>::nan();
        } else if (sizeof(G) == 4) {
            r = lerp_numpy<T, float, O>(colp[qq.lo], colp[qq.hi], qq.gamma_f);
        } else {
            r = lerp_numpy<T, double, O>(colp[qq.lo], colp[qq.hi], qq.gamma_d);
        }
        out[(int64_t)k * Q + c0 + c] = r;
    }
}

// ------------------------------------------------------------------------------------------
// Percentiles of short columns (N <= 1024, e.g. the reference's 50 realisations of a 4693 x 14 map,
// ECD.py:870-872): one WARP sorts one column in registers.  The CTA (8 warps) first loads a tile of
// CT = 8 * CPW adjacent columns coalesced (CT * sizeof(T) contiguous bytes per member) into shared memory; each
// warp then takes CPW columns in turn: lane l holds elements l, l+32, ..., l+32(E-1) (padding = +inf), the bitonic
// network runs on registers -- compare-exchanges between registers for strides >= 32, xor-shuffles below -- with
// no block-wide barrier, and lane k interpolates query k (numpy `_lerp`), fetching its two order statistics with
// indexed shuffles.  Same results as k_percentiles, bit for bit.
template <typename T>
__device__ __forceinline__ void cmpx(T& a, T& b, bool asc) {      // ascending: a <= b afterwards
    const T x = a, y = b;
    if ((x > y) == asc) { a = y; b = x; }
}

template <typename T, typename G, typename O, int E>
__global__ void __launch_bounds__(256)
k_percentiles_warp(const T* __restrict__ a, int64_t N, int64_t Q, int CPW,
                   const __grid_constant__ PctlQueryPack qs, O* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char pct_smem_raw[];
    T* tile = reinterpret_cast<T*>(pct_smem_raw);                 // [N][CT + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CT = 8 * CPW, LD = CT + 1;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    const int n = (int)N;
    for (int idx = tid; idx < n * CT; idx += 256) {
        const int c = idx % CT, i = idx / CT;
        tile[i * LD + c] = (c0 + c < Q) ? a[(int64_t)i * Q + c0 + c] : (T)0;
    }
    __syncthreads();
    constexpr int LOG = (E == 1 ? 0 : E == 2 ? 1 : E == 4 ? 2 : E == 8 ? 3 : E == 16 ? 4 : 5) + 5;   // log2(32 E)
    for (int cw = 0; cw < CPW; ++cw) {
        const int c = warp * CPW + cw;
        if (c0 + c >= Q) break;                                   // (warp-uniform)
        T v[E];
        bool has_nan = false;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int i = lane + 32 * r;
            T x = i < n ? tile[i * LD + c] : RN<T>::inf();
            if (x != x) { has_nan = true; x = RN<T>::inf(); }
            v[r] = x;
        }
        has_nan = __any_sync(0xffffffffu, has_nan);
#pragma unroll
        for (int lk = 1; lk <= LOG; ++lk) {
            const int k = 1 << lk;
#pragma unroll
            for (int lj = lk - 1; lj >= 0; --lj) {
                const int j = 1 << lj;
                if (j >= 32) {                                    // partner element lives in another register of this lane
                    const int rj = j >> 5;
#pragma unroll
                    for (int r = 0; r < E; ++r)
                        if ((r & rj) == 0) cmpx(v[r], v[r | rj], (((lane + 32 * r) & k) == 0));
                } else {                                          // partner element lives in lane ^ j, same register
                    const bool lower = (lane & j) == 0;
#pragma unroll
                    for (int r = 0; r < E; ++r) {
                        const T other = __shfl_xor_sync(0xffffffffu, v[r], j);
                        const bool asc = ((lane + 32 * r) & k) == 0;
                        // the lower element of an ascending pair keeps the smaller value (ties: either copy)
                        const bool take_other = lower == asc ? (v[r] > other) : (other > v[r]);
                        if (take_other) v[r] = other;
                    }
                }
            }
        }
        // element e of the sorted column sits in lane e % 32, register e / 32; lane k serves query k
        for (int q0 = 0; q0 < qs.n; q0 += 32) {
            const int k = q0 + lane;
            const PctlQuery qq = qs.q[k < qs.n ? k : 0];
            T A = v[0], Bv = v[0];
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const T fa = __shfl_sync(0xffffffffu, v[r], qq.lo & 31);
                const T fb = __shfl_sync(0xffffffffu, v[r], qq.hi & 31);
                if ((qq.lo >> 5) == r) A = fa;
                if ((qq.hi >> 5) == r) Bv = fb;
            }
            if (k < qs.n) {
                O res;
                if (has_nan) res = RN<O>::nan();
                else if (sizeof(G) == 4) res = lerp_numpy<T, float, O>(A, Bv, qq.gamma_f);
                else res = lerp_numpy<T, double, O>(A, Bv, qq.gamma_d);
                out[(int64_t)k * Q + c0 + c] = res;
            }
        }
    }
}

// order-preserving unsigned keys of floating-point values (radix selection, sorted runs)
template <typename T> struct SortKey;
template <> struct SortKey<float> {
    using K = uint32_t;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ K pad() { return 0xFFFFFFFFu; }
    static __device__ __forceinline__ K of(float v) {
        const uint32_t u = __float_as_uint(v);
        return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    }
    static __device__ __forceinline__ float back(K k) {
        return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
    }
};
template <> struct SortKey<double> {
    using K = unsigned long long;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ K pad() { return 0xFFFFFFFFFFFFFFFFull; }
    static __device__ __forceinline__ K of(double v) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        return u ^ ((u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double back(K k) {
        return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull)));
    }
};

// ------------------------------------------------------------------------------------------
// Percentiles of medium-length columns of many-column arrays with a few quantiles (an ensemble of 1024 .. 8192
// simulated maps, 25/50/75: SURVEY.md §8 d) WITHOUT sorting: exact selection by radix.  A CTA loads CT adjacent
// columns (coalesced) into shared memory as order-preserving unsigned keys and treats them one at a time:
//   * one 256-bin histogram of the first byte below the bits all keys share (block-wide min / max) -- shared by all
//     the column's ranks;
//   * per rank: the bucket holding it is compacted into a small buffer (or, when it is too large for the buffer,
//     narrowed in place by one more histogram round over the column with a prefix test) and the procedure repeats on
//     the next byte until at most 256 candidates are left, where every thread ranks one candidate by counting.
// Work per column ~ N (histogram) + ranks * N (one compaction pass each) + small rounds, against N log^2 N / 2
// compare-exchanges for the bitonic sort: 8192 members x 65,702 pixels, float64: 43 ms -> a few ms.
// The selected order statistics are the same keys a sort would deliver, so the results are bit-identical.
constexpr int PS_CAP = 512;         // candidates a compaction buffer holds
constexpr int PS_FINAL = 64;        // candidates left when the last step ranks them by counting

template <typename K>
struct SelectSmem {
    int hist[256];
    int wsum[8];
    int bucket, before, count, counter, have2;
    K result, result2, kmin[8], kmax[8];
};

// one shared-memory atomic per distinct bin and warp instead of one per lane: a digit of floating-point keys is
// often shared by most of a column (a binade holds a quarter of a lognormal sample), and 32 lanes hitting one
// address serialise
__device__ __forceinline__ void hist_add_aggregated(int* hist, int bin, bool active) {
    // the lanes that share the first active lane's bin add once, together; the others add for themselves
    // (a full match.any costs more than the conflicts it removes when most lanes hold different bins)
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (act == 0) return;
    const int leader = __ffs(act) - 1;
    const int b0 = __shfl_sync(0xffffffffu, bin, leader);
    const unsigned same = __ballot_sync(0xffffffffu, active && bin == b0);
    if (active) {
        if (bin != b0) atomicAdd(&hist[bin], 1);
        else if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[b0], __popc(same));
    }
}
// append `key` of the lanes with `take` to dst, one atomic per warp
template <typename K>
__device__ __forceinline__ void append_aggregated(K* dst, int* counter, K key, bool take) {
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (take) dst[base + __popc(m & ((1u << lane) - 1))] = key;
}

// exact k-th smallest (0-based) of keys[0..n): every thread of the 256-thread CTA calls; returns the key in all.
// With `want_next` the (k+1)-th smallest is delivered in sm.result2 as well whenever it sits among the same final
// candidates (sm.have2 = 1): the two order statistics numpy interpolates between are neighbours.
// State: bits [shift, BITS) of the answer are decided and held in `prefix`; a key is a candidate when it agrees with
// the prefix on those bits.  `cur` holds exactly the candidates (physical) or a superset that is prefix-tested.
template <typename K, int BITS>
__device__ K block_select(const K* __restrict__ keys, int n, int k, int shift0, const int* __restrict__ cum0 /* [257] */,
                          K* bufA, K* bufB, SelectSmem<K>& sm, bool want_next = false) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // round 0 comes from the column's shared histogram of byte [shift0, shift0 + 8): thread t tests bucket t
    if (cum0[tid] <= k && k < cum0[tid + 1]) { sm.bucket = tid; sm.before = cum0[tid]; sm.count = cum0[tid + 1] - cum0[tid]; }
    __syncthreads();
    int kk = k - sm.before, c = sm.count, shift = shift0;
    const K high = (shift0 + 8 >= BITS) ? (K)0 : (K)((keys[0] >> (shift0 + 8)) << (shift0 + 8));   // the bits all keys share
    K prefix = high | ((K)sm.bucket << shift0);
    const K* cur = keys;
    int n_cur = n;
    bool physical = false;
    __syncthreads();
    for (;;) {
        if (!physical && c <= PS_CAP) {         // the bucket fits a buffer: compact it
            K* dst = (cur == bufA) ? bufB : bufA;
            if (tid == 0) sm.counter = 0;
            __syncthreads();
            for (int i0 = 0; i0 < n_cur; i0 += 256) {           // (uniform trip count: warp-wide votes inside)
                const int i = i0 + tid;
                const K key = i < n_cur ? cur[i] : (K)0;
                append_aggregated(dst, &sm.counter, key, i < n_cur && ((key ^ prefix) >> shift) == 0);
            }
            __syncthreads();
            cur = dst; n_cur = c; physical = true;
        }
        if (physical && n_cur <= PS_FINAL) break;
        if (shift == 0) {                       // every bit decided: all candidates equal the prefix
            if (tid == 0) { sm.have2 = (want_next && kk + 1 < c) ? 1 : 0; sm.result2 = prefix; }
            __syncthreads();
            return prefix;
        }
        const int above = shift;                // candidates agree with the prefix on bits [above, BITS)
        shift -= 8;
        // ---- histogram of the next byte among the candidates ----------------------------------------------
        sm.hist[tid] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n_cur; i0 += 256) {
            const int i = i0 + tid;
            const K key = i < n_cur ? cur[i] : (K)0;
            hist_add_aggregated(sm.hist, (int)((key >> shift) & 255), i < n_cur && (physical || ((key ^ prefix) >> above) == 0));
        }
        __syncthreads();
        const int v = sm.hist[tid];             // inclusive scan of the 256 bins: thread t owns bin t
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) sm.wsum[warp] = incl;
        __syncthreads();
        int off = 0;
        for (int w = 0; w < warp; ++w) off += sm.wsum[w];
        incl += off;
        if (incl - v <= kk && kk < incl) { sm.bucket = tid; sm.before = incl - v; sm.count = v; }
        __syncthreads();
        prefix |= (K)sm.bucket << shift;
        kk -= sm.before; c = sm.count;
        physical = false;                       // the next round's candidates are a subset of cur
        __syncthreads();
    }
    // ---- at most PS_FINAL candidates: thread t ranks candidate t by counting (two warps at most) -------------------
    if (tid < n_cur) {
        const K mine = cur[tid];
        int less = 0;
        for (int j = 0; j < n_cur; ++j) {
            const K o = cur[j];
            less += (o < mine || (o == mine && j < tid)) ? 1 : 0;
        }
        if (less == kk) sm.result = mine;
        if (want_next && less == kk + 1) sm.result2 = mine;
    }
    if (tid == 0) sm.have2 = (want_next && kk + 1 < n_cur) ? 1 : 0;
    __syncthreads();
    const K r = sm.result;
    __syncthreads();
    return r;
}

// grid = column groups of CT; 256 threads; dynamic shared memory: CT*N keys + 2 buffers of PS_CAP keys + 257 ints
template <typename T, typename G, typename O>
__global__ void __launch_bounds__(256)
k_percentiles_select(const T* __restrict__ a, int64_t N, int64_t Q, int CT,
                     const __grid_constant__ PctlQueryPack qs, O* __restrict__ out) {
    using K = typename SortKey<T>::K;
    constexpr int BITS = SortKey<T>::BITS;
    extern __shared__ __align__(16) unsigned char pct_smem_raw[];
    const int n = (int)N;
    K* keys = reinterpret_cast<K*>(pct_smem_raw);                 // [CT][N]
    K* bufA = keys + (size_t)CT * n;
    K* bufB = bufA + PS_CAP;
    int* cum0 = reinterpret_cast<int*>(bufB + PS_CAP);            // [257]
    int* nanflag = cum0 + 257;                                    // [CT]
    __shared__ SelectSmem<K> sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * CT;
    for (int c = tid; c < CT; c += 256) nanflag[c] = 0;
    __syncthreads();
    for (int idx = tid; idx < n * CT; idx += 256) {
        const int c = idx % CT, i = idx / CT;
        T v = (c0 + c < Q) ? a[(int64_t)i * Q + c0 + c] : (T)0;
        if (v != v) { nanflag[c] = 1; v = RN<T>::inf(); }
        keys[(size_t)c * n + i] = SortKey<T>::of(v);
    }
    __syncthreads();
    for (int c = 0; c < CT; ++c) {
        if (c0 + c >= Q) break;
        const K* col = keys + (size_t)c * n;
        if (nanflag[c]) {
            for (int k = tid; k < qs.n; k += 256) out[(int64_t)k * Q + c0 + c] = RN<O>::nan();
            continue;
        }
        // ---- bits shared by all keys, and the histogram of the first byte below them ------------------------
        K kmn = ~(K)0, kmx = 0;
        for (int i = tid; i < n; i += 256) { const K key = col[i]; kmn = key < kmn ? key : kmn; kmx = key > kmx ? key : kmx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const K a1 = __shfl_xor_sync(0xffffffffu, kmn, o), a2 = __shfl_xor_sync(0xffffffffu, kmx, o);
            kmn = a1 < kmn ? a1 : kmn; kmx = a2 > kmx ? a2 : kmx;
        }
        if (lane == 0) { sm.kmin[warp] = kmn; sm.kmax[warp] = kmx; }
        sm.hist[tid] = 0;
        __syncthreads();
        for (int w = 0; w < 8; ++w) { kmn = sm.kmin[w] < kmn ? sm.kmin[w] : kmn; kmx = sm.kmax[w] > kmx ? sm.kmax[w] : kmx; }
        const K diff = kmn ^ kmx;
        int top = 0;                                              // highest differing bit (0 when all keys are equal)
        if (diff) top = BITS - 1 - (BITS == 64 ? __clzll((long long)diff) : __clz((int)diff));
        const int shift0 = (top / 8) * 8;
        for (int i0 = 0; i0 < n; i0 += 256) {
            const int i = i0 + tid;
            hist_add_aggregated(sm.hist, i < n ? (int)((col[i] >> shift0) & 255) : 0, i < n);
        }
        __syncthreads();
        {   // exclusive prefix of the 256 bins -> cum0[0..256]
            const int v = sm.hist[tid];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            int off = 0;
            for (int w = 0; w < warp; ++w) off += sm.wsum[w];
            cum0[tid + 1] = incl + off;
            if (tid == 0) cum0[0] = 0;
        }
        __syncthreads();
        for (int k = 0; k < qs.n; ++k) {
            const PctlQuery qq = qs.q[k];
            const K ka = block_select<K, BITS>(col, n, qq.lo, shift0, cum0, bufA, bufB, sm, qq.hi != qq.lo);
            K kb = ka;
            if (qq.hi != qq.lo) {               // usually the neighbour came with it; otherwise it heads the next bucket
                const bool have = sm.have2 != 0;
                const K k2 = sm.result2;
                __syncthreads();
                kb = have ? k2 : block_select<K, BITS>(col, n, qq.hi, shift0, cum0, bufA, bufB, sm);
            }
            if (tid == 0) {
                const T A = SortKey<T>::back(ka), Bv = SortKey<T>::back(kb);
                O r;
                if (sizeof(G) == 4) r = lerp_numpy<T, float, O>(A, Bv, qq.gamma_f);
                else r = lerp_numpy<T, double, O>(A, Bv, qq.gamma_d);
                out[(int64_t)k * Q + c0 + c] = r;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Percentiles of columns that do not fit one CTA's shared memory (any N): two kernels.
//   k_sort_runs     the column is cut into runs of CH members; a CTA sorts one run of CT adjacent columns
//                   in shared memory (same bitonic network) and writes it, as order-preserving unsigned
//                   keys, to a column-major scratch array runs[col][N] (run r = [r*CH, min(N, (r+1)*CH)))
//   k_select_runs   one warp per (column, query): the exact order statistic s[lo] is found by deciding the
//                   key bit by bit -- "how many members are < try?" is a sum of lower bounds over the
//                   sorted runs, each lane binary-searching its runs inside a window that shrinks with every
//                   decided bit -- then s[hi] is either the same value (ties) or the smallest successor over
//                   the runs, and numpy's `_lerp` finishes as in k_percentiles.  No full merge is needed:
//                   a query costs O(R * (bits + log CH)) L2 reads.
// Results are bit-identical to k_percentiles (and numpy) for any N.
// grid = (runs, column groups of CT); columns [col0, col0 + ncols) of `a`; CH = run length (power of two)
template <typename T>
__global__ void __launch_bounds__(1024)
k_sort_runs(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0, int ncols, int CH, int CT,
            typename SortKey<T>::K* __restrict__ runs, int* __restrict__ nanflag /* [ncols] */) {
    using K = typename SortKey<T>::K;
    extern __shared__ __align__(16) unsigned char pct_smem_raw[];
    K* sm = reinterpret_cast<K*>(pct_smem_raw);               // [CT][CH]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * CH;              // first member of this run
    const int cbase = blockIdx.y * CT;
    const int total = CH * CT;
    for (int idx = tid; idx < total; idx += nthr) {
        const int c = idx % CT, i = idx / CT;
        K k = SortKey<T>::pad();
        if (r0 + i < N && cbase + c < ncols) {
            T v = a[(r0 + i) * Q + col0 + cbase + c];
            if (v != v) { nanflag[cbase + c] = 1; v = RN<T>::inf(); }
            k = SortKey<T>::of(v);
        }
        sm[c * CH + i] = k;
    }
    __syncthreads();
    const int pairs = total >> 1, half = CH >> 1;
    for (int k = 2; k <= CH; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int tp = tid; tp < pairs; tp += nthr) {
                const int c = tp / half, r = tp % half;
                const int i = ((r & ~(j - 1)) << 1) | (r & (j - 1));
                const int ip = i | j;
                K* colp = sm + c * CH;
                const K x = colp[i], y = colp[ip];
                const bool asc = (i & k) == 0;
                if ((x > y) == asc) { colp[i] = y; colp[ip] = x; }
            }
            __syncthreads();
        }
    }
    const int len = (int)((N - r0) < CH ? (N - r0) : CH);
    for (int idx = tid; idx < len * CT; idx += nthr) {
        const int c = idx / len, i = idx % len;
        if (cbase + c < ncols) runs[(int64_t)(cbase + c) * N + r0 + i] = sm[c * CH + i];
    }
}

// first index in [lo, hi) of the sorted run whose key is >= v (hi if none)
template <typename K>
__device__ __forceinline__ int run_lower_bound(const K* __restrict__ run, int lo, int hi, K v) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(run + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// grid.x covers (ncols * nq) warps, 4 warps per CTA; dynamic shared memory: 4 warps x 3 x R ints (the windows)
template <typename T, typename G, typename O>
__global__ void __launch_bounds__(128)
k_select_runs(const typename SortKey<T>::K* __restrict__ runs, int64_t N, int64_t Q, int64_t col0, int ncols,
              int CH, int R, const int* __restrict__ nanflag, const __grid_constant__ PctlQueryPack qs,
              O* __restrict__ out) {
    using K = typename SortKey<T>::K;
    extern __shared__ int sel_win[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * 4 + warp;
    if (w >= (int64_t)ncols * qs.n) return;
    const int c = (int)(w % ncols), kq = (int)(w / ncols);
    const PctlQuery qq = qs.q[kq];
    O* dst = out + (int64_t)kq * Q + col0 + c;
    if (nanflag[c]) { if (lane == 0) *dst = RN<O>::nan(); return; }
    const K* col = runs + (int64_t)c * N;
    int* wa = sel_win + warp * 3 * R;       // window [wa[i], wb[i]] brackets lower_bound(run i, answer)
    int* wb = wa + R;
    int* wp = wb + R;                       // lower bound of the current trial key
    for (int i = lane; i < R; i += 32) {
        const int64_t r0 = (int64_t)i * CH;
        wa[i] = 0; wb[i] = (int)((N - r0) < CH ? (N - r0) : CH);
    }
    __syncwarp();
    const int64_t rank = qq.lo;
    K ans = 0;
    for (int bit = SortKey<T>::BITS - 1; bit >= 0; --bit) {
        const K trial = ans | ((K)1 << bit);
        int64_t below = 0;
        for (int i = lane; i < R; i += 32) {
            const int pos = run_lower_bound(col + (int64_t)i * CH, wa[i], wb[i], trial);
            below += pos;
            wp[i] = pos;                     // (entry i is only ever touched by this lane)
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        const bool take = below <= rank;     // the rank-th member is >= trial
        for (int i = lane; i < R; i += 32) {
            if (take) wa[i] = wp[i]; else wb[i] = wp[i];
        }
        if (take) ans = trial;
    }
    const T A = SortKey<T>::back(ans);
    T Bv = A;
    if (qq.hi != qq.lo) {
        // members <= ans: upper bound = lower bound of the next key
        int64_t le = 0;
        K succ = SortKey<T>::pad();
        for (int i = lane; i < R; i += 32) {
            const int64_t r0 = (int64_t)i * CH;
            const int len = (int)((N - r0) < CH ? (N - r0) : CH);
            const K* run = col + r0;
            const int ub = (ans == SortKey<T>::pad()) ? len : run_lower_bound(run, wa[i], len, (K)(ans + 1));
            le += ub;
            if (ub < len) { const K nx = __ldg(run + ub); succ = nx < succ ? nx : succ; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            le += __shfl_xor_sync(0xffffffffu, le, o);
            const K other = __shfl_xor_sync(0xffffffffu, succ, o);
            succ = other < succ ? other : succ;
        }
        if (le <= rank + 1) Bv = SortKey<T>::back(succ);     // no tie reaches rank + 1: the successor
    }
    if (lane == 0) {
        if (sizeof(G) == 4) *dst = lerp_numpy<T, float, O>(A, Bv, qq.gamma_f);
        else *dst = lerp_numpy<T, double, O>(A, Bv, qq.gamma_d);
    }
}

// ------------------------------------------------------------------------------------------
// Gaussian-KDE mode, ECD.py:751-762: per column, argmax over a common grid of
//   pdf(g) = sum_i exp(-(g - x_i)^2 / (2 h^2)),   h^2 = var_ddof1 * N^(-2/5)   (Scott),
// first maximum.  (The normalisation constant is common to a column's grid points and is
// skipped.)  grid[i] = lo + i*step with separately rounded multiply and add, grid[G-1] = hi,
// exactly as np.linspace builds it.
//
// The argmax is decided in float64, but float64 exp is only evaluated where it can matter:
//   k_kde_scan32   every (column, grid point) in fp32 on data centred at the column mean
//                  (ex2.approx, 64-term fp32 partial sums added up in fp64): N*G cheap terms.
//   k_kde_select64 per column: the fp32 maximum M; every grid point whose fp32 value reaches
//                  M*(1 - KDE_TOL) is re-evaluated in float64 (members summed lane-strided in
//                  order, then a fixed shuffle tree) and the first float64 maximum wins.
// KDE_TOL = 1e-3 is ~50x the worst-case relative error of the fp32 scan (centred arguments:
// <= 2e-5), so the true float64 maximum -- and every exact tie with it -- is always among the
// candidates; typically a few dozen of the 5000 grid points are re-evaluated.
// Columns with zero variance have no KDE (scipy raises): mode = NaN, index = -1.
constexpr float KDE_TOL = 1e-3f;
constexpr int KDE_MAX_CAND = 2048;

struct KdeColumn {          // per-column constants, written by k_kde_prepare
    double mean;
    double neg_inv_2h2;     // -1 / (2 h^2)
};

// one warp per column: mean, ddof-1 variance, bandwidth
template <typename T>
__global__ void k_kde_prepare(const T* __restrict__ a, int64_t N, int64_t Q,
                              double scott_factor_sq, KdeColumn* __restrict__ cols) {
    const int lane = threadIdx.x & 31;
    const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= Q) return;
    double s = 0.0;
    for (int64_t i = lane; i < N; i += 32) s += (double)a[i * Q + col];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double mean = s / (double)N;
    double ss = 0.0;
    for (int64_t i = lane; i < N; i += 32) {
        const double d = (double)a[i * Q + col] - mean;
        ss += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
        const double h2 = (ss / (double)(N - 1)) * scott_factor_sq;
        cols[col].mean = mean;
        cols[col].neg_inv_2h2 = -0.5 / h2;        // -inf when the column is constant
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ double kde_grid_point(int g, int G, double lo, double hi, double step) {
    return (g == G - 1) ? hi : __dadd_rn(__dmul_rn((double)g, step), lo);
}

// Shared-memory layout of a column's centred members for the fp32 scan: 4 words of padding after every 64 members, so
// that lanes working on neighbouring 64-member blocks (the member-split scan) hit different banks with their 128-bit
// loads.  kde_pad(i) = word index of member i; kde_padded(n) = words for n members.
__host__ __device__ __forceinline__ int64_t kde_pad(int64_t i) { return i + ((i >> 6) << 2); }
__host__ __device__ __forceinline__ int64_t kde_padded(int64_t n) { return n + (((n + 63) >> 6) << 2); }

// How far from the column's members the fp32 scan has to look (in data units).
//  * Always: a term ex2.approx.ftz(d^2 * c2) is EXACTLY +0 once d^2/(2h^2) * log2(e) > 126, i.e. |d| > 13.22 h, so a
//    grid point further than 14 h from every member has a scan value of exactly zero -- skipping it changes
//    nothing, whatever the grid spacing.  This is what keeps a column much narrower than the common grid's step
//    (ECD.py:749-751 spans the GLOBAL min..max) from costing a full 5000-point scan.
//  * If the grid is not coarser than 7 bandwidths and a member lies inside it, the maximum is >= 1e-3 (some
//    grid point is within step/2 of a member) while points further than R h, R^2 = 2 ln(1e7 N), sum to < 1e-7:
//    they cannot hold the maximum and are written as zero.
__device__ __forceinline__ double kde_scan_reach(double h, double step, int64_t N, bool member_inside_grid) {
    double reach = 14.0 * h;
    if (member_inside_grid && step <= 7.0 * h) reach = fmin(reach, h * sqrt(2.0 * log(1e7 * (double)N)));
    return reach;
}

// fp32 partial sums of one block of <= 64 centred members (a padded block starts 16-byte aligned: full blocks are
// read with 128-bit loads) for one / two grid points; the terms are added in member order.  A term is
// ex2(-(v*sc - x*sc)^2) with sc = sqrt(log2(e) / (2 h^2)): one FFMA, one FMUL, the MUFU and the accumulating FADD --
// four issue slots per term (va, vb arrive pre-multiplied by sc; nsc = -sc)
__device__ __forceinline__ float kde_term(float vs, float x, float nsc) {
    const float d = fmaf(x, nsc, vs);
    return ex2_approx(-d * d);
}
__device__ __forceinline__ void kde_block_sum2(const float* __restrict__ blk, int len, float va, float vb, float nsc,
                                               float& pa, float& pb) {
    pa = 0.f; pb = 0.f;
    if (len == 64) {
        const float4* __restrict__ b4 = reinterpret_cast<const float4*>(blk);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const float4 x4 = b4[q];
            pa += kde_term(va, x4.x, nsc); pb += kde_term(vb, x4.x, nsc);
            pa += kde_term(va, x4.y, nsc); pb += kde_term(vb, x4.y, nsc);
            pa += kde_term(va, x4.z, nsc); pb += kde_term(vb, x4.z, nsc);
            pa += kde_term(va, x4.w, nsc); pb += kde_term(vb, x4.w, nsc);
        }
    } else {
        for (int i = 0; i < len; ++i) {
            const float xi = blk[i];
            pa += kde_term(va, xi, nsc);
            pb += kde_term(vb, xi, nsc);
        }
    }
}
__device__ __forceinline__ float kde_block_sum1(const float* __restrict__ blk, int len, float va, float nsc) {
    float pa = 0.f;
    if (len == 64) {
        const float4* __restrict__ b4 = reinterpret_cast<const float4*>(blk);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const float4 x4 = b4[q];
            pa += kde_term(va, x4.x, nsc);
            pa += kde_term(va, x4.y, nsc);
            pa += kde_term(va, x4.z, nsc);
            pa += kde_term(va, x4.w, nsc);
        }
    } else {
        for (int i = 0; i < len; ++i) pa += kde_term(va, blk[i], nsc);
    }
    return pa;
}

// ---- coarse-to-fine scan ---------------------------------------------------------------------------------
// The scan only has to deliver, exactly, the grid points whose value reaches (1 - KDE_TOL) of the column's maximum
// (k_kde_select64 re-evaluates those in float64); everything else may be written as zero.  The sum of kernels
// S(x) = sum_i 2^-((x - x_i) sc)^2 has a bounded second derivative, |S''(x)| <= 2 ln 2 N sc^2 (each term's is
// largest at its centre), so between two evaluated points c < c' = c + s step it cannot rise above their larger
// value by more than (c' - c)^2 / 8 times that bound (the error of linear interpolation).  Where the grid is fine
// against the bandwidth a CTA therefore first evaluates every s-th point of its chunk ("coarse" points, plus the
// chunk's end), takes M = the largest value seen so far for the column -- its own coarse maximum and, when several
// CTAs share the column, a global cell they all atomicMax into: whatever is in the cell is an evaluated value, i.e.
// a valid lower bound of the maximum, no matter how far the other CTAs have got (the cell is (launch epoch << 32 |
// float bits): a value left by an earlier launch is ignored, nothing needs resetting) -- and skips every interval with
//       max(S(c), S(c')) + 0.17329 N (sc s step)^2  <  M (1 - 2 KDE_TOL)
// because no point inside it can then reach the candidates' threshold (the factor 2 covers the fp32 scan's own
// error, <= 2e-5 relative).  An evaluated point is the same sum of the same 64-term fp32 blocks as in a full scan
// (how many lanes share it only regroups the float64 additions of the block sums); which of the NON-candidate
// points are evaluated may vary from run to run (the cell), the candidates -- and the mode -- never do.
// s = the largest power of two <= max_stride with 3 s step <= h: the bound then costs <= 1.4 % of N in height; a
// unimodal column of the chain's output (step / h ~ 0.01 .. 0.02) evaluates 1/16 .. 1/32 of its active range plus the
// band around the mode -- about a tenth of the points of the full scan (numerical check of the rule on normal,
// bimodal, log-normal, uniform and tied samples is part of the CPU test suite).
constexpr int KDE_MAX_COARSE = 2560;       // coarse values of one CTA's grid chunk, in shared memory

__device__ __forceinline__ int kde_coarse_stride(double h, double step, int max_stride) {
    int s = 1;
    while (2 * s <= max_stride && 6.0 * (double)s * step <= h) s *= 2;
    return s;
}

// lanes per grid point for a pass over `npoints` points: all of a CTA's threads on the points there are (a coarse
// pass of a dozen points, or the few fine points around the mode, would otherwise leave most threads idle while a
// few walk all N members), never fewer than the launch's own `ms`, never more than there are 64-member blocks
__device__ __forceinline__ int kde_lanes_for(int npoints, int nthr, int ms_launch, int64_t N) {
    int ms = ms_launch;
    while (ms < 32 && (int64_t)npoints * (2 * ms) <= nthr && (int64_t)64 * (2 * ms) <= N + 63) ms *= 2;
    return ms;
}

// This CTA's (`part` of `nparts`) share of the scan of one column's active grid range [ga, gb].
// accumulate(va, vb, has0, has1, sa, sb, sub, ms): member sums of up to two grid points (scaled coordinates va / vb)
// into sa / sb, by lane `sub` of the `ms` lanes that share the points (lane `sub` takes the 64-member blocks sub,
// sub + ms, ...; the partial sums meet in a fixed xor-shuffle tree).  cta_uniform: accumulate contains block barriers
// (the tiled form) and must be called by every thread.
// Without a coarse pass (s = 1) a part takes a contiguous chunk of the range.  With one, the unit of work is the
// interval between two coarse points, dealt to the parts ROUND-ROBIN: the intervals that need their interior evaluated
// are neighbours on the grid (the band around the mode), and contiguous chunks would leave them all to one or two
// CTAs.  A part evaluates both ends of each of its intervals (the right end is also the next part's left end: the
// coarse pass is done twice, it is a sixteenth or less of the points); alone on the column it shares the ends.
template <typename Acc>
__device__ __forceinline__ void kde_scan_points(Acc&& accumulate, bool cta_uniform, int ga, int gb, int part, int nparts, int s,
                                                int64_t N, const KdeColumn kc, double lo, double hi, double step, int G,
                                                float sc, int ms_launch, unsigned long long* cell, unsigned int epoch,
                                                float* __restrict__ out) {
    __shared__ float cv[KDE_MAX_COARSE];       // coarse values of this part's intervals
    __shared__ int ivl[KDE_MAX_COARSE];        // (local) intervals that may hold a candidate
    __shared__ float s_red[8];
    __shared__ float s_thr;
    __shared__ int s_nact;
    (void)cta_uniform;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    auto coord = [&](int g) { return (float)(kde_grid_point(g, G, lo, hi, step) - kc.mean) * sc; };
    auto reduce_lanes = [&](double& sa, double& sb, int ms) {
        for (int o = ms >> 1; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
    };
    // intervals of the active range: k = 0 .. n_iv - 1, interval k = [ga + k s, min(ga + (k + 1) s, gb)]
    const int n_iv = (s > 1 && ga < gb) ? (gb - ga + s - 1) / s : 0;
    const int n_own = n_iv > part ? (n_iv - part + nparts - 1) / nparts : 0;          // this part's: k = part + i nparts
    const bool shared_ends = nparts == 1;
    const int n_pts = shared_ends ? n_own + 1 : 2 * n_own;                            // coarse points this part evaluates
    // (the same decision in every CTA of the column: taken on the largest share, not on this part's)
    const bool coarse = n_iv > 0 && (shared_ends ? n_iv + 1 : 2 * ((n_iv + nparts - 1) / nparts)) <= KDE_MAX_COARSE;
    if (!coarse) {
        // ---- every point of a contiguous chunk ----------------------------------------------------------------
        const int n_act = gb - ga + 1;
        const int chunk = (n_act + nparts - 1) / nparts;
        const int g_begin = ga + part * chunk;
        const int g_end = min(gb + 1, g_begin + chunk);
        const int ms = kde_lanes_for(g_end - g_begin, nthr, ms_launch, N);
        const int sub = tid & (ms - 1), slot = tid / ms, slots = nthr / ms;
        for (int gbase = g_begin; gbase < g_end; gbase += 2 * slots) {        // (uniform trip count: shuffles inside)
            const int g0 = gbase + slot, g1 = g0 + slots;
            const bool has0 = g0 < g_end, has1 = g1 < g_end;
            const float va = has0 ? coord(g0) : 0.f, vb = has1 ? coord(g1) : 0.f;
            double sa = 0.0, sb = 0.0;
            accumulate(va, vb, has0, has1, sa, sb, sub, ms);
            reduce_lanes(sa, sb, ms);
            if (sub == 0) {
                if (has1) out[g1] = (float)sb;
                if (has0) out[g0] = (float)sa;
            }
        }
        return;
    }
    // coarse point j of this part -> grid point; who writes it; the ends of local interval i
    auto point_g = [&](int j) {
        const int k = shared_ends ? j : part + (j >> 1) * nparts + (j & 1);           // index of the coarse point in the range
        return min(ga + k * s, gb);
    };
    auto point_written_here = [&](int j, int g) { return shared_ends || !(j & 1) || g == gb; };
    auto left_of = [&](int i) { return shared_ends ? i : 2 * i; };
    // ---- coarse pass ------------------------------------------------------------------------------------------
    {
        const int ms = kde_lanes_for(n_pts, nthr, ms_launch, N);
        const int sub = tid & (ms - 1), slot = tid / ms, slots = nthr / ms;
        for (int jb = 0; jb < n_pts; jb += 2 * slots) {                  // (uniform trip count)
            const int j0 = jb + slot, j1 = j0 + slots;
            const bool has0 = j0 < n_pts, has1 = j1 < n_pts;
            const int g0 = has0 ? point_g(j0) : 0, g1 = has1 ? point_g(j1) : 0;
            const float va = has0 ? coord(g0) : 0.f, vb = has1 ? coord(g1) : 0.f;
            double sa = 0.0, sb = 0.0;
            accumulate(va, vb, has0, has1, sa, sb, sub, ms);
            reduce_lanes(sa, sb, ms);
            if (sub == 0) {
                if (has0) { cv[j0] = (float)sa; if (point_written_here(j0, g0)) out[g0] = (float)sa; }
                if (has1) { cv[j1] = (float)sb; if (point_written_here(j1, g1)) out[g1] = (float)sb; }
            }
        }
    }
    if (tid == 0) s_nact = 0;
    __syncthreads();
    float m = 0.f;
    for (int j = tid; j < n_pts; j += nthr) m = fmaxf(m, cv[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w) m = fmaxf(m, s_red[w]);
        if (cell) {                           // sums are >= 0: their bit patterns order as integers
            const unsigned long long old = atomicMax(cell, ((unsigned long long)epoch << 32) | __float_as_uint(m));
            if ((unsigned int)(old >> 32) == epoch) m = fmaxf(m, __uint_as_float((unsigned int)old));
        }
        const double w = (double)sc * (double)s * step;
        const float add = (float)(0.17329 * (double)N * w * w) * 1.001f;          // (s step)^2 / 8 * max |S''|
        s_thr = m * (1.0f - 2.0f * KDE_TOL) - add;
    }
    __syncthreads();
    const float thr = s_thr;
    // ---- the intervals that may hold a candidate are listed, the others' interior points written as zero --------
    for (int i = tid; i < n_own; i += nthr) {
        const int k = part + i * nparts;
        const int c0 = ga + k * s, c1 = min(c0 + s, gb);
        const int l = left_of(i);
        if (fmaxf(cv[l], cv[l + 1]) < thr) {                                      // (a NaN never compares below)
            for (int g = c0 + 1; g < c1; ++g) out[g] = 0.f;
        } else if (c1 - c0 > 1) {
            ivl[atomicAdd(&s_nact, 1)] = i;
        }
    }
    __syncthreads();
    // ---- fine pass over the listed intervals' interior points, densely mapped onto the threads -----------------
    const int n_fine = s_nact * (s - 1);             // (the range's last interval may be shorter: its surplus slots idle)
    if (n_fine == 0) return;
    const int ms = kde_lanes_for(n_fine, nthr, ms_launch, N);
    const int sub = tid & (ms - 1), slot = tid / ms, slots = nthr / ms;
    auto fine_g = [&](int p, bool& has) {
        const int k = part + ivl[p / (s - 1)] * nparts;
        const int g = ga + k * s + 1 + p % (s - 1);
        has = g < min(ga + (k + 1) * s, gb);
        return g;
    };
    for (int pb = 0; pb < n_fine; pb += 2 * slots) {                     // (uniform trip count)
        const int p0 = pb + slot, p1 = p0 + slots;
        bool has0 = p0 < n_fine, has1 = p1 < n_fine;
        const int g0 = has0 ? fine_g(p0, has0) : 0, g1 = has1 ? fine_g(p1, has1) : 0;
        const float va = has0 ? coord(g0) : 0.f, vb = has1 ? coord(g1) : 0.f;
        double sa = 0.0, sb = 0.0;
        accumulate(va, vb, has0, has1, sa, sb, sub, ms);
        reduce_lanes(sa, sb, ms);
        if (sub == 0) {
            if (has1) out[g1] = (float)sb;
            if (has0) out[g0] = (float)sa;
        }
    }
}

// The fp32 scan of one column, split over `nparts` CTAs (this one is `part`).  `xs` = the column's N
// members centred at kc.mean (fp32, shared memory).
//
// Grid points further than `kde_scan_reach` from every member are written as zero and only the "active" range
// [x_min - reach, x_max + reach] is evaluated, split evenly over the parts: a column that occupies a
// small part of the common grid (ECD.py:749-751 spans the GLOBAL min..max) costs proportionally less.
// Inside the active range every member is summed for every evaluated point, so the values equal the full scan's bit
// for bit; with max_stride > 1 the range is scanned coarse to fine (above), `cell` = the column's shared maximum
// (nullptr when this CTA scans the column alone), `epoch` = this launch's tag.
__device__ __forceinline__ void kde_scan_column(const float* __restrict__ xs, int64_t N, const KdeColumn kc,
                                                double lo, double hi, int G, int part, int nparts,
                                                float* __restrict__ out, int ms = 1, int max_stride = 1,
                                                unsigned long long* cell = nullptr, unsigned int epoch = 0) {
    __shared__ float s_mn[8], s_mx[8];
    __shared__ int s_in[8];
    __shared__ int s_range[3];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const double step = (hi - lo) / (double)(G - 1);
    // ---- the column's extent and whether a member lies inside the grid ------------------------------
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int inside = 0;
    const float glo = (float)(lo - kc.mean), ghi = (float)(hi - kc.mean);
    for (int64_t i = tid; i < N; i += nthr) {
        const float v = xs[kde_pad(i)];
        mn = fminf(mn, v); mx = fmaxf(mx, v);
        inside |= (v >= glo && v <= ghi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        inside |= __shfl_xor_sync(0xffffffffu, inside, o);
    }
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_in[warp] = inside; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); inside |= s_in[w]; }
        int ga = 0, gb = G - 1, stride = 1;
        const double h = sqrt(-0.5 / kc.neg_inv_2h2);               // 0 for a constant column, NaN for NaN data
        if (h > 0.0 && step > 0.0) {
            const double reach = kde_scan_reach(h, step, N, inside != 0);
            const double a = ((kc.mean + (double)mn - reach) - lo) / step;
            const double b = ((kc.mean + (double)mx + reach) - lo) / step;
            if (a > 1.0) ga = (int)fmin(a - 1.0, (double)(G - 1));  // one grid point of slack on both sides
            if (b < (double)(G - 2)) gb = (int)fmax(b + 1.0, 0.0);
            if (gb < ga) { ga = 0; gb = G - 1; }
            stride = kde_coarse_stride(h, step, max_stride);
        }
        s_range[0] = ga; s_range[1] = gb; s_range[2] = stride;
    }
    __syncthreads();
    const int ga = s_range[0], gb = s_range[1], stride = s_range[2];
    // ---- zeros outside the active range (each part clears its static slice of the row) --------------
    {
        const int z0 = (int)((int64_t)G * part / nparts), z1 = (int)((int64_t)G * (part + 1) / nparts);
        for (int g = z0 + tid; g < z1; g += nthr)
            if (g < ga || g > gb) out[g] = 0.f;
    }
    // ---- this part's share of the active range ---------------------------------------------------------
    // exponent in base 2: -(g - x)^2 log2(e) / (2 h^2) = -((g - x) sc)^2
    const float sc = (float)sqrt(-kc.neg_inv_2h2 * 1.4426950408889634), nsc = -sc;
    auto accumulate = [&](float va, float vb, bool has0, bool has1, double& sa, double& sb, int sub, int ms) {
        if (has0 && has1) {                        // two grid points per thread: one shared-memory read feeds both
            for (int64_t i0 = 64 * sub; i0 < N; i0 += 64 * ms) {
                float pa, pb;
                kde_block_sum2(xs + kde_pad(i0), (int)((i0 + 64 <= N) ? 64 : N - i0), va, vb, nsc, pa, pb);
                sa += (double)pa;
                sb += (double)pb;
            }
        } else if (has0 || has1) {                 // a single point: no wasted second evaluation
            const float v = has0 ? va : vb;
            double t = 0.0;
            for (int64_t i0 = 64 * sub; i0 < N; i0 += 64 * ms)
                t += (double)kde_block_sum1(xs + kde_pad(i0), (int)((i0 + 64 <= N) ? 64 : N - i0), v, nsc);
            if (has0) sa = t; else sb = t;
        }
    };
    kde_scan_points(accumulate, false, ga, gb, part, nparts, stride, N, kc, lo, hi, step, G, sc, ms, cell, epoch, out);
}

// grid = (columns of this batch, n_gchunks); CTA (c, gc) scans grid points
// [gc*gchunk, (gc+1)*gchunk) of column col0 + c and writes s32[c*G + g].
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_scan32(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0,
             const double* __restrict__ lohi, int G, int ms /* lanes per grid point, a power of two <= 32 */,
             const KdeColumn* __restrict__ cols, float* __restrict__ s32,
             int max_stride /* coarse-to-fine: largest coarse stride, 1 = scan every point */,
             unsigned long long* __restrict__ cells /* per column of the launch: shared maximum (gridDim.y > 1) */,
             unsigned int epoch /* tag of this launch in the cells */) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    float* xs = reinterpret_cast<float*>(kde_smem_raw);        // [N] centred members, fp32
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t col = col0 + blockIdx.x;
    const KdeColumn kc = cols[col];
    for (int64_t i = tid; i < N; i += nthr) xs[kde_pad(i)] = (float)((double)a[i * Q + col] - kc.mean);
    __syncthreads();
    kde_scan_column(xs, N, kc, lohi[0], lohi[1], G, (int)blockIdx.y, (int)gridDim.y, s32 + (int64_t)blockIdx.x * G, ms,
                    max_stride, gridDim.y > 1 ? cells + blockIdx.x : nullptr, epoch);
}

// All-grid fallback of the float64 decision: exp(-d^2/(2h^2)) is exactly 0 in float64 once d^2/(2h^2) > 745.2
// (d > 38.6 h), so a grid point further than 40 h from every member has a KDE sum of exactly zero and only the
// grid range [g_lo, g_hi] around the column's extent [xmin, xmax] needs evaluating.  If nothing in it is positive
// every grid point is zero and the first maximum is index 0 (what an argmax over the full grid returns).
__device__ __forceinline__ void kde_nonzero_range(double xmin, double xmax, const KdeColumn kc, double lo, double step,
                                                  int G, int& g_lo, int& g_hi) {
    g_lo = 0; g_hi = G - 1;
    const double h = sqrt(-0.5 / kc.neg_inv_2h2);
    if (h > 0.0 && step > 0.0 && xmin <= xmax) {
        const double a = ((xmin - 40.0 * h) - lo) / step, b = ((xmax + 40.0 * h) - lo) / step;
        if (a > 1.0) g_lo = (int)fmin(a - 1.0, (double)(G - 1));
        if (b < (double)(G - 2)) g_hi = (int)fmax(b + 1.0, 0.0);
    }
}

// min / max of a value over a 256-thread CTA (all threads call; result in every thread)
__device__ __forceinline__ void block_minmax(double& mn, double& mx, double* red16 /* shared, 16 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { red16[warp] = mn; red16[8 + warp] = mx; }
    __syncthreads();
    mn = red16[0]; mx = red16[8];
    for (int w = 1; w < nwarps; ++w) { mn = fmin(mn, red16[w]); mx = fmax(mx, red16[8 + w]); }
}

// The float64 decision for one column (see above): `row` = the column's fp32 scan, `xs` = its N members
// as float64 in shared memory.  Called by every thread of a 256-thread CTA.
// With nparts > 1 several CTAs share a column: part p re-evaluates the candidates whose grid index is
// congruent to p (the candidates of a mode are neighbours on the grid), and the last part to finish (atomic ticket, `slot` = the column's index in
// the launch) combines the partial maxima (first index wins ties, as in the single-CTA form).
__device__ __forceinline__ void kde_select_column(int64_t N, int64_t col, double lo, double hi, int G,
                                                  const KdeColumn kc, const float* row,
                                                  const double* __restrict__ xs, double* __restrict__ mode_out,
                                                  int64_t* __restrict__ index_out, int part = 0, int nparts = 1,
                                                  int64_t slot = 0, double* partials = nullptr,
                                                  unsigned int* tickets = nullptr) {
    __shared__ float redf[8];
    __shared__ double redv[8];
    __shared__ int redi[8];
    __shared__ int cand[KDE_MAX_CAND];
    __shared__ int ncand;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = nthr >> 5;
    if (tid == 0) ncand = 0;
    // fp32 maximum of the column's scan
    float mx = 0.f;
    // (__ldcg: in the fused kernel other CTAs of this launch wrote the scan -- read it through L2)
    for (int g = tid; g < G; g += nthr) mx = fmaxf(mx, __ldcg(row + g));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) redf[warp] = mx;
    __syncthreads();
    mx = redf[0];
    for (int w = 1; w < nwarps; ++w) mx = fmaxf(mx, redf[w]);
    const bool degenerate = !(kc.neg_inv_2h2 > -CUDART_INF) || !(kc.neg_inv_2h2 == kc.neg_inv_2h2);
    const float thr = mx * (1.0f - KDE_TOL);
    __shared__ int ntotal;
    if (tid == 0) ntotal = 0;
    __syncthreads();
    for (int g = tid; g < G; g += nthr) {
        if (__ldcg(row + g) >= thr) {
            atomicAdd(&ntotal, 1);
            if (g % nparts == part) {
                const int k = atomicAdd(&ncand, 1);
                if (k < KDE_MAX_CAND) cand[k] = g;
            }
        }
    }
    __syncthreads();
    // a flat scan (more candidates than the list holds, or an all-zero scan) falls back to
    // evaluating every grid point in float64 (the decision is the same in every part)
    const bool all = ntotal > KDE_MAX_CAND || !(mx > 0.f);
    const double step = (hi - lo) / (double)(G - 1);
    int g_first = part, n_eval = ncand;
    if (all) {                                   // (uniform over the CTA)
        __shared__ double red16[16];
        double xmn = CUDART_INF, xmx = -CUDART_INF;
        for (int64_t i = tid; i < N; i += nthr) { xmn = fmin(xmn, xs[i]); xmx = fmax(xmx, xs[i]); }
        block_minmax(xmn, xmx, red16);
        int g_lo, g_hi;
        kde_nonzero_range(xmn, xmx, kc, lo, step, G, g_lo, g_hi);
        g_first = g_lo + ((part - g_lo) % nparts + nparts) % nparts;      // first g >= g_lo with g % nparts == part
        n_eval = (degenerate || g_first > g_hi) ? 0 : (g_hi - g_first) / nparts + 1;
    }
    // long columns: every candidate is summed by the whole CTA (thread t takes members t, t + 256, ...; a fixed shuffle
    // tree, then the warps' partial sums in warp order), so that a part with two or three candidates uses all its warps
    double best = -1.0;
    int besti = 0x7fffffff;
    if (N >= 4096) {
        for (int k = 0; k < n_eval; ++k) {
            const int g = all ? g_first + k * nparts : cand[k];
            const double gv = kde_grid_point(g, G, lo, hi, step);
            double acc = 0.0;
            for (int64_t i = tid; i < N; i += nthr) {
                const double d = gv - xs[i];
                acc += exp(d * d * kc.neg_inv_2h2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            __syncthreads();
            if (lane == 0) redv[warp] = acc;
            __syncthreads();
            acc = redv[0];
            for (int w = 1; w < nwarps; ++w) acc += redv[w];
            if (acc > best || (acc == best && g < besti)) { best = acc; besti = g; }      // (identical in every thread)
        }
    } else {
        // short columns: one warp per candidate (eight candidates in flight hide the float64 exp latency that a
        // single member per thread would expose)
        for (int k = warp; k < n_eval; k += nwarps) {
            const int g = all ? g_first + k * nparts : cand[k];
            const double gv = kde_grid_point(g, G, lo, hi, step);
            double acc = 0.0;
            for (int64_t i = lane; i < N; i += 32) {
                const double d = gv - xs[i];
                acc += exp(d * d * kc.neg_inv_2h2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (acc > best || (acc == best && g < besti)) { best = acc; besti = g; }
        }
        __syncthreads();
        if (lane == 0) { redv[warp] = best; redi[warp] = besti; }
        __syncthreads();
        for (int w = 0; w < nwarps; ++w)
            if (redv[w] > best || (redv[w] == best && redi[w] < besti)) { best = redv[w]; besti = redi[w]; }
    }
    if (tid == 0) {
        if (nparts > 1) {
            double* mine = partials + (slot * nparts + part) * 2;
            mine[0] = best; mine[1] = (double)besti;
            __threadfence();
            const unsigned int t = atomicAdd(&tickets[slot], 1u);
            if (t != (unsigned int)(nparts - 1)) return;        // not the last part of this column
            tickets[slot] = 0u;                                   // self-cleaning
            __threadfence();
            best = -1.0; besti = 0x7fffffff;
            for (int p = 0; p < nparts; ++p) {
                const double v = __ldcg(partials + (slot * nparts + p) * 2);
                const int gi = (int)__ldcg(partials + (slot * nparts + p) * 2 + 1);
                if (v > best || (v == best && gi < besti)) { best = v; besti = gi; }
            }
        }
        if (!(best > 0.0)) besti = 0;            // every grid point is zero: the first one is the maximum
        if (degenerate) {
            if (index_out) index_out[col] = -1;
            if (mode_out) mode_out[col] = CUDART_NAN;
        } else {
            if (index_out) index_out[col] = besti;
            if (mode_out) mode_out[col] = kde_grid_point(besti, G, lo, hi, step);
        }
    }
}


// grid = (columns of this batch, parts), 256 threads.
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_select64(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0,
               const double* __restrict__ lohi, int G, const KdeColumn* __restrict__ cols,
               const float* __restrict__ s32, double* __restrict__ mode_out,
               int64_t* __restrict__ index_out, double* __restrict__ partials,
               unsigned int* __restrict__ tickets) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    double* xs = reinterpret_cast<double*>(kde_smem_raw);      // [N] members, float64
    const int64_t col = col0 + blockIdx.x;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) xs[i] = (double)a[i * Q + col];
    kde_select_column(N, col, lohi[0], lohi[1], G, cols[col], s32 + (int64_t)blockIdx.x * G, xs, mode_out, index_out,
                      (int)blockIdx.y, (int)gridDim.y, (int64_t)blockIdx.x, partials, tickets);
}

// ------------------------------------------------------------------------------------------
// Columns longer than one CTA's shared memory (N > 25,600 members): the same two steps with the members
// streamed through shared memory a tile at a time.  Per grid point the scan adds the same 64-term fp32 partial
// sums in the same order as k_kde_scan32 would (tiles are multiples of 64 members), so the two forms agree bit
// for bit; the float64 selection keeps one accumulator per candidate across the tiles.
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_scan32_tiled(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0,
                   const double* __restrict__ lohi, int G, const KdeColumn* __restrict__ cols,
                   float* __restrict__ s32, int tile /* members per tile, multiple of 64 */, int ms /* lanes per grid point */,
                   int max_stride, unsigned long long* __restrict__ cells, unsigned int epoch /* as in k_kde_scan32 */) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    float* xs = reinterpret_cast<float*>(kde_smem_raw);        // [tile] centred members, fp32
    __shared__ float s_mn[8], s_mx[8];
    __shared__ int s_in[8];
    __shared__ int s_range[3];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int64_t col = col0 + blockIdx.x;
    const int part = blockIdx.y, nparts = gridDim.y;
    const KdeColumn kc = cols[col];
    const double lo = lohi[0], hi = lohi[1];
    const double step = (hi - lo) / (double)(G - 1);
    float* out = s32 + (int64_t)blockIdx.x * G;
    // ---- the column's extent (one pass over the column, no staging) -------------------------------------
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int inside = 0;
    const float glo = (float)(lo - kc.mean), ghi = (float)(hi - kc.mean);
    for (int64_t i = tid; i < N; i += nthr) {
        const float v = (float)((double)a[i * Q + col] - kc.mean);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
        inside |= (v >= glo && v <= ghi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        inside |= __shfl_xor_sync(0xffffffffu, inside, o);
    }
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_in[warp] = inside; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); inside |= s_in[w]; }
        int ga = 0, gb = G - 1, stride = 1;
        const double h = sqrt(-0.5 / kc.neg_inv_2h2);
        if (h > 0.0 && step > 0.0) {                                    // see kde_scan_column
            stride = kde_coarse_stride(h, step, max_stride);
            const double reach = kde_scan_reach(h, step, N, inside != 0);
            const double aa = ((kc.mean + (double)mn - reach) - lo) / step;
            const double bb = ((kc.mean + (double)mx + reach) - lo) / step;
            if (aa > 1.0) ga = (int)fmin(aa - 1.0, (double)(G - 1));
            if (bb < (double)(G - 2)) gb = (int)fmax(bb + 1.0, 0.0);
            if (gb < ga) { ga = 0; gb = G - 1; }
        }
        s_range[0] = ga; s_range[1] = gb; s_range[2] = stride;
    }
    __syncthreads();
    const int ga = s_range[0], gb = s_range[1], stride = s_range[2];
    {
        const int z0 = (int)((int64_t)G * part / nparts), z1 = (int)((int64_t)G * (part + 1) / nparts);
        for (int g = z0 + tid; g < z1; g += nthr)
            if (g < ga || g > gb) out[g] = 0.f;
    }
    // exponent in base 2: -(g - x)^2 log2(e) / (2 h^2) = -((g - x) sc)^2
    const float sc = (float)sqrt(-kc.neg_inv_2h2 * 1.4426950408889634), nsc = -sc;
    // the members stream through shared memory once per pass of up to two grid points per slot: block barriers inside,
    // so every thread calls it
    auto accumulate = [&](float va, float vb, bool has0, bool has1, double& sa, double& sb, int sub, int ms) {
        const float v1 = has0 ? va : vb;
        double t = 0.0;
        for (int64_t t0 = 0; t0 < N; t0 += tile) {
            const int n = (int)(N - t0 < tile ? N - t0 : tile);
            __syncthreads();
            for (int i = tid; i < n; i += nthr) xs[kde_pad(i)] = (float)((double)a[(t0 + i) * Q + col] - kc.mean);
            __syncthreads();
            if (has0 && has1) {
                for (int i0 = 64 * sub; i0 < n; i0 += 64 * ms) {
                    float pa, pb;
                    kde_block_sum2(xs + kde_pad(i0), (i0 + 64 <= n) ? 64 : n - i0, va, vb, nsc, pa, pb);
                    sa += (double)pa;
                    sb += (double)pb;
                }
            } else if (has0 || has1) {
                for (int i0 = 64 * sub; i0 < n; i0 += 64 * ms)
                    t += (double)kde_block_sum1(xs + kde_pad(i0), (i0 + 64 <= n) ? 64 : n - i0, v1, nsc);
            }
        }
        if (has0 != has1) { if (has0) sa = t; else sb = t; }
    };
    kde_scan_points(accumulate, true, ga, gb, part, nparts, stride, N, kc, lo, hi, step, G, sc, ms,
                    nparts > 1 ? cells + blockIdx.x : nullptr, epoch, out);
}

// grid = (columns of this batch, parts), 256 threads; dynamic shared memory: tile doubles + 8 x n_acc doubles
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_select64_tiled(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0,
                     const double* __restrict__ lohi, int G, const KdeColumn* __restrict__ cols,
                     const float* __restrict__ s32, double* __restrict__ mode_out,
                     int64_t* __restrict__ index_out, double* __restrict__ partials,
                     unsigned int* __restrict__ tickets, int tile /* multiple of 32 */, int n_acc) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    double* xs = reinterpret_cast<double*>(kde_smem_raw);      // [tile] members, float64
    double* accs = xs + tile;                                  // [8 warps][n_acc] running sums per candidate
    __shared__ float redf[8];
    __shared__ double redv[8];
    __shared__ int redi[8];
    __shared__ int cand[KDE_MAX_CAND];
    __shared__ int ncand, ntotal;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int64_t col = col0 + blockIdx.x, slot = blockIdx.x;
    const int part = blockIdx.y, nparts = gridDim.y;
    const KdeColumn kc = cols[col];
    const double lo = lohi[0], hi = lohi[1];
    const float* row = s32 + (int64_t)blockIdx.x * G;
    if (tid == 0) { ncand = 0; ntotal = 0; }
    float mx = 0.f;
    for (int g = tid; g < G; g += nthr) mx = fmaxf(mx, __ldcg(row + g));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) redf[warp] = mx;
    __syncthreads();
    mx = redf[0];
    for (int w = 1; w < nwarps; ++w) mx = fmaxf(mx, redf[w]);
    const bool degenerate = !(kc.neg_inv_2h2 > -CUDART_INF) || !(kc.neg_inv_2h2 == kc.neg_inv_2h2);
    const float thr = mx * (1.0f - KDE_TOL);
    for (int g = tid; g < G; g += nthr) {
        if (__ldcg(row + g) >= thr) {
            atomicAdd(&ntotal, 1);
            if (g % nparts == part) {
                const int k = atomicAdd(&ncand, 1);
                if (k < KDE_MAX_CAND) cand[k] = g;
            }
        }
    }
    __syncthreads();
    const bool all = ntotal > KDE_MAX_CAND || !(mx > 0.f);
    const double step = (hi - lo) / (double)(G - 1);
    int g_first = part, n_eval = ncand;                                       // <= n_acc by construction
    if (all) {                                   // (uniform over the CTA) one extra pass over the column for its extent
        __shared__ double red16[16];
        double xmn = CUDART_INF, xmx = -CUDART_INF;
        for (int64_t i = tid; i < N; i += nthr) { const double v = (double)a[i * Q + col]; xmn = fmin(xmn, v); xmx = fmax(xmx, v); }
        block_minmax(xmn, xmx, red16);
        int g_lo, g_hi;
        kde_nonzero_range(xmn, xmx, kc, lo, step, G, g_lo, g_hi);
        g_first = g_lo + ((part - g_lo) % nparts + nparts) % nparts;
        n_eval = (degenerate || g_first > g_hi) ? 0 : (g_hi - g_first) / nparts + 1;
    }
    // every warp sums its slice of the members (t, t + 256, ...) for EVERY candidate and keeps one running sum per
    // candidate (accs[warp][k], no barrier inside a tile): a part with two or three candidates still uses all its
    // warps.  Candidates are taken n_acc at a time (one batch unless they all fell to this part).
    (void)redv; (void)redi;
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int k0 = 0; k0 < n_eval; k0 += n_acc) {
        const int nb = n_eval - k0 < n_acc ? n_eval - k0 : n_acc;
        __syncthreads();
        for (int k = tid; k < nwarps * n_acc; k += nthr) accs[k] = 0.0;
        for (int64_t t0 = 0; t0 < N; t0 += tile) {
            const int n = (int)(N - t0 < tile ? N - t0 : tile);
            __syncthreads();
            for (int i = tid; i < n; i += nthr) xs[i] = (double)a[(t0 + i) * Q + col];
            __syncthreads();
            for (int k = 0; k < nb; ++k) {
                const int g = all ? g_first + (k0 + k) * nparts : cand[k0 + k];
                const double gv = kde_grid_point(g, G, lo, hi, step);
                double acc = 0.0;
                for (int i = tid; i < n; i += nthr) {
                    const double d = gv - xs[i];
                    acc += exp(d * d * kc.neg_inv_2h2);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) accs[warp * n_acc + k] += acc;
            }
        }
        __syncthreads();
        if (tid == 0) {
            for (int k = 0; k < nb; ++k) {
                const int g = all ? g_first + (k0 + k) * nparts : cand[k0 + k];
                double acc = accs[k];
                for (int w = 1; w < nwarps; ++w) acc += accs[w * n_acc + k];
                if (acc > best || (acc == best && g < besti)) { best = acc; besti = g; }
            }
        }
    }
    if (tid == 0) {
        if (nparts > 1) {
            double* mine = partials + (slot * nparts + part) * 2;
            mine[0] = best; mine[1] = (double)besti;
            __threadfence();
            const unsigned int t = atomicAdd(&tickets[slot], 1u);
            if (t != (unsigned int)(nparts - 1)) return;
            tickets[slot] = 0u;
            __threadfence();
            best = -1.0; besti = 0x7fffffff;
            for (int p = 0; p < nparts; ++p) {
                const double v = __ldcg(partials + (slot * nparts + p) * 2);
                const int gi = (int)__ldcg(partials + (slot * nparts + p) * 2 + 1);
                if (v > best || (v == best && gi < besti)) { best = v; besti = gi; }
            }
        }
        if (!(best > 0.0)) besti = 0;            // every grid point is zero: the first one is the maximum
        if (degenerate) {
            if (index_out) index_out[col] = -1;
            if (mode_out) mode_out[col] = CUDART_NAN;
        } else {
            if (index_out) index_out[col] = besti;
            if (mode_out) mode_out[col] = kde_grid_point(besti, G, lo, hi, step);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Small ensembles (N*Q <= 64 K values, the chain's own (members, 29) output): the five dependent
// launches above (two for the global range, column constants, scan, select) cost more in start-up
// latency than in work, so ONE launch does it all.  grid = (Q, n_gchunks), 256 threads:
//   1. every CTA takes the global min / max of the whole (N, Q) array itself (a few dozen loads per thread); the
//      columns of the launch are the window [col0, col0 + gridDim.x), outputs are indexed within the window
//   2. its column's members -> shared memory (float64), mean / ddof-1 variance / bandwidth
//   3. the fp32 scan of its chunk of grid points -> s32
//   4. the last CTA of a column to finish (atomic ticket) runs the float64 selection
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_small(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0, int compute_range, double* __restrict__ lohi,
            int G, int gchunk, double scott_factor_sq, float* __restrict__ s32,
            unsigned int* __restrict__ tickets, double* __restrict__ mode_out, int64_t* __restrict__ index_out) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    double* xs = reinterpret_cast<double*>(kde_smem_raw);                       // [N] members, float64
    float* xc = reinterpret_cast<float*>(kde_smem_raw + (((size_t)N * sizeof(double) + 15) & ~(size_t)15));   // [padded N] centred, fp32 (16-byte aligned)
    __shared__ double rlo[8], rhi[8], rsum[8];
    __shared__ int rnan[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t col = blockIdx.x;            // index in this launch's column window [col0, col0 + gridDim.x) of the array
    // ---- 1. grid range (np.min / np.max of the whole array; NaN propagates) ---------------------
    double lo, hi;
    if (compute_range) {
        lo = CUDART_INF; hi = -CUDART_INF;
        int has_nan = 0;
        const int64_t n = N * Q;
        for (int64_t i = tid; i < n; i += 256) {
            const double v = (double)a[i];
            has_nan |= (v != v);
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
        }
        if (lane == 0) { rlo[warp] = lo; rhi[warp] = hi; rnan[warp] = has_nan; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; ++w) { lo = fmin(lo, rlo[w]); hi = fmax(hi, rhi[w]); has_nan |= rnan[w]; }
        if (has_nan) { lo = CUDART_NAN; hi = CUDART_NAN; }
        if (col == 0 && blockIdx.y == 0 && tid == 0) { lohi[0] = lo; lohi[1] = hi; }
    } else {
        lo = lohi[0]; hi = lohi[1];
    }
    // ---- 2. column constants ------------------------------------------------------------------------
    double sum = 0.0;
    for (int64_t i = tid; i < N; i += 256) { const double v = (double)a[i * Q + col0 + col]; xs[i] = v; sum += v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) rsum[warp] = sum;
    __syncthreads();
    sum = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += rsum[w];
    const double mean = sum / (double)N;
    double ss = 0.0;
    for (int64_t i = tid; i < N; i += 256) { const double d = xs[i] - mean; ss += d * d; xc[kde_pad(i)] = (float)d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    __syncthreads();
    if (lane == 0) rsum[warp] = ss;
    __syncthreads();
    ss = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) ss += rsum[w];
    KdeColumn kc;
    kc.mean = mean;
    kc.neg_inv_2h2 = -0.5 / ((ss / (double)(N - 1)) * scott_factor_sq);        // -inf when the column is constant
    // ---- 3. fp32 scan of this CTA's share of the column's active grid range ------------------------------
    (void)gchunk;
    float* out = s32 + col * G;
    __syncthreads();                           // xc complete
    kde_scan_column(xc, N, kc, lo, hi, G, (int)blockIdx.y, (int)gridDim.y, out);
    // ---- 4. the last CTA of the column selects -----------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&tickets[col], 1u);
        s_last = (t == gridDim.y - 1);
        if (s_last) tickets[col] = 0u;            // self-cleaning for the next call
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    kde_select_column(N, col, lo, hi, G, kc, out, xs, mode_out, index_out);
}

// ------------------------------------------------------------------------------------------
// Stable ascending argsort of a short vector (the misfit ranking of ECD.py:786, `np.argsort` of the
// per-member totals; n = number of simulated maps, 50 in the reference): rank by counting,
// rank(i) = #{j : v_j < v_i or (v_j == v_i and j < i)}, NaN last (numpy's order), order[rank(i)] = i.
// O(n^2) compares on the whole machine beat any sort's launch count at these sizes.
template <typename T>
__global__ void __launch_bounds__(256)
k_argsort_count(const T* __restrict__ v, int64_t n, int64_t* __restrict__ order) {
    __shared__ T tile[1024];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const T vi = i < n ? v[i] : (T)0;
    const bool nan_i = vi != vi;
    int64_t rank = 0;
    for (int64_t j0 = 0; j0 < n; j0 += 1024) {
        const int m = (int)(n - j0 < 1024 ? n - j0 : 1024);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) tile[j] = v[j0 + j];
        __syncthreads();
        if (i < n) {
            for (int j = 0; j < m; ++j) {
                const T vj = tile[j];
                const bool nan_j = vj != vj;
                const bool less = nan_i ? !nan_j : (!nan_j && vj < vi);
                const bool equal = nan_i ? nan_j : (vj == vi);
                rank += (less || (equal && j0 + j < i)) ? 1 : 0;
            }
        }
    }
    if (i < n) order[rank] = i;
}

// ------------------------------------------------------------------------------------------
// logits -> physical parameters -> bounds check (ECD.py:42-53, 402-406,
// 183-218).  One warp per member; lane p handles parameter p.
__global__ void k_untransform_bounds(const float* __restrict__ u, int64_t B, int P, float a,
                                     float b, const double* __restrict__ smin,
                                     const double* __restrict__ sscale,
                                     const double* __restrict__ lim_lo,
                                     const double* __restrict__ lim_hi, float* __restrict__ phys,
                                     uint8_t* __restrict__ valid, int32_t* __restrict__ first_bad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    bool bad = false;
    if (lane < P) {
        const float uu = u[row * P + lane];
        const float sg = 1.0f / (1.0f + expf(-uu));                 // torch.sigmoid, fp32
        float v = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), sg));     // a + (b-a)*sigmoid(u)
        if (smin) {
            v = (float)__dsub_rn((double)v, smin[lane]);            // X -= min_   (f64 math, f32 store)
            v = (float)__ddiv_rn((double)v, sscale[lane]);          // X /= scale_
        }
        if (phys) phys[row * P + lane] = v;
        if (lim_lo) bad = ((double)v < lim_lo[lane]) || ((double)v > lim_hi[lane]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (valid) valid[row] = mask == 0;
        if (first_bad) first_bad[row] = mask ? (__ffs(mask) - 1) : -1;
    }
}

// Row vectors of mixed dtype (float32 / float64 / int64) -> one float64 block, out[c * ld + r] = row r, column c
// (column-major: one contiguous record per column).  The column-sharded statistics pack everything a rank
// computed for its columns with this single launch before the (one) result all-gather.
constexpr int kMaxPackRows = 64;
struct PackRows {
    const void* src[kMaxPackRows];
    int32_t dtype[kMaxPackRows];      // ERTDIFF_F32, ERTDIFF_F64, 2 = int64
    int32_t n;
};
__global__ void k_pack_rows(const __grid_constant__ PackRows p, int64_t ncols, int64_t ld, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncols * p.n) return;
    const int r = (int)(idx / ncols);
    const int64_t c = idx % ncols;
    double v;
    if (p.dtype[r] == ERTDIFF_F32) v = (double)static_cast<const float*>(p.src[r])[c];
    else if (p.dtype[r] == ERTDIFF_F64) v = static_cast<const double*>(p.src[r])[c];
    else v = (double)static_cast<const long long*>(p.src[r])[c];
    out[c * ld + r] = v;
}

// columns [col0, col0 + ncols) of a row-major (N, Q) array -> a contiguous (N, ncols) copy
template <typename T>
__global__ void k_slice_columns(const T* __restrict__ a, int64_t N, int64_t Q, int64_t col0, int64_t ncols, T* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * ncols) return;
    const int64_t i = idx / ncols, c = idx % ncols;
    out[idx] = a[i * Q + col0 + c];
}

// check_param_bounds alone (ECD.py:183-218) on values of either dtype: a row is dropped when any parameter is
// `< min or > max` (so a NaN never drops a row, as in the reference); first_bad = the parameter the reference's
// loop reports before it breaks.  One warp per row.
template <typename T>
__global__ void k_check_bounds(const T* __restrict__ v, int64_t B, int P, const double* __restrict__ lim_lo,
                               const double* __restrict__ lim_hi, uint8_t* __restrict__ valid,
                               int32_t* __restrict__ first_bad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    bool bad = false;
    if (lane < P) {
        const double x = (double)v[row * P + lane];
        bad = (x < lim_lo[lane]) || (x > lim_hi[lane]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (valid) valid[row] = mask == 0;
        if (first_bad) first_bad[row] = mask ? (__ffs(mask) - 1) : -1;
    }
}

}  // namespace ertdiff
