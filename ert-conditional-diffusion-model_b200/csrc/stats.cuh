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
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double lo = CUDART_INF, hi = -CUDART_INF, nn = 0.0;
        for (int b = 0; b < nblocks; ++b) {
            lo = fmin(lo, part[3 * b]); hi = fmax(hi, part[3 * b + 1]); nn += part[3 * b + 2];
        }
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
// Gaussian-KDE mode, ECD.py:751-762.  One CTA per column: the N members sit in shared memory
// as float64, every thread evaluates pdf(g) = sum_i exp(-(g - x_i)^2 / (2 h^2)) for a strided
// set of grid points (members in order, float64 throughout, as scipy does), and the CTA reduces
// to the FIRST maximum.  h^2 = var_ddof1 * N^(-2/5) (Scott).  The normalisation constant is
// common to all grid points of a column and does not change the argmax, so it is skipped.
// grid[i] = lo + i*step (separately rounded multiply and add, as np.linspace), grid[G-1] = hi.
//
// Launch: grid = (Q, n_gchunks); CTA (col, gc) scans grid points [gc*gchunk, (gc+1)*gchunk) and
// writes its (best value, best index) to part_val/part_idx[col*n_gchunks + gc]; k_kde_final
// picks the first maximum per column.  Splitting the grid keeps all SMs busy when Q is small
// (the (N, 29) parameter posteriors); for map-shaped inputs n_gchunks = 1.
template <typename T>
__global__ void __launch_bounds__(256)
k_kde_mode(const T* __restrict__ a, int64_t N, int64_t Q, const double* __restrict__ lohi,
           int G, int gchunk, double scott_factor_sq, double* __restrict__ part_val,
           int* __restrict__ part_idx) {
    extern __shared__ __align__(16) unsigned char kde_smem_raw[];
    double* xs = reinterpret_cast<double*>(kde_smem_raw);    // [N]
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ double s_bcast[2];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = nthr >> 5;
    const int64_t col = blockIdx.x;
    const int g_begin = blockIdx.y * gchunk;
    const int g_end = min(G, g_begin + gchunk);

    double lsum = 0.0;
    for (int64_t i = tid; i < N; i += nthr) {
        const double v = (double)a[i * Q + col];
        xs[i] = v;
        lsum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red[w];
        s_bcast[0] = s / (double)N;
    }
    __syncthreads();
    const double mean = s_bcast[0];
    double lss = 0.0;
    for (int64_t i = tid; i < N; i += nthr) {
        const double d = xs[i] - mean;
        lss += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lss += __shfl_xor_sync(0xffffffffu, lss, o);
    __syncthreads();
    if (lane == 0) red[warp] = lss;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red[w];
        const double h2 = (s / (double)(N - 1)) * scott_factor_sq;
        s_bcast[1] = -0.5 / h2;
    }
    __syncthreads();
    const double neg_inv_2h2 = s_bcast[1];

    const double lo = lohi[0], hi = lohi[1];
    const double step = (hi - lo) / (double)(G - 1);
    double best = -1.0;
    int besti = 0x7fffffff;
    // two grid points per thread per pass: independent exp chains
    for (int g0 = g_begin + tid; g0 < g_end; g0 += 2 * nthr) {
        const int g1 = g0 + nthr;
        const double ga = (g0 == G - 1) ? hi : __dadd_rn(__dmul_rn((double)g0, step), lo);
        const double gb = (g1 >= g_end) ? ga : (g1 == G - 1) ? hi
                                                              : __dadd_rn(__dmul_rn((double)g1, step), lo);
        double pa = 0.0, pb = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            const double xi = xs[i];
            const double da = ga - xi, db = gb - xi;
            pa += exp(da * da * neg_inv_2h2);
            pb += exp(db * db * neg_inv_2h2);
        }
        if (pa > best) { best = pa; besti = g0; }
        if (g1 < g_end && pb > best) { best = pb; besti = g1; }
    }
    // first maximum: larger value wins, ties go to the smaller index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = best; redi[warp] = besti; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w)
            if (red[w] > best || (red[w] == best && redi[w] < besti)) { best = red[w]; besti = redi[w]; }
        part_val[col * gridDim.y + blockIdx.y] = best;
        part_idx[col * gridDim.y + blockIdx.y] = besti;
    }
}

__global__ void k_kde_final(const double* __restrict__ part_val, const int* __restrict__ part_idx,
                            int64_t Q, int n_gchunks, const double* __restrict__ lohi, int G,
                            double* __restrict__ mode_out, int64_t* __restrict__ index_out) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= Q) return;
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int c = 0; c < n_gchunks; ++c) {            // chunks are in grid order: strict > keeps the first
        const double v = part_val[col * n_gchunks + c];
        const int i = part_idx[col * n_gchunks + c];
        if (v > best || (v == best && i < besti)) { best = v; besti = i; }
    }
    const double lo = lohi[0], hi = lohi[1];
    const double step = (hi - lo) / (double)(G - 1);
    if (index_out) index_out[col] = besti;
    if (mode_out)
        mode_out[col] = (besti == G - 1) ? hi : __dadd_rn(__dmul_rn((double)besti, step), lo);
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

}  // namespace ertdiff
