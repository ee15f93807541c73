// Per-member data misfit of an ensemble of simulated ERT maps against the observed map:
//   WSSE per (member, survey)   ECD.py:764-783   np.average((pred-obs)**2 / (A*|obs|+B)**2) over the L measurements
//   WSSE total per member       ECD.py:785        WSSE_sim.sum(axis=1)
//   MSE per member              ECD.py:927-930    sklearn mean_squared_error over the flattened (L*C) map
//
// All three are numpy add-reductions of a contiguous 1-D array, i.e. numpy's *pairwise* summation
// (numpy/_core/src/umath/loops_utils.h.src, `@TYPE@_pairwise_sum`: fewer than 8 elements are added
// in order; up to 128 elements go through 8 interleaved accumulators that are combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and followed by the n%8 tail in order; longer arrays are split
// at n/2 rounded down to a multiple of 8 and the two halves' sums added).  The result depends on that
// tree, so the kernels walk the same tree: the host flattens the recursion for a given n into a
// leaf table (start, length <= 128) and a list of internal nodes ordered by height
// (`PairwisePlan`); 8 lanes own the 8 accumulators of a leaf, the tree above the leaves is evaluated
// level by level in shared memory.  Every element is computed with the non-contracting
// __f*_rn / __d*_rn intrinsics in numpy's operation order, so the sums are bit-identical to numpy's
// for f32 and f64 maps alike.  HBM-bound: each map is read once from DRAM (the survey columns of one
// member are interleaved in memory, the CTA's neighbouring lanes share the lines through L1).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ertdiff {

struct PairwisePlan {
    const int2* leaves;     // (start, length)
    const int2* nodes;      // internal node k -> value indices of its children; its own value index is n_leaves + k
    const int* level_off;   // nodes [level_off[h], level_off[h+1]) depend only on lower levels
    int n_leaves, n_nodes, n_levels;
};

__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }

// sum of one leaf by an aligned group of 8 lanes (k = lane & 7); valid in the group's lane 0
template <typename T, typename F>
__device__ __forceinline__ T pairwise_leaf(int start, int len, int k, F elem) {
    // every lane of the warp reaches the shuffles, whatever its group's leaf length
    T r = T(0);
    const int body = len >= 8 ? len - (len & 7) : 0;        // fewer than 8 elements: added in order from 0
    if (body) {
        // a leaf has at most 128 elements = 16 per lane: fetch them all before the dependent add chain, so
        // that a lane keeps up to 16 (x2 maps) loads in flight instead of one round trip per add
        T e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = elem(start + (8 * j < body ? 8 * j : 0) + k);   // unpredicated: slots past the leaf re-read its first row
        r = e[0];
#pragma unroll
        for (int j = 1; j < 16; ++j)
            if (8 * j < body) r = rn_add(r, e[j]);
    }
    r = rn_add(r, __shfl_xor_sync(0xffffffffu, r, 1));      // (r0+r1) ... : IEEE addition commutes, so both
    r = rn_add(r, __shfl_xor_sync(0xffffffffu, r, 2));      // lanes of a pair hold the same bits
    r = rn_add(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (k == 0)
        for (int i = body; i < len; ++i) r = rn_add(r, elem(start + i));
    return r;
}

// numpy's pairwise sum of n <= 128 values by one thread (the survey totals)
template <typename T>
__device__ T pairwise_small(const T* a, int n) {
    if (n < 8) {
        T r = T(0);
        for (int i = 0; i < n; ++i) r = rn_add(r, a[i]);
        return r;
    }
    T r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    const int body = n - (n & 7);
    for (int i = 8; i < body; i += 8)
        for (int k = 0; k < 8; ++k) r[k] = rn_add(r[k], a[i + k]);
    T res = rn_add(rn_add(rn_add(r[0], r[1]), rn_add(r[2], r[3])), rn_add(rn_add(r[4], r[5]), rn_add(r[6], r[7])));
    for (int i = body; i < n; ++i) res = rn_add(res, a[i]);
    return res;
}

// the tree above the leaves, for `nseq` independent sequences whose values sit `stride` apart in smem
template <typename T>
__device__ __forceinline__ void pairwise_combine(const PairwisePlan& pl, T* val, int nseq, int stride) {
    for (int h = 0; h < pl.n_levels; ++h) {
        __syncthreads();
        const int lo = pl.level_off[h], cnt = pl.level_off[h + 1] - lo;
        for (int i = threadIdx.x; i < cnt * nseq; i += blockDim.x) {
            const int s = i / cnt, k = lo + i % cnt;
            const int2 ch = pl.nodes[k];
            val[s * stride + pl.n_leaves + k] = rn_add(val[s * stride + ch.x], val[s * stride + ch.y]);
        }
    }
    __syncthreads();
}

// one CTA per member: WSSE of each of the C surveys and their total.  sims (N, L, C), obs (L, C).
// The map is walked in chunks of whole leaves (at most `rcap` rows): phase 1 reads the chunk's rows of
// both maps fully coalesced and parks the weighted squared errors survey-major in shared memory, phase 2
// sums each (leaf, survey) from there with the 8 accumulators of a lane group.
constexpr int kMisfitThreads = 512;
constexpr int kMisfitBatch = 8;       // elements per thread whose loads are issued together in phase 1
template <typename T>
__global__ void __launch_bounds__(kMisfitThreads) k_misfit_wsse(const T* __restrict__ sims, const T* __restrict__ obs, int L, int C,
                                                                T A, T B, int rcap, PairwisePlan pl, T* __restrict__ wsse,
                                                                T* __restrict__ wsse_total) {
    extern __shared__ __align__(16) unsigned char misfit_smem[];
    const int stride = pl.n_leaves + pl.n_nodes;
    const int rs = rcap | 1;                                  // odd row pitch: the survey-major stores spread over the banks
    T* val = reinterpret_cast<T*>(misfit_smem);               // [C][stride] tree values, then C survey means
    T* W = val + (size_t)C * stride + C;                      // [C][rs]
    const T* sim = sims + (int64_t)blockIdx.x * L * C;
    const int k = threadIdx.x & 7, group = threadIdx.x >> 3, n_groups = kMisfitThreads >> 3;
    const int dl = kMisfitThreads / C, de = kMisfitThreads % C;   // (row, survey) advance of one CTA-wide stride
    int leaf_lo = 0;
    while (leaf_lo < pl.n_leaves) {                           // uniform over the CTA
        const int r0 = pl.leaves[leaf_lo].x;
        int leaf_hi = leaf_lo, r1 = r0;
        while (leaf_hi < pl.n_leaves) {
            const int2 lf = pl.leaves[leaf_hi];
            if (lf.x + lf.y - r0 > rcap) break;
            r1 = lf.x + lf.y;
            ++leaf_hi;
        }
        // phase 1: flat elements [r0*C, r1*C), consecutive threads on consecutive addresses
        {
            const int n = (r1 - r0) * C;
            int l = (int)threadIdx.x / C, es = (int)threadIdx.x % C;
            const T* so = obs + (int64_t)r0 * C;
            const T* sp = sim + (int64_t)r0 * C;
            for (int f0 = threadIdx.x; f0 < n; f0 += kMisfitBatch * kMisfitThreads) {
                T o[kMisfitBatch], p[kMisfitBatch];
#pragma unroll
                for (int u = 0; u < kMisfitBatch; ++u) {      // all loads of the batch first
                    const int f = f0 + u * kMisfitThreads;
                    o[u] = f < n ? so[f] : T(1);
                    p[u] = f < n ? sp[f] : T(1);
                }
#pragma unroll
                for (int u = 0; u < kMisfitBatch; ++u) {
                    const T sd = rn_add(rn_mul(A, fabs(o[u])), B);             // A*np.abs(observations)+B
                    const T d = rn_sub(p[u], o[u]);
                    if (f0 + u * kMisfitThreads < n)
                        W[es * rs + l] = rn_div(rn_mul(d, d), rn_mul(sd, sd)); // (predictions - observations)**2/(sd)**2
                    l += dl; es += de;
                    if (es >= C) { es -= C; ++l; }
                }
            }
        }
        __syncthreads();
        // phase 2: one lane group per (leaf, survey)
        const int n_tasks = (leaf_hi - leaf_lo) * C;
        for (int base = 0; base < n_tasks; base += n_groups) {   // uniform trip count: the leaf sum uses warp shuffles
            const int task = base + group;
            const bool live = task < n_tasks;
            const int leaf = leaf_lo + (live ? task / C : 0), es = live ? task % C : 0;
            const int2 lf = pl.leaves[leaf];
            const T* w = W + es * rs - r0;
            const T v = pairwise_leaf<T>(lf.x, lf.y, k, [&](int l) { return w[l]; });
            if (live && k == 0) val[es * stride + leaf] = v;
        }
        __syncthreads();
        leaf_lo = leaf_hi;
    }
    pairwise_combine(pl, val, C, stride);
    T* out = val + (size_t)C * stride;
    if (threadIdx.x < C) {
        const T m = rn_div(val[threadIdx.x * stride + stride - 1], (T)L);
        out[threadIdx.x] = m;
        wsse[(int64_t)blockIdx.x * C + threadIdx.x] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) wsse_total[blockIdx.x] = pairwise_small(out, C);
}

// one CTA per member: mean over the flattened map of (obs - sim)^2
template <typename T>
__global__ void __launch_bounds__(kMisfitThreads) k_misfit_mse(const T* __restrict__ sims, const T* __restrict__ obs, int64_t n,
                                                    PairwisePlan pl, T* __restrict__ mse) {
    extern __shared__ __align__(16) unsigned char misfit_smem[];
    T* val = reinterpret_cast<T*>(misfit_smem);
    const T* sim = sims + (int64_t)blockIdx.x * n;
    const int k = threadIdx.x & 7, group = threadIdx.x >> 3, n_groups = blockDim.x >> 3;
    for (int base = 0; base < pl.n_leaves; base += n_groups) {
        const int leaf = base + group;
        const bool live = leaf < pl.n_leaves;
        const int2 lf = pl.leaves[live ? leaf : 0];
        const T v = pairwise_leaf<T>(lf.x, lf.y, k, [&](int i) {
            const T d = rn_sub(obs[i], sim[i]);                        // y_true - y_pred
            return rn_mul(d, d);
        });
        if (live && k == 0) val[leaf] = v;
    }
    pairwise_combine(pl, val, 1, 0);
    if (threadIdx.x == 0) mse[blockIdx.x] = rn_div(val[pl.n_leaves + pl.n_nodes - 1], (T)n);
}

// ---- 1-D Wasserstein distance between a simulated map and the observed one -----------------------
// ECD.py:860, 898-899: scipy.stats.wasserstein_distance(sim.flatten(), obs.flatten()), i.e. scipy's
// _cdf_distance with p = 1 (scipy/stats/_stats_py.py): both samples as float64; all_values = the sorted
// union; deltas = diff(all_values); u_cdf[k] = #{u <= all_values[k]} / n, v_cdf likewise;
// result = vecdot(|u_cdf - v_cdf|, deltas).  The dot product runs through BLAS in scipy, so its
// summation order is the host library's: parity is to a relative 1e-12, not bitwise.
//
// k_sort_rows_f64: one CTA per row, bitonic sort in global memory (the rows are L2-resident; padded to a
// power of two with +inf).  k_wasserstein: one CTA per pair.  Ranks instead of a merge: u_i lands at
// i + #{v < u_i}, v_j at j + #{u <= v_j} (ties: u first, as the stable sort of concat(u, v) leaves
// them).  The count of u-elements up to position k equals scipy's searchsorted(..., 'right') wherever
// all_values[k+1] > all_values[k]; inside a run of equal values delta is 0 and the term vanishes either way.
template <typename T>
__global__ void __launch_bounds__(1024) k_sort_rows_f64(const T* __restrict__ in, int64_t row_stride, int n, int npad,
                                                        double* __restrict__ out /* (rows, npad) */) {
    double* a = out + (int64_t)blockIdx.x * npad;
    const T* src = in + (int64_t)blockIdx.x * row_stride;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) a[i] = i < n ? (double)src[i] : __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int idx = threadIdx.x; idx < (npad >> 1); idx += blockDim.x) {
                const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));     // element whose bit j is 0
                const int p = i | j;
                const double x = a[i], y = a[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int lower_bound_f64(const double* a, int n, double v) {   // #{a < v}
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_f64(const double* a, int n, double v) {   // #{a <= v}
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(1024) k_wasserstein(const double* __restrict__ u_sorted /* (N, upad) */, int n, int upad,
                                                      const double* __restrict__ v_sorted, int m,
                                                      double* __restrict__ merged /* (N, n+m) */,
                                                      int* __restrict__ cnt_u /* (N, n+m) */, double* __restrict__ out) {
    __shared__ double red[32];
    const double* u = u_sorted + (int64_t)blockIdx.x * upad;
    double* mg = merged + (int64_t)blockIdx.x * (n + m);
    int* cu = cnt_u + (int64_t)blockIdx.x * (n + m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = u[i];
        const int r = i + lower_bound_f64(v_sorted, m, x);
        mg[r] = x; cu[r] = i + 1;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double y = v_sorted[j];
        const int c = upper_bound_f64(u, n, y);
        mg[j + c] = y; cu[j + c] = c;
    }
    __syncthreads();
    double acc = 0.0;
    const double dn = (double)n, dm = (double)m;
    for (int k = threadIdx.x; k < n + m - 1; k += blockDim.x) {
        const int c = cu[k];
        const double ucdf = __ddiv_rn((double)c, dn), vcdf = __ddiv_rn((double)(k + 1 - c), dm);
        acc += fabs(ucdf - vcdf) * (mg[k + 1] - mg[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[blockIdx.x] = t;
    }
}

}  // namespace ertdiff
