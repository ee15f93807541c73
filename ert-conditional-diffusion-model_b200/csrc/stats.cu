// Ensemble statistics, misfit metrics and the un-transform epilogue: C-ABI entry points that take no
// model handle (include/ertdiff_b200.h) and their per-device scratch.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "chain_params.cuh"
#include "stats.cuh"
#include "misfit.cuh"

namespace ertdiff {

// ---- per-device scratch for the statistics entry points (they take no model handle) --------
// Grown on demand, kept for the life of the process: no allocation on the hot path.  The scratch (and the KDE
// tickets) is shared by every call on a device, so calls are serialised by a lease: host threads take turns (the
// mutex is held while a call enqueues its kernels), and a call issued on another stream than the previous one
// first waits, on the device, for that call's work -- the scratch is never in use by two calls at once.
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;       // recorded after the last call's kernels
    cudaStream_t last = nullptr;
    bool used = false;
    unsigned int* tickets = nullptr;  // persistent, self-cleaning per-column counters (last-CTA-done pattern of the KDE kernels)
    unsigned long long* kde_cells = nullptr;   // per column of a scan launch: (launch epoch << 32 | running fp32 maximum)
    unsigned int kde_epoch = 0;
};
// two independent regions per device: 0 = the KDE kernels, min/max and everything else, 1 = the sorted runs of the
// percentile path -- so that the fused summary can run the percentiles beside the KDE kernels on another stream
static Workspace g_ws[64][2];
static std::recursive_mutex g_ws_mutex;

class WorkspaceLease {
 public:
    explicit WorkspaceLease(cudaStream_t st, int region = 0) : st_(st), lock_(g_ws_mutex) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) w_ = &g_ws[dev][region];
        if (w_ && w_->used && w_->last != st_ && w_->done) cudaStreamWaitEvent(st_, w_->done, 0);
    }
    ~WorkspaceLease() {
        if (!w_ || !touched_) return;
        if (!w_->done && cudaEventCreateWithFlags(&w_->done, cudaEventDisableTiming) != cudaSuccess) { w_->done = nullptr; return; }
        cudaEventRecord(w_->done, st_);
        w_->last = st_; w_->used = true;
    }
    int get(size_t bytes, void** out) {
        if (!w_) return fail(ERTDIFF_ERR_ARG, "workspace: device index out of range");
        touched_ = true;
        if (w_->bytes < bytes) {
            if (w_->ptr) { ERT_CUDA(cudaDeviceSynchronize()); cudaFree(w_->ptr); w_->ptr = nullptr; w_->bytes = 0; }
            const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
            ERT_CUDA(cudaMalloc(&w_->ptr, want));
            w_->bytes = want;
        }
        *out = w_->ptr;
        return 0;
    }
    int tickets(unsigned int** out) {
        if (!w_) return fail(ERTDIFF_ERR_ARG, "kde: device index out of range");
        touched_ = true;
        if (!w_->tickets) {
            ERT_CUDA(cudaMalloc(&w_->tickets, kTicketSlots * sizeof(unsigned int)));
            ERT_CUDA(cudaMemset(w_->tickets, 0, kTicketSlots * sizeof(unsigned int)));
        }
        *out = w_->tickets;
        return 0;
    }
    // the coarse-to-fine KDE scan's per-column cells and a fresh tag for the launch that is about to use them (a
    // cell tagged by an earlier launch is ignored by the kernel, so nothing is ever reset -- except when the 32-bit
    // tag wraps)
    int kde_cells(unsigned long long** out, unsigned int* epoch) {
        if (!w_) return fail(ERTDIFF_ERR_ARG, "kde: device index out of range");
        touched_ = true;
        if (!w_->kde_cells) {
            ERT_CUDA(cudaMalloc(&w_->kde_cells, kTicketSlots * sizeof(unsigned long long)));
            ERT_CUDA(cudaMemset(w_->kde_cells, 0, kTicketSlots * sizeof(unsigned long long)));
        }
        if (++w_->kde_epoch == 0) {
            ERT_CUDA(cudaMemsetAsync(w_->kde_cells, 0, kTicketSlots * sizeof(unsigned long long), st_));
            w_->kde_epoch = 1;
        }
        *out = w_->kde_cells;
        *epoch = w_->kde_epoch;
        return 0;
    }
    static constexpr int kTicketSlots = 4096;

 private:
    cudaStream_t st_;
    std::unique_lock<std::recursive_mutex> lock_;
    Workspace* w_ = nullptr;
    bool touched_ = false;
};

static int colstats_attr() {
    static PerDeviceOnce once;
    bool& done = *once.slot();
    if (!done) {
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        ERT_CUDA(cudaFuncSetAttribute(k_colstats_smallq<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM_BYTES));
        done = true;
    }
    return 0;
}

}  // namespace ertdiff

namespace ertdiff {

// Percentiles of columns of any length: k_sort_runs + k_select_runs (stats.cuh).  Columns are processed in
// batches so that the sorted-run scratch stays within ~1 GiB.
template <typename T, typename G, typename O>
static int percentiles_by_runs_t(const void* d_a, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs, void* d_out,
                                 int CH, cudaStream_t st) {
    using K = typename SortKey<T>::K;
    const int64_t R64 = (N + CH - 1) / CH;
    ERT_REQUIRE(R64 <= 4096, "ensemble_percentiles: more than 4096 runs per column (N > 33.5 M members)");
    const int R = (int)R64;
    int64_t qb = (int64_t)((size_t(1) << 30) / ((size_t)N * sizeof(K)));
    if (qb < 1) qb = 1;
    if (qb > Q) qb = Q;
    int CT = (int)((128 * 1024) / ((size_t)CH * sizeof(K)));      // columns per sorting CTA (coalescing)
    if (CT > 8) CT = 8;
    if (CT < 1) CT = 1;
    while (CT > 1 && (R64 * ((qb + CT - 1) / CT) < 2 * kNumSMs || CT > qb)) CT >>= 1;
    const size_t smem_sort = (size_t)CH * CT * sizeof(K);
    const size_t smem_sel = (size_t)4 * 3 * R * sizeof(int);
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_sort_runs<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_select_runs<T, G, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    WorkspaceLease lease(st, 1);
    char* ws = nullptr;
    const size_t runs_bytes = ((size_t)qb * N * sizeof(K) + 255) & ~(size_t)255;
    if (int rc = lease.get(runs_bytes + (size_t)qb * sizeof(int), (void**)&ws)) return rc;
    K* runs = (K*)ws;
    int* nanflag = (int*)(ws + runs_bytes);
    const int threads = (int64_t)CH * CT / 2 >= 1024 ? 1024 : ((int64_t)CH * CT / 2 >= 256 ? 256 : 64);
    const int nq = (int)qs.size();
    for (int64_t c0 = 0; c0 < Q; c0 += qb) {
        const int nc = (int)((Q - c0) < qb ? (Q - c0) : qb);
        ERT_CUDA(cudaMemsetAsync(nanflag, 0, (size_t)nc * sizeof(int), st));
        k_sort_runs<T><<<dim3((unsigned)R, (unsigned)((nc + CT - 1) / CT)), threads, smem_sort, st>>>(
            (const T*)d_a, N, Q, c0, nc, CH, CT, runs, nanflag);
        ERT_LAUNCH_CHECK("k_sort_runs");
        for (int k0 = 0; k0 < nq; k0 += kMaxPctlQueries) {
            PctlQueryPack pack{};
            pack.n = (nq - k0) < kMaxPctlQueries ? (nq - k0) : kMaxPctlQueries;
            for (int k = 0; k < pack.n; ++k) pack.q[k] = qs[k0 + k];
            const int64_t warps = (int64_t)nc * pack.n;
            k_select_runs<T, G, O><<<(unsigned)((warps + 3) / 4), 128, smem_sel, st>>>(
                runs, N, Q, c0, nc, CH, R, nanflag, pack, (O*)d_out + (size_t)k0 * Q);
            ERT_LAUNCH_CHECK("k_select_runs");
        }
    }
    return 0;
}

template <typename T, typename G, typename O, int E>
static int percentiles_by_warps_e(const void* d_a, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs, void* d_out,
                                  cudaStream_t st) {
    // columns per warp: as many as keep the tile within ~96 KB, fewer when that would leave SMs without a CTA
    int CPW = 4;
    while (CPW > 1 && ((size_t)N * (8 * CPW + 1) * sizeof(T) > 96 * 1024 || (Q + 8 * CPW - 1) / (8 * CPW) < 2 * kNumSMs)) CPW >>= 1;
    const size_t smem = (size_t)N * (8 * CPW + 1) * sizeof(T);
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_percentiles_warp<T, G, O, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    const unsigned grid = (unsigned)((Q + 8 * CPW - 1) / (8 * CPW));
    const int nq = (int)qs.size();
    for (int k0 = 0; k0 < nq; k0 += kMaxPctlQueries) {
        PctlQueryPack pack{};
        pack.n = (nq - k0) < kMaxPctlQueries ? (nq - k0) : kMaxPctlQueries;
        for (int k = 0; k < pack.n; ++k) pack.q[k] = qs[k0 + k];
        k_percentiles_warp<T, G, O, E><<<grid, 256, smem, st>>>((const T*)d_a, N, Q, CPW, pack, (O*)d_out + (size_t)k0 * Q);
        ERT_LAUNCH_CHECK("k_percentiles_warp");
    }
    return 0;
}

template <typename T, typename G, typename O>
static int percentiles_by_warps_t(const void* d_a, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs, void* d_out,
                                  cudaStream_t st) {
    if (N <= 32) return percentiles_by_warps_e<T, G, O, 1>(d_a, N, Q, qs, d_out, st);
    if (N <= 64) return percentiles_by_warps_e<T, G, O, 2>(d_a, N, Q, qs, d_out, st);
    if (N <= 128) return percentiles_by_warps_e<T, G, O, 4>(d_a, N, Q, qs, d_out, st);
    if (N <= 256) return percentiles_by_warps_e<T, G, O, 8>(d_a, N, Q, qs, d_out, st);
    if (N <= 512) return percentiles_by_warps_e<T, G, O, 16>(d_a, N, Q, qs, d_out, st);
    return percentiles_by_warps_e<T, G, O, 32>(d_a, N, Q, qs, d_out, st);
}

static int percentiles_by_warps(const void* d_a, int dtype, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs,
                                int index_dtype, void* d_out, cudaStream_t st) {
    if (dtype == ERTDIFF_F32 && index_dtype == ERTDIFF_F32)
        return percentiles_by_warps_t<float, float, float>(d_a, N, Q, qs, d_out, st);
    if (dtype == ERTDIFF_F32) return percentiles_by_warps_t<float, double, double>(d_a, N, Q, qs, d_out, st);
    return percentiles_by_warps_t<double, double, double>(d_a, N, Q, qs, d_out, st);
}

// exact selection by radix in shared memory (k_percentiles_select): medium-length columns, many columns, few quantiles
template <typename T, typename G, typename O>
static int percentiles_by_select_t(const void* d_a, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs, void* d_out,
                                   cudaStream_t st) {
    using K = typename SortKey<T>::K;
    int CT = (int)((64 * 1024) / ((size_t)N * sizeof(K)));
    if (CT > 16) CT = 16;
    if (CT < 1) CT = 1;
    while (CT > 1 && (Q + CT - 1) / CT < 4 * kNumSMs) CT >>= 1;
    const size_t smem = ((size_t)CT * N + 2 * PS_CAP) * sizeof(K) + (257 + CT) * sizeof(int);
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_percentiles_select<T, G, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    PctlQueryPack pack{};
    pack.n = (int)qs.size();
    for (int k = 0; k < pack.n; ++k) pack.q[k] = qs[k];
    k_percentiles_select<T, G, O><<<(unsigned)((Q + CT - 1) / CT), 256, smem, st>>>((const T*)d_a, N, Q, CT, pack, (O*)d_out);
    ERT_LAUNCH_CHECK("k_percentiles_select");
    return 0;
}

static int percentiles_by_select(const void* d_a, int dtype, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs,
                                 int index_dtype, void* d_out, cudaStream_t st) {
    if (dtype == ERTDIFF_F32 && index_dtype == ERTDIFF_F32)
        return percentiles_by_select_t<float, float, float>(d_a, N, Q, qs, d_out, st);
    if (dtype == ERTDIFF_F32) return percentiles_by_select_t<float, double, double>(d_a, N, Q, qs, d_out, st);
    return percentiles_by_select_t<double, double, double>(d_a, N, Q, qs, d_out, st);
}

static int percentiles_by_runs(const void* d_a, int dtype, int64_t N, int64_t Q, const std::vector<PctlQuery>& qs,
                               int index_dtype, void* d_out, int CH, cudaStream_t st) {
    if (dtype == ERTDIFF_F32 && index_dtype == ERTDIFF_F32)
        return percentiles_by_runs_t<float, float, float>(d_a, N, Q, qs, d_out, CH, st);
    if (dtype == ERTDIFF_F32) return percentiles_by_runs_t<float, double, double>(d_a, N, Q, qs, d_out, CH, st);
    return percentiles_by_runs_t<double, double, double>(d_a, N, Q, qs, d_out, CH, st);
}

// members below which the one-launch KDE kernel (k_kde_small) is used (ERTDIFF_KDE_SMALL_MAX overrides, for sweeps)
static int64_t kde_small_max_members() {
    if (const char* e = std::getenv("ERTDIFF_KDE_SMALL_MAX")) return std::atoll(e);
    return 1024;
}

// Run length of the sorted-runs percentile path for an (N, Q) array, or 0 when a shared-memory kernel serves it
// (those need no scratch, so the fused summary can run them beside the KDE kernels).
static int percentile_run_length(int dtype, int64_t N, int64_t Q) {
    // (ERTDIFF_PCTL_RUN_LEN: forces the run path with that run length, for tests)
    if (const char* e = std::getenv("ERTDIFF_PCTL_RUN_LEN")) {
        const int v = std::atoi(e);
        if (v >= 2 && v <= 8192 && (v & (v - 1)) == 0) return v;
    }
    const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
    int64_t NP64 = 1;
    while (NP64 < N) NP64 <<= 1;
    // too long for one CTA's shared memory -> runs.  Few columns (the chain's own (members, 29) output) also go
    // through runs as soon as a column exceeds 2048 members: a single CTA per column would leave the machine to
    // Q CTAs sorting 8K..32K elements each (18,944 x 29: 506 -> 92 us), while short runs spread the sort over
    // R x Q CTAs and the selection costs a warp per (column, query)
    if ((size_t)NP64 * esz + 64 > 200 * 1024) return Q <= 1024 ? 2048 : 8192;
    if (N > 2048 && Q <= 1024) return 2048;
    return 0;
}

}  // namespace ertdiff

using namespace ertdiff;

#pragma GCC visibility push(default)
extern "C" {

int ertdiff_ensemble_moments(const void* d_a, int dtype, int64_t N, int64_t Q, void* d_mean,
                             void* d_std, void* d_var, void* stream) {
    ERT_REQUIRE(d_a && N > 0 && Q > 0, "ensemble_moments: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((Q + 127) / 128);
    if (Q <= SQ_MAXQ && N >= 64 && (dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64)) {
        // few columns: one CTA streams the rows through shared memory (see k_colstats_smallq)
        const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
        const size_t smem = (size_t)N * esz < (size_t)SQ_SMEM_BYTES ? (size_t)N * esz : (size_t)SQ_SMEM_BYTES;
        if (int rc = colstats_attr()) return rc;
        if (dtype == ERTDIFF_F32)
            k_colstats_smallq<float, 0><<<(unsigned)Q, 256, smem, st>>>((const float*)d_a, N, Q, (float*)d_mean, (float*)d_std, (float*)d_var, 0.0, nullptr);
        else
            k_colstats_smallq<double, 0><<<(unsigned)Q, 256, smem, st>>>((const double*)d_a, N, Q, (double*)d_mean, (double*)d_std, (double*)d_var, 0.0, nullptr);
        ERT_LAUNCH_CHECK("k_colstats_smallq");
        return 0;
    }
    if (dtype == ERTDIFF_F32)
        k_moments<float><<<grid, 128, 0, st>>>((const float*)d_a, N, Q, (float*)d_mean,
                                               (float*)d_std, (float*)d_var);
    else if (dtype == ERTDIFF_F64)
        k_moments<double><<<grid, 128, 0, st>>>((const double*)d_a, N, Q, (double*)d_mean,
                                                (double*)d_std, (double*)d_var);
    else
        return fail(ERTDIFF_ERR_ARG, "ensemble_moments: bad dtype");
    ERT_LAUNCH_CHECK("k_moments");
    return 0;
}

int ertdiff_ensemble_percentiles(const void* d_a, int dtype, int64_t N, int64_t Q,
                                 const double* h_q, int32_t nq, int index_dtype, void* d_out,
                                 void* stream) {
    ERT_REQUIRE(d_a && d_out && h_q && N > 0 && Q > 0 && nq > 0, "ensemble_percentiles: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_percentiles: bad dtype");
    ERT_REQUIRE(!(dtype == ERTDIFF_F64 && index_dtype == ERTDIFF_F32),
                "ensemble_percentiles: float64 data always uses float64 index arithmetic");
    cudaStream_t st = (cudaStream_t)stream;
    // numpy's index arithmetic (function_base._quantile, method 'linear'), in index_dtype
    std::vector<PctlQuery> qs(nq);
    for (int k = 0; k < nq; ++k) {
        ERT_REQUIRE(h_q[k] >= 0.0 && h_q[k] <= 100.0, "ensemble_percentiles: q outside [0,100]");
        PctlQuery& q = qs[k];
        if (index_dtype == ERTDIFF_F32) {
            const float quant = (float)h_q[k] / 100.0f;
            const float vi = (float)(N - 1) * quant;
            const float lo = floorf(vi);
            q.gamma_f = vi - lo; q.gamma_d = 0.0;
            q.lo = (int32_t)lo; q.hi = q.lo + 1;
            if (vi >= (float)(N - 1)) q.lo = q.hi = (int32_t)(N - 1);
        } else {
            const double quant = h_q[k] / 100.0;
            const double vi = (double)(N - 1) * quant;
            const double lo = floor(vi);
            q.gamma_d = vi - lo; q.gamma_f = 0.f;
            q.lo = (int32_t)lo; q.hi = q.lo + 1;
            if (vi >= (double)(N - 1)) q.lo = q.hi = (int32_t)(N - 1);
        }
    }
    const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
    const size_t budget = 200 * 1024;
    int64_t NP64 = 1;
    while (NP64 < N) NP64 <<= 1;
    // many medium-length columns, few quantiles (maps of 1024 .. 16K members, 25/50/75): exact radix selection, no sort
    const bool select_forced = std::getenv("ERTDIFF_PCTL_SELECT") != nullptr;
    if (!std::getenv("ERTDIFF_PCTL_RUN_LEN") && !std::getenv("ERTDIFF_PCTL_NO_SELECT") && nq <= 4 && N < (int64_t(1) << 30) &&
        (size_t)N * esz <= 128 * 1024 && (select_forced || (N > 512 && Q > 1024)))
        return percentiles_by_select(d_a, dtype, N, Q, qs, index_dtype, d_out, st);
    const int run_len = percentile_run_length(dtype, N, Q);
    if (run_len) return percentiles_by_runs(d_a, dtype, N, Q, qs, index_dtype, d_out, run_len, st);
    // short columns of many-column arrays (the reference's 50 realisations of a 65,702-pixel map): one warp sorts a
    // column in registers.  (Measured: beyond 256 members the shuffle count makes it slower than the shared-memory
    // kernel -- 1024 x 65,702 f64: 4.2 vs 3.5 ms -- and with few columns it leaves the machine to Q/8 CTAs.)
    const bool warp_ok = std::getenv("ERTDIFF_PCTL_WARP") ? N <= 1024 : (N <= 256 && Q >= 16 * kNumSMs);
    if (warp_ok && !std::getenv("ERTDIFF_PCTL_NO_WARP"))
        return percentiles_by_warps(d_a, dtype, N, Q, qs, index_dtype, d_out, st);
    const int NP = (int)NP64;
    int CT = 32;
    while (CT > 1 && ((size_t)CT * NP * esz + CT * sizeof(int) > budget / 2 || (int64_t)CT > Q)) CT >>= 1;
    // keep enough CTAs in flight
    while (CT > 1 && (Q + CT - 1) / CT < 2 * kNumSMs) CT >>= 1;
    const size_t smem = (size_t)CT * NP * esz + CT * sizeof(int);
    const int threads = (int64_t)CT * NP / 2 >= 1024 ? 1024 : ((int64_t)CT * NP / 2 >= 256 ? 256 : 64);
    const unsigned grid = (unsigned)((Q + CT - 1) / CT);
    cudaError_t e = cudaSuccess;
    // queries travel as a by-value kernel argument, kMaxPctlQueries per launch
    for (int k0 = 0; k0 < nq && e == cudaSuccess; k0 += kMaxPctlQueries) {
        PctlQueryPack pack{};
        pack.n = (nq - k0) < kMaxPctlQueries ? (nq - k0) : kMaxPctlQueries;
        for (int k = 0; k < pack.n; ++k) pack.q[k] = qs[k0 + k];
#define ERT_PCT_LAUNCH(T, G, O)                                                              \
    do {                                                                                     \
        e = cudaFuncSetAttribute(k_percentiles<T, G, O>,                                     \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);  \
        if (e == cudaSuccess)                                                                \
            k_percentiles<T, G, O><<<grid, threads, smem, st>>>((const T*)d_a, N, Q, NP, CT, \
                                                               pack, (O*)d_out + (size_t)k0 * Q); \
    } while (0)
        if (dtype == ERTDIFF_F32 && index_dtype == ERTDIFF_F32) ERT_PCT_LAUNCH(float, float, float);
        else if (dtype == ERTDIFF_F32) ERT_PCT_LAUNCH(float, double, double);
        else ERT_PCT_LAUNCH(double, double, double);
#undef ERT_PCT_LAUNCH
        if (e != cudaSuccess) return fail(ERTDIFF_ERR_CUDA, std::string("percentiles attr: ") + cudaGetErrorString(e));
        ERT_LAUNCH_CHECK("k_percentiles");
    }
    return 0;
}

int ertdiff_interval_coverage(const double* d_low, const double* d_upp, const double* d_truth,
                              int32_t n_intervals, int64_t Q, int32_t P, int32_t* d_counts, void* stream) {
    ERT_REQUIRE(d_low && d_upp && d_truth && d_counts, "interval_coverage: NULL pointer");
    ERT_REQUIRE(n_intervals > 0 && Q > 0 && P > 0 && P <= kPPad && Q % P == 0,
                "interval_coverage: need n_intervals > 0, 0 < P <= 32 and Q a multiple of P");
    k_interval_coverage<<<(unsigned)n_intervals, 256, 0, (cudaStream_t)stream>>>(d_low, d_upp, d_truth, Q, P, d_counts);
    ERT_LAUNCH_CHECK("k_interval_coverage");
    return 0;
}

int ertdiff_minmax(const void* d_a, int dtype, int64_t n, double* d_out2, void* stream) {
    ERT_REQUIRE(d_a && d_out2 && n > 0, "minmax: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (blocks < 1) blocks = 1;
    WorkspaceLease lease(st);
    double* part = nullptr;
    if (int rc = lease.get((size_t)blocks * 3 * sizeof(double), (void**)&part)) return rc;
    if (dtype == ERTDIFF_F32) k_minmax_partial<float><<<blocks, 256, 0, st>>>((const float*)d_a, n, part);
    else if (dtype == ERTDIFF_F64) k_minmax_partial<double><<<blocks, 256, 0, st>>>((const double*)d_a, n, part);
    else return fail(ERTDIFF_ERR_ARG, "minmax: bad dtype");
    ERT_LAUNCH_CHECK("k_minmax_partial");
    k_minmax_final<<<1, 32, 0, st>>>(part, blocks, d_out2);
    ERT_LAUNCH_CHECK("k_minmax_final");
    return 0;
}

int ertdiff_ensemble_kde_mode(const void* d_a, int dtype, int64_t N, int64_t Q,
                              const double* d_lohi, int32_t n_grid, double* d_mode,
                              int64_t* d_index, void* stream) {
    ERT_REQUIRE(d_a && d_lohi && N > 1 && Q > 0 && n_grid > 1, "ensemble_kde_mode: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_kde_mode: bad dtype");
    // columns that do not fit one CTA's shared memory are streamed through it tile by tile
    // (ERTDIFF_KDE_TILE: tile length override, for tests)
    bool tiled = (size_t)N * 8 > 200 * 1024;
    int tile_override = 0;
    if (const char* e = std::getenv("ERTDIFF_KDE_TILE")) {
        tile_override = std::atoi(e) / 64 * 64;
        if (tile_override >= 64) tiled = true; else tile_override = 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int G = n_grid;
    // scipy: factor = neff**(-1/(d+4)) with d = 1, neff = N; covariance = data_cov * factor**2
    const double factor = std::pow((double)N, -1.0 / 5.0);
    const double f2 = factor * factor;
    // columns per batch: the fp32 scan of a batch lives in the workspace (<= 128 MiB)
    int64_t qb = (int64_t)((128u << 20) / ((size_t)G * sizeof(float)));
    if (qb < 1) qb = 1;
    if (qb > Q) qb = Q;
    if (qb > 65535 * 16) qb = 65535 * 16;
    WorkspaceLease lease(st);
    void* ws = nullptr;
    const size_t cols_bytes = ((size_t)Q * sizeof(KdeColumn) + 255) & ~(size_t)255;
    // the float64 selection of a column is shared by several CTAs when there are few columns and many
    // members (one CTA would re-evaluate ~100 candidates x N members alone)
    int sel_parts = (qb <= 4096 && N >= 1024) ? (qb * 8 <= 4 * kNumSMs ? 8 : (qb * 2 <= 4 * kNumSMs ? 2 : 1)) : 1;
    // a handful of long columns (a rank's share of a big ensemble): more parts, so that the candidates' float64
    // sums spread over the whole machine
    while (sel_parts >= 8 && sel_parts < 32 && qb * sel_parts * 2 <= 4 * kNumSMs && N >= 16384) sel_parts *= 2;
    const size_t part_bytes = (size_t)qb * sel_parts * 2 * sizeof(double);
    if (int rc = lease.get(cols_bytes + part_bytes + (size_t)qb * G * sizeof(float), &ws)) return rc;
    unsigned int* tk = nullptr;
    if (int rc = lease.tickets(&tk)) return rc;
    KdeColumn* cols = (KdeColumn*)ws;
    double* partials = (double*)((char*)ws + cols_bytes);
    float* s32 = (float*)((char*)ws + cols_bytes + part_bytes);
    const bool f32in = dtype == ERTDIFF_F32;
    if (Q <= SQ_MAXQ && N >= 64) {
        static_assert(sizeof(KdeColumn) == 2 * sizeof(double), "KdeColumn is written as two doubles");
        const size_t esz = f32in ? 4 : 8;
        const size_t smem = (size_t)N * esz < (size_t)SQ_SMEM_BYTES ? (size_t)N * esz : (size_t)SQ_SMEM_BYTES;
        if (int rc = colstats_attr()) return rc;
        if (f32in) k_colstats_smallq<float, 1><<<(unsigned)Q, 256, smem, st>>>((const float*)d_a, N, Q, nullptr, nullptr, nullptr, f2, (double*)cols);
        else k_colstats_smallq<double, 1><<<(unsigned)Q, 256, smem, st>>>((const double*)d_a, N, Q, nullptr, nullptr, nullptr, f2, (double*)cols);
        ERT_LAUNCH_CHECK("k_colstats_smallq");
    } else {
        const unsigned grid = (unsigned)((Q * 32 + 255) / 256);
        if (f32in) k_kde_prepare<float><<<grid, 256, 0, st>>>((const float*)d_a, N, Q, f2, cols);
        else k_kde_prepare<double><<<grid, 256, 0, st>>>((const double*)d_a, N, Q, f2, cols);
        ERT_LAUNCH_CHECK("k_kde_prepare");
    }
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32_tiled<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_scan32_tiled<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64_tiled<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_kde_select64_tiled<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    for (int64_t c0 = 0; c0 < Q; c0 += qb) {
        const int64_t nc = (Q - c0) < qb ? (Q - c0) : qb;
        // CTAs per column: one 256-thread pass over the grid each while nc * n_gchunks stays within ~4
        // CTAs per SM; with many columns (maps) one CTA walks the whole grid of its column
        int n_gchunks = (G + 255) / 256;
        while (n_gchunks > 1 && nc * n_gchunks > 4 * kNumSMs) n_gchunks = (n_gchunks + 1) / 2;
        // few columns of a long ensemble (a rank's share of the chain's (members, 29) output: 4 columns x 20 chunks
        // = 80 CTAs, each thread summing every member for its two grid points) leave most of the machine idle: let
        // ms lanes share a grid point and split the members, with ms times as many CTAs per column
        int ms = 1;
        while (ms < 32 && nc * n_gchunks * 2 <= 4 * kNumSMs && N / (2 * ms) >= 512) { ms *= 2; n_gchunks *= 2; }
        // (evening the CTA count out to whole CTAs per SM -- 444 instead of 320 -- was measured slower: every CTA
        // streams its whole column, so more CTAs cost more than the imbalance they remove)
        const int threads = 256;
        // coarse-to-fine scan (stats.cuh): largest coarse stride (ERTDIFF_KDE_COARSE: 1 scans every point), and the
        // cells through which the CTAs of a column share its running maximum
        int max_stride = 32;
        if (const char* e = std::getenv("ERTDIFF_KDE_COARSE")) { const int v = std::atoi(e); max_stride = v >= 1 && v <= 64 ? v : 32; }
        unsigned long long* cells = nullptr;
        unsigned int epoch = 0;
        if (n_gchunks > 1) {
            ERT_REQUIRE(nc <= WorkspaceLease::kTicketSlots, "ensemble_kde_mode: internal: more shared columns than cells");
            if (int rc = lease.kde_cells(&cells, &epoch)) return rc;
        }
        const dim3 grid((unsigned)nc, (unsigned)n_gchunks);
        const dim3 sgrid((unsigned)nc, (unsigned)sel_parts);
        if (tiled) {
            // scan: 44 K fp32 members per tile (16 K when lanes share grid points: more CTAs per SM); selection: what
            // is left of 200 KB after one float64 accumulator per candidate (or per grid point of this part, should
            // the scan be flat)
            const int stile = tile_override ? tile_override : (ms > 1 ? 16 * 1024 : 44 * 1024);
            // running sums: one per candidate of this part and warp.  Candidates are consecutive grid indices dealt
            // round-robin to the parts, so a part holds at most ceil(KDE_MAX_CAND / parts) + 1 of them -- or, should
            // the scan be flat, its share of the grid
            const int n_acc = std::max((KDE_MAX_CAND + sel_parts - 1) / sel_parts + 1, (G + sel_parts - 1) / sel_parts);
            ERT_REQUIRE((size_t)8 * n_acc * 8 + 64 * 8 <= 160 * 1024, "ensemble_kde_mode: n_grid too large");
            int dtile = (int)((200 * 1024 - (size_t)8 * n_acc * 8) / 8) / 32 * 32;
            if (dtile > 8192) dtile = 8192;                      // 64 KB tiles: three CTAs per SM
            if (tile_override && tile_override < dtile) dtile = tile_override;
            const size_t dsmem = ((size_t)dtile + 8 * (size_t)n_acc) * 8;
            if (f32in) {
                k_kde_scan32_tiled<float><<<grid, threads, (size_t)kde_padded(stile) * 4, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, cols, s32, stile, ms, max_stride, cells, epoch);
                ERT_LAUNCH_CHECK("k_kde_scan32_tiled");
                k_kde_select64_tiled<float><<<sgrid, 256, dsmem, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk, dtile, n_acc);
            } else {
                k_kde_scan32_tiled<double><<<grid, threads, (size_t)kde_padded(stile) * 4, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, cols, s32, stile, ms, max_stride, cells, epoch);
                ERT_LAUNCH_CHECK("k_kde_scan32_tiled");
                k_kde_select64_tiled<double><<<sgrid, 256, dsmem, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk, dtile, n_acc);
            }
            ERT_LAUNCH_CHECK("k_kde_select64_tiled");
        } else if (f32in) {
            k_kde_scan32<float><<<grid, threads, (size_t)kde_padded(N) * 4, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, ms, cols, s32, max_stride, cells, epoch);
            ERT_LAUNCH_CHECK("k_kde_scan32");
            k_kde_select64<float><<<sgrid, 256, (size_t)N * 8, st>>>((const float*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk);
            ERT_LAUNCH_CHECK("k_kde_select64");
        } else {
            k_kde_scan32<double><<<grid, threads, (size_t)kde_padded(N) * 4, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, ms, cols, s32, max_stride, cells, epoch);
            ERT_LAUNCH_CHECK("k_kde_scan32");
            k_kde_select64<double><<<sgrid, 256, (size_t)N * 8, st>>>((const double*)d_a, N, Q, c0, d_lohi, G, cols, s32, d_mode, d_index, partials, tk);
            ERT_LAUNCH_CHECK("k_kde_select64");
        }
    }
    return 0;
}

int ertdiff_ensemble_kde_mode_auto(const void* d_a, int dtype, int64_t N, int64_t Q, int32_t n_grid,
                                   double* d_lohi, double* d_mode, int64_t* d_index, void* stream) {
    ERT_REQUIRE(d_a && d_lohi && N > 1 && Q > 0 && n_grid > 1, "ensemble_kde_mode_auto: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_kde_mode_auto: bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    // (from ~1000 members on, the staged kernels win: their float64 selection is shared by 8 CTAs per column)
    const bool small = N * Q <= 65536 && N < kde_small_max_members() && Q <= 4096;
    if (!small) {
        if (int rc = ertdiff_minmax(d_a, dtype, N * Q, d_lohi, stream)) return rc;
        return ertdiff_ensemble_kde_mode(d_a, dtype, N, Q, d_lohi, n_grid, d_mode, d_index, stream);
    }
    WorkspaceLease lease(st);
    // one fused launch (k_kde_small); tickets: a persistent, self-cleaning counter per column
    unsigned int* tk = nullptr;
    if (int rc = lease.tickets(&tk)) return rc;
    const int G = n_grid;
    void* ws = nullptr;
    if (int rc = lease.get((size_t)Q * G * sizeof(float), &ws)) return rc;
    // CTAs per column: one 256-thread pass over the grid each (every thread scans one point) when that
    // still fits in one wave of co-resident CTAs (a second, nearly empty wave would double the
    // duration); otherwise fewer, longer chunks
    int n_gchunks = (G + 255) / 256;
    if (const char* e = std::getenv("ERTDIFF_KDE_GCHUNKS")) n_gchunks = std::atoi(e) > 0 ? std::atoi(e) : 1;
    else
        while (n_gchunks > 1 && Q * n_gchunks > 4 * kNumSMs) n_gchunks = (n_gchunks + 1) / 2;
    const int gchunk = (G + n_gchunks - 1) / n_gchunks;
    const double factor = std::pow((double)N, -1.0 / 5.0);
    const dim3 grid((unsigned)Q, (unsigned)n_gchunks);
    const size_t smem = (((size_t)N * 8 + 15) & ~(size_t)15) + (size_t)kde_padded(N) * 4;
    if (dtype == ERTDIFF_F32)
        k_kde_small<float><<<grid, 256, smem, st>>>((const float*)d_a, N, Q, 0, 1, d_lohi, G, gchunk, factor * factor,
                                                    (float*)ws, tk, d_mode, d_index);
    else
        k_kde_small<double><<<grid, 256, smem, st>>>((const double*)d_a, N, Q, 0, 1, d_lohi, G, gchunk, factor * factor,
                                                     (float*)ws, tk, d_mode, d_index);
    ERT_LAUNCH_CHECK("k_kde_small");
    return 0;
}

// the fused small-ensemble KDE launch on a column window of the array (grid range from the WHOLE array)
static int kde_small_window(const void* d_a, int dtype, int64_t N, int64_t Q, int64_t col0, int64_t ncols, int G,
                            double* d_lohi, double* d_mode, int64_t* d_index, cudaStream_t st) {
    WorkspaceLease lease(st);
    unsigned int* tk = nullptr;
    if (int rc = lease.tickets(&tk)) return rc;
    void* ws = nullptr;
    if (int rc = lease.get((size_t)ncols * G * sizeof(float), &ws)) return rc;
    int n_gchunks = (G + 255) / 256;
    while (n_gchunks > 1 && ncols * n_gchunks > 4 * kNumSMs) n_gchunks = (n_gchunks + 1) / 2;
    const int gchunk = (G + n_gchunks - 1) / n_gchunks;
    const double factor = std::pow((double)N, -1.0 / 5.0);
    const dim3 grid((unsigned)ncols, (unsigned)n_gchunks);
    const size_t smem = (((size_t)N * 8 + 15) & ~(size_t)15) + (size_t)kde_padded(N) * 4;
    if (dtype == ERTDIFF_F32)
        k_kde_small<float><<<grid, 256, smem, st>>>((const float*)d_a, N, Q, col0, 1, d_lohi, G, gchunk, factor * factor,
                                                    (float*)ws, tk, d_mode, d_index);
    else
        k_kde_small<double><<<grid, 256, smem, st>>>((const double*)d_a, N, Q, col0, 1, d_lohi, G, gchunk, factor * factor,
                                                     (float*)ws, tk, d_mode, d_index);
    ERT_LAUNCH_CHECK("k_kde_small");
    return 0;
}

int ertdiff_argsort_stable(const void* d_v, int dtype, int64_t n, int64_t* d_order, void* stream) {
    ERT_REQUIRE(d_v && d_order && n > 0, "argsort_stable: bad arguments");
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == ERTDIFF_F32) k_argsort_count<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)d_v, n, d_order);
    else if (dtype == ERTDIFF_F64) k_argsort_count<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)d_v, n, d_order);
    else return fail(ERTDIFF_ERR_ARG, "argsort_stable: bad dtype");
    ERT_LAUNCH_CHECK("k_argsort_count");
    return 0;
}

int ertdiff_untransform_bounds(const float* d_u, int64_t B, int32_t P, float a, float b,
                               const double* d_scaler_min, const double* d_scaler_scale,
                               const double* d_lim_lo, const double* d_lim_hi, float* d_phys,
                               uint8_t* d_valid, int32_t* d_first_bad, void* stream) {
    ERT_REQUIRE(d_u && B > 0 && P > 0 && P <= 32, "untransform_bounds: bad arguments");
    ERT_REQUIRE((d_scaler_min == nullptr) == (d_scaler_scale == nullptr), "untransform_bounds: give both scaler arrays or neither");
    ERT_REQUIRE((d_lim_lo == nullptr) == (d_lim_hi == nullptr), "untransform_bounds: give both limit arrays or neither");
    const int64_t threads = B * 32;
    k_untransform_bounds<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_u, B, P, a, b, d_scaler_min, d_scaler_scale, d_lim_lo, d_lim_hi, d_phys, d_valid, d_first_bad);
    ERT_LAUNCH_CHECK("k_untransform_bounds");
    return 0;
}

int ertdiff_pack_rows_f64(const void* const* h_rows, const int32_t* h_dtypes, int32_t n_rows, int64_t ncols,
                          int64_t ld, double* d_out, void* stream) {
    ERT_REQUIRE(h_rows && h_dtypes && d_out && n_rows > 0 && n_rows <= kMaxPackRows && ncols >= 0 && ld >= n_rows,
                "pack_rows_f64: bad arguments (at most 64 rows, ld >= n_rows)");
    if (ncols == 0) return 0;
    PackRows p{};
    p.n = n_rows;
    for (int r = 0; r < n_rows; ++r) {
        ERT_REQUIRE(h_rows[r] && h_dtypes[r] >= 0 && h_dtypes[r] <= 2, "pack_rows_f64: NULL row or bad dtype");
        p.src[r] = h_rows[r]; p.dtype[r] = h_dtypes[r];
    }
    const int64_t n = ncols * n_rows;
    k_pack_rows<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, ncols, ld, d_out);
    ERT_LAUNCH_CHECK("k_pack_rows");
    return 0;
}

// Everything the path reports about an ensemble, for a window of columns, in ONE call (the column-sharded multi-GPU
// statistics issue this once per step: no per-statistic host round trips, no allocations).
int ertdiff_ensemble_summary(const void* d_a, int dtype, int64_t N, int64_t Q, int64_t col0, int64_t ncols,
                             const double* h_q, int32_t nq, int32_t n_grid, double* d_lohi_out, double* d_out, int64_t ld,
                             void* stream) {
    ERT_REQUIRE(d_a && d_out && N > 1 && Q > 0 && col0 >= 0 && ncols > 0 && col0 + ncols <= Q, "ensemble_summary: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "ensemble_summary: bad dtype");
    ERT_REQUIRE(nq >= 0 && 5 + nq <= kMaxPackRows && ld >= 5 + nq && (nq == 0 || h_q) && n_grid > 1, "ensemble_summary: bad query / ld");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == ERTDIFF_F32 ? 4 : 8;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // private scratch (the statistics entry points called below use the shared workspace themselves)
    static void* scratch[64] = {};
    static size_t scratch_bytes[64] = {};
    int dev = 0;
    ERT_CUDA(cudaGetDevice(&dev));
    ERT_REQUIRE(dev >= 0 && dev < 64, "ensemble_summary: device index out of range");
    WorkspaceLease lease(st);           // serialises concurrent callers on this device (the nested calls re-enter it)
    const bool whole = ncols == Q;
    const size_t b_a = whole ? 0 : up((size_t)N * ncols * esz), b_m = up((size_t)ncols * esz), b_p = up((size_t)(nq ? nq : 1) * ncols * 8),
                 b_v = up((size_t)ncols * 8);
    const size_t need = b_a + 3 * b_m + b_p + 2 * b_v + 256;
    if (scratch_bytes[dev] < need) {
        if (scratch[dev]) { ERT_CUDA(cudaDeviceSynchronize()); cudaFree(scratch[dev]); scratch[dev] = nullptr; scratch_bytes[dev] = 0; }
        ERT_CUDA(cudaMalloc(&scratch[dev], need));
        scratch_bytes[dev] = need;
    }
    char* base = (char*)scratch[dev];
    void* cols = whole ? const_cast<void*>(d_a) : (void*)base;
    char* p_mean = base + b_a; char* p_std = p_mean + b_m; char* p_var = p_std + b_m; char* p_pct = p_var + b_m;
    double* p_mode = (double*)(p_pct + b_p);
    int64_t* p_idx = (int64_t*)((char*)p_mode + b_v);
    double* p_lohi = (double*)((char*)p_idx + b_v);
    if (!whole) {
        const int64_t n = N * ncols;
        if (dtype == ERTDIFF_F32) k_slice_columns<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)d_a, N, Q, col0, ncols, (float*)cols);
        else k_slice_columns<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const double*)d_a, N, Q, col0, ncols, (double*)cols);
        ERT_LAUNCH_CHECK("k_slice_columns");
    }
    // the KDE grid spans the min / max of the WHOLE array (ECD.py:749-751), not of this window.  Small ensembles: the
    // one-launch KDE kernel takes the range itself (every CTA scans the whole array: N*Q <= 64 K values)
    // (measured: beyond ~1000 members the staged kernels win even for a 4-column window -- 2048 x 4: 43 us of
    // dependent launches against 90 us for the one-launch form, whose every CTA scans the whole array for the range
    // and whose last CTA selects alone)
    const bool kde_small = N * Q <= 65536 && N < kde_small_max_members() && ncols <= ertdiff::WorkspaceLease::kTicketSlots;
    if (!kde_small)
        if (int rc = ertdiff_minmax(d_a, dtype, N * Q, p_lohi, stream)) return rc;
    // the moments are a dependent add chain per column (numpy's order; 170 us at 18,944 members): they run on a side
    // stream beside the percentile and KDE kernels and are joined before the packing launch
    // the three statistics are independent per-column reductions.  The moments are a dependent add chain per column
    // (numpy's order; 170 us at 18,944 members): they run on a side stream beside the percentile and KDE kernels.  The
    // percentiles join them on a second side stream (their run path has its own scratch region).  Both are joined
    // before the packing launch.
    static cudaStream_t side[64][2] = {};
    static cudaEvent_t ev_fork[64] = {}, ev_join[64][2] = {};
    if (!side[dev][0]) {
        for (int k = 0; k < 2; ++k) {
            ERT_CUDA(cudaStreamCreateWithFlags(&side[dev][k], cudaStreamNonBlocking));
            ERT_CUDA(cudaEventCreateWithFlags(&ev_join[dev][k], cudaEventDisableTiming));
        }
        ERT_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
    }
    ERT_CUDA(cudaEventRecord(ev_fork[dev], st));
    ERT_CUDA(cudaStreamWaitEvent(side[dev][0], ev_fork[dev], 0));
    if (int rc = ertdiff_ensemble_moments(cols, dtype, N, ncols, p_mean, p_std, p_var, side[dev][0])) return rc;
    ERT_CUDA(cudaEventRecord(ev_join[dev][0], side[dev][0]));
    const bool pct_beside = nq > 0;      // (the run path has its own scratch region: it, too, runs beside the KDE kernels)
    if (pct_beside) {
        ERT_CUDA(cudaStreamWaitEvent(side[dev][1], ev_fork[dev], 0));
        if (int rc = ertdiff_ensemble_percentiles(cols, dtype, N, ncols, h_q, nq, ERTDIFF_F64, p_pct, side[dev][1])) return rc;
        ERT_CUDA(cudaEventRecord(ev_join[dev][1], side[dev][1]));
    } else if (nq) {
        if (int rc = ertdiff_ensemble_percentiles(cols, dtype, N, ncols, h_q, nq, ERTDIFF_F64, p_pct, stream)) return rc;
    }
    if (kde_small) {
        if (int rc = kde_small_window(d_a, dtype, N, Q, col0, ncols, n_grid, p_lohi, p_mode, p_idx, st)) return rc;
    } else {
        if (int rc = ertdiff_ensemble_kde_mode(cols, dtype, N, ncols, p_lohi, n_grid, p_mode, p_idx, stream)) return rc;
    }
    if (d_lohi_out) ERT_CUDA(cudaMemcpyAsync(d_lohi_out, p_lohi, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    ERT_CUDA(cudaStreamWaitEvent(st, ev_join[dev][0], 0));
    if (pct_beside) ERT_CUDA(cudaStreamWaitEvent(st, ev_join[dev][1], 0));
    const void* rows[kMaxPackRows];
    int32_t dts[kMaxPackRows];
    rows[0] = p_mean; rows[1] = p_std; rows[2] = p_var;
    dts[0] = dts[1] = dts[2] = dtype;
    for (int k = 0; k < nq; ++k) { rows[3 + k] = p_pct + (size_t)k * ncols * 8; dts[3 + k] = ERTDIFF_F64; }
    rows[3 + nq] = p_mode; dts[3 + nq] = ERTDIFF_F64;
    rows[4 + nq] = p_idx; dts[4 + nq] = 2;
    return ertdiff_pack_rows_f64(rows, dts, 5 + nq, ncols, ld, d_out, stream);
}

int ertdiff_check_bounds(const void* d_v, int dtype, int64_t B, int32_t P, const double* d_lim_lo,
                         const double* d_lim_hi, uint8_t* d_valid, int32_t* d_first_bad, void* stream) {
    ERT_REQUIRE(d_v && d_lim_lo && d_lim_hi && B > 0 && P > 0 && P <= 32, "check_bounds: bad arguments");
    const int64_t threads = B * 32;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (dtype == ERTDIFF_F32)
        k_check_bounds<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)d_v, B, P, d_lim_lo, d_lim_hi, d_valid, d_first_bad);
    else if (dtype == ERTDIFF_F64)
        k_check_bounds<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)d_v, B, P, d_lim_lo, d_lim_hi, d_valid, d_first_bad);
    else return fail(ERTDIFF_ERR_ARG, "check_bounds: bad dtype");
    ERT_LAUNCH_CHECK("k_check_bounds");
    return 0;
}

// ---- per-member misfit metrics (ECD.py:764-785, 927-930) ---------------------------------------
// numpy's pairwise-summation tree for n elements, flattened once per (device, n) and kept on the device
extern "C++" {
struct PairwiseNodes {
    std::vector<int2> leaves, nodes;      // nodes: children as (is_node ? -1 - k : leaf index) until fixed up
    std::vector<int> height;
    int build(int64_t start, int64_t n) {
        if (n <= 128) {
            leaves.push_back(make_int2((int)start, (int)n));
            return (int)leaves.size() - 1;
        }
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        const int l = build(start, n2), r = build(start + n2, n - n2);
        const int hl = l < 0 ? height[-1 - l] : 0, hr = r < 0 ? height[-1 - r] : 0;
        nodes.push_back(make_int2(l, r));
        height.push_back((hl > hr ? hl : hr) + 1);
        return -(int)nodes.size();
    }
};

static int pairwise_plan(int64_t n, PairwisePlan* out) {
    static std::map<std::pair<int, int64_t>, PairwisePlan> cache;
    int dev = 0;
    ERT_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::recursive_mutex> lock(g_ws_mutex);
    auto it = cache.find({dev, n});
    if (it != cache.end()) { *out = it->second; return 0; }
    PairwiseNodes t;
    t.build(0, n);
    const int nl = (int)t.leaves.size(), nn = (int)t.nodes.size();
    int n_levels = 0;
    for (int h : t.height) n_levels = h > n_levels ? h : n_levels;
    // order the internal nodes by height (children always sit on a lower level), remap the child indices
    std::vector<int> order, pos(nn), level_off(n_levels + 1, 0);
    for (int h = 1; h <= n_levels; ++h) {
        level_off[h - 1] = (int)order.size();
        for (int k = 0; k < nn; ++k)
            if (t.height[k] == h) { pos[k] = (int)order.size(); order.push_back(k); }
    }
    level_off[n_levels] = nn;
    std::vector<int2> nodes(nn);
    for (int i = 0; i < nn; ++i) {
        const int2 c = t.nodes[order[i]];
        nodes[i] = make_int2(c.x < 0 ? nl + pos[-1 - c.x] : c.x, c.y < 0 ? nl + pos[-1 - c.y] : c.y);
    }
    int2 *d_leaves = nullptr, *d_nodes = nullptr;
    int* d_off = nullptr;
    ERT_CUDA(cudaMalloc(&d_leaves, sizeof(int2) * nl));
    ERT_CUDA(cudaMalloc(&d_nodes, sizeof(int2) * (nn > 0 ? nn : 1)));
    ERT_CUDA(cudaMalloc(&d_off, sizeof(int) * (n_levels + 1)));
    ERT_CUDA(cudaMemcpy(d_leaves, t.leaves.data(), sizeof(int2) * nl, cudaMemcpyHostToDevice));
    if (nn) ERT_CUDA(cudaMemcpy(d_nodes, nodes.data(), sizeof(int2) * nn, cudaMemcpyHostToDevice));
    ERT_CUDA(cudaMemcpy(d_off, level_off.data(), sizeof(int) * (n_levels + 1), cudaMemcpyHostToDevice));
    PairwisePlan pl{d_leaves, d_nodes, d_off, nl, nn, n_levels};
    cache[{dev, n}] = pl;
    *out = pl;
    return 0;
}

template <typename T>
static int launch_misfit(const void* d_sims, const void* d_obs, int64_t N, int64_t L, int64_t C, double A, double B,
                         void* d_wsse, void* d_wsse_total, void* d_mse, cudaStream_t st) {
    static PerDeviceOnce once;
    bool& attr_set = *once.slot();
    if (!attr_set) {
        ERT_CUDA(cudaFuncSetAttribute(k_misfit_wsse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ERT_CUDA(cudaFuncSetAttribute(k_misfit_mse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (d_wsse) {
        PairwisePlan pl;
        if (int rc = pairwise_plan(L, &pl)) return rc;
        // rows per chunk: whole leaves (<= 128 rows each), sized so that several CTAs stay resident per SM
        int rcap = (int)(L < 512 ? (L < 128 ? 128 : L) : 512);
        while (rcap > 128 && (size_t)C * (rcap | 1) * sizeof(T) > 64 * 1024) rcap = rcap / 2 < 128 ? 128 : rcap / 2;
        const size_t smem = ((size_t)C * (pl.n_leaves + pl.n_nodes) + C + (size_t)C * (rcap | 1)) * sizeof(T);
        ERT_REQUIRE(smem <= 200 * 1024, "misfit_metrics: L x C too large for the on-chip summation tree");
        k_misfit_wsse<T><<<(unsigned)N, kMisfitThreads, smem, st>>>((const T*)d_sims, (const T*)d_obs, (int)L, (int)C, (T)A, (T)B,
                                                                    rcap, pl, (T*)d_wsse, (T*)d_wsse_total);
        ERT_LAUNCH_CHECK("k_misfit_wsse");
    }
    if (d_mse) {
        PairwisePlan pl;
        if (int rc = pairwise_plan(L * C, &pl)) return rc;
        const size_t smem = (size_t)(pl.n_leaves + pl.n_nodes) * sizeof(T);
        ERT_REQUIRE(smem <= 200 * 1024, "misfit_metrics: map too large for the on-chip summation tree");
        k_misfit_mse<T><<<(unsigned)N, kMisfitThreads, smem, st>>>((const T*)d_sims, (const T*)d_obs, L * C, pl, (T*)d_mse);
        ERT_LAUNCH_CHECK("k_misfit_mse");
    }
    return 0;
}
}  // extern "C++"

int ertdiff_misfit_metrics(const void* d_sims, const void* d_obs, int dtype, int64_t N, int64_t L, int64_t C,
                           double A, double B, void* d_wsse, void* d_wsse_total, void* d_mse, void* stream) {
    ERT_REQUIRE(d_sims && d_obs && N > 0 && L > 0 && C > 0, "misfit_metrics: bad arguments");
    ERT_REQUIRE(C <= 128 && L * C < (int64_t(1) << 31), "misfit_metrics: need C <= 128 and L*C < 2^31");
    ERT_REQUIRE((d_wsse == nullptr) == (d_wsse_total == nullptr), "misfit_metrics: give both WSSE outputs or neither");
    ERT_REQUIRE(d_wsse || d_mse, "misfit_metrics: no output requested");
    if (dtype == ERTDIFF_F32) return launch_misfit<float>(d_sims, d_obs, N, L, C, A, B, d_wsse, d_wsse_total, d_mse, (cudaStream_t)stream);
    if (dtype == ERTDIFF_F64) return launch_misfit<double>(d_sims, d_obs, N, L, C, A, B, d_wsse, d_wsse_total, d_mse, (cudaStream_t)stream);
    return fail(ERTDIFF_ERR_ARG, "misfit_metrics: bad dtype");
}

extern "C++" {
template <typename T>
static int launch_sort_rows(const void* d_in, int64_t rows, int64_t stride, int n, int npad, double* d_out, cudaStream_t st) {
    k_sort_rows_f64<T><<<(unsigned)rows, 1024, 0, st>>>((const T*)d_in, stride, n, npad, d_out);
    ERT_LAUNCH_CHECK("k_sort_rows_f64");
    return 0;
}
}  // extern "C++"

int ertdiff_wasserstein_distance(const void* d_u, const void* d_v, int dtype, int64_t N, int64_t n, int64_t m,
                                 double* d_out, void* stream) {
    ERT_REQUIRE(d_u && d_v && d_out && N > 0 && n > 0 && m > 0, "wasserstein_distance: bad arguments");
    ERT_REQUIRE(dtype == ERTDIFF_F32 || dtype == ERTDIFF_F64, "wasserstein_distance: bad dtype");
    ERT_REQUIRE(n <= (1 << 24) && m <= (1 << 24), "wasserstein_distance: at most 2^24 values per sample");
    cudaStream_t st = (cudaStream_t)stream;
    int upad = 1, vpad = 1;
    while (upad < n) upad <<= 1;
    while (vpad < m) vpad <<= 1;
    // scratch: sorted u rows, sorted v, the merged values and the u-counts of every pair
    const size_t b_us = (size_t)N * upad * 8, b_vs = (size_t)vpad * 8, b_mg = (size_t)N * (n + m) * 8, b_cu = (size_t)N * (n + m) * 4;
    WorkspaceLease lease(st);
    char* ws = nullptr;
    if (int rc = lease.get(b_us + b_vs + b_mg + b_cu, (void**)&ws)) return rc;
    double* us = (double*)ws;
    double* vs = (double*)(ws + b_us);
    double* mg = (double*)(ws + b_us + b_vs);
    int* cu = (int*)(ws + b_us + b_vs + b_mg);
    int rc = dtype == ERTDIFF_F32 ? launch_sort_rows<float>(d_u, N, n, (int)n, upad, us, st) : launch_sort_rows<double>(d_u, N, n, (int)n, upad, us, st);
    if (rc) return rc;
    rc = dtype == ERTDIFF_F32 ? launch_sort_rows<float>(d_v, 1, m, (int)m, vpad, vs, st) : launch_sort_rows<double>(d_v, 1, m, (int)m, vpad, vs, st);
    if (rc) return rc;
    k_wasserstein<<<(unsigned)N, 1024, 0, st>>>(us, (int)n, upad, vs, (int)m, mg, cu, d_out);
    ERT_LAUNCH_CHECK("k_wasserstein");
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
