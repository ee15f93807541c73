// All-gather over NVLink peer memory, hand-written: every rank owns a buffer that its peers map through CUDA IPC;
// ONE kernel per collective stores this rank's slice straight into every peer's buffer (P2P stores through
// NVLink / NVSwitch), publishes an epoch flag to every peer after a system-scope fence, and waits until every
// peer's flag for this epoch has landed in its own memory.  The path's two collectives (the final fields,
// (B/G, 29) fp32 per rank, and the packed statistics records) are a few KB to a few MB: what matters is latency,
// and this kernel replaces NCCL's launch + protocol (~20 us each on 8 GPUs, measured) by one launch and one
// NVLink round trip.  One process per GPU; the IPC handles travel through torch.distributed once, at set-up.
//
// Buffers are double-buffered by epoch parity.  A rank cannot complete collective e+1 before every rank has
// published e+1, which each does only after it has consumed the data of collective e -- so the buffer of parity
// (e & 1) is never overwritten (by collective e+2) while a slower rank still reads it.
#include <cstring>

#include "common.cuh"

constexpr int kPeerMaxWorld = 16;

struct ertdiff_peer {
    int device = 0, rank = 0, world = 1;
    size_t bytes = 0;                 // capacity of ONE gathered buffer (all ranks' slices)
    char* local = nullptr;            // [2 x bytes data][flags: world x uint32, padded][ticket][status]
    char* peers[kPeerMaxWorld] = {};  // peers[rank] == local
    bool opened[kPeerMaxWorld] = {};
    uint32_t epoch = 0;
};

namespace ertdiff {

struct PeerPtrs { char* p[kPeerMaxWorld]; };

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// grid = a few CTAs; src = this rank's slice (nbytes); data lands at peer[r] + buf_off + rank * slot_bytes
__global__ void __launch_bounds__(256)
k_peer_all_gather(const PeerPtrs peers, int rank, int world, const char* __restrict__ src, size_t nbytes,
                  size_t slot_bytes, size_t buf_off, size_t flags_off, uint32_t epoch, long long timeout_cycles) {
    __shared__ int s_last;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | nbytes | slot_bytes | buf_off) & 15) == 0;
    for (int r = 0; r < world; ++r) {
        const int peer = (rank + r) % world;                  // start with the own copy, then walk the ring
        char* dst = peers.p[peer] + buf_off + (size_t)rank * slot_bytes;
        if (vec) {
            for (size_t i = tid * 16; i < nbytes; i += nthr * 16)
                *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
        } else {
            for (size_t i = tid; i < nbytes; i += nthr) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    uint32_t* flags = reinterpret_cast<uint32_t*>(peers.p[rank] + flags_off);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(flags + kPeerMaxWorld);
    int* status = reinterpret_cast<int*>(flags + kPeerMaxWorld + 1);
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    // the last CTA: publish this rank's flag in every peer's memory, then wait for every peer's flag in ours
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        st_release_sys(reinterpret_cast<uint32_t*>(peers.p[peer] + flags_off) + rank, epoch);
        const long long t0 = clock64();
        // (epochs only grow, and a peer may already be one collective ahead: >=, on the wrapping difference)
        while ((int32_t)(ld_acquire_sys(flags + peer) - epoch) < 0) {
            if (clock64() - t0 > timeout_cycles) { *status = 1; break; }     // a peer never arrived: do not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace ertdiff

using namespace ertdiff;

static size_t peer_flags_off(const ertdiff_peer* p) { return 2 * p->bytes; }
static size_t peer_alloc_bytes(size_t bytes) { return 2 * bytes + (kPeerMaxWorld + 2) * sizeof(uint32_t) + 256; }

#pragma GCC visibility push(default)
extern "C" {

int ertdiff_peer_create(ertdiff_peer** out, int device, int rank, int world, size_t bytes, void* h_ipc_handle64) {
    ERT_REQUIRE(out && h_ipc_handle64, "peer_create: NULL pointer");
    *out = nullptr;
    ERT_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && bytes > 0, "peer_create: bad rank / world (<= 16) / size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
    DeviceGuard g(device);
    if (!g.ok) return fail(ERTDIFF_ERR_CUDA, "peer_create: cudaSetDevice failed");
    auto* p = new ertdiff_peer();
    p->device = device; p->rank = rank; p->world = world;
    p->bytes = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p->local, peer_alloc_bytes(p->bytes));
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, peer_alloc_bytes(p->bytes));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        if (p->local) cudaFree(p->local);
        delete p;
        return fail(ERTDIFF_ERR_CUDA, std::string("peer_create: ") + cudaGetErrorString(e));
    }
    std::memcpy(h_ipc_handle64, &h, 64);
    p->peers[rank] = p->local;
    *out = p;
    return 0;
}

int ertdiff_peer_connect(ertdiff_peer* p, const void* h_all_handles) {
    ERT_REQUIRE(p && h_all_handles, "peer_connect: NULL pointer");
    DeviceGuard g(p->device);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank || p->opened[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(h_all_handles) + 64 * (size_t)r, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(ERTDIFF_ERR_CUDA, std::string("peer_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        p->peers[r] = static_cast<char*>(ptr);
        p->opened[r] = true;
    }
    return 0;
}

int ertdiff_peer_all_gather(ertdiff_peer* p, const void* d_src, size_t nbytes, size_t slot_bytes, void** d_gathered, void* stream) {
    ERT_REQUIRE(p && d_src && d_gathered, "peer_all_gather: NULL pointer");
    ERT_REQUIRE(nbytes > 0 && nbytes <= slot_bytes && slot_bytes * (size_t)p->world <= p->bytes, "peer_all_gather: slice does not fit the buffer");
    for (int r = 0; r < p->world; ++r) ERT_REQUIRE(p->peers[r], "peer_all_gather: peers not connected");
    DeviceGuard g(p->device);
    const uint32_t epoch = ++p->epoch;
    const size_t buf_off = (epoch & 1u) ? p->bytes : 0;
    PeerPtrs pp{};
    for (int r = 0; r < p->world; ++r) pp.p[r] = p->peers[r];
    int blocks = (int)((nbytes + 16 * 1024 - 1) / (16 * 1024));
    blocks = blocks < 1 ? 1 : (blocks > 32 ? 32 : blocks);
    const long long timeout = 4000000000LL;          // ~2 s of SM clocks
    k_peer_all_gather<<<blocks, 256, 0, (cudaStream_t)stream>>>(pp, p->rank, p->world, static_cast<const char*>(d_src), nbytes,
                                                               slot_bytes, buf_off, peer_flags_off(p), epoch, timeout);
    ERT_LAUNCH_CHECK("k_peer_all_gather");
    *d_gathered = p->local + buf_off;
    return 0;
}

int ertdiff_peer_status(ertdiff_peer* p, int* h_status) {
    ERT_REQUIRE(p && h_status, "peer_status: NULL pointer");
    DeviceGuard g(p->device);
    ERT_CUDA(cudaMemcpy(h_status, p->local + peer_flags_off(p) + (kPeerMaxWorld + 1) * sizeof(uint32_t), sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int ertdiff_peer_destroy(ertdiff_peer* p) {
    if (!p) return 0;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->peers[r]);
    cudaFree(p->local);
    delete p;
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
