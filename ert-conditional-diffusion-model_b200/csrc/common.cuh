// Shared helpers for the ertdiff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/ertdiff_b200.h"

namespace ertdiff {

constexpr int kInChannels = 14;   // ECD.py:134 (hard-coded in the reference)
constexpr int kConv1Out = 32;     // ECD.py:134
constexpr int kConv2Out = 64;     // ECD.py:136
constexpr int kPPad = 32;         // param_dim padded to one warp
constexpr int kNumSMs = 148;      // B200

std::string& last_error();
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}

#define ERT_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            return ::ertdiff::fail(ERTDIFF_ERR_CUDA, std::string(#expr) + ": " +           \
                                                         cudaGetErrorString(_e));          \
        }                                                                                  \
    } while (0)

#define ERT_LAUNCH_CHECK(name)                                                             \
    do {                                                                                   \
        ::ertdiff::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            return ::ertdiff::fail(ERTDIFF_ERR_CUDA,                                       \
                                   std::string("launch ") + name + ": " +                  \
                                       cudaGetErrorString(_e));                            \
        }                                                                                  \
    } while (0)

#define ERT_REQUIRE(cond, msg)                                                             \
    do {                                                                                   \
        if (!(cond)) return ::ertdiff::fail(ERTDIFF_ERR_ARG, std::string(msg));            \
    } while (0)

// conv output length for kernel 3, stride 2, padding 1 (ECD.py:134,136)
__host__ __device__ inline int64_t conv_out_len(int64_t L) { return (L + 2 - 3) / 2 + 1; }

// cudaFuncSetAttribute is per device: every "set the dynamic shared memory limit once" site keeps one
// flag per device (a process may drive several GPUs through several handles)
struct PerDeviceOnce {
    bool done[64] = {};
    bool* slot() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        return &done[dev];
    }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        want = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    int want = -1;
};

}  // namespace ertdiff
