// Development microbenchmark: issue cost of the device RNG pieces at the chain kernels' occupancy
// (one CTA of 512 threads per SM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/rng_bench rng_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ert-conditional-diffusion-model_b200/csrc/denoiser.cuh"
namespace ertdiff { std::string& last_error() { static std::string e; return e; } std::atomic<int64_t> g_launches{0}; }
using namespace ertdiff;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_bench(PhiloxKeys ks, int iters, float* out, long long* cycles) {
    float acc = 0.f;
    uint32_t ctr = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // philox only (2 calls)
            uint4 a = philox4x32_10(make_uint4(ctr, 1, it, 3), ks);
            uint4 b = philox4x32_10(make_uint4(ctr, 1, it, 4), ks);
            acc += __uint_as_float((a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w) & 0x3fffffff);
        } else if (MODE == 1) {     // box-muller only (4 pairs)
            float n[8];
            box_muller(ctr * 2654435761u + it, ctr ^ (it * 40503u), n[0], n[1]);
            box_muller(ctr * 2246822519u + it, ctr ^ (it * 30011u), n[2], n[3]);
            box_muller(ctr * 3266489917u + it, ctr ^ (it * 20011u), n[4], n[5]);
            box_muller(ctr * 668265263u + it, ctr ^ (it * 10007u), n[6], n[7]);
            acc += n[0] + n[1] + n[2] + n[3] + n[4] + n[5] + n[6] + n[7];
        } else {                    // the full 8 normals
            float n[8];
            philox_normal8(ks, 5, ctr, it, 2, n);
            acc += n[0] + n[1] + n[2] + n[3] + n[4] + n[5] + n[6] + n[7];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    PhiloxKeys ks = make_philox_keys(1234);
    const int iters = 2000;
    for (int threads : {128, 256, 512}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k_bench<0><<<148, threads>>>(ks, iters, out, cyc);
                if (mode == 1) k_bench<1><<<148, threads>>>(ks, iters, out, cyc);
                if (mode == 2) k_bench<2><<<148, threads>>>(ks, iters, out, cyc);
                cudaDeviceSynchronize();
            }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const char* names[] = {"philox x2", "box-muller x4", "philox_normal8"};
            printf("threads/SM %4d  %-15s %8.1f cycles per iteration per warp-slot (%.1f per SMSP-warp)\n", threads, names[mode],
                   (double)h / iters, (double)h / iters / (threads / 128.0));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
