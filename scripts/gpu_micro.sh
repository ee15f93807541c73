cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/rng_bench rng_bench.cu && timeout 120 /tmp/rng_bench
