// Single-thread latencies of the operations on the ordered-response hand-off path (L2-resident data):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_microbench tools/lat_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chase_ld32(const uint32_t* next, uint32_t start, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) i = __ldcg(next + (size_t)i * 8);
    out[0] = clock64() - t0;
    out[1] = i;
}
__global__ void chase_ld256(const uint32_t* next, uint32_t start, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
        asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                     : "l"(next + (size_t)i * 8));
        i = r0 + (r7 >> 31);
    }
    out[0] = clock64() - t0;
    out[1] = i;
}
__global__ void chase_atom(uint32_t* arr, uint32_t start, uint32_t mask, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        uint32_t v = atomicAdd(arr + (size_t)i * 8 + 1, 1u);
        i = (i * 1664525u + 1013904223u + (v >> 31)) & mask;   // v < 2^31: a real dependency the compiler cannot fold
    }
    out[0] = clock64() - t0;
    out[1] = i;
}
__global__ void chase_st_ld(uint32_t* arr, uint32_t start, uint32_t mask, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        arr[(size_t)i * 8 + 2] = k;
        uint32_t v = __ldcg(arr + (size_t)i * 8 + 2);      // read back own store from L2
        i = (i * 1664525u + 1013904223u + (v >> 31)) & mask;
    }
    out[0] = clock64() - t0;
    out[1] = i;
}
__global__ void chase_fence(uint32_t* arr, uint32_t start, uint32_t mask, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        arr[(size_t)i * 8 + 3] = k;
        __threadfence();
        i = (i * 1664525u + 1013904223u) & mask;
    }
    out[0] = clock64() - t0;
    out[1] = i;
}
// the hand-off bundle: 2 stores, then 2 atomics + 6 loads in flight together, wait for all
__global__ void bundle(uint32_t* arr, uint32_t start, uint32_t mask, int n, long long* out) {
    uint32_t i = start;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        uint32_t a = (i * 1664525u + 1013904223u) & mask, b = (a * 1664525u + 1013904223u) & mask;
        arr[(size_t)a * 8 + 4] = k;
        arr[(size_t)b * 8 + 4] = k;
        uint32_t r1 = atomicAdd(arr + (size_t)a * 8 + 1, 1u), r2 = atomicAdd(arr + (size_t)b * 8 + 1, 1u);
        uint32_t s = 0;
        for (int q = 0; q < 6; ++q) {
            uint32_t c = (b * (2u * q + 3u) + 12345u) & mask;
            s += __ldcg(arr + (size_t)c * 8);
        }
        i = (b + ((r1 + r2 + s) >> 31)) & mask;
    }
    out[0] = clock64() - t0;
    out[1] = i;
}

int main() {
    const uint32_t n = 1u << 20;   // 32 MB of 32-byte entries: L2 resident
    uint32_t* h = new uint32_t[(size_t)n * 8]();
    uint32_t x = 1;
    for (uint32_t k = 0; k < n; ++k) {   // random cycle-ish chase
        x = x * 1664525u + 1013904223u;
        h[(size_t)k * 8] = (k * 2654435761u + 12345u) & (n - 1);
    }
    uint32_t* d;
    long long* o;
    cudaMalloc(&d, (size_t)n * 32);
    cudaMemcpy(d, h, (size_t)n * 32, cudaMemcpyHostToDevice);
    cudaMallocManaged(&o, 16);
    const int iters = 20000;
    auto report = [&](const char* name) {
        cudaDeviceSynchronize();
        printf("%-34s %8.1f cycles per dependent op\n", name, (double)o[0] / iters);
    };
    for (int rep = 0; rep < 2; ++rep) {
        chase_ld32<<<1, 1>>>(d, 1, iters, o); report("ld.cg 32-bit chase");
        chase_ld256<<<1, 1>>>(d, 1, iters, o); report("ld.cg 256-bit chase");
        chase_atom<<<1, 1>>>(d, 1, n - 1, iters, o); report("atomicAdd (returning) chase");
        chase_st_ld<<<1, 1>>>(d, 1, n - 1, iters, o); report("store + ld.cg read-back");
        chase_fence<<<1, 1>>>(d, 1, n - 1, iters, o); report("store + __threadfence");
        bundle<<<1, 1>>>(d, 1, n - 1, iters, o); report("2 st, 2 atom + 6 ld together");
    }
    return 0;
}
