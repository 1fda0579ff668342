// Does scalar FFMA co-issue with packed FFMA2 on sm_100a?  R packed + S scalar independent
// chains per thread; prints the combined fp32 lane-op rate.  (Answer feeds DESIGN.md K1.)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int R, int S>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* out, int iters) {
    float b = in[threadIdx.x & 31], c = in[32 + (threadIdx.x & 31)];
    u64 b2 = pk(b, b + 1.f), c2 = pk(c, c + 1.f);
    u64 p[R > 0 ? R : 1]; float s[S > 0 ? S : 1];
#pragma unroll
    for (int i = 0; i < R; ++i) p[i] = pk(in[64 + i] + threadIdx.x, in[64 + i]);
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = in[80 + i] + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (R > S ? R : S); ++i) {
            if (i < R) p[i] = fma2(p[i], b2, c2);
            if (i < S) s[i] = fma1(s[i], b, c);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) { float x, y; upk(p[i], x, y); r += x + y; }
#pragma unroll
    for (int i = 0; i < S; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int R, int S> void run(int nsm, const float* din, float* dout) {
    const int iters = 100000, blocks = nsm * 4;
    k<R, S><<<blocks, 256>>>(din, dout, 100); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) { CK(cudaEventRecord(e0)); k<R, S><<<blocks, 256>>>(din, dout, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    double ops = (2.0 * R + S) * iters * 256.0 * blocks;
    printf("packed chains %2d scalar chains %2d : %7.2f Tlaneop/s  (%.1f laneop/clk/SM @1.96GHz)\n", R, S, ops / (best * 1e-3) / 1e12, ops / (best * 1e-3) / nsm / 1.96e9);
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); int nsm = pr.multiProcessorCount;
    float h[128]; for (int i = 0; i < 128; ++i) h[i] = 0.5f + 0.001f * i;
    float *din, *dout; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice)); CK(cudaMalloc(&dout, 4 * 256 * nsm * 4));
    run<16, 0>(nsm, din, dout); run<0, 16>(nsm, din, dout); run<8, 8>(nsm, din, dout); run<12, 4>(nsm, din, dout);
    run<8, 4>(nsm, din, dout); run<8, 2>(nsm, din, dout); run<12, 2>(nsm, din, dout); run<8, 16>(nsm, din, dout); run<4, 16>(nsm, din, dout);
    return 0;
}
