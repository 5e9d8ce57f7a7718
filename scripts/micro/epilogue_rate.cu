// Micro-benchmark of the tensor-mode epilogue in isolation: per 128-column tile each of 4 warps does 4 x (tcgen05.ld
// 32x32b.x32 + wait, 32 MUFU.EX2, 32 FADD).  Variants: 0 = ld + ex2 + add, 1 = ex2 + add only, 2 = ld only,
// 3 = ld + add (no MUFU).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../recogym_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "rg_tensor.cuh"
using namespace rg::tc;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int VAR>
__global__ void __launch_bounds__(128) k(float *out, long long *cyc, int tiles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    float part = 0.f;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = -1.0f - 0.01f * (float)(lane + i);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
            if (VAR != 1) tmem_ld32(base + (uint32_t)((t & 1) * 128 + c), v);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (VAR == 0 || VAR == 1) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float a = VAR == 1 ? v[i] * 0.999f : v[i], b = VAR == 1 ? v[i + 1] * 0.999f : v[i + 1];
                    s0 += ex2(a); s1 += ex2(b); s2 += ex2(v[i + 2]); s3 += ex2(v[i + 3]);
                }
            } else if (VAR == 3) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
            } else {
                s0 = v[0] + v[31];
            }
            part += (s0 + s1) + (s2 + s3);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = part;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 256);
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 128);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    const int tiles = 2000;
    const char *names[4] = {"ld + ex2 + add", "ex2 + add", "ld only", "ld + add"};
    for (int var = 0; var < 4; ++var) {
        if (var == 0) k<0><<<148, 128>>>(out, cyc, tiles);
        if (var == 1) k<1><<<148, 128>>>(out, cyc, tiles);
        if (var == 2) k<2><<<148, 128>>>(out, cyc, tiles);
        if (var == 3) k<3><<<148, 128>>>(out, cyc, tiles);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        printf("%-16s: %.0f cycles per 128-column tile per warp (%s)\n", names[var], avg / tiles, cudaGetErrorString(e));
    }
    return 0;
}
