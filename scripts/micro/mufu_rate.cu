// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition on this GPU, with 1, 2, 4 and 8 warps per sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int WITH_FADD>
__global__ void k(float *out, long long *cyc, int iters) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = -1.0f - 0.001f * (float)(threadIdx.x + i);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ex2(v[i] * 0.5f - 1.0f);   // FMUL+FADD feed (FMA pipe) + one MUFU
        if (WITH_FADD) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
        }
    }
    const long long t1 = clock64();
    float s = s0 + s1 + s2 + s3;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    const int iters = 2000;
    for (int fadd = 0; fadd < 2; ++fadd)
        for (int warps = 4; warps <= 32; warps *= 2) {   // warps per CTA, one CTA per SM: warps/4 per sub-partition
            if (fadd) k<1><<<148, warps * 32>>>(out, cyc, iters); else k<0><<<148, warps * 32>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
            const double mufu_per_subpart = (double)iters * 32 * (warps / 4);
            printf("fadd=%d warps/subpartition=%d: %.2f cycles per warp-MUFU per sub-partition (%.1f ex2/clk/SM)\n", fadd,
                   warps / 4, avg / mufu_per_subpart, 32.0 * 4 * mufu_per_subpart / avg);
        }
    return 0;
}
