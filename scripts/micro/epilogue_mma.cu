// Micro-benchmark: does a concurrent stream of tcgen05.mma (accumulators in TMEM) slow the epilogue's TMEM reads / MUFU
// work down?  Epilogue = per 128-column tile and lane quarter: tcgen05.ld 32x32b.x32 + wait, 32 MUFU.EX2 + 32 FADD per
// 32-column chunk, done by NE warps per quarter (NE = 1: one warp reads all 128 columns; NE = 2: two warps, 64 columns
// each).  MMA stream = one extra warp issuing `mmas_per_tile`-deep batches of 128x128x8 tf32 UMMAs (zero operands in
// shared memory) into the accumulator buffer the epilogue is NOT reading, free-running (commit + wait per batch).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../recogym_b200/csrc -o epilogue_mma epilogue_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rg_tensor.cuh"
using namespace rg::tc;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// MODE: 0 ld + ex2 + add, 1 ld + add (no MUFU), 2 ex2 + add on registers (no TMEM read)
template <int NE, int MODE>
__global__ void k(float *out, long long *cyc, int tiles, int mmas_per_tile, int *stop_flag) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nepi = 4 * NE;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;   // A (8 KB: K = 16) + B (8 KB)
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        done = 0;
    }
    if (warp == 0) tmem_alloc(&slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp < nepi) {
        const int q = warp & 3, half = warp >> 2;
        const uint32_t base = tm + ((uint32_t)(q * 32) << 16);
        constexpr int NCOL = 128 / NE;
        float part = 0.f;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = -1.0f - 0.01f * (float)(lane + i);
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
            for (int c = 0; c < NCOL; c += 32) {
                if (MODE != 2) tmem_ld32(base + (uint32_t)((t & 1) * 128 + half * NCOL + c), v);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                if (MODE == 0 || MODE == 2) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (MODE == 2) { v[i] = v[i] * 0.999f; v[i + 1] *= 0.999f; v[i + 2] *= 0.999f; v[i + 3] *= 0.999f; }
                        s0 += ex2(v[i]); s1 += ex2(v[i + 1]); s2 += ex2(v[i + 2]); s3 += ex2(v[i + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
                }
                part += (s0 + s1) + (s2 + s3);
            }
        }
        const long long t1 = clock64();
        out[blockIdx.x * 512 + threadIdx.x] = part;
        if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; done = 1; }
    } else if (warp == nepi && mmas_per_tile > 0) {
        // free-running MMA stream into the two accumulator buffers alternately
        const uint32_t a_s = s32(smem), b_s = s32(smem + 8192);
        const uint64_t da = umma_desc(a_s, CHUNK_BYTES), db = umma_desc(b_s, CHUNK_BYTES_B);
        unsigned ph = 0;
        long long n = 0;
        while (!done) {
            if (lane == 0) {
                for (int m = 0; m < mmas_per_tile; ++m) umma_tf32(tm + (uint32_t)((n & 1) * 128), da, db, kIdescTf32, m > 0 ? 1u : 0u);
                umma_commit(&bar);
                mbar_wait(s32(&bar), ph);
            }
            ph ^= 1u;
            ++n;
            __syncwarp();
        }
        if (lane == 0) cyc[148 + blockIdx.x] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

template <int NE, int MODE>
void run(const char *name, int mmas, float *out, long long *cyc) {
    const int tiles = 2000;
    cudaMemset(cyc, 0, sizeof(long long) * 296);
    cudaFuncSetAttribute(k<NE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k<NE, MODE><<<148, (4 * NE + 1) * 32, 16384>>>(out, cyc, tiles, mmas, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, nb = 0;
    for (int i = 0; i < 148; ++i) { avg += (double)h[i]; nb += (double)h[148 + i]; }
    avg /= 148; nb /= 148;
    printf("%-34s NE=%d mma batch=%d: %7.0f cycles per 128-column tile per quarter; MMA batches per tile %.2f (%s)\n", name, NE, mmas,
           avg / tiles, nb / tiles, cudaGetErrorString(e));
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 512);
    cudaMalloc(&cyc, sizeof(long long) * 296);
    for (int mmas : {0, 9, 3}) {
        run<1, 0>("ld + ex2 + add", mmas, out, cyc);
        run<2, 0>("ld + ex2 + add", mmas, out, cyc);
        run<1, 1>("ld + add", mmas, out, cyc);
        run<2, 1>("ld + add", mmas, out, cyc);
        run<1, 2>("ex2 + add (registers)", mmas, out, cyc);
        run<2, 2>("ex2 + add (registers)", mmas, out, cyc);
    }
    return 0;
}
