// Micro-benchmark: cost of a CHAIN of tcgen05.mma (M = 128, operands in shared memory, K-major no-swizzle, accumulators
// in TMEM) as the rollout's pass 1 issues them -- one thread, `chain` MMAs per accumulator tile, commit, mbarrier wait.
// Question: is a tile's chain bound by the tensor pipe's throughput or by the dependent accumulation into ONE
// accumulator (then two tiles interleaved cost the same as one), and what does kind::f16 (K = 16 per instruction) buy
// over kind::tf32 (K = 8)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../recogym_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rg_tensor.cuh"
using namespace rg::tc;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// lean issue, as in the rollout kernel: converged warp, one elected lane, descriptors = constant low word + compile-time
// offsets, kind and shape fixed at compile time
template <int F16>
__device__ __forceinline__ void mma_pred(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, uint32_t issue) {
    if (F16)
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(issue)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(issue)
                     : "memory");
}

// CHAIN MMAs into each of NTILE accumulators (round-robin over the tiles inside a step), then ONE commit + wait
template <int F16, int N, int CHAIN, int NTILE>
__global__ void k(long long *cyc, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lo = umma_desc_lo(s32(smem), 2048u), b_lo = umma_desc_lo(s32(smem + 32768), (uint32_t)(N / 8 * 128));
        const uint32_t el = elect_one();
        const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
        unsigned ph = 0;
        long long tot = 0, iss = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int m = 0; m < CHAIN; ++m)
#pragma unroll
                for (int t = 0; t < NTILE; ++t)
                    mma_pred<F16>(tmu + (uint32_t)(t * N), umma_desc_from_lo(a_lo + (uint32_t)((m % 4) * 4096 / 16)),
                                  umma_desc_from_lo(b_lo + (uint32_t)((m % 4) * 4096 / 16)), idesc, m > 0 ? 1u : 0u, el);
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(s32(&bar)), "r"(el) : "memory");
            const long long t1 = clock64();
            mbar_wait(s32(&bar), ph);
            ph ^= 1u;
            const long long t2 = clock64();
            if (r > 0) { tot += t2 - t0; iss += t1 - t0; }
        }
        if (threadIdx.x == 0) { cyc[blockIdx.x * 4 + 0] = tot / (reps - 1); cyc[blockIdx.x * 4 + 2] = iss / (reps - 1); }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int F16, int N, int CHAIN, int NTILE>
void run(const char *name, long long *cyc) {
    cudaFuncSetAttribute(k<F16, N, CHAIN, NTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<F16, N, CHAIN, NTILE><<<148, 128, 65536>>>(cyc, 50);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, iss = 0;
    for (int i = 0; i < 148; ++i) { avg += (double)h[i * 4]; iss += (double)h[i * 4 + 2]; }
    printf("%-40s: %7.0f cycles per batch (issue %5.0f) = %6.1f per MMA  (%s)\n", name, avg / 148, iss / 148, avg / 148 / (CHAIN * NTILE),
           cudaGetErrorString(e));
}

int main() {
    long long *cyc;
    cudaMalloc(&cyc, sizeof(long long) * 148 * 4);
    run<0, 128, 9, 1>("tf32 N=128 chain 9, 1 tile", cyc);
    run<0, 128, 18, 1>("tf32 N=128 chain 18, 1 tile", cyc);
    run<0, 128, 36, 1>("tf32 N=128 chain 36, 1 tile", cyc);
    run<0, 128, 9, 2>("tf32 N=128 chain 9, 2 tiles interleaved", cyc);
    run<0, 128, 9, 4>("tf32 N=128 chain 9, 4 tiles interleaved", cyc);
    run<0, 64, 9, 1>("tf32 N=64 chain 9, 1 tile", cyc);
    run<0, 64, 36, 1>("tf32 N=64 chain 36, 1 tile", cyc);
    run<0, 64, 9, 4>("tf32 N=64 chain 9, 4 tiles interleaved", cyc);
    run<0, 256, 9, 1>("tf32 N=256 chain 9, 1 tile", cyc);
    run<0, 256, 36, 1>("tf32 N=256 chain 36, 1 tile", cyc);
    run<1, 128, 6, 1>("f16 N=128 chain 6, 1 tile", cyc);
    run<1, 128, 36, 1>("f16 N=128 chain 36, 1 tile", cyc);
    run<1, 128, 6, 2>("f16 N=128 chain 6, 2 tiles interleaved", cyc);
    run<1, 256, 36, 1>("f16 N=256 chain 36, 1 tile", cyc);
    return 0;
}
