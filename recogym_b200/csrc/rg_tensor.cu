// rg_tensor.cu -- tensor-core (tcgen05) contraction of the organic-view logits.
//
// tc_logits_kernel is the stand-alone form used by the tests: one CTA computes the shifted log2-logits
//   D[u][p] = log2e * (Gamma_p . omega_u + mu_p) - shift_u      for 128 users x all products
// with UMMA tiles (operands and layout: rg_tensor.cuh) and writes them to global memory, so that the
// descriptor encodings, the tf32 hi/lo split and the TMEM read-back can be checked against binary64.
#include "rg_host.h"
#include "rg_tensor.cuh"

namespace rg {
namespace tc {

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// Writes the A operand (omega tile) of 128 users into shared memory: hi and lo tf32 terms.
// omega: [128][K] binary64 (row stride ldo), shift: [128] fp32 (log2 units).  All threads of the CTA.
__device__ void build_a_tile(opnd_t *a_tile, int KC, int K, const double *omega, int ldo, const float *shift,
                             int nrows, int tid, int nthreads) {
    const int KT = KC * EPC;
    for (int idx = tid; idx < TILE * KT; idx += nthreads) {
        const int row = idx / KT, k = idx - row * KT;
        float x = 0.f;
        if (row < nrows) {
            if (k < K) x = (float)omega[(size_t)row * ldo + k];
            else if (k == K) x = 1.f;
            else if (k == K + 1) x = -shift[row];
        }
        split_hi_lo(x, a_tile[tile_index(0, KC, row, k)], a_tile[tile_index(1, KC, row, k)]);
    }
}

__global__ void __launch_bounds__(128) tc_logits_kernel(const opnd_t *__restrict__ GTC, int nsub, int KC, int KCs,
                                                        int nslab, int K,
                                                        const double *__restrict__ omega,
                                                        const float *__restrict__ shift, float *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned a_bytes = 2u * (unsigned)KC * CHUNK_BYTES, tile_bytes = 2u * (unsigned)KCs * CHUNK_BYTES_B;  // one slab
    opnd_t *a_tile = (opnd_t *)smem;
    opnd_t *b_tile = (opnd_t *)(smem + a_bytes);
    uint64_t *bars = (uint64_t *)(smem + a_bytes + tile_bytes);  // [0] TMA, [1] MMA
    uint32_t *tmem_slot = (uint32_t *)(bars + 2);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, TILE_N);
    build_a_tile(a_tile, KC, K, omega, K, shift, TILE, tid, 128);
    fence_proxy_async();  // generic-proxy writes of the A tile -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    unsigned unit = 0;
    for (int sub = 0; sub < nsub * (128 / TILE_N); ++sub) {  // TILE_N-product tiles, each in nslab K slabs
        for (int sl = 0; sl < nslab; ++sl, ++unit) {
            const unsigned parity = unit & 1u;
            if (tid == 0) {
                mbar_expect_tx(&bars[0], tile_bytes);
                bulk_g2s(b_tile, (const unsigned char *)GTC + (size_t)unit * tile_bytes, tile_bytes, &bars[0]);
            }
            mbar_wait(&bars[0], parity);
            if (tid == 0) {
                tc_fence_after();
                issue_slab_mmas(tmem_base, smem_addr(a_tile), KC, sl, smem_addr(b_tile), KCs);
                umma_commit(&bars[1]);
            }
            mbar_wait(&bars[1], parity);   // the slab buffer is free again (and, after the last slab, D is complete)
        }
        tc_fence_after();
        const int row = warp * 32 + lane;  // TMEM lane == user row
#pragma unroll 1
        for (int c0 = 0; c0 < TILE_N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) out[(size_t)row * ((size_t)nsub * TILE) + (size_t)sub * TILE_N + c0 + i] = v[i];
        }
        tc_fence_before();
        __syncthreads();  // TMEM tile and B tile are free again
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem_base, TILE_N);
}

}  // namespace tc

cudaError_t launch_tc_logits(const rg_env_impl &e, const double *omega, const float *shift, float *out,
                             cudaStream_t stream) {
    using namespace tc;
    const int KC = e.dev.KC, KCs = e.dev.KCs, nslab = e.dev.nslab;
    const size_t smem = 2 * (size_t)KC * CHUNK_BYTES + 2 * (size_t)KCs * CHUNK_BYTES_B + 64;
    cudaError_t ce = cudaFuncSetAttribute(tc_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    tc_logits_kernel<<<1, 128, smem, stream>>>(e.d_GTC, e.dev.nsub, KC, KCs, nslab, e.dev.K, omega, shift, out);
    return cudaGetLastError();
}

}  // namespace rg
