// rg_tensor.cuh -- tcgen05 / TMEM primitives (inline PTX) and the operand layout of the tensor-core
// contraction used by the organic-view sampler.
//
// Logits of 128 users x TILE_N products are one UMMA tile:  D[u][p] = sum_k A[u][k] * B[p][k]  with
//   A = omega tile   (128 x KT operand elements, K-major), rebuilt by the CTA every round,
//   B = Gamma tile   (TILE_N x KT, K-major), precomputed on the host (GTC array),
// both in the canonical no-swizzle K-major layout of the UMMA shared-memory descriptor, in 16-byte K CHUNKS of EPC
// elements (8 fp16 / 4 tf32):
//   element (row r, column k) -> chunk kc = k / EPC, row group g = r/8:
//   byte offset = kc * (rows/8 * 128) + g * 128 + (r % 8) * 16 + (k % EPC) * sizeof(element)
// so LBO (between the two 16-byte K chunks of one MMA step: K = 16 fp16 / 8 tf32) = 2048 B for A and TILE_N * 16 B for B,
// and SBO (between 8-row groups) = 128 B.
// fp32-grade accuracy comes from a two-term split of both operands, x = hi + lo (11 significant bits each):
//   D = A_hi B_hi + A_hi B_lo + A_lo B_hi   (the dropped lo*lo term is ~2^-22 relative), accumulated in fp32 in TMEM.
// Columns: k < K: log2e-scaled Gamma / omega;  k = K: mu' / 1;  k = K+1: 1 / -shift_u;  rest 0.
#pragma once
#include <cstdint>
#ifndef RG_TC_F16
#define RG_TC_F16 1   // operand element type of the UMMA: 1 = fp16 hi/lo (kind::f16, K = 16 per MMA), 0 = tf32 hi/lo (K = 8)
#endif
#if RG_TC_F16
#include <cuda_fp16.h>
#endif

namespace rg {
namespace tc {

// Operand elements.  Both forms split a value into two 11-significant-bit terms x = hi + lo and contract three of the
// four products (the dropped lo*lo term is ~2^-22 relative), accumulating in fp32.  fp16 elements carry the same
// significand as tf32 in half the bytes: one MMA step covers K = 16 instead of 8, and a B tile is half the L2 -> shared
// memory traffic -- which, not the tensor pipe, is what pass 1 waits for with two CTAs per SM.  The price
// is fp16's exponent range: |x| must stay below 65504 (checked: Gamma', mu' at rg_env_create, omega and the shift when the
// A tile is built), and a `lo` term below 2^-14 is subnormal, i.e. carries an ABSOLUTE error of up to 2^-25 instead of a
// relative one (a term of the error bound: rollout_cta_body, phase A).
#if RG_TC_F16
typedef __half opnd_t;
constexpr int EPC = 8;               // elements per 16-byte K chunk
constexpr float kOpndMax = 65000.f;  // largest magnitude an operand may have
constexpr float kPadLogit = -60000.f; // mu' of the padding products (exactly representable; ex2 of it is 0)
#else
typedef float opnd_t;
constexpr int EPC = 4;
constexpr float kOpndMax = 3.0e38f;
constexpr float kPadLogit = -1e30f;
#endif

constexpr int TILE = 128;            // users per tile (UMMA M)
#ifndef RG_TC_TILE_N
#define RG_TC_TILE_N 128
#endif
constexpr int TILE_N = RG_TC_TILE_N; // products per tile (UMMA N): 64 or 128
constexpr int CHUNK_BYTES = 2048;    // A: one 16-byte K chunk of all 128 rows
constexpr int CHUNK_BYTES_B = TILE_N / 8 * 128;  // B: one 16-byte K chunk of all TILE_N rows

__host__ __device__ inline int kt_chunks(int K) {  // number of 16-byte K chunks (even: one MMA eats two)
    int kc = (K + 2 + EPC - 1) / EPC;
    return (kc + 1) & ~1;
}

// K slabs.  The K dimension (K + 2 columns, padded) is cut into `nslab` slabs of `KCs` 16-byte chunks: one slab of a B
// tile is the unit the TMA ring streams and one group of UMMAs consumes, all slabs of a tile accumulating into the same
// TMEM columns.  Small K (<= 8 chunks, e.g. K = 20 -> 6) is a single slab; larger K uses 6-chunk slabs, so that the ring
// stage stays 24 KB (tf32: K = 64 -> 3 slabs; fp16: one slab up to 10 chunks, K <= 78).  KC = nslab * KCs is the padded
// total the A tile is built with.
__host__ __device__ inline void kt_slabs(int K, int *KC, int *KCs, int *nslab) {
    const int kc0 = kt_chunks(K);
    if (kc0 <= (RG_TC_F16 ? 10 : 8)) { *KCs = kc0; *nslab = 1; }   // fp16: K = 64 is one 40 KB slab
    else { *KCs = 6; *nslab = (kc0 + 5) / 6; }
    *KC = *KCs * *nslab;
}

__host__ __device__ inline float tf32_hi(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u &= 0xFFFFE000u;
    return v.f;
}
#if defined(__CUDACC__)
// x ~= hi + lo in operand elements
__host__ __device__ inline void split_hi_lo(float x, opnd_t &hi, opnd_t &lo) {
#if RG_TC_F16
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
#else
    hi = tf32_hi(x);
    lo = tf32_hi(x - hi);
#endif
}
#endif

// element index of (row, k) of term t inside one operand tile [2 terms][KC][rows/8][8][EPC]
__host__ __device__ inline size_t tile_index(int t, int KC, int row, int k) {  // A: 128 rows
    return ((((size_t)t * KC + (k / EPC)) * 16 + (row >> 3)) * 8 + (row & 7)) * EPC + (k % EPC);
}
__host__ __device__ inline size_t tile_index_b(int t, int KC, int row, int k) {  // B: TILE_N rows
    return ((((size_t)t * KC + (k / EPC)) * (TILE_N / 8) + (row >> 3)) * 8 + (row & 7)) * EPC + (k % EPC);
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D fp32 (bit 4), A/B format (bits 7, 10: 0 = f16, 2 = tf32), both K-major, M = 128, N = TILE_N
// (cute::UMMA::InstrDescriptor)
constexpr uint32_t kOpndFmt = RG_TC_F16 ? 0u : 2u;
constexpr uint32_t kIdescTf32 = (1u << 4) | (kOpndFmt << 7) | (kOpndFmt << 10) | (((uint32_t)TILE_N >> 3) << 17) | ((128u >> 4) << 24);
#if RG_TC_F16
#define RG_UMMA_KIND "kind::f16"
#else
#define RG_UMMA_KIND "kind::tf32"
#endif

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as the alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, one MMA step (two K chunks); issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1." RG_UMMA_KIND " [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    // the wait names every destination register as in/out so that no consumer can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Lean descriptor arithmetic for the MMA issuer (ONE thread issues every UMMA of the kernel: each scalar instruction on
// its path costs ~5 cycles of latency).  Only the address field of a descriptor changes from MMA to MMA: the low word is
// (saddr >> 4) | (LBO >> 4) << 16, the high word (SBO = 128 B, version 1) is the same for every operand tile, and an
// operand offset is an addition to the low word (shared-memory addresses stay below 2^18, so nothing carries into LBO).
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
__host__ __device__ inline uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo) { return (uint64_t)lo | ((uint64_t)kDescHi << 32); }
// tcgen05.mma executed by a CONVERGED warp with warp-uniform operands, issued by the lane(s) whose `issue` is non-zero
// (exactly one: elect_one()).  Keeping the warp converged lets the compiler move the descriptors to uniform registers
// once instead of running a per-lane uniformisation loop (ELECT / R2UR / BRA.U.ANY) around every MMA.
__device__ __forceinline__ void umma_tf32_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate, uint32_t issue) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1." RG_UMMA_KIND " [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {   // 1 in exactly one lane of the (converged) warp
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}
// The three-term contraction of one K slab (KCS chunks) with everything folded at compile time: a_lo / b_lo are the
// descriptor low words of chunk 0 of this slab, "hi" term; a_term16 = 16-byte units between A's hi and lo terms.
// Called by a converged warp; `issue` selects the issuing lane.
template <int KCS>
__device__ __forceinline__ void issue_slab_mmas_lean(uint32_t tmem_d, uint32_t a_lo, uint32_t a_term16, uint32_t b_lo,
                                                     uint32_t accumulate, uint32_t issue) {
    constexpr uint32_t b_term16 = (uint32_t)KCS * CHUNK_BYTES_B / 16u;
    constexpr uint32_t a_step16 = 2u * CHUNK_BYTES / 16u, b_step16 = 2u * CHUNK_BYTES_B / 16u;
    uint32_t acc = accumulate;
    // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll
    for (int pair = 0; pair < 3; ++pair) {
        const uint32_t a0 = a_lo + (pair == 0 ? a_term16 : 0u);
        const uint32_t b0 = b_lo + (pair == 1 ? b_term16 : 0u);
#pragma unroll
        for (int s = 0; s < KCS / 2; ++s) {
            umma_tf32_pred(tmem_d, umma_desc_from_lo(a0 + (uint32_t)s * a_step16),
                           umma_desc_from_lo(b0 + (uint32_t)s * b_step16), kIdescTf32, acc, issue);
            acc = 1u;
        }
    }
}
#endif

// Split form of tmem_ld32 for a software-pipelined epilogue: the load of the NEXT 32 columns is in flight while the
// exponentials of the current ones issue.  tmem_ld32_issue only starts the load; the destination registers must not be
// read before tmem_ld32_wait, which names them as in/out operands so that no consumer is scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// One thread: the three-term tf32 contraction of ONE K SLAB of one 128 x TILE_N tile into TMEM columns
// [tmem_d, tmem_d + TILE_N).  a_saddr: the whole A tile ([2 terms][KC chunks]); b_saddr: this slab of the B tile
// ([2 terms][KCs chunks]).  Slab 0 overwrites the accumulator, the others add to it.
__device__ __forceinline__ void issue_slab_mmas(uint32_t tmem_d, uint32_t a_saddr, int KC, int slab, uint32_t b_saddr,
                                                int KCs) {
    const uint32_t aterm = (uint32_t)KC * CHUNK_BYTES, bterm = (uint32_t)KCs * CHUNK_BYTES_B;  // hi -> lo term
    const uint32_t aoff = (uint32_t)(slab * KCs) * CHUNK_BYTES;
    uint32_t acc = slab > 0 ? 1u : 0u;
    // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll 1
    for (int pair = 0; pair < 3; ++pair) {
        const uint32_t a0 = a_saddr + aoff + (pair == 0 ? aterm : 0u);
        const uint32_t b0 = b_saddr + (pair == 1 ? bterm : 0u);
        for (int s = 0; s < KCs / 2; ++s) {
            umma_tf32(tmem_d, umma_desc(a0 + (uint32_t)s * 2u * CHUNK_BYTES, CHUNK_BYTES),
                      umma_desc(b0 + (uint32_t)s * 2u * CHUNK_BYTES_B, CHUNK_BYTES_B), kIdescTf32, acc);
            acc = 1;
        }
    }
}
#endif

}  // namespace tc
}  // namespace rg
