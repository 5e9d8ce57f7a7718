// rg_rollout_fast.cu -- RG_MATH_FAST / RG_MATH_TENSOR: the CTA-batched fused rollout kernel.
//
// A persistent CTA owns TU=128 user slots.  Work proceeds in rounds; at the start of a round every
// live slot is in the ORGANIC state, so the expensive part of the model -- the product-view draw
// from Categorical(softmax(Gamma omega + mu)) (reco_env_v1.py:119-128) -- is one dense
// [128 users] x [P products] x [K] contraction per round:
//
//   pass 1  CUDA cores (RG_MATH_FAST): Gamma tiles (128 products x (K+1) floats, pre-scaled by log2 e, mu as row K)
//           stream L2 -> shared memory through a ring of TMA bulk copies (cp.async.bulk + mbarrier); each thread owns
//           an 8-user x 8-product register tile.  Tensor cores (RG_MATH_TENSOR): the contraction is a tcgen05 UMMA
//           (tf32 hi/lo split, three terms, K slabs for large K) with accumulators in TMEM; warp 4 issues the MMAs,
//           warp 5 refills the B ring, warps 0-3 run the epilogue (thread = TMEM lane = user).  Either way the epilogue
//           turns logits into masses with a bare ex2 (shifted by a per-user upper bound of the logits, so no running
//           maximum is needed) and keeps per-user partial sums per "record tile" in shared memory (tensor mode: also
//           per 32-product group in an L2-resident scratch).
//   pass 2  total mass, tau = u * Z, locate the record tile (thread per user), then recompute only the crossing tile
//           (CUDA cores) / the crossing 32-product group (tensor mode, four users interleaved per warp) in natural
//           product order with a warp scan -> candidate index + its CDF margins.
//   filter  the margins are compared with a rigorous bound of the fp32 error in which the errors common to the prefix
//           sum and the total cancel (DESIGN.md "three-tier sampler"); a draw that could be affected waits in a queue
//           until 8 of them share one re-sampling sweep: tier 2 (binary64 logits on DMMA, fp32 ex2, a-posteriori
//           bound), tier 3 (all binary64).  Every discrete result therefore equals RG_MATH_EXACT.
//   advance D1 (thread per slot): bookkeeping of the view, transition; D2 (warp per user): the whole run of cheap
//           bandit events (policy action, click ~ Bernoulli(ff(beta_a.omega + mu_a)), transition), 32 candidate events
//           per turn; stopped users are written out and the slot claims the next user from a global ticket queue
//           (longest-first schedule for large launches); omega updates are batched over the CTA.
//
// All draws are keyed by (seed, user, event time, slot): the result is independent of scheduling.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rg_device.cuh"
#include "rg_host.h"
#include "rg_tensor.cuh"

namespace rg {
namespace fast {

// RG_PROF (experiment builds only, `make prof`): cycle counters of the pass-1 hand-shakes and a time line of one round
#ifdef RG_PROF
__device__ unsigned long long g_prof[64];
__device__ long long g_trace[8192];
#define RG_PROF_ONLY(...) __VA_ARGS__
#else
#define RG_PROF_ONLY(...)
#endif

constexpr int TU = 128;     // user slots per CTA
constexpr int SUB = 128;    // products per shared-memory tile
constexpr int NT = 256;     // threads per CTA
constexpr int CHUNK = 256;  // products per partial sum of the cooperative binary64 sampler
constexpr int kMaxStages = 6;
constexpr int NBUF = tc::TILE_N == 64 ? 3 : 2;  // tensor mode: accumulator tiles (TILE_N TMEM columns each) in flight
constexpr int FB = 8;       // users per sweep of the cooperative binary64 sampler

struct Layout {  // byte offsets into dynamic shared memory
    int clock, dt,   // NormalTimeGenerator only: per-slot clock value and the omega_k of the slot's pending drift
        mbar, om64, csum, user, hash, t, state, last_view, n_org, n_ban, n_clk, view, live, pos2slot, shift, eps,
        nview, amb_list, fb_slot, fb_pos, fb_unres, fb_u, fb_lse, fb_om, lse, pz, pu, rq_u, rq_lse, ptau, pcum, prsel, wait, rq_slot, q3_slot, q3_u, pend, pend_list, misc, om32, tilesum, gtile, atile,
        total;
};

struct Params {
    EnvDev env;
    PolicyDev pol;
    long long user_begin, n_organic_users, n_total;
    RolloutOutDev out;
    unsigned long long *queue;
    const int32_t *perm;    // longest-first schedule: ticket q -> user index perm[q] (nullptr: identity)
    int nstage, rec_sub, nrec, nrec_pad, kp, ncs;
    int tc;                 // 1: tensor-core pass 1 (RG_MATH_TENSOR)
    int nhalf;              // partial sums per record tile and user: 2 in tensor mode (one per column half), else 1
    double *om_scratch;     // tc (or spill >= 2): the binary64 omega rows live in global memory, [grid][TU][kp]
    double *csum_scratch;   // spill >= 1: partial sums of the re-sampling sweeps in global memory, [grid][ncs * FB * 2]
    int spill;              // CUDA-core form, large P x K: 0 all in shared memory, 1 csum global, 2 csum + omega global
    float *gsum;            // tc: pass-1 mass sums per 32-product group, [grid][nsub * 4][TU] (L2-resident scratch)
    Layout L;
};

__host__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ inline void make_layout(Params &p) {
    const int K = p.env.K, K1 = K + 1;
    Layout &L = p.L;
    int o = 0;
    int KC = 0, KCs = 0, nslab = 0;
    tc::kt_slabs(K, &KC, &KCs, &nslab);
    const int tc_tile_bytes = 2 * KCs * tc::CHUNK_BYTES_B;  // one K slab of a B tile (hi + lo): the ring unit
    const int tc_a_bytes = 2 * KC * tc::CHUNK_BYTES;       // the 128-user A tile (hi + lo)
    L.mbar = o; o += 8 * (kMaxStages + 2 * NBUF + 1);  // ring full, NBUF TMEM-full, NBUF TMEM-empty, TMEM address
    L.om64 = o; if (!p.tc && p.spill < 2) o += 8 * TU * p.kp;
    L.csum = o; if (!p.tc && p.spill < 1) o += 8 * p.ncs * FB * 2;
    L.fb_u = o; o += 8 * FB;
    L.fb_lse = o; o += 8 * FB;
    L.fb_om = o; o += 8 * FB * p.kp;
    L.lse = o; if (!p.tc) o += 8 * TU;
    L.pz = o; if (!p.tc) o += 8 * TU;
    L.pu = o; if (!p.tc) o += 8 * TU;
    L.rq_u = o; o += 8 * TU;
    L.rq_lse = o; o += 8 * TU;
    L.q3_u = o; o += 8 * TU;
    L.clock = o; if (p.env.time_mode) o += 8 * TU;
    L.dt = o; if (p.env.time_mode) o += 8 * TU;
    L.ptau = o; if (!p.tc) o += 8 * TU;
    L.pcum = o; if (!p.tc) o += 8 * TU;
    L.user = o; o += 8 * TU;
    L.hash = o; o += 8 * TU;
    L.t = o; o += 4 * TU;
    L.state = o; o += 4 * TU;
    L.last_view = o; o += 4 * TU;
    L.n_org = o; o += 4 * TU;
    L.n_ban = o; o += 4 * TU;
    L.n_clk = o; o += 4 * TU;
    L.view = o; o += 4 * TU;
    L.live = o; o += 4 * TU;
    L.pos2slot = o; o += 4 * TU;
    L.shift = o; o += 4 * TU;
    L.eps = o; o += 4 * TU;
    L.nview = o; o += 4 * TU;
    L.amb_list = o; o += 4 * TU;
    L.fb_slot = o; o += 4 * FB;
    L.fb_pos = o; o += 4 * FB;
    L.fb_unres = o; o += 4 * FB;
    L.prsel = o; o += 4 * TU;
    L.wait = o; o += 4 * TU;
    L.rq_slot = o; o += 4 * TU;
    L.q3_slot = o; o += 4 * TU;
    L.pend = o; o += 4 * TU;
    L.pend_list = o; o += 4 * TU;
    L.misc = o; o += 64;
    o = align_up(o, 16);
    L.om32 = o; if (!p.tc) o += 4 * K1 * TU;
    L.tilesum = o; o += 4 * TU * p.nrec_pad;
    o = align_up(o, 128);
    L.gtile = o;
    if (!p.tc) {
        o += 4 * p.nstage * K1 * SUB;
        L.atile = o;
    } else {
        // the B ring is idle between pass 1 and the next round's first load: the fp32 omega tile of pass 2 and
        // the partial sums of the re-sampling sweeps live in the same bytes
        int ring = p.nstage * tc_tile_bytes;
        const int om32_bytes = align_up(4 * K1 * TU, 16);
        const int alias = om32_bytes + 5 * 8 * TU + 8 * p.ncs * FB * 2;
        if (alias > ring) ring = align_up(alias, 128);
        L.om32 = o;
        L.lse = o + om32_bytes;
        L.pz = L.lse + 8 * TU; L.pu = L.pz + 8 * TU; L.ptau = L.pu + 8 * TU; L.pcum = L.ptau + 8 * TU;
        L.csum = L.pcum + 8 * TU;
        o += ring;
        L.atile = o; o += tc_a_bytes;
    }
    L.total = o;
}

// ---- mbarrier / TMA bulk-copy primitives (PTX) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waiting threads are SUSPENDED by the hardware up to this many ns per try_wait instead of re-issuing the probe: the
// issuer / loader warps share their scheduler with an epilogue warp, and their polling cost that warp ~10 % of its issue
// slots (r02 profile: epilogue warp 0 took 2200 cycles per tile where warp 3 took 2050, and the slowest warp paces the tile).
#ifndef RG_SUSPEND_NS
#define RG_SUSPEND_NS 1000000
#endif
constexpr unsigned kSuspendHint = RG_SUSPEND_NS;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(kSuspendHint)
        : "memory");
}
// the same on a raw shared-memory address (the single-thread issuer / loader loops keep their barrier addresses in registers)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar_saddr),
        "r"(parity), "r"(kSuspendHint)
        : "memory");
}
// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test_s(uint32_t bar_saddr, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_saddr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_commit_s(uint32_t bar_saddr, uint32_t issue) {   // converged warp, one issuing lane
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_saddr), "r"(issue)
        : "memory");
}
// Barrier of the NT compute threads (named barrier 1).  The tensor-mode kernel carries a ninth warp -- the MMA issuer --
// that never takes part in it; barrier 2 (NT + 32 threads) releases that warp into a round's pass 1.
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void issuer_sync() { asm volatile("bar.sync 2, 288;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (One CTA per SM -- no co-resident CTA at all -- still needs 2280 cycles per 128 x 128 tile in pass 1 (two CTAs: 3035 each),
// against 1100 of MUFU and 610 of tensor time: pass 1 is a latency chain, not a throughput limit.  Re-measured after the
// TMA loader thread was decoupled: software-pipelined tcgen05.ld and the exponentials-first ordering still change nothing.)
// (Measured and dropped, all on 1xB200 at C3 with 5.9e8 user-steps/s as the baseline: software-pipelining the tcgen05.ld
// of the next 32 columns under the exponentials of the current ones (5.86e8); issuing the 32 exponentials of a chunk
// before any addition (5.90e8); 64-product UMMA tiles with 4-6 ring stages (4.4e8); one pass-1 token per SM so that the
// two co-resident CTAs alternate on the MUFU pipe (5.56e8 vs 5.82e8 in the same run).  The epilogue warp issues one
// MUFU.EX2 per ~14 cycles whether or not the other CTA is in pass 1.)
// (Measured and dropped: sending 3 of every 8 exponentials of the tensor-core epilogue through a degree-5 polynomial
// on the FMA pipe -- max relative error 2.2e-7, inside the filter's allowance -- takes MUFU time off pass 1 but costs 10
// issue slots per exponential; with two co-resident CTAs sharing the schedulers the kernel got 10 % slower, 4.6e8 vs
// 5.1e8 user-steps/s at C3.)

// (Measured and dropped in round 2, all parity-green on B200 at C3: issuing the Gamma loads of all five k steps of a sweep
// step before its first DMMA, with the epilogue's mu values prefetched -- 43.0k vs 37.8k cycles per round: the sweep is bound
// by the FP64 pipe, 6250 DMMA per sweep, not by its load latency; a dedicated tenth warp that ran the tier 2/3 sweeps
// asynchronously, off the round's critical path, with a ring of queued draws -- one scheduler then carries all the FP64
// work and starves its two epilogue warps: 5.2e8 vs 7.0e8 user-steps/s; gathering the tail branch of the normal quantile
// onto a few lanes of each warp in the omega phase -- 22.6k vs 18.2k cycles per round.)
// ---- scalar (one thread per user) pieces --------------------------------------------------------
// omega init / drift by ONE thread: same operation sequence as warp_omega_update.
__device__ __forceinline__ void thread_omega_update(double *omega, int K, double scale, bool init, uint64_t seed,
                                                    uint64_t user, uint32_t t) {
    for (int j = 0; 2 * j < K; ++j) {
        const u32x4 w = draw(seed, user, t, SLOT_NORMAL + (uint32_t)j);
        const double z0 = normal_from_words(w.x, w.y);
        const double b0 = init ? 0.0 : omega[2 * j];
        omega[2 * j] = b0 + scale * z0;
        if (2 * j + 1 < K) {
            const double z1 = normal_from_words(w.z, w.w);
            const double b1 = init ? 0.0 : omega[2 * j + 1];
            omega[2 * j + 1] = b1 + scale * z1;
        }
    }
}

// Re-sampling of the draws the fp32 filter could not certify: up to FB users per sweep of Gamma, whole CTA.
// Every Gamma column (binary64, K-major) is loaded once from L2 and used for all users of the batch; partial
// sums per CHUNK products go to csum[f][c]; the inverse CDF of user f is then located by warp f.
//
//   TIER 2 (T2 = true): binary64 logits (same fma chains as the oracle), but the mass is a bare
//     ex2.approx of the logit relative to the user's log-sum-exp `lse` known from the fp32 pass:
//     m = ex2(float(x)), x = log2e*(l) - lse <= ~0.  Its relative error is at most
//     ln2*2^-24*|x| (rounding x to fp32) + 2^-21 (ex2.approx), so the error of ANY prefix sum is at most
//     E = sum_p m_p*(c1*|x_p| + c2), accumulated alongside in esum[f][c] (an a-posteriori bound, about
//     6e-7 of the total: ~250x tighter than the fp32 pass).  A draw whose margins exceed 2E is final.
//   TIER 3 (T2 = false): everything in binary64 (exp64), the reference arithmetic; final.
template <int NB, int PP, bool T2>
__device__ __forceinline__ void batch_partial_sums(const EnvDev &env, const double *om_s, int KP, double *csum,
                                                   double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                   int warp, int nwarps) {
    // om_s: the omega rows of the batch staged in shared memory, [FB][KP].  Each lane carries PP products x NB
    // users = PP*NB independent fma chains and PP loads in flight per k (the sweep is latency bound otherwise).
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    double sv[NB];
#pragma unroll
    for (int f = 0; f < NB; ++f) sv[f] = shiftv[f < nb ? f : nb - 1];
    const int nch = (Ppad + CHUNK - 1) / CHUNK;
    for (int c = warp; c < nch; c += nwarps) {
        double part[NB], perr[NB];
#pragma unroll
        for (int f = 0; f < NB; ++f) { part[f] = 0.0; perr[f] = 0.0; }
        for (int g0 = 0; g0 < CHUNK / 32; g0 += PP) {
            double acc[PP][NB];
            const double *gp[PP];
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int j = c * CHUNK + (g0 + q) * 32 + lane;
                gp[q] = env.GammaT + (j < Ppad ? j : Ppad - 1);
#pragma unroll
                for (int f = 0; f < NB; ++f) acc[q][f] = 0.0;
            }
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                double gk[PP];
#pragma unroll
                for (int q = 0; q < PP; ++q) gk[q] = __ldg(gp[q] + (size_t)k * Ppad);
#pragma unroll
                for (int f = 0; f < NB; ++f) {
                    const double w = om_s[f * KP + k];
#pragma unroll
                    for (int q = 0; q < PP; ++q) acc[q][f] = fma(gk[q], w, acc[q][f]);
                }
            }
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int j = c * CHUNK + (g0 + q) * 32 + lane;
                if (j < P) {
                    const double mu = __ldg(env.mu_o + j);
#pragma unroll
                    for (int f = 0; f < NB; ++f) {
                        if (T2) {
                            const double x = (acc[q][f] + mu) * 1.4426950408889634074 - sv[f];
                            const double m = (double)ex2_approx((float)x);
                            part[f] += m;
                            perr[f] = fma(m, fma(fabs(x), 4.14e-8, 4.8e-7), perr[f]);
                        } else {
                            part[f] += exp64(acc[q][f] + mu - sv[f]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NB; ++f) {
            const double tot = warp_sum(part[f]);
            if (lane == 0 && f < nb) csum[(size_t)f * ncs + c] = tot;
            if (T2) {
                const double te = warp_sum(perr[f]);
                if (lane == 0 && f < nb) esum[(size_t)f * ncs + c] = te;
            }
        }
    }
}

// TIER 2 sweep on the FP64 tensor path: one DMMA (mma.sync m8n8k4, f64) is the [8 users] x [8 products] x [4 k]
// block of the logit contraction -- 256 fma per warp instruction with ONE Gamma element per lane as operand, instead
// of 8 DFMA instructions and PP loads per k.  A = omega (row = user, lane>>2; col = k, lane&3) stays in registers for
// the whole sweep; B = Gamma^T (row = k, lane&3; col = product, lane>>2) straight from L2, K-major; D (row = user;
// cols 2*(lane&3), +1).  The accumulation order inside the DMMA is the hardware's; tier 2 does not need the oracle's
// fma chain, only a bounded error: 2^-40 relative to the logit bound is added to the a-posteriori bound (a binary64
// dot product of K+1 terms is within (K+1) 2^-53 of it).  Tier 3 keeps the oracle's chains.
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int KS>
__device__ __forceinline__ void batch_partial_sums_mma(const EnvDev &env, const double *om_s, int KP, double *csum,
                                                       double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                       int warp, int nwarps) {
    constexpr int G = 4;                         // 8-product groups in flight (independent accumulators)
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int g = lane >> 2, q = lane & 3;
    const int gu = g < nb ? g : nb - 1;          // rows beyond the batch repeat the last user (results discarded)
    double a[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int k = 4 * ks + q;
        a[ks] = k < K ? om_s[gu * KP + k] : 0.0;
    }
    // B operands from the blocked copy of Gamma: the 8 products x KQ rows of a product group are contiguous
    const double *gb = env.GammaB + q * 8 + g;
    const size_t blk = (size_t)env.KQ * 8;
    const double sv = shiftv[gu];
    const int nch = (Ppad + CHUNK - 1) / CHUNK;
    for (int c = warp; c < nch; c += nwarps) {
        double part = 0.0, perr = 0.0;
        for (int g0 = 0; g0 < CHUNK / 8; g0 += G) {
            double acc[G][2];
            const double *pb[G];
#pragma unroll
            for (int i = 0; i < G; ++i) {
                acc[i][0] = 0.0; acc[i][1] = 0.0;
                pb[i] = gb + (size_t)(c * (CHUNK / 8) + g0 + i) * blk;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double b[G];
#pragma unroll
                for (int i = 0; i < G; ++i) b[i] = __ldg(pb[i] + 32 * ks);
#pragma unroll
                for (int i = 0; i < G; ++i) dmma_m8n8k4(acc[i][0], acc[i][1], a[ks], b[i]);
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int p0 = c * CHUNK + (g0 + i) * 8 + 2 * q;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = p0 + h;
                    if (j < P) {
                        const double x = (acc[i][h] + __ldg(env.mu_o + j)) * 1.4426950408889634074 - sv;
                        const double m = (double)ex2_approx((float)x);
                        part += m;
                        perr = fma(m, fma(fabs(x), 4.14e-8, 4.8e-7), perr);
                    }
                }
            }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        perr += __shfl_xor_sync(0xffffffffu, perr, 1);
        perr += __shfl_xor_sync(0xffffffffu, perr, 2);
        if (q == 0 && g < nb) {
            csum[(size_t)g * ncs + c] = part;
            esum[(size_t)g * ncs + c] = perr;
        }
    }
}

// The same sweep for LARGE K (K > 20: KS >= 8 k-steps per product group), where it is bound by the latency of the Gamma
// loads instead of the FP64 pipe: at C4 (P = 100 000, K = 64) the binary64 Gamma is 51 MB, mostly out of L2 by the time a
// sweep starts, and the compiler schedules the form above as load, DMMA, load, DMMA (one or two loads in flight per lane:
// 8 B/clk per SM).  Here the loads run DEPTH - 1 = 7 k-steps ahead of the DMMAs in a register ring (14 loads in flight per
// lane), across product groups and chunks, in an order pinned by volatile asm; omega comes from shared memory per k-step
// (16 steps of fragments do not fit the register budget next to the ring).
__device__ __forceinline__ double ldg_f64_pinned(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int KS>
__device__ __forceinline__ void batch_partial_sums_mma_deep(const EnvDev &env, const double *om_s, int KP, double *csum,
                                                            double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                            int warp, int nwarps) {
    constexpr int G = 2;       // 8-product groups per step (independent accumulators)
    constexpr int DEPTH = 8;   // ring of k-steps: step ks + DEPTH - 1 is loaded while step ks is multiplied
    static_assert(KS % DEPTH == 0, "a group's k-steps must map onto the same ring entries in every group");
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int g = lane >> 2, q = lane & 3;
    const int gu = g < nb ? g : nb - 1;          // rows beyond the batch repeat the last user (results discarded)
    const double *arow = om_s + gu * KP;
    const double *gb = env.GammaB + q * 8 + g;   // blocked copy of Gamma: one product group = one contiguous run
    const size_t blk = (size_t)env.KQ * 8;
    const double sv = shiftv[gu];
    const int nch = (Ppad + CHUNK - 1) / CHUNK;
    constexpr int NG = CHUNK / 8;   // 8-product groups per chunk
    int c = warp, g0 = 0;
    if (c >= nch) return;
    auto group_products = [&](int cc, int gg, const double *(&pbv)[G]) {
#pragma unroll
        for (int i = 0; i < G; ++i) pbv[i] = gb + (size_t)(cc * NG + gg + i) * blk;
    };
    double b[DEPTH][G];
    const double *pb[G], *pbn[G];
    group_products(c, 0, pb);
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) {
#pragma unroll
        for (int i = 0; i < G; ++i) b[d][i] = ldg_f64_pinned(pb[i] + 32 * d);
    }
    double part = 0.0, perr = 0.0;
    for (;;) {
        int cn = c, gn = g0 + G;
        if (gn >= NG) { gn = 0; cn = c + nwarps; }
        const bool has_next = cn < nch;
        group_products(has_next ? cn : c, has_next ? gn : g0, pbn);   // (no next group: harmless re-loads)
        double acc[G][2];
#pragma unroll
        for (int i = 0; i < G; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int sn = ks + DEPTH - 1;   // the step loaded now: of this group, or of the next one
            const int kn = sn < KS ? sn : sn - KS;
#pragma unroll
            for (int i = 0; i < G; ++i) b[sn % DEPTH][i] = ldg_f64_pinned((sn < KS ? pb[i] : pbn[i]) + 32 * kn);
            const double a = 4 * ks + q < K ? arow[4 * ks + q] : 0.0;
#pragma unroll
            for (int i = 0; i < G; ++i) dmma_m8n8k4(acc[i][0], acc[i][1], a, b[ks % DEPTH][i]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int p0 = c * CHUNK + (g0 + i) * 8 + 2 * q;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = p0 + h;
                if (j < P) {
                    const double x = (acc[i][h] + __ldg(env.mu_o + j)) * 1.4426950408889634074 - sv;
                    const double m = (double)ex2_approx((float)x);
                    part += m;
                    perr = fma(m, fma(fabs(x), 4.14e-8, 4.8e-7), perr);
                }
            }
        }
        if (gn == 0) {   // last group of chunk c
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            perr += __shfl_xor_sync(0xffffffffu, perr, 1);
            perr += __shfl_xor_sync(0xffffffffu, perr, 2);
            if (q == 0 && g < nb) {
                csum[(size_t)g * ncs + c] = part;
                esum[(size_t)g * ncs + c] = perr;
            }
            part = 0.0; perr = 0.0;
        }
        if (!has_next) break;
        c = cn; g0 = gn;
#pragma unroll
        for (int i = 0; i < G; ++i) pb[i] = pbn[i];
    }
}

// slots/us/shiftv: shared arrays [FB] (slot index, view uniform, tier-2 log-sum-exp or unused);
// unresolved[f] (shared) is set to 1 when tier 2 cannot certify the draw of user f.
template <bool T2>
__device__ __noinline__ void block_sample_views(const EnvDev &env, const double *om64, int KP, double *csum, int ncs,
                                                const int *slots, const double *us, const double *lse, int nb,
                                                int *s_view, int *unresolved, double *om_s, int tid, int gn,
                                                int bar_id, unsigned &ambig) {
    // Called by a GROUP of gn threads (tid = index inside the group): the whole CTA (bar_id 0, __syncthreads)
    // or the three idle warps of the tensor-core pass 1 (named barrier bar_id).
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int lane = tid & 31, warp = tid >> 5, nwarps = gn >> 5;
    auto gsync = [&]() {
        if (bar_id == 0) cta_sync();
        else if (bar_id < 0) __syncwarp();                       // a single warp (the tensor-mode re-sampling warp)
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gn) : "memory");
    };
    double *esum = csum + (size_t)FB * ncs;
    __shared__ double sh_shift[FB];
    if (tid < FB) {
        double v = 0.0;
        if (tid < nb) {
            if (T2) {
                v = lse[tid];
            } else {  // upper bound of the logits (nats), as in warp_sample_view_exact
                const double *om = om64 + (size_t)slots[tid] * KP;
                double n2 = 0.0;
                for (int k = 0; k < K; ++k) n2 = fma(om[k], om[k], n2);
                v = sqrt(n2) * env.gmax + env.mumax;
            }
        }
        sh_shift[tid] = v;
    }
    for (int idx = tid; idx < FB * KP; idx += gn) {  // stage the omega rows of the batch (rows >= nb repeat the last)
        const int f = idx / KP, k = idx - f * KP;
        om_s[idx] = k < K ? om64[(size_t)slots[f < nb ? f : nb - 1] * KP + k] : 0.0;
    }
    gsync();
#ifdef RG_T2_DFMA
    if (false) {
#else
    if (T2) {   // FP64 tensor path (any nb <= 8 in one pass)
#endif
        if (K <= 20) batch_partial_sums_mma<5>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps);
        else if (K <= 32) batch_partial_sums_mma_deep<8>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps);
        else if (K <= 64) batch_partial_sums_mma_deep<16>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps);
        else batch_partial_sums_mma_deep<32>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps);
    } else
    switch (nb) {
        case 1: batch_partial_sums<1, 8, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 2: batch_partial_sums<2, 8, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 3: batch_partial_sums<3, 4, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 4: batch_partial_sums<4, 4, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 5: batch_partial_sums<5, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 6: batch_partial_sums<6, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        case 7: batch_partial_sums<7, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
        default: batch_partial_sums<8, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;
    }
    gsync();
    for (int f = warp; f < nb; f += nwarps) {
        const double *omega = om_s + (size_t)f * KP;
        const double *cs = csum + (size_t)f * ncs;
        const double sf = sh_shift[f];
        const int nch = (Ppad + CHUNK - 1) / CHUNK;
        double zl = 0.0, el = 0.0;
        for (int c = lane; c < nch; c += 32) {
            zl += cs[c];
            if (T2) el += esum[(size_t)f * ncs + c];
        }
        const double Z = warp_sum(zl);
        const double E = T2 ? warp_sum(el) : 0.0;
        const double tau = us[f] * Z;
        // chunk containing tau: 32 chunks per step, inclusive scan over lanes
        double cum = 0.0;
        int csel = nch - 1;
        bool have = false;
        for (int c0 = 0; c0 < nch && !have; c0 += 32) {
            const int c = c0 + lane;
            double x = c < nch ? cs[c] : 0.0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            const double pre = cum + x;
            const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (c < nch));
            if (m) {
                const int fl = __ffs(m) - 1;
                csel = c0 + fl;
                const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
                cum = fl > 0 ? below : cum;
                have = true;
            } else {
                cum = __shfl_sync(0xffffffffu, pre, 31);
            }
        }
        if (!have) {  // tau >= every prefix (rounding-level): walk the last chunk from its start
            double c2 = 0.0;
            for (int c = 0; c < nch - 1; ++c) c2 += cs[c];
            cum = c2;
        }
        int found = -1;
        double lower = 0.0, upper = 0.0;
        const int end = min((csel + 1) * CHUNK, Ppad);
        {   // the whole chunk at once: 8 masses per lane (8 independent chains), 8 butterfly totals, one scan
            constexpr int G = CHUNK / 32;
            double acc[G], xs[G], tq[G];
            const double *gp[G];
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int j = csel * CHUNK + q * 32 + lane;
                gp[q] = env.GammaT + (j < Ppad ? j : Ppad - 1);
                acc[q] = 0.0;
            }
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                const double w = omega[k];
#pragma unroll
                for (int q = 0; q < G; ++q) acc[q] = fma(__ldg(gp[q] + (size_t)k * Ppad), w, acc[q]);
            }
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int j = csel * CHUNK + q * 32 + lane;
                double m = 0.0;
                if (j < P) {
                    const double l = acc[q] + __ldg(env.mu_o + j);
                    m = T2 ? (double)ex2_approx((float)(l * 1.4426950408889634074 - sf)) : exp64(l - sf);
                }
                xs[q] = m;
                tq[q] = m;
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
                for (int q = 0; q < G; ++q) tq[q] += __shfl_xor_sync(0xffffffffu, tq[q], d);
            }
            bool done = false;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                if (done) break;
                if (cum + tq[q] > tau || q == G - 1) {
                    double x = xs[q];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double y = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= d) x += y;
                    }
                    const double pre = cum + x;
                    const int base = csel * CHUNK + q * 32;
                    const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (base + lane < P));
                    if (m) {
                        const int fl = __ffs(m) - 1;
                        found = base + fl;
                        upper = __shfl_sync(0xffffffffu, pre, fl);
                        const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
                        lower = fl > 0 ? below : cum;
                    }
                    done = true;
                } else {
                    cum += tq[q];
                }
            }
        }
        bool amb = !T2 && env.force_t4 != 0;
        if (found < 0) {
            found = min(end, P) - 1;
            amb = true;
        } else {
            // tier 2: the a-posteriori bound E of sum_p |e_p| enters S^ - u Z^ as (1-u) e(<=j) - u e(>j), i.e. at most
            // max(u, 1-u) E, plus twice the crossing chunk's share (its masses are recomputed here, the sweep's
            // values sit in Z^); tier 3: binary64 summation-order level
            double tol = fmax(1e-11, 4.0 * (double)P * 1.1102230246251565e-16) * Z;
            if (T2) {
                const double uu = us[f];
                tol = 1.01 * (fmax(uu, 1.0 - uu) * E + 2.0 * uu * esum[(size_t)f * ncs + csel]) + 1e-12 * Z;
            }
            amb = amb || (tau - lower < tol && found > 0) || (upper - tau < tol && found < P - 1);
        }
        if (!T2 && amb)   // warp-uniform: neither binary64 order can be trusted -> the oracle's own order (tier 4)
            found = warp_sample_view_oracle_order(env, omega, us[f], lane);
        if (lane == 0) {
            if (T2) {
                unresolved[f] = amb ? 1 : 0;
                if (!amb) s_view[slots[f]] = found;
            } else {
                s_view[slots[f]] = found;
                if (amb) ambig += 1;
            }
        }
    }
    gsync();
}

// ---- longest-first schedule --------------------------------------------------------------------------------------------
// A launch ends with a tail: when the ticket queue is empty every CTA still has to finish the episodes its slots hold, and
// a round costs the same with 3 live users as with 128 (13 % of a 1M-user launch at C3).  Episodes that start last should
// therefore be the shortest ones.  The number of organic draws of a user is ESTIMATED by replaying its Markov chain alone
// -- the transition uniforms are a function of (user, t) -- with clicks guessed from the click uniform (a click needs
// u >= 1 - ff, ff in [0.0045, 0.0293]; the guess uses the policy's mean click probability at omega = 0), users are bucketed by
// the estimate, and tickets are handed out from the longest bucket down.  Only the ORDER in which users are simulated
// changes: every result is a pure function of (seed, user).
constexpr int LPT_BUCKETS = 256;

// mean click probability of the acting policy at omega = 0: (1/P) sum_v ff(mu_bandit[a(v)]) with a(v) the table's action
// (table policies) or v itself (uniform / user_views) -- the click-guess threshold of the estimate
__global__ void lpt_ctr_kernel(const EnvDev env, const PolicyDev pol, double *__restrict__ ctr_mean) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int v = threadIdx.x; v < env.P; v += blockDim.x) {
        const int a = (pol.kind == RG_POLICY_TABLE || pol.kind == RG_POLICY_EPS_TABLE) ? __ldg(pol.table + v) : v;
        s += ff64(__ldg(env.mu_b + a));
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        *ctr_mean = t / (double)env.P;
    }
}

__global__ void lpt_estimate_kernel(const EnvDev env, long long user_begin, long long n_organic_users, long long n_total,
                                    const double *__restrict__ ctr_mean, unsigned char *__restrict__ est,
                                    unsigned *__restrict__ hist) {
    const double click_thr = 1.0 - *ctr_mean;
    __shared__ unsigned sh[LPT_BUCKETS];
    for (int b = threadIdx.x; b < LPT_BUCKETS; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) {
        const uint64_t user = (uint64_t)(user_begin + i);
        const bool organic_only = i < n_organic_users;
        int state = RG_ORGANIC, n_org = 0;
        for (uint32_t t = 0; t < 8192u && state != RG_STOP; ++t) {
            const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
            const int was = state;
            state = next_state(env, was, uniform53(w.z, w.w));
            if (was == RG_ORGANIC) {
                n_org += 1;
                if (organic_only && state != RG_ORGANIC) break;
            } else if (uniform53(w.x, w.y) >= click_thr) {
                state = RG_ORGANIC;   // guessed click (abstract.py:180-181)
            }
        }
        const int b = n_org < LPT_BUCKETS ? n_org : LPT_BUCKETS - 1;
        est[i] = (unsigned char)b;
        atomicAdd(&sh[b], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < LPT_BUCKETS; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[b], sh[b]);
}

__global__ void lpt_cursor_kernel(unsigned *__restrict__ hist) {   // in place: first ticket of each bucket, longest first
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int b = LPT_BUCKETS - 1; b >= 0; --b) { const unsigned c = hist[b]; hist[b] = run; run += c; }
    }
}

__global__ void lpt_scatter_kernel(const unsigned char *__restrict__ est, long long n_total, unsigned *__restrict__ cursor,
                                   int32_t *__restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) perm[atomicAdd(&cursor[est[i]], 1u)] = (int32_t)i;
}

template <bool TC>
__device__ __forceinline__ void rollout_cta_body(const Params &p) {
    // NT compute threads (8 warps).  TC: a ninth warp is the MMA issuer of pass 1; warps 0-7 run the epilogue (thread =
    // TMEM lane = user; warps w and w+4 share a lane quarter and take one column half of every accumulator tile each, so
    // that every scheduler has two warps feeding the MUFU pipe), warp 0 also refills the B ring.
    constexpr int NTH = NT, NWARP = NTH / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const EnvDev &env = p.env;
    const Layout &L = p.L;
    const int K = env.K, K1 = K + 1, KP = p.kp, P = env.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint64_t *mbar = (uint64_t *)(smem + L.mbar);
    // (the spill levels only exist in the CUDA-core form: for TC both selections are compile-time)
    double *om64 = (TC || p.spill >= 2) ? p.om_scratch + (size_t)blockIdx.x * TU * p.kp : (double *)(smem + L.om64);
    double *csum = (!TC && p.spill >= 1) ? p.csum_scratch + (size_t)blockIdx.x * (size_t)p.ncs * FB * 2
                                         : (double *)(smem + L.csum);
    float *const gsum_cta = TC ? p.gsum + (size_t)blockIdx.x * (size_t)p.env.nsub * 4 * TU : nullptr;
    uint64_t *s_user = (uint64_t *)(smem + L.user);
    uint64_t *s_hash = (uint64_t *)(smem + L.hash);
    uint32_t *s_t = (uint32_t *)(smem + L.t);
    int *s_state = (int *)(smem + L.state);
    int *s_last_view = (int *)(smem + L.last_view);
    int *s_n_org = (int *)(smem + L.n_org);
    int *s_n_ban = (int *)(smem + L.n_ban);
    int *s_n_clk = (int *)(smem + L.n_clk);
    int *s_view = (int *)(smem + L.view);
    int *s_live = (int *)(smem + L.live);
    int *pos2slot = (int *)(smem + L.pos2slot);
    float *s_shift = (float *)(smem + L.shift);
    float *s_eps = (float *)(smem + L.eps);
    int *amb_list = (int *)(smem + L.amb_list);
    int *s_nview = (int *)(smem + L.nview);   // RG_POLICY_USER_VIEWS: distinct products viewed by the slot's user
    const bool track_views = p.pol.kind == RG_POLICY_USER_VIEWS;
    int2 *const vlists = track_views ? p.pol.views + (size_t)blockIdx.x * TU * p.pol.view_cap : nullptr;
    // acting policy at (user, t): Agent.act of the compiled agent (include/recogym_b200.h RG_POLICY_*)
    auto policy_act = [&](int slot, uint64_t user, uint32_t t, int last_view, int n_org) -> int {
        if (p.pol.kind == RG_POLICY_TABLE) return __ldg(p.pol.table + last_view);
        const u32x4 w = draw(p.env.seed, user, t, SLOT_POLICY);
        if (p.pol.kind == RG_POLICY_EPS_TABLE) return eps_table_action(p.pol, p.env.P, last_view, w);
        if (p.pol.kind == RG_POLICY_USER_VIEWS)
            return views_action(vlists + (size_t)slot * p.pol.view_cap, s_nview[slot], n_org, uniform53(w.z, w.w));
        return uniform_action(w.x, p.env.P);
    };
    int *fb_slot = (int *)(smem + L.fb_slot);
    double *fb_u = (double *)(smem + L.fb_u);
    double *fb_lse = (double *)(smem + L.fb_lse);
    double *fb_om = (double *)(smem + L.fb_om);
    double *s_Z = (double *)(smem + L.pz);
    double *s_u = (double *)(smem + L.pu);
    double *rq_u = (double *)(smem + L.rq_u);       // re-sampling queues [2][TU] (double-buffered by round parity)
    double *rq_lse = (double *)(smem + L.rq_lse);
    int *rq_slot = (int *)(smem + L.rq_slot);
    int *q3_slot = (int *)(smem + L.q3_slot);      // draws tier 2 could not certify, waiting for a tier-3 sweep
    double *q3_u = (double *)(smem + L.q3_u);
    double *s_clock = (double *)(smem + L.clock);   // only touched when env.time_mode != 0
    double *s_dt = (double *)(smem + L.dt);
    const int tmode = env.time_mode;
    int *s_wait = (int *)(smem + L.wait);           // 0 none, 2 queued this round, 1 being re-sampled this round
    double *s_tau = (double *)(smem + L.ptau);
    double *s_cum = (double *)(smem + L.pcum);
    int *s_rsel = (int *)(smem + L.prsel);
    int *fb_unres = (int *)(smem + L.fb_unres);
    int *s_pend = (int *)(smem + L.pend);      // 0 none, 1 drift into s_t, 2 init of a fresh user
    int *pend_list = (int *)(smem + L.pend_list);
    int *misc = (int *)(smem + L.misc);  // [0] nlive, [1] amb_count, [2] exact result
    float *om32 = (float *)(smem + L.om32);
    float *tilesum = (float *)(smem + L.tilesum);
    float *gtile = (float *)(smem + L.gtile);
    tc::opnd_t *a_tile = (tc::opnd_t *)(smem + L.atile);
    uint64_t *dfull = mbar + kMaxStages;                   // TC: accumulator tile ready, one per TMEM buffer
    uint64_t *dempty = dfull + NBUF;                       // TC: accumulator tile drained by the 4 epilogue warps that read it
    uint32_t *tmem_slot = (uint32_t *)(dempty + NBUF);
    const int KC = env.KC, KCs = env.KCs, nslab = env.nslab;
    const unsigned tc_tile_bytes = 2u * (unsigned)KCs * tc::CHUNK_BYTES_B;   // one K slab of a B tile

    const unsigned tile_bytes = TC ? tc_tile_bytes : (unsigned)(4 * K1 * SUB);
    const int nsub = env.nsub, nstage = p.nstage;
    const int ntile = TC ? nsub * (128 / tc::TILE_N) : nsub;  // product tiles per round
    const int nunit = TC ? ntile * nslab : nsub;               // ring units streamed per round (TC: K slabs of tiles)

    if constexpr (TC) {
        if (warp == NWARP) {
            // ---- MMA issuer (ninth warp; converged, warp-uniform values, one elected lane issues).  Per ring unit (K
            // slab of a product tile): B slab landed -> accumulator buffer drained (first slab of a tile) -> three-term
            // MMA -> commit "accumulator ready" after the last slab.  Every scalar instruction of this loop is on the
            // critical path of pass 1: ring stage, accumulator buffer and their phase bits are carried incrementally,
            // descriptors are additions to a precomputed low word, the MMA sequence is unrolled for the slab size.
            issuer_sync();   // set-up done (barriers initialised, TMEM allocated)
            tc::tc_fence_after();
            const uint32_t tmem_u = *(volatile uint32_t *)tmem_slot;
            const uint32_t mbar_s = smem_u32(mbar), dfull_s = smem_u32(dfull), dempty_s = smem_u32(dempty);
            const uint32_t a_lo0 = tc::umma_desc_lo(tc::smem_addr(a_tile), tc::CHUNK_BYTES);
            const uint32_t b_lo0 = tc::umma_desc_lo(tc::smem_addr(gtile), tc::CHUNK_BYTES_B);
            const uint32_t a_term16 = (uint32_t)KC * tc::CHUNK_BYTES / 16u, a_slab16 = (uint32_t)KCs * tc::CHUNK_BYTES / 16u;
            const uint32_t b_stage16 = tile_bytes / 16u;
            const uint32_t el = tc::elect_one();
            unsigned st = 0u, stph = 0u, buf = 0u, use = 0u;   // carried across rounds (nunit units per round)
            RG_PROF_ONLY(unsigned long long pf[4] = {0ull, 0ull, 0ull, 0ull}; int round_no = 0;)
            for (;;) {
                RG_PROF_ONLY(++round_no; const bool tr_on = blockIdx.x == 0 && round_no == 30;)
                issuer_sync();   // a round's A tile is complete, or the launch is over
                if (*(volatile int *)&misc[9]) break;
                tc::tc_fence_after();
                int sl = 0;
                for (int uu = 0; uu < nunit; ++uu) {
                    RG_PROF_ONLY(const long long pa = clock64();)
                    mbar_wait_s(mbar_s + st * 8u, stph);
                    RG_PROF_ONLY(const long long pb = clock64();)
                    if (sl == 0 && use > 0) mbar_wait_s(dempty_s + buf * 8u, (use - 1u) & 1u);
                    RG_PROF_ONLY(const long long pc = clock64();)
                    tc::tc_fence_after();
                    const uint32_t td = tmem_u + buf * (uint32_t)tc::TILE_N;
                    const uint32_t a_lo = a_lo0 + (uint32_t)sl * a_slab16, b_lo = b_lo0 + st * b_stage16;
                    const uint32_t accf = sl > 0 ? 1u : 0u;
                    switch (KCs) {
                        case 2: tc::issue_slab_mmas_lean<2>(td, a_lo, a_term16, b_lo, accf, el); break;
                        case 4: tc::issue_slab_mmas_lean<4>(td, a_lo, a_term16, b_lo, accf, el); break;
                        case 6: tc::issue_slab_mmas_lean<6>(td, a_lo, a_term16, b_lo, accf, el); break;
                        case 8: tc::issue_slab_mmas_lean<8>(td, a_lo, a_term16, b_lo, accf, el); break;
                        default: tc::issue_slab_mmas_lean<10>(td, a_lo, a_term16, b_lo, accf, el); break;
                    }
                    if (sl == nslab - 1) umma_commit_s(dfull_s + buf * 8u, el);   // the accumulator tile is complete
                    RG_PROF_ONLY(const long long pd = clock64(); pf[0] += pd - pc; pf[1] += pc - pb; pf[2] += pb - pa; pf[3] += 1ull;
                                 if (tr_on && uu < 1000 && lane == 0) { g_trace[uu * 8 + 0] = pb; g_trace[uu * 8 + 1] = pc; g_trace[uu * 8 + 2] = pd; })
                    if (++st == (unsigned)nstage) { st = 0u; stph ^= 1u; }
                    if (++sl == nslab) { sl = 0; if (++buf == NBUF) { buf = 0u; ++use; } }
                }
            }
            RG_PROF_ONLY(if (lane == 0) { atomicAdd(&g_prof[16], pf[2]); atomicAdd(&g_prof[17], pf[1]); atomicAdd(&g_prof[18], pf[0]); atomicAdd(&g_prof[19], pf[3]); })
            return;
        }
    }
    // ---- set-up: barriers, first claims, first tile loads
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) mbar_init(&mbar[s], 1);
        if (TC) for (int b = 0; b < NBUF; ++b) { mbar_init(&dfull[b], 1); mbar_init(&dempty[b], NWARP / 2); }
        fence_barrier_init();
        misc[1] = 0;
        misc[9] = 0;   // TC: 1 = the launch is over (read by the issuer warp after barrier 2)
    }
    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // NBUF accumulator tiles of TILE_N columns
    unsigned long long tot_users = 0, tot_events = 0, tot_org = 0, tot_ban = 0, tot_clk = 0, tot_ph = 0;
    unsigned long long n_fallback = 0, n_tier3 = 0;
    unsigned long long ph_a = 0, ph_b = 0, ph_c = 0, ph_c2 = 0, ph_d = 0, ph_o = 0;
    unsigned ambig64 = 0;

    // RG_PROF: sub-phase cycle counters of thread 0 (g_prof[48..]); a shared-memory read orders the clock read behind the
    // release of the preceding barrier (a bare clock read after bar.sync captures the barrier's ISSUE time)
    RG_PROF_ONLY(unsigned long long pf2[12]; for (int i = 0; i < 12; ++i) pf2[i] = 0ull; long long pm_last = clock64();)
#ifdef RG_PROF
#define RG_MARK(idx) do { if (tid == 0) { (void)*(volatile int *)misc; const long long pm_now = clock64(); pf2[idx] += (unsigned long long)(pm_now - pm_last); pm_last = pm_now; } } while (0)
#else
#define RG_MARK(idx) do { } while (0)
#endif
    auto claim = [&](int slot) -> bool {  // next user id from the global queue -> reset()
        const unsigned long long q = atomicAdd(p.queue, 1ull);
        if (q >= (unsigned long long)p.n_total) {
            s_live[slot] = 0;
            return false;
        }
        const uint64_t user = (uint64_t)(p.user_begin + (long long)(p.perm ? p.perm[q] : (long long)q));
        s_user[slot] = user;
        s_hash[slot] = kHashInit;
        s_t[slot] = 0u;
        s_state[slot] = RG_ORGANIC;
        s_last_view[slot] = -1;
        s_n_org[slot] = 0; s_n_ban[slot] = 0; s_n_clk[slot] = 0;
        s_nview[slot] = 0;
        s_live[slot] = 1;
        s_wait[slot] = 0;
        s_pend[slot] = 2;  // omega = sigma0 * z(user, 0) (reco_env_v1.py:80-82), done by omega_phase()
        if (tmode) s_clock[slot] = 0.0;   // time_generator.reset(); current_time = new_time() = 0
        return true;
    };
    // Pending omega updates of all slots, spread over the whole CTA: task = (slot, normal pair j).
    auto omega_phase = [&]() {
        if (warp == 0) {
            int base = 0;
            for (int c = 0; c < TU / 32; ++c) {
                const int slot = c * 32 + lane;
                const bool pd = s_pend[slot] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, pd);
                if (pd) pend_list[base + __popc(m & ((1u << lane) - 1u))] = slot;
                base += __popc(m);
            }
            if (lane == 0) misc[3] = base;
        }
        cta_sync();
        RG_MARK(8);   // omega: list of pending slots
        const int npend = misc[3], npairs = (K + 1) / 2;
        for (int task = tid; task < npend * npairs; task += NTH) {
            const int slot = pend_list[task / npairs], j = task - (task / npairs) * npairs;
            const bool init = s_pend[slot] == 2;
            double *om = om64 + (size_t)slot * KP;
            const double scale = init ? env.sigma0
                                      : env.sigma * (tmode ? s_dt[slot] : 1.0);
            const u32x4 w = draw(env.seed, s_user[slot], s_t[slot], SLOT_NORMAL + (uint32_t)j);
            const double z0 = normal_from_words(w.x, w.y);
            om[2 * j] = (init ? 0.0 : om[2 * j]) + scale * z0;
            if (2 * j + 1 < K) {
                const double z1 = normal_from_words(w.z, w.w);
                om[2 * j + 1] = (init ? 0.0 : om[2 * j + 1]) + scale * z1;
            }
        }
        cta_sync();
        RG_MARK(9);   // omega: normals + updates
        if (tid < TU) s_pend[tid] = 0;
    };
    if (tid < TU) { s_pend[tid] = 0; s_wait[tid] = 0; claim(tid); }
    if (tid == 0) { misc[6] = 0; misc[4] = 0; }
    cta_sync();
    omega_phase();
    cta_sync();

    unsigned seq_use = 0;   // sequence number of the next ring unit to consume (same in every thread)
    auto issue_load = [&](unsigned seq) {  // thread 0 only: unit (seq % nunit) -> ring stage (seq % nstage)
        const int stage = (int)(seq % (unsigned)nstage);
        const int sub = (int)(seq % (unsigned)nunit);
        mbar_expect_tx(&mbar[stage], tile_bytes);
        if (TC)
            tma_bulk_g2s((unsigned char *)gtile + (size_t)stage * tile_bytes, (const unsigned char *)env.GTC + (size_t)sub * tile_bytes,
                         tile_bytes, &mbar[stage]);
        else
            tma_bulk_g2s(gtile + (size_t)stage * K1 * SUB, env.G32 + (size_t)sub * K1 * SUB, tile_bytes, &mbar[stage]);
    };
    // non-TC: the ring is always nstage-1 tiles ahead, across rounds.  TC: the ring bytes are reused between
    // pass 1 and the next round, so each round starts with its own prologue of min(nstage, nsub) loads.
    const int tc_prologue = nstage < nunit ? nstage : nunit;
    if (tid == 0) {
        if (TC) for (int s = 0; s < tc_prologue; ++s) issue_load((unsigned)s);
        else for (int s = 0; s < nstage - 1; ++s) issue_load((unsigned)s);
    }
    uint32_t tmem_base = 0;
    if (TC) {
        tc::tc_fence_before();
        issuer_sync();     // barriers initialised, TMEM allocated: the issuer warp may read the TMEM address
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }

    const int tx = tid & 15, ty = tid >> 4;

    RG_PROF_ONLY(int round_no = 0; unsigned long long pf[16]; for (int i = 0; i < 16; ++i) pf[i] = 0ull;)
    for (;;) {
        RG_PROF_ONLY(++round_no; const bool tr_on = blockIdx.x == 0 && round_no == 30;)
        long long tk0 = clock64();   // diagnostic: cycles per phase as seen by thread 0 (totals[10..15])
        // ---- A: compact the live slots into positions 0..nlive-1 (no data moves: pos2slot indirection)
        if (warp == 0) {
            int base = 0;
            for (int c = 0; c < TU / 32; ++c) {
                const int slot = c * 32 + lane;
                const bool lv = s_live[slot] != 0 && s_wait[slot] == 0;   // s_wait: the slot's draw is queued for a re-sampling sweep
                const unsigned m = __ballot_sync(0xffffffffu, lv);
                if (lv) pos2slot[base + __popc(m & ((1u << lane) - 1u))] = slot;
                base += __popc(m);
            }
            if (lane == 0) misc[0] = base;
        }
        cta_sync();
        RG_MARK(0);   // A: compaction (+ whatever was left of the previous round's last barrier)
        const int nlive = misc[0];
        RG_PROF_ONLY(if (tid == 0) { pf[12] += (unsigned long long)nlive; pf[13] += 1ull; })
        const int nq = misc[6];  // draws waiting for a re-sampling sweep (tier 2) ...
        const int nq3 = misc[4]; // ... and for a tier-3 sweep
        if (nlive == 0 && nq == 0 && nq3 == 0) {
            if (TC) {   // release the issuer warp into its exit
                if (tid == 0) misc[9] = 1;
                issuer_sync();
            }
            break;
        }

        // fp32 operand tile [K+1][TU] (row K = 1 multiplies the mu row of the Gamma tile), per-user shift and
        // error bound.  Dead positions get zeros.  (TC: the fp32 tile is only needed by pass 2 and shares its
        // bytes with the B ring, so it is built after pass 1.)
        auto build_om32 = [&]() {   // thread = (user position, half of the rows): the omega loads of a user pipeline
            const int i = tid & (TU - 1), kh = tid >> 7, kmid = (K1 + 1) / 2;
            const bool alive = i < nlive;
            const double *om = om64 + (size_t)(alive ? pos2slot[i] : 0) * KP;
            const int k0 = kh ? kmid : 0, k1 = kh ? K1 : kmid;
#pragma unroll 4
            for (int k = k0; k < k1; ++k) om32[k * TU + i] = alive ? (k < K ? (float)om[k] : 1.f) : 0.f;
        };
        if (!TC) build_om32();
        if (tid < TU) {
            float sh = 0.f, ep = 1.f;
            if (tid < nlive) {
                const double *om = om64 + (size_t)pos2slot[tid] * KP;
                double n2 = 0.0;
                for (int k = 0; k < K; ++k) n2 = fma(om[k], om[k], n2);
                const double nrm = sqrt(n2);
                const double kLog2e = 1.4426950408889634074;
                const double bl2 = (nrm * env.gmax + env.mumax) * kLog2e;    // >= every log2-logit
                const double shift = bl2 - 100.0;                              // top mass <= 2^100
                const double aub = (nrm * env.gmax + env.muabsmax) * kLog2e;  // >= sum_k |G'_pk w_k| + |mu'_p|
                const double u24 = 5.9604644775390625e-08;                    // 2^-24
                // log2-logit error: input rounding + K+1 fma roundings + the shift subtraction
                double dl2 = (K + 4) * u24 * aub + u24 * (aub + fabs(shift)) + u24 * fabs(shift);
                if (TC) {
                    // tensor-core tile sums: dropped lo*lo term (2^-22 per product term, doubled) and the fp32
                    // accumulation of the UMMA, whose rounding is undocumented: 8 roundings at the largest
                    // magnitude.  Measured (tests): <= 0.5 of this bound for K = 5..64.
                    double dtc = aub * 4.76837158203125e-07 + 8.0 * u24 * (aub + fabs(shift));
#if RG_TC_F16
                    // fp16 operand elements: a `lo` term below 2^-14 is subnormal and carries an absolute error of up to
                    // 2^-25 per element: sum_k (|dG_k| |w_k| + |G'_k| |dw_k|) + the mu' and shift columns
                    dtc += 2.98023223876953125e-08 * (sqrt((double)K) * (nrm + env.gmax * kLog2e) + 2.0);
#endif
                    dl2 = fmax(dl2, dtc);
                }
                // relative mass error: exp of that, ex2.approx (2^-21 covers its 2 ulp), fp32 partial sums of pass 1
                // fp32 partial sums of pass 1: longest chain of additions behind one record-tile sum
                const double nadd = TC ? (double)(tc::TILE_N / 4 + 4 + p.rec_sub * (128 / tc::TILE_N)) : 8.0 * p.rec_sub + 6.0;
                const double delta = 0.6931471805599453 * dl2 * 1.01 + 4.76837158203125e-07 + nadd * u24;
                sh = (float)shift;
                ep = (float)(delta * 1.05 + 5e-8);   // bound of the relative error of every computed mass
            }
            s_shift[tid] = sh;
            s_eps[tid] = ep;
        }
        cta_sync();

        RG_MARK(1);   // A: logit bounds, shifts
        { const long long tk1 = clock64(); ph_a += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
        // ---- B: pass 1, dense contraction + exponentials + per-record-tile partial sums
        if constexpr (TC) {
            // Tensor-core form.  A = omega tile (tf32 hi/lo, UMMA K-major layout) written by the CTA, B = Gamma
            // tiles streamed by TMA, D = shifted log2-logits of 128 users x 128 products in TMEM (two buffers).
            // Thread 0 issues the three-term MMA of tile s+1 before the CTA runs the epilogue of tile s:
            // thread (lane quarter q, column half h) owns user 32q+lane and 64 columns; per logit the epilogue is
            // one ex2 and one add.
            {   // thread = (row, half of the K chunks): one 16-byte chunk (4 consecutive k of the row) per step, the
                // chunk's four omega loads in flight together, float4 stores (consecutive rows -> consecutive 16 bytes)
                static_assert(NT == 2 * TU, "A-tile build: two threads per row");
                const int row = tid & (TU - 1), c0 = (tid >> 7) * (KC / 2);   // KC is even
                const bool alive = row < nlive;
                const double *om = om64 + (size_t)(alive ? pos2slot[row] : 0) * KP;
                const float nshift = alive ? -s_shift[row] : 0.f;
                bool in_range = true;
                for (int c = c0; c < c0 + KC / 2; ++c) {
                    union { tc::opnd_t e[tc::EPC]; uint4 v; } hi, lo;
#pragma unroll
                    for (int j = 0; j < tc::EPC; ++j) {
                        const int k = tc::EPC * c + j;
                        float v = 0.f;
                        if (alive) v = k < K ? (float)om[k] : (k == K ? 1.f : (k == K + 1 ? nshift : 0.f));
                        if (!(fabsf(v) <= tc::kOpndMax)) { in_range = false; v = 0.f; }
                        tc::split_hi_lo(v, hi.e[j], lo.e[j]);
                    }
                    *reinterpret_cast<uint4 *>(a_tile + tc::tile_index(0, KC, row, tc::EPC * c)) = hi.v;
                    *reinterpret_cast<uint4 *>(a_tile + tc::tile_index(1, KC, row, tc::EPC * c)) = lo.v;
                }
                // a latent state (or shift) outside the operand element range cannot be sampled by this kernel: counted in
                // totals[9] ("must be 0"), like the overflow of a view list
                if (!in_range && p.out.totals) atomicAdd(p.out.totals + 9, 1ull);
            }
            tc::fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core
            tc::tc_fence_before();
            issuer_sync();            // the A tile is complete: the issuer warp starts this round's MMAs
            tc::tc_fence_after();
            RG_PROF_ONLY(const long long pq0 = clock64(); if (tid == 0) pf[10] += (unsigned long long)(pq0 - tk0);)
            {
                // ---- epilogue, all 8 warps: thread = TMEM lane = user 32 q + lane.  The two warps of a lane quarter (w and
                // w + 4, same scheduler) take ALTERNATE accumulator tiles -- group g = w / 4 reads the tiles tt = g, g + 2, ...,
                // i.e. always the same TMEM buffer -- so that their hand-shakes and loads never coincide and the scheduler
                // always has one warp in its run of exponentials (two warps per scheduler saturate the MUFU pipe, one does
                // not: scripts/micro/epilogue_mma.cu).  Per logit: one ex2 and one add.
                const int q = warp & 3, grp = warp >> 2;
                const int pos = q * 32 + lane;
                const bool act = pos < nlive;
                const bool warp_act = q * 32 < nlive;   // a warp without live users only keeps the hand-shakes going
                const int tiles_per_rec = p.rec_sub * (128 / tc::TILE_N);
                const unsigned tile_use = seq_use / (unsigned)nslab;
                const uint32_t dfull_s = smem_u32(dfull), dempty_s = smem_u32(dempty);
                const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
                constexpr int NCH = tc::TILE_N / 32;     // 32-column chunks per tile
                static_assert(NBUF == 2, "one accumulator buffer per epilogue warp group");
                const int ring_tiles = nstage / nslab;   // product tiles the B ring holds (nstage is a multiple of nslab)
                float *gs_row = gsum_cta + pos;
                float *ts_row = tilesum + (size_t)pos * p.nrec_pad + grp;
                int tt = grp;
#ifdef RG_X_DEPHASE
                // experiment: start the second warp group late, so that the two groups' waits for their accumulators do
                // not coincide (they are released into pass 1 by the same barrier and would otherwise stay in lockstep)
                if (grp) { const long long d0 = clock64(); while (clock64() - d0 < (long long)(RG_X_DEPHASE)) { } }
#endif
                for (int rec = 0; rec < p.nrec; ++rec) {
                    const int rec_end = min((rec + 1) * tiles_per_rec, ntile);
                    float part = 0.f;
                    for (; tt < rec_end; tt += 2) {
                        const unsigned gs = tile_use + (unsigned)tt;
                        const unsigned buf = gs % NBUF;
                        const uint32_t taddr = lane_taddr + buf * (uint32_t)tc::TILE_N;
                        RG_PROF_ONLY(const long long pa = clock64();)
                        mbar_wait_s(dfull_s + buf * 8u, (gs / NBUF) & 1u);
                        tc::tc_fence_after();
                        RG_PROF_ONLY(const long long pb = clock64(); long long pld = 0;)
                        // Accumulator tile tt is complete, i.e. every UMMA that read its ring stages has completed: refill
                        // them with the tile that comes ring_tiles later (no "stage free" barrier is needed).
                        if (q == 0 && lane == 0 && tt + ring_tiles < ntile) {
                            for (int sl = 0; sl < nslab; ++sl) issue_load(seq_use + (unsigned)((tt + ring_tiles) * nslab + sl));
                        }
                        float *gs_out = gs_row + (size_t)tt * (NCH * TU);
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            float v[32];
                            RG_PROF_ONLY(const long long pl0 = clock64();)
#ifdef RG_X_NOLD
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = -3.f - 0.01f * (float)(lane + i + tt);
#else
                            if (warp_act) tc::tmem_ld32(taddr + (uint32_t)(32 * c), v);
#endif
                            RG_PROF_ONLY(pld += clock64() - pl0;)
                            if (c == NCH - 1) {   // last read of this tile: hand the accumulator back before the math
                                tc::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dempty_s + buf * 8u) : "memory");
                            }
                            if (warp_act) {
                                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
#ifdef RG_X_NOMUFU
                                    s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3];
#else
                                    s0 += ex2_approx(v[i]); s1 += ex2_approx(v[i + 1]);
                                    s2 += ex2_approx(v[i + 2]); s3 += ex2_approx(v[i + 3]);
#endif
                                }
                                // the masses of this 32-product group: kept (coalesced over the 32 users of the warp) so that
                                // pass 2b only has to recompute ONE group of 32 products instead of the whole record tile
                                const float cs = (s0 + s1) + (s2 + s3);
#ifndef RG_X_NOSTG
                                if (act) gs_out[c * TU] = cs;
#endif
                                part += cs;
                            }
                        }
                        RG_PROF_ONLY(if (lane == 0) { const long long pe = clock64(); pf[6] += pb - pa; pf[7] += pld; pf[8] += pe - pb; pf[9] += 1ull;
                                         if (tr_on && tt < 1000 && (warp == 0 || warp == 4)) { g_trace[tt * 8 + 3] = pb; g_trace[tt * 8 + 4] = pe; g_trace[tt * 8 + 5] = warp; } })
                    }
                    if (act) ts_row[2 * rec] = part;   // this group's share of the record tile (0 if it owns none of its tiles)
                }
            }
            RG_PROF_ONLY(const long long pq1 = clock64(); if (tid == 0) pf[11] += (unsigned long long)(pq1 - pq0);)
            tc::tc_fence_before();
            cta_sync();
            tc::tc_fence_after();
            seq_use += (unsigned)nunit;
            build_om32();     // pass 2 operand, in the (now idle) ring bytes
            cta_sync();
            RG_PROF_ONLY(if (tid == 0) pf[14] += (unsigned long long)(clock64() - pq1);)
        } else {
            const bool warp_active = warp * 16 < nlive;
            float sh[8], part[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { sh[u] = s_shift[ty * 8 + u]; part[u] = 0.f; }
            int rec = 0, sub_in_rec = 0;
            for (int sub = 0; sub < nsub; ++sub) {
                const unsigned seq = seq_use + (unsigned)sub;
                const int stage = (int)(seq % (unsigned)nstage);
                mbar_wait(&mbar[stage], (seq / (unsigned)nstage) & 1u);
                if (warp_active) {
                    const float *gt = gtile + (size_t)stage * K1 * SUB;
                    float acc[8][8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
#pragma unroll 3
                    for (int k = 0; k < K1; ++k) {
                        const float4 w0 = *reinterpret_cast<const float4 *>(om32 + k * TU + ty * 8);
                        const float4 w1 = *reinterpret_cast<const float4 *>(om32 + k * TU + ty * 8 + 4);
                        const float4 g0 = *reinterpret_cast<const float4 *>(gt + k * SUB + tx * 4);
                        const float4 g1 = *reinterpret_cast<const float4 *>(gt + k * SUB + 64 + tx * 4);
                        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                        for (int u = 0; u < 8; ++u)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[u][j] = fmaf(w[u], g[j], acc[u][j]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int j = 0; j < 8; ++j) part[u] += ex2_approx(acc[u][j] - sh[u]);
                }
                if (++sub_in_rec == p.rec_sub || sub == nsub - 1) {
                    if (warp_active) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float v = part[u];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            if (tx == 0) tilesum[(size_t)(ty * 8 + u) * p.nrec_pad + rec] = v;
                            part[u] = 0.f;
                        }
                    }
                    ++rec;
                    sub_in_rec = 0;
                }
                cta_sync();  // everyone is done with this ring stage
                if (tid == 0) issue_load(seq + (unsigned)nstage - 1u);
            }
            seq_use += (unsigned)nsub;
        }

        RG_MARK(2);   // B: pass 1 (A tile, tiles, omega tile of pass 2)
        { const long long tk1 = clock64(); ph_b += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
        if constexpr (TC) {
            // ---- C: pass 2 (tensor mode).  Stage A, one thread per live user: view uniform, total mass, tau = u Z, the
            // record tile that contains tau (sequential binary64 sums of the fp32 partial sums).  Stage B, EIGHT LANES per
            // user (four users per warp at a time, 32 per CTA pass): the crossing 32-product GROUP inside the record from
            // the pass-1 group sums the epilogue kept, then the group's 32 masses recomputed in natural product order --
            // lane q owns products 4q .. 4q+3, one 16-byte load per k, so that the eight lanes of a user read one whole
            // 128-byte row segment of the Gamma tile per k (fully coalesced: a thread-private layout costs 8x the L1
            // wavefronts) -- fp32 FMA chains over k as in pass 1's error model -> index, CDF margins, filter.
            if (tid < nlive) {
                const int i = tid, slot = pos2slot[i];
                const u32x4 w = draw(env.seed, s_user[slot], s_t[slot], SLOT_EVENT);
                const double u = uniform53(w.x, w.y);
                const float *ts = tilesum + (size_t)i * p.nrec_pad;
                double Z = 0.0;
                for (int r = 0; r < p.nrec; ++r) Z += (double)ts[2 * r] + (double)ts[2 * r + 1];
                const double tau = u * Z;
                double cum = 0.0;
                int rsel = -1;
                for (int r = 0; r < p.nrec; ++r) {
                    const double nx = cum + ((double)ts[2 * r] + (double)ts[2 * r + 1]);
                    if (nx > tau) { rsel = r; break; }
                    cum = nx;
                }
                s_Z[i] = Z; s_tau[i] = tau; s_cum[i] = cum; s_rsel[i] = rsel; s_u[i] = u;
            }
            cta_sync();
            RG_MARK(3);   // C: pass 2, stage A
            const int sub = lane & 7, seg = lane & ~7;
            for (int ib = 0; ib < nlive; ib += NTH / 8) {
                const int i = ib + (tid >> 3);
                if (i >= nlive) continue;                       // whole 8-lane segments drop out
                const unsigned pm = 0xffu << seg;               // every collective below is over this user's eight lanes only
                const int rsel = s_rsel[i];
                const double tau = s_tau[i];
                double cum = s_cum[i];
                const int p0 = max(rsel, 0) * p.rec_sub * SUB;
                const int ngr = rsel < 0 ? 0 : (min(p0 + p.rec_sub * SUB, nsub * SUB) - p0) / 32;
                const float *gs = gsum_cta + (size_t)(p0 / 32) * TU + i;
                // crossing group: 32 groups per step, lane q holds the groups q, q+8, q+16, q+24 of the step
                int gsel = -1;
                double gmass = 0.0;
                for (int g0 = 0; g0 < ngr && gsel < 0; g0 += 32) {
                    float x[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) { const int g = g0 + 8 * b + sub; x[b] = g < ngr ? __ldcg(gs + (size_t)g * TU) : 0.f; }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (gsel >= 0 || g0 + 8 * b >= ngr) continue;          // uniform over the segment
                        const double own = (double)x[b];
                        double xs = own;
#pragma unroll
                        for (int d = 1; d < 8; d <<= 1) {
                            const double y = __shfl_up_sync(pm, xs, d, 8);
                            if (sub >= d) xs += y;
                        }
                        const unsigned m = (__ballot_sync(pm, cum + xs > tau && g0 + 8 * b + sub < ngr) >> seg) & 0xffu;
                        if (m) {
                            const int fl = __ffs(m) - 1;
                            gsel = g0 + 8 * b + fl;
                            cum += __shfl_sync(pm, xs - own, seg + fl);          // groups before the crossing one
                            gmass = __shfl_sync(pm, own, seg + fl);
                        } else {
                            cum += __shfl_sync(pm, xs, seg + 7);
                        }
                    }
                }
                // recompute the group: one float4 of Gamma per k and lane, FFMA chains over k
                const int pj0 = p0 + max(gsel, 0) * 32 + 4 * sub;
                const float *gp = env.G32 + (size_t)(pj0 >> 7) * K1 * SUB + (pj0 & (SUB - 1));
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 7
                for (int k = 0; k < K1; ++k) {
                    const float wk = om32[k * TU + i];
                    const float4 g = __ldg(reinterpret_cast<const float4 *>(gp + k * SUB));
                    a0 = fmaf(wk, g.x, a0); a1 = fmaf(wk, g.y, a1); a2 = fmaf(wk, g.z, a2); a3 = fmaf(wk, g.w, a3);
                }
                const float shf = s_shift[i];
                const double m0 = (double)ex2_approx(a0 - shf), m1 = (double)ex2_approx(a1 - shf);
                const double m2 = (double)ex2_approx(a2 - shf), m3 = (double)ex2_approx(a3 - shf);
                const double l0 = m0, l1 = l0 + m1, l2 = l1 + m2, l3 = l2 + m3;     // this lane's inclusive prefixes
                double xs = l3;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    const double y = __shfl_up_sync(pm, xs, d, 8);
                    if (sub >= d) xs += y;
                }
                const double gtot = __shfl_sync(pm, xs, seg + 7);       // recomputed mass of the whole group
                const double base = cum + (xs - l3);                    // mass before this lane's first product
                int fnd = -1;
                double lo = 0.0, up = 0.0;
                if (base + l0 > tau && pj0 < P) { fnd = pj0; lo = base; up = base + l0; }
                else if (base + l1 > tau && pj0 + 1 < P) { fnd = pj0 + 1; lo = base + l0; up = base + l1; }
                else if (base + l2 > tau && pj0 + 2 < P) { fnd = pj0 + 2; lo = base + l1; up = base + l2; }
                else if (base + l3 > tau && pj0 + 3 < P) { fnd = pj0 + 3; lo = base + l2; up = base + l3; }
                const unsigned mf = (__ballot_sync(pm, fnd >= 0) >> seg) & 0xffu;
                const int first = mf ? __ffs(mf) - 1 : 0;
                if (sub == first) {      // the lane that holds the crossing (lane 0 of the segment if there is none)
                    const int slot = pos2slot[i];
                    bool amb = gsel < 0 || mf == 0u;   // (a recomputed group lighter than its pass-1 sum by a rounding: ambiguous)
                    const double Z = s_Z[i], uu = s_u[i];
                    if (!amb) {
                        // Filter.  Every computed mass is within d (relative) of the exact one; the SAME pass-1 masses enter
                        // the prefix sum S^ and the total Z^, so their errors largely cancel in S^ - u Z^:
                        //   |(S^ - u Z^) - (S - u Z)| <= d [(1-u) S + u (Z - S)] + 2 u d m(G) + u g m(T)
                        // G = the crossing group, whose products are counted with their pass-1 value in Z^ and with their
                        // recomputed value in S^; T = the crossing record tile, whose earlier groups enter S^ as group sums
                        // and Z^ through the fp32 chain of the record sum: g = (adds of that chain) 2^-24.
                        const double dlt = (double)s_eps[i];
                        const float *ts = tilesum + (size_t)i * p.nrec_pad + 2 * rsel;
                        const double tmass = (double)ts[0] + (double)ts[1];
                        const double gm = fmax(gmass, gtot);
                        const double grec = (double)(p.rec_sub * 4 + 4) * 5.9604644775390625e-08 * tmass;
                        const double tl = dlt * ((1.0 - uu) * lo + uu * fmax(Z - lo, 0.0) + 2.0 * uu * gm) + uu * grec + 1e-10 * Z;
                        const double tu = dlt * ((1.0 - uu) * up + uu * fmax(Z - up, 0.0) + 2.0 * uu * gm) + uu * grec + 1e-10 * Z;
                        amb = (tau - lo < tl && fnd > 0) || (up - tau < tu && fnd < P - 1);
                    }
                    s_view[slot] = fnd;
                    if (amb) {   // queued for a re-sampling sweep; the slot sits out until its draw has been resolved
                        const double lse = log64(Z) * 1.4426950408889634074 + (double)shf;
                        const int e = atomicAdd(&misc[6], 1);
                        rq_slot[e] = slot; rq_u[e] = uu; rq_lse[e] = lse;
                        s_wait[slot] = 1;
                    }
                }
            }
            __syncwarp();
        }
        // ---- C: pass 2a (CUDA-core form), one thread per live user: view uniform, total mass, tau = u*Z, the record tile
        // that contains tau and the mass before it (sequential binary64 sums of the fp32 tile sums)
        if (!TC && tid < nlive) {
            const int i = tid, slot = pos2slot[i];
            const u32x4 w = draw(env.seed, s_user[slot], s_t[slot], SLOT_EVENT);
            const double u = uniform53(w.x, w.y);
            const float *ts = tilesum + (size_t)i * p.nrec_pad;
            double Z = 0.0;
            for (int r = 0; r < p.nrec; ++r) Z += (double)ts[r];
            const double tau = u * Z;
            double cum = 0.0;
            int rsel = -1;
            for (int r = 0; r < p.nrec; ++r) {
                const double nx = cum + (double)ts[r];
                if (nx > tau) { rsel = r; break; }
                cum = nx;
            }
            s_Z[i] = Z; s_tau[i] = tau; s_cum[i] = cum; s_rsel[i] = rsel; s_u[i] = u;
        }
        if (!TC) cta_sync();
        RG_MARK(3);   // C: pass 2a

        // ---- C: pass 2b (CUDA-core form), one warp per live user: recompute the located record tile, index and margins
        for (int i = warp; i < (TC ? 0 : nlive); i += NWARP) {
            const int slot = pos2slot[i];
            const double Z = s_Z[i], tau = s_tau[i];
            double cum = s_cum[i];
            const int rsel = s_rsel[i];
            int found = -1;
            bool amb = false;
            double lower = 0.0, upper = 0.0;
            if (rsel < 0) {
                amb = true;
            } else {
                const float shf = s_shift[i];
                const int p0 = rsel * p.rec_sub * SUB;
                const int p1 = min(p0 + p.rec_sub * SUB, nsub * SUB);
              {
                // two 128-product tiles per step: 8 masses per lane (8 independent FFMA chains keep 8 x K1 L2 loads
                // in flight), the 8 group totals by butterfly sums, a scan only in the crossing group
                bool done = false;
                for (int base = p0; base < p1 && !done; base += 2 * SUB) {
                    const bool two = base + SUB < p1;
                    const float *g0 = env.G32 + (size_t)(base >> 7) * K1 * SUB + lane;
                    const float *g1 = two ? g0 + (size_t)K1 * SUB : g0;
                    float a[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) a[q] = 0.f;
#pragma unroll 3
                    for (int k = 0; k < K1; ++k) {
                        const float wv = om32[k * TU + i];
                        const float *gk0 = g0 + k * SUB, *gk1 = g1 + k * SUB;
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { v[q] = __ldg(gk0 + q * 32); v[4 + q] = __ldg(gk1 + q * 32); }
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[q] = fmaf(wv, v[q], a[q]);
                    }
                    double xs[8], tq[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        xs[q] = (q < 4 || two) ? (double)ex2_approx(a[q] - shf) : 0.0;
                        tq[q] = xs[q];
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) tq[q] += __shfl_xor_sync(0xffffffffu, tq[q], d);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (done) break;
                        if (cum + tq[q] > tau) {  // the crossing is in this group of 32 (or a rounding-level tie)
                            double x = xs[q];
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const double y = __shfl_up_sync(0xffffffffu, x, d);
                                if (lane >= d) x += y;
                            }
                            const double pre = cum + x;
                            const int pj = base + q * 32 + lane;
                            const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (pj < P));
                            if (m) {
                                const int fl = __ffs(m) - 1;
                                found = base + q * 32 + fl;
                                upper = __shfl_sync(0xffffffffu, pre, fl);
                                const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
                                lower = fl > 0 ? below : cum;
                            }
                            done = true;  // found, or inconsistent rounding (found < 0 -> ambiguous)
                        } else {
                            cum += tq[q];
                        }
                    }
                }
              }
                if (found < 0) {
                    amb = true;
                } else {
                    // Filter.  Every computed mass is within d (relative) of the exact one; the SAME pass-1 masses
                    // enter both the prefix sum S^ and the total Z^, so their errors largely cancel in S^ - u Z^:
                    //   |(S^ - u Z^) - (S - u Z)| <= d [ (1-u) S + u (Z - S) + 2u m(T) ]
                    // where the last term covers the products of the crossing record tile that are counted once
                    // with their pass-1 value (in Z^) and once with their recomputed value (in S^).  At the
                    // crossing S ~ u Z, i.e. the bound is ~ 2u(1-u) d Z instead of the uncorrelated 2 d Z.
                    const double dlt = (double)s_eps[i], u = s_u[i];
                    const double tmass = (double)tilesum[(size_t)i * p.nrec_pad + rsel];
                    const double tl = dlt * ((1.0 - u) * lower + u * fmax(Z - lower, 0.0) + 2.0 * u * tmass) + 1e-10 * Z;
                    const double tu = dlt * ((1.0 - u) * upper + u * fmax(Z - upper, 0.0) + 2.0 * u * tmass) + 1e-10 * Z;
                    amb = (tau - lower < tl && found > 0) || (upper - tau < tu && found < P - 1);
                }
            }
            if (lane == 0) {
                s_view[slot] = found;
                if (amb) {
                    const double lse = log64(Z) * 1.4426950408889634074 + (double)s_shift[i];
                    // queued for a re-sampling sweep; the slot sits out until its draw has been resolved
                    const int e = atomicAdd(&misc[6], 1);
                    rq_slot[e] = slot; rq_u[e] = s_u[i]; rq_lse[e] = lse;
                    s_wait[slot] = 1;
                }
            }
        }
        cta_sync();

        RG_MARK(4);   // C: pass 2b
        { const long long tk1 = clock64(); ph_c += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
        // ---- C2: re-sampling of the draws the fp32 filter could not certify: tier 2 (binary64 logits, fp32 ex2,
        // a-posteriori bound), tier 3 (all binary64) for what tier 2 leaves open.  One sweep streams the whole binary64
        // Gamma (L2 -> SM) whatever the number of users in it, so uncertified draws WAIT in a queue (their slots sit the
        // rounds out) until FB of them share a sweep; in the tail of the launch (few live slots) they are resolved at once.
        {
            const bool tail = nlive < 16;
            for (int tier = 2; tier <= 3; ++tier) {
                int *qs = tier == 2 ? rq_slot : q3_slot;
                double *qu = tier == 2 ? rq_u : q3_u;
                for (;;) {
                    cta_sync();
                    const int n = tier == 2 ? misc[6] : misc[4];
                    if (n == 0 || (n < FB && !tail)) break;
                    const int nb = min(FB, n);
                    if (tid < nb) {
                        fb_slot[tid] = qs[tid];
                        fb_u[tid] = qu[tid];
                        fb_lse[tid] = tier == 2 ? rq_lse[tid] : 0.0;
                        fb_unres[tid] = 0;
                    }
                    cta_sync();
                    if (tier == 2)
                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 fb_om, tid, NTH, 0, ambig64);
                    else
                        block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view,
                                                  fb_unres, fb_om, tid, NTH, 0, ambig64);
                    if (tid == 0) {   // bookkeeping of <= FB results, then pop the front of the queue
                        for (int f = 0; f < nb; ++f) {
                            if (tier == 2 && fb_unres[f]) {
                                const int e3 = misc[4]++;
                                q3_slot[e3] = fb_slot[f]; q3_u[e3] = fb_u[f];
                                n_tier3 += 1;
                            } else {
                                s_wait[fb_slot[f]] = 0;   // its view is in s_view: advanced by D below
                            }
                        }
                        if (tier == 2) n_fallback += (unsigned long long)nb;
                        for (int e = nb; e < n; ++e) {
                            qs[e - nb] = qs[e]; qu[e - nb] = qu[e];
                            if (tier == 2) rq_lse[e - nb] = rq_lse[e];
                        }
                        if (tier == 2) misc[6] = n - nb; else misc[4] = n - nb;
                    }
                }
            }
            cta_sync();
        }

        if (TC && tid == 0) {
            // pass 2 and the re-sampling sweeps are done with the bytes they share with the ring:
            // start streaming the first tiles of the next round while the users are advanced
            for (int s2 = 0; s2 < tc_prologue; ++s2) issue_load(seq_use + (unsigned)s2);
        }

        RG_MARK(5);   // C2: re-sampling
        { const long long tk1 = clock64(); ph_c2 += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
        // ---- D: advance the users.  D1 (one thread per slot) finishes the organic event: bookkeeping of the
        // view, transition.  D2 (one warp per user) runs the bandit events: the draws of an event depend only
        // on (user, t), so lane j evaluates the event at time t+j assuming the run is still alive; the run ends
        // at the first lane whose transition leaves the bandit state or whose click fires.  omega updates that
        // precede an ORGANIC event are deferred to omega_phase(); with change_omega_for_bandits every event
        // changes omega, so that configuration keeps the serial per-thread loop.
        auto finish_user = [&](int slot, uint64_t user, uint64_t h, uint32_t t, int last_view, int n_org, int n_ban,
                               int n_clk) {  // episode over: write the user out, claim the next one
            const long long idx = (long long)user - p.user_begin;
            const double *om = om64 + (size_t)slot * KP;
            h = hash_fold(h, ((uint64_t)t << 2) | 3u);
            if (p.out.n_events) p.out.n_events[idx] = (int32_t)t;
            if (p.out.n_organic) p.out.n_organic[idx] = n_org;
            if (p.out.n_bandit) p.out.n_bandit[idx] = n_ban;
            if (p.out.n_clicks) p.out.n_clicks[idx] = n_clk;
            if (p.out.last_view) p.out.last_view[idx] = last_view;
            if (p.out.hash) p.out.hash[idx] = h;
            if (p.out.omega)
                for (int k = 0; k < K; ++k) p.out.omega[(size_t)idx * K + k] = om[k];
            tot_users += 1; tot_events += t; tot_org += n_org; tot_ban += n_ban; tot_clk += n_clk;
            tot_ph += idx < p.n_organic_users ? 0 : 1;
            if (p.out.log && idx >= p.n_organic_users) {  // phantom action after done (abstract.py:311-316)
                const int a = policy_act(slot, user, t, last_view, n_org);
                log_write(p.out, log_reserve(p.out, 1), idx, t, a, 1 | 4);
            }
            claim(slot);
        };
        const bool serial_bandit = env.change_omega_for_bandits != 0;
        if (tid < TU && s_live[tid] && s_wait[tid] == 0) {
            const int slot = tid;
            double *om = om64 + (size_t)slot * KP;
            const uint64_t user = s_user[slot];
            uint64_t h = s_hash[slot];
            uint32_t t = s_t[slot];
            int last_view, n_org = s_n_org[slot], n_ban = s_n_ban[slot], n_clk = s_n_clk[slot];
            const bool organic_only = (long long)user - p.user_begin < p.n_organic_users;
            {   // abstract.py:110-119: the view was drawn above; now update_state
                const int v = s_view[slot];
                last_view = v;
                n_org += 1;
                h = hash_fold(h, ((uint64_t)(uint32_t)v << 2) | 0u);
                if (p.out.view_hist) atomicAdd(p.out.view_hist + v, 1ull);
                if (p.out.log) log_write(p.out, log_reserve(p.out, 1), (long long)user - p.user_begin, t, v, 0);
                if (track_views && !organic_only) {
                    int nd = s_nview[slot];
                    views_insert(vlists + (size_t)slot * p.pol.view_cap, nd, p.pol.view_cap, v, p.pol.view_overflow);
                    s_nview[slot] = nd;
                }
            }
            const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
            int st = next_state(env, RG_ORGANIC, uniform53(w.z, w.w));
            t += 1;
            // NormalTimeGenerator: the user's accumulated clock lives in shared memory (no registers on the default clock)
            auto tick = [&](uint32_t t_new) -> double {
                if (!tmode) return 1.0;
                double clock = s_clock[slot];
                const double dt = clock_advance(1, env.seed, user, t_new, env.time_mu, env.time_sigma, clock);
                s_clock[slot] = clock;
                s_dt[slot] = dt;
                return dt;
            };
            double omega_k = tick(t);
            int pend = 0;
            if (st == RG_ORGANIC) pend = 1;
            else if (env.change_omega_for_bandits) thread_omega_update(om, K, env.sigma * omega_k, false, env.seed, user, t);
            while (serial_bandit && st == RG_BANDIT && !organic_only) {  // abstract.py:174-181, one event per turn
                const int a = policy_act(slot, user, t, last_view, n_org);
                const u32x4 wb = draw(env.seed, user, t, SLOT_EVENT);
                const double uc = uniform53(wb.x, wb.y);
                int c = 0;
                if (uc >= 0.9706) c = click_from_uniform(ctr_of_action(env, om, a), uc);
                st = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                if (p.out.log) log_write(p.out, log_reserve(p.out, 1), (long long)user - p.user_begin, t, a, 1 | (c << 1));
                t += 1;
                omega_k = tick(t);
                if (st == RG_ORGANIC) pend = 1;  // the loop ends here: deferred
                else thread_omega_update(om, K, env.sigma * omega_k, false, env.seed, user, t);
                if (c) st = RG_ORGANIC;
                n_ban += 1;
                n_clk += c;
                h = hash_fold(h, ((uint64_t)(uint32_t)a << 2) | ((uint64_t)c << 1) | 1u);
                if (p.out.action_hist) atomicAdd(p.out.action_hist + a, 1ull);
                if (c && p.out.click_hist) atomicAdd(p.out.click_hist + a, 1ull);
            }
            if (st == RG_STOP || (st == RG_BANDIT && organic_only)) {
                finish_user(slot, user, h, t, last_view, n_org, n_ban, n_clk);
            } else {
                s_hash[slot] = h; s_t[slot] = t; s_last_view[slot] = last_view;
                s_n_org[slot] = n_org; s_n_ban[slot] = n_ban; s_n_clk[slot] = n_clk;
                s_pend[slot] = pend;
                s_state[slot] = st;  // ORGANIC, or BANDIT: the run is played by D2
            }
        }
        cta_sync();
        RG_MARK(6);   // D1: view bookkeeping, transition
        if (!serial_bandit) {
            if (warp == 0) {  // list of the slots that enter a bandit run (amb_list is free at this point)
                int base = 0;
                for (int c = 0; c < TU / 32; ++c) {
                    const int slot = c * 32 + lane;
                    const bool bd = s_live[slot] != 0 && s_state[slot] == RG_BANDIT;
                    const unsigned m = __ballot_sync(0xffffffffu, bd);
                    if (bd) amb_list[base + __popc(m & ((1u << lane) - 1u))] = slot;
                    base += __popc(m);
                }
                if (lane == 0) misc[5] = base;
            }
            cta_sync();
            const int nban = misc[5];
            for (int b = warp; b < nban; b += NWARP) {
                const int slot = amb_list[b];
                const double *om = om64 + (size_t)slot * KP;
                const uint64_t user = s_user[slot];
                uint64_t h = s_hash[slot];
                uint32_t t = s_t[slot];
                const int last_view = s_last_view[slot];
                int n_ban = s_n_ban[slot], n_clk = s_n_clk[slot];
                int st = RG_BANDIT, pend = 0;
                while (st == RG_BANDIT) {  // 32 candidate events per turn (a run is ~14 events on average)
                    const uint32_t tt = t + (uint32_t)lane;
                    const int a = policy_act(slot, user, tt, last_view, s_n_org[slot]);
                    const u32x4 wb = draw(env.seed, user, tt, SLOT_EVENT);
                    const double uc = uniform53(wb.x, wb.y);
                    // ff() never exceeds sig(5*0.5-6) = 0.02931, so the click threshold (1-ctr)/((1-ctr)+ctr) is
                    // above 0.9706: below that the click is 0 whatever omega is (97% of the bandit events)
                    // The few candidate clicks are evaluated ONE AT A TIME BY THE WHOLE WARP, in event order, and only up to
                    // the event whose transition ends the run: lane k fetches beta[a][k] and omega[k] (two coalesced loads
                    // instead of a chain of 2K dependent L2 round trips in one lane), the dot product is then accumulated in
                    // the oracle's order (sequential fma over k, ctr_of_action) from shuffled operands.
                    int c = 0;
                    const int nst = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                    const unsigned mleave = __ballot_sync(0xffffffffu, nst != RG_BANDIT);
                    const int nev_t = mleave ? __ffs(mleave) : 32;   // events up to the first transition out of the bandit state
                    unsigned mcand = __ballot_sync(0xffffffffu, uc >= 0.9706 && lane < nev_t);
                    while (mcand) {   // warp-uniform
                        const int j = __ffs(mcand) - 1;
                        mcand &= mcand - 1u;
                        const int aj = __shfl_sync(0xffffffffu, a, j);
                        const double *row = env.beta + (size_t)aj * K;
                        double acc = 0.0;
                        for (int k0 = 0; k0 < K; k0 += 32) {
                            const int k = k0 + lane;
                            const double bk = k < K ? __ldg(row + k) : 0.0;
                            const double wk = k < K ? om[k] : 0.0;
                            const int kn = min(32, K - k0);
                            for (int kk = 0; kk < kn; ++kk)
                                acc = fma(__shfl_sync(0xffffffffu, bk, kk), __shfl_sync(0xffffffffu, wk, kk), acc);
                        }
                        const int cj = click_from_uniform(ff64(acc + __ldg(env.mu_b + aj)), __shfl_sync(0xffffffffu, uc, j));
                        if (cj) {     // the run ends with this click: later candidates never happen
                            if (lane == j) c = 1;
                            break;
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, nst != RG_BANDIT || c != 0);
                    const int nev = m ? __ffs(m) : 32;  // events of this turn that really happen
                    const uint64_t rec = ((uint64_t)(uint32_t)a << 2) | ((uint64_t)c << 1) | 1u;
                    for (int j = 0; j < nev; ++j) h = hash_fold(h, __shfl_sync(0xffffffffu, rec, j));
                    if (p.out.log) {  // one reservation per turn, lane j writes the row of its event
                        unsigned long long row0 = 0;
                        if (lane == 0) row0 = log_reserve(p.out, nev);
                        row0 = __shfl_sync(0xffffffffu, row0, 0);
                        if (lane < nev)
                            log_write(p.out, row0 + (unsigned long long)lane, (long long)user - p.user_begin, tt, a, 1 | (c << 1));
                    }
                    if (lane < nev) {
                        if (p.out.action_hist) atomicAdd(p.out.action_hist + a, 1ull);
                        if (c && p.out.click_hist) atomicAdd(p.out.click_hist + a, 1ull);
                    }
                    if (tmode) {   // every event of the run advances the clock: the steps in parallel, their SEQUENTIAL sum
                        const double stp = lane < nev ? time_step(env.seed, user, tt, env.time_mu, env.time_sigma) : 0.0;
                        double clock = s_clock[slot], omega_k = 1.0;
                        for (int j = 0; j < nev; ++j) {
                            const double prev = clock;
                            clock = clock + __shfl_sync(0xffffffffu, stp, j);
                            const double d = clock - prev;
                            omega_k = d == 0.0 ? 1.0 : d;
                        }
                        __syncwarp();
                        if (lane == 0) { s_clock[slot] = clock; s_dt[slot] = omega_k; }   // s_dt: read by the omega phase if the run ends organic
                        __syncwarp();
                    }
                    n_ban += nev;
                    t += (uint32_t)nev;
                    if (m) {
                        const int cf = __shfl_sync(0xffffffffu, c, nev - 1);
                        const int nstf = __shfl_sync(0xffffffffu, nst, nev - 1);
                        n_clk += cf;
                        pend = nstf == RG_ORGANIC ? 1 : 0;   // drift only if the TRANSITION said organic
                        st = cf ? RG_ORGANIC : nstf;          // abstract.py:180-181: a click overrides it
                    }
                }
                if (lane == 0) {
                    if (st == RG_ORGANIC) {
                        s_hash[slot] = h; s_t[slot] = t; s_n_ban[slot] = n_ban; s_n_clk[slot] = n_clk;
                        s_pend[slot] = pend;
                        s_state[slot] = RG_ORGANIC;
                    } else {
                        finish_user(slot, user, h, t, last_view, s_n_org[slot], n_ban, n_clk);
                    }
                }
                __syncwarp();
            }
        }
        cta_sync();
        RG_MARK(7);   // D2: bandit runs
        { const long long tk1 = clock64(); ph_d += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
        omega_phase();
        cta_sync();
        { const long long tk1 = clock64(); ph_o += (unsigned long long)(tk1 - tk0); tk0 = tk1; }
    }

    // ---- drain the prefetched tiles (a CTA must not exit with bulk copies in flight)
    for (int s = 0; s < (TC ? tc_prologue : nstage - 1); ++s) {
        const unsigned seq = seq_use + (unsigned)s;
        mbar_wait(&mbar[seq % (unsigned)nstage], (seq / (unsigned)nstage) & 1u);
    }
    if (TC) {
        tc::tc_fence_before();
        cta_sync();
        if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
    }

    // ---- hand the user queue back clean: the last CTA to get here zeroes the ticket and the arrival counter
    if (tid == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(p.queue + 1, 1ull);
        if (done == (unsigned long long)gridDim.x - 1ull) {
            p.queue[0] = 0ull;
            p.queue[1] = 0ull;
            __threadfence();
        }
    }

    RG_PROF_ONLY(if (lane == 0) { for (int i = 0; i < 16; ++i) if (pf[i]) atomicAdd(&g_prof[i], pf[i]); })
    RG_PROF_ONLY(if (tid == 0) { for (int i = 0; i < 12; ++i) if (pf2[i]) atomicAdd(&g_prof[48 + i], pf2[i]); })
    // ---- totals: warp reduction, one atomic per counter per warp
    if (p.out.totals) {
        if (tid == 0) {   // SM cycles thread 0 of this CTA spent in each phase (diagnostic, summed over the CTAs)
            atomicAdd(p.out.totals + 10, ph_a); atomicAdd(p.out.totals + 11, ph_b); atomicAdd(p.out.totals + 12, ph_c);
            atomicAdd(p.out.totals + 13, ph_c2); atomicAdd(p.out.totals + 14, ph_d); atomicAdd(p.out.totals + 15, ph_o);
        }
        unsigned long long v[9] = {tot_users, tot_events, tot_org, tot_ban, tot_clk, tot_ph, n_fallback,
                                   (unsigned long long)ambig64, n_tier3};
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            unsigned long long x = v[c];
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            if (lane == 0 && x) atomicAdd(p.out.totals + c, x);
        }
    }
}

// CUDA-core form: 256 threads, two CTAs per SM at 128 registers.  Tensor form: 288 threads (the ninth warp is the MMA
// issuer), two CTAs per SM at 96 registers (the register file is per scheduler: 5 warps x 32 x 96 regs fit its 16 K, 5 x 112 do not).
__global__ void __launch_bounds__(NT, 2) rollout_cta_kernel_cc(const Params p) { rollout_cta_body<false>(p); }
__global__ void __launch_bounds__(NT + 32, 2) rollout_cta_kernel_tc(const Params p) { rollout_cta_body<true>(p); }

}  // namespace fast

#ifdef RG_PROF
extern "C" int rg_prof_read(unsigned long long *prof64, long long *trace8192, int reset) {
    if (prof64) cudaMemcpyFromSymbol(prof64, fast::g_prof, sizeof(fast::g_prof));
    if (trace8192) cudaMemcpyFromSymbol(trace8192, fast::g_trace, sizeof(fast::g_trace));
    if (reset) {
        static unsigned long long z[64];
        cudaMemcpyToSymbol(fast::g_prof, z, sizeof(z));
    }
    return (int)cudaGetLastError();
}
#endif

cudaError_t launch_rollout_fast(rg_env_impl &e, const PolicyDev &pol, long long user_begin,
                                long long n_organic_users, long long n_total, const RolloutOutDev &out,
                                cudaStream_t stream, bool tensor) {
    using namespace fast;
    if (n_total <= 0) return cudaSuccess;
    // parameters outside the range of the UMMA operand elements: same kernel with the CUDA-core contraction
    // (identical discrete results)
    if (tensor && !e.tc_operands_ok) return launch_rollout_fast(e, pol, user_begin, n_organic_users, n_total, out, stream, false);
    Params p;
    p.env = e.dev;
    p.pol = pol;
    p.user_begin = user_begin;
    p.n_organic_users = n_organic_users;
    p.n_total = n_total;
    p.out = out;
    p.tc = tensor ? 1 : 0;
    p.nhalf = tensor ? 2 : 1;
    p.om_scratch = nullptr;
    p.csum_scratch = nullptr;
    p.spill = 0;
    p.gsum = nullptr;
    const int K1 = e.dev.K + 1;
    p.kp = e.dev.K | 1;  // odd row stride of the binary64 omega rows: conflict-free per-slot access
    p.ncs = (e.dev.Ppad + CHUNK - 1) / CHUNK + 1;
    const int nsub = e.dev.nsub;
    // record tiles: at most 64 partial sums per user, at least 2 sub-tiles each
    int rec_sub = 2;
    static const int rec_max_env = [] { const char *v = getenv("RG_REC_MAX"); const int x = v ? atoi(v) : 0; return x > 0 ? x : 0; }();
    const int rec_max = rec_max_env ? rec_max_env : (tensor ? 20 : 64);   // tensor mode keeps two partial sums per record (same bytes as 40 records)
    while ((nsub + rec_sub - 1) / rec_sub > rec_max) ++rec_sub;
    if (rec_sub > nsub) rec_sub = nsub;
    p.rec_sub = rec_sub;
    p.nrec = (nsub + rec_sub - 1) / rec_sub;
    p.nrec_pad = (p.nrec * p.nhalf) | 1;
    // ring depth: as deep as shared memory allows (2 CTAs/SM when possible)
    const int hi_stage = tensor ? kMaxStages : 3, lo_stage = tensor ? e.dev.nslab : 2;
    const int two_ctas = 113 * 1024;  // (228 KB - 1 KB reserved per CTA) / 2
    int best_stage = 0;
    // (CUDA-core form) when P x K is too large for one CTA's shared memory, the partial sums of the re-sampling sweeps,
    // then also the binary64 omega rows, move to an L2-resident global scratch (C4: P = 100 000, K = 64)
    static const bool force_one = [] { const char *v = getenv("RG_FORCE_ONE_CTA"); return v && v[0] == '1'; }();   // experiment knob
    for (int spill = 0; spill <= (tensor ? 0 : 2) && !best_stage; ++spill)
    for (int pass = force_one ? 1 : 0; pass < 2 && !best_stage; ++pass) {
        p.spill = spill;
        const int limit = pass == 0 ? two_ctas : 227 * 1024;
        // (asynchronous re-sampling on idle warps underneath pass 1 was measured SLOWER at C3 -- 3.2e8 vs 4.1e8 steps/s: three
        // warps take longer over the sweep than pass 1 lasts -- and has been removed)
        for (int nstage = hi_stage; nstage >= lo_stage && !best_stage; --nstage) {
            if (tensor && nstage % e.dev.nslab != 0) continue;   // the B ring holds whole product tiles (all K slabs)
            p.nstage = nstage;
            make_layout(p);
            if (p.L.total <= limit) best_stage = nstage;
        }
    }
    if (!best_stage && tensor) {
        // the UMMA operand tiles (2 x 2 x KC x 2 KB per stage + the A tile) do not fit for this K:
        // same kernel with the CUDA-core contraction (identical discrete results)
        return launch_rollout_fast(e, pol, user_begin, n_organic_users, n_total, out, stream, false);
    }
    if (!best_stage) return cudaErrorInvalidConfiguration;  // K/P too large for the CTA-batched kernel
    p.nstage = best_stage;
    make_layout(p);
    (void)K1;
    static thread_local unsigned launch_seq = 0;
    // queue[0] = next user ticket, queue[1] = CTAs finished; the last CTA to finish zeroes both, so a
    // launch never needs a memset (and profiler kernel replays see a clean queue).  30 pairs in rotation
    // keep launches on different streams apart.
    p.queue = e.diag + 4 + 2 * (launch_seq++ % 30u);
    cudaError_t ce = tensor ? cudaFuncSetAttribute(rollout_cta_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, p.L.total)
                            : cudaFuncSetAttribute(rollout_cta_kernel_cc, cudaFuncAttributeMaxDynamicSharedMemorySize, p.L.total);
    if (ce != cudaSuccess) return ce;
    int ctas_per_sm = p.L.total <= two_ctas ? 2 : 1;
    if (force_one && ctas_per_sm == 2) {   // deepest ring that fits, padded so that only one CTA is resident per SM
        ctas_per_sm = 1;
        p.L.total = two_ctas + 1024;
        ce = tensor ? cudaFuncSetAttribute(rollout_cta_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, p.L.total)
                    : cudaFuncSetAttribute(rollout_cta_kernel_cc, cudaFuncAttributeMaxDynamicSharedMemorySize, p.L.total);
        if (ce != cudaSuccess) return ce;
    }
    long long blocks = (n_total + TU - 1) / TU;
    const long long max_blocks = (long long)e.num_sms * ctas_per_sm;
    static const bool dbg_launch = [] { const char *v = getenv("RG_DEBUG_LAUNCH"); return v && v[0] == '1'; }();
    if (dbg_launch) {
        static thread_local int printed = 0;
        if (printed++ < 4)
            fprintf(stderr, "[rg] rollout_cta %s: P=%d K=%d nstage=%d nslab=%d rec_sub=%d nrec=%d smem=%d B ctas/SM=%d\n",
                    tensor ? "tensor" : "cuda-core", e.dev.P, e.dev.K, p.nstage, e.dev.nslab, p.rec_sub, p.nrec, p.L.total, ctas_per_sm);
    }
    if (blocks > max_blocks) blocks = max_blocks;
    if (p.spill >= 1) {
        const size_t need_c = sizeof(double) * (size_t)max_blocks * (size_t)p.ncs * FB * 2;
        if (e.csum_scratch_bytes < need_c) {
            if (e.d_csum_scratch) cudaFree(e.d_csum_scratch);
            e.d_csum_scratch = nullptr;
            e.csum_scratch_bytes = 0;
            ce = cudaMalloc((void **)&e.d_csum_scratch, need_c);
            if (ce != cudaSuccess) return ce;
            e.csum_scratch_bytes = need_c;
        }
        p.csum_scratch = e.d_csum_scratch;
    }
    if (tensor || p.spill >= 2) {
        const size_t need = sizeof(double) * (size_t)max_blocks * TU * p.kp;
        if (e.om_scratch_bytes < need) {   // binary64 omega rows of every resident slot (grown once per env)
            if (e.d_om_scratch) cudaFree(e.d_om_scratch);
            e.d_om_scratch = nullptr;
            e.om_scratch_bytes = 0;
            ce = cudaMalloc((void **)&e.d_om_scratch, need);
            if (ce != cudaSuccess) return ce;
            e.om_scratch_bytes = need;
        }
        p.om_scratch = e.d_om_scratch;
    }
    // longest-first schedule for launches that fill the machine several times over (see lpt_estimate_kernel)
    static const bool no_lpt = [] { const char *v = getenv("RG_NO_LPT"); return v && v[0] == '1'; }();
    char *lpt_arena = nullptr;
    p.perm = nullptr;
    if (!no_lpt && n_total >= 4 * max_blocks * TU && n_total < 2000000000ll) {
        const size_t o_est = 2048, o_perm = (o_est + (size_t)n_total + 255) & ~(size_t)255;
        ce = cudaMallocAsync((void **)&lpt_arena, o_perm + 4 * (size_t)n_total, stream);
        if (ce != cudaSuccess) return ce;
        unsigned *hist = (unsigned *)lpt_arena;
        unsigned char *est = (unsigned char *)(lpt_arena + o_est);
        int32_t *perm = (int32_t *)(lpt_arena + o_perm);
        cudaMemsetAsync(hist, 0, 1024, stream);
        const unsigned nb = (unsigned)((n_total + 255) / 256);
        lpt_ctr_kernel<<<1, 256, 0, stream>>>(e.dev, pol, (double *)(lpt_arena + 1024));
        lpt_estimate_kernel<<<nb, 256, 0, stream>>>(e.dev, user_begin, n_organic_users, n_total,
                                                    (const double *)(lpt_arena + 1024), est, hist);
        lpt_cursor_kernel<<<1, 32, 0, stream>>>(hist);
        lpt_scatter_kernel<<<nb, 256, 0, stream>>>(est, n_total, hist, perm);
        p.perm = perm;
    }
    if (tensor) {
        const size_t need_g = sizeof(float) * (size_t)max_blocks * (size_t)e.dev.nsub * 4 * TU;
        if (e.gsum_bytes < need_g) {       // pass-1 group sums of every resident slot (48 MB at C3: L2-resident)
            if (e.d_gsum) cudaFree(e.d_gsum);
            e.d_gsum = nullptr;
            e.gsum_bytes = 0;
            ce = cudaMalloc((void **)&e.d_gsum, need_g);
            if (ce != cudaSuccess) return ce;
            e.gsum_bytes = need_g;
        }
        p.gsum = e.d_gsum;
        // Experiment knob (RG_L2_PERSIST=1): mark the group-sum scratch as L2-persisting for this launch.  Every CTA rewrites
        // its 162 KB of it each round; with two CTAs per SM the 48 MB do not survive in L2 between rounds and are written
        // back (28 GB of DRAM writes per 1M-user launch at C3, profiles/r02_roofline_traffic.json) -- 4 % of the HBM
        // bandwidth, no effect on the kernel's speed.
        static const bool l2_persist = [] { const char *v = getenv("RG_L2_PERSIST"); return v && v[0] == '1'; }();
        bool window_set = false;
        if (l2_persist) {
            int max_win = 0, max_persist = 0;
            cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e.device);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e.device);
            if (max_win > 0 && max_persist > 0) {
                const size_t want = need_g < (size_t)max_persist ? need_g : (size_t)max_persist;
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                cudaStreamAttrValue av;
                std::memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = e.d_gsum;
                av.accessPolicyWindow.num_bytes = need_g < (size_t)max_win ? need_g : (size_t)max_win;
                av.accessPolicyWindow.hitRatio = want >= need_g ? 1.0f : (float)((double)want / (double)need_g);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                window_set = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                if (!window_set) (void)cudaGetLastError();
            }
        }
        rollout_cta_kernel_tc<<<(unsigned)blocks, NT + 32, (size_t)p.L.total, stream>>>(p);
        if (window_set) {
            cudaStreamAttrValue av;
            std::memset(&av, 0, sizeof(av));
            cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
        }
    } else {
        rollout_cta_kernel_cc<<<(unsigned)blocks, NT, (size_t)p.L.total, stream>>>(p);
    }
    ce = cudaGetLastError();
    if (lpt_arena) cudaFreeAsync(lpt_arena, stream);
    return ce;
}

}  // namespace rg
