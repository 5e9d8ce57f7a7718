// rg_device.cuh -- device-side data layout and warp-level building blocks shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include "rg_math.cuh"
#include "../../include/recogym_b200.h"

namespace rg {

constexpr int kMaxK = 128;          // latent dimension limit of the kernels
constexpr int kSuperChunk = 1024;   // products per recorded partial sum in the exact sampler

// k-steps (of 4) of the DMMA sweep template instantiated for K (rg_rollout_fast.cu, block_sample_views)
__host__ __device__ inline int dmma_ksteps(int K) { return K <= 20 ? 5 : (K <= 32 ? 8 : (K <= 64 ? 16 : 32)); }

// Static parameters as the kernels see them (all pointers are device memory owned by rg_env).
struct EnvDev {
    int P, K, Ppad;            // Ppad = P rounded up to 32
    const double *GammaT;      // [K][Ppad]  K-major f64: lane-consecutive products are contiguous
    const double *GammaB;      // [ceil(P/256)*32][KQ][8] f64: the (8 products x KQ) operand block of a DMMA product group is
                               // contiguous (KQ = 4 * dmma_ksteps(K), zero rows; zero products beyond P): tier-2 sweeps
    int KQ;
    const double *mu_o;        // [Ppad]
    const double *beta;        // [P][K] row-major f64 (one row gathered per bandit event)
    const double *mu_b;        // [P]
    const float *GammaT32;     // [K][Ppad] fp32 copy (RG_MATH_FAST)
    const float *mu_o32;       // [Ppad]
    double thr[3][2];          // numpy-choice thresholds of each transition row: cdf0/cdf2, cdf1/cdf2
    double sigma0, sigma;
    double gmax, mumax;        // max_p |Gamma_p|_2 and max_p mu_organic[p]: logit upper bound
    double muabsmax;           // max_p |mu_organic[p]|
    const float *G32;          // [nsub][K+1][128] fp32 tiles of log2(e)*Gamma (row K = log2(e)*mu_organic)
    int nsub;                  // number of 128-product tiles
    const void *GTC;           // [tiles][nslab][2][KCs][TILE_N/8][8][EPC] hi/lo UMMA operand tiles in K slabs (rg_tensor.cuh)
    int KC, KCs, nslab;        // 16-byte K chunks per operand row: padded total, per slab, slabs (KC = KCs * nslab)
    int change_omega_for_bandits;
    int time_mode;             // 0 default clock, 1 NormalTimeGenerator (include/recogym_b200.h rg_env_desc.time_generator)
    double time_mu, time_sigma;
    uint64_t seed;
    // tier 4 (oracle-order re-draw of the draws no binary64 summation order can separate, warp_sample_view_oracle_order)
    double *t4_buf;            // [kT4Slots][Ppad] scratch of masses, one slot per concurrent re-draw
    int *t4_lock;              // [kT4Slots] spin locks (0 free)
    int force_t4;              // test hook (rg_debug_force_tier4): every binary64 draw takes the oracle-order path
};
constexpr int kT4Slots = 32;

struct PolicyDev {
    int kind;
    const int32_t *table;  // device [P]
    // RG_POLICY_USER_VIEWS: one list of (product, views) pairs, sorted by product, per resident user
    // (ViewsFeaturesProvider of the reference, agents/abstract.py:318-409, kept sparse)
    int2 *views;           // [lists][view_cap]
    int view_cap;
    unsigned long long *view_overflow;  // counts users whose distinct views exceeded view_cap (results invalid)
    // RG_POLICY_EPS_TABLE: coin threshold eps / (eps + (1 - eps)) and the hole-CDF table of the exploring draw
    double eps_c0;
    const double *hole_cdf;             // [P]: S[m] = sequential sum of m copies of 1 / (P - 1) (oracle.hole_cdf_table)
};

// EpsilonGreedy.act around a table agent (agents/epsilon_greedy.py:30-68; oracle/rg_oracle.c eps_table_act): the coin on
// the first uniform of the policy stream, the exploring action by numpy's choice(P, p) on the second one, with
// p = ones(P), p[greedy] = 0, p /= sum.  The cdf of that p is cdf[j] = S[j + 1] for j < greedy and S[j] for j >= greedy
// (adding the hole's exact 0 changes nothing), so the inverse cdf is a binary search in ONE table per P.
__device__ __forceinline__ int eps_table_action(const PolicyDev &pol, int P, int last_view, const u32x4 &w) {
    const int g = __ldg(pol.table + last_view);
    if (pol.eps_c0 <= uniform53(w.x, w.y)) return g;          // searchsorted(right) == 1 -> False -> exploit
    const double u = uniform53(w.z, w.w), last = __ldg(pol.hole_cdf + P - 1);
    int lo = 0, hi = P;                                        // smallest j with cdf[j] / last > u
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double cdf = __ldg(pol.hole_cdf + (mid < g ? mid + 1 : mid));
        if (cdf / last <= u) lo = mid + 1; else hi = mid;
    }
    return lo < P ? lo : P - 1;
}

// ViewsFeaturesProvider.observe (agents/abstract.py:340-351): views[0, v] += 1 on the sparse sorted list.
__device__ __forceinline__ void views_insert(int2 *list, int &d, int cap, int v, unsigned long long *overflow) {
    int lo = 0, hi = d;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid].x < v) lo = mid + 1; else hi = mid;
    }
    if (lo < d && list[lo].x == v) { list[lo].y += 1; return; }
    if (d >= cap) { atomicAdd(overflow, 1ull); return; }
    for (int k = d; k > lo; --k) list[k] = list[k - 1];
    list[lo] = make_int2(v, 1);
    d += 1;
}

// OrganicUserEventCounterModel.act with exploit_explore, epsilon = 0, select_randomly
// (agents/organic_user_count.py:49-75): action_proba = features / sum(features); numpy choice(P, p) on ONE
// uniform = cumsum(p) (sequential; products never viewed add an exact 0), /= cdf[-1], searchsorted right.
// n = total views of the user so far, d = distinct products.
__device__ __forceinline__ int views_action(const int2 *list, int d, int n, double u) {
    const double dn = (double)n;
    double T = 0.0;
    for (int k = 0; k < d; ++k) T += (double)list[k].y / dn;
    double S = 0.0;
    for (int k = 0; k < d; ++k) {
        S += (double)list[k].y / dn;
        if (S / T > u) return list[k].x;
    }
    return list[d - 1].x;  // not reached: cdf[-1] / cdf[-1] == 1 > u
}

struct RolloutOutDev {
    int32_t *n_events, *n_organic, *n_bandit, *n_clicks, *last_view;
    uint64_t *hash;
    double *omega;
    unsigned long long *view_hist, *action_hist, *click_hist, *totals;
    int4 *log;                     // event log records {u, t, x, flags}
    long long log_capacity;
    unsigned long long *log_count;
};

// append n consecutive log rows: returns the first row index (rows >= capacity must not be written)
__device__ __forceinline__ unsigned long long log_reserve(const RolloutOutDev &out, int n) {
    return atomicAdd(out.log_count, (unsigned long long)n);
}
__device__ __forceinline__ void log_write(const RolloutOutDev &out, unsigned long long row, long long u, uint32_t t,
                                          int x, int flags) {
    if (row < (unsigned long long)out.log_capacity) out.log[row] = make_int4((int)u, (int)t, x, flags);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// omega init (t=0) or drift into time t: omega[k] = base[k] + scale * z_k(user, t).
// reco_env_v1.py:80-82 (base 0, scale sigma_omega_initial) and :96-99 (base omega, scale sigma_omega*1).
// `omega` lives in shared memory of the calling warp; all 32 lanes participate.
__device__ __forceinline__ void warp_omega_update(double *omega, int K, double scale, bool init, uint64_t seed,
                                                  uint64_t user, uint32_t t, int lane) {
    for (int j = lane; 2 * j < K; j += 32) {
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
    __syncwarp();
}

// numpy RandomState.choice(3, p=row) on uniform u (reco_env_v1.py:87): searchsorted(cdf, u, 'right')
__device__ __forceinline__ int next_state(const EnvDev &env, int state, double u) {
    return (env.thr[state][0] <= u ? 1 : 0) + (env.thr[state][1] <= u ? 1 : 0);
}

// Personalised CTR of action a (reco_env_v1.py:106-111); sequential fma chain over k.
__device__ __forceinline__ double ctr_of_action(const EnvDev &env, const double *omega, int a) {
    const double *row = env.beta + (size_t)a * env.K;
    double acc = 0.0;
    for (int k = 0; k < env.K; ++k) acc = fma(__ldg(row + k), omega[k], acc);
    return ff64(acc + __ldg(env.mu_b + a));
}

// numpy's DOUBLE_pairwise_sum (contiguous input; numpy/core/src/umath/loops_utils.h) by ONE thread: blocks of at most
// 128 elements are summed with 8 strided accumulators, larger ranges are split at n/2 rounded down to a multiple of 8.
// The recursion is unrolled on an explicit stack (depth <= log2(n / 128) + 2).
__device__ inline double np_pairwise_leaf(const double *a, long long n) {
    if (n < 8) {
        double res = 0.0;
        for (long long i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    long long i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
__device__ inline double np_pairwise_sum_dev(const double *a, int n) {
    int off[24], len[24];      // depth <= log2(n / 128) + 2 <= 24 for n < 2^31 / 128
    double left[24];
    int stage[24];
    int sp = 0;
    double ret = 0.0;
    off[0] = 0; len[0] = n; stage[0] = 0; sp = 1;
    while (sp > 0) {
        const int f = sp - 1;
        if (len[f] <= 128) { ret = np_pairwise_leaf(a + off[f], len[f]); --sp; continue; }
        int n2 = len[f] / 2;
        n2 -= n2 % 8;
        if (stage[f] == 0) {
            stage[f] = 1;
            off[sp] = off[f]; len[sp] = n2; stage[sp] = 0; ++sp;
        } else if (stage[f] == 1) {
            left[f] = ret;
            stage[f] = 2;
            off[sp] = off[f] + n2; len[sp] = len[f] - n2; stage[sp] = 0; ++sp;
        } else {
            ret = left[f] + ret;
            --sp;
        }
    }
    return ret;
}

// TIER 4: the organic view of ONE user by ONE warp in the ORACLE'S OWN operation order (oracle/rg_oracle.c
// user_product_view, i.e. reco_env_v1.py:119-128 + numpy's choice): logits by sequential fma chains, shift by their
// MAXIMUM, exp, s = numpy pairwise sum, cdf = sequential cumsum of e/s, cdf /= cdf[-1], first index with cdf > u.
// Used for the draws whose uniform lies so close to a CDF boundary that the summation order of the parallel samplers
// could decide the index (about 1 per 10^7 draws): with it every discrete output equals the oracle by construction.
// The masses go through a global scratch slot (env.t4_buf) guarded by a spin lock; the sequential part runs on lane 0.
static __device__ __noinline__ int warp_sample_view_oracle_order(const EnvDev &env, const double *omega, double u, int lane) {
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int slot = (int)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kT4Slots);
    double *e = env.t4_buf + (size_t)slot * Ppad;
    if (lane == 0) {
        while (atomicCAS(env.t4_lock + slot, 0, 1) != 0) { __nanosleep(200); }
        __threadfence();
    }
    __syncwarp();
    double mx = -INFINITY;
    for (int p = lane; p < P; p += 32) {
        const double *g = env.GammaT + p;
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(__ldg(g + (size_t)k * Ppad), omega[k], acc);
        const double l = acc + __ldg(env.mu_o + p);
        e[p] = l;
        mx = fmax(mx, l);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    for (int p = lane; p < P; p += 32) e[p] = exp64(e[p] - mx);
    __syncwarp();
    int found = P - 1;
    if (lane == 0) {
        const double s = np_pairwise_sum_dev(e, P);
        double c = 0.0;
        for (int p = 0; p < P; ++p) c += e[p] / s;
        const double last = c;
        c = 0.0;
        for (int p = 0; p < P; ++p) {
            c += e[p] / s;
            if (c / last > u) { found = p; break; }
        }
        __threadfence();
        atomicExch(env.t4_lock + slot, 0);
    }
    found = __shfl_sync(0xffffffffu, found, 0);
    return found;
}

// Exact (binary64) organic view sample for ONE user by ONE warp (reco_env_v1.py:119-128).
// All products are visited in natural order, lane = product mod 32.  exp is shifted by an upper
// bound of the logits instead of their maximum (the draw is invariant to the shift in exact
// arithmetic; binary64 keeps ~1e-16 relative accuracy either way).  The inverse CDF is located in
// two levels: 1024-product partial sums recorded in `sc` (shared, >= Ppad/1024+1 doubles), then a
// 32-wide inclusive scan inside the crossing super-chunk.
// When the uniform falls within 1e-11 (relative to the total mass; binary64 sums of P <= 10^5 terms agree to
// P 2^-53 <= 1.2e-11) of a CDF boundary -- only there could a different binary64 summation order change the index --
// the draw is repeated in the oracle's own order (tier 4, warp_sample_view_oracle_order) and counted in *ambig.
__device__ inline int warp_sample_view_exact(const EnvDev &env, const double *omega, double *sc, double u,
                                             int lane, unsigned &ambig) {
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    double n2 = 0.0;
    for (int k = 0; k < K; ++k) n2 = fma(omega[k], omega[k], n2);
    const double B = sqrt(n2) * env.gmax + env.mumax;

    auto mass = [&](int j) -> double {
        if (j >= P) return 0.0;
        const double *g = env.GammaT + j;
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(__ldg(g + (size_t)k * Ppad), omega[k], acc);
        return exp64(acc + __ldg(env.mu_o + j) - B);
    };

    double part = 0.0;
    int nsc = 0;
    for (int base = 0; base < Ppad; base += 32) {
        part += mass(base + lane);
        if (((base + 32) % kSuperChunk) == 0 || base + 32 >= Ppad) {
            const double tot = warp_sum(part);
            if (lane == 0) sc[nsc] = tot;
            ++nsc;
            part = 0.0;
        }
    }
    __syncwarp();
    double Z = 0.0;
    for (int c = 0; c < nsc; ++c) Z += sc[c];
    const double tau = u * Z;
    double cum = 0.0;
    int c = 0;
    for (; c < nsc - 1; ++c) {
        const double nx = cum + sc[c];
        if (nx > tau) break;
        cum = nx;
    }
    __syncwarp();
    int found = -1;
    double lower = 0.0, upper = 0.0;
    const int end = min((c + 1) * kSuperChunk, Ppad);
    for (int base = c * kSuperChunk; base < end; base += 32) {
        double x = mass(base + lane);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        const double pre = cum + x;
        const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (base + lane < P));
        if (m) {
            const int fl = __ffs(m) - 1;
            found = base + fl;
            upper = __shfl_sync(0xffffffffu, pre, fl);
            const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
            lower = fl > 0 ? below : cum;
            break;
        }
        cum = __shfl_sync(0xffffffffu, pre, 31);
    }
    bool amb = env.force_t4 != 0;
    if (found < 0) {  // tau >= recomputed mass of the crossing super-chunk: rounding-level tie
        amb = true;
    } else {
        const double tol = fmax(1e-11, 4.0 * (double)P * 1.1102230246251565e-16) * Z;
        // a boundary at distance < tol matters only if it is a real boundary (index > 0 below, index < P-1 above)
        if ((tau - lower < tol && found > 0) || (upper - tau < tol && found < P - 1)) amb = true;
    }
    if (amb) {   // warp-uniform
        found = warp_sample_view_oracle_order(env, omega, u, lane);
        ambig += 1;
    }
    return found;
}

}  // namespace rg
