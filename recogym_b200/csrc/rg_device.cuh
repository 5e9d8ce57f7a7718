// rg_device.cuh -- device-side data layout and warp-level building blocks shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include "rg_math.cuh"
#include "../../include/recogym_b200.h"

namespace rg {

constexpr int kMaxK = 128;          // latent dimension limit of the kernels
constexpr int kSuperChunk = 1024;   // products per recorded partial sum in the exact sampler

// Static parameters as the kernels see them (all pointers are device memory owned by rg_env).
struct EnvDev {
    int P, K, Ppad;            // Ppad = P rounded up to 32
    const double *GammaT;      // [K][Ppad]  K-major f64: lane-consecutive products are contiguous
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
    const float *GTC;          // [tiles][nslab][2][KCs][TILE_N/8][8][4] tf32 hi/lo UMMA tiles in K slabs (rg_tensor.cuh)
    int KC, KCs, nslab;        // 16-byte K chunks per operand row: padded total, per slab, slabs (KC = KCs * nslab)
    int change_omega_for_bandits;
    uint64_t seed;
};

struct PolicyDev {
    int kind;
    const int32_t *table;  // device [P]
    // RG_POLICY_USER_VIEWS: one list of (product, views) pairs, sorted by product, per resident user
    // (ViewsFeaturesProvider of the reference, agents/abstract.py:318-409, kept sparse)
    int2 *views;           // [lists][view_cap]
    int view_cap;
    unsigned long long *view_overflow;  // counts users whose distinct views exceeded view_cap (results invalid)
};

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

// Exact (binary64) organic view sample for ONE user by ONE warp (reco_env_v1.py:119-128).
// All products are visited in natural order, lane = product mod 32.  exp is shifted by an upper
// bound of the logits instead of their maximum (the draw is invariant to the shift in exact
// arithmetic; binary64 keeps ~1e-16 relative accuracy either way).  The inverse CDF is located in
// two levels: 1024-product partial sums recorded in `sc` (shared, >= Ppad/1024+1 doubles), then a
// 32-wide inclusive scan inside the crossing super-chunk.
// *ambig is incremented when the uniform falls within 1e-11 (relative to the total mass) of a CDF
// boundary: only there could a different binary64 summation order change the index.
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
    if (found < 0) {  // tau >= recomputed mass of the crossing super-chunk: rounding-level tie
        found = min(end, P) - 1;
        ambig += 1;
    } else {
        const double tol = 1e-11 * Z;
        if (tau - lower < tol || upper - tau < tol) {
            // a boundary at distance < tol matters only if it is a real boundary (index > 0 below,
            // index < P-1 above)
            if ((tau - lower < tol && found > 0) || (upper - tau < tol && found < P - 1)) ambig += 1;
        }
    }
    return found;
}

}  // namespace rg
