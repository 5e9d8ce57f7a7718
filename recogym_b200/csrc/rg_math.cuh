// rg_math.cuh -- RNG + elementary-function spec of the reco-gym-v1 hot path (DESIGN.md "RNG spec").
//
// Everything here is __host__ __device__ and built only from IEEE-754 binary64 +, *, /, sqrt, fma
// and integer ops, so the device result is bit-identical on any box; the translation unit is
// compiled with -fmad=false (no implicit contraction) and every fused operation is an explicit
// fma().  The host instantiations are exported as rg_spec_* for the CPU tests.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define RG_HD __host__ __device__ __forceinline__
#else
#define RG_HD inline
#endif

namespace rg {

constexpr uint32_t SLOT_EVENT = 0u;     // words 0,1: view/click uniform; words 2,3: transition uniform
constexpr uint32_t SLOT_POLICY = 1u;    // word 0: uniform logging-policy action; RG_POLICY_USER_VIEWS: words 0,1 =
                                        // the explore coin's uniform, words 2,3 = the action's uniform
constexpr uint32_t SLOT_TIME = 2u;      // NormalTimeGenerator: words 0,1 -> the normal behind the clock step of event index t
constexpr uint32_t SLOT_NORMAL = 0x100u; // + j : normals 2j, 2j+1 of omega init (t=0) / drift into t

struct u32x4 { uint32_t x, y, z, w; };

RG_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: SC'11).
RG_HD u32x4 philox4x32_10(uint32_t k0, uint32_t k1, u32x4 c) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

RG_HD u32x4 draw(uint64_t seed, uint64_t user, uint32_t t, uint32_t slot) {
    return philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32),
                         u32x4{(uint32_t)user, (uint32_t)(user >> 32), t, slot});
}

// 53-bit uniform in [0,1) from two words -- numpy random_sample()'s recipe.
RG_HD double uniform53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

RG_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; __builtin_memcpy(&d, &b, 8); return d;
#endif
}
RG_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; __builtin_memcpy(&b, &d, 8); return b;
#endif
}

constexpr double kLn2Hi = 6.93147180369123816490e-01;
constexpr double kLn2Lo = 1.90821492927058770002e-10;
constexpr double kInvLn2 = 1.44269504088896338700e+00;

// exp: k = rint(x/ln2), two-step Cody-Waite reduction, degree-13 Taylor/Horner, 2^k by bits.
RG_HD double exp64(double x) {
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    const double k = rint(x * kInvLn2);
    double r = fma(-k, kLn2Hi, x);
    r = fma(-k, kLn2Lo, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const long long ki = (long long)k;
    return p * bits_to_double((uint64_t)(ki + 1023) << 52);
}

// log of a normal positive double: atanh series in s=(m-1)/(m+1), m in [sqrt(1/2), sqrt(2)).
RG_HD double log64(double x) {
    const uint64_t b = double_to_bits(x);
    long long e = (long long)(b >> 52) - 1023;
    double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    const double two_s = s + s;
    const double lm = fma(two_s * z, p, two_s);
    const double de = (double)e;
    return fma(de, kLn2Hi, fma(de, kLn2Lo, lm));
}

// Standard normal quantile, Wichura's AS241 (PPND16), rational evaluation with fma.  Split into its central branch
// (|p - 1/2| <= 0.425, 85 % of the draws) and its tail branches so that a warp can evaluate the two separately (the rollout
// gathers the tail cases of a warp onto a few lanes instead of running both branches under divergence); norm_ppf is the
// composition, operation for operation.
RG_HD double norm_ppf_central(double q) {
    const double r = 0.180625 - q * q;
    double n = 2509.0809287301226727;
    n = fma(n, r, 33430.575583588128105);
    n = fma(n, r, 67265.770927008700853);
    n = fma(n, r, 45921.953931549871457);
    n = fma(n, r, 13731.693765509461125);
    n = fma(n, r, 1971.5909503065514427);
    n = fma(n, r, 133.14166789178437745);
    n = fma(n, r, 3.387132872796366608);
    double d = 5226.495278852545925;
    d = fma(d, r, 28729.085735721942674);
    d = fma(d, r, 39307.89580009271061);
    d = fma(d, r, 21213.794301586595867);
    d = fma(d, r, 5394.1960214247511077);
    d = fma(d, r, 687.1870074920579083);
    d = fma(d, r, 42.313330701600911252);
    d = fma(d, r, 1.0);
    return q * n / d;
}
RG_HD double norm_ppf_tail(double p, double q) {
    double r = (q < 0.0) ? p : 1.0 - p;
    r = sqrt(-log64(r));
    double val;
    if (r <= 5.0) {
        r -= 1.6;
        double n = 7.7454501427834140764e-4;
        n = fma(n, r, 0.0227238449892691845833);
        n = fma(n, r, 0.24178072517745061177);
        n = fma(n, r, 1.27045825245236838258);
        n = fma(n, r, 3.64784832476320460504);
        n = fma(n, r, 5.7694972214606914055);
        n = fma(n, r, 4.6303378461565452959);
        n = fma(n, r, 1.42343711074968357734);
        double d = 1.05075007164441684324e-9;
        d = fma(d, r, 5.475938084995344946e-4);
        d = fma(d, r, 0.0151986665636164571966);
        d = fma(d, r, 0.14810397642748007459);
        d = fma(d, r, 0.68976733498510000455);
        d = fma(d, r, 1.6763848301838038494);
        d = fma(d, r, 2.05319162663775882187);
        d = fma(d, r, 1.0);
        val = n / d;
    } else {
        r -= 5.0;
        double n = 2.01033439929228813265e-7;
        n = fma(n, r, 2.71155556874348757815e-5);
        n = fma(n, r, 0.0012426609473880784386);
        n = fma(n, r, 0.026532189526576123093);
        n = fma(n, r, 0.29656057182850489123);
        n = fma(n, r, 1.7848265399172913358);
        n = fma(n, r, 5.4637849111641143699);
        n = fma(n, r, 6.6579046435011037772);
        double d = 2.04426310338993978564e-15;
        d = fma(d, r, 1.4215117583164458887e-7);
        d = fma(d, r, 1.8463183175100546818e-5);
        d = fma(d, r, 7.868691311456132591e-4);
        d = fma(d, r, 0.0148753612908506148525);
        d = fma(d, r, 0.13692988092273580531);
        d = fma(d, r, 0.59983220655588793769);
        d = fma(d, r, 1.0);
        val = n / d;
    }
    return (q < 0.0) ? -val : val;
}
RG_HD double norm_ppf(double p) {
    const double q = p - 0.5;
    if (fabs(q) <= 0.425) return norm_ppf_central(q);
    return norm_ppf_tail(p, q);
}

// standard normal from two words: 52-bit m, p = (2m+1)/2^53 lies strictly inside (0,1).
RG_HD double normal_p_from_words(uint32_t a, uint32_t b) {   // the probability whose quantile is the normal
    const uint64_t m = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
    return (double)(2 * m + 1) * (1.0 / 9007199254740992.0);
}
RG_HD double normal_from_words(uint32_t a, uint32_t b) { return norm_ppf(normal_p_from_words(a, b)); }

// NormalTimeGenerator.new_time() (normal_time_generator.py:25-28): the clock step taken by the `call`-th call of this user,
// |mu + sigma * z|; call i moves the clock from event index i to i + 1.
RG_HD double time_step(uint64_t seed, uint64_t user, uint32_t call, double mu, double sigma) {
    const u32x4 w = draw(seed, user, call, SLOT_TIME);
    return fabs(mu + sigma * normal_from_words(w.x, w.y));
}
// RecoEnv1.update_state's clock handling (reco_env_v1.py:89-93) for the update INTO event index t (t >= 1): the clock is the
// SEQUENTIAL binary64 sum of the steps taken so far (the generator accumulates `current_time += step`), time_delta is the
// DIFFERENCE of two accumulated values -- not the raw step -- and omega_k = 1 if that difference is exactly 0.
// `clock` holds the value at event index t - 1 and is advanced to event index t; returns omega_k.  Default clock: 1.
RG_HD double clock_advance(int time_mode, uint64_t seed, uint64_t user, uint32_t t, double mu, double sigma, double &clock) {
    if (time_mode == 0) return 1.0;
    const double old = clock;
    clock = clock + time_step(seed, user, t - 1u, mu, sigma);
    const double dt = clock - old;
    return dt == 0.0 ? 1.0 : dt;
}
// the clock at event index t from scratch (sequential sum of t steps)
RG_HD double clock_at(int time_mode, uint64_t seed, uint64_t user, uint32_t t, double mu, double sigma) {
    double c = 0.0;
    if (time_mode != 0)
        for (uint32_t i = 0; i < t; ++i) c = c + time_step(seed, user, i, mu, sigma);
    return c;
}

RG_HD double sigmoid64(double x) { return 1.0 / (1.0 + exp64(-x)); }  // reco_env_v1.py:32-34

// click-through rate map, reco_env_v1.py:38-41 with aa=5 bb=2 cc=0.3 dd=2 ee=6
RG_HD double ff64(double x) { return sigmoid64(5.0 * sigmoid64(2.0 * sigmoid64(0.3 * x) - 2.0) - 6.0); }

// uniform action of the default logging policy (abstract.py:214): floor(word * P / 2^32)
RG_HD int32_t uniform_action(uint32_t w, int32_t P) { return (int32_t)mulhi32(w, (uint32_t)P); }

// numpy choice([0,1], p=[1-ctr, ctr]) on uniform u (reco_env_v1.py:112-115):
// cdf = [1-ctr, (1-ctr)+ctr] / cdf[-1]; click = searchsorted(u, 'right')
RG_HD int click_from_uniform(double ctr, double u) {
    const double c0 = 1.0 - ctr;
    const double c1 = c0 + ctr;
    return (c0 / c1 <= u) ? 1 : 0;
}

constexpr uint64_t kHashInit = 0xcbf29ce484222325ull;
RG_HD uint64_t hash_fold(uint64_t h, uint64_t rec) { return (h ^ rec) * 0x100000001B3ull; }

}  // namespace rg
