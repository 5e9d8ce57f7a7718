// rg_capi.cu -- the C ABI of include/recogym_b200.h (host side: validation, parameter upload,
// launches, host<->device copies of the *_host entry points).  No CPU compute path exists here:
// without a usable GPU every entry point returns RG_ENODEVICE.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "rg_host.h"
#include "rg_tensor.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define RG_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return fail(RG_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),  \
                        __FILE__, __LINE__);                                                   \
    } while (0)

template <typename T>
cudaError_t upload(T **dst, const std::vector<T> &src) {
    cudaError_t e = cudaMalloc((void **)dst, sizeof(T) * src.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
}

void free_env(rg_env_impl &e) {
    cudaFree(e.d_GammaT); cudaFree(e.d_GammaB); cudaFree(e.d_mu_o); cudaFree(e.d_beta); cudaFree(e.d_mu_b);
    cudaFree(e.d_GammaT32); cudaFree(e.d_mu_o32); cudaFree(e.d_Gamma32); cudaFree(e.d_GTC); cudaFree(e.d_om_scratch); cudaFree(e.d_csum_scratch); cudaFree(e.d_gsum); cudaFree(e.d_views); cudaFree(e.diag); cudaFree(e.d_t4); cudaFree(e.d_host_arena); cudaFree(e.d_hole_cdf);
}

int g_force_t4 = 0;   // rg_debug_force_tier4

rg::RolloutOutDev to_dev(const rg_rollout_out &o) {
    rg::RolloutOutDev d;
    d.n_events = o.n_events; d.n_organic = o.n_organic; d.n_bandit = o.n_bandit; d.n_clicks = o.n_clicks;
    d.last_view = o.last_view; d.hash = o.hash; d.omega = o.omega;
    d.view_hist = (unsigned long long *)o.view_hist;
    d.action_hist = (unsigned long long *)o.action_hist;
    d.click_hist = (unsigned long long *)o.click_hist;
    d.totals = (unsigned long long *)o.totals;
    d.log = (o.log && o.log_count && o.log_capacity > 0) ? (int4 *)o.log : nullptr;
    d.log_capacity = d.log ? o.log_capacity : 0;
    d.log_count = (unsigned long long *)o.log_count;
    return d;
}

int check_rollout_args(const rg_env *env, const rg_policy *policy, int64_t n_organic_users, int64_t n_users,
                       const rg_rollout_out *out, int math_mode) {
    if (!env || !policy || !out) return fail(RG_EINVAL, "null argument");
    if (n_organic_users < 0 || n_users < 0) return fail(RG_EINVAL, "negative user count");
    if (policy->kind != RG_POLICY_UNIFORM && policy->kind != RG_POLICY_TABLE && policy->kind != RG_POLICY_USER_VIEWS &&
        policy->kind != RG_POLICY_EPS_TABLE)
        return fail(RG_EINVAL, "unknown policy kind %d", policy->kind);
    if ((policy->kind == RG_POLICY_TABLE || policy->kind == RG_POLICY_EPS_TABLE) && !policy->table)
        return fail(RG_EINVAL, "table policy without table");
    if (policy->kind == RG_POLICY_EPS_TABLE) {
        if (!(policy->epsilon > 0.0 && policy->epsilon < 1.0)) return fail(RG_EINVAL, "epsilon must lie in (0, 1), got %g", policy->epsilon);
        if (env->impl.dev.P < 2) return fail(RG_EINVAL, "an exploring policy needs at least two products");
    }
    if (math_mode != RG_MATH_EXACT && math_mode != RG_MATH_FAST && math_mode != RG_MATH_TENSOR)
        return fail(RG_EINVAL, "unknown math mode %d", math_mode);
    return RG_OK;
}

}  // namespace

namespace rg {
int set_error(int code, const char *msg) { return fail(code, "%s", msg); }   // for the other translation units (rg_env0.cu)
}

extern "C" {

const char *rg_last_error(void) { return g_err; }
int rg_version(void) { return 100; }

int rg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rg_env_create(const rg_env_desc *d, int device, rg_env **out) {
    if (!d || !out) return fail(RG_EINVAL, "null argument");
    *out = nullptr;
    const int P = d->num_products, K = d->K;
    if (P < 1) return fail(RG_EINVAL, "num_products must be >= 1 (got %d)", P);
    if (K < 1 || K > rg::kMaxK) return fail(RG_EINVAL, "K must be in [1,%d] (got %d)", rg::kMaxK, K);
    if (!d->Gamma || !d->mu_organic || !d->beta || !d->mu_bandit) return fail(RG_EINVAL, "null parameter array");
    if (!(d->sigma_omega_initial >= 0.0) || !(d->sigma_omega >= 0.0)) return fail(RG_EINVAL, "negative sigma");
    for (int r = 0; r < 3; ++r) {
        double s = 0.0;
        for (int c = 0; c < 3; ++c) {
            const double p = d->state_transition[3 * r + c];
            // numpy RandomState.choice: "probabilities are not non-negative" / "do not sum to 1"
            if (!(p >= 0.0) || !(p <= 1.0)) return fail(RG_EINVAL, "state_transition[%d][%d]=%g is not a probability", r, c, p);
            s += p;
        }
        if (std::fabs(s - 1.0) > 1.4901161193847656e-08)
            return fail(RG_EINVAL, "state_transition row %d does not sum to 1 (%.17g)", r, s);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RG_ENODEVICE, "no CUDA device: recogym_b200 has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(RG_EINVAL, "device %d out of range (%d devices)", device, ndev);
    RG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RG_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);

    {   // temporaries of the log pipeline come from the stream-ordered allocator: keep freed blocks in the pool
        // instead of returning them to the driver at every synchronisation (a 3 GB arena costs ~100 ms to re-map)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    rg_env *env = new (std::nothrow) rg_env();
    if (!env) return fail(RG_ENOMEM, "out of host memory");
    rg_env_impl &e = env->impl;
    e.device = device;
    e.num_sms = prop.multiProcessorCount;
    const int Ppad = (P + 31) & ~31;

    std::vector<double> GT((size_t)K * Ppad, 0.0), mu(Ppad, 0.0);
    std::vector<float> GT32((size_t)K * Ppad, 0.f), mu32(Ppad, 0.f);
    double gmax = 0.0, mumax = -INFINITY;
    for (int p = 0; p < P; ++p) {
        double n2 = 0.0;
        for (int k = 0; k < K; ++k) {
            const double g = d->Gamma[(size_t)p * K + k];
            GT[(size_t)k * Ppad + p] = g;
            GT32[(size_t)k * Ppad + p] = (float)g;
            n2 += g * g;
        }
        mu[p] = d->mu_organic[p];
        mu32[p] = (float)d->mu_organic[p];
        gmax = std::fmax(gmax, std::sqrt(n2));
        mumax = std::fmax(mumax, d->mu_organic[p]);
    }
    std::vector<double> beta(d->beta, d->beta + (size_t)P * K), mub(d->mu_bandit, d->mu_bandit + P);
    // fp32 tiles of the fast sampler: [nsub][K+1][128], pre-scaled by log2(e) so that the kernel's
    // exponential is a bare ex2; row K carries mu_organic; products past P get mass 0 (mu = -1e30).
    const int nsub = (P + 127) / 128;
    const double kLog2e = 1.4426950408889634074;
    double muabsmax = 0.0;
    std::vector<float> G32((size_t)nsub * (K + 1) * 128, 0.f);
    for (int p = 0; p < nsub * 128; ++p) {
        const size_t base = (size_t)(p / 128) * (K + 1) * 128 + (p % 128);
        if (p < P) {
            for (int k = 0; k < K; ++k) G32[base + (size_t)k * 128] = (float)(d->Gamma[(size_t)p * K + k] * kLog2e);
            G32[base + (size_t)K * 128] = (float)(d->mu_organic[p] * kLog2e);
            muabsmax = std::fmax(muabsmax, std::fabs(d->mu_organic[p]));
        } else {
            G32[base + (size_t)K * 128] = -1e30f;
        }
    }

    // tf32 hi/lo UMMA tiles of the tensor-core sampler (layout: rg_tensor.cuh)
    int KC = 0, KCs = 0, nslab = 0;
    rg::tc::kt_slabs(K, &KC, &KCs, &nslab);
    using rg::tc::EPC;
    using rg::tc::opnd_t;
    const size_t slab_elems = (size_t)2 * KCs * (rg::tc::CHUNK_BYTES_B / sizeof(opnd_t));   // one K slab of one B tile, hi + lo
    const size_t tile_elems = slab_elems * nslab;
    std::vector<opnd_t> GTC((size_t)nsub * (128 / rg::tc::TILE_N) * tile_elems);
    std::memset(GTC.data(), 0, GTC.size() * sizeof(opnd_t));
    bool opnd_range_ok = true;   // fp16 operands: the tensor mode is refused for parameters outside the element range
    for (int p = 0; p < nsub * 128; ++p) {
        opnd_t *tile0 = GTC.data() + (size_t)(p / rg::tc::TILE_N) * tile_elems;
        const int row = p % rg::tc::TILE_N;
        for (int k = 0; k < KC * EPC; ++k) {
            opnd_t *tile = tile0 + (size_t)(k / (EPC * KCs)) * slab_elems;
            const int kk = k % (EPC * KCs);
            float x = 0.f;
            if (p < P) {
                if (k < K) x = (float)(d->Gamma[(size_t)p * K + k] * kLog2e);
                else if (k == K) x = (float)(d->mu_organic[p] * kLog2e);
                else if (k == K + 1) x = 1.f;
                if (!(std::fabs(x) <= rg::tc::kOpndMax / 2)) { opnd_range_ok = false; x = 0.f; }
            } else if (k == K) {
                x = rg::tc::kPadLogit;
            }
            rg::tc::split_hi_lo(x, tile[rg::tc::tile_index_b(0, KCs, row, kk)], tile[rg::tc::tile_index_b(1, KCs, row, kk)]);
        }
    }

    cudaError_t ce = cudaSuccess;
    e.tc_operands_ok = opnd_range_ok;
    if (ce == cudaSuccess) ce = upload(&e.d_GTC, GTC);
    // the same binary64 Gamma in 8-product blocks for the DMMA sweeps of tier 2 (rg_rollout_fast.cu): one product group's
    // operands are one contiguous run of KQ * 64 bytes instead of 64-byte pieces of K rows 8 * Ppad bytes apart
    const int KQ = 4 * rg::dmma_ksteps(K);
    const size_t nblk = (size_t)((P + 255) / 256) * 32;
    std::vector<double> GB(nblk * KQ * 8, 0.0);
    for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k) GB[((size_t)(p >> 3) * KQ + k) * 8 + (p & 7)] = d->Gamma[(size_t)p * K + k];
    if (ce == cudaSuccess) ce = upload(&e.d_GammaB, GB);
    if (ce == cudaSuccess) ce = upload(&e.d_GammaT, GT);
    if (ce == cudaSuccess) ce = upload(&e.d_mu_o, mu);
    if (ce == cudaSuccess) ce = upload(&e.d_beta, beta);
    if (ce == cudaSuccess) ce = upload(&e.d_mu_b, mub);
    if (ce == cudaSuccess) ce = upload(&e.d_GammaT32, GT32);
    if (ce == cudaSuccess) ce = upload(&e.d_mu_o32, mu32);
    if (ce == cudaSuccess) ce = upload(&e.d_Gamma32, G32);
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&e.diag, sizeof(unsigned long long) * 64);
    if (ce == cudaSuccess) ce = cudaMemset(e.diag, 0, sizeof(unsigned long long) * 64);
    const size_t t4_bytes = sizeof(double) * (size_t)rg::kT4Slots * Ppad + sizeof(int) * rg::kT4Slots;
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&e.d_t4, t4_bytes);
    if (ce == cudaSuccess) ce = cudaMemset(e.d_t4, 0, t4_bytes);
    if (ce != cudaSuccess) {
        free_env(e);
        delete env;
        return fail(ce == cudaErrorMemoryAllocation ? RG_ENOMEM : RG_ECUDA, "parameter upload failed: %s",
                    cudaGetErrorString(ce));
    }

    rg::EnvDev &v = e.dev;
    v.P = P; v.K = K; v.Ppad = Ppad;
    v.GammaB = e.d_GammaB; v.KQ = KQ;
    v.GammaT = e.d_GammaT; v.mu_o = e.d_mu_o; v.beta = e.d_beta; v.mu_b = e.d_mu_b;
    v.GammaT32 = e.d_GammaT32; v.mu_o32 = e.d_mu_o32;
    for (int r = 0; r < 3; ++r) {
        // numpy legacy choice: cdf = p.cumsum(); cdf /= cdf[-1]
        const double c0 = d->state_transition[3 * r], c1 = c0 + d->state_transition[3 * r + 1];
        const double c2 = c1 + d->state_transition[3 * r + 2];
        v.thr[r][0] = c0 / c2;
        v.thr[r][1] = c1 / c2;
    }
    v.sigma0 = d->sigma_omega_initial;
    v.sigma = d->sigma_omega;
    // slack so that rounding in the bound itself can never put a logit above it
    v.gmax = gmax * (1.0 + 1e-12);
    v.mumax = mumax + 1e-9 * std::fabs(mumax) + 1e-12;
    v.muabsmax = muabsmax;
    v.G32 = e.d_Gamma32;
    v.nsub = nsub;
    v.GTC = e.d_GTC;
    v.KC = KC;
    v.KCs = KCs;
    v.nslab = nslab;
    v.change_omega_for_bandits = d->change_omega_for_bandits ? 1 : 0;
    v.time_mode = d->time_generator == 1 ? 1 : 0;
    v.time_mu = d->normal_time_mu;
    v.time_sigma = d->normal_time_sigma;
    v.seed = d->seed;
    v.t4_buf = e.d_t4;
    v.t4_lock = (int *)(e.d_t4 + (size_t)rg::kT4Slots * Ppad);
    v.force_t4 = 0;
    *out = env;
    return RG_OK;
}

int rg_debug_force_tier4(int on) {
    const int was = g_force_t4;
    g_force_t4 = on ? 1 : 0;
    return was;
}

int rg_env_set_seed(rg_env *env, uint64_t seed) {
    if (!env) return fail(RG_EINVAL, "null env");
    env->impl.dev.seed = seed;
    return RG_OK;
}

int rg_env_destroy(rg_env *env) {
    if (!env) return RG_OK;
    cudaSetDevice(env->impl.device);
    free_env(env->impl);
    delete env;
    return RG_OK;
}

int rg_rollout_device(rg_env *env, const rg_policy *policy, int64_t user_begin, int64_t n_organic_users,
                      int64_t n_users, const rg_rollout_out *out, int math_mode, void *stream) {
    int rc = check_rollout_args(env, policy, n_organic_users, n_users, out, math_mode);
    if (rc != RG_OK) return rc;
    RG_CUDA(cudaSetDevice(env->impl.device));
    env->impl.dev.force_t4 = g_force_t4;
    rg::PolicyDev pol{policy->kind, policy->table, nullptr, 0, nullptr, 0.0, nullptr};
    const rg::RolloutOutDev od = to_dev(*out);
    if (policy->kind == RG_POLICY_EPS_TABLE) {
        // numpy: choice([True, False], p = [eps, 1 - eps]) -> cdf = [eps, eps + (1 - eps)] / cdf[-1]; explore iff u < cdf[0]
        const double c0 = policy->epsilon, c1 = c0 + (1.0 - policy->epsilon);
        pol.eps_c0 = c0 / c1;
        rg_env_impl &e = env->impl;
        if (!e.d_hole_cdf) {      // S[m] = sequential binary64 sum of m copies of 1 / (P - 1): numpy's cumsum, once per env
            const int P = e.dev.P;
            std::vector<double> S((size_t)P, 0.0);
            const double c = 1.0 / (double)(P - 1);
            double acc = 0.0;
            for (int m = 1; m < P; ++m) { acc = acc + c; S[(size_t)m] = acc; }
            RG_CUDA(cudaMalloc((void **)&e.d_hole_cdf, sizeof(double) * (size_t)P));
            RG_CUDA(cudaMemcpy(e.d_hole_cdf, S.data(), sizeof(double) * (size_t)P, cudaMemcpyHostToDevice));
        }
        pol.hole_cdf = e.d_hole_cdf;
    }
    if (policy->kind == RG_POLICY_USER_VIEWS) {
        // one sorted (product, views) list per resident user: 2 CTAs x 128 slots per SM in the CTA-batched kernel,
        // 64 warps per SM in the warp-per-user kernel.  A user with more than 1024 distinct views would need an
        // episode of > ~4000 events (probability 0.99^4000 per user); it is counted in totals[9] if it ever happens.
        rg_env_impl &e = env->impl;
        const int cap = e.dev.P < 1024 ? e.dev.P : 1024;
        const size_t lists = (size_t)e.num_sms * 2 * 128;
        if (e.views_lists < lists || e.view_cap != cap) {
            if (e.d_views) cudaFree(e.d_views);
            e.d_views = nullptr; e.views_lists = 0;
            RG_CUDA(cudaMalloc((void **)&e.d_views, sizeof(int2) * lists * (size_t)cap));
            e.views_lists = lists; e.view_cap = cap;
        }
        pol.views = e.d_views;
        pol.view_cap = cap;
        pol.view_overflow = od.totals ? od.totals + 9 : e.diag + 1;
    }
    if (math_mode == RG_MATH_FAST || math_mode == RG_MATH_TENSOR)
        RG_CUDA(rg::launch_rollout_fast(env->impl, pol, user_begin, n_organic_users, n_organic_users + n_users, od,
                                        (cudaStream_t)stream, math_mode == RG_MATH_TENSOR));
    else
        RG_CUDA(rg::launch_rollout_exact(env->impl, pol, user_begin, n_organic_users, n_organic_users + n_users, od,
                                         (cudaStream_t)stream));
    return RG_OK;
}

int rg_rollout_host(rg_env *env, const rg_policy *policy, int64_t user_begin, int64_t n_organic_users,
                    int64_t n_users, const rg_rollout_out *out, int math_mode) {
    int rc = check_rollout_args(env, policy, n_organic_users, n_users, out, math_mode);
    if (rc != RG_OK) return rc;
    RG_CUDA(cudaSetDevice(env->impl.device));
    const int64_t n = n_organic_users + n_users;
    const int P = env->impl.dev.P, K = env->impl.dev.K;
    // one device arena for the table and every requested output
    struct Seg { void *host; size_t bytes; size_t off; bool accumulate; };
    std::vector<Seg> segs;
    size_t total = 0;
    auto add = [&](void *host, size_t bytes, bool acc) -> size_t {
        const size_t off = total;
        if (host) { segs.push_back(Seg{host, bytes, off, acc}); total += (bytes + 255) & ~(size_t)255; }
        return off;
    };
    const size_t o_ne = add(out->n_events, 4 * n, false), o_no = add(out->n_organic, 4 * n, false);
    const size_t o_nb = add(out->n_bandit, 4 * n, false), o_nc = add(out->n_clicks, 4 * n, false);
    const size_t o_lv = add(out->last_view, 4 * n, false), o_h = add(out->hash, 8 * n, false);
    const size_t o_om = add(out->omega, 8 * (size_t)n * K, false);
    const size_t o_vh = add(out->view_hist, 8 * (size_t)P, true), o_ah = add(out->action_hist, 8 * (size_t)P, true);
    const size_t o_ch = add(out->click_hist, 8 * (size_t)P, true), o_t = add(out->totals, 8 * 16, true);
    const size_t o_log = add(out->log && out->log_count && out->log_capacity > 0 ? out->log : nullptr,
                             16 * (size_t)(out->log_capacity > 0 ? out->log_capacity : 0), false);
    const size_t o_lc = add(out->log && out->log_count ? out->log_count : nullptr, 8, true);
    const size_t o_tab = total;
    const bool has_table = policy->kind == RG_POLICY_TABLE || policy->kind == RG_POLICY_EPS_TABLE;
    if (has_table) total += ((size_t)4 * P + 255) & ~(size_t)255;
    // The staging arena belongs to the env and only ever grows: no cudaMalloc / cudaFree (an implicit device
    // synchronisation each) on the call path.  Only the ACCUMULATED outputs need zeroing -- every per-user entry and
    // every log row below *log_count is written by the kernel.
    if (total == 0) total = 256;
    rg_env_impl &e = env->impl;
    if (e.host_arena_bytes < total) {
        if (e.d_host_arena) cudaFree(e.d_host_arena);
        e.d_host_arena = nullptr;
        e.host_arena_bytes = 0;
        const size_t want = total + total / 4 + (1u << 20);
        cudaError_t ae = cudaMalloc((void **)&e.d_host_arena, want);
        if (ae != cudaSuccess) return fail(RG_ENOMEM, "rollout staging arena (%zu bytes): %s", want, cudaGetErrorString(ae));
        e.host_arena_bytes = want;
    }
    char *arena = e.d_host_arena;
    cudaError_t ce = cudaSuccess;
    for (const Seg &s : segs)
        if (s.accumulate && ce == cudaSuccess) ce = cudaMemsetAsync(arena + s.off, 0, s.bytes, 0);
    rg_policy pd = *policy;
    if (ce == cudaSuccess && has_table) {
        ce = cudaMemcpyAsync(arena + o_tab, policy->table, (size_t)4 * P, cudaMemcpyHostToDevice, 0);
        pd.table = (const int32_t *)(arena + o_tab);
    }
    rg_rollout_out dv;
    std::memset(&dv, 0, sizeof(dv));
    if (out->n_events) dv.n_events = (int32_t *)(arena + o_ne);
    if (out->n_organic) dv.n_organic = (int32_t *)(arena + o_no);
    if (out->n_bandit) dv.n_bandit = (int32_t *)(arena + o_nb);
    if (out->n_clicks) dv.n_clicks = (int32_t *)(arena + o_nc);
    if (out->last_view) dv.last_view = (int32_t *)(arena + o_lv);
    if (out->hash) dv.hash = (uint64_t *)(arena + o_h);
    if (out->omega) dv.omega = (double *)(arena + o_om);
    if (out->view_hist) dv.view_hist = (int64_t *)(arena + o_vh);
    if (out->action_hist) dv.action_hist = (int64_t *)(arena + o_ah);
    if (out->click_hist) dv.click_hist = (int64_t *)(arena + o_ch);
    if (out->totals) dv.totals = (int64_t *)(arena + o_t);
    if (out->log && out->log_count && out->log_capacity > 0) {
        dv.log = (int32_t *)(arena + o_log);
        dv.log_capacity = out->log_capacity;
        dv.log_count = (int64_t *)(arena + o_lc);
    }
    if (ce == cudaSuccess) {
        rc = rg_rollout_device(env, &pd, user_begin, n_organic_users, n_users, &dv, math_mode, nullptr);
        if (rc != RG_OK) return rc;
    }
    // D2H on the same (legacy default) stream: each copy waits for the rollout and returns when the bytes are on the host
    std::vector<int64_t> tmp;
    for (const Seg &s : segs) {
        if (ce != cudaSuccess) break;
        if (!s.accumulate) {
            ce = cudaMemcpy(s.host, arena + s.off, s.bytes, cudaMemcpyDeviceToHost);
        } else {
            tmp.resize(s.bytes / 8);
            ce = cudaMemcpy(tmp.data(), arena + s.off, s.bytes, cudaMemcpyDeviceToHost);
            if (ce == cudaSuccess)
                for (size_t i = 0; i < tmp.size(); ++i) ((int64_t *)s.host)[i] += tmp[i];
        }
    }
    if (ce == cudaSuccess && segs.empty()) ce = cudaStreamSynchronize(0);
    if (ce != cudaSuccess) return fail(RG_ECUDA, "rollout failed: %s", cudaGetErrorString(ce));
    return RG_OK;
}

static int current_num_sms() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

int rg_generate_logs_device(rg_env *env, const rg_policy *policy, const double *ps_table_device, int64_t user_begin,
                            int64_t n_organic_users, int64_t n_users, int32_t prev_last_view, int train_mode,
                            int32_t *carry_device, const rg_log_columns *cols, int64_t *row_offset_device, int32_t *last_view_device,
                            int64_t *rows_device, int64_t *totals_device, int math_mode, void *stream_) {
    if (!cols || !rows_device) return fail(RG_EINVAL, "null argument");
    if (cols->capacity < 0) return fail(RG_EINVAL, "negative capacity");
    if (train_mode < 0 || train_mode > 2) return fail(RG_EINVAL, "unknown train mode %d", train_mode);
    rg_rollout_out probe;
    std::memset(&probe, 0, sizeof(probe));
    int rc = check_rollout_args(env, policy, n_organic_users, n_users, &probe, math_mode);
    if (rc != RG_OK) return rc;
    RG_CUDA(cudaSetDevice(env->impl.device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t n = n_organic_users + n_users, cap = cols->capacity;
    const int64_t ntiles = (n + 4095) / 4096;
    // one stream-ordered arena: unordered records | ordered records | n_events | last_view | row_offset | scan tiles | count
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_log = take(16 * (size_t)(cap > 0 ? cap : 1)), o_ord = take(16 * (size_t)(cap > 0 ? cap : 1));
    const size_t o_ne = take(4 * (size_t)(n + 1)), o_lv = take(4 * (size_t)(n + 1));
    const size_t o_ro = take(8 * (size_t)(n + 1)), o_ts = take(8 * (size_t)(ntiles + 2)), o_cnt = take(16);
    char *arena = nullptr;
    RG_CUDA(cudaMallocAsync((void **)&arena, off, stream));
    cudaError_t ce = cudaMemsetAsync(arena + o_cnt, 0, 16, stream);
    int32_t *lv = last_view_device ? last_view_device : (int32_t *)(arena + o_lv);
    long long *ro = row_offset_device ? (long long *)row_offset_device : (long long *)(arena + o_ro);
    if (ce == cudaSuccess) {
        rg_rollout_out o;
        std::memset(&o, 0, sizeof(o));
        o.n_events = (int32_t *)(arena + o_ne);
        o.last_view = lv;
        o.totals = totals_device;
        o.log = (int32_t *)(arena + o_log);
        o.log_capacity = cap;
        o.log_count = (int64_t *)(arena + o_cnt);
        rc = cap > 0 || n == 0 ? rg_rollout_device(env, policy, user_begin, n_organic_users, n_users, &o, math_mode, stream_)
                               : fail(RG_EINVAL, "columns without capacity");
        if (rc != RG_OK) { cudaFreeAsync(arena, stream); return rc; }
        ce = rg::launch_row_offsets((const int32_t *)(arena + o_ne), n, n_organic_users, ro, (long long *)(arena + o_ts), stream);
    }
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(rows_device, ro + n, 8, cudaMemcpyDeviceToDevice, stream);
    if (ce == cudaSuccess && n > 0)
        ce = rg::launch_order((const int4 *)(arena + o_log), (const unsigned long long *)(arena + o_cnt), cap, ro,
                              (int4 *)(arena + o_ord), cap, env->impl.num_sms, stream);
    if (ce == cudaSuccess)
        ce = rg::launch_columns((const int4 *)(arena + o_ord), ro, n, n_organic_users, user_begin, lv, prev_last_view,
                                policy->kind, env->impl.dev.P, ps_table_device, policy->table, policy->epsilon,
                                env->impl.dev.time_mode, env->impl.dev.seed, env->impl.dev.time_mu, env->impl.dev.time_sigma, *cols,
                                train_mode, carry_device,
                                env->impl.num_sms, stream);
    cudaFreeAsync(arena, stream);
    if (ce != cudaSuccess) return fail(RG_ECUDA, "generate_logs failed: %s", cudaGetErrorString(ce));
    return RG_OK;
}

int rg_count_stats_device(const rg_log_columns *cols, int64_t rows, int32_t P, int train_mode, int64_t *co_counts,
                          int64_t *pulls, int64_t *clicks, void *stream) {
    if (!cols || rows < 0 || P < 1 || train_mode < 0 || train_mode > 2) return fail(RG_EINVAL, "bad argument");
    if (rows > cols->capacity) return fail(RG_EINVAL, "rows exceed the column capacity");
    if (!cols->z || !cols->v || !cols->a || !cols->c || !cols->lvp || !cols->sess || !cols->flags)
        return fail(RG_EINVAL, "count statistics need the columns z, v, a, c, lvp, sess, flags");
    if ((pulls == nullptr) != (clicks == nullptr)) return fail(RG_EINVAL, "pulls and clicks come together");
    RG_CUDA(rg::launch_count_stats(*cols, rows, P, train_mode, co_counts, pulls, clicks, current_num_sms(),
                                   (cudaStream_t)stream));
    return RG_OK;
}

int rg_count_policy_device(int kind, int32_t P, const int64_t *A, const int64_t *B, int32_t *table, double *ps,
                           void *stream) {
    if ((kind != 0 && kind != 1) || P < 1 || !A || !table || (kind == 1 && !B)) return fail(RG_EINVAL, "bad argument");
    RG_CUDA(rg::launch_row_argmax(kind, P, A, B, table, ps, (cudaStream_t)stream));
    return RG_OK;
}

static int check_sparse(const rg_sparse_counts *s, bool bandit) {
    if (!s->keys || !s->a || !s->stats || s->capacity < 2 || (s->capacity & (s->capacity - 1)))
        return fail(RG_EINVAL, "sparse counts: keys / a / stats and a power-of-two capacity are required");
    if (bandit && (!s->b || !s->row_a || !s->row_b)) return fail(RG_EINVAL, "sparse bandit counts need b, row_a, row_b");
    return RG_OK;
}

int rg_sparse_count_stats_device(const rg_log_columns *cols, int64_t rows, int train_mode, const rg_sparse_counts *co,
                                 const rg_sparse_counts *bandit, void *stream) {
    if (!cols || rows < 0 || train_mode < 0 || train_mode > 2) return fail(RG_EINVAL, "bad argument");
    if (rows > cols->capacity) return fail(RG_EINVAL, "rows exceed the column capacity");
    if (!cols->z || !cols->v || !cols->a || !cols->c || !cols->lvp || !cols->sess || !cols->flags)
        return fail(RG_EINVAL, "count statistics need the columns z, v, a, c, lvp, sess, flags");
    int rc = RG_OK;
    if (co && (rc = check_sparse(co, false)) != RG_OK) return rc;
    if (bandit && (rc = check_sparse(bandit, true)) != RG_OK) return rc;
    RG_CUDA(rg::launch_sparse_count_stats(*cols, rows, train_mode, co, bandit, current_num_sms(), (cudaStream_t)stream));
    return RG_OK;
}

int rg_sparse_rehash_device(const rg_sparse_counts *from, const rg_sparse_counts *to, void *stream) {
    if (!from || !to) return fail(RG_EINVAL, "null argument");
    int rc = check_sparse(from, false);
    if (rc == RG_OK) rc = check_sparse(to, false);
    if (rc != RG_OK) return rc;
    if ((from->b != nullptr) != (to->b != nullptr)) return fail(RG_EINVAL, "both tables must have (or lack) the b values");
    RG_CUDA(rg::launch_sparse_rehash(*from, *to, current_num_sms(), (cudaStream_t)stream));
    return RG_OK;
}

int rg_sparse_count_policy_device(int kind, int32_t P, const rg_sparse_counts *s, int32_t *table, double *ps, void *stream_) {
    if ((kind != 0 && kind != 1) || P < 1 || !s || !table) return fail(RG_EINVAL, "bad argument");
    int rc = check_sparse(s, kind == 1);
    if (rc != RG_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long *scratch = nullptr;
    RG_CUDA(cudaMallocAsync((void **)&scratch, sizeof(unsigned long long) * 3 * (size_t)P, stream));
    cudaError_t ce = rg::launch_sparse_policy(kind, P, *s, table, ps, scratch, current_num_sms(), stream);
    cudaFreeAsync(scratch, stream);
    if (ce != cudaSuccess) return fail(RG_ECUDA, "sparse policy failed: %s", cudaGetErrorString(ce));
    return RG_OK;
}

int rg_ips_device(const rg_log_columns *cols, int64_t rows, int64_t u_end, const int32_t *target_table,
                  double *out4_device, void *stream_) {
    if (!cols || rows < 0 || !target_table || !out4_device) return fail(RG_EINVAL, "bad argument");
    if (rows > cols->capacity) return fail(RG_EINVAL, "rows exceed the column capacity");
    if (!cols->z || !cols->u || !cols->a || !cols->c || !cols->ps || !cols->lv)
        return fail(RG_EINVAL, "IPS needs the columns z, u, a, c, ps, lv");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nblocks = current_num_sms() * 4;
    double *part = nullptr;
    RG_CUDA(cudaMallocAsync((void **)&part, sizeof(double) * 4 * (size_t)nblocks, stream));
    cudaError_t ce = rg::launch_ips(*cols, rows, u_end, target_table, part, nblocks, out4_device, stream);
    cudaFreeAsync(part, stream);
    if (ce != cudaSuccess) return fail(RG_ECUDA, "ips failed: %s", cudaGetErrorString(ce));
    return RG_OK;
}

int rg_recall_device(const rg_log_columns *cols, int64_t rows, int64_t u_end, const int32_t *target_table,
                     const int32_t *topk_table, int32_t k, int64_t *out2_device, void *stream_) {
    if (!cols || rows < 0 || !target_table || !topk_table || !out2_device || k < 1) return fail(RG_EINVAL, "bad argument");
    if (rows > cols->capacity) return fail(RG_EINVAL, "rows exceed the column capacity");
    if (!cols->z || !cols->u || !cols->v || !cols->c || !cols->lv) return fail(RG_EINVAL, "recall@k needs the columns z, u, v, c, lv");
    cudaError_t ce = rg::launch_recall(*cols, rows, u_end, target_table, topk_table, k, (unsigned long long *)out2_device,
                                       current_num_sms(), (cudaStream_t)stream_);
    if (ce != cudaSuccess) return fail(RG_ECUDA, "recall failed: %s", cudaGetErrorString(ce));
    return RG_OK;
}

static int batch_common(rg_env *env, int mode, int64_t N, const int64_t *user_id, const int32_t *action,
                        int32_t *state, int32_t *t, int32_t *last_view, double *omega, int32_t *reward,
                        int32_t *sess_v, int32_t *sess_len, int32_t sess_cap, int32_t *n_unfinished, int math_mode,
                        void *stream) {
    if (!env) return fail(RG_EINVAL, "null env");
    if (N < 0) return fail(RG_EINVAL, "negative batch size");
    if (!user_id || !state || !t || !last_view || !omega || !sess_v || !sess_len)
        return fail(RG_EINVAL, "null state/session buffer");
    if (mode == 1 && (!action || !reward)) return fail(RG_EINVAL, "step needs action and reward buffers");
    if (sess_cap < 1) return fail(RG_EINVAL, "sess_cap must be >= 1");
    if (math_mode != RG_MATH_EXACT && math_mode != RG_MATH_FAST) return fail(RG_EINVAL, "unknown math mode");
    RG_CUDA(cudaSetDevice(env->impl.device));
    env->impl.dev.force_t4 = g_force_t4;
    RG_CUDA(rg::launch_batch(env->impl, mode, N, (const long long *)user_id, action, state, t, last_view, omega,
                             reward, sess_v, sess_len, sess_cap, n_unfinished, (cudaStream_t)stream));
    return RG_OK;
}

int rg_batch_reset_device(rg_env *env, int64_t N, const int64_t *user_id, int32_t *state, int32_t *t,
                          int32_t *last_view, double *omega, int32_t *sess_v, int32_t *sess_len, int32_t sess_cap,
                          int32_t *n_unfinished, int math_mode, void *stream) {
    return batch_common(env, 0, N, user_id, nullptr, state, t, last_view, omega, nullptr, sess_v, sess_len, sess_cap,
                        n_unfinished, math_mode, stream);
}

int rg_batch_step_device(rg_env *env, int64_t N, const int64_t *user_id, const int32_t *action, int32_t *state,
                         int32_t *t, int32_t *last_view, double *omega, int32_t *reward, int32_t *sess_v,
                         int32_t *sess_len, int32_t sess_cap, int32_t *n_unfinished, int math_mode, void *stream) {
    return batch_common(env, 1, N, user_id, action, state, t, last_view, omega, reward, sess_v, sess_len, sess_cap,
                        n_unfinished, math_mode, stream);
}

int rg_batch_continue_device(rg_env *env, int64_t N, const int64_t *user_id, int32_t *state, int32_t *t,
                             int32_t *last_view, double *omega, int32_t *sess_v, int32_t *sess_len,
                             int32_t sess_cap, int32_t *n_unfinished, int math_mode, void *stream) {
    return batch_common(env, 2, N, user_id, nullptr, state, t, last_view, omega, nullptr, sess_v, sess_len, sess_cap,
                        n_unfinished, math_mode, stream);
}

int rg_spec_eval_device(int kind, int64_t n, const double *in_device, double *out_device, void *stream) {
    if (kind < 0 || kind > 6 || n < 0 || !in_device || !out_device) return fail(RG_EINVAL, "bad spec_eval argument");
    RG_CUDA(rg::launch_spec_eval(kind, n, in_device, out_device, (cudaStream_t)stream));
    return RG_OK;
}

int rg_tc_logits_device(rg_env *env, const double *omega_device, const float *shift_device, float *out_device,
                        void *stream) {
    if (!env || !omega_device || !shift_device || !out_device) return fail(RG_EINVAL, "null argument");
    if (!env->impl.tc_operands_ok) return fail(RG_EINVAL, "Gamma / mu_organic outside the range of the tensor-core operand elements");
    RG_CUDA(cudaSetDevice(env->impl.device));
    RG_CUDA(rg::launch_tc_logits(env->impl, omega_device, shift_device, out_device, (cudaStream_t)stream));
    return RG_OK;
}

// ---- scalar spec functions, host instantiations of rg_math.cuh ---------------------------------
void rg_spec_philox(uint64_t seed, uint64_t user, uint32_t t, uint32_t slot, uint32_t out[4]) {
    const rg::u32x4 w = rg::draw(seed, user, t, slot);
    out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
}
double rg_spec_exp(double x) { return rg::exp64(x); }
double rg_spec_log(double x) { return rg::log64(x); }
double rg_spec_norm_ppf(double p) { return rg::norm_ppf(p); }
double rg_spec_ff(double x) { return rg::ff64(x); }
double rg_spec_event_uniform(uint64_t seed, uint64_t user, uint32_t t, int which) {
    const rg::u32x4 w = rg::draw(seed, user, t, rg::SLOT_EVENT);
    return which ? rg::uniform53(w.z, w.w) : rg::uniform53(w.x, w.y);
}
void rg_spec_normals(uint64_t seed, uint64_t user, uint32_t t, int K, double *z) {
    for (int j = 0; 2 * j < K; ++j) {
        const rg::u32x4 w = rg::draw(seed, user, t, rg::SLOT_NORMAL + (uint32_t)j);
        z[2 * j] = rg::normal_from_words(w.x, w.y);
        if (2 * j + 1 < K) z[2 * j + 1] = rg::normal_from_words(w.z, w.w);
    }
}
double rg_spec_time_step(uint64_t seed, uint64_t user, uint32_t call, double mu, double sigma) {
    return rg::time_step(seed, user, call, mu, sigma);
}
int32_t rg_spec_policy_uniform_action(uint64_t seed, uint64_t user, uint32_t t, int32_t P) {
    return rg::uniform_action(rg::draw(seed, user, t, rg::SLOT_POLICY).x, P);
}

}  // extern "C"
