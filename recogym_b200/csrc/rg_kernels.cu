// rg_kernels.cu -- warp-per-user kernels of the reco-gym-v1 hot path (RG_MATH_EXACT) and the
// batched gym protocol (reset / step / continue).
//
// One warp owns one simulated user at a time: omega lives in the warp's slice of shared memory,
// the Markov state, clock and counters in registers.  Every random draw is a pure function of
// (seed, user, event time, slot) -- DESIGN.md "RNG spec" -- so results do not depend on the grid,
// on the order in which users are processed, or on how users are sharded over GPUs.
#include "rg_device.cuh"
#include "rg_host.h"

namespace rg {

constexpr int kWarpsPerBlock = 8;

struct WarpCtx {
    double *omega;  // [Kpad] shared
    double *sc;     // [nsc_cap] shared
};

__device__ __forceinline__ WarpCtx warp_ctx(const EnvDev &env, double *smem, int warp_in_block) {
    const int kpad = (env.K + 1) & ~1;
    const int nsc_cap = env.Ppad / kSuperChunk + 1;
    double *base = smem + (size_t)warp_in_block * (kpad + nsc_cap);
    return WarpCtx{base, base + kpad};
}

size_t warp_kernel_smem_bytes(const EnvDev &env) {
    const int kpad = (env.K + 1) & ~1;
    const int nsc_cap = env.Ppad / kSuperChunk + 1;
    return sizeof(double) * (size_t)kWarpsPerBlock * (kpad + nsc_cap);
}

// One Markov event in the organic state (abstract.py:110-119): view, then update_state.
// Returns the view; updates state, t, omega.
__device__ __forceinline__ int organic_event(const EnvDev &env, const WarpCtx &wc, uint64_t user, int &state,
                                             uint32_t &t, double &clock, int lane, unsigned &ambig) {
    const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
    const int v = warp_sample_view_exact(env, wc.omega, wc.sc, uniform53(w.x, w.y), lane, ambig);
    state = next_state(env, RG_ORGANIC, uniform53(w.z, w.w));
    t += 1;
    const double omega_k = clock_advance(env.time_mode, env.seed, user, t, env.time_mu, env.time_sigma, clock);   // every update_state advances the clock
    if (env.change_omega_for_bandits || state == RG_ORGANIC)
        warp_omega_update(wc.omega, env.K, env.sigma * omega_k, false, env.seed, user, t, lane);
    return v;
}

// One Markov event in the bandit state (abstract.py:174-181): click, update_state, click override.
__device__ __forceinline__ int bandit_event(const EnvDev &env, const WarpCtx &wc, uint64_t user, int a, int &state,
                                            uint32_t &t, double &clock, int lane) {
    const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
    const int click = click_from_uniform(ctr_of_action(env, wc.omega, a), uniform53(w.x, w.y));
    state = next_state(env, RG_BANDIT, uniform53(w.z, w.w));
    t += 1;
    const double omega_k = clock_advance(env.time_mode, env.seed, user, t, env.time_mu, env.time_sigma, clock);   // every update_state advances the clock
    if (env.change_omega_for_bandits || state == RG_ORGANIC)
        warp_omega_update(wc.omega, env.K, env.sigma * omega_k, false, env.seed, user, t, lane);
    if (click) state = RG_ORGANIC;  // abstract.py:180-181: a click overrides any transition, incl. stop
    return click;
}

__device__ __forceinline__ int policy_action(const EnvDev &env, const PolicyDev &pol, uint64_t user, uint32_t t,
                                             int last_view, const int2 *vlist, int nd, int nviews) {
    if (pol.kind == RG_POLICY_TABLE) return __ldg(pol.table + last_view);
    const u32x4 w = draw(env.seed, user, t, SLOT_POLICY);
    if (pol.kind == RG_POLICY_EPS_TABLE) return eps_table_action(pol, env.P, last_view, w);
    if (pol.kind == RG_POLICY_USER_VIEWS) return views_action(vlist, nd, nviews, uniform53(w.z, w.w));
    return uniform_action(w.x, env.P);
}

// generate_logs (abstract.py:241-327) reduced to counters, RG_MATH_EXACT.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rollout_warp_kernel(const EnvDev env, const PolicyDev pol, long long user_begin, long long n_organic_users,
                    long long n_total, const RolloutOutDev out) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const WarpCtx wc = warp_ctx(env, smem, wib);
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + wib;
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;

    unsigned long long tot_users = 0, tot_events = 0, tot_org = 0, tot_ban = 0, tot_clk = 0, tot_ph = 0;
    unsigned ambig = 0;

    for (long long i = warp0; i < n_total; i += nwarps) {
        const uint64_t user = (uint64_t)(user_begin + i);
        int state = RG_ORGANIC;
        uint32_t t = 0;
        double clock = 0.0;     // current_time of the user's clock (NormalTimeGenerator; unused under the default clock)
        int last_view = -1, n_org = 0, n_ban = 0, n_clk = 0;
        uint64_t h = kHashInit;
        warp_omega_update(wc.omega, env.K, env.sigma0, true, env.seed, user, 0u, lane);
        const bool organic_only = i < n_organic_users;
        const bool track_views = pol.kind == RG_POLICY_USER_VIEWS;
        int2 *vlist = track_views ? pol.views + (size_t)warp0 * pol.view_cap : nullptr;
        int nd = 0;  // distinct products viewed by this user (per-user features reset, organic_user_count.py:111-116)
        for (;;) {
            if (state == RG_ORGANIC) {
                const int v = organic_event(env, wc, user, state, t, clock, lane, ambig);
                last_view = v;
                n_org += 1;
                h = hash_fold(h, ((uint64_t)(uint32_t)v << 2) | 0u);
                if (out.view_hist && lane == 0) atomicAdd(out.view_hist + v, 1ull);
                if (out.log && lane == 0) log_write(out, log_reserve(out, 1), i, t - 1u, v, 0);
                if (track_views) {
                    if (lane == 0) views_insert(vlist, nd, pol.view_cap, v, pol.view_overflow);
                    nd = __shfl_sync(0xffffffffu, nd, 0);
                    __syncwarp();
                }
            } else if (state == RG_BANDIT && !organic_only) {
                const int a = policy_action(env, pol, user, t, last_view, vlist, nd, n_org);
                const int c = bandit_event(env, wc, user, a, state, t, clock, lane);
                if (out.log && lane == 0) log_write(out, log_reserve(out, 1), i, t - 1u, a, 1 | (c << 1));
                n_ban += 1;
                n_clk += c;
                h = hash_fold(h, ((uint64_t)(uint32_t)a << 2) | ((uint64_t)c << 1) | 1u);
                if (lane == 0) {
                    if (out.action_hist) atomicAdd(out.action_hist + a, 1ull);
                    if (c && out.click_hist) atomicAdd(out.click_hist + a, 1ull);
                }
            } else {
                break;
            }
        }
        h = hash_fold(h, ((uint64_t)t << 2) | 3u);
        if (out.log && lane == 0 && !organic_only)  // phantom action after done (abstract.py:311-316), reward 0
            log_write(out, log_reserve(out, 1), i, t, policy_action(env, pol, user, t, last_view, vlist, nd, n_org), 1 | 4);
        if (lane == 0) {
            if (out.n_events) out.n_events[i] = (int32_t)t;
            if (out.n_organic) out.n_organic[i] = n_org;
            if (out.n_bandit) out.n_bandit[i] = n_ban;
            if (out.n_clicks) out.n_clicks[i] = n_clk;
            if (out.last_view) out.last_view[i] = last_view;
            if (out.hash) out.hash[i] = h;
        }
        if (out.omega)
            for (int k = lane; k < env.K; k += 32) out.omega[(size_t)i * env.K + k] = wc.omega[k];
        __syncwarp();
        tot_users += 1; tot_events += t; tot_org += n_org; tot_ban += n_ban; tot_clk += n_clk;
        tot_ph += organic_only ? 0 : 1;  // phantom row of generate_logs (abstract.py:311-316)
    }
    if (out.totals && lane == 0 && tot_users) {
        atomicAdd(out.totals + 0, tot_users);
        atomicAdd(out.totals + 1, tot_events);
        atomicAdd(out.totals + 2, tot_org);
        atomicAdd(out.totals + 3, tot_ban);
        atomicAdd(out.totals + 4, tot_clk);
        atomicAdd(out.totals + 5, tot_ph);
        if (ambig) atomicAdd(out.totals + 7, (unsigned long long)ambig);
    }
}

// ---------------------------------------------------------------------------------------------
// Batched gym protocol.  mode 0: reset + step(None); mode 1: step(action); mode 2: continue session.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
batch_step_kernel(const EnvDev env, int mode, long long N, const long long *__restrict__ user_id,
                  const int32_t *__restrict__ action, int32_t *__restrict__ state_io, int32_t *__restrict__ t_io,
                  int32_t *__restrict__ last_view_io, double *__restrict__ omega_io, int32_t *__restrict__ reward,
                  int32_t *__restrict__ sess_v, int32_t *__restrict__ sess_len, int sess_cap,
                  int32_t *__restrict__ n_unfinished, unsigned long long *__restrict__ diag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const WarpCtx wc = warp_ctx(env, smem, wib);
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + wib;
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    unsigned ambig = 0;
    int unfinished = 0;

    for (long long i = warp0; i < N; i += nwarps) {
        const uint64_t user = (uint64_t)user_id[i];
        int state;
        uint32_t t;
        int last_view;
        int len = 0;
        double clock = 0.0;
        if (mode == 0) {
            state = RG_ORGANIC; t = 0; last_view = -1;
            warp_omega_update(wc.omega, env.K, env.sigma0, true, env.seed, user, 0u, lane);
        } else {
            state = state_io[i]; t = (uint32_t)t_io[i]; last_view = last_view_io[i];
            // the caller's state holds the event index only: the clock value is re-accumulated (compatibility path)
            clock = clock_at(env.time_mode, env.seed, user, t, env.time_mu, env.time_sigma);
            for (int k = lane; k < env.K; k += 32) wc.omega[k] = omega_io[(size_t)i * env.K + k];
            __syncwarp();
        }
        if (mode == 1) {
            int r = 0;
            if (state == RG_BANDIT) r = bandit_event(env, wc, user, action[i], state, t, clock, lane);
            if (lane == 0) reward[i] = r;
        }
        if (mode != 1 || state == RG_ORGANIC) {
            // generate_organic_sessions (abstract.py:105-121), at most sess_cap views per call
            while (state == RG_ORGANIC && len < sess_cap) {
                const int v = organic_event(env, wc, user, state, t, clock, lane, ambig);
                last_view = v;
                if (lane == 0) sess_v[(size_t)i * sess_cap + len] = v;
                ++len;
            }
        }
        if (state == RG_ORGANIC) unfinished += 1;
        if (lane == 0) {
            state_io[i] = state; t_io[i] = (int32_t)t; last_view_io[i] = last_view; sess_len[i] = len;
        }
        for (int k = lane; k < env.K; k += 32) omega_io[(size_t)i * env.K + k] = wc.omega[k];
        __syncwarp();
    }
    if (lane == 0) {
        if (unfinished && n_unfinished) atomicAdd(n_unfinished, unfinished);
        if (ambig && diag) atomicAdd(diag, (unsigned long long)ambig);
    }
}

// ---------------------------------------------------------------------------------------------
// Device evaluation of the scalar spec functions (tests: device == host == oracle bit-for-bit).
// kind 0 exp, 1 log, 2 norm_ppf, 3 ff, 4 normal from (seed=in[3i], user=in[3i+1], t=in[3i+2]) first normal,
// 5 event uniform 0, 6 event uniform 1, 7 uniform policy action with P = (int)in[3i+2]... (see rg_capi.cu)
__global__ void spec_eval_kernel(int kind, long long n, const double *__restrict__ in, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind <= 3) {
        const double x = in[i];
        out[i] = kind == 0 ? exp64(x) : kind == 1 ? log64(x) : kind == 2 ? norm_ppf(x) : ff64(x);
        return;
    }
    const uint64_t seed = (uint64_t)in[3 * i], user = (uint64_t)in[3 * i + 1];
    const uint32_t t = (uint32_t)in[3 * i + 2];
    if (kind == 4) {
        const u32x4 w = draw(seed, user, t, SLOT_NORMAL);
        out[2 * i] = normal_from_words(w.x, w.y);
        out[2 * i + 1] = normal_from_words(w.z, w.w);
    } else if (kind == 5) {
        const u32x4 w = draw(seed, user, t, SLOT_EVENT);
        out[2 * i] = uniform53(w.x, w.y);
        out[2 * i + 1] = uniform53(w.z, w.w);
    } else {
        const u32x4 w = draw(seed, user, t, SLOT_POLICY);
        out[2 * i] = (double)uniform_action(w.x, 1000003);
        out[2 * i + 1] = (double)w.w;
    }
}

cudaError_t launch_spec_eval(int kind, long long n, const double *in, double *out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    spec_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(kind, n, in, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launch helpers (called from rg_capi.cu)
cudaError_t launch_rollout_exact(const rg_env_impl &e, const PolicyDev &pol, long long user_begin,
                                 long long n_organic_users, long long n_total, const RolloutOutDev &out,
                                 cudaStream_t stream) {
    if (n_total <= 0) return cudaSuccess;
    const size_t smem = warp_kernel_smem_bytes(e.dev);
    long long blocks = (n_total + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long max_blocks = (long long)e.num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    rollout_warp_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(e.dev, pol, user_begin,
                                                                                n_organic_users, n_total, out);
    return cudaGetLastError();
}

cudaError_t launch_batch(const rg_env_impl &e, int mode, long long N, const long long *user_id,
                         const int32_t *action, int32_t *state, int32_t *t, int32_t *last_view, double *omega,
                         int32_t *reward, int32_t *sess_v, int32_t *sess_len, int sess_cap, int32_t *n_unfinished,
                         cudaStream_t stream) {
    if (N <= 0) return cudaSuccess;
    const size_t smem = warp_kernel_smem_bytes(e.dev);
    long long blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long max_blocks = (long long)e.num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    batch_step_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, stream>>>(
        e.dev, mode, N, user_id, action, state, t, last_view, omega, reward, sess_v, sess_len, sess_cap,
        n_unfinished, e.diag);
    return cudaGetLastError();
}

}  // namespace rg
