// rg_env0.cu -- reco-gym-v0 (recogym/envs/reco_env_v0.py:17-69 on the loops of recogym/envs/abstract.py:90-197).
// The toy environment has no latent state: an episode is a chain of table look-ups driven by the counter RNG, so one
// THREAD simulates one user; the tables (cumulative distributions, click thresholds) are built by the host with NumPy
// exactly as the reference builds them and only searched here (include/recogym_b200.h, "reco-gym-v0").
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "rg_host.h"
#include "rg_math.cuh"

namespace rg {
namespace v0 {

constexpr uint32_t SLOT_V0_INIT = 3u;   // counter 0, words 0,1: the first product view of reset()
constexpr int ORGANIC = 0, BANDIT = 1, STOP = 2;

struct Env0Dev {
    int P, C;
    const double *init_cdf, *view_cdf, *click_thr;
    const int32_t *cluster_of;
    double trans_cdf[9];
    uint64_t seed;
};
struct Pol0 {
    int kind;
    const int32_t *table;   // RG0_POLICY_TABLE: [P]; RG0_POLICY_REPLAY: the actions
    long long n;
};
struct Out0 {
    unsigned long long *totals;
    int32_t *n_events, *n_organic, *n_bandit, *n_clicks, *last_view, *state;
    int32_t *log;
    long long log_capacity;
    unsigned long long *log_count;
};

// numpy.searchsorted(cdf, u, side='right') == the number of entries <= u
RG_HD int cdf_index(const double *cdf, int n, double u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Row sink of an episode: {user - user_begin, event index, product | action, flags}.  On the device the rows of different
// users interleave (one atomic slot reservation per row); the host restores (user, event index) order.
struct LogSink {
    int32_t *log;
    long long capacity;
    unsigned long long *count;
    long long u_rel;
    RG_HD void operator()(uint32_t e, int x, int flags) const {
        if (!log) return;
#ifdef __CUDA_ARCH__
        const unsigned long long i = atomicAdd(count, 1ull);
#else
        const unsigned long long i = (*count)++;
#endif
        if ((long long)i < capacity) {
            int32_t *row = log + 4 * i;
            row[0] = (int32_t)u_rel; row[1] = (int32_t)e; row[2] = x; row[3] = flags;
        }
    }
};

// One user's episode; `emit` receives the rows of generate_logs in order.
RG_HD void user_episode(const Env0Dev &env, const Pol0 &pol, uint64_t user, bool organic_only, const LogSink &emit, int &n_org,
                        int &n_ban, int &n_clk, int &n_phantom, int &view_out, int &state_out) {
    int state = ORGANIC;
    uint32_t e = 0;          // event index: update_state() calls so far
    long long na = 0;        // replayed actions consumed
    n_org = n_ban = n_clk = n_phantom = 0;
    // reset (reco_env_v0.py:50-56): the first product view from initial_product_probs
    const u32x4 w0 = draw(env.seed, user, 0u, SLOT_V0_INIT);
    int view = cdf_index(env.init_cdf, env.P, uniform53(w0.x, w0.y));
    bool waiting = false;
    for (;;) {
        while (state == ORGANIC) {   // generate_organic_sessions (abstract.py:105-121)
            const u32x4 w = draw(env.seed, user, e, SLOT_EVENT);
            view = cdf_index(env.view_cdf + (size_t)env.cluster_of[view] * env.P, env.P, uniform53(w.x, w.y));   // :67-69
            emit(e, view, 0);
            n_org += 1;
            state = cdf_index(env.trans_cdf + 3 * ORGANIC, 3, uniform53(w.z, w.w));                              // :58-60
            e += 1;
        }
        if (state == STOP || organic_only) break;
        // one bandit event: step_offline -> act, step -> draw_click, update_state (abstract.py:199-239, 174-197)
        int a;
        if (pol.kind == RG0_POLICY_REPLAY) {
            if (na >= pol.n) { waiting = true; break; }
            a = pol.table[na++];
        } else if (pol.kind == RG0_POLICY_TABLE) {
            a = pol.table[view];
        } else {
            a = uniform_action(draw(env.seed, user, e, SLOT_POLICY).x, env.P);
        }
        const u32x4 w = draw(env.seed, user, e, SLOT_EVENT);
        const double thr = env.click_thr[(size_t)a * env.P + view], u = uniform53(w.x, w.y);
        const int c = thr < 0.0 ? (u > -thr ? 0 : 1) : (u > thr ? 1 : 0);   // rng.binomial(1, p) (:62-65)
        state = cdf_index(env.trans_cdf + 3 * BANDIT, 3, uniform53(w.z, w.w));
        if (c) state = ORGANIC;       // abstract.py:180-181: after a click organic events always follow
        emit(e, a, 1 | (c << 1));
        e += 1;
        n_ban += 1;
        n_clk += c;
    }
    if (!organic_only && !waiting && pol.kind != RG0_POLICY_REPLAY) {
        // the phantom row of generate_logs: after `done` the agent acts once more, reward 0 (abstract.py:311-316)
        const int a = pol.kind == RG0_POLICY_TABLE ? pol.table[view] : uniform_action(draw(env.seed, user, e, SLOT_POLICY).x, env.P);
        emit(e, a, 1 | 4);
        n_phantom = 1;
    }
    view_out = view;
    state_out = waiting ? BANDIT : (organic_only && state != STOP ? state : STOP);
}

__global__ void __launch_bounds__(128) rollout_v0_kernel(Env0Dev env, Pol0 pol, long long user_begin,
                                                          long long n_organic_users, long long n_total, Out0 out) {
    unsigned long long t_ev = 0, t_org = 0, t_ban = 0, t_clk = 0, t_ph = 0, t_users = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
        int n_org, n_ban, n_clk, n_ph, view, state;
        user_episode(env, pol, (uint64_t)(user_begin + i), i < n_organic_users, LogSink{out.log, out.log_capacity, out.log_count, i},
                     n_org, n_ban, n_clk, n_ph, view, state);
        if (out.n_events) out.n_events[i] = n_org + n_ban;
        if (out.n_organic) out.n_organic[i] = n_org;
        if (out.n_bandit) out.n_bandit[i] = n_ban;
        if (out.n_clicks) out.n_clicks[i] = n_clk;
        if (out.last_view) out.last_view[i] = view;
        if (out.state) out.state[i] = state;
        t_users += 1; t_ev += n_org + n_ban; t_org += n_org; t_ban += n_ban; t_clk += n_clk; t_ph += n_ph;
    }
    // warp-level reduction, one atomic per warp and counter
    const unsigned long long vals[6] = {t_users, t_ev, t_org, t_ban, t_clk, t_ph};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        unsigned long long v = vals[k];
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out.totals + k, v);
    }
}

}  // namespace v0
}  // namespace rg

struct rg_env0 {
    rg::v0::Env0Dev dev{};
    int device = 0, num_sms = 0;
    double *d_init = nullptr, *d_view = nullptr, *d_thr = nullptr;
    int32_t *d_cluster = nullptr;
};

namespace {

using rg::v0::Env0Dev;
using rg::v0::Out0;
using rg::v0::Pol0;

#define RG0_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess) {                                                                          \
            char msg__[384];                                                                                 \
            snprintf(msg__, sizeof(msg__), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return rg::set_error(RG_ECUDA, msg__);                                                           \
        }                                                                                                    \
    } while (0)

int check_desc(const rg_env0_desc *d) {
    if (!d || !d->init_cdf || !d->view_cdf || !d->cluster_of || !d->click_thr) return rg::set_error(RG_EINVAL, "env0: null table");
    if (d->num_products < 1 || d->num_clusters < 1 || d->num_clusters > d->num_products)
        return rg::set_error(RG_EINVAL, "env0: num_products >= num_clusters >= 1 is required");
    for (int p = 0; p < d->num_products; ++p)
        if (d->cluster_of[p] < 0 || d->cluster_of[p] >= d->num_clusters) return rg::set_error(RG_EINVAL, "env0: cluster_of out of range");
    for (int s = 0; s < 3; ++s)
        if (!(d->trans_cdf[3 * s + 2] == 1.0) || !(d->trans_cdf[3 * s] <= d->trans_cdf[3 * s + 1] && d->trans_cdf[3 * s + 1] <= 1.0))
            return rg::set_error(RG_EINVAL, "env0: trans_cdf rows must be non-decreasing and end at exactly 1");
    if (!(d->init_cdf[d->num_products - 1] == 1.0)) return rg::set_error(RG_EINVAL, "env0: init_cdf must end at exactly 1");
    return RG_OK;
}

int check_call(int policy_kind, const int32_t *table, int64_t n_table, int P, int64_t n_organic_users, int64_t n_users,
               const rg_env0_out *out) {
    if (!out || !out->totals) return rg::set_error(RG_EINVAL, "env0: out and out->totals are required");
    if (n_users < 0 || n_organic_users < 0) return rg::set_error(RG_EINVAL, "env0: negative user count");
    if (policy_kind == RG0_POLICY_TABLE) {
        if (!table || n_table != P) return rg::set_error(RG_EINVAL, "env0: a table policy needs table[P]");
    } else if (policy_kind == RG0_POLICY_REPLAY) {
        if (n_users != 1 || n_organic_users != 0 || n_table < 0 || (n_table > 0 && !table))
            return rg::set_error(RG_EINVAL, "env0: a replay is ONE regular user with its list of actions");
    } else if (policy_kind != RG0_POLICY_UNIFORM) {
        return rg::set_error(RG_EINVAL, "env0: unknown policy kind");
    }
    if (out->log && (!out->log_count || out->log_capacity < 0)) return rg::set_error(RG_EINVAL, "env0: log needs log_count and a capacity");
    return RG_OK;
}

Out0 to_out(const rg_env0_out *o) {
    Out0 r{};
    r.totals = (unsigned long long *)o->totals;
    r.n_events = o->n_events; r.n_organic = o->n_organic; r.n_bandit = o->n_bandit; r.n_clicks = o->n_clicks;
    r.last_view = o->last_view; r.state = o->state;
    r.log = o->log; r.log_capacity = o->log_capacity; r.log_count = (unsigned long long *)o->log_count;
    return r;
}

}  // namespace

int rg_env0_create(const rg_env0_desc *d, rg_env0 **out) {
    if (!out) return rg::set_error(RG_EINVAL, "env0: null out");
    *out = nullptr;
    int rc = check_desc(d);
    if (rc != RG_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || d->device < 0 || d->device >= ndev)
        return rg::set_error(RG_ENODEVICE, "env0: no usable CUDA device (there is no CPU path)");
    RG0_CUDA(cudaSetDevice(d->device));
    cudaDeviceProp prop;
    RG0_CUDA(cudaGetDeviceProperties(&prop, d->device));
    rg_env0 *e = new (std::nothrow) rg_env0();
    if (!e) return rg::set_error(RG_ENOMEM, "env0: out of memory");
    const size_t P = (size_t)d->num_products, Cn = (size_t)d->num_clusters;
    e->device = d->device;
    e->num_sms = prop.multiProcessorCount;
    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t ce = cudaMalloc(dst, bytes);
        if (ce != cudaSuccess) return ce;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t ce = up((void **)&e->d_init, d->init_cdf, sizeof(double) * P);
    if (ce == cudaSuccess) ce = up((void **)&e->d_view, d->view_cdf, sizeof(double) * Cn * P);
    if (ce == cudaSuccess) ce = up((void **)&e->d_thr, d->click_thr, sizeof(double) * P * P);
    if (ce == cudaSuccess) ce = up((void **)&e->d_cluster, d->cluster_of, sizeof(int32_t) * P);
    if (ce != cudaSuccess) {
        rg_env0_destroy(e);
        char msg[256];
        snprintf(msg, sizeof(msg), "env0: parameter upload failed: %s", cudaGetErrorString(ce));
        return rg::set_error(RG_ECUDA, msg);
    }
    e->dev.P = d->num_products; e->dev.C = d->num_clusters;
    e->dev.init_cdf = e->d_init; e->dev.view_cdf = e->d_view; e->dev.click_thr = e->d_thr; e->dev.cluster_of = e->d_cluster;
    std::memcpy(e->dev.trans_cdf, d->trans_cdf, sizeof(e->dev.trans_cdf));
    e->dev.seed = d->seed;
    *out = e;
    return RG_OK;
}

void rg_env0_destroy(rg_env0 *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_init); cudaFree(e->d_view); cudaFree(e->d_thr); cudaFree(e->d_cluster);
    delete e;
}

int rg_env0_set_seed(rg_env0 *e, uint64_t seed) {
    if (!e) return rg::set_error(RG_EINVAL, "env0: null env");
    e->dev.seed = seed;
    return RG_OK;
}

int rg_env0_rollout_device(rg_env0 *e, int policy_kind, const int32_t *table, int64_t n_table, int64_t user_begin,
                           int64_t n_organic_users, int64_t n_users, const rg_env0_out *out, void *stream) {
    if (!e) return rg::set_error(RG_EINVAL, "env0: null env");
    int rc = check_call(policy_kind, table, n_table, e->dev.P, n_organic_users, n_users, out);
    if (rc != RG_OK) return rc;
    const long long n_total = n_organic_users + n_users;
    if (n_total == 0) return RG_OK;
    RG0_CUDA(cudaSetDevice(e->device));
    Pol0 pol{policy_kind, table, (long long)n_table};
    const long long blocks = (n_total + 127) / 128;
    const unsigned grid = (unsigned)(blocks < (long long)e->num_sms * 16 ? blocks : (long long)e->num_sms * 16);
    rg::v0::rollout_v0_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(e->dev, pol, user_begin, n_organic_users, n_total, to_out(out));
    RG0_CUDA(cudaGetLastError());
    return RG_OK;
}

int rg_env0_rollout_host(rg_env0 *e, int policy_kind, const int32_t *table, int64_t n_table, int64_t user_begin,
                         int64_t n_organic_users, int64_t n_users, const rg_env0_out *out) {
    if (!e) return rg::set_error(RG_EINVAL, "env0: null env");
    int rc = check_call(policy_kind, table, n_table, e->dev.P, n_organic_users, n_users, out);
    if (rc != RG_OK) return rc;
    const long long n = n_organic_users + n_users;
    RG0_CUDA(cudaSetDevice(e->device));
    // one device slab: totals[8] | log_count | per-user arrays [6][n] | table | log
    const size_t n_tab = n_table > 0 ? (size_t)n_table : 0, cap = out->log ? (size_t)out->log_capacity : 0;
    const size_t off_cnt = 8 * 8, off_pu = off_cnt + 8, off_tab = off_pu + 6 * 4 * (size_t)n;
    const size_t off_log = (off_tab + 4 * n_tab + 15) & ~(size_t)15, bytes = off_log + 16 * cap;
    unsigned char *slab = nullptr;
    RG0_CUDA(cudaMalloc((void **)&slab, bytes ? bytes : 16));
    struct Free { unsigned char *p; ~Free() { cudaFree(p); } } guard{slab};
    RG0_CUDA(cudaMemset(slab, 0, off_pu));
    if (n_tab) RG0_CUDA(cudaMemcpy(slab + off_tab, table, 4 * n_tab, cudaMemcpyHostToDevice));
    rg_env0_out d{};
    d.totals = (int64_t *)slab;
    d.log_count = (int64_t *)(slab + off_cnt);
    int32_t *pu = (int32_t *)(slab + off_pu);
    int32_t *const host_pu[6] = {out->n_events, out->n_organic, out->n_bandit, out->n_clicks, out->last_view, out->state};
    int32_t **dev_pu[6] = {&d.n_events, &d.n_organic, &d.n_bandit, &d.n_clicks, &d.last_view, &d.state};
    for (int k = 0; k < 6; ++k) *dev_pu[k] = host_pu[k] ? pu + (size_t)k * n : nullptr;
    d.log = out->log ? (int32_t *)(slab + off_log) : nullptr;
    d.log_capacity = out->log_capacity;
    rc = rg_env0_rollout_device(e, policy_kind, n_tab ? (const int32_t *)(slab + off_tab) : nullptr, n_table, user_begin,
                                n_organic_users, n_users, &d, nullptr);
    if (rc != RG_OK) return rc;
    RG0_CUDA(cudaDeviceSynchronize());
    int64_t tot[8];
    RG0_CUDA(cudaMemcpy(tot, slab, sizeof(tot), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out->totals[k] += tot[k];
    for (int k = 0; k < 6; ++k)
        if (host_pu[k] && n) RG0_CUDA(cudaMemcpy(host_pu[k], pu + (size_t)k * n, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    if (out->log) {
        int64_t rows = 0;
        RG0_CUDA(cudaMemcpy(&rows, slab + off_cnt, 8, cudaMemcpyDeviceToHost));
        *out->log_count += rows;
        const size_t have = (size_t)(rows < out->log_capacity ? rows : out->log_capacity);
        if (have) RG0_CUDA(cudaMemcpy(out->log, slab + off_log, 16 * have, cudaMemcpyDeviceToHost));
    }
    return RG_OK;
}

// ---- test hook: the same per-user body on the host (CPU tests of the episode logic; not a product path) ---------------
int rg_spec_env0_rollout(const rg_env0_desc *d, int policy_kind, const int32_t *table, int64_t n_table, int64_t user_begin,
                         int64_t n_organic_users, int64_t n_users, const rg_env0_out *out) {
    int rc = check_desc(d);
    if (rc != RG_OK) return rc;
    rc = check_call(policy_kind, table, n_table, d->num_products, n_organic_users, n_users, out);
    if (rc != RG_OK) return rc;
    Env0Dev env{};
    env.P = d->num_products; env.C = d->num_clusters;
    env.init_cdf = d->init_cdf; env.view_cdf = d->view_cdf; env.click_thr = d->click_thr; env.cluster_of = d->cluster_of;
    std::memcpy(env.trans_cdf, d->trans_cdf, sizeof(env.trans_cdf));
    env.seed = d->seed;
    Pol0 pol{policy_kind, table, (long long)n_table};
    const long long n = n_organic_users + n_users;
    for (long long i = 0; i < n; ++i) {
        int n_org, n_ban, n_clk, n_ph, view, state;
        const rg::v0::LogSink emit{out->log, out->log_capacity, (unsigned long long *)out->log_count, i};
        rg::v0::user_episode(env, pol, (uint64_t)(user_begin + i), i < n_organic_users, emit, n_org, n_ban, n_clk, n_ph, view, state);
        if (out->n_events) out->n_events[i] = n_org + n_ban;
        if (out->n_organic) out->n_organic[i] = n_org;
        if (out->n_bandit) out->n_bandit[i] = n_ban;
        if (out->n_clicks) out->n_clicks[i] = n_clk;
        if (out->last_view) out->last_view[i] = view;
        if (out->state) out->state[i] = state;
        out->totals[0] += 1; out->totals[1] += n_org + n_ban; out->totals[2] += n_org; out->totals[3] += n_ban;
        out->totals[4] += n_clk; out->totals[5] += n_ph;
    }
    return RG_OK;
}
