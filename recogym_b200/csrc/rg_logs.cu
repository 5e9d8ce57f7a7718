// rg_logs.cu -- what sits either side of the rollout, all HBM-bound integer work:
//
//   order    the rollout kernels emit one int32x4 record per generate_logs row in completion order; a user's rows
//            have t = 0,1,2,... without gaps (the phantom row after `done` carries t = n_events), so the reference's
//            row order (abstract.py:293-316: users in id order, rows in time order) is position
//            row_offset[u] + t -- an exclusive scan of the per-user row counts and ONE scatter, no sort.
//   columns  one warp per user walks its ordered rows and writes the DataFrame columns (t,u,z,v,a,c,ps) plus the
//            derived per-row context every consumer needs: the acting last view `lv` (Agent.update_lpv), the last
//            view of the PREVIOUS train call `lvp` (BanditCount.train's row index, bandit_count.py:54-57) and the
//            distance `sess` to the first row of the organic session a view belongs to.
//   train    sufficient statistics of the count agents as scatter-adds: co_counts += b b^T per session
//            (organic_count.py:74-82), pulls/clicks[lvp, a] (bandit_count.py:49-62); row arg-max -> policy table.
//   ips      IPS / SNIPS sums of a deterministic (table) target policy over the logged bandit rows
//            (evaluate_agent.py:753-835), fixed-order reduction.
#include <cmath>

#include "rg_device.cuh"
#include "rg_host.h"

namespace rg {
namespace logs {

constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ long long rows_of_user(const int32_t *n_events, long long i, long long n_organic_users) {
    return (long long)n_events[i] + (i >= n_organic_users ? 1 : 0);  // + the phantom row (abstract.py:311-316)
}

// block-wide inclusive scan of one long long per thread (1024 threads)
__device__ __forceinline__ long long block_scan_incl(long long x, long long *warp_tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        long long t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += y;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    return x + (w > 0 ? warp_tot[w - 1] : 0);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int32_t *__restrict__ n_events, long long n,
                                                                long long n_organic_users, long long *tile_sum) {
    __shared__ long long wt[32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += rows_of_user(n_events, base + k, n_organic_users);
    const long long incl = block_scan_incl(s, wt);
    if (threadIdx.x == SCAN_THREADS - 1) tile_sum[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets(long long *tile_sum, int ntiles, long long *total) {
    // exclusive scan of the tile sums by ONE block, chunk after chunk
    __shared__ long long wt[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < ntiles; c0 += SCAN_THREADS) {
        const int i = c0 + threadIdx.x;
        const long long x = i < ntiles ? tile_sum[i] : 0;
        const long long incl = block_scan_incl(x, wt);
        const long long cbase = carry;
        __syncthreads();
        if (i < ntiles) tile_sum[i] = cbase + incl - x;
        if (threadIdx.x == SCAN_THREADS - 1) carry = cbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const int32_t *__restrict__ n_events, long long n,
                                                            long long n_organic_users,
                                                            const long long *__restrict__ tile_sum,
                                                            long long *__restrict__ row_offset) {
    __shared__ long long wt[32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    long long r[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        r[k] = base + k < n ? rows_of_user(n_events, base + k, n_organic_users) : 0;
        s += r[k];
    }
    long long off = tile_sum[blockIdx.x] + block_scan_incl(s, wt) - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) row_offset[base + k] = off;
        off += r[k];
        if (base + k == n - 1) row_offset[n] = off;
    }
}

// unordered records -> ordered records (16-byte scatter: two rows per 32-byte sector)
__global__ void order_scatter(const int4 *__restrict__ log, const unsigned long long *__restrict__ log_count,
                              long long log_capacity, const long long *__restrict__ row_offset, int4 *__restrict__ ordered,
                              long long ordered_capacity) {
    const long long n = min((long long)*log_count, log_capacity);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int4 r = log[i];
        const long long pos = row_offset[r.x] + (long long)r.y;
        if (pos < ordered_capacity) ordered[pos] = r;
    }
}

struct Columns {
    float *t; int32_t *u; int8_t *z; int32_t *v; int32_t *a; float *c; double *ps; int32_t *lv; int32_t *lvp; int32_t *sess;
    int8_t *flags;
};

// Which observations a training loop of the reference hands to agent.train (train_mode):
//   0  test_agent            (bench_agents.py:169-190): every session, every bandit row incl. the phantom one
//   1  evaluate_agent steps  (evaluate_agent.py:100-139): the loop ends at `done` -- no phantom call, the user's final
//                            session is never trained on
//   2  evaluate_agent warm-up (evaluate_agent.py:56-78): as 1, except that a user whose FIRST observation is already
//                            `done` gets the phantom call (with its only session)
constexpr int FLAG_PHANTOM = 1, FLAG_FINAL_SESSION = 2, FLAG_NO_REAL_BANDIT = 4;

// warp-cooperative: largest row index in [lo, hi) whose record satisfies pred, or -1
template <class F>
__device__ __forceinline__ long long warp_find_last(const int4 *__restrict__ ordered, long long lo, long long hi, int lane,
                                                    F pred) {
    for (long long e = hi; e > lo; e -= 32) {
        const long long r = e - 1 - lane;
        bool hit = false;
        if (r >= lo) hit = pred(ordered[r]);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) return e - 1 - (long long)(__ffs(m) - 1);
    }
    return -1;
}

// BanditCount.last_product_viewed after the train calls of user j under train_mode, or -2 when no call of user j
// updated it (mode 1, user without a real bandit row).
__device__ __forceinline__ int trained_last_view(const int4 *__restrict__ ordered, const long long *__restrict__ row_offset,
                                                 const int32_t *__restrict__ last_view, long long j, long long capacity,
                                                 int train_mode, int lane) {
    if (train_mode == 0) return last_view[j];
    const long long r0 = row_offset[j], r1 = min(row_offset[j + 1], capacity);
    const long long lr = warp_find_last(ordered, r0, r1, lane, [](const int4 &q) { return (q.w & 1) && !(q.w & 4); });
    if (lr < 0) return train_mode == 2 ? last_view[j] : -2;
    const long long lo = warp_find_last(ordered, r0, lr, lane, [](const int4 &q) { return !(q.w & 1); });
    return ordered[lo].z;   // a user's first row is always a view: lo >= r0
}

// One warp per user: walks the user's ordered rows 32 at a time.
//   lv   = view of the latest organic row at or before this row (the acting last view of a bandit row)
//   lvp  = lv of the previous bandit row of this user; for the user's first bandit row the final last view of the
//          previous regular user (BanditCount keeps last_product_viewed across users, bandit_count.py:20,54-57);
//          -1 = None (nothing trained yet).  Organic rows: -1.
//   sess = distance to the first row of the organic session this view belongs to (0 for the first); bandit rows: 0.
__global__ void __launch_bounds__(256) columns_kernel(const int4 *__restrict__ ordered,
                                                      const long long *__restrict__ row_offset, long long n_total,
                                                      long long n_organic_users, long long user_begin,
                                                      const int32_t *__restrict__ last_view, int prev_last_view,
                                                      int policy_kind, int P, const double *__restrict__ ps_table,
                                                      const int32_t *__restrict__ table, double epsilon,
                                                      int time_mode, uint64_t seed, double time_mu, double time_sigma,
                                                      long long capacity, int train_mode, Columns col) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float fnan = __int_as_float(0x7fc00000);
    const double dnan = __longlong_as_double(0x7ff8000000000000ll);
    // The log did not fit (the caller reads the row count and regenerates into larger columns): order_scatter placed only
    // the first `capacity` records, the rest of `ordered` is uninitialised -- nothing here may be derived from it.
    if (row_offset[n_total] > capacity) return;
    for (long long i = warp; i < n_total; i += nwarps) {
        const long long r0 = row_offset[i], r1 = min(row_offset[i + 1], capacity);
        int carry_lv = -1;                               // last view before this chunk
        // lv of the previous train call: the trained last view of the closest earlier regular user that has one
        int carry_blv = -1;
        if (i >= n_organic_users) {                      // organic-only users never reach a bandit row
            carry_blv = prev_last_view;
            for (long long j = i - 1; j >= n_organic_users; --j) {
                const int x = trained_last_view(ordered, row_offset, last_view, j, capacity, train_mode, lane);
                if (x != -2) { carry_blv = x; break; }
            }
        }
        const long long last_real = warp_find_last(ordered, r0, r1, lane, [](const int4 &q) { return (q.w & 1) && !(q.w & 4); });
        int run = 0;                                     // organic rows immediately before this chunk (open session)
        double clock = 0.0;                              // NormalTimeGenerator: the clock after the rows before this chunk
        for (long long b = r0; b < r1; b += 32) {
            const long long r = b + lane;
            const bool ok = r < r1;
            int4 rec = make_int4(0, 0, 0, 1);
            if (ok) rec = ordered[r];
            // The row's clock value.  Default clock: the event index.  NormalTimeGenerator (normal_time_generator.py:25-28,
            // abstract.py:100, reco_env_v1.py:91): the clock of event index t is the SEQUENTIAL binary64 sum of the steps
            // of the calls 0 .. t-1 -- every lane draws its row's step, the warp then accumulates them one after the
            // other, each lane keeping the value at its own row.
            float tval = (float)rec.y;
            if (time_mode == 1) {
                const double st = (ok && rec.y > 0) ? time_step(seed, (uint64_t)(user_begin + i), (uint32_t)rec.y - 1u, time_mu, time_sigma) : 0.0;
                double mine = clock;
                for (int k = 0; k < 32; ++k) {
                    clock += __shfl_sync(0xffffffffu, st, k);
                    if (lane == k) mine = clock;
                }
                tval = (float)mine;
            }
            const bool org = ok && !(rec.w & 1);
            const bool ban = ok && (rec.w & 1);
            const unsigned om = __ballot_sync(0xffffffffu, org);
            const unsigned bm = __ballot_sync(0xffffffffu, ban);
            const unsigned le = 0xffffffffu >> (31 - lane);   // lanes <= me
            const unsigned lt = le >> 1;                        // lanes < me
            // acting last view
            const unsigned omle = om & le;
            const int src = omle ? 31 - __clz(omle) : 0;
            const int vsrc = __shfl_sync(0xffffffffu, rec.z, src);
            const int lv = omle ? vsrc : carry_lv;
            // previous bandit row's lv
            const unsigned bmlt = bm & lt;
            const int bsrc = bmlt ? 31 - __clz(bmlt) : 0;
            const int lv_at_b = __shfl_sync(0xffffffffu, lv, bsrc);
            const int lvp = bmlt ? lv_at_b : carry_blv;
            // distance to the session start: organic rows since the last bandit row
            const int since = bmlt ? lane - (31 - __clz(bmlt)) - 1 : lane + run;
            if (ok) {
                if (col.t) col.t[r] = tval;
                if (col.u) col.u[r] = (int32_t)(user_begin + i);
                if (col.z) col.z[r] = org ? 0 : 1;
                if (col.v) col.v[r] = org ? rec.z : -1;
                if (col.a) col.a[r] = org ? -1 : rec.z;
                if (col.c) col.c[r] = org ? fnan : (float)((rec.w >> 1) & 1);
                if (col.lv) col.lv[r] = lv;
                if (col.lvp) col.lvp[r] = org ? -1 : lvp;
                if (col.sess) col.sess[r] = org ? since : 0;
                if (col.flags)
                    col.flags[r] = (int8_t)(((rec.w & 4) ? FLAG_PHANTOM : 0) | ((org && r > last_real) ? FLAG_FINAL_SESSION : 0) |
                                            (last_real < 0 ? FLAG_NO_REAL_BANDIT : 0));
                if (col.ps) {
                    double ps = dnan;
                    if (!org) {
                        if (policy_kind == RG_POLICY_UNIFORM) ps = 1.0 / (double)P;   // abstract.py:219
                        else if (policy_kind == RG_POLICY_TABLE) ps = ps_table ? (lv >= 0 && lv < P ? ps_table[lv] : dnan) : 1.0;
                        else if (policy_kind == RG_POLICY_EPS_TABLE && lv >= 0 && lv < P) {
                            // epsilon_greedy.py:44-68: exploit -> (1 - eps) * the wrapped agent's ps; explore -> eps * p[a],
                            // p = ones with a hole / sum = 1 / (P - 1).  An action is greedy iff it is table[lv].
                            const double ps_greedy = ps_table ? ps_table[lv] : 1.0;
                            ps = rec.z == table[lv] ? (1.0 - epsilon) * ps_greedy : epsilon * (1.0 / (double)(P - 1));
                        }
                    }
                    col.ps[r] = ps;
                }
            }
            // carries into the next chunk
            const int last_lv = __shfl_sync(0xffffffffu, lv, 31);
            carry_lv = last_lv;
            if (bm) {
                const int hb = 31 - __clz(bm);
                carry_blv = __shfl_sync(0xffffffffu, lv, hb);
                run = 31 - hb;
            } else {
                run += 32;
            }
        }
    }
}

// BanditCount.last_product_viewed after the whole log (the next log's prev_last_view): one warp
__global__ void carry_kernel(const int4 *__restrict__ ordered, const long long *__restrict__ row_offset,
                             const int32_t *__restrict__ last_view, long long n_total, long long n_organic_users,
                             int prev_last_view, long long capacity, int train_mode, int32_t *__restrict__ carry_out) {
    const int lane = threadIdx.x & 31;
    int x = prev_last_view;
    if (row_offset[n_total] > capacity) {   // incomplete log (see columns_kernel): the caller regenerates it
        if (lane == 0) carry_out[0] = prev_last_view;
        return;
    }
    for (long long j = n_total - 1; j >= n_organic_users; --j) {
        const int t = trained_last_view(ordered, row_offset, last_view, j, capacity, train_mode, lane);
        if (t != -2) { x = t; break; }
    }
    if (lane == 0) carry_out[0] = x;
}

// ---- training statistics ----------------------------------------------------------------------------------------
// OrganicCount.train (organic_count.py:74-82): per observation, b = view counts of its sessions, co_counts += b b^T.
// Row r (a view, `sess` rows into its session) adds the ordered pairs (r, j) and (j, r) for every earlier row j of the
// session, and (r, r).
__global__ void organic_pairs_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ v,
                                     const int32_t *__restrict__ sess, const int8_t *__restrict__ flags, int train_mode,
                                     long long rows, int P, unsigned long long *__restrict__ co) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        if (z[r] != 0) continue;
        if (train_mode != 0) {   // the final session is only trained on by the warm-up loop's phantom call
            const int f = flags[r];
            if ((f & FLAG_FINAL_SESSION) && !(train_mode == 2 && (f & FLAG_NO_REAL_BANDIT))) continue;
        }
        const int vr = v[r], s = sess[r];
        atomicAdd(co + (size_t)vr * P + vr, 1ull);
        for (int k = 1; k <= s; ++k) {
            const int vj = v[r - k];
            atomicAdd(co + (size_t)vr * P + vj, 1ull);
            atomicAdd(co + (size_t)vj * P + vr, 1ull);
        }
    }
}

// BanditCount.train (bandit_count.py:49-62): pulls[ix, a] += 1, clicks[ix, a] += reward with ix = lvp.
// ix None (lvp < 0, only before the very first train call): numpy reads pulls_a[None, a] as pulls_a[np.newaxis, a], i.e.
// the whole ROW a gets the increment (pinned by tests/golden/t2_train_*.npz).
__global__ void bandit_stats_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ a,
                                    const float *__restrict__ c, const int32_t *__restrict__ lvp,
                                    const int8_t *__restrict__ flags, int train_mode, long long rows, int P,
                                    unsigned long long *__restrict__ pulls, unsigned long long *__restrict__ clicks) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        if (z[r] != 1) continue;
        if (train_mode != 0) {
            const int f = flags[r];
            if ((f & FLAG_PHANTOM) && !(train_mode == 2 && (f & FLAG_NO_REAL_BANDIT))) continue;
        }
        const int ar = a[r], ix = lvp[r];
        const int rew = c[r] > 0.5f ? 1 : 0;
        if (ix >= 0) {
            atomicAdd(pulls + (size_t)ix * P + ar, 1ull);
            if (rew) atomicAdd(clicks + (size_t)ix * P + ar, 1ull);
        } else {
            for (int p = 0; p < P; ++p) {
                atomicAdd(pulls + (size_t)ar * P + p, 1ull);
                if (rew) atomicAdd(clicks + (size_t)ar * P + p, 1ull);
            }
        }
    }
}

// Row arg-max (first maximum wins, as numpy.argmax): one warp per row.
//   kind 0: OrganicCount.act  -> argmax_j co[v][j]                       (organic_count.py:56)
//   kind 1: BanditCount.act   -> argmax_j (clicks[v][j]+1)/(pulls[v][j]+2), ps = that ratio (bandit_count.py:30-33)
__global__ void __launch_bounds__(256) row_argmax_kernel(int kind, int P, const int64_t *__restrict__ A,
                                                         const int64_t *__restrict__ B, int32_t *__restrict__ table,
                                                         double *__restrict__ ps) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= P) return;
    double best = -1.0;
    int bj = 0x7fffffff;
    for (int j = lane; j < P; j += 32) {
        double x;
        if (kind == 0) x = (double)A[(size_t)row * P + j];
        else x = ((double)B[(size_t)row * P + j] + 1.0) / ((double)A[(size_t)row * P + j] + 2.0);
        if (x > best) { best = x; bj = j; }     // j increases per lane: the first maximum of this lane
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
        if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) {
        table[row] = bj;
        if (ps) ps[row] = kind == 0 ? 1.0 : best;
    }
}

// ---- the same statistics in a device hash table (num_products too large for [P][P]: BASELINE configs[3], P = 100 000) ----
// Open addressing, linear probing; key = row * 2^32 + col, empty = ~0; values int64 {a, b} accumulated with atomics.
// stats[0] counts occupied slots, stats[1] insertions that found no slot (the caller grows the table and re-trains).
__device__ __forceinline__ void sparse_add(const rg_sparse_counts &t, int row, int col, unsigned long long da,
                                           unsigned long long db) {
    const unsigned long long key = ((unsigned long long)(unsigned)row << 32) | (unsigned)col;
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    const unsigned long long mask = (unsigned long long)t.capacity - 1ull;
    unsigned long long slot = h & mask;
    for (long long probe = 0; probe < t.capacity; ++probe) {
        unsigned long long k = t.keys[slot];
        if (k != key) {
            if (k != ~0ull) { slot = (slot + 1ull) & mask; continue; }
            k = atomicCAS((unsigned long long *)t.keys + slot, ~0ull, key);
            if (k == ~0ull) atomicAdd((unsigned long long *)t.stats, 1ull);
            else if (k != key) { slot = (slot + 1ull) & mask; continue; }
        }
        if (da) atomicAdd((unsigned long long *)t.a + slot, da);
        if (db && t.b) atomicAdd((unsigned long long *)t.b + slot, db);
        return;
    }
    atomicAdd((unsigned long long *)t.stats + 1, 1ull);
}

__global__ void sparse_organic_pairs_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ v,
                                            const int32_t *__restrict__ sess, const int8_t *__restrict__ flags,
                                            int train_mode, long long rows, const rg_sparse_counts co) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        if (z[r] != 0) continue;
        if (train_mode != 0) {
            const int f = flags[r];
            if ((f & FLAG_FINAL_SESSION) && !(train_mode == 2 && (f & FLAG_NO_REAL_BANDIT))) continue;
        }
        const int vr = v[r], s = sess[r];
        sparse_add(co, vr, vr, 1ull, 0ull);
        for (int k = 1; k <= s; ++k) {
            const int vj = v[r - k];
            sparse_add(co, vr, vj, 1ull, 0ull);
            sparse_add(co, vj, vr, 1ull, 0ull);
        }
    }
}

__global__ void sparse_bandit_stats_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ a,
                                           const float *__restrict__ c, const int32_t *__restrict__ lvp,
                                           const int8_t *__restrict__ flags, int train_mode, long long rows,
                                           const rg_sparse_counts t) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        if (z[r] != 1) continue;
        if (train_mode != 0) {
            const int f = flags[r];
            if ((f & FLAG_PHANTOM) && !(train_mode == 2 && (f & FLAG_NO_REAL_BANDIT))) continue;
        }
        const int ar = a[r], ix = lvp[r];
        const unsigned long long rew = c[r] > 0.5f ? 1ull : 0ull;
        if (ix >= 0) {
            sparse_add(t, ix, ar, 1ull, rew);
        } else {   // last_product_viewed still None: numpy increments the whole ROW ar (see bandit_stats_kernel)
            atomicAdd((unsigned long long *)t.row_a + ar, 1ull);
            if (rew) atomicAdd((unsigned long long *)t.row_b + ar, 1ull);
        }
    }
}

// growth: every occupied cell of `from` is added to `to` (a larger table) through the same insertion path
__global__ void sparse_rehash_kernel(const rg_sparse_counts from, const rg_sparse_counts to) {
    for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < from.capacity;
         slot += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = from.keys[slot];
        if (key == ~0ull) continue;
        sparse_add(to, (int)(key >> 32), (int)(key & 0xFFFFFFFFull), (unsigned long long)from.a[slot],
                   from.b ? (unsigned long long)from.b[slot] : 0ull);
    }
}

// Acting table from the sparse cells == numpy's argmax over the dense row (first maximum wins):
//   kind 0  OrganicCount: the largest count of the row, lowest column among ties; a row without cells -> 0;
//   kind 1  BanditCount: every untouched cell of row i holds the base value b_i = (rc_i + 1) / (rp_i + 2) (rc, rp = the
//           whole-row increments; 0.5 without them).  The best touched cell wins when it beats b_i or when every cell
//           of the row has been touched (lowest column among ties); otherwise the answer is the lowest column whose cell
//           is not BELOW b_i -- found with a bitmap of the "below" cells among the first 64 columns (a row whose first
//           64 cells are all below base is reported in stats[1]: it cannot be resolved here).
__device__ __forceinline__ double sparse_value(int kind, const rg_sparse_counts &s, long long slot, int row) {
    if (kind == 0) return (double)s.a[slot];
    const double rp = s.row_a ? (double)s.row_a[row] : 0.0, rc = s.row_b ? (double)s.row_b[row] : 0.0;
    return ((double)s.b[slot] + rc + 1.0) / ((double)s.a[slot] + rp + 2.0);
}
__global__ void sparse_policy_init_kernel(int P, unsigned long long *__restrict__ colcnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) colcnt[i] = 0xFFFFFFFF00000000ull;    // high word: best column (min), low word: touched cells (count)
}
__global__ void sparse_policy_best_kernel(int kind, const rg_sparse_counts s, unsigned long long *__restrict__ best,
                                          unsigned long long *__restrict__ blocked, unsigned long long *__restrict__ colcnt) {
    for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < s.capacity;
         slot += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = s.keys[slot];
        if (key == ~0ull) continue;
        const int row = (int)(key >> 32), col = (int)(key & 0xFFFFFFFFull);
        const double val = sparse_value(kind, s, slot, row);
        atomicMax(best + row, (unsigned long long)__double_as_longlong(val));   // val > 0: the bit pattern is monotone
        atomicAdd(colcnt + row, 1ull);
        if (kind == 1 && col < 64) {
            const double rp = s.row_a ? (double)s.row_a[row] : 0.0, rc = s.row_b ? (double)s.row_b[row] : 0.0;
            if (val < (rc + 1.0) / (rp + 2.0)) atomicOr(blocked + row, 1ull << col);
        }
    }
}
__global__ void sparse_policy_col_kernel(int kind, const rg_sparse_counts s, const unsigned long long *__restrict__ best,
                                         unsigned long long *__restrict__ colcnt) {
    for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < s.capacity;
         slot += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = s.keys[slot];
        if (key == ~0ull) continue;
        const int row = (int)(key >> 32);
        const unsigned long long col = key & 0xFFFFFFFFull;
        if ((unsigned long long)__double_as_longlong(sparse_value(kind, s, slot, row)) != best[row]) continue;
        // min over the high word, the low word (count) is left untouched: CAS loop
        unsigned long long old = colcnt[row];
        while ((old >> 32) > col) {
            const unsigned long long want = (col << 32) | (old & 0xFFFFFFFFull);
            const unsigned long long got = atomicCAS(colcnt + row, old, want);
            if (got == old) break;
            old = got;
        }
    }
}
__global__ void sparse_policy_final_kernel(int kind, int P, const rg_sparse_counts s, const unsigned long long *__restrict__ best,
                                           const unsigned long long *__restrict__ blocked,
                                           const unsigned long long *__restrict__ colcnt, int32_t *__restrict__ table,
                                           double *__restrict__ ps) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= P) return;
    const unsigned long long cc = colcnt[row];
    const long long touched = (long long)(cc & 0xFFFFFFFFull);
    const int bcol = (int)(cc >> 32);
    const double vmax = __longlong_as_double((long long)best[row]);
    int out = 0;
    double p = 1.0;
    if (kind == 0) {
        out = (touched > 0 && vmax > 0.0) ? bcol : 0;
    } else {
        const double rp = s.row_a ? (double)s.row_a[row] : 0.0, rc = s.row_b ? (double)s.row_b[row] : 0.0;
        const double base = (rc + 1.0) / (rp + 2.0);
        if (touched > 0 && (vmax > base || touched >= (long long)P)) {
            out = bcol; p = vmax;
        } else {
            const unsigned long long free_cols = ~blocked[row];
            if (free_cols == 0ull) { atomicAdd((unsigned long long *)s.stats + 1, 1ull); out = 64 < P ? 64 : P - 1; }
            else out = __ffsll((long long)free_cols) - 1;
            p = base;
        }
    }
    table[row] = out;
    if (ps) ps[row] = p;
}

// IPS / SNIPS of a deterministic target policy pi(a | lv) = [a == target[lv]] over the bandit rows with u < u_end
// (evaluate_agent.py:753-835; `for u in range(max(reco_log.u))` leaves the last user out): per-block partial sums of
// ee = c*pi/ps, ee^2, pi/ps and the row count, reduced in fixed order by ips_final.
__global__ void __launch_bounds__(256) ips_partial_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ u,
                                                          const int32_t *__restrict__ a, const float *__restrict__ c,
                                                          const double *__restrict__ ps, const int32_t *__restrict__ lv,
                                                          long long rows, long long u_end,
                                                          const int32_t *__restrict__ target, double *__restrict__ part) {
    __shared__ double sh[4][8];
    const long long per = (rows + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(rows, lo + per);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (long long r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        if (z[r] != 1 || (long long)u[r] >= u_end) continue;
        const double ratio = (a[r] == target[lv[r]] ? 1.0 : 0.0) / ps[r];
        const double ee = (double)c[r] * ratio;
        s0 += ee; s1 += ee * ee; s2 += ratio; s3 += 1.0;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) { sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; sh[3][w] = s3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[threadIdx.x][k];
        part[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}
__global__ void ips_final_kernel(const double *__restrict__ part, int nblocks, double *__restrict__ out) {
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int b = 0; b < nblocks; ++b) t += part[(size_t)b * 4 + threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// recall@k of a deterministic table policy (evaluate_agent.py:852-908): for every bandit row of the users < u_end that was
// not clicked and is directly followed by an organic row of the same user, is the product viewed next among the top k of
// the policy's action probabilities?  pi(. | lv) is one-hot at g = target[lv]; its "top k" is whatever numpy.argpartition
// returns for that vector (k - 1 zero-probability products beside g), tabulated per action by the caller: topk[g][0..k).
// Integer counters: out2 = {hits, eligible rows}.
__global__ void __launch_bounds__(256) recall_kernel(const int8_t *__restrict__ z, const int32_t *__restrict__ u,
                                                     const int32_t *__restrict__ v, const float *__restrict__ c,
                                                     const int32_t *__restrict__ lv, long long rows, long long u_end,
                                                     const int32_t *__restrict__ target, const int32_t *__restrict__ topk,
                                                     int k, unsigned long long *__restrict__ out2) {
    unsigned long long hits = 0ull, n = 0ull;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r + 1 < rows; r += (long long)gridDim.x * blockDim.x) {
        if (z[r] != 1 || (long long)u[r] >= u_end) continue;
        if (u[r + 1] != u[r] || z[r + 1] != 0 || c[r] != 0.f) continue;
        const int32_t *tk = topk + (size_t)target[lv[r]] * k;
        const int32_t next = v[r + 1];
        int hit = 0;
        for (int j = 0; j < k; ++j) hit |= (tk[j] == next) ? 1 : 0;
        hits += (unsigned long long)hit;
        n += 1ull;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        hits += __shfl_xor_sync(0xffffffffu, hits, d);
        n += __shfl_xor_sync(0xffffffffu, n, d);
    }
    if ((threadIdx.x & 31) == 0 && n) { atomicAdd(out2, hits); atomicAdd(out2 + 1, n); }
}

}  // namespace logs

using namespace logs;

cudaError_t launch_row_offsets(const int32_t *n_events, long long n, long long n_organic_users, long long *row_offset,
                               long long *tile_scratch, cudaStream_t stream) {
    // tile_scratch: [ceil(n / 4096) + 1] long long; the last element receives the total
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (n <= 0) return cudaMemsetAsync(row_offset, 0, sizeof(long long), stream);
    scan_tile_sums<<<ntiles, SCAN_THREADS, 0, stream>>>(n_events, n, n_organic_users, tile_scratch);
    scan_tile_offsets<<<1, SCAN_THREADS, 0, stream>>>(tile_scratch, ntiles, tile_scratch + ntiles);
    scan_apply<<<ntiles, SCAN_THREADS, 0, stream>>>(n_events, n, n_organic_users, tile_scratch, row_offset);
    return cudaGetLastError();
}

cudaError_t launch_order(const int4 *log, const unsigned long long *log_count, long long log_capacity,
                         const long long *row_offset, int4 *ordered, long long ordered_capacity, int num_sms,
                         cudaStream_t stream) {
    order_scatter<<<num_sms * 8, 256, 0, stream>>>(log, log_count, log_capacity, row_offset, ordered, ordered_capacity);
    return cudaGetLastError();
}

cudaError_t launch_columns(const int4 *ordered, const long long *row_offset, long long n_total,
                           long long n_organic_users, long long user_begin, const int32_t *last_view, int prev_last_view,
                           int policy_kind, int P, const double *ps_table, const int32_t *table, double epsilon,
                           int time_mode, uint64_t seed, double time_mu, double time_sigma,
                           const rg_log_columns &c, int train_mode,
                           int32_t *carry_out, int num_sms, cudaStream_t stream) {
    if (n_total <= 0) {
        if (carry_out) return cudaMemcpyAsync(carry_out, &prev_last_view, 4, cudaMemcpyHostToDevice, stream);
        return cudaSuccess;
    }
    Columns col{c.t, c.u, c.z, c.v, c.a, c.c, c.ps, c.lv, c.lvp, c.sess, c.flags};
    long long blocks = (n_total + 7) / 8;
    if (blocks > (long long)num_sms * 16) blocks = (long long)num_sms * 16;
    columns_kernel<<<(unsigned)blocks, 256, 0, stream>>>(ordered, row_offset, n_total, n_organic_users, user_begin,
                                                         last_view, prev_last_view, policy_kind, P, ps_table, table, epsilon,
                                                         time_mode, seed, time_mu, time_sigma, c.capacity,
                                                         train_mode, col);
    if (carry_out)
        carry_kernel<<<1, 32, 0, stream>>>(ordered, row_offset, last_view, n_total, n_organic_users, prev_last_view,
                                           c.capacity, train_mode, carry_out);
    return cudaGetLastError();
}

cudaError_t launch_count_stats(const rg_log_columns &c, long long rows, int P, int train_mode, int64_t *co,
                               int64_t *pulls, int64_t *clicks, int num_sms, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    if (co) organic_pairs_kernel<<<num_sms * 8, 256, 0, stream>>>(c.z, c.v, c.sess, c.flags, train_mode, rows, P,
                                                                  (unsigned long long *)co);
    if (pulls && clicks)
        bandit_stats_kernel<<<num_sms * 8, 256, 0, stream>>>(c.z, c.a, c.c, c.lvp, c.flags, train_mode, rows, P,
                                                             (unsigned long long *)pulls, (unsigned long long *)clicks);
    return cudaGetLastError();
}

cudaError_t launch_sparse_count_stats(const rg_log_columns &c, long long rows, int train_mode, const rg_sparse_counts *co,
                                      const rg_sparse_counts *bandit, int num_sms, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    if (co) sparse_organic_pairs_kernel<<<num_sms * 8, 256, 0, stream>>>(c.z, c.v, c.sess, c.flags, train_mode, rows, *co);
    if (bandit)
        sparse_bandit_stats_kernel<<<num_sms * 8, 256, 0, stream>>>(c.z, c.a, c.c, c.lvp, c.flags, train_mode, rows, *bandit);
    return cudaGetLastError();
}

cudaError_t launch_sparse_rehash(const rg_sparse_counts &from, const rg_sparse_counts &to, int num_sms, cudaStream_t stream) {
    sparse_rehash_kernel<<<num_sms * 8, 256, 0, stream>>>(from, to);
    return cudaGetLastError();
}

cudaError_t launch_sparse_policy(int kind, int P, const rg_sparse_counts &s, int32_t *table, double *ps,
                                 unsigned long long *scratch, int num_sms, cudaStream_t stream) {
    // scratch: [3 P] best value bits | blocked-column bitmap | (best column, touched cells) packed
    unsigned long long *best = scratch, *blocked = scratch + P, *colcnt = scratch + 2 * (size_t)P;
    cudaError_t ce = cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * 2 * (size_t)P, stream);
    if (ce != cudaSuccess) return ce;
    sparse_policy_init_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, colcnt);
    sparse_policy_best_kernel<<<num_sms * 8, 256, 0, stream>>>(kind, s, best, blocked, colcnt);
    sparse_policy_col_kernel<<<num_sms * 8, 256, 0, stream>>>(kind, s, best, colcnt);
    sparse_policy_final_kernel<<<(P + 255) / 256, 256, 0, stream>>>(kind, P, s, best, blocked, colcnt, table, ps);
    return cudaGetLastError();
}

cudaError_t launch_row_argmax(int kind, int P, const int64_t *A, const int64_t *B, int32_t *table, double *ps,
                              cudaStream_t stream) {
    const long long blocks = ((long long)P * 32 + 255) / 256;
    row_argmax_kernel<<<(unsigned)blocks, 256, 0, stream>>>(kind, P, A, B, table, ps);
    return cudaGetLastError();
}

cudaError_t launch_ips(const rg_log_columns &c, long long rows, long long u_end, const int32_t *target, double *part,
                       int nblocks, double *out, cudaStream_t stream) {
    ips_partial_kernel<<<nblocks, 256, 0, stream>>>(c.z, c.u, c.a, c.c, c.ps, c.lv, rows, u_end, target, part);
    ips_final_kernel<<<1, 32, 0, stream>>>(part, nblocks, out);
    return cudaGetLastError();
}

cudaError_t launch_recall(const rg_log_columns &c, long long rows, long long u_end, const int32_t *target,
                          const int32_t *topk, int k, unsigned long long *out2, int num_sms, cudaStream_t stream) {
    cudaError_t ce = cudaMemsetAsync(out2, 0, 16, stream);
    if (ce != cudaSuccess || rows < 2) return ce;
    long long blocks = (rows + 255) / 256;
    if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
    recall_kernel<<<(unsigned)blocks, 256, 0, stream>>>(c.z, c.u, c.v, c.c, c.lv, rows, u_end, target, topk, k, out2);
    return cudaGetLastError();
}

}  // namespace rg
