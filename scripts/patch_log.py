def rw(path, pairs):
    s = open(path).read()
    for old, new in pairs:
        assert s.count(old) == 1, (path, s.count(old), old[:60])
        s = s.replace(old, new)
    open(path, 'w').write(s)

rw('include/recogym_b200.h', [
 ('''                            samples (diagnostic, ~1e-7 of the views), [8] draws tier 2 sent to tier 3, rest reserved */
} rg_rollout_out;''', '''                            samples (diagnostic, ~1e-7 of the views), [8] draws tier 2 sent to tier 3, rest reserved */
    /* Event log == the rows of the generate_logs DataFrame (abstract.py:256-327), unordered: one int32x4
     * record {u = user_id - user_begin, t, x = product viewed | action, flags} per row, flags bit 0: z
     * (0 organic, 1 bandit), bit 1: click, bit 2: phantom row after done (abstract.py:311-316).  Rows of one
     * user have distinct t, so sorting by (u, t) restores the reference order.  *log_count accumulates the
     * rows produced; rows beyond log_capacity are dropped (call again with a larger buffer). */
    int32_t *log;        /* [log_capacity][4] or NULL */
    int64_t log_capacity;
    int64_t *log_count;  /* [1] */
} rg_rollout_out;'''),
])
rw('recogym_b200/csrc/rg_device.cuh', [
 ('''    unsigned long long *view_hist, *action_hist, *click_hist, *totals;
};''', '''    unsigned long long *view_hist, *action_hist, *click_hist, *totals;
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
}'''),
])
rw('recogym_b200/csrc/rg_capi.cu', [
 ('''    d.totals = (unsigned long long *)o.totals;
    return d;''', '''    d.totals = (unsigned long long *)o.totals;
    d.log = (o.log && o.log_count && o.log_capacity > 0) ? (int4 *)o.log : nullptr;
    d.log_capacity = d.log ? o.log_capacity : 0;
    d.log_count = (unsigned long long *)o.log_count;
    return d;'''),
 ('''    const size_t o_tab = total;''', '''    const size_t o_log = add(out->log && out->log_count && out->log_capacity > 0 ? out->log : nullptr,
                             16 * (size_t)(out->log_capacity > 0 ? out->log_capacity : 0), false);
    const size_t o_lc = add(out->log && out->log_count ? out->log_count : nullptr, 8, true);
    const size_t o_tab = total;'''),
 ('''    if (out->totals) dv.totals = (int64_t *)(arena + o_t);''', '''    if (out->totals) dv.totals = (int64_t *)(arena + o_t);
    if (out->log && out->log_count && out->log_capacity > 0) {
        dv.log = (int32_t *)(arena + o_log);
        dv.log_capacity = out->log_capacity;
        dv.log_count = (int64_t *)(arena + o_lc);
    }'''),
])
# exact kernel
rw('recogym_b200/csrc/rg_kernels.cu', [
 ('''                if (out.view_hist && lane == 0) atomicAdd(out.view_hist + v, 1ull);
            } else if (state == RG_BANDIT && !organic_only) {
                const int a = policy_action(env, pol, user, t, last_view);
                const int c = bandit_event(env, wc, user, a, state, t, lane);''', '''                if (out.view_hist && lane == 0) atomicAdd(out.view_hist + v, 1ull);
                if (out.log && lane == 0) log_write(out, log_reserve(out, 1), i, t - 1u, v, 0);
            } else if (state == RG_BANDIT && !organic_only) {
                const int a = policy_action(env, pol, user, t, last_view);
                const int c = bandit_event(env, wc, user, a, state, t, lane);
                if (out.log && lane == 0) log_write(out, log_reserve(out, 1), i, t - 1u, a, 1 | (c << 1));'''),
 ('''        h = hash_fold(h, ((uint64_t)t << 2) | 3u);
        if (lane == 0) {
            if (out.n_events) out.n_events[i] = (int32_t)t;''', '''        h = hash_fold(h, ((uint64_t)t << 2) | 3u);
        if (out.log && lane == 0 && !organic_only)  // phantom action after done (abstract.py:311-316), reward 0
            log_write(out, log_reserve(out, 1), i, t, policy_action(env, pol, user, t, last_view), 1 | 4);
        if (lane == 0) {
            if (out.n_events) out.n_events[i] = (int32_t)t;'''),
])
# fast kernel
rw('recogym_b200/csrc/rg_rollout_fast.cu', [
 ('''            tot_ph += idx < p.n_organic_users ? 0 : 1;
            claim(slot);
        };''', '''            tot_ph += idx < p.n_organic_users ? 0 : 1;
            if (p.out.log && idx >= p.n_organic_users) {  // phantom action after done (abstract.py:311-316)
                int a;
                if (p.pol.kind == RG_POLICY_TABLE) a = __ldg(p.pol.table + last_view);
                else a = uniform_action(draw(env.seed, user, t, SLOT_POLICY).x, P);
                log_write(p.out, log_reserve(p.out, 1), idx, t, a, 1 | 4);
            }
            claim(slot);
        };'''),
 ('''                if (p.out.view_hist) atomicAdd(p.out.view_hist + v, 1ull);
            }
            const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
            int st = next_state(env, RG_ORGANIC, uniform53(w.z, w.w));''', '''                if (p.out.view_hist) atomicAdd(p.out.view_hist + v, 1ull);
                if (p.out.log) log_write(p.out, log_reserve(p.out, 1), (long long)user - p.user_begin, t, v, 0);
            }
            const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
            int st = next_state(env, RG_ORGANIC, uniform53(w.z, w.w));'''),
 ('''                st = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                t += 1;
                if (st == RG_ORGANIC) pend = 1;  // the loop ends here: deferred''', '''                st = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                if (p.out.log) log_write(p.out, log_reserve(p.out, 1), (long long)user - p.user_begin, t, a, 1 | (c << 1));
                t += 1;
                if (st == RG_ORGANIC) pend = 1;  // the loop ends here: deferred'''),
 ('''                    if (lane < nev) {
                        if (p.out.action_hist) atomicAdd(p.out.action_hist + a, 1ull);''', '''                    if (p.out.log) {  // one reservation per turn, lane j writes the row of its event
                        unsigned long long row0 = 0;
                        if (lane == 0) row0 = log_reserve(p.out, nev);
                        row0 = __shfl_sync(0xffffffffu, row0, 0);
                        if (lane < nev)
                            log_write(p.out, row0 + (unsigned long long)lane, (long long)user - p.user_begin, tt, a, 1 | (c << 1));
                    }
                    if (lane < nev) {
                        if (p.out.action_hist) atomicAdd(p.out.action_hist + a, 1ull);'''),
])
rw('recogym_b200/_lib.py', [
 ('''class RolloutOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "hash", "omega",
        "view_hist", "action_hist", "click_hist", "totals")]''', '''class RolloutOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_events", "n_organic", "n_bandit", "n_clicks", "last_view", "hash", "omega",
        "view_hist", "action_hist", "click_hist", "totals")] + [
        ("log", C.c_void_p), ("log_capacity", C.c_int64), ("log_count", C.c_void_p)]'''),
])
print('ok')
