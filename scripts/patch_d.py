F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
a = s.index("        // ---- D: one thread per slot: finish the organic event, run the bandit events, recycle the slot.")
b = s.index("        __syncthreads();\n        omega_phase();\n        __syncthreads();\n    }\n", a)
new = r'''        // ---- D: advance the users.  D1 (one thread per slot) finishes the organic event: bookkeeping of the
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
            claim(slot);
        };
        const bool serial_bandit = env.change_omega_for_bandits != 0;
        if (tid < TU && s_live[tid]) {
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
            }
            const u32x4 w = draw(env.seed, user, t, SLOT_EVENT);
            int st = next_state(env, RG_ORGANIC, uniform53(w.z, w.w));
            t += 1;
            int pend = 0;
            if (st == RG_ORGANIC) pend = 1;
            else if (env.change_omega_for_bandits) thread_omega_update(om, K, env.sigma * 1.0, false, env.seed, user, t);
            while (serial_bandit && st == RG_BANDIT && !organic_only) {  // abstract.py:174-181, one event per turn
                int a;
                if (p.pol.kind == RG_POLICY_TABLE) a = __ldg(p.pol.table + last_view);
                else a = uniform_action(draw(env.seed, user, t, SLOT_POLICY).x, P);
                const u32x4 wb = draw(env.seed, user, t, SLOT_EVENT);
                const double uc = uniform53(wb.x, wb.y);
                int c = 0;
                if (uc >= 0.9706) c = click_from_uniform(ctr_of_action(env, om, a), uc);
                st = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                t += 1;
                if (st == RG_ORGANIC) pend = 1;  // the loop ends here: deferred
                else thread_omega_update(om, K, env.sigma * 1.0, false, env.seed, user, t);
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
        __syncthreads();
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
            __syncthreads();
            const int nban = misc[5];
            for (int b = warp; b < nban; b += NT / 32) {
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
                    int a;
                    if (p.pol.kind == RG_POLICY_TABLE) a = __ldg(p.pol.table + last_view);
                    else a = uniform_action(draw(env.seed, user, tt, SLOT_POLICY).x, P);
                    const u32x4 wb = draw(env.seed, user, tt, SLOT_EVENT);
                    const double uc = uniform53(wb.x, wb.y);
                    // ff() never exceeds sig(5*0.5-6) = 0.02931, so the click threshold (1-ctr)/((1-ctr)+ctr) is
                    // above 0.9706: below that the click is 0 whatever omega is (97% of the bandit events)
                    int c = 0;
                    if (uc >= 0.9706) c = click_from_uniform(ctr_of_action(env, om, a), uc);
                    const int nst = next_state(env, RG_BANDIT, uniform53(wb.z, wb.w));
                    const unsigned m = __ballot_sync(0xffffffffu, nst != RG_BANDIT || c != 0);
                    const int nev = m ? __ffs(m) : 32;  // events of this turn that really happen
                    const uint64_t rec = ((uint64_t)(uint32_t)a << 2) | ((uint64_t)c << 1) | 1u;
                    for (int j = 0; j < nev; ++j) h = hash_fold(h, __shfl_sync(0xffffffffu, rec, j));
                    if (lane < nev) {
                        if (p.out.action_hist) atomicAdd(p.out.action_hist + a, 1ull);
                        if (c && p.out.click_hist) atomicAdd(p.out.click_hist + a, 1ull);
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
'''
s = s[:a] + new + s[b:]
open(F, 'w').write(s)
print('ok')
