F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:70])
    s = s.replace(old, new)

# ---- params / layout
sub("    int tc;                 // 1: tensor-core pass 1 (RG_MATH_TENSOR)", "    int tc;                 // 1: tensor-core pass 1 (RG_MATH_TENSOR)\n    int async2;             // tc: tier 2/3 re-sampling runs on warps 5-7 underneath the next round's pass 1")
sub('''        L.om32 = o;
        L.csum = o + align_up(4 * K1 * TU, 16);
        o += ring;''', '''        L.om32 = o;
        if (!p.async2) L.csum = o + align_up(4 * K1 * TU, 16);
        o += ring;''')
sub("    L.csum = o; if (!p.tc) o += 8 * p.ncs * FB * 2;", "    L.csum = o; if (!p.tc || p.async2) o += 8 * p.ncs * FB * 2;")
sub('''        const int alias = align_up(4 * K1 * TU, 16) + 8 * p.ncs * FB * 2;''', '''        const int alias = align_up(4 * K1 * TU, 16) + (p.async2 ? 0 : 8 * p.ncs * FB * 2);''')
sub("    L.pz = o; o += 8 * TU;", "    L.pz = o; o += 8 * TU;\n    L.pu = o; o += 8 * TU;\n    L.rq_u = o; o += 8 * 2 * TU;\n    L.rq_lse = o; o += 8 * 2 * TU;")
sub("    L.prsel = o; o += 4 * TU;", "    L.prsel = o; o += 4 * TU;\n    L.wait = o; o += 4 * TU;\n    L.rq_slot = o; o += 4 * 2 * TU;")
sub("lse, pz, ptau, pcum, prsel, pend,", "lse, pz, pu, rq_u, rq_lse, ptau, pcum, prsel, wait, rq_slot, pend,")
sub("    double *s_Z = (double *)(smem + L.pz);", "    double *s_Z = (double *)(smem + L.pz);\n    double *s_u = (double *)(smem + L.pu);\n    double *rq_u = (double *)(smem + L.rq_u);       // re-sampling queues [2][TU] (double-buffered by round parity)\n    double *rq_lse = (double *)(smem + L.rq_lse);\n    int *rq_slot = (int *)(smem + L.rq_slot);\n    int *s_wait = (int *)(smem + L.wait);           // 0 none, 2 queued this round, 1 being re-sampled this round")

# ---- claim clears wait; set-up clears arrays
sub("        s_pend[slot] = 2;  // omega = sigma0 * z(user, 0)", "        s_wait[slot] = 0;\n        s_pend[slot] = 2;  // omega = sigma0 * z(user, 0)")
sub("    if (tid < TU) { s_pend[tid] = 0; claim(tid); }", "    if (tid < TU) { s_pend[tid] = 0; s_wait[tid] = 0; claim(tid); }\n    if (tid == 0) { misc[6] = 0; misc[7] = 0; }\n    const bool ASYNC = TC && p.async2 != 0;\n    int par = 0;  // queue consumed in this round: par; queue filled in this round: par ^ 1")

# ---- phase A: waiting slots are not positions
sub('''                const bool lv = s_live[slot] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, lv);
                if (lv) pos2slot[base + __popc(m & ((1u << lane) - 1u))] = slot;''', '''                if (s_wait[slot] == 2) s_wait[slot] = 1;  // queued last round: re-sampled during this pass 1
                const bool lv = s_live[slot] != 0 && s_wait[slot] == 0;
                const unsigned m = __ballot_sync(0xffffffffu, lv);
                if (lv) pos2slot[base + __popc(m & ((1u << lane) - 1u))] = slot;''')
sub('''        const int nlive = misc[0];
        if (nlive == 0) break;
''', '''        const int nlive = misc[0];
        const int nq = misc[6 + par];
        if (nlive == 0 && nq == 0) break;
''')

# ---- pass 1: warps 5-7 re-sample the queue
sub('''            } else if (warp < 4) {
                // ---- epilogue warps: thread = TMEM lane = user 32*warp+lane, all TILE_N columns of every tile''', '''            } else if (warp >= 5) {
                // ---- re-sampling warps: the draws queued in the previous round, tier 2 then tier 3, on the
                // (otherwise idle) FP64 pipe while warps 0-4 stream the contraction of this round
                if (ASYNC) {
                    const int gtid = tid - 5 * 32, gn = 3 * 32;
                    const int *qs = rq_slot + par * TU;
                    const double *qu = rq_u + par * TU, *ql = rq_lse + par * TU;
                    for (int b0 = 0; b0 < nq; b0 += FB) {
                        int nb = min(FB, nq - b0);
                        if (gtid < nb) {
                            fb_slot[gtid] = qs[b0 + gtid];
                            fb_u[gtid] = qu[b0 + gtid];
                            fb_lse[gtid] = ql[b0 + gtid];
                            fb_unres[gtid] = 0;
                        }
                        asm volatile("bar.sync 1, 96;" ::: "memory");
                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 fb_om, gtid, gn, 1, ambig64);
                        if (gtid == 0) {  // compact the draws tier 2 left open (serial: at most FB entries)
                            int n3 = 0;
                            for (int f = 0; f < nb; ++f)
                                if (fb_unres[f]) { fb_slot[n3] = fb_slot[f]; fb_u[n3] = fb_u[f]; ++n3; }
                            misc[8] = n3;
                            n_fallback += (unsigned long long)nb;
                            n_tier3 += (unsigned long long)n3;
                        }
                        asm volatile("bar.sync 1, 96;" ::: "memory");
                        const int n3 = misc[8];
                        if (n3 > 0)
                            block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, n3, s_view,
                                                      fb_unres, fb_om, gtid, gn, 1, ambig64);
                        asm volatile("bar.sync 1, 96;" ::: "memory");
                    }
                    for (int e = gtid; e < nq; e += gn) s_wait[qs[e]] = 0;  // their views are in s_view: advance in D
                }
            } else if (warp < 4) {
                // ---- epilogue warps: thread = TMEM lane = user 32*warp+lane, all TILE_N columns of every tile''')

# ---- pass 2a stores u; pass 2b queues ambiguous draws
sub("            s_Z[i] = Z; s_tau[i] = tau; s_cum[i] = cum; s_rsel[i] = rsel;", "            s_Z[i] = Z; s_tau[i] = tau; s_cum[i] = cum; s_rsel[i] = rsel; s_u[i] = u;")
sub('''                if (amb) {
                    amb_list[atomicAdd(&misc[1], 1)] = i;
                    s_lse[i] = log64(Z) * 1.4426950408889634074 + (double)s_shift[i];
                }''', '''                if (amb) {
                    const double lse = log64(Z) * 1.4426950408889634074 + (double)s_shift[i];
                    if (ASYNC) {  // re-sampled underneath the next round's pass 1; the slot sits that round out
                        const int e = (par ^ 1) * TU + atomicAdd(&misc[6 + (par ^ 1)], 1);
                        rq_slot[e] = slot; rq_u[e] = s_u[i]; rq_lse[e] = lse;
                        s_wait[slot] = 2;
                    } else {
                        amb_list[atomicAdd(&misc[1], 1)] = i;
                        s_lse[i] = lse;
                    }
                }''')
# ---- D1 skips slots that wait; end of round flips the queues
sub("        if (tid < TU && s_live[tid]) {\n            const int slot = tid;\n            double *om = om64 + (size_t)slot * KP;\n            const uint64_t user = s_user[slot];\n            uint64_t h = s_hash[slot];",
    "        if (tid < TU && s_live[tid] && s_wait[tid] == 0) {\n            const int slot = tid;\n            double *om = om64 + (size_t)slot * KP;\n            const uint64_t user = s_user[slot];\n            uint64_t h = s_hash[slot];")
sub('''        __syncthreads();
        omega_phase();
        __syncthreads();
    }
''', '''        __syncthreads();
        omega_phase();
        if (tid == 0) misc[6 + par] = 0;  // the consumed queue is empty again
        par ^= 1;
        __syncthreads();
    }
''')
# ---- launch: async when the dedicated partial-sum arrays fit
sub('''    p.tc = tensor ? 1 : 0;
    p.om_scratch = nullptr;''', '''    p.tc = tensor ? 1 : 0;
    p.async2 = tensor ? 1 : 0;
    p.om_scratch = nullptr;''')
open(F, 'w').write(s)
print('ok')
