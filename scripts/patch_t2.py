F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new):
    global s
    assert s.count(old) == 1, (s.count(old), old[:70])
    s = s.replace(old, new)

a = s.index("template <int NB, bool T2>\n__device__ __forceinline__ void batch_partial_sums(")
b = s.index("// slots/us/shiftv: shared arrays [FB]")
new = r'''template <int NB, int PP, bool T2>
__device__ __forceinline__ void batch_partial_sums(const EnvDev &env, const double *om_s, int KP, double *csum,
                                                   double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                   int warp) {
    // om_s: the omega rows of the batch staged in shared memory, [FB][KP].  Each lane carries PP products x NB
    // users = PP*NB independent fma chains and PP loads in flight per k (the sweep is latency bound otherwise).
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    double sv[NB];
#pragma unroll
    for (int f = 0; f < NB; ++f) sv[f] = shiftv[f < nb ? f : nb - 1];
    const int nch = (Ppad + CHUNK - 1) / CHUNK;
    for (int c = warp; c < nch; c += NT / 32) {
        double part[NB], perr[NB];
#pragma unroll
        for (int f = 0; f < NB; ++f) { part[f] = 0.0; perr[f] = 0.0; }
        for (int g0 = 0; g0 < CHUNK / 32; g0 += PP) {
            double acc[PP][NB];
            const double *gp[PP];
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int j = c * CHUNK + (g0 + q) * 32 + lane;
                gp[q] = env.GammaT + (j < Ppad ? j : Ppad - 1);
#pragma unroll
                for (int f = 0; f < NB; ++f) acc[q][f] = 0.0;
            }
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                double gk[PP];
#pragma unroll
                for (int q = 0; q < PP; ++q) gk[q] = __ldg(gp[q] + (size_t)k * Ppad);
#pragma unroll
                for (int f = 0; f < NB; ++f) {
                    const double w = om_s[f * KP + k];
#pragma unroll
                    for (int q = 0; q < PP; ++q) acc[q][f] = fma(gk[q], w, acc[q][f]);
                }
            }
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const int j = c * CHUNK + (g0 + q) * 32 + lane;
                if (j < P) {
                    const double mu = __ldg(env.mu_o + j);
#pragma unroll
                    for (int f = 0; f < NB; ++f) {
                        if (T2) {
                            const double x = (acc[q][f] + mu) * 1.4426950408889634074 - sv[f];
                            const double m = (double)ex2_approx((float)x);
                            part[f] += m;
                            perr[f] = fma(m, fma(fabs(x), 4.14e-8, 4.8e-7), perr[f]);
                        } else {
                            part[f] += exp64(acc[q][f] + mu - sv[f]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NB; ++f) {
            const double tot = warp_sum(part[f]);
            if (lane == 0 && f < nb) csum[(size_t)f * ncs + c] = tot;
            if (T2) {
                const double te = warp_sum(perr[f]);
                if (lane == 0 && f < nb) esum[(size_t)f * ncs + c] = te;
            }
        }
    }
}

'''
s = s[:a] + new + s[b:]

# block_sample_views: stage omega rows, new dispatch, faster locate
sub('''                                                const int *slots, const double *us, const double *lse, int nb,
                                                int *s_view, int *unresolved, int tid, unsigned &ambig) {''',
    '''                                                const int *slots, const double *us, const double *lse, int nb,
                                                int *s_view, int *unresolved, double *om_s, int tid,
                                                unsigned &ambig) {''')
sub('''        sh_shift[tid] = v;
    }
    __syncthreads();
    switch (nb) {
        case 1: batch_partial_sums<1, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 2: batch_partial_sums<2, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 3: batch_partial_sums<3, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 4: batch_partial_sums<4, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 5: batch_partial_sums<5, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 6: batch_partial_sums<6, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        case 7: batch_partial_sums<7, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
        default: batch_partial_sums<8, T2>(env, om64, KP, csum, esum, ncs, slots, nb, sh_shift, lane, warp); break;
    }
    __syncthreads();''', '''        sh_shift[tid] = v;
    }
    for (int idx = tid; idx < FB * KP; idx += NT) {  // stage the omega rows of the batch (rows >= nb repeat the last)
        const int f = idx / KP, k = idx - f * KP;
        om_s[idx] = k < K ? om64[(size_t)slots[f < nb ? f : nb - 1] * KP + k] : 0.0;
    }
    __syncthreads();
    switch (nb) {
        case 1: batch_partial_sums<1, 8, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 2: batch_partial_sums<2, 8, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 3: batch_partial_sums<3, 4, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 4: batch_partial_sums<4, 4, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 5: batch_partial_sums<5, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 6: batch_partial_sums<6, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        case 7: batch_partial_sums<7, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
        default: batch_partial_sums<8, 2, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;
    }
    __syncthreads();''')
sub('''        const double *omega = om64 + (size_t)slots[f] * KP;
        const double *cs = csum + (size_t)f * ncs;''', '''        const double *omega = om_s + (size_t)f * KP;
        const double *cs = csum + (size_t)f * ncs;''')
# locate walk: all 8 groups of the chunk at once
a = s.index("        int found = -1;\n        double lower = 0.0, upper = 0.0;\n        const int end = min((csel + 1) * CHUNK, Ppad);")
b = s.index("        bool amb = false;\n        if (found < 0) {\n            found = min(end, P) - 1;")
new = '''        int found = -1;
        double lower = 0.0, upper = 0.0;
        const int end = min((csel + 1) * CHUNK, Ppad);
        {   // the whole chunk at once: 8 masses per lane (8 independent chains), 8 butterfly totals, one scan
            constexpr int G = CHUNK / 32;
            double acc[G], xs[G], tq[G];
            const double *gp[G];
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int j = csel * CHUNK + q * 32 + lane;
                gp[q] = env.GammaT + (j < Ppad ? j : Ppad - 1);
                acc[q] = 0.0;
            }
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                const double w = omega[k];
#pragma unroll
                for (int q = 0; q < G; ++q) acc[q] = fma(__ldg(gp[q] + (size_t)k * Ppad), w, acc[q]);
            }
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int j = csel * CHUNK + q * 32 + lane;
                double m = 0.0;
                if (j < P) {
                    const double l = acc[q] + __ldg(env.mu_o + j);
                    m = T2 ? (double)ex2_approx((float)(l * 1.4426950408889634074 - sf)) : exp64(l - sf);
                }
                xs[q] = m;
                tq[q] = m;
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
                for (int q = 0; q < G; ++q) tq[q] += __shfl_xor_sync(0xffffffffu, tq[q], d);
            }
            bool done = false;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                if (done) break;
                if (cum + tq[q] > tau || q == G - 1) {
                    double x = xs[q];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double y = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= d) x += y;
                    }
                    const double pre = cum + x;
                    const int base = csel * CHUNK + q * 32;
                    const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (base + lane < P));
                    if (m) {
                        const int fl = __ffs(m) - 1;
                        found = base + fl;
                        upper = __shfl_sync(0xffffffffu, pre, fl);
                        const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
                        lower = fl > 0 ? below : cum;
                    }
                    done = true;
                } else {
                    cum += tq[q];
                }
            }
        }
'''
s = s[:a] + new + s[b:]
# the old mass lambda is now unused in block_sample_views: remove it
a = s.index("        auto mass = [&](int j) -> double {\n            if (j >= P) return 0.0;\n            const double *g = env.GammaT + j;\n            double acc = 0.0;\n            for (int k = 0; k < K; ++k) acc = fma(__ldg(g + (size_t)k * Ppad), omega[k], acc);\n            const double l = acc + __ldg(env.mu_o + j);")
b = s.index("        const int nch = (Ppad + CHUNK - 1) / CHUNK;\n        double zl = 0.0, el = 0.0;")
s = s[:a] + s[b:]
# staging buffer in the layout + call sites
sub("    L.fb_lse = o; o += 8 * FB;", "    L.fb_lse = o; o += 8 * FB;\n    L.fb_om = o; o += 8 * FB * p.kp;")
sub("fb_u, fb_lse, lse,", "fb_u, fb_lse, fb_om, lse,")
sub("    double *fb_lse = (double *)(smem + L.fb_lse);", "    double *fb_lse = (double *)(smem + L.fb_lse);\n    double *fb_om = (double *)(smem + L.fb_om);")
sub('''                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 tid, ambig64);''', '''                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 fb_om, tid, ambig64);''')
sub('''                        block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view,
                                                  fb_unres, tid, ambig64);''', '''                        block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view,
                                                  fb_unres, fb_om, tid, ambig64);''')
open(F, 'w').write(s)
print('ok')
