F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:70])
    s = s.replace(old, new)

# sweep: stride by the number of warps of the calling group
sub('''                                                   double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                   int warp) {''', '''                                                   double *esum, int ncs, int nb, const double *shiftv, int lane,
                                                   int warp, int nwarps) {''')
sub("    for (int c = warp; c < nch; c += (int)(blockDim.x >> 5)) {", "    for (int c = warp; c < nch; c += nwarps) {")
# group-callable sampler
sub('''                                                int *s_view, int *unresolved, double *om_s, int tid,
                                                unsigned &ambig) {
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int lane = tid & 31, warp = tid >> 5;
    double *esum = csum + (size_t)FB * ncs;
    __shared__ double sh_shift[FB];''', '''                                                int *s_view, int *unresolved, double *om_s, int tid, int gn,
                                                int bar_id, unsigned &ambig) {
    // Called by a GROUP of gn threads (tid = index inside the group): the whole CTA (bar_id 0, __syncthreads)
    // or the three idle warps of the tensor-core pass 1 (named barrier bar_id).
    const int K = env.K, P = env.P, Ppad = env.Ppad;
    const int lane = tid & 31, warp = tid >> 5, nwarps = gn >> 5;
    auto gsync = [&]() {
        if (bar_id == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gn) : "memory");
    };
    double *esum = csum + (size_t)FB * ncs;
    __shared__ double sh_shift[FB];''')
sub("    for (int idx = tid; idx < FB * KP; idx += (int)blockDim.x) {  // stage the omega rows", "    for (int idx = tid; idx < FB * KP; idx += gn) {  // stage the omega rows")
sub('''        om_s[idx] = k < K ? om64[(size_t)slots[f < nb ? f : nb - 1] * KP + k] : 0.0;
    }
    __syncthreads();''', '''        om_s[idx] = k < K ? om64[(size_t)slots[f < nb ? f : nb - 1] * KP + k] : 0.0;
    }
    gsync();''')
for nbv, pp in ((1,8),(2,8),(3,4),(4,4),(5,2),(6,2),(7,2),(8,2)):
    sub("batch_partial_sums<%d, %d, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp); break;" % (nbv, pp),
        "batch_partial_sums<%d, %d, T2>(env, om_s, KP, csum, esum, ncs, nb, sh_shift, lane, warp, nwarps); break;" % (nbv, pp))
sub('''    __syncthreads();
    if (warp < nb) {
        const int f = warp;
        const double *omega = om_s + (size_t)f * KP;''', '''    gsync();
    for (int f = warp; f < nb; f += nwarps) {
        const double *omega = om_s + (size_t)f * KP;''')
sub('''                if (amb) ambig += 1;
            }
        }
    }
    __syncthreads();
}''', '''                if (amb) ambig += 1;
            }
        }
    }
    gsync();
}''')
# existing synchronous call sites
sub('''                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 fb_om, tid, ambig64);''', '''                        block_sample_views<true>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view, fb_unres,
                                                 fb_om, tid, NTH, 0, ambig64);''')
sub('''                        block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view,
                                                  fb_unres, fb_om, tid, ambig64);''', '''                        block_sample_views<false>(env, om64, KP, csum, p.ncs, fb_slot, fb_u, fb_lse, nb, s_view,
                                                  fb_unres, fb_om, tid, NTH, 0, ambig64);''')
open(F, 'w').write(s)
print('ok')
