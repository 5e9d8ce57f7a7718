F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
old = '''#pragma unroll 2
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
'''
new = '''            // memory-level parallelism: KB x PP binary64 loads of Gamma are issued before any of them is used
            // (the sweep streams 8*K*P bytes from L2 and is bound by the bytes a CTA keeps in flight)
            constexpr int KB = PP >= 8 ? 2 : PP >= 4 ? 5 : 10;
            for (int k0 = 0; k0 < K; k0 += KB) {
                double gk[KB][PP];
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) {
                    const int k = k0 + kk < K ? k0 + kk : K - 1;
#pragma unroll
                    for (int q = 0; q < PP; ++q) gk[kk][q] = __ldg(gp[q] + (size_t)k * Ppad);
                }
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) {
                    if (k0 + kk < K) {
#pragma unroll
                        for (int f = 0; f < NB; ++f) {
                            const double w = om_s[f * KP + k0 + kk];
#pragma unroll
                            for (int q = 0; q < PP; ++q) acc[q][f] = fma(gk[kk][q], w, acc[q][f]);
                        }
                    }
                }
            }
'''
assert s.count(old) == 1
s = s.replace(old, new)
open(F, 'w').write(s)
print('ok')
