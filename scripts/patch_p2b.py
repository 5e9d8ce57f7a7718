F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
a = s.index("                // one 128-product tile per step: 4 masses per lane (4 independent FFMA chains keep 4 x K1 L2")
b = s.index("                if (found < 0) {", a)
new = '''                // two 128-product tiles per step: 8 masses per lane (8 independent FFMA chains keep 8 x K1 L2 loads
                // in flight), the 8 group totals by butterfly sums, a scan only in the crossing group
                bool done = false;
                for (int base = p0; base < p1 && !done; base += 2 * SUB) {
                    const bool two = base + SUB < p1;
                    const float *g0 = env.G32 + (size_t)(base >> 7) * K1 * SUB + lane;
                    const float *g1 = two ? g0 + (size_t)K1 * SUB : g0;
                    float a[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) a[q] = 0.f;
#pragma unroll 3
                    for (int k = 0; k < K1; ++k) {
                        const float wv = om32[k * TU + i];
                        const float *gk0 = g0 + k * SUB, *gk1 = g1 + k * SUB;
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { v[q] = __ldg(gk0 + q * 32); v[4 + q] = __ldg(gk1 + q * 32); }
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[q] = fmaf(wv, v[q], a[q]);
                    }
                    double xs[8], tq[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        xs[q] = (q < 4 || two) ? (double)ex2_approx(a[q] - shf) : 0.0;
                        tq[q] = xs[q];
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) tq[q] += __shfl_xor_sync(0xffffffffu, tq[q], d);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (done) break;
                        if (cum + tq[q] > tau) {  // the crossing is in this group of 32 (or a rounding-level tie)
                            double x = xs[q];
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const double y = __shfl_up_sync(0xffffffffu, x, d);
                                if (lane >= d) x += y;
                            }
                            const double pre = cum + x;
                            const int pj = base + q * 32 + lane;
                            const unsigned m = __ballot_sync(0xffffffffu, (pre > tau) && (pj < P));
                            if (m) {
                                const int fl = __ffs(m) - 1;
                                found = base + q * 32 + fl;
                                upper = __shfl_sync(0xffffffffu, pre, fl);
                                const double below = __shfl_sync(0xffffffffu, pre, fl > 0 ? fl - 1 : 0);
                                lower = fl > 0 ? below : cum;
                            }
                            done = true;  // found, or inconsistent rounding (found < 0 -> ambiguous)
                        } else {
                            cum += tq[q];
                        }
                    }
                }
'''
s = s[:a] + new + s[b:]
open(F, 'w').write(s)
print('ok')
