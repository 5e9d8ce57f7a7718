F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new):
    global s
    assert s.count(old) == 1, (s.count(old), old[:70])
    s = s.replace(old, new)
a = s.index("        // ---- C: pass 2, one warp per live user: total mass, locate, recompute one record tile, margins")
b = s.index("            int found = -1;\n            bool amb = false;\n            double lower = 0.0, upper = 0.0;\n            if (rsel < 0) {")
new = '''        // ---- C: pass 2a, one thread per live user: view uniform, total mass, tau = u*Z, the record tile that
        // contains tau and the mass before it (sequential binary64 sums of the fp32 tile sums)
        if (tid < nlive) {
            const int i = tid, slot = pos2slot[i];
            const u32x4 w = draw(env.seed, s_user[slot], s_t[slot], SLOT_EVENT);
            const double u = uniform53(w.x, w.y);
            const float *ts = tilesum + (size_t)i * p.nrec_pad;
            double Z = 0.0;
            for (int r = 0; r < p.nrec; ++r) Z += (double)ts[r];
            const double tau = u * Z;
            double cum = 0.0;
            int rsel = -1;
            for (int r = 0; r < p.nrec; ++r) {
                const double nx = cum + (double)ts[r];
                if (nx > tau) { rsel = r; break; }
                cum = nx;
            }
            s_Z[i] = Z; s_tau[i] = tau; s_cum[i] = cum; s_rsel[i] = rsel;
        }
        __syncthreads();

        // ---- C: pass 2b, one warp per live user: recompute the located record tile, index and margins
        for (int i = warp; i < nlive; i += NT / 32) {
            const int slot = pos2slot[i];
            const double Z = s_Z[i], tau = s_tau[i];
            double cum = s_cum[i];
            const int rsel = s_rsel[i];
'''
s = s[:a] + new + s[b:]
sub("    L.fb_lse = o; o += 8 * FB;\n    L.lse = o; o += 8 * TU;", "    L.fb_lse = o; o += 8 * FB;\n    L.lse = o; o += 8 * TU;\n    L.pz = o; o += 8 * TU;\n    L.ptau = o; o += 8 * TU;\n    L.pcum = o; o += 8 * TU;")
sub("    L.fb_unres = o; o += 4 * FB;", "    L.fb_unres = o; o += 4 * FB;\n    L.prsel = o; o += 4 * TU;")
sub("fb_u, fb_lse, lse, pend,", "fb_u, fb_lse, lse, pz, ptau, pcum, prsel, pend,")
sub("    int *fb_pos = (int *)(smem + L.fb_pos);", "    double *s_Z = (double *)(smem + L.pz);\n    double *s_tau = (double *)(smem + L.ptau);\n    double *s_cum = (double *)(smem + L.pcum);\n    int *s_rsel = (int *)(smem + L.prsel);\n    int *fb_pos = (int *)(smem + L.fb_pos);")
open(F, 'w').write(s)
print('ok')
