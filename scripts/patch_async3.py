F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:70])
    s = s.replace(old, new)

# single queue
sub("    L.pu = o; o += 8 * TU;\n    L.rq_u = o; o += 8 * 2 * TU;\n    L.rq_lse = o; o += 8 * 2 * TU;", "    L.pu = o; o += 8 * TU;\n    L.rq_u = o; o += 8 * TU;\n    L.rq_lse = o; o += 8 * TU;")
sub("    L.wait = o; o += 4 * TU;\n    L.rq_slot = o; o += 4 * 2 * TU;", "    L.wait = o; o += 4 * TU;\n    L.rq_slot = o; o += 4 * TU;")
sub("    if (tid == 0) { misc[6] = 0; misc[7] = 0; }\n    const bool ASYNC = TC && p.async2 != 0;\n    int par = 0;  // queue consumed in this round: par; queue filled in this round: par ^ 1",
    "    if (tid == 0) misc[6] = 0;\n    const bool ASYNC = TC && p.async2 != 0;")
sub("        const int nq = misc[6 + par];", "        const int nq = misc[6];  // draws queued in the previous round")
sub('''                    const int *qs = rq_slot + par * TU;
                    const double *qu = rq_u + par * TU, *ql = rq_lse + par * TU;''', '''                    const int *qs = rq_slot;
                    const double *qu = rq_u, *ql = rq_lse;''')
sub('''                        const int e = (par ^ 1) * TU + atomicAdd(&misc[6 + (par ^ 1)], 1);''', '''                        const int e = atomicAdd(&misc[6], 1);''')
sub('''        if (tid == 0) misc[6 + par] = 0;  // the consumed queue is empty again
        par ^= 1;
        __syncthreads();''', '''        __syncthreads();''')
# the queue is consumed in pass 1 and refilled in pass 2b of the same round: reset its counter in between
sub('''            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();
            seq_use += (unsigned)ntile;''', '''            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();
            if (tid == 0) misc[6] = 0;  // the queue has been consumed; pass 2b refills it (after the next barrier)
            seq_use += (unsigned)ntile;''')
# per-position doubles of pass 2 live in the ring bytes in tensor mode
sub("    L.lse = o; o += 8 * TU;\n    L.pz = o; o += 8 * TU;\n    L.pu = o; o += 8 * TU;", "    L.lse = o; if (!p.tc) o += 8 * TU;\n    L.pz = o; if (!p.tc) o += 8 * TU;\n    L.pu = o; if (!p.tc) o += 8 * TU;")
sub("    L.ptau = o; o += 8 * TU;\n    L.pcum = o; o += 8 * TU;", "    L.ptau = o; if (!p.tc) o += 8 * TU;\n    L.pcum = o; if (!p.tc) o += 8 * TU;")
sub('''        const int alias = align_up(4 * K1 * TU, 16) + (p.async2 ? 0 : 8 * p.ncs * FB * 2);
        if (alias > ring) ring = align_up(alias, 128);
        L.om32 = o;
        if (!p.async2) L.csum = o + align_up(4 * K1 * TU, 16);
        o += ring;''', '''        const int om32_bytes = align_up(4 * K1 * TU, 16);
        const int alias = om32_bytes + 5 * 8 * TU + (p.async2 ? 0 : 8 * p.ncs * FB * 2);
        if (alias > ring) ring = align_up(alias, 128);
        L.om32 = o;
        L.lse = o + om32_bytes;
        L.pz = L.lse + 8 * TU; L.pu = L.pz + 8 * TU; L.ptau = L.pu + 8 * TU; L.pcum = L.ptau + 8 * TU;
        if (!p.async2) L.csum = L.pcum + 8 * TU;
        o += ring;''')
# launch: prefer the asynchronous layout when it still allows 2 CTAs per SM
a = s.index("    const int hi_stage = tensor ? kMaxStages : 3, lo_stage = tensor ? NBUF : 2;")
b = s.index("    p.nstage = best_stage;\n    make_layout(p);")
new = '''    const int hi_stage = tensor ? kMaxStages : 3, lo_stage = tensor ? NBUF : 2;
    const int two_ctas = 113 * 1024;  // (228 KB - 1 KB reserved per CTA) / 2
    int best_stage = 0;
    for (int pass = 0; pass < 2 && !best_stage; ++pass) {
        const int limit = pass == 0 ? two_ctas : 227 * 1024;
        for (int a2 = tensor ? 1 : 0; a2 >= 0 && !best_stage; --a2) {
            p.async2 = a2;
            for (int nstage = hi_stage; nstage >= lo_stage && !best_stage; --nstage) {
                p.nstage = nstage;
                make_layout(p);
                if (p.L.total <= limit) best_stage = nstage;
            }
        }
    }
    if (!best_stage && tensor) {
        // the UMMA operand tiles (2 x 2 x KC x 2 KB per stage + the A tile) do not fit for this K:
        // same kernel with the CUDA-core contraction (identical discrete results)
        return launch_rollout_fast(e, pol, user_begin, n_organic_users, n_total, out, stream, false);
    }
    if (!best_stage) return cudaErrorInvalidConfiguration;  // K/P too large for the CTA-batched kernel
'''
s = s[:a] + new + s[b:]
sub("    int ctas_per_sm = p.L.total <= 110 * 1024 ? 2 : 1;", "    int ctas_per_sm = p.L.total <= two_ctas ? 2 : 1;")
open(F, 'w').write(s)
print('ok')
