F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new):
    global s
    assert s.count(old) == 1, (s.count(old), old[:70])
    s = s.replace(old, new)

sub("constexpr int kMaxStages = 4;", "constexpr int kMaxStages = 6;\nconstexpr int NBUF = 3;      // tensor mode: accumulator tiles (64 TMEM columns each) in flight")
sub('''    const int tc_tile_bytes = 2 * KC * tc::CHUNK_BYTES;
    L.mbar = o; o += 8 * (kMaxStages + 4);  // ring barriers, then [kMaxStages..+1] TMEM-full, [+2] TMEM address word''',
    '''    const int tc_tile_bytes = 2 * KC * tc::CHUNK_BYTES_B;  // one 64-product B tile (hi + lo)
    const int tc_a_bytes = 2 * KC * tc::CHUNK_BYTES;       // the 128-user A tile (hi + lo)
    L.mbar = o; o += 8 * (kMaxStages + NBUF + 1);  // ring barriers, NBUF TMEM-full barriers, TMEM address word''')
sub("        L.atile = o; o += tc_tile_bytes;", "        L.atile = o; o += tc_a_bytes;")
sub('''    uint32_t *tmem_slot = (uint32_t *)(mbar + kMaxStages + 2);''', '''    uint32_t *tmem_slot = (uint32_t *)(mbar + kMaxStages + NBUF);''')
sub('''    const unsigned tc_tile_bytes = 2u * (unsigned)KC * tc::CHUNK_BYTES;
''', '''    const unsigned tc_tile_bytes = 2u * (unsigned)KC * tc::CHUNK_BYTES_B;
''')
sub('''        if (TC) { mbar_init(&dfull[0], 1); mbar_init(&dfull[1], 1); }''', '''        if (TC) for (int b = 0; b < NBUF; ++b) mbar_init(&dfull[b], 1);''')
sub('''    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // two 128-column accumulator tiles''', '''    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // NBUF 64-column accumulator tiles''')
sub('''        const int stage = (int)(seq % (unsigned)nstage);
        const int sub = (int)(seq % (unsigned)nsub);
        mbar_expect_tx(&mbar[stage], tile_bytes);''', '''        const int stage = (int)(seq % (unsigned)nstage);
        const int sub = (int)(seq % (unsigned)ntile);
        mbar_expect_tx(&mbar[stage], tile_bytes);''')
sub('''    const int tc_prologue = nstage < nsub ? nstage : nsub;''', '''    const int tc_prologue = nstage < ntile ? nstage : ntile;''')
sub('''    const int nsub = env.nsub, nstage = p.nstage;
''', '''    const int nsub = env.nsub, nstage = p.nstage;
    const int ntile = TC ? 2 * nsub : nsub;  // tiles streamed per round: 64 products (TC) or 128 (CUDA cores)
''')
# the TC pass-1 loop
a = s.index("            auto issue_mma = [&](unsigned gs) {  // thread 0: tile gs -> TMEM buffer gs & 1")
b = s.index("            seq_use += (unsigned)nsub;\n            build_om32();     // pass 2 operand, in the (now idle) ring bytes")
new = '''            auto issue_mma = [&](unsigned gs) {  // thread 0: 64-product tile gs -> TMEM buffer gs % NBUF
                const unsigned st = gs % (unsigned)nstage;
                mbar_wait(&mbar[st], (gs / (unsigned)nstage) & 1u);
                tc::tc_fence_after();
                tc::issue_tile_mmas(tmem_base + (gs % NBUF) * 64u, a_saddr, ring_saddr + st * tile_bytes, KC);
                tc::umma_commit(&dfull[gs % NBUF]);
            };
            const int q = warp & 3, h = warp >> 2;
            const int pos = q * 32 + lane;
            const bool act = pos < nlive;
            float part = 0.f;
            int rec = 0;
            if (tid == 0)
                for (int j = 0; j < NBUF - 1 && j < ntile; ++j) issue_mma(seq_use + (unsigned)j);
            for (int tt = 0; tt < ntile; ++tt) {
                const unsigned gs = seq_use + (unsigned)tt;
                if (tid == 0 && tt + NBUF - 1 < ntile) issue_mma(gs + (unsigned)(NBUF - 1));
                mbar_wait(&dfull[gs % NBUF], (gs / NBUF) & 1u);
                tc::tc_fence_after();
                // MMA(gs) is complete: its ring stage is free for tile gs + nstage of this round
                if (tid == 0 && tt + nstage < ntile) issue_load(gs + (unsigned)nstage);
                __syncwarp();
                {
                    float v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (gs % NBUF) * 64u + (uint32_t)(h * 32), v);
                    if (act) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            s0 += ex2_approx(v[i]); s1 += ex2_approx(v[i + 1]);
                            s2 += ex2_approx(v[i + 2]); s3 += ex2_approx(v[i + 3]);
                        }
                        part += (s0 + s1) + (s2 + s3);
                    }
                }
                if ((tt + 1) % (2 * p.rec_sub) == 0 || tt == ntile - 1) {
                    if (act) atomicAdd(&tilesum[(size_t)pos * p.nrec_pad + rec], part);
                    part = 0.f;
                    ++rec;
                }
                tc::tc_fence_before();
                __syncthreads();  // TMEM buffer gs % NBUF may be overwritten by MMA(gs + NBUF)
            }
'''
s = s[:a] + new + s[b:]
sub("            seq_use += (unsigned)nsub;\n            build_om32();     // pass 2 operand, in the (now idle) ring bytes", "            seq_use += (unsigned)ntile;\n            build_om32();     // pass 2 operand, in the (now idle) ring bytes")
# launch: stage selection
sub('''    for (int nstage = 3; nstage >= 2 && !best_stage; --nstage) {
        p.nstage = nstage;
        make_layout(p);
        if (p.L.total <= 110 * 1024) best_stage = nstage;
    }
    if (!best_stage) {
        for (int nstage = 3; nstage >= 2 && !best_stage; --nstage) {''', '''    const int hi_stage = tensor ? kMaxStages : 3, lo_stage = tensor ? NBUF : 2;
    for (int nstage = hi_stage; nstage >= lo_stage && !best_stage; --nstage) {
        p.nstage = nstage;
        make_layout(p);
        if (p.L.total <= 110 * 1024) best_stage = nstage;
    }
    if (!best_stage) {
        for (int nstage = hi_stage; nstage >= lo_stage && !best_stage; --nstage) {''')
open(F, 'w').write(s)
print('ok')
