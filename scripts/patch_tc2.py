"""one-off source patch, part 2: the tensor-core pass 1 and its hooks."""
F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()


def sub(old, new):
    global s
    assert s.count(old) == 1, (s.count(old), old[:70])
    s = s.replace(old, new)


sub('''        // ---- B: pass 1, dense contraction + exponentials + per-record-tile partial sums
        {
            const bool warp_active = warp * 16 < nlive;''', '''        // ---- B: pass 1, dense contraction + exponentials + per-record-tile partial sums
        if constexpr (TC) {
            // Tensor-core form.  A = omega tile (tf32 hi/lo, UMMA K-major layout) written by the CTA, B = Gamma
            // tiles streamed by TMA, D = shifted log2-logits of 128 users x 128 products in TMEM (two buffers).
            // Thread 0 issues the three-term MMA of tile s+1 before the CTA runs the epilogue of tile s:
            // thread (lane quarter q, column half h) owns user 32q+lane and 64 columns; per logit the epilogue is
            // one ex2 and one add.
            for (int idx = tid; idx < TU * p.nrec_pad; idx += NT) tilesum[idx] = 0.f;
            const int KT = KC * 4;
            for (int idx = tid; idx < TU * KT; idx += NT) {
                const int row = idx / KT, k = idx - row * KT;
                float x = 0.f;
                if (row < nlive) {
                    if (k < K) x = (float)om64[(size_t)pos2slot[row] * KP + k];
                    else if (k == K) x = 1.f;
                    else if (k == K + 1) x = -s_shift[row];
                }
                const float hi = tc::tf32_hi(x);
                a_tile[tc::tile_index(0, KC, row, k)] = hi;
                a_tile[tc::tile_index(1, KC, row, k)] = tc::tf32_hi(x - hi);
            }
            tc::fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core
            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();

            const uint32_t a_saddr = tc::smem_addr(a_tile);
            const uint32_t ring_saddr = tc::smem_addr(gtile);
            auto issue_mma = [&](unsigned gs) {  // thread 0: tile gs -> TMEM buffer gs & 1
                const unsigned st = gs % (unsigned)nstage;
                mbar_wait(&mbar[st], (gs / (unsigned)nstage) & 1u);
                tc::tc_fence_after();
                tc::issue_tile_mmas(tmem_base + (gs & 1u) * 128u, a_saddr, ring_saddr + st * tile_bytes, KC);
                tc::umma_commit(&dfull[gs & 1u]);
            };
            const int q = warp & 3, h = warp >> 2;
            const int pos = q * 32 + lane;
            const bool act = pos < nlive;
            float part = 0.f;
            int rec = 0, sub_in_rec = 0;
            if (tid == 0) issue_mma(seq_use);
            for (int sub = 0; sub < nsub; ++sub) {
                const unsigned gs = seq_use + (unsigned)sub;
                if (tid == 0 && sub + 1 < nsub) issue_mma(gs + 1u);
                mbar_wait(&dfull[gs & 1u], (gs >> 1) & 1u);
                tc::tc_fence_after();
                // MMA(gs) is complete: its ring stage is free for tile gs + nstage of this round
                if (tid == 0 && sub + nstage < nsub) issue_load(gs + (unsigned)nstage);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (gs & 1u) * 128u + (uint32_t)(h * 64 + c * 32), v);
                    if (act) {
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) { s0 += ex2_approx(v[i]); s1 += ex2_approx(v[i + 1]); }
                        part += s0 + s1;
                    }
                }
                if (++sub_in_rec == p.rec_sub || sub == nsub - 1) {
                    if (act) atomicAdd(&tilesum[(size_t)pos * p.nrec_pad + rec], part);
                    part = 0.f;
                    ++rec;
                    sub_in_rec = 0;
                }
                tc::tc_fence_before();
                __syncthreads();  // TMEM buffer gs & 1 may be overwritten by MMA(gs + 2)
            }
            seq_use += (unsigned)nsub;
            build_om32();     // pass 2 operand, in the (now idle) ring bytes
            __syncthreads();
        } else {
            const bool warp_active = warp * 16 < nlive;''')

# prologue of the next round's ring (TC) right before phase D; C2 no longer needs the aliased bytes
sub('''        // ---- D: one thread per slot: finish the organic event, run the bandit events, recycle the slot.''',
    '''        if (TC && tid == 0) {
            // pass 2 and the re-sampling sweeps are done with the bytes they share with the ring:
            // start streaming the first tiles of the next round while the users are advanced
            for (int s2 = 0; s2 < tc_prologue; ++s2) issue_load(seq_use + (unsigned)s2);
        }

        // ---- D: one thread per slot: finish the organic event, run the bandit events, recycle the slot.''')

# drain + teardown
sub('''    for (int s = 0; s < nstage - 1; ++s) {
        const unsigned seq = seq_use + (unsigned)s;
        mbar_wait(&mbar[seq % (unsigned)nstage], (seq / (unsigned)nstage) & 1u);
    }
''', '''    for (int s = 0; s < (TC ? tc_prologue : nstage - 1); ++s) {
        const unsigned seq = seq_use + (unsigned)s;
        mbar_wait(&mbar[seq % (unsigned)nstage], (seq / (unsigned)nstage) & 1u);
    }
    if (TC) {
        tc::tc_fence_before();
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
    }
''')
open(F, 'w').write(s)
print("part 2 ok")
