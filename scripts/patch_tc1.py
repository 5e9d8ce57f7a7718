"""one-off source patch: turn rollout_cta_kernel into a template<bool TC> (tensor-core pass 1)."""
F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()


def sub(old, new):
    global s
    assert s.count(old) == 1, (s.count(old), old[:70])
    s = s.replace(old, new)


sub('#include "rg_device.cuh"\n#include "rg_host.h"\n', '#include "rg_device.cuh"\n#include "rg_host.h"\n#include "rg_tensor.cuh"\n')
sub('''        amb_list, fb_slot, fb_pos, fb_unres, fb_u, fb_lse, lse, pend, pend_list, misc, om32, tilesum, gtile, total;''',
    '''        amb_list, fb_slot, fb_pos, fb_unres, fb_u, fb_lse, lse, pend, pend_list, misc, om32, tilesum, gtile, atile,
        total;''')
sub('''    int nstage, rec_sub, nrec, nrec_pad, kp, ncs;
    Layout L;''', '''    int nstage, rec_sub, nrec, nrec_pad, kp, ncs;
    int tc;                 // 1: tensor-core pass 1 (RG_MATH_TENSOR)
    double *om_scratch;     // tc: the binary64 omega rows live in global memory, [grid][TU][kp]
    Layout L;''')
sub('''    L.mbar = o; o += 8 * kMaxStages;
    L.om64 = o; o += 8 * TU * p.kp;
    L.csum = o; o += 8 * p.ncs * FB * 2;''', '''    const int KC = tc::kt_chunks(K);
    const int tc_tile_bytes = 2 * KC * tc::CHUNK_BYTES;
    L.mbar = o; o += 8 * (kMaxStages + 4);  // ring barriers, then [kMaxStages..+1] TMEM-full, [+2] TMEM address word
    L.om64 = o; if (!p.tc) o += 8 * TU * p.kp;
    L.csum = o; if (!p.tc) o += 8 * p.ncs * FB * 2;''')
sub('''    L.om32 = o; o += 4 * K1 * TU;
    L.tilesum = o; o += 4 * TU * p.nrec_pad;
    o = align_up(o, 128);
    L.gtile = o; o += 4 * p.nstage * K1 * SUB;
    L.total = o;''', '''    L.om32 = o; if (!p.tc) o += 4 * K1 * TU;
    L.tilesum = o; o += 4 * TU * p.nrec_pad;
    o = align_up(o, 128);
    L.gtile = o;
    if (!p.tc) {
        o += 4 * p.nstage * K1 * SUB;
        L.atile = o;
    } else {
        // the B ring is idle between pass 1 and the next round's first load: the fp32 omega tile of pass 2 and
        // the partial sums of the re-sampling sweeps live in the same bytes
        int ring = p.nstage * tc_tile_bytes;
        const int alias = align_up(4 * K1 * TU, 16) + 8 * p.ncs * FB * 2;
        if (alias > ring) ring = align_up(alias, 128);
        L.om32 = o;
        L.csum = o + align_up(4 * K1 * TU, 16);
        o += ring;
        L.atile = o; o += tc_tile_bytes;
    }
    L.total = o;''')
sub('__global__ void __launch_bounds__(NT, 2) rollout_cta_kernel(const Params p) {',
    'template <bool TC>\n__global__ void __launch_bounds__(NT, 2) rollout_cta_kernel(const Params p) {')
sub('    double *om64 = (double *)(smem + L.om64);\n',
    '    double *om64 = TC ? p.om_scratch + (size_t)blockIdx.x * TU * p.kp : (double *)(smem + L.om64);\n')
sub('''    float *gtile = (float *)(smem + L.gtile);

    const unsigned tile_bytes = (unsigned)(4 * K1 * SUB);
    const int nsub = env.nsub, nstage = p.nstage;
''', '''    float *gtile = (float *)(smem + L.gtile);
    float *a_tile = (float *)(smem + L.atile);
    uint64_t *dfull = mbar + kMaxStages;                   // TC: accumulator tile ready, one per TMEM buffer
    uint32_t *tmem_slot = (uint32_t *)(mbar + kMaxStages + 2);
    const int KC = env.KC;
    const unsigned tc_tile_bytes = 2u * (unsigned)KC * tc::CHUNK_BYTES;

    const unsigned tile_bytes = TC ? tc_tile_bytes : (unsigned)(4 * K1 * SUB);
    const int nsub = env.nsub, nstage = p.nstage;
''')
sub('''        for (int s = 0; s < nstage; ++s) mbar_init(&mbar[s], 1);
        fence_barrier_init();
        misc[1] = 0;
    }''', '''        for (int s = 0; s < nstage; ++s) mbar_init(&mbar[s], 1);
        if (TC) { mbar_init(&dfull[0], 1); mbar_init(&dfull[1], 1); }
        fence_barrier_init();
        misc[1] = 0;
    }
    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // two 128-column accumulator tiles''')
sub('''    auto issue_load = [&](unsigned seq) {  // thread 0 only: tile (seq % nsub) -> ring stage (seq % nstage)
        const int stage = (int)(seq % (unsigned)nstage);
        const int sub = (int)(seq % (unsigned)nsub);
        mbar_expect_tx(&mbar[stage], tile_bytes);
        tma_bulk_g2s(gtile + (size_t)stage * K1 * SUB, env.G32 + (size_t)sub * K1 * SUB, tile_bytes, &mbar[stage]);
    };
    if (tid == 0)
        for (int s = 0; s < nstage - 1; ++s) issue_load((unsigned)s);
''', '''    auto issue_load = [&](unsigned seq) {  // thread 0 only: tile (seq % nsub) -> ring stage (seq % nstage)
        const int stage = (int)(seq % (unsigned)nstage);
        const int sub = (int)(seq % (unsigned)nsub);
        mbar_expect_tx(&mbar[stage], tile_bytes);
        if (TC)
            tma_bulk_g2s((unsigned char *)gtile + (size_t)stage * tile_bytes, env.GTC + (size_t)sub * (tile_bytes / 4),
                         tile_bytes, &mbar[stage]);
        else
            tma_bulk_g2s(gtile + (size_t)stage * K1 * SUB, env.G32 + (size_t)sub * K1 * SUB, tile_bytes, &mbar[stage]);
    };
    // non-TC: the ring is always nstage-1 tiles ahead, across rounds.  TC: the ring bytes are reused between
    // pass 1 and the next round, so each round starts with its own prologue of min(nstage, nsub) loads.
    const int tc_prologue = nstage < nsub ? nstage : nsub;
    if (tid == 0) {
        if (TC) for (int s = 0; s < tc_prologue; ++s) issue_load((unsigned)s);
        else for (int s = 0; s < nstage - 1; ++s) issue_load((unsigned)s);
    }
    uint32_t tmem_base = 0;
    if (TC) {
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }
''')

# ---- phase A: shift/eps first, then the operand tile of the selected path
sub('''        // fp32 operand tile [K+1][TU] (row K = 1 multiplies the mu row of the Gamma tile), per-user shift and
        // error bound.  Dead positions get zeros.
        for (int idx = tid; idx < K1 * TU; idx += NT) {
            const int k = idx / TU, i = idx - k * TU;
            float v = 0.f;
            if (i < nlive) v = (k < K) ? (float)om64[(size_t)pos2slot[i] * KP + k] : 1.f;
            om32[idx] = v;
        }
        if (tid < TU) {''', '''        // fp32 operand tile [K+1][TU] (row K = 1 multiplies the mu row of the Gamma tile), per-user shift and
        // error bound.  Dead positions get zeros.  (TC: the fp32 tile is only needed by pass 2 and shares its
        // bytes with the B ring, so it is built after pass 1.)
        auto build_om32 = [&]() {
            for (int idx = tid; idx < K1 * TU; idx += NT) {
                const int k = idx / TU, i = idx - k * TU;
                float v = 0.f;
                if (i < nlive) v = (k < K) ? (float)om64[(size_t)pos2slot[i] * KP + k] : 1.f;
                om32[idx] = v;
            }
        };
        if (!TC) build_om32();
        if (tid < TU) {''')
open(F, 'w').write(s)
print("part 1 ok")
