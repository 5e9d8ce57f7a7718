F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:70])
    s = s.replace(old, new)

# helpers: thread count from blockDim
sub("    for (int c = warp; c < nch; c += NT / 32) {", "    for (int c = warp; c < nch; c += (int)(blockDim.x >> 5)) {")
sub("    for (int idx = tid; idx < FB * KP; idx += NT) {  // stage the omega rows", "    for (int idx = tid; idx < FB * KP; idx += (int)blockDim.x) {  // stage the omega rows")
# kernel: thread count depends on the mode
sub("template <bool TC>\n__global__ void __launch_bounds__(NT, 2) rollout_cta_kernel(const Params p) {",
    "template <bool TC>\n__global__ void __launch_bounds__(TC ? NT + 32 : NT, 2) rollout_cta_kernel(const Params p) {\n    // TC: warp 8 is the TMA producer / MMA issuer of pass 1; it takes part in every other phase as a 9th worker\n    constexpr int NTH = TC ? NT + 32 : NT, NWARP = NTH / 32;")
sub("        for (int task = tid; task < npend * npairs; task += NT) {", "        for (int task = tid; task < npend * npairs; task += NTH) {")
sub("            for (int idx = tid; idx < K1 * TU; idx += NT) {", "            for (int idx = tid; idx < K1 * TU; idx += NTH) {")
sub("            for (int idx = tid; idx < TU * p.nrec_pad; idx += NT) tilesum[idx] = 0.f;", "            for (int idx = tid; idx < TU * p.nrec_pad; idx += NTH) tilesum[idx] = 0.f;")
sub("            for (int idx = tid; idx < TU * KT; idx += NT) {", "            for (int idx = tid; idx < TU * KT; idx += NTH) {")
sub("        for (int i = warp; i < nlive; i += NT / 32) {", "        for (int i = warp; i < nlive; i += NWARP) {")
sub("            for (int b = warp; b < nban; b += NT / 32) {", "            for (int b = warp; b < nban; b += NWARP) {")
sub("        rollout_cta_kernel<true><<<(unsigned)blocks, NT, (size_t)p.L.total, stream>>>(p);", "        rollout_cta_kernel<true><<<(unsigned)blocks, NT + 32, (size_t)p.L.total, stream>>>(p);")
# barriers: dfull + dempty
sub("    L.mbar = o; o += 8 * (kMaxStages + NBUF + 1);  // ring barriers, NBUF TMEM-full barriers, TMEM address word",
    "    L.mbar = o; o += 8 * (kMaxStages + 2 * NBUF + 1);  // ring, NBUF TMEM-full, NBUF TMEM-empty barriers, TMEM address")
sub("    uint32_t *tmem_slot = (uint32_t *)(mbar + kMaxStages + NBUF);", "    uint64_t *dempty = mbar + kMaxStages + NBUF;           // TC: accumulator tile drained by the 8 epilogue warps\n    uint32_t *tmem_slot = (uint32_t *)(mbar + kMaxStages + 2 * NBUF);")
sub("        if (TC) for (int b = 0; b < NBUF; ++b) mbar_init(&dfull[b], 1);", "        if (TC) for (int b = 0; b < NBUF; ++b) { mbar_init(&dfull[b], 1); mbar_init(&dempty[b], 8); }")
# pass 1 TC loop -> warp specialised
a = s.index("            auto issue_mma = [&](unsigned gs) {  // thread 0: 64-product tile gs -> TMEM buffer gs % NBUF")
b = s.index("            seq_use += (unsigned)ntile;\n            build_om32();")
new = '''            if (warp == 8) {
                // ---- producer / MMA issuer (one thread): B tile landed -> accumulator buffer drained -> three-term
                // MMA -> commit (accumulator ready).  The ring stage of tile gs-1 is refilled as soon as its MMA
                // has completed, i.e. independently of the epilogue.
                if (lane == 0) {
                    for (int tt = 0; tt < ntile; ++tt) {
                        const unsigned gs = seq_use + (unsigned)tt;
                        const unsigned st = gs % (unsigned)nstage, buf = gs % NBUF, use = gs / NBUF;
                        mbar_wait(&mbar[st], (gs / (unsigned)nstage) & 1u);
                        if (use > 0) mbar_wait(&dempty[buf], (use - 1u) & 1u);
                        tc::tc_fence_after();
                        tc::issue_tile_mmas(tmem_base + buf * 64u, a_saddr, ring_saddr + st * tile_bytes, KC);
                        tc::umma_commit(&dfull[buf]);
                        if (tt > 0 && tt - 1 + nstage < ntile) {
                            // MMA(gs-1) complete <=> dfull of its buffer completed its phase
                            const unsigned gp = gs - 1u;
                            mbar_wait(&dfull[gp % NBUF], (gp / NBUF) & 1u);
                            issue_load(gp + (unsigned)nstage);
                        }
                    }
                }
                __syncwarp();
            } else {
                // ---- epilogue warps: thread (lane quarter q, column half h) owns user 32q+lane and 32 columns
                const int q = warp & 3, h = warp >> 2;
                const int pos = q * 32 + lane;
                const bool act = pos < nlive;
                float part = 0.f;
                int rec = 0;
                for (int tt = 0; tt < ntile; ++tt) {
                    const unsigned gs = seq_use + (unsigned)tt;
                    const unsigned buf = gs % NBUF;
                    mbar_wait(&dfull[buf], (gs / NBUF) & 1u);
                    tc::tc_fence_after();
                    float v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64u + (uint32_t)(h * 32), v);
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dempty[buf]);  // this warp is done with the accumulator tile
                    if (act) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            s0 += ex2_approx(v[i]); s1 += ex2_approx(v[i + 1]);
                            s2 += ex2_approx(v[i + 2]); s3 += ex2_approx(v[i + 3]);
                        }
                        part += (s0 + s1) + (s2 + s3);
                    }
                    if ((tt + 1) % (2 * p.rec_sub) == 0 || tt == ntile - 1) {
                        if (act) atomicAdd(&tilesum[(size_t)pos * p.nrec_pad + rec], part);
                        part = 0.f;
                        ++rec;
                    }
                }
            }
            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();
'''
s = s[:a] + new + s[b:]
# mbar_arrive primitive
sub("__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {", '''__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {''')
open(F, 'w').write(s)
print('ok')
