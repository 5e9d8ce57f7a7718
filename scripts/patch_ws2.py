F = 'recogym_b200/csrc/rg_rollout_fast.cu'
s = open(F).read()
def sub(old, new, cnt=1):
    global s
    assert s.count(old) == cnt, (s.count(old), old[:70])
    s = s.replace(old, new)
sub("template <bool TC>\n__global__ void __launch_bounds__(TC ? NT + 32 : NT, 2) rollout_cta_kernel(const Params p) {\n    // TC: warp 8 is the TMA producer / MMA issuer of pass 1; it takes part in every other phase as a 9th worker\n    constexpr int NTH = TC ? NT + 32 : NT, NWARP = NTH / 32;",
    "template <bool TC>\n__global__ void __launch_bounds__(NT, 2) rollout_cta_kernel(const Params p) {\n    // TC pass 1: warps 0-3 run the epilogue (thread = TMEM lane = user), warp 4 is the TMA producer / MMA issuer,\n    // warps 5-7 idle (288 threads would cap the kernel at 96 registers)\n    constexpr int NTH = NT, NWARP = NTH / 32;")
sub("        rollout_cta_kernel<true><<<(unsigned)blocks, NT + 32, (size_t)p.L.total, stream>>>(p);", "        rollout_cta_kernel<true><<<(unsigned)blocks, NT, (size_t)p.L.total, stream>>>(p);")
sub("mbar_init(&dempty[b], 8); }", "mbar_init(&dempty[b], 4); }")
sub("    uint64_t *dempty = mbar + kMaxStages + NBUF;           // TC: accumulator tile drained by the 8 epilogue warps", "    uint64_t *dempty = mbar + kMaxStages + NBUF;           // TC: accumulator tile drained by the 4 epilogue warps")
sub("            if (warp == 8) {", "            if (warp == 4) {")
sub('''            } else {
                // ---- epilogue warps: thread (lane quarter q, column half h) owns user 32q+lane and 32 columns
                const int q = warp & 3, h = warp >> 2;
                const int pos = q * 32 + lane;''', '''            } else if (warp < 4) {
                // ---- epilogue warps: thread = TMEM lane = user 32*warp+lane, all TILE_N columns of every tile
                const int q = warp;
                const int pos = q * 32 + lane;''')
sub('''                    constexpr int HALF = tc::TILE_N / 2;  // columns of this thread''', '''                    constexpr int HALF = tc::TILE_N;  // columns of this thread (the whole tile row)''')
sub("                                          (uint32_t)(h * HALF + c), v);", "                                          (uint32_t)c, v);")
sub("                        if (act) atomicAdd(&tilesum[(size_t)pos * p.nrec_pad + rec], part);", "                        if (act) tilesum[(size_t)pos * p.nrec_pad + rec] = part;")
sub("            for (int idx = tid; idx < TU * p.nrec_pad; idx += NTH) tilesum[idx] = 0.f;\n", "")
open(F, 'w').write(s)
print('ok')
