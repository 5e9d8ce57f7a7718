def rw(path, pairs):
    s = open(path).read()
    for old, new, cnt in pairs:
        assert s.count(old) == cnt, (path, s.count(old), old[:60])
        s = s.replace(old, new)
    open(path, 'w').write(s)

rw('recogym_b200/csrc/rg_tensor.cuh', [
 ('''constexpr int TILE_N = 64;           // products per tile (UMMA N)
constexpr int CHUNK_BYTES = 2048;    // A: one 16-byte K chunk of all 128 rows
constexpr int CHUNK_BYTES_B = 1024;  // B: one 16-byte K chunk of all 64 rows''',
  '''#ifndef RG_TC_TILE_N
#define RG_TC_TILE_N 128
#endif
constexpr int TILE_N = RG_TC_TILE_N; // products per tile (UMMA N): 64 or 128
constexpr int CHUNK_BYTES = 2048;    // A: one 16-byte K chunk of all 128 rows
constexpr int CHUNK_BYTES_B = TILE_N / 8 * 128;  // B: one 16-byte K chunk of all TILE_N rows''', 1),
 ('''    return ((((size_t)t * KC + (k >> 2)) * 8 + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}''', '''    return ((((size_t)t * KC + (k >> 2)) * (TILE_N / 8) + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}''', 1),
 ('constexpr uint32_t kIdescTf32M128N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);',
  'constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)TILE_N >> 3) << 17) | ((128u >> 4) << 24);', 1),
 ('kIdescTf32M128N64, acc);', 'kIdescTf32, acc);', 1),
])
rw('recogym_b200/csrc/rg_capi.cu', [
 ('''    std::vector<float> GTC((size_t)nsub * 2 * tile_floats, 0.f);
    for (int p = 0; p < nsub * 128; ++p) {
        float *tile = GTC.data() + (size_t)(p / 64) * tile_floats;
        const int row = p % 64;''', '''    std::vector<float> GTC((size_t)nsub * (128 / rg::tc::TILE_N) * tile_floats, 0.f);
    for (int p = 0; p < nsub * 128; ++p) {
        float *tile = GTC.data() + (size_t)(p / rg::tc::TILE_N) * tile_floats;
        const int row = p % rg::tc::TILE_N;''', 1),
])
rw('recogym_b200/csrc/rg_tensor.cu', [
 ("    if (warp == 0) tmem_alloc(tmem_slot, 64);", "    if (warp == 0) tmem_alloc(tmem_slot, TILE_N);", 1),
 ("    for (int sub = 0; sub < 2 * nsub; ++sub) {  // 64-product tiles", "    for (int sub = 0; sub < nsub * (128 / TILE_N); ++sub) {  // TILE_N-product tiles", 1),
 ("    if (warp == 0) tmem_dealloc(tmem_base, 64);", "    if (warp == 0) tmem_dealloc(tmem_base, TILE_N);", 1),
])
F = 'recogym_b200/csrc/rg_rollout_fast.cu'
rw(F, [
 ("constexpr int NBUF = 3;      // tensor mode: accumulator tiles (64 TMEM columns each) in flight",
  "constexpr int NBUF = tc::TILE_N == 64 ? 3 : 2;  // tensor mode: accumulator tiles (TILE_N TMEM columns each) in flight", 1),
 ("    const int ntile = TC ? 2 * nsub : nsub;  // tiles streamed per round: 64 products (TC) or 128 (CUDA cores)",
  "    const int ntile = TC ? nsub * (128 / tc::TILE_N) : nsub;  // tiles streamed per round", 1),
 ("                        tc::issue_tile_mmas(tmem_base + buf * 64u, a_saddr, ring_saddr + st * tile_bytes, KC);",
  "                        tc::issue_tile_mmas(tmem_base + buf * (uint32_t)tc::TILE_N, a_saddr, ring_saddr + st * tile_bytes, KC);", 1),
 ('''                    float v[32];
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
                    if ((tt + 1) % (2 * p.rec_sub) == 0 || tt == ntile - 1) {''',
  '''                    constexpr int HALF = tc::TILE_N / 2;  // columns of this thread
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int c = 0; c < HALF; c += 32) {
                        float v[32];
                        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)tc::TILE_N +
                                          (uint32_t)(h * HALF + c), v);
                        if (c + 32 >= HALF) {  // last read of this tile: hand the accumulator back before the math
                            tc::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&dempty[buf]);
                        }
                        if (act) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                s0 += ex2_approx(v[i]); s1 += ex2_approx(v[i + 1]);
                                s2 += ex2_approx(v[i + 2]); s3 += ex2_approx(v[i + 3]);
                            }
                        }
                    }
                    part += (s0 + s1) + (s2 + s3);
                    if ((tt + 1) % (p.rec_sub * (128 / tc::TILE_N)) == 0 || tt == ntile - 1) {''', 1),
 ("    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // NBUF 64-column accumulator tiles", "    if (TC && warp == 0) tc::tmem_alloc(tmem_slot, 256);  // NBUF accumulator tiles of TILE_N columns", 1),
])
print('ok')
