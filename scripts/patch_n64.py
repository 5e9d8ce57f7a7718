import re
def rw(path, fn):
    s = open(path).read(); s2 = fn(s); open(path, 'w').write(s2)
def sub1(s, old, new):
    assert s.count(old) == 1, (s.count(old), old[:70]); return s.replace(old, new)

# ---------------- rg_tensor.cuh
def f(s):
    s = sub1(s, '''// Logits of 128 users x 128 products are one UMMA tile:  D[u][p] = sum_k A[u][k] * B[p][k]  with
//   A = omega tile   (128 x KT tf32, K-major), rebuilt by the CTA every round,
//   B = Gamma tile   (128 x KT tf32, K-major), precomputed on the host (GTC array),
// both in the canonical no-swizzle K-major layout of the UMMA shared-memory descriptor:
//   element (row r, column k) -> chunk kc = k/4 (16 bytes), row group g = r/8:
//   byte offset = kc * 2048 + g * 128 + (r % 8) * 16 + (k % 4) * 4        (2048 = 16 row groups x 128 B)
// so LBO (between the two 16-byte K chunks of one K=8 MMA step) = 2048 B and SBO (between 8-row groups) = 128 B.''',
'''// Logits of 128 users x 64 products are one UMMA tile:  D[u][p] = sum_k A[u][k] * B[p][k]  with
//   A = omega tile   (128 x KT tf32, K-major), rebuilt by the CTA every round,
//   B = Gamma tile   ( 64 x KT tf32, K-major), precomputed on the host (GTC array),
// both in the canonical no-swizzle K-major layout of the UMMA shared-memory descriptor:
//   element (row r, column k) -> chunk kc = k/4 (16 bytes), row group g = r/8:
//   byte offset = kc * (rows/8 * 128) + g * 128 + (r % 8) * 16 + (k % 4) * 4
// so LBO (between the two 16-byte K chunks of one K=8 MMA step) = 2048 B for A, 1024 B for B, and SBO (between
// 8-row groups) = 128 B.  (64-product tiles keep four accumulator tiles in 256 TMEM columns and make the
// L2 -> shared-memory stream of B fine grained: 12 KB per stage at K = 20.)''')
    s = sub1(s, '''constexpr int TILE = 128;            // users per tile (UMMA M) and products per tile (UMMA N)
constexpr int CHUNK_BYTES = 2048;    // one 16-byte K chunk of all 128 rows''',
'''constexpr int TILE = 128;            // users per tile (UMMA M)
constexpr int TILE_N = 64;           // products per tile (UMMA N)
constexpr int CHUNK_BYTES = 2048;    // A: one 16-byte K chunk of all 128 rows
constexpr int CHUNK_BYTES_B = 1024;  // B: one 16-byte K chunk of all 64 rows''')
    s = sub1(s, '''// float index of element (row, k) of term t inside one operand tile [2 terms][KC][16][8][4]
__host__ __device__ inline size_t tile_index(int t, int KC, int row, int k) {
    return ((((size_t)t * KC + (k >> 2)) * 16 + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}''', '''// float index of element (row, k) of term t inside one operand tile [2 terms][KC][rows/8][8][4]
__host__ __device__ inline size_t tile_index(int t, int KC, int row, int k) {  // A: 128 rows
    return ((((size_t)t * KC + (k >> 2)) * 16 + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}
__host__ __device__ inline size_t tile_index_b(int t, int KC, int row, int k) {  // B: 64 rows
    return ((((size_t)t * KC + (k >> 2)) * 8 + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}''')
    s = sub1(s, '''__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(CHUNK_BYTES >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdescTf32M128N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);''',
'''__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = 64 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdescTf32M128N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);''')
    a = s.index("// One thread: the three-term tf32 contraction of one 128 x 128 tile")
    b = s.index("#endif\n\n}  // namespace tc")
    s = s[:a] + '''// One thread: the three-term tf32 contraction of one 128 x 64 tile into TMEM columns [tmem_d, tmem_d+64)
__device__ __forceinline__ void issue_tile_mmas(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, int KC) {
    const uint32_t aterm = (uint32_t)KC * CHUNK_BYTES, bterm = (uint32_t)KC * CHUNK_BYTES_B;  // hi -> lo term
    uint32_t acc = 0;
    // small terms first: lo*hi, hi*lo, then hi*hi
#pragma unroll 1
    for (int pair = 0; pair < 3; ++pair) {
        const uint32_t a0 = a_saddr + (pair == 0 ? aterm : 0u);
        const uint32_t b0 = b_saddr + (pair == 1 ? bterm : 0u);
        for (int s = 0; s < KC / 2; ++s) {
            umma_tf32(tmem_d, umma_desc(a0 + (uint32_t)s * 2u * CHUNK_BYTES, CHUNK_BYTES),
                      umma_desc(b0 + (uint32_t)s * 2u * CHUNK_BYTES_B, CHUNK_BYTES_B), kIdescTf32M128N64, acc);
            acc = 1;
        }
    }
}
''' + s[b:]
    return s
rw('recogym_b200/csrc/rg_tensor.cuh', f)

# ---------------- rg_capi.cu: GTC per 64-product tile
def f(s):
    s = sub1(s, '''    const size_t tile_floats = (size_t)2 * KC * (rg::tc::CHUNK_BYTES / 4);
    std::vector<float> GTC((size_t)nsub * tile_floats, 0.f);
    for (int p = 0; p < nsub * 128; ++p) {
        float *tile = GTC.data() + (size_t)(p / 128) * tile_floats;
        const int row = p % 128;''', '''    const size_t tile_floats = (size_t)2 * KC * (rg::tc::CHUNK_BYTES_B / 4);   // one 64-product tile, hi + lo
    std::vector<float> GTC((size_t)nsub * 2 * tile_floats, 0.f);
    for (int p = 0; p < nsub * 128; ++p) {
        float *tile = GTC.data() + (size_t)(p / 64) * tile_floats;
        const int row = p % 64;''')
    s = sub1(s, '''            tile[rg::tc::tile_index(0, KC, row, k)] = hi;
            tile[rg::tc::tile_index(1, KC, row, k)] = rg::tc::tf32_hi(x - hi);''', '''            tile[rg::tc::tile_index_b(0, KC, row, k)] = hi;
            tile[rg::tc::tile_index_b(1, KC, row, k)] = rg::tc::tf32_hi(x - hi);''')
    return s
rw('recogym_b200/csrc/rg_capi.cu', f)

# ---------------- rg_tensor.cu: test kernel on 64-product tiles
def f(s):
    s = sub1(s, '''    const unsigned tile_bytes = 2u * (unsigned)KC * CHUNK_BYTES;
    float *a_tile = (float *)smem;
    float *b_tile = (float *)(smem + tile_bytes);
    uint64_t *bars = (uint64_t *)(smem + 2 * tile_bytes);  // [0] TMA, [1] MMA''', '''    const unsigned a_bytes = 2u * (unsigned)KC * CHUNK_BYTES, tile_bytes = 2u * (unsigned)KC * CHUNK_BYTES_B;
    float *a_tile = (float *)smem;
    float *b_tile = (float *)(smem + a_bytes);
    uint64_t *bars = (uint64_t *)(smem + a_bytes + tile_bytes);  // [0] TMA, [1] MMA''')
    s = sub1(s, "    if (warp == 0) tmem_alloc(tmem_slot, 128);", "    if (warp == 0) tmem_alloc(tmem_slot, 64);")
    s = sub1(s, "    for (int sub = 0; sub < nsub; ++sub) {\n        const unsigned parity = (unsigned)sub & 1u;", "    for (int sub = 0; sub < 2 * nsub; ++sub) {  // 64-product tiles\n        const unsigned parity = (unsigned)sub & 1u;")
    s = sub1(s, '''        for (int c0 = 0; c0 < TILE; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) out[(size_t)row * ((size_t)nsub * TILE) + (size_t)sub * TILE + c0 + i] = v[i];
        }''', '''        for (int c0 = 0; c0 < TILE_N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) out[(size_t)row * ((size_t)nsub * TILE) + (size_t)sub * TILE_N + c0 + i] = v[i];
        }''')
    s = sub1(s, "    if (warp == 0) tmem_dealloc(tmem_base, 128);", "    if (warp == 0) tmem_dealloc(tmem_base, 64);")
    s = sub1(s, "    const size_t smem = 2 * 2 * (size_t)KC * CHUNK_BYTES + 64;", "    const size_t smem = 2 * (size_t)KC * (CHUNK_BYTES + CHUNK_BYTES_B) + 64;")
    return s
rw('recogym_b200/csrc/rg_tensor.cu', f)
print('ok')
