// TEST INFRASTRUCTURE (tests/test_pack_roundtrip.py): random columns through aqc_pack::pack_columns, decoded and compared byte for byte;
// pools of 1..12 threads are created and destroyed between rounds.
#include "aqc_pack.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main() {
    uint64_t s = 12345;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    size_t checked = 0;
    for (int round = 0; round < 12; round++) {
        int T = 1 + (int)(rnd() % 12);
        auto *p = aqc_pack::pool_create(T);
        for (int it = 0; it < 250; it++) {
            size_t n = rnd() % 3 == 0 ? rnd() % 100 : rnd() % 200000;
            int ncols = 1 + (int)(rnd() % 4);
            std::vector<std::vector<uint8_t>> src(ncols), dst(ncols), xv(ncols);
            std::vector<std::vector<uint32_t>> xp(ncols);
            aqc_pack::Column cols[4];
            for (int c = 0; c < ncols; c++) {
                int kind = (int)(rnd() & 1);
                size_t m = n + rnd() % 50;
                src[c].resize(m + 1); dst[c].assign(aqc_pack::packed_bytes(kind, m) + 64, 0xEE); xp[c].resize(m / 16 + 2048); xv[c].resize(m / 16 + 2048);
                int flavour = (int)(rnd() % 4);
                for (size_t i = 0; i < m; i++) {
                    uint64_t r = rnd();
                    if (kind == 0) src[c][i] = (flavour == 3 || r % 500 == 0) ? (uint8_t)(r >> 8) : "ACGT"[(r >> 20) & 3];
                    else src[c][i] = (flavour == 3 || r % 500 == 0) ? (uint8_t)(r >> 8) : (uint8_t)(33 + ((r >> 20) % 42));
                }
                cols[c] = aqc_pack::Column{kind, src[c].data(), m, dst[c].data(), xp[c].data(), xv[c].data(), m / 16 + 1024, 0, true};
            }
            aqc_pack::pack_columns(p, cols, ncols);
            for (int c = 0; c < ncols; c++) {
                if (!cols[c].ok) continue;
                // decode and compare
                std::vector<uint8_t> out(cols[c].n);
                for (size_t i = 0; i < cols[c].n; i++) {
                    if (cols[c].kind == 0) out[i] = "ACTG"[(dst[c][i >> 2] >> (2 * (i & 3))) & 3];
                    else { const uint8_t *g = dst[c].data() + 3 * (i >> 2); uint32_t v = g[0] | (g[1] << 8) | (g[2] << 16); out[i] = (uint8_t)(((v >> (6 * (i & 3))) & 63) + 33); }
                }
                for (size_t k = 0; k < cols[c].n_exc; k++) out[xp[c][k]] = xv[c][k];
                if (memcmp(out.data(), src[c].data(), cols[c].n) != 0) { printf("MISMATCH round %d it %d col %d kind %d n %zu\n", round, it, c, cols[c].kind, cols[c].n); return 1; }
                checked++;
            }
        }
        aqc_pack::pool_destroy(p);
    }
    printf("stress ok: %zu columns round-tripped\n", checked);
}
