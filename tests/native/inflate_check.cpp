// Host check of csrc/fast_inflate.h against zlib: raw DEFLATE streams of BGZF size made by zlib at every level and
// strategy (dynamic, fixed and stored blocks; text-like, BAM-like, repetitive and random data; empty and one-byte
// inputs) must decode to the same bytes, and damaged streams must be refused or fail the length check, never overrun.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <zlib.h>
#ifndef FAST_INFLATE_HEADER
#define FAST_INFLATE_HEADER "../../mgatk2_b200/csrc/fast_inflate.h"
#endif
#include FAST_INFLATE_HEADER

static uint64_t rng = 0x853c49e6748fea9bull;
static uint32_t rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 20); }

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &src, int level, int strategy) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)src.size()) + 64);
    zs.next_in = (Bytef *)src.data(); zs.avail_in = (uInt)src.size();
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(out.size() - zs.avail_out);
    deflateEnd(&zs);
    return out;
}

int main() {
    static mgatk_inflate::Tables tables;
    int bad = 0, refused = 0, n = 0;
    for (int it = 0; it < 3000 && !bad; it++) {
        const int kind = it % 6;
        size_t len = it < 40 ? (size_t)it : 1 + rnd() % 65280;
        std::vector<uint8_t> src(len);
        if (kind == 0) for (auto &b : src) b = (uint8_t)rnd();                                   // incompressible
        else if (kind == 1) for (auto &b : src) b = (uint8_t)("ACGT"[rnd() & 3]);                 // four symbols
        else if (kind == 2) for (size_t i = 0; i < len; i++) src[i] = (uint8_t)(i % 7 ? 'x' : rnd());   // long matches
        else if (kind == 3) for (size_t i = 0; i < len; i++) src[i] = (uint8_t)((i / 190) * 3 + (rnd() % 5 == 0 ? rnd() : i % 190));   // record-like
        else if (kind == 4) { if (len) memset(src.data(), 0, len); }
        else for (auto &b : src) b = (uint8_t)(rnd() % 40 + (rnd() % 50 == 0 ? 128 : 30));        // quality-like, long codes
        const int level = it % 10, strategy = (it / 10) % 4 == 3 ? Z_FIXED : (it / 10) % 4 == 2 ? Z_HUFFMAN_ONLY : Z_DEFAULT_STRATEGY;
        std::vector<uint8_t> z = deflate_raw(src, level, strategy);
        z.resize(z.size() + 8, 0xA5);                              // the BGZF trailer follows the stream
        std::vector<uint8_t> got(len + 16, 0xEE);
        n++;
        if (!mgatk_inflate::inflate_raw(z.data(), z.size(), got.data(), len, tables)) { refused++; printf("refused: it=%d len=%zu level=%d strategy=%d\n", it, len, level, strategy); continue; }
        if (len && memcmp(got.data(), src.data(), len)) { printf("mismatch: it=%d len=%zu level=%d\n", it, len, level); bad++; }
        for (int k = 0; k < 16; k++) if (got[len + k] != 0xEE) { printf("overrun: it=%d len=%zu at +%d\n", it, len, k); bad++; break; }
        // wrong output length: must not succeed, must not write past the buffer it was given
        if (len > 4) {
            std::vector<uint8_t> small(len - 3 + 16, 0xEE);
            if (mgatk_inflate::inflate_raw(z.data(), z.size(), small.data(), len - 3, tables)) { printf("short buffer accepted: it=%d\n", it); bad++; }
            for (int k = 0; k < 16; k++) if (small[len - 3 + k] != 0xEE) { printf("overrun (short): it=%d\n", it); bad++; break; }
        }
        // damaged stream: any outcome but an overrun or a crash
        if (z.size() > 12) {
            std::vector<uint8_t> d = z;
            d[rnd() % (d.size() - 8)] ^= (uint8_t)(1u << (rnd() & 7));
            std::vector<uint8_t> g2(len + 16, 0xEE);
            mgatk_inflate::inflate_raw(d.data(), d.size(), g2.data(), len, tables);
            for (int k = 0; k < 16; k++) if (g2[len + k] != 0xEE) { printf("overrun (damaged): it=%d\n", it); bad++; break; }
        }
    }
    if (refused) { printf("%d of %d valid streams refused\n", refused, n); bad++; }
    if (bad) { printf("FAILED\n"); return 1; }
    printf("inflate ok (%d streams)\n", n);
    return 0;
}
