// Host check of build_query_masks / query_window against the per-base rule of pileup.py:67-86.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../mgatk2_b200/csrc/bitplane.cuh"
using namespace mgatk;

static uint64_t rng = 88172645463325252ull;
static u32 rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (u32)(rng >> 11); }

struct HostMem {
    uint8_t *p;
    u32 ld32(u32 a) const { u32 v; memcpy(&v, p + a, 4); return v; }
    void st128(u32 a, u32 x, u32 y, u32 z, u32 w) const { u32 v[4] = {x, y, z, w}; memcpy(p + a, v, 16); }
    void ld128(u32 a, u32 (&v)[4]) const { memcpy(v, p + a, 16); }
};

int main() {
    int bad = 0;
    std::vector<uint8_t> buf(8192);
    for (int it = 0; it < 200000 && !bad; it++) {
        for (auto &b : buf) b = (uint8_t)rnd();
        const int ncig = rnd() % 5, L = 1 + rnd() % ((it & 7) ? 160 : 700);
        const int d = (it % 5 == 0) ? 0 : (int)(rnd() % 12);
        const int minq = (it % 11 == 0) ? (int)(rnd() % 300) - 150 : (int)(rnd() % 45);
        const u32 blob = 16 * (rnd() % 8);
        const u32 seq = blob + 4 * ncig, qual = seq + (L + 1) / 2;
        // plausible content: mostly ACGT, quals 0..60, some odd bytes
        for (int i = 0; i < (L + 1) / 2; i++) {
            auto nib = [&]() { u32 r = rnd() % 100; return r < 90 ? (1u << (rnd() & 3)) : r < 95 ? 15u : (rnd() & 15); };
            buf[seq + i] = (uint8_t)((nib() << 4) | nib());
        }
        for (int i = 0; i < L; i++) buf[qual + i] = (uint8_t)((rnd() % 50 == 0) ? rnd() : rnd() % 61);
        const int blob16 = (4 * ncig + (L + 1) / 2 + L + 15) & ~15;
        const u32 out = blob + blob16;
        const int nq = (L + 31) / 32;
        QualGe qg = make_qual_ge(minq);
        int lo = d > 0 ? d : 0, hi = d > 0 ? L - d : L;
        if (qg.none) hi = lo;
        std::vector<uint8_t> seqc(buf.begin() + seq, buf.begin() + seq + (L + 1) / 2), qualc(buf.begin() + qual, buf.begin() + qual + L);
        HostMem mem{buf.data()};
        build_query_masks(mem, seq, out, L, lo, hi, qg);
        // expected
        std::vector<u32> want(4 * nq, 0);
        for (int q = 0; q < L; q++) {
            if (!(lo <= q && q < hi)) continue;
            if ((int)(int8_t)qualc[q] < minq) continue;
            const u32 nibv = (q & 1) ? (seqc[q >> 1] & 15) : (seqc[q >> 1] >> 4);
            for (int x = 0; x < 4; x++) if (nibv == (1u << x)) want[4 * (q >> 5) + x] |= 1u << (q & 31);
        }
        std::vector<u32> planes(4 * nq, 0);                  // V, B0, B1, 0 per group of 32 bases
        for (int w = 0; w < nq; w++) {
            planes[4 * w] = want[4 * w] | want[4 * w + 1] | want[4 * w + 2] | want[4 * w + 3];
            planes[4 * w + 1] = want[4 * w + 1] | want[4 * w + 3];
            planes[4 * w + 2] = want[4 * w + 2] | want[4 * w + 3];
        }
        if (memcmp(planes.data(), buf.data() + out, 16 * nq)) { printf("build_query_masks mismatch L=%d d=%d minq=%d ncig=%d\n", L, d, minq, ncig); bad++; break; }
        for (int k = 0; k < 8; k++) {
            const int qb = (int)(rnd() % (L + 100)) - 50;
            u32 got3[3], got[4];
            query_window(mem, out, nq, qb, got3);
            planes_to_bases(got3[0], got3[1], got3[2], got);
            for (int x = 0; x < 4; x++) {
                u32 w = 0;
                for (int b = 0; b < 32; b++) { const int q = qb + b; if (q >= 0 && q < 32 * nq && ((want[4 * (q >> 5) + x] >> (q & 31)) & 1)) w |= 1u << b; }
                if (w != got[x]) { printf("query_window mismatch qb=%d\n", qb); bad++; }
            }
        }
    }
    if (bad) { printf("FAILED\n"); return 1; }
    printf("query masks ok\n");
    return 0;
}
