// Host check of one phase-B round of k_pileup: 32 lanes simulated, the four bit matrices (V, B0, B1, Tn5) of up to 32
// candidate reads go through one, two or four 32x32 transposes (bitplane.cuh: quarter_row / half_row / *_columns /
// count_columns, the same helpers the kernel calls) and must give the per-position counters of the obvious loop.
#include <stdio.h>
#include <string.h>
#ifndef BITPLANE_HEADER
#define BITPLANE_HEADER "../../mgatk2_b200/csrc/bitplane.cuh"
#endif
#include BITPLANE_HEADER
using namespace mgatk;

static uint64_t rng = 0x2545f4914f6cdd1dull;
static u32 rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (u32)(rng >> 13); }

static void transpose32(u32 (&x)[32]) {                    // what warp_transpose does across the lanes
    u32 y[32];
    for (int j = 16; j >= 1; j >>= 1) {
        for (int l = 0; l < 32; l++)
            y[l] = j >= 8 ? transpose_stage_bytes(x[l], x[l ^ j], transpose_sel(l, j))
                          : transpose_stage(x[l], x[l ^ j], transpose_keep(l, j), transpose_amt(l, j));
        memcpy(x, y, sizeof(y));
    }
}

int main() {
    int bad = 0;
    for (int it = 0; it < 300000 && !bad; it++) {
        u32 pv[32], p0[32], p1[32], m5[32];
        u32 cands = 0, rev = 0;
        const int k = it % 3 == 0 ? 1 + (int)(rnd() % 8) : it % 3 == 1 ? 1 + (int)(rnd() % 16) : 1 + (int)(rnd() % 32);
        for (int l = 0; l < 32; l++) {
            const bool cand = l < k && (rnd() % 10) != 0;          // candidates in the low lanes, with holes
            pv[l] = p0[l] = p1[l] = m5[l] = 0;
            if (!cand) continue;
            cands |= 1u << l;
            if (rnd() & 1) rev |= 1u << l;
            pv[l] = (it & 4) ? rnd() : rnd() & rnd();
            p0[l] = rnd() & pv[l]; p1[l] = rnd() & pv[l];
            if (rnd() % 3) m5[l] = 1u << (rnd() & 31);
        }
        u32 want[32][10];
        memset(want, 0, sizeof(want));
        for (int l = 0; l < 32; l++) {
            const int s = (rev >> l) & 1;
            for (int p = 0; p < 32; p++) {
                if ((pv[l] >> p) & 1) want[p][2 * ((int)((p0[l] >> p) & 1) | (int)(((p1[l] >> p) & 1) << 1)) + s]++;
                if ((m5[l] >> p) & 1) want[p][8 + s]++;
            }
        }
        u32 got[32][10];
        memset(got, 0, sizeof(got));
        if (cands <= 0xffu) {
            u32 x[32];
            for (int l = 0; l < 32; l++) x[l] = quarter_row(l, pv[l], p0[l & 7], p1[l & 7], m5[l & 7]);
            transpose32(x);
            for (int p = 0; p < 32; p++) { u32 v, b0, b1, t5; quarter_columns(x[p], v, b0, b1, t5); count_columns(v, b0, b1, t5, rev, got[p]); }
        } else if (cands <= 0xffffu) {
            u32 xa[32], xb[32];
            for (int l = 0; l < 32; l++) { xa[l] = half_row(l, pv[l], p0[l ^ 16]); xb[l] = half_row(l, p1[l], m5[l ^ 16]); }
            transpose32(xa); transpose32(xb);
            for (int p = 0; p < 32; p++) { u32 v, b0, b1, t5; half_columns(xa[p], xb[p], v, b0, b1, t5); count_columns(v, b0, b1, t5, rev, got[p]); }
        } else {
            u32 a[32], b[32], c[32], d[32];
            memcpy(a, pv, sizeof(a)); memcpy(b, p0, sizeof(b)); memcpy(c, p1, sizeof(c)); memcpy(d, m5, sizeof(d));
            transpose32(a); transpose32(b); transpose32(c); transpose32(d);
            for (int p = 0; p < 32; p++) count_columns(a[p], b[p], c[p], d[p], rev, got[p]);
        }
        if (memcmp(want, got, sizeof(want))) { printf("round mismatch: %d candidates (mask %08x)\n", k, cands); bad++; }
    }
    if (bad) { printf("FAILED\n"); return 1; }
    printf("rounds ok\n");
    return 0;
}
