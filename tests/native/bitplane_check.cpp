// Host check of mgatk2_b200/csrc/bitplane.cuh against per-base loops. Exit code 0 = all good.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../mgatk2_b200/csrc/bitplane.cuh"
using namespace mgatk;

static uint64_t rng = 0x9e3779b97f4a7c15ull;
static u32 rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (u32)(rng >> 16); }

int main() {
    int bad = 0;
    // qual compare: every threshold, every byte value in every byte lane
    for (int minq = -200; minq <= 200 && !bad; minq++) {
        QualGe g = make_qual_ge(minq);
        for (int v = 0; v < 256; v++) {
            const bool want = !g.none && (int)(int8_t)v >= minq;
            if (g.none) { if ((int)(int8_t)v >= minq) { printf("none wrong minq=%d v=%d\n", minq, v); bad++; } continue; }
            for (int lane = 0; lane < 4; lane++) {
                u32 q = rnd(); q &= ~(0xffu << (8 * lane)); q |= (u32)v << (8 * lane);
                const bool got = (qual_ge4(q, g) >> (8 * lane + 7)) & 1;
                if (got != want) { printf("qual_ge4 minq=%d v=%d lane=%d\n", minq, v, lane); bad++; break; }
            }
        }
    }
    for (int it = 0; it < 2000000 && !bad; it++) {
        const int minq = (int)(rnd() % 300) - 150;
        QualGe g = make_qual_ge(minq);
        if (g.none) continue;
        const u32 q0 = rnd(), q1 = rnd();
        u32 want = 0;
        for (int i = 0; i < 8; i++) { const int8_t b = (int8_t)(((i < 4 ? q0 : q1) >> (8 * (i & 3))) & 255); if ((int)b >= minq) want |= 1u << (24 + i); }
        if (qual_ok8_top(q0, q1, g) != want) { printf("qual_ok8_top mismatch\n"); bad++; }
        if ((qual_ok8_raw(q0, q1, g) & 0xff000000u) != want) { printf("qual_ok8_raw mismatch\n"); bad++; }
        if (g.sel == 0u && (qual_ok8_raw<true>(q0, q1, g) & 0xff000000u) != want) { printf("qual_ok8_raw<and> mismatch minq=%d\n", minq); bad++; }
    }
    // seq: all nibble patterns by random words + structured words
    for (int it = 0; it < 4000000 && !bad; it++) {
        u32 s = rnd();
        if (it & 1) { u32 t = 0; for (int n = 0; n < 8; n++) t |= (1u << (rnd() & 3)) << (4 * n); s = (it & 2) ? t : (t ^ ((rnd() & rnd() & rnd()) & 0xffffffffu)); }
        u32 want[4] = {0, 0, 0, 0};
        for (int i = 0; i < 8; i++) {
            const u32 byte = (s >> (8 * (i >> 1))) & 255;
            const u32 nib = (i & 1) ? (byte & 15) : (byte >> 4);
            for (int x = 0; x < 4; x++) if (nib == (1u << x)) want[x] |= 1u << (24 + i);
        }
        const Planes8 e = seq_planes8_top(s);
        const u32 v = want[0] | want[1] | want[2] | want[3];
        if (e.v != v || (e.b0 & v) != (want[1] | want[3]) || (e.b1 & v) != (want[2] | want[3])) { printf("seq_planes8_top mismatch %08x\n", s); bad++; }
        const Planes8 t = seq_planes8_lut(s);                 // the table-lookup form: b0, b1 exact, scrap below the top byte
        if ((t.v & 0xff000000u) != v || (t.b0 & 0xff000000u) != (want[1] | want[3]) || (t.b1 & 0xff000000u) != (want[2] | want[3])) { printf("seq_planes8_lut mismatch %08x\n", s); bad++; }
        u32 back[4];
        planes_to_bases(e.v, e.b0 & e.v, e.b1 & e.v, back);
        for (int x = 0; x < 4; x++) if (back[x] != want[x]) { printf("planes_to_bases mismatch %08x\n", s); bad++; break; }
    }
    for (u32 h = 0; h < 65536 && !bad; h++) {                // every pair of SEQ bytes in both halves of the word
        const u32 s = h * 0x00010001u;
        const Planes8 e = seq_planes8_top(s), t = seq_planes8_lut(s);
        if ((t.v & 0xff000000u) != e.v || (t.b0 & 0xff000000u) != (e.b0 & e.v) || (t.b1 & 0xff000000u) != (e.b1 & e.v)) { printf("seq_planes8_lut exhaustive %08x\n", s); bad++; }
    }
    // every quality byte against every threshold through the raw form (mask applied by the caller)
    for (int minq = -200; minq <= 200 && !bad; minq++) {
        QualGe g = make_qual_ge(minq);
        if (g.none) continue;
        for (int v = 0; v < 256; v++) for (int lane = 0; lane < 4; lane++) {
            u32 q = rnd(); q &= ~(0xffu << (8 * lane)); q |= (u32)v << (8 * lane);
            if (((qual_ge4_raw(q, g) >> (8 * lane + 7)) & 1) != (u32)((int)(int8_t)v >= minq)) { printf("qual_ge4_raw minq=%d v=%d\n", minq, v); bad++; }
        }
    }
    // bit_range / funnel_r
    for (int lo = -40; lo <= 40; lo++) for (int hi = -40; hi <= 72; hi++) {
        u32 want = 0; for (int b = 0; b < 32; b++) if (b >= lo && b < hi) want |= 1u << b;
        if (bit_range(lo, hi) != want) { printf("bit_range %d %d\n", lo, hi); bad++; }
    }
    for (int it = 0; it < 100000; it++) {
        const u32 lo = rnd(), hi = rnd(), sh = rnd() & 31;
        const uint64_t v = ((uint64_t)hi << 32) | lo;
        if (funnel_r(lo, hi, sh) != (u32)(v >> sh)) { printf("funnel_r\n"); bad++; }
    }
    // transpose: simulate the 32 lanes
    for (int it = 0; it < 2000 && !bad; it++) {
        u32 x[32], y[32], org[32];
        for (int l = 0; l < 32; l++) org[l] = x[l] = rnd() & ((it % 3) ? 0xffffffffu : rnd());
        for (int j = 16; j >= 1; j >>= 1) {
            for (int l = 0; l < 32; l++) {
                y[l] = transpose_stage(x[l], x[l ^ j], transpose_keep(l, j), transpose_amt(l, j));
                if (j >= 8 && transpose_stage_bytes(x[l], x[l ^ j], transpose_sel(l, j)) != y[l]) { printf("transpose_stage_bytes j=%d lane=%d\n", j, l); bad++; }
            }
            memcpy(x, y, sizeof(x));
        }
        for (int p = 0; p < 32; p++) for (int r = 0; r < 32; r++)
            if (((x[p] >> r) & 1) != ((org[r] >> p) & 1)) { printf("transpose\n"); bad++; p = 32; break; }
    }
    // planning: lower_bound_near against the plain definition, any guess, ties, ranges of every size
    {
        struct Starts { const int *p; int pos(int i) const { return p[i]; } };
        static int arr[5000];
        for (int it = 0; it < 300000 && !bad; it++) {
            const int n = (it % 7 == 0) ? (int)(rnd() % 5000) : (int)(rnd() % 60);
            const int spread = 1 + (int)(rnd() % ((it & 1) ? 40 : 20000));
            int v = (int)(rnd() % 100) - 50;
            for (int i = 0; i < n; i++) { v += (int)(rnd() % spread) * (int)(rnd() & 1); arr[i] = v; }
            const int lo = n ? (int)(rnd() % (n + 1)) : 0, hi = lo + (n - lo ? (int)(rnd() % (n - lo + 1)) : 0);
            const int key = (n ? arr[rnd() % n] : 0) + (int)(rnd() % 5) - 2;
            const int guess = (int)(rnd() % (n + 20)) - 10;
            int want = lo;
            while (want < hi && arr[want] < key) want++;
            const Starts sl{arr};
            if (lower_bound_near(sl, lo, hi, guess, key) != want) { printf("lower_bound_near n=%d lo=%d hi=%d guess=%d key=%d\n", n, lo, hi, guess, key); bad++; }
        }
    }
    if (bad) { printf("FAILED %d\n", bad); return 1; }
    printf("bitplane helpers ok\n");
    return 0;
}
