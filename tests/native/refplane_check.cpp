// Host check of the query -> reference coordinate conversion (bitplane.cuh: query_mask_group, ref_group,
// QueryPlanes64, QueryPlanesMem, cigar_ref_span) against the per-base CIGAR walk of pileup.py:52-95.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../mgatk2_b200/csrc/bitplane.cuh"
using namespace mgatk;

static uint64_t rng = 0x2545F4914F6CDD1Dull;
static u32 rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (u32)(rng >> 11); }

struct HostMem {
    uint8_t *p;
    u32 ld32(u32 a) const { u32 v; memcpy(&v, p + a, 4); return v; }
    void st128(u32 a, u32 x, u32 y, u32 z, u32 w) const { u32 v[4] = {x, y, z, w}; memcpy(p + a, v, 16); }
    void ld128(u32 a, u32 (&v)[4]) const { memcpy(v, p + a, 16); }
};

int main() {
    int bad = 0;
    std::vector<uint8_t> buf(16384);
    for (int it = 0; it < 300000 && !bad; it++) {
        for (auto &b : buf) b = (uint8_t)rnd();
        const bool small = it & 1;                          // small: the 64-bit register form (reads of <= 56 bases)
        const int L = 1 + rnd() % (small ? 56 : 250);
        const int d = (it % 5 == 0) ? 0 : (int)(rnd() % 12);
        const int minq = (it % 11 == 0) ? (int)(rnd() % 300) - 150 : (int)(rnd() % 45);
        // random CIGAR: query-consuming ops roughly add up to L, sometimes not (malformed inputs must not crash)
        std::vector<u32> cig;
        const int kind = rnd() % 4;
        if (kind == 0) cig.push_back(((u32)L << 4) | 0);
        else {
            int left = L;
            const int nops = 1 + rnd() % 9;
            for (int o = 0; o < nops && left > 0; o++) {
                static const int ops[] = {0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8};
                const int op = ops[rnd() % 11];
                int n = 1 + rnd() % (op == 0 || op == 7 || op == 8 ? (left > 1 ? left : 1) : 12);
                if (o == nops - 1 && (rnd() & 3)) { cig.push_back(((u32)left << 4) | 0); left = 0; break; }
                cig.push_back(((u32)n << 4) | (u32)op);
                if (op == 0 || op == 7 || op == 8 || op == 4 || op == 1) left -= n;
            }
            if (it % 97 == 0) cig.push_back(((u32)(1 + rnd() % 100) << 4) | 0);        // overlong
        }
        const int ncig = (int)cig.size();
        const u32 blob = 16 * (rnd() % 8);
        const u32 seq = blob + 4 * ncig, qual = seq + (L + 1) / 2;
        memcpy(buf.data() + blob, cig.data(), 4 * ncig);
        for (int i = 0; i < (L + 1) / 2; i++) {
            auto nib = [&]() { u32 r = rnd() % 100; return r < 90 ? (1u << (rnd() & 3)) : r < 95 ? 15u : (rnd() & 15); };
            buf[seq + i] = (uint8_t)((nib() << 4) | nib());
        }
        for (int i = 0; i < L; i++) buf[qual + i] = (uint8_t)((rnd() % 50 == 0) ? rnd() : rnd() % 61);
        std::vector<uint8_t> seqc(buf.begin() + seq, buf.begin() + seq + (L + 1) / 2), qualc(buf.begin() + qual, buf.begin() + qual + L);
        const int blob16 = (4 * ncig + (L + 1) / 2 + L + 15) & ~15;
        const u32 scratch = blob + blob16 + 64;
        const int nq = (L + 31) / 32;
        QualGe qg = make_qual_ge(minq);
        int lo = d > 0 ? d : 0, hi = d > 0 ? L - d : L;
        if (qg.none) hi = lo;
        HostMem mem{buf.data()};
        // expected: per-base walk (pileup.py:52-95), reference offsets relative to reference_start
        long long span = 0;
        std::vector<int> want;                               // base index 0..3 or -1 per reference offset
        {
            long long ref = 0, qp = 0;
            for (u32 w : cig) {
                const int op = w & 15; const long long n = w >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    for (long long i = 0; i < n; i++) {
                        const long long q = qp + i, r = ref + i;
                        if ((long long)want.size() <= r) want.resize(r + 1, -1);
                        if (!(lo <= q && q < hi) || q >= L) continue;
                        if ((int)(int8_t)qualc[q] < minq) continue;
                        const u32 nibv = (q & 1) ? (seqc[q >> 1] & 15) : (seqc[q >> 1] >> 4);
                        for (int x = 0; x < 4; x++) if (nibv == (1u << x)) want[r] = x;
                    }
                    ref += n; qp += n;
                } else if (op == 2 || op == 3) ref += n;
                else if (op == 4) qp += n;
            }
            span = ref;
            if ((long long)want.size() < ref) want.resize(ref, -1);
        }
        if (cigar_ref_span(mem, blob, ncig) != (int)span) { printf("span mismatch\n"); bad++; break; }
        const int W = (int)((span + 31) / 32) + 1;
        std::vector<u32> got(3 * W, 0);
        for (int w = 0; w < nq; w++) {                          // the straight-line form of a group gives the same planes
            u32 a0, a1, a2, b0, b1, b2;
            query_mask_group(mem, seq, L, w, lo, hi, qg, a0, a1, a2);
            query_mask_group_straight(mem, seq, L, w, lo, hi, qg, b0, b1, b2);
            if (a0 != b0 || a1 != b1 || a2 != b2) { printf("query_mask_group_straight mismatch L=%d w=%d d=%d minq=%d\n", L, w, d, minq); bad++; break; }
            query_mask_group_lut(mem, seq, L, w, lo, hi, qg, b0, b1, b2);      // and so does the table-lookup form
            if (a0 != b0 || a1 != b1 || a2 != b2) { printf("query_mask_group_lut mismatch L=%d w=%d d=%d minq=%d\n", L, w, d, minq); bad++; break; }
            if (qg.sel == 0u && !qg.none) {
                query_mask_group_lut<true>(mem, seq, L, w, lo, hi, qg, b0, b1, b2);
                if (a0 != b0 || a1 != b1 || a2 != b2) { printf("query_mask_group_lut<and> mismatch L=%d w=%d d=%d minq=%d\n", L, w, d, minq); bad++; break; }
            }
        }
        if (small) {
            u32 g[2][3] = {{0, 0, 0}, {0, 0, 0}};
            for (int w = 0; w < nq; w++) query_mask_group(mem, seq, L, w, lo, hi, qg, g[w][0], g[w][1], g[w][2]);
            u32 g56[2][3];
            query_planes56(mem, seq, L, lo, hi, qg, g56);           // the all-at-once form gives the same planes
            if (memcmp(g56, g, sizeof(g))) { printf("query_planes56 mismatch L=%d d=%d minq=%d\n", L, d, minq); bad++; break; }
            query_planes56_straight<7>(mem, seq, L, lo, hi, qg, g56);  // and so does the form without per-group conditions
            if (memcmp(g56, g, sizeof(g))) { printf("query_planes56_straight mismatch L=%d d=%d minq=%d\n", L, d, minq); bad++; break; }
            if (hi <= 40) { query_planes56_straight<5>(mem, seq, L, lo, hi, qg, g56); if (memcmp(g56, g, sizeof(g))) { printf("straight<5> mismatch\n"); bad++; break; } }
            if (hi <= 48) { query_planes56_straight<6>(mem, seq, L, lo, hi, qg, g56); if (memcmp(g56, g, sizeof(g))) { printf("straight<6> mismatch\n"); bad++; break; } }
            query_planes56_lut<7>(mem, seq, L, lo, hi, qg, g56);       // table-lookup planes (PRMT), quality mask applied once
            if (memcmp(g56, g, sizeof(g))) { printf("query_planes56_lut mismatch L=%d d=%d minq=%d\n", L, d, minq); bad++; break; }
            if (hi <= 40) { query_planes56_lut<5>(mem, seq, L, lo, hi, qg, g56); if (memcmp(g56, g, sizeof(g))) { printf("lut<5> mismatch\n"); bad++; break; } }
            if (hi <= 48) { query_planes56_lut<6>(mem, seq, L, lo, hi, qg, g56); if (memcmp(g56, g, sizeof(g))) { printf("lut<6> mismatch\n"); bad++; break; } }
            if (qg.sel == 0u && !qg.none) {                          // min_baseq in [0, 127]: the form that knows it
                query_planes56_lut<7, true>(mem, seq, L, lo, hi, qg, g56);
                if (memcmp(g56, g, sizeof(g))) { printf("query_planes56_lut<7, and> mismatch L=%d d=%d minq=%d\n", L, d, minq); bad++; break; }
            }
            QueryPlanes64 q;
            q.v = ((unsigned long long)g[1][0] << 32) | g[0][0]; q.b0 = ((unsigned long long)g[1][1] << 32) | g[0][1];
            q.b1 = ((unsigned long long)g[1][2] << 32) | g[0][2];
            for (int k = 0; k < W; k++) { u32 o[3]; ref_group(mem, blob, ncig, q, k, o); got[3 * k] = o[0]; got[3 * k + 1] = o[1]; got[3 * k + 2] = o[2]; }
        } else {
            build_query_masks(mem, seq, scratch, L, lo, hi, qg);
            QueryPlanesMem<HostMem> q{mem, scratch, nq};
            for (int k = 0; k < W; k++) { u32 o[3]; ref_group(mem, blob, ncig, q, k, o); got[3 * k] = o[0]; got[3 * k + 1] = o[1]; got[3 * k + 2] = o[2]; }
        }
        for (int r = 0; r < 32 * W && !bad; r++) {
            const int wb = r < (int)want.size() ? want[r] : -1;
            const u32 v = (got[3 * (r >> 5)] >> (r & 31)) & 1, b0 = (got[3 * (r >> 5) + 1] >> (r & 31)) & 1, b1 = (got[3 * (r >> 5) + 2] >> (r & 31)) & 1;
            const int gb = v ? (int)(b0 | (b1 << 1)) : -1;
            if ((!v && (b0 | b1)) || gb != wb) { printf("ref plane mismatch it=%d r=%d want=%d got=%d L=%d ncig=%d small=%d\n", it, r, wb, gb, L, ncig, (int)small); bad++; }
        }
    }
    if (bad) { printf("FAILED\n"); return 1; }
    printf("reference planes ok\n");
    return 0;
}
