// mgatk2_b200 — word-parallel helpers of the bit-plane pileup (stage 3 of the hot path).
//
// A read's bases are turned once into three bit planes in QUERY coordinates: bit q of V is set iff
// SEQ[q] is one of A, C, G, T, the base quality passes and q lies inside the distance-from-end window
// (pileup.py:67-86); B0 and B1 hold the two bits of the base code (A=0, C=1, G=2, T=3) under V.
// Counting a chunk of 32 reference positions then reduces to picking a 32-bit window out of these
// planes per aligned block (pileup.py:55-65) and a 32x32 bit transpose across the warp, so that each
// lane (= position) can popcount the reads that carry a given base there. Three planes instead of one
// mask per base: one transpose less per round, and the spare quarter of the matrix takes the Tn5 bit.
//
// Everything here is plain integer arithmetic and compiles for the host too: tests/test_bitplane_host.py
// checks each helper exhaustively / on random words against the obvious per-base loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace mgatk {

typedef unsigned int u32;

MG_HD u32 funnel_r(u32 lo, u32 hi, u32 sh) {            // low word of (hi:lo) >> sh, sh in [0, 31]
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

MG_HD u32 rotl32(u32 x, u32 n) {                        // n in [0, 31]
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, n);
#else
    return n ? (x << n) | (x >> (32 - n)) : x;
#endif
}

// bits [lo, hi) of a 32-bit word; lo/hi may lie outside [0, 32]
MG_HD u32 bit_range(int lo, int hi) {
    if (lo < 0) lo = 0;
    if (hi > 32) hi = 32;
    if (hi <= lo) return 0u;
    const u32 upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return upto & (0xffffffffu << lo);
}

// ---- base quality: (int8)qual >= min_baseq (pileup.py:80-81; quals are stored np.int8, readers.py:158) ----
// Compare on u = q ^ 0x80 (unsigned order == signed order) against t = min_baseq + 128, four bytes per word.
struct QualGe { u32 add, sel; int none; };

MG_HD QualGe make_qual_ge(int min_baseq) {
    QualGe g;
    int t = min_baseq + 128;
    g.none = t > 255;                                   // nothing passes: the caller empties the window instead
    if (t > 255) t = 255;
    if (t < 0) t = 0;                                   // everything passes
    g.add = (u32)(0x80 - (t & 0x7f)) * 0x01010101u;     // bit 7 of (low7 + add) set iff low7 >= t's low seven bits
    g.sel = (t & 0x80) ? 0u : 0xffffffffu;              // t >= 128: need the top bit AND the low part; else OR
    return g;
}

// bit 7 of each byte set iff that quality passes
MG_HD u32 qual_ge4(u32 q, QualGe g) {
    const u32 a = (q & 0x7f7f7f7fu) + g.add;
    const u32 u = ~q;                                   // bit 7 of q ^ 0x80
    return ((u & a) | ((u | a) & g.sel)) & 0x80808080u;
}

// pass flags of eight consecutive quality bytes (two words) as the top byte: bit 24+i = byte i passes.
// The flags of the first word move to bits 3, 11, 19, 27, next to those of the second word at 7, 15, 23, 31; one
// multiplication then lines all eight up in the top byte (every partial product lands on its own bit).
MG_HD u32 qual_ok8_top(u32 q0, u32 q1, QualGe g) {
    const u32 f = (qual_ge4(q0, g) >> 4) | qual_ge4(q1, g);
    return (f * 0x00204081u) & 0xff000000u;
}

// The same with the byte mask applied once, after the two words met: bit 7 of each byte of qual_ge4_raw is the flag, the
// other bits are scrap. (~q, a, sel) -> flag is ONE three-input function: sel = 0 gives ~q & a, sel = ~0 gives ~q | a.
MG_HD u32 qual_ge4_raw(u32 q, QualGe g) {
    const u32 a = (q & 0x7f7f7f7fu) + g.add;
    return ((~q ^ g.sel) & (a ^ g.sel)) ^ g.sel;
}
// kAnd: the caller knows g.sel == 0 (min_baseq in [0, 127], every real run): flag = ~q & a, which takes the byte mask
// along in the same operation - two instructions fewer per eight bases.
template <bool kAnd = false>
MG_HD u32 qual_ok8_raw(u32 q0, u32 q1, QualGe g) {       // top byte = the eight flags, lower bits scrap
    if (kAnd) {
        const u32 r0 = ((q0 & 0x7f7f7f7fu) + g.add) & ~q0 & 0x80808080u, r1 = ((q1 & 0x7f7f7f7fu) + g.add) & ~q1 & 0x80808080u;
        return ((r0 >> 4) | r1) * 0x00204081u;
    }
    const u32 x = qual_ge4_raw(q0, g) >> 4, y = qual_ge4_raw(q1, g);
#if defined(__CUDA_ARCH__)
    u32 f;                                               // (x & 0x0f..) | (y & 0xf0..): flags at bits 3 and 7 of every byte
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(f) : "r"(x), "r"(y), "r"(0x0f0f0f0fu));
#else
    const u32 f = (x & 0x0f0f0f0fu) | (y & 0xf0f0f0f0u);
#endif
    return (f & 0x88888888u) * 0x00204081u;
}

// ---- bases: one little-endian word of BAM 4-bit SEQ = eight bases, high nibble first in every byte ----
// Returns the eight flags of each plane as the top byte (bit 24+i = base i): v = SEQ[i] is A, C, G or T (BAM codes
// 1, 2, 4, 8; any other code (=, N, IUPAC) counts nowhere, pileup.py:83-86), b0 = code has bit 1 or 3 (C, T; scrap
// outside v), b1 = code has bit 2 or 3 (G, T; scrap outside v).
struct Planes8 { u32 v, b0, b1; };

// `e` holds one flag per base at the low bit of its nibble (bits 4n, nothing else set), in BAM order: byte k carries
// base 2k in its HIGH nibble and base 2k + 1 in its low one. Returns the eight flags as the top byte (bit 24 + i = base i;
// lower bits are scrap). The two flags of a byte first meet in its bits 0, 1 (high nibble >> 4, low nibble << 1 - an
// add on the FMA pipe, the sets are disjoint), then one multiplication lines the four pairs up in the top byte. Working
// on the bytes as stored spares swapping the nibbles of every SEQ word first.
MG_HD u32 pack_nibble_flags_raw(u32 e) {
    const u32 y = ((e >> 4) + e * 2u) & 0x03030303u;    // byte k: bit 0 = base 2k, bit 1 = base 2k + 1
    return y * 0x01041040u;
}

MG_HD Planes8 seq_planes8_raw(u32 s) {                  // top byte = flags, lower bits scrap (AND with a top-byte mask)
    const u32 t1 = s >> 1, t2 = s >> 2, t3 = s >> 3;     // bit 4n of tK = bit K of nibble n
    const u32 k = 0x11111111u;
    const u32 one3 = (s ^ t1 ^ t2) & ~(s & t1 & t2);     // exactly one of the low three code bits
    const u32 none3 = ~(s | t1 | t2);
    Planes8 r;
    r.v = pack_nibble_flags_raw(((one3 & ~t3) | (none3 & t3)) & k);       // one-hot code
    r.b0 = pack_nibble_flags_raw((t1 | t3) & k);
    r.b1 = pack_nibble_flags_raw((t2 | t3) & k);
    return r;
}

// ---- the same planes by table lookup: PRMT (byte permute) as a 16-entry table over the BAM base code ----
// prmt.b32 in its generic mode picks, for each of the four selector nibbles of `sel` (low 16 bits), one byte of the pool
// {b, a}: the low three bits index the byte, the top bit asks for the byte's SIGN replicated over the result byte instead
// (0x00 / 0xff). With the pool below a selector nibble that is a BAM base code n turns into its plane flags
// (bit 0 = V, bit 1 = B0, bit 2 = B1): A = 1 -> pool[1] = V, C = 2 -> pool[2] = V|B0, G = 4 -> pool[4] = V|B1,
// T = 8 -> sign of pool[0] = 0xff (all three flags; bits 3-7 scrap); '=' = 0 -> pool[0] = 0x80 (no flag), every other
// code -> 0 (pool[3], [5..7] = 0; codes 9..15 replicate the clear sign of pool[1..7]). Two PRMTs turn the eight bases of a
// SEQ word into eight flag bytes - no per-nibble logic at all.
MG_HD u32 prmt_generic(u32 a, u32 b, u32 sel) {
#if defined(__CUDA_ARCH__)
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; i++) {
        const u32 c = (sel >> (4 * i)) & 15u;
        u32 by = (u32)((ab >> (8 * (c & 7u))) & 0xffu);
        if (c & 8u) by = (by & 0x80u) ? 0xffu : 0u;
        r |= by << (8 * i);
    }
    return r;
#endif
}
constexpr u32 kBaseLutLo = 0x00030180u, kBaseLutHi = 0x00000005u;

// Flags of the eight bases of one SEQ word, two bases per byte: byte i of the result carries base [1, 0, 3, 2][i] in its
// bits 0-2 (V, B0, B1) and base [5, 4, 7, 6][i] in its bits 4-6 (selector nibble 0 is the LOW nibble of SEQ byte 0 =
// base 1); bits 3 and 7 are scrap.
MG_HD u32 seq_flags8(u32 s) {
    const u32 lo = prmt_generic(kBaseLutLo, kBaseLutHi, s), hi = prmt_generic(kBaseLutLo, kBaseLutHi, s >> 16);
#if defined(__CUDA_ARCH__)
    u32 x;                                               // (lo & m) | ((hi << 4) & ~m) as ONE three-input operation
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(x) : "r"(lo), "r"(hi << 4), "r"(0x0f0f0f0fu));
    return x;
#else
    return (lo & 0x0f0f0f0fu) | ((hi << 4) & 0xf0f0f0f0u);
#endif
}
// One plane of seq_flags8 (j = 0, 1, 2) as the top byte, bit 24 + i = base i: the eight flags sit at bits 8i + j (base
// [1, 0, 3, 2][i]) and 8i + 4 + j; one multiplication moves each byte's pair to its place (shifts 25, 16, 11, 2 for
// i = 0..3, less j) and no two partial products below bit 32 meet, so nothing carries. Lower bits are scrap.
template <int J>
MG_HD u32 gather_plane8(u32 x) { return (x & (0x11111111u << J)) * (0x02010804u >> J); }

MG_HD Planes8 seq_planes8_lut(u32 s) {                  // same contract as seq_planes8_raw (b0, b1 exact, not only under v)
    const u32 x = seq_flags8(s);
    Planes8 r;
    r.v = gather_plane8<0>(x); r.b0 = gather_plane8<1>(x); r.b1 = gather_plane8<2>(x);
    return r;
}

MG_HD Planes8 seq_planes8_top(u32 s) {
    Planes8 r = seq_planes8_raw(s);
    r.v &= 0xff000000u; r.b0 &= 0xff000000u; r.b1 &= 0xff000000u;
    return r;
}

// per-base masks back from the planes (b0, b1 inside v): out = A, C, G, T
MG_HD void planes_to_bases(u32 v, u32 b0, u32 b1, u32 (&out)[4]) {
    out[0] = v & ~(b0 | b1); out[1] = b0 & ~b1; out[2] = b1 & ~b0; out[3] = b0 & b1;
}

// byte G of m <- top byte of v
template <int G>
MG_HD u32 insert_top_byte(u32 m, u32 v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(m, v, G == 0 ? 0x3217 : G == 1 ? 0x3270 : G == 2 ? 0x3710 : 0x7210);
#else
    return (m & ~(0xffu << (8 * G))) | ((v >> 24) << (8 * G));
#endif
}

// ---- 32x32 bit transpose across a warp: one butterfly stage ----
// Lane r holds row r; after the five stages (j = 16, 8, 4, 2, 1) lane p holds column p (bit r = row r's bit p).
// `mine` is the word of this lane, `other` the word of lane ^ j. keep / amt are the lane's constants of the stage.
MG_HD u32 transpose_keep(int lane, int j) {
    const u32 m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
    return (lane & j) ? ~m : m;
}
MG_HD u32 transpose_amt(int lane, int j) { return (lane & j) ? 32 - j : j; }
MG_HD u32 transpose_stage(u32 mine, u32 other, u32 keep, u32 amt) {
    const u32 moved = rotl32(other, amt);                  // wrapped-around bits fall outside ~keep
    return (mine & keep) | (moved & ~keep);
}
// The stages j = 16 and j = 8 move whole bytes: one byte permute of (mine, other) each instead of rotate + select.
MG_HD u32 byte_perm(u32 a, u32 b, u32 sel) {            // PRMT, selector nibbles 0..7 only
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (u32)((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}
MG_HD u32 transpose_sel(int lane, int j) {              // j = 16 or 8
    return j == 16 ? ((lane & 16) ? 0x3276u : 0x5410u) : ((lane & 8) ? 0x3715u : 0x6240u);
}
MG_HD u32 transpose_stage_bytes(u32 mine, u32 other, u32 sel) { return byte_perm(mine, other, sel); }

// ---- one round of phase B: four bit matrices (V, B0, B1, Tn5 site; row = read = lane) through as few transposes as
// possible. When the candidate reads sit in lanes 0..7 the four matrices share ONE transpose: lanes 8..15 carry the B0
// rows of lanes 0..7, lanes 16..23 their B1 rows, lanes 24..31 their Tn5 rows (s0, s1, s5 = those rows fetched from lane
// & 7); after the transpose every byte of a lane's word is the column of one matrix. With candidates in lanes 0..15 two
// transposes carry (V | B0) and (B1 | Tn5) in their half words.
MG_HD u32 quarter_row(int lane, u32 v, u32 s0, u32 s1, u32 s5) { return lane < 16 ? (lane < 8 ? v : s0) : (lane < 24 ? s1 : s5); }
MG_HD void quarter_columns(u32 t, u32 &v, u32 &b0, u32 &b1, u32 &t5) {
    v = t & 0xffu; b0 = byte_perm(t, 0u, 0x4441); b1 = byte_perm(t, 0u, 0x4442); t5 = t >> 24;
}
MG_HD u32 half_row(int lane, u32 mine, u32 from_lane_xor_16) { return lane < 16 ? mine : from_lane_xor_16; }
MG_HD void half_columns(u32 ta, u32 tb, u32 &v, u32 &b0, u32 &b1, u32 &t5) {
    v = ta & 0xffffu; b0 = ta >> 16; b1 = tb & 0xffffu; t5 = tb >> 16;
}
// columns (bit r = read r) -> the ten counters of a position: A, C, G, T x (forward, reverse), Tn5 forward / reverse.
// `rev` = reads on the reverse strand (pileup.py:88); b0, b1 lie inside v.
MG_HD u32 popc32(u32 x) {
#if defined(__CUDA_ARCH__)
    return (u32)__popc(x);
#else
    return (u32)__builtin_popcount(x);
#endif
}
MG_HD void count_columns(u32 v, u32 b0, u32 b1, u32 t5, u32 rev, u32 (&cnt)[10]) {
    const u32 any = b0 | b1;
    cnt[0] += popc32(v & ~any & ~rev); cnt[1] += popc32(v & ~any & rev);       // A: valid, code 0
    cnt[2] += popc32(b0 & ~b1 & ~rev); cnt[3] += popc32(b0 & ~b1 & rev);       // C
    cnt[4] += popc32(b1 & ~b0 & ~rev); cnt[5] += popc32(b1 & ~b0 & rev);       // G
    cnt[6] += popc32(b0 & b1 & ~rev); cnt[7] += popc32(b0 & b1 & rev);         // T
    cnt[8] += popc32(t5 & ~rev); cnt[9] += popc32(t5 & rev);
}

}  // namespace mgatk

// ---------------------------------------------------------------------------------------------
// Planning helper (pileup.cuh: k_plan_units). `S` gives `int pos(int i)`, the start of slot i (sorted).
// ---------------------------------------------------------------------------------------------
namespace mgatk {

// First index in [lo, hi) whose start is >= key (hi when there is none); the slots are sorted by start. The answer is
// expected a little below `guess` (a tile border lies less than a chunk, the first read of a tile less than a halo below
// the read the border was taken from): the probes guess - 1, - 2, - 4 ... - 512 are loaded at once, the answer's bracket is
// read off them and closed by bisection - four or five dependent loads instead of the thirteen of a bisection over a
// whole cell. Exact for any guess.
template <class S>
MG_HD int lower_bound_near(const S &sl, int lo, int hi, int guess, int key) {
    if (lo >= hi) return lo;
    const int g = guess < lo ? lo : guess > hi - 1 ? hi - 1 : guess;
    int a = g + 1, b = hi;                                   // the answer lies in [a, b]
    if (sl.pos(g) >= key) {
        int p[10];
#pragma unroll
        for (int k = 0; k < 10; k++) { const int j = g - (1 << k); p[k] = j >= lo ? sl.pos(j) : 0; }
        a = lo; b = g;
        bool closed = false;
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const int j = g - (1 << k);
            if (!closed) {
                if (j < lo) closed = true;                   // out of range: [lo, b]
                else if (p[k] < key) { a = j + 1; closed = true; }
                else b = j;
            }
        }
    }
    while (a < b) { const int mid = (a + b) >> 1; if (sl.pos(mid) < key) a = mid + 1; else b = mid; }
    return a;
}

}  // namespace mgatk

// ---------------------------------------------------------------------------------------------
// Query masks of one read. `M` gives word access to the staging memory holding the read's
// cigar|seq|qual blob (shared memory on the device, a byte array in the host test):
//     u32  ld32(u32 addr)                    4-byte aligned load
//     void st128(u32 addr, u32 a, u32 c, u32 g, u32 t)   16-byte aligned store
//     void ld128(u32 addr, u32 (&v)[4])      16-byte aligned load
// Layout written at `out`: ceil(L/32) groups of four words (V, B0, B1 planes of query bases 32w..32w+31, B0 and B1
// cut to V; the fourth word is zero).
// Loads may run up to 40 bytes past the end of QUAL; whatever they return is masked by the window.
// ---------------------------------------------------------------------------------------------
namespace mgatk {

// planes of query bases 32w .. 32w+31 of one read, in registers (B0 and B1 cut to V)
template <class M>
MG_HD void query_mask_group(const M &mem, u32 seq_addr /*4-aligned*/, int L, int w, int q_lo, int q_hi, QualGe qg, u32 &oV, u32 &o0, u32 &o1) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1) + 32u * (u32)w;
    const u32 qsh = (qual_addr & 3u) * 8u;
    const u32 qa = qual_addr & ~3u;
    u32 mV = 0, m0 = 0, m1 = 0;
    const int lo = q_lo - 32 * w, hi = (q_hi < L ? q_hi : L) - 32 * w;   // window inside this group of 32 bases
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (8 * g >= hi) break;                              // eight bases at a time; those outside the window are skipped
        if (8 * g + 8 <= lo) continue;
        const u32 s = mem.ld32(seq_addr + 16u * w + 4u * g);
        const u32 w0 = mem.ld32(qa + 8u * g), w1 = mem.ld32(qa + 8u * g + 4u), w2 = mem.ld32(qa + 8u * g + 8u);
        const u32 ok = qual_ok8_top(funnel_r(w0, w1, qsh), funnel_r(w1, w2, qsh), qg);
        const Planes8 e = seq_planes8_raw(s);
        if (g == 0) { mV = insert_top_byte<0>(mV, e.v & ok); m0 = insert_top_byte<0>(m0, e.b0); m1 = insert_top_byte<0>(m1, e.b1); }
        else if (g == 1) { mV = insert_top_byte<1>(mV, e.v & ok); m0 = insert_top_byte<1>(m0, e.b0); m1 = insert_top_byte<1>(m1, e.b1); }
        else if (g == 2) { mV = insert_top_byte<2>(mV, e.v & ok); m0 = insert_top_byte<2>(m0, e.b0); m1 = insert_top_byte<2>(m1, e.b1); }
        else { mV = insert_top_byte<3>(mV, e.v & ok); m0 = insert_top_byte<3>(m0, e.b0); m1 = insert_top_byte<3>(m1, e.b1); }
    }
    mV &= bit_range(lo, (q_hi < L ? q_hi : L) - 32 * w);      // pileup.py:67-78 (also cuts bases >= L)
    oV = mV; o0 = m0 & mV; o1 = m1 & mV;
}

// query_mask_group without its per-subgroup conditions: the four words of SEQ and the nine of QUAL of the 32 bases are
// always loaded and turned into flags, the window is one mask at the end (straight-line code; loads reach up to 44 bytes
// past the end of QUAL). Same result.
template <class M>
MG_HD void query_mask_group_straight(const M &mem, u32 seq_addr /*4-aligned*/, int L, int w, int q_lo, int q_hi, QualGe qg, u32 &oV, u32 &o0, u32 &o1) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1) + 32u * (u32)w;
    const u32 qsh = (qual_addr & 3u) * 8u, qa = qual_addr & ~3u;
    u32 qw[9];
#pragma unroll
    for (int k = 0; k < 9; k++) qw[k] = mem.ld32(qa + 4u * (u32)k);
    u32 mV = 0, m0 = 0, m1 = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 s = mem.ld32(seq_addr + 16u * (u32)w + 4u * (u32)g);
        const u32 ok = qual_ok8_top(funnel_r(qw[2 * g], qw[2 * g + 1], qsh), funnel_r(qw[2 * g + 1], qw[2 * g + 2], qsh), qg);
        const Planes8 e = seq_planes8_raw(s);
        if (g == 0) { mV = insert_top_byte<0>(mV, e.v & ok); m0 = insert_top_byte<0>(m0, e.b0); m1 = insert_top_byte<0>(m1, e.b1); }
        else if (g == 1) { mV = insert_top_byte<1>(mV, e.v & ok); m0 = insert_top_byte<1>(m0, e.b0); m1 = insert_top_byte<1>(m1, e.b1); }
        else if (g == 2) { mV = insert_top_byte<2>(mV, e.v & ok); m0 = insert_top_byte<2>(m0, e.b0); m1 = insert_top_byte<2>(m1, e.b1); }
        else { mV = insert_top_byte<3>(mV, e.v & ok); m0 = insert_top_byte<3>(m0, e.b0); m1 = insert_top_byte<3>(m1, e.b1); }
    }
    mV &= bit_range(q_lo - 32 * w, (q_hi < L ? q_hi : L) - 32 * w);      // pileup.py:67-78 (also cuts bases >= L)
    oV = mV; o0 = m0 & mV; o1 = m1 & mV;
}

// query_mask_group_straight with the table-lookup planes (seq_planes8_lut) and the raw quality flags. Same result.
template <bool kAnd = false, class M>
MG_HD void query_mask_group_lut(const M &mem, u32 seq_addr /*4-aligned*/, int L, int w, int q_lo, int q_hi, QualGe qg, u32 &oV, u32 &o0, u32 &o1) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1) + 32u * (u32)w;
    const u32 qsh = (qual_addr & 3u) * 8u, qa = qual_addr & ~3u;
    u32 qw[9];
#pragma unroll
    for (int k = 0; k < 9; k++) qw[k] = mem.ld32(qa + 4u * (u32)k);
    u32 mV = 0, m0 = 0, m1 = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 s = mem.ld32(seq_addr + 16u * (u32)w + 4u * (u32)g);
        const u32 ok = qual_ok8_raw<kAnd>(funnel_r(qw[2 * g], qw[2 * g + 1], qsh), funnel_r(qw[2 * g + 1], qw[2 * g + 2], qsh), qg);
        const Planes8 e = seq_planes8_lut(s);
        if (g == 0) { mV = insert_top_byte<0>(mV, e.v & ok); m0 = insert_top_byte<0>(m0, e.b0); m1 = insert_top_byte<0>(m1, e.b1); }
        else if (g == 1) { mV = insert_top_byte<1>(mV, e.v & ok); m0 = insert_top_byte<1>(m0, e.b0); m1 = insert_top_byte<1>(m1, e.b1); }
        else if (g == 2) { mV = insert_top_byte<2>(mV, e.v & ok); m0 = insert_top_byte<2>(m0, e.b0); m1 = insert_top_byte<2>(m1, e.b1); }
        else { mV = insert_top_byte<3>(mV, e.v & ok); m0 = insert_top_byte<3>(m0, e.b0); m1 = insert_top_byte<3>(m1, e.b1); }
    }
    mV &= bit_range(q_lo - 32 * w, (q_hi < L ? q_hi : L) - 32 * w);      // pileup.py:67-78 (also cuts bases >= L)
    oV = mV; o0 = m0 & mV; o1 = m1 & mV;
}

// All query planes of a read of at most 56 bases at once (the compact slot, common.cuh): every word of SEQ and QUAL
// is loaded once, groups of eight bases outside the window are skipped as a whole, the window itself is one mask at the
// end. g[w] = { V, B0, B1 } of query bases 32w .. 32w+31.
template <class M>
MG_HD void query_planes56(const M &mem, u32 seq_addr /*4-aligned*/, int L, int q_lo, int q_hi, QualGe qg, u32 (&g)[2][3]) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1);
    const u32 qsh = (qual_addr & 3u) * 8u, qa = qual_addr & ~3u;
    const int hi = q_hi < L ? q_hi : L;                              // window [q_lo, hi), hi <= 56
    const int lo = q_lo > 0 ? q_lo : 0;
    const int g_lo = lo >> 3;
    int g_n = ((hi + 7) >> 3) - g_lo;                                // groups g_lo .. g_lo + g_n - 1 meet the window
    if (g_n < 0 || hi <= lo) g_n = 0;
    u32 qw[15];
#pragma unroll
    for (int k = 0; k < 15; k++) qw[k] = ((u32)(k - 2 * g_lo) < (u32)(2 * g_n + (g_n > 0))) ? mem.ld32(qa + 4u * (u32)k) : 0u;
    u32 mV[2] = {0u, 0u}, m0[2] = {0u, 0u}, m1[2] = {0u, 0u};
#pragma unroll
    for (int gi = 0; gi < 7; gi++) {
        if ((u32)(gi - g_lo) < (u32)g_n) {
            const u32 s = mem.ld32(seq_addr + 4u * (u32)gi);
            const u32 ok = qual_ok8_top(funnel_r(qw[2 * gi], qw[2 * gi + 1], qsh), funnel_r(qw[2 * gi + 1], qw[2 * gi + 2], qsh), qg);
            const Planes8 e = seq_planes8_raw(s);
            const int w = gi >> 2, b = gi & 3;
            if (b == 0) { mV[w] = insert_top_byte<0>(mV[w], e.v & ok); m0[w] = insert_top_byte<0>(m0[w], e.b0); m1[w] = insert_top_byte<0>(m1[w], e.b1); }
            else if (b == 1) { mV[w] = insert_top_byte<1>(mV[w], e.v & ok); m0[w] = insert_top_byte<1>(m0[w], e.b0); m1[w] = insert_top_byte<1>(m1[w], e.b1); }
            else if (b == 2) { mV[w] = insert_top_byte<2>(mV[w], e.v & ok); m0[w] = insert_top_byte<2>(m0[w], e.b0); m1[w] = insert_top_byte<2>(m1[w], e.b1); }
            else { mV[w] = insert_top_byte<3>(mV[w], e.v & ok); m0[w] = insert_top_byte<3>(m0[w], e.b0); m1[w] = insert_top_byte<3>(m1[w], e.b1); }
        }
    }
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const u32 v = mV[w] & bit_range(lo - 32 * w, hi - 32 * w);   // pileup.py:67-78 (also cuts bases >= L)
        g[w][0] = v; g[w][1] = m0[w] & v; g[w][2] = m1[w] & v;
    }
}

// The same without any per-group condition: all seven groups of eight bases are computed whatever the window (what lies
// outside it is cut by the mask at the end), so the builder is straight-line code. Loads reach up to 64 bytes past the
// start of QUAL + 56, i.e. at most 48 bytes beyond a record's blob: the caller's staging memory must allow that.
// kGroups = groups of eight bases to compute: 7 covers any compact read; fewer when the caller knows that no window of the
// batch reaches further (window end <= 8 kGroups for every read).
template <int kGroups, class M>
MG_HD void query_planes56_straight(const M &mem, u32 seq_addr /*4-aligned*/, int L, int q_lo, int q_hi, QualGe qg, u32 (&g)[2][3]) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1);
    const u32 qsh = (qual_addr & 3u) * 8u, qa = qual_addr & ~3u;
    const int hi = q_hi < L ? q_hi : L;
    const int lo = q_lo > 0 ? q_lo : 0;
    u32 qw[2 * kGroups + 1];
#pragma unroll
    for (int k = 0; k < 2 * kGroups + 1; k++) qw[k] = mem.ld32(qa + 4u * (u32)k);
    u32 mV[2] = {0u, 0u}, m0[2] = {0u, 0u}, m1[2] = {0u, 0u};
#pragma unroll
    for (int gi = 0; gi < kGroups; gi++) {
        const u32 s = mem.ld32(seq_addr + 4u * (u32)gi);
        const u32 ok = qual_ok8_top(funnel_r(qw[2 * gi], qw[2 * gi + 1], qsh), funnel_r(qw[2 * gi + 1], qw[2 * gi + 2], qsh), qg);
        const Planes8 e = seq_planes8_raw(s);
        const int w = gi >> 2, b = gi & 3;
        if (b == 0) { mV[w] = insert_top_byte<0>(mV[w], e.v & ok); m0[w] = insert_top_byte<0>(m0[w], e.b0); m1[w] = insert_top_byte<0>(m1[w], e.b1); }
        else if (b == 1) { mV[w] = insert_top_byte<1>(mV[w], e.v & ok); m0[w] = insert_top_byte<1>(m0[w], e.b0); m1[w] = insert_top_byte<1>(m1[w], e.b1); }
        else if (b == 2) { mV[w] = insert_top_byte<2>(mV[w], e.v & ok); m0[w] = insert_top_byte<2>(m0[w], e.b0); m1[w] = insert_top_byte<2>(m1[w], e.b1); }
        else { mV[w] = insert_top_byte<3>(mV[w], e.v & ok); m0[w] = insert_top_byte<3>(m0[w], e.b0); m1[w] = insert_top_byte<3>(m1[w], e.b1); }
    }
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const u32 v = hi > lo ? mV[w] & bit_range(lo - 32 * w, hi - 32 * w) : 0u;
        g[w][0] = v; g[w][1] = m0[w] & v; g[w][2] = m1[w] & v;
    }
}

// query_planes56_straight with the table-lookup planes (seq_planes8_lut) and the quality mask applied once per eight
// bases: about 40 % fewer integer instructions per read, same result.
template <int kGroups, bool kAnd = false, class M>
MG_HD void query_planes56_lut(const M &mem, u32 seq_addr /*4-aligned*/, int L, int q_lo, int q_hi, QualGe qg, u32 (&g)[2][3]) {
    const u32 qual_addr = seq_addr + (u32)((L + 1) >> 1);
    const u32 qsh = (qual_addr & 3u) * 8u, qa = qual_addr & ~3u;
    const int hi = q_hi < L ? q_hi : L;
    const int lo = q_lo > 0 ? q_lo : 0;
    u32 qw[2 * kGroups + 1];
#pragma unroll
    for (int k = 0; k < 2 * kGroups + 1; k++) qw[k] = mem.ld32(qa + 4u * (u32)k);
    u32 mV[2] = {0u, 0u}, m0[2] = {0u, 0u}, m1[2] = {0u, 0u};
#pragma unroll
    for (int gi = 0; gi < kGroups; gi++) {
        const u32 s = mem.ld32(seq_addr + 4u * (u32)gi);
        const u32 ok = qual_ok8_raw<kAnd>(funnel_r(qw[2 * gi], qw[2 * gi + 1], qsh), funnel_r(qw[2 * gi + 1], qw[2 * gi + 2], qsh), qg);
        const Planes8 e = seq_planes8_lut(s);
        const int w = gi >> 2, b = gi & 3;
        if (b == 0) { mV[w] = insert_top_byte<0>(mV[w], e.v & ok); m0[w] = insert_top_byte<0>(m0[w], e.b0); m1[w] = insert_top_byte<0>(m1[w], e.b1); }
        else if (b == 1) { mV[w] = insert_top_byte<1>(mV[w], e.v & ok); m0[w] = insert_top_byte<1>(m0[w], e.b0); m1[w] = insert_top_byte<1>(m1[w], e.b1); }
        else if (b == 2) { mV[w] = insert_top_byte<2>(mV[w], e.v & ok); m0[w] = insert_top_byte<2>(m0[w], e.b0); m1[w] = insert_top_byte<2>(m1[w], e.b1); }
        else { mV[w] = insert_top_byte<3>(mV[w], e.v & ok); m0[w] = insert_top_byte<3>(m0[w], e.b0); m1[w] = insert_top_byte<3>(m1[w], e.b1); }
    }
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const u32 v = hi > lo ? mV[w] & bit_range(lo - 32 * w, hi - 32 * w) : 0u;
        g[w][0] = v; g[w][1] = m0[w] & v; g[w][2] = m1[w] & v;
    }
}

// the same -> four words at out + 16w
template <class M>
MG_HD void build_query_mask_group(const M &mem, u32 seq_addr /*4-aligned*/, u32 out /*16-aligned*/, int L, int w, int q_lo, int q_hi, QualGe qg) {
    u32 mV, m0, m1;
    query_mask_group(mem, seq_addr, L, w, q_lo, q_hi, qg, mV, m0, m1);
    mem.st128(out + 16u * w, mV, m0, m1, 0u);
}

template <class M>
MG_HD void build_query_masks(const M &mem, u32 seq_addr /*4-aligned*/, u32 out /*16-aligned*/, int L, int q_lo, int q_hi, QualGe qg) {
    const int nq = (L + 31) >> 5;
    for (int w = 0; w < nq; w++) build_query_mask_group(mem, seq_addr, out, L, w, q_lo, q_hi, qg);
}

// 32-bit windows of the three planes (V, B0, B1) starting at query bit qb (may be negative or beyond the read)
template <class M>
MG_HD void query_window(const M &mem, u32 masks /*16-aligned*/, int nq, int qb, u32 (&out)[3]) {
    const int w0 = qb >> 5;                                // floor
    const u32 sh = (u32)qb & 31u;
    u32 lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    if (w0 >= 0 && w0 < nq) mem.ld128(masks + 16u * (u32)w0, lo);
    if (w0 + 1 >= 0 && w0 + 1 < nq) mem.ld128(masks + 16u * (u32)(w0 + 1), hi);
#pragma unroll
    for (int x = 0; x < 3; x++) out[x] = funnel_r(lo[x], hi[x], sh);
}

// ---------------------------------------------------------------------------------------------
// Query coordinates -> reference coordinates. The CIGAR walk of pileup.py:52-95 moves runs of query
// bases to reference positions: an aligned block (M, =, X) of length n places query bases
// [qp, qp + n) on reference offsets [ref, ref + n) (relative to reference_start) and advances both;
// D / N advance the reference only, S the query only; I, H and P advance NOTHING (the reference's
// quirk, SURVEY Q3: the bases after an insertion are read from the inserted query positions).
// ref_group() builds one group of 32 reference offsets [32k, 32k + 32) of the three planes from the
// query planes: per overlapping block, a 32-bit window of the query planes cut to the block.
//   M: ld32(addr) of the cigar words;  Q: window(qb, out[3]) = query-plane bits [qb, qb + 32)
// After this, stages 4-6 never see a CIGAR: a read is three bit planes anchored at its start.
// ---------------------------------------------------------------------------------------------
constexpr int kOpCap = 1 << 20;     // cigar lengths above this cannot be valid for a short-read batch
constexpr int kRefCap = 1 << 28;    // running offsets saturate here (65535 ops x 2^20 would overflow an int)

MG_HD bool cigar_op_aligned(int op) { return op == 0 || op == 7 || op == 8; }      // pileup.py:56
MG_HD bool cigar_op_ref_only(int op) { return op == 2 || op == 3; }                 // pileup.py:92-93

// reference span of a read (sum of M/=/X/D/N lengths), saturating
template <class M>
MG_HD int cigar_ref_span(const M &mem, u32 cig_addr, int ncig) {
    int span = 0;
    for (int ci = 0; ci < ncig; ci++) {
        const u32 w = mem.ld32(cig_addr + 4u * (u32)ci);
        const int op = (int)(w & 15u);
        if (cigar_op_aligned(op) || cigar_op_ref_only(op)) { span += (int)(w >> 4) < kOpCap ? (int)(w >> 4) : kOpCap; if (span > kRefCap) span = kRefCap; }
    }
    return span;
}

template <class M, class Q>
MG_HD void ref_group(const M &mem, u32 cig_addr, int ncig, const Q &qplanes, int k, u32 (&out)[3]) {
    out[0] = 0u; out[1] = 0u; out[2] = 0u;
    const int lo = 32 * k, hi = lo + 32;
    int ref = 0, qp = 0;
    for (int ci = 0; ci < ncig; ci++) {
        const u32 w = mem.ld32(cig_addr + 4u * (u32)ci);
        const int op = (int)(w & 15u);
        const int n = (int)(w >> 4) < kOpCap ? (int)(w >> 4) : kOpCap;
        if (cigar_op_aligned(op)) {
            if (ref < hi && ref + n > lo) {
                u32 wv[3];
                qplanes.window(lo - ref + qp, wv);             // query bit of reference offset `lo`
                const u32 rm = bit_range(ref - lo, ref + n - lo);
                out[0] |= wv[0] & rm; out[1] |= wv[1] & rm; out[2] |= wv[2] & rm;
            }
            ref += n; qp += n;                                 // pileup.py:90-91
            if (ref > kRefCap) ref = kRefCap;
            if (qp > kRefCap) qp = kRefCap;
        } else if (cigar_op_ref_only(op)) {
            ref += n;
            if (ref > kRefCap) ref = kRefCap;
        } else if (op == 4) {                                  // pileup.py:94-95
            qp += n;
            if (qp > kRefCap) qp = kRefCap;
        }
        if (ref >= hi) break;                                  // blocks only move right
    }
}

// query planes of a read of at most 64 bases held in three 64-bit words
struct QueryPlanes64 {
    unsigned long long v, b0, b1;
    MG_HD static u32 win(unsigned long long x, int qb) {
        if (qb >= 64 || qb <= -32) return 0u;
        return qb >= 0 ? (u32)(x >> qb) : (u32)(x << (-qb));
    }
    MG_HD void window(int qb, u32 (&out)[3]) const { out[0] = win(v, qb); out[1] = win(b0, qb); out[2] = win(b1, qb); }
};

// query planes stored as groups of four words (V, B0, B1, 0 per 32 bases) behind an accessor with ld128
template <class M>
struct QueryPlanesMem {
    const M &mem; u32 addr; int nq;
    MG_HD void window(int qb, u32 (&out)[3]) const {
        if (qb >= 32 * nq || qb <= -32) { out[0] = 0u; out[1] = 0u; out[2] = 0u; return; }
        query_window(mem, addr, nq, qb, out);
    }
};

}  // namespace mgatk
