// mgatk2_b200 — DEFLATE decoder for BGZF blocks (RFC 1951), used by the BAM ingest (csrc/bamio.cpp).
//
// A BGZF block is a complete raw DEFLATE stream of at most 64 KB with its CRC-32 and length in the trailer, so the
// decoder can be specialised: whole input and output in memory, 64-bit bit buffer refilled once per symbol pair,
// two-level lookup tables (11 root bits for literals / lengths, 8 for distances) whose entries carry base value, extra
// bit count and code length, word-wise match copies. Anything irregular (a code that does not decode, output that does
// not fit, a stream that ends early) makes it return false and the caller falls back to zlib - the decoder is an
// accelerator, not the authority; the caller also checks the CRC-32 of what it got.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace mgatk_inflate {

constexpr int kLitRoot = 11, kDistRoot = 8, kMaxLen = 15;
constexpr uint32_t kLiteral = 1u << 12, kEndOfBlock = 2u << 12, kSubtable = 4u << 12, kInvalid = 8u << 12;
constexpr uint32_t kPair = 1u << 8;        // with kLiteral: two literals (value = first | second << 8)
// entry: bits 0..7 code length (bits to drop), 8..11 extra bits (or subtable bits), 12..15 flags, 16..31 value

struct Tables {
    uint32_t lit[(1 << kLitRoot) + 288 * 16];
    uint32_t dist[(1 << kDistRoot) + 32 * 128];
};

struct Rev8 { uint8_t r[256]; Rev8() { for (int i = 0; i < 256; i++) { int v = i, x = 0; for (int b = 0; b < 8; b++) { x = (x << 1) | (v & 1); v >>= 1; } r[i] = (uint8_t)x; } } };
inline uint32_t reverse_bits(uint32_t v, int n) {          // the low n <= 16 bits of v, reversed
    static const Rev8 t;
    return (((uint32_t)t.r[v & 0xff] << 8) | t.r[(v >> 8) & 0xff]) >> (16 - n);
}

// canonical Huffman code of `n` symbols with lengths lens[] -> lookup table; payload[sym] = value << 16 | extra << 8 | flags
inline bool build_table(const uint8_t *lens, int n, const uint32_t *payload, uint32_t *table, int root, size_t capacity,
                        bool pair_literals = false) {
    int count[kMaxLen + 1] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    if (count[0] == n) {                                   // no codes at all: every lookup is invalid
        for (int i = 0; i < (1 << root); i++) table[i] = kInvalid | 1u;
        return true;
    }
    // over-subscribed codes are refused; an incomplete code is legal only with a single code of one bit (RFC 1951 3.2.7)
    int left = 1;
    for (int len = 1; len <= kMaxLen; len++) { left = (left << 1) - count[len]; if (left < 0) return false; }
    int used = 0;
    for (int len = 1; len <= kMaxLen; len++) used += count[len];
    if (left > 0 && !(used == 1 && count[1] == 1)) return false;
    uint32_t next_code[kMaxLen + 2];
    {   // code of the first symbol of each length (RFC 1951 3.2.2)
        uint32_t c = 0;
        int prev = 0;
        for (int len = 1; len <= kMaxLen; len++) { c = (c + (uint32_t)prev) << 1; next_code[len] = c; prev = count[len]; }
    }
    const uint32_t root_mask = (1u << root) - 1u;
    if (left > 0) for (int i = 0; i < (1 << root); i++) table[i] = kInvalid | 1u;     // a complete code fills every entry
    // longest code below every root prefix that needs a second level
    uint8_t sub_bits[1 << kLitRoot];
    int long_codes = 0;
    for (int len = root + 1; len <= kMaxLen; len++) long_codes += count[len];
    if (long_codes) memset(sub_bits, 0, (size_t)1 << root);
    if (long_codes) {
        uint32_t nc[kMaxLen + 2];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n; s++) {
            const int len = lens[s];
            if (!len) continue;
            const uint32_t c = nc[len]++;
            if (len > root) {
                const uint32_t prefix = reverse_bits(c, len) & root_mask;
                if ((int)sub_bits[prefix] < len - root) sub_bits[prefix] = (uint8_t)(len - root);
            }
        }
    }
    size_t next_sub = (size_t)1 << root;
    for (uint32_t p = 0; long_codes && p <= root_mask; p++) {
        if (!sub_bits[p]) continue;
        const size_t size = (size_t)1 << sub_bits[p];
        if (next_sub + size > capacity) return false;
        table[p] = kSubtable | ((uint32_t)next_sub << 16) | ((uint32_t)sub_bits[p] << 8) | (uint32_t)root;
        for (size_t i = 0; i < size; i++) table[next_sub + i] = kInvalid | 1u;
        next_sub += size;
        if (next_sub > 0xffff) return false;
    }
    for (int s = 0; s < n; s++) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t rev = reverse_bits(next_code[len]++, len);
        if (len <= root) {
            const uint32_t e = payload[s] | (uint32_t)len;
            for (uint32_t i = rev; i <= root_mask; i += 1u << len) table[i] = e;
        } else {
            const uint32_t head = table[rev & root_mask];
            const size_t start = head >> 16;
            const int sb = (int)((head >> 8) & 15);
            const uint32_t e = payload[s] | (uint32_t)(len - root);
            for (uint32_t i = rev >> root; i < (1u << sb); i += 1u << (len - root)) table[start + i] = e;
        }
    }
    if (pair_literals) {
        // two short literal codes that fit the root bits together decode with one lookup: entry = both bytes, summed
        // length, kPair. Walking down keeps table[i >> len] (a smaller index) in its single form while it is read.
        for (uint32_t i = root_mask;; i--) {
            const uint32_t e = table[i];
            const int l1 = (int)(e & 0xff);
            if ((e & kLiteral) && l1 < root) {
                const uint32_t e2 = table[i >> l1];
                const int l2 = (int)(e2 & 0xff);
                if ((e2 & kLiteral) && !(e2 & kPair) && l1 + l2 <= root)
                    table[i] = kLiteral | kPair | (uint32_t)(l1 + l2) | (e & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8);
            }
            if (i == 0) break;
        }
    }
    return true;
}

struct Payloads {
    uint32_t lit[288], dist[32];
    Payloads() {
        static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (int s = 0; s < 256; s++) lit[s] = ((uint32_t)s << 16) | kLiteral;
        lit[256] = kEndOfBlock;
        for (int s = 257; s < 286; s++) lit[s] = ((uint32_t)len_base[s - 257] << 16) | ((uint32_t)len_extra[s - 257] << 8);
        lit[286] = lit[287] = kInvalid;
        for (int s = 0; s < 30; s++) dist[s] = ((uint32_t)dist_base[s] << 16) | ((uint32_t)dist_extra[s] << 8);
        dist[30] = dist[31] = kInvalid;
    }
};

inline const Payloads &payloads() { static const Payloads p; return p; }

struct BitReader {
    const uint8_t *in, *in_end;
    uint64_t buf = 0;
    int cnt = 0;
    // at least 56 bits when eight bytes can be loaded; otherwise byte by byte (zeros past the end)
    inline void refill() {
        if (in_end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < in_end) { buf |= (uint64_t)*in++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & (((uint64_t)1 << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
};

// raw DEFLATE stream in[0..in_len) -> exactly out_len bytes at out. false = not decoded (fall back to zlib).
inline bool inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, Tables &t) {
    BitReader br{in, in + in_len};
    uint8_t *o = out, *const o_end = out + out_len;
    const Payloads &pl = payloads();
    bool last = false;
    while (!last) {
        br.refill();
        if (br.cnt < 3) return false;
        last = br.peek(1);
        const uint32_t type = (br.peek(3) >> 1);
        br.drop(3);
        if (type == 0) {                                       // stored
            br.drop(br.cnt & 7);
            // the bit buffer holds whole bytes now: hand them back
            br.in -= br.cnt >> 3; br.buf = 0; br.cnt = 0;
            if (br.in_end - br.in < 4) return false;
            const uint32_t len = br.in[0] | (br.in[1] << 8), nlen = br.in[2] | (br.in[3] << 8);
            if ((len ^ 0xffffu) != nlen) return false;
            br.in += 4;
            if ((size_t)(br.in_end - br.in) < len || (size_t)(o_end - o) < len) return false;
            memcpy(o, br.in, len);
            o += len; br.in += len;
            continue;
        }
        if (type == 3) return false;
        if (type == 1) {                                       // fixed code
            uint8_t lens[288 + 32];
            for (int i = 0; i < 144; i++) lens[i] = 8;
            for (int i = 144; i < 256; i++) lens[i] = 9;
            for (int i = 256; i < 280; i++) lens[i] = 7;
            for (int i = 280; i < 288; i++) lens[i] = 8;
            for (int i = 0; i < 32; i++) lens[288 + i] = 5;
            if (!build_table(lens, 288, pl.lit, t.lit, kLitRoot, sizeof(t.lit) / 4, true) ||
                !build_table(lens + 288, 32, pl.dist, t.dist, kDistRoot, sizeof(t.dist) / 4)) return false;
        } else {                                               // dynamic code
            br.refill();
            if (br.cnt < 14) return false;
            const int hlit = (int)br.peek(5) + 257; br.drop(5);
            const int hdist = (int)br.peek(5) + 1; br.drop(5);
            const int hclen = (int)br.peek(4) + 4; br.drop(4);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pre_lens[19] = {0};
            for (int i = 0; i < hclen; i++) {
                br.refill();
                if (br.cnt < 3) return false;
                pre_lens[order[i]] = (uint8_t)br.peek(3); br.drop(3);
            }
            uint32_t pre_payload[19], pre_table[1 << 7];
            for (int i = 0; i < 19; i++) pre_payload[i] = (uint32_t)i << 16;
            if (!build_table(pre_lens, 19, pre_payload, pre_table, 7, 1 << 7)) return false;
            uint8_t lens[288 + 32];
            memset(lens, 0, sizeof(lens));
            int i = 0;
            while (i < hlit + hdist) {
                br.refill();
                const uint32_t e = pre_table[br.peek(7)];
                if (e & kInvalid) return false;
                const int cl = (int)(e & 0xff);
                if (br.cnt < cl + 7) return false;
                br.drop(cl);
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)br.peek(2); br.drop(2); }
                else if (sym == 17) { rep = 3 + (int)br.peek(3); br.drop(3); }
                else { rep = 11 + (int)br.peek(7); br.drop(7); }
                if (i + rep > hlit + hdist) return false;
                while (rep--) lens[i++] = (uint8_t)val;
            }
            if (lens[256] == 0) return false;                  // no end-of-block code
            uint8_t dl[32];
            memset(dl, 0, sizeof(dl));
            memcpy(dl, lens + hlit, (size_t)hdist);
            memset(lens + hlit, 0, sizeof(lens) - (size_t)hlit);
            if (!build_table(lens, 288, pl.lit, t.lit, kLitRoot, sizeof(t.lit) / 4, true) ||
                !build_table(dl, 32, pl.dist, t.dist, kDistRoot, sizeof(t.dist) / 4)) return false;
        }
        // ---- symbols ----
        for (;;) {
            br.refill();
            uint32_t e = t.lit[br.peek(kLitRoot)];
            if (e & kSubtable) {
                br.drop(kLitRoot);
                e = t.lit[(e >> 16) + br.peek((int)((e >> 8) & 15))];
            }
            if (e & kLiteral) {
                if (o_end - o >= 8) {
                    // up to three lookups (six literals) per refill: 15 + 11 + 11 bits; every store writes two bytes
#define MGATK_PUT_LITERALS { br.drop((int)(e & 0xff)); o[0] = (uint8_t)(e >> 16); o[1] = (uint8_t)(e >> 24); o += 1 + ((e >> 8) & 1); }
                    MGATK_PUT_LITERALS
                    e = t.lit[br.peek(kLitRoot)];
                    if (e & kLiteral) {
                        MGATK_PUT_LITERALS
                        e = t.lit[br.peek(kLitRoot)];
                        if (e & kLiteral) MGATK_PUT_LITERALS
                    }
#undef MGATK_PUT_LITERALS
                    if (br.cnt < 0) return false;
                    continue;                                  // a non-literal entry is decoded after the refill
                }
                const int nl = 1 + (int)((e >> 8) & 1);        // last bytes of the output: exact stores
                if (o_end - o < nl) return false;
                br.drop((int)(e & 0xff));
                o[0] = (uint8_t)(e >> 16);
                if (nl == 2) o[1] = (uint8_t)(e >> 24);
                o += nl;
                if (br.cnt < 0) return false;
                continue;
            }
            if (e & kInvalid) return false;
            br.drop((int)(e & 0xff));
            if (e & kEndOfBlock) { if (br.cnt < 0) return false; break; }
            const int le = (int)((e >> 8) & 15);
            const uint32_t length = (e >> 16) + br.peek(le);
            br.drop(le);
            if (br.cnt < 32) br.refill();                      // at most 15 + 13 more bits
            uint32_t d = t.dist[br.peek(kDistRoot)];
            if (d & kSubtable) {
                br.drop(kDistRoot);
                d = t.dist[(d >> 16) + br.peek((int)((d >> 8) & 15))];
            }
            if (d & kInvalid) return false;
            br.drop((int)(d & 0xff));
            const int de = (int)((d >> 8) & 15);
            const uint32_t offset = (d >> 16) + br.peek(de);
            br.drop(de);
            if (br.cnt < 0) return false;
            if (offset > (size_t)(o - out) || length > (size_t)(o_end - o)) return false;
            const uint8_t *src = o - offset;
            if (offset >= 8 && (size_t)(o_end - o) >= length + 8) {         // word copies, may write up to 7 bytes beyond
                uint8_t *dst = o;
                const uint8_t *const stop = o + length;
                do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < stop);
            } else if (offset == 1) {
                memset(o, *src, length);
            } else {
                for (uint32_t k = 0; k < length; k++) o[k] = src[k];
            }
            o += length;
        }
    }
    return o == o_end;
}

}  // namespace mgatk_inflate
