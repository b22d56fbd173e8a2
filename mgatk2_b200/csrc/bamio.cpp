// mgatk2_b200 — native host BAM ingest: replaces pysam at readers.py:85-88.
//
// Decodes every record placed on one contig of a BGZF-compressed BAM, in file order, straight into the
// structure-of-arrays batch of include/mgatk2_b200.h. Semantics follow what the reference relies on from
// pysam's AlignmentFile.fetch(contig): records placed on the contig INCLUDING unmapped mates placed
// there; reference_start = POS (0-based); SEQ/QUAL/CIGAR as stored (the cigar|seq|qual region of a BAM record is
// copied verbatim: it is the blob layout of the batch); template_length = TLEN; the barcode tag as a Z string
// compared verbatim. Host code only (g++, zlib, std::thread): BGZF blocks are inflated in parallel batches, records
// are parsed sequentially from the inflated stream. A .bai next to the file is used to start at the first chunk of
// the contig; without it the file is scanned from the first record.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/mgatk2_bamio.h"
#include "fast_inflate.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Ref { std::string name; int64_t len; };

inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Block { size_t off; uint32_t csize; uint32_t isize; };      // offset in the file, total block size, inflated size

struct Bytes {                                                     // growable byte buffer that is never zero-filled
    uint8_t *p = nullptr; size_t n = 0, cap = 0;
    ~Bytes() { free(p); }
    bool grow(size_t want) {
        if (want <= cap) return true;
        size_t c = cap ? cap : (1u << 20);
        while (c < want) c *= 2;
        uint8_t *q = (uint8_t *)realloc(p, c);
        if (!q) return false;
        p = q; cap = c;
        return true;
    }
    uint8_t *data() { return p; }
    size_t size() const { return n; }
    void drop_front(size_t k) { memmove(p, p + k, n - k); n -= k; }
};

// growable array that is never value-initialised: its pages are first touched by the decoding threads, not by a
// zero fill on the calling thread (which used to cost more than the decode itself)
template <class T> struct Arr {
    T *p = nullptr; size_t n = 0, cap = 0;
    Arr() = default;
    Arr(const Arr &) = delete;
    Arr &operator=(const Arr &) = delete;
    ~Arr() { free(p); }
    bool resize(size_t want) {
        if (want > cap) {
            size_t c = cap ? cap : (size_t)1 << 16;
            while (c < want) c += c / 2;
            T *q = (T *)realloc(p, c * sizeof(T));       // large blocks move by mremap, not by copying
            if (!q) return false;
            p = q; cap = c;
        }
        n = want;
        return true;
    }
    void clear() { n = 0; }
    T *release() {                                   // ownership to the caller, trimmed to size
        T *q = p;
        if (q && n < cap) { T *r = (T *)realloc(q, (n ? n : 1) * sizeof(T)); if (r) q = r; }
        if (q && n == 0) { free(q); q = nullptr; }
        p = nullptr; n = cap = 0;
        return q;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// distinct tag values of one thread: open addressing on a 64-bit hash, no std::string per record
struct TagTable {
    std::vector<int32_t> slot;                       // id + 1, 0 = empty
    std::vector<uint64_t> hash;
    std::vector<std::string> names;                  // first-appearance order
    size_t mask = 0;
    static uint64_t hash_of(const char *z, size_t n) {
        uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)n;
        size_t i = 0;
        for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, z + i, 8); h = (h ^ w) * 0x9e3779b97f4a7c15ull; h ^= h >> 29; }
        uint64_t w = 0;
        if (i < n) memcpy(&w, z + i, n - i);
        h = (h ^ w) * 0x9e3779b97f4a7c15ull;
        return h ^ (h >> 32);
    }
    void rehash(size_t cap) {
        slot.assign(cap, 0); hash.assign(cap, 0); mask = cap - 1;
        for (size_t id = 0; id < names.size(); id++) {
            const uint64_t h = hash_of(names[id].data(), names[id].size());
            size_t k = (size_t)h & mask;
            while (slot[k]) k = (k + 1) & mask;
            slot[k] = (int32_t)id + 1; hash[k] = h;
        }
    }
    int32_t get(const char *z, size_t n) {
        if (names.size() * 2 >= slot.size()) rehash(slot.empty() ? 1024 : slot.size() * 4);
        const uint64_t h = hash_of(z, n);
        size_t k = (size_t)h & mask;
        while (slot[k]) {
            if (hash[k] == h) {
                const std::string &s = names[(size_t)slot[k] - 1];
                if (s.size() == n && memcmp(s.data(), z, n) == 0) return slot[k] - 1;
            }
            k = (k + 1) & mask;
        }
        names.emplace_back(z, n);
        slot[k] = (int32_t)names.size(); hash[k] = h;
        return (int32_t)names.size() - 1;
    }
};

}  // namespace

struct mgatk_bam {
    std::string path, err;
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0;
    std::vector<Ref> refs;
    bool coordinate_sorted = false;
    size_t first_record_coff = 0;       // BGZF block that holds the first alignment record
    uint32_t first_record_uoff = 0;
    // decoded records of the last fetch
    Arr<int32_t> pos, tlen, bc_id;
    Arr<uint16_t> flag, l_seq, n_cigar;
    Arr<uint8_t> mapq, qual_missing;
    Arr<uint32_t> blob_off;
    Arr<uint8_t> blob;
    std::vector<std::string> barcodes;  // distinct tag values in order of first appearance (of the whole fetch, all parts)
    bool align_parts = false;          // mgatk_bam_align_parts: a part that reaches max_records runs on to the next start border
    // where a fetch that stopped at max_records goes on (mgatk_bam_fetch_more)
    struct Resume {
        Bytes s;                        // inflated stream; s[cur..] is not scanned yet
        size_t cur = 0, coff = 0, skip = 0;
        bool first = true, finished = true;
        int ref_id = -1;
        char tag[2] = {0, 0};
    } resume;
    std::unordered_map<std::string, int32_t> barcode_ids;
};

namespace {

int fail(mgatk_bam *h, int code, const std::string &msg) { h->err = msg; return code; }

// header of the BGZF block at `off`; false at a clean end of file or on a malformed block
bool block_at(const mgatk_bam *h, size_t off, Block *b) {
    if (off + 18 > h->size) return false;
    const uint8_t *p = h->map + off;
    if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return false;
    const uint32_t xlen = rd16(p + 10);
    if (off + 12 + xlen > h->size) return false;
    uint32_t bsize = 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {                 // extra subfields: SI1 SI2 SLEN data
        const uint8_t *s = p + 12 + x;
        const uint32_t slen = rd16(s + 2);
        if (4 + slen > xlen - x) return false;             // subfield runs past the extra field
        if (s[0] == 'B' && s[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(s + 4) + 1;
        x += 4 + slen;
    }
    if (bsize < 12 + xlen + 8 || off + bsize > h->size) return false;      // bsize <= 65536 by construction (16-bit + 1)
    b->off = off; b->csize = bsize; b->isize = rd32(p + bsize - 4);
    if (b->isize > 65536u) return false;                   // BGZF: at most 64 KiB of data per block
    return true;
}

std::atomic<long long> g_blocks_fast{0}, g_blocks_zlib{0};      // blocks decoded by fast_inflate.h / by zlib (MGATK_BAM_TIMING)

bool inflate_block(const mgatk_bam *h, const Block &b, uint8_t *dst) {
    const uint8_t *p = h->map + b.off;
    const uint32_t xlen = rd16(p + 10);
    const uint8_t *data = p + 12 + xlen;
    const size_t data_len = b.csize - 12 - xlen - 8;
    if (b.isize == 0) return true;
    // own decoder first (fast_inflate.h); the block trailer's CRC-32 decides whether its output stands. The eight
    // trailer bytes are handed over as slack for the bit reader, they are never part of a valid stream.
    static const bool use_fast = getenv("MGATK_BAM_ZLIB_ONLY") == nullptr;
    if (use_fast) {
        thread_local mgatk_inflate::Tables tables;
        if (mgatk_inflate::inflate_raw(data, data_len + 8, dst, b.isize, tables) &&
            (uint32_t)crc32(0L, dst, b.isize) == rd32(data + data_len)) {
            g_blocks_fast.fetch_add(1, std::memory_order_relaxed);
            return true;
        }
    }
    g_blocks_zlib.fetch_add(1, std::memory_order_relaxed);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(data);
    zs.avail_in = (uInt)data_len;
    zs.next_out = dst;
    zs.avail_out = b.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0 && (uint32_t)crc32(0L, dst, b.isize) == rd32(data + data_len);
}

// Inflates consecutive blocks starting at file offset `off` until `want` more bytes are available in `out`
// (appended) or the file ends. Blocks of a batch are inflated in parallel.
bool inflate_more(const mgatk_bam *h, size_t *off, Bytes *out, size_t want, int n_threads, std::string *err) {
    std::vector<Block> batch;
    size_t total = 0;
    while (total < want || batch.size() < 64) {
        Block b;
        if (!block_at(h, *off, &b)) {
            if (*off != h->size) { *err = "malformed BGZF block at offset " + std::to_string(*off); return false; }
            break;
        }
        batch.push_back(b);
        total += b.isize;
        *off += b.csize;
        if (batch.size() >= 4096) break;
    }
    if (batch.empty()) return true;
    const size_t base = out->size();
    if (!out->grow(base + total)) { *err = "out of memory"; return false; }
    out->n = base + total;
    std::vector<size_t> dst(batch.size());
    size_t o = base;
    for (size_t i = 0; i < batch.size(); i++) { dst[i] = o; o += batch[i].isize; }
    const int T = std::max(1, std::min<int>(n_threads, (int)batch.size()));
    std::vector<char> ok(T, 1);
    auto work = [&](int t) {
        for (size_t i = t; i < batch.size(); i += T)
            if (!inflate_block(h, batch[i], out->data() + dst[i])) ok[t] = 0;
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (char c : ok) if (!c) { *err = "BGZF block does not inflate (corrupt file?)"; return false; }
    return true;
}

// One reference of a .bai (SAM spec 5.2): smallest chunk offset over its real bins, the chunk and interval counts, and the
// two chunks of the metadata pseudo-bin 37450 that htslib writes (file range of the reference, mapped / unmapped-placed
// record counts). `has_meta` is false for indexes written without the pseudo-bin.
struct BaiRef { int32_t n_ref = 0, n_bin = 0, n_chunk = 0, n_intv = 0; uint64_t min_voff = ~0ull, ref_beg = 0, ref_end = 0, n_mapped = 0, n_unmapped = 0, first_intv = 0; bool has_meta = false; };

bool read_file(const std::string &p, std::vector<uint8_t> &d) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return false;
    uint8_t buf[1 << 16];
    size_t n;
    d.clear();
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    return true;
}

// walks the index up to `ref_id`; false when the bytes are not a well-formed BAI covering that reference
bool parse_bai(const std::vector<uint8_t> &d, int ref_id, BaiRef *out) {
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return false;
    size_t o = 4;
    BaiRef R;
    R.n_ref = rdi32(d.data() + o); o += 4;
    if (ref_id < 0 || ref_id >= R.n_ref) return false;
    for (int r = 0; r <= ref_id; r++) {
        const bool mine = r == ref_id;
        if (o + 4 > d.size()) return false;
        const int32_t n_bin = rdi32(d.data() + o); o += 4;
        if (n_bin < 0) return false;
        for (int b = 0; b < n_bin; b++) {
            if (o + 8 > d.size()) return false;
            const uint32_t bin = rd32(d.data() + o);
            const int32_t n_chunk = rdi32(d.data() + o + 4);
            o += 8;
            if (n_chunk < 0 || o + 16 * (size_t)n_chunk > d.size()) return false;
            if (mine) {
                if (bin == 37450) {                          // metadata pseudo-bin: (ref_beg, ref_end), (n_mapped, n_unmapped)
                    if (n_chunk == 2) {
                        R.has_meta = true;
                        R.ref_beg = rd64(d.data() + o); R.ref_end = rd64(d.data() + o + 8);
                        R.n_mapped = rd64(d.data() + o + 16); R.n_unmapped = rd64(d.data() + o + 24);
                    }
                } else {
                    R.n_bin++; R.n_chunk += n_chunk;
                    for (int c = 0; c < n_chunk; c++) R.min_voff = std::min(R.min_voff, rd64(d.data() + o + 16 * (size_t)c));
                }
            }
            o += 16 * (size_t)n_chunk;
        }
        if (o + 4 > d.size()) return false;
        const int32_t n_intv = rdi32(d.data() + o); o += 4;
        if (n_intv < 0 || o + 8 * (size_t)n_intv > d.size()) return false;
        if (mine) {
            R.n_intv = n_intv;
            for (int i = 0; i < n_intv && !R.first_intv; i++) R.first_intv = rd64(d.data() + o + 8 * (size_t)i);
        }
        o += 8 * (size_t)n_intv;
    }
    *out = R;
    return true;
}

// smallest virtual offset of the chunks of `ref_id` in the .bai next to the BAM; false when there is no usable index
bool bai_start(const mgatk_bam *h, int ref_id, uint64_t *voff, bool *empty) {
    std::string cand[2] = {h->path + ".bai", h->path};
    if (h->path.size() > 4 && h->path.substr(h->path.size() - 4) == ".bam") cand[1] = h->path.substr(0, h->path.size() - 4) + ".bai";
    for (const std::string &p : cand) {
        std::vector<uint8_t> d;
        BaiRef R;
        if (!read_file(p, d) || !parse_bai(d, ref_id, &R)) continue;
        *empty = R.min_voff == ~0ull;
        *voff = R.min_voff;
        return true;
    }
    return false;
}

int parse_header(mgatk_bam *h) {
    // the header may span several blocks: inflate until it is complete
    std::vector<uint8_t> s;
    size_t off = 0, consumed_blocks_end = 0;
    std::vector<size_t> block_end_u;            // inflated offset at which each block ends
    std::vector<size_t> block_off;              // file offset of each block
    auto more = [&]() -> bool {
        Block b;
        if (!block_at(h, off, &b)) return false;
        const size_t base = s.size();
        s.resize(base + b.isize);
        if (!inflate_block(h, b, s.data() + base)) return false;
        block_off.push_back(off);
        off += b.csize;
        block_end_u.push_back(s.size());
        consumed_blocks_end = off;
        return true;
    };
    auto need = [&](size_t n) -> bool { while (s.size() < n) if (!more()) return false; return true; };
    if (!need(12) || memcmp(s.data(), "BAM\1", 4) != 0) return fail(h, 2, "not a BAM file (bad magic)");
    const uint32_t l_text = rd32(s.data() + 4);
    if (!need(12 + (size_t)l_text)) return fail(h, 2, "truncated BAM header");
    const std::string text((const char *)s.data() + 8, l_text);
    const size_t hd = text.find("@HD");
    if (hd != std::string::npos) {
        const size_t eol = text.find('\n', hd);
        h->coordinate_sorted = text.substr(hd, eol == std::string::npos ? std::string::npos : eol - hd).find("SO:coordinate") != std::string::npos;
    }
    size_t o = 8 + l_text;
    const int32_t n_ref = rdi32(s.data() + o); o += 4;
    if (n_ref < 0) return fail(h, 2, "negative reference count");
    for (int r = 0; r < n_ref; r++) {
        if (!need(o + 4)) return fail(h, 2, "truncated reference list");
        const uint32_t l_name = rd32(s.data() + o); o += 4;
        if (!need(o + l_name + 4)) return fail(h, 2, "truncated reference list");
        Ref ref;
        ref.name.assign((const char *)s.data() + o, l_name ? l_name - 1 : 0);
        o += l_name;
        ref.len = rd32(s.data() + o); o += 4;
        h->refs.push_back(ref);
    }
    // virtual offset of the first alignment record
    size_t bi = 0;
    while (bi < block_end_u.size() && block_end_u[bi] <= o) bi++;
    if (bi < block_end_u.size()) {
        h->first_record_coff = block_off[bi];
        h->first_record_uoff = (uint32_t)(o - (bi ? block_end_u[bi - 1] : 0));
    } else {
        h->first_record_coff = consumed_blocks_end;
        h->first_record_uoff = 0;
    }
    return 0;
}

// value of the Z tag `tag` in the aux region; has=1 if the tag is present with any type
void find_tag(const uint8_t *a, const uint8_t *end, const char tag[2], bool *has, const char **z, size_t *zlen) {
    *has = false; *z = nullptr; *zlen = 0;
    while (a + 3 <= end) {
        const bool hit = a[0] == (uint8_t)tag[0] && a[1] == (uint8_t)tag[1];
        const char t = (char)a[2];
        a += 3;
        size_t n = 0;
        switch (t) {
            case 'A': case 'c': case 'C': n = 1; break;
            case 's': case 'S': n = 2; break;
            case 'i': case 'I': case 'f': n = 4; break;
            case 'Z': case 'H': {
                const uint8_t *e = (const uint8_t *)memchr(a, 0, end - a);
                if (!e) return;
                if (hit) { *has = true; if (t == 'Z') { *z = (const char *)a; *zlen = e - a; } return; }
                a = e + 1;
                continue;
            }
            case 'B': {
                if (a + 5 > end) return;
                const char st = (char)a[0];
                const uint32_t cnt = rd32(a + 1);
                const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                n = 5 + es * (size_t)cnt;
                break;
            }
            default: return;                                   // unknown type: stop scanning this record
        }
        if (hit) { *has = true; return; }
        if ((size_t)(end - a) < n) return;
        a += n;
    }
}

}  // namespace

extern "C" {

int mgatk_bam_open(const char *path, mgatk_bam **out) {
    if (!path || !out) return 1;
    mgatk_bam *h = new mgatk_bam();
    *out = h;
    h->path = path;
    h->fd = open(path, O_RDONLY);
    if (h->fd < 0) return fail(h, 1, std::string("cannot open: ") + strerror(errno));
    struct stat st;
    if (fstat(h->fd, &st) != 0 || st.st_size < 28) return fail(h, 2, "file too small to be a BAM");
    h->size = (size_t)st.st_size;
    void *m = mmap(nullptr, h->size, PROT_READ, MAP_PRIVATE, h->fd, 0);
    if (m == MAP_FAILED) return fail(h, 1, std::string("mmap failed: ") + strerror(errno));
    h->map = (const uint8_t *)m;
    madvise(m, h->size, MADV_SEQUENTIAL);
    return parse_header(h);
}

void mgatk_bam_close(mgatk_bam *h) {
    if (!h) return;
    if (h->map) munmap(const_cast<uint8_t *>(h->map), h->size);
    if (h->fd >= 0) close(h->fd);
    delete h;
}

const char *mgatk_bam_error(const mgatk_bam *h) { return h ? h->err.c_str() : "null handle"; }
int mgatk_bam_n_refs(const mgatk_bam *h) { return h ? (int)h->refs.size() : 0; }
const char *mgatk_bam_ref_name(const mgatk_bam *h, int i) { return (h && i >= 0 && i < (int)h->refs.size()) ? h->refs[i].name.c_str() : ""; }
int64_t mgatk_bam_ref_len(const mgatk_bam *h, int i) { return (h && i >= 0 && i < (int)h->refs.size()) ? h->refs[i].len : -1; }
int mgatk_bam_coordinate_sorted(const mgatk_bam *h) { return h && h->coordinate_sorted; }

// Decode every record placed on ref_id, in file order; at most max_records (< 0: all). tag: two characters.
static int fetch_run(mgatk_bam *h, int n_threads, int64_t max_records);

int mgatk_bam_fetch(mgatk_bam *h, int ref_id, const char *tag, int n_threads, int64_t max_records) {
    if (!h || !tag || ref_id < 0 || ref_id >= (int)h->refs.size()) return 1;
    h->err.clear();
    h->barcodes.clear(); h->barcode_ids.clear();
    mgatk_bam::Resume &st = h->resume;
    st.s.n = 0; st.cur = 0; st.first = true; st.finished = false; st.ref_id = ref_id; st.tag[0] = tag[0]; st.tag[1] = tag[1];
    st.coff = h->first_record_coff;
    st.skip = h->first_record_uoff;
    uint64_t voff = 0;
    bool empty = false;
    const bool indexed = h->coordinate_sorted && bai_start(h, ref_id, &voff, &empty);
    if (indexed) {
        if (empty) { st.finished = true; return fetch_run(h, n_threads, max_records); }
        st.coff = (size_t)(voff >> 16);
        st.skip = (size_t)(voff & 0xffff);
    }
    return fetch_run(h, n_threads, max_records);
}

int mgatk_bam_align_parts(mgatk_bam *h, int on) {
    if (!h) return 1;
    h->align_parts = on != 0;
    return 0;
}

int mgatk_bam_fetch_more(mgatk_bam *h, int n_threads, int64_t max_records) {
    if (!h || h->resume.ref_id < 0) return 1;
    h->err.clear();
    return fetch_run(h, n_threads, max_records);
}

// the records of the next part: up to max_records (all that are left if negative) into the arrays of the handle
static int fetch_run(mgatk_bam *h, int n_threads, int64_t max_records) {
    h->pos.clear(); h->tlen.clear(); h->bc_id.clear(); h->flag.clear(); h->l_seq.clear(); h->n_cigar.clear();
    h->mapq.clear(); h->qual_missing.clear(); h->blob_off.clear(); h->blob.clear();
    mgatk_bam::Resume &st = h->resume;
    if (st.finished || max_records == 0) return 0;
    const int ref_id = st.ref_id;
    const char *tag = st.tag;
    Bytes &s = st.s;
    size_t &cur = st.cur, &coff = st.coff;
    bool done = false, limit_hit = false;
    int32_t limit_pos = 0;
    struct Loc { size_t off; uint32_t bs; size_t blob_at; };      // record body in s, its size, its place in the blob
    std::vector<Loc> locs;
    const int T = std::max(1, n_threads);
    const bool timing = getenv("MGATK_BAM_TIMING") != nullptr;          // phase times on stderr (tools/bench_bamio.py)
    double t_inflate = 0, t_scan = 0, t_decode = 0, t_merge = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (!done) {
        double tp = now();
        // ---- inflate the next batch of blocks (in parallel) unless a whole record is still waiting (resumed fetch) ----
        if (st.first || s.size() - cur < 4 || s.size() - cur < 4 + (size_t)rd32(s.data() + cur)) {
            if (cur > 0) { s.drop_front(cur); cur = 0; }
            const size_t before = s.size();
            const size_t want = (s.size() - cur >= 4) ? 4 + (size_t)rd32(s.data() + cur) : 0;
            if (!inflate_more(h, &coff, &s, std::max<size_t>(want, 1u << 25), n_threads, &h->err)) return 2;
            if (st.first) { cur = std::min(st.skip, s.size()); st.first = false; }
            if (s.size() == before) {
                st.finished = true;
                if (s.size() - cur == 0) break;      // clean end of file
                return fail(h, 2, "truncated BAM record at end of file");
            }
        }
        t_inflate += now() - tp; tp = now();
        // ---- scan record borders (sequential, a few loads per record) ----
        locs.clear();
        size_t blob_at = h->blob.size();
        while (s.size() - cur >= 4 && s.size() - cur >= 4 + (size_t)rd32(s.data() + cur)) {
            const uint32_t bs = rd32(s.data() + cur);
            const uint8_t *r = s.data() + cur + 4;
            if (bs < 32) return fail(h, 2, "BAM record shorter than its fixed part");
            // the walk is a chain of dependent loads, one or two cache lines per record: records of a file have
            // similar sizes, so the headers a few records ahead are fetched on a guess
            if (cur + 6 * (size_t)(bs + 4) + 64 < s.size()) {
                __builtin_prefetch(s.data() + cur + 4 * (size_t)(bs + 4));
                __builtin_prefetch(s.data() + cur + 6 * (size_t)(bs + 4));
            }
            const int32_t rid = rdi32(r);
            cur += 4 + (size_t)bs;
            if (rid != ref_id) {
                if (h->coordinate_sorted && (rid > ref_id || rid < 0)) { done = true; st.finished = true; break; }    // past the contig
                continue;
            }
            if (limit_hit && rdi32(r + 4) != limit_pos) {     // the part ends before the first record of another start
                cur -= 4 + (size_t)bs;
                done = true;
                break;
            }
            const uint32_t l_name = r[8], ncig = rd16(r + 12), lseq = rd32(r + 16);
            const size_t nbytes = 4 * (size_t)ncig + (lseq + 1) / 2 + lseq;
            if (32 + (size_t)l_name + nbytes > bs) return fail(h, 2, "BAM record fields exceed the record size");
            if (lseq > 65535) return fail(h, 3, "read longer than 65535 bases: not supported by this batch layout");
            if (blob_at / 16 > 0xffffffffull) return fail(h, 3, "more than 64 GiB of cigar|seq|qual in one fetch");
            locs.push_back({(size_t)(r - s.data()), bs, blob_at});
            blob_at += (nbytes + 15) & ~(size_t)15;
            if (!limit_hit && max_records >= 0 && (int64_t)(h->pos.size() + locs.size()) >= max_records) {
                if (!h->align_parts) { done = true; break; }
                limit_hit = true;                             // go on while reference_start stays the same
                limit_pos = rdi32(r + 4);
            }
        }
        t_scan += now() - tp; tp = now();
        if (locs.empty()) continue;
        // ---- decode the records of the batch (in parallel: fields, blob copy, barcode tag) ----
        const size_t n0 = h->pos.size(), k = locs.size();
        if (!(h->pos.resize(n0 + k) && h->tlen.resize(n0 + k) && h->bc_id.resize(n0 + k) && h->flag.resize(n0 + k) &&
              h->l_seq.resize(n0 + k) && h->n_cigar.resize(n0 + k) && h->mapq.resize(n0 + k) &&
              h->qual_missing.resize(n0 + k) && h->blob_off.resize(n0 + k) && h->blob.resize(blob_at)))
            return fail(h, 2, "out of memory");
        const int nt = (int)std::min<size_t>((size_t)T, (k + 4095) / 4096);
        std::vector<TagTable> tables(nt);                                 // distinct tag values per thread, first appearance order
        auto work = [&](int t) {
            TagTable &ids = tables[t];
            const size_t i0 = k * (size_t)t / nt, i1 = k * (size_t)(t + 1) / nt;
            for (size_t i = i0; i < i1; i++) {
                const uint8_t *r = s.data() + locs[i].off;
                const uint32_t l_name = r[8], ncig = rd16(r + 12), lseq = rd32(r + 16);
                const size_t nbytes = 4 * (size_t)ncig + (lseq + 1) / 2 + lseq;
                const uint8_t *blob = r + 32 + l_name;                      // cigar | seq | qual, contiguous in the record
                h->pos[n0 + i] = rdi32(r + 4);
                h->mapq[n0 + i] = r[9];
                h->flag[n0 + i] = rd16(r + 14);
                h->n_cigar[n0 + i] = (uint16_t)ncig;
                h->l_seq[n0 + i] = (uint16_t)lseq;
                h->tlen[n0 + i] = rdi32(r + 28);
                h->blob_off[n0 + i] = (uint32_t)(locs[i].blob_at / 16);
                memcpy(h->blob.data() + locs[i].blob_at, blob, nbytes);
                memset(h->blob.data() + locs[i].blob_at + nbytes, 0, (size_t)(-(ptrdiff_t)nbytes & 15));   // padding to 16 bytes
                h->qual_missing[n0 + i] = lseq > 0 && blob[4 * (size_t)ncig + (lseq + 1) / 2] == 0xff;
                bool has;
                const char *z;
                size_t zl;
                find_tag(blob + nbytes, r + locs[i].bs, tag, &has, &z, &zl);
                int32_t id = has ? -2 : -1;                                 // -1 no tag, -2 tag that is not a string
                if (z) id = ids.get(z, zl);
                h->bc_id[n0 + i] = id;                                      // local id, made global below
            }
        };
        if (nt == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        t_decode += now() - tp; tp = now();
        // distinct barcodes in order of first appearance in the file: threads own consecutive record ranges
        for (int t = 0; t < nt; t++) {
            const std::vector<std::string> &names = tables[t].names;
            std::vector<int32_t> global(names.size());
            for (size_t j = 0; j < names.size(); j++) {
                auto it = h->barcode_ids.find(names[j]);
                if (it == h->barcode_ids.end()) {
                    global[j] = (int32_t)h->barcodes.size();
                    h->barcode_ids.emplace(names[j], global[j]);
                    h->barcodes.push_back(names[j]);
                } else global[j] = it->second;
            }
            const size_t i0 = k * (size_t)t / nt, i1 = k * (size_t)(t + 1) / nt;
            for (size_t i = i0; i < i1; i++) if (h->bc_id[n0 + i] >= 0) h->bc_id[n0 + i] = global[h->bc_id[n0 + i]];
        }
        t_merge += now() - tp;
    }
    if (timing) fprintf(stderr, "[bamio] inflate %.3f s, scan %.3f s, decode %.3f s, barcode merge %.3f s (%zu records, %d threads; "
                        "blocks so far: %lld own decoder, %lld zlib)\n",
                        t_inflate, t_scan, t_decode, t_merge, h->pos.size(), T, g_blocks_fast.load(), g_blocks_zlib.load());
    return 0;
}

int64_t mgatk_bam_n_records(const mgatk_bam *h) { return h ? (int64_t)h->pos.size() : 0; }
int64_t mgatk_bam_blob_bytes(const mgatk_bam *h) { return h ? (int64_t)h->blob.size() : 0; }
int64_t mgatk_bam_n_barcodes(const mgatk_bam *h) { return h ? (int64_t)h->barcodes.size() : 0; }
int64_t mgatk_bam_barcode_bytes(const mgatk_bam *h) {
    int64_t n = 0;
    if (h) for (const auto &b : h->barcodes) n += (int64_t)b.size();
    return n;
}

// copies the decoded records into caller arrays; barcode strings are concatenated, barcode_end[i] = end offset of string i
int mgatk_bam_export(const mgatk_bam *h, int32_t *pos, int32_t *tlen, uint16_t *flag, uint8_t *mapq, int32_t *bc_id,
                     uint16_t *l_seq, uint16_t *n_cigar, uint32_t *blob_off, uint8_t *blob, uint8_t *qual_missing,
                     char *barcode_chars, int64_t *barcode_end) {
    if (!h) return 1;
    const size_t n = h->pos.size();
    if (n) {
        memcpy(pos, h->pos.data(), 4 * n); memcpy(tlen, h->tlen.data(), 4 * n); memcpy(flag, h->flag.data(), 2 * n);
        memcpy(mapq, h->mapq.data(), n); memcpy(bc_id, h->bc_id.data(), 4 * n); memcpy(l_seq, h->l_seq.data(), 2 * n);
        memcpy(n_cigar, h->n_cigar.data(), 2 * n); memcpy(blob_off, h->blob_off.data(), 4 * n);
        memcpy(qual_missing, h->qual_missing.data(), n);
    }
    if (!h->blob.empty()) memcpy(blob, h->blob.data(), h->blob.size());
    int64_t o = 0;
    for (size_t i = 0; i < h->barcodes.size(); i++) {
        memcpy(barcode_chars + o, h->barcodes[i].data(), h->barcodes[i].size());
        o += (int64_t)h->barcodes[i].size();
        barcode_end[i] = o;
    }
    return 0;
}

int mgatk_bam_detach(mgatk_bam *h, void *arrays[10], char *barcode_chars, int64_t *barcode_end) {
    if (!h || !arrays) return 1;
    arrays[0] = h->pos.release(); arrays[1] = h->tlen.release(); arrays[2] = h->flag.release(); arrays[3] = h->mapq.release();
    arrays[4] = h->bc_id.release(); arrays[5] = h->l_seq.release(); arrays[6] = h->n_cigar.release();
    arrays[7] = h->blob_off.release(); arrays[8] = h->blob.release(); arrays[9] = h->qual_missing.release();
    int64_t o = 0;
    for (size_t i = 0; i < h->barcodes.size(); i++) {
        memcpy(barcode_chars + o, h->barcodes[i].data(), h->barcodes[i].size());
        o += (int64_t)h->barcodes[i].size();
        barcode_end[i] = o;
    }
    return 0;
}

void mgatk_bam_free(void *p) { free(p); }

int mgatk_bai_inspect(const char *bai_path, int ref_id, int64_t out[10]) {
    std::vector<uint8_t> d;
    BaiRef R;
    if (!bai_path || !out || !read_file(bai_path, d)) return 1;
    if (!parse_bai(d, ref_id, &R)) return 2;
    out[0] = R.n_ref; out[1] = R.n_bin; out[2] = R.n_chunk; out[3] = R.n_intv;
    out[4] = R.min_voff == ~0ull ? -1 : (int64_t)R.min_voff; out[5] = R.has_meta ? 1 : 0;
    out[6] = (int64_t)R.ref_beg; out[7] = (int64_t)R.ref_end; out[8] = (int64_t)R.n_mapped; out[9] = (int64_t)R.n_unmapped;
    return 0;
}

}  // extern "C"
