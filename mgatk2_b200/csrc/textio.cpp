// mgatk2_b200 — dense-plane text writer (include/mgatk2_textio.h): the gzip text files of the
// reference's IncrementalTextWriter (src/file_io/writers.py:440-486) written straight from the uint16 planes.
//
// The listed cells are cut into groups; worker threads format the rows of a group ("pos,barcode,fwd,rev\n",
// positions ascending, 1-based) and deflate them into one gzip member each; the members are written in group order.
// Saturated plane entries (65535) take their exact value from the overflow list, as the reference's text files are
// not saturated (writers.py:440-466 write the Python ints).
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../../include/mgatk2_textio.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(const std::string &msg) { g_err = msg; return 1; }

// decimal digits of v at p, returns the end
inline char *put_u32(char *p, uint32_t v) {
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

struct OverflowKey { uint64_t key; uint32_t value; };      // key = cell << 32 | plane_pos

inline uint64_t ovf_key(int32_t cell, int plane, int pos) { return ((uint64_t)(uint32_t)cell << 32) | ((uint32_t)plane << 24) | (uint32_t)pos; }

struct Job {
    const uint16_t *planes; int32_t pos_pad, mito_length; int32_t plane_a, plane_b;
    const int32_t *cells; const char *names; const int64_t *name_end;
    const std::vector<OverflowKey> *ovf;
    int level;
};

inline uint32_t exact(const Job &j, int32_t cell, int plane, int pos, uint16_t v) {
    if (v != 65535 || j.ovf->empty()) return v;
    const uint64_t k = ovf_key(cell, plane, pos);
    auto it = std::lower_bound(j.ovf->begin(), j.ovf->end(), k, [](const OverflowKey &a, uint64_t b) { return a.key < b; });
    return (it != j.ovf->end() && it->key == k) ? it->value : (uint32_t)v;
}

// rows of the listed cells [i0, i1) appended to `text`; returns the number of rows
int64_t format_group(const Job &j, int64_t i0, int64_t i1, std::vector<char> &text) {
    int64_t rows = 0;
    size_t used = 0;
    for (int64_t i = i0; i < i1; i++) {
        const int32_t cell = j.cells[i];
        const char *name = j.names + (i ? j.name_end[i - 1] : 0);
        const size_t name_len = (size_t)(j.name_end[i] - (i ? j.name_end[i - 1] : 0));
        const uint16_t *a = j.planes + ((size_t)cell * MGATK_N_PLANES + (size_t)j.plane_a) * (size_t)j.pos_pad;
        const uint16_t *b = j.plane_b >= 0 ? j.planes + ((size_t)cell * MGATK_N_PLANES + (size_t)j.plane_b) * (size_t)j.pos_pad : nullptr;
        // worst case of a row: 10 + 1 + name + 1 + 10 + 1 + 10 + 1
        const size_t row_max = name_len + 36;
        for (int32_t p = 0; p < j.mito_length; p++) {
            const uint16_t va = a[p], vb = b ? b[p] : 0;
            if (!(va | vb)) continue;
            if (text.size() - used < row_max) text.resize(std::max(text.size() * 2, used + row_max + (1u << 16)));
            char *q = text.data() + used;
            q = put_u32(q, (uint32_t)p + 1u);
            *q++ = ',';
            memcpy(q, name, name_len); q += name_len;
            *q++ = ',';
            q = put_u32(q, exact(j, cell, j.plane_a, p, va));
            if (b) { *q++ = ','; q = put_u32(q, exact(j, cell, j.plane_b, p, vb)); }
            *q++ = '\n';
            used = (size_t)(q - text.data());
            rows++;
        }
    }
    text.resize(used);
    return rows;
}

// one gzip member holding `text`
bool gzip_member(const std::vector<char> &text, int level, std::vector<unsigned char> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)text.size()) + 64);
    zs.next_in = (Bytef *)text.data();
    zs.avail_in = (uInt)text.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    const size_t n = out.size() - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    out.resize(n);
    return true;
}

}  // namespace

extern "C" {

const char *mgatk_text_error(void) { return g_err.c_str(); }

int mgatk_text_write_plane_file(const char *path, const uint16_t *planes, int32_t n_cells, int32_t pos_pad,
                                int32_t mito_length, const mgatk_overflow *overflow, int64_t n_overflow,
                                int32_t plane_a, int32_t plane_b, const int32_t *cells, int64_t n_listed,
                                const char *names, const int64_t *name_end, int32_t level, int32_t n_threads,
                                int64_t *rows_out) {
    g_err.clear();
    if (!path || n_cells < 0 || pos_pad <= 0 || mito_length <= 0 || mito_length > pos_pad || n_listed < 0 || n_overflow < 0)
        return fail("bad argument");
    if (plane_a < 0 || plane_a >= MGATK_N_PLANES || plane_b >= MGATK_N_PLANES) return fail("plane index out of range");
    if (n_listed > 0 && (!planes || !cells || !names || !name_end)) return fail("null array");
    if (n_overflow > 0 && !overflow) return fail("null overflow list");
    if (level < 0 || level > 9) return fail("gzip level must be 0..9");
    for (int64_t i = 0; i < n_listed; i++)
        if (cells[i] < 0 || cells[i] >= n_cells) return fail("cell index out of range");
    std::vector<OverflowKey> ovf;
    for (int64_t i = 0; i < n_overflow; i++) {
        const int pl = (int)(overflow[i].plane_pos >> 24);
        if (pl == plane_a || pl == plane_b) ovf.push_back({((uint64_t)(uint32_t)overflow[i].cell << 32) | overflow[i].plane_pos, overflow[i].value});
    }
    std::sort(ovf.begin(), ovf.end(), [](const OverflowKey &a, const OverflowKey &b) { return a.key < b.key; });
    const Job job{planes, pos_pad, mito_length, plane_a, plane_b, cells, names, name_end, &ovf, level};

    FILE *f = fopen(path, "wb");
    if (!f) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
    // groups of cells: about 4 MB of text each (a member that small still compresses within 1 % of one stream)
    const int64_t group = std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)(4000000 / ((int64_t)mito_length * 12 + 1)) + 1));
    const int64_t n_groups = (n_listed + group - 1) / group;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n_groups));
    std::vector<std::vector<unsigned char>> done((size_t)n_groups);
    std::vector<char> ready((size_t)n_groups, 0);
    std::atomic<int64_t> next{0}, rows{0};
    std::atomic<bool> failed{false};
    std::mutex mu;
    std::condition_variable cv;
    int64_t written = 0;                                     // groups written so far (writer = the calling thread)
    auto work = [&]() {
        std::vector<char> text;
        for (;;) {
            const int64_t g = next.fetch_add(1);
            if (g >= n_groups || failed.load()) break;
            {   // at most 4 T finished members wait for the writer
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return g < written + 4 * (int64_t)T || failed.load(); });
            }
            text.clear();
            text.resize(1u << 20);
            rows.fetch_add(format_group(job, g * group, std::min(n_listed, (g + 1) * group), text));
            std::vector<unsigned char> z;
            if (!gzip_member(text, level, z)) failed.store(true);
            std::lock_guard<std::mutex> lk(mu);
            done[(size_t)g] = std::move(z);
            ready[(size_t)g] = 1;
            cv.notify_all();
        }
        std::lock_guard<std::mutex> lk(mu);
        cv.notify_all();
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work);
    bool io_error = false;
    for (int64_t g = 0; g < n_groups; g++) {
        std::vector<unsigned char> z;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready[(size_t)g] || failed.load(); });
            if (!ready[(size_t)g]) break;
            z = std::move(done[(size_t)g]);
        }
        if (!z.empty() && fwrite(z.data(), 1, z.size(), f) != z.size()) { io_error = true; failed.store(true); }
        {
            std::lock_guard<std::mutex> lk(mu);
            written = g + 1;
            cv.notify_all();
        }
        if (io_error) break;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        written = n_groups + 1;                              // release anyone still waiting
        cv.notify_all();
    }
    for (auto &x : th) x.join();
    if (n_groups == 0) {                                     // an empty file is still a valid gzip stream
        std::vector<unsigned char> z;
        if (!gzip_member(std::vector<char>(), level, z) || fwrite(z.data(), 1, z.size(), f) != z.size()) io_error = true;
    }
    if (fclose(f) != 0) io_error = true;
    if (io_error) return fail(std::string("write error on ") + path);
    if (failed.load()) return fail("deflate failed");
    if (rows_out) *rows_out = rows.load();
    return 0;
}

}  // extern "C"
