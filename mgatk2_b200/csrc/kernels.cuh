// mgatk2_b200 — sm_100a kernels of the per-cell chrM pileup hot path.
//
// Pipeline (one stream, no host synchronisation between stages):
//   1  k_hist / k_scan* / k_scatter   stage 1 (flag + whitelist filter, readers.py:96-111) fused into a
//                                     stable counting partition of the coordinate-sorted records by cell
//   2  k_cell_start, k_dedup          stage 2 (dedup, readers.py:118-150) on (cell,start) runs, per-cell
//                                     n_reads / n_paired (processors.py:33-34), global counters (readers.py:193-199)
//   3  k_plan*                        cut every cell into position tiles ("units") of bounded read count
//   4  k_pileup                       stages 3-6: a CTA stages the reads of a (cell, tile) in shared memory (cp.async);
//                                     CIGAR walk, base-quality / distance-from-end masks, per-base per-strand counting
//                                     and Tn5 sites (pileup.py:32-95) gathered per position in registers; strand-bias filter, coverage, Tn5 gating (pileup.py:128-154)
//                                     and depth statistics in the same pass; planes written once
//   5  k_base_totals, k_median        reference-allele vote input and median depth (writers.py:187-197,220-222)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgatk2_b200.h"
#include "bitplane.cuh"

namespace mgatk {

typedef unsigned long long u64;

#ifndef MGATK_PILEUP_WARPS
#define MGATK_PILEUP_WARPS 8
#endif
constexpr int kWarpsPerCta = MGATK_PILEUP_WARPS;       // k_pileup: warps per CTA; 32 warps per SM in all
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kScanGroup = 16;        // chunks per scan group
constexpr u32 kFull = 0xffffffffu;
constexpr int kSplitChunks = 4;       // k_pileup "deep" units: at most this many chunks of 32 positions; their reads come in batches

constexpr u32 ERR_UNSORTED = 1, ERR_EXTENT = 2, ERR_OVERFLOW_CAP = 4, ERR_SATURATED = 8;

// ReadRec.flags bits written by k_dedup, read by k_pileup
constexpr int GF_PROCESS = 1, GF_STRAND = 2, GF_KEEP = 4, GF_CELL_SHIFT = 8;   // cell index in bits 8..31
// set by k_pileup on its shared-memory copy: the read has query masks in its slot; one aligned block covers all of SEQ
constexpr int GF_MASKS = 8, GF_SIMPLE = 16;
// KeyRec.mq layout: mapq | strand<<8 | paired<<9
constexpr int GMQ_STRAND = 0x100, GMQ_PAIRED = 0x200;

// Records that passed stage 1, grouped by cell, BAM order inside a cell: one 32-byte sector per record, so
// the scattered side of the partition writes whole sectors (no read-modify-write of partial sectors in L2).
struct __align__(32) GroupRec {
    int32_t cell; int32_t pos; u32 tlen; u32 mq;       // dedup key (+ mapq | strand << 8 | paired << 9)
    u32 off; u32 len; u32 pad0; u32 pad1;              // blob offset (16 B units), l_seq | n_cigar << 16
};
struct __align__(16) ReadRec { int32_t pos; u32 off; u32 len; u32 flags; };     // what k_pileup consumes (from k_dedup)

struct Unit { int32_t cell, t0, t1, rbeg, rend; };

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// Stable counting partition by a digit of the cell index: per-CTA digit histograms of contiguous chunks
// (k_hist), an exclusive scan in (digit, chunk) order (k_scan_*), then a scatter in which every CTA walks
// its chunk in order against its own running offsets (k_scatter).
// ---------------------------------------------------------------------------------------------
struct SrcUser {          // pass 0: reads the caller's SoA batch and applies the stage-1 filter
    mgatk_batch b;
    int32_t n_cells;
    int vec4;             // bc_idx is 16-byte and flag 8-byte aligned: k_hist loads four records at a time
    __device__ __forceinline__ void cell4(int64_t i, int (&c)[4]) const {      // i % 4 == 0, i + 3 < count()
        const int4 bc = *reinterpret_cast<const int4 *>(b.bc_idx + i);
        const uint2 fl = *reinterpret_cast<const uint2 *>(b.flag + i);
        const int cc[4] = {bc.x, bc.y, bc.z, bc.w};
        const u32 ff[4] = {fl.x & 0xffffu, fl.x >> 16, fl.y & 0xffffu, fl.y >> 16};
#pragma unroll
        for (int k = 0; k < 4; k++) c[k] = ((ff[k] & 0x904u) || cc[k] < 0 || cc[k] >= n_cells) ? -1 : cc[k];
    }
    __device__ __forceinline__ int64_t count() const { return b.n_records; }
    __device__ __forceinline__ int cell(int64_t i) const {
        // readers.py:96-97 unmapped/secondary/supplementary; :104-111 tag absent or not whitelisted
        const int c = b.bc_idx[i];
        return ((b.flag[i] & 0x904) || c < 0 || c >= n_cells) ? -1 : c;
    }
    struct Raw { int32_t pos, tlen, bc, prev; u32 off; uint16_t flag, lseq, ncig; uint8_t mapq; bool valid; };
    __device__ __forceinline__ Raw load_raw(int64_t i, bool valid) const {     // loads only; combined one step later
        Raw r; r.valid = valid;
        r.pos = 0; r.tlen = 0; r.bc = -1; r.off = 0; r.flag = 0x4; r.lseq = 0; r.ncig = 0; r.mapq = 0; r.prev = 0x80000000;
        if (valid) {
            if ((threadIdx.x & 31) == 0 && i > 0) r.prev = b.pos[i - 1];     // sortedness check across the warp border
            r.pos = b.pos[i]; r.tlen = b.tlen[i]; r.bc = b.bc_idx[i]; r.off = b.blob_off[i];
            r.flag = b.flag[i]; r.lseq = b.l_seq[i]; r.ncig = b.n_cigar[i]; r.mapq = b.mapq[i];
        }
        return r;
    }
    // records must come sorted by reference_start (coordinate-sorted BAM): compare with the record before
    __device__ __forceinline__ bool out_of_order(const Raw &w) const {
        int32_t before = __shfl_up_sync(0xffffffffu, w.valid ? w.pos : 0x7fffffff, 1);
        if ((threadIdx.x & 31) == 0) before = w.prev;
        return w.valid && w.pos < before;
    }
    __device__ __forceinline__ GroupRec finish(const Raw &w) const {
        GroupRec r;
        const int32_t t = w.tlen;
        const uint16_t f = w.flag;
        const int c = w.bc;
        r.cell = (!w.valid || (f & 0x904) || c < 0 || c >= n_cells) ? -1 : c;
        r.pos = w.pos;
        r.tlen = t < 0 ? (u32)(-(int64_t)t) : (u32)t;       // abs(read.template_length), readers.py:124
        r.mq = (u32)w.mapq | ((f & 0x10) ? GMQ_STRAND : 0) | ((f & 0x1) ? GMQ_PAIRED : 0);
        r.off = w.off;
        r.len = (u32)w.lseq | ((u32)w.ncig << 16);
        r.pad0 = 0; r.pad1 = 0;
        return r;
    }
};

struct SrcGrouped {       // pass 1 of a two-digit partition: already filtered, count on the device
    const GroupRec *a;
    const int64_t *m;
    static constexpr int vec4 = 0;
    __device__ __forceinline__ void cell4(int64_t, int (&)[4]) const {}
    __device__ __forceinline__ int64_t count() const { return *m; }
    __device__ __forceinline__ int cell(int64_t i) const { return a[i].cell; }
    struct Raw { uint4 lo, hi; bool valid; };
    __device__ __forceinline__ Raw load_raw(int64_t i, bool valid) const {
        Raw r; r.valid = valid;
        r.lo = make_uint4(0xffffffffu, 0u, 0u, 0u); r.hi = make_uint4(0u, 0u, 0u, 0u);
        if (valid) { r.lo = reinterpret_cast<const uint4 *>(a + i)[0]; r.hi = reinterpret_cast<const uint4 *>(a + i)[1]; }
        return r;
    }
    __device__ __forceinline__ bool out_of_order(const Raw &) const { return false; }
    __device__ __forceinline__ GroupRec finish(const Raw &w) const {
        GroupRec r;
        r.cell = w.valid ? (int32_t)w.lo.x : -1; r.pos = (int32_t)w.lo.y; r.tlen = w.lo.z; r.mq = w.lo.w;
        r.off = w.hi.x; r.len = w.hi.y; r.pad0 = 0; r.pad1 = 0;
        return r;
    }
};

#ifndef MGATK_PART_THREADS
#define MGATK_PART_THREADS 256
#endif
constexpr int kPartThreads = MGATK_PART_THREADS;  // records per CTA step of the scatter
constexpr int kHistAhead = 4;      // steps loaded before counting (k_hist)

// Per-CTA digit histogram of a contiguous chunk of records (order does not matter for counting).
constexpr int kHistThreads = 1024;
template <class Src>
__global__ void __launch_bounds__(kHistThreads)
k_hist(Src src, int64_t chunk, int nchunks, int shift, int bins, u32 *__restrict__ mat) {
    extern __shared__ u32 smem[];
    u32 *h = smem;
    for (int b = threadIdx.x; b < bins; b += kHistThreads) h[b] = 0;
    __syncthreads();
    const int64_t n = src.count();
    int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk;
    if (end > n) end = n;
    if (src.vec4) {                                          // four records per thread and load, two loads in flight
        const int64_t end4 = beg + ((end - beg) & ~(int64_t)3);
        for (int64_t i0 = beg + 4 * (int64_t)threadIdx.x; i0 < end4; i0 += 8 * kHistThreads) {
            int c[2][4];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int64_t i = i0 + 4 * (int64_t)kHistThreads * k;
#pragma unroll
                for (int q = 0; q < 4; q++) c[k][q] = -1;
                if (i < end4) src.cell4(i, c[k]);
            }
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int q = 0; q < 4; q++) if (c[k][q] >= 0) atomicAdd(&h[(c[k][q] >> shift) & (bins - 1)], 1u);
        }
        for (int64_t i = end4 + threadIdx.x; i < end; i += kHistThreads) {
            const int c = src.cell(i);
            if (c >= 0) atomicAdd(&h[(c >> shift) & (bins - 1)], 1u);
        }
    } else
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += kHistThreads * kHistAhead) {
        int d[kHistAhead];
#pragma unroll
        for (int k = 0; k < kHistAhead; k++) {
            const int64_t i = i0 + (int64_t)kHistThreads * k;
            d[k] = -1;
            if (i < end) {
                const int c = src.cell(i);
                if (c >= 0) d[k] = (c >> shift) & (bins - 1);
            }
        }
#pragma unroll
        for (int k = 0; k < kHistAhead; k++) if (d[k] >= 0) atomicAdd(&h[d[k]], 1u);
    }
    __syncthreads();
    u32 *row = mat + (size_t)blockIdx.x * bins;
    for (int b = threadIdx.x; b < bins; b += kHistThreads) row[b] = h[b];
}

// scan of mat[chunk][bin] in (bin, chunk) order: S1 group sums, S2 bases, S3 in-place exclusive prefixes
__global__ void k_scan_group_sums(const u32 *__restrict__ mat, int nchunks, int bins, u32 *__restrict__ part) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (b >= bins) return;
    const int w0 = g * kScanGroup;
    u32 v[kScanGroup];
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) v[k] = w0 + k < nchunks ? mat[(size_t)(w0 + k) * bins + b] : 0u;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) s += v[k];
    part[(size_t)g * bins + b] = s;
}

__global__ void __launch_bounds__(1024)
k_scan_bases(u32 *__restrict__ part, int ngroups, int bins, int64_t *__restrict__ total_out) {
    __shared__ u32 warp_sums[32];
    __shared__ u32 carry_s;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < bins; b0 += 1024) {
        const int b = b0 + t;
        u32 tot = 0;
        if (b < bins) {
#pragma unroll 8
            for (int g = 0; g < ngroups; g++) tot += part[(size_t)g * bins + b];       // independent loads, eight in flight
        }
        u32 inc = tot;                                   // inclusive block scan of bin totals
        for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            u32 v = warp_sums[lane], s = v;
            for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(kFull, s, o); if (lane >= o) s += x; }
            warp_sums[lane] = s - v;                     // exclusive
        }
        __syncthreads();
        const u32 carry = carry_s;
        u32 run = carry + warp_sums[wid] + inc - tot;    // exclusive base of bin b
        if (b < bins) for (int g0 = 0; g0 < ngroups; g0 += 8) {
            u32 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = g0 + k < ngroups ? part[(size_t)(g0 + k) * bins + b] : 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) if (g0 + k < ngroups) { part[(size_t)(g0 + k) * bins + b] = run; run += v[k]; }
        }
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_sums[31] + inc;
        __syncthreads();
    }
    if (t == 0 && total_out) *total_out = carry_s;
}

__global__ void k_scan_apply(u32 *__restrict__ mat, int nchunks, int bins, const u32 *__restrict__ part) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (b >= bins) return;
    const int w0 = g * kScanGroup;
    u32 v[kScanGroup];
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) v[k] = w0 + k < nchunks ? mat[(size_t)(w0 + k) * bins + b] : 0u;
    u32 run = part[(size_t)g * bins + b];
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) if (w0 + k < nchunks) { mat[(size_t)(w0 + k) * bins + b] = run; run += v[k]; }
}

// Stable scatter. A CTA walks its chunk in steps of 256 records in BAM order, one record per thread (the
// next step's loads are issued before the current one is ranked). Ranking is parallel over the warps:
// match_any groups the lanes of a warp by digit, the group leaders add their group size into the byte of
// their warp in a packed per-digit counter (8 warps x 8 bits, two words), and after one barrier every thread
// reads "records of my digit in earlier warps" out of the lower bytes - ranks follow record order with
// one shared-memory atomic per (warp, digit) and two barriers per step (the packed counters are double
// buffered). Every record leaves as one full 32-byte sector.
constexpr int kPartWarps = kPartThreads / 32;
static_assert(kPartWarps == 8 || kPartWarps == 4, "the packed per-digit counter holds one byte per warp");

// packed per-digit counter of a step: one byte per warp
template <int kWarps> struct Packed;
template <> struct Packed<4> {
    u32 x;
    __device__ __forceinline__ void clear() { x = 0u; }
    __device__ __forceinline__ static void add(Packed *p, int wid, u32 n) { atomicAdd(&p->x, n << (8 * wid)); }
    __device__ __forceinline__ u32 before(int wid) const { return __dp4a(x & (wid == 0 ? 0u : (0xffffffffu >> (32 - 8 * wid))), 0x01010101u, 0u); }
    __device__ __forceinline__ bool first(int wid) const { return (x & (wid == 0 ? 0u : (0xffffffffu >> (32 - 8 * wid)))) == 0u; }
    __device__ __forceinline__ u32 total() const { return __dp4a(x, 0x01010101u, 0u); }
};
template <> struct Packed<8> {
    u32 x, y;                                              // x: warps 0-3, y: warps 4-7
    __device__ __forceinline__ void clear() { x = 0u; y = 0u; }
    __device__ __forceinline__ static void add(Packed *p, int wid, u32 n) { atomicAdd(wid < 4 ? &p->x : &p->y, n << (8 * (wid & 3))); }
    __device__ __forceinline__ u32 mx(int wid) const { return wid >= 4 ? 0xffffffffu : wid == 0 ? 0u : (0xffffffffu >> (32 - 8 * wid)); }
    __device__ __forceinline__ u32 my(int wid) const { return wid <= 4 ? 0u : (0xffffffffu >> (32 - 8 * (wid - 4))); }
    __device__ __forceinline__ u32 before(int wid) const { return __dp4a(x & mx(wid), 0x01010101u, __dp4a(y & my(wid), 0x01010101u, 0u)); }
    __device__ __forceinline__ bool first(int wid) const { return ((x & mx(wid)) | (y & my(wid))) == 0u; }
    __device__ __forceinline__ u32 total() const { return __dp4a(x, 0x01010101u, __dp4a(y, 0x01010101u, 0u)); }
};
typedef Packed<kPartWarps> Pack;
__host__ __device__ inline size_t scatter_smem_bytes(int bins) { return (size_t)bins * (2 * sizeof(Pack) + 4); }

#ifndef MGATK_SCATTER_CTAS
#define MGATK_SCATTER_CTAS 4
#endif
template <class Src>
__global__ void __launch_bounds__(kPartThreads, MGATK_SCATTER_CTAS)
k_scatter(Src src, int64_t chunk, int nchunks, int shift, int bins, const u32 *__restrict__ mat, GroupRec *__restrict__ dst,
          u64 *__restrict__ error_bits) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Pack *packed = reinterpret_cast<Pack *>(smem_raw);                 // [2][bins] per-warp byte counters of the step
    u32 *off = reinterpret_cast<u32 *>(smem_raw + (size_t)2 * bins * sizeof(Pack));   // [bins] running destination offsets
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const u32 *row = mat + (size_t)blockIdx.x * bins;
    for (int b = t; b < bins; b += kPartThreads) { off[b] = row[b]; packed[b].clear(); packed[bins + b].clear(); }
    const int64_t n = src.count();
    int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk;
    if (end > n) end = n;
    const u32 lt = (1u << lane) - 1;
    typename Src::Raw r1 = src.load_raw(beg + t, beg + t < end);                        // two steps of loads in flight
    typename Src::Raw r2 = src.load_raw(beg + kPartThreads + t, beg + kPartThreads + t < end);
    __syncthreads();
    int buf = 0;
    bool unsorted = false;
    for (int64_t i0 = beg; i0 < end; i0 += kPartThreads, buf ^= 1) {
        unsorted |= src.out_of_order(r1);
        const GroupRec cur = src.finish(r1);
        r1 = r2;
        r2 = src.load_raw(i0 + 2 * kPartThreads + t, i0 + 2 * kPartThreads + t < end);
        const int c = cur.cell;
        const int d = c >= 0 ? (c >> shift) & (bins - 1) : -1;
        Pack *pk = packed + (size_t)buf * bins;
        const u32 peers = __match_any_sync(kFull, d);
        const bool leader = d >= 0 && lane == __ffs(peers) - 1;
        if (leader) Pack::add(&pk[d], wid, (u32)__popc(peers));
        __syncthreads();
        Pack v; v.clear(); u32 base = 0;
        if (d >= 0) { v = pk[d]; base = off[d]; }
        __syncthreads();
        if (d >= 0) {
            if (leader && v.first(wid)) { off[d] = base + v.total(); pk[d].clear(); }   // first warp that holds the digit
            const size_t dd = (size_t)base + v.before(wid) + __popc(peers & lt);
            // one 256-bit store per record: a scattered store costs the LSU one pass per lane whatever its width
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + dd), "r"((u32)cur.cell), "r"((u32)cur.pos),
                         "r"(cur.tlen), "r"(cur.mq), "r"(cur.off), "r"(cur.len), "r"(0u), "r"(0u) : "memory");
        }
    }
    if (unsorted) atomicOr(error_bits, (u64)ERR_UNSORTED);
}

// ---------------------------------------------------------------------------------------------
// Stage 2: dedup. Inside a cell the records are sorted by start, so all candidates for a duplicate
// of record i sit directly before it in the same (cell, start) run; the first record of a key in
// BAM order survives (readers.py:129-150). Both key sets are evaluated for every stage-1 survivor
// (readers.py:128-144) so both duplicate counters are exact whichever strategy is selected.
// The kernel also applies the mapq gate (pileup.py:33-34: after dedup, a low-mapq first read still
// shadows its duplicates) and emits, compacted and in order, the 16-byte record k_pileup consumes:
// one pass, block-wise exclusive scan chained over blocks by decoupled look-back (blocks take
// their index from a ticket, so a block only ever waits for blocks that are already running).
// ---------------------------------------------------------------------------------------------
#ifndef MGATK_DEDUP_THREADS
#define MGATK_DEDUP_THREADS 256
#endif
#ifndef MGATK_DEDUP_ROUNDS
#define MGATK_DEDUP_ROUNDS 8
#endif
constexpr int kDedupThreads = MGATK_DEDUP_THREADS;
constexpr int kDedupRounds = MGATK_DEDUP_ROUNDS;             // records per thread
constexpr int kDedupTile = kDedupThreads * kDedupRounds;
constexpr int kDedupCells = 64;                             // consecutive cells of a tile counted in shared memory
constexpr u64 kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62, kScanValue = (1ull << 62) - 1;

__global__ void __launch_bounds__(kDedupThreads, 1536 / kDedupThreads)
k_dedup(const GroupRec *__restrict__ g, const int64_t *__restrict__ m_ptr, ReadRec *__restrict__ recs, int dedup_mode, int min_mapq,
        mgatk_cell_qc *__restrict__ qc, mgatk_stats *__restrict__ stats, u32 *__restrict__ ticket,
        u64 *__restrict__ scan_state, int64_t *__restrict__ n_proc_out) {
    __shared__ u32 s_cnt[4];
    __shared__ u32 s_warp[kDedupRounds][kDedupThreads / 32];
    __shared__ u32 s_blk;
    __shared__ u64 s_prefix;
    __shared__ u32 s_cell[kDedupCells][3];                   // per-cell counts of the tile's first cells: one global atomic per (tile, cell)
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    for (int e = threadIdx.x; e < kDedupCells * 3; e += kDedupThreads) (&s_cell[0][0])[e] = 0;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 blk = s_blk;
    const int64_t m = *m_ptr;
    // the records are grouped by cell: a tile holds a short run of consecutive cells, counted in shared memory
    const int cell0 = (int64_t)blk * kDedupTile < m ? g[(int64_t)blk * kDedupTile].cell : 0;
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    ReadRec rr[kDedupRounds];
    u32 pm[kDedupRounds];
    u32 n_keep = 0, n_len = 0, n_pos = 0, n_empty = 0;
#pragma unroll
    for (int k = 0; k < kDedupRounds; k++) {
        const int64_t i = (int64_t)blk * kDedupTile + k * kDedupThreads + threadIdx.x;
        int cell = -1;
        bool keep = false, paired = false, process = false;
        rr[k].pos = 0; rr[k].off = 0; rr[k].len = 0; rr[k].flags = 0;
        uint4 me = make_uint4(0xffffffffu, 0u, 0u, 0u);                      // cell, pos, |tlen|, mq
        uint2 lc = make_uint2(0u, 0u);                                       // off, len
        if (i < m) { me = reinterpret_cast<const uint4 *>(g + i)[0]; lc = reinterpret_cast<const uint2 *>(g + i)[2]; }
        // the record before this one is the neighbouring lane's (lane 0 fetches it)
        uint4 prev;
        prev.x = __shfl_up_sync(kFull, me.x, 1); prev.y = __shfl_up_sync(kFull, me.y, 1);
        prev.z = __shfl_up_sync(kFull, me.z, 1); prev.w = __shfl_up_sync(kFull, me.w, 1);
        if (lane == 0 && i > 0 && i < m) prev = reinterpret_cast<const uint4 *>(g + i - 1)[0];
        if (i < m) {
            cell = (int)me.x;
            const u32 strand = me.w & GMQ_STRAND;
            paired = me.w & GMQ_PAIRED;
            bool len_dup = false, pos_dup = false;
            if (dedup_mode != MGATK_DEDUP_NONE) {
                uint4 o = prev;
                for (int64_t j = i - 1; j >= 0; ) {
                    if (o.y != me.y || o.x != me.x) break;
                    if ((o.w & GMQ_STRAND) == strand) {
                        pos_dup = true;
                        if (o.z == me.z) { len_dup = true; break; }
                    }
                    if (--j >= 0) o = reinterpret_cast<const uint4 *>(g + j)[0];
                }
            }
            keep = dedup_mode == MGATK_DEDUP_FRAGMENT_LENGTH ? !len_dup : dedup_mode == MGATK_DEDUP_POSITION_ONLY ? !pos_dup : true;
            // pileup.py:33-34 mapq gate (after dedup, Q2). An empty SEQ makes the reference raise
            // (readers.py:157); such survivors are reported in stats.n_empty_seq and not piled up.
            process = keep && (int)(me.w & 0xff) >= min_mapq && (lc.y & 0xffff) != 0;
            rr[k].pos = (int32_t)me.y; rr[k].off = lc.x; rr[k].len = lc.y;
            rr[k].flags = GF_PROCESS | (strand ? GF_STRAND : 0) | GF_KEEP | ((u32)cell << GF_CELL_SHIFT);
            n_len += len_dup; n_pos += pos_dup; n_keep += keep; n_empty += keep && (lc.y & 0xffff) == 0;
        }
        // per-cell survivors: lanes of a warp mostly share one cell
        const u32 peers = __match_any_sync(kFull, cell);
        const u32 kept = __ballot_sync(kFull, keep), paird = __ballot_sync(kFull, keep && paired);
        pm[k] = __ballot_sync(kFull, process);
        if (cell >= 0 && lane == __ffs(peers) - 1) {
            const u32 nk = __popc(kept & peers), np = __popc(paird & peers), npr = __popc(pm[k] & peers);
            const u32 rel = (u32)(cell - cell0);
            if (rel < (u32)kDedupCells) {
                if (nk) atomicAdd(&s_cell[rel][0], nk);
                if (np) atomicAdd(&s_cell[rel][1], np);
                if (npr) atomicAdd(&s_cell[rel][2], npr);
            } else {
                if (nk) atomicAdd(&qc[cell].n_reads, nk);
                if (np) atomicAdd(&qc[cell].n_paired, np);
                if (npr) atomicAdd(&qc[cell].median_lo, npr);
            }
        }
        if (lane == 0) s_warp[k][wid] = __popc(pm[k]);
    }
    for (int o = 16; o; o >>= 1) {
        n_keep += __shfl_xor_sync(kFull, n_keep, o); n_len += __shfl_xor_sync(kFull, n_len, o);
        n_pos += __shfl_xor_sync(kFull, n_pos, o); n_empty += __shfl_xor_sync(kFull, n_empty, o);
    }
    if (lane == 0) {
        if (n_keep) atomicAdd(&s_cnt[0], n_keep);
        if (n_len) atomicAdd(&s_cnt[1], n_len);
        if (n_pos) atomicAdd(&s_cnt[2], n_pos);
        if (n_empty) atomicAdd(&s_cnt[3], n_empty);
    }
    __syncthreads();
    // per-cell survivors (processors.py:33-34) and reads to pile up per cell (the cell borders of the compacted records;
    // parked in median_lo, which shares a sector with the two counters and is only written by k_median at the very end)
    for (int e = threadIdx.x; e < kDedupCells; e += kDedupThreads) {
        if (s_cell[e][0]) atomicAdd(&qc[cell0 + e].n_reads, s_cell[e][0]);
        if (s_cell[e][1]) atomicAdd(&qc[cell0 + e].n_paired, s_cell[e][1]);
        if (s_cell[e][2]) atomicAdd(&qc[cell0 + e].median_lo, s_cell[e][2]);
    }
    // stable compaction of the reads that are piled up: position inside the tile, then the tile's prefix
    if (wid == 0) {
        // exclusive scan of the kDedupRounds x (warps) warp counts (round-major = record order), kPer per lane
        constexpr int kCounts = kDedupRounds * (kDedupThreads / 32), kPer = kCounts / 32;
        static_assert(kCounts % 32 == 0, "whole counts per lane");
        u32 *flat = &s_warp[0][0];
        u32 c[kPer], incl = 0;
#pragma unroll
        for (int q = 0; q < kPer; q++) { c[q] = flat[kPer * lane + q]; incl += c[q]; }
        const u32 mine = incl;
        for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
        const u32 total = __shfl_sync(kFull, incl, 31);
        u32 run = incl - mine;
#pragma unroll
        for (int q = 0; q < kPer; q++) { flat[kPer * lane + q] = run; run += c[q]; }
        // decoupled look-back, 128 predecessors at a time
        u64 prefix = 0;
        if (lane == 0) atomicExch((unsigned long long *)&scan_state[blk], (blk == 0 ? kScanPrefix : kScanAggregate) | (u64)total);
        for (int64_t top = (int64_t)blk - 1; top >= 0; top -= 128) {
            u64 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {                    // lane-major: lane l looks at top - 4l - k
                const int64_t b = top - 4 * lane - k;
                v[k] = kScanPrefix;                          // below block 0: an empty prefix
                if (b >= 0) v[k] = *(volatile u64 *)&scan_state[b];
            }
            bool stop_here = false;                          // this lane holds the nearest full prefix
            u64 add = 0;
            bool seen = false;                               // a full prefix was seen at a nearer block of this lane
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t b = top - 4 * lane - k;
                while (b >= 0 && (v[k] & ~kScanValue) == 0) v[k] = *(volatile u64 *)&scan_state[b];
                if (!seen) { add += v[k] & kScanValue; if (v[k] & kScanPrefix) { seen = true; stop_here = true; } }
            }
            const u32 has_prefix = __ballot_sync(kFull, stop_here);
            const int stop = __ffs(has_prefix) - 1;          // nearest lane with a full prefix (or -1)
            if (stop >= 0 && lane > stop) add = 0;
            for (int o = 16; o; o >>= 1) add += __shfl_xor_sync(kFull, add, o);
            prefix += add;
            if (stop >= 0) break;
        }
        if (lane == 0) {
            if (blk != 0) atomicExch((unsigned long long *)&scan_state[blk], kScanPrefix | (prefix + total));
            s_prefix = prefix;
            if ((int64_t)(blk + 1) * kDedupTile >= m && (int64_t)blk * kDedupTile < (m > 0 ? m : 1)) *n_proc_out = (int64_t)(prefix + total);
            if (s_cnt[0]) atomicAdd((u64 *)&stats->filtered_reads, (u64)s_cnt[0]);
            if (s_cnt[1]) atomicAdd((u64 *)&stats->dup_with_length, (u64)s_cnt[1]);
            if (s_cnt[2]) atomicAdd((u64 *)&stats->dup_position_only, (u64)s_cnt[2]);
            if (s_cnt[3]) atomicAdd((u64 *)&stats->n_empty_seq, (u64)s_cnt[3]);
        }
    }
    __syncthreads();
    const u64 prefix = s_prefix;
#pragma unroll
    for (int k = 0; k < kDedupRounds; k++)
        if ((pm[k] >> lane) & 1u) recs[prefix + s_warp[k][wid] + __popc(pm[k] & ((1u << lane) - 1u))] = rr[k];
}

// ---------------------------------------------------------------------------------------------
// Work planning: a unit is (cell, position tile). Tile borders sit at every `unit_reads`-th read of
// the cell (rounded down to a chunk of 32 positions), so units carry about the same number of reads
// wherever the cell's coverage is dense or sparse. Dead cells (processors.py:22) and cells without
// reads to pile up get one unit that only writes zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tiles_for(int cnt, int unit_reads, int ppad) {
    int nt = cnt <= 0 ? 1 : (cnt + unit_reads - 1) / unit_reads;
    const int max_nt = ppad / 32;
    return nt > max_nt ? max_nt : nt;
}

__device__ __forceinline__ bool cell_dead(const mgatk_cell_qc &q, int min_reads) {
    return q.n_reads == 0 || (int64_t)q.n_reads < (int64_t)min_reads;
}

// one pass over the cells: first compacted record of every cell (exclusive scan of the per-cell counts of k_dedup)
// and first unit of every cell (exclusive scan of the tile counts), both in one 64-bit scan
__global__ void __launch_bounds__(1024)
k_plan_scan(const mgatk_cell_qc *__restrict__ qc, int n_cells, int min_reads,
            int unit_reads, int ppad, int32_t *__restrict__ cell_start, int32_t *__restrict__ unit_start, int32_t *__restrict__ n_units) {
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry_s;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_cells; c0 += 1024) {
        const int c = c0 + t;
        u64 mine = 0;                                        // records << 32 | tiles
        if (c < n_cells) {
            const u32 np = qc[c].median_lo;                  // parked there by k_dedup
            const int cnt = cell_dead(qc[c], min_reads) ? 0 : (int)np;
            mine = ((u64)np << 32) | (u32)tiles_for(cnt, unit_reads, ppad);
        }
        u64 inc = mine;
        for (int o = 1; o < 32; o <<= 1) { const u64 v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const u64 v = warp_sums[lane];
            u64 sc = v;
            for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(kFull, sc, o); if (lane >= o) sc += x; }
            warp_sums[lane] = sc - v;
        }
        __syncthreads();
        const u64 carry = carry_s;
        const u64 excl = carry + warp_sums[wid] + inc - mine;
        if (c < n_cells) { cell_start[c] = (int32_t)(excl >> 32); unit_start[c] = (int32_t)(u32)excl; }
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_sums[31] + inc;
        __syncthreads();
    }
    if (t == 0) { cell_start[n_cells] = (int32_t)(carry_s >> 32); unit_start[n_cells] = (int32_t)(u32)carry_s; *n_units = (int32_t)(u32)carry_s; }
}

__global__ void k_plan_units(const int32_t *__restrict__ cell_start, const mgatk_cell_qc *__restrict__ qc,
                             const ReadRec *__restrict__ recs, const int32_t *__restrict__ unit_start, int n_cells,
                             int min_reads, int unit_reads, int ppad, int halo, Unit *__restrict__ units,
                             int cap_reads, int overflow_slack, Unit *__restrict__ units_overflow, int32_t *__restrict__ n_overflow_units) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= unit_start[n_cells]) return;
    int lo = 0, hi = n_cells;                              // last cell with unit_start[c] <= u
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (unit_start[mid] <= u) lo = mid; else hi = mid; }
    const int c = lo;
    const bool dead = cell_dead(qc[c], min_reads);
    const int cs = cell_start[c], ce = cell_start[c + 1];
    const int cnt = dead ? 0 : ce - cs;
    const int nt = tiles_for(cnt, unit_reads, ppad), k = u - unit_start[c];
    const int per = cnt > 0 ? (cnt + nt - 1) / nt : 0;
    Unit un;
    un.cell = c;
    un.t0 = 0; un.t1 = ppad;
    if (k > 0) un.t0 = min(max(recs[cs + min(k * per, cnt - 1)].pos, 0), ppad) & ~31;
    if (k + 1 < nt) un.t1 = min(max(recs[cs + min((k + 1) * per, cnt - 1)].pos, 0), ppad) & ~31;
    if (un.t1 < un.t0) un.t1 = un.t0;
    if (cnt == 0) { un.rbeg = un.rend = 0; }
    else {
        const int first = un.t0 - halo + 1;                // reads starting before cannot reach t0
        int a = cs, b = un.t0 == 0 ? cs : ce;              // the leftmost tile also takes (and extent-checks) reads left of 0
        while (a < b) { int mid = (a + b) >> 1; if (recs[mid].pos < first) a = mid + 1; else b = mid; }
        un.rbeg = a;
        b = ce;
        while (a < b) { int mid = (a + b) >> 1; if (recs[mid].pos < un.t1) a = mid + 1; else b = mid; }
        un.rend = a;
    }
    // tiles with clearly more reads than a CTA has mask slots go to the overflow list (walked in sub-tiles by the kSplit
    // kernel); a handful of reads beyond the slots is cheaper on the per-base path than a second pass over the halo
    if (un.rend - un.rbeg > cap_reads + overflow_slack && un.t1 - un.t0 > 32 * kSplitChunks) {
        units_overflow[atomicAdd(n_overflow_units, 1)] = un;
        un.t1 = un.t0;                                      // empty here
    }
    units[u] = un;
}

// ---------------------------------------------------------------------------------------------
// Stages 3-6 as a bit-plane gather. One CTA per unit (cell, position tile):
//   phase A every warp takes 32 reads of the unit at a time, one per lane: record from global memory, the
//           cigar|seq|qual blob through a small per-warp staging buffer (cp.async, L2 -> shared, no L1
//           pollution), then SEQ/QUAL turned into three bit planes in query coordinates (bitplane.cuh: V =
//           base is A/C/G/T, base quality, distance-from-end window, pileup.py:67-86; B0, B1 = the two bits of
//           the base code), word-parallel; only the planes (16 bytes per 32 bases) and the record stay in shared
//           memory. The declared extent is verified and the per-chunk "first candidate" table is filled in the
//           same pass; warps run independently.
//   phase B the warps take chunks of 32 positions. The candidate reads of a chunk are those starting in
//           (chunk - extent, chunk + 32); 32 candidates at a time, one per lane: a read with one aligned block
//           over all of SEQ cuts the 32-bit window at query offset chunk - start out of its planes, any other
//           read walks its CIGAR (pileup.py:52-95) and ORs the windows of its blocks; 32x32 bit transposes
//           across the warp turn "lane = read" into "lane = position" for the four bit matrices of the round
//           (V, B0, B1 and the Tn5 sites, pileup.py:43-50; one or two transposes when the candidates fit 8 or
//           16 lanes), and one LOP3 + popcount per base and strand adds up the ten counters in registers.
//           Nothing is shared between warps, so there are no atomics
//           on counters; when the chunk's reads are exhausted the counts are final and the strand-bias filter,
//           coverage, Tn5 gating (pileup.py:128-154) and the depth statistics are applied in registers and the
//           11 planes are written once.
// Reads that do not fit (blob larger than the warp buffer, more reads than mask slots) take a per-base path
// from global memory (same results).
// ---------------------------------------------------------------------------------------------
struct PileupArgs {
    const ReadRec *recs;
    const uint8_t *blob;
    const Unit *units; const int32_t *n_units; int32_t *work_counter;
    uint16_t *planes; mgatk_cell_qc *qc; mgatk_stats *stats;
    mgatk_overflow *ovf; int64_t ovf_cap;
    int P, ppad, min_baseq, dist, apply_bias, extent, raw;
    int accumulate;                  // streaming: add the counts of this batch to the planes, nothing else (MGATK_FLAG_ACCUMULATE)
#ifdef MGATK_TIMING
    unsigned long long *dbg;          // -DMGATK_TIMING: warp-cycles per phase (profiling build only, see profiles/README.md)
#endif
    int mask_stride;                 // bytes of one read's mask slot: 16 * ceil(extent / 32)
    int cap_reads;                   // mask slots per CTA (<= kStageReads)
    int group_reads;                 // reads a warp stages at a time: 32, fewer when 32 average blobs exceed the warp buffer
    double max_bias;
};

constexpr int kOpCap = 1 << 20;     // cigar lengths above this cannot be valid for a short-read batch

__device__ __forceinline__ void cp_async16(u32 dst_shared, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct SharedMem {                   // word access to shared memory by 32-bit shared address (bitplane.cuh's `M`)
    __device__ __forceinline__ u32 ld32(u32 a) const { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
    __device__ __forceinline__ void st128(u32 a, u32 x, u32 y, u32 z, u32 w) const {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
    __device__ __forceinline__ void ld128(u32 a, u32 (&v)[4]) const {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
    }
};

constexpr int kStageReads = 64 * kWarpsPerCta;       // records of a batch kept in shared memory
constexpr int kMaskBytes = 32 * kStageReads;         // query masks of a batch (32-byte slots at 2x50 bp)
constexpr int kWarpBuf = 2560;       // per-warp blob staging buffer: 32 blobs of a 50 bp read
constexpr int kWarpBufSlack = 64;    // phase A may load this far past the last staged byte
constexpr int kMaxItems = 256;       // 32 reads x 8 groups of 32 bases per staging pass
constexpr int kChunkSeg = 576;       // chunks per pass over a unit (the 518 chunks of chrM in one)
constexpr int kAccWords = 10 * 32;   // deep units: counts of one chunk, [8 base x strand + 2 Tn5][32]

// dynamic shared memory of k_pileup: records, query masks, per-warp staging buffers
__host__ __device__ inline size_t pileup_smem_bytes() {
    return (size_t)kStageReads * sizeof(ReadRec) + kMaskBytes + (size_t)kWarpsPerCta * (kWarpBuf + kWarpBufSlack);
}

// Per-base form of one aligned block for a read without query masks: chunk bits [pa, pa + span)
// for query bases q0.. (already clipped to the distance-from-end window). m = planes V, B0, B1.
__device__ __noinline__ void block_masks_global(const uint8_t *seq, int L, int pa, int span, int q0, int min_baseq, u32 (&m)[3]) {
    const int8_t *qual = reinterpret_cast<const int8_t *>(seq) + ((L + 1) >> 1);
    const int b0 = max(pa, 0), b1 = min(pa + span, 32);
    for (int b = b0; b < b1; b++) {
        const int q = q0 + (b - pa);                           // pileup.py:75
        if ((int)__ldg(qual + q) < min_baseq) continue;        // int8 compare, pileup.py:80
        const u32 by = __ldg(seq + (q >> 1));
        const u32 nib = (q & 1) ? (by & 15u) : (by >> 4);
        const u32 bit = 1u << b;
        if (nib == 1) m[0] |= bit;                             // pileup.py:83-86: A, C, G, T only
        else if (nib == 2) { m[0] |= bit; m[1] |= bit; }
        else if (nib == 4) { m[0] |= bit; m[2] |= bit; }
        else if (nib == 8) { m[0] |= bit; m[1] |= bit; m[2] |= bit; }
    }
}

// per-lane constants of the five butterfly stages: byte selectors for j = 16, 8 (whole bytes move), keep mask and
// rotate amount for j = 4, 2, 1
struct TransposeConst { u32 sel[2], keep[3], amt[3]; };
__device__ __forceinline__ TransposeConst make_transpose_const(int lane) {
    TransposeConst tc;
    tc.sel[0] = transpose_sel(lane, 16); tc.sel[1] = transpose_sel(lane, 8);
    asm volatile("" : "+r"(tc.sel[0]), "+r"(tc.sel[1]));
#pragma unroll
    for (int s = 0; s < 3; s++) {
        tc.keep[s] = transpose_keep(lane, 4 >> s); tc.amt[s] = transpose_amt(lane, 4 >> s);
        // opaque to the compiler: it would otherwise recompute both with two SELs per stage on the saturated ALU pipe
        asm volatile("" : "+r"(tc.keep[s]), "+r"(tc.amt[s]));
    }
    return tc;
}
__device__ __forceinline__ u32 warp_transpose(u32 x, const TransposeConst &tc) {
    x = transpose_stage_bytes(x, __shfl_xor_sync(kFull, x, 16), tc.sel[0]);
    x = transpose_stage_bytes(x, __shfl_xor_sync(kFull, x, 8), tc.sel[1]);
#pragma unroll
    for (int s = 0; s < 3; s++) x = transpose_stage(x, __shfl_xor_sync(kFull, x, 4 >> s), tc.keep[s], tc.amt[s]);
    return x;
}

// reference span check of one read against the declared extent (cigar words from global memory)
__device__ __forceinline__ bool span_exceeds(const u32 *cig, int ncig, int L, int extent) {
    if (L > extent) return true;
    int span = 0;
    for (int ci = 0; ci < ncig; ci++) {
        const u32 w = __ldg(cig + ci);
        const int op = w & 15;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += min((int)(w >> 4), kOpCap);
        if (span > extent) return true;
    }
    return false;
}

// Phase A for the reads [j0, j0 + 32) of a batch of nb reads (g_rec[0..nb)), one per lane: record, blob through
// the warp's staging buffer, query masks into slot j, first-candidate table entries. ns = reads with a slot.
__device__ __forceinline__ void stage_reads(const PileupArgs &a, const ReadRec *g_rec, int nb, int ns, int j0, int lane,
                                            ReadRec *s_rec, u32 mask_addr, u32 wbuf_addr, uint8_t *s_items, int *s_first, int seg0, int nseg,
                                            int q_lo, QualGe qg, bool &extent_err, int gs) {
    const SharedMem smem;
    // a group holds `gs` <= 32 reads: as many as fit the staging buffer in one pass, so that long reads spread over all
    // warps of the CTA instead of queueing for the buffer of one
    const bool mine = lane < gs;
    const int j = mine ? j0 + lane : 0x3fffffff;
    ReadRec rr; rr.pos = 0; rr.off = 0; rr.len = 0; rr.flags = 0;
    if (j < ns) rr = g_rec[j];
    const int L = rr.len & 0xffff, ncig = rr.len >> 16;
    const int nq = (L + 31) >> 5;
    const u32 blob16 = (u32)((4 * ncig + ((L + 1) >> 1) + L + 15) & ~15);
    const u32 *cig = reinterpret_cast<const u32 *>(a.blob + 16 * (size_t)rr.off);
    const bool live = j < ns && (rr.flags & GF_PROCESS);
    bool todo = live && blob16 <= (u32)kWarpBuf && 16 * nq <= a.mask_stride;     // gets query masks
    if (live && !todo && span_exceeds(cig, ncig, L, a.extent)) extent_err = true;
    u32 flags = rr.flags;
    while (__any_sync(kFull, todo)) {                        // as many reads per pass as the buffer holds (all 32 at 50 bp)
        const u32 sz = todo ? blob16 : 0u;
        u32 incl = sz;
        for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
        const bool now = todo && incl <= (u32)kWarpBuf;
        const u32 sb = wbuf_addr + incl - sz;
        if (now) for (u32 o = 0; o < blob16; o += 16) cp_async16(sb + o, reinterpret_cast<const uint8_t *>(cig) + o);
        cp_async_wait_all();
        __syncwarp();
        if (now) {
            if (L > a.extent) extent_err = true;
            int span = 0;
            u32 w0 = 0;
            for (int ci = 0; ci < ncig; ci++) {
                const u32 w = smem.ld32(sb + 4 * ci);
                if (ci == 0) w0 = w;
                const int op = w & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += min((int)(w >> 4), kOpCap);
                if (span > a.extent) { extent_err = true; break; }
            }
            flags |= GF_MASKS;
            const int op0 = w0 & 15;                         // one aligned block over all of SEQ: phase B needs no CIGAR walk
            if (ncig == 1 && (op0 == 0 || op0 == 7 || op0 == 8) && (int)(w0 >> 4) >= L) flags |= GF_SIMPLE;
        }
        // the groups of 32 bases of the staged reads are dealt to the lanes (a 150 bp read has five, and only a few such
        // reads fit the buffer at a time): item t = (owner lane, group)
        // (when every read of the warp fits one pass - 32 reads of 50 bp - one read per lane is already dense)
        if ((gs < 32 || __any_sync(kFull, todo && !now)) && !__any_sync(kFull, now && nq > 8)) {   // n_items <= kMaxItems, three bits hold the group
            const u32 ng = now ? (u32)nq : 0u;
            u32 gincl = ng;
            for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(kFull, gincl, o); if (lane >= o) gincl += v; }
            const int n_items = (int)__shfl_sync(kFull, gincl, 31);
            for (u32 g = 0; g < ng; g++) s_items[gincl - ng + g] = (uint8_t)((lane << 3) | (int)g);
            __syncwarp();
            const u32 seq_a = sb + 4 * ncig;
            for (int t0 = 0; t0 < n_items; t0 += 32) {
                const int t = t0 + lane;
                const int item = t < n_items ? s_items[t] : 0;
                const int owner = item >> 3, w = item & 7;
                const u32 o_seq = __shfl_sync(kFull, seq_a, owner);
                const int o_L = __shfl_sync(kFull, L, owner);
                if (t < n_items) {
                    int q_hi = a.dist > 0 ? o_L - a.dist : o_L;
                    if (qg.none) q_hi = q_lo;
                    build_query_mask_group(smem, o_seq, mask_addr + (u32)(j0 + owner) * (u32)a.mask_stride, o_L, w, q_lo, q_hi, qg);
                }
            }
        } else if (now) {                                    // one read per lane
            int q_hi = a.dist > 0 ? L - a.dist : L;
            if (qg.none) q_hi = q_lo;
            build_query_masks(smem, sb + 4 * ncig, mask_addr + (u32)j * (u32)a.mask_stride, L, q_lo, q_hi, qg);
        }
        if (now) todo = false;
        __syncwarp();                                        // the buffer is reused by the next pass
    }
    if (j < ns) { rr.flags = flags; s_rec[j] = rr; }
    // s_first[ch] = first read with pos + extent - 1 >= first position of chunk ch of the segment (reads are sorted by start)
    int pj = 0;
    if (j < nb) pj = max(j < ns ? rr.pos : g_rec[j].pos, -a.extent - 1);
    int pp = __shfl_up_sync(kFull, pj, 1);
    if (j < nb) {
        if (lane == 0 && j > 0) pp = max(g_rec[j - 1].pos, -a.extent - 1);
        int prev = -1;
        if (j > 0) { const int v = pp + a.extent - 1 - seg0; prev = v < 0 ? -1 : min(v >> 5, nseg - 1); }
        const int vj = pj + a.extent - 1 - seg0;
        const int cur = vj < 0 ? -1 : min(vj >> 5, nseg - 1);
        for (int ch = prev + 1; ch <= cur; ch++) s_first[ch] = j;
        if (j == nb - 1) for (int ch = cur + 1; ch < nseg; ch++) s_first[ch] = nb;
    }
}

// Counts of chunk `ch` of the unit from a batch of nb reads (g_rec[0..nb), the first `ns` with a record in
// shared memory, below the plane slots at mask_addr), part `part` of `nparts`: lane = position on return.
// `first` is the first read of the batch that can reach the chunk; `first_batch` is set for the batch that holds
// the unit's leftmost reads.
__device__ __forceinline__ void count_chunk(const PileupArgs &a, const Unit &un, const ReadRec *g_rec,
                                            u32 mask_addr, int first, bool first_batch, int nb, int ns, int ch, int part,
                                            int nparts, int lane, int q_lo, const TransposeConst &tc, u32 (&cnt)[10], bool &extent_err) {
    const SharedMem smem;
    const int c0 = un.t0 + 32 * ch, c1 = c0 + 32;
    const int skip_le = c0 - a.extent;               // reads starting at or before this cannot reach the chunk
    first = min(max(first, 0), nb);                  // (stale table entries if the batch was not sorted: flagged elsewhere)
    for (int r = first + 32 * part; r < nb; r += 32 * nparts) {
        const int j = r + lane;
        ReadRec rr; rr.pos = 0x7fffffff; rr.off = 0; rr.len = 0; rr.flags = 0;
        if (j < ns) {                                        // the records sit below the mask slots
            u32 w[4];
            smem.ld128(mask_addr - (u32)(kStageReads * sizeof(ReadRec)) + 16u * (u32)j, w);
            rr.pos = (int)w[0]; rr.off = w[1]; rr.len = w[2]; rr.flags = w[3];
        }
        else if (j < nb) rr = g_rec[j];
        const int pos = rr.pos;
        const u32 m_after = __ballot_sync(kFull, pos >= c1);
        const bool cand = pos < c1 && pos > skip_le && (rr.flags & GF_PROCESS);   // implies j < nb
        const u32 cands = __ballot_sync(kFull, cand);
        if (cands) {
            const int strand = (rr.flags & GF_STRAND) ? 1 : 0;
            const int L = rr.len & 0xffff;
            const int nq = (L + 31) >> 5;
            const bool simple = (rr.flags & GF_SIMPLE) != 0;
            const u32 mask_s = mask_addr + (u32)j * (u32)a.mask_stride;
            u32 pv = 0u, p0 = 0u, p1 = 0u;                   // planes of this read inside the chunk: valid base, code bits
            u32 m5 = 0u;
            if (cand) {
                // Tn5 site (pileup.py:43-50): reverse = start + len(SEQ) - 1, forward = start
                const int t5 = strand ? pos + L - 1 : pos;
                if ((u32)(t5 - c0) < 32u && t5 < a.P) m5 = 1u << (t5 - c0);
                // one aligned block over all of SEQ: reference position p <-> query base p - pos; the planes are
                // already empty outside the distance-from-end window and beyond SEQ
                if (simple) {
                    u32 wv[3];
                    query_window(smem, mask_s, nq, c0 - pos, wv);
                    pv = wv[0]; p0 = wv[1]; p1 = wv[2];
                }
            }
            if (__any_sync(kFull, cand && !simple)) {        // soft clips, indels, reads without query masks
                if (cand && !simple) {
                    const int ncig = rr.len >> 16;
                    const bool masks = (rr.flags & GF_MASKS) != 0;
                    const u32 *cig = reinterpret_cast<const u32 *>(a.blob + 16 * (size_t)rr.off);
                    // no record in shared memory: the declared extent is verified here, once per unit
                    if (j >= ns && (pos >= c0 || (ch == 0 && first_batch)) && span_exceeds(cig, ncig, L, a.extent)) extent_err = true;
                    // aligned blocks overlapping this chunk (pileup.py:55-95)
                    const int q_hi = a.dist > 0 ? L - a.dist : L;
                    int ref = pos, qp = 0;
                    for (int ci = 0; ci < ncig; ci++) {
                        const u32 w = __ldg(cig + ci);
                        const int op = w & 15;
                        const int n = min((int)(w >> 4), kOpCap);
                        if (op == 0 || op == 7 || op == 8) {             // pileup.py:56
                            const int va = max(q_lo - qp, 0), vb = min(q_hi - qp, n);
                            const int r0 = ref, q00 = qp;
                            ref += n; qp = min(qp + n, kOpCap);          // pileup.py:90-91
                            if (vb > va && r0 + va < c1 && r0 + vb > c0) {
                                const int pa = r0 + va - c0, span = vb - va, q0 = q00 + va;
                                u32 wv[3] = {0u, 0u, 0u};
                                if (masks) {
                                    query_window(smem, mask_s, nq, q0 - pa, wv);
                                    const u32 rm = bit_range(pa, pa + span);
                                    wv[0] &= rm; wv[1] &= rm; wv[2] &= rm;
                                } else {
                                    block_masks_global(reinterpret_cast<const uint8_t *>(cig + ncig), L, pa, span, q0, a.min_baseq, wv);
                                }
                                pv |= wv[0]; p0 |= wv[1]; p1 |= wv[2];
                            }
                        } else if (op == 2 || op == 3) ref += n;         // pileup.py:92-93
                        else if (op == 4) qp = min(qp + n, kOpCap);      // pileup.py:94-95; I, H, P: nothing (sic)
                        if (ref >= c1) break;                            // blocks only move right
                    }
                }
            }
            // lane = read -> lane = position. The four bit matrices of a round (V, B0, B1, Tn5; rows = reads) share one
            // 32x32 transpose when its candidates sit in lanes 0..7 (one matrix per byte), two when they sit in lanes
            // 0..15; forward and reverse reads are counted apart (pileup.py:88)
            u32 rev = __ballot_sync(kFull, cand && strand);
            u32 v, b0, b1, t5;
            if (cands <= 0xffu) {
                const int src = lane & 7;
                const u32 s0 = __shfl_sync(kFull, p0, src), s1 = __shfl_sync(kFull, p1, src), s5 = __shfl_sync(kFull, m5, src);
                quarter_columns(warp_transpose(quarter_row(lane, pv, s0, s1, s5), tc), v, b0, b1, t5);
            } else if (cands <= 0xffffu) {
                const u32 s0 = __shfl_xor_sync(kFull, p0, 16), s5 = __shfl_xor_sync(kFull, m5, 16);
                const u32 ta = warp_transpose(half_row(lane, pv, s0), tc);
                const u32 tb = warp_transpose(half_row(lane, p1, s5), tc);
                half_columns(ta, tb, v, b0, b1, t5);
            } else {
                v = warp_transpose(pv, tc); b0 = warp_transpose(p0, tc); b1 = warp_transpose(p1, tc); t5 = warp_transpose(m5, tc);
            }
            count_columns(v, b0, b1, t5, rev, cnt);
        }
        if (m_after) break;
    }
}

// The counts of a chunk are final: strand-bias filter, coverage, Tn5 gating (pileup.py:128-154), depth
// statistics, saturation (writers.py:205-218) and the one write of the 11 planes.
template <int kPpad>
__device__ __forceinline__ void finish_chunk(const PileupArgs &a, int cell, int c0, int lane, u32 (&cnt)[10],
                                             u64 &sum, u32 &covered, u32 &maxd) {
    const int ppad = kPpad ? kPpad : a.ppad;             // compile-time plane pitch: the 11 stores share one address
    const int p = c0 + lane;
    if (p >= a.P) {                                  // pileup.py:58 end_refpos = min(.., mito_length): padding stays zero
#pragma unroll
        for (int k = 0; k < 10; k++) cnt[k] = 0;
    }
    if (a.accumulate) {                              // streamed batches: raw counts add up in the planes; filters, coverage
        uint16_t *acc = a.planes + (size_t)cell * MGATK_N_PLANES * ppad + p;      // and statistics come in k_stream_finish
        bool sat = false;
#pragma unroll
        for (int pl = 0; pl < 10; pl++) {
            if (cnt[pl]) {
                u32 v = (u32)acc[(size_t)pl * ppad] + cnt[pl];
                if (v > 65535u) { v = 65535u; sat = true; }
                acc[(size_t)pl * ppad] = (uint16_t)v;
            }
        }
        if (sat) atomicOr((u64 *)&a.stats->error_bits, (u64)ERR_SATURATED);
        return;
    }
    if (a.apply_bias) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u32 f = cnt[2 * b], r = cnt[2 * b + 1], t = f + r;
            if (t > 0) {                             // pileup.py:143-148, IEEE double, strict >
                const double bias = (double)max(f, r) / (double)t;
                if (bias > a.max_bias) { cnt[2 * b] = 0; cnt[2 * b + 1] = 0; }
            }
        }
    }
    const u32 cov = ((cnt[0] + cnt[1]) + (cnt[2] + cnt[3])) + ((cnt[4] + cnt[5]) + (cnt[6] + cnt[7]));  // pileup.py:150
    if (cov == 0) { if (!a.raw) { cnt[8] = 0; cnt[9] = 0; } }   // pileup.py:152-153: dropped with its Tn5 counts
    else { sum += cov; covered++; maxd = max(maxd, cov); }
    u32 vals[MGATK_N_PLANES];
#pragma unroll
    for (int k = 0; k < 10; k++) vals[k] = cnt[k];
    vals[10] = cov;
    if (max(max(cov, cnt[8]), cnt[9]) > 65535u) {    // rare: exact value kept aside
#pragma unroll
        for (int pl = 0; pl < MGATK_N_PLANES; pl++) {
            if (vals[pl] > 65535u) {
                const u64 idx = atomicAdd((u64 *)&a.stats->n_overflow, 1ull);
                if ((int64_t)idx < a.ovf_cap) {
                    a.ovf[idx].cell = cell;
                    a.ovf[idx].plane_pos = ((u32)pl << 24) | (u32)p;
                    a.ovf[idx].value = vals[pl];
                } else atomicOr((u64 *)&a.stats->error_bits, (u64)ERR_OVERFLOW_CAP);
                vals[pl] = 65535u;
            }
        }
    }
    uint16_t *out = a.planes + (size_t)cell * MGATK_N_PLANES * ppad + p;
#pragma unroll
    for (int pl = 0; pl < MGATK_N_PLANES; pl++) out[(size_t)pl * ppad] = (uint16_t)vals[pl];
}

// Number of records among g[0..n) (sorted by start) whose start is below `key`: every thread of the CTA counts a
// strided share, one shared-memory counter collects the warp sums.
__device__ __forceinline__ int block_count_below(const ReadRec *g, int n, int key, int *s_count) {
    if (threadIdx.x == 0) *s_count = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < n; i += kThreads) c += g[i].pos < key;
    c = __reduce_add_sync(kFull, c);
    if (lane_id() == 0 && c) atomicAdd(s_count, c);
    __syncthreads();
    const int total = *s_count;
    __syncthreads();                                         // everybody has read it before the counter is reused
    return total;
}

template <int kPpad, bool kSplit>
__global__ void __launch_bounds__(kThreads, 32 / kWarpsPerCta)
k_pileup(PileupArgs a, int batch_reads) {
    __shared__ int s_unit, s_count;
    __shared__ u32 s_acc[kSplitChunks * kAccWords];          // deep units: counts of every chunk, summed over warps and batches
    __shared__ int s_first[kChunkSeg];                       // per chunk of the segment: first read of the batch that can reach it
    __shared__ uint8_t s_items[kWarpsPerCta][kMaxItems];     // phase A: (lane << 3 | group) work items of a warp's staging pass
    extern __shared__ __align__(16) uint8_t dyn[];
    ReadRec *s_rec = reinterpret_cast<ReadRec *>(dyn);                                  // [kStageReads]
    uint8_t *s_mask = dyn + kStageReads * sizeof(ReadRec);                              // [kMaskBytes]
    uint8_t *s_wbuf = s_mask + kMaskBytes;                                              // [warps][kWarpBuf + slack]
    for (int e = threadIdx.x; e < kSplitChunks * kAccWords; e += blockDim.x) s_acc[e] = 0;
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    u32 mask_addr = (u32)__cvta_generic_to_shared(s_mask);
    asm volatile("" : "+r"(mask_addr));                      // kept in a register: recomputing it costs S2R + 4 ALU ops per use
    const u32 wbuf_addr = (u32)__cvta_generic_to_shared(s_wbuf) + (u32)wid * (kWarpBuf + kWarpBufSlack);
    const int n_units = *a.n_units;
    const QualGe qg = make_qual_ge(a.min_baseq);
    const int q_lo = a.dist > 0 ? a.dist : 0;                // pileup.py:67-72
    const TransposeConst tc = make_transpose_const(lane);
#ifdef MGATK_TIMING
    // BAR.SYNC only blocks at the first instruction that needs the barrier: the wait of a barrier shows up in the tick
    // AFTER the next dependent instruction, hence the tick behind the shared-memory read of the unit index
    long long tt[6] = {0, 0, 0, 0, 0, 0}; long long tc0 = clock64(), tc1;
#define TICK(k) { tc1 = clock64(); tt[k] += tc1 - tc0; tc0 = tc1; }
#else
#define TICK(k)
#endif
    int next_unit = 0;
    if (threadIdx.x == 0) next_unit = atomicAdd(a.work_counter, 1);
    for (;;) {
        TICK(5)
        __syncthreads();                                     // previous unit fully consumed (also covers the s_acc init)
        if (threadIdx.x == 0) s_unit = next_unit;
        __syncthreads();
        const int u = s_unit;
        if (u >= n_units) break;
        TICK(0)
        if (threadIdx.x == 0) next_unit = atomicAdd(a.work_counter, 1);   // in flight while this unit is processed
        const Unit un0 = a.units[u];
        if (un0.t1 - un0.t0 < 32) continue;                  // empty tile (its reads belong to the tile before)
        // A tile whose reads do not fit the mask slots (a hot spot inside a wide tile) is walked in sub-tiles, each cut
        // where the slots are full; a sub-tile that cannot be cut any narrower is a deep unit and takes its reads in batches.
        // (k_plan_units sends such tiles to a list of their own, processed by the kSplit instance of this kernel, so the
        // common instance does not carry the registers of the walk)
        const bool split = kSplit && un0.rend - un0.rbeg > a.cap_reads && un0.t1 - un0.t0 > 32 * kSplitChunks;
        int sub_t0 = un0.t0, sub_r = un0.rbeg;
      for (;;) {
        Unit un = un0;
        if (split) {
            if (sub_t0 >= un0.t1) break;
            if (sub_t0 > un0.t0) {
                __syncthreads();                             // the previous sub-tile is consumed
                sub_r += block_count_below(a.recs + sub_r, un0.rend - sub_r, sub_t0 - a.extent + 1, &s_count);
            }
            un.t0 = sub_t0; un.rbeg = sub_r;
            if (un0.rend - sub_r > a.cap_reads) {
                const int cut = min(max(a.recs[sub_r + a.cap_reads].pos, 0), un0.t1) & ~31;     // first read that finds no slot
                un.t1 = cut > sub_t0 ? cut : min(sub_t0 + 32 * kSplitChunks, un0.t1);
                un.rend = sub_r + block_count_below(a.recs + sub_r, un0.rend - sub_r, un.t1, &s_count);
            }
            sub_t0 = un.t1;
        }
        const int n_chunks = (un.t1 - un.t0) >> 5;
        const int n_reads = un.rend - un.rbeg;
        // deep unit (few chunks): the reads come in batches and every chunk's candidates are split over `nparts`
        // warps; partial counts meet in s_acc and are finished after the last batch
        const bool deep = n_chunks <= kSplitChunks;
        int nparts = 1, part_shift = 0;
        if (deep) while (nparts * 2 * n_chunks <= kWarpsPerCta) { nparts *= 2; part_shift++; }
        bool extent_err = false;
        u64 sum = 0; u32 covered = 0, maxd = 0;

        for (int rb = 0; rb == 0 || rb < n_reads; ) {
            const int nb = deep ? min(batch_reads, n_reads - rb) : n_reads;   // reads of this batch
            const int ns = min(nb, a.cap_reads);                              // reads with a record + mask slot in shared memory
            const ReadRec *g_rec = a.recs + un.rbeg + rb;
            for (int cs = 0; cs < n_chunks; cs += kChunkSeg) {                // kChunkSeg chunks per pass (one pass for chrM)
                const int nseg = min(kChunkSeg, n_chunks - cs);
                if (rb > 0 || cs > 0) __syncthreads();       // the previous batch / pass is consumed
                // ---- phase A (the first pass also builds the masks) ----
                TICK(4)
                if (nb == 0) for (int ch = threadIdx.x; ch < nseg; ch += kThreads) s_first[ch] = 0;
                for (int j0 = a.group_reads * wid; j0 < nb; j0 += a.group_reads * kWarpsPerCta)
                    stage_reads(a, g_rec, nb, cs == 0 ? ns : 0, j0, lane, s_rec, mask_addr, wbuf_addr, s_items[wid], s_first, un.t0 + 32 * cs, nseg,
                                q_lo, qg, extent_err, a.group_reads);
                TICK(1)
                __syncthreads();
#ifdef MGATK_TIMING
                if (s_first[0] < 0) break;                   // (never: makes the barrier's wait land in the next tick)
#endif
                TICK(2)
                // ---- phase B: chunks of 32 positions (x parts) dealt to the warps round-robin ----
                for (int item = wid; item < nseg * nparts; item += kWarpsPerCta) {
                    const int chl = item >> part_shift, part = item & (nparts - 1);
                    u32 cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // 8 base x strand counters, Tn5 fwd / rev; lane = position
                    count_chunk(a, un, g_rec, mask_addr, s_first[chl], rb == 0 && un.t0 == un0.t0, nb, ns, cs + chl, part, nparts, lane, q_lo, tc, cnt, extent_err);
                    if (deep) {
#pragma unroll
                        for (int k = 0; k < 10; k++) if (cnt[k]) atomicAdd(&s_acc[chl * kAccWords + k * 32 + lane], cnt[k]);
                    } else {
                        finish_chunk<kPpad>(a, un.cell, un.t0 + 32 * (cs + chl), lane, cnt, sum, covered, maxd);
                    }
                }
                TICK(3)
            }
            rb += nb > 0 ? nb : 1;
        }
        if (deep) {
            __syncthreads();                                 // all partial counts are in s_acc
            if (wid < n_chunks) {
                u32 cnt[10];
#pragma unroll
                for (int k = 0; k < 10; k++) { cnt[k] = s_acc[wid * kAccWords + k * 32 + lane]; s_acc[wid * kAccWords + k * 32 + lane] = 0; }
                finish_chunk<kPpad>(a, un.cell, un.t0 + 32 * wid, lane, cnt, sum, covered, maxd);
            }
        }
        // per-cell depth statistics (processors.py:36-39, writers.py:187-193): hardware warp reductions
        if (__any_sync(kFull, covered != 0)) {
            u64 tot;
            if (__any_sync(kFull, (sum >> 32) != 0)) {       // cannot happen below 2^32 counted bases per lane and unit
                tot = sum;
                for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
            } else {
                const u32 lo = (u32)sum;
                tot = (u64)__reduce_add_sync(kFull, lo & 0xffffu) + ((u64)__reduce_add_sync(kFull, lo >> 16) << 16);
            }
            const u32 cv = __reduce_add_sync(kFull, covered), mx = __reduce_max_sync(kFull, maxd);
            if (lane == 0) {
                atomicAdd((u64 *)&a.qc[un.cell].sum_depth, tot);
                atomicAdd(&a.qc[un.cell].covered, cv);
                atomicMax(&a.qc[un.cell].max_depth, mx);
            }
        }
        if (__any_sync(kFull, extent_err) && lane == 0) atomicOr((u64 *)&a.stats->error_bits, (u64)ERR_EXTENT);
        if (!split) break;
      }
    }
#ifdef MGATK_TIMING
    if (lane == 0) for (int k = 0; k < 6; k++) atomicAdd(&a.dbg[k], (unsigned long long)tt[k]);
#endif
}

// ---------------------------------------------------------------------------------------------
// Reference-allele vote input (writers.py:220-222): per position and base, the sum over cells of
// fwd+rev after filtering. One thread owns two positions and a group of 32 cells.
// ---------------------------------------------------------------------------------------------
constexpr int kTotalsCellGroup = 32;
__global__ void __launch_bounds__(128)
k_base_totals(const uint16_t *__restrict__ planes, int n_cells, int P, int ppad, u64 *__restrict__ totals) {
    const int pp = blockIdx.x * blockDim.x + threadIdx.x;          // position pair
    if (2 * pp >= ppad) return;
    const int c0 = blockIdx.y * kTotalsCellGroup, c1 = min(n_cells, c0 + kTotalsCellGroup);
    u32 s[4][2] = {};
    for (int c = c0; c < c1; c++) {
        const u32 *row = reinterpret_cast<const u32 *>(planes + (size_t)c * MGATK_N_PLANES * ppad) + pp;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u32 f = __ldg(row + (size_t)(2 * b) * (ppad / 2)), r = __ldg(row + (size_t)(2 * b + 1) * (ppad / 2));
            s[b][0] += (f & 0xffff) + (r & 0xffff);
            s[b][1] += (f >> 16) + (r >> 16);
        }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int p = 2 * pp + k;
        if (p < P)
#pragma unroll
            for (int b = 0; b < 4; b++) if (s[b][k]) atomicAdd(&totals[(size_t)p * 4 + b], (u64)s[b][k]);
    }
}

__global__ void k_base_totals_overflow(const mgatk_overflow *__restrict__ ovf, const mgatk_stats *__restrict__ stats,
                                       int64_t cap, int P, u64 *__restrict__ totals) {
    const int64_t n = min((int64_t)stats->n_overflow, cap);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int pl = ovf[k].plane_pos >> 24, p = ovf[k].plane_pos & 0xffffff;
        if (pl < 8 && p < P) atomicAdd(&totals[(size_t)p * 4 + pl / 2], (u64)(ovf[k].value - 65535u));
    }
}

// ---------------------------------------------------------------------------------------------
// PileupGenerator.filter_strand_bias (pileup.py:128-154) as a stand-alone pass over raw planes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_filter_planes(uint16_t *__restrict__ planes, int n_cells, int P, int ppad, double max_bias) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (p >= P || c >= n_cells) return;
    uint16_t *row = planes + (size_t)c * MGATK_N_PLANES * ppad + p;
    u32 cov = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        u32 f = row[(size_t)(2 * b) * ppad], r = row[(size_t)(2 * b + 1) * ppad];
        const u32 t = f + r;
        if (t > 0) {
            const double bias = (double)max(f, r) / (double)t;
            if (bias > max_bias) { f = 0; r = 0; row[(size_t)(2 * b) * ppad] = 0; row[(size_t)(2 * b + 1) * ppad] = 0; }
        }
        cov += f + r;
    }
    row[(size_t)MGATK_PLANE_COVERAGE * ppad] = (uint16_t)min(cov, 65535u);
    if (cov == 0) { row[(size_t)MGATK_PLANE_TN5_FWD * ppad] = 0; row[(size_t)MGATK_PLANE_TN5_REV * ppad] = 0; }
}

// ---------------------------------------------------------------------------------------------
// Streaming (batches cut on reference_start borders, planes resident and accumulating): the per-batch kernels only add
// raw counts; this pass turns the accumulated planes into the final ones exactly as a one-batch run would have written
// them: cell gate (processors.py:22), strand-bias filter, coverage, Tn5 gating (pileup.py:128-154), depth statistics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_stream_finish(PileupArgs a, int n_cells, int min_reads) {
    const int lane = lane_id();
    const int warps = (int)((gridDim.x * (size_t)blockDim.x) >> 5), gw = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int chunks = a.ppad >> 5;
    for (long long w = gw; w < (long long)n_cells * chunks; w += warps) {
        const int cell = (int)(w / chunks), ch = (int)(w - (long long)cell * chunks);
        const bool dead = cell_dead(a.qc[cell], min_reads);
        const uint16_t *in = a.planes + (size_t)cell * MGATK_N_PLANES * a.ppad + 32 * ch + lane;
        u32 cnt[10];
#pragma unroll
        for (int k = 0; k < 10; k++) cnt[k] = dead ? 0u : (u32)in[(size_t)k * a.ppad];
        u64 sum = 0; u32 covered = 0, maxd = 0;
        finish_chunk<0>(a, cell, 32 * ch, lane, cnt, sum, covered, maxd);
        if (__any_sync(kFull, covered != 0)) {
            u64 tot = sum;
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
            const u32 cv = __reduce_add_sync(kFull, covered), mx = __reduce_max_sync(kFull, maxd);
            if (lane == 0) {
                atomicAdd((u64 *)&a.qc[cell].sum_depth, tot);
                atomicAdd(&a.qc[cell].covered, cv);
                atomicMax(&a.qc[cell].max_depth, mx);
            }
        }
    }
}

// k_dedup parks the per-cell count of reads to pile up in median_lo: cleared before every streamed batch
__global__ void k_clear_parked(mgatk_cell_qc *__restrict__ qc, int n_cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cells) qc[c].median_lo = 0;
}

// ---------------------------------------------------------------------------------------------
// Median depth over covered positions (writers.py:190): the two middle order statistics by a
// two-level (high byte, low byte) counting select on the coverage plane. One CTA per cell.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_median(const uint16_t *__restrict__ planes, int P, int ppad, mgatk_cell_qc *__restrict__ qc) {
    __shared__ u32 hist[256];
    __shared__ u32 sel[4];                                   // bin, remainder for lo / hi
    const int c = blockIdx.x, t = threadIdx.x;
    const uint16_t *cov = planes + ((size_t)c * MGATK_N_PLANES + MGATK_PLANE_COVERAGE) * ppad;
    hist[t] = 0;
    __syncthreads();
    for (int p = t; p < P; p += 256) { const u32 v = cov[p]; if (v) atomicAdd(&hist[v >> 8], 1u); }
    __syncthreads();
    u32 res[2] = {0, 0};
    if (t == 0) {
        u32 n = 0;
        for (int b = 0; b < 256; b++) n += hist[b];
        sel[0] = sel[2] = 0xffffffffu;
        if (n) {
            const u32 k[2] = {(n - 1) / 2, n / 2};
            for (int s = 0; s < 2; s++) {
                u32 acc = 0;
                for (int b = 0; b < 256; b++) { if (k[s] < acc + hist[b]) { sel[2 * s] = b; sel[2 * s + 1] = k[s] - acc; break; } acc += hist[b]; }
            }
        }
    }
    __syncthreads();
    if (sel[0] == 0xffffffffu) { if (t == 0) { qc[c].median_lo = 0; qc[c].median_hi = 0; } return; }
    for (int s = 0; s < 2; s++) {
        const u32 bin = sel[2 * s], rem = sel[2 * s + 1];
        if (s == 1 && bin == sel[0]) {                       // same high byte: low-byte histogram is still valid
            if (t == 0) { u32 acc = 0; for (int b = 0; b < 256; b++) { if (rem < acc + hist[b]) { res[1] = (bin << 8) | b; break; } acc += hist[b]; } }
            break;
        }
        __syncthreads();
        hist[t] = 0;
        __syncthreads();
        for (int p = t; p < P; p += 256) { const u32 v = cov[p]; if (v && (v >> 8) == bin) atomicAdd(&hist[v & 255], 1u); }
        __syncthreads();
        if (t == 0) { u32 acc = 0; for (int b = 0; b < 256; b++) { if (rem < acc + hist[b]) { res[s] = (bin << 8) | b; break; } acc += hist[b]; } }
    }
    if (t == 0) { qc[c].median_lo = res[0]; qc[c].median_hi = res[1]; }
}

}  // namespace mgatk
