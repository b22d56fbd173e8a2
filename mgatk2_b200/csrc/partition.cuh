// mgatk2_b200 — stage 1 + stage 3 of the hot path in ONE streaming pass over the batch (sm_100a).
//
//   k_hist / k_scan_*        per-chunk histograms of the cell index of the records that pass the flag + whitelist filter
//                            (readers.py:96-111) and their exclusive scan in (cell, chunk) order
//   k_scatter_planes         walks the batch in BAM order. Every warp streams the cigar|seq|qual blobs of its next 32
//                            records into shared memory with ONE bulk asynchronous copy (TMA, mbarrier-signalled, one step
//                            ahead), turns SEQ / QUAL / CIGAR of each surviving record into reference-coordinate bit
//                            planes (pileup.py:52-86: aligned blocks, base quality, distance-from-end window, A/C/G/T only)
//                            and writes the finished slot (common.cuh) at the record's rank inside its cell: a stable
//                            counting partition by cell, so dedup and counting see (cell, start, BAM order) order.
//   k_scatter_slots          second digit pass of the partition for more than 2048 cells (moves finished slots).
#pragma once
#include "common.cuh"

namespace mgatk {

constexpr int kScanGroup = 16;        // chunks per scan group

// ---------------------------------------------------------------------------------------------
// Sources of the partition passes
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
};

struct SrcSlots {         // second pass of a two-digit partition: finished slots, already filtered, count on the device
    const uint8_t *a;
    const int64_t *m;
    int slot_bytes;
    int compact, hi_shift;    // compact slots carry only the high digit of their cell (common.cuh), `hi_shift` = bits of the low one
    static constexpr int vec4 = 0;
    __device__ __forceinline__ void cell4(int64_t, int (&)[4]) const {}
    __device__ __forceinline__ int64_t count() const { return *m; }
    __device__ __forceinline__ int cell(int64_t i) const {
        const u32 *w = reinterpret_cast<const u32 *>(a + (size_t)i * slot_bytes);
        if (compact) return (int)(((w[7] >> 24) | ((w[3] >> 29) << 8)) << hi_shift);
        return (int)(w[2] & 0xffffffu);
    }
};

#ifndef MGATK_PART_THREADS
#define MGATK_PART_THREADS 256
#endif
constexpr int kPartThreads = MGATK_PART_THREADS;  // records per CTA step of the scatter
constexpr int kHistAhead = 4;      // steps loaded before counting (k_hist)

// Per-CTA digit histogram of a contiguous chunk of records (order does not matter for counting).
#ifndef MGATK_HIST_THREADS
#define MGATK_HIST_THREADS 1024
#endif
constexpr int kHistThreads = MGATK_HIST_THREADS;
// `bintot` (may be null): per-bin totals that k_scan_group_sums adds up afterwards - cleared here by the first CTA.
template <class Src>
__global__ void __launch_bounds__(kHistThreads)
k_hist(Src src, int64_t chunk, int nchunks, int shift, int bins, u32 *__restrict__ mat, u32 *__restrict__ bintot) {
    extern __shared__ u32 smem[];
    u32 *h = smem;
    if (bintot && blockIdx.x == 0) for (int b = threadIdx.x; b <= bins; b += kHistThreads) bintot[b] = 0;
    for (int b = threadIdx.x; b < bins; b += kHistThreads) h[b] = 0;
    __syncthreads();
    const int64_t n = src.count();
    int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk;
    if (end > n) end = n;
    if (end < beg) end = beg;                                // a chunk past the end (chunks are rounded up): nothing to count -
                                                             // (end - beg) & ~3 below must not see a negative length
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
// With `bintot` the bin totals are added up here (one atomic per group and bin), the scan over the bins is a single pass
// over `bins` numbers (k_scan_cells) and k_scan_apply sums the groups before its own: the one-CTA pass over the whole
// [groups][bins] matrix (k_scan_bases, 18 us on C2) drops out.
__global__ void k_scan_group_sums(const u32 *__restrict__ mat, int nchunks, int bins, u32 *__restrict__ part, u32 *__restrict__ bintot) {
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
    if (bintot && s) atomicAdd(&bintot[b], s);
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

// `binbase` (may be null): exclusive scan of the bin totals; then part[][] still holds the plain group sums and the base
// of group g is binbase[b] + the sums of the groups before it.
__global__ void k_scan_apply(u32 *__restrict__ mat, int nchunks, int bins, const u32 *__restrict__ part, const u32 *__restrict__ binbase) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (b >= bins) return;
    const int w0 = g * kScanGroup;
    u32 v[kScanGroup];
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) v[k] = w0 + k < nchunks ? mat[(size_t)(w0 + k) * bins + b] : 0u;
    u32 run;
    if (binbase) {
        run = binbase[b];
#pragma unroll 4
        for (int q = 0; q < g; q++) run += part[(size_t)q * bins + b];
    } else run = part[(size_t)g * bins + b];
#pragma unroll
    for (int k = 0; k < kScanGroup; k++) if (w0 + k < nchunks) { mat[(size_t)(w0 + k) * bins + b] = run; run += v[k]; }
}

// Slots per cell and their exclusive scan: where every cell starts in the partitioned array. A one-pass partition has this
// for free (row 0 of its scanned histogram); compact slots of a two-digit partition carry no cell index, so their
// consumer (k_dedup) needs the table. Cells below `smem_cells` are counted in shared memory first.
__global__ void __launch_bounds__(1024)
k_cell_counts(SrcUser src, int smem_cells, u32 *__restrict__ counts) {
    extern __shared__ u32 smem[];
    for (int c = threadIdx.x; c < smem_cells; c += blockDim.x) smem[c] = 0;
    __syncthreads();
    const int64_t n = src.count();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = src.cell(i);
        if (c >= 0) atomicAdd(c < smem_cells ? &smem[c] : &counts[c], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < smem_cells; c += blockDim.x) if (smem[c]) atomicAdd(&counts[c], smem[c]);
}

__global__ void __launch_bounds__(1024)
k_scan_cells(u32 *__restrict__ counts, int n, int64_t *__restrict__ total_out = nullptr) {   // in place, exclusive; counts[n] = total
    __shared__ u32 warp_sums[32];
    __shared__ u32 carry_s;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int c = c0 + t;
        const u32 mine = c < n ? counts[c] : 0u;
        u32 inc = mine;
        for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const u32 v = warp_sums[lane];
            u32 sc = v;
            for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(kFull, sc, o); if (lane >= o) sc += x; }
            warp_sums[lane] = sc - v;
        }
        __syncthreads();
        const u32 carry = carry_s;
        if (c < n) counts[c] = carry + warp_sums[wid] + inc - mine;
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_sums[31] + inc;
        __syncthreads();
    }
    if (t == 0) { counts[n] = carry_s; if (total_out) *total_out = (int64_t)carry_s; }
}

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
template <int kWarps>
__host__ __device__ inline size_t rank_smem_bytes_t(int bins) { return ((size_t)bins * (2 * sizeof(Packed<kWarps>) + 4) + 127) / 128 * 128; }
__host__ __device__ inline size_t rank_smem_bytes(int bins) { return rank_smem_bytes_t<kPartWarps>(bins); }

// Rank of a record inside its digit for one step of kPartThreads records in BAM order, and the digit's running
// destination offset: match_any groups the lanes of a warp by digit, the group leaders add their group size into the
// byte of their warp in a packed per-digit counter (8 warps x 8 bits, two words), and after one barrier every thread
// reads "records of my digit in earlier warps" out of the lower bytes. One shared-memory atomic per (warp, digit) and
// two barriers per step (the packed counters are double buffered). Returns the destination index (d < 0: none).
template <class Pack>
__device__ __forceinline__ size_t rank_step(Pack *packed, u32 *off, int bins, int buf, int d, int lane, int wid) {
    Pack *pk = packed + (size_t)buf * bins;
    const u32 peers = __match_any_sync(kFull, d);
    const bool leader = d >= 0 && lane == __ffs(peers) - 1;
    if (leader) Pack::add(&pk[d], wid, (u32)__popc(peers));
    __syncthreads();
    Pack v; v.clear(); u32 base = 0;
    if (d >= 0) { v = pk[d]; base = off[d]; }
    __syncthreads();
    if (d < 0) return 0;
    if (leader && v.first(wid)) { off[d] = base + v.total(); pk[d].clear(); }   // first warp that holds the digit
    return (size_t)base + v.before(wid) + __popc(peers & ((1u << lane) - 1u));
}

#ifndef MGATK_LUT_PLANES
#define MGATK_LUT_PLANES 1           // 0: the per-nibble logic of round 2's first builder (kept for A/B runs, tools/stage_sweep.py)
#endif
#if MGATK_LUT_PLANES
#define MGATK_MASK_GROUP query_mask_group_lut<kQualAnd>
#else
#define MGATK_MASK_GROUP query_mask_group_straight
#endif
#ifndef MGATK_QUAL_AND
#define MGATK_QUAL_AND 1             // a second instance of the scatter for min_base_quality in [0, 127] (api.cu: launch_scatter)
#endif
#ifndef MGATK_SCATTER_CTAS
#define MGATK_SCATTER_CTAS 3
#endif
constexpr int kWarpBufSlack = 64;    // the plane builder may load this far past the last staged byte

struct ScatterArgs {
    mgatk_batch b;
    int32_t n_cells;
    int64_t chunk; int nchunks, shift, bins;
    const u32 *mat;
    uint8_t *dst;
    u64 *error_bits;
    int words, slot_bytes;            // plane words per read (W), bytes per slot
    int min_baseq, dist, min_mapq, extent;
    int wbuf;                         // bytes of one per-warp blob staging buffer (multiple of 16)
    int hi_shift;                     // two-digit partition: bits of the first digit (compact slots carry cell >> hi_shift); else 31
};

// Wide slots (reads beyond 56 positions) run 128-thread CTAs: their staging buffers (32 blobs of a few hundred bytes per
// warp) and plane scratch would otherwise leave one 256-thread CTA per SM (stress shape: partition 1.95 -> 1.55 ms).
#ifndef MGATK_WIDE_THREADS
#define MGATK_WIDE_THREADS 128
#endif
constexpr int kWideThreads = MGATK_WIDE_THREADS;
template <int kT>
__host__ __device__ inline size_t scatter_smem_bytes(int bins, int wbuf, int words, bool compact) {
    return rank_smem_bytes_t<kT / 32>(bins) + (size_t)(kT / 32) * (wbuf + kWarpBufSlack) + (compact ? 0 : (size_t)kT * 16 * words);
}

struct RawRec { int32_t pos, tlen, bc, prev; u32 off; uint16_t flag, lseq, ncig; uint8_t mapq; bool valid; };

__device__ __forceinline__ RawRec load_raw(const mgatk_batch &b, int64_t i, bool valid) {     // loads only; used one or two steps later
    RawRec r; r.valid = valid;
    r.pos = 0; r.tlen = 0; r.bc = -1; r.off = 0; r.flag = 0x4; r.lseq = 0; r.ncig = 0; r.mapq = 0; r.prev = 0x80000000;
    if (valid) {
        if ((threadIdx.x & 31) == 0 && i > 0) r.prev = b.pos[i - 1];     // sortedness check across the warp border
        r.pos = b.pos[i]; r.tlen = b.tlen[i]; r.bc = b.bc_idx[i]; r.off = b.blob_off[i];
        r.flag = b.flag[i]; r.lseq = b.l_seq[i]; r.ncig = b.n_cigar[i]; r.mapq = b.mapq[i];
    }
    return r;
}

// The blobs of a warp's 32 records of one step lie in [min offset, max end) of the caller's blob - back to back when the
// host packs records in file order. When that range fits the warp's buffer, lane 0 arms the barrier and issues ONE bulk
// copy of it. Returns whether the step is staged; `beg16` = blob offset (16-byte units) of the first byte in the buffer.
__device__ __forceinline__ bool stage_blobs(const mgatk_batch &b, const RawRec &r, int lane, u32 buf_addr, u32 bar, int wbuf, u32 &beg16) {
    const u32 sz16 = (u32)((4 * (int)r.ncig + (((int)r.lseq + 1) >> 1) + (int)r.lseq + 15) >> 4);
    u32 end = r.off + sz16;
    if (end < r.off) end = 0xffffffffu;                      // offsets near 2^32: too large for the buffer below
    const u32 lo = __reduce_min_sync(kFull, r.valid ? r.off : 0xffffffffu), hi = __reduce_max_sync(kFull, r.valid ? end : 0u);
    beg16 = lo;
    const bool ok = hi > lo && hi - lo <= (u32)(wbuf >> 4) && 16ull * hi <= (u64)b.blob_bytes;
    if (ok && lane == 0) {
        fence_proxy_async();                                 // the warp's reads of the buffer (the step before) come first
        mbar_expect_tx(bar, 16u * (hi - lo));
        bulk_load(buf_addr, b.blob + 16 * (size_t)lo, 16u * (hi - lo), bar);
    }
    return ok;
}

// What a slot carries besides its planes
struct SlotHead { u32 pos, tlen_abs, meta, tn5off; int L, ncig, span; bool beyond, simple; };

template <class M>
__device__ __forceinline__ SlotHead slot_head(const ScatterArgs &a, const M &mem, u32 cig_addr, const RawRec &r, bool &extent_err) {
    SlotHead h;
    h.L = r.lseq; h.ncig = r.ncig;
    const int32_t t = r.tlen;
    h.pos = (u32)r.pos;
    h.tlen_abs = t < 0 ? (u32)(-(int64_t)t) : (u32)t;                    // abs(read.template_length), readers.py:124
    h.meta = ((r.flag & 0x10) ? SM_STRAND : 0u) | ((r.flag & 0x1) ? SM_PAIRED : 0u) | ((int)r.mapq >= a.min_mapq ? SM_MAPQ_OK : 0u) |
             (h.L == 0 ? SM_EMPTY : 0u);
    const u32 w0 = h.ncig > 0 ? mem.ld32(cig_addr) : 0u;
    // one aligned block over all of SEQ: reference offset i <-> query base i, the query planes are the slot's planes
    h.simple = h.ncig == 1 && cigar_op_aligned((int)(w0 & 15u)) && (int)(w0 >> 4) >= h.L;
    h.span = h.ncig == 1 ? (cigar_op_aligned((int)(w0 & 15u)) || cigar_op_ref_only((int)(w0 & 15u)) ? ((int)(w0 >> 4) < kOpCap ? (int)(w0 >> 4) : kOpCap) : 0)
                         : cigar_ref_span(mem, cig_addr, h.ncig);
    h.beyond = h.L > a.extent || h.span > a.extent;                        // MGATK_ERR_EXTENT: the slot stays empty
    if (h.beyond) extent_err = true;
    h.tn5off = h.L > 0 ? (u32)(h.L - 1) : 0u;
    return h;
}

// planes of a compact slot (at most 56 reference positions) in registers
template <int kGroups, bool kQualAnd, class M>
__device__ __forceinline__ void compact_planes(const ScatterArgs &a, const M &mem, u32 cig_addr, const SlotHead &h, QualGe qg, u32 (&g)[2][3]) {
    const int q_lo = a.dist > 0 ? a.dist : 0;                             // pileup.py:67-72
    int q_hi = a.dist > 0 ? h.L - a.dist : h.L;
    if (qg.none) q_hi = q_lo;
    if (h.beyond || h.L == 0 || !(h.meta & SM_MAPQ_OK)) {                // (a read below min_mapq only takes part in dedup, Q2)
#pragma unroll
        for (int w = 0; w < 2; w++) { g[w][0] = 0u; g[w][1] = 0u; g[w][2] = 0u; }
        return;
    }
    // straight-line builder (all seven groups, no per-group conditions; what lies outside the window is masked at the end):
    // partition 0.797 -> 0.714 ms on C2 against the form that skips groups outside the window (query_planes56).
    // Planes by table lookup (two PRMTs per eight bases, bitplane.cuh): a third fewer integer instructions per read.
#if MGATK_LUT_PLANES
    query_planes56_lut<kGroups, kQualAnd>(mem, cig_addr + 4u * (u32)h.ncig, h.L, q_lo, q_hi, qg, g);
#else
    query_planes56_straight<kGroups>(mem, cig_addr + 4u * (u32)h.ncig, h.L, q_lo, q_hi, qg, g);
#endif
    if (!h.simple) {
        QueryPlanes64 q;
        q.v = ((u64)g[1][0] << 32) | g[0][0]; q.b0 = ((u64)g[1][1] << 32) | g[0][1]; q.b1 = ((u64)g[1][2] << 32) | g[0][2];
#pragma unroll
        for (int k = 0; k < 2; k++) ref_group(mem, cig_addr, h.ncig, q, k, g[k]);
    }
}

__device__ __forceinline__ void store_compact(uint8_t *dst, const SlotHead &h, const u32 (&g)[2][3], u32 hi) {     // one 256-bit store: a whole sector
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(h.pos), "r"(h.tlen_abs), "r"(g[0][0]),
                 "r"((g[1][0] & 0xffffffu) | (h.meta << 24) | ((hi >> 8) << 29)), "r"(g[0][1]),
                 "r"((g[1][1] & 0xffffffu) | ((h.tn5off & 63u) << 24)), "r"(g[0][2]), "r"((g[1][2] & 0xffffffu) | ((hi & 255u) << 24)) : "memory");
}

// 16-byte pieces of a wide slot leave in pairs: one 256-bit store per 32-byte sector of the slot
struct SectorWriter {
    uint8_t *dst; uint4 held; bool have;
    __device__ __forceinline__ void put(const uint4 &v) {
        if (!have) { held = v; have = true; return; }
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(held.x), "r"(held.y), "r"(held.z), "r"(held.w),
                     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        dst += 32; have = false;
    }
    __device__ __forceinline__ void finish() { if (have) put(make_uint4(0u, 0u, 0u, 0u)); }
};

// a wide slot, written sector by sector
template <bool kQualAnd, class M>
__device__ __forceinline__ void store_wide(const ScatterArgs &a, const M &mem, u32 cig_addr, const SlotHead &h, const RawRec &r, int cell,
                                           uint8_t *dst, QualGe qg, u32 scratch_addr) {
    const int q_lo = a.dist > 0 ? a.dist : 0;
    int q_hi = a.dist > 0 ? h.L - a.dist : h.L;
    if (qg.none) q_hi = q_lo;
    const int W = a.words, L = h.L;
    const bool indirect = L > 32 * W || h.span > 32 * W;
    const u32 seq_addr = cig_addr + 4u * (u32)h.ncig;
    SectorWriter out{dst, make_uint4(0u, 0u, 0u, 0u), false};
    out.put(make_uint4(h.pos, h.tlen_abs, (u32)cell | ((h.meta | (indirect ? SM_INDIRECT : 0u)) << 24), h.tn5off & 0xffffu));
    const int nq = (L + 31) >> 5;
    if (indirect || L == 0 || h.beyond || !(h.meta & SM_MAPQ_OK)) {      // (a read below min_mapq only takes part in dedup, Q2)
        out.put(indirect ? make_uint4(r.off, (u32)L | ((u32)h.ncig << 16), 0u, 0u) : make_uint4(0u, 0u, 0u, 0u));
        for (int k = 1; k < W; k++) out.put(make_uint4(0u, 0u, 0u, 0u));
    } else if (h.simple) {
        for (int k = 0; k < W; k++) {
            u32 v = 0u, b0 = 0u, b1 = 0u;
            if (k < nq) MGATK_MASK_GROUP(mem, seq_addr, L, k, q_lo, q_hi, qg, v, b0, b1);
            out.put(make_uint4(v, b0, b1, 0u));
        }
    } else {
        const SharedMem smem;
        for (int w = 0; w < nq; w++) {
            u32 v, b0, b1;
            MGATK_MASK_GROUP(mem, seq_addr, L, w, q_lo, q_hi, qg, v, b0, b1);
            smem.st128(scratch_addr + 16u * (u32)w, v, b0, b1, 0u);
        }
        const QueryPlanesMem<SharedMem> q{smem, scratch_addr, nq};
        for (int k = 0; k < W; k++) {
            u32 o[3];
            ref_group(mem, cig_addr, h.ncig, q, k, o);
            out.put(make_uint4(o[0], o[1], o[2], 0u));
        }
    }
    out.finish();
}

// One step = kPartThreads records in BAM order, one per thread. Per warp: the blobs of the step sit in the warp's
// staging buffer (bulk copy issued one step earlier); compact slots are built in registers straight away, the buffer is
// handed back to the TMA unit for the next step, and only then the CTA ranks the step (two barriers) and stores.
// kGroups (compact slots): groups of eight query bases the plane builder computes = ceil(longest window end / 8), where
// no window of the batch ends beyond max_read_extent - min_distance_from_end (5, 6 or 7; wide slots: unused)
// kQualAnd: min_base_quality lies in [0, 127] (every real run), the quality test needs no mode select (bitplane.cuh)
template <bool kCompact, int kPartThreads, int kGroups, bool kQualAnd>
__global__ void __launch_bounds__(kPartThreads, MGATK_SCATTER_CTAS)
k_scatter_planes(ScatterArgs a) {
    constexpr int kPartWarps = kPartThreads / 32;
    typedef Packed<kPartWarps> Pack;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) u64 s_bar[kPartWarps];
    const int bins = a.bins;
    Pack *packed = reinterpret_cast<Pack *>(smem_raw);                 // [2][bins] per-warp byte counters of the step
    u32 *off = reinterpret_cast<u32 *>(smem_raw + (size_t)2 * bins * sizeof(Pack));   // [bins] running destination offsets
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const u32 stage_stride = (u32)(a.wbuf + kWarpBufSlack);
    const u32 wbuf_addr = (u32)__cvta_generic_to_shared(smem_raw + rank_smem_bytes_t<kPartWarps>(bins)) + (u32)wid * stage_stride;
    const u32 scratch_addr = kCompact ? 0u
        : (u32)__cvta_generic_to_shared(smem_raw + rank_smem_bytes_t<kPartWarps>(bins) + (size_t)kPartWarps * stage_stride) + (u32)t * 16u * (u32)a.words;
    const u32 bar_addr = (u32)__cvta_generic_to_shared(&s_bar[wid]);
    const u32 *row = a.mat + (size_t)blockIdx.x * bins;
    for (int b = t; b < bins; b += kPartThreads) { off[b] = row[b]; packed[b].clear(); packed[bins + b].clear(); }
    if (lane == 0) { mbar_init(bar_addr, 1); mbar_fence_init(); }
    const int64_t n = a.b.n_records;
    int64_t beg = (int64_t)blockIdx.x * a.chunk, end = beg + a.chunk;
    if (end > n) end = n;
    const QualGe qg = make_qual_ge(a.min_baseq);
    RawRec r1 = load_raw(a.b, beg + t, beg + t < end);                                  // two steps of loads in flight
    RawRec r2 = load_raw(a.b, beg + kPartThreads + t, beg + kPartThreads + t < end);
    __syncthreads();
    u32 beg16 = 0u, parity = 0u;
    bool have = stage_blobs(a.b, r1, lane, wbuf_addr, bar_addr, a.wbuf, beg16);         // the buffer holds (will hold) the current step
    int buf = 0;
    bool unsorted = false, extent_err = false;
    for (int64_t i0 = beg; i0 < end; i0 += kPartThreads, buf ^= 1) {
        {   // records must come sorted by reference_start (coordinate-sorted BAM): compare with the record before
            int32_t before = __shfl_up_sync(kFull, r1.valid ? r1.pos : 0x7fffffff, 1);
            if (lane == 0) before = r1.prev;
            unsorted |= r1.valid && r1.pos < before;
        }
        const RawRec cur = r1;
        r1 = r2;
        r2 = load_raw(a.b, i0 + 2 * kPartThreads + t, i0 + 2 * kPartThreads + t < end);
        // readers.py:96-97 unmapped / secondary / supplementary; :104-111 tag absent or not whitelisted
        const int cell = (!cur.valid || (cur.flag & 0x904) || cur.bc < 0 || cur.bc >= a.n_cells) ? -1 : cur.bc;
        const int d = cell >= 0 ? (cell >> a.shift) & (bins - 1) : -1;
        auto global_blob = [&]() {                           // the record's blob in the caller's memory (when the step is not staged)
            const u64 byte_off = 16ull * cur.off;
            const u64 left = (u64)a.b.blob_bytes > byte_off ? (u64)a.b.blob_bytes - byte_off : 0ull;
            return GlobalBlob{a.b.blob + byte_off, left > 0xffffffffull ? 0xffffffffu : (u32)left};
        };
        if (kCompact) {
            SlotHead h;
            u32 g[2][3];
            if (have) {
                const u32 tok = mbar_wait_token(bar_addr, parity);
                parity ^= 1u;
                if (d >= 0) {
                    const SharedMemPure smem;
                    const u32 sb = wbuf_addr + 16u * (cur.off - beg16) + tok;
                    h = slot_head(a, smem, sb, cur, extent_err);
                    compact_planes<kGroups, kQualAnd>(a, smem, sb, h, qg, g);
                }
            } else if (d >= 0) {
                const GlobalBlob gb = global_blob();
                h = slot_head(a, gb, 0u, cur, extent_err);
                compact_planes<kGroups, kQualAnd>(a, gb, 0u, h, qg, g);
            }
            __syncwarp();                                    // every lane has read the buffer: it goes back to the TMA unit
            have = stage_blobs(a.b, r1, lane, wbuf_addr, bar_addr, a.wbuf, beg16);
            const size_t dd = rank_step(packed, off, bins, buf, d, lane, wid);
            if (d >= 0) store_compact(a.dst + dd * 32, h, g, (u32)cell >> a.hi_shift);
        } else {
            const size_t dd = rank_step(packed, off, bins, buf, d, lane, wid);
            u32 tok = 0u;
            if (have) { tok = mbar_wait_token(bar_addr, parity); parity ^= 1u; }
            if (d >= 0) {
                uint8_t *dst = a.dst + dd * (size_t)a.slot_bytes;
                if (have) {
                    const SharedMem smem;
                    const u32 sb = wbuf_addr + 16u * (cur.off - beg16) + tok;
                    const SlotHead h = slot_head(a, smem, sb, cur, extent_err);
                    store_wide<kQualAnd>(a, smem, sb, h, cur, cell, dst, qg, scratch_addr);
                } else {
                    const GlobalBlob gb = global_blob();
                    const SlotHead h = slot_head(a, gb, 0u, cur, extent_err);
                    store_wide<kQualAnd>(a, gb, 0u, h, cur, cell, dst, qg, scratch_addr);
                }
            }
            __syncwarp();
            have = stage_blobs(a.b, r1, lane, wbuf_addr, bar_addr, a.wbuf, beg16);
        }
    }
    if (unsorted) atomicOr(a.error_bits, (u64)ERR_UNSORTED);
    if (extent_err) atomicOr(a.error_bits, (u64)ERR_EXTENT);
}

// Second digit pass (more than 2048 cells): finished wide slots move to their rank inside the high digit.
__global__ void __launch_bounds__(kPartThreads, 4)
k_scatter_slots(SrcSlots src, int64_t chunk, int nchunks, int shift, int bins, const u32 *__restrict__ mat, uint8_t *__restrict__ dst) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Pack *packed = reinterpret_cast<Pack *>(smem_raw);
    u32 *off = reinterpret_cast<u32 *>(smem_raw + (size_t)2 * bins * sizeof(Pack));
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const u32 *row = mat + (size_t)blockIdx.x * bins;
    for (int b = t; b < bins; b += kPartThreads) { off[b] = row[b]; packed[b].clear(); packed[bins + b].clear(); }
    const int64_t n = src.count();
    int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk;
    if (end > n) end = n;
    const int q = src.slot_bytes >> 4;
    __syncthreads();
    int buf = 0;
    for (int64_t i0 = beg; i0 < end; i0 += kPartThreads, buf ^= 1) {
        const int64_t i = i0 + t;
        const uint4 *in = reinterpret_cast<const uint4 *>(src.a + (size_t)i * src.slot_bytes);
        uint4 head = make_uint4(0u, 0u, 0u, 0u);
        if (i < end) head = in[0];
        const int d = i < end ? (src.cell(i) >> shift) & (bins - 1) : -1;
        const size_t dd = rank_step(packed, off, bins, buf, d, lane, wid);
        if (d >= 0) {                                        // whole sectors: 256-bit stores
            uint8_t *out = dst + dd * (size_t)src.slot_bytes;
            uint4 lo = head;
            for (int k = 1; k < q; k += 2) {
                const uint4 hi = in[k];
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + 16 * (k - 1)), "r"(lo.x), "r"(lo.y), "r"(lo.z),
                             "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                if (k + 1 < q) lo = in[k + 1];
            }
        }
    }
}

}  // namespace mgatk
