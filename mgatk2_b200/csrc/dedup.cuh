// mgatk2_b200 — stage 2: dedup on the partitioned slots (readers.py:118-150) and compaction of the reads to pile up.
//
// Inside a cell the slots are sorted by start, so all candidates for a duplicate of record i sit directly before it in
// the same (cell, start) run; the first record of a key in BAM order survives (readers.py:129-150). Both key sets are
// evaluated for every stage-1 survivor (readers.py:128-144) so both duplicate counters are exact whichever strategy is
// selected. The mapq gate comes AFTER the survivor is chosen (pileup.py:33-34: a low-mapq first read still shadows its
// duplicates). Survivors that will be piled up are copied, compacted and in order, to the array k_pileup reads:
// one pass, block-wise exclusive scan chained over blocks by decoupled look-back (blocks take their index from a ticket,
// so a block only ever waits for blocks that are already running).
#pragma once
#include "common.cuh"

namespace mgatk {

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

struct DedupArgs {
    const uint8_t *slots; uint8_t *out; int slot_bytes;
    const int64_t *m_ptr;
    const u32 *cell_first; int n_first;      // compact slots carry no cell: first slot of every cell (row 0 of the scanned histogram)
    int dedup_mode;
    mgatk_cell_qc *qc; mgatk_stats *stats;
    u32 *ticket; u64 *scan_state; int64_t *n_proc_out;
};

template <bool kCompact>
__global__ void __launch_bounds__(kDedupThreads, 1536 / kDedupThreads)
k_dedup(DedupArgs a) {
    __shared__ u32 s_cnt[4];
    __shared__ u32 s_warp[kDedupRounds][kDedupThreads / 32];
    __shared__ u32 s_blk;
    __shared__ u64 s_prefix;
    __shared__ u32 s_cell[kDedupCells][3];                   // per-cell counts of the tile's first cells: one global atomic per (tile, cell)
    __shared__ u32 s_first[kDedupCells + 1];                 // compact: first slot of cells cell0 .. cell0 + 64
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    for (int e = threadIdx.x; e < kDedupCells * 3; e += kDedupThreads) (&s_cell[0][0])[e] = 0;
    if (threadIdx.x == 0) s_blk = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 blk = s_blk;
    const int64_t m = *a.m_ptr;
    const int64_t tile0 = (int64_t)blk * kDedupTile;
    const size_t sb = (size_t)a.slot_bytes;
    // the slots are grouped by cell: a tile holds a short run of consecutive cells, counted in shared memory
    int cell0 = 0;
    if (tile0 < m) {
        if (kCompact) {                                      // last cell that starts at or before the tile
            int lo = 0, hi = a.n_first;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int64_t)__ldg(a.cell_first + mid) <= tile0) lo = mid; else hi = mid; }
            cell0 = lo;
        } else {
            cell0 = SlotKey<false>::cell(*reinterpret_cast<const uint4 *>(a.slots + (size_t)tile0 * sb));
        }
    }
    if (kCompact) {
        for (int e = threadIdx.x; e <= kDedupCells; e += kDedupThreads)
            s_first[e] = cell0 + e < a.n_first ? __ldg(a.cell_first + cell0 + e) : 0xffffffffu;
        __syncthreads();
    }
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    u32 pm[kDedupRounds];
    u32 n_keep = 0, n_len = 0, n_pos = 0, n_empty = 0;
#pragma unroll
    for (int k = 0; k < kDedupRounds; k++) {
        const int64_t i = tile0 + k * kDedupThreads + threadIdx.x;
        int cell = -1;
        bool keep = false, paired = false, process = false;
        uint4 me = make_uint4(0u, 0u, 0u, 0u);                               // pos, |tlen|, then the word with the flags
        if (i < m) me = *reinterpret_cast<const uint4 *>(a.slots + (size_t)i * sb);
        // the record before this one is the neighbouring lane's (lane 0 fetches it)
        uint4 prev;
        prev.x = __shfl_up_sync(kFull, me.x, 1); prev.y = __shfl_up_sync(kFull, me.y, 1);
        prev.z = __shfl_up_sync(kFull, me.z, 1); prev.w = __shfl_up_sync(kFull, me.w, 1);
        if (lane == 0 && i > 0 && i < m) prev = *reinterpret_cast<const uint4 *>(a.slots + (size_t)(i - 1) * sb);
        if (i < m) {
            int64_t first = 0;                               // first slot of this record's cell
            if (kCompact) {
                int e = 0;
                while (e < kDedupCells && (int64_t)s_first[e + 1] <= i) e++;
                if (e < kDedupCells) { cell = cell0 + e; first = s_first[e]; }
                else {                                       // a run of empty cells: search the whole table
                    int lo = cell0 + e, hi = a.n_first;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int64_t)__ldg(a.cell_first + mid) <= i) lo = mid; else hi = mid; }
                    cell = lo; first = __ldg(a.cell_first + lo);
                }
            } else cell = SlotKey<false>::cell(me);
            const u32 meta = SlotKey<kCompact>::meta(me);
            const u32 strand = meta & SM_STRAND;
            paired = meta & SM_PAIRED;
            bool len_dup = false, pos_dup = false;
            if (a.dedup_mode != MGATK_DEDUP_NONE) {
                uint4 o = prev;
                for (int64_t j = i - 1; j >= first; ) {
                    if (o.x != me.x) break;                                  // another start
                    if (!kCompact && SlotKey<false>::cell(o) != cell) break;
                    if ((SlotKey<kCompact>::meta(o) & SM_STRAND) == strand) {
                        pos_dup = true;
                        if (o.y == me.y) { len_dup = true; break; }
                    }
                    if (--j >= first) o = *reinterpret_cast<const uint4 *>(a.slots + (size_t)j * sb);
                }
            }
            keep = a.dedup_mode == MGATK_DEDUP_FRAGMENT_LENGTH ? !len_dup : a.dedup_mode == MGATK_DEDUP_POSITION_ONLY ? !pos_dup : true;
            // pileup.py:33-34 mapq gate (after dedup, Q2). An empty SEQ makes the reference raise
            // (readers.py:157); such survivors are reported in stats.n_empty_seq and not piled up.
            process = keep && (meta & SM_MAPQ_OK) && !(meta & SM_EMPTY);
            n_len += len_dup; n_pos += pos_dup; n_keep += keep; n_empty += keep && (meta & SM_EMPTY);
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
                if (nk) atomicAdd(&a.qc[cell].n_reads, nk);
                if (np) atomicAdd(&a.qc[cell].n_paired, np);
                if (npr) atomicAdd(&a.qc[cell].median_lo, npr);
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
    // per-cell survivors (processors.py:33-34) and reads to pile up per cell (the cell borders of the compacted slots;
    // parked in median_lo, which shares a sector with the two counters and is only written by k_median at the very end)
    for (int e = threadIdx.x; e < kDedupCells; e += kDedupThreads) {
        if (s_cell[e][0]) atomicAdd(&a.qc[cell0 + e].n_reads, s_cell[e][0]);
        if (s_cell[e][1]) atomicAdd(&a.qc[cell0 + e].n_paired, s_cell[e][1]);
        if (s_cell[e][2]) atomicAdd(&a.qc[cell0 + e].median_lo, s_cell[e][2]);
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
        if (lane == 0) atomicExch((unsigned long long *)&a.scan_state[blk], (blk == 0 ? kScanPrefix : kScanAggregate) | (u64)total);
        for (int64_t top = (int64_t)blk - 1; top >= 0; top -= 128) {
            u64 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {                    // lane-major: lane l looks at top - 4l - k
                const int64_t b = top - 4 * lane - k;
                v[k] = kScanPrefix;                          // below block 0: an empty prefix
                if (b >= 0) v[k] = *(volatile u64 *)&a.scan_state[b];
            }
            bool stop_here = false;                          // this lane holds the nearest full prefix
            u64 add = 0;
            bool seen = false;                               // a full prefix was seen at a nearer block of this lane
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t b = top - 4 * lane - k;
                while (b >= 0 && (v[k] & ~kScanValue) == 0) v[k] = *(volatile u64 *)&a.scan_state[b];
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
            if (blk != 0) atomicExch((unsigned long long *)&a.scan_state[blk], kScanPrefix | (prefix + total));
            s_prefix = prefix;
            if ((int64_t)(blk + 1) * kDedupTile >= m && (int64_t)blk * kDedupTile < (m > 0 ? m : 1)) *a.n_proc_out = (int64_t)(prefix + total);
            if (s_cnt[0]) atomicAdd((u64 *)&a.stats->filtered_reads, (u64)s_cnt[0]);
            if (s_cnt[1]) atomicAdd((u64 *)&a.stats->dup_with_length, (u64)s_cnt[1]);
            if (s_cnt[2]) atomicAdd((u64 *)&a.stats->dup_position_only, (u64)s_cnt[2]);
            if (s_cnt[3]) atomicAdd((u64 *)&a.stats->n_empty_seq, (u64)s_cnt[3]);
        }
    }
    __syncthreads();
    const u64 prefix = s_prefix;
    const int q16 = a.slot_bytes >> 4;
#pragma unroll
    for (int k = 0; k < kDedupRounds; k++)
        if ((pm[k] >> lane) & 1u) {
            const int64_t i = tile0 + k * kDedupThreads + threadIdx.x;
            const size_t rank = (size_t)(prefix + s_warp[k][wid] + __popc(pm[k] & ((1u << lane) - 1u)));
            const uint4 *in = reinterpret_cast<const uint4 *>(a.slots + (size_t)i * sb);
            uint8_t *dst = a.out + rank * sb;
            if (kCompact) {                                  // one 256-bit store per slot (a whole sector)
                const uint4 lo = in[0], hi = in[1];
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
                             "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
            } else {
                for (int w = 0; w < q16; w++) reinterpret_cast<uint4 *>(dst)[w] = in[w];
            }
        }
}

}  // namespace mgatk
