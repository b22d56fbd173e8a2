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
constexpr int kDedupPrivateSteps = 64;                      // long-run mode: look-back steps a record takes on its own before the warp helps
constexpr int kLongRun = 2048;                              // a (cell, start) run of at least twice this many slots is always noticed
constexpr int kDedupCells = 64;                             // consecutive cells of a tile counted in shared memory
constexpr u64 kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62, kScanValue = (1ull << 62) - 1;

struct DedupArgs {
    const uint8_t *slots; uint8_t *out; int slot_bytes;
    const int64_t *m_ptr;
    const u32 *cell_first; int n_first;      // compact slots carry no cell: first slot of every cell (row 0 of the scanned histogram)
    int dedup_mode;
    mgatk_cell_qc *qc; mgatk_stats *stats;
    u32 *ticket; u64 *scan_state; int64_t *n_proc_out;
    const u32 *long_runs;                    // set by k_find_long_runs: some (cell, start) run is longer than kLongRun slots
};

// The rest of the look-backs of a warp's records (`resume` >= 0: where a record's own scan stopped), one pending record
// after the other, 32 predecessors per step. Returns for the calling lane: bit 0 = an earlier record of the run has the
// same strand, bit 1 = ... and the same template length.
// Does the partitioned array hold a very long (cell, start) run? Every kLongRun-th slot is compared with the one kLongRun
// before it: a run of 2 kLongRun slots or more cannot hide. Sets *flag (cleared by the caller).
template <bool kCompact>
__global__ void k_find_long_runs(const uint8_t *__restrict__ slots, int slot_bytes, const int64_t *__restrict__ m_ptr,
                                 const u32 *__restrict__ cell_first, int n_first, u32 *__restrict__ flag,
                                 mgatk_stats *__restrict__ publish, int accumulate) {
    if (publish && blockIdx.x == 0 && threadIdx.x == 0)      // readers.py:111 survivors of stage 1 (streamed batches keep counting)
        publish->stage1_reads = (accumulate ? publish->stage1_reads : 0) + (uint64_t)*m_ptr;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1) * kLongRun;
    if (i >= *m_ptr) return;
    const uint4 a = *reinterpret_cast<const uint4 *>(slots + (size_t)i * slot_bytes);
    const uint4 b = *reinterpret_cast<const uint4 *>(slots + (size_t)(i - kLongRun) * slot_bytes);
    if (a.x != b.x) return;
    if (kCompact) {                                          // same cell: the cell of slot i starts at or before i - kLongRun
        int lo = 0, hi = n_first;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int64_t)__ldg(cell_first + mid) <= i) lo = mid; else hi = mid; }
        if ((int64_t)__ldg(cell_first + lo) > i - kLongRun) return;
    } else if (SlotKey<false>::cell(a) != SlotKey<false>::cell(b)) return;
    *flag = 1u;
}

template <bool kCompact>
__device__ __noinline__ u32 warp_lookback(const uint8_t *slots, size_t sb, u32 pos, u32 tlen, u32 strand, int cell, int64_t first, int64_t resume, int lane) {
    u32 mine = 0u;
    for (u32 pend = __ballot_sync(kFull, resume >= 0); pend; pend &= pend - 1) {
        const int L = __ffs(pend) - 1;
        const u32 kx = __shfl_sync(kFull, pos, L), ky = __shfl_sync(kFull, tlen, L), kstrand = __shfl_sync(kFull, strand, L);
        const int kcell = __shfl_sync(kFull, cell, L);
        const int64_t kfirst = __shfl_sync(kFull, first, L);
        bool any_pos = false, any_len = false;
        for (int64_t base = __shfl_sync(kFull, resume, L);; base -= 32) {
            const int64_t j = base - lane;
            bool stop = true, sp = false, sl = false;
            if (j >= kfirst) {
                const uint4 o = *reinterpret_cast<const uint4 *>(slots + (size_t)j * sb);
                stop = o.x != kx || (!kCompact && SlotKey<false>::cell(o) != kcell);
                sp = !stop && (SlotKey<kCompact>::meta(o) & SM_STRAND) == kstrand;
                sl = sp && o.y == ky;
            }
            const u32 ms = __ballot_sync(kFull, stop), mp = __ballot_sync(kFull, sp), ml = __ballot_sync(kFull, sl);
            const u32 before = ms ? ((1u << (__ffs(ms) - 1)) - 1u) : kFull;      // the run ends at the first stop
            any_pos |= (mp & before) != 0u;
            any_len |= (ml & before) != 0u;
            if (ms || any_len) break;
        }
        if (lane == L) mine = (any_pos ? 1u : 0u) | (any_len ? 2u : 0u);
    }
    return mine;
}

// Resident CTAs per SM the kernel is compiled for: six of 256 threads at 40 registers. Measured on C2 at 4 / 5 / 6 / 8 per
// SM: 0.467 / 0.383 / 0.293 / 0.315 ms (fewer: too little latency hiding; eight at 32 registers: the tiles in flight
// outgrow what the L2 keeps for the copy pass) - profiles/r2_sweep_dedup_ctas_per_sm.log, r2_sweep_occupancy_c2.log.
#ifndef MGATK_DEDUP_CTAS
#define MGATK_DEDUP_CTAS (1536 / MGATK_DEDUP_THREADS)
#endif
// kGuard: the instance for batches with a very long (cell, start) run (bounded private look-back, then the whole warp).
// Both instances are launched; the one whose turn it is not leaves at once (the choice is made on the device by
// k_find_long_runs, no host round trip). Compiled into one kernel behind a run-time flag the guard costs the common path
// 0.035 ms on C2 (dedup 0.315 -> 0.350: its registers), as two instances about 0.01 ms (an empty launch).
template <bool kCompact, bool kGuard>
__global__ void __launch_bounds__(kDedupThreads, MGATK_DEDUP_CTAS)
k_dedup(DedupArgs a) {
    if ((*a.long_runs != 0u) != kGuard) return;
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
    const size_t sb = kCompact ? (size_t)32 : (size_t)a.slot_bytes;
    // the slots are grouped by cell: a tile holds a short run of consecutive cells, counted in shared memory
    int cell0 = 0;
    if (tile0 < m) {
        if (kCompact) {                                      // last cell that starts at or before the tile: 32 probes per step
            int lo = 0, hi = a.n_first;                      // across the warp (3 dependent loads for 2048 cells instead of 11)
            const int ln = threadIdx.x & 31;
            while (hi - lo > 1) {
                const int step = (hi - lo + 31) >> 5, probe = lo + ln * step;
                const bool le = probe < hi && (int64_t)__ldg(a.cell_first + probe) <= tile0;
                const int f = __popc(__ballot_sync(kFull, le));              // a prefix of the lanes (lane 0: always)
                hi = min(lo + f * step, hi);
                lo = lo + (f - 1) * step;
            }
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
    constexpr bool guard = kGuard;
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
        int64_t first = 0;                                   // first slot of this record's cell
        u32 meta = 0u;
        bool len_dup = false, pos_dup = false;
        int64_t resume = -1;                                 // >= 0: the look-back of this record goes on there, by the whole warp
        if (i < m) {
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
            meta = SlotKey<kCompact>::meta(me);
            const u32 strand = meta & SM_STRAND;
            paired = meta & SM_PAIRED;
            if (a.dedup_mode != MGATK_DEDUP_NONE) {
                uint4 o = prev;
                if (!guard) {                                // the common case: every record walks its own (short) run
                    const uint8_t *tile_ptr = a.slots + (size_t)tile0 * sb;      // 32-bit offsets from the tile (may go below it)
                    const int64_t fr = first - tile0;
                    const int first_rel = fr < -0x40000000 ? -0x40000000 : (int)fr;
                    int j = k * kDedupThreads + (int)threadIdx.x - 1;
                    while (j >= first_rel && o.x == me.x && (kCompact || SlotKey<false>::cell(o) == cell)) {   // same cell, same start
                        if ((SlotKey<kCompact>::meta(o) & SM_STRAND) == strand) {
                            pos_dup = true;
                            if (o.y == me.y) { len_dup = true; break; }
                        }
                        if (--j >= first_rel) o = *reinterpret_cast<const uint4 *>(tile_ptr + (ptrdiff_t)j * (ptrdiff_t)sb);
                    }
                } else {                                     // the batch holds a very long run: bounded private scan, then the warp
                    for (int64_t j = i - 1; j >= first; ) {
                        if (o.x != me.x) break;
                        if (!kCompact && SlotKey<false>::cell(o) != cell) break;
                        if ((SlotKey<kCompact>::meta(o) & SM_STRAND) == strand) {
                            pos_dup = true;
                            if (o.y == me.y) { len_dup = true; break; }
                        }
                        if (--j < first) break;
                        if (j < i - kDedupPrivateSteps) { resume = j; break; }
                        o = *reinterpret_cast<const uint4 *>(a.slots + (size_t)j * sb);
                    }
                }
            }
        }
        // Long runs (a hot spot: very many reads of one cell at one start with distinct template lengths): the rest of a
        // record's look-back is walked by the whole warp, 32 predecessors per step - K / 32 steps instead of K for a run of K
        // (K^2 / 2 key loads per run otherwise). Only when k_find_long_runs saw such a run: the guard costs the common path
        // 0.04 ms on C2 when compiled in unconditionally.
        if (guard && __any_sync(kFull, resume >= 0)) {
            const u32 found = warp_lookback<kCompact>(a.slots, sb, me.x, me.y, meta & SM_STRAND, cell, first, resume, lane);
            pos_dup |= (found & 1u) != 0u; len_dup |= (found & 2u) != 0u;
        }
        if (i < m) {
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
