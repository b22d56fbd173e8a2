// mgatk2_b200 — sm_100a kernels of the per-cell chrM pileup hot path.
//
// Pipeline (one stream, no host synchronisation between stages):
//   1  k_hist / k_scan* / k_scatter   stage 1 (flag + whitelist filter, readers.py:96-111) fused into a
//                                     stable counting partition of the coordinate-sorted records by cell
//   2  k_cell_start, k_dedup          stage 2 (dedup, readers.py:118-150) on (cell,start) runs, per-cell
//                                     n_reads / n_paired (processors.py:33-34), global counters (readers.py:193-199)
//   3  k_plan*                        cut every cell into position tiles ("units") of bounded read count
//   4  k_pileup                       stages 3-6: CIGAR walk, base-quality / distance-from-end masks, per-base
//                                     per-strand counting and Tn5 sites (pileup.py:32-95) in a per-warp
//                                     shared-memory position ring; strand-bias filter, coverage, Tn5 gating
//                                     (pileup.py:128-154) and depth statistics at flush; planes written once
//   5  k_base_totals, k_median        reference-allele vote input and median depth (writers.py:187-197,220-222)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgatk2_b200.h"

namespace mgatk {

typedef unsigned int u32;
typedef unsigned long long u64;

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kScanGroup = 64;        // chunks per scan group
constexpr u32 kFull = 0xffffffffu;

constexpr u32 ERR_UNSORTED = 1, ERR_EXTENT = 2, ERR_OVERFLOW_CAP = 4;

// gflags bits written by k_dedup, read by k_pileup
constexpr int GF_PROCESS = 1, GF_STRAND = 2, GF_KEEP = 4;
// gmq layout: mapq | strand<<8 | paired<<9
constexpr int GMQ_STRAND = 0x100, GMQ_PAIRED = 0x200;

struct Grouped {          // records that passed stage 1, grouped by cell, BAM order inside a cell
    int32_t *cell;
    int32_t *pos;
    u32 *tlen;            // |template_length|
    u32 *off;             // blob offset, 16-byte units
    u32 *len;             // l_seq | n_cigar<<16
    uint16_t *mq;
};

struct Unit { int32_t cell, t0, t1, rbeg, rend; };

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// Stable counting partition by a digit of the cell index.
// Each warp owns a contiguous chunk of records and a private histogram in shared memory, so ranks
// follow record order without atomics: match_any groups the lanes of one 32-record step by digit,
// the lowest lane of each group bumps the private counter.
// ---------------------------------------------------------------------------------------------
struct SrcUser {          // pass 0: reads the caller's SoA batch and applies the stage-1 filter
    mgatk_batch b;
    int32_t n_cells;
    __device__ __forceinline__ int64_t count() const { return b.n_records; }
    __device__ __forceinline__ int cell(int64_t i) const {
        // readers.py:96-97 unmapped/secondary/supplementary; :104-111 tag absent or not whitelisted
        if (b.flag[i] & 0x904) return -1;
        int c = b.bc_idx[i];
        return (c < 0 || c >= n_cells) ? -1 : c;
    }
    __device__ __forceinline__ void emit(int64_t i, int c, const Grouped &g, int64_t d) const {
        int32_t t = b.tlen[i];
        uint16_t f = b.flag[i];
        g.cell[d] = c;
        g.pos[d] = b.pos[i];
        g.tlen[d] = t < 0 ? (u32)(-(int64_t)t) : (u32)t;     // abs(read.template_length), readers.py:124
        g.off[d] = b.blob_off[i];
        g.len[d] = (u32)b.l_seq[i] | ((u32)b.n_cigar[i] << 16);
        g.mq[d] = (uint16_t)(b.mapq[i] | ((f & 0x10) ? GMQ_STRAND : 0) | ((f & 0x1) ? GMQ_PAIRED : 0));
    }
};

struct SrcGrouped {       // pass 1 of a two-digit partition: already filtered, count on the device
    Grouped a;
    const int64_t *m;
    __device__ __forceinline__ int64_t count() const { return *m; }
    __device__ __forceinline__ int cell(int64_t i) const { return a.cell[i]; }
    __device__ __forceinline__ void emit(int64_t i, int c, const Grouped &g, int64_t d) const {
        g.cell[d] = c; g.pos[d] = a.pos[i]; g.tlen[d] = a.tlen[i];
        g.off[d] = a.off[i]; g.len[d] = a.len[i]; g.mq[d] = a.mq[i];
    }
};

template <class Src>
__global__ void __launch_bounds__(kThreads)
k_hist(Src src, int64_t chunk, int nchunks, int shift, int bins, u32 *__restrict__ mat,
       const int32_t *__restrict__ sorted_check_pos, u64 *__restrict__ error_bits) {
    extern __shared__ u32 smem[];
    const int w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= nchunks) return;
    const int lane = lane_id();
    u32 *h = smem + (threadIdx.x >> 5) * bins;
    for (int b = lane; b < bins; b += 32) h[b] = 0;
    __syncwarp();
    const int64_t n = src.count();
    int64_t beg = (int64_t)w * chunk, end = beg + chunk;
    if (end > n) end = n;
    bool unsorted = false;
    for (int64_t i0 = beg; i0 < end; i0 += 32) {
        const int64_t i = i0 + lane;
        int d = -1;
        if (i < end) {
            int c = src.cell(i);
            if (c >= 0) d = (c >> shift) & (bins - 1);
            if (sorted_check_pos && i > 0 && sorted_check_pos[i] < sorted_check_pos[i - 1]) unsorted = true;
        }
        const u32 peers = __match_any_sync(kFull, d);
        if (d >= 0 && lane == __ffs(peers) - 1) h[d] += __popc(peers);
        __syncwarp();
    }
    if (unsorted) atomicOr(error_bits, (u64)ERR_UNSORTED);
    u32 *row = mat + (size_t)w * bins;
    for (int b = lane; b < bins; b += 32) row[b] = h[b];
}

// scan of mat[chunk][bin] in (bin, chunk) order: S1 group sums, S2 bases, S3 in-place exclusive prefixes
__global__ void k_scan_group_sums(const u32 *__restrict__ mat, int nchunks, int bins, u32 *__restrict__ part) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (b >= bins) return;
    const int w0 = g * kScanGroup, w1 = min(nchunks, w0 + kScanGroup);
    u32 s = 0;
    for (int w = w0; w < w1; w++) s += mat[(size_t)w * bins + b];
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
        if (b < bins) for (int g = 0; g < ngroups; g++) tot += part[(size_t)g * bins + b];
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
        if (b < bins) for (int g = 0; g < ngroups; g++) {
            u32 v = part[(size_t)g * bins + b]; part[(size_t)g * bins + b] = run; run += v;
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
    const int w0 = g * kScanGroup, w1 = min(nchunks, w0 + kScanGroup);
    u32 run = part[(size_t)g * bins + b];
    for (int w = w0; w < w1; w++) { u32 v = mat[(size_t)w * bins + b]; mat[(size_t)w * bins + b] = run; run += v; }
}

template <class Src>
__global__ void __launch_bounds__(kThreads)
k_scatter(Src src, int64_t chunk, int nchunks, int shift, int bins, const u32 *__restrict__ mat, Grouped dst) {
    extern __shared__ u32 smem[];
    const int w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= nchunks) return;
    const int lane = lane_id();
    u32 *off = smem + (threadIdx.x >> 5) * bins;
    const u32 *row = mat + (size_t)w * bins;
    for (int b = lane; b < bins; b += 32) off[b] = row[b];
    __syncwarp();
    const int64_t n = src.count();
    int64_t beg = (int64_t)w * chunk, end = beg + chunk;
    if (end > n) end = n;
    const u32 lt = (1u << lane) - 1;
    for (int64_t i0 = beg; i0 < end; i0 += 32) {
        const int64_t i = i0 + lane;
        int d = -1, c = -1;
        if (i < end) { c = src.cell(i); if (c >= 0) d = (c >> shift) & (bins - 1); }
        const u32 peers = __match_any_sync(kFull, d);
        u32 base = 0;
        if (d >= 0) base = off[d];
        __syncwarp();
        if (d >= 0 && lane == __ffs(peers) - 1) off[d] = base + __popc(peers);
        __syncwarp();
        if (d >= 0) src.emit(i, c, dst, (int64_t)base + __popc(peers & lt));
    }
}

// first grouped index of every cell: lower_bound on the cell column (records are grouped by cell)
__global__ void k_cell_start(const int32_t *__restrict__ gcell, const int64_t *__restrict__ m_ptr, int n_cells,
                             int32_t *__restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    int lo = 0, hi = (int)*m_ptr;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (gcell[mid] < c) lo = mid + 1; else hi = mid; }
    cell_start[c] = lo;
}

// ---------------------------------------------------------------------------------------------
// Stage 2: dedup. Inside a cell the records are sorted by start, so all candidates for a duplicate
// of record i sit directly before it in the same (cell, start) run; the first record of a key in
// BAM order survives (readers.py:129-150). Both key sets are evaluated for every stage-1 survivor
// (readers.py:128-144) so both duplicate counters are exact whichever strategy is selected.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dedup(Grouped g, const int64_t *__restrict__ m_ptr, uint8_t *__restrict__ gflags, int dedup_mode, int min_mapq,
        mgatk_cell_qc *__restrict__ qc, mgatk_stats *__restrict__ stats) {
    __shared__ u32 s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t m = *m_ptr;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = lane_id();
    int cell = -1;
    bool keep = false, paired = false;
    if (i < m) {
        cell = g.cell[i];
        const int pos = g.pos[i];
        const u32 tl = g.tlen[i];
        const u32 mq = g.mq[i];
        const u32 strand = mq & GMQ_STRAND;
        paired = mq & GMQ_PAIRED;
        bool len_dup = false, pos_dup = false;
        if (dedup_mode != MGATK_DEDUP_NONE) {
            for (int64_t j = i - 1; j >= 0; j--) {
                if (g.pos[j] != pos || g.cell[j] != cell) break;
                if ((g.mq[j] & GMQ_STRAND) == strand) {
                    pos_dup = true;
                    if (g.tlen[j] == tl) { len_dup = true; break; }
                }
            }
        }
        keep = dedup_mode == MGATK_DEDUP_FRAGMENT_LENGTH ? !len_dup : dedup_mode == MGATK_DEDUP_POSITION_ONLY ? !pos_dup : true;
        // pileup.py:33-34 mapq gate (after dedup, Q2). An empty SEQ makes the reference raise
        // (readers.py:157); such survivors are reported in stats.n_empty_seq and not piled up.
        const bool process = keep && (int)(mq & 0xff) >= min_mapq && (g.len[i] & 0xffff) != 0;
        gflags[i] = (uint8_t)((process ? GF_PROCESS : 0) | (strand ? GF_STRAND : 0) | (keep ? GF_KEEP : 0));
        if (len_dup) atomicAdd(&s_cnt[1], 1u);
        if (pos_dup) atomicAdd(&s_cnt[2], 1u);
        if (keep) { atomicAdd(&s_cnt[0], 1u); if ((g.len[i] & 0xffff) == 0) atomicAdd(&s_cnt[3], 1u); }
    }
    // per-cell survivors: lanes of a warp mostly share one cell
    const u32 peers = __match_any_sync(kFull, cell);
    const u32 kept = __ballot_sync(kFull, keep), paird = __ballot_sync(kFull, keep && paired);
    if (cell >= 0 && lane == __ffs(peers) - 1) {
        const u32 nk = __popc(kept & peers), np = __popc(paird & peers);
        if (nk) atomicAdd(&qc[cell].n_reads, nk);
        if (np) atomicAdd(&qc[cell].n_paired, np);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd((u64 *)&stats->filtered_reads, (u64)s_cnt[0]);
        if (s_cnt[1]) atomicAdd((u64 *)&stats->dup_with_length, (u64)s_cnt[1]);
        if (s_cnt[2]) atomicAdd((u64 *)&stats->dup_position_only, (u64)s_cnt[2]);
        if (s_cnt[3]) atomicAdd((u64 *)&stats->n_empty_seq, (u64)s_cnt[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Work planning: a unit is (cell, position tile); the tile width shrinks with the cell's read count
// so units carry a bounded number of reads. Dead cells (processors.py:22) get one empty unit that
// only writes zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tiles_for(int cnt, int unit_reads, int ppad, int *width_out) {
    int nt = cnt <= 0 ? 1 : (cnt + unit_reads - 1) / unit_reads;
    const int max_nt = ppad / 64;
    if (nt > max_nt) nt = max_nt;
    int width = ((ppad + nt - 1) / nt + 63) / 64 * 64;
    *width_out = width;
    return (ppad + width - 1) / width;
}

__device__ __forceinline__ bool cell_dead(const mgatk_cell_qc &q, int min_reads) {
    return q.n_reads == 0 || (int64_t)q.n_reads < (int64_t)min_reads;
}

__global__ void __launch_bounds__(1024)
k_plan_scan(const int32_t *__restrict__ cell_start, const mgatk_cell_qc *__restrict__ qc, int n_cells, int min_reads,
            int unit_reads, int ppad, int32_t *__restrict__ unit_start, int32_t *__restrict__ n_units) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_cells; c0 += 1024) {
        const int c = c0 + t;
        int nt = 0, width;
        if (c < n_cells) {
            const int cnt = cell_dead(qc[c], min_reads) ? 0 : cell_start[c + 1] - cell_start[c];
            nt = tiles_for(cnt, unit_reads, ppad, &width);
        }
        int inc = nt;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int v = warp_sums[lane], s = v;
            for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(kFull, s, o); if (lane >= o) s += x; }
            warp_sums[lane] = s - v;
        }
        __syncthreads();
        const int carry = carry_s;
        if (c < n_cells) unit_start[c] = carry + warp_sums[wid] + inc - nt;
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_sums[31] + inc;
        __syncthreads();
    }
    if (t == 0) { unit_start[n_cells] = carry_s; *n_units = carry_s; }
}

__global__ void k_plan_units(const int32_t *__restrict__ cell_start, const mgatk_cell_qc *__restrict__ qc,
                             const int32_t *__restrict__ gpos, const int32_t *__restrict__ unit_start, int n_cells,
                             int min_reads, int unit_reads, int ppad, int halo, Unit *__restrict__ units) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= unit_start[n_cells]) return;
    int lo = 0, hi = n_cells;                              // last cell with unit_start[c] <= u
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (unit_start[mid] <= u) lo = mid; else hi = mid; }
    const int c = lo;
    const bool dead = cell_dead(qc[c], min_reads);
    const int cs = cell_start[c], ce = cell_start[c + 1];
    int width;
    tiles_for(dead ? 0 : ce - cs, unit_reads, ppad, &width);
    Unit un;
    un.cell = c;
    un.t0 = (u - unit_start[c]) * width;
    un.t1 = min(ppad, un.t0 + width);
    if (dead) { un.rbeg = un.rend = 0; }
    else {
        const int first = un.t0 - halo + 1;                // reads starting before cannot reach t0
        int a = cs, b = ce;
        while (a < b) { int mid = (a + b) >> 1; if (gpos[mid] < first) a = mid + 1; else b = mid; }
        un.rbeg = a;
        b = ce;
        while (a < b) { int mid = (a + b) >> 1; if (gpos[mid] < un.t1) a = mid + 1; else b = mid; }
        un.rend = a;
    }
    units[u] = un;
}

// ---------------------------------------------------------------------------------------------
// Stages 3-6. One warp per unit. The warp streams the unit's reads in start order; lanes take
// consecutive reference positions of an aligned block, so every shared-memory update of one
// instruction hits 32 different banks of the ring[plane][slot] layout and needs no atomics (the
// warp is the only writer). Positions left of the current read start are final: they are filtered,
// reduced and written out in 64-position chunks (one 128-byte store per plane), and the ring slots
// are recycled.
// ---------------------------------------------------------------------------------------------
struct PileupArgs {
    const int32_t *gpos; const u32 *goff; const u32 *glen; const uint8_t *gflags;
    const uint8_t *blob;
    const Unit *units; const int32_t *n_units; int32_t *work_counter;
    uint16_t *planes; mgatk_cell_qc *qc; mgatk_stats *stats;
    mgatk_overflow *ovf; int64_t ovf_cap;
    int P, ppad, min_baseq, dist, apply_bias;
    double max_bias;
};

template <int R>
__device__ __forceinline__ void flush_chunk(const PileupArgs &a, u32 *ring, int cell, int base, bool dirty,
                                            u64 &sum, u32 &covered, u32 &maxd) {
    const int lane = lane_id();
    const int p0 = base + 2 * lane;
    u32 *out = reinterpret_cast<u32 *>(a.planes + ((size_t)cell * MGATK_N_PLANES) * a.ppad + p0);
    const size_t pstride = (size_t)a.ppad / 2;             // plane stride in u32
    if (!dirty) {
#pragma unroll
        for (int pl = 0; pl < MGATK_N_PLANES; pl++) out[pl * pstride] = 0u;
        return;
    }
    const int slot = p0 & (R - 1);
    u32 v[10][2];
#pragma unroll
    for (int pl = 0; pl < 10; pl++) {
        uint2 t = *reinterpret_cast<uint2 *>(&ring[pl * R + slot]);
        v[pl][0] = t.x; v[pl][1] = t.y;
        *reinterpret_cast<uint2 *>(&ring[pl * R + slot]) = make_uint2(0u, 0u);
    }
    u32 cov[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        u32 c = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            u32 f = v[2 * b][k], r = v[2 * b + 1][k];
            const u32 t = f + r;
            if (a.apply_bias && t > 0) {                   // pileup.py:143-148, IEEE double, strict >
                const double bias = (double)max(f, r) / (double)t;
                if (bias > a.max_bias) { f = 0; r = 0; v[2 * b][k] = 0; v[2 * b + 1][k] = 0; }
            }
            c += f + r;                                    // pileup.py:150
        }
        cov[k] = c;
        if (c == 0) { v[8][k] = 0; v[9][k] = 0; }          // pileup.py:152-153: position dropped with its Tn5 counts
        else { sum += c; covered++; maxd = max(maxd, c); }
    }
    u32 all[MGATK_N_PLANES][2];
#pragma unroll
    for (int pl = 0; pl < 10; pl++) { all[pl][0] = v[pl][0]; all[pl][1] = v[pl][1]; }
    all[10][0] = cov[0]; all[10][1] = cov[1];
#pragma unroll
    for (int pl = 0; pl < MGATK_N_PLANES; pl++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (all[pl][k] > 65535u) {                     // writers.py:205-218 saturation; exact value kept aside
                const u64 idx = atomicAdd((u64 *)&a.stats->n_overflow, 1ull);
                if ((int64_t)idx < a.ovf_cap) {
                    a.ovf[idx].cell = cell;
                    a.ovf[idx].plane_pos = ((u32)pl << 24) | (u32)(p0 + k);
                    a.ovf[idx].value = all[pl][k];
                } else atomicOr((u64 *)&a.stats->error_bits, (u64)ERR_OVERFLOW_CAP);
                all[pl][k] = 65535u;
            }
        }
        out[pl * pstride] = all[pl][0] | (all[pl][1] << 16);
    }
}

template <int R>
__global__ void __launch_bounds__(kThreads)
k_pileup(PileupArgs a) {
    extern __shared__ u32 smem[];
    const int lane = lane_id();
    u32 *ring = smem + (threadIdx.x >> 5) * (10 * R);
    for (int k = lane; k < 10 * R; k += 32) ring[k] = 0;
    __syncwarp();
    const int n_units = *a.n_units;
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(a.work_counter, 1);
        u = __shfl_sync(kFull, u, 0);
        if (u >= n_units) break;
        const Unit un = a.units[u];
        const int T0 = un.t0, T1 = un.t1, T1c = min(un.t1, a.P);
        int base = T0, dirty_hi = T0;
        u64 sum = 0; u32 covered = 0, maxd = 0;
        bool extent_err = false;

        for (int b0 = un.rbeg; b0 < un.rend; b0 += 32) {
            const int i = b0 + lane;
            int my_pos = 0; u32 my_off = 0, my_len = 0, my_fl = 0;
            if (i < un.rend) {
                my_fl = a.gflags[i];
                if (my_fl & GF_PROCESS) {
                    my_pos = a.gpos[i]; my_off = a.goff[i]; my_len = a.glen[i];
                    const uint8_t *bl = a.blob + 16 * (size_t)my_off;
                    const int L = my_len & 0xffff, nbytes = 4 * (int)(my_len >> 16) + (L + 1) / 2 + L;
                    for (int o = 0; o < nbytes + 127 && o < 1024; o += 128)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(bl + o));
                }
            }
            u32 active = __ballot_sync(kFull, my_fl & GF_PROCESS);
            while (active) {
                const int j = __ffs(active) - 1;
                active &= active - 1;
                const int pos = __shfl_sync(kFull, my_pos, j);
                const u32 off = __shfl_sync(kFull, my_off, j);
                const u32 ln = __shfl_sync(kFull, my_len, j);
                const int strand = (__shfl_sync(kFull, my_fl, j) & GF_STRAND) ? 1 : 0;
                const int L = ln & 0xffff, ncig = ln >> 16;

                // positions left of this read's start are final
                const int lim = min(pos, T1);
                while (base + 64 <= lim) {
                    flush_chunk<R>(a, ring, un.cell, base, base < dirty_hi, sum, covered, maxd);
                    base += 64;
                }
                __syncwarp();
                const long long ring_end = (long long)base + R;   // first position the ring cannot hold

                // Tn5 site (pileup.py:43-50): reverse = start + len(SEQ) - 1, forward = start
                const long long t5 = strand ? (long long)pos + L - 1 : (long long)pos;
                if (t5 >= T0 && t5 < T1c) {
                    if (t5 < ring_end) {
                        if (lane == 0) ring[(8 + strand) * R + ((int)t5 & (R - 1))] += 1;
                        dirty_hi = max(dirty_hi, (int)t5 + 1);
                    } else extent_err = true;
                }

                const uint8_t *bl = a.blob + 16 * (size_t)off;
                const u32 *cig = reinterpret_cast<const u32 *>(bl);
                const uint8_t *seq = bl + 4 * ncig;
                const int8_t *qual = reinterpret_cast<const int8_t *>(seq + (L + 1) / 2);
                const long long q_lo = a.dist > 0 ? a.dist : 0;            // pileup.py:67-72
                const long long q_hi = a.dist > 0 ? (long long)L - a.dist : (long long)L;
                long long ref = pos, qp = 0;                                // pileup.py:52-53
                for (int ci = 0; ci < ncig; ci++) {
                    const u32 w = __ldg(cig + ci);
                    const int op = w & 15;
                    const long long n = w >> 4;
                    if (op == 0 || op == 7 || op == 8) {                    // pileup.py:56
                        long long i_lo = max(max(0ll, (long long)T0 - ref), q_lo - qp);
                        long long i_hi = min(min(n, (long long)T1c - ref), q_hi - qp);
                        if (i_hi > ring_end - ref) { i_hi = ring_end - ref; extent_err = true; }
                        if (i_hi > i_lo) {
                            dirty_hi = max(dirty_hi, (int)(ref + i_hi));
                            const uint8_t *sq = seq; const int8_t *ql = qual;
                            for (long long i0 = i_lo; i0 < i_hi; i0 += 32) {
                                const long long ii = i0 + lane;
                                if (ii < i_hi) {
                                    const long long q = qp + ii;            // pileup.py:75
                                    const int qv = __ldg(ql + q);           // int8 compare, pileup.py:80
                                    const u32 nib = (__ldg(sq + (q >> 1)) >> ((~q & 1) << 2)) & 15u;
                                    if (qv >= a.min_baseq && __popc(nib) == 1) {   // A,C,G,T = 1,2,4,8 (pileup.py:83-86)
                                        const int code = 31 - __clz(nib);
                                        ring[(code * 2 + strand) * R + ((int)(ref + ii) & (R - 1))] += 1;  // :88
                                    }
                                }
                            }
                        }
                        ref += n; qp += n;                                   // pileup.py:90-91
                    } else if (op == 2 || op == 3) ref += n;                 // :92-93
                    else if (op == 4) qp += n;                               // :94-95 (I, H, P: nothing, sic)
                }
                __syncwarp();
            }
        }
        while (base < T1) {
            flush_chunk<R>(a, ring, un.cell, base, base < dirty_hi, sum, covered, maxd);
            base += 64;
        }
        __syncwarp();
        // per-cell depth statistics (processors.py:36-39, writers.py:187-193)
        for (int o = 16; o; o >>= 1) {
            sum += __shfl_xor_sync(kFull, sum, o);
            covered += __shfl_xor_sync(kFull, covered, o);
            maxd = max(maxd, __shfl_xor_sync(kFull, maxd, o));
        }
        if (lane == 0 && covered) {
            atomicAdd((u64 *)&a.qc[un.cell].sum_depth, sum);
            atomicAdd(&a.qc[un.cell].covered, covered);
            atomicMax(&a.qc[un.cell].max_depth, maxd);
        }
        if (extent_err && lane == 0) atomicOr((u64 *)&a.stats->error_bits, (u64)ERR_EXTENT);
    }
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
