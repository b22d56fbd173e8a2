/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for the per-cell chrM pileup hot path: a plain-C restatement of the
 * reference's (ollieeknight/mgatk2) Python algorithm, following it function by
 * function. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (mgatk2_b200/csrc) never does.
 *
 * Parity pin: this restatement is checked in tests/test_oracle_golden.py against
 * golden vectors produced by executing the reference's own Python code
 * (tests/golden/make_golden.py, run where /root/reference is mounted) and against
 * the known-answer vectors of SURVEY.md Appendix A.
 *
 * Reference lines restated (paths relative to the reference repo):
 *   read loop, flag/barcode filter, dedup, stats  src/processing/readers.py:87-165,193-199
 *   cell gate, QC                                 src/processing/processors.py:20-55
 *   pileup                                        src/processing/pileup.py:18-126
 *   strand-bias filter                            src/processing/pileup.py:128-154
 *   depth stats / reference-allele vote           src/file_io/writers.py:187-197,220-222,345-349
 *
 * The input is the same structure-of-arrays batch the CUDA path consumes
 * (include/mgatk2_b200.h); the output is the reference's natural dense layout,
 * unsaturated uint32.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mgatk2_b200.h"

typedef struct oracle_outputs {
    uint32_t *counts;      /* [n_cells][P][4][2] after strand filter; may be NULL */
    uint32_t *tn5;         /* [n_cells][P][2]    gated by depth>0;    may be NULL */
    uint32_t *coverage;    /* [n_cells][P]                            may be NULL */
    mgatk_cell_qc *cell_qc;/* [n_cells] (n_reads==0 / sum_depth==0 marks dead)    */
    mgatk_stats *stats;    /* [1]                                                 */
    int64_t *base_totals;  /* [P][4]                                  may be NULL */
    uint8_t *keep;         /* [n_records] 1 = record survived stages 1-2; may be NULL */
} oracle_outputs;

/* ---- dedup key sets: readers.py:75-76 keeps one Python set per barcode; a single
 * open-addressing set keyed by (barcode index, key) is the same relation. ---- */
typedef struct { uint64_t a, b; uint8_t used; } slot_t;
typedef struct { slot_t *s; uint64_t mask; } set_t;

static int set_init(set_t *t, uint64_t n) {
    uint64_t cap = 16;
    while (cap < 2 * n + 1) cap <<= 1;
    t->s = (slot_t *)calloc(cap, sizeof(slot_t));
    t->mask = cap - 1;
    return t->s ? 0 : -1;
}
static uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
/* returns 1 if (a,b) was already present ("key in seen", readers.py:129-134), then adds it (:137-138) */
static int set_test_and_add(set_t *t, uint64_t a, uint64_t b) {
    uint64_t h = mix(a * 0x9e3779b97f4a7c15ULL ^ mix(b)) & t->mask;
    for (;;) {
        slot_t *s = &t->s[h];
        if (!s->used) { s->used = 1; s->a = a; s->b = b; return 0; }
        if (s->a == a && s->b == b) return 1;
        h = (h + 1) & t->mask;
    }
}

typedef struct job {
    const mgatk_params *p;
    const mgatk_batch *b;
    const oracle_outputs *o;
    const int64_t *cell_start; /* [n_cells+1] into cell_reads */
    const int64_t *cell_reads; /* record indices, BAM order within a cell */
    int64_t next_cell;         /* atomic work counter */
    int64_t *base_totals_acc;  /* shared, guarded by mu */
    pthread_mutex_t mu;
} job_t;

static int cmp_u32(const void *x, const void *y) {
    uint32_t a = *(const uint32_t *)x, b = *(const uint32_t *)y;
    return (a > b) - (a < b);
}

/* One cell: processors.py:20-55 process_barcode_worker */
static void process_cell(job_t *J, int64_t c, uint32_t *base_counts, uint32_t *tn5_cuts,
                         uint32_t *depth_buf, int64_t *bt_local) {
    const mgatk_params *p = J->p;
    const mgatk_batch *b = J->b;
    const int64_t P = p->mito_length;
    const int64_t n_reads = J->cell_start[c + 1] - J->cell_start[c];
    mgatk_cell_qc *qc = &J->o->cell_qc[c];
    memset(qc, 0, sizeof(*qc));

    /* n_reads / n_paired are reported for every barcode that kept reads (len(reads_by_barcode[bc]),
       readers.py:164,196), also for cells the gates below drop; dead cells are recognised by
       n_reads < max(1, min_reads_per_cell) or sum_depth == 0. */
    uint32_t n_paired = 0;
    for (int64_t k = 0; k < n_reads; k++)
        if (b->flag[J->cell_reads[J->cell_start[c] + k]] & 1) n_paired++;   /* processors.py:34 */
    qc->n_reads = (uint32_t)n_reads;                       /* processors.py:33 */
    qc->n_paired = n_paired;

    /* processors.py:22  if not reads or len(reads) < config.min_reads_per_cell: return None */
    if (n_reads == 0 || n_reads < p->min_reads_per_cell) return;

    /* pileup.py:24-25 */
    memset(base_counts, 0, sizeof(uint32_t) * P * 8);
    memset(tn5_cuts, 0, sizeof(uint32_t) * P * 2);

    const int min_mapq = p->min_mapq, min_baseq = p->min_baseq, d = p->min_distance_from_end;

    for (int64_t k = 0; k < n_reads; k++) {            /* pileup.py:32 */
        const int64_t i = J->cell_reads[J->cell_start[c] + k];
        if ((int)b->mapq[i] < min_mapq) continue;      /* pileup.py:33-34 */

        const int strand_idx = (b->flag[i] & 0x10) ? 1 : 0;       /* :36-37 */
        const uint8_t *blob = b->blob + 16 * (int64_t)b->blob_off[i];
        const int n_cigar = b->n_cigar[i];
        const int64_t read_length = b->l_seq[i];                   /* :41 len(sequence) */
        const uint32_t *cigar = (const uint32_t *)blob;
        const uint8_t *seq = blob + 4 * (int64_t)n_cigar;
        const int8_t *qual = (const int8_t *)(seq + (read_length + 1) / 2); /* np.int8, readers.py:158 */
        const int64_t reference_start = b->pos[i];

        if (strand_idx) {                              /* pileup.py:43-50 */
            int64_t start_pos = reference_start + read_length - 1;
            if (0 <= start_pos && start_pos < P) tn5_cuts[start_pos * 2 + 1]++;
        } else {
            int64_t start_pos = reference_start;
            if (0 <= start_pos && start_pos < P) tn5_cuts[start_pos * 2 + 0]++;
        }

        int64_t ref_pos = reference_start, query_pos = 0; /* :52-53 */
        for (int ci = 0; ci < n_cigar; ci++) {             /* :55 */
            const int op = cigar[ci] & 0xf;
            const int64_t length = cigar[ci] >> 4;
            if (op == 0 || op == 7 || op == 8) {           /* :56 */
                int64_t start_refpos = ref_pos > 0 ? ref_pos : 0;               /* :57 */
                int64_t end_refpos = ref_pos + length < P ? ref_pos + length : P; /* :58 */
                if (start_refpos >= end_refpos) {          /* :60-63 */
                    query_pos += length; ref_pos += length; continue;
                }
                int64_t offset = start_refpos - ref_pos;   /* :65 */
                int64_t valid_q_start, valid_q_end;        /* :67-72 */
                if (d > 0) { valid_q_start = d; valid_q_end = read_length - d; }
                else { valid_q_start = 0; valid_q_end = read_length; }
                for (int64_t j = 0; j < end_refpos - start_refpos; j++) { /* :74 */
                    int64_t current_qpos = query_pos + offset + j;        /* :75 */
                    if (!(valid_q_start <= current_qpos && current_qpos < valid_q_end)) continue; /* :77 */
                    if ((int)qual[current_qpos] < min_baseq) continue;    /* :80 */
                    int nib = (seq[current_qpos >> 1] >> ((current_qpos & 1) ? 0 : 4)) & 0xf;
                    int base_idx;                          /* :83-86: "=ACMGRSVTWYHKDBN"[nib].upper() in ACGT */
                    if (nib == 1) base_idx = 0; else if (nib == 2) base_idx = 1;
                    else if (nib == 4) base_idx = 2; else if (nib == 8) base_idx = 3;
                    else continue;
                    base_counts[(start_refpos + j) * 8 + base_idx * 2 + strand_idx]++; /* :88 */
                }
                query_pos += length; ref_pos += length;    /* :90-91 */
            } else if (op == 2 || op == 3) {               /* :92-93 */
                ref_pos += length;
            } else if (op == 4) {                          /* :94-95 */
                query_pos += length;
            }                                              /* ops 1,5,6: nothing (sic) */
        }
    }

    /* pileup.py:100-124 (sparse dict) + :128-154 (strand filter), done densely:
       an entry exists iff depth>0 or tn5>0; the filter drops entries whose filtered depth is 0. */
    const double max_bias = p->max_strand_bias;
    const int raw = (p->flags & MGATK_FLAG_RAW_PILEUP) != 0;   /* generate_pileup() only, no filter_strand_bias() */
    uint64_t sum_depth = 0; uint32_t covered = 0, max_depth = 0;
    uint32_t *out_counts = J->o->counts ? J->o->counts + c * P * 8 : NULL;
    uint32_t *out_tn5 = J->o->tn5 ? J->o->tn5 + c * P * 2 : NULL;
    uint32_t *out_cov = J->o->coverage ? J->o->coverage + c * P : NULL;
    for (int64_t pos = 0; pos < P; pos++) {
        uint32_t *bc = &base_counts[pos * 8];
        uint32_t depth = 0;
        for (int base = 0; base < 4; base++) {             /* pileup.py:138-148 */
            uint32_t fwd = bc[base * 2], rev = bc[base * 2 + 1];
            uint32_t total = fwd + rev;
            if (total > 0) {
                double bias = (double)(fwd > rev ? fwd : rev) / (double)total;
                if (!raw && bias > max_bias) { bc[base * 2] = 0; bc[base * 2 + 1] = 0; }
            }
            depth += bc[base * 2] + bc[base * 2 + 1];      /* :150 */
        }
        if (depth > 0) {                                   /* :152-153 */
            sum_depth += depth; depth_buf[covered++] = depth;
            if (depth > max_depth) max_depth = depth;
            for (int base = 0; base < 4; base++) bt_local[pos * 4 + base] += (int64_t)bc[base * 2] + bc[base * 2 + 1];
        } else if (!raw) {
            tn5_cuts[pos * 2] = tn5_cuts[pos * 2 + 1] = 0; /* position dropped with its Tn5 counts */
            memset(bc, 0, 8 * sizeof(uint32_t));
        }
        if (out_cov) out_cov[pos] = depth;
    }

    if (covered == 0 && !raw) {                            /* processors.py:30-31 if not pileup: return None */
        /* undo nothing: bt_local only received covered positions */
        return;
    }
    if (out_counts) memcpy(out_counts, base_counts, sizeof(uint32_t) * P * 8);
    if (out_tn5) memcpy(out_tn5, tn5_cuts, sizeof(uint32_t) * P * 2);

    qsort(depth_buf, covered, sizeof(uint32_t), cmp_u32);  /* writers.py:190 np.median */
    qc->sum_depth = sum_depth;
    qc->covered = covered;
    qc->max_depth = max_depth;
    qc->median_lo = covered ? depth_buf[(covered - 1) / 2] : 0;
    qc->median_hi = covered ? depth_buf[covered / 2] : 0;
}

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    const int64_t P = J->p->mito_length;
    uint32_t *base_counts = (uint32_t *)malloc(sizeof(uint32_t) * P * 8);
    uint32_t *tn5_cuts = (uint32_t *)malloc(sizeof(uint32_t) * P * 2);
    uint32_t *depth_buf = (uint32_t *)malloc(sizeof(uint32_t) * P);
    int64_t *bt_local = (int64_t *)calloc(P * 4, sizeof(int64_t));
    for (;;) {
        int64_t c = __atomic_fetch_add(&J->next_cell, 1, __ATOMIC_RELAXED);
        if (c >= J->p->n_cells) break;
        process_cell(J, c, base_counts, tn5_cuts, depth_buf, bt_local);
    }
    if (J->base_totals_acc) {
        pthread_mutex_lock(&J->mu);
        for (int64_t k = 0; k < P * 4; k++) J->base_totals_acc[k] += bt_local[k];
        pthread_mutex_unlock(&J->mu);
    }
    free(base_counts); free(tn5_cuts); free(depth_buf); free(bt_local);
    return NULL;
}

int mgatk_oracle_run(const mgatk_params *p, const mgatk_batch *b, const oracle_outputs *o, int n_threads) {
    if (!p || !b || !o || !o->cell_qc || !o->stats || p->n_cells < 0 || b->n_records < 0) return MGATK_ERR_BAD_ARG;
    const int64_t N = b->n_records, C = p->n_cells, P = p->mito_length;
    memset(o->stats, 0, sizeof(*o->stats));
    if (o->counts) memset(o->counts, 0, sizeof(uint32_t) * C * P * 8);
    if (o->tn5) memset(o->tn5, 0, sizeof(uint32_t) * C * P * 2);
    if (o->coverage) memset(o->coverage, 0, sizeof(uint32_t) * C * P);
    if (o->base_totals) memset(o->base_totals, 0, sizeof(int64_t) * P * 4);

    /* ---- readers.py:87-165: the single-threaded read loop ---- */
    uint8_t *keep = o->keep ? o->keep : (uint8_t *)malloc(N > 0 ? N : 1);
    set_t with_len = {0, 0}, pos_only = {0, 0};
    const int skip = (p->dedup_mode == MGATK_DEDUP_NONE);
    if (!skip) { if (set_init(&with_len, N) || set_init(&pos_only, N)) return MGATK_ERR_BAD_ARG; }
    mgatk_stats st; memset(&st, 0, sizeof(st));
    int64_t *cell_count = (int64_t *)calloc(C + 1, sizeof(int64_t));
    for (int64_t i = 0; i < N; i++) {
        keep[i] = 0;
        st.total_reads++;                                              /* :93 */
        if (b->flag[i] & (0x4 | 0x100 | 0x800)) continue;              /* :96-97 */
        const int32_t cell = b->bc_idx[i];
        if (cell < 0 || cell >= C) continue;                           /* :104-111 */
        st.stage1_reads++;
        if (!skip) {                                                   /* :118 */
            const uint64_t ref_start = (uint32_t)b->pos[i];            /* :119 */
            const uint64_t is_rev = (b->flag[i] & 0x10) ? 1 : 0;       /* :120 */
            const int64_t t = b->tlen[i];
            const uint64_t abs_tlen = (uint64_t)(t < 0 ? -t : t);      /* :124 */
            const uint64_t ka = ((uint64_t)(uint32_t)cell << 32) | ref_start;
            const int is_fragment_length_dup = set_test_and_add(&with_len, ka, (abs_tlen << 1) | is_rev); /* :129-131,137 */
            const int is_position_only_dup = set_test_and_add(&pos_only, ka, is_rev);                     /* :132-134,138 */
            if (is_fragment_length_dup) st.dup_with_length++;          /* :141-142 */
            if (is_position_only_dup) st.dup_position_only++;          /* :143-144 */
            if (p->dedup_mode == MGATK_DEDUP_FRAGMENT_LENGTH && is_fragment_length_dup) continue; /* :147-148 */
            if (p->dedup_mode == MGATK_DEDUP_POSITION_ONLY && is_position_only_dup) continue;     /* :149-150 */
        }
        keep[i] = 1;                                                   /* :153-165 */
        if (b->l_seq[i] == 0) st.n_empty_seq++;
        cell_count[cell]++;
        st.filtered_reads++;
    }
    if (!skip) { free(with_len.s); free(pos_only.s); }

    /* reads_by_barcode[barcode].append(...) — lists in BAM order */
    int64_t *cell_start = (int64_t *)malloc(sizeof(int64_t) * (C + 1));
    cell_start[0] = 0;
    for (int64_t c = 0; c < C; c++) cell_start[c + 1] = cell_start[c] + cell_count[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (C + 1));
    memcpy(fill, cell_start, sizeof(int64_t) * (C + 1));
    int64_t *cell_reads = (int64_t *)malloc(sizeof(int64_t) * (st.filtered_reads > 0 ? st.filtered_reads : 1));
    for (int64_t i = 0; i < N; i++) if (keep[i]) cell_reads[fill[b->bc_idx[i]]++] = i;
    free(fill); free(cell_count);

    /* ---- processors.py: one task per barcode (the reference's only parallelism) ---- */
    job_t J; memset(&J, 0, sizeof(J));
    J.p = p; J.b = b; J.o = o; J.cell_start = cell_start; J.cell_reads = cell_reads;
    J.base_totals_acc = o->base_totals;
    pthread_mutex_init(&J.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) {
        worker(&J);
    } else {
        pthread_t th[256];
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker, &J);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&J.mu);

    *o->stats = st;
    free(cell_start); free(cell_reads);
    if (!o->keep) free(keep);
    return MGATK_OK;
}

int mgatk_oracle_abi_version(void) { return MGATK_ABI_VERSION; }
