/*
 * mgatk2_b200 — C ABI of the B200-native per-cell chrM pileup hot path.
 *
 * This is the drop-in boundary for the three-call seam inside the reference's
 * MtDNAPipeline.run() (reference src/core/pipeline.py:80-81,102-103):
 *
 *     BAMReader.collect_reads_by_barcode()           src/processing/readers.py:63-201
 *     CellProcessor.process_cells_progressive()      src/processing/processors.py:87-144
 *       -> process_barcode_worker()                  src/processing/processors.py:20-55
 *       -> PileupGenerator.generate_pileup()         src/processing/pileup.py:18-126
 *       -> PileupGenerator.filter_strand_bias()      src/processing/pileup.py:128-154
 *     per-cell depth statistics / reference allele   src/file_io/writers.py:187-197,345-349,493-500
 *
 * The reference has no FFI for this path (it is pure Python); the entry points
 * below are what a ctypes/cffi binding inside those three functions would call
 * (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers are caller-owned;
 *   - "_device" entry points take DEVICE pointers and a cudaStream_t passed as
 *     void*; they only enqueue work (no host synchronisation);
 *   - "_host" entry points take HOST pointers, do H2D, the kernels, D2H and
 *     synchronise before returning;
 *   - every entry point returns an mgatk_status; 0 is success. No exceptions
 *     cross the boundary. mgatk_last_error() gives a human readable string.
 *   - one handle per GPU; handles are independent (no global mutable state);
 *     a handle must be used from one thread at a time.
 */
#ifndef MGATK2_B200_H
#define MGATK2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGATK_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------- */
typedef enum mgatk_status {
    MGATK_OK = 0,
    MGATK_ERR_BAD_ARG = 1,       /* null pointer, negative size, bad enum           */
    MGATK_ERR_CUDA = 2,          /* a CUDA runtime call failed (see last_error)     */
    MGATK_ERR_WORKSPACE = 3,     /* caller workspace smaller than workspace_bytes() */
    MGATK_ERR_UNSORTED = 4,      /* records are not sorted by reference_start       */
    MGATK_ERR_EXTENT = 5,        /* a read exceeds params.max_read_extent           */
    MGATK_ERR_OVERFLOW_CAP = 6,  /* more >65535 entries than overflow_capacity      */
    MGATK_ERR_NO_DEVICE = 7,     /* no usable CUDA device                           */
    MGATK_ERR_RANGE = 8,         /* n_cells / n_records / blob outside limits       */
    MGATK_ERR_STREAM_SATURATED = 9 /* streaming: more cells with entries beyond 65535 than carry-plane sets */
} mgatk_status;

/* ---- dedup strategies (reference src/cli/utils.py:164-169) ---------------- */
#define MGATK_DEDUP_FRAGMENT_LENGTH 0 /* alignment_and_fragment_length: (start, strand, |tlen|) */
#define MGATK_DEDUP_POSITION_ONLY   1 /* alignment_start: (start, strand)                       */
#define MGATK_DEDUP_NONE            2 /* none: skip                                             */

/* ---- output plane order -------------------------------------------------- */
/* planes[cell][plane][pos_padded], uint16, saturated at 65535 exactly like the
 * reference's HDF5 writer (src/file_io/writers.py:205-218); values above 65535
 * are additionally listed exactly in the overflow list so that the unsaturated
 * text layout (src/file_io/writers.py:442-462) stays bit-exact. */
#define MGATK_PLANE_A_FWD 0
#define MGATK_PLANE_A_REV 1
#define MGATK_PLANE_C_FWD 2
#define MGATK_PLANE_C_REV 3
#define MGATK_PLANE_G_FWD 4
#define MGATK_PLANE_G_REV 5
#define MGATK_PLANE_T_FWD 6
#define MGATK_PLANE_T_REV 7
#define MGATK_PLANE_TN5_FWD 8
#define MGATK_PLANE_TN5_REV 9
#define MGATK_PLANE_COVERAGE 10
#define MGATK_N_PLANES 11

/* positions are padded to a multiple of 64 so every flush is a full 128-byte store */
#define MGATK_POS_PAD(P) ((((int64_t)(P)) + 63) / 64 * 64)

/* ---- parameters (reference src/core/config.py:8-23,111-114) --------------- */
typedef struct mgatk_params {
    int32_t min_baseq;           /* QualityThresholds.min_baseq   (int8 compare, pileup.py:80) */
    int32_t min_mapq;            /* QualityThresholds.min_mapq    (pileup.py:33)               */
    int32_t min_distance_from_end; /* QualityThresholds.min_distance_from_end (pileup.py:67-72) */
    int32_t dedup_mode;          /* MGATK_DEDUP_*                 (readers.py:118-150)         */
    double  max_strand_bias;     /* QualityThresholds.max_strand_bias (pileup.py:143-145)      */
    int32_t min_reads_per_cell;  /* PipelineConfig.min_reads_per_cell (processors.py:22)       */
    int32_t mito_length;         /* PipelineConfig.mito_length, 16569                          */
    int32_t n_cells;             /* whitelist length; bc_idx in [0,n_cells) or -1              */
    int32_t max_read_extent;     /* upper bound on max(reference span, l_seq) over the batch;
                                    verified on the device (MGATK_ERR_EXTENT)                  */
    int32_t flags;               /* MGATK_FLAG_*                                               */
} mgatk_params;

/* generate_pileup() without filter_strand_bias(): no strand-bias zeroing and Tn5 counts are
 * kept at positions without coverage, i.e. the dict PileupGenerator.generate_pileup returns
 * (pileup.py:100-124) before pileup.py:128-154 is applied. */
#define MGATK_FLAG_RAW_PILEUP 1

/* Streaming over several batches (BASELINE configs[4]: inputs larger than HBM). The records of a coordinate-sorted
 * BAM are cut into batches on reference_start borders (all duplicates of a read share its start, so dedup needs no
 * state across batches, readers.py:118-150); the planes stay resident and every batch ADDS its raw counts:
 *     mgatk_stream_begin_device(...)                            zero planes / QC / counters
 *     mgatk_pileup_device(..., flags | MGATK_FLAG_ACCUMULATE)   once per batch, in file order
 *     mgatk_stream_finish_device(...)                           cell gate, strand-bias filter, coverage, Tn5 gating,
 *                                                               depth statistics, base totals, medians
 * The result equals the one-batch result bit for bit, whatever the depth: a cell whose plane entry passes 65535 while the
 * batches add up (bulk mode, deep piles) gets a set of 32-bit carry planes inside the handle (64 sets, allocated by
 * mgatk_stream_begin_device; one stream per handle at a time), folded back in by mgatk_stream_finish_device.
 * MGATK_ERR_STREAM_SATURATED only when every set is taken. */
#define MGATK_FLAG_ACCUMULATE 2

/* ---- one batch of records, structure-of-arrays, BAM (coordinate) order ---- */
/* One entry per record returned by fetch(chrM) (readers.py:87-93), i.e. also
 * unmapped-placed / secondary / supplementary records and records without a
 * (whitelisted) barcode: the flag and whitelist tests are stage 1 on the GPU.
 * blob: for record i, starting at byte 16*blob_off[i]:
 *        n_cigar[i] uint32 BAM cigar words (len<<4|op), then (l_seq+1)/2 bytes of
 *        4-bit packed SEQ (BAM nibble order), then l_seq bytes of raw phred QUAL —
 *        the same contiguous cigar|seq|qual region a BAM record carries. */
typedef struct mgatk_batch {
    int64_t n_records;
    const int32_t  *pos;      /* reference_start, 0-based                              */
    const int32_t  *tlen;     /* template_length (signed)                              */
    const uint16_t *flag;     /* BAM FLAG                                              */
    const uint8_t  *mapq;     /* mapping_quality                                       */
    const int32_t  *bc_idx;   /* whitelist index of the barcode tag, -1 = absent/unknown */
    const uint16_t *l_seq;    /* len(query_sequence)                                   */
    const uint16_t *n_cigar;  /* number of cigar operations                            */
    const uint32_t *blob_off; /* offset of the record's blob in 16-byte units          */
    const uint8_t  *blob;     /* cigar|seq|qual blobs, each 16-byte aligned            */
    int64_t blob_bytes;       /* total size of blob (multiple of 16)                   */
} mgatk_batch;

/* ---- per-cell QC row, 32 bytes (processors.py:33-39, writers.py:187-197) --- */
typedef struct mgatk_cell_qc {
    uint32_t n_reads;     /* len(reads) after dedup                (processors.py:33)   */
    uint32_t n_paired;    /* reads with flag&1                     (processors.py:34)   */
    uint64_t sum_depth;   /* sum of depth over covered positions   (writers.py:192)     */
    uint32_t covered;     /* positions with depth>0                (processors.py:38)   */
    uint32_t max_depth;   /* unsaturated                           (writers.py:191)     */
    uint32_t median_lo;   /* the two middle order statistics of the covered depths;    */
    uint32_t median_hi;   /* np.median == (lo+hi)/2                (writers.py:190)     */
} mgatk_cell_qc;

/* ---- global counters (readers.py:193-199) --------------------------------- */
typedef struct mgatk_stats {
    uint64_t total_reads;      /* every fetched record                                  */
    uint64_t stage1_reads;     /* passed flag + barcode filter                          */
    uint64_t filtered_reads;   /* survivors of dedup (== stats["filtered_reads"])       */
    uint64_t dup_with_length;  /* stats["duplicate_reads_with_length"]                  */
    uint64_t dup_position_only;/* stats["duplicate_reads_position_only"]                */
    uint64_t n_empty_seq;      /* dedup survivors with l_seq==0 (reference raises)      */
    uint64_t n_overflow;       /* entries written to the overflow list                  */
    uint64_t error_bits;       /* bit0 unsorted, bit1 extent, bit2 overflow capacity, bit3 stream saturated */
} mgatk_stats;

/* exact value of one saturated plane entry */
typedef struct mgatk_overflow {
    int32_t  cell;
    uint32_t plane_pos;  /* plane<<24 | position */
    uint32_t value;
} mgatk_overflow;

/* ---- outputs -------------------------------------------------------------- */
typedef struct mgatk_outputs {
    uint16_t        *planes;       /* [n_cells][MGATK_N_PLANES][MGATK_POS_PAD(P)]; fully written */
    mgatk_cell_qc   *cell_qc;      /* [n_cells]                                                 */
    mgatk_stats     *stats;        /* [1]                                                       */
    int64_t         *base_totals;  /* [P][4] sum over cells of fwd+rev after filtering
                                      (reference allele vote, writers.py:220-222,345-349)      */
    mgatk_overflow  *overflow;     /* [overflow_capacity], may be NULL if capacity 0           */
    int64_t          overflow_capacity;
} mgatk_outputs;

typedef struct mgatk_handle mgatk_handle;

/* library / ABI */
int         mgatk_abi_version(void);
const char *mgatk_status_string(int status);

/* handle: binds to CUDA device `device`; owns nothing but small scratch and the
 * buffers used by the _host entry point */
int         mgatk_create(mgatk_handle **out, int device);
int         mgatk_destroy(mgatk_handle *h);
const char *mgatk_last_error(const mgatk_handle *h);

/* size of the device workspace the _device entry point needs for a batch shape; max_read_extent as in mgatk_params
 * (it selects the slot layout: 32 bytes per read up to an extent of 56, 16 + 16 * ceil(extent / 32) bytes beyond, at
 * most 144). -1 when the shape is outside the limits. */
int64_t     mgatk_workspace_bytes(int64_t n_records, int32_t n_cells, int32_t max_read_extent);

/* Stages 1-6 on device-resident buffers; enqueues on `stream`, never blocks.
 * After the stream has drained, mgatk_check_stats turns a host copy of
 * stats (its error_bits) into a status code. */
int mgatk_pileup_device(mgatk_handle *h, const mgatk_params *params,
                        const mgatk_batch *batch_dev, const mgatk_outputs *out_dev,
                        void *workspace_dev, int64_t workspace_bytes, void *stream);

int mgatk_check_stats(const mgatk_stats *stats_host);

/* streaming: see MGATK_FLAG_ACCUMULATE. params->flags of the finish call selects raw / filtered output. */
int mgatk_stream_begin_device(mgatk_handle *h, const mgatk_params *params, const mgatk_outputs *out_dev, void *stream);
int mgatk_stream_finish_device(mgatk_handle *h, const mgatk_params *params, const mgatk_outputs *out_dev, void *stream);

/* PileupGenerator.filter_strand_bias (pileup.py:128-154) on raw planes, in place:
 * planes_dev is [n_cells][MGATK_N_PLANES][MGATK_POS_PAD(P)] as written with MGATK_FLAG_RAW_PILEUP
 * (values above 65535 are not representable here; use the 32-bit variant below for those). */
int mgatk_filter_strand_bias_device(mgatk_handle *h, uint16_t *planes_dev, int32_t n_cells, int32_t mito_length,
                                    double max_strand_bias, void *stream);
/* The same rule on 32-bit planes (same shape): exact for any depth. This is what PileupGenerator.filter_strand_bias
 * is given when a caller hands over the dict generate_pileup() returned (pileup.py:128: plain Python ints there). */
int mgatk_filter_strand_bias_u32_device(mgatk_handle *h, uint32_t *planes_dev, int32_t n_cells, int32_t mito_length,
                                        double max_strand_bias, void *stream);

/* Same, but batch and outputs are HOST buffers (pinned for full speed): H2D copy,
 * kernels, D2H copy, synchronise. Device buffers are cached inside the handle. */
int mgatk_pileup_host(mgatk_handle *h, const mgatk_params *params,
                      const mgatk_batch *batch_host, const mgatk_outputs *out_host);

/* The same call in two halves, for callers with more than one batch (one BAM per sample or per
 * cell, as in the reference's `call` command, cli/options.py:310,421): submit enqueues the upload, the kernels and the
 * download on three streams and returns; wait blocks until out_host of that ticket is complete
 * (overflow list included) and returns the batch's status. Up to TWO tickets may be in flight:
 * the upload of batch k+1 then runs while batch k computes and downloads (PCIe is full duplex),
 * which is what bounds throughput with host buffers. batch_host / out_host must stay valid and
 * untouched until wait returns; pinned memory is needed for the copies to overlap. Tickets may be
 * waited for in any order. mgatk_pileup_host == submit + wait. */
int mgatk_pileup_host_submit(mgatk_handle *h, const mgatk_params *params, const mgatk_batch *batch_host,
                             const mgatk_outputs *out_host, int64_t *ticket);
int mgatk_pileup_host_wait(mgatk_handle *h, int64_t ticket);

/* number of kernel launches issued by the last mgatk_pileup_* call on this handle */
int64_t mgatk_last_launch_count(const mgatk_handle *h);

/* device time (ms, CUDA events on the launch stream) of each stage of the last
 * _host call; names/ms arrays of length *n (n<=16). For profiling only. */
int mgatk_last_stage_times(const mgatk_handle *h, const char **names, float *ms, int *n);

#ifdef __cplusplus
}
#endif
#endif /* MGATK2_B200_H */
