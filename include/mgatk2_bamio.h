/*
 * mgatk2_bamio — C ABI of the native BAM ingest (host only; mgatk2_b200/csrc/bamio.cpp -> libmgatk2_bamio.so).
 *
 * Replaces what the reference's read path takes from pysam / htslib (reference src/processing/readers.py and
 * src/file_io/barcode_extraction.py; pysam is a third-party dependency pinned only as ">=0.19", pyproject.toml:15):
 *
 *     pysam.AlignmentFile(path, "rb")              readers.py:37,85      -> mgatk_bam_open / mgatk_bam_close
 *     bam.references                               readers.py:42          -> mgatk_bam_n_refs / _ref_name / _ref_len
 *     bam.fetch(mito_chr) + per-record attributes  readers.py:54-55,87-162; barcode_extraction.py:22-32
 *         reference_start, mapping_quality, flag, template_length, cigartuples, query_sequence,
 *         query_qualities, has_tag / get_tag(barcode_tag)                 -> mgatk_bam_fetch + mgatk_bam_export / _detach
 *
 * fetch(contig) semantics kept: every record placed on the contig in file order, unmapped mates placed there
 * included; reference_start = POS (0-based); the cigar|seq|qual region of each record is copied verbatim into the
 * 16-byte aligned blob of mgatk_batch (include/mgatk2_b200.h); the barcode tag is a Z string compared verbatim.
 * All calls return 0 on success; mgatk_bam_error() gives the reason otherwise. One handle per file, one thread at a time
 * (the library spawns its own inflate / decode threads inside mgatk_bam_fetch).
 */
#ifndef MGATK2_BAMIO_H
#define MGATK2_BAMIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgatk_bam mgatk_bam;

/* 1 = cannot open, 2 = not a BAM / malformed, 3 = outside the limits of the batch layout */
int         mgatk_bam_open(const char *path, mgatk_bam **out);
void        mgatk_bam_close(mgatk_bam *h);
const char *mgatk_bam_error(const mgatk_bam *h);

int         mgatk_bam_n_refs(const mgatk_bam *h);
const char *mgatk_bam_ref_name(const mgatk_bam *h, int i);
int64_t     mgatk_bam_ref_len(const mgatk_bam *h, int i);
int         mgatk_bam_coordinate_sorted(const mgatk_bam *h);       /* @HD SO:coordinate */

/* Decode every record placed on reference ref_id (at most max_records, < 0 = all) into the handle. `tag` = two
 * characters (config.barcode_tag). Uses <path>.bai / <stem>.bai for the start offset when present. */
int         mgatk_bam_fetch(mgatk_bam *h, int ref_id, const char *tag, int n_threads, int64_t max_records);

/* Goes on where a fetch that stopped at max_records left off: the next up to max_records records of the same contig
 * (0 records once the contig is exhausted). bc_id keeps indexing one table of distinct tag values that grows over the
 * parts. For inputs larger than host memory: the parts, cut on reference_start borders by the caller
 * (mgatk2_b200.bamio.iter_bam_chrM), feed the accumulating device path (MGATK_FLAG_ACCUMULATE, include/mgatk2_b200.h). */
int         mgatk_bam_fetch_more(mgatk_bam *h, int n_threads, int64_t max_records);

/* on != 0: a fetch / fetch_more that reaches max_records does not stop there but after the last record with the same
 * reference_start, so that every part ends on a start border (all candidates for a duplicate of a read share its start,
 * readers.py:118-150, and the accumulating device path needs no state across parts). Off by default. */
int         mgatk_bam_align_parts(mgatk_bam *h, int on);

int64_t     mgatk_bam_n_records(const mgatk_bam *h);
int64_t     mgatk_bam_blob_bytes(const mgatk_bam *h);
int64_t     mgatk_bam_n_barcodes(const mgatk_bam *h);              /* distinct tag values, first-appearance order */
int64_t     mgatk_bam_barcode_bytes(const mgatk_bam *h);

/* Copy the decoded records into caller arrays (sizes from the four calls above). bc_id = index into the table of
 * distinct tag values, -1 = no tag, -2 = tag present but not a string. qual_missing[i] = QUAL absent (0xFF).
 * barcode_chars = the distinct values concatenated, barcode_end[i] = end offset of value i. */
int         mgatk_bam_export(const mgatk_bam *h, int32_t *pos, int32_t *tlen, uint16_t *flag, uint8_t *mapq, int32_t *bc_id,
                             uint16_t *l_seq, uint16_t *n_cigar, uint32_t *blob_off, uint8_t *blob, uint8_t *qual_missing,
                             char *barcode_chars, int64_t *barcode_end);

/* The same hand-over without the copy: the ten record arrays of the last fetch become the caller's (one malloc block
 * each, exactly n_records / blob_bytes elements; NULL when empty), to be released with mgatk_bam_free(). The handle is
 * left without records. Writing every record twice costs first-touch page faults that bound the ingest on hosts where
 * those are slow (virtual machines); arrays[] order: pos, tlen, flag, mapq, bc_id, l_seq, n_cigar, blob_off, blob,
 * qual_missing. barcode_chars / barcode_end as in mgatk_bam_export. */
int         mgatk_bam_detach(mgatk_bam *h, void *arrays[10], char *barcode_chars, int64_t *barcode_end);
void        mgatk_bam_free(void *p);

/* The entry of reference `ref_id` in a .bai file (the same parser the fetch uses for its start offset; the reference
 * relies on pysam.fetch(contig) -> htslib's index walk, readers.py:85-88): out = { n_ref, real bins, chunks in them,
 * linear-index intervals, smallest chunk virtual offset (-1: no chunk), pseudo-bin 37450 present, then its four
 * values ref_beg, ref_end, n_mapped, n_unmapped }. 1 = cannot read, 2 = not a well-formed BAI covering ref_id. */
int         mgatk_bai_inspect(const char *bai_path, int ref_id, int64_t out[10]);

#ifdef __cplusplus
}
#endif
#endif /* MGATK2_BAMIO_H */
