/* mgatk2_b200 — C ABI of the dense-plane text writer.
 *
 * Writes the gzip text files of the reference's IncrementalTextWriter straight from the dense planes the pileup
 * path returns (include/mgatk2_b200.h, mgatk_outputs.planes), instead of one Python dict per covered position:
 *
 *     output.{A,C,G,T}.txt.gz   rows "pos,barcode,fwd,rev"  where fwd > 0 or rev > 0      src/file_io/writers.py:449-466,471-486
 *     output.coverage.txt.gz    rows "pos,barcode,depth"    where depth > 0               src/file_io/writers.py:440-447,471-486
 *
 * pos is 1-based, rows are grouped by cell in the order of the cell list, positions ascending inside a cell, no
 * header, "\n" line ends (src/file_io/formats.py). The file is a sequence of gzip members (one per group of cells,
 * compressed in parallel), which every gzip reader takes as one stream; the inflated bytes are identical to what the
 * reference writes, the compressed bytes are not (they are not for the reference either: the gzip header carries a
 * time stamp).
 *
 * All calls return 0 on success; mgatk_text_error() gives the reason otherwise. Host code only (g++, zlib, threads).
 */
#ifndef MGATK2_TEXTIO_H
#define MGATK2_TEXTIO_H

#include <stdint.h>

#include "mgatk2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* planes      [n_cells][MGATK_N_PLANES][pos_pad] uint16, saturated at 65535 (host memory)
 * overflow    exact values of the saturated entries (mgatk_outputs.overflow), n_overflow of them; may be NULL
 * plane_a     plane written as the third column (MGATK_PLANE_*: A_fwd ... T_fwd, or MGATK_PLANE_COVERAGE)
 * plane_b     plane written as the fourth column (the _rev plane), or -1 for the three-column coverage file
 * cells       indices of the cells to write, in output order (the live cells in first-seen order), n_listed of them
 * names       their barcodes concatenated, name_end[i] = end offset of the i-th listed barcode
 * level       gzip level (the reference uses 9); n_threads <= 0 means all cores
 * rows_out    if not NULL: number of rows written */
int mgatk_text_write_plane_file(const char *path, const uint16_t *planes, int32_t n_cells, int32_t pos_pad,
                                int32_t mito_length, const mgatk_overflow *overflow, int64_t n_overflow,
                                int32_t plane_a, int32_t plane_b, const int32_t *cells, int64_t n_listed,
                                const char *names, const int64_t *name_end, int32_t level, int32_t n_threads,
                                int64_t *rows_out);

/* reason of the last failure on the calling thread */
const char *mgatk_text_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MGATK2_TEXTIO_H */
