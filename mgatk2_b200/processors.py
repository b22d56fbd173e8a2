"""`CellProcessor` / `process_barcode_worker` — second call of the drop-in seam
(reference src/processing/processors.py:20-144). The counting already happened on the GPU inside
`BAMReader.collect_reads_by_barcode`; here the per-cell gates (processors.py:22,30-31) and the QC row
(processors.py:33-52) are applied to the dense result, and cells are handed to the writer in the
reference's sequential order (first-seen barcode order, processors.py:63-80)."""
from __future__ import annotations

import logging

import numpy as np

from .pileup import PileupGenerator, cell_pileup_dict
from .readers import CellReads, ReadsByBarcode

logger = logging.getLogger(__name__)


def cell_qc_row(barcode: str, qc_row, mito_length: int) -> dict:
    """processors.py:33-52 from the device QC row; float64 on exact integers, hence bit-identical."""
    n_reads = int(qc_row["n_reads"])
    depth_sum, covered = int(qc_row["sum_depth"]), int(qc_row["covered"])
    return {"barcode": barcode, "total_reads": n_reads,
            "total_fragments": n_reads // 2 if int(qc_row["n_paired"]) > 0 else n_reads,
            "mean_depth": np.float64(depth_sum) / np.float64(covered),
            "coverage_breadth": covered / mito_length if mito_length > 0 else 0}


def process_barcode_worker(args):
    """(barcode, reads, config) -> result dict or None, as processors.py:20-55.
    `reads` is a `CellReads` from our BAMReader, or the reference's list of SimpleRead objects."""
    barcode, reads, config = args
    if not reads or len(reads) < config.min_reads_per_cell:
        return None
    try:
        if isinstance(reads, CellReads):
            res, c = reads.result, reads.index
            if res.cell_qc["sum_depth"][c] == 0:
                return None
            pileup = cell_pileup_dict(res, c)
            qc_row = res.cell_qc[c]
        else:                                           # reference-style list of reads: run this one cell
            gen = PileupGenerator(config)
            from .pileup import reads_to_batch
            r = gen._run(reads_to_batch(reads), 0)
            if r.cell_qc["sum_depth"][0] == 0:
                return None
            pileup = cell_pileup_dict(r, 0)
            qc_row = r.cell_qc[0]
        return {"barcode": barcode, "pileup": pileup, "n_reads": len(reads),
                "qc": cell_qc_row(barcode, qc_row, config.mito_length)}
    except Exception as e:                              # processors.py:53-55
        logger.error("Error processing %s: %s", barcode, e)
        return None


class CellProcessor:
    def __init__(self, config, output_dir):
        self.config = config
        self.output_dir = output_dir

    def process_cells_direct(self, reads_by_barcode, incremental_writer=None):
        results, failed = [], 0
        for bc in list(reads_by_barcode.keys()):
            result = process_barcode_worker((bc, reads_by_barcode[bc], self.config))
            if result:
                if incremental_writer:
                    incremental_writer.write_cell(result)
                    results.append({"barcode": result["barcode"], "n_reads": result["n_reads"]})
                else:
                    results.append(result)
            else:
                failed += 1
        if failed:
            logger.warning("%d cells failed", failed)
        return results

    def process_cells_progressive(self, reads_by_barcode, incremental_writer=None):
        """The reference chooses between a process pool and a sequential loop here (processors.py:87-110);
        with the counting on the GPU there is one path, in the sequential (deterministic) cell order."""
        n_cells = len(reads_by_barcode)
        total = sum(len(r) for r in reads_by_barcode.values())
        logger.info("Processing %d cells at an average of %.0f reads/cell", n_cells, total / n_cells if n_cells else 0)
        if isinstance(reads_by_barcode, ReadsByBarcode) and hasattr(incremental_writer, "write_result"):
            return incremental_writer.write_result(reads_by_barcode, self.config)     # dense-plane writers (no dicts)
        return self.process_cells_direct(reads_by_barcode, incremental_writer)
