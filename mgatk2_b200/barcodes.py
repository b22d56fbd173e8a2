"""Whitelist loading and barcode discovery: the callers that define the column order of the planes.

Same names, arguments and error behaviour as the reference's `utils.utils.load_singlecell_csv` (utils.py:14-69),
`file_io.barcode_extraction.extract_barcodes_from_bam` (barcode_extraction.py:12-46) and the dispatch inside
`core.pipeline.run_pipeline` (pipeline.py:214-230). The order of the returned list is the column order
(`writers.py:41`)."""
from __future__ import annotations

import csv
import logging

import numpy as np

from .exceptions import InvalidInputError

logger = logging.getLogger(__name__)


def _coerce(text: str):
    """One metadata cell: empty -> 0, integer text -> int, other numeric text -> float, anything else unchanged."""
    if not text:
        return 0
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            continue
    return text


_TEXT_COLUMNS = frozenset(("barcode", "excluded_reason"))


def load_singlecell_csv(csv_file):
    """cellranger-atac singlecell.csv -> (barcodes of the rows with is__cell_barcode == "1" in file order,
    {column: values of those rows}); numeric columns are coerced cell by cell, `barcode` and `excluded_reason`
    stay text (utils.py:14-69). (None, None) without a file; InvalidInputError for a missing file, a file without
    header or without the `is__cell_barcode` column, or no cell row."""
    if csv_file is None:
        return None, None
    try:
        with open(csv_file, newline="") as f:
            rows = csv.reader(f)
            header = next(rows, None)
            if not header:
                raise InvalidInputError("CSV file has no headers")
            if "is__cell_barcode" not in header:
                raise InvalidInputError("singlecell.csv missing 'is__cell_barcode' column")
            flag_col = header.index("is__cell_barcode")
            width = len(header)
            # DictReader semantics: short rows are padded with None, long rows are cut to the header
            cells = [(r + [None] * width)[:width] for r in rows if r and len(r) > flag_col and r[flag_col] == "1"]
    except FileNotFoundError as e:
        raise InvalidInputError(f"singlecell.csv file not found: {csv_file}") from e
    if not cells:
        raise InvalidInputError(f"No cells found with is__cell_barcode == 1 in {csv_file}")
    columns = list(zip(*cells))
    metadata = {}
    for name, column in zip(header, columns):                # a repeated column name keeps its last occurrence ...
        values = list(column) if name in _TEXT_COLUMNS else [_coerce(v) for v in column]
        k = header.count(name)                               # ... and, as in the reference, lists every row k times
        metadata[name] = values if k == 1 else [v for v in values for _ in range(k)]
    return list(metadata["barcode"]), metadata


def extract_barcodes_from_bam(bam_path: str, barcode_tag: str = "CB", mito_chr: str = "chrM", min_reads: int = 10) -> list:
    """Distinct barcode tag values on `mito_chr` carried by at least `min_reads` records that are neither unmapped
    (0x4) nor flagged duplicate (0x400), sorted (barcode_extraction.py:26-41). Non-string tag values are counted under
    their str() in the reference; the native reader reports them as absent (they never occur for CB)."""
    from .bamio import BamFile
    logger.info("Extracting barcodes from BAM file...")
    with BamFile(bam_path) as bam:
        batch, names, _ = bam.fetch(mito_chr, barcode_tag)
    ok = ((batch.flag & 0x404) == 0) & (batch.bc_idx >= 0)
    counts = np.bincount(batch.bc_idx[ok], minlength=len(names)) if len(names) else np.zeros(0, np.int64)
    barcodes = sorted(n for n, c in zip(names, counts.tolist()) if c >= min_reads)
    logger.info("  Found %d total barcodes", int((counts > 0).sum()))
    logger.info("  Retained %d barcodes with >= %d reads", len(barcodes), min_reads)
    return barcodes


def load_barcodes(barcode_file, bam_path: str, barcode_tag: str = "CB", mito_chr: str = "chrM", min_barcode_reads: int = 10):
    """pipeline.py:214-230: no file -> discover from the BAM; *.csv -> singlecell.csv; else one barcode per line.
    Returns (barcodes, metadata or None)."""
    if barcode_file is None:
        barcodes = extract_barcodes_from_bam(bam_path, barcode_tag=barcode_tag, mito_chr=mito_chr, min_reads=min_barcode_reads)
        if not barcodes:
            raise InvalidInputError(f"No barcodes found in BAM file with tag '{barcode_tag}' and minimum {min_barcode_reads} reads")
        return barcodes, None
    if str(barcode_file).endswith(".csv"):
        return load_singlecell_csv(barcode_file)
    with open(barcode_file) as f:
        return [line.strip() for line in f if line.strip()], None
