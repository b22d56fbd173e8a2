"""`MtDNAPipeline` — the caller of the three-call seam (reference src/core/pipeline.py:23-181), so a BAM file goes
in and the reference's text outputs come out with the counting on the GPU:

    BAMReader.collect_reads_by_barcode()  ->  CellProcessor.process_cells_progressive()  ->  writer.finalize()

Same constructor keywords, same return value of `run()`, same files (`output/output.{A,C,G,T,coverage}.txt.gz`,
`output/output.depthTable.txt`, `output/<mito>_refAllele.txt`, `qc/cell_stats.csv`, `qc/summary.txt`), or with
`output_format="hdf5"` the reference's default `output/counts.h5` + `output/metadata.h5` (written by `h5lite`; the HTML
report that follows them in the reference is out of scope).
No BAM index is required (the native reader scans without one; the reference would call `pysam.index`)."""
from __future__ import annotations

import datetime
import logging
import time
from pathlib import Path

from .barcodes import load_barcodes
from .config import PipelineConfig
from .exceptions import InvalidInputError

logger = logging.getLogger(__name__)
VERSION = "mgatk2_b200 round 1"


def write_run_summary(run_metadata: dict, output_path: Path):
    """`qc/summary.txt` in the layout of file_io/formats.py:49-57: title, rule, `key: value` lines, then the indented
    `parameters` block when there is one."""
    params = run_metadata.get("parameters")
    lines = ["mgatk2 Run Summary", "=" * 20]
    lines += [f"{k}: {v}" for k, v in run_metadata.items() if k != "parameters"]
    if "parameters" in run_metadata:
        lines += ["", "Parameters:"] + [f"  {k}: {v}" for k, v in params.items()]
    Path(output_path).write_text("\n".join(lines) + "\n")


class MtDNAPipeline:
    def __init__(self, bam_path: str, barcodes: list, output_dir, config: PipelineConfig | None = None,
                 output_format: str = "standard", barcode_metadata=None, sample_name: str = "mgatk2",
                 report_title: str | None = None, report_subtitle: str | None = None, working_directory: str | None = None,
                 device: int = 0, devices: list | None = None, max_batch_records: int | None = None):
        self.bam_path = Path(bam_path)
        self.barcodes = set(barcodes)
        self.barcode_list = list(barcodes)
        self.output_dir = Path(output_dir)
        self.config = config or PipelineConfig()
        self.output_format = output_format.lower()
        self.barcode_metadata = barcode_metadata
        self.sample_name = sample_name
        self.devices = list(devices) if devices else [device]      # GPUs of this box; the cells are split between them
        self.device = self.devices[0]
        self.max_batch_records = max_batch_records                  # larger contigs are streamed in parts (dispatch.py)
        if not self.bam_path.exists():
            raise InvalidInputError(f"BAM file not found: {bam_path}")
        from .bamio import BamFile, pick_mito_contig
        with BamFile(str(self.bam_path)) as bam:                          # pipeline.py:62-76
            if self.config.mito_chr not in bam.references:
                alt = pick_mito_contig(bam.references)
                if alt is None:
                    raise InvalidInputError(f"Mitochondrial chromosome '{self.config.mito_chr}' not found. "
                                            f"Available: {', '.join(bam.references[:10])}")
                logger.warning("Using '%s' instead of '%s'", alt, self.config.mito_chr)
                self.config.mito_chr = alt
        self.output_dir.mkdir(parents=True, exist_ok=True)

    def run(self) -> dict:
        from .processors import CellProcessor
        from .readers import BAMReader
        from .writers import DenseHDF5Writer, DenseTextWriter
        start = time.time()
        logger.info("Collecting reads from BAM by barcode...")
        reader = BAMReader(str(self.bam_path), self.config, self.barcodes, barcode_list=self.barcode_list,
                           devices=self.devices, max_batch_records=self.max_batch_records)
        reads_by_barcode, stats = reader.collect_reads_by_barcode()
        if not reads_by_barcode:
            logger.error("No reads found for any barcodes!")
            return {}
        n_cells_input = len(reads_by_barcode)
        t_write = time.time()
        if self.output_format == "hdf5":                                  # pipeline.py:88-94, the default of `mgatk2 run`
            writer = DenseHDF5Writer(self.output_dir, self.config, self.barcode_list, barcode_metadata=self.barcode_metadata)
        else:
            writer = DenseTextWriter(self.output_dir, self.config, self.barcode_list)
        cell_results = CellProcessor(self.config, self.output_dir).process_cells_progressive(reads_by_barcode, writer)
        if not cell_results:
            logger.error("No cells passed quality filters")
            return {}
        qc_dir = self.output_dir / "qc"
        writer.finalize(qc_dir)
        self.timings = dict(getattr(reader, "timings", {}), write_s=time.time() - t_write, total_s=time.time() - start)
        c = self.config
        write_run_summary({                                               # analysis/qc.py:68-93
            "mgatk_version": VERSION, "run_date": datetime.datetime.now().isoformat(), "input_bam": str(self.bam_path),
            "output_dir": str(self.output_dir), "reference": c.mito_chr, "reference_length": c.mito_length,
            "cells_total": n_cells_input, "cells_passed_qc": len(cell_results),
            "cells_failed_qc": n_cells_input - len(cell_results),
            "parameters": {"min_base_quality": c.quality.min_baseq, "min_mapping_quality": c.quality.min_mapq,
                           "min_reads_per_cell": c.min_reads_per_cell, "max_strand_bias": c.quality.max_strand_bias,
                           "skip_deduplication": c.dedup.skip, "use_fragment_length_dedup": c.dedup.use_fragment_length,
                           "barcode_tag": c.barcode_tag, "mito_chr": c.mito_chr, "n_cores": c.performance.n_cores},
        }, qc_dir / "summary.txt")
        logger.info("Pipeline complete")
        logger.info("Elapsed time: %ds", int(time.time() - start))
        self.stats = stats
        return {"cells_processed": n_cells_input, "cells_passed_qc": len(cell_results),
                "mean_reads": sum(r["n_reads"] for r in cell_results) / len(cell_results) if cell_results else 0}


def run_pipeline(bam_path: str, output_dir: str, barcode_file: str | None = None, min_barcode_reads: int = 10,
                 barcode_tag: str = "CB", mito_chr: str = "chrM", output_format: str = "standard", device: int = 0,
                 devices: list | None = None, max_batch_records: int | None = None, **config_kwargs) -> dict:
    """core.pipeline.run_pipeline (pipeline.py:184-270): whitelist from file / singlecell.csv / the BAM itself, then
    `MtDNAPipeline.run()`. `config_kwargs` are `PipelineConfig` keywords (min_baseq, min_mapq, max_strand_bias,
    skip_deduplication, use_fragment_length_dedup, min_reads_per_cell, ...)."""
    barcodes, metadata = load_barcodes(barcode_file, bam_path, barcode_tag=barcode_tag, mito_chr=mito_chr,
                                       min_barcode_reads=min_barcode_reads)
    config = PipelineConfig(barcode_tag=barcode_tag, mito_chr=mito_chr, **config_kwargs)
    return MtDNAPipeline(bam_path, barcodes, Path(output_dir), config, output_format=output_format,
                         barcode_metadata=metadata, device=device, devices=devices,
                         max_batch_records=max_batch_records).run()
