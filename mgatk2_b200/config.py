"""Configuration objects crossing the drop-in boundary.

Field names and defaults follow the reference's `PipelineConfig` and its three nested records
(src/core/config.py:8-34,77-114) so `config.quality.min_baseq`, `config.dedup.skip`, ... read the same
on both sides. `to_params()` is the only new piece: it flattens the subset the kernels need into the
`mgatk_params` C struct.
"""
from __future__ import annotations

from dataclasses import dataclass

DEDUP_FRAGMENT_LENGTH, DEDUP_POSITION_ONLY, DEDUP_NONE = 0, 1, 2
DEDUP_MODE_NAMES = {"alignment_and_fragment_length": DEDUP_FRAGMENT_LENGTH,
                    "alignment_start": DEDUP_POSITION_ONLY, "none": DEDUP_NONE}


@dataclass
class QualityThresholds:
    min_baseq: int = 20
    min_mapq: int = 30
    max_strand_bias: float = 1.0
    min_distance_from_end: int = 5


@dataclass
class DeduplicationConfig:
    skip: bool = False
    use_fragment_length: bool = True

    @property
    def mode(self) -> int:
        return DEDUP_NONE if self.skip else (DEDUP_FRAGMENT_LENGTH if self.use_fragment_length else DEDUP_POSITION_ONLY)


@dataclass
class PerformanceConfig:
    n_cores: int = 8
    worker_batch_size: int = 8
    io_batch_size: int = 100
    max_memory_gb: float = 128.0
    sequential: bool = False


class PipelineConfig:
    """Same constructor keywords as the reference (config.py:80-96), unknown keywords swallowed by `**kwargs` as there.
    `min_distance_from_end` is not a constructor argument: it stays 5 unless set on `config.quality` directly (SURVEY
    Q1). Like the reference's class, only the three nested records and the four scalar fields below are attributes."""

    def __init__(self, min_baseq: int = 20, min_mapq: int = 30, max_strand_bias: float = 0.9,
                 skip_deduplication: bool = False, use_fragment_length_dedup: bool = True, n_cores: int = 8,
                 worker_batch_size: int | None = None, io_batch_size: int | None = None, max_memory_gb: float = 128.0,
                 sequential: bool = False, min_reads_per_cell: int = 1, barcode_tag: str = "CB", mito_chr: str = "chrM",
                 mito_length: int = 16569, **kwargs):
        self.quality = QualityThresholds(min_baseq, min_mapq, max_strand_bias)
        self.dedup = DeduplicationConfig(skip_deduplication, use_fragment_length_dedup)
        self.performance = PerformanceConfig(n_cores, worker_batch_size or n_cores, io_batch_size or 100,
                                             max_memory_gb, sequential)
        self.min_reads_per_cell = min_reads_per_cell
        self.barcode_tag = barcode_tag
        self.mito_chr = mito_chr
        self.mito_length = mito_length

    def __repr__(self):
        return (f"PipelineConfig(quality={self.quality}, dedup={self.dedup}, performance={self.performance}, "
                f"min_reads_per_cell={self.min_reads_per_cell}, barcode_tag={self.barcode_tag!r}, "
                f"mito_chr={self.mito_chr!r}, mito_length={self.mito_length})")

    def to_params(self, n_cells: int, max_read_extent: int, flags: int = 0):
        from ._lib import ParamsC
        q = self.quality
        return ParamsC(int(q.min_baseq), int(q.min_mapq), int(q.min_distance_from_end), int(self.dedup.mode),
                       float(q.max_strand_bias), int(self.min_reads_per_cell), int(self.mito_length),
                       int(n_cells), int(max_read_extent), int(flags))
