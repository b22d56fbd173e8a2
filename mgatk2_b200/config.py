"""Configuration objects crossing the drop-in boundary.

Field names and defaults follow the reference's `PipelineConfig` and its three nested records
(src/core/config.py:8-34,77-114) so `config.quality.min_baseq`, `config.dedup.skip`, ... read the same
on both sides. `to_params()` is the only new piece: it flattens the subset the kernels need into the
`mgatk_params` C struct.
"""
from __future__ import annotations

from dataclasses import dataclass, field

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


@dataclass
class PipelineConfig:
    """Same constructor keywords as the reference (config.py:80-96). As there, `min_distance_from_end`
    is not a constructor argument: it stays 5 unless set on `config.quality` directly."""

    min_baseq: int = 20
    min_mapq: int = 30
    max_strand_bias: float = 0.9
    skip_deduplication: bool = False
    use_fragment_length_dedup: bool = True
    n_cores: int = 8
    worker_batch_size: int | None = None
    io_batch_size: int | None = None
    max_memory_gb: float = 128.0
    sequential: bool = False
    min_reads_per_cell: int = 1
    barcode_tag: str = "CB"
    mito_chr: str = "chrM"
    mito_length: int = 16569
    quality: QualityThresholds = field(init=False)
    dedup: DeduplicationConfig = field(init=False)
    performance: PerformanceConfig = field(init=False)

    def __post_init__(self):
        self.quality = QualityThresholds(self.min_baseq, self.min_mapq, self.max_strand_bias)
        self.dedup = DeduplicationConfig(self.skip_deduplication, self.use_fragment_length_dedup)
        self.performance = PerformanceConfig(self.n_cores, self.worker_batch_size or self.n_cores,
                                             self.io_batch_size or 100, self.max_memory_gb, self.sequential)

    def to_params(self, n_cells: int, max_read_extent: int, flags: int = 0):
        from ._lib import ParamsC
        q = self.quality
        return ParamsC(int(q.min_baseq), int(q.min_mapq), int(q.min_distance_from_end), int(self.dedup.mode),
                       float(q.max_strand_bias), int(self.min_reads_per_cell), int(self.mito_length),
                       int(n_cells), int(max_read_extent), int(flags))
