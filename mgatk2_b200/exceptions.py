"""Error types of the pileup path, named after the reference's (src/core/exceptions.py:4-86) so callers
that catch `ProcessingError` / `BAMReadError` keep working. `PileupKernelError` is new: a non-zero
C-ABI status mapped onto `ProcessingError` (SURVEY §8b error convention)."""
from __future__ import annotations


class MgatkError(Exception):
    pass


class InvalidInputError(MgatkError):
    pass


class ProcessingError(MgatkError):
    pass


class BAMReadError(ProcessingError):
    def __init__(self, bam_path: str, message: str):
        super().__init__(f"BAM read error for {bam_path}: {message}")
        self.bam_path = bam_path


class BAMFormatError(InvalidInputError):
    """The file is not a readable BAM (src/core/exceptions.py:66-73; raised by readers.py:38-39 when the file cannot
    be opened). An input error like the reference's, not a `BAMReadError`."""

    def __init__(self, bam_path: str, details: str = ""):
        message = f"BAM file appears corrupted or is not a valid BAM format: {bam_path}"
        if details:
            message += f"\n{details}"
        super().__init__(message)


class NoChrMReadsError(BAMReadError):
    def __init__(self, bam_path: str, available_chromosomes: list):
        self.available_chromosomes = available_chromosomes
        chrs = ", ".join(available_chromosomes[:10])
        super().__init__(bam_path, "No mitochondrial chromosome (chrM, MT, or M) found.\n"
                                   f"Available: {chrs}{'...' if len(available_chromosomes) > 10 else ''}")


class NoBarcodeTagsError(BAMReadError):
    def __init__(self, bam_path: str, barcode_tag: str, total_reads_checked: int):
        super().__init__(bam_path, f"No reads with barcode tag '{barcode_tag}' found "
                                   f"(checked {total_reads_checked:,} reads).")
        self.barcode_tag, self.total_reads_checked = barcode_tag, total_reads_checked


class PileupKernelError(ProcessingError):
    def __init__(self, status: int, message: str):
        super().__init__(f"mgatk2_b200 kernel path failed (status {status}): {message}")
        self.status = status
