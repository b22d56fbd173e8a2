"""Stand-in for pysam (absent from this image) so that the UNMODIFIED reference in baseline/_ref can be imported and run.

Only what the reference's hot path touches exists: `AlignmentFile(path).fetch(contig)` yields read objects with the pysam
attribute names (`processing/readers.py:85-165`), `.references`, context-manager protocol. The reads come from a
registry filled by `baseline/ref_harness.py` (objects built BEFORE any timing starts: pysam decodes records in C, so
decoding must not be charged to the reference)."""
from __future__ import annotations

import array

REGISTRY: dict[str, tuple[list, tuple]] = {}      # path -> (reads, references)


class AlignedSegment:
    __slots__ = ("reference_start", "mapping_quality", "query_sequence", "query_qualities", "cigartuples",
                 "template_length", "flag", "_cb")

    def __init__(self, rec: dict, barcode):
        self.reference_start = rec["pos"]
        self.mapping_quality = rec["mapq"]
        self.query_sequence = rec["seq"]
        self.query_qualities = array.array("B", rec["qual"])
        self.cigartuples = rec["cigar"] or None
        self.template_length = rec["tlen"]
        self.flag = rec["flag"]
        self._cb = barcode

    is_paired = property(lambda s: bool(s.flag & 0x1))
    is_proper_pair = property(lambda s: bool(s.flag & 0x2))
    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    is_reverse = property(lambda s: bool(s.flag & 0x10))
    is_secondary = property(lambda s: bool(s.flag & 0x100))
    is_duplicate = property(lambda s: bool(s.flag & 0x400))
    is_supplementary = property(lambda s: bool(s.flag & 0x800))

    def has_tag(self, tag):
        return tag == "CB" and self._cb is not None

    def get_tag(self, tag):
        if not self.has_tag(tag):
            raise KeyError(tag)
        return self._cb


class AlignmentFile:
    def __init__(self, path, mode="rb", **kw):
        self._reads, self.references = REGISTRY[str(path)]

    def fetch(self, contig=None, **kw):
        return iter(self._reads)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def index(*a, **k):
    return None
