"""Native BAM ingest (mgatk2_b200/csrc/bamio.cpp + bamio.py) against the BAM writer: every field of the
structure-of-arrays batch survives a BGZF BAM round trip, with pysam's fetch(contig) semantics (records of other
contigs skipped, unmapped mates placed on chrM kept, unplaced reads at the end ignored), with and without a .bai,
for sorted and unsorted headers, plus the validation errors of readers.py:35-61."""
import numpy as np
import pytest

from mgatk2_b200.bamio import BamFile, read_bam_chrM, write_bam
from mgatk2_b200.batch import ReadBatch
from mgatk2_b200.config import PipelineConfig
from mgatk2_b200.exceptions import BAMFormatError, BAMReadError, NoBarcodeTagsError, NoChrMReadsError
from mgatk2_b200.synth import synth_batch

FIELDS = ("pos", "tlen", "flag", "mapq", "l_seq", "n_cigar")


def assert_same_records(a: ReadBatch, b: ReadBatch):
    assert a.n_records == b.n_records
    for f in FIELDS:
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    def rec(x, i):
        r = x.record(i)
        r.pop("bc_idx")                      # the generator marks "no tag" / "not whitelisted" apart, a BAM reader cannot
        return r
    for i in list(range(0, a.n_records, max(1, a.n_records // 997))) + [a.n_records - 1]:   # blobs byte for byte (sampled + end)
        assert rec(a, i) == rec(b, i), i
    la = a.l_seq.astype(np.int64) + (a.l_seq.astype(np.int64) + 1) // 2 + 4 * a.n_cigar.astype(np.int64)
    assert int(((la + 15) // 16 * 16).sum()) == len(b.blob)


@pytest.mark.parametrize("profile,index,sorted_header", [("atac50", True, True), ("stress150", False, True), ("atac70", True, False)])
def test_round_trip(tmp_path, profile, index, sorted_header):
    n_cells = 50
    batch = synth_batch(n_cells, 30_000, profile, seed=3)
    barcodes = [f"BC{i:05d}-1" for i in range(n_cells)]
    extra = [("chr1", 100), ("chr1", 5000), ("chrX", 7), (None, -1), (None, -1)]
    path = str(tmp_path / "t.bam")
    write_bam(path, batch, barcodes, extra=extra, write_index=index, sorted_header=sorted_header, block_bytes=20000)
    cfg = PipelineConfig()
    got, mito = read_bam_chrM(path, cfg, {b: i for i, b in enumerate(barcodes)}, threads=4)
    assert mito == "chrM"
    assert_same_records(batch, got)
    np.testing.assert_array_equal(got.bc_idx, np.maximum(batch.bc_idx, -1))
    with BamFile(path) as bam:
        assert bam.references == ["chr1", "chrM", "chrX"] and bam.coordinate_sorted == sorted_header
        few, names, _ = bam.fetch("chrM", "CB", threads=1, max_records=10)
        assert few.n_records == 10 and all(isinstance(x, str) for x in names)
        other, _, _ = bam.fetch("chr1")
        assert other.n_records == 2 and other.pos.tolist() == [100, 5000]


def test_mito_names_bulk_and_missing_tags(tmp_path):
    batch = synth_batch(4, 3000, "atac50", seed=5)
    barcodes = ["A-1", "B-1", "C-1", "D-1"]
    cfg = PipelineConfig()
    p = str(tmp_path / "mt.bam")
    write_bam(p, batch, barcodes, ref_names=("1", "MT"), ref_lens=(1000, 16569), mito="MT")
    got, mito = read_bam_chrM(p, cfg, {b: i for i, b in enumerate(barcodes)})
    assert mito == "MT" and got.n_records == batch.n_records
    bulk, _ = read_bam_chrM(p, cfg, {"bulk": 0})                       # readers.py:72,100-102
    assert (bulk.bc_idx == 0).all()
    # another tag name
    cfg2 = PipelineConfig(barcode_tag="XC")
    p2 = str(tmp_path / "xc.bam")
    write_bam(p2, batch, barcodes, tag="XC")
    got2, _ = read_bam_chrM(p2, cfg2, {b: i for i, b in enumerate(barcodes)})
    np.testing.assert_array_equal(got2.bc_idx, np.maximum(batch.bc_idx, -1))
    # no barcode tag in the first 1001 records -> refused; fewer records than that -> accepted (readers.py:54-59)
    p3 = str(tmp_path / "notag.bam")
    write_bam(p3, batch, barcodes, cb_strings=[None] * batch.n_records)
    with pytest.raises(NoBarcodeTagsError):
        read_bam_chrM(p3, cfg, {b: i for i, b in enumerate(barcodes)})
    small = batch.take(np.arange(500))
    p4 = str(tmp_path / "notag_small.bam")
    write_bam(p4, small, barcodes, cb_strings=[None] * 500)
    got4, _ = read_bam_chrM(p4, cfg, {b: i for i, b in enumerate(barcodes)})
    assert (got4.bc_idx == -1).all()


def test_errors(tmp_path):
    cfg = PipelineConfig()
    batch = synth_batch(2, 200, "atac50", seed=1)
    p = str(tmp_path / "nochrm.bam")
    write_bam(p, batch, ["A-1", "B-1"], ref_names=("chr1", "chr2"), ref_lens=(1000, 16569), mito="chr2")
    with pytest.raises(NoChrMReadsError):
        read_bam_chrM(p, cfg, {"A-1": 0})
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"this is not a bam file, not even close........")
    with pytest.raises(BAMFormatError):
        read_bam_chrM(str(junk), cfg, {"A-1": 0})
    # truncated file: an error, never a silently short batch
    good = str(tmp_path / "good.bam")
    write_bam(good, synth_batch(2, 5000, "atac50", seed=2), ["A-1", "B-1"], write_index=False, block_bytes=8000)
    data = open(good, "rb").read()
    cut = tmp_path / "cut.bam"
    cut.write_bytes(data[: len(data) // 2])
    with pytest.raises(BAMReadError):
        read_bam_chrM(str(cut), cfg, {"A-1": 0, "B-1": 1})
    # QUAL absent (0xFF) on a record that passes the filters (readers.py:158)
    b = ReadBatch.from_records([dict(pos=10, flag=0, mapq=60, seq="ACGT" * 5, qual=[255] * 20, cigar=[(0, 20)], bc_idx=0)])
    q = str(tmp_path / "noqual.bam")
    write_bam(q, b, ["A-1"])
    with pytest.raises(BAMReadError):
        read_bam_chrM(q, cfg, {"A-1": 0})
    # ... but a QUAL-less record that dedup drops never becomes a SimpleRead: the reference accepts the file (readers.py:146-158)
    good_first = dict(pos=10, flag=0, mapq=60, seq="ACGT" * 5, qual=[30] * 20, cigar=[(0, 20)], tlen=77, bc_idx=0)
    b2 = ReadBatch.from_records([good_first, dict(good_first, qual=[255] * 20)])
    q2 = str(tmp_path / "noqual_dup.bam")
    write_bam(q2, b2, ["A-1"])
    got, _ = read_bam_chrM(q2, cfg, {"A-1": 0})
    assert got.n_records == 2
    cfg_nodedup = PipelineConfig(skip_deduplication=True)
    with pytest.raises(BAMReadError):
        read_bam_chrM(q2, cfg_nodedup, {"A-1": 0})


def test_many_distinct_barcodes_first_appearance_order(tmp_path):
    """The tag table of the decoder (open addressing, one per thread, merged in record order): thousands of distinct
    18-character barcodes, more than one thread, table growth; ids must follow first appearance in the file and the
    padding bytes between blobs must be zero whatever memory the decoder was handed."""
    n_cells, n = 6000, 24_000
    batch = synth_batch(n_cells, n, "atac50", seed=11)
    rng = np.random.default_rng(5)
    barcodes = ["".join("ACGT"[k] for k in rng.integers(0, 4, 16)) + "-1" for _ in range(n_cells)]
    assert len(set(barcodes)) == n_cells
    cb = [barcodes[int(c)] if c >= 0 else (None if i % 2 else "NNNNNNNNNNNNNNNN-1") for i, c in enumerate(batch.bc_idx)]
    path = str(tmp_path / "many.bam")
    write_bam(path, batch, barcodes, cb_strings=cb, block_bytes=30000)
    with BamFile(path) as bam:
        for threads in (1, 4):
            got, names, _ = bam.fetch("chrM", "CB", threads=threads)
            first = []
            seen = set()
            for s in cb:
                if s is not None and s not in seen:
                    seen.add(s)
                    first.append(s)
            assert names == first
            idx = {s: k for k, s in enumerate(first)}
            want = np.array([idx[s] if s is not None else -1 for s in cb], dtype=np.int32)
            np.testing.assert_array_equal(got.bc_idx, want)
            sizes = 4 * got.n_cigar.astype(np.int64) + (got.l_seq.astype(np.int64) + 1) // 2 + got.l_seq.astype(np.int64)
            starts = got.blob_off.astype(np.int64) * 16
            for s0, sz in zip(starts[:: max(1, n // 500)], sizes[:: max(1, n // 500)]):
                pad = got.blob[s0 + sz: s0 + (sz + 15) // 16 * 16]
                assert not pad.any()


@pytest.mark.parametrize("max_records,index", [(1500, True), (7000, False), (100000, True), (1, True)])
def test_parts_equal_one_read(tmp_path, max_records, index):
    """iter_bam_chrM (mgatk_bam_fetch + mgatk_bam_fetch_more): the parts, in order, are the records of the one-shot
    read; every border lies between two different reference_start values (what the accumulating device path needs);
    barcode indices stay consistent over the parts although the tag table grows."""
    from mgatk2_b200.bamio import iter_bam_chrM
    n_cells, n = 300, 20_000 if max_records > 1 else 300
    batch = synth_batch(n_cells, n, "stress150", seed=21)
    barcodes = [f"CELL{i:04d}AAAACCCCGGGG-1" for i in range(n_cells)]
    path = str(tmp_path / "p.bam")
    write_bam(path, batch, barcodes, extra=[("chr1", 5), ("chrX", 9), (None, -1)], write_index=index, block_bytes=25000)
    cfg = PipelineConfig()
    wl = {b: i for i, b in enumerate(barcodes)}
    whole, _ = read_bam_chrM(path, cfg, wl, threads=3)
    parts = list(iter_bam_chrM(path, cfg, wl, max_records=max_records, threads=3))
    assert sum(p.n_records for p in parts) == whole.n_records
    if max_records < n:
        assert len(parts) > 1
    for a, b in zip(parts[:-1], parts[1:]):
        assert a.n_records and a.pos[-1] < b.pos[0]
    joined = ReadBatch.concat(parts)
    assert_same_records(whole, joined)
    np.testing.assert_array_equal(joined.bc_idx, whole.bc_idx)
    np.testing.assert_array_equal(joined.blob, whole.blob)
    np.testing.assert_array_equal(joined.blob_off, whole.blob_off)
    # slices of a batch are views with their own offsets
    s = whole.slice(10, 200)
    assert s.n_records == 190 and s.blob_off[0] == 0 and s.record(0) == whole.record(10) and s.record(189) == whole.record(199)


def test_reference_bai_known_answers():
    """A pin at the htslib boundary that does not come from this repo's own writer: the index the reference ships next
    to its (missing) test BAM, `tests/outs/possorted_bam.bam.bai`, copied to tests/golden/reference_files/. samtools
    wrote it; the native parser must read what SURVEY §4 recovered from it: 194 references, only reference 22 (chrM of
    the 10x GRCh38 layout) populated, 228 149 mapped + 1 619 unmapped-placed records (= the ~230 k records
    `fetch("chrM")` returns, readers.py:85-93), records from virtual offset 2959 << 16 (right behind the header) to
    about 11.6 MB. Cross-checked against a plain struct walk of the same bytes."""
    import os
    import struct
    from mgatk2_b200.bamio import inspect_bai
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_files", "possorted_bam.bam.bai")
    raw = open(path, "rb").read()
    assert raw[:4] == b"BAI\1" and struct.unpack_from("<i", raw, 4)[0] == 194
    o, populated = 8, {}
    for r in range(194):                                          # independent walk (SAM spec 5.2)
        n_bin = struct.unpack_from("<i", raw, o)[0]; o += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", raw, o); o += 8
            bins[b] = [struct.unpack_from("<QQ", raw, o + 16 * c) for c in range(n_chunk)]; o += 16 * n_chunk
        n_intv = struct.unpack_from("<i", raw, o)[0]; o += 4 + 8 * n_intv
        if n_bin:
            populated[r] = (bins, n_intv)
    assert list(populated) == [22] and struct.unpack_from("<Q", raw, o)[0] == 0 and o + 8 == len(raw)
    bins, n_intv = populated[22]
    real = [c for b, cs in bins.items() if b != 37450 for c in cs]
    got = inspect_bai(path, 22)
    assert got == dict(n_ref=194, n_bin=3, n_chunk=len(real), n_intv=n_intv, min_voff=min(c[0] for c in real), has_meta=1,
                       ref_beg=bins[37450][0][0], ref_end=bins[37450][0][1], n_mapped=228149, n_unmapped=1619)
    assert got["n_chunk"] == 10 and got["n_intv"] == 2
    assert got["min_voff"] == got["ref_beg"] == 2959 << 16 and got["ref_end"] >> 16 == 11588294
    for r in (0, 21, 23, 193):                                    # unpopulated references: parsed, no chunk
        e = inspect_bai(path, r)
        assert e["n_ref"] == 194 and e["n_bin"] == 0 and e["min_voff"] == -1 and e["has_meta"] == 0
    assert inspect_bai(path, 194) is None and inspect_bai(path, -1) is None
    assert inspect_bai(__file__, 0) is None                       # not a BAI
