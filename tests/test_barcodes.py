"""Whitelist loaders and barcode discovery (mgatk2_b200/barcodes.py) against the reference's rules
(utils.py:14-69, barcode_extraction.py:12-46, pipeline.py:214-230)."""
import pytest

from mgatk2_b200.bamio import write_bam
from mgatk2_b200.barcodes import extract_barcodes_from_bam, load_barcodes, load_singlecell_csv
from mgatk2_b200.batch import ReadBatch
from mgatk2_b200.exceptions import InvalidInputError


def test_singlecell_csv(tmp_path):
    p = tmp_path / "singlecell.csv"
    p.write_text("barcode,total,frac,is__cell_barcode,excluded_reason,note\n"
                 "NO_BARCODE,10,0.5,0,x,a\nCCC-1,7,0.25,1,,b\nAAA-1,,x,1,1,\nBBB-1,3,1e-3,0,,c\n")
    barcodes, meta = load_singlecell_csv(str(p))
    assert barcodes == ["CCC-1", "AAA-1"]                       # file order, only is__cell_barcode == 1
    assert meta["total"] == [7, 0] and meta["frac"] == [0.25, "x"] and meta["excluded_reason"] == ["", "1"]
    assert meta["is__cell_barcode"] == [1, 1] and meta["note"] == ["b", 0]
    assert load_singlecell_csv(None) == (None, None)
    bad = tmp_path / "bad.csv"
    bad.write_text("barcode,total\nA,1\n")
    with pytest.raises(InvalidInputError):
        load_singlecell_csv(str(bad))
    none = tmp_path / "none.csv"
    none.write_text("barcode,is__cell_barcode\nA,0\n")
    with pytest.raises(InvalidInputError):
        load_singlecell_csv(str(none))
    with pytest.raises(InvalidInputError):
        load_singlecell_csv(str(tmp_path / "missing.csv"))


def test_extract_and_dispatch(tmp_path):
    recs, cbs = [], []
    plan = [("GGG-1", 12, 0), ("AAA-1", 10, 0), ("TTT-1", 9, 0), ("CCC-1", 15, 0x400), ("CCC-1", 4, 0), ("UUU-1", 11, 0x4), (None, 20, 0)]
    for cb, n, flag in plan:
        for _ in range(n):
            recs.append(dict(pos=100 + len(recs), flag=flag, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=0))
            cbs.append(cb)
    bam = str(tmp_path / "x.bam")
    write_bam(bam, ReadBatch.from_records(recs), ["unused"], cb_strings=cbs)
    assert extract_barcodes_from_bam(bam) == ["AAA-1", "GGG-1"]             # >= 10 countable reads, sorted
    assert extract_barcodes_from_bam(bam, min_reads=4) == ["AAA-1", "CCC-1", "GGG-1", "TTT-1"]
    assert load_barcodes(None, bam) == (["AAA-1", "GGG-1"], None)
    with pytest.raises(InvalidInputError):
        load_barcodes(None, bam, min_barcode_reads=1000)
    txt = tmp_path / "barcodes.tsv"
    txt.write_text("B-1\n\n A-1 \nB-1\n")
    assert load_barcodes(str(txt), bam) == (["B-1", "A-1", "B-1"], None)      # order and duplicates kept (last index wins downstream)
