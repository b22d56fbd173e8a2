"""The host-side mirror of the reference seam (BAMReader -> CellProcessor -> writer; PileupGenerator;
process_barcode_worker) on the GPU, against the reference-generated goldens."""
import gzip
import types

import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def make_config(params):
    from mgatk2_b200 import PipelineConfig
    cfg = PipelineConfig(min_baseq=params["min_baseq"], min_mapq=params["min_mapq"],
                         max_strand_bias=params["max_strand_bias"], skip_deduplication=params["dedup_mode"] == 2,
                         use_fragment_length_dedup=params["dedup_mode"] == 0,
                         min_reads_per_cell=params["min_reads_per_cell"], sequential=True)
    cfg.quality.min_distance_from_end = params["min_distance_from_end"]
    return cfg


@pytest.mark.parametrize("name", ["synth_run_default", "synth_tenx"])
def test_seam_text_outputs_match_reference(golden_dir, tmp_path, name):
    """MtDNAPipeline.run()'s three calls (pipeline.py:80-81,102-103,113) with our classes: the files the
    reference's IncrementalTextWriter wrote for the same records are reproduced byte for byte."""
    from mgatk2_b200 import BAMReader, CellProcessor
    from mgatk2_b200.writers import DenseTextWriter
    d, batch, barcodes, params = load_golden(f"{golden_dir}/{name}.npz")
    cfg = make_config(params)
    reader = BAMReader("in-memory.bam", cfg, set(barcodes), barcode_list=barcodes, batch=batch)
    reads_by_barcode, stats = reader.collect_reads_by_barcode()
    total, filtered, n_bc, dl, dp = (int(x) for x in d["exp_stats"])
    assert stats == {"total_reads": total, "filtered_reads": filtered, "n_barcodes": n_bc,
                     "duplicate_reads_with_length": dl, "duplicate_reads_position_only": dp}
    assert [barcodes.index(b) for b in reads_by_barcode] == \
        [c for c in dict.fromkeys(batch.bc_idx[((batch.flag & 0x904) == 0) & (batch.bc_idx >= 0)].tolist())]
    writer = DenseTextWriter(tmp_path, cfg, barcodes)
    results = CellProcessor(cfg, tmp_path).process_cells_progressive(reads_by_barcode, writer)
    writer.finalize(tmp_path / "qc")
    assert [barcodes.index(r["barcode"]) for r in results] == d["exp_cell_order"].tolist()
    for base in ("A", "C", "G", "T", "coverage"):
        got = gzip.open(tmp_path / "output" / f"output.{base}.txt.gz").read()
        assert got == d[f"txt_{base}"].tobytes(), base
    assert (tmp_path / "output" / "output.depthTable.txt").read_bytes() == d["txt_depthTable"].tobytes()
    assert (tmp_path / "output" / "chrM_refAllele.txt").read_bytes() == d["txt_refAllele"].tobytes()
    assert (tmp_path / "qc" / "cell_stats.csv").read_bytes() == d["txt_cell_stats"].tobytes()


def test_process_barcode_worker_dicts(golden_dir):
    """Result dicts (processors.py:41-52) for every cell: same keys, same integers, same float64 QC."""
    from mgatk2_b200 import BAMReader
    from mgatk2_b200.processors import process_barcode_worker
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_stress150.npz")
    cfg = make_config(params)
    rbb, _ = BAMReader("x.bam", cfg, set(barcodes), barcode_list=barcodes, batch=batch).collect_reads_by_barcode()
    seen = 0
    for bc, reads in rbb.items():
        res = process_barcode_worker((bc, reads, cfg))
        c = barcodes.index(bc)
        if not d["exp_alive"][c]:
            assert res is None
            continue
        seen += 1
        assert set(res) == {"barcode", "pileup", "n_reads", "qc"}
        pos = sorted(res["pileup"])
        assert pos == np.nonzero(d["exp_coverage"][c])[0].tolist()
        for p in pos[:: max(1, len(pos) // 50)]:
            e = res["pileup"][p]
            assert e["depth"] == d["exp_coverage"][c, p]
            assert (e["tn5_cuts_fwd"], e["tn5_cuts_rev"]) == tuple(d["exp_tn5"][c, p])
            for bi, base in enumerate("ACGT"):
                assert (e[f"{base}_fwd"], e[f"{base}_rev"]) == tuple(d["exp_counts"][c, p, bi])
                assert e[base] == int(d["exp_counts"][c, p, bi].sum())
        q = res["qc"]
        assert (q["total_reads"], q["total_fragments"], q["mean_depth"]) == tuple(d["exp_qc"][c])
        assert q["coverage_breadth"] == d["exp_breadth"][c]
    assert seen == int(d["exp_alive"].sum())


def test_pileup_generator_two_step(golden_dir):
    """generate_pileup() then filter_strand_bias() (two device calls) equals the fused path, and
    generate_pileup() alone keeps uncovered Tn5 sites like pileup.py:105-106."""
    from mgatk2_b200 import PileupGenerator
    d, batch, barcodes, params = load_golden(f"{golden_dir}/kat4_strand_bias.npz")
    cfg = make_config(params)
    reads = []
    for i in range(batch.n_records):
        r = batch.record(i)
        reads.append(types.SimpleNamespace(reference_start=r["pos"], is_reverse=bool(r["flag"] & 16),
                                           mapping_quality=r["mapq"], query_sequence=r["seq"].encode(),
                                           query_qualities=np.array(r["qual"], np.int8), cigar=r["cigar"],
                                           is_paired=bool(r["flag"] & 1), template_length=r["tlen"]))
    # the reference de-duplicates before the pileup: drop what readers.py would drop (position-only dups here are kept: len mode)
    gen = PileupGenerator(cfg)
    raw = gen.generate_pileup(reads)
    assert raw[400]["tn5_cuts_fwd"] == 5 and raw[400]["C_fwd"] == 5 and raw[500]["G_fwd"] == 3      # unfiltered
    flt = gen.filter_strand_bias(raw)
    assert sorted(flt) == list(range(300, 310))
    assert flt[300]["tn5_cuts_fwd"] == 4 and flt[309]["tn5_cuts_rev"] == 1 and flt[300]["depth"] == 5


def test_bam_file_through_the_seam(golden_dir, tmp_path):
    """BAMReader on a real BGZF BAM (native ingest, csrc/bamio.cpp) gives what the in-memory batch gives, and the
    reference's constructor-time validation (readers.py:35-61) is in place."""
    from mgatk2_b200 import BAMReader
    from mgatk2_b200.bamio import write_bam
    from mgatk2_b200.exceptions import BAMReadError
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    path = str(tmp_path / "possorted_bam.bam")
    write_bam(path, batch, barcodes, extra=[("chr1", 5), ("chrX", 9), (None, -1)])
    cfg = make_config(params)
    reads_a, stats_a = BAMReader(path, cfg, set(barcodes), barcode_list=barcodes).collect_reads_by_barcode()
    reads_b, stats_b = BAMReader("in-memory.bam", make_config(params), set(barcodes), barcode_list=barcodes,
                                 batch=batch).collect_reads_by_barcode()
    assert stats_a == stats_b and list(reads_a) == list(reads_b)
    np.testing.assert_array_equal(reads_a.result.planes, reads_b.result.planes)
    np.testing.assert_array_equal(reads_a.result.cell_qc, reads_b.result.cell_qc)
    with pytest.raises(BAMReadError):
        BAMReader(str(tmp_path / "missing.bam"), cfg, set(barcodes))


def test_pipeline_from_bam_to_text_outputs(golden_dir, tmp_path):
    """MtDNAPipeline.run() on a BGZF BAM: the files the reference's IncrementalTextWriter wrote for the same records
    come out byte for byte (pipeline.py:76-181)."""
    from mgatk2_b200 import MtDNAPipeline, run_pipeline
    from mgatk2_b200.bamio import write_bam
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    bam = str(tmp_path / "possorted_bam.bam")
    write_bam(bam, batch, barcodes, write_index=False)
    out = tmp_path / "run"
    res = MtDNAPipeline(bam, barcodes, out, make_config(params)).run()
    assert res["cells_passed_qc"] == int(d["exp_alive"].sum()) and res["cells_processed"] >= res["cells_passed_qc"]
    for base in ("A", "C", "G", "T", "coverage"):
        assert gzip.open(out / "output" / f"output.{base}.txt.gz").read() == d[f"txt_{base}"].tobytes(), base
    assert (out / "output" / "output.depthTable.txt").read_bytes() == d["txt_depthTable"].tobytes()
    assert (out / "output" / "chrM_refAllele.txt").read_bytes() == d["txt_refAllele"].tobytes()
    assert (out / "qc" / "cell_stats.csv").read_bytes() == d["txt_cell_stats"].tobytes()
    assert "cells_passed_qc" in (out / "qc" / "summary.txt").read_text()
    # whitelist discovered from the BAM itself (pipeline.py:214-224): at least the cells with >= 10 countable reads
    res2 = run_pipeline(bam, str(tmp_path / "run2"), barcode_file=None, min_barcode_reads=10)
    assert res2["cells_passed_qc"] > 0
