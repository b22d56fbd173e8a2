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


def _reader_result(batch, barcodes, params, **kw):
    from mgatk2_b200 import BAMReader
    reader = BAMReader("in-memory.bam", make_config(params), set(barcodes), barcode_list=barcodes, batch=batch, **kw)
    rbb, stats = reader.collect_reads_by_barcode()
    return reader, rbb, stats


def _same_cells(rbb_a, rbb_b):
    """Same barcodes in the same order with the same planes / QC rows, whatever the column layout of either result."""
    assert list(rbb_a) == list(rbb_b)
    ra, rb = rbb_a.result, rbb_b.result
    ia = np.array([rbb_a[b].index for b in rbb_a], np.int64)
    ib = np.array([rbb_b[b].index for b in rbb_b], np.int64)
    for k in range(11):
        np.testing.assert_array_equal(ra.plane(k)[ia], rb.plane(k)[ib], err_msg=f"plane {k}")
    np.testing.assert_array_equal(ra.cell_qc[ia], rb.cell_qc[ib])
    np.testing.assert_array_equal(ra.base_totals, rb.base_totals)


def test_columns_only_for_observed_barcodes(golden_dir):
    """A whitelist far longer than the barcodes on the contig (10x lists have 737 k entries): planes are allocated for
    the observed ones only, results unchanged (the reference only does set membership, readers.py:104-111)."""
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    _, rbb_small, stats_small = _reader_result(batch, barcodes, params)
    pad = [f"PAD{i:07d}-1" for i in range(300_000)]
    long_list = pad[:150_000] + barcodes + pad[150_000:]
    import copy
    shifted = copy.copy(batch)
    shifted.bc_idx = np.where(batch.bc_idx >= 0, batch.bc_idx + 150_000, batch.bc_idx).astype(np.int32)
    _, rbb_long, stats_long = _reader_result(shifted, long_list, params)
    assert stats_small == stats_long
    assert len(rbb_long.result.planes) <= len(barcodes)
    assert [long_list[c] for c in rbb_long.result.columns] == [b for b in barcodes if b in set(long_list[c] for c in rbb_long.result.columns)]
    _same_cells(rbb_small, rbb_long)


@pytest.mark.parametrize("name", ["synth_run_default", "synth_stress150"])
def test_streamed_through_the_seam(golden_dir, name):
    """A contig longer than `max_batch_records` goes through the seam in parts (device-resident accumulation, prefetch
    thread) and gives what one batch gives."""
    d, batch, barcodes, params = load_golden(f"{golden_dir}/{name}.npz")
    r1, rbb_one, stats_one = _reader_result(batch, barcodes, params)
    r2, rbb_str, stats_str = _reader_result(batch, barcodes, params, max_batch_records=max(batch.n_records // 7, 1))
    assert not r1.streamed and r2.streamed
    assert stats_one == stats_str
    _same_cells(rbb_one, rbb_str)


def test_two_handles_split_the_cells(golden_dir, tmp_path):
    """`devices=[...]`: the columns are cut into one range per entry, each counted through its own handle on its own host
    thread, written into one result. With one GPU in the box the same device is listed twice (two handles); one batch and
    streamed; the text outputs still equal the reference's byte for byte."""
    import torch
    from mgatk2_b200 import MtDNAPipeline
    from mgatk2_b200.bamio import write_bam
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    _, rbb_one, stats_one = _reader_result(batch, barcodes, params)
    _, rbb_two, stats_two = _reader_result(batch, barcodes, params, devices=devs)
    assert stats_one == stats_two
    _same_cells(rbb_one, rbb_two)
    _, rbb_two_s, stats_two_s = _reader_result(batch, barcodes, params, devices=devs, max_batch_records=max(batch.n_records // 5, 1))
    assert stats_one == stats_two_s
    _same_cells(rbb_one, rbb_two_s)
    bam = str(tmp_path / "possorted_bam.bam")
    write_bam(bam, batch, barcodes, write_index=False)
    out = tmp_path / "run"
    MtDNAPipeline(bam, barcodes, out, make_config(params), devices=devs).run()
    for base in ("A", "C", "G", "T", "coverage"):
        assert gzip.open(out / "output" / f"output.{base}.txt.gz").read() == d[f"txt_{base}"].tobytes(), base
    assert (out / "output" / "chrM_refAllele.txt").read_bytes() == d["txt_refAllele"].tobytes()
    assert (out / "qc" / "cell_stats.csv").read_bytes() == d["txt_cell_stats"].tobytes()


def test_filter_strand_bias_above_uint16():
    """PileupGenerator.filter_strand_bias on a dict with counts beyond 65535 (bulk depths): same device path as any other
    dict (32-bit planes), same rule as pileup.py:128-154 - IEEE double, strict >, coverage from what is left, positions
    without coverage dropped with their Tn5 counts."""
    from mgatk2_b200 import PileupGenerator, PipelineConfig
    cfg = PipelineConfig(max_strand_bias=0.8)

    def entry(a=(0, 0), c=(0, 0), g=(0, 0), t=(0, 0), tn5=(0, 0)):
        d = {"tn5_cuts_fwd": tn5[0], "tn5_cuts_rev": tn5[1]}
        for base, (f, r) in zip("ACGT", (a, c, g, t)):
            d[base], d[f"{base}_fwd"], d[f"{base}_rev"] = f + r, f, r
        d["depth"] = sum(d[b] for b in "ACGT")
        return d
    raw = {10: entry(a=(400_000, 100_000), c=(70_000, 3), tn5=(5, 6)),          # A: bias 0.8 exactly, kept; C: dropped
           11: entry(g=(100_001, 25_000), tn5=(9, 9)),                          # 0.80000... > 0.8: dropped, position vanishes
           12: entry(t=(3, 3))}
    out = PileupGenerator(cfg).filter_strand_bias(raw)
    assert sorted(out) == [10, 12]
    assert out[10]["A_fwd"] == 400_000 and out[10]["A"] == 500_000 and out[10]["C"] == 0 and out[10]["depth"] == 500_000
    assert (out[10]["tn5_cuts_fwd"], out[10]["tn5_cuts_rev"]) == (5, 6) and out[12]["depth"] == 6


def test_pipeline_writes_the_default_hdf5_layout(golden_dir, tmp_path):
    """MtDNAPipeline(output_format="hdf5") - the default of `mgatk2 run` (cli/options.py:14-78): counts.h5 / metadata.h5
    hold what the unmodified reference computed for the same records (golden: its per-cell pileups), laid out and typed as
    IncrementalHDF5Writer does (writers.py:60-134,187-229)."""
    from mgatk2_b200 import MtDNAPipeline
    from mgatk2_b200.bamio import write_bam
    from mgatk2_b200.h5lite import H5Reader
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    bam = str(tmp_path / "possorted_bam.bam")
    write_bam(bam, batch, barcodes, write_index=False)
    out = tmp_path / "run"
    res = MtDNAPipeline(bam, barcodes, out, make_config(params), output_format="hdf5").run()
    alive = d["exp_alive"].astype(bool)
    assert res["cells_passed_qc"] == int(alive.sum())
    c = H5Reader(out / "output" / "counts.h5")
    assert c.attrs == {"n_cells": len(barcodes), "n_positions": 16569, "mito_chr": "chrM"}
    assert c.objects["barcode"].read().tolist() == [b.encode() for b in barcodes]
    for bi, base in enumerate("ACGT"):
        for si, strand in enumerate(("fwd", "rev")):
            want = np.minimum(d["exp_counts"][:, :, bi, si], 65535).astype(np.uint16) * alive[:, None]
            np.testing.assert_array_equal(c.objects[f"{base}_{strand}"].read(), want.T, err_msg=f"{base}_{strand}")
    for si, strand in enumerate(("fwd", "rev")):
        want = np.minimum(d["exp_tn5"][:, :, si], 65535).astype(np.uint16) * alive[:, None]
        np.testing.assert_array_equal(c.objects[f"tn5_cuts_{strand}"].read(), want.T)
    m = H5Reader(out / "output" / "metadata.h5")
    cov = d["exp_coverage"] * alive[:, None]
    np.testing.assert_array_equal(m.objects["coverage"].read(), np.minimum(cov, 65535).astype(np.uint16).T)
    dep = [cov[k][cov[k] > 0] for k in range(len(barcodes))]
    np.testing.assert_array_equal(m.objects["mean_depth"].read(), np.array([x.mean() if len(x) else 0 for x in dep], np.float32))
    np.testing.assert_array_equal(m.objects["median_depth"].read(), np.array([np.median(x) if len(x) else 0 for x in dep], np.float32))
    np.testing.assert_array_equal(m.objects["max_depth"].read(), np.array([min(x.max(), 65535) if len(x) else 0 for x in dep], np.uint16))
    np.testing.assert_array_equal(m.objects["total_bases"].read(), np.array([x.sum() if len(x) else 0 for x in dep], np.float32))
    np.testing.assert_array_equal(m.objects["genome_coverage"].read(), np.array([len(x) / 16569 * 100 for x in dep], np.float32))
    assert m.objects["reference"].read().tobytes() == b"".join(l.split(b"\t")[1] for l in d["txt_refAllele"].tobytes().splitlines()[1:])
    assert (out / "qc" / "cell_stats.csv").read_bytes() == d["txt_cell_stats"].tobytes()


def test_bam_file_streamed_through_the_seam(golden_dir, tmp_path):
    """A BAM longer than `max_batch_records` is decoded in parts (native reader, parts cut on reference_start borders, next
    part decoded on the prefetch thread) and counted through the accumulating device path: same result as the one-shot
    read of the same file."""
    from mgatk2_b200 import BAMReader
    from mgatk2_b200.bamio import write_bam
    d, batch, barcodes, params = load_golden(f"{golden_dir}/synth_run_default.npz")
    path = str(tmp_path / "possorted_bam.bam")
    write_bam(path, batch, barcodes, extra=[("chr1", 5), ("chrX", 9), (None, -1)])
    one = BAMReader(path, make_config(params), set(barcodes), barcode_list=barcodes)
    rbb_one, stats_one = one.collect_reads_by_barcode()
    parts = BAMReader(path, make_config(params), set(barcodes), barcode_list=barcodes, max_batch_records=max(batch.n_records // 6, 1))
    rbb_parts, stats_parts = parts.collect_reads_by_barcode()
    assert not one.streamed and parts.streamed
    assert stats_one == stats_parts
    _same_cells(rbb_one, rbb_parts)
