"""`counts.h5` / `metadata.h5` without libhdf5 (mgatk2_b200/h5lite.py, writers.DenseHDF5Writer).

No HDF5 library exists in this image, so the pin is the reference's own committed run output
(/root/reference/tests/run_hdf5_output/output/{counts,metadata}.h5, copied to tests/golden/ref_hdf5/): the reader must
parse what libhdf5 wrote - every checksum verified, names / shapes / types as R/mgatk2_functions.R:18-61 reads them - and
the writer's files must come back through the very same code paths."""
import os

import numpy as np
import pytest

from mgatk2_b200.h5lite import H5Reader, H5Writer, lookup3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref_hdf5")
PLANES = ["A_fwd", "A_rev", "C_fwd", "C_rev", "G_fwd", "G_rev", "T_fwd", "T_rev", "tn5_cuts_fwd", "tn5_cuts_rev"]


def test_lookup3_known_answers():
    """Bob Jenkins' self-test values of hashlittle (lookup3.c driver5) - the checksum of every HDF5 metadata block."""
    assert lookup3(b"") == 0xDEADBEEF
    assert lookup3(b"", 0xDEADBEEF) == 0xBD5B7DDE
    assert lookup3(b"Four score and seven years ago", 0) == 0x17770551
    assert lookup3(b"Four score and seven years ago", 1) == 0xCD628161


def test_reads_the_reference_files():
    """libhdf5's own output for the reference's test run: 141 barcodes x 16569 positions."""
    c = H5Reader(os.path.join(REF, "counts.h5"))
    assert c.version == 3 and c.attrs == {"n_cells": 141, "n_positions": 16569, "mito_chr": "chrM"}
    assert set(c.objects) == set(PLANES) | {"barcode"}
    assert c.structures["OHDR"] == 12 and c.structures["dense link storage"] == 1     # a root group libhdf5 stored densely
    bc = c.objects["barcode"].read()
    assert bc.shape == (141,) and bc.dtype == np.dtype("S18") and bc[0] == b"AAAGGTACACTAAGGC-1"
    planes = {}
    for name in PLANES:
        ds = c.objects[name]
        assert ds.shape == (16569, 141) and ds.dtype == np.uint16 and ds.filters == [(1, (4,))]     # gzip level 4
        planes[name] = ds.read()
    assert c.structures["FAHD"] == 10 and c.structures["FADB"] == 10                   # chunks (1000, 100): 17 x 2 per dataset
    m = H5Reader(os.path.join(REF, "metadata.h5"))
    assert m.attrs == {"mito_chr": "chrM", "mito_length": 16569}
    cov = m.objects["coverage"].read()
    # the files are consistent with each other the way the pipeline makes them (pileup.py:150, writers.py:187-197)
    np.testing.assert_array_equal(cov, sum(planes[n].astype(np.uint32) for n in PLANES[:8]).astype(np.uint16))
    covered = (cov > 0).sum(0)
    np.testing.assert_array_equal(m.objects["max_depth"].read(), cov.max(0))
    np.testing.assert_allclose(m.objects["total_bases"].read(), cov.sum(0, dtype=np.float64), rtol=1e-6)
    np.testing.assert_allclose(m.objects["genome_coverage"].read(), covered / 16569 * 100, rtol=1e-6)
    np.testing.assert_allclose(m.objects["mean_depth"].read(), cov.sum(0) / np.maximum(covered, 1), rtol=1e-6)
    med = np.array([np.median(cov[:, j][cov[:, j] > 0]) if covered[j] else 0 for j in range(141)], np.float32)
    np.testing.assert_array_equal(m.objects["median_depth"].read(), med)
    ref = m.objects["reference"].read()
    assert ref.shape == (16569,) and ref.dtype == np.dtype("S1") and set(ref.tolist()) <= {b"A", b"C", b"G", b"T", b"N"}
    tot = np.stack([planes[f"{b}_fwd"].astype(np.int64).sum(1) + planes[f"{b}_rev"].sum(1) for b in "ACGT"], 1)
    want = np.where(tot.max(1) > 0, np.array(list("ACGT"))[tot.argmax(1)], "N").astype("S1")      # writers.py:341-346
    np.testing.assert_array_equal(ref, want)
    assert "barcode_metadata" in m.groups and m.objects["barcode_metadata/barcode"].read()[0] == bc[0]
    cols = [k for k in m.objects if k.startswith("barcode_metadata/")]
    assert len(cols) == 19 and all(m.objects[k].read().shape == (141,) for k in cols)
    assert m.structures["FHIB"] == 1 and m.structures["single-chunk index"] >= 19      # a dense group with an indirect root block


def test_writer_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    a = (rng.random((16569, 141)) < 0.05) * rng.integers(1, 65536, (16569, 141))
    a = a.astype(np.uint16)
    many = rng.integers(0, 3, (1650, 730)).astype(np.uint16)                            # 17 x 73 = 1241 chunks: a paged index
    sparse = np.zeros((3000, 500), np.uint16)
    sparse[1200:1300, 250:260] = 7
    path = tmp_path / "t.h5"
    with H5Writer(path) as f:
        f.attr("n_cells", 141)
        f.attr("mito_chr", "chrM")
        f.dataset("barcode", np.array([f"BC{i:05d}-1" for i in range(141)], dtype="S"))
        f.dataset("A_fwd", a, chunks=(1000, 100), gzip=4)
        f.dataset("many", many, chunks=(100, 10), gzip=1)
        f.dataset("sparse", None, chunks=(1000, 100), gzip=4, shape=sparse.shape, dtype=np.uint16,
                  chunk_source=lambda i0, i1, j0, j1: sparse[i0:i1, j0:j1] if sparse[i0:i1, j0:j1].any() else None)
        f.dataset("mean_depth", rng.random(141).astype(np.float32))
        f.dataset("reference", np.array(list("ACGTN" * 20), dtype="S1"), chunks=(100,), gzip=4)
        f.dataset("empty", np.zeros((16569, 0), np.uint16), chunks=(1000, 1), gzip=4)
        g = f.group("barcode_metadata")
        f.dataset("total", np.arange(141), chunks=(141,), gzip=4, parent=g)
    r = H5Reader(path)                                                                  # verifies all checksums on the way
    assert r.attrs == {"n_cells": 141, "mito_chr": "chrM"} and isinstance(r.attrs["mito_chr"], str)
    np.testing.assert_array_equal(r.objects["A_fwd"].read(), a)
    np.testing.assert_array_equal(r.objects["many"].read(), many)
    np.testing.assert_array_equal(r.objects["sparse"].read(), sparse)
    np.testing.assert_array_equal(r.objects["barcode_metadata/total"].read(), np.arange(141))
    assert r.objects["reference"].read().tobytes() == b"ACGTN" * 20 and r.objects["empty"].read().shape == (16569, 0)
    assert r.structures["paged FADB"] == 1 and r.structures["single-chunk index"] == 2
    # the same message encodings libhdf5 chose for the reference's file (datatype, filter pipeline, layout of a
    # (1000, 100) uint16 gzip-4 dataset): byte for byte, apart from the index address
    ref = H5Reader(os.path.join(REF, "counts.h5"))
    mine, theirs = r.objects["A_fwd"], ref.objects["A_fwd"]
    assert mine.layout[:-8] == theirs.layout[:-8] and mine.filters == theirs.filters and mine.dtype == theirs.dtype
    # a damaged metadata block is noticed
    blob = bytearray(open(path, "rb").read())
    for at in (9, 14):                                                                  # the size field, a message body
        blob2 = bytearray(blob)
        blob2[blob.rfind(b"OHDR") + at] ^= 1
        bad = tmp_path / "bad.h5"
        bad.write_bytes(blob2)
        with pytest.raises(ValueError):
            H5Reader(bad)


def test_dense_hdf5_writer_files(tmp_path):
    """DenseHDF5Writer on a hand-made device result: names, shapes, types and attributes as the reference's files, the
    values those of the planes (dead cells and barcodes without reads: zero columns), laid out over the whole barcode
    list although the result only has columns for three of its barcodes."""
    from mgatk2_b200 import PipelineConfig
    from mgatk2_b200.engine import CELL_QC_DTYPE, OVERFLOW_DTYPE, PileupResult
    from mgatk2_b200.readers import ReadsByBarcode
    from mgatk2_b200.writers import DenseHDF5Writer
    rng = np.random.default_rng(9)
    P, ppad = 16569, 16576
    barcodes = [f"CELL{i:04d}-1" for i in range(250)]
    columns = np.array([3, 120, 249])
    planes = np.zeros((3, 11, ppad), np.uint16)
    planes[:, :8, :P] = (rng.random((3, 8, P)) < 0.3) * rng.integers(1, 40, (3, 8, P))
    planes[1, 0, 77] = 65535
    planes[:, 8:10, :P] = rng.integers(0, 2, (3, 2, P))
    planes[:, 10, :P] = np.minimum(planes[:, :8, :P].astype(np.uint32).sum(1), 65535)
    planes[2] = 0                                                                       # a cell whose reads all fell to the filters
    qc = np.zeros(3, CELL_QC_DTYPE)
    for c in range(3):
        cov = planes[c, 10, :P].astype(np.int64)
        dep = np.sort(cov[cov > 0])
        qc[c] = (5 + c, 4, cov.sum(), len(dep), dep.max() if len(dep) else 0, dep[(len(dep) - 1) // 2] if len(dep) else 0,
                 dep[len(dep) // 2] if len(dep) else 0)
    totals = np.stack([planes[:, 2 * b, :P].astype(np.int64).sum(0) + planes[:, 2 * b + 1, :P].sum(0) for b in range(4)], 1)
    res = PileupResult(planes, qc, {}, totals, np.zeros(0, OVERFLOW_DTYPE), P, 1, columns=columns)
    cfg = PipelineConfig()
    meta = {"barcode": barcodes[::-1], "passed_filters": list(range(250)), "excluded_reason": ["0"] * 250}
    w = DenseHDF5Writer(tmp_path, cfg, barcodes, barcode_metadata=meta)
    results = w.write_result(ReadsByBarcode(barcodes, np.array([1, 0, 2]), res), cfg)
    assert [r["barcode"] for r in results] == [barcodes[120], barcodes[3]]
    w.finalize(tmp_path / "qc")
    c = H5Reader(tmp_path / "output" / "counts.h5")
    assert c.attrs == {"n_cells": 250, "n_positions": P, "mito_chr": "chrM"}
    assert c.objects["barcode"].read().tolist() == [b.encode() for b in barcodes]
    for k, name in enumerate(PLANES):
        got = c.objects[name].read()
        assert got.shape == (P, 250) and got.dtype == np.uint16
        want = np.zeros((P, 250), np.uint16)
        want[:, [3, 120]] = planes[:2, k, :P].T
        np.testing.assert_array_equal(got, want)
    m = H5Reader(tmp_path / "output" / "metadata.h5")
    assert m.attrs == {"mito_chr": "chrM", "mito_length": P}
    np.testing.assert_array_equal(m.objects["coverage"].read()[:, 120], planes[1, 10, :P])
    cov = planes[:2, 10, :P].astype(np.float64)
    for name, want in (("mean_depth", cov.sum(1) / (cov > 0).sum(1)), ("total_bases", cov.sum(1)),
                       ("genome_coverage", (cov > 0).sum(1) / P * 100), ("max_depth", cov.max(1)),
                       ("median_depth", [np.median(x[x > 0]) for x in cov])):
        got = m.objects[name].read()
        full = np.zeros(250, got.dtype)
        full[[3, 120]] = np.asarray(want).astype(got.dtype)
        np.testing.assert_array_equal(got, full, err_msg=name)
    assert m.objects["reference"].read().tobytes() == res.reference_alleles().astype("S1").tobytes()
    np.testing.assert_array_equal(m.objects["barcode_metadata/passed_filters"].read(), np.arange(250)[::-1])
    assert m.objects["barcode_metadata/excluded_reason"].read().dtype == np.dtype("S1")
    assert (tmp_path / "qc" / "cell_stats.csv").read_text().splitlines()[1].startswith(barcodes[120])
