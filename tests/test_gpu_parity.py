"""Parity of the CUDA path (through the C-ABI) with the oracle and the reference-generated goldens."""
import os

import numpy as np
import pytest

from mgatk2_b200.batch import ReadBatch
from mgatk2_b200.synth import synth_batch
from tests.helpers import (GOLDEN, assert_matches_golden, assert_result_equals_oracle, golden_ids, load_golden)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from mgatk2_b200.engine import PileupEngine
    eng = PileupEngine(0)
    yield eng
    eng.close()


def make_params(batch, n_cells, **kw):
    from oracle.oracle import make_params as mk
    return mk(n_cells, max_read_extent=batch.max_read_extent(), **kw)


def to_lib_params(p):
    from mgatk2_b200._lib import ParamsC
    return ParamsC(*[getattr(p, f) for f, _ in p._fields_])


def run_both(engine, batch, n_cells, device_path=False, threads=8, **kw):
    from oracle.oracle import run_oracle
    p = make_params(batch, n_cells, **kw)
    ora = run_oracle(batch, p, n_threads=threads)
    lp = to_lib_params(p)
    if device_path:
        db = engine.upload(batch)
        do = engine.alloc_device_outputs(n_cells, lp.mito_length, batch.n_records, overflow_capacity=1 << 16)
        engine.run_device(db, lp, do)
        res = engine.download(do, lp)
    else:
        res = engine.run_host(batch, lp, overflow_capacity=1 << 16)
    return res, ora


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
@pytest.mark.parametrize("device_path", [False, True], ids=["host_abi", "device_abi"])
def test_golden_vectors(engine, path, device_path):
    d, batch, barcodes, params = load_golden(path)
    res, ora = run_both(engine, batch, len(barcodes), device_path=device_path, **params)
    assert_matches_golden(d, params, res.counts(), res.tn5(), res.coverage(), res.cell_qc, res.stats)
    np.testing.assert_array_equal(res.base_totals, d["exp_counts"].sum(axis=(0, 3)).astype(np.int64))
    assert_result_equals_oracle(res, ora)


CASES = {
    "run_defaults": dict(profile="atac50", n_cells=200, n=300_000, kw=dict()),
    "tenx": dict(profile="atac50", n_cells=200, n=300_000,
                 kw=dict(min_baseq=0, min_mapq=0, dedup_mode=1, min_reads_per_cell=0)),
    "dedup_none_gate": dict(profile="atac50", n_cells=150, n=100_000, kw=dict(dedup_mode=2, min_reads_per_cell=400)),
    "stress150": dict(profile="stress150", n_cells=100, n=120_000,
                      kw=dict(max_strand_bias=0.8, min_distance_from_end=10)),
    "atac70_d0": dict(profile="atac70", n_cells=64, n=80_000, kw=dict(min_distance_from_end=0, max_strand_bias=0.6)),
    "two_pass_partition": dict(profile="atac50", n_cells=5000, n=250_000, kw=dict()),
    "one_cell": dict(profile="atac50", n_cells=1, n=60_000, kw=dict(dedup_mode=1)),
    "raw_pileup_flag": dict(profile="atac70", n_cells=40, n=30_000, kw=dict(max_strand_bias=0.7, flags=1)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_synthetic_vs_oracle(engine, name):
    c = CASES[name]
    batch = synth_batch(c["n_cells"], c["n"], c["profile"], seed=20261018 + len(name))
    res, ora = run_both(engine, batch, c["n_cells"], device_path=(len(name) % 2 == 0), **c["kw"])
    assert ora.stats["filtered_reads"] > 0
    assert_result_equals_oracle(res, ora)
    assert res.launches > 0


def test_large_checksums(engine):
    """BASELINE configs[1] shape at 1/5 scale: per-cell QC rows, global counters and base totals against the
    oracle (no dense arrays), plus the self-consistency invariants the reference's goldens obey."""
    n_cells = 2000
    batch = synth_batch(n_cells, 4_000_000, "atac50", seed=20261019)
    from oracle.oracle import run_oracle
    p = make_params(batch, n_cells)
    ora = run_oracle(batch, p, n_threads=16, dense=False)
    res = engine.run_host(batch, to_lib_params(p))
    assert_result_equals_oracle(res, ora, dense=False)
    pl = res.planes[:, :, :16569].astype(np.int64)
    np.testing.assert_array_equal(pl[:, 10], pl[:, :8].sum(axis=1))            # coverage == sum of base planes
    assert not ((pl[:, 10] == 0) & ((pl[:, 8] > 0) | (pl[:, 9] > 0))).any()     # Tn5 only where covered
    np.testing.assert_array_equal(pl[:, 10].sum(axis=1), res.cell_qc["sum_depth"].astype(np.int64))
    assert res.stats["filtered_reads"] == res.stats["stage1_reads"] - res.stats["dup_with_length"]


def test_saturation_and_overflow_list(engine):
    """>65535 reads over one position: planes saturate like the HDF5 writer (writers.py:205-218), the
    overflow list keeps the exact text-format values."""
    n = 70_000
    recs_pos = np.full(n, 5000, np.int32)
    b = ReadBatch.from_records([dict(pos=5000, flag=0, mapq=60, seq="ACGTACGTAC" * 2, cigar=[(0, 20)], tlen=0, bc_idx=0)])
    batch = ReadBatch(pos=recs_pos, tlen=np.arange(n, dtype=np.int32), flag=np.zeros(n, np.uint16),
                      mapq=np.full(n, 60, np.uint8), bc_idx=np.zeros(n, np.int32), l_seq=np.full(n, 20, np.uint16),
                      n_cigar=np.ones(n, np.uint16), blob_off=np.zeros(n, np.uint32), blob=b.blob)
    res, ora = run_both(engine, batch, 2, min_distance_from_end=0)
    assert res.stats["n_overflow"] > 0 and res.planes.max() == 65535
    assert_result_equals_oracle(res, ora)
    assert res.coverage()[0, 5000] == n and res.tn5()[0, 5000, 0] == n


def test_median_straddles_the_saturation_border(engine):
    """Median depth (writers.py:190) of a cell whose middle order statistics sit on either side of 65535, and of one
    where both lie above it: the device takes them from the exact depths (overflow list), not from the saturated plane."""
    one = ReadBatch.from_records([dict(pos=0, flag=0, mapq=60, seq="ACGTACGTAC" * 2, cigar=[(0, 20)], tlen=0, bc_idx=0)])
    short = ReadBatch.from_records([dict(pos=0, flag=0, mapq=60, seq="ACGTACGTAC", cigar=[(0, 10)], tlen=0, bc_idx=0)])

    def pile(pos, n, cell, proto, L):
        return dict(pos=np.full(n, pos, np.int32), tlen=np.arange(n, dtype=np.int32), flag=np.zeros(n, np.uint16),
                    mapq=np.full(n, 60, np.uint8), bc_idx=np.full(n, cell, np.int32), l_seq=np.full(n, L, np.uint16),
                    n_cigar=np.ones(n, np.uint16), blob_off=np.full(n, proto, np.uint32))
    blob = np.concatenate([one.blob, short.blob])                      # two prototype blobs: 20 bases at unit 0, 10 bases after it
    off_short = len(one.blob) // 16
    piles = [pile(100, 10, 0, 0, 20), pile(200, 7, 0, off_short, 10), pile(5000, 70_000, 0, 0, 20), pile(5010, 66_000, 0, 0, 20),
             pile(300, 3, 1, 0, 20), pile(7000, 80_000, 1, 0, 20), pile(7020, 90_000, 1, 0, 20), pile(7040, 100_000, 1, 0, 20)]
    cat = {k: np.concatenate([p[k] for p in piles]) for k in piles[0]}
    order = np.argsort(cat["pos"], kind="stable")
    batch = ReadBatch(blob=blob, **{k: v[order] for k, v in cat.items()})
    res, ora = run_both(engine, batch, 3, min_distance_from_end=0)
    assert_result_equals_oracle(res, ora)
    assert (int(ora.cell_qc["median_lo"][0]), int(ora.cell_qc["median_hi"][0])) == (10, 66_000)     # 30 shallow, 30 deep positions
    assert (int(res.cell_qc["median_lo"][1]), int(res.cell_qc["median_hi"][1])) == (80_000, 90_000)    # both above the border
    assert res.cell_qc["median_lo"][2] == 0


@pytest.mark.parametrize("dedup_mode", [0, 1, 2])
def test_long_start_runs_in_dedup(engine, dedup_mode):
    """Thousands of reads of one cell at one start (a hot spot): a record's look-back is finished by the whole warp after
    its first 16 private steps. Duplicates at every distance, both strands, template lengths from small and large sets, runs
    that end at a cell border or at another start; both duplicate counters against the oracle in all three modes."""
    rng = np.random.default_rng(12)
    one = ReadBatch.from_records([dict(pos=0, flag=0, mapq=60, seq="ACGTACGTAC" * 2, cigar=[(0, 20)], tlen=0, bc_idx=0)])
    runs = [(100, 0, 3000, 40), (100, 1, 700, 5000), (101, 1, 40, 3), (4000, 0, 9000, 100000), (4000, 2, 17, 2), (4001, 2, 5000, 7)]
    cols = {k: [] for k in ("pos", "bc", "tlen", "flag")}
    for pos, cell, n, n_tlen in runs:
        cols["pos"].append(np.full(n, pos)); cols["bc"].append(np.full(n, cell))
        cols["tlen"].append(rng.integers(1, n_tlen + 1, n) * rng.choice([-1, 1], n))
        cols["flag"].append(rng.choice([99, 147, 83, 163, 0, 16], n))
    cat = {k: np.concatenate(v) for k, v in cols.items()}
    order = np.argsort(cat["pos"], kind="stable")
    n = len(order)
    batch = ReadBatch(pos=cat["pos"][order].astype(np.int32), tlen=cat["tlen"][order].astype(np.int32),
                      flag=cat["flag"][order].astype(np.uint16), mapq=np.full(n, 60, np.uint8), bc_idx=cat["bc"][order].astype(np.int32),
                      l_seq=np.full(n, 20, np.uint16), n_cigar=np.ones(n, np.uint16), blob_off=np.zeros(n, np.uint32), blob=one.blob)
    res, ora = run_both(engine, batch, 3, dedup_mode=dedup_mode)
    if dedup_mode != 2:                                      # (no counters without dedup, readers.py:118)
        assert ora.stats["dup_with_length"] > 1000 and ora.stats["dup_position_only"] > ora.stats["dup_with_length"]
    assert_result_equals_oracle(res, ora)


def test_record_count_with_chunks_past_the_end(engine):
    """1 818 625 records: the partition's 444 chunks of 4 352 records overshoot the batch by 26 chunks and the count is
    1 modulo 4. Those empty chunks must count nothing (a negative length rounded down to a multiple of four once made each
    of them count the last record again: 9 phantom slots in a 2 000 477-record batch, found by the streamed stress run)."""
    n = 1_818_625
    batch = synth_batch(50, n, "atac50", seed=31)
    assert batch.n_records == n
    batch.flag[-1] = 99
    batch.bc_idx[-1] = 0                                     # the last record passes stage 1
    res, ora = run_both(engine, batch, 50, device_path=True)
    assert res.stats["stage1_reads"] == ora.stats["stage1_reads"]
    assert_result_equals_oracle(res, ora)


def test_edge_cases(engine):
    from mgatk2_b200.exceptions import PileupKernelError
    empty = ReadBatch.from_records([])
    res, ora = run_both(engine, empty, 3)
    assert_result_equals_oracle(res, ora)
    assert not res.planes.any() and res.stats["total_reads"] == 0
    # everything filtered at stage 1
    recs = [dict(pos=10 + i, flag=4, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=0) for i in range(5)]
    recs += [dict(pos=20 + i, flag=0, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=-1) for i in range(5)]
    res, ora = run_both(engine, ReadBatch.from_records(recs), 3)
    assert_result_equals_oracle(res, ora)
    assert res.stats["total_reads"] == 10 and res.stats["stage1_reads"] == 0
    # unsorted input is refused, not silently mis-deduplicated
    recs = [dict(pos=50, flag=0, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=0),
            dict(pos=40, flag=0, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=0)]
    with pytest.raises(PileupKernelError) as e:
        run_both(engine, ReadBatch.from_records(recs), 1)
    assert e.value.status == 4
    # a read longer than the declared extent is refused
    b = ReadBatch.from_records([dict(pos=50, flag=0, mapq=60, seq="ACGT" * 50, cigar=[(0, 200)], bc_idx=0)])
    from oracle.oracle import make_params as mk
    with pytest.raises(PileupKernelError) as e:
        engine.run_host(b, to_lib_params(mk(1, max_read_extent=20)))
    assert e.value.status == 5
    # empty SEQ survivors are reported (the reference raises on them, readers.py:157)
    b = ReadBatch.from_records([dict(pos=50, flag=0, mapq=60, seq="", cigar=[], bc_idx=0),
                                dict(pos=60, flag=16, mapq=60, seq="ACGT" * 5, cigar=[(0, 20)], bc_idx=0)])
    res, ora = run_both(engine, b, 1)
    assert res.stats["n_empty_seq"] == 1
    assert_result_equals_oracle(res, ora)


def test_long_extent_ring_sizes(engine):
    """Deletions / skips up to ~1500 bp exercise the larger position rings."""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(3000):
        gap = int(rng.choice([0, 0, 0, 100, 400, 1500]))
        cig = [(0, 25), (3, gap), (0, 25)] if gap else [(0, 50)]
        recs.append(dict(pos=int(rng.integers(0, 16569)), flag=int(rng.choice([0, 16])), mapq=60,
                         seq="".join(rng.choice(list("ACGT"), size=50)), cigar=cig, tlen=int(rng.integers(0, 300)),
                         bc_idx=int(rng.integers(0, 4))))
    recs.sort(key=lambda r: r["pos"])
    for cap in (60, 500, 2000):
        sub = [r for r in recs if sum(l for o, l in r["cigar"]) <= cap]
        res, ora = run_both(engine, ReadBatch.from_records(sub), 4)
        assert_result_equals_oracle(res, ora)


def _random_read(rng, pos, L, bc, kind):
    """One record with a CIGAR of the requested kind whose query length matches SEQ."""
    seq = "".join(rng.choice(list("ACGTN="), p=[0.24, 0.24, 0.24, 0.24, 0.03, 0.01], size=L))
    qual = rng.choice([40, 30, 20, 19, 5, 200], p=[0.5, 0.2, 0.1, 0.1, 0.05, 0.05], size=L).tolist()
    if kind == "simple" or L < 12:
        cig = [(0, L)]
    elif kind == "clip":
        a, b = int(rng.integers(1, 5)), int(rng.integers(0, 5))
        cig = [(4, a), (0, L - a - b)] + ([(4, b)] if b else [])
    elif kind == "indel":
        a = int(rng.integers(3, L - 6))
        i, d = int(rng.integers(1, 3)), int(rng.integers(1, 9))
        cig = [(0, a), (1, i), (0, 2), (2, d), (0, L - a - i - 2)]
    else:                                # many small blocks: S M (N M)* H
        cig, left = [(4, 2)], L - 2
        while left > 6:
            cig += [(0, 3), (3, int(rng.integers(1, 4)))]
            left -= 3
        cig += [(0, left), (5, 7)]
    return dict(pos=pos, flag=int(rng.choice([0, 16, 99, 147])), mapq=int(rng.choice([60, 29])), seq=seq, qual=qual,
                cigar=cig, tlen=int(rng.integers(-400, 400)), bc_idx=bc)


def test_staging_paths(engine):
    """The corners of the pileup kernel's staging: a hot spot inside a wide tile (more reads than mask slots), mixed
    read lengths that need several staging passes per warp, reads with more than eight 32-base groups, reads whose
    blob exceeds the warp buffer (per-base path), long CIGARs, and the base-quality extremes of the int8 compare."""
    rng = np.random.default_rng(11)
    recs = []
    for _ in range(2500):                # cell 0: hot spot of 50 bp reads in 150 positions + sparse background
        recs.append(_random_read(rng, int(rng.integers(8000, 8150)), 50, 0, rng.choice(["simple", "clip", "indel"])))
    for _ in range(400):
        recs.append(_random_read(rng, int(rng.integers(0, 16569)), 50, 0, "simple"))
    for _ in range(1500):                # cell 1: 50 / 150 / 300 bp mixed
        L = int(rng.choice([50, 150, 300]))
        recs.append(_random_read(rng, int(rng.integers(0, 16400)), L, 1, rng.choice(["simple", "clip", "indel", "blocks"])))
    for _ in range(40):                  # cell 2: blobs beyond the warp buffer
        recs.append(_random_read(rng, int(rng.integers(0, 14000)), 2100, 2, rng.choice(["simple", "indel"])))
    for _ in range(300):
        recs.append(_random_read(rng, int(rng.integers(0, 16569)), 36, 2, "simple"))
    recs.sort(key=lambda r: r["pos"])
    short = [r for r in recs if len(r["seq"]) < 1000]    # extent of a few hundred: mask slots for ~100 reads per CTA
    for kw in (dict(), dict(min_baseq=-100, min_distance_from_end=0), dict(min_baseq=127, dedup_mode=2),
               dict(min_baseq=200), dict(max_strand_bias=0.75, min_distance_from_end=12, dedup_mode=1)):
        res, ora = run_both(engine, ReadBatch.from_records(short), 3, device_path=bool(kw.get("dedup_mode", 0)), **kw)
        assert_result_equals_oracle(res, ora)
    plain = [r for r in short if len(r["seq"]) == 50 and len(r["cigar"]) <= 3]   # extent 50: 512 slots, hot spot overflows
    res, ora = run_both(engine, ReadBatch.from_records(plain), 3)
    assert_result_equals_oracle(res, ora)
    res, ora = run_both(engine, ReadBatch.from_records(recs), 3)                 # with the 2100 bp reads: per-base path
    assert_result_equals_oracle(res, ora)


def test_compact_slots_complex_cigars(engine):
    """Reads of at most 56 reference positions travel as 32-byte slots whose planes are already in reference coordinates:
    soft clips, insertions (the reference's query-position quirk), deletions, hard clips and padding must land
    exactly where pileup.py:52-95 puts them, on both strands, with hot spots deeper than a stage of the pileup kernel."""
    rng = np.random.default_rng(23)
    recs = []
    for _ in range(6000):
        L = int(rng.choice([20, 36, 48]))
        kind = rng.choice(["simple", "clip", "indel"])
        pos = int(rng.integers(0, 16569)) if rng.random() < 0.6 else int(rng.integers(3000, 3100))
        recs.append(_random_read(rng, pos, L, int(rng.integers(0, 3)), kind))
    for _ in range(300):                                  # H and P operations, reads left of 0 and over the end of chrM
        L = 30
        cig = [(5, 3), (0, 10), (6, 2), (1, 2), (0, 8), (2, 4), (0, 10), (5, 1)]
        recs.append(dict(pos=int(rng.choice([-20, -3, 0, 16540, 16560, 16568])), flag=int(rng.choice([0, 16])), mapq=60,
                         seq="".join(rng.choice(list("ACGT"), size=L)), cigar=cig, tlen=int(rng.integers(0, 50)), bc_idx=1))
    recs.sort(key=lambda r: r["pos"])
    batch = ReadBatch.from_records(recs)
    assert batch.max_read_extent() <= 56
    for kw in (dict(), dict(min_distance_from_end=0, dedup_mode=2, max_strand_bias=0.8), dict(min_baseq=30, dedup_mode=1, min_distance_from_end=9)):
        res, ora = run_both(engine, batch, 3, device_path=bool(kw.get("dedup_mode", 0) == 2), **kw)
        assert_result_equals_oracle(res, ora)
    # the same reads with a declared extent beyond 56 take the wide slots: same result
    from oracle.oracle import make_params as mk, run_oracle
    p = mk(3, max_read_extent=100)
    assert_result_equals_oracle(engine.run_host(batch, to_lib_params(p), overflow_capacity=1 << 16), run_oracle(batch, p, n_threads=8))


@pytest.mark.parametrize("profile,kw", [("atac50", dict()), ("stress150", dict(max_strand_bias=0.8, min_distance_from_end=10, min_reads_per_cell=300)),
                                        ("atac70", dict(dedup_mode=1, flags=1))])
def test_streamed_batches_equal_one_batch(engine, profile, kw):
    """BASELINE configs[4] streams its input: batches cut on reference_start borders accumulate into resident planes and
    the finish pass gives, bit for bit, what one batch gives (cell gate included: a cell below min_reads_per_cell in every
    single batch can still pass it in total)."""
    from mgatk2_b200.exceptions import PileupKernelError
    from oracle.oracle import run_oracle
    n_cells = 120
    batch = synth_batch(n_cells, 150_000, profile, seed=77)
    p = make_params(batch, n_cells, **kw)
    ora = run_oracle(batch, p, n_threads=8)
    lp = to_lib_params(p)
    parts = batch.split_on_start_borders(5)
    assert len(parts) >= 4 and sum(b.n_records for b in parts) == batch.n_records
    for a, b in zip(parts[:-1], parts[1:]):
        assert a.pos[-1] < b.pos[0]
    dout = engine.alloc_device_outputs(n_cells, lp.mito_length, max(b.n_records for b in parts), overflow_capacity=1 << 16)
    res = engine.run_stream(parts, lp, dout)
    assert_result_equals_oracle(res, ora)
    # entries that pass 65535 while the batches add up (bulk depths): the cell gets carry planes and the result stays exact -
    # one pile in one batch, and two piles in two batches that only overlap to more than 65535 together
    from oracle.oracle import run_oracle
    one = ReadBatch.from_records([dict(pos=0, flag=0, mapq=60, seq="ACGTACGTAC" * 2, cigar=[(0, 20)], tlen=0, bc_idx=0)])

    def pile(pos, n, cell):
        return ReadBatch(pos=np.full(n, pos, np.int32), tlen=np.arange(n, dtype=np.int32), flag=np.zeros(n, np.uint16),
                         mapq=np.full(n, 60, np.uint8), bc_idx=np.full(n, cell, np.int32), l_seq=np.full(n, 20, np.uint16),
                         n_cigar=np.ones(n, np.uint16), blob_off=np.zeros(n, np.uint32), blob=one.blob)
    for parts2, cells in (([pile(5000, 70_000, 0)], 1), ([pile(5000, 40_000, 1), pile(5010, 45_000, 1), pile(9000, 66_000, 2)], 3)):
        whole = ReadBatch.concat(parts2)
        p2 = make_params(whole, cells, min_distance_from_end=0)
        lp2 = to_lib_params(p2)
        dout2 = engine.alloc_device_outputs(cells, lp2.mito_length, max(b.n_records for b in parts2), overflow_capacity=1 << 16)
        res2 = engine.run_stream(parts2, lp2, dout2)
        assert res2.stats["n_overflow"] > 0
        assert_result_equals_oracle(res2, run_oracle(whole, p2, n_threads=4))
    # without carry planes left the batch is refused, not silently saturated
    os.environ["MGATK_DEEP_SETS"] = "0"
    try:
        with pytest.raises(PileupKernelError) as e:
            engine.run_stream([pile(5000, 70_000, 0)], lp2, dout2)
        assert e.value.status == 9
    finally:
        del os.environ["MGATK_DEEP_SETS"]


def test_submit_wait_two_batches_in_flight(engine):
    """mgatk_pileup_host_submit / _wait: different batches (sizes, cell counts, parameters) in flight together give what the
    oracle gives for each; tickets are waited for out of order; a third submit is refused while two are in flight; an
    unsorted batch reports its status at wait and leaves the other ticket intact."""
    from mgatk2_b200.exceptions import PileupKernelError
    from oracle.oracle import run_oracle
    cases = [(synth_batch(64, 60_000, "atac50", seed=5), 64, dict()),
             (synth_batch(200, 150_000, "stress150", seed=6), 200, dict(max_strand_bias=0.8, min_reads_per_cell=300)),
             (synth_batch(16, 9_000, "atac70", seed=7), 16, dict(dedup_mode=0, min_baseq=30)),
             (synth_batch(300, 40_000, "atac50", seed=8), 300, dict(flags=1))]
    params = [make_params(b, c, **kw) for b, c, kw in cases]
    oracles = [run_oracle(b, p, n_threads=8) for (b, _, _), p in zip(cases, params)]
    lps = [to_lib_params(p) for p in params]
    outs = [engine.alloc_host_outputs(c, 16569, 1 << 16) for _, c, _ in cases]
    # pairs in flight, second ticket waited for first
    for i in (0, 2):
        t0 = engine.submit_host(cases[i][0], lps[i], outs[i])
        t1 = engine.submit_host(cases[i + 1][0], lps[i + 1], outs[i + 1])
        with pytest.raises(PileupKernelError):
            engine.submit_host(cases[i][0], lps[i], outs[i])
        r1 = engine.wait_host(t1)
        r0 = engine.wait_host(t0)
        assert_result_equals_oracle(r0, oracles[i])
        assert_result_equals_oracle(r1, oracles[i + 1])
        with pytest.raises(PileupKernelError):
            engine.wait_host(t0)                                      # not in flight any more
    # the generator form: same cell count (the two output sets are reused alternately), parameters per batch
    same = [synth_batch(64, 20_000 + 7_000 * k, "atac50", seed=20 + k) for k in range(5)]
    sp = [make_params(b, 64, min_baseq=10 + 5 * k) for k, b in enumerate(same)]
    by_id = {id(b): to_lib_params(p) for b, p in zip(same, sp)}
    pair = [engine.alloc_host_outputs(64, 16569, 1 << 16) for _ in range(2)]
    seen = []
    for j, res in engine.run_host_many(same, lambda b: by_id[id(b)], pair):
        assert_result_equals_oracle(res, run_oracle(same[j], sp[j], n_threads=8))
        seen.append(j)
    assert seen == [0, 1, 2, 3, 4]
    # an error in one ticket
    bad = cases[0][0]
    bad = ReadBatch(**{**{f: getattr(bad, f) for f in ("tlen", "flag", "mapq", "bc_idx", "l_seq", "n_cigar", "blob_off", "blob")},
                       "pos": bad.pos[::-1].copy()})
    t_bad = engine.submit_host(bad, lps[0], outs[0])
    t_ok = engine.submit_host(cases[1][0], lps[1], outs[1])
    with pytest.raises(PileupKernelError) as e:
        engine.wait_host(t_bad)
    assert e.value.status == 4
    assert_result_equals_oracle(engine.wait_host(t_ok), oracles[1])
    # and the blocking call still works afterwards
    assert_result_equals_oracle(engine.run_host(cases[2][0], lps[2], out=outs[2]), oracles[2])


def test_run_sharded_single_rank(engine):
    """multi.run_sharded (the torchrun form of one input on N GPUs) at world size 1 against the oracle: routing,
    renumbering and recombination are the ones the N-rank run uses (tools/check_multi_gpu.py is the same check under
    torchrun; the 2-rank variant below runs when the box has two GPUs)."""
    from mgatk2_b200._lib import ParamsC
    from mgatk2_b200.multi import run_sharded
    from mgatk2_b200.synth import synth_batch
    from oracle.oracle import make_params, run_oracle
    n_cells = 120
    batch = synth_batch(n_cells, 150_000, "atac50", seed=321)
    p = make_params(n_cells, max_read_extent=batch.max_read_extent(), max_strand_bias=0.9)
    lp = ParamsC(*[getattr(p, f) for f, _ in p._fields_])
    res, cols, combined = run_sharded(batch, lp, engine, 0, 1)
    ora = run_oracle(batch, p, n_threads=4)
    np.testing.assert_array_equal(res.counts(), ora.counts[cols])
    np.testing.assert_array_equal(res.coverage(), ora.coverage[cols])
    for f in ("n_reads", "n_paired", "sum_depth", "covered", "max_depth", "median_lo", "median_hi"):
        np.testing.assert_array_equal(combined["cell_qc"][f], ora.cell_qc[f])
    np.testing.assert_array_equal(combined["base_totals"], ora.base_totals)
    assert all(combined["stats"][k] == ora.stats[k] for k in ("total_reads", "filtered_reads", "dup_with_length", "dup_position_only"))


def test_run_sharded_two_ranks_when_two_gpus():
    """The same check under torchrun with two ranks (NCCL all-reduce of the base totals); skipped on a one-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29571", os.path.join(root, "tools", "check_multi_gpu.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("parity OK") == 2
