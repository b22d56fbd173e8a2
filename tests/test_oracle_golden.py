"""The C oracle against golden vectors produced by the reference's own Python code
(tests/golden/make_golden.py) and the SURVEY Appendix A known answers."""
import numpy as np
import pytest

from oracle.oracle import make_params, run_oracle
from tests.helpers import GOLDEN, assert_matches_golden, golden_ids, load_golden


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_matches_reference_golden(path, threads):
    d, batch, barcodes, params = load_golden(path)
    res = run_oracle(batch, make_params(len(barcodes), **params), n_threads=threads)
    assert_matches_golden(d, params, res.counts, res.tn5, res.coverage, res.cell_qc, res.stats)
    # reference-allele vote input: sum over live cells of fwd+rev (writers.py:220-222)
    np.testing.assert_array_equal(res.base_totals, d["exp_counts"].sum(axis=(0, 3)).astype(np.int64))


def test_appendix_a_known_answers(golden_dir):
    """Spot values from SURVEY.md Appendix A, independent of the fixtures' own expectations."""
    d, batch, bcs, p = load_golden(f"{golden_dir}/kat1_insertion_softclip.npz")
    res = run_oracle(batch, make_params(2, **p))
    got = {(int(pos), "ACGT"[b]) for pos, b in zip(*np.nonzero(res.counts[0, :, :, 0]))}
    assert got == {(103, "C"), (104, "G"), (105, "T"), (106, "A"), (107, "C"),
                   (108, "G"), (109, "T"), (110, "A"), (111, "C"), (112, "G")}
    assert res.tn5.sum() == 0 and res.cell_qc["covered"][0] == 10

    d, batch, bcs, p = load_golden(f"{golden_dir}/kat2_reverse_overhang.npz")
    res = run_oracle(batch, make_params(2, **p))
    got = {(int(pos), "ACGT"[b]) for pos, b in zip(*np.nonzero(res.counts[0, :, :, 1]))}
    assert got == {(16565, "C"), (16566, "G"), (16567, "T"), (16568, "A")} and res.tn5.sum() == 0

    d, batch, bcs, p = load_golden(f"{golden_dir}/kat4_strand_bias.npz")
    res = run_oracle(batch, make_params(2, **p))
    assert res.coverage[0, 300:310].tolist() == [5] * 10 and res.coverage[0].sum() == 50
    assert res.tn5[0, 300, 0] == 4 and res.tn5[0, 309, 1] == 1 and res.tn5.sum() == 5
    assert res.stats["filtered_reads"] == 14 and res.stats["dup_position_only"] == 9

    d, batch, bcs, p = load_golden(f"{golden_dir}/kat5_flags_gate.npz")
    res = run_oracle(batch, make_params(2, **p))
    assert res.stats["total_reads"] == 8 and res.stats["filtered_reads"] == 3
    assert res.coverage[0, 66] == 0 and res.cell_qc["covered"][0] == 19 and res.coverage[1].sum() == 0
    assert res.tn5[0, 60, 0] == 1 and res.tn5.sum() == 1
