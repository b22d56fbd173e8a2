#!/usr/bin/env python
"""Generate golden input/output vectors by EXECUTING THE REFERENCE'S OWN PYTHON CODE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The unmodified reference modules `processing.readers.BAMReader`,
`processing.processors.process_barcode_worker` (-> `PileupGenerator.generate_pileup`,
`.filter_strand_bias`) and `file_io.writers.IncrementalTextWriter` are imported from
/root/reference/src. `pysam`, `h5py` and `matplotlib` are absent from this image, so stub modules
are injected; the fake `pysam.AlignmentFile.fetch()` yields read objects decoded from the same
structure-of-arrays batch the CUDA path and the C oracle consume (recipe: SURVEY.md §8c).
Nothing from the reference is copied: its code is only imported and run.
"""
from __future__ import annotations

import array
import gzip
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
REFERENCE_SRC = "/root/reference/src"

from mgatk2_b200.batch import ReadBatch  # noqa: E402
from mgatk2_b200.synth import make_whitelist, synth_batch  # noqa: E402

FAKE_BAMS: dict[str, tuple[ReadBatch, list[str]]] = {}


# ------------------------------------------------------------------ fake pysam
class FakeRead:
    __slots__ = ("reference_start", "mapping_quality", "query_sequence", "query_qualities", "cigartuples",
                 "template_length", "flag", "_cb")

    def __init__(self, rec: dict, barcodes: list[str]):
        self.reference_start = rec["pos"]
        self.mapping_quality = rec["mapq"]
        self.query_sequence = rec["seq"]
        self.query_qualities = array.array("B", rec["qual"])   # what pysam returns
        self.cigartuples = rec["cigar"] or None
        self.template_length = rec["tlen"]
        self.flag = rec["flag"]
        b = rec["bc_idx"]
        self._cb = barcodes[b] if b >= 0 else (None if b == -1 else "NOTINWHITELIST-1")

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


class FakeAlignmentFile:
    def __init__(self, path, mode="rb", **kw):
        self.batch, self.barcodes = FAKE_BAMS[str(path)]
        self.references = ("chr1", "chrM")

    def fetch(self, contig=None, **kw):
        for i in range(self.batch.n_records):
            yield FakeRead(self.batch.record(i), self.barcodes)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install_stubs():
    pysam = types.ModuleType("pysam")
    pysam.AlignmentFile = FakeAlignmentFile
    pysam.index = lambda *a, **k: None
    sys.modules["pysam"] = pysam
    h5py = types.ModuleType("h5py")
    h5py.File = object
    sys.modules["h5py"] = h5py
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    for sub in ("pyplot", "ticker", "colors", "patches", "gridspec"):
        m = types.ModuleType("matplotlib." + sub)
        sys.modules["matplotlib." + sub] = m
        setattr(mpl, sub, m)
    sys.path.insert(0, REFERENCE_SRC)
    import core  # noqa: F401  (must precede `processing`: circular import otherwise)


# ------------------------------------------------------------------ run the reference
def run_reference(batch: ReadBatch, barcodes: list[str], *, min_baseq=20, min_mapq=30, min_distance_from_end=5,
                  dedup_mode=0, max_strand_bias=1.0, min_reads_per_cell=1, with_text_writer=False) -> dict:
    from core.config import PipelineConfig
    from processing.processors import process_barcode_worker
    from processing.readers import BAMReader

    cfg = PipelineConfig(min_baseq=min_baseq, min_mapq=min_mapq, max_strand_bias=max_strand_bias,
                         skip_deduplication=(dedup_mode == 2), use_fragment_length_dedup=(dedup_mode == 0),
                         min_reads_per_cell=min_reads_per_cell, sequential=True)
    cfg.quality.min_distance_from_end = min_distance_from_end   # not reachable through the CLI (Q1)
    P = cfg.mito_length
    C = len(barcodes)
    bc_index = {b: i for i, b in enumerate(barcodes)}

    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "fake.bam")
        open(bam, "wb").close()
        FAKE_BAMS[bam] = (batch, barcodes)
        reader = BAMReader(bam, cfg, set(barcodes))
        reads_by_barcode, stats = reader.collect_reads_by_barcode()

        writer = None
        if with_text_writer:
            from file_io.writers import IncrementalTextWriter
            writer = IncrementalTextWriter(Path(tmp) / "out", cfg, barcodes)

        counts = np.zeros((C, P, 4, 2), np.uint32)
        tn5 = np.zeros((C, P, 2), np.uint32)
        cov = np.zeros((C, P), np.uint32)
        qc = np.zeros((C, 3), np.float64)          # total_reads, total_fragments, mean_depth
        breadth = np.zeros(C, np.float64)
        n_reads_in = np.zeros(C, np.int64)
        alive = np.zeros(C, np.uint8)
        order = []
        for bc, reads in reads_by_barcode.items():  # processors.py:63-80 process_cells_direct
            c = bc_index[bc]
            n_reads_in[c] = len(reads)
            res = process_barcode_worker((bc, reads, cfg))
            if not res:
                continue
            alive[c] = 1
            order.append(c)
            if writer:
                writer.write_cell(res)
            for pos, d in res["pileup"].items():
                cov[c, pos] = d["depth"]
                tn5[c, pos, 0] = d["tn5_cuts_fwd"]
                tn5[c, pos, 1] = d["tn5_cuts_rev"]
                for bi, base in enumerate("ACGT"):
                    counts[c, pos, bi, 0] = d[f"{base}_fwd"]
                    counts[c, pos, bi, 1] = d[f"{base}_rev"]
            q = res["qc"]
            qc[c] = (q["total_reads"], q["total_fragments"], q["mean_depth"])
            breadth[c] = q["coverage_breadth"]
        out = dict(exp_counts=counts, exp_tn5=tn5, exp_coverage=cov, exp_qc=qc, exp_breadth=breadth,
                   exp_n_reads_in=n_reads_in, exp_alive=alive, exp_cell_order=np.array(order, np.int64),
                   exp_stats=np.array([stats["total_reads"], stats["filtered_reads"], stats["n_barcodes"],
                                       stats["duplicate_reads_with_length"], stats["duplicate_reads_position_only"]],
                                      np.int64))
        if writer:
            writer.finalize(Path(tmp) / "out" / "qc")
            outdir = Path(tmp) / "out" / "output"
            for name in ("A", "C", "G", "T", "coverage"):
                out[f"txt_{name}"] = np.frombuffer(gzip.open(outdir / f"output.{name}.txt.gz").read(), np.uint8)
            out["txt_depthTable"] = np.frombuffer((outdir / "output.depthTable.txt").read_bytes(), np.uint8)
            out["txt_refAllele"] = np.frombuffer((outdir / "chrM_refAllele.txt").read_bytes(), np.uint8)
            out["txt_cell_stats"] = np.frombuffer((Path(tmp) / "out" / "qc" / "cell_stats.csv").read_bytes(), np.uint8)
        del FAKE_BAMS[bam]
    return out


def save(name: str, batch: ReadBatch, barcodes: list[str], params: dict, exp: dict):
    path = REPO / "tests" / "golden" / f"{name}.npz"
    np.savez_compressed(path, **batch.to_npz_dict(), barcodes=np.array(barcodes),
                        **{"param_" + k: np.array(v) for k, v in params.items()}, **exp)
    print(f"{name}: {batch.n_records} records, {len(barcodes)} cells -> {path.stat().st_size / 1024:.0f} KiB; "
          f"stats={exp['exp_stats'].tolist()} alive={int(exp['exp_alive'].sum())} cov={int(exp['exp_coverage'].sum())}")


# ------------------------------------------------------------------ SURVEY Appendix A known-answer inputs
S20 = "ACGTACGTACGTACGTACGT"


def kat_cases():
    W = ["A-1", "B-1"]
    run = dict(min_baseq=20, min_mapq=30, min_distance_from_end=5, dedup_mode=0, max_strand_bias=1.0, min_reads_per_cell=1)
    yield "kat1_insertion_softclip", W, run, [
        dict(pos=100, flag=0x63, mapq=60, seq=S20, cigar=[(4, 2), (0, 8), (1, 2), (0, 8)], tlen=150, bc_idx=0)]
    yield "kat2_reverse_overhang", W, run, [
        dict(pos=16560, flag=0x93, mapq=60, seq="ACGTNCGTACGTACGTACGT", cigar=[(0, 20)], tlen=-150, bc_idx=0)]
    k3 = [dict(pos=200, flag=0x63, mapq=10, seq=S20, cigar=[(0, 20)], tlen=120, bc_idx=0),
          dict(pos=200, flag=0x63, mapq=60, seq="T" * 20, cigar=[(0, 20)], tlen=120, bc_idx=0),
          dict(pos=200, flag=0x63, mapq=60, seq="G" * 20, cigar=[(0, 20)], tlen=121, bc_idx=0),
          dict(pos=200, flag=0x53, mapq=60, seq="C" * 20, cigar=[(0, 20)], tlen=-120, bc_idx=0)]
    for mode, tag in ((0, "fraglen"), (1, "posonly"), (2, "none")):
        yield f"kat3_dedup_{tag}", W, dict(run, dedup_mode=mode), k3
    k4 = []
    for i in range(4):
        k4.append(dict(pos=300, flag=0x63, mapq=60, seq="A" * 10, cigar=[(0, 10)], tlen=100 + i, bc_idx=0))
    k4.append(dict(pos=300, flag=0x53, mapq=60, seq="A" * 10, cigar=[(0, 10)], tlen=-100, bc_idx=0))
    for i in range(5):
        k4.append(dict(pos=400, flag=0x63, mapq=60, seq="C" * 10, cigar=[(0, 10)], tlen=100 + i, bc_idx=0))
    k4.append(dict(pos=400, flag=0x53, mapq=60, seq="C" * 10, cigar=[(0, 10)], tlen=-100, bc_idx=0))
    for i in range(3):
        k4.append(dict(pos=500, flag=0x63, mapq=60, seq="G" * 10, cigar=[(0, 10)], tlen=100 + i, bc_idx=0))
    yield "kat4_strand_bias", W, dict(run, max_strand_bias=0.8, min_distance_from_end=0), k4
    q = [30] * 20
    q[6], q[7] = 19, 20
    k5 = [dict(pos=50, flag=0x163, mapq=60, seq=S20, cigar=[(0, 20)], tlen=90, bc_idx=0),
          dict(pos=50, flag=0x863, mapq=60, seq=S20, cigar=[(0, 20)], tlen=91, bc_idx=0),
          dict(pos=50, flag=0x67, mapq=60, seq=S20, cigar=[(0, 20)], tlen=92, bc_idx=0),
          dict(pos=50, flag=0x463, mapq=60, seq=S20, cigar=[(0, 20)], tlen=93, bc_idx=0),
          dict(pos=50, flag=0x63, mapq=60, seq=S20, cigar=[(0, 20)], tlen=94, bc_idx=-1),
          dict(pos=50, flag=0x63, mapq=60, seq=S20, cigar=[(0, 20)], tlen=95, bc_idx=-2),
          dict(pos=60, flag=0x63, mapq=60, seq=S20, qual=q, cigar=[(0, 10), (2, 3), (0, 10)], tlen=96, bc_idx=0),
          dict(pos=70, flag=0x63, mapq=60, seq=S20, cigar=[(0, 20)], tlen=97, bc_idx=1)]
    yield "kat5_flags_gate", W, dict(run, min_reads_per_cell=2), k5


def adversarial_records(rng, n, n_cells):
    """Random small records over odd CIGARs / flags / positions (property-test style inputs)."""
    recs = []
    for _ in range(n):
        L = int(rng.integers(1, 40))
        ops, remaining = [], L
        if rng.random() < 0.3:
            s = int(rng.integers(0, remaining)); ops.append((4, s)); remaining -= s
        while remaining > 0:
            r = rng.random()
            ln = int(rng.integers(1, remaining + 1))
            if r < 0.55:
                ops.append((int(rng.choice([0, 7, 8])), ln)); remaining -= ln
            elif r < 0.65:
                ops.append((1, ln)); remaining -= ln
            elif r < 0.8:
                ops.append((int(rng.choice([2, 3])), int(rng.integers(1, 30))))
            elif r < 0.9:
                ops.append((4, ln)); remaining -= ln
            else:
                ops.append((int(rng.choice([5, 6])), int(rng.integers(1, 5))))
        ops = [(o, l) for o, l in ops if l > 0]
        pos = int(rng.choice([rng.integers(0, 40), rng.integers(16500, 16569), rng.integers(0, 16569)]))
        flag = int(rng.choice([0, 16, 99, 147, 83, 163, 0x400 | 99, 0x100 | 99, 0x800 | 16, 4 | 99, 0x200 | 147]))
        recs.append(dict(pos=pos, flag=flag, mapq=int(rng.choice([0, 10, 29, 30, 31, 60, 255])),
                         seq="".join(rng.choice(list("ACGTNRY="), size=L, p=[.22, .22, .22, .22, .06, .02, .02, .02])),
                         qual=rng.choice([0, 2, 19, 20, 21, 30, 40, 93, 200], size=L).tolist(),
                         cigar=ops, tlen=int(rng.choice([0, 1, -1, 100, -100, 101, 2 ** 31 - 1, -(2 ** 31) + 1])),
                         bc_idx=int(rng.choice([-2, -1] + list(range(n_cells))))))
    recs.sort(key=lambda r: r["pos"])
    return recs


def main():
    install_stubs()
    # Appendix A known-answer tests
    for name, wl, params, recs in kat_cases():
        b = ReadBatch.from_records(recs)
        save(name, b, wl, params, run_reference(b, wl, **params))

    # synthetic, BASELINE.json config shapes at reference-feasible size
    wl30 = make_whitelist(30, seed=1)
    b = synth_batch(30, 6000, "atac50", seed=20261018)
    run = dict(min_baseq=20, min_mapq=30, min_distance_from_end=5, dedup_mode=0, max_strand_bias=1.0, min_reads_per_cell=1)
    save("synth_run_default", b, wl30, run, run_reference(b, wl30, with_text_writer=True, **run))
    tenx = dict(min_baseq=0, min_mapq=0, min_distance_from_end=5, dedup_mode=1, max_strand_bias=1.0, min_reads_per_cell=0)
    save("synth_tenx", b, wl30, tenx, run_reference(b, wl30, with_text_writer=True, **tenx))
    none = dict(run, dedup_mode=2, min_reads_per_cell=100)
    save("synth_dedup_none_gate100", b, wl30, none, run_reference(b, wl30, **none))

    wl20 = make_whitelist(20, seed=2)
    b = synth_batch(20, 4000, "stress150", seed=20261022)
    stress = dict(min_baseq=20, min_mapq=30, min_distance_from_end=10, dedup_mode=0, max_strand_bias=0.8, min_reads_per_cell=1)
    save("synth_stress150", b, wl20, stress, run_reference(b, wl20, **stress))

    b = synth_batch(12, 3000, "atac70", seed=20261019)
    wl12 = make_whitelist(12, seed=3)
    odd = dict(min_baseq=30, min_mapq=10, min_distance_from_end=0, dedup_mode=1, max_strand_bias=0.9, min_reads_per_cell=3)
    save("synth_atac70_d0", b, wl12, odd, run_reference(b, wl12, **odd))

    # deep single-position pile (long dedup runs, saturating-ish depth in few cells)
    rng = np.random.default_rng(7)
    recs = []
    for i in range(1500):
        recs.append(dict(pos=1000 + int(rng.integers(0, 3)), flag=int(rng.choice([99, 147])), mapq=60,
                         seq="".join(rng.choice(list("ACGT"), size=30)), cigar=[(0, 30)],
                         tlen=int(rng.choice([1, -1])) * int(rng.integers(100, 140)), bc_idx=int(rng.integers(0, 2))))
    recs.sort(key=lambda r: r["pos"])
    b = ReadBatch.from_records(recs)
    deep = dict(run, max_strand_bias=0.7)
    save("deep_pile", b, ["A-1", "B-1"], deep, run_reference(b, ["A-1", "B-1"], **deep))

    # adversarial CIGARs / flags / positions
    rng = np.random.default_rng(11)
    wl5 = make_whitelist(5, seed=4)
    for k, params in enumerate((run, dict(tenx, min_distance_from_end=0), dict(stress, min_distance_from_end=3, dedup_mode=2))):
        recs = adversarial_records(rng, 1200, 5)
        b = ReadBatch.from_records(recs)
        save(f"adversarial_{k}", b, wl5, params, run_reference(b, wl5, **params))


if __name__ == "__main__":
    main()
