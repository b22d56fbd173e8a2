"""Dense-plane text writer: the original mgatk text layout the reference's `IncrementalTextWriter` produces
(src/file_io/writers.py:409-510, src/file_io/formats.py:9-24), written straight from the device planes
instead of per-cell dict-of-dicts. Same file names, row syntax, row order (cells in first-seen order,
positions ascending, 1-based), gzip level (the files are sequences of gzip members, written by the native
writer csrc/textio.cpp on all host cores); HDF5 (`counts.h5` / `metadata.h5`) needs h5py, which this image
does not have — `hdf5_datasets()` returns the exact arrays those files hold (writers.py:60-134,205-229)."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from .engine import PLANE_NAMES, PileupResult
from .processors import cell_qc_row
from .textio import write_plane_file


class DenseTextWriter:
    def __init__(self, output_dir: Path, config, barcodes: list[str], compresslevel: int | None = None):
        # gzip level: the reference's 9 (writers.py:471-486) unless asked otherwise; the inflated bytes do not depend
        # on it; level 7 writes three times as fast and 2 % more bytes, level 4 fifteen times as fast and 17 % more
        if compresslevel is None:
            compresslevel = int(os.environ.get("MGATK_TXT_GZIP_LEVEL", "9"))
        self.compresslevel = int(compresslevel)
        self.output_dir = Path(output_dir) / "output"
        self.output_dir.mkdir(exist_ok=True, parents=True)
        self.config = config
        self.barcodes = list(barcodes)
        self.cell_stats: list[dict] = []
        self.cell_depths: dict[str, float] = {}
        self._result: PileupResult | None = None

    def write_result(self, reads_by_barcode, config) -> list[dict]:
        """Consumes the whole device result at once; returns what process_cells_progressive returns when a
        writer is given (processors.py:75)."""
        res: PileupResult = reads_by_barcode.result
        self._result = res
        P = res.mito_length
        alive = res.alive()
        results, cells, names = [], [], []
        for bc, cell in reads_by_barcode.items():
            c = cell.index
            if not alive[c]:
                continue
            qc = cell_qc_row(bc, res.cell_qc[c], P)
            self.cell_stats.append(qc)
            self.cell_depths[bc] = qc["mean_depth"]               # writers.py:437-438
            cells.append(c)
            names.append(bc)
            results.append({"barcode": bc, "n_reads": len(cell)})
        # rows and gzip members on all host cores, straight from the uint16 planes (csrc/textio.cpp); writers.py:471-486
        for name, plane_a, plane_b in (("A", 0, 1), ("C", 2, 3), ("G", 4, 5), ("T", 6, 7), ("coverage", 10, -1)):
            write_plane_file(self.output_dir / f"output.{name}.txt.gz", res.planes, P, res.overflow, plane_a, plane_b,
                             cells, names, level=self.compresslevel)
        return results

    def finalize(self, qc_dir: Path):
        res = self._result
        with open(self.output_dir / "output.depthTable.txt", "w") as f:           # writers.py:488-491
            for cell, depth in sorted(self.cell_depths.items()):
                f.write(f"{cell}\t{depth:.2f}\n")
        ref = res.reference_alleles() if res is not None else np.full(self.config.mito_length, "N")
        with open(self.output_dir / f"{self.config.mito_chr}_refAllele.txt", "w") as f:   # writers.py:502-506
            f.write("pos\tref\n")
            f.write("".join(f"{p}\t{b}\n" for p, b in enumerate(ref.tolist(), start=1)))
        qc_dir = Path(qc_dir)
        qc_dir.mkdir(exist_ok=True, parents=True)
        if self.cell_stats:                                                       # formats.py:9-24
            cols = ["barcode", "mean_depth", "coverage_breadth", "total_fragments", "total_reads"]
            with open(qc_dir / "cell_stats.csv", "w") as f:
                f.write(",".join(cols) + "\n")
                for st in self.cell_stats:
                    f.write(",".join(str(st.get(k, "NA")) for k in cols) + "\n")


def hdf5_datasets(res: PileupResult) -> dict[str, np.ndarray]:
    """The arrays counts.h5 / metadata.h5 hold (writers.py:79-134): uint16 [n_positions, n_barcodes] planes
    saturated at 65535, per-cell float32 depth statistics, S1 reference. Dead cells are all-zero columns."""
    P = res.mito_length
    alive = res.alive()
    out = {}
    for k, name in enumerate(PLANE_NAMES):
        out[name] = np.ascontiguousarray(res.planes[:, k, :P].T)
    qc = res.cell_qc
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = np.where(alive, qc["sum_depth"] / np.maximum(qc["covered"], 1), 0.0)
    out["mean_depth"] = mean.astype(np.float32)
    out["median_depth"] = np.where(alive, (qc["median_lo"].astype(np.float64) + qc["median_hi"]) / 2, 0).astype(np.float32)
    out["max_depth"] = np.where(alive, np.minimum(qc["max_depth"], 65535), 0).astype(np.uint16)
    out["genome_coverage"] = np.where(alive, qc["covered"] / P * 100, 0).astype(np.float32)
    out["total_bases"] = np.where(alive, qc["sum_depth"], 0).astype(np.float32)
    out["reference"] = res.reference_alleles().astype("S1")
    return out
