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


def hdf5_datasets(res: PileupResult, n_barcodes: int | None = None) -> dict[str, np.ndarray]:
    """The arrays counts.h5 / metadata.h5 hold (writers.py:79-134): uint16 [n_positions, n_barcodes] planes
    saturated at 65535, per-cell float32 depth statistics, S1 reference. Dead cells are all-zero columns. With
    compacted result columns (`res.columns`) and `n_barcodes` given, the arrays are laid out over the whole whitelist."""
    P = res.mito_length
    alive = res.alive()
    rows = np.arange(len(res.planes)) if res.columns is None else np.asarray(res.columns)
    n = len(res.planes) if n_barcodes is None else int(n_barcodes)

    def spread(v, dtype):
        out = np.zeros(n, dtype)
        out[rows] = v
        return out
    out = {}
    for k, name in enumerate(PLANE_NAMES):
        a = np.zeros((P, n), np.uint16)
        a[:, rows[alive]] = res.planes[alive, k, :P].T
        out[name] = a
    qc = res.cell_qc
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = np.where(alive, qc["sum_depth"] / np.maximum(qc["covered"], 1), 0.0)
    out["mean_depth"] = spread(mean, np.float32)
    out["median_depth"] = spread(np.where(alive, (qc["median_lo"].astype(np.float64) + qc["median_hi"]) / 2, 0), np.float32)
    out["max_depth"] = spread(np.where(alive, np.minimum(qc["max_depth"], 65535), 0), np.uint16)
    out["genome_coverage"] = spread(np.where(alive, qc["covered"] / P * 100, 0), np.float32)
    out["total_bases"] = spread(np.where(alive, qc["sum_depth"], 0), np.float32)
    out["reference"] = res.reference_alleles().astype("S1")
    return out


class DenseHDF5Writer:
    """`output/counts.h5` + `output/metadata.h5`, the default layout of `mgatk2 run` (IncrementalHDF5Writer,
    writers.py:36-395; read by R/mgatk2_functions.R:18-61): eleven uint16 [n_positions, n_barcodes] datasets over the WHOLE
    barcode list (gzip 4, chunks (1000, 100), saturated at 65535, untouched chunks not allocated), `barcode`, the per-cell
    float32 / uint16 depth statistics, `reference` (S1) and the `barcode_metadata` group of a singlecell.csv; attributes
    n_cells / n_positions / mito_chr / mito_length. Written by `h5lite` (no libhdf5 in this image) straight from the device
    planes, chunk by chunk, on all host cores."""

    def __init__(self, output_dir: Path, config, barcodes: list[str], barcode_metadata: dict | None = None):
        self.output_dir = Path(output_dir) / "output"
        self.output_dir.mkdir(exist_ok=True, parents=True)
        self.config = config
        self.barcodes = list(barcodes)
        self.barcode_to_idx = {bc: i for i, bc in enumerate(self.barcodes)}      # last index wins (writers.py:41)
        self.barcode_metadata = barcode_metadata
        self.cell_stats: list[dict] = []
        self.cell_depths: dict[str, float] = {}
        self._result = None
        self._cells = None

    def write_result(self, reads_by_barcode, config) -> list[dict]:
        res: PileupResult = reads_by_barcode.result
        self._result = res
        P = res.mito_length
        alive = res.alive()
        results, rows, cols = [], [], []
        for bc, cell in reads_by_barcode.items():
            c = cell.index
            if not alive[c] or bc not in self.barcode_to_idx:                     # writers.py:140-141
                continue
            qc = cell_qc_row(bc, res.cell_qc[c], P)
            self.cell_stats.append(qc)
            self.cell_depths[bc] = qc["mean_depth"]
            rows.append(c)
            cols.append(self.barcode_to_idx[bc])
            results.append({"barcode": bc, "n_reads": len(cell)})
        self._cells = (np.asarray(rows, np.int64), np.asarray(cols, np.int64))
        return results

    def finalize(self, qc_dir: Path):
        import os

        from .h5lite import H5Writer
        res, (rows, cols) = self._result, self._cells if self._cells is not None else (np.zeros(0, np.int64),) * 2
        P, n = int(self.config.mito_length), len(self.barcodes)
        order = np.argsort(cols, kind="stable")
        rows, cols = rows[order], cols[order]
        threads = os.cpu_count() or 1

        def source(k):
            def chunk(i0, i1, j0, j1):
                a, b = np.searchsorted(cols, [j0, j1])
                if a == b:
                    return None                              # no written cell in this column range: the chunk is not allocated
                out = np.zeros((i1 - i0, j1 - j0), np.uint16)
                out[:, cols[a:b] - j0] = res.planes[rows[a:b], k, i0:i1].T
                return out
            return chunk
        chunks = (min(1000, P), min(100, n))
        with H5Writer(self.output_dir / "counts.h5", threads=threads) as f:       # writers.py:68-104
            f.attr("n_cells", n)
            f.attr("n_positions", P)
            f.attr("mito_chr", str(self.config.mito_chr))
            f.dataset("barcode", np.array(self.barcodes, dtype="S") if n else np.zeros(0, "S1"))
            for k, name in enumerate(PLANE_NAMES[:10]):
                f.dataset(name, None, chunks=chunks, gzip=4, shape=(P, n), dtype=np.uint16, chunk_source=source(k))
        qc = res.cell_qc if res is not None else None

        def per_cell(values, dtype):
            out = np.zeros(n, dtype)
            if len(rows):
                out[cols] = values[rows]
            return out
        with H5Writer(self.output_dir / "metadata.h5", threads=threads) as f:     # writers.py:106-134
            f.attr("mito_chr", str(self.config.mito_chr))
            f.attr("mito_length", P)
            f.dataset("coverage", None, chunks=chunks, gzip=4, shape=(P, n), dtype=np.uint16, chunk_source=source(10))
            if len(rows):
                with np.errstate(divide="ignore", invalid="ignore"):
                    mean = qc["sum_depth"] / np.maximum(qc["covered"], 1)
                f.dataset("mean_depth", per_cell(mean, np.float32))                                  # writers.py:187-197
                f.dataset("median_depth", per_cell((qc["median_lo"].astype(np.float64) + qc["median_hi"]) / 2, np.float32))
                f.dataset("max_depth", per_cell(np.minimum(qc["max_depth"], 65535), np.uint16))
                f.dataset("genome_coverage", per_cell(qc["covered"] / P * 100, np.float32))
                f.dataset("total_bases", per_cell(qc["sum_depth"].astype(np.float64), np.float32))
            else:
                for name, dt in (("mean_depth", np.float32), ("median_depth", np.float32), ("max_depth", np.uint16),
                                 ("genome_coverage", np.float32), ("total_bases", np.float32)):
                    f.dataset(name, np.zeros(n, dt))
            ref = res.reference_alleles() if res is not None else np.full(P, "N")
            f.dataset("reference", ref.astype("S1"), chunks=(P,), gzip=4)                            # writers.py:349-353
            if self.barcode_metadata is not None:                                                     # writers.py:355-385
                g = f.group("barcode_metadata")
                listed = list(self.barcode_metadata.get("barcode", []))
                at = {bc: i for i, bc in enumerate(listed)}
                pick = [at[bc] for bc in self.barcodes if bc in at]
                for col, values in self.barcode_metadata.items():
                    try:
                        picked = [values[i] for i in pick]
                        arr = np.array(picked)
                        if arr.dtype == object or arr.dtype.kind == "U" or (len(picked) and isinstance(picked[0], str)):
                            arr = np.array(picked, dtype="S")
                        if arr.dtype.kind == "b":
                            arr = arr.astype(np.int8)
                        f.dataset(col, arr, chunks=(max(len(arr), 1),), gzip=4, parent=g)
                    except Exception as e:                   # the reference logs and carries on
                        import logging
                        logging.getLogger(__name__).warning("Could not store metadata column '%s': %s", col, e)
        qc_dir = Path(qc_dir)
        qc_dir.mkdir(exist_ok=True, parents=True)
        if self.cell_stats:                                                       # formats.py:9-24
            names = ["barcode", "mean_depth", "coverage_breadth", "total_fragments", "total_reads"]
            with open(qc_dir / "cell_stats.csv", "w") as f:
                f.write(",".join(names) + "\n")
                for st in self.cell_stats:
                    f.write(",".join(str(st.get(k, "NA")) for k in names) + "\n")
