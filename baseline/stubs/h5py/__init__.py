"""Stand-in for h5py (absent from this image): lets `file_io.writers` of the reference import. Nothing on the timed
path (reader -> cell processor) opens an HDF5 file."""


class File:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py is not installed in this image (baseline/stubs)")
