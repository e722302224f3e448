"""Stand-in for h5py (test tooling only; the reference's Davidson only subclasses h5py.File for
out-of-core mode, which the golden-vector generation never triggers)."""


class File:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("h5py shim: out-of-core mode is not available")
