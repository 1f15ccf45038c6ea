"""Minimal stand-in for h5py so that `import seqm` (the upstream reference, which
imports h5py at module load for its MD writers) works in containers without h5py.
Used ONLY by tools/ scripts that import the reference to generate golden vectors."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py stub: HDF5 output is not available in this environment")


File = Group = Dataset = _Unavailable
