"""Shared helpers for the parity tests: seeded mrandn-style inputs and bit comparisons."""
import numpy as np

SEED_A = 0x6A424C41  # "jBLA"
SEED_X = SEED_A + 1


def randn_f(shape, dtype=np.float64, seed=SEED_A, ld=None):
    """iid N(0,1) drawn in float64 then rounded to dtype (what `x[i] = randn()` does, src/randmat.jl:5-10),
    column-major; with ld > rows the matrix is a view into a taller NaN-filled parent (strided leading dim)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows, cols = shape
    vals = rng.standard_normal((cols, rows)).T.astype(dtype)  # Fortran-order fill: column by column
    if ld is None or ld == rows:
        return np.asfortranarray(vals)
    parent = np.full((ld, cols), np.nan, dtype=dtype, order="F")
    parent[:rows, :] = vals
    return parent[:rows, :]


def nan_f(shape, dtype=np.float64, ld=None):
    rows, cols = shape
    parent = np.full((ld or rows, cols), np.nan, dtype=dtype, order="F")
    return parent[:rows, :]


def bits_equal(a, b) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def to_dev(a):
    """numpy column-major (possibly strided) -> torch CUDA tensor with the same shape and strides (1, ld)."""
    import torch

    rows, cols = a.shape
    ld = a.strides[1] // a.itemsize if cols > 1 else max(rows, 1)
    if ld == rows or cols <= 1:
        parent = np.asfortranarray(a)
        ld = max(rows, 1)
    else:  # strided view: rebuild the (ld, cols) column-major parent, padding rows included (NaN where unknown)
        parent = np.full((ld, cols), np.nan, dtype=a.dtype, order="F")
        b = a.base
        if (isinstance(b, np.ndarray) and b.shape == (ld, cols) and b.flags.f_contiguous and b.dtype == a.dtype
                and b.ctypes.data == a.ctypes.data):
            parent[...] = b
        else:
            parent[:rows, :] = a
    store = torch.from_numpy(np.ascontiguousarray(parent.T)).cuda()  # (cols, ld) row-major == column-major (ld, cols)
    return store.t()[:rows, :]


def to_host(t):
    return np.asfortranarray(t.detach().cpu().numpy())
