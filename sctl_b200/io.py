"""SCTL's raw binary container format (reference: include/sctl/vector.txx:132-178, matrix.txx Write/Read):
a header of two uint64 (dim0, dim1 -- a Vector writes (Dim, 1)) followed by the payload in native byte order.
Used to exchange golden vectors and bench inputs with C++ SCTL programs (`Vector<Real>::Write/Read`)."""
import numpy as np


def write_vector(path, a):
    a = np.ascontiguousarray(a).reshape(-1)
    with open(path, "wb") as f:
        np.array([a.size, 1], dtype=np.uint64).tofile(f)
        a.tofile(f)


def write_matrix(path, m):
    m = np.ascontiguousarray(m)
    assert m.ndim == 2
    with open(path, "wb") as f:
        np.array(m.shape, dtype=np.uint64).tofile(f)
        m.tofile(f)


def read(path, dtype=np.float64):
    """Returns a 1-D array for a Vector file (dim1 == 1) and a 2-D row-major array otherwise."""
    with open(path, "rb") as f:
        d0, d1 = (int(x) for x in np.fromfile(f, dtype=np.uint64, count=2))
        data = np.fromfile(f, dtype=dtype, count=d0 * d1)
    if data.size != d0 * d1:
        raise ValueError("%s: header says %d x %d, payload has %d values" % (path, d0, d1, data.size))
    return data if d1 == 1 else data.reshape(d0, d1)
