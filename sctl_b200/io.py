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


# ---- QuadElemList `.geom` text files (reference: include/sctl/experimental/quad_element.cpp:292-397) -----------------
# One line per node: X Y Z in scientific notation, the first node of every element also carries the element order; lines
# starting with '#' are comments.  Nodes of an element are its order x order tensor grid, element after element.

def read_geom(path, dtype=np.float64):
    """Returns (order, coord) with coord AoS [(element, node, xyz)] -- the layout QuadElemList::Init takes (:21-40) --
    applying the reference reader's checks (:340-383): a positive order on the first node, the same order (or none) on
    every other node, a whole number of elements."""
    coord, marks = [], []
    with open(path) as f:
        for line in f:
            t = line.strip()
            if not t or t[0] == "#":
                continue
            tok = t.split()
            if len(tok) < 3:
                raise ValueError("%s: a node line needs three coordinates: %r" % (path, line))
            coord.extend(float(x) for x in tok[:3])
            marks.append(int(tok[3]) if len(tok) > 3 else -1)
    if not marks or marks[0] <= 0:
        raise ValueError("%s: the first node must carry a positive element order" % path)
    order = marks[0]
    nn = order * order
    if len(marks) % nn:
        raise ValueError("%s: %d nodes is not a whole number of order-%d elements" % (path, len(marks), order))
    for e in range(0, len(marks), nn):
        if marks[e] != order or any(m not in (order, -1) for m in marks[e + 1:e + nn]):
            raise ValueError("%s: element %d does not have order %d" % (path, e // nn, order))
    return order, np.array(coord, dtype=dtype)


def write_geom(path, order, coord):
    """Writes the bytes QuadElemList<double>::Write produces (:292-338) for the same elements (coord AoS as above)."""
    coord = np.asarray(coord, dtype=np.float64).reshape(-1, 3)
    nn = order * order
    if coord.shape[0] % nn:
        raise ValueError("%d nodes is not a whole number of order-%d elements" % (coord.shape[0], order))
    precision = int(np.ceil(-np.log(np.finfo(np.float64).eps) / np.log(10.0)))  # 16
    width = precision + 8
    with open(path, "w") as f:
        f.write("#" + "X".rjust(width - 1) + "Y".rjust(width) + "Z".rjust(width) + "ElemOrder".rjust(width) + "\n")
        for i, x in enumerate(coord):
            f.write("".join(("%.*e" % (precision, v)).rjust(width) for v in x))
            if i % nn == 0:
                f.write(str(order).rjust(width))
            f.write("\n")


def geom_nodes(order, coord):
    """Node coordinates of a .geom element list as the flat AoS [N*3] array the kernels take (GetNodeCoord order)."""
    return np.ascontiguousarray(coord).reshape(-1)
