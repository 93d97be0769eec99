"""SCTL's binary Vector/Matrix container (vector.txx:132-178): sctl_b200.io must read the files the reference writes
(tests/golden/sctl_*_f64.bin, produced by Vector::Write / Matrix::Write through oracle/_ref) and write identical bytes."""
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = np.array([1.5, -2.25, 3.0, 0.0, 1e-300, -7.0, 42.0])
MAT = np.arange(6, dtype=np.float64).reshape(2, 3) - 2.5


def test_reads_reference_files():
    from sctl_b200 import io
    v = io.read(os.path.join(HERE, "sctl_vector_f64.bin"))
    m = io.read(os.path.join(HERE, "sctl_matrix_f64.bin"))
    assert v.shape == (7,) and np.array_equal(v, VEC)
    assert m.shape == (2, 3) and np.array_equal(m, MAT)


def test_writes_identical_bytes(tmp_path):
    from sctl_b200 import io
    io.write_vector(str(tmp_path / "v.bin"), VEC)
    io.write_matrix(str(tmp_path / "m.bin"), MAT)
    assert open(tmp_path / "v.bin", "rb").read() == open(os.path.join(HERE, "sctl_vector_f64.bin"), "rb").read()
    assert open(tmp_path / "m.bin", "rb").read() == open(os.path.join(HERE, "sctl_matrix_f64.bin"), "rb").read()
    f = np.linspace(0, 1, 5, dtype=np.float32)
    io.write_vector(str(tmp_path / "f.bin"), f)
    assert np.array_equal(io.read(str(tmp_path / "f.bin"), dtype=np.float32), f)


def test_geom_reader_and_writer_match_the_reference(tmp_path):
    """QuadElemList .geom (experimental/quad_element.cpp:292-397): read the file the unmodified reference wrote
    (tests/golden/make_geom_fixture.sh), compare with the arrays it was written from and with what the reference itself
    reads back, and write byte-identical text."""
    import pytest
    from sctl_b200 import io
    order, coord = io.read_geom(os.path.join(HERE, "torus_patch.geom"))
    assert order == 4 and coord.shape == (3 * 16 * 3,)
    assert np.array_equal(coord, io.read(os.path.join(HERE, "torus_patch_coord_f64.bin")))       # 17 significant digits round-trip
    assert np.array_equal(io.geom_nodes(order, coord), io.read(os.path.join(HERE, "torus_patch_nodes_f64.bin")))
    io.write_geom(str(tmp_path / "w.geom"), order, coord)
    assert open(tmp_path / "w.geom").read() == open(os.path.join(HERE, "torus_patch.geom")).read()
    # the reader's checks (:362-383)
    bad = tmp_path / "bad.geom"
    bad.write_text("# c\n 0 0 0\n")
    with pytest.raises(ValueError, match="order"):
        io.read_geom(str(bad))
    bad.write_text(" 0 0 0 2\n 1 0 0\n 0 1 0\n")
    with pytest.raises(ValueError, match="whole number"):
        io.read_geom(str(bad))
    bad.write_text(" 0 0 0 1\n 1 0 0 2\n")
    with pytest.raises(ValueError, match="order 1"):
        io.read_geom(str(bad))
