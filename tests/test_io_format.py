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
