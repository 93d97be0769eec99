"""GPU tests of the C++ host layer: the standalone mirror test and the REFERENCE's own test programs compiled,
unmodified, against the reference headers with `-include sctl_cuda/sctl_dropin.hpp` (binaries are built by
__graft_entry__.build() into tests/_bin in the container that has /root/reference; they travel to the GPU box)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_bin")


def _run(name, timeout=300):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("%s not built (run __graft_entry__.build())" % exe)
    return subprocess.run([exe], capture_output=True, text=True, timeout=timeout)


def test_standalone_host_layer(sb):
    r = _run("test_generic_kernel_cuda")
    assert r.returncode == 0 and "PASS (0 check(s) failed)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["ref-test-generic-kernel-dropin", "ref-test-kernel_functions-dropin"])
def test_reference_test_programs_pass_through_the_dropin(sb, name):
    r = _run(name)
    assert r.returncode == 0 and "PASS (0 check(s) failed)" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_reference_particle_fmm_through_the_dropin(sb):
    """ParticleFMM<double,3>::test (fmm-wrapper.txx:36-94): 5000 x 5000 Stokes double-layer through EvalDirect and the
    stored runtime-digits trampoline; both arms now run on the GPU, so the printed error must be ~0."""
    r = _run("ref-test-fmm-dropin")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"Maximum relative error:\s*([0-9.eE+-]+)", r.stdout)
    assert m and float(m.group(1)) < 1e-12, r.stdout[-2000:]
