"""Guards the vendored-at-build-time reference suite (see conftest.py in this directory)."""
import os

import pytest

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["test_resample.py", "test_random.py"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", FILES)
def test_reference_suite_is_present_on_the_gpu_box(name):
    path = os.path.join(_REF, name)
    assert os.path.exists(path), (
        "%s is missing: run __graft_entry__.build() where /root/reference exists so the reference's own "
        "tests travel with the snapshot" % path)
    with open(path) as f:
        src = f.read()
    assert "import soxr" in src and "def test_" in src


@pytest.mark.parametrize("name", FILES)
def test_reference_suite_copy_is_unmodified(name):
    """Where the reference tree is available (this container), the copies are byte-identical to it."""
    src = os.path.join("/root/reference/tests", name)
    dst = os.path.join(_REF, name)
    if not os.path.exists(src):
        pytest.skip("reference tree not present on this machine")
    if not os.path.exists(dst):
        pytest.skip("build() has not run yet")
    with open(src, "rb") as a, open(dst, "rb") as b:
        assert a.read() == b.read()
