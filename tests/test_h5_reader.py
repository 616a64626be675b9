"""apps/h5_reader.h: the `nlogprobs` loader of the run_cityscapes harness without libhdf5
(H5Segmentation.cpp:25-49).  Independent fixture: a file written by the real HDF5 library (MATLAB 7.3 format =
HDF5 with a 512-byte user block, shipped with scipy's tests) whose content scipy reads from the MAT-5 twin of the
same variable; plus files from tools/write_h5.py in the layouts h5py produces (contiguous / chunked / deflate)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
EXE = os.path.join(ROOT, "tests", "cpp", "h5_check")


@pytest.fixture(scope="module")
def h5_check():
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "apps"),
                    os.path.join(ROOT, "tests", "cpp", "h5_check.cpp"), "-lz", "-o", EXE], check=True)

    def run(path, dataset, tmp, raw=False):
        out = os.path.join(tmp, "out.bin")
        p = subprocess.run([EXE, path, dataset, out] + (["raw"] if raw else []), capture_output=True, text=True)
        return p, out
    return run


def test_file_written_by_the_hdf5_library(h5_check, tmp_path):
    scipy_io = pytest.importorskip("scipy.io")
    data_dir = os.path.join(os.path.dirname(scipy_io.__file__), "matlab", "tests", "data")
    h5 = os.path.join(data_dir, "testhdf5_7.4_GLNX86.mat")
    twin = os.path.join(data_dir, "testdouble_7.4_GLNX86.mat")
    if not (os.path.exists(h5) and os.path.exists(twin)):
        pytest.skip("scipy's MATLAB test files are not installed")
    want = scipy_io.loadmat(twin)["testdouble"]
    p, out = h5_check(h5, "testdouble", str(tmp_path), raw=True)
    assert p.returncode == 0, p.stderr
    head = p.stdout.split()
    rank = int(head[0])
    shape = [int(x) for x in head[1:1 + rank]]
    assert "class 1 elem 8" in p.stdout                     # IEEE double
    got = np.fromfile(out, dtype="<f8").reshape(shape)
    assert got.size == want.size and np.array_equal(got.ravel(), want.ravel())   # MATLAB stores column-major
    # the integer entry point refuses it like the reference's assert(H5T_INTEGER)
    p, _ = h5_check(h5, "testdouble", str(tmp_path))
    assert p.returncode == 1 and "not of an integer type" in p.stderr
    p, _ = h5_check(h5, "nlogprobs", str(tmp_path))
    assert p.returncode == 1 and "no dataset named" in p.stderr


@pytest.mark.parametrize("layout", ["contiguous", "chunked", "chunked_deflate", "chunked_shuffle_deflate", "userblock"])
@pytest.mark.parametrize("dtype", ["<i4", "<i2", "<u1", ">i4", "<i8"])
def test_files_in_the_layouts_h5py_writes(h5_check, tmp_path, layout, dtype):
    import write_h5
    rng = np.random.default_rng(len(layout) + len(dtype))
    shape = (5, 21, 37)                  # [C][channels][rows/8]-like, ragged against the chunks
    info = np.iinfo(np.dtype(dtype))
    a = rng.integers(max(info.min, -30000), min(info.max, 30000), size=shape).astype(dtype)
    kw = {}
    if layout.startswith("chunked"):
        kw = dict(chunks=(2, 8, 16), deflate="deflate" in layout, shuffle="shuffle" in layout)
    if layout == "userblock":
        kw = dict(userblock=512)
    path = str(tmp_path / "x_probs.h5")
    write_h5.write_h5(path, "nlogprobs", a, **kw)
    p, out = h5_check(path, "nlogprobs", str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert p.stdout.split() == ["3", "5", "21", "37"]
    got = np.fromfile(out, dtype=np.int32).reshape(shape)
    assert np.array_equal(got, a.astype(np.int64).astype(np.int32))


def test_not_an_hdf5_file(h5_check, tmp_path):
    path = tmp_path / "junk.h5"
    path.write_bytes(b"not hdf5" * 100)
    p, _ = h5_check(str(path), "nlogprobs", str(tmp_path))
    assert p.returncode == 1 and "not an HDF5 file" in p.stderr
