"""The C++ drop-in: include/InstanceStixels/Stixels.hpp compiles against caller code written like
apps/run_cityscapes.cu and, on a GPU, produces the same stixels as the C ABI called from Python."""
import os
import subprocess

import numpy as np
import pytest

from instance_stixels_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dropin_harness.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_harness")


def build_harness():
    libdir = os.path.join(ROOT, "instance_stixels_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include", "InstanceStixels"), SRC,
           "-L" + libdir, "-linstance_stixels_b200", "-Wl,-rpath," + libdir, "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_dropin_headers_compile_with_reference_style_caller():
    build_harness()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("pairwise", [0, 1])
def test_dropin_class_matches_c_abi(tmp_path, pairwise):
    build_harness()
    rows, cols = 128, 256
    fr = synth.make_frame(9, rows=rows, cols=cols)
    dpath, spath, out = tmp_path / "d.f32", tmp_path / "s.i32", tmp_path / "o.stixels"
    fr.disparity.tofile(dpath)
    fr.segmentation.tofile(spath)
    p = subprocess.run([EXE, str(dpath), str(spath), str(rows), str(cols), str(pairwise), str(fr.road["vhor"]),
                        repr(fr.road["alpha_ground"]), str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    st = api.make_stixels(synth.preset("pairwise" if pairwise else "unary", rows, cols, 8))
    st.SetDisparityImage(fr.disparity)
    st.SetSegmentation(fr.segmentation)
    st.SetRoadParameters(**fr.road)
    data = st.Compute(bool(pairwise))
    inst = st.GetInstanceStixels()
    st.Finish()
    lines = out.read_text().splitlines()
    assert len(lines) == cols // 8 + 1 and lines[-1].startswith("groundplane")
    n_sec = 0
    for c, line in enumerate(lines[:-1]):
        for j, item in enumerate(x for x in line.split(";") if x):
            f = item.split(",")
            s = data.sections[c, j]
            assert (int(f[0]), int(f[1]), int(f[2]), int(f[4])) == (s["type"], s["vB"], s["vT"], s["semantic_class"])
            assert np.isclose(float(f[3]), s["disparity"], rtol=1e-5) and np.isclose(float(f[5]), s["cost"], rtol=1e-5)
            assert (len(f) == 9) == ((c, j) in inst)
            if len(f) == 9:
                assert int(f[8]) == inst[(c, j)]
            n_sec += 1
        assert data.sections[c, j + 1]["type"] == -1
    assert p.stdout.split() == ["sections", str(n_sec), "instances", str(len(inst))]


@pytest.mark.gpu
def test_dropin_road_estimation_class(tmp_path):
    """The drop-in `class RoadEstimation` (include/InstanceStixels/RoadEstimation.h) in the caller's sequence
    of apps/run_cityscapes.cu:390-407, against the CPU restatement."""
    from oracle import road_cpu
    build_harness()
    rows, cols = 256, 512
    fr = synth.make_frame(12, rows=rows, cols=cols)
    dpath, spath, out = tmp_path / "d.f32", tmp_path / "s.i32", tmp_path / "o.stixels"
    fr.disparity.tofile(dpath)
    fr.segmentation.tofile(spath)
    p = subprocess.run([EXE, str(dpath), str(spath), str(rows), str(cols), "1", "-1", "0", str(out)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    road = [l for l in p.stdout.splitlines() if l.startswith("road ")][0].split()
    est, _ = road_cpu.estimate(fr.disparity, 128, 512.0, 0.209313, 2262.52)
    assert est["ok"] and int(road[1]) == est["horizon_point"]
    got = np.array([float(x) for x in road[2:5]], dtype=np.float32)
    want = np.array([est["pitch"], est["camera_height"], est["slope"]], dtype=np.float32)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
