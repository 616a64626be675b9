"""The CPU restatement of the road estimation (oracle/road_cpu.py) against OpenCV itself and against the
committed golden vectors (tools/make_road_golden.py: the reference pipeline with the real cv2.HoughLines)."""
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from instance_stixels_b200 import synth
from oracle import road_cpu
import make_road_golden

HAVE_CV2 = importlib.util.find_spec("cv2") is not None


def golden_road_files():
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "road_*.npz")))
    assert len(files) >= 5
    return files


def test_restatement_reproduces_the_golden_vectors():
    for path in golden_road_files():
        z = np.load(path)
        disp = make_road_golden.frame_disparity(str(z["name"]), int(z["rows"]), int(z["cols"]), int(z["frame"]))
        est, inter = road_cpu.estimate(disp, int(z["max_dis"]), 512.0, 0.209313, 2262.52)
        assert int(inter["vdisp"].sum()) == int(z["vdisp_sum"]) and int((inter["binary"] > 0).sum()) == int(z["binary_count"])
        lines = np.array([(l[0], l[1]) for l in inter["lines"]], dtype=np.float32).reshape(-1, 2)
        assert len(lines) == int(z["n_lines"])
        assert np.array_equal(lines[:64].view(np.int32), z["lines"].view(np.int32)), path   # rho, theta: bit-exact
        assert est["ok"] == bool(z["ok"]) and est["horizon_point"] == int(z["horizon_point"])
        got = np.array([est["pitch"], est["camera_height"], est["slope"], est["rho"], est["theta"]], dtype=np.float32)
        assert np.array_equal(got.view(np.int32), z["floats"].view(np.int32)), path


def test_first_candidate_can_be_rejected():
    """A vertical v-disparity line (theta = 0: a wall) collects the most votes but has an infinite horizon
    point; ComputeHough moves on to the next line (RoadEstimation.cu:155-167)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "road_tilted_f5.npz"))
    assert z["lines"][0, 1] == 0.0 and z["floats"][4] != 0.0


@pytest.mark.skipif(not HAVE_CV2, reason="cv2 not installed")
@pytest.mark.parametrize("seed", range(8))
def test_hough_restatement_equals_cv2(seed):
    import cv2
    rng = np.random.default_rng(100 + seed)
    rows = int(rng.integers(60, 400))
    binary = (rng.random((rows, 128)) < rng.uniform(0.005, 0.05)).astype(np.uint8) * 255
    for _ in range(int(rng.integers(1, 4))):      # a few straight segments
        a, b = rng.uniform(-0.6, 0.6), rng.uniform(0, 127)
        r = np.arange(rows)
        c = np.clip(np.rint(a * r + b).astype(int), 0, 127)
        keep = rng.random(rows) < 0.8
        binary[r[keep], c[keep]] = 255
    ref = cv2.HoughLines(binary, 1.0, np.pi / 180, road_cpu.HOUGH_ACCUM_THRESHOLD)
    ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
    lines, _ = road_cpu.hough_lines(binary)
    mine = np.array([(l[0], l[1]) for l in lines], dtype=np.float32).reshape(-1, 2)
    assert mine.shape == ref.shape and np.array_equal(mine.view(np.int32), ref.view(np.int32))


def test_empty_and_degenerate_images():
    est, inter = road_cpu.estimate(np.zeros((64, 128), np.float32), 128, 512.0, 0.2, 2262.0)
    assert not est["ok"] and inter["vdisp"].sum() == 0 and len(inter["lines"]) == 0
    # out-of-range disparities are ignored (the reference would write out of bounds)
    d = np.full((64, 128), 500.0, np.float32)
    d[:, :4] = -3.0
    assert road_cpu.vdisparity(d, 128).sum() == 0
