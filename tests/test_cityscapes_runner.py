"""apps/cityscapes_runner: the reference's batch runner apps/run_cityscapes.cu on the drop-in
classes, with its own PNG / camera-JSON / .npy loaders.  CPU: the loaders against cv2 / json / numpy.  GPU: a small
Cityscapes-shaped dataset directory through the executable, its .stixels files against the C ABI called from
Python on the same inputs, and the closing line against the regular expression of tools/run_cityscapes.py:314-325."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from instance_stixels_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "apps", "cityscapes_runner")
CHECK = os.path.join(ROOT, "tests", "cpp", "loaders_check")


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "apps")], check=True, capture_output=True, text=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "apps"),
                    os.path.join(ROOT, "tests", "cpp", "loaders_check.cpp"), "-lz", "-o", CHECK],
                   check=True, capture_output=True, text=True)


def check(kind, path):
    p = subprocess.run([CHECK, kind, str(path)], capture_output=True, text=True)
    return p.returncode, p.stdout.split(), p.stderr


def test_png_loader_matches_cv2(tmp_path):
    cv2 = pytest.importorskip("cv2")
    build()
    rng = np.random.default_rng(0)
    smooth = (np.add.outer(np.arange(97), np.arange(211)) * 37 % 30000).astype(np.uint16)   # exercises the filters
    for name, img in [("noise16", rng.integers(0, 65536, (64, 80), dtype=np.uint16)), ("smooth16", smooth),
                      ("gray8", rng.integers(0, 256, (33, 47), dtype=np.uint8)),
                      ("flat8", np.full((20, 31), 7, np.uint8))]:
        for level in (0, 3, 9):
            path = tmp_path / f"{name}_{level}.png"
            assert cv2.imwrite(str(path), img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            rc, out, err = check("png", path)
            assert rc == 0, err
            assert [int(x) for x in out] == [img.shape[0], img.shape[1], img.dtype.itemsize * 8, int(img.sum()),
                                             int(img[0, 0]), int(img[-1, -1])]
    rgb = tmp_path / "rgb.png"
    cv2.imwrite(str(rgb), rng.integers(0, 256, (8, 8, 3), dtype=np.uint8))
    rc, _, err = check("png", rgb)
    assert rc == 1 and "grayscale" in err
    rc, _, err = check("png", tmp_path / "missing.png")
    assert rc == 1 and "Couldn't read the file" in err


def test_camera_and_npy_loaders(tmp_path):
    build()
    cam = {"extrinsic": {"baseline": 0.209313, "pitch": 0.038, "roll": 0.0, "x": 1.7, "y": 0.1, "yaw": -0.0195, "z": 1.22},
           "intrinsic": {"fx": 2262.52, "fy": 2265.3017905988554, "u0": 1096.98, "v0": 513.137}}
    path = tmp_path / "a_camera.json"
    path.write_text(json.dumps(cam, indent=4))
    rc, out, err = check("camera", path)
    assert rc == 0, err
    got = np.array([float(x) for x in out[:3]], np.float32)
    want = np.array([cam["extrinsic"]["baseline"], cam["intrinsic"]["fy"], cam["intrinsic"]["v0"]], np.float32)
    assert np.array_equal(got, want) and out[3] == "1"
    rc, out, _ = check("camera", tmp_path / "none.json")      # the reference's UEYE fallback
    assert rc == 0 and out[3] == "0" and abs(float(out[1]) - 1495.46) < 1e-3
    bad = tmp_path / "bad.json"
    bad.write_text('{"intrinsic": {"fy": 1.0}}')
    assert check("camera", bad)[0] == 1
    a = np.random.default_rng(1).integers(-50, 500, (16, 21, 32), dtype=np.int32)
    np.save(tmp_path / "x_probs.npy", a)
    rc, out, err = check("npy", tmp_path / "x_probs.npy")
    assert rc == 0, err
    assert [int(x) for x in out] == [3, 16, 21, 32, int(a.sum())]
    np.save(tmp_path / "f.npy", a.astype(np.float32))
    assert check("npy", tmp_path / "f.npy")[0] == 1


def _read_stixels(path):
    cols = []
    for line in open(path).read().splitlines():
        if line.startswith("groundplane"):
            continue
        cols.append([item.split(",") for item in line.split(";") if item])
    return cols


@pytest.mark.gpu
@pytest.mark.parametrize("pairwise", [0, 1])
def test_runner_on_a_dataset_directory(tmp_path, pairwise):
    cv2 = pytest.importorskip("cv2")
    build()
    rows, cols = 512, 1024
    for d in ("disparities", "camera", "probs", "stixels"):
        (tmp_path / d).mkdir()
    cam = {"extrinsic": {"baseline": 0.209313}, "intrinsic": {"fx": 2262.52, "fy": 2262.52, "u0": 512.0, "v0": 256.0}}
    frames = {}
    for i in range(3):
        fr = synth.make_frame(20 + i, rows=rows, cols=cols)
        base = f"town_{i:06d}_000019"
        q = np.clip(np.rint(fr.disparity * 256.0), 0, 65535).astype(np.uint16)
        cv2.imwrite(str(tmp_path / "disparities" / f"{base}_disparity.png"), q)
        (tmp_path / "camera" / f"{base}_camera.json").write_text(json.dumps(cam))
        if i == 1:   # the reference's own input format: dataset "nlogprobs" of an HDF5 file (H5Segmentation.cpp:25-49)
            import sys
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import write_h5
            write_h5.write_h5(str(tmp_path / "probs" / f"{base}_probs.h5"), "nlogprobs", fr.segmentation)
        else:
            np.save(tmp_path / "probs" / f"{base}_probs.npy", fr.segmentation)
        frames[base] = (q.astype(np.float32) / 256.0, fr.segmentation)
    pre = synth.preset("pairwise" if pairwise else "unary", rows, cols, 8)
    args = [APP, str(tmp_path), "128", repr(pre["segmentation_weight"]), repr(pre["instance_weight"]),
            repr(pre["disparity_weight"]), str(pairwise), "8", repr(pre["eps"]), str(pre["min_pts"]),
            str(pre["size_filter"])]
    p = subprocess.run(args, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    # what tools/run_cityscapes.py:314-325 does with the output
    time_line = [l for l in p.stdout.split("\n")[-3:] if l.startswith("It took an average")][0]
    m = re.search(r"([0-9]*\.[0-9]*) milliseconds", time_line)
    assert m and 0.0 < float(m.group(1)) < 1000.0 and time_line.rstrip().endswith("fps")
    assert p.stdout.count("Done. Time elapsed (s):") == 3 and p.stdout.count("New camera parameters") == 1

    # the same frames through the C ABI from Python, with exactly the fields the runner sets (everything else keeps
    # the StixelConfig defaults, e.g. pground = pobject = psky = 1/3, types.h:120-122)
    st = api.Stixels()
    st.SetConfig(api.StixelConfig(
        rows=rows, cols=cols, max_dis=128, column_step=8, invalid_disparity=0.0, n_semantic_classes=19,
        n_offset_channels=2, prior_weight=1.0 if pairwise else 1e4, segmentation_weight=pre["segmentation_weight"],
        instance_weight=pre["instance_weight"], disparity_weight=pre["disparity_weight"], eps=pre["eps"],
        min_pts=pre["min_pts"], size_filter=pre["size_filter"], baseline=0.209313, focal=2262.52,
        camera_center_y=256.0))
    st.Initialize()
    road = api.RoadEstimation()
    road.Initialize(256.0, 0.209313, 2262.52, rows, cols, 128)
    for base, (disp, seg) in frames.items():
        assert road.Compute(disp)
        st.SetDisparityImage(disp)
        st.SetSegmentation(seg)
        st.SetRoadParameters(road.GetHorizonPoint(), road.GetPitch(), road.GetCameraHeight(), road.GetSlope())
        data = st.Compute(bool(pairwise))
        inst = st.GetInstanceStixels()
        got = _read_stixels(tmp_path / "stixels" / f"{base}.stixels")
        assert len(got) == cols // 8
        for c, items in enumerate(got):
            for j, f in enumerate(items):
                s = data.sections[c, j]
                assert (int(f[0]), int(f[1]), int(f[2]), int(f[4])) == (s["type"], s["vB"], s["vT"], s["semantic_class"])
                assert np.isclose(float(f[3]), s["disparity"], rtol=1e-5) and np.isclose(float(f[5]), s["cost"], rtol=1e-5)
                assert (len(f) == 9) == ((c, j) in inst)
            assert data.sections[c, len(items)]["type"] == -1
    st.Finish()


@pytest.mark.gpu
@pytest.mark.parametrize("pairwise", [0, 1])
def test_batched_runner_writes_the_same_files(tmp_path, pairwise):
    """--batch B / --gpus G (StixelsPool over the batched C entry points): byte-identical .stixels files to the
    reference-style one-frame loop, and the same closing line format."""
    cv2 = pytest.importorskip("cv2")
    import shutil
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import write_h5
    build()
    rows, cols, n = 256, 512, 7
    for d in ("disparities", "camera", "probs", "stixels"):
        (tmp_path / d).mkdir()
    cam = {"extrinsic": {"baseline": 0.209313}, "intrinsic": {"fx": 2262.52, "fy": 2262.52, "u0": 256.0, "v0": 128.0}}
    for i in range(n):
        fr = synth.make_frame(50 + i, rows=rows, cols=cols)
        base = f"city_{i:06d}_000019"
        q = np.clip(np.rint(fr.disparity * 256.0), 0, 65535).astype(np.uint16)
        cv2.imwrite(str(tmp_path / "disparities" / f"{base}_disparity.png"), q)
        (tmp_path / "camera" / f"{base}_camera.json").write_text(json.dumps(cam))
        write_h5.write_h5(str(tmp_path / "probs" / f"{base}_probs.h5"), "nlogprobs", fr.segmentation)
    pre = synth.preset("pairwise" if pairwise else "unary", rows, cols, 8)
    args = [APP, str(tmp_path), "128", repr(pre["segmentation_weight"]), repr(pre["instance_weight"]),
            repr(pre["disparity_weight"]), str(pairwise), "8", repr(pre["eps"]), str(pre["min_pts"]),
            str(pre["size_filter"])]
    p = subprocess.run(args, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    single = {f: open(tmp_path / "stixels" / f, "rb").read() for f in sorted(os.listdir(tmp_path / "stixels"))}
    assert len(single) == n
    gpus = min(2, torch.cuda.device_count())
    for extra in (["--batch", "3"], ["--batch", "2", "--gpus", str(gpus)], ["--batch", "16"]):
        shutil.rmtree(tmp_path / "stixels")
        (tmp_path / "stixels").mkdir()
        p = subprocess.run(args + extra, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
        batched = {f: open(tmp_path / "stixels" / f, "rb").read() for f in sorted(os.listdir(tmp_path / "stixels"))}
        assert batched.keys() == single.keys()
        for f in single:
            assert batched[f] == single[f], (extra, f)
        time_line = [l for l in p.stdout.split("\n")[-3:] if l.startswith("It took an average")][0]
        assert re.search(r"([0-9]*\.[0-9]*) milliseconds", time_line) and time_line.rstrip().endswith("fps")
