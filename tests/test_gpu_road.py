"""Road estimation on the device (csrc/road.cu through the C ABI) against the CPU restatement
(oracle/road_cpu.py, itself pinned on cv2.HoughLines) and the golden vectors: integer stages (v-disparity,
binary image, Hough accumulator) and the selected line are bit-exact; the camera properties use the same
glibc sinf/cosf/atanf as the reference's host code and are bit-exact too."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from instance_stixels_b200 import api, synth
from oracle import road_cpu
import make_road_golden

pytestmark = pytest.mark.gpu
CAM = dict(camera_center_y=512.0, baseline=0.209313, focal=2262.52)


def _check(est_ours, est):
    assert est_ours["ok"] == est["ok"]
    if est["ok"]:
        assert est_ours["vhor"] == est["horizon_point"]
        a = np.array([est_ours[k] for k in ("camera_tilt", "camera_height", "alpha_ground", "rho", "theta")], np.float32)
        b = np.array([est[k] for k in ("pitch", "camera_height", "slope", "rho", "theta")], np.float32)
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), (a, b)


def _single(re, arg):
    ok = re.Compute(arg)
    rho, theta = re.line()
    return dict(ok=ok, vhor=re.GetHorizonPoint(), camera_tilt=re.GetPitch(), camera_height=re.GetCameraHeight(),
                alpha_ground=re.GetSlope(), rho=rho, theta=theta)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "road_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_golden_vectors_and_stage_tensors(path):
    z = np.load(path)
    rows, cols, D = int(z["rows"]), int(z["cols"]), int(z["max_dis"])
    disp = make_road_golden.frame_disparity(str(z["name"]), rows, cols, int(z["frame"]))
    re = api.RoadEstimation()
    re.Initialize(CAM["camera_center_y"], CAM["baseline"], CAM["focal"], rows, cols, D)
    got = _single(re, disp)
    assert got["ok"] == bool(z["ok"]) and got["vhor"] == int(z["horizon_point"])
    f = np.array([got["camera_tilt"], got["camera_height"], got["alpha_ground"], got["rho"], got["theta"]], np.float32)
    assert np.array_equal(f.view(np.int32), z["floats"].view(np.int32))
    est, inter = road_cpu.estimate(disp, D, *CAM.values())
    assert np.array_equal(re.read_tensor(0), inter["vdisp"])
    assert np.array_equal(re.read_tensor(1), inter["binary"])
    assert np.array_equal(re.read_tensor(2), inter["accum"])
    re.Finish()
    assert not re.IsInitialized()


def test_host_device_and_batch_entry_points_agree():
    import torch
    rows, cols, n = 256, 512, 5
    disp = np.stack([make_road_golden.frame_disparity("tilted" if i == 2 else "small", rows, cols, 30 + i)
                     for i in range(n)])
    disp[4] = 0.0                                   # nothing to vote with: Compute returns false
    re = api.RoadEstimation()
    re.Initialize(CAM["camera_center_y"], CAM["baseline"], CAM["focal"], rows, cols, 128, 0.2, max_batch=8)
    d = torch.from_numpy(disp).cuda()
    batch = re.ComputeBatchDevice(n, d.data_ptr())
    for i in range(n):
        est, _ = road_cpu.estimate(disp[i], 128, *CAM.values())
        _check(batch[i], est)
        if est["ok"]:                               # the getters keep the last accepted estimate (RoadEstimation.cu:123-133)
            _check(_single(re, disp[i]), est)
            _check(_single(re, d[i].data_ptr()), est)
        else:
            assert not re.Compute(disp[i])
    assert not batch[4]["ok"] and batch[2]["theta"] != 0.0
    with pytest.raises(api.StixelsError):
        re.ComputeBatchDevice(9, d.data_ptr())      # more than max_batch
    with pytest.raises(api.InvalidArgument):
        re.Compute(disp[0][:10])                    # wrong image size
    re.Finish()


def test_estimated_road_feeds_the_stixel_path():
    """apps/run_cityscapes.cu:390-407: RoadEstimation on the disparity image Stixels already holds on the
    device, then SetRoadParameters(horizon, pitch, height, slope)."""
    rows, cols = 256, 512
    fr = synth.make_frame(4, rows=rows, cols=cols)
    st = api.make_stixels(synth.preset("pairwise", rows, cols, 8))
    st.SetDisparityImage(fr.disparity)
    st.SetSegmentation(fr.segmentation)
    re = api.RoadEstimation()
    re.Initialize(CAM["camera_center_y"], CAM["baseline"], CAM["focal"], rows, cols, 128)
    assert re.Compute(st.GetInputDisparityImageOnDevice())
    est, _ = road_cpu.estimate(fr.disparity, 128, *CAM.values())
    assert re.GetHorizonPoint() == est["horizon_point"]
    st.SetRoadParameters(re.GetHorizonPoint(), re.GetPitch(), re.GetCameraHeight(), re.GetSlope())
    data = st.Compute(True)
    assert (data.sections["type"][:, 0] >= 0).all()
    re.Finish()
    st.Finish()
