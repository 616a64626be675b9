"""Segmentation ingest (FlipAndPad): CPU restatement against the torch-generated golden
vectors, and the device kernel against the restatement (bit-exact integer output)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from instance_stixels_b200 import api, synth
from oracle import ingest_cpu
import make_ingest_golden


def test_restatement_reproduces_torch_golden_vectors():
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ingest_*.npz")))
    assert len(files) >= 3
    for path in files:
        z = np.load(path)
        x = make_ingest_golden.make_input(int(z["seed"]), int(z["hs"]), int(z["ws"]))
        got = ingest_cpu.flip_and_pad(x)
        assert got.dtype == np.int32 and np.array_equal(got, z["out"]), path
        assert (got[:, 19:] < 0).any() and (got[:, :, int(z["hs"]):] == 0).all()   # truncation of negatives, padding


def test_ingest_inverts_the_synthetic_generator():
    """synth.make_frame builds the tensor directly; un-flipping / un-scaling it gives a CNN-shaped input whose
    ingest is the tensor again (and twice the columns for stixel width 4)."""
    rows, cols = 128, 256
    fr = synth.make_frame(2, rows=rows, cols=cols)
    hs = rows // 8
    cnn = (fr.segmentation[:, :, :hs][:, :, ::-1].transpose(1, 2, 0) / 8.0).astype(np.float32)   # [21][Hs][Ws]
    assert np.array_equal(ingest_cpu.flip_and_pad(cnn), fr.segmentation)
    w4 = ingest_cpu.flip_and_pad(cnn, column_step=4)
    assert w4.shape[0] == 2 * fr.segmentation.shape[0] and np.array_equal(w4[1::2], fr.segmentation)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,step", [(128, 192, 8), (200, 328, 8), (784, 1792, 8), (256, 512, 4)])
def test_device_ingest_matches_restatement_and_feeds_compute(rows, cols, step):
    import torch
    hs, ws = rows // 8, cols // 8
    n = 3
    cnn = np.stack([make_ingest_golden.make_input(10 + i, hs, ws) for i in range(n)])
    cnn[:, :19] = np.minimum(cnn[:, :19], 30.0)
    cnn[:, 19:] = np.clip(np.rint(cnn[:, 19:]), -50, 50) / 8.0
    want = np.stack([ingest_cpu.flip_and_pad(cnn[i], step) for i in range(n)])
    pre = synth.preset("pairwise", rows, cols, step)
    st = api.make_stixels(pre, max_batch=4)
    d_cnn = torch.from_numpy(cnn).cuda()
    d_seg = torch.full(want.shape, -7, dtype=torch.int32, device="cuda")
    st.FlipAndPadBatchDevice(n, d_cnn.data_ptr(), hs, ws, d_seg.data_ptr())
    st.Synchronize()
    assert np.array_equal(d_seg.cpu().numpy(), want)
    # single-frame entry point == SetSegmentation with the restated tensor
    fr = synth.make_frame(1, rows=rows, cols=cols, column_step=step)
    st.SetDisparityImage(fr.disparity)
    st.SetRoadParameters(**fr.road)
    st.SetSegmentation(want[1])
    a = st.Compute(True).sections.copy()
    st.SetSegmentation(np.zeros_like(want[1]))
    st.SetSegmentationFromCNN(d_cnn[1].data_ptr(), hs, ws)
    b = st.Compute(True).sections
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    with pytest.raises(api.InvalidArgument):
        st.SetSegmentationFromCNN(d_cnn[1].data_ptr(), hs + 1, ws)
    st.Finish()
