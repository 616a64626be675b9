"""Result images: CPU restatement of the reference's rectangle drawing against cv2.rectangle
itself, and the device rasteriser against the restatement (bit-exact)."""
import importlib.util

import numpy as np
import pytest

from instance_stixels_b200 import _lib as L, api, synth
from oracle import raster_cpu


def _toy_sections():
    sec = np.zeros((6, 200), dtype=L.SECTION_DTYPE)
    sec["type"] = -1
    rng = np.random.default_rng(3)
    inst = {}
    for c in range(6):
        cuts = [40] + sorted(rng.choice(np.arange(1, 40), size=3, replace=False).tolist(), reverse=True) + [0]
        for j in range(4):
            cls = int(rng.integers(0, 19))
            sec[c, j] = (j % 3, cuts[j + 1], cuts[j] - 1, rng.uniform(0, 100), cls, 1.0, 0.0, 0.0)
            if cls >= 11:
                inst[(c, j)] = int(rng.integers(-1, 4))
    return sec, inst


@pytest.mark.skipif(importlib.util.find_spec("cv2") is None, reason="cv2 not installed")
def test_restatement_equals_cv2_rectangles():
    sec, inst = _toy_sections()
    a = raster_cpu.draw(sec, inst, 40, 48)
    b = raster_cpu.draw(sec, inst, 40, 48, use_cv2=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[0] > 0).all() and (a[1] > 0).any() and set(np.unique(a[1])) <= {0} | set(range(11000, 19000))


@pytest.mark.gpu
@pytest.mark.parametrize("mode,rows,cols,step", [("pairwise", 256, 512, 8), ("unary", 200, 328, 8), ("pairwise", 256, 512, 4)])
def test_device_rasteriser_matches_restatement(mode, rows, cols, step):
    import torch
    n = 3
    pre = synth.preset(mode, rows, cols, step)
    disp, seg, roads = synth.make_batch(n, start=40, rows=rows, cols=cols, column_step=step)
    st = api.make_stixels(pre, max_batch=4)
    sec, inst, offs = st.ComputeBatch(mode == "pairwise", disp, seg, roads)
    lab = torch.full((n, rows, cols), 255, dtype=torch.uint8, device="cuda")
    ins = torch.full((n, rows, cols), -5, dtype=torch.int32, device="cuda")
    dsp = torch.full((n, rows, cols), -1.0, dtype=torch.float32, device="cuda")
    st.RasterizeBatchDevice(0, n, lab.data_ptr(), ins.data_ptr(), dsp.data_ptr())
    st.Synchronize()
    for f in range(n):
        r = inst[offs[f]:offs[f + 1]]
        m = {(int(x["column"]), int(x["index"])): int(x["label"]) for x in r}
        want = raster_cpu.draw(sec[f], m, rows, cols)
        assert np.array_equal(lab[f].cpu().numpy(), want[0])
        assert np.array_equal(ins[f].cpu().numpy(), want[1])
        assert np.array_equal(dsp[f].cpu().numpy().view(np.int32), want[2].view(np.int32))
    # a sub-range of the batch, label image only
    lab2 = torch.zeros((1, rows, cols), dtype=torch.uint8, device="cuda")
    st.RasterizeBatchDevice(2, 1, lab2.data_ptr())
    st.Synchronize()
    assert torch.equal(lab2[0], lab[2])
    with pytest.raises(api.InvalidArgument):
        st.RasterizeBatchDevice(2, 2, lab2.data_ptr())
    st.Finish()
