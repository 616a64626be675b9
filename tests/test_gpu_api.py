"""Error behaviour and call-sequence semantics of the drop-in boundary (needs a device because the
object allocates in Initialize like the reference)."""
import numpy as np
import pytest

from instance_stixels_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_setconfig_rejects_unset_fields_like_the_reference():
    """Stixels.cu:294-313: std::invalid_argument for every group of unset (-1) fields."""
    st = api.Stixels()
    full = synth.preset("unary", 64, 64, 8)
    for drop, msg in ((("rows",), "rows or columns"), (("max_dis",), "Maximum disparity"),
                      (("eps",), "Clustering"), (("prior_weight",), "weights"),
                      (("column_step",), "Stixel width"), (("focal",), "Camera")):
        pre = {k: v for k, v in full.items() if k not in drop}
        with pytest.raises(api.InvalidArgument, match=msg):
            st.SetConfig(api.StixelConfig(**pre))
    st.SetConfig(api.StixelConfig(**full))
    assert not st.IsInitialized()
    st.Initialize()
    assert st.IsInitialized() and st.GetRealCols() == 8 and st.GetMaxSections() == 200
    st.Finish()
    assert not st.IsInitialized()


def test_call_order_and_capacity_errors():
    pre = synth.preset("pairwise", 64, 64, 8)
    fr = synth.make_frame(0, rows=64, cols=64)
    st = api.Stixels()
    st.SetConfig(api.StixelConfig(**pre))
    with pytest.raises(api.StixelsError):          # not initialised
        st.SetDisparityImage(fr.disparity)
    st.Initialize(max_batch=2)
    with pytest.raises(api.InvalidArgument):       # wrong image size
        st.SetDisparityImage(fr.disparity[:32])
    st.SetDisparityImage(fr.disparity)
    st.SetSegmentation(fr.segmentation)
    with pytest.raises(api.InvalidArgument):       # road parameters missing
        st.Compute(True)
    st.SetRoadParameters(**fr.road)
    a = st.Compute(True)
    b = st.Compute(True)                           # inputs are not consumed: Compute is repeatable
    assert np.array_equal(a.sections.view(np.uint8), b.sections.view(np.uint8))
    assert a.vhor == 64 - fr.road["vhor"] - 1 and a.realcols == 8 and a.max_sections == 200
    m = st.GetInstanceStixels()
    assert all(isinstance(k, tuple) and len(k) == 2 for k in m)
    disp, seg, roads = synth.make_batch(3, rows=64, cols=64)
    with pytest.raises(api.StixelsError):          # batch larger than Initialize(max_batch)
        st.ComputeBatch(True, disp, seg, roads)
    st.Finish()
    st.Initialize(max_batch=3)                     # re-initialise after Finish like run_cityscapes.cu:328-343
    st.ComputeBatch(True, disp, seg, roads)
    st.Finish()


def test_fine_grained_setters_equal_setconfig():
    pre = synth.preset("pairwise", 64, 128, 8)
    fr = synth.make_frame(2, rows=64, cols=128)
    a = api.make_stixels(pre)
    b = api.Stixels()
    c = api.StixelConfig(**pre)
    b.SetDisparityParameters(64, 128, c.max_dis, c.invalid_disparity, c.sigma_disparity_object,
                             c.sigma_disparity_ground, c.sigma_sky)
    b.SetSegmentationParameters(19, 2)
    b.SetClusteringParameters(c.eps, c.min_pts, c.size_filter)
    b.SetWeightParameters(c.prior_weight, c.disparity_weight, c.segmentation_weight, c.instance_weight)
    b.SetProbabilities(c.pout, c.pout_sky, c.pground_given_nexist, c.pobject_given_nexist, c.psky_given_nexist,
                       c.pnexist_dis, c.pground, c.pobject, c.psky, c.pord, c.pgrav, c.pblg)
    b.SetModelParameters(c.column_step, False, c.epsilon, c.range_objects_z, c.width_margin)
    b.SetCameraParameters(c.focal, c.baseline, c.sigma_camera_tilt, c.sigma_camera_height, c.camera_center_x,
                          c.camera_center_y)
    b.Initialize()
    out = []
    for s in (a, b):
        s.SetDisparityImage(fr.disparity)
        s.SetSegmentation(fr.segmentation)
        s.SetRoadParameters(**fr.road)
        out.append(s.Compute(True).sections)
        s.Finish()
    assert np.array_equal(out[0].view(np.uint8), out[1].view(np.uint8))
