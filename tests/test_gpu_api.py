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


def _narrow(disp, seg, rows):
    """What a caller with 16-bit data holds: disparity as u16 * 1/256 and the unpadded int16 tensor."""
    used = (rows + 7) // 8
    d16 = np.rint(disp * 256.0).astype(np.uint16)
    s16 = np.ascontiguousarray(seg[..., :used]).astype(np.int16)
    wide_d = (d16.astype(np.float32) / np.float32(256.0)).astype(np.float32)    # apps/run_cityscapes.cu:141-147
    wide_s = np.zeros_like(seg)
    wide_s[..., :used] = s16
    return d16, s16, wide_d, wide_s


def test_narrow_host_inputs_equal_the_float_api_on_the_widened_arrays():
    import parity
    rows, cols, n = 200, 328, 5       # rows/8 below rows_power2_segmentation, ragged width
    pre = synth.preset("pairwise", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=40, rows=rows, cols=cols)
    d16, s16, wide_d, wide_s = _narrow(disp, seg, rows)
    import os
    os.environ["ISX_CHUNK"] = "2"
    try:
        st = api.make_stixels(pre, max_batch=8)
    finally:
        del os.environ["ISX_CHUNK"]
    assert s16[0].size == st.narrow_segmentation_elems()
    want_sec, want_inst, want_offs = st.ComputeBatch(True, wide_d, wide_s, roads)
    sec, inst, offs = st.ComputeBatchU16(True, d16, 1.0 / 256.0, s16, roads)
    assert all(parity.same_used_sections(want_sec[f], sec[f]) for f in range(n))
    assert np.array_equal(want_inst.view(np.uint8), inst.view(np.uint8)) and np.array_equal(want_offs, offs)
    # streaming form, float and narrow batches interleaved
    C_, S = st.GetRealCols(), st.GetMaxSections()
    outs = [np.zeros((n, C_, S), dtype=api.L.SECTION_DTYPE) for _ in range(3)]
    st.SubmitBatchU16(True, d16, 1.0 / 256.0, s16, roads, outs[0])
    st.SubmitBatch(True, wide_d, wide_s, roads, outs[1])
    r0 = st.WaitBatch()
    st.SubmitBatchU16(True, d16, 1.0 / 256.0, s16, roads, outs[2])
    r1 = st.WaitBatch()
    r2 = st.WaitBatch()
    for sec_k, inst_k, offs_k in (r0, r1, r2):
        assert all(parity.same_used_sections(want_sec[f], sec_k[f]) for f in range(n))
        assert np.array_equal(want_inst.view(np.uint8), inst_k.view(np.uint8)) and np.array_equal(want_offs, offs_k)
    st.Finish()


@pytest.mark.parametrize("budget", [None, "1"], ids=["packed", "overflow_fallback"])
def test_packed_results_equal_the_padded_arrays(budget, monkeypatch):
    """The device packs the results into pinned host memory; the padded [C][200] array the reference's callers
    index is expanded from that.  ISX_PACK_BUDGET=1 sizes the packed arrays for one stixel per column, so every
    frame overflows and is delivered from the padded device arrays instead: same results."""
    import parity
    import torch
    rows, cols, n = 256, 512, 6
    if budget:
        monkeypatch.setenv("ISX_PACK_BUDGET", budget)
    monkeypatch.setenv("ISX_CHUNK", "4")
    pre = synth.preset("pairwise", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=60, rows=rows, cols=cols)
    st = api.make_stixels(pre, max_batch=8)
    sec, inst, offs = st.ComputeBatch(True, disp, seg, roads)
    # a pinned (mapped) Section array is filled by the device itself: no packed copy, no expansion on the host
    pinned = torch.empty((n, st.GetRealCols(), 200, 32), dtype=torch.uint8).pin_memory()
    pinned.fill_(0x5A)
    sec_pin = pinned.numpy().view(api.L.SECTION_DTYPE).reshape(n, st.GetRealCols(), 200)
    sec_p, inst_p, offs_p = st.ComputeBatch(True, disp, seg, roads, sections_out=sec_pin)
    assert all(parity.same_used_sections(sec[f], sec_p[f]) for f in range(n))
    assert np.array_equal(inst.view(np.uint8), inst_p.view(np.uint8)) and np.array_equal(offs, offs_p)
    pinned.fill_(0x5A)
    st.SubmitBatch(True, disp, seg, roads, sec_pin)
    sec_p2, inst_p2, _ = st.WaitBatch()
    assert all(parity.same_used_sections(sec[f], sec_p2[f]) for f in range(n))
    # frame by frame through the reference call sequence (pageable and pinned result arrays)
    for f in range(n):
        st.SetDisparityImage(disp[f]); st.SetSegmentation(seg[f]); st.SetRoadParameters(**roads[f])
        data = st.Compute(True)
        assert parity.same_used_sections(data.sections, sec[f]), f
        data_p = st.Compute(True, sections_out=sec_pin[0])
        assert parity.same_used_sections(data_p.sections, sec[f]), f
        assert np.array_equal(st.instance_records().view(np.uint8), inst[offs[f]:offs[f + 1]].view(np.uint8))
    # device batch + fetch
    d_disp, d_seg = torch.from_numpy(disp).cuda(), torch.from_numpy(seg).cuda()    # kept alive until the fetch
    st.ComputeBatchDevice(True, n, d_disp.data_ptr(), d_seg.data_ptr(), roads)
    st.Synchronize()
    sec_d, inst_d, offs_d = st.FetchBatchResults(n)
    assert all(parity.same_used_sections(sec[f], sec_d[f]) for f in range(n))
    assert np.array_equal(inst.view(np.uint8), inst_d.view(np.uint8)) and np.array_equal(offs, offs_d)
    # the zero-copy form: packed arrays + descriptors
    st.SubmitBatch(True, disp, seg, roads, None)
    psec, counts, pinst, frames = st.WaitBatchPacked()
    lens = parity.column_lengths(sec.reshape(n * st.GetRealCols(), -1)).reshape(n, -1)
    assert np.array_equal(counts, lens)
    for f in range(n):
        d = frames[f]
        assert d["error"] == 0 and d["section_count"] == lens[f].sum() and d["instance_count"] == offs[f + 1] - offs[f]
        if budget:
            assert d["overflow"] != 0      # nothing fits one stixel per column
            continue
        assert d["overflow"] == 0
        mine = psec[d["section_offset"]:d["section_offset"] + d["section_count"]]
        want = np.concatenate([sec[f, c, :lens[f, c]] for c in range(sec.shape[1])])
        assert np.array_equal(mine.view(np.uint8), want.view(np.uint8))
        assert np.array_equal(pinst[d["instance_offset"]:d["instance_offset"] + d["instance_count"]].view(np.uint8),
                              inst[offs[f]:offs[f + 1]].view(np.uint8))
    st.Finish()


def test_frame_pool_equals_one_context():
    """isx_pool_*: contiguous frame blocks over several workers (here two and three workers on GPU 0, and every GPU
    of the box), sub-batches smaller than a block, idle workers taking sub-batches from the back of the fullest block:
    same Sections and instance records, in frame order, whoever computed them."""
    import parity
    import torch
    rows, cols, n = 128, 256, 11
    pre = synth.preset("pairwise", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=80, rows=rows, cols=cols)
    roads[3] = dict(roads[3], vhor=roads[3]["vhor"] + 5)
    st = api.make_stixels(pre, max_batch=n)
    want_sec, want_inst, want_offs = st.ComputeBatch(True, disp, seg, roads)
    st.Finish()
    device_sets = [[0], [0, 0], [0, 0, 0]]
    if torch.cuda.device_count() > 1:
        device_sets.append(list(range(torch.cuda.device_count())))
    for devices in device_sets:
        pool = api.StixelsPool(api.StixelConfig(**pre), devices, max_batch=2)
        assert pool.size() == len(devices) and pool.GetRealCols() == cols // 8
        for _ in range(2):
            sec, inst, offs = pool.ComputeBatch(True, disp, seg, roads)
            assert all(parity.same_used_sections(want_sec[f], sec[f]) for f in range(n)), devices
            assert np.array_equal(want_offs, offs), devices
            assert np.array_equal(want_inst.view(np.uint8), inst.view(np.uint8)), devices
            by_worker = pool.frames_by_worker()        # blocks, less or plus what a faster worker took over
            assert len(by_worker) == len(devices) and sum(by_worker) == n and min(by_worker) >= 0, by_worker
        sec1, inst1, offs1 = pool.ComputeBatch(True, disp[:1], seg[:1], roads[:1])   # fewer frames than workers
        assert parity.same_used_sections(want_sec[0], sec1[0]) and offs1[1] == want_offs[1]
        pool.close()


def test_two_devices_in_one_process():
    """Kernel attributes and SM counts are per device: a second context on another GPU of the box."""
    import parity
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    pre = synth.preset("unary", 256, 512, 8)
    fr = synth.make_frame(3, rows=256, cols=512)
    outs = []
    for dev in (0, 1):
        st = api.make_stixels(pre, max_batch=1, device=dev)
        st.SetDisparityImage(fr.disparity); st.SetSegmentation(fr.segmentation); st.SetRoadParameters(**fr.road)
        outs.append(st.Compute(False).sections.copy())
        st.Finish()
    assert parity.same_used_sections(outs[0], outs[1])


def test_errors_belong_to_their_frame_and_do_not_stick():
    """A frame the path refuses (instance offsets beyond the exact-float range of the column sums; a column with
    >= 200 stixels is reported the same way) fails ITS batch and names the frame; the results of the other frames
    are delivered, and the next call on the same handle is clean again (the flag used to be sticky)."""
    import parity
    rows, cols, n = 256, 128, 3
    pre = synth.preset("pairwise", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=5, rows=rows, cols=cols)
    st = api.make_stixels(pre, max_batch=4)
    want, want_inst, want_offs = st.ComputeBatch(True, disp, seg, roads)
    bad = seg.copy()
    bad[1, :, 20, :rows // 8] = 400000          # x offsets of 50 km: sum of the instance means >= 2^24
    out = np.zeros_like(want)
    with pytest.raises(api.StixelsError, match="frame 1"):
        st.ComputeBatch(True, disp, bad, roads, sections_out=out)
    assert parity.same_used_sections(want[0], out[0]) and parity.same_used_sections(want[2], out[2])
    # streaming form: the failing batch between two clean ones
    outs = [np.zeros_like(want) for _ in range(3)]
    st.SubmitBatch(True, disp, seg, roads, outs[0])
    st.SubmitBatch(True, disp, bad, roads, outs[1])
    sec0, inst0, offs0 = st.WaitBatch()
    st.SubmitBatch(True, disp, seg, roads, outs[2])
    with pytest.raises(api.StixelsError, match="frame 1"):
        st.WaitBatch()
    sec2, inst2, offs2 = st.WaitBatch()
    for sec_k, inst_k, offs_k in ((sec0, inst0, offs0), (sec2, inst2, offs2)):
        assert all(parity.same_used_sections(want[f], sec_k[f]) for f in range(n))
        assert np.array_equal(want_inst.view(np.uint8), inst_k.view(np.uint8)) and np.array_equal(want_offs, offs_k)
    # single-frame call sequence on the same handle
    st.SetDisparityImage(disp[1]); st.SetSegmentation(bad[1]); st.SetRoadParameters(**roads[1])
    with pytest.raises(api.StixelsError):
        st.Compute(True)
    st.SetSegmentation(seg[1])
    assert parity.same_used_sections(st.Compute(True).sections, want[1])
    st.Finish()


def test_single_frame_compute_after_an_asynchronous_device_batch():
    """Compute() right behind ComputeBatchDevice (no Synchronize) must not overwrite the road tables
    the batch's copies still read from pinned staging."""
    import parity
    import torch
    rows, cols, n = 256, 512, 4
    pre = synth.preset("pairwise", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=7, rows=rows, cols=cols)
    other = synth.make_frame(99, rows=rows, cols=cols)
    other_road = dict(other.road, vhor=other.road["vhor"] - 17, camera_tilt=0.02, alpha_ground=other.road["alpha_ground"] * 1.1)
    st = api.make_stixels(pre, max_batch=4)
    want, want_inst, _ = st.ComputeBatch(True, disp, seg, roads)
    d_disp, d_seg = torch.from_numpy(disp).cuda(), torch.from_numpy(seg).cuda()
    for _ in range(3):
        st.ComputeBatchDevice(True, n, d_disp.data_ptr(), d_seg.data_ptr(), roads)
        st.Synchronize()
        got, got_inst, _ = st.FetchBatchResults(n)
        assert all(parity.same_used_sections(want[f], got[f]) for f in range(n))
        st.ComputeBatchDevice(True, n, d_disp.data_ptr(), d_seg.data_ptr(), roads)
        st.SetDisparityImage(other.disparity); st.SetSegmentation(other.segmentation); st.SetRoadParameters(**other_road)
        single = st.Compute(True)
        assert single.vhor == rows - other_road["vhor"] - 1
    st.Finish()


@pytest.mark.parametrize("mode", ["unary", "pairwise"])
def test_launch_sizes_do_not_change_results(mode, monkeypatch):
    """How a batch is cut into launches is invisible in the results: the default chunks (64 frames per pairwise
    launch, 32 per unary launch), the quarter pieces at the head of an idle pipeline and at the tail of a synchronous
    call, three batches in flight -- all against launches of 5 frames (ISX_CHUNK below 32: no pieces)."""
    import parity
    rows, cols, n = 96, 128, 80
    pairwise = mode == "pairwise"
    pre = synth.preset(mode, rows, cols, 8)
    disp, seg, roads = synth.make_batch(16, start=3, rows=rows, cols=cols)
    disp, seg, roads = np.concatenate([disp] * 5), np.concatenate([seg] * 5), roads * 5
    monkeypatch.setenv("ISX_CHUNK", "5")
    ref = api.make_stixels(pre, max_batch=n)
    monkeypatch.delenv("ISX_CHUNK")
    want_sec, want_inst, want_offs = ref.ComputeBatch(pairwise, disp, seg, roads)
    assert ref.chunk_frames() == 5
    ref.Finish()
    st = api.make_stixels(pre, max_batch=n)
    sec, inst, offs = st.ComputeBatch(pairwise, disp, seg, roads)          # head and tail pieces around whole chunks
    assert st.chunk_frames() == (64 if pairwise else 32)
    assert all(parity.same_used_sections(want_sec[f], sec[f]) for f in range(n))
    assert np.array_equal(want_inst.view(np.uint8), inst.view(np.uint8)) and np.array_equal(want_offs, offs)
    C_, S = st.GetRealCols(), st.GetMaxSections()
    outs = [np.zeros((n, C_, S), dtype=api.L.SECTION_DTYPE) for _ in range(3)]
    got = []
    for i in range(5):                                                     # idle head, then a filled pipeline
        st.SubmitBatch(pairwise, disp, seg, roads, outs[i % 3])
        if i >= 2:
            s_, i_, o_ = st.WaitBatch()
            got.append((s_.copy(), i_, o_))
    for _ in range(2):
        s_, i_, o_ = st.WaitBatch()
        got.append((s_.copy(), i_, o_))
    for s_, i_, o_ in got:
        assert all(parity.same_used_sections(want_sec[f], s_[f]) for f in range(n))
        assert np.array_equal(want_inst.view(np.uint8), i_.view(np.uint8)) and np.array_equal(want_offs, o_)
    st.Finish()


def test_reserve_in_flight():
    """isx_reserve_in_flight: one to three batches, before or between submits; a fourth batch in flight is refused
    whether or not the sets were reserved."""
    import parity
    rows, cols, n = 64, 128, 3
    pre = synth.preset("pairwise", rows, cols, 8)
    st = api.make_stixels(pre, max_batch=4)
    for bad in (0, 4):
        with pytest.raises(api.InvalidArgument):
            st.ReserveInFlight(bad)
    st.ReserveInFlight(3)
    disp, seg, roads = synth.make_batch(n, start=11, rows=rows, cols=cols)
    want, _, _ = st.ComputeBatch(True, disp, seg, roads)
    outs = [np.zeros((n, st.GetRealCols(), st.GetMaxSections()), dtype=api.L.SECTION_DTYPE) for _ in range(3)]
    for o in outs:
        st.SubmitBatch(True, disp, seg, roads, o)
    st.ReserveInFlight(2)            # nothing left to allocate: a no-op while batches are in flight
    with pytest.raises(api.StixelsError):
        st.SubmitBatch(True, disp, seg, roads, outs[0])
    for o in outs:
        sec, _, _ = st.WaitBatch()
        assert all(parity.same_used_sections(want[f], sec[f]) for f in range(n))
    st.Finish()


def test_frame_pool_work_is_taken_over_from_a_slow_worker():
    """A worker whose own block is empty (more workers than frames in its share) or used up takes sub-batches from the
    others: with one frame per sub-batch and many more frames than workers every frame is still computed exactly once
    and lands at its place."""
    import parity
    rows, cols, n = 96, 128, 23
    pre = synth.preset("unary", rows, cols, 8)
    disp, seg, roads = synth.make_batch(n, start=5, rows=rows, cols=cols)
    st = api.make_stixels(pre, max_batch=n)
    want_sec, want_inst, want_offs = st.ComputeBatch(False, disp, seg, roads)
    st.Finish()
    pool = api.StixelsPool(api.StixelConfig(**pre), [0, 0, 0, 0], max_batch=1)
    for _ in range(3):
        sec, inst, offs = pool.ComputeBatch(False, disp, seg, roads)
        assert all(parity.same_used_sections(want_sec[f], sec[f]) for f in range(n))
        assert np.array_equal(want_offs, offs) and np.array_equal(want_inst.view(np.uint8), inst.view(np.uint8))
        assert sum(pool.frames_by_worker()) == n
    pool.close()
