"""The CPU oracle (oracle/stixels_cpu.cpp) pinned against outputs of the reference itself:
tests/golden/*.npz were produced by the unmodified reference CUDA sources (oracle/_ref, sm_100a)
on a B200 with tools/gpu_parity_report.py --golden.  No GPU needed here."""
import numpy as np
import pytest

import parity
from instance_stixels_b200 import synth
from oracle import cpubind


def _load_golden(path):
    z = np.load(path)
    mode = str(z["mode"])
    rows, cols, step, frame = int(z["rows"]), int(z["cols"]), int(z["step"]), int(z["frame"])
    pre = synth.preset(mode, rows, cols, step)
    pre["invalid_disparity"] = float(z["invalid"])
    fr = synth.make_frame(frame, rows=rows, cols=cols, column_step=step)
    ref = np.zeros((cols // step, 200), dtype=parity.SECTION_DTYPE)
    ref["type"] = -1
    ref[:, :z["sections"].shape[1]] = z["sections"]
    assert np.array_equal(parity.column_lengths(ref), z["lengths"])
    return mode, pre, fr, ref, z["instances"]


def test_oracle_matches_reference_golden_vectors(golden_files):
    """Column-exact structure on every golden frame; float fields within 1e-4 relative (the host has
    no MUFU.RCP/LG2, so bit equality with the GPU is not expected -- SURVEY.md 8c)."""
    assert len(golden_files) >= 7
    for path in golden_files:
        mode, pre, fr, ref, ref_inst = _load_golden(path)
        cfg = cpubind.default_config(**pre)
        sec, inst, _ = cpubind.compute_frame(cfg, mode == "pairwise", fr.disparity, fr.segmentation, fr.road)
        r = parity.compare_sections(sec, ref, rtol=1e-4)
        assert r["exact"] >= 0.999, (path, r)
        assert r["close"] >= 0.999, (path, r)
        assert all(v <= 1e-4 for v in r["max_rel"].values()), (path, r)
        ri = parity.compare_instances(inst, ref_inst)
        assert ri["same_keys"] and ri["same_partition"], (path, ri)


def test_oracle_is_deterministic_and_thread_count_independent():
    pre = synth.preset("pairwise", 128, 256, 8)
    fr = synth.make_frame(7, rows=128, cols=256)
    cfg = cpubind.default_config(**pre)
    a, ia, _ = cpubind.compute_frame(cfg, True, fr.disparity, fr.segmentation, fr.road, nthreads=1)
    b, ib, _ = cpubind.compute_frame(cfg, True, fr.disparity, fr.segmentation, fr.road, nthreads=4)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert np.array_equal(ia.view(np.uint8), ib.view(np.uint8))


def test_oracle_intermediates_are_consistent():
    """Stage tables: joined disparity is the flipped row mean of the valid pixels, LUT rows are
    monotone prefix sums, costs on the backtracked path equal Section.cost."""
    rows, cols = 96, 128
    pre = synth.preset("unary", rows, cols, 8)
    fr = synth.make_frame(5, rows=rows, cols=cols)
    cfg = cpubind.default_config(**pre)
    sec, _, ex = cpubind.compute_frame(cfg, False, fr.disparity, fr.segmentation, fr.road, tables=True)
    d = fr.disparity.reshape(rows, cols // 8, 8)
    valid = d != 0
    cnt = valid.sum(axis=2)
    mean = np.where(cnt > 0, (d * valid).sum(axis=2) / np.maximum(cnt, 1), 0.0)[::-1].T
    assert np.allclose(ex["joined"], mean, rtol=1e-5, atol=1e-5)
    lut = ex["object_lut"]
    assert np.all(lut[:, :, 0] == 0) and np.all(np.diff(lut, axis=2) >= 0)
    n = parity.column_lengths(sec)
    for c in range(sec.shape[0]):
        for s in sec[c, :n[c]]:
            row = np.minimum(ex["cost_table"][c, s["vT"]], np.float32(1e4))
            # far OBJECT stixels are relabelled SKY after the DP (StixelsKernels.cu:894-902)
            assert s["cost"] in ((row[2], row[1]) if s["type"] == 2 else (row[s["type"]],))
        # stixels tile the column top to bottom without gaps
        assert sec[c, 0]["vT"] == rows - 1 and sec[c, n[c] - 1]["vB"] == 0
        assert np.all(sec[c, :n[c] - 1]["vB"] == sec[c, 1:n[c]]["vT"] + 1)


def test_oracle_edge_cases():
    rows, cols = 64, 64
    cfg = cpubind.default_config(**synth.preset("pairwise", rows, cols, 8))
    fr = synth.make_frame(1, rows=rows, cols=cols)
    # all-invalid disparity: every column still gets a gap-free tiling
    sec, inst, _ = cpubind.compute_frame(cfg, True, np.zeros_like(fr.disparity), fr.segmentation, fr.road)
    n = parity.column_lengths(sec)
    assert np.all(n >= 1) and np.all(sec[np.arange(sec.shape[0]), n - 1]["vB"] == 0)
    # zero segmentation tensor (empty CNN output) is accepted
    sec, inst, _ = cpubind.compute_frame(cfg, True, fr.disparity, np.zeros_like(fr.segmentation), fr.road)
    assert np.all(parity.column_lengths(sec) >= 1) and len(inst) == 0


def test_dbscan_definition_matches_sklearn():
    """oracle/dbscan_def.h == classic DBSCAN when every point is a core candidate; with a size
    filter, filtered points never seed clusters (cuML fork semantics, UNPINNED w.r.t. cuML)."""
    sklearn = pytest.importorskip("sklearn.cluster")
    rng = np.random.default_rng(3)
    for trial in range(5):
        centres = rng.uniform(0, 400, size=(4, 2))
        xy = np.concatenate([c + rng.normal(0, 6, size=(30, 2)) for c in centres] +
                            [rng.uniform(0, 400, size=(15, 2))]).astype(np.float32)
        eps, min_pts = 12.0 + trial, 3 + trial % 2
        lab = cpubind.dbscan(xy, eps, min_pts, np.ones(len(xy), np.uint8))
        sk = sklearn.DBSCAN(eps=eps, min_samples=min_pts).fit(xy.astype(np.float64))
        core = np.zeros(len(xy), bool)
        core[sk.core_sample_indices_] = True
        # identical core partition and identical noise set; border points may legally differ
        assert np.array_equal(lab == -1, sk.labels_ == -1)
        pairs = {}
        for a, b in zip(lab[core], sk.labels_[core]):
            assert pairs.setdefault(a, b) == b
        assert len(set(pairs.values())) == len(pairs)
        # cluster ids are ranked by lowest member core index
        firsts = [np.flatnonzero((lab == k) & core)[0] for k in range(lab.max() + 1)]
        assert firsts == sorted(firsts)
    cand = np.zeros(len(xy), np.uint8)
    assert np.all(cpubind.dbscan(xy, 15.0, 3, cand) == -1)
