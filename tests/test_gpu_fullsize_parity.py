"""Parity at BASELINE.json's size ON THE FRAMES THE BENCH TIMES: all 64 frames of unary_b64, pairwise_b64 and
pairwise_w4_b64 (1024 x 2048) through the batched host entry point against the reference CUDA build
(oracle/_ref), frame by frame.  Bar (north star): boundaries / types / classes / instance partitions identical on
every column, costs / disparities / instance means within 1e-4 relative on every column.

Plus a seeded fuzz of the two pruning DP kernels against the exhaustive ones (full cost / argmin tables) over random
small shapes, weights and horizons."""
import os
import sys

import numpy as np
import pytest

import parity
from instance_stixels_b200 import _lib as L, api, synth
from oracle import refbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload", ["unary_b64", "pairwise_b64", "pairwise_w4_b64"])
def test_all_bench_frames_against_reference_cuda_build(workload):
    import fullsize_parity
    rep = fullsize_parity.run_workload(workload, frames=64)
    assert rep["columns"] == 64 * (2048 // rep["column_step"])
    # stixel boundaries, types and classes: identical on EVERY column of every frame (no ties to document)
    assert rep["columns_exact"] == rep["columns"], rep["differing_columns"]
    assert rep["stixels_ours"] == rep["stixels_ref"]
    # DP costs, stixel disparities, instance means: within 1e-4 relative (north-star tolerance) on every column
    assert rep["columns_close_1e4"] == rep["columns"], rep["field_bit_mismatches"]
    # instance ids up to label permutation (same candidates AND same partition), frame by frame
    assert rep["frames_same_keys"] == 64 and rep["frames_same_partition"] == 64
    # ... and in fact bit-identical: every float field of every stixel (profiles/r2_fullsize_parity.json: 2.8 million
    # stixels, no differing bit).  The kernels spell every float operation in the order of the reference's SASS.
    assert rep["columns_bitwise"] == rep["columns"], rep["field_bit_mismatches"]
    for f, m in rep["field_bit_mismatches"].items():
        assert m["stixels"] == 0, (f, m)
    if workload == "unary_b64":
        assert rep["dp_units_evaluated_frac"] < 0.5


def _fuzz_case(rng, mode):
    step = int(rng.choice([4, 8]))
    rows = int(rng.integers(33, 417))
    cols = step * int(rng.integers(4, 25))
    pre = synth.preset(mode, rows, cols, step)
    # random non-negative weights around (and far from) the tuned ones
    pre["prior_weight"] = float(10.0 ** rng.uniform(-1.0, 4.3)) if mode == "unary" else float(10.0 ** rng.uniform(-1.0, 1.0))
    pre["segmentation_weight"] = float(10.0 ** rng.uniform(-2.0, 1.5))
    pre["instance_weight"] = float(10.0 ** rng.uniform(-4.0, -1.0)) * rng.integers(0, 2)
    pre["disparity_weight"] = float(10.0 ** rng.uniform(-4.0, 0.5))
    pre["invalid_disparity"] = float(rng.choice([0.0, -1.0]))
    pre["max_dis"] = int(rng.choice([64, 128]))
    vhor = int(rng.integers(0, rows - 12))            # anywhere, including the top border
    fr = synth.make_frame(int(rng.integers(0, 1000)), rows=rows, cols=cols, column_step=step, max_dis=pre["max_dis"],
                          vhor=vhor)
    seg = fr.segmentation.copy()
    kind = int(rng.integers(0, 4))
    used = (rows + 7) // 8
    if kind == 1:      # confident CNN: many zero costs, ties everywhere
        seg[:, :19, :used] = np.where(rng.random(seg[:, :19, :used].shape) < 0.7, 0, seg[:, :19, :used])
    elif kind == 2:    # structureless
        seg[:, :19, :used] = rng.integers(0, 80, size=seg[:, :19, :used].shape)
    return pre, fr.disparity, seg, fr.road


@pytest.mark.parametrize("mode", ["unary", "pairwise"])
def test_pruning_fuzz_against_exhaustive_tables(mode, monkeypatch):
    """>= 200 random (shape, weights, horizon, input statistics) cases per mode: the pruning kernel's full
    (cost, argmin) tables and Sections are byte-identical to the exhaustive kernel's."""
    pairwise = mode == "pairwise"
    rng = np.random.default_rng(0xF0221 + (1 if pairwise else 0))
    monkeypatch.setenv("ISX_DP_WARPS", "4")      # the walk variants even for a single small frame
    envs = {"pruning": {}, "exhaustive": {"ISX_UNARY_EXHAUSTIVE": "1", "ISX_PAIRWISE_WALK": "0"}}
    pruned_units = total_units = 0
    n_cases = 200
    for case in range(n_cases):
        pre, disp, seg, road = _fuzz_case(rng, mode)
        outs = {}
        for name, env in envs.items():
            for k in ("ISX_UNARY_EXHAUSTIVE", "ISX_PAIRWISE_WALK"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            st = api.make_stixels(pre, max_batch=1)
            st.SetDisparityImage(disp)
            st.SetSegmentation(seg)
            st.SetRoadParameters(**road)
            try:
                data = st.Compute(pairwise)
                outs[name] = (data.sections.copy(), st.read_tensor(L.T_COST_TABLE).copy(),
                              st.read_tensor(L.T_INDEX_TABLE).copy())
            except api.StixelsError as e:     # >= 200 stixels in a column: both kernels must agree on that too
                outs[name] = ("error", str(e))
            if name == "pruning":
                ev, tot = st.dp_units()
                pruned_units += ev
                total_units += tot
            st.Finish()
        a, b = outs["pruning"], outs["exhaustive"]
        ctx = (case, {k: pre[k] for k in ("rows", "cols", "column_step", "prior_weight", "segmentation_weight",
                                          "instance_weight", "disparity_weight", "invalid_disparity", "max_dis")}, road)
        assert isinstance(a[0], str) == isinstance(b[0], str), ctx
        if isinstance(a[0], str):
            continue
        assert np.array_equal(a[1].view(np.int32), b[1].view(np.int32)), ctx
        assert np.array_equal(a[2], b[2]), ctx
        assert parity.same_used_sections(a[0], b[0]), ctx
    assert 0 < pruned_units < total_units      # the bounds did fire somewhere
    print(f"fuzz {mode}: {n_cases} cases, {pruned_units} of {total_units} units evaluated")
