"""Full-size parity on the frames the benchmark times (GPU box): every frame of a bench workload at
1024 x 2048 through the batched host entry point against the reference CUDA build (oracle/_ref), frame by frame.

  python tools/fullsize_parity.py [--frames 64] [--out gpurun_out/r2_fullsize_parity.json]

tests/test_gpu_fullsize_parity.py asserts on the same report: boundaries / types / classes identical on every
column, floats within 1e-4 relative on every column, instance partitions identical; the report also says, field by
field, how many stixels are not bit-identical and why.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from instance_stixels_b200 import api, synth  # noqa: E402
from oracle import refbind  # noqa: E402
import parity  # noqa: E402

ROWS, COLS = 1024, 2048
# the bench's workloads (bench.py WORKLOADS): mode, column_step
WORKLOADS = {"unary_b64": ("unary", 8), "pairwise_b64": ("pairwise", 8), "pairwise_w4_b64": ("pairwise", 4)}
FLOAT_FIELDS = ("disparity", "cost", "instance_meanx", "instance_meany")


def field_mismatches(ours: np.ndarray, ref: np.ndarray) -> dict:
    """Per float field: stixels of structurally identical columns whose bits differ, and the largest relative
    difference among them."""
    n_o, n_r = parity.column_lengths(ours), parity.column_lengths(ref)
    out = {f: dict(stixels=0, max_rel=0.0, max_ulp=0) for f in FLOAT_FIELDS}
    total = 0
    for c in range(ours.shape[0]):
        if n_o[c] != n_r[c]:
            continue
        a, b = ours[c, :n_o[c]], ref[c, :n_r[c]]
        total += len(a)
        for f in FLOAT_FIELDS:
            x, y = a[f], b[f]
            xi, yi = x.view(np.int32).astype(np.int64), y.view(np.int32).astype(np.int64)
            bad = xi != yi
            if bad.any():
                out[f]["stixels"] += int(bad.sum())
                with np.errstate(invalid="ignore", divide="ignore"):
                    rel = np.abs(x[bad] - y[bad]) / np.maximum(np.abs(y[bad]), 1e-6)
                rel = rel[np.isfinite(rel)]
                if rel.size:
                    out[f]["max_rel"] = max(out[f]["max_rel"], float(rel.max()))
                out[f]["max_ulp"] = max(out[f]["max_ulp"], int(np.abs(xi[bad] - yi[bad]).max()))
    out["stixels_compared"] = total
    return out


def run_workload(name: str, frames: int = 64, start: int = 0) -> dict:
    mode, step = WORKLOADS[name]
    pairwise = mode == "pairwise"
    pre = synth.preset(mode, ROWS, COLS, step)
    disp, seg, roads = synth.make_batch(frames, start=start, rows=ROWS, cols=COLS, column_step=step)
    st = api.make_stixels(pre, max_batch=frames)
    sec, inst, offs = st.ComputeBatch(pairwise, disp, seg, roads)
    ev, tot = st.dp_units()
    st.Finish()
    ref = refbind.RefStixels(api.StixelConfig(**pre))
    rep = dict(workload=name, mode=mode, column_step=step, rows=ROWS, cols=COLS, frames=frames,
               first_frame=start, dp_units_evaluated_frac=ev / max(tot, 1),
               columns=0, columns_exact=0, columns_close_1e4=0, columns_bitwise=0, stixels_ours=0, stixels_ref=0,
               instance_stixels=0, frames_same_partition=0, frames_same_keys=0,
               field_bit_mismatches={f: dict(stixels=0, max_rel=0.0, max_ulp=0) for f in FLOAT_FIELDS},
               stixels_compared=0, differing_columns=[])
    for f in range(frames):
        rsec, rinst, _ = ref.compute(pairwise, disp[f], seg[f], roads[f])
        r = parity.compare_sections(sec[f], rsec, rtol=1e-4)
        C_ = r["columns"]
        rep["columns"] += C_
        rep["columns_exact"] += int(round(r["exact"] * C_))
        rep["columns_close_1e4"] += int(round(r["close"] * C_))
        rep["columns_bitwise"] += int(round(r["bitwise"] * C_))
        rep["stixels_ours"] += r["stixels_ours"]
        rep["stixels_ref"] += r["stixels_ref"]
        for bad in r["first_bad"]:
            if len(rep["differing_columns"]) < 32:
                rep["differing_columns"].append(dict(frame=start + f, detail=[str(x) for x in bad]))
        ri = parity.compare_instances(inst[offs[f]:offs[f + 1]], rinst)
        rep["instance_stixels"] += ri["n_ref"]
        rep["frames_same_partition"] += int(ri["same_partition"])
        rep["frames_same_keys"] += int(ri["same_keys"])
        fm = field_mismatches(sec[f], rsec)
        rep["stixels_compared"] += fm["stixels_compared"]
        for k in FLOAT_FIELDS:
            rep["field_bit_mismatches"][k]["stixels"] += fm[k]["stixels"]
            rep["field_bit_mismatches"][k]["max_rel"] = max(rep["field_bit_mismatches"][k]["max_rel"], fm[k]["max_rel"])
            rep["field_bit_mismatches"][k]["max_ulp"] = max(rep["field_bit_mismatches"][k]["max_ulp"], fm[k]["max_ulp"])
    ref.close()
    rep["column_exact_frac"] = rep["columns_exact"] / rep["columns"]
    rep["column_close_frac"] = rep["columns_close_1e4"] / rep["columns"]
    rep["column_bitwise_frac"] = rep["columns_bitwise"] / rep["columns"]
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_fullsize_parity.json"))
    ap.add_argument("--workloads", default=",".join(WORKLOADS))
    args = ap.parse_args()
    reports = []
    for name in args.workloads.split(","):
        rep = run_workload(name, args.frames)
        reports.append(rep)
        print(json.dumps({k: rep[k] for k in ("workload", "frames", "column_exact_frac", "column_close_frac",
                                              "column_bitwise_frac", "frames_same_partition",
                                              "field_bit_mismatches", "stixels_compared")}), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(tool="tools/fullsize_parity.py", checker="oracle/_ref (reference CUDA build, sm_100a)",
                       tolerance="1e-4 relative on disparity / cost / instance means", reports=reports), f, indent=1)


if __name__ == "__main__":
    main()
