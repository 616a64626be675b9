"""Developer tool (GPU box): stage-by-stage parity of the CUDA product against
the reference CUDA build (oracle/_ref), written to gpurun_out/parity_report.json.
Also dumps small reference outputs as golden fixtures (gpurun_out/golden/).

  python tools/gpu_parity_report.py [--quick] [--golden]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from instance_stixels_b200 import _lib as L, api, synth  # noqa: E402
from oracle import refbind  # noqa: E402
import parity  # noqa: E402


def run_case(name, mode, rows, cols, step, frames, dump_golden=None, invalid=0.0):
    pre = synth.preset(mode, rows, cols, step)
    pre["invalid_disparity"] = invalid
    cfg = api.StixelConfig(**pre)
    pairwise = mode == "pairwise"
    ref = refbind.RefStixels(cfg)
    ours = api.make_stixels(pre, max_batch=max(len(frames), 1))
    out = dict(case=name, mode=mode, rows=rows, cols=cols, step=step, frames=[])
    # init tables
    rl, rr, _ = ref.init_tables()
    for fi in frames:
        fr = synth.make_frame(fi, rows=rows, cols=cols, column_step=step)
        t0 = time.time()
        rsec, rinst, rmeta = ref.compute(pairwise, fr.disparity, fr.segmentation, fr.road)
        t_ref = time.time() - t0
        ours.SetDisparityImage(fr.disparity)
        ours.SetSegmentation(fr.segmentation)
        ours.SetRoadParameters(**fr.road)
        t0 = time.time()
        data = ours.Compute(pairwise)
        t_ours = time.time() - t0
        oinst = ours.instance_records()
        rep = dict(frame=fi, t_ref_s=t_ref, t_ours_s=t_ours)
        # stage tensors
        rep["obj_cost_lut_bits"] = parity.bit_equal(ours.read_tensor(L.T_OBJ_COST_LUT), rl.ravel())
        rep["odr_bits"] = parity.bit_equal(ours.read_tensor(L.T_OBJECT_DISPARITY_RANGE), rr)
        g = np.concatenate(ref.ground_tables())
        og = ours.read_tensor(L.T_GROUND_TABLES)
        rep["ground_tables_bits"] = parity.bit_equal(og, g)
        rj = ref.read_tensor(L.T_JOINED_DISPARITY)
        oj = ours.read_tensor(L.T_JOINED_DISPARITY)
        rep["joined_bits"] = parity.bit_equal(oj, rj)
        rlut = ref.read_tensor(L.T_OBJECT_LUT)
        olut = ours.read_tensor(L.T_OBJECT_LUT)
        rep["object_lut_bits"] = parity.bit_equal(olut, rlut)
        rep["sections"] = parity.compare_sections(data.sections, rsec)
        rep["instances"] = parity.compare_instances(oinst, rinst)
        rep["meta_equal"] = all(getattr(rmeta, n) == getattr(data, n) for n, _ in L.FrameMeta._fields_)
        out["frames"].append(rep)
        print(json.dumps(rep, default=str)[:1500], flush=True)
        if dump_golden:
            os.makedirs(dump_golden, exist_ok=True)
            np.savez_compressed(os.path.join(dump_golden, f"{name}_f{fi}.npz"),
                                sections=rsec[:, :32].copy(), lengths=parity.column_lengths(rsec),
                                instances=rinst, rows=rows, cols=cols, step=step, mode=mode, frame=fi,
                                invalid=invalid)
    ref.close()
    ours.Finish()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    gold = os.path.join(a.out, "golden") if a.golden else None
    cases = [
        ("small_unary", "unary", 256, 512, 8, [0, 1], gold, 0.0),
        ("small_pairwise", "pairwise", 256, 512, 8, [0, 1], gold, 0.0),
        ("small_pairwise_w4", "pairwise", 256, 512, 4, [2], gold, 0.0),
        ("ragged_pairwise", "pairwise", 200, 328, 8, [3], gold, 0.0),
        ("ragged_unary_noinvalid", "unary", 200, 328, 8, [4], gold, -1.0),
    ]
    if not a.quick:
        cases += [
            ("crop_pairwise", "pairwise", 784, 1792, 8, [0], None, 0.0),
            ("full_unary", "unary", 1024, 2048, 8, [0, 1], None, 0.0),
            ("full_pairwise", "pairwise", 1024, 2048, 8, [0, 1], None, 0.0),
            ("full_pairwise_w4", "pairwise", 1024, 2048, 4, [0], None, 0.0),
        ]
    report = []
    for c in cases:
        try:
            report.append(run_case(*c[:6], dump_golden=c[6], invalid=c[7]))
        except Exception as e:  # keep going: one report for all cases
            import traceback
            traceback.print_exc()
            report.append(dict(case=c[0], error=repr(e)))
        with open(os.path.join(a.out, "parity_report.json"), "w") as f:
            json.dump(report, f, indent=1, default=str)


if __name__ == "__main__":
    main()
