"""Batch-1 latency of the reference call sequence, per call and per kernel stage.
  python tools/latency_probe.py [--mode pairwise] [--step 8]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from instance_stixels_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="pairwise")
ap.add_argument("--step", type=int, default=8)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--pinned", action="store_true", help="inputs and the Section buffer in pinned host memory")
a = ap.parse_args()
pre = synth.preset(a.mode, 1024, 2048, a.step)
st = api.make_stixels(pre, max_batch=1)
fr = synth.make_frame(0, rows=1024, cols=2048, column_step=a.step)
sec_out = None
if a.pinned:
    import torch
    from instance_stixels_b200 import _lib as L
    keep = [torch.from_numpy(fr.disparity).pin_memory(), torch.from_numpy(fr.segmentation).pin_memory(),
            torch.empty(st.GetRealCols() * 200 * 32, dtype=torch.uint8).pin_memory()]
    fr.disparity, fr.segmentation = keep[0].numpy(), keep[1].numpy()
    sec_out = keep[2].numpy().view(L.SECTION_DTYPE)
names = ["SetDisparityImage", "SetSegmentation", "SetRoadParameters", "Compute", "GetInstanceStixels"]
acc = {n: [] for n in names}
tot = []
st.set_profiling(True)
for i in range(a.reps + 5):
    t = [time.perf_counter()]
    st.SetDisparityImage(fr.disparity); t.append(time.perf_counter())
    st.SetSegmentation(fr.segmentation); t.append(time.perf_counter())
    st.SetRoadParameters(**fr.road); t.append(time.perf_counter())
    st.Compute(a.mode == "pairwise", sections_out=sec_out); t.append(time.perf_counter())
    st.GetInstanceStixels(); t.append(time.perf_counter())
    if i >= 5:
        for k, n in enumerate(names):
            acc[n].append(1e3 * (t[k + 1] - t[k]))
        tot.append(1e3 * (t[-1] - t[0]))
    elif i == 4:
        st.stage_times(reset=True)
print(a.mode, "w", a.step, "pinned" if a.pinned else "pageable", "p50 total %.3f ms  p99 %.3f" % (np.median(tot), np.max(tot)))
for n in names:
    print("  %-20s p50 %.3f ms" % (n, np.median(acc[n])))
print("  kernel stages (ms per frame):", {k: round(v[0] / max(v[1], 1), 3) for k, v in st.stage_times().items()})
st.Finish()
