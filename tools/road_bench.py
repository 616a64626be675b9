"""Road estimation: device throughput / latency, the v-disparity kernel against the HBM
roofline, and the reference's host step (cv2.HoughLines on the binary image, numpy histogram) timed beside it.
  python tools/road_bench.py [--batch 64] [--reps 20]"""
import argparse, importlib.util, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from instance_stixels_b200 import api, synth
from oracle import road_cpu

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rows, cols, D = 1024, 2048, 128
disp, _, _ = synth.make_batch(min(a.batch, 16), rows=rows, cols=cols)
disp = np.tile(disp, ((a.batch + len(disp) - 1) // len(disp), 1, 1))[:a.batch]
d = torch.from_numpy(disp).cuda()
re = api.RoadEstimation()
re.Initialize(512.0, 0.209313, 2262.52, rows, cols, D, 0.2, max_batch=a.batch)
for _ in range(3):
    re.ComputeBatchDevice(a.batch, d.data_ptr())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.reps):
    est = re.ComputeBatchDevice(a.batch, d.data_ptr())
dt = (time.perf_counter() - t0) / a.reps
lat = []
for i in range(30):
    t0 = time.perf_counter(); re.Compute(d[i % a.batch].data_ptr()); lat.append(1e3 * (time.perf_counter() - t0))
out = dict(component="road_estimation", frames_per_s_batch=a.batch / dt, batch=a.batch, ms_per_batch=1e3 * dt,
           latency_ms_single_p50=float(np.median(lat[5:])), input_gbs=a.batch * rows * cols * 4 / dt / 1e9,
           all_ok=all(e["ok"] for e in est))
if importlib.util.find_spec("cv2"):
    import cv2
    t0 = time.perf_counter()
    for i in range(8):
        vd = road_cpu.vdisparity(disp[i], D)
        b = road_cpu.binary_image(vd, 0.2)
        t1 = time.perf_counter()
        cv2.HoughLines(b, 1.0, np.pi / 180, 25)
        hough = time.perf_counter() - t1
    out["cpu_reference_step"] = dict(ms_per_frame_total=1e3 * (time.perf_counter() - t0) / 8, ms_cv2_houghlines=1e3 * hough,
                                     what="numpy v-disparity + cv2.HoughLines (the reference's host step), 1 core")
print(json.dumps(out))
re.Finish()
