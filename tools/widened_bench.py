"""The widened rows on the device, one JSON line each:
  road estimation: throughput / latency, and the reference's host step (numpy histogram + cv2.HoughLines) beside it;
  segmentation ingest and result images: GB/s of the bytes they must move against the HBM peak.
  python tools/widened_bench.py [--batch 64] [--reps 20]"""
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

# ---- ingest + rasteriser: CUDA events on the library's stream ----
HBM_PEAK = 6650.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    HBM_PEAK = json.load(open(pk)).get("hbm_gbs", HBM_PEAK)
B = min(a.batch, 16)
pre = synth.preset("pairwise", rows, cols, 8)
st = api.make_stixels(pre, max_batch=B)
stream = torch.cuda.ExternalStream(st.stream())
_, seg, roads = synth.make_batch(2, rows=rows, cols=cols)
cnn = torch.randn((B, 21, rows // 8, cols // 8), device="cuda")
d_seg = torch.empty((B,) + seg.shape[1:], dtype=torch.int32, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.Synchronize()
    e0.record(stream)
    for _ in range(a.reps):
        fn()
    e1.record(stream)
    st.Synchronize()
    return e0.elapsed_time(e1) * 1e-3 / a.reps


t = timed(lambda: st.FlipAndPadBatchDevice(B, cnn.data_ptr(), rows // 8, cols // 8, d_seg.data_ptr()))
nbytes = cnn.numel() * 4 + d_seg.numel() * 4
print(json.dumps(dict(component="segmentation_ingest", frames=B, us_per_frame=1e6 * t / B, gbs=nbytes / t / 1e9,
                      hbm_frac=nbytes / t / 1e9 / HBM_PEAK, bytes_per_frame=nbytes // B)))
dd = torch.from_numpy(np.tile(disp[:2], (B // 2, 1, 1))).cuda()
ds = torch.from_numpy(np.tile(seg, (B // 2, 1, 1, 1))).cuda()
st.ComputeBatchDevice(True, B, dd.data_ptr(), ds.data_ptr(), roads * (B // 2))
st.Synchronize()
lab = torch.empty((B, rows, cols), dtype=torch.uint8, device="cuda")
ins = torch.empty((B, rows, cols), dtype=torch.int32, device="cuda")
dsp = torch.empty((B, rows, cols), dtype=torch.float32, device="cuda")
t = timed(lambda: st.RasterizeBatchDevice(0, B, lab.data_ptr(), ins.data_ptr(), dsp.data_ptr()))
nbytes = lab.numel() + ins.numel() * 4 + dsp.numel() * 4
print(json.dumps(dict(component="result_images", frames=B, us_per_frame=1e6 * t / B, gbs=nbytes / t / 1e9,
                      hbm_frac=nbytes / t / 1e9 / HBM_PEAK, bytes_per_frame=nbytes // B)))
st.Finish()
