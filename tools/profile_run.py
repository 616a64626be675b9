"""Minimal driver for ncu: a few device-resident batches of one workload, nothing else.
  python tools/profile_run.py --mode pairwise --step 8 --batch 16 --steps 2"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from instance_stixels_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="pairwise")
ap.add_argument("--step", type=int, default=8)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--distinct", type=int, default=4, help="distinct synthetic frames (tiled to the batch)")
a = ap.parse_args()
pre = synth.preset(a.mode, 1024, 2048, a.step)
st = api.make_stixels(pre, max_batch=a.batch)
disp, seg, roads = synth.make_batch(min(a.distinct, a.batch), rows=1024, cols=2048, column_step=a.step)
reps = (a.batch + len(roads) - 1) // len(roads)
disp = np.tile(disp, (reps, 1, 1))[:a.batch]
seg = np.tile(seg, (reps, 1, 1, 1))[:a.batch]
roads = (roads * reps)[:a.batch]
d_disp, d_seg = torch.from_numpy(disp).cuda(), torch.from_numpy(seg).cuda()
torch.cuda.synchronize()
for _ in range(a.steps):
    st.ComputeBatchDevice(a.mode == "pairwise", a.batch, d_disp.data_ptr(), d_seg.data_ptr(), roads)
st.Synchronize()
st.Finish()
print("done")
