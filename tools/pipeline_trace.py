"""Developer probe (GPU box): the host-batch pipeline chunk by chunk.  Streams STEPS batches of B frames through
isx_submit_batch_host / isx_wait_batch_host with DEPTH batches in flight and prints, per chunk, the device time stamps
of isx_get_chunk_trace (input copy begin | end; join | frame tables | column tables | tables end | DP begin | DP end; emission
begin | grouping begin | emission end) and, per batch, how long the host thread spent in submit and wait.

  MODE=pairwise B=64 DEPTH=3 STEPS=8 KIND=float python tools/pipeline_trace.py
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instance_stixels_b200 import api, synth

MODE = os.environ.get("MODE", "pairwise")
B = int(os.environ.get("B", "64"))
DEPTH = int(os.environ.get("DEPTH", "3"))
STEPS = int(os.environ.get("STEPS", "8"))
KIND = os.environ.get("KIND", "float")
pairwise = MODE == "pairwise"

pre = synth.preset(MODE, 1024, 2048, 8)
st = api.make_stixels(pre, max_batch=B)
disp, seg, roads = synth.make_batch(min(B, 16), rows=1024, cols=2048)
reps = B // len(roads)
disp, seg, roads = np.concatenate([disp] * reps), np.concatenate([seg] * reps), roads * reps
h_disp, h_seg = torch.from_numpy(disp).pin_memory(), torch.from_numpy(seg).pin_memory()
d16 = torch.from_numpy(np.rint(disp * 256.0).astype(np.uint16)).pin_memory()
s16 = torch.from_numpy(np.ascontiguousarray(seg[..., :128]).astype(np.int16)).pin_memory()
C_ = st.GetRealCols()
out = [torch.empty((B, C_, 200, 32), dtype=torch.uint8).pin_memory().numpy().view(api.L.SECTION_DTYPE).reshape(B, C_, 200)
       for _ in range(DEPTH)]


def submit(i):
    if KIND == "u16":
        st.SubmitBatchU16(pairwise, d16.numpy(), 1.0 / 256.0, s16.numpy(), roads, out[i % DEPTH])
    else:
        st.SubmitBatch(pairwise, h_disp.numpy(), h_seg.numpy(), roads, out[i % DEPTH])


for w in range(DEPTH + 1):
    submit(w)
    if w >= DEPTH - 1:
        st.WaitBatch()
for _ in range(DEPTH - 1):
    st.WaitBatch()
torch.cuda.synchronize()
st.set_profiling(True)
host = []
t0 = time.perf_counter()
for i in range(STEPS):
    a = time.perf_counter()
    submit(i)
    b = time.perf_counter()
    if i >= DEPTH - 1:
        st.WaitBatch()
    c = time.perf_counter()
    host.append((1e3 * (a - t0), 1e3 * (b - a), 1e3 * (c - b)))
for _ in range(min(DEPTH - 1, STEPS)):
    st.WaitBatch()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
buf = (ctypes.c_double * (11 * 256))()
nch = st._lib.isx_get_chunk_trace(st._h, buf, 256)
print(f"{MODE} {KIND} B={B} depth={DEPTH}: {B * STEPS / dt:.0f} frames/s, {1e3 * dt / STEPS:.2f} ms/batch")
print("chunk   h2d0    h2d1    join    ftab    ctab  tabend     dp0     dp1   emit0   group   emit1")
for ch in range(nch):
    print("%4d " % ch + " ".join("%7.2f" % buf[ch * 11 + k] for k in range(11)), flush=True)
print("batch  submit_at  submit_ms  wait_ms")
for i, (a, s_, w_) in enumerate(host):
    print("%4d  %9.2f  %9.2f  %7.2f" % (i, a, s_, w_))
st.Finish()
