"""Developer probe (GPU box): one pairwise batch through the float and the narrow host entry points, for an ncu
launch list / a quick A-B of the two paths."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instance_stixels_b200 import api, synth

B = int(os.environ.get("PROBE_B", "64"))
pre = synth.preset("pairwise", 1024, 2048, 8)
st = api.make_stixels(pre, max_batch=B)
disp, seg, roads = synth.make_batch(min(B, 16), rows=1024, cols=2048)
reps = B // len(roads)
disp, seg, roads = np.concatenate([disp] * reps), np.concatenate([seg] * reps), roads * reps
h_disp, h_seg = torch.from_numpy(disp).pin_memory(), torch.from_numpy(seg).pin_memory()
d16 = torch.from_numpy(np.rint(disp * 256.0).astype(np.uint16)).pin_memory()
s16 = torch.from_numpy(np.ascontiguousarray(seg[..., :128]).astype(np.int16)).pin_memory()
C_ = st.GetRealCols()
out = [torch.empty((B, C_, 200, 32), dtype=torch.uint8).pin_memory().numpy().view(api.L.SECTION_DTYPE).reshape(B, C_, 200) for _ in range(2)]
h_deq = torch.from_numpy((d16.numpy().astype(np.float32) / np.float32(256.0)).astype(np.float32)).pin_memory()
st.set_profiling(True)
for kind in ("float", "float_dequantized", "u16", "float", "u16"):
    def submit(i):
        if kind == "u16":
            st.SubmitBatchU16(True, d16.numpy(), 1.0 / 256.0, s16.numpy(), roads, out[i & 1])
        elif kind == "float_dequantized":
            st.SubmitBatch(True, h_deq.numpy(), h_seg.numpy(), roads, out[i & 1])
        else:
            st.SubmitBatch(True, h_disp.numpy(), h_seg.numpy(), roads, out[i & 1])
    submit(0); st.WaitBatch()
    torch.cuda.synchronize()
    n = int(os.environ.get("PROBE_STEPS", "6"))
    t0 = time.perf_counter()
    for i in range(n):
        submit(i)
        if i:
            st.WaitBatch()
    st.WaitBatch()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    import ctypes
    buf = (ctypes.c_double * (11 * 64))()
    nch = st._lib.isx_get_chunk_trace(st._h, buf, 64)
    if os.environ.get("PROBE_TRACE"):
        for c in range(nch):
            print("   chunk", c, " ".join("%7.2f" % buf[c * 11 + k] for k in range(11)), flush=True)
    stages = st.stage_times(reset=True)
    print(kind, "%.0f frames/s" % (B * n / dt), {k: round(v[0] / max(v[1], 1), 2) for k, v in stages.items()}, st.dp_units(), flush=True)
st.Finish()
