"""Host-side timeline of the submit/wait pipeline: duration of every call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from instance_stixels_b200 import api, synth, _lib as L
mode = sys.argv[1] if len(sys.argv) > 1 else "unary"
B = 64; pw = mode == "pairwise"
pre = synth.preset(mode, 1024, 2048, 8)
st = api.make_stixels(pre, max_batch=B)
disp, seg, roads = synth.make_batch(B, rows=1024, cols=2048)
hd, hs = torch.from_numpy(disp).pin_memory(), torch.from_numpy(seg).pin_memory()
Cc = st.GetRealCols()
secs = [torch.empty((B, Cc, 200, 32), dtype=torch.uint8).pin_memory() for _ in range(2)]
sec_np = [s.numpy().view(L.SECTION_DTYPE).reshape(B, Cc, 200) for s in secs]
for w in range(2): st.ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np[w])
torch.cuda.synchronize()
T0 = time.perf_counter(); log = []
def stamp(name, t): log.append((name, 1e3 * (t - T0), 1e3 * (time.perf_counter() - t)))
N = 8
for i in range(N):
    t = time.perf_counter(); st.SubmitBatch(pw, hd.numpy(), hs.numpy(), roads, sec_np[i & 1]); stamp(f"submit {i}", t)
    if i > 0:
        t = time.perf_counter(); r = st.WaitBatch(); stamp(f"wait {i-1} ({len(r[1])} inst)", t)
t = time.perf_counter(); st.WaitBatch(); stamp(f"wait {N-1}", t)
for name, at, dur in log: print(f"{at:8.2f} ms  {name:28s} {dur:7.2f} ms")
print("total", 1e3 * (time.perf_counter() - T0) / N, "ms/batch")
