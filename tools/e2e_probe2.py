"""Multi-worker throughput of the batch entry points, with and without the host copies.
  python tools/e2e_probe2.py [unary|pairwise] [workers] [batches per worker]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from instance_stixels_b200 import api, synth, _lib as L
mode = sys.argv[1] if len(sys.argv) > 1 else "unary"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B = 64
pw = mode == "pairwise"
pre = synth.preset(mode, 1024, 2048, 8)
sts = [api.make_stixels(pre, max_batch=B) for _ in range(W)]
disp, seg, roads = synth.make_batch(B, rows=1024, cols=2048)
hd, hs = torch.from_numpy(disp).pin_memory(), torch.from_numpy(seg).pin_memory()
dd, ds = hd.cuda(), hs.cuda()
Cc = sts[0].GetRealCols()
secs = [torch.empty((B, Cc, 200, 32), dtype=torch.uint8).pin_memory() for _ in range(W)]
sec_np = [s.numpy().view(L.SECTION_DTYPE).reshape(B, Cc, 200) for s in secs]

def run(name, fn, stagger=True):
    for w in range(W): fn(w)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    def loop(w):
        torch.cuda.set_device(0)
        if stagger: time.sleep(0.033 * w / W)
        for _ in range(NB): fn(w)
    th = [threading.Thread(target=loop, args=(w,)) for w in range(W)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode} {W} workers x {NB} batches  {name:44s} {1e3*dt/(W*NB):6.2f} ms/batch  {B*W*NB/dt:6.0f} frames/s", flush=True)

def dev_only(w):
    sts[w].ComputeBatchDevice(pw, B, dd.data_ptr(), ds.data_ptr(), roads); sts[w].Synchronize()
def dev_fetch(w):
    sts[w].ComputeBatchDevice(pw, B, dd.data_ptr(), ds.data_ptr(), roads); sts[w].FetchBatchResults(B)
def host_full(w):
    sts[w].ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np[w])
def host_noinst(w):
    sts[w].ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np[w], want_instances=False)
run("device-resident + Synchronize", dev_only)
run("device-resident + FetchBatchResults", dev_fetch)
run("host in/out (e2e)", host_full)
run("host in/out, no instance records", host_noinst)
# one worker, back to back without host syncs: the kernel-only bound
st = sts[0]
for _ in range(2): st.ComputeBatchDevice(pw, B, dd.data_ptr(), ds.data_ptr(), roads)
st.Synchronize(); t0 = time.perf_counter()
for _ in range(W * NB): st.ComputeBatchDevice(pw, B, dd.data_ptr(), ds.data_ptr(), roads)
st.Synchronize(); dt = time.perf_counter() - t0
print(f"{mode} 1 context back to back, no host sync: {1e3*dt/(W*NB):6.2f} ms/batch  {B*W*NB/dt:6.0f} frames/s")
