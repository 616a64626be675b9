"""Where the host-buffer batch path spends its time beyond the kernels: variants of one ComputeBatch call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
from instance_stixels_b200 import api, synth, _lib as L
mode = sys.argv[1] if len(sys.argv) > 1 else "unary"
B = 64
pre = synth.preset(mode, 1024, 2048, 8)
st = api.make_stixels(pre, max_batch=B)
disp, seg, roads = synth.make_batch(B, rows=1024, cols=2048)
hd, hs = torch.from_numpy(disp).pin_memory(), torch.from_numpy(seg).pin_memory()
dd, ds = hd.cuda(), hs.cuda()
Cc = st.GetRealCols()
sec = torch.empty((B, Cc, 200, 32), dtype=torch.uint8).pin_memory()
sec_np = sec.numpy().view(L.SECTION_DTYPE).reshape(B, Cc, 200)
pw = mode == "pairwise"
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
def dev():
    st.ComputeBatchDevice(pw, B, dd.data_ptr(), ds.data_ptr(), roads); st.Synchronize()
print(mode, "device-resident          %.2f ms" % t(dev))
print(mode, "host full                %.2f ms" % t(lambda: st.ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np)))
print(mode, "host, no instance fetch  %.2f ms" % t(lambda: st.ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np, want_instances=False)))
lib = st._lib
r = api._roads(roads)
def raw(sections, inst):
    rc = lib.isx_compute_batch_host(st._h, int(pw), B, hd.data_ptr(), hs.data_ptr(), r, sections, None, 0, None)
    assert rc == 0
print(mode, "host, no sections D2H    %.2f ms" % t(lambda: raw(None, None)))
for ch in (8, 32):
    os.environ["ISX_CHUNK"] = str(ch)
    s2 = api.make_stixels(pre, max_batch=B)
    print(mode, "host full, chunk %2d      %.2f ms" % (ch, t(lambda: s2.ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=sec_np))))
    s2.Finish()
# ---- several host threads, one context each, alternate batches ----
import threading
del os.environ["ISX_CHUNK"]
ctxs = [st] + [api.make_stixels(pre, max_batch=B) for _ in range(3)]
pinned = [sec] + [torch.empty_like(sec).pin_memory() for _ in range(3)]
secs = [p.numpy().view(L.SECTION_DTYPE).reshape(B, Cc, 200) for p in pinned]
for nw, stagger in ((2, False), (2, True), (3, False), (3, True), (4, False), (4, True)):
    steps = 24
    def loop(w):
        if stagger:
            time.sleep(0.036 * w / nw)
        for _ in range(w, steps, nw):
            ctxs[w].ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=secs[w])
    for w in range(nw): ctxs[w].ComputeBatch(pw, hd.numpy(), hs.numpy(), roads, sections_out=secs[w])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ths = [threading.Thread(target=loop, args=(w,)) for w in range(nw)]
    [x.start() for x in ths]; [x.join() for x in ths]
    torch.cuda.synchronize()
    print(mode, "threads %d stagger %-5s %.2f ms/batch  %.0f frames/s" % (nw, stagger, 1e3 * (time.perf_counter() - t0) / steps, B * steps / (time.perf_counter() - t0)))
