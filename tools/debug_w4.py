import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from instance_stixels_b200 import api, synth, _lib as L
from oracle import refbind
import parity
pre = synth.preset("pairwise", 1024, 2048, 4)
fr = synth.make_frame(0, rows=1024, cols=2048, column_step=4)
ref = refbind.RefStixels(api.StixelConfig(**pre))
r1, _, _ = ref.compute(True, fr.disparity, fr.segmentation, fr.road)
r2, _, _ = ref.compute(True, fr.disparity, fr.segmentation, fr.road)
print("ref deterministic:", np.array_equal(r1.view(np.uint8), r2.view(np.uint8)))
st = api.make_stixels(pre)
outs = []
for i in range(2):
    st.SetDisparityImage(fr.disparity); st.SetSegmentation(fr.segmentation); st.SetRoadParameters(**fr.road)
    outs.append(st.Compute(True).sections.copy())
print("ours deterministic:", np.array_equal(outs[0].view(np.uint8), outs[1].view(np.uint8)))
o = outs[0]
n = parity.column_lengths(o)
for c in range(o.shape[0]):
    a, b = o[c, :n[c]], r1[c, :n[c]]
    bad = np.flatnonzero(a["cost"].view(np.int32) != b["cost"].view(np.int32))
    for j in bad:
        print("col", c, "j", j, "ours", a[j], "ref", b[j], a[j]["cost"].view(np.int32) - b[j]["cost"].view(np.int32))
        print("  neighbours ours:", a[max(0, j - 1):j + 2])
cost = st.read_tensor(L.T_COST_TABLE).reshape(o.shape[0], 1024, 3)
idx = st.read_tensor(L.T_INDEX_TABLE).reshape(o.shape[0], 1024, 3)
for c in range(o.shape[0]):
    a, b = o[c, :n[c]], r1[c, :n[c]]
    bad = np.flatnonzero(a["cost"].view(np.int32) != b["cost"].view(np.int32))
    for j in bad:
        vT, vB = a[j]["vT"], a[j]["vB"]
        print("cost row vT", cost[c, vT], idx[c, vT], "prev row", cost[c, vB - 1] if vB > 0 else None, idx[c, vB - 1] if vB > 0 else None)
