"""Top stall locations of one kernel of an .ncu-rep (source page, SASS view).
  python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[iN]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][iN]))[:N]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[j]), hdr[j]) for j in stall), reverse=True)[:2]
    print(f"{i:5d} {int(r[iN]):7d} {100*int(r[iN])/tot:5.1f}% exec={r[iE]:>10s} {r[iS].strip()[:70]:70s} {top}")
