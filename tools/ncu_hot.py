"""Top stall locations of the kernels of an .ncu-rep (source page, SASS view).
  python tools/ncu_hot.py report.ncu-rep [min_percent] [kernel substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# sections: a "Kernel Name" line, a header line containing "Source", then the instructions
sections, cur, name = [], None, ""
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1] if len(r) > 1 else ""
        continue
    if "Source" in r and "# Samples" in r:
        cur = dict(name=name, hdr=r, body=[])
        sections.append(cur)
        continue
    if cur is not None and len(r) == len(cur["hdr"]):
        cur["body"].append(r)
for k, sec in enumerate(sections):
    hdr, body = sec["hdr"], sec["body"]
    if want and want not in sec["name"]:
        continue
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iN]) for r in body) or 1
    print(f"== section {k} {sec['name'][:70]}: samples {tot}, SASS lines {len(body)}, warp-instr {sum(int(r[iE]) for r in body)}")
    agg = {}
    for r in body:
        for i in stall:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print({k2: v for k2, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 50 > tot})
    cum = 0
    for i, r in enumerate(body):
        n = int(r[iN]); cum += n
        if 100 * n / tot >= minpct:
            top = sorted(((int(r[j]), hdr[j]) for j in stall), reverse=True)[:2]
            print(f"{i:5d} {100*n/tot:5.1f}% cum={100*cum/tot:5.1f} exec={r[iE]:>10s} {r[iS].strip()[:64]:64s} {top}")
