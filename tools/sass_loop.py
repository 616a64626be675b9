"""Instruction mix of the longest backward-branch loops of one kernel in the built library (no GPU needed).
  python tools/sass_loop.py 'dp_kernelILb0ELb1' [N loops]"""
import collections, re, subprocess, sys
pat = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
so = "instance_stixels_b200/libinstance_stixels_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
blk = [b for b in blocks if pat in b.split("\n")[0]][0]
ins = []
for line in blk.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel instructions:", len(ins))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx: loops.append((i - addr_idx[tgt] + 1, addr_idx[tgt], i))
loops = sorted(l for l in loops if l[0] >= 60)
for n, s, e in loops[:nl]:
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins[s:e + 1])
    print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}: {n} instr", dict(mix.most_common()))
