#!/usr/bin/env python
"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines via nvdisasm line info.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top N]
Instructions are matched by their order inside the kernel's .text section (the report lists one row per SASS
instruction in address order, nvdisasm prints them in the same order with '//## File "...", line N' markers).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
inst = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.strip().startswith(".section") and ".text." in l and kern in l)
lines = []   # (file, line) per instruction
cur = ("?", 0)
for l in sass[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
print(f"{len(inst)} SASS rows in report, {len(lines)} instructions in disassembly")
n = min(len(inst), len(lines))
samp = collections.Counter(); execd = collections.Counter(); local = collections.Counter()
tot = 0
for k in range(n):
    r = inst[k]
    s = int(r[col["# Samples"]] or 0)
    e = int(r[col["Instructions Executed"]] or 0)
    samp[lines[k]] += s; execd[lines[k]] += e; tot += s
    if "LDL" in r[col["Source"]] or "STL" in r[col["Source"]]:
        local[lines[k]] += e
print(f"total samples {tot}, total warp-instructions {sum(execd.values())}")
print("---- top lines by stall samples")
for (f, ln), s in samp.most_common(top):
    print(f"{100*s/tot:5.1f}%  inst {execd[(f,ln)]:>12}  {f}:{ln}")
print("---- top lines by local-memory instructions")
for (f, ln), e in local.most_common(15):
    print(f"{e:>12}  {f}:{ln}")
byfile = collections.Counter()
for (f, ln), s in samp.items():
    byfile[f] += s
print("---- by file:", {f: f"{100*s/tot:.1f}%" for f, s in byfile.items()})
