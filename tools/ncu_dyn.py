import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; col = {n: i for i, n in enumerate(hdr)}
inst = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.strip().startswith(".section") and ".text." in l and kern in l)
lines = []; cur = ("?", 0); stack=[]
# use inline info: nvdisasm prints '//## File "x", line N inlined at "y", line M' ; take outermost non-math file
for l in sass[start + 1:]:
    if l.strip().startswith(".section"): break
    m = re.findall(r'File "([^"]+)", line (\d+)', l)
    if m and '//##' in l:
        cur = [(os.path.basename(f), int(n)) for f, n in m]; continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): lines.append(cur)
B = 10
ex = collections.Counter(); sm = collections.Counter(); tot=0; tots=0
for k in range(min(len(inst), len(lines))):
    e = int(inst[k][col["Instructions Executed"]] or 0); s = int(inst[k][col["# Samples"]] or 0)
    chain = lines[k] if isinstance(lines[k], list) else [lines[k]]
    # attribute to the outermost frame in solve/collide/step/kernels
    pick = chain[-1]
    for f in chain:
        if f[0] in ("avsim_solve.cuh", "avsim_collide.cuh", "avsim_step.cuh", "avsim_kernels.cuh"): pick = f; break
    key = (pick[0], pick[1] // B * B)
    ex[key] += e; sm[key] += s; tot += e; tots += s
print("total inst", tot)
for (f, ln), e in sorted(ex.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{100*e/tot:5.1f}% inst  {100*sm[(f,ln)]/tots:5.1f}% samples   {f}:{ln}")
