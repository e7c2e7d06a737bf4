#!/usr/bin/env python
"""Static SASS size of a kernel attributed to source lines (nvdisasm line info): python tools/sass_static.py lib.so kernel [bucket]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.strip().startswith(".section") and ".text." in l and kern in l)
cnt = collections.Counter(); cur = ("?", 0); n = 0
for l in sass[start + 1:]:
    if l.strip().startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2)) // bucket * bucket); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): cnt[cur] += 1; n += 1
print(n, "instructions =", n * 16 // 1024, "KB")
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{c:6d}  {f}:{ln}-{ln + bucket - 1}")
