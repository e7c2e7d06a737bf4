#!/usr/bin/env python
"""Static SASS size per (cloned) device function: python tools/sass_funcs.py lib.so [kernel-substring]"""
import os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1]); filt = sys.argv[2] if len(sys.argv) > 2 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", cubin], capture_output=True, text=True).stdout.split("\n")
cur = None; cnt = {}; order = []
for l in sass:
    m = re.match(r'^(\$?[_A-Za-z][\w$]*):', l)
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1)
        if cur not in cnt: cnt[cur] = 0; order.append(cur)
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l) and cur: cnt[cur] += 1
for k in order:
    if cnt[k] > 60 and filt in k: print(f"{cnt[k]:6d} instr {cnt[k]*16/1024:6.1f} KB  {k[-70:]}")
