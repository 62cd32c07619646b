#!/usr/bin/env python
"""Instruction count per source line of one kernel (code-size / I-cache budget aid).

    python tools/sass_lines.py deft_b200/lib/libdeft_b200.so attn_umma stage1_umma_kernelILi128ELi4 [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib, unit, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(unit) and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside, cur, cnt = False, None, collections.Counter()
for l in sass:
    if l.startswith("//---------------------"):
        inside = pat in l and ".text." in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print(f"{tot} instructions = {tot * 16 / 1024:.1f} KiB")
src = {}
for (f, n), v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    path = os.path.join("deft_b200/csrc", f)
    if f not in src and os.path.exists(path):
        src[f] = open(path).read().split("\n")
    text = src[f][n - 1].strip()[:90] if f in src else ""
    print(f"{v:6d}  {f}:{n}  {text}")
