#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills, smem)."""
import re, subprocess, sys
src = sys.argv[1] if len(sys.argv) > 1 else "ell_kernels.cu"
cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
       "-ccbin", "/usr/bin/g++", "-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
spill = ""
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ellspmv::", "")
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = f"stack={m.group(1)} spill_st={m.group(2)} spill_ld={m.group(3)}"
        continue
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m and name:
        smem = re.search(r"(\d+) bytes smem", m.group(2))
        print(f"{name:70s} regs={m.group(1):>3s} {spill} smem={smem.group(1) if smem else 0}")
        name = None
