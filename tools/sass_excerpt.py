#!/usr/bin/env python
"""Counts of the SASS mnemonics that show how the hot kernels move data (no GPU needed):
vector loads/stores, L2 bulk prefetch, bulk-async (TMA) copies, mbarrier ops, system-scope
acquire/release, warp syncs.  python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ellspmv_b200", "lib", "libellspmv_cuda.so")
WANT = [r"ell_thread_kernel<int, 2, 5, false, true, 0, 1, false, false>", r"ell_thread_kernel<int, 1, 27, false, true, 0, 1, false, false>",
        r"ell_thread_kernel<int, 4, 5, false, true, 0, 1, false, false>", r"ell_thread_kernel<int, 2, 5, false, true, 0, 1, false, true>",
        r"ell_thread_kernel<int, 2, 5, false, true, 0, 2, false, false>", r"ell_thread_kernel<int, 1, 0, false, true, 0, 0, true, false>",
        r"ell_longrow_kernel<int>", r"ell_longrow_ring_kernel<int, 4, 2, 2, true>", r"ell_longrow_ring_kernel<int, 4, 2, 2, false>", r"sg_gather_kernel<int>", r"sg_sum_kernel", r"sell_spmv_kernel<int, false>", r"sell_long_kernel<int>",
        r"csr_stream_kernel<int, false>", r"ell_bulk_kernel<int, 1, false>", r"peer_sync_kernel", r"peer_barrier_kernel"]
KEYS = re.compile(r"\b(LDGSTS[\w.]*|LDGDEPBAR|DEPBAR[\w.]*|LDG\.E[\w.]*|STG\.E[\w.]*|UBLKPF[\w.]*|UBLKCP[\w.]*|SYNCS[\w.]*|ATOMG[\w.]*|RED[\w.]*|MEMBAR[\w.]*|CCTL[\w.]*|"
                  r"WARPSYNC[\w.]*|BAR\.SYNC[\w.]*|SHFL[\w.]*|DMUL|DADD|DFMA|LDS[\w.]*|STS[\w.]*|NANOSLEEP|REDUX[\w.]*|MATCH[\w.]*)")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
mangled = re.findall(r"Function : (\S+)", out)
demangle = dict(zip(mangled, names))
blocks = re.split(r"\n\s*Function : ", out)[1:]
print(f"# SASS of {os.path.relpath(LIB, ROOT)} (sm_100a), key mnemonics per kernel: count x mnemonic")
for b in blocks:
    m = b.split("\n", 1)[0].strip()
    d = demangle.get(m, m)
    d = d.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "")
    short = re.sub(r"^(void )?ellspmv::", "", d)
    short = short[: short.rindex("(")] if "(" in short else short
    if not any(short == w or short.startswith(w + "<") and w.endswith("kernel") for w in WANT) and short not in WANT:
        continue
    c = collections.Counter(KEYS.findall(b))
    regs = ""
    print(f"\n{short}")
    for k, v in sorted(c.items()):
        print(f"    {v:4d} x {k}")
