import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import ellspmv_b200 as E
from oracle.pyoracle import Oracle
orc = Oracle()
sync_each = len(sys.argv) > 1 and sys.argv[1] == "sync"
FLAGS = E.FUSED_SYNC if (len(sys.argv) > 2 and sys.argv[2] == "fused") else 0
K, ncols, ec, ea, _ = orc.gen_ell("laplace2d", (700, 97), (0.25, -0.125), bits=32)
rows = len(ea) // K
x0 = np.random.default_rng(5).uniform(-1, 1, rows)
steps = 4
want = orc.ell_iterate(rows, x0, steps, K, ec, ea)
cut = rows // 2 + 3
parts = [(0, cut), (cut, rows)]
S = [E.EllMatrix.upload(b - a, ncols, K, ec[a * K:b * K], ea[a * K:b * K], FLAGS, global_rows=rows, row_begin=a, device=0) for a, b in parts]
needs = [(S[r].info().min_col, S[r].info().max_col + 1) for r in range(2)]
print("parts", parts, "needs", needs, flush=True)
xb = [[torch.zeros(rows, dtype=torch.float64, device="cuda") for _ in range(2)] for _ in range(2)]
flags = [torch.zeros(32, dtype=torch.int64, device="cuda") for _ in range(2)]
for r in range(2):
    xb[r][0].copy_(torch.from_numpy(x0))
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
cur = 0
t0 = time.time()
for k in range(steps):
    for r in range(2):
        o = 1 - r
        a, b = parts[r]
        lo, hi = max(a, needs[o][0]), min(b, needs[o][1])
        S[r].spmv_exchange(xb[r][1 - cur][a:b], xb[r][cur], E.OVERWRITE, [xb[o][1 - cur].data_ptr()], [lo], [hi],
                           r, [o], [flags[o].data_ptr()], flags[r].data_ptr(), k + 1, streams[r].cuda_stream)
        print("launched", k, r, round(time.time() - t0, 2), flush=True)
        if sync_each:
            torch.cuda.synchronize()
            print("  flags0", flags[0][:3].tolist(), flags[0][16].item(), "flags1", flags[1][:3].tolist(), flags[1][16].item(), flush=True)
    cur = 1 - cur
    if len(sys.argv) > 3 and sys.argv[3] == "stepsync":
        torch.cuda.synchronize()
torch.cuda.synchronize()
print("done", round(time.time() - t0, 2), "flags0", flags[0][:3].tolist(), flags[0][16].item(), "flags1", flags[1][:3].tolist(), flags[1][16].item(), flush=True)
got = torch.cat([xb[0][cur][:cut], xb[1][cur][cut:]]).cpu().numpy()
eq = got.view(np.uint64) == want.view(np.uint64)
print("bit-equal:", eq.all(), "mismatches", int((~eq).sum()), "first at", np.flatnonzero(~eq)[:10], "cut", cut)
if not eq.all():
    i = np.flatnonzero(~eq)
    print("rows mod 97:", (i % 97)[:20], "row/97:", (i // 97)[:20], "max abs diff", np.abs(got - want)[i].max())
