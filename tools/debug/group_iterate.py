"""One process, 2 GPUs: ellspmv_cuda_generate_sharded + ITERATE (the fused SpMV + push + step
hand-shake kernel) -- the target of the ncu capture with NVLink / peer-aperture counters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import ellspmv_b200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = E.EllMatrix.generate(E.GEN_STENCIL27, (n, n, n), (0.5, 1.0 / 52), 42, 64, num_gpus=2)
rows = n ** 3
x = np.ones(rows)
y = np.zeros(rows)
secs = A.spmv(y, x, 6, E.ITERATE)
i = A.info()
print("rows", rows, "gpus", i.num_gpus, "ms/step", [round(s * 1e3, 3) for s in secs], "pattern rows", i.pattern_rows / rows)
print("y[0], y[mid]", y[0], y[rows // 2])
A.free()
