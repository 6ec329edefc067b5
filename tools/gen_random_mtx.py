import numpy as np, sys
n=int(sys.argv[2]); rows=int(sys.argv[3])
rng=np.random.default_rng(0)
ri=rng.integers(1,rows+1,n); ci=rng.integers(1,rows+1,n); a=rng.standard_normal(n)
with open(sys.argv[1],'w') as f:
    f.write("%%MatrixMarket matrix coordinate real general\n{} {} {}\n".format(rows,rows,n))
    f.write("".join(f"{r} {c} {v:.17g}\n" for r,c,v in zip(ri,ci,a)))
