"""CPU checks of the measurement tools' host-side logic (no GPU): the stencil builder of
tools/stencil_sweep.py produces, for BASELINE's own two stencils, exactly the matrices the oracle's
generators produce -- so its other stencils are 'the same kind of matrix, another shape'."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ell_structure_from_coo(n, rowidx, colidx, K):
    """columns of each row in file order, -1 where the row has fewer than K entries"""
    out = np.full((n, K), -1, dtype=np.int64)
    fill = np.zeros(n, dtype=np.int64)
    for r, c in zip(rowidx - 1, colidx - 1):
        out[r, fill[r]] = c
        fill[r] += 1
    return out, fill


def test_stencil_builder_matches_the_oracle_generators(oracle):
    ss = load_tool("stencil_sweep")
    for kind, dims, offsets in (("laplace2d", (7, 9), ss.star(2, 1)), ("stencil27", (4, 5, 3), ss.box(3, 3))):
        n, ri, ci, va = ss.stencil_coo(dims, offsets)
        K, ncols, ec, ea, real = oracle.gen_ell(kind, dims, (4.0, -1.0), bits=32)
        assert n == len(ea) // K == ncols and len(offsets) == K and len(va) == real
        got, fill = ell_structure_from_coo(n, ri, ci, K)
        ec = ec.reshape(n, K)
        ea = ea.reshape(n, K)
        for r in range(n):
            # the generator's stored entries of a row (values != 0) are the stencil points inside the
            # grid, in ascending column order -- what the builder emits in file order
            assert list(got[r, :fill[r]]) == list(ec[r, :fill[r]]), (kind, r)
            assert (ea[r, :fill[r]] != 0).all() and (ea[r, fill[r]:] == 0).all()
        assert (np.diff(ri) >= 0).all()                      # row-major file order


def test_star_and_box_sizes():
    ss = load_tool("stencil_sweep")
    assert [len(ss.star(3, 1)), len(ss.box(2, 2)), len(ss.star(2, 3)), len(ss.star(3, 2)), len(ss.box(3, 2))] == [7, 9, 13, 13, 19]
    assert len(ss.box(3, 3)) == 27 and len(ss.star(2, 1)) == 5
