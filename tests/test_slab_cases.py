"""Host-side decomposition logic on CPU: the rank-local cases of a decomposed run are a
partition of the global case (fluid particles exactly once, bit-identical coordinates, wall
pieces that cover what a rank can reach), slabs are at least a halo thick, and the halo
width covers every pass of the step."""
import math

import numpy as np
import pytest

from titsolver_b200 import cases
from titsolver_b200.slab import halo_width


@pytest.mark.parametrize("mode,world,n_col", [("weak_z", 2, 11), ("weak_z", 3, 11), ("strong_x", 2, 11), ("strong_x", 4, 22)])
def test_slab_cases_partition_the_global_case(mode, world, n_col):
    tank = (5.366, 4.0, 1.0 * world) if mode == "weak_z" else (5.366, 4.0, 1.0)
    glob = cases.dam_break_3d(n_col, tank=tank)
    nf = glob.n_fluid
    seen = np.zeros(nf, dtype=int)
    wall_seen = np.zeros(glob.n_fixed, dtype=int)
    wall_key = {v.tobytes(): i for i, v in enumerate(glob.verts)}  # coordinates are bit-identical by construction
    for rank in range(world):
        case, edges, axis = cases.dam_break_3d_slab(n_col, world, rank, mode=mode)
        assert axis == (2 if mode == "weak_z" else 0)
        gid = case.meta["gid"]
        assert case.meta["n_fluid_global"] == nf and case.meta["n_fixed_global"] == glob.n_fixed
        # bit-identical coordinates and densities, owned range [lo, hi)
        assert np.array_equal(case.r[: case.n_fluid], glob.r[gid])
        assert np.array_equal(case.rho[: case.n_fluid], glob.rho[gid])
        lo, hi = edges[rank], edges[rank + 1]
        x = case.r[: case.n_fluid, axis]
        assert np.all((x >= lo) & (x < hi))
        seen[gid] += 1
        # wall vertices are global ones, and everything within halo + support radius of the slab is there
        halo, R, dw = halo_width(case)
        ids = np.array([wall_key[v.tobytes()] for v in case.verts])
        assert np.array_equal(case.verts, glob.verts[ids])
        wall_seen[ids] += 1
        need = (glob.verts[:, axis] >= lo - halo - R - dw) & (glob.verts[:, axis] <= hi + halo + R + dw)
        assert set(np.nonzero(need)[0]).issubset(set(ids.tolist()))
        # faces refer to local vertices and keep the orientation (normals into the tank)
        f = case.faces.astype(np.int64)
        assert f.max() < len(case.verts)
        # interior slabs are thick enough for the adjacent-slab-only exchange
        if math.isfinite(lo) and math.isfinite(hi):
            assert hi - lo >= halo
    assert np.all(seen == 1)
    assert np.all(wall_seen >= 1)


def test_halo_width_formula():
    case = cases.dam_break_3d(6)
    halo, R, dw = halo_width(case)
    assert R == pytest.approx(4 * case.dr) and dw == pytest.approx(math.sqrt(2) * case.dr)
    assert halo == pytest.approx(2 * R + dw + case.dr)
    # the wider kernels widen the halo
    assert halo_width(case, kernel_id=2)[1] == pytest.approx(6 * case.dr)


def test_rebalanced_edges_equalise_the_measured_cost():
    from titsolver_b200.slab import rebalanced_edges

    edges = [-math.inf] + [92.5 * k for k in range(1, 8)] + [math.inf]
    costs = [319.0, 261.0, 261.0, 261.0, 261.0, 261.0, 261.0, 332.0]  # profiles/r2d: the end slabs carry the end walls
    new = rebalanced_edges(edges, costs, 1.0, 736.0, min_width=10.4, max_shift=30.0)
    assert new[0] == -math.inf and new[-1] == math.inf and all(a < b for a, b in zip(new[:-1], new[1:]))
    # cost per unit length of the old slabs, integrated over the new ones: equal shares
    knots = [1.0] + edges[1:-1] + [736.0]
    dens = [c / (b - a) for c, a, b in zip(costs, knots[:-1], knots[1:])]
    def cost_of(lo, hi):
        return sum(d * max(0.0, min(hi, b) - max(lo, a)) for d, a, b in zip(dens, knots[:-1], knots[1:]))
    nk = [1.0] + new[1:-1] + [736.0]
    shares = [cost_of(a, b) for a, b in zip(nk[:-1], nk[1:])]
    assert max(shares) - min(shares) <= 1e-9 * sum(shares)
    assert new[1] < edges[1] and new[-2] > edges[-2]  # the end slabs shrink
    # limits: no edge farther than max_shift from the reference cut, no slab thinner than the halo
    tight = rebalanced_edges(edges, costs, 1.0, 736.0, min_width=10.4, max_shift=5.0)
    assert all(abs(a - b) <= 5.0 + 1e-12 for a, b in zip(tight[1:-1], edges[1:-1]))
    skew = rebalanced_edges([-math.inf, 50.0, 100.0, math.inf], [1.0, 1.0, 1000.0], 0.0, 150.0, min_width=20.0)
    assert skew[2] - skew[1] >= 20.0 - 1e-9 and 150.0 - skew[2] >= 20.0 - 1e-9
    assert rebalanced_edges([-math.inf, math.inf], [5.0], 0.0, 1.0, 0.1) == [-math.inf, math.inf]
    with pytest.raises(ValueError):
        rebalanced_edges([-math.inf, 1.0, 2.0, math.inf], [1.0, 1.0, 1.0], 0.0, 3.0, min_width=2.0)
