"""The oracle's restatement of OAK's grid interpolation (oracle/oak_ndgrid.c: split, interp_tetrahedron, interp_cube,
the databox search, cinterp) pinned on the reference's own test, test/test_ndgrid.F90:11-32: analytic linear fields on
1- to 5-dimensional grids, including the degenerate ones with singleton dimensions, interpolated at the middle of the
domain through the coefficients of cinterp, tolerance 1e-6."""
import numpy as np
import pytest

import oracle


def _nd_case(sz):
    # test_ndgrid_nd (test/test_ndgrid.F90:199-281): x(j,i) = 0-based subscript i of point j, f = sum_i 2 i x_i
    sz = tuple(sz)
    n = len(sz)
    sub = np.indices(sz).reshape(n, -1, order="F").astype(np.float64)  # first dimension fastest
    f = sum(2 * (i + 1) * sub[i] for i in range(n))
    xi = sub.mean(axis=1)
    return sub, f, xi


@pytest.mark.parametrize("sz", [(10,), (10, 20), (2, 2, 2), (2, 2, 2, 2),
                                (1, 2), (2, 1, 2), (10, 1, 20), (10, 3, 1, 20), (10, 3, 1), (1, 10, 3, 1),
                                (1, 1, 1, 10), (1, 1, 1, 10, 1)])
def test_cinterp_reference_cases_nd(sz):
    # test/test_ndgrid.F90:16-31
    coord, f, xi = _nd_case(sz)
    idx, co, nbp = oracle.cinterp(sz, coord, xi[None, :])
    n = len(sz)
    assert nbp[0] == 2 ** n
    ioff = np.cumprod((1,) + tuple(sz[:-1]))
    lin = ((idx[0] - 1) * ioff).sum(axis=1)
    fi = float((co[0] * f[lin]).sum())
    ref = float(sum(2 * (i + 1) * xi[i] for i in range(n)))
    assert abs(fi - ref) < 1e-6
    assert abs(co[0].sum() - 1.0) < 1e-6


@pytest.mark.parametrize("m,n", [(10, 20), (1, 3), (3, 1)])
def test_cinterp_reference_cases_2d(m, n):
    # test_ndgrid_2d (test/test_ndgrid.F90:41-112): x = i+1, y = j+2, f = 2x + 4y, point = mean of the coordinates
    ii, jj = np.meshgrid(np.arange(1, m + 1), np.arange(1, n + 1), indexing="ij")
    x = (ii + 1.0).ravel(order="F"); y = (jj + 2.0).ravel(order="F")
    f = 2 * x + 4 * y
    xi = np.array([x.mean(), y.mean()])
    idx, co, nbp = oracle.cinterp((m, n), np.stack([x, y]), xi[None, :])
    assert nbp[0] == 4
    lin = (idx[0, :, 0] - 1) + m * (idx[0, :, 1] - 1)
    assert abs(float((co[0] * f[lin]).sum()) - (2 * xi[0] + 4 * xi[1])) < 1e-6


def test_split_gives_the_simplices_of_a_cell():
    # split (ndgrid.F90:357-435): n! 2^(n-1) simplices, every vertex a convex combination of cube corners, the last
    # vertex the centre of the cell, and together they tile the cell (volumes add up to 1 on the unit cube)
    for n in (1, 2, 3, 4):
        t = oracle.tetrahedra(n)
        nb = {1: 1, 2: 4, 3: 24, 4: 192}[n]
        assert t.shape == (nb, n + 1, 2 ** n)
        assert np.allclose(t.sum(axis=2), 1.0) and (t >= 0).all()
        if n > 1:
            assert np.allclose(t[:, n, :], 1.0 / 2 ** n)
        corners = np.array([[(j >> k) & 1 for k in range(n)] for j in range(2 ** n)], dtype=float)  # [2^n][n]
        vol = 0.0
        for l in range(nb):
            V = t[l] @ corners                      # [n+1][n] vertices
            vol += abs(np.linalg.det(V[1:] - V[0])) / float(np.prod(range(1, n + 1)))
        assert abs(vol - 1.0) < 1e-12


def test_cinterp_out_of_grid_and_masked_points():
    # nbp = 0 outside the grid and when a corner of the cell is masked (ndgrid.F90:1205-1233), which genObservationOper
    # turns into a zero row with model index -1 (assimilation.F90:2597-2611)
    sz = (6, 5)
    coord = oracle.ndgrid_full_coords(sz, axes=[np.arange(6.0), 10.0 + 2.0 * np.arange(5.0)])
    masked = np.zeros(30, np.uint8)
    masked[2 + 6 * 1] = 1
    pts = np.array([[0.5, 10.5], [-0.1, 12.0], [2.5, 12.5], [1.5, 11.0], [5.0, 18.0], [3.3, 18.1]])
    idx, co, nbp = oracle.cinterp(sz, coord, pts, masked=masked)
    assert list(nbp) == [4, 0, 0, 0, 4, 0]   # inside; outside; (2.5,12.5) and (1.5,11.0) touch the masked node; corner; outside
    assert abs(co[0].sum() - 1) < 1e-12


def test_simplex_of_a_point_follows_from_the_order_of_its_offsets_from_the_centre():
    # what k_cinterp's fast path relies on (oak_b200/csrc/hgen.cu): in the numbering of split (ndgrid.F90:357-435) the
    # simplex that contains u is found level by level from the free dimension with the largest |u_k - 1/2| and its side
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4):
        t = oracle.tetrahedra(n)
        nb = t.shape[0]
        corners = np.array([[(q >> k) & 1 for k in range(n)] for q in range(2 ** n)], dtype=float)
        V = t @ corners                                  # [nb][n+1][n] vertices of the simplices of the unit cell
        for _ in range(300):
            u = rng.uniform(0, 1, n)
            lstar, sub, fixed = 0, nb, 0
            for level in range(n - 1):
                free = [k for k in range(n) if not (fixed >> k) & 1]
                best = max(free, key=lambda k: (abs(u[k] - 0.5), -k))
                sub //= 2 * (n - level)
                lstar += (2 * free.index(best) + (1 if u[best] > 0.5 else 0)) * sub
                fixed |= 1 << best
            M = np.vstack([np.ones(n + 1), V[lstar].T])
            c = np.linalg.solve(M, np.concatenate([[1.0], u]))
            assert (c >= -1e-12).all()
